#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python tools/probe_select.py > gpurun_out/s16_select_12.log 2>&1
QSMRT_LIB=build/variants/libqsmrt_m0q10.so timeout 600 python tools/probe_select.py > gpurun_out/s16_select_10.log 2>&1
grep -h -v "^+" gpurun_out/s16_select_*.log
