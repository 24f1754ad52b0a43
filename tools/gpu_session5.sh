#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python tools/probe_ctas.py > gpurun_out/s5_ctas.log 2>&1
grep -v "^+" gpurun_out/s5_ctas.log
