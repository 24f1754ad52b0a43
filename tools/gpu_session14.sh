#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for v in stkl stkl10 stkl16; do QSMRT_LIB=build/variants/libqsmrt_$v.so timeout 300 python tools/probe_perf.py c2 > gpurun_out/s14_perf_$v.log 2>&1; done
QSMRT_LIB=build/variants/libqsmrt_stkl.so timeout 300 python tools/probe_perf.py c1 > gpurun_out/s14_perf_stkl_c1.log 2>&1
grep -h -v "^+" gpurun_out/s14_perf_*.log
