#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python tools/probe_big.py > gpurun_out/s11_big_base.log 2>&1
timeout 300 python tools/probe_perf.py c2 sun > gpurun_out/s11_perf_base.log 2>&1
for v in pfq1 pfq2; do QSMRT_LIB=build/variants/libqsmrt_$v.so timeout 300 python tools/probe_perf.py c2 sun > gpurun_out/s11_perf_$v.log 2>&1; QSMRT_LIB=build/variants/libqsmrt_$v.so timeout 300 python tools/probe_big.py c3 > gpurun_out/s11_big_$v.log 2>&1; done
for v in pff1 pff2; do QSMRT_LIB=build/variants/libqsmrt_$v.so timeout 600 python tools/probe_big.py c4 > gpurun_out/s11_big_$v.log 2>&1; done
grep -h -v "^+" gpurun_out/s11_*.log
