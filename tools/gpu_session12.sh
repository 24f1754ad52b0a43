#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python tools/probe_big.py > gpurun_out/s12_big_base.log 2>&1
timeout 300 python tools/probe_perf.py c2 cnt > gpurun_out/s12_perf_base.log 2>&1
for v in setl; do QSMRT_LIB=build/variants/libqsmrt_$v.so timeout 300 python tools/probe_perf.py cnt > gpurun_out/s12_perf_$v.log 2>&1; QSMRT_LIB=build/variants/libqsmrt_$v.so timeout 300 python tools/probe_big.py c3 > gpurun_out/s12_big_$v.log 2>&1; done
for v in minb12 setl12; do QSMRT_LIB=build/variants/libqsmrt_$v.so timeout 300 python tools/probe_perf.py c2 cnt > gpurun_out/s12_perf_$v.log 2>&1; QSMRT_LIB=build/variants/libqsmrt_$v.so timeout 600 python tools/probe_big.py > gpurun_out/s12_big_$v.log 2>&1; done
QSMRT_LIB=build/variants/libqsmrt_setl.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "count or list or soup or c1" > gpurun_out/s12_pytest_setl.log 2>&1; tail -3 gpurun_out/s12_pytest_setl.log
grep -h -v "^+" gpurun_out/s12_*.log
