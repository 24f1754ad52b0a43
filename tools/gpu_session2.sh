#!/bin/bash
# round-2 GPU session 2: correctness of the new list path / widening, then A/B of kernel variants
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_environment.py tests/test_gpu_api.py -m gpu -x -q > gpurun_out/s2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s2_pytest.log
tail -8 gpurun_out/s2_pytest.log
timeout 600 python tools/probe_perf.py > gpurun_out/s2_perf_base.log 2>&1
for v in m3s8 m3s2; do QSMRT_LIB=build/variants/libqsmrt_$v.so timeout 300 python tools/probe_perf.py sun > gpurun_out/s2_perf_$v.log 2>&1; done
for v in m4s2 m4s8; do QSMRT_LIB=build/variants/libqsmrt_$v.so timeout 300 python tools/probe_perf.py sky > gpurun_out/s2_perf_$v.log 2>&1; done
QSMRT_LIB=build/variants/libqsmrt_m0s4.so timeout 300 python tools/probe_perf.py c2 c1 > gpurun_out/s2_perf_m0s4.log 2>&1
cat gpurun_out/s2_perf_*.log | grep -v "^+" | tail -80
