#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_api.py tests/test_gpu_parity.py tests/test_gpu_environment.py -m gpu -x -q > gpurun_out/s6_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s6_pytest.log
tail -30 gpurun_out/s6_pytest.log
timeout 600 python tools/probe_perf.py c1 c2 > gpurun_out/s6_perf.log 2>&1
timeout 300 python tools/probe_build.py 2 > gpurun_out/s6_build.log 2>&1
grep -v "^+" gpurun_out/s6_perf.log gpurun_out/s6_build.log
