#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python tools/probe_mix.py > gpurun_out/s10_mix.log 2>&1
timeout 300 python tools/probe_build.py 2 10 > gpurun_out/s10_build.log 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "builder or c1 or golden" > gpurun_out/s10_pytest.log 2>&1; tail -3 gpurun_out/s10_pytest.log
grep -v "^+" gpurun_out/s10_mix.log gpurun_out/s10_build.log
