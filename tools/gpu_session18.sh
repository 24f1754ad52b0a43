#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_api.py -m gpu -x -q -k "scene_file or host or lazy" > gpurun_out/s18_pytest.log 2>&1; tail -3 gpurun_out/s18_pytest.log
timeout 600 python tools/probe_e2e.py > gpurun_out/s18_e2e.log 2>&1
grep -v "^+" gpurun_out/s18_e2e.log
