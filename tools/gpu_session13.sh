#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/s13_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s13_pytest.log
tail -5 gpurun_out/s13_pytest.log
timeout 300 python tools/probe_perf.py c2 sun sky > gpurun_out/s13_perf_base.log 2>&1
QSMRT_LIB=build/variants/libqsmrt_m34_12.so timeout 300 python tools/probe_perf.py sun sky > gpurun_out/s13_perf_m34.log 2>&1
grep -h -v "^+" gpurun_out/s13_perf_*.log
