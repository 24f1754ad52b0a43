#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "list or soup or property or kat or reference_call or edge_cases or count or scene_file or host" > gpurun_out/s19_pytest.log 2>&1; tail -5 gpurun_out/s19_pytest.log
timeout 300 python tools/probe_perf.py cnt > gpurun_out/s19_perf.log 2>&1
timeout 300 python tools/probe_small.py > gpurun_out/s19_small.log 2>&1
grep -h -v "^+" gpurun_out/s19_perf.log; grep -h "list_inter\|count_inter" gpurun_out/s19_small.log
