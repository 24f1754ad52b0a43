#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/s7_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s7_pytest.log
tail -12 gpurun_out/s7_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/s7_bench.json 2> gpurun_out/s7_bench.err; echo "bench rc=$?"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s7_smoke.log 2>&1; tail -2 gpurun_out/s7_smoke.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/s7_bench_ref.json 2> gpurun_out/s7_bench_ref.err; echo "ref rc=$?"
