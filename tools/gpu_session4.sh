#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/s4_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s4_pytest.log
tail -6 gpurun_out/s4_pytest.log
timeout 600 python tools/probe_perf.py > gpurun_out/s4_perf_base.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,sm__cycles_active.avg,sm__cycles_active.max,sm__cycles_elapsed.max,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct \
   --clock-control none -k regex:"k_trace5|k_cast_rays|k_count" --csv --log-file gpurun_out/s4_c1_ncu.csv python tools/probe_c1.py 2 > gpurun_out/s4_c1.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/s4_bench.json 2> gpurun_out/s4_bench.err; echo "bench rc=$?"
grep -v "^+" gpurun_out/s4_perf_base.log
