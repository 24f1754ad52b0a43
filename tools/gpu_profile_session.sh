#!/bin/bash
# final N=1 session of round 2: GPU tests, smoke, bench (both arms), launch list, full ncu of the cast kernel,
# per-angle profile, the five configs, call shapes, C1 tail
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
M1=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
M2=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum
M3=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,sm__cycles_active.avg,sm__cycles_active.max,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed
timeout 1200 python -m pytest tests -x -q -m gpu > $O/s15_pytest.log 2>&1; tail -3 $O/s15_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/s15_smoke.log 2>&1; tail -1 $O/s15_smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > $O/s15_bench_n1.json 2> $O/s15_bench_n1.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $O/s15_bench_ref.json 2> $O/s15_bench_ref.err
timeout 900 ncu --metrics $M1 --clock-control none -c 400 --csv --log-file $O/s15_final_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/s15_ncu_a.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_trace5 -s 4 -c 1 -o $O/s15_trace5 -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/s15_ncu_b.log 2>&1
ncu -i $O/s15_trace5.ncu-rep --page raw --csv > $O/s15_trace5_raw.csv 2>/dev/null
ncu -i $O/s15_trace5.ncu-rep --page source --csv > $O/s15_trace5_src.csv 2>/dev/null
rm -f $O/s15_trace5.ncu-rep
timeout 600 python tools/profile_angles.py --counters $O/s15_angle_counters.json > $O/s15_angles.log 2>&1
timeout 900 ncu --metrics $M2 --clock-control none -k regex:k_trace5 --csv --log-file $O/s15_angle_launches.csv python tools/profile_angles.py > $O/s15_angles_ncu.log 2>&1
timeout 1500 python tests/measure/run_configs.py > $O/s15_configs.json 2> $O/s15_configs.err
timeout 1500 ncu --metrics $M3 --clock-control none -k regex:k_trace5 -c 60 --csv --log-file $O/s15_configs_ncu.csv python tests/measure/run_configs.py c1 c3 c4 > $O/s15_configs_ncu.log 2>&1
timeout 600 python tools/probe_small.py > $O/s15_small.log 2>&1
timeout 600 python tools/probe_perf.py > $O/s15_perf.log 2>&1
timeout 600 python tools/probe_build.py 2 10 50 > $O/s15_build.log 2>&1
ls -la $O | grep s15
