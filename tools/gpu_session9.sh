#!/bin/bash
# profiling session for profiles/r02_*: launch lists, full ncu of the cast kernel, per-angle profile, configs, build, C1 tail
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
M1=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
M2=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum
M3=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,sm__cycles_active.avg,sm__cycles_active.max,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/s9_bench_short.json 2> $O/s9_bench_short.err
timeout 900 ncu --metrics $M1 --clock-control none -c 400 --csv --log-file $O/s9_final_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/s9_ncu_a.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_trace5 -s 4 -c 1 -o $O/s9_trace5 -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/s9_ncu_b.log 2>&1
ncu -i $O/s9_trace5.ncu-rep --page raw --csv > $O/s9_trace5_raw.csv 2>/dev/null
ncu -i $O/s9_trace5.ncu-rep --page source --csv > $O/s9_trace5_src.csv 2>/dev/null
rm -f $O/s9_trace5.ncu-rep
timeout 600 python tools/profile_angles.py --counters $O/s9_angle_counters.json > $O/s9_angles.log 2>&1
timeout 900 ncu --metrics $M2 --clock-control none -k regex:k_trace5 --csv --log-file $O/s9_angle_launches.csv python tools/profile_angles.py > $O/s9_angles_ncu.log 2>&1
timeout 1500 python tests/measure/run_configs.py > $O/s9_configs.json 2> $O/s9_configs.err
timeout 1500 ncu --metrics $M3 --clock-control none -k regex:k_trace5 -c 60 --csv --log-file $O/s9_configs_ncu.csv python tests/measure/run_configs.py c1 c3 c4 > $O/s9_configs_ncu.log 2>&1
timeout 600 python tools/probe_build.py 2 10 50 > $O/s9_build.log 2>&1
timeout 600 ncu --metrics $M1 --clock-control none -c 200 --csv --log-file $O/s9_build2m_launches.csv python tools/probe_build.py 2 > $O/s9_ncu_c.log 2>&1
for sm in 1 8; do timeout 600 ncu --metrics $M3 --clock-control none -k regex:"k_trace5|k_cast_rays" --csv --log-file $O/s9_c1_split$sm.csv python tools/probe_c1.py 2 $sm > $O/s9_c1_split$sm.log 2>&1; done
timeout 600 python tools/probe_small.py > $O/s9_small.log 2>&1
timeout 600 python tools/probe_perf.py > $O/s9_perf.log 2>&1
ls -la $O | tail -30
