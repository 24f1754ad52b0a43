#!/bin/bash
# round-2 GPU session 1: tests, bench, launch list with instruction counts, full ncu of the cast kernel, lane census
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/s1_smi.txt
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/s1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s1_pytest.log
tail -5 gpurun_out/s1_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/s1_bench.json 2> gpurun_out/s1_bench.err; echo "bench rc=$?"
tail -c 600 gpurun_out/s1_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum \
   --clock-control none -k regex:k_trace5 -c 40 --csv --log-file gpurun_out/s1_trace_launches.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/s1_ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_trace5 -s 4 -c 1 -o gpurun_out/s1_trace5 -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/s1_ncu2.log 2>&1
ncu -i gpurun_out/s1_trace5.ncu-rep --page raw --csv > gpurun_out/s1_trace5_raw.csv 2>/dev/null
ncu -i gpurun_out/s1_trace5.ncu-rep --page source --csv > gpurun_out/s1_trace5_src.csv 2>/dev/null
timeout 600 python tools/probe_tune.py --mesh c2 --angles 2 --combos "2:12,16,1,2,0,1" > gpurun_out/s1_tune_c2.log 2>&1
timeout 600 python tools/probe_tune.py --mesh c1 --grid 1000 --angles 2 --combos "2:12,16,1,2,0,1;2:12,16,1,1,0,1;2:12,16,1,4,0,1" > gpurun_out/s1_tune_c1.log 2>&1
ls -la gpurun_out | tail -20
