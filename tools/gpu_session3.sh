#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py -m gpu -x -q > gpurun_out/s3_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s3_pytest.log
tail -6 gpurun_out/s3_pytest.log
timeout 600 python tools/probe_build.py 2 10 50 > gpurun_out/s3_build_base.log 2>&1
for v in rf5 rf6; do QSMRT_LIB=build/variants/libqsmrt_$v.so timeout 600 python tools/probe_build.py 2 10 50 > gpurun_out/s3_build_$v.log 2>&1; done
timeout 300 python tools/probe_perf.py sun sky > gpurun_out/s3_perf_base.log 2>&1
QSMRT_LIB=build/variants/libqsmrt_m3s1.so timeout 300 python tools/probe_perf.py sun > gpurun_out/s3_perf_m3s1.log 2>&1
QSMRT_LIB=build/variants/libqsmrt_m4s1.so timeout 300 python tools/probe_perf.py sky > gpurun_out/s3_perf_m4s1.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,sm__cycles_active.avg,sm__cycles_active.max,sm__cycles_elapsed.max,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct \
   --clock-control none -k regex:"k_trace5|k_cast_rays|k_count" --csv --log-file gpurun_out/s3_c1_ncu.csv python tools/probe_c1.py 2 > gpurun_out/s3_c1.log 2>&1
timeout 600 python tools/profile_angles.py --counters gpurun_out/s3_angle_counters.json > gpurun_out/s3_angles.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum \
   --clock-control none -k regex:k_trace5 --csv --log-file gpurun_out/s3_angle_launches.csv python tools/profile_angles.py > gpurun_out/s3_angles_ncu.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/s3_bench.json 2> gpurun_out/s3_bench.err; echo "bench rc=$?"
cat gpurun_out/s3_build_*.log gpurun_out/s3_perf_*.log gpurun_out/s3_c1.log gpurun_out/s3_angles.log | grep -v "^+"
