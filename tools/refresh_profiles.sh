#!/bin/bash
# turns the raw outputs of tools/gpu_profile_session.sh (gpurun_out/s15_*) into the committed files under profiles/
set -e
cd "$(dirname "$0")/.."
O=gpurun_out
cp $O/s15_bench_n1.json profiles/r02_bench_n1.json
cp $O/s15_bench_ref.json profiles/r02_bench_ref.json
cp $O/s15_final_launches.csv profiles/r02_final_launches.csv
python tools/ncu_launches.py $O/s15_final_launches.csv > profiles/r02_final_launches_summary.txt
{ echo "ncu --set full --clock-control none --import-source on -k regex:k_trace5 -s 4 -c 1 python bench.py --steps 2 --warmup 3 --no-cpu-baseline  (B200, final kernel of round 2:"
  echo "32-byte quantised nodes, 4 node steps per vote, one-register leaf parking, slab widening, 10 CTAs per SM); summarised by tools/ncu_summary.py"
  python tools/ncu_summary.py $O/s15_trace5_raw.csv $O/s15_trace5_src.csv; } > profiles/r02_k_trace5_ncu.txt
{ echo "tools/ncu_regions.py on the source page of the same capture: SASS lines grouped into regions of equal execution count"
  echo "(= loop bodies) with their share of the issued warp instructions and average active lanes.  The ~135 + 64 + 65 line"
  echo "regions at 24-26 lanes are the node phase (4 unrolled node steps of 64-67 instructions + the phase vote), the two 52-line"
  echo "regions at ~11 lanes the triangle phase (two tests per vote), the rest retire / refill."
  python tools/ncu_regions.py $O/s15_trace5_src.csv 0.8; } > profiles/r02_k_trace5_regions.txt
python tools/profile_angles.py --merge $O/s15_angle_counters.json $O/s15_angle_launches.csv profiles/r02_cast_rays_profile.json
cp $O/s15_configs.json profiles/r02_configs.json
{ echo "tests/measure/run_configs.py c1 c3 c4 under ncu (k_trace5 launches only; --clock-control none): which resource each"
  echo "configuration loads.  Launch order: C1 cast x5+, C1 count, C3 count (10 launches of 10M rays), C4 cast."
  python tools/ncu_table.py $O/s15_configs_ncu.csv; } > profiles/r02_configs_ncu.txt
grep -v "^+" $O/s15_small.log > profiles/r02_reference_call_shapes.txt
grep -v "^+" $O/s15_perf.log > profiles/r02_probe_perf.txt
echo refreshed
