#!/bin/bash
# multi-GPU session (run with gpurun --gpus N): NCCL test, bench at N, C4 at N
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
nvidia-smi topo -m > gpurun_out/mg_topo_n$N.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_api.py -m gpu -q -k "two_gpu or current_device" > gpurun_out/mg_pytest_n$N.log 2>&1; tail -3 gpurun_out/mg_pytest_n$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/mg_bench_n$N.json 2> gpurun_out/mg_bench_n$N.err; echo "bench rc=$?"
tail -c 400 gpurun_out/mg_bench_n$N.err
C4_CANOPIES=${C4_CANOPIES:-25} timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 tests/measure/run_c4_multi.py > gpurun_out/mg_c4_n$N.json 2> gpurun_out/mg_c4_n$N.err; echo "c4 rc=$?"
tail -c 300 gpurun_out/mg_c4_n$N.err; cat gpurun_out/mg_c4_n$N.json
