#!/bin/bash
# round 2, 2-GPU call: the sharded path on real peers (NCCL default group, gather + nccl exchanges) and the N=2 bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2c_gpus.txt
timeout 900 python -m pytest tests/test_gpu_distributed.py -m gpu -x -q -p no:cacheprovider > gpurun_out/r2c_pytest_dist.log 2>&1; echo "pytest dist rc=$?"; tail -n 8 gpurun_out/r2c_pytest_dist.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2c_bench_n2.json 2> gpurun_out/r2c_bench_n2.err; echo "bench n2 rc=$?"; tail -c 1500 gpurun_out/r2c_bench_n2.err; head -c 6000 gpurun_out/r2c_bench_n2.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 3 --warmup 3 --exchange nccl > gpurun_out/r2c_bench_n2_nccl.json 2> gpurun_out/r2c_bench_n2_nccl.err; echo "bench n2 nccl rc=$?"; tail -c 600 gpurun_out/r2c_bench_n2_nccl.err; head -c 1500 gpurun_out/r2c_bench_n2_nccl.json
