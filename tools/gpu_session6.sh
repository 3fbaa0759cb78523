#!/bin/bash
# N-GPU session: VAE row-sharding parity (simulated ranks on one GPU, then real ranks over NCCL), Ulysses check, bench at N GPUs
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus.txt
timeout 600 python -m pytest tests/test_vae_gpu.py -m gpu -x -q 2>&1 | tail -5 > gpurun_out/t_vae.log; echo "== vae tests"; tail -3 gpurun_out/t_vae.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/ulysses_check.py > gpurun_out/mgpu_check_n$N.log 2>&1
echo "== multi-gpu check rc=$?"; grep -E "rank [0-9]" gpurun_out/mgpu_check_n$N.log | sort | head -30
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "== bench rc=$?"; cat gpurun_out/bench_n$N.json; tail -5 gpurun_out/bench_n$N.err
