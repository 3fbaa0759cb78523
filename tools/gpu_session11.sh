#!/bin/bash
# N-GPU: multi-GPU parity check + the bench exactly as the driver launches it
N=${1:-8}
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/ulysses_check.py > gpurun_out/mgpu_check_n$N.log 2>&1
echo "== multi-gpu check rc=$?"; grep -E "rank 0/|Error|error" gpurun_out/mgpu_check_n$N.log | sort | head -8 | cut -c1-250
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps ${STEPS:-10} --warmup ${WARMUP:-3} > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "== bench rc=$?"; cut -c1-1500 gpurun_out/bench_n$N.json; tail -3 gpurun_out/bench_n$N.err | cut -c1-300
