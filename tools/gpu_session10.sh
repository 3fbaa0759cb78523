#!/bin/bash
# N-GPU: peer-memory Ulysses check, then bench A/B (NCCL all-to-all vs peer-memory exchange) on the same box
N=${1:-2}
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/ulysses_check.py > gpurun_out/mgpu_check_n$N.log 2>&1
echo "== multi-gpu check rc=$?"; grep -E "rank [0-9]|Error|error" gpurun_out/mgpu_check_n$N.log | sort | head -20 | cut -c1-250
for mode in nccl peer; do
  WF_ULYSSES=$mode timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps ${STEPS:-6} --warmup 3 --no-e2e > gpurun_out/bench_n${N}_$mode.json 2> gpurun_out/bench_n${N}_$mode.err
  echo "== bench $mode rc=$?"; python -c "
import json,sys
d=json.loads(open('gpurun_out/bench_n${N}_$mode.json').read())
print('$mode', round(d['ms_per_step'],1), 'ms/step', round(d['value'],4), 'steps/s attn', round(d['roofline']['mean_launch_ms'],2), 'ms clk', d['clocks']['sm_mhz'])" 2>&1 | tail -1
  tail -2 gpurun_out/bench_n${N}_$mode.err | cut -c1-300
done
