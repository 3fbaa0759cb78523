#!/bin/bash
# ncu --set full of the self-attention kernel, variants 7 (128-key blocks) and 4 (64-key blocks), with source-level stall sampling
mkdir -p gpurun_out
for v in 7 4; do
  WF_ATTN=$v WF_ITERS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_tcgen05 -s 1 -c 1 -f -o gpurun_out/attn_v$v \
    python tools/attn_probe.py > gpurun_out/ncu_attn_v$v.log 2>&1
  echo "== ncu v$v rc=$?"; tail -2 gpurun_out/ncu_attn_v$v.log | cut -c1-200
done
ls -la gpurun_out/*.ncu-rep
