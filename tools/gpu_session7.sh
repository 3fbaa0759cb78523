#!/bin/bash
# attention variants: 4 (64-key blocks, barrier fix) vs 7/8/9 (128-key blocks, 0 / 25% / 37.5% polynomial ex2); BSA tests after the barrier fix
mkdir -p gpurun_out
for v in 4 7 8 9; do
  WF_ATTN=$v timeout 600 python -m pytest tests/test_dit_kernels_gpu.py -m gpu -q -k "attention" 2>&1 | tail -4 > gpurun_out/t_attn_v$v.log
  echo "== tests WF_ATTN=$v"; tail -3 gpurun_out/t_attn_v$v.log
  WF_ATTN=$v timeout 300 python tools/attn_probe.py > gpurun_out/attn_probe_v$v.log 2>&1; tail -1 gpurun_out/attn_probe_v$v.log | cut -c1-400
done
timeout 900 python -m pytest tests/test_bsa_gpu.py -m gpu -x -q 2>&1 | tail -4 > gpurun_out/t_bsa.log; echo "== bsa tests"; tail -3 gpurun_out/t_bsa.log
