#!/bin/bash
# attention variants: parity tests + clock-independent comparison (elapsed SM cycles and tensor-pipe activity from a light ncu pass)
mkdir -p gpurun_out
for v in ${VARIANTS:-4 7 8 9}; do
  WF_ATTN=$v timeout 600 python -m pytest tests/test_dit_kernels_gpu.py -m gpu -q -k "attention" 2>&1 | tail -1
  WF_ATTN=$v WF_ITERS=2 timeout 600 ncu --metrics sm__cycles_elapsed.max,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,gpu__time_duration.sum \
     --clock-control none -k regex:attention_tcgen05 -s 1 -c 2 --csv python tools/attn_probe.py 2>/dev/null | grep -E "sm__cycles_elapsed.max|tensor_cycles|pipe_xu|gpu__time" | awk -F'","' -v v=$v '{printf "v%s %s %s\n", v, $(NF-2), $NF}' | tr -d '"'
done
