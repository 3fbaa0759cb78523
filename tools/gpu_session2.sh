#!/bin/bash
mkdir -p gpurun_out
for v in 2 3; do
  WF_ATTN=$v timeout 600 python -m pytest tests/test_dit_kernels_gpu.py tests/test_dit_forward_gpu.py -m gpu -q -k "attention or forward" 2>&1 | tail -5 > gpurun_out/t_attn_v$v.log
  echo "== WF_ATTN=$v"; tail -3 gpurun_out/t_attn_v$v.log
done
timeout 300 python -m pytest tests/test_dit_kernels_gpu.py -m gpu -q -k gemm 2>&1 | tail -3
for v in 1 2 3; do
  WF_ATTN=$v timeout 300 python tools/perf_probe.py > gpurun_out/perf_probe_v$v.log 2>&1
  echo "== perf WF_ATTN=$v"; grep -E "self_attn|cross512|qkv|ffn0|o_resid|ffn2" gpurun_out/perf_probe_v$v.log | cut -c1-200
done
timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum --clock-control none -k regex:"gemm_bf16|attention" -c 8 python tools/perf_probe.py 2>&1 | grep -E "gemm_bf16|attention_tc|dram__|tensor_cycles|gpu__time" | head -60 > gpurun_out/ncu_quick.log
cat gpurun_out/ncu_quick.log
