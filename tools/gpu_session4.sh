#!/bin/bash
# BSA / refine-pass tests, BSA probe at config-5 shape, attention softmax variants (4 = default, 5/6 = +polynomial ex2)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_bsa_gpu.py tests/test_longcat_refine_gpu.py -m gpu -x -q 2>&1 | tail -25 > gpurun_out/t_bsa.log
echo "== bsa tests"; tail -25 gpurun_out/t_bsa.log
timeout 600 python tools/bsa_probe.py > gpurun_out/bsa_probe.log 2>&1; echo "== bsa probe rc=$?"; cat gpurun_out/bsa_probe.log | cut -c1-260
for v in 4 5 6; do
  WF_ATTN=$v timeout 300 python tools/perf_probe.py > gpurun_out/perf_probe_v$v.log 2>&1
  echo "== perf WF_ATTN=$v"; grep -E "self_attn|flash_attn2" gpurun_out/perf_probe_v$v.log | cut -c1-200
done
for v in 5 6; do
  WF_ATTN=$v timeout 600 python -m pytest tests/test_dit_kernels_gpu.py tests/test_dit_forward_gpu.py -m gpu -q -k "attention or forward" 2>&1 | tail -3 > gpurun_out/t_attn_v$v.log
  echo "== tests WF_ATTN=$v"; tail -2 gpurun_out/t_attn_v$v.log
done
