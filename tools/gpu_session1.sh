#!/bin/bash
# first GPU session: kernel parity tests, then a perf probe
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; free -g >> gpurun_out/gpu.txt
timeout 600 python -m pytest tests/test_sampler_ops_gpu.py -m gpu -x -q 2>&1 | tail -30 > gpurun_out/t_sampler.log
timeout 900 python -m pytest tests/test_dit_kernels_gpu.py -m gpu -q 2>&1 | tail -120 > gpurun_out/t_kernels.log
timeout 600 python -m pytest tests/test_dit_forward_gpu.py -m gpu -q 2>&1 | tail -60 > gpurun_out/t_forward.log
timeout 600 python tools/perf_probe.py > gpurun_out/perf_probe.log 2>&1
tail -5 gpurun_out/t_sampler.log; tail -40 gpurun_out/t_kernels.log; tail -15 gpurun_out/t_forward.log; tail -20 gpurun_out/perf_probe.log
