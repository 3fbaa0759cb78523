#!/bin/bash
# last call of round 2 (79 s of budget left): the bench line after the config refactor on a tiny configuration, and the
# resize branch of fuse_latents on the device
mkdir -p gpurun_out
WF_CPU_GRID=1x10x13 WF_CPU_REPS=1 timeout 45 python bench.py --steps 1 --warmup 1 --layers 1 --frames 9 --no-e2e --no-gpu-reference \
    > gpurun_out/r02_last_bench.json 2> gpurun_out/r02_last_bench.err
echo "bench rc=$?"; head -c 1500 gpurun_out/r02_last_bench.json
timeout 25 python tools/gpu_last_check.py > gpurun_out/r02_last_check.log 2>&1
echo "check rc=$?"; tail -n 3 gpurun_out/r02_last_check.log
