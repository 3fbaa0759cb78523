#!/bin/bash
# compute-sanitizer over the CI-size GPU tests (memcheck + racecheck on every kernel).  NOT run in rounds 1-2
# (the GPU budget went to parity, benches and ncu); the tcgen05 / TMA kernels carry a 4-second barrier watchdog instead.
# Usage on a B200 box:  bash tools/sanitize.sh [memcheck|racecheck|synccheck]   (expect 20-100x slowdown; toy shapes only)
tool=${1:-memcheck}
mkdir -p gpurun_out
compute-sanitizer --tool "$tool" --error-exitcode 9 --launch-timeout 120 \
  python -m pytest tests/test_dit_kernels_gpu.py tests/test_sampler_ops_gpu.py tests/test_flow_gpu.py tests/test_inputs.py \
         tests/test_vae_gpu.py tests/test_bsa_gpu.py -m gpu -x -q > "gpurun_out/sanitize_$tool.log" 2>&1
echo "compute-sanitizer $tool rc=$?"; tail -n 15 "gpurun_out/sanitize_$tool.log"
