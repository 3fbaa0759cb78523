#!/bin/bash
# round-end evidence: GPU tests, launch list of one guided step, ncu --set full of the default attention kernel and of a
# full-resolution 96-channel VAE convolution, VAE probe
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 7000 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 3 --warmup 0 --no-e2e --no-cpu-baseline > gpurun_out/launches_run.log 2>&1; echo "launch list rc=$?"
WF_ITERS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_tcgen05 -s 1 -c 1 -f -o gpurun_out/attn_v3_full \
    python tools/attn_probe.py > gpurun_out/ncu_attn_v3.log 2>&1; echo "attn ncu rc=$?"
WF_F=9 timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv333_halo -s 24 -c 1 -f -o gpurun_out/conv_multi_full \
    python tools/vae_probe.py > gpurun_out/ncu_conv_multi.log 2>&1; echo "conv ncu rc=$?"
timeout 300 python tools/vae_probe.py > gpurun_out/vae_probe.log 2>&1; echo "vae probe rc=$?"
ls gpurun_out | tail -20
