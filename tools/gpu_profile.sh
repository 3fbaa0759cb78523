#!/bin/bash
# ncu captures for profiles/: (1) launch list of a guided + two plain steps of the bench, (2) full sets of the top kernels
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 7000 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 3 --warmup 0 --no-e2e --no-cpu-baseline > gpurun_out/launches_run.log 2>&1
WF_L=32760 timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_tcgen05 -s 1 -c 1 -o gpurun_out/attn_full \
    python tools/perf_probe.py > gpurun_out/ncu_attn.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tcgen05 -s 2 -c 1 -o gpurun_out/gemm_full \
    python tools/perf_probe.py > gpurun_out/ncu_gemm.log 2>&1
WF_F=17 timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tf32 -s 60 -c 1 -o gpurun_out/conv_full \
    python tools/vae_probe.py > gpurun_out/ncu_conv.log 2>&1
timeout 600 python tools/vae_probe.py > gpurun_out/vae_probe.log 2>&1
ls -la gpurun_out | tail -15; tail -40 gpurun_out/vae_probe.log
