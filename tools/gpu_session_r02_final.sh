#!/bin/bash
# Round-2 single-GPU evidence: the GPU test-suite, the bench, the ncu launch list of a guided + two plain steps and
# `ncu --set full` of the 3x3x3 VAE convolution (row epilogue) and of the CTA-pair GEMM.
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > $O/r02_final_gputests.log 2>&1; tail -3 $O/r02_final_gputests.log | cut -c1-200
timeout 600 python bench.py --steps 4 --warmup 3 > $O/r02_final_bench_n1.json 2> $O/r02_final_bench_n1.err; cut -c1-300 $O/r02_final_bench_n1.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 7500 --csv --log-file $O/r02_launches.csv \
    python bench.py --steps 3 --warmup 0 --no-e2e --no-cpu-baseline --no-gpu-reference > $O/r02_launches_run.log 2>&1; echo "launch list rc=$?"
WF_F=9 timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv333_halo -s 24 -c 1 -f -o $O/r02_conv_rows_full \
    python tools/vae_probe.py > $O/r02_ncu_conv.log 2>&1; echo "conv ncu rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tcgen05_pair -s 2 -c 1 -f -o $O/r02_gemm_pair_full \
    python tools/perf_probe.py > $O/r02_ncu_gemm.log 2>&1; echo "gemm ncu rc=$?"
ls -la $O | tail -8
