#!/bin/bash
# N=8 after the per-level VAE stages: the driver's launch line at 480p, a phase trace, and 720p
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541"
O=gpurun_out
timeout 300 $TR bench.py --gpus 8 --steps 10 --warmup 3 > $O/r02_final_bench_n8.json 2> $O/r02_final_bench_n8.err; cut -c1-330 $O/r02_final_bench_n8.json
WF_TRACE=1 timeout 200 $TR bench.py --gpus 8 --steps 4 --warmup 3 --no-e2e > $O/r02_final_bench_n8_trace.json 2> $O/r02_final_bench_n8_trace.err; grep trace $O/r02_final_bench_n8_trace.err; cut -c1-200 $O/r02_final_bench_n8_trace.json
timeout 300 $TR bench.py --gpus 8 --height 720 --width 1280 --steps 4 --warmup 3 --no-e2e > $O/r02_final_bench_720p_n8.json 2> $O/r02_final_bench_720p_n8.err; cut -c1-330 $O/r02_final_bench_720p_n8.json
