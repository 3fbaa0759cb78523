#!/bin/bash
# Round-2 evidence on 8 GPUs of one box (gpurun --gpus 8): real-rank parity of every sharded path, the 480p and 720p benches
# in the CFG x Ulysses layout (and the plain Ulysses layout for A/B), and a phase trace of the timed steps.
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531"
O=gpurun_out
timeout 300 $TR tools/ulysses_check.py > $O/r02_mgpu_check_n8.log 2>&1; grep -c "True" $O/r02_mgpu_check_n8.log; grep -E "False|Error" $O/r02_mgpu_check_n8.log | head -5
timeout 300 $TR bench.py --gpus 8 --steps 10 --warmup 3 > $O/r02_bench_n8_cfg.json 2> $O/r02_bench_n8_cfg.err; cut -c1-330 $O/r02_bench_n8_cfg.json
WF_TRACE=1 timeout 200 $TR bench.py --gpus 8 --steps 4 --warmup 3 --no-e2e > $O/r02_bench_n8_trace.json 2> $O/r02_bench_n8_trace.err; grep trace $O/r02_bench_n8_trace.err
WF_LAYOUT=ulysses timeout 200 $TR bench.py --gpus 8 --steps 4 --warmup 3 --no-e2e > $O/r02_bench_n8_uly.json 2> $O/r02_bench_n8_uly.err; cut -c1-330 $O/r02_bench_n8_uly.json
timeout 400 $TR bench.py --gpus 8 --height 720 --width 1280 --steps 4 --warmup 3 > $O/r02_bench_720p_n8.json 2> $O/r02_bench_720p_n8.err; cut -c1-1200 $O/r02_bench_720p_n8.json; tail -2 $O/r02_bench_720p_n8.err
