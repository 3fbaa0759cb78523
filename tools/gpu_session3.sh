#!/bin/bash
# full GPU suite + smoke + default bench + refreshed ncu captures (attention v4 default, a real conv launch)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/t_all.log
echo "== tests"; tail -6 gpurun_out/t_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "== smoke rc=$?"; tail -3 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "== bench rc=$?"; cat gpurun_out/bench_default.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "== ref rc=$?"; cat gpurun_out/bench_ref.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 7000 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 3 --warmup 0 --no-e2e --no-cpu-baseline > gpurun_out/launches_run.log 2>&1
WF_L=32760 timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_tcgen05 -s 1 -c 1 -f -o gpurun_out/attn_full \
    python tools/perf_probe.py > gpurun_out/ncu_attn.log 2>&1
WF_F=5 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum --clock-control none -k regex:conv_tf32 -c 200 --csv --log-file gpurun_out/conv_list.csv \
    python tools/vae_probe.py > gpurun_out/ncu_conv.log 2>&1
ls -la gpurun_out | tail -12
