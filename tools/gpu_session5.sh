#!/bin/bash
# step breakdown at the benchmark size + the UMMA start-offset probe
mkdir -p gpurun_out
nproc > gpurun_out/host.txt; free -g >> gpurun_out/host.txt
timeout 120 build/umma_probe > gpurun_out/umma_probe.log 2>&1; echo "== umma probe rc=$?"; grep -c OK gpurun_out/umma_probe.log; grep -v OK gpurun_out/umma_probe.log | head -20
timeout 900 python tools/step_breakdown.py > gpurun_out/step_breakdown.log 2>&1; echo "== breakdown rc=$?"; tail -20 gpurun_out/step_breakdown.log
