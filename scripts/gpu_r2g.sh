#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -k "queries or fused" 2>&1 | tail -8
timeout 300 python scripts/stage_probe.py queries fused
export PROBE_WARM=1 PROBE_ITERS=1
CMD="python scripts/stage_probe.py queries fused"
$CMD > gpurun_out/plain_stage.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"ad2_queries|fused_" -c 6 -o gpurun_out/prof_stage5 -f $CMD > gpurun_out/ncu_stage5.log 2>&1
echo ncu_rc=$?
