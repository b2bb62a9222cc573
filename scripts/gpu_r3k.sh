#!/bin/bash
set -x
mkdir -p gpurun_out
CMD="python scripts/stage_probe.py fused"
export PROBE_WARM=1 PROBE_ITERS=2
timeout 300 $CMD > gpurun_out/plain_fused3.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:"fused_" -c 6 -o gpurun_out/prof_fused3 -f $CMD > gpurun_out/ncu_fused3.log 2>&1
echo ncu_rc=$?
