#!/bin/bash
# ncu capture of the lane<->row Gumbel kernels (after the same command ran clean without ncu)
set -x
mkdir -p gpurun_out
CMD="python scripts/stage_probe.py gumbel"
export PROBE_WARM=1 PROBE_ITERS=2
timeout 300 $CMD > gpurun_out/plain_gumbel.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gumbel_rows" -c 4 -o gpurun_out/prof_gumbel_rows -f $CMD > gpurun_out/ncu_gumbel.log 2>&1
echo ncu_rc=$?
tail -3 gpurun_out/ncu_gumbel.log
