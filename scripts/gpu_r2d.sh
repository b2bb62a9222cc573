#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -k "literal or queries or mario or wide" 2>&1 | tail -40 > gpurun_out/r2d_pytest.log; tail -12 gpurun_out/r2d_pytest.log
timeout 300 python scripts/stage_probe.py lt queries
export PROBE_WARM=1 PROBE_ITERS=1
CMD="python scripts/stage_probe.py lt queries"
$CMD > gpurun_out/plain_stage.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"lt_|queries" -c 8 -o gpurun_out/prof_stage3 -f $CMD > gpurun_out/ncu_stage3.log 2>&1
echo ncu_rc=$?
