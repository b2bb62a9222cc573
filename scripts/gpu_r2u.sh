#!/bin/bash
# round-2 launch lists: the bench command and one full train step (fused logic kernel in it)
set -x
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-train --no-cpu --no-e2e --no-secondary"
timeout 300 $CMD > gpurun_out/plain_r2.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2.csv $CMD > gpurun_out/ncu_list_r2.log 2>&1
echo ncu_list_rc=$?
SCMD="python scripts/step_launches.py 4096 cl"
timeout 300 $SCMD > gpurun_out/plain_step_r2.log 2>&1 && timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/step_launches_r2.csv $SCMD > gpurun_out/ncu_step_r2.log 2>&1
echo ncu_step_rc=$?
