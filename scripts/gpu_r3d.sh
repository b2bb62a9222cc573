#!/bin/bash
# ncu capture of the column-chunked generated kernels (3-digit circuit) with 2-D tensor-map box copies
set -x
mkdir -p gpurun_out
CMD="python scripts/stage_probe.py jitwide"
export PROBE_WARM=1 PROBE_ITERS=2
timeout 300 $CMD > gpurun_out/plain_jitwide.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:"jit_" -c 4 -o gpurun_out/prof_jitwide -f $CMD > gpurun_out/ncu_jitwide.log 2>&1
echo ncu_rc=$?
