#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err; echo bench_rc=$?; tail -3 gpurun_out/r2k_bench.err
export PROBE_WARM=1 PROBE_ITERS=1
CMD="python scripts/stage_probe.py jit jitwide generic"
$CMD > gpurun_out/plain_stage.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"jit_|generic_" -c 14 -o gpurun_out/prof_jit -f $CMD > gpurun_out/ncu_jit.log 2>&1
echo ncu_rc=$?
