#!/bin/bash
# Round-2 GPU visit B: full GPU tests (no -x), stage timings, ncu captures of the stage kernels.
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -60 > gpurun_out/r2b_pytest.log; tail -30 gpurun_out/r2b_pytest.log
timeout 300 python scripts/stage_probe.py > gpurun_out/r2b_stage.json 2> gpurun_out/r2b_stage.err; echo stage_rc=$?; cat gpurun_out/r2b_stage.json; tail -3 gpurun_out/r2b_stage.err
CMD="python scripts/stage_probe.py fused lt gumbel facts elbo"
export PROBE_WARM=1 PROBE_ITERS=1
$CMD > gpurun_out/plain_stage.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"fused_|lt_|gumbel_|facts_tile|elbo_terms" -c 20 -o gpurun_out/prof_stage -f $CMD > gpurun_out/ncu_stage.log 2>&1
echo ncu_rc=$?; tail -3 gpurun_out/ncu_stage.log
