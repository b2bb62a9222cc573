#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -k "fused or literal or dropin or clipped or elbo or query or mario" 2>&1 | tail -80 > gpurun_out/r2c_pytest.log; tail -15 gpurun_out/r2c_pytest.log
for mb in 4 1 3 5; do VAEL_FUSED_BWD_MINB=$mb timeout 300 python scripts/stage_probe.py fused > gpurun_out/r2c_fused_mb$mb.json 2>/dev/null; echo mb=$mb; cat gpurun_out/r2c_fused_mb$mb.json; done
VAEL_FUSED_GENERAL=1 timeout 300 python scripts/stage_probe.py fused; echo
timeout 300 python scripts/stage_probe.py lt facts
export PROBE_WARM=1 PROBE_ITERS=1
CMD="python scripts/stage_probe.py fused lt"
$CMD > gpurun_out/plain_stage.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"fused_|lt_" -c 8 -o gpurun_out/prof_stage2 -f $CMD > gpurun_out/ncu_stage2.log 2>&1
echo ncu_rc=$?
