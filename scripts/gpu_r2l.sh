#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 -k "facts or epoch or stages or model or dropin" 2>&1 | tail -30 > gpurun_out/r2l_pytest.log; tail -12 gpurun_out/r2l_pytest.log
timeout 300 python scripts/stage_probe.py facts
