#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 -k "jit or random" -x 2>&1 | tail -40 > gpurun_out/r2h_pytest.log; tail -25 gpurun_out/r2h_pytest.log
timeout 300 python scripts/stage_probe.py jit
