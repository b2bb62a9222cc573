#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -k "jit" -x 2>&1 | tail -40 > gpurun_out/r2i_pytest.log; tail -25 gpurun_out/r2i_pytest.log
timeout 600 python scripts/stage_probe.py jitwide
