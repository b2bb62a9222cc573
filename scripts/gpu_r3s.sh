#!/bin/bash
set -x
timeout 600 python scripts/stage_probe.py jit jitwide 2>&1 | tail -1
VAEL_JIT_PERSISTENT=1 timeout 600 python scripts/stage_probe.py jit jitwide 2>&1 | tail -1
timeout 1500 python -m pytest tests/test_gpu_jit.py tests/test_gpu_random_circuits.py tests/test_gpu_guard_bands.py -m gpu -q --timeout 900 2>&1 | tail -4
