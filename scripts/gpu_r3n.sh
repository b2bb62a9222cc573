#!/bin/bash
set -x
timeout 1500 python -m pytest tests/test_gpu_fused.py tests/test_gpu_model.py tests/test_gpu_guard_bands.py tests/test_gpu_stages.py -m gpu -q --timeout 600 2>&1 | tail -8
timeout 300 python scripts/stage_probe.py fused gumbel
VAEL_GUMBEL_ROWS=0 timeout 300 python scripts/stage_probe.py gumbel
