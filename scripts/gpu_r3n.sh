#!/bin/bash
set -x
timeout 300 python scripts/stage_probe.py fused
VAEL_FUSED_TILE_WARPS=1 timeout 300 python scripts/stage_probe.py fused
timeout 1500 python -m pytest tests/test_gpu_fused.py tests/test_gpu_guard_bands.py tests/test_gpu_model.py -m gpu -q --timeout 900 2>&1 | tail -5
