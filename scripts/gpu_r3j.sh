#!/bin/bash
set -x
timeout 1500 python -m pytest tests/test_gpu_fused.py tests/test_gpu_model.py tests/test_gpu_guard_bands.py -m gpu -q --timeout 600 2>&1 | tail -25
timeout 300 python scripts/stage_probe.py fused
timeout 300 python scripts/fused_small_batch.py
