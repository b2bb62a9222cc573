#!/bin/bash
set -x
timeout 1500 python -m pytest tests/test_gpu_stages.py tests/test_gpu_guard_bands.py tests/test_gpu_model.py -m gpu -q --timeout 900 2>&1 | tail -5
timeout 300 python scripts/stage_probe.py gumbel
