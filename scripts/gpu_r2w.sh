#!/bin/bash
set -x
timeout 1200 python -m pytest tests/test_gpu_stages.py tests/test_gpu_guard_bands.py tests/test_gpu_model.py -m gpu -q --timeout 600 2>&1 | tail -12
timeout 300 python scripts/sanitize_case.py 2>&1 | tail -2
timeout 300 python scripts/gumbel_small_batch.py
