#!/bin/bash
set -x
timeout 1200 python -m pytest tests/test_gpu_guard_bands.py -m gpu -q --timeout 600 2>&1 | tail -30
