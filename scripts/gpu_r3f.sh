#!/bin/bash
set -x
timeout 1500 python -m pytest tests/test_gpu_circuit.py tests/test_gpu_guard_bands.py tests/test_gpu_model.py tests/test_gpu_jit.py tests/test_gpu_reference_dropin.py -m gpu -q --timeout 900 2>&1 | tail -6
