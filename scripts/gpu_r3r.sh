#!/bin/bash
set -x
timeout 1500 python -m pytest tests/test_gpu_circuit.py tests/test_gpu_guard_bands.py tests/test_gpu_model.py -m gpu -q --timeout 900 2>&1 | tail -4
timeout 120 python scripts/quick_bench.py 8388608 | head -1
timeout 120 python scripts/quick_bench.py 262144 | head -1; VAEL_AD2_FWD_ONCE=0 VAEL_AD2_BWD_ONCE=0 timeout 120 python scripts/quick_bench.py 262144 | head -1
timeout 120 python scripts/quick_bench.py 524288 | head -1; VAEL_AD2_FWD_ONCE=0 VAEL_AD2_BWD_ONCE=0 timeout 120 python scripts/quick_bench.py 524288 | head -1
