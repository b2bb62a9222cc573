#!/bin/bash
set -x
timeout 600 python scripts/stage_probe.py jitwide 2>&1 | tail -1
VAEL_ADW_ONCE=0 timeout 600 python scripts/stage_probe.py jitwide 2>&1 | tail -1
for w in 1 3 4; do VAEL_ADW_FWD_WARPS=$w VAEL_ADW_BWD_WARPS=$w timeout 600 python scripts/stage_probe.py jitwide 2>&1 | tail -1; done
timeout 900 python -m pytest tests/test_gpu_wide.py tests/test_gpu_guard_bands.py -m gpu -q --timeout 900 2>&1 | tail -4
