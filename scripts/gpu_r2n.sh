#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_stages.py tests/test_gpu_model.py tests/test_gpu_fused.py -m gpu -q --timeout 600 -x 2>&1 | tail -30
timeout 300 python scripts/stage_probe.py gumbel
VAEL_GUMBEL_ROWS=0 timeout 300 python scripts/stage_probe.py gumbel
