#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 200 python scripts/e2e_probe.py
VAEL_HOST_FIXED_CHUNKS=1 timeout 200 python scripts/e2e_probe.py
timeout 200 python scripts/e2e_probe.py
timeout 600 python -m pytest tests -m gpu -q --timeout 600 -x -k "host or e2e or capi or thread" 2>&1 | tail -5
timeout 300 python scripts/sanitize_case.py 2>&1 | tail -1
