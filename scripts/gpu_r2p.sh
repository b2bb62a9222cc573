#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 200 python scripts/e2e_probe.py
VAEL_HOST_FIXED_CHUNKS=1 VAEL_HOST_CHUNK_ROWS=524288 timeout 200 python scripts/e2e_probe.py
VAEL_HOST_FIXED_CHUNKS=1 VAEL_HOST_CHUNK_ROWS=196608 timeout 200 python scripts/e2e_probe.py
timeout 600 python -m pytest tests -m gpu -q --timeout 600 -x -k "host or e2e or capi or thread" 2>&1 | tail -5
