#!/bin/bash
set -x
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -x 2>&1 | tail -8
timeout 300 python scripts/stage_probe.py facts fused
