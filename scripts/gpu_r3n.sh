#!/bin/bash
set -x
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -12
timeout 300 python scripts/stage_probe.py fused facts
