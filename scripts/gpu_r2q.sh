#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python scripts/sanitize_case.py > gpurun_out/sanitize_plain.log 2>&1; echo plain_rc=$?; tail -3 gpurun_out/sanitize_plain.log
