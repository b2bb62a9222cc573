#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -40 > gpurun_out/r2e_pytest.log; tail -8 gpurun_out/r2e_pytest.log
timeout 300 python scripts/stage_probe.py lt queries
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -5
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; echo bench_rc=$?; tail -3 gpurun_out/r2e_bench.err
