#!/bin/bash
# Round-2 first GPU visit: all GPU tests (verbose on failure), smoke, bench both arms.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -40 > gpurun_out/r2a_pytest.log; tail -40 gpurun_out/r2a_pytest.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -8 | tee gpurun_out/r2a_smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo bench_rc=$?; tail -5 gpurun_out/r2a_bench.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2a_bench_ref.json 2> gpurun_out/r2a_ref.err; echo ref_rc=$?; tail -3 gpurun_out/r2a_ref.err
