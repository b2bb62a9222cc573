#!/bin/bash
# full verification of the tree: GPU suite, smoke, both bench arms
set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -40 > gpurun_out/r2m_pytest.log; tail -8 gpurun_out/r2m_pytest.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -12
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2m_ref.json 2> gpurun_out/r2m_ref.err; echo ref_rc=$?; cat gpurun_out/r2m_ref.json | cut -c1-600
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err; echo bench_rc=$?; tail -3 gpurun_out/r2m_bench.err
