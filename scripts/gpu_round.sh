#!/bin/bash
# One GPU visit: parity tests, smoke, bench, then the ncu launch list and one full capture.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -30
python __graft_entry__.py --smoke 2>&1 | tail -8
python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo bench_rc=$?; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
if [ "$1" == "ncu" ]; then
  CMD="python bench.py --steps 2 --warmup 3 --no-train --no-cpu --no-e2e"
  $CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
  echo ncu_list_rc=$?
  $CMD > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:ad2_ -s 6 -c 2 -o gpurun_out/prof_ad2 -f $CMD > gpurun_out/ncu_full.log 2>&1
  echo ncu_full_rc=$?
  tail -3 gpurun_out/ncu_full.log
fi
