#!/bin/bash
# final numbers of the round: bench (both arms), launch list and full ncu capture of the headline kernels
set -x
mkdir -p gpurun_out
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r3v_ref.json 2> gpurun_out/r3v_ref.err; echo ref_rc=$?
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r3v_bench.json 2> gpurun_out/r3v_bench.err; echo bench_rc=$?; tail -2 gpurun_out/r3v_bench.err
CMD="python bench.py --steps 2 --warmup 3 --no-train --no-cpu --no-e2e --no-secondary"
timeout 300 $CMD > gpurun_out/plain_r3v.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r3v.csv $CMD > gpurun_out/ncu_list_r3v.log 2>&1
echo ncu_list_rc=$?
timeout 300 $CMD > gpurun_out/plain2_r3v.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:ad2_ -s 6 -c 2 -o gpurun_out/prof_ad2_r3v -f $CMD > gpurun_out/ncu_full_r3v.log 2>&1
echo ncu_full_rc=$?
