#!/bin/bash
# One GPU visit: parity tests, smoke, bench, then (arg "ncu") the ncu launch list and one full capture of the
# headline kernels.  Everything lands in gpurun_out/; scripts/ncu_summary.py turns it into profiles/.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -6
python __graft_entry__.py --smoke 2>&1 | tail -6
python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo bench_rc=$?; tail -3 gpurun_out/bench.err
python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err; echo ref_rc=$?
if [ "$1" == "ncu" ]; then
  CMD="python bench.py --steps 2 --warmup 3 --no-train --no-cpu --no-e2e --no-secondary"
  $CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
  echo ncu_list_rc=$?
  $CMD > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:ad2_ -s 6 -c 2 -o gpurun_out/prof_ad2 -f $CMD > gpurun_out/ncu_full.log 2>&1
  echo ncu_full_rc=$?
  # the generic CSR kernels (2-digit circuit, VAEL_FLAG_FORCE_GENERIC) and the all-queries kernel
  GCMD="python scripts/quick_bench.py 2097152 generic"
  $GCMD > gpurun_out/gen_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:generic_ -s 12 -c 2 -o gpurun_out/prof_gen -f $GCMD > gpurun_out/ncu_gen.log 2>&1
  echo ncu_gen_rc=$?
  QCMD="python bench.py --steps 2 --warmup 3 --no-train --no-cpu --no-e2e"
  ncu --set full --clock-control none --import-source on -k regex:ad2_queries -s 3 -c 1 -o gpurun_out/prof_q -f $QCMD > gpurun_out/ncu_q.log 2>&1
  echo ncu_q_rc=$?
fi
