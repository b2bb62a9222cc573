#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
run() { VAEL_AD2_FWD_STAGES=$1 VAEL_AD2_FWD_WARPS=$2 VAEL_AD2_FWD_CTAS_PER_SM=$3 python scripts/quick_bench.py 8388608 2>/dev/null | head -1 | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('stages=$1 warps=$2 cps=$3', round(d['fwd_GBs']), round(d['bwd_GBs']), round(d['copy_GBs']))"; }
for w in 2 3 4 5; do run 22 $w 1; done
for w in 2 3; do run 22 $w 2; done
run 22 2 3
for w in 3 4 5; do run 23 $w 1; done
for w in 4 6 8; do run 21 $w 1; done
for w in 4 6 8; do run 31 $w 1; done
run 21 4 2; run 21 3 3; run 21 2 4
