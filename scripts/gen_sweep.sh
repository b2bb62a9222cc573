#!/bin/bash
# Generic-kernel sweep (2-digit circuit, 2 Mi rows): rows per lane (VAEL_GEN_R) x resident CTAs per SM with chunked
# staging (VAEL_GEN_CTAS) x resident warps per SM (VAEL_GEN_WARPS_PER_SM).  Usage: gen_sweep.sh ["R CTAS WARPS" ...]
cfgs=("$@")
if [ ${#cfgs[@]} -eq 0 ]; then
  cfgs=("4 1 16" "4 2 24" "2 1 16" "2 1 24" "2 1 32" "2 4 24" "1 1 16" "1 1 24" "1 6 24")
fi
for cfg in "${cfgs[@]}"; do set -- $cfg
  VAEL_GEN_R=$1 VAEL_GEN_CTAS=$2 VAEL_GEN_WARPS_PER_SM=$3 python scripts/quick_bench.py 2097152 generic 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('R=$1 ctas=$2 warps=$3', 'fwd %.3f ms %.0f GB/s  bwd %.3f ms %.0f GB/s' % (d['fwd_ms'], d['fwd_GBs'], d['bwd_ms'], d['bwd_GBs']))"
done
