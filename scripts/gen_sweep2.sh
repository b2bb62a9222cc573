#!/bin/bash
# Generic-kernel sweep: rows per lane x resident CTAs per SM (chunked staging) x warps per SM
for cfg in "4 2 16" "4 2 24" "4 2 32" "2 3 24" "2 4 24" "2 4 32" "2 5 30" "1 6 24" "1 8 32"; do set -- $cfg
  VAEL_GEN_R=$1 VAEL_GEN_CTAS=$2 VAEL_GEN_WARPS_PER_SM=$3 python scripts/quick_bench.py 2097152 generic 2>&1 | head -1 | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('R=$1 ctas=$2 warps=$3', 'fwd %.3f ms %.0f GB/s  bwd %.3f ms %.0f GB/s' % (d['fwd_ms'], d['fwd_GBs'], d['bwd_ms'], d['bwd_GBs']))"
done
