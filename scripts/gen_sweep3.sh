#!/bin/bash
for cfg in "2 1 16" "2 1 24" "2 1 32" "4 1 8" "1 1 24" "1 1 32"; do set -- $cfg
  VAEL_GEN_R=$1 VAEL_GEN_CTAS=$2 VAEL_GEN_WARPS_PER_SM=$3 python scripts/quick_bench.py 2097152 generic 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('R=$1 ctas=$2 warps=$3', 'fwd %.3f ms %.0f GB/s  bwd %.3f ms %.0f GB/s' % (d['fwd_ms'], d['fwd_GBs'], d['bwd_ms'], d['bwd_GBs']))"
done
