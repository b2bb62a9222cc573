#!/bin/bash
# Generic-kernel sweep: rows per lane (VAEL_GEN_R) x resident warps per SM (VAEL_GEN_WARPS_PER_SM)
for R in 4 2 1; do for W in 8 16 24 32; do
  VAEL_GEN_R=$R VAEL_GEN_WARPS_PER_SM=$W python scripts/quick_bench.py 2097152 generic 2>&1 | head -1 | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('R=$R warps=$W', 'fwd %.3f ms %.0f GB/s  bwd %.3f ms %.0f GB/s' % (d['fwd_ms'], d['fwd_GBs'], d['bwd_ms'], d['bwd_GBs']))"
done; done
