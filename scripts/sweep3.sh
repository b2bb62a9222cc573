#!/bin/bash
w() { env "$@" python scripts/bench_wide.py 1048576 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('adw', d['env'], round(d['fwd_GBs']), round(d['bwd_GBs']))"; }
for cfg in 222 232 231; do for nw in 2 3 4; do w VAEL_ADW_FWD_CFG=$cfg VAEL_ADW_FWD_WARPS=$nw; done; done
