#!/bin/bash
# knob sweep for the ad2 / adw kernels (one JSON line per configuration)
w() { env "$@" python scripts/bench_wide.py 1048576 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('adw', d['env'], round(d['fwd_GBs']), round(d['bwd_GBs']))"; }
a() { env "$@" python scripts/quick_bench.py 8388608 2>/dev/null | head -1 | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('ad2', d['env'], round(d['fwd_GBs']), round(d['bwd_GBs']))"; }
a VAEL_L2_HINT=0
a VAEL_L2_HINT=1
a VAEL_L2_HINT=2
a VAEL_L2_HINT=1 VAEL_AD2_FWD_WARPS=5
a VAEL_L2_HINT=0 VAEL_AD2_FWD_WARPS=5
w VAEL_L2_HINT=0
w VAEL_L2_HINT=1
w VAEL_L2_HINT=2
for st in 22 23 24; do for nw in 3 4 5 6; do w VAEL_ADW_FWD_STAGES=$st VAEL_ADW_FWD_WARPS=$nw; done; done
for st in 2 3 4; do for nw in 3 4 5; do w VAEL_ADW_BWD_STAGES=$st VAEL_ADW_BWD_WARPS=$nw; done; done
