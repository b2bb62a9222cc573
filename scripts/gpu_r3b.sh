#!/bin/bash
set -x
for cfg in "4 2" "5 2" "6 2" "6 1" "5 3"; do set -- $cfg; echo "== nsf $1 warps $2"; VAEL_JIT_WIDE_NSO=$1 VAEL_JIT_FWD_WARPS=$2 timeout 600 python scripts/stage_probe.py jitwide 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print({k:v['GBs'] for k,v in d.items() if k.startswith('jit_') or k.startswith('structured')})"; done
echo "== default"; timeout 600 python scripts/stage_probe.py jitwide 2>/dev/null | tail -1
timeout 900 python -m pytest tests/test_gpu_jit.py tests/test_gpu_guard_bands.py tests/test_gpu_wide.py -m gpu -q --timeout 900 2>&1 | tail -4
