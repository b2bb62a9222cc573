#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -30 > gpurun_out/r2j_pytest.log; tail -8 gpurun_out/r2j_pytest.log
for fw in 3 4 5 6 7 8; do for bw in $fw; do echo "== warps fwd=$fw bwd=$bw"; VAEL_JIT_FWD_WARPS=$fw VAEL_JIT_BWD_WARPS=$bw timeout 300 python scripts/stage_probe.py jit jitwide 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print({k:v['GBs'] for k,v in d.items() if k.startswith('jit_')})"; done; done
