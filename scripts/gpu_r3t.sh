#!/bin/bash
set -x
timeout 600 python scripts/stage_probe.py lt gumbel facts queries fused 2>&1 | tail -1
for w in 4 6; do VAEL_JIT_FWD_WARPS=$w timeout 600 python scripts/stage_probe.py queries 2>&1 | tail -1; done
