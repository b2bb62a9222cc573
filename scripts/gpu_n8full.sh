#!/bin/bash
# N = 2 exactly as the driver launches it: reference arm, then the full bench line (secondary probes included)
set -x
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --impl reference --gpus 8 --steps 3 --warmup 1 > gpurun_out/r2_n8_ref.json 2> gpurun_out/r2_n8_ref.err
echo ref_rc=$?; tail -c 600 gpurun_out/r2_n8_ref.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2_n8_full.json 2> gpurun_out/r2_n8_full.err
echo rc=$?
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_n8_full.json").read().strip().splitlines()[-1])
c=d["config"]; e=d["e2e"]
print("N",d["n_gpus"],"value",d["value"],"frac",d["roofline"]["frac"],"train",c["train_samples_per_sec"],c["train_ms_per_step"],c["train_mode"],"eager",d["train"]["eager"]["ms_per_step"])
print("e2e",e["value"],e["link"],"diet rows/s",e["query_grad_rows_per_sec"],e["query_grad_ms_per_step"])
print("secondary keys", list(d.get("secondary",{}).keys()))
PY
grep -c "NCCL INFO" gpurun_out/r2_n8_full.err; grep -i "nranks\|NVLS\|Traceback\|Error" gpurun_out/r2_n8_full.err | head -8
tail -n 1 gpurun_out/r2_n8_full.json | cut -c1-200
