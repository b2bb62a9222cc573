#!/bin/bash
# N ranks on one box: the bench line (no secondary probes) incl. the graphed data-parallel train step
N=$1
set -x
mkdir -p gpurun_out
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 --no-secondary > gpurun_out/r2_n$N.json 2> gpurun_out/r2_n$N.err
echo rc=$?
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_n$N.json").read().strip().splitlines()[-1])
c=d["config"]; e=d["e2e"]
print("N",d["n_gpus"],"value",d["value"],"frac",d["roofline"]["frac"],"train",c["train_samples_per_sec"],c["train_ms_per_step"],c["train_mode"],"eager",d["train"]["eager"]["ms_per_step"])
print("e2e",e["value"],e["link"],"diet rows/s",e["query_grad_rows_per_sec"],e["query_grad_ms_per_step"])
PY
grep -c "NCCL INFO" gpurun_out/r2_n$N.err; grep -i "nranks\|NVLS" gpurun_out/r2_n$N.err | head -6
