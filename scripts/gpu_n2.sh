#!/bin/bash
# N = 2: the graphed data-parallel train step (two CUDA graphs + one NCCL all-reduce) and the bench line
set -x
mkdir -p gpurun_out
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 --no-secondary > gpurun_out/r2_n2.json 2> gpurun_out/r2_n2.err
echo rc=$?
tail -c 3000 gpurun_out/r2_n2.json
grep -i "nranks\|NVLS\|error\|Traceback" gpurun_out/r2_n2.err | head -20
tail -5 gpurun_out/r2_n2.err
