#!/bin/bash
mkdir -p gpurun_out
cat > /tmp/nccl_dbg.py <<'PY'
import os, sys
print("env NCCL_DEBUG at start:", os.environ.get("NCCL_DEBUG"), file=sys.stderr)
if not os.environ.get("NCCL_DEBUG"):
    os.environ["NCCL_DEBUG"] = "INFO"; os.environ["NCCL_DEBUG_SUBSYS"] = "INIT"
import torch, torch.distributed as dist
local = int(os.environ["LOCAL_RANK"]); torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
t = torch.ones(4, device="cuda"); dist.all_reduce(t); torch.cuda.synchronize()
print("rank", dist.get_rank(), "sum", t[0].item(), file=sys.stderr)
dist.destroy_process_group()
PY
echo "== A: env unset, set inside python before import torch"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 /tmp/nccl_dbg.py > gpurun_out/dbgA.out 2> gpurun_out/dbgA.err; grep -c "NCCL INFO" gpurun_out/dbgA.out gpurun_out/dbgA.err; grep -h "nranks" gpurun_out/dbgA.out gpurun_out/dbgA.err | head -3
echo "== B: NCCL_DEBUG=INFO exported by the shell"
NCCL_DEBUG=INFO python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 /tmp/nccl_dbg.py > gpurun_out/dbgB.out 2> gpurun_out/dbgB.err; grep -c "NCCL INFO" gpurun_out/dbgB.out gpurun_out/dbgB.err; grep -h "nranks" gpurun_out/dbgB.out gpurun_out/dbgB.err | head -3
echo "== C: exported + NCCL_DEBUG_FILE=/dev/stderr"
NCCL_DEBUG=INFO NCCL_DEBUG_FILE=/dev/stderr python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 /tmp/nccl_dbg.py > gpurun_out/dbgC.out 2> gpurun_out/dbgC.err; grep -c "NCCL INFO" gpurun_out/dbgC.out gpurun_out/dbgC.err
head -5 gpurun_out/dbgA.err
