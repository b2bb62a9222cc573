#!/bin/bash
# generic CSR knobs on the 3-digit circuit (schedule in shared memory) + Mario queries profile
set -x
mkdir -p gpurun_out
for lim in 12288 65536; do
  echo "== VAEL_GEN_SCHED_SMEM=$lim"
  VAEL_GEN_SCHED_SMEM=$lim timeout 300 python - <<'PY'
import torch, json
from vael_b200 import _lib
from vael_b200.mnist_task import compile_addition_family
from vael_b200.ops import DeviceCircuit, _p
lib=_lib.load(); dev=torch.device("cuda",0); sp=torch.cuda.current_stream().cuda_stream
dc3=DeviceCircuit(compile_addition_family(3,10).circuit, dev)
B=1<<18
g=torch.Generator(device=dev).manual_seed(0)
p=torch.softmax(3*torch.randn(B,3,10,device=dev,generator=g),-1).reshape(B,30).contiguous()
q=torch.randint(0,28,(B,),device=dev,dtype=torch.int32,generator=g)
w=torch.empty(B,1000,device=dev); qo=torch.empty(B,device=dev); gf=torch.empty(B,30,device=dev)
gw=torch.randn(B,1000,device=dev,generator=g); gq=torch.randn(B,device=dev,generator=g)
def t(fn,n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/n
f=t(lambda:_lib.check(lib.vael_circuit_forward(dc3.handle,_p(p),_p(q),-1,B,_p(w),_p(qo),4,sp)))
b=t(lambda:_lib.check(lib.vael_circuit_backward(dc3.handle,_p(p),_p(q),-1,B,_p(gw),_p(gq),_p(gf),4,sp)))
print(json.dumps({"fwd_ms":f,"fwd_GBs":B*4128/f/1e6,"bwd_ms":b,"bwd_GBs":B*4248/b/1e6}))
PY
done
export PROBE_WARM=1 PROBE_ITERS=1
CMD="python scripts/stage_probe.py queries"
$CMD > gpurun_out/plain_stage.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"ad2_queries" -c 2 -o gpurun_out/prof_stage4 -f $CMD > gpurun_out/ncu_stage4.log 2>&1
echo ncu_rc=$?
