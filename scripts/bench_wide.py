"""Scratch timing of the 3-digit (W=1000) circuit kernels: device-resident inputs, CUDA events."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vael_b200 import _lib
from vael_b200.mnist_task import compile_addition_family
from vael_b200.ops import DeviceCircuit, _p, _stream_ptr

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
flags = 4 if (len(sys.argv) > 2 and sys.argv[2] == "generic") else 0
dc = DeviceCircuit(compile_addition_family(3, 10).circuit, "cuda:0")
lib = dc.lib
torch.manual_seed(0)
p = torch.softmax(3 * torch.randn(B, 3, 10, device="cuda"), -1) + 1e-5
facts = (p / p.sum(-1, keepdim=True)).reshape(B, 30).contiguous()
qidx = torch.randint(0, 28, (B,), device="cuda", dtype=torch.int32)
worlds = torch.empty(B, 1000, device="cuda"); query = torch.empty(B, device="cuda")
gw = torch.randn(B, 1000, device="cuda"); gq = torch.randn(B, device="cuda"); gf = torch.empty(B, 30, device="cuda")
st = _stream_ptr()
def fwd(): _lib.check(lib.vael_circuit_forward(dc.handle, _p(facts), _p(qidx), -1, B, _p(worlds), _p(query), flags, st))
def bwd(): _lib.check(lib.vael_circuit_backward(dc.handle, _p(facts), _p(qidx), -1, B, _p(gw), _p(gq), _p(gf), flags, st))
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
tf, tb = timeit(fwd), timeit(bwd)
print(json.dumps(dict(B=B, kernel=_lib.last_kernel(), fwd_ms=tf, bwd_ms=tb, fwd_GBs=B * 4128 / tf / 1e6,
                      bwd_GBs=B * 4248 / tb / 1e6, env={k: v for k, v in os.environ.items() if k.startswith("VAEL_")})))
