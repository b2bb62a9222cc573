"""Fused logic kernels at small batches (graph-captured launches): a lane runs ~5 k instructions for its row whatever
the batch, so below ~19 k rows the kernel time is a latency floor.
    [VAEL_FUSED_FWD_MINB=1|6] [VAEL_FUSED_BWD_MINB=1|4] python scripts/fused_small_batch.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vael_b200 import _lib
from vael_b200.mnist_task import compile_addition_family
from vael_b200.ops import DeviceCircuit, _p
lib = _lib.load(); dev = torch.device("cuda:0")
sp = lambda: torch.cuda.current_stream().cuda_stream
dc = DeviceCircuit(compile_addition_family(2, 10).circuit, dev)
out = {"fwd_minb": os.environ.get("VAEL_FUSED_FWD_MINB", "auto"), "bwd_minb": os.environ.get("VAEL_FUSED_BWD_MINB", "auto")}
for B in (32, 1024, 4096, 8192, 16384, 32768, 131072):
    x = 3 * torch.randn(B, 20, device=dev); q = torch.randint(0, 19, (B,), device=dev, dtype=torch.int32)
    facts, query, wh = torch.empty(B, 20, device=dev), torch.empty(B, device=dev), torch.empty(B, 20, device=dev)
    idx, gl = torch.empty(B, device=dev, dtype=torch.int32), torch.empty(B, 20, device=dev)
    gh, gq = torch.randn(B, 20, device=dev), torch.randn(B, device=dev)
    f = lambda: _lib.check(lib.vael_logic_train_forward(dc.handle, _p(x), _p(q), -1, B, 1e-5, 1.0, None, 1234, None, _p(facts),
                                                        _p(query), _p(idx), _p(wh), None, sp()))
    b = lambda: _lib.check(lib.vael_logic_train_backward(dc.handle, _p(x), _p(q), -1, B, 1e-5, 1.0, None, 1234, None, _p(gh),
                                                         _p(gq), None, _p(gl), sp()))
    res = []
    for fn in (f, b):
        for _ in range(5): fn()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(20): fn()
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        res.append(round(e0.elapsed_time(e1) / 20 * 1e3, 2))
    out[B] = {"fwd_us": res[0], "bwd_us": res[1]}
print(json.dumps(out))
