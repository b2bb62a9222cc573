"""Stand-alone Gumbel kernels at small batches (W = 100, H = 20, Philox): the lane <-> row kernels need >= 32 rows per
warp and ~4.6 k instructions per lane, so below some batch the one-warp-per-row kernels have the shorter latency.
    VAEL_GUMBEL_ROWS=0|1 python scripts/gumbel_small_batch.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vael_b200 import _lib
from vael_b200.ops import _p
lib = _lib.load(); dev = torch.device("cuda:0")
sp = lambda: torch.cuda.current_stream().cuda_stream   # inside a graph capture: the capturing stream
hb = torch.cat([torch.eye(10).repeat_interleave(10, 0), torch.eye(10).repeat(10, 1)], 1).to(dev).contiguous()
out = {"rows_kernels": os.environ.get("VAEL_GUMBEL_ROWS", "default")}
for B in (32, 1024, 4096, 8192, 16384, 32768, 65536, 262144):
    p = torch.softmax(torch.randn(B, 100, device=dev), 1)
    ys, idx, wh = torch.empty(B, 100, device=dev), torch.empty(B, device=dev, dtype=torch.int32), torch.empty(B, 20, device=dev)
    gh, gs = torch.randn(B, 20, device=dev), torch.empty(B, 100, device=dev)
    f = lambda: _lib.check(lib.vael_gumbel_world_forward(_p(p), None, 7, B, 100, 0, 1.0, _p(hb), 20, _p(ys), _p(idx), _p(wh), sp()))
    b = lambda: _lib.check(lib.vael_gumbel_world_backward(_p(p), _p(ys), _p(gh), B, 100, 0, 1.0, _p(hb), 20, _p(gs), sp()))
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
