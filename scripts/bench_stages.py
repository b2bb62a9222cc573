"""Scratch timing of the stage kernels around the circuit (facts head, gumbel, ELBO): CUDA events."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vael_b200.ops import elbo_terms, facts_from_logits, gumbel_world
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
out = {}
for B in (4096, 65536):
    r, x = torch.rand(B, 1, 28, 56, device="cuda"), torch.rand(B, 1, 28, 56, device="cuda")
    mu, lv, q = torch.randn(B, 23, device="cuda"), torch.randn(B, 23, device="cuda"), torch.rand(B, 1, device="cuda")
    ms = timeit(lambda: elbo_terms(r, x, mu, lv, q, 0.1, 1e-5, 1.0))
    out[f"elbo_B{B}"] = dict(ms=ms, GBs=B * 1568 * 12 / ms / 1e6)
for B in (1 << 20,):
    lg = torch.randn(B, 20, device="cuda")
    ms = timeit(lambda: facts_from_logits(lg, 2))
    out[f"facts_B{B}"] = dict(ms=ms, GBs=B * 160 / ms / 1e6)
    p = torch.softmax(torch.randn(B, 100, device="cuda"), 1)
    hb = torch.cat([torch.eye(10).repeat_interleave(10, 0), torch.eye(10).repeat(10, 1)], 1).cuda()
    ms = timeit(lambda: gumbel_world(p, hb, seed=1))
    out[f"gumbel_B{B}"] = dict(ms=ms, GBs=B * (400 + 400 + 4 + 80) / ms / 1e6)
    pg = p.clone().requires_grad_(True)
    wh, idx, ys = gumbel_world(pg, hb, seed=1)
    gwh = torch.randn_like(wh)
    ms = timeit(lambda: torch.autograd.grad(wh, pg, gwh, retain_graph=True))
    out[f"gumbel_bwd_B{B}"] = dict(ms=ms, GBs=B * (400 + 80 + 400 + 400) / ms / 1e6)
print(json.dumps(out))
