"""Smallest run that touches every kernel of libvael_b200.so once (written for compute-sanitizer under gpurun; the
tool has since been closed on the GPU pool, the script stays as an every-kernel smoke run):
ragged batch sizes so that the tail paths (plain-copy fallbacks, clipped TMA boxes) execute too."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from tests.util import synthetic_facts
from vael_b200.mnist_task import compile_addition_family
from vael_b200.mario_task import compile_mario_family, mario_tables, admissible_circuit
from vael_b200.ops import (DeviceCircuit, circuit_forward, circuit_queries_all, facts_from_logits,
                           facts_from_logits_clipped, gumbel_world, elbo_terms, logic_train)

dev = torch.device("cuda", 0)
for k, groups, B in ((2, (10, 10), 77), (3, (10, 10, 10), 45)):
    fam = compile_addition_family(k, 10)
    dc = DeviceCircuit(fam.circuit, dev)
    W, Q = fam.circuit.n_worlds, fam.circuit.n_queries
    for generic in (False, True):
        for norm in (False, True):
            f = torch.from_numpy(synthetic_facts(B, groups=groups, seed=1)).to(dev).requires_grad_(True)
            q = torch.arange(B, device=dev) % Q
            w, qq = circuit_forward(dc, f, q, normalize=norm, force_generic=generic)
            (w.sum() + 2 * qq.sum()).backward()
        with torch.no_grad():
            circuit_forward(dc, f.detach(), 3, evidence=True, force_generic=generic)
        circuit_queries_all(dc, f.detach(), force_generic=generic)
    if dc.jit_status()[0]:   # generated kernels (narrow: 2-digit, wide: 3-digit)
        for norm in (False, True):
            f = torch.from_numpy(synthetic_facts(B, groups=groups, seed=2)).to(dev).requires_grad_(True)
            w, qq = circuit_forward(dc, f, torch.arange(B, device=dev) % Q, normalize=norm, force_jit=True)
            (w.sum() + 2 * qq.sum()).backward()
        if k == 2:           # the column-chunked (wide) generated kernels have no evidence mode
            with torch.no_grad():
                circuit_forward(dc, f.detach(), 3, evidence=True, force_jit=True)
        circuit_queries_all(dc, f.detach(), force_jit=True)
    if k == 2:               # fused training-mode logic kernels: Philox (FAST) and injected noise (GENERAL)
        x = torch.randn(B, 20, device=dev, requires_grad=True)
        for noise in (None, -torch.empty(B, 100, device=dev).exponential_().log()):
            fo, qo, wh, idx, _ = logic_train(dc, x, torch.arange(B, device=dev, dtype=torch.int32) % Q, noise=noise, seed=5)
            (fo.sum() + qo.sum() + (wh * torch.randn_like(wh)).sum()).backward()
fam = compile_mario_family((3, 3))
dcm = DeviceCircuit(fam.circuit, dev)
f = torch.rand(37, 18, device=dev).clamp(1e-5, 0.9999).requires_grad_(True)
w, qq = circuit_forward(dcm, f, torch.arange(37, device=dev) % 4)
(w.sum() + qq.sum()).backward()
adm = DeviceCircuit(admissible_circuit(mario_tables(fam)[2]), dev)
f2 = torch.rand(37, 18, device=dev).clamp(1e-5, 0.9999).requires_grad_(True)
lw, _ = circuit_forward(adm, f2, None)
lw.sum().backward()
circuit_queries_all(dcm, f.detach())                       # member-list all-queries kernel
z = torch.randn(37, 18, device=dev, requires_grad=True)
facts_from_logits_clipped(z, 2, 1e-5, 0.9999).sum().backward()
tabs = mario_tables(fam)
os.environ["VAEL_GUMBEL_ROWS"] = "1"   # lane <-> row Gumbel kernels at these small batches too
w_adm = torch.from_numpy(tabs[2].astype("float32")).to(dev)   # [24, 18] Herbrand table of the admissible worlds
wh, idx, ys = gumbel_world(lw.detach().requires_grad_(True), w_adm, is_log=True, seed=3)   # lane <-> row, log domain
wh.sum().backward()
from oracle import dense_path   # test infrastructure: the reference's Herbrand base
pp = torch.softmax(torch.randn(45, 100, device=dev), 1).requires_grad_(True)
for nz in (None, -torch.empty(45, 100, device=dev).exponential_().log()):
    wh, idx, ys = gumbel_world(pp, dense_path.herbrand_base(20).to(dev), seed=4, noise=nz)   # lane <-> row, padded lists
    (wh * torch.randn_like(wh)).sum().backward()
os.environ["VAEL_GUMBEL_ROWS"] = "0"
p81 = torch.softmax(torch.randn(33, 81, device=dev), 1).requires_grad_(True)                 # one warp per row
wh, idx, ys = gumbel_world(p81, (torch.rand(81, 18, device=dev) < 0.3).float(), seed=1)
wh.sum().backward()
# host-buffer pipeline (two chunks per slot)
import ctypes
from vael_b200 import _lib
from vael_b200.ops import _p
lib = _lib.load()
dc2 = DeviceCircuit(compile_addition_family(2, 10).circuit, dev)
Bh = 5 * 32768 + 17
hf = torch.from_numpy(synthetic_facts(Bh, groups=(10, 10), seed=3)).pin_memory()
hq = (torch.arange(Bh, dtype=torch.int32) % 19).pin_memory()
hgw, hgq = torch.randn(Bh, 100).pin_memory(), torch.randn(Bh).pin_memory()
hw, hqo, hgf = torch.empty(Bh, 100).pin_memory(), torch.empty(Bh).pin_memory(), torch.empty(Bh, 20).pin_memory()
_lib.check(lib.vael_circuit_fwdbwd_host(dc2.handle, _p(hf), _p(hq), -1, Bh, _p(hgw), _p(hgq), _p(hw), _p(hqo), _p(hgf), 0))
_lib.check(lib.vael_circuit_query_grad_host(dc2.handle, _p(hf), _p(hq), -1, Bh, _p(hgq), _p(hqo), _p(hgf), 0))
x = torch.randn(33, 20, device=dev, requires_grad=True)
facts_from_logits(x, 2).sum().backward()
p = torch.softmax(torch.randn(33, 100, device=dev), 1).requires_grad_(True)
hb = (torch.rand(100, 20, device=dev) < 0.3).float()
wh, idx, ys = gumbel_world(p, hb, seed=1)
wh.sum().backward()
r = torch.rand(5, 1, 28, 56, device=dev, requires_grad=True)
mu = torch.randn(5, 23, device=dev, requires_grad=True); lv = torch.randn(5, 23, device=dev, requires_grad=True)
qp = torch.rand(5, 1, device=dev, requires_grad=True)
loss, terms = elbo_terms(r, torch.rand(5, 1, 28, 56, device=dev), mu, lv, qp, 0.1, 1e-5, 1.0)
loss.backward()
torch.cuda.synchronize()
print("sanitize_case ok")
