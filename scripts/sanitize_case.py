"""Smallest run that touches every kernel of libvael_b200.so once (for compute-sanitizer under gpurun):
ragged batch sizes so that the tail paths (plain-copy fallbacks, clipped TMA boxes) execute too."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from tests.util import synthetic_facts
from vael_b200.mnist_task import compile_addition_family
from vael_b200.mario_task import compile_mario_family, mario_tables, admissible_circuit
from vael_b200.ops import DeviceCircuit, circuit_forward, circuit_queries_all, facts_from_logits, gumbel_world, elbo_terms

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
fam = compile_mario_family((3, 3))
dcm = DeviceCircuit(fam.circuit, dev)
f = torch.rand(37, 18, device=dev).clamp(1e-5, 0.9999).requires_grad_(True)
w, qq = circuit_forward(dcm, f, torch.arange(37, device=dev) % 4)
(w.sum() + qq.sum()).backward()
adm = DeviceCircuit(admissible_circuit(mario_tables(fam)[2]), dev)
f2 = torch.rand(37, 18, device=dev).clamp(1e-5, 0.9999).requires_grad_(True)
lw, _ = circuit_forward(adm, f2, None)
lw.sum().backward()
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
