"""Run by tests/test_gpu_random_circuits.py in a subprocess under VAEL_GEN_R / VAEL_GEN_CTAS (the generic
kernels read these once per process): random circuits, the 2-digit and the 3-digit circuit through the generic
CSR kernels at the forced tile shape, forward bit-exact against the fp32 oracle, adjoint against fp64."""
import sys

import numpy as np
import torch

from oracle import fp_interp, torch_interp
from tests.test_gpu_random_circuits import random_circuit
from tests.util import rel_err, synthetic_facts


def check(circ, facts, qi, normalize, log_semiring=False):
    from vael_b200 import _lib
    from vael_b200.ops import DeviceCircuit, circuit_forward, circuit_queries_all
    dc = DeviceCircuit(circ, torch.device("cuda", 0))
    B = facts.shape[0]
    rng = np.random.default_rng(B)
    f = torch.from_numpy(facts).cuda().requires_grad_(True)
    w, q = circuit_forward(dc, f, torch.from_numpy(qi).cuda(), normalize=normalize, force_generic=True)
    assert _lib.last_kernel() == "generic-csr"
    w_ref, q_ref = fp_interp.forward(circ, facts, query=qi, normalize=normalize)
    if log_semiring:
        np.testing.assert_allclose(w.detach().cpu().numpy(), w_ref, rtol=2e-5, atol=2e-6)
    else:
        np.testing.assert_array_equal(w.detach().cpu().numpy(), w_ref)
        np.testing.assert_array_equal(q.detach().cpu().numpy()[:, 0], q_ref)
        qa = circuit_queries_all(dc, f.detach(), normalize=normalize, force_generic=True)
        np.testing.assert_array_equal(qa.cpu().numpy(), fp_interp.queries_all(circ, facts, normalize=normalize))
    gw = rng.standard_normal(w.shape).astype(np.float32)
    gq = rng.standard_normal(B).astype(np.float32)
    ((w * torch.from_numpy(gw).cuda()).sum() + (q[:, 0] * torch.from_numpy(gq).cuda()).sum()).backward()
    f64 = torch.from_numpy(facts.astype(np.float64)).requires_grad_(True)
    w64, q64 = torch_interp.forward(circ, f64, torch.from_numpy(qi), normalize=normalize)
    ((w64 * torch.from_numpy(gw).double()).sum() + (q64 * torch.from_numpy(gq).double()).sum()).backward()
    g_ref = f64.grad.numpy()
    err = rel_err(f.grad.cpu().numpy(), g_ref, floor=1e-3 * np.abs(g_ref).max() + 1e-12)
    assert err < 1e-4, err


def main():
    from vael_b200.mnist_task import compile_addition_family
    rng = np.random.default_rng(7)
    for seed, log_semiring, width in ((0, False, 9), (1, True, 9), (2, False, 10), (3, False, 30)):
        circ = random_circuit(seed, log_semiring, width=width)
        for B in (1, 129, 300):
            facts = rng.uniform(0.05, 0.95, size=(B, circ.n_facts)).astype(np.float32)
            qi = rng.integers(0, circ.n_queries, size=B)
            for normalize in ([False] if log_semiring else [False, True]):
                check(circ, facts, qi, normalize, log_semiring)
    for k, groups, B in ((2, (10, 10), 333), (3, (10, 10, 10), 70)):
        circ = compile_addition_family(k, 10).circuit
        facts = synthetic_facts(B, groups=groups, seed=k)
        qi = rng.integers(0, circ.n_queries, size=B)
        for normalize in (False, True):
            check(circ, facts, qi, normalize)
    print("variant_check ok")


if __name__ == "__main__":
    sys.exit(main())
