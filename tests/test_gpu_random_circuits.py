"""Randomised circuits through the two paths a circuit without a recognised shape can take — the circuit-specialised
straight-line kernels (generated + NVRTC-compiled per circuit; probability semiring) and the level-ordered CSR
interpreter: every node kind (negative literals, constants, unary / binary / n-ary products, sums and complements),
both semirings, with and without the normaliser, ragged batches — forward against the fp32 CSR-order oracle
(bit-exact in the probability semiring), adjoint against fp64 autograd through the node-by-node semantics."""
import numpy as np
import pytest
import torch

from oracle import fp_interp, torch_interp
from tests.util import rel_err

pytestmark = pytest.mark.gpu


def random_circuit(seed, log_semiring, n_facts=7, n_levels=4, width=9):
    from vael_b200.circuit import SEMIRING_LOG, SEMIRING_PROB, CircuitBuilder
    rng = np.random.default_rng(seed)
    b = CircuitBuilder(n_facts, SEMIRING_LOG if log_semiring else SEMIRING_PROB)
    pool = [b.leaf(f, True) for f in range(n_facts)] + [b.leaf(f, False) for f in range(0, n_facts, 2)]
    if not log_semiring:
        pool += [b.const(0.25), b.const(1.0)]
    levels = [list(pool)]
    for L in range(1, n_levels):
        new = []
        for _ in range(width):
            lower = [n for lv in levels for n in lv]
            fan = int(rng.choice([1, 2, 2, 2, 3, 5, 7]))
            ch = [levels[-1][int(rng.integers(len(levels[-1])))]] + \
                 [lower[int(i)] for i in rng.integers(0, len(lower), size=fan - 1)]
            kinds = ["prod", "sum"] if log_semiring else ["prod", "prod", "sum", "sum", "compl"]
            kind = kinds[int(rng.integers(len(kinds)))]
            if kind == "compl":      # keep complements in range: 1 - (a scaled-down sum)
                s = b.sum(tuple(ch))
                new.append(b.compl((b.prod((s, b.const(0.1))),)))
            else:
                new.append(getattr(b, kind)(tuple(ch)))
        levels.append(new)
    tops = [n for lv in levels[1:] for n in lv]
    worlds = list(dict.fromkeys(tops[-width:] + tops[:3]))          # distinct world nodes
    queries = [tops[int(i)] for i in rng.integers(0, len(tops), size=5)]
    norm = -1
    if not log_semiring:
        norm = b.sum((b.prod((levels[1][0], b.const(0.5))), b.const(1.0)))
    return b.finish(worlds, queries, norm, fact_labels=[f"f{i}" for i in range(n_facts)],
                    world_atoms=[f"w{i}" for i in range(len(worlds))], query_atoms=[f"q{i}" for i in range(5)])


@pytest.mark.parametrize("path", ["default", "interpreter"])
@pytest.mark.parametrize("log_semiring", [False, True])
@pytest.mark.parametrize("seed", range(6))
def test_random_circuit_forward_and_adjoint(cuda_device, seed, log_semiring, path):
    from vael_b200 import _lib
    from vael_b200.ops import DeviceCircuit, circuit_forward, circuit_queries_all
    circ = random_circuit(seed, log_semiring)
    dc = DeviceCircuit(circ, cuda_device)
    assert dc.fast_kind == 0
    generic = path == "interpreter"
    if log_semiring and not generic:
        pytest.skip("the log semiring always runs the interpreter")
    jit_ok, why = dc.jit_status()
    assert jit_ok == (not log_semiring), why
    kernel = "generic-csr" if (generic or log_semiring) else "jit-straightline"
    rng = np.random.default_rng(100 + seed)
    for B in (1, 33, 200):
        facts = rng.uniform(0.05, 0.95, size=(B, circ.n_facts)).astype(np.float32)
        qi = rng.integers(0, circ.n_queries, size=B)
        for normalize in ([False] if log_semiring else [False, True]):
            f = torch.from_numpy(facts).cuda().requires_grad_(True)
            w, q = circuit_forward(dc, f, torch.from_numpy(qi).cuda(), normalize=normalize, force_generic=generic)
            assert _lib.last_kernel() == kernel
            w_ref, q_ref = fp_interp.forward(circ, facts, query=qi, normalize=normalize)
            if log_semiring:   # logf / expf differ by a few ulp between libm and the device
                np.testing.assert_allclose(w.detach().cpu().numpy(), w_ref, rtol=2e-5, atol=2e-6)
                np.testing.assert_allclose(q.detach().cpu().numpy()[:, 0], q_ref, rtol=2e-5, atol=2e-6)
            else:
                np.testing.assert_array_equal(w.detach().cpu().numpy(), w_ref)
                np.testing.assert_array_equal(q.detach().cpu().numpy()[:, 0], q_ref)
            gw = rng.standard_normal(w.shape).astype(np.float32)
            gq = rng.standard_normal(B).astype(np.float32)
            ((w * torch.from_numpy(gw).cuda()).sum() + (q[:, 0] * torch.from_numpy(gq).cuda()).sum()).backward()
            f64 = torch.from_numpy(facts.astype(np.float64)).requires_grad_(True)
            w64, q64 = torch_interp.forward(circ, f64, torch.from_numpy(qi), normalize=normalize)
            ((w64 * torch.from_numpy(gw).double()).sum() + (q64 * torch.from_numpy(gq).double()).sum()).backward()
            g_ref = f64.grad.numpy()
            assert _lib.last_kernel() == kernel
            assert rel_err(f.grad.cpu().numpy(), g_ref, floor=1e-3 * np.abs(g_ref).max() + 1e-12) < 1e-4
            if not log_semiring:   # every query node of every row, and P(w | query) — same bits on both paths
                qa = circuit_queries_all(dc, f.detach(), normalize=normalize, force_generic=generic)
                assert _lib.last_kernel() == kernel
                np.testing.assert_array_equal(qa.cpu().numpy(), fp_interp.queries_all(circ, facts, normalize=normalize))


@pytest.mark.parametrize("env", [{"VAEL_GEN_R": "1"}, {"VAEL_GEN_R": "4"}, {"VAEL_GEN_R": "4", "VAEL_GEN_CTAS": "3"},
                                 {"VAEL_GEN_R": "2", "VAEL_GEN_CTAS": "6", "VAEL_GEN_WARPS_PER_SM": "12"}],
                         ids=lambda e: "-".join(f"{k[9:]}{v}" for k, v in e.items()))
def test_generic_kernel_tile_shapes(cuda_device, env):
    """The generic kernels pick rows-per-lane (R) and the staged chunk width once per process: every forced
    shape (1 / 2 / 4 rows per lane, whole-tile and chunked row-by-row staging) must give the same bits."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "tests.variant_check"], cwd=root, env={**os.environ, **env},
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "variant_check ok" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
