"""Parity of the wide-level path (3-digit addition: F=30, W=1000, Q=28 — BASELINE.json configs[4])
through the C-ABI: the TMA 2-D tiled kernels (adw) against the fp32/fp64 CSR oracle and against the
generic CSR kernel, plus size-independent properties at a batch the oracle could not finish."""
import numpy as np
import pytest
import torch

from oracle import fp_interp
from tests.util import rel_err, synthetic_facts

pytestmark = pytest.mark.gpu
G3 = (10, 10, 10)


@pytest.fixture(scope="module")
def fam3():
    from vael_b200.mnist_task import compile_addition_family
    return compile_addition_family(3, 10)


@pytest.fixture(scope="module")
def dc3(fam3, cuda_device):
    from vael_b200.ops import DeviceCircuit
    d = DeviceCircuit(fam3.circuit, cuda_device)
    assert d.fast_kind == 3, "3-digit addition must take the three-AD wide path"
    assert d.hash == fam3.circuit.fnv1a64()
    return d


@pytest.mark.parametrize("B", [1, 31, 32, 33, 100, 1000 + 3])
def test_wide_forward_bit_exact(dc3, fam3, B):
    from vael_b200 import _lib
    from vael_b200.ops import circuit_forward
    circ = fam3.circuit
    facts = synthetic_facts(B, groups=G3, seed=B)
    f = torch.from_numpy(facts).cuda()
    for q in (0, 13, 27):
        w_ref, q_ref = fp_interp.forward(circ, facts, query=q)
        w, qq = circuit_forward(dc3, f, q)
        assert _lib.last_kernel() == "adw-tma2d"
        np.testing.assert_array_equal(w.cpu().numpy(), w_ref)
        np.testing.assert_array_equal(qq.cpu().numpy()[:, 0], q_ref)
    w64, q64 = fp_interp.forward(circ, facts, query=27, dtype=np.float64)
    assert rel_err(w.cpu().numpy(), w64, floor=1e-30) < 1e-5 and rel_err(qq.cpu().numpy()[:, 0], q64, floor=1e-30) < 1e-5
    wg, qg = circuit_forward(dc3, f, 27, force_generic=True)
    assert _lib.last_kernel() == "generic-csr"
    assert torch.equal(wg, w) and torch.equal(qg, qq)


@pytest.mark.parametrize("normalize", [False, True])
def test_wide_forward_query_per_row(dc3, fam3, normalize):
    from vael_b200.ops import circuit_forward
    circ, B = fam3.circuit, 333
    facts = synthetic_facts(B, groups=G3, seed=5)
    qi = np.random.default_rng(1).integers(0, 28, size=B)
    w_ref, q_ref = fp_interp.forward(circ, facts, query=qi, normalize=normalize)
    w, q = circuit_forward(dc3, torch.from_numpy(facts).cuda(), torch.from_numpy(qi).cuda(), normalize=normalize)
    np.testing.assert_array_equal(w.cpu().numpy(), w_ref)
    np.testing.assert_array_equal(q.cpu().numpy()[:, 0], q_ref)
    # query only (no worlds tensor at all)
    _, q2 = circuit_forward(dc3, torch.from_numpy(facts).cuda(), torch.from_numpy(qi).cuda(), normalize=normalize,
                            want_worlds=False)
    np.testing.assert_array_equal(q2.cpu().numpy()[:, 0], q_ref)


def test_wide_forward_evidence(dc3, fam3):
    from vael_b200.ops import circuit_forward
    circ, B = fam3.circuit, 130
    facts = synthetic_facts(B, groups=G3, seed=6)
    for e in (0, 11, 27):
        w_ref, _ = fp_interp.forward(circ, facts, query=e, evidence=True)
        with torch.no_grad():
            w, _ = circuit_forward(dc3, torch.from_numpy(facts).cuda(), e, evidence=True)
        np.testing.assert_array_equal(w.cpu().numpy(), w_ref)
        np.testing.assert_allclose(w.cpu().numpy().sum(1), 1.0, rtol=1e-5)


@pytest.mark.parametrize("force_generic", [False, True])
@pytest.mark.parametrize("normalize", [False, True])
@pytest.mark.parametrize("B", [1, 33, 300])
def test_wide_backward(dc3, fam3, B, normalize, force_generic):
    from vael_b200.ops import circuit_forward
    circ = fam3.circuit
    rng = np.random.default_rng(B)
    facts = synthetic_facts(B, groups=G3, seed=B + 1)
    qi = rng.integers(0, 28, size=B)
    gw = rng.standard_normal((B, 1000)).astype(np.float32)
    gq = rng.standard_normal((B,)).astype(np.float32)
    g_ref = fp_interp.backward(circ, facts, gw, gq, qi, normalize=normalize)
    f = torch.from_numpy(facts).cuda().requires_grad_(True)
    w, q = circuit_forward(dc3, f, torch.from_numpy(qi).cuda(), normalize=normalize, force_generic=force_generic)
    (w * torch.from_numpy(gw).cuda()).sum().add((q[:, 0] * torch.from_numpy(gq).cuda()).sum()).backward()
    assert rel_err(f.grad.cpu().numpy(), g_ref, floor=1e-3 * np.abs(g_ref).max()) < 1e-4
    # scalar query and only the query gradient (no [B,W] gradient tensor: no tensor map at all)
    f2 = torch.from_numpy(facts).cuda().requires_grad_(True)
    _, q2 = circuit_forward(dc3, f2, 13, normalize=normalize, force_generic=force_generic, want_worlds=False)
    q2.sum().backward()
    g2 = fp_interp.backward(circ, facts, None, np.ones(B), 13, normalize=normalize)
    assert rel_err(f2.grad.cpu().numpy(), g2, floor=1e-3 * np.abs(g2).max()) < 1e-4


def test_wide_unaligned_world_buffer_falls_back_to_the_generic_kernel(dc3, fam3):
    """A [B,W] base address that is not 16-byte aligned cannot be described by a tensor map."""
    from vael_b200 import _lib
    from vael_b200.ops import _p, _stream_ptr
    B = 40
    facts = synthetic_facts(B, groups=G3, seed=3)
    f = torch.from_numpy(facts).cuda()
    buf = torch.zeros(B * 1000 + 1, device="cuda")
    out = buf[1:].reshape(B, 1000)
    assert out.data_ptr() % 16 != 0
    q = torch.empty(B, device="cuda")
    _lib.check(dc3.lib.vael_circuit_forward(dc3.handle, _p(f), None, 5, B, _p(out), _p(q), 0, _stream_ptr()))
    assert _lib.last_kernel() == "generic-csr"
    w_ref, q_ref = fp_interp.forward(fam3.circuit, facts, query=5)
    np.testing.assert_array_equal(out.cpu().numpy(), w_ref)
    np.testing.assert_array_equal(q.cpu().numpy(), q_ref)


def test_wide_full_size_properties(dc3, fam3):
    """B = 2^18 + 5 rows (1 GB of worlds): rows sum to 1, the query is the masked sum of the worlds,
    the adjoint is linear in the upstream gradient and equals the closed form for a constant one."""
    from vael_b200.ops import circuit_forward
    B = (1 << 18) + 5
    g = torch.Generator(device="cuda").manual_seed(0)
    p = torch.softmax(3 * torch.randn(B, 3, 10, device="cuda", generator=g), -1) + 1e-5
    facts = (p / p.sum(-1, keepdim=True)).reshape(B, 30).requires_grad_(True)
    qi = torch.randint(0, 28, (B,), device="cuda", generator=g)
    w, q = circuit_forward(dc3, facts, qi)
    assert torch.allclose(w.sum(1), torch.ones(B, device="cuda"), atol=2e-5)
    wq = torch.from_numpy(fam3.circuit.worlds_queries_matrix().astype(np.float32)).cuda()   # [1000, 28]
    q_dense = (w.detach().double() * wq.double().T[qi]).sum(1)
    assert torch.allclose(q.detach()[:, 0].double(), q_dense, rtol=2e-6, atol=1e-12)
    # world (i,j,k) = p1[i] p2[j] p3[k] at a few columns, bit for bit
    f3 = facts.detach().reshape(B, 3, 10)
    for (i, j, k) in ((0, 0, 0), (3, 1, 4), (9, 9, 9)):
        assert torch.equal(w.detach()[:, (i * 10 + j) * 10 + k], (f3[:, 0, i] * f3[:, 1, j]) * f3[:, 2, k])
    gw = torch.randn(B, 1000, device="cuda", generator=g)
    (g1,) = torch.autograd.grad((w * gw).sum(), [facts], retain_graph=True)
    (g2,) = torch.autograd.grad((w * (2 * gw)).sum(), [facts], retain_graph=True)
    assert torch.allclose(g2, 2 * g1, rtol=1e-5, atol=1e-6)
    (gs,) = torch.autograd.grad(w.sum(), [facts])     # d/dp1[i] sum_w = sum(p2) * sum(p3)
    expect = torch.cat([(f3[:, 1].sum(1) * f3[:, 2].sum(1))[:, None].expand(B, 10),
                        (f3[:, 0].sum(1) * f3[:, 2].sum(1))[:, None].expand(B, 10),
                        (f3[:, 0].sum(1) * f3[:, 1].sum(1))[:, None].expand(B, 10)], 1)
    assert torch.allclose(gs, expect, rtol=1e-5)


@pytest.mark.parametrize("normalize", [False, True])
@pytest.mark.parametrize("B", [1, 45, 1000])
def test_wide_queries_all(dc3, fam3, B, normalize):
    """All 28 sums of the 3-digit circuit per row (the `worlds @ w_q` loop of metrics.py:71-75 scaled to three
    digits).  The generic CSR kernel is bit-exact against the CSR-order fp32 oracle; the register-resident kernel
    contracts in two stages (sum_i p1[i] * (sum_{j+k=s-i} p2[j] p3[k]): 290 FMAs per row instead of 2100 operations),
    i.e. the same terms in another association and order — held to 1e-6 relative (floor 1e-9) against the fp64
    oracle, an order of magnitude inside the north star's 1e-5; worlds are never materialised."""
    from oracle import fp_interp
    from tests.util import synthetic_facts
    from vael_b200 import _lib
    from vael_b200.ops import circuit_queries_all
    circ = fam3.circuit
    facts = synthetic_facts(B, groups=(10, 10, 10), seed=11 + B)
    ref = fp_interp.queries_all(circ, facts, normalize=normalize)
    f = torch.from_numpy(facts).cuda()
    out = circuit_queries_all(dc3, f, normalize=normalize)
    assert _lib.last_kernel() == "adw-queries"
    p64 = facts.astype(np.float64).reshape(B, 3, 10)
    w64 = (p64[:, 0, :, None, None] * p64[:, 1, None, :, None] * p64[:, 2, None, None, :]).reshape(B, 1000)
    ref64 = w64 @ circ.worlds_queries_matrix().astype(np.float64)
    if normalize:   # ProbLog's Z = prod_AD (sum p + (1 - sum p)) is 1 up to rounding
        assert rel_err(out.cpu().numpy(), ref, floor=1e-9) < 2e-6
    else:
        assert rel_err(out.cpu().numpy(), ref64, floor=1e-9) < 1e-6
    out_g = circuit_queries_all(dc3, f, normalize=normalize, force_generic=True)
    assert _lib.last_kernel() == "generic-csr"
    np.testing.assert_array_equal(out_g.cpu().numpy(), ref)
    if not normalize:
        assert np.allclose(ref.sum(1), 1.0, atol=2e-5)   # every world belongs to exactly one sum
