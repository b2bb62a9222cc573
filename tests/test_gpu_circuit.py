"""Parity of the CUDA circuit kernels (through the C-ABI) against the CPU oracle.

Tolerances (BASELINE.json north_star): world/query probabilities 1e-5 relative, gradients
1e-4 relative, fp32.  The forward is in fact required to be BIT-exact against the fp32
sequential interpreter (same IEEE ops in the same order); the documented 1e-5 bound is
checked against the fp64 interpreter.
"""
import numpy as np
import pytest
import torch

from oracle import fp_interp
from tests.util import rel_err, synthetic_facts

pytestmark = pytest.mark.gpu

BATCHES = [1, 31, 32, 33, 100, 1000, 4096 + 7]


@pytest.fixture(scope="module")
def dc(addition_family, cuda_device):
    from vael_b200.ops import DeviceCircuit
    d = DeviceCircuit(addition_family.circuit, cuda_device)
    assert d.fast_kind == 2, "2-digit addition must take the two-AD fast path"
    assert d.hash == addition_family.circuit.fnv1a64()
    return d


@pytest.mark.parametrize("force_generic", [False, True])
@pytest.mark.parametrize("B", BATCHES)
def test_forward_query_scalar_bit_exact(dc, addition_family, B, force_generic):
    from vael_b200 import _lib
    from vael_b200.ops import circuit_forward
    circ = addition_family.circuit
    facts = synthetic_facts(B, seed=B)
    for q in (0, 9, 18):
        w_ref, q_ref = fp_interp.forward(circ, facts, query=q)
        w, qq = circuit_forward(dc, torch.from_numpy(facts).cuda(), q, force_generic=force_generic)
        assert _lib.last_kernel() == ("generic-csr" if force_generic else "ad2-tma")
        np.testing.assert_array_equal(w.cpu().numpy(), w_ref)
        np.testing.assert_array_equal(qq.cpu().numpy()[:, 0], q_ref)
        w64, q64 = fp_interp.forward(circ, facts, query=q, dtype=np.float64)
        assert rel_err(w.cpu().numpy(), w64) < 1e-5 and rel_err(qq.cpu().numpy()[:, 0], q64) < 1e-5


@pytest.mark.parametrize("force_generic", [False, True])
@pytest.mark.parametrize("normalize", [False, True])
def test_forward_query_per_row(dc, addition_family, force_generic, normalize):
    from vael_b200.ops import circuit_forward
    circ = addition_family.circuit
    B = 777
    facts = synthetic_facts(B, seed=5)
    qi = np.random.default_rng(1).integers(0, 19, size=B)
    w_ref, q_ref = fp_interp.forward(circ, facts, query=qi, normalize=normalize)
    w, q = circuit_forward(dc, torch.from_numpy(facts).cuda(), torch.from_numpy(qi).cuda(),
                           normalize=normalize, force_generic=force_generic)
    np.testing.assert_array_equal(w.cpu().numpy(), w_ref)
    np.testing.assert_array_equal(q.cpu().numpy()[:, 0], q_ref)


@pytest.mark.parametrize("force_generic", [False, True])
def test_forward_evidence(dc, addition_family, force_generic):
    from vael_b200.ops import circuit_forward
    circ = addition_family.circuit
    B = 130
    facts = synthetic_facts(B, seed=6)
    for e in (0, 7, 18):
        w_ref, _ = fp_interp.forward(circ, facts, query=e, evidence=True)
        with torch.no_grad():
            w, _ = circuit_forward(dc, torch.from_numpy(facts).cuda(), e, evidence=True, force_generic=force_generic)
        np.testing.assert_array_equal(w.cpu().numpy(), w_ref)
        np.testing.assert_allclose(w.cpu().numpy().sum(1), 1.0, rtol=1e-5)


@pytest.mark.parametrize("force_generic", [False, True])
@pytest.mark.parametrize("normalize", [False, True])
@pytest.mark.parametrize("B", [1, 33, 1000])
def test_backward(dc, addition_family, B, normalize, force_generic):
    from vael_b200.ops import circuit_forward
    circ = addition_family.circuit
    rng = np.random.default_rng(B)
    facts = synthetic_facts(B, seed=B + 1)
    qi = rng.integers(0, 19, size=B)
    gw = rng.standard_normal((B, 100)).astype(np.float32)
    gq = rng.standard_normal((B,)).astype(np.float32)
    g_ref = fp_interp.backward(circ, facts, gw, gq, qi, normalize=normalize)
    f = torch.from_numpy(facts).cuda().requires_grad_(True)
    w, q = circuit_forward(dc, f, torch.from_numpy(qi).cuda(), normalize=normalize, force_generic=force_generic)
    (w * torch.from_numpy(gw).cuda()).sum().add((q[:, 0] * torch.from_numpy(gq).cuda()).sum()).backward()
    scale = np.abs(g_ref).max()
    assert rel_err(f.grad.cpu().numpy(), g_ref, floor=1e-3 * scale) < 1e-4
    # scalar query, only one of the two upstream gradients present
    f2 = torch.from_numpy(facts).cuda().requires_grad_(True)
    w2, q2 = circuit_forward(dc, f2, 9, normalize=normalize, force_generic=force_generic)
    q2.sum().backward()
    g2 = fp_interp.backward(circ, facts, None, np.ones(B), 9, normalize=normalize)
    assert rel_err(f2.grad.cpu().numpy(), g2, floor=1e-3 * np.abs(g2).max()) < 1e-4


@pytest.mark.parametrize("normalize", [False, True])
@pytest.mark.parametrize("B", [1, 33, 257, 5000])
def test_queries_all(dc, addition_family, B, normalize):
    """[B,19] = worlds @ w_q (utils/mnist_utils/metrics.py:71-75) from the fused kernel and from the
    generic CSR kernel, both bit-exact against the CSR-order fp32 oracle."""
    from vael_b200 import _lib
    from vael_b200.ops import circuit_queries_all
    circ = addition_family.circuit
    facts = synthetic_facts(B, seed=3 + B)
    ref = fp_interp.queries_all(circ, facts, normalize=normalize)
    f = torch.from_numpy(facts).cuda()
    out = circuit_queries_all(dc, f, normalize=normalize)
    assert _lib.last_kernel() == "ad2-queries"
    np.testing.assert_array_equal(out.cpu().numpy(), ref)
    out_g = circuit_queries_all(dc, f, normalize=normalize, force_generic=True)
    assert _lib.last_kernel() == "generic-csr"
    np.testing.assert_array_equal(out_g.cpu().numpy(), ref)
    if not normalize:   # the reference's dense form: worlds @ w_q
        w = (facts[:, :10, None] * facts[:, None, 10:]).reshape(B, 100).astype(np.float64)
        np.testing.assert_allclose(out.cpu().numpy(), w @ circ.worlds_queries_matrix().astype(np.float64), rtol=1e-5)


@pytest.mark.parametrize("B", [1, 97, 4001])
def test_queries_all_overlapping_members_mario(cuda_device, B, monkeypatch):
    """Mario's five queries (four moves + `constraint`, which is the union of the moves: worlds belong to TWO
    queries, and the member lists have ragged lengths) through the hand-written member-list kernel (ad2<9,9>), the
    circuit's generated kernel (what the dispatcher prefers for member-list circuits) and the generic kernel: all
    bit-exact against the CSR-order fp32 oracle, with and without the normaliser."""
    from vael_b200 import _lib
    from vael_b200.mario_task import compile_mario_family
    from vael_b200.ops import DeviceCircuit, circuit_queries_all
    circ = compile_mario_family((3, 3)).circuit
    dcm = DeviceCircuit(circ, cuda_device)
    rng = np.random.default_rng(B)
    p = rng.dirichlet(np.ones(9), size=(B, 2)).astype(np.float32).clip(1e-5, 0.9999).reshape(B, 18)
    f = torch.from_numpy(p).cuda()
    for normalize in (False, True):
        ref = fp_interp.queries_all(circ, p, normalize=normalize)
        monkeypatch.setenv("VAEL_QUERIES_MEMBERLIST", "1")
        out = circuit_queries_all(dcm, f, normalize=normalize)
        assert _lib.last_kernel() == "ad2-queries"
        np.testing.assert_array_equal(out.cpu().numpy(), ref)
        monkeypatch.delenv("VAEL_QUERIES_MEMBERLIST")
        out_j = circuit_queries_all(dcm, f, normalize=normalize)
        assert _lib.last_kernel() == ("jit-straightline" if dcm.jit_status()[0] else "ad2-queries")
        np.testing.assert_array_equal(out_j.cpu().numpy(), ref)
        out_g = circuit_queries_all(dcm, f, normalize=normalize, force_generic=True)
        assert _lib.last_kernel() == "generic-csr"
        np.testing.assert_array_equal(out_g.cpu().numpy(), ref)
    assert ref.shape == (B, 5)
    # worlds [B,81] / per-row query / adjoint: W and F are not multiples of four (scalar staging paths of the generic
    # kernel, 324-byte rows in the structured one); structured and generic kernels must agree bit for bit forward
    from vael_b200.ops import circuit_forward
    qi = torch.from_numpy(rng.integers(0, 5, size=B)).cuda()
    gw = torch.from_numpy(rng.standard_normal((B, 81)).astype(np.float32)).cuda()
    res = []
    for generic in (False, True):
        fg = f.clone().requires_grad_(True)
        w, q = circuit_forward(dcm, fg, qi, force_generic=generic)
        assert _lib.last_kernel() == ("generic-csr" if generic else "ad2-tma")
        ((w * gw).sum() + 3.0 * q.sum()).backward()
        res.append((w.detach().cpu().numpy(), q.detach().cpu().numpy(), fg.grad.cpu().numpy()))
    w_ref, q_ref = fp_interp.forward(circ, p, query=qi.cpu().numpy())
    np.testing.assert_array_equal(res[0][0], w_ref)
    np.testing.assert_array_equal(res[1][0], w_ref)
    np.testing.assert_array_equal(res[0][1][:, 0], q_ref)
    np.testing.assert_array_equal(res[1][1][:, 0], q_ref)
    g_ref = fp_interp.backward(circ, p, gw.cpu().numpy(), 3.0 * np.ones(B), qi.cpu().numpy())
    for r in res:
        assert rel_err(r[2], g_ref, floor=1e-3 * np.abs(g_ref).max()) < 1e-4


def test_empty_batch_and_errors(dc):
    from vael_b200._lib import VaelError
    from vael_b200.ops import circuit_forward
    w, q = circuit_forward(dc, torch.zeros(0, 20, device="cuda"), 3)
    assert w.shape == (0, 100) and q.shape == (0, 1)
    with pytest.raises(KeyError):
        circuit_forward(dc, torch.zeros(4, 20, device="cuda"), 19)
    with pytest.raises(RuntimeError):
        circuit_forward(dc, torch.zeros(4, 20), 3)            # CPU tensor: no fallback
    with pytest.raises(RuntimeError):
        circuit_forward(dc, torch.zeros(4, 21, device="cuda"), 3)
    with pytest.raises(RuntimeError):
        circuit_forward(dc, torch.zeros(4, 20, device="cuda", dtype=torch.float64), 3)


def test_unaligned_views_take_the_plain_copy_path(dc, addition_family):
    """Bulk copies need 16-byte aligned spans; a buffer offset by one float must still work."""
    from vael_b200.ops import circuit_forward
    circ = addition_family.circuit
    B = 70
    facts = synthetic_facts(B, seed=11)
    buf = torch.zeros(B * 20 + 1, device="cuda")
    buf[1:] = torch.from_numpy(facts).cuda().reshape(-1)
    view = buf[1:].reshape(B, 20)
    assert view.data_ptr() % 16 != 0
    w, q = circuit_forward(dc, view, 5)
    w_ref, q_ref = fp_interp.forward(circ, facts, query=5)
    np.testing.assert_array_equal(w.cpu().numpy(), w_ref)
    np.testing.assert_array_equal(q.cpu().numpy()[:, 0], q_ref)


@pytest.mark.parametrize("B", [(1 << 21) + 3, 10_000_003])
def test_full_size_properties_and_host_path(dc, addition_family, B):
    """BASELINE.json microbenchmark scale (B = 2^21 + 3 and the full 1e7 rows of configs[3], 4 GB of worlds; the oracle
    could not finish these): size-independent properties of the reference's closed form (models/vael.py:208-225) — rows sum
    to 1, world (j,k) = p1[j] * p2[k] bit for bit, query = sum of the member worlds, evidence rows
    renormalise to 1 and vanish off the evidence, the adjoint is linear and matches its closed form — and
    the C-ABI host-buffer path (copies inside) returns exactly what the device path does."""
    import ctypes as C
    from vael_b200 import _lib
    from vael_b200.ops import _p, circuit_forward
    g = torch.Generator(device="cuda").manual_seed(0)
    p = torch.softmax(3 * torch.randn(B, 2, 10, device="cuda", generator=g), -1) + 1e-5
    facts = (p / p.sum(-1, keepdim=True)).reshape(B, 20).contiguous().requires_grad_(True)
    qi = torch.randint(0, 19, (B,), device="cuda", generator=g)
    w, q = circuit_forward(dc, facts, qi)
    f3 = facts.detach().reshape(B, 2, 10)
    assert torch.allclose(w.sum(1), torch.ones(B, device="cuda"), atol=1e-5)
    assert torch.equal(w.detach().reshape(B, 10, 10), f3[:, 0, :, None] * f3[:, 1, None, :])
    wq = torch.from_numpy(addition_family.circuit.worlds_queries_matrix().astype(np.float32)).cuda()
    q_dense = (w.detach().double() * wq.double().T[qi]).sum(1)
    assert torch.allclose(q.detach()[:, 0].double(), q_dense, rtol=2e-6, atol=1e-12)
    with torch.no_grad():
        we, _ = circuit_forward(dc, facts.detach(), qi, evidence=True)
    assert torch.allclose(we.sum(1), torch.ones(B, device="cuda"), atol=1e-5)
    assert float((we * (1 - wq.T[qi])).abs().max()) == 0.0
    # NORMALIZE at this size (the non-persistent kernels of streaming batches): Z = prod_AD (sum p + (1 - sum p)) is 1
    # up to rounding, so the normalised outputs are the plain ones within an ulp or two, and so is the adjoint
    fn = facts.detach().clone().requires_grad_(True)
    wn, qn = circuit_forward(dc, fn, qi, normalize=True)
    assert torch.allclose(wn.detach(), w.detach(), rtol=5e-7, atol=0) and torch.allclose(qn.detach(), q.detach(), rtol=5e-7, atol=0)
    gwn = torch.randn(B, 100, device="cuda", generator=g)
    (gn,) = torch.autograd.grad((wn * gwn).sum() + qn.sum(), [fn])
    (gp,) = torch.autograd.grad((w * gwn).sum() + q.sum(), [facts], retain_graph=True)
    assert torch.allclose(gn, gp, rtol=1e-4, atol=1e-4 * float(gp.abs().max()))
    del fn, wn, qn, gwn, gn, gp
    gw = torch.randn(B, 100, device="cuda", generator=g)
    (g1,) = torch.autograd.grad((w * gw).sum(), [facts], retain_graph=True)
    (g2,) = torch.autograd.grad((w * (3 * gw)).sum(), [facts], retain_graph=True)
    assert torch.allclose(g2, 3 * g1, rtol=1e-5, atol=1e-6)
    (gs,) = torch.autograd.grad(w.sum() + q.sum(), [facts], retain_graph=True)   # closed form of d/dp of sum_w + query
    member = wq.T[qi].reshape(B, 10, 10)
    expect = torch.cat([f3[:, 1].sum(1, keepdim=True) + (member * f3[:, 1, None, :]).sum(2),
                        f3[:, 0].sum(1, keepdim=True) + (member * f3[:, 0, :, None]).sum(1)], 1)
    assert torch.allclose(gs, expect, rtol=1e-5, atol=1e-6)

    # host-buffer entry points on a slice with a ragged chunk tail
    n = (1 << 17) + 77
    hf = facts.detach()[:n].cpu().pin_memory(); hq = qi[:n].to(torch.int32).cpu().pin_memory()
    hgw = gw[:n].cpu().pin_memory(); hgq = torch.ones(n).pin_memory()
    hw, hqo, hgf = torch.empty(n, 100).pin_memory(), torch.empty(n).pin_memory(), torch.empty(n, 20).pin_memory()
    _lib.check(dc.lib.vael_circuit_fwdbwd_host(dc.handle, _p(hf), _p(hq), -1, n, _p(hgw), _p(hgq), _p(hw), _p(hqo), _p(hgf), 0))
    assert torch.equal(hw, w.detach()[:n].cpu()) and torch.equal(hqo, q.detach()[:n, 0].cpu())
    (gd,) = torch.autograd.grad((w[:n] * gw[:n]).sum() + q[:n].sum(), [facts])
    assert torch.equal(hgf, gd[:n].cpu())
    hw2 = torch.empty(n, 100)          # pageable host memory works too
    _lib.check(dc.lib.vael_circuit_forward_host(dc.handle, _p(hf), None, 9, n, _p(hw2), None, 0))
    assert torch.equal(hw2, hw)


@pytest.mark.parametrize("B", [1, 77, 3000])
def test_admissible_worlds_literal_table_kernel(cuda_device, B):
    """Mario's admissible-world log-probabilities (models/vael.py:368-375) through the literal-table kernel
    (fast_kind 4): log-probabilities within 1e-5 * max(1, |ref|) of the reference's two-matmul form in fp64 — an
    absolute error of a log-probability is a relative error of the probability, which is what the 1e-5 bar is about
    (the base + delta form cancels on peaked rows) —, the generic CSR kernel (accurate logf, CSR
    order) to 1e-5 relative, adjoint against fp64 autograd; tables with mostly-positive rows take the base+ form; a
    table the kernel does not cover stays on the generic kernel."""
    from vael_b200 import _lib
    from vael_b200.compiler import circuit_from_world_table
    from vael_b200.mario_task import admissible_circuit, compile_mario_family, mario_tables
    from vael_b200.ops import DeviceCircuit, circuit_forward
    W_adm = mario_tables(compile_mario_family((3, 3)))[2]
    dca = DeviceCircuit(admissible_circuit(W_adm), cuda_device)
    assert dca.fast_kind == 4
    rng = np.random.default_rng(B)
    p = rng.dirichlet(np.ones(9), size=(B, 2)).astype(np.float32).clip(1e-5, 0.9999).reshape(B, 18)
    if B > 10:   # peaked rows: one position at the clip ceiling, the others at the floor (the cancelling case)
        p[:5] = 1e-5
        p[np.arange(5), np.arange(5)] = 0.9999
        p[np.arange(5), 9 + np.arange(5)] = 0.9999
    ga = rng.standard_normal((B, 24)).astype(np.float32)
    outs = []
    for generic in (False, True):
        f = torch.from_numpy(p).cuda().requires_grad_(True)
        lw, _ = circuit_forward(dca, f, None, force_generic=generic)
        assert _lib.last_kernel() == ("generic-csr" if generic else "lt-log")
        (lw * torch.from_numpy(ga).cuda()).sum().backward()
        assert _lib.last_kernel() == ("generic-csr" if generic else "lt-log")
        outs.append((lw.detach().cpu().numpy(), f.grad.cpu().numpy()))
    p64 = torch.from_numpy(p.astype(np.float64)).requires_grad_(True)
    Wt = torch.from_numpy(W_adm.astype(np.float64))
    ref = torch.log(p64) @ Wt.T + torch.log(1 - p64) @ (1 - Wt).T          # models/vael.py:371-374
    (ref * torch.from_numpy(ga).double()).sum().backward()
    g_ref = p64.grad.numpy()
    refn = ref.detach().numpy()
    assert np.all(np.abs(outs[0][0] - refn) <= 1e-5 * np.maximum(1.0, np.abs(refn)))
    assert rel_err(outs[1][0], refn) < 1e-5
    for o in outs:
        assert rel_err(o[1], g_ref, floor=1e-3 * np.abs(g_ref).max()) < 1e-4
    # a random table with sparse, dense and half-set rows: base- and base+ forms side by side
    tab = np.zeros((12, 18), dtype=np.int8)
    for w, k in enumerate((0, 1, 2, 3, 5, 8, 9, 10, 14, 16, 17, 18)):
        tab[w, rng.permutation(18)[:k]] = 1
    dct = DeviceCircuit(circuit_from_world_table(tab), cuda_device)
    assert dct.fast_kind == 4
    f = torch.from_numpy(p).cuda().requires_grad_(True)
    gt = rng.standard_normal((B, 12)).astype(np.float32)
    lw, _ = circuit_forward(dct, f, None)
    assert _lib.last_kernel() == "lt-log"
    (lw * torch.from_numpy(gt).cuda()).sum().backward()
    p64 = torch.from_numpy(p.astype(np.float64)).requires_grad_(True)
    Tt = torch.from_numpy(tab.astype(np.float64))
    ref = torch.log(p64) @ Tt.T + torch.log(1 - p64) @ (1 - Tt).T
    (ref * torch.from_numpy(gt).double()).sum().backward()
    refn = ref.detach().numpy()
    assert np.all(np.abs(lw.detach().cpu().numpy() - refn) <= 1e-5 * np.maximum(1.0, np.abs(refn)))
    assert rel_err(f.grad.cpu().numpy(), p64.grad.numpy(), floor=1e-3 * np.abs(p64.grad.numpy()).max()) < 1e-4
    # 19 fact columns: not a shape the kernel is built for; 40 worlds all setting column 0: the per-column list
    # does not fit the kernel's tables
    odd = DeviceCircuit(circuit_from_world_table(rng.integers(0, 2, size=(5, 19))), cuda_device)
    assert odd.fast_kind == 0
    long_col = rng.integers(0, 2, size=(40, 18)); long_col[:, 0] = 1
    assert DeviceCircuit(circuit_from_world_table(long_col), cuda_device).fast_kind == 0


@pytest.mark.parametrize("name", ["addition", "mario"])
def test_streaming_batch_kernels_agree_with_interpreter(cuda_device, addition_family, name):
    """From 2^18 rows up the two-AD kernels run on non-persistent grids (one tile per warp, single-buffered stages) and
    so do the generated kernels: at B = 2^18 + 77 (ragged last tile) all three paths — structured, generated, CSR
    interpreter — give the same bits forward (worlds, per-row query, NORMALIZE too) and the same adjoint to 1e-6 of its
    largest entry."""
    from vael_b200.mario_task import compile_mario_family
    from vael_b200.ops import DeviceCircuit, circuit_forward
    circ, groups = (addition_family.circuit, (10, 10)) if name == "addition" else (compile_mario_family((3, 3)).circuit, (9, 9))
    dc = DeviceCircuit(circ, cuda_device)
    B = (1 << 18) + 77 if name == "addition" else (1 << 16) + 77   # Mario: the adjoint alone is non-persistent here
    g = torch.Generator(device="cuda").manual_seed(9)
    n = groups[0]
    p = torch.softmax(2 * torch.randn(B, 2, n, device="cuda", generator=g), -1).clamp(1e-5, 0.9999)
    qi = torch.randint(0, circ.n_queries, (B,), device="cuda", generator=g)
    gw = torch.randn(B, circ.n_worlds, device="cuda", generator=g)
    paths = [dict(), dict(force_generic=True)] + ([dict(force_jit=True)] if dc.jit_status()[0] else [])
    for normalize in (False, True):
        res = []
        for kw in paths:
            f = p.reshape(B, 2 * n).clone().requires_grad_(True)
            w, q = circuit_forward(dc, f, qi, normalize=normalize, **kw)
            (gf,) = torch.autograd.grad((w * gw).sum() + 2 * q.sum(), [f])
            res.append((w.detach(), q.detach(), gf))
        for w, q, gf in res[1:]:
            assert torch.equal(w, res[0][0]) and torch.equal(q, res[0][1])
            assert float((gf - res[0][2]).abs().max()) <= 1e-6 * float(res[0][2].abs().max())
