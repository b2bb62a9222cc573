"""The circuit-specialised straight-line kernels (vael_b200/csrc/circuit_jit.cu: generated per circuit, compiled with
NVRTC on first use) against the hand-written structured kernels, the CSR interpreter and the oracles, on the circuits
VAEL ships: same bits forward (the same __fmul_rn / __fadd_rn folds in CSR order), adjoint within 1e-4 of fp64."""
import numpy as np
import pytest
import torch

from oracle import fp_interp
from tests.util import rel_err, synthetic_facts

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def circuits(cuda_device):
    from vael_b200.mario_task import compile_mario_family
    from vael_b200.mnist_task import compile_addition_family
    from vael_b200.ops import DeviceCircuit
    out = {}
    for name, circ, groups in (("addition", compile_addition_family(2, 10).circuit, (10, 10)),
                               ("mario", compile_mario_family((3, 3)).circuit, (9, 9))):
        dc = DeviceCircuit(circ, cuda_device)
        ok, log = dc.jit_status()
        assert ok, log
        out[name] = (circ, dc, groups)
    return out


@pytest.mark.parametrize("name", ["addition", "mario"])
@pytest.mark.parametrize("B", [1, 33, 4097])
def test_jit_forward_is_bit_identical(circuits, name, B):
    from vael_b200 import _lib
    from vael_b200.ops import circuit_forward, circuit_queries_all
    circ, dc, groups = circuits[name]
    facts = synthetic_facts(B, groups=groups, seed=B)
    f = torch.from_numpy(facts).cuda()
    qi = torch.from_numpy(np.arange(B) % circ.n_queries).cuda()
    for kw in (dict(), dict(normalize=True), dict(evidence=True)):
        for query in (qi, 3, None) if "evidence" not in kw else (qi, 3):
            with torch.no_grad():
                wj, qj = circuit_forward(dc, f, query, force_jit=True, **kw)
                assert _lib.last_kernel() == "jit-straightline"
                ws, qs = circuit_forward(dc, f, query, **kw)
                assert _lib.last_kernel() == "ad2-tma"
                wg, qg = circuit_forward(dc, f, query, force_generic=True, **kw)
                assert _lib.last_kernel() == "generic-csr"
            assert torch.equal(wj, ws) and torch.equal(wj, wg)
            if query is not None:
                assert torch.equal(qj, qs) and torch.equal(qj, qg)
    w_ref, q_ref = fp_interp.forward(circ, facts, query=qi.cpu().numpy())
    wj, qj = circuit_forward(dc, f, qi, force_jit=True)
    np.testing.assert_array_equal(wj.cpu().numpy(), w_ref)
    np.testing.assert_array_equal(qj.cpu().numpy()[:, 0], q_ref)
    for normalize in (False, True):
        qa = circuit_queries_all(dc, f, normalize=normalize, force_jit=True)
        assert _lib.last_kernel() == "jit-straightline"
        np.testing.assert_array_equal(qa.cpu().numpy(), fp_interp.queries_all(circ, facts, normalize=normalize))
    _, q_only = circuit_forward(dc, f, qi, force_jit=True, want_worlds=False)     # no worlds requested
    assert torch.equal(q_only, qj)


@pytest.mark.parametrize("name", ["addition", "mario"])
@pytest.mark.parametrize("B", [1, 77, 5000])
def test_jit_adjoint(circuits, name, B):
    from vael_b200 import _lib
    from vael_b200.ops import circuit_forward
    circ, dc, groups = circuits[name]
    rng = np.random.default_rng(B)
    facts = synthetic_facts(B, groups=groups, seed=B + 1)
    qi = rng.integers(0, circ.n_queries, size=B)
    gw = rng.standard_normal((B, circ.n_worlds)).astype(np.float32)
    gq = rng.standard_normal(B).astype(np.float32)
    for normalize in (False, True):
        grads = {}
        for path in ("jit", "structured"):
            f = torch.from_numpy(facts).cuda().requires_grad_(True)
            w, q = circuit_forward(dc, f, torch.from_numpy(qi).cuda(), normalize=normalize, force_jit=path == "jit")
            ((w * torch.from_numpy(gw).cuda()).sum() + (q[:, 0] * torch.from_numpy(gq).cuda()).sum()).backward()
            assert _lib.last_kernel() == ("jit-straightline" if path == "jit" else "ad2-tma")
            grads[path] = f.grad.cpu().numpy()
        g_ref = fp_interp.backward(circ, facts, gw, gq, qi, normalize=normalize)
        floor = 1e-3 * np.abs(g_ref).max()
        assert rel_err(grads["jit"], g_ref, floor=floor) < 1e-4
        assert rel_err(grads["jit"], grads["structured"], floor=floor) < 1e-5
    # query gradient only / world gradient only (null upstream pointers)
    f = torch.from_numpy(facts).cuda().requires_grad_(True)
    _, q = circuit_forward(dc, f, torch.from_numpy(qi).cuda(), force_jit=True, want_worlds=False)
    (q[:, 0] * torch.from_numpy(gq).cuda()).sum().backward()
    g_ref = fp_interp.backward(circ, facts, None, gq, qi)
    assert rel_err(f.grad.cpu().numpy(), g_ref, floor=1e-3 * np.abs(g_ref).max()) < 1e-4
    f = torch.from_numpy(facts).cuda().requires_grad_(True)
    w, _ = circuit_forward(dc, f, None, force_jit=True)
    (w * torch.from_numpy(gw).cuda()).sum().backward()
    g_ref = fp_interp.backward(circ, facts, gw, None, None)
    assert rel_err(f.grad.cpu().numpy(), g_ref, floor=1e-3 * np.abs(g_ref).max()) < 1e-4


def test_jit_unaligned_views_and_empty_batch(circuits):
    from vael_b200.ops import circuit_forward
    circ, dc, _ = circuits["addition"]
    B = 130
    base = torch.from_numpy(synthetic_facts(B + 1, seed=5)).cuda().reshape(-1)
    f = base[1:1 + B * 20].view(B, 20)                 # 4-byte aligned only: the plain-copy path of the tile ring
    assert f.data_ptr() % 16 != 0
    w1, q1 = circuit_forward(dc, f, 9, force_jit=True)
    w2, q2 = circuit_forward(dc, f.clone(), 9, force_jit=True)
    assert torch.equal(w1, w2) and torch.equal(q1, q2)
    e, _ = circuit_forward(dc, torch.empty(0, 20).cuda(), 3, force_jit=True)
    assert e.shape == (0, 100)


@pytest.fixture(scope="module")
def wide(cuda_device):
    from vael_b200.mnist_task import compile_addition_family
    from vael_b200.ops import DeviceCircuit
    circ = compile_addition_family(3, 10).circuit
    dc = DeviceCircuit(circ, cuda_device)
    ok, log = dc.jit_status()
    assert ok, log
    return circ, dc


@pytest.mark.parametrize("B", [1, 45, 2049])
def test_jit_wide_circuit(wide, B):
    """W = 1000: the generated kernels sweep the world level in column chunks, every lane moving its own row segment
    with a bulk copy.  Same bits as the hand-written three-AD kernels and the interpreter; adjoint against fp64."""
    from vael_b200 import _lib
    from vael_b200._lib import VaelError
    from vael_b200.ops import circuit_forward, circuit_queries_all
    circ, dc = wide
    rng = np.random.default_rng(B)
    facts = synthetic_facts(B, groups=(10, 10, 10), seed=B)
    f = torch.from_numpy(facts).cuda()
    qi = torch.from_numpy(rng.integers(0, 28, size=B)).cuda()
    for normalize in (False, True):
        for query in (qi, 13, None):
            with torch.no_grad():
                wj, qj = circuit_forward(dc, f, query, normalize=normalize, force_jit=True)
                assert _lib.last_kernel() == "jit-straightline"
                ws, qs = circuit_forward(dc, f, query, normalize=normalize)
                assert _lib.last_kernel() == "adw-tma2d"
            assert torch.equal(wj, ws)
            if query is not None:
                assert torch.equal(qj, qs)
        qa = circuit_queries_all(dc, f, normalize=normalize, force_jit=True)
        assert _lib.last_kernel() == "jit-straightline"
        np.testing.assert_array_equal(qa.cpu().numpy(), fp_interp.queries_all(circ, facts, normalize=normalize))
    w_ref, q_ref = fp_interp.forward(circ, facts, query=qi.cpu().numpy())
    wj, qj = circuit_forward(dc, f, qi, force_jit=True)
    np.testing.assert_array_equal(wj.cpu().numpy(), w_ref)
    np.testing.assert_array_equal(qj.cpu().numpy()[:, 0], q_ref)
    _, q_only = circuit_forward(dc, f, qi, force_jit=True, want_worlds=False)
    assert torch.equal(q_only, qj)
    gw = rng.standard_normal((B, 1000)).astype(np.float32)
    gq = rng.standard_normal(B).astype(np.float32)
    for normalize in (False, True):
        fr = torch.from_numpy(facts).cuda().requires_grad_(True)
        w, q = circuit_forward(dc, fr, qi, normalize=normalize, force_jit=True)
        ((w * torch.from_numpy(gw).cuda()).sum() + (q[:, 0] * torch.from_numpy(gq).cuda()).sum()).backward()
        assert _lib.last_kernel() == "jit-straightline"
        g_ref = fp_interp.backward(circ, facts, gw, gq, qi.cpu().numpy(), normalize=normalize)
        assert rel_err(fr.grad.cpu().numpy(), g_ref, floor=1e-3 * np.abs(g_ref).max()) < 1e-4
    fr = torch.from_numpy(facts).cuda().requires_grad_(True)     # query gradient only: no [B,W] upstream at all
    _, q = circuit_forward(dc, fr, qi, force_jit=True, want_worlds=False)
    (q[:, 0] * torch.from_numpy(gq).cuda()).sum().backward()
    g_ref = fp_interp.backward(circ, facts, None, gq, qi.cpu().numpy())
    assert rel_err(fr.grad.cpu().numpy(), g_ref, floor=1e-3 * np.abs(g_ref).max()) < 1e-4
    # what the chunked kernels do not cover is refused when forced (and served by other kernels when not)
    with pytest.raises(VaelError) as e:
        with torch.no_grad():
            circuit_forward(dc, f, 5, evidence=True, force_jit=True)
    assert e.value.code == 3


def test_jit_scope_and_fallback(cuda_device):
    """Out of the generator's scope (log semiring; a wide world level whose size is not a multiple of four): no
    specialised kernel, the interpreter runs, and forcing the JIT is an error rather than a silent substitute."""
    from vael_b200 import _lib
    from vael_b200._lib import VaelError
    from vael_b200.circuit import CircuitBuilder, SEMIRING_PROB
    from vael_b200.mario_task import admissible_circuit, compile_mario_family, mario_tables
    from vael_b200.ops import DeviceCircuit, circuit_forward
    dca = DeviceCircuit(admissible_circuit(mario_tables(compile_mario_family((3, 3)))[2]), cuda_device)
    ok, why = dca.jit_status()
    assert not ok and "scope" in why
    b = CircuitBuilder(14, SEMIRING_PROB)                       # 13 x 10 + 1 = 131 worlds
    leaves = [b.leaf(i, True) for i in range(14)]
    worlds = [b.prod((leaves[i % 13], leaves[13 - (i % 2)])) for i in range(0)]
    worlds = [b.prod((leaves[i], leaves[j])) for i in range(12) for j in range(12) if i < j][:65] + \
             [b.sum((leaves[i], leaves[j])) for i in range(12) for j in range(12) if i < j][:66]
    circ = b.finish(worlds, [worlds[0]], -1, fact_labels=[f"f{i}" for i in range(14)],
                    world_atoms=[f"w{i}" for i in range(131)], query_atoms=["q0"])
    dcw = DeviceCircuit(circ, cuda_device)
    assert dcw.fast_kind == 0 and not dcw.jit_status()[0]
    f = torch.rand(50, 14).cuda() * 0.9 + 0.05
    with pytest.raises(VaelError) as e:
        circuit_forward(dcw, f, 0, force_jit=True)
    assert e.value.code == 3
    w, q = circuit_forward(dcw, f, 0)
    assert _lib.last_kernel() == "generic-csr"
    w_ref, q_ref = fp_interp.forward(circ, f.cpu().numpy(), query=np.zeros(50, dtype=np.int64))
    np.testing.assert_array_equal(w.cpu().numpy(), w_ref)


def test_jit_cubin_cache(cuda_device, tmp_path, monkeypatch):
    """VAEL_JIT_CACHE_DIR: the first handle compiles and stores the cubin, the second loads it (no NVRTC run) and
    computes the same bits; a cache file that does not load is dropped and rebuilt."""
    import os
    from vael_b200.mario_task import compile_mario_family
    from vael_b200.ops import DeviceCircuit, circuit_forward
    monkeypatch.setenv("VAEL_JIT_CACHE_DIR", str(tmp_path))
    circ = compile_mario_family((3, 3)).circuit
    f = torch.from_numpy(synthetic_facts(77, groups=(9, 9), seed=5)).cuda()
    q = torch.arange(77, device="cuda") % circ.n_queries

    def run():
        dc = DeviceCircuit(circ, cuda_device)
        ok, log = dc.jit_status()
        assert ok, log
        with torch.no_grad():
            w, qq = circuit_forward(dc, f, q, force_jit=True)
        return log, w, qq

    log1, w1, q1 = run()
    files = [p for p in os.listdir(tmp_path) if p.endswith(".cubin")]
    assert len(files) == 1 and "from cache" not in log1, (files, log1)
    log2, w2, q2 = run()
    assert "from cache" in log2, log2
    assert torch.equal(w1, w2) and torch.equal(q1, q2)
    with open(tmp_path / files[0], "wb") as fh:
        fh.write(b"not a cubin")
    log3, w3, q3 = run()   # the unloadable file is dropped, the circuit compiled again and the cache rewritten
    assert "from cache" not in log3 and "replaced" in log3, log3
    assert torch.equal(w1, w3) and torch.equal(q1, q3)
    assert os.path.getsize(tmp_path / files[0]) > 1000
