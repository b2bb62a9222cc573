"""GPU parity of the stages around the circuit: facts head, Gumbel world sampling + Herbrand
projection, ELBO terms — against the goldens written by the reference and the torch oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import dense_path
from tests.util import rel_err

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


def load(name):
    return np.load(os.path.join(G, name + ".npz"))


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_facts_head_matches_reference_golden(cuda_device):
    from vael_b200.ops import facts_from_logits
    g = load("mnist_dense")
    logits = cu(g["logits"]).requires_grad_(True)
    facts = facts_from_logits(logits, 2, 1e-5)
    assert rel_err(facts.detach().cpu().numpy(), g["facts"].reshape(32, 20)) < 1e-5
    # gradient at the logits (SURVEY.md hard part 3): upstream = golden d loss / d facts
    gf = cu(g["grad_facts_q9"].reshape(32, 20))
    (gl,) = torch.autograd.grad((facts * gf).sum(), [logits])
    ref = g["grad_logits_q9"]
    assert rel_err(gl.cpu().numpy(), ref, floor=1e-3 * np.abs(ref).max()) < 1e-4


@pytest.mark.parametrize("B,groups,n", [(1, 2, 10), (4099, 2, 10), (257, 3, 10), (64, 1, 9)])
def test_facts_head_matches_oracle(cuda_device, B, groups, n):
    from vael_b200.ops import facts_from_logits
    x = 3 * torch.randn(B, groups * n, generator=torch.Generator().manual_seed(B))
    up = torch.randn(B, groups * n, generator=torch.Generator().manual_seed(1))

    def oracle(dtype):
        xr = x.to(dtype).requires_grad_(True)
        ref = dense_path.facts_probability(xr, n_groups=groups).reshape(B, -1)
        (g,) = torch.autograd.grad((ref * up.to(dtype)).sum(), [xr])
        return ref.detach().numpy(), g.numpy()

    ref, gref = oracle(torch.float32)      # what the reference computes
    _, g64 = oracle(torch.float64)         # what it means to compute
    xc = x.cuda().requires_grad_(True)
    out = facts_from_logits(xc, groups, 1e-5)
    (gout,) = torch.autograd.grad((out * up.cuda()).sum(), [xc])
    assert rel_err(out.detach().cpu().numpy(), ref) < 1e-5
    # y_i * (a_i - sum_j y_j a_j) cancels: the reference's own fp32 autograd is ~1.5e-4 away from the fp64
    # gradient under a floor of 1e-3*max, so measure both fp32 evaluations against fp64 and require ours to be
    # within 1e-4 relative (floor 1e-2*max, i.e. a few ulp of the largest entry) and no worse than the reference
    floor = 1e-2 * np.abs(g64).max()
    ours, theirs = rel_err(gout.cpu().numpy(), g64, floor=floor), rel_err(gref, g64, floor=floor)
    assert ours < 1e-4 and ours < 2 * theirs + 1e-5, (ours, theirs)


def test_gumbel_world_matches_reference_golden(cuda_device):
    from vael_b200.ops import gumbel_world
    g = load("gumbel")
    worlds = cu(g["worlds"]).requires_grad_(True)
    hb = dense_path.herbrand_base(20).cuda()
    world_h, idx, y_soft = gumbel_world(worlds, hb, tau=1.0, is_log=False, noise=cu(g["noise"]))
    np.testing.assert_array_equal(idx.cpu().numpy(), g["world"].argmax(1))
    np.testing.assert_allclose(world_h.detach().cpu().numpy(), g["world_h"], rtol=0, atol=2.4e-7)  # (1-y)+y is 1 +- 1 ulp
    (gw,) = torch.autograd.grad((world_h * cu(g["gh"])).sum(), [worlds])
    ref = g["grad_worlds"]
    assert rel_err(gw.cpu().numpy(), ref, floor=1e-4 * np.abs(ref).max()) < 1e-4


# (100, 20), (24, 18), (128, 32), (16, 5), (64, 7): shapes the lane <-> row kernels (gumbel_rows.cu) take when
# VAEL_GUMBEL_ROWS=1 forces them at this batch size; "0": the one-warp-per-row kernels for every shape
@pytest.mark.parametrize("family", ["0", "1"])
@pytest.mark.parametrize("W,H,is_log", [(100, 20, False), (24, 18, True), (81, 18, False), (1000, 30, False), (7, 3, True),
                                        (128, 32, False), (16, 5, True), (64, 7, False), (100, 33, False)])
def test_gumbel_world_matches_oracle(cuda_device, W, H, is_log, family, monkeypatch):
    from vael_b200.ops import gumbel_world
    monkeypatch.setenv("VAEL_GUMBEL_ROWS", family)
    B = 193
    gen = torch.Generator().manual_seed(W)
    p = torch.softmax(2 * torch.randn(B, W, generator=gen), 1)
    scores = torch.log(p) if is_log else p
    noise = -torch.empty(B, W).exponential_(generator=gen).log()
    hb = (torch.rand(W, H, generator=gen) < 0.3).float()
    up = torch.randn(B, H, generator=gen)
    sr = scores.clone().requires_grad_(True)
    world, y_soft, idx = dense_path.gumbel_hard(sr if is_log else torch.log(sr), noise, tau=1.0)
    wh_ref = world @ hb
    (g_ref,) = torch.autograd.grad((wh_ref * up).sum(), [sr])
    sc = scores.cuda().requires_grad_(True)
    wh, i2, ys = gumbel_world(sc, hb.cuda(), tau=1.0, is_log=is_log, noise=noise.cuda())
    np.testing.assert_array_equal(i2.cpu().numpy(), idx.numpy())
    np.testing.assert_allclose(ys.cpu().numpy(), y_soft.detach().numpy(), rtol=2e-5, atol=1e-9)
    np.testing.assert_allclose(wh.detach().cpu().numpy(), wh_ref.detach().numpy(), atol=2.4e-7)
    (g2,) = torch.autograd.grad((wh * up.cuda()).sum(), [sc])
    assert rel_err(g2.cpu().numpy(), g_ref.numpy(), floor=1e-3 * g_ref.abs().max().item()) < 1e-4


@pytest.mark.parametrize("tau", [1.0, 0.4, 2.5])
def test_gumbel_rows_kernels_equal_warp_per_row_kernels(cuda_device, tau, monkeypatch):
    """The two kernel families on the same inputs (real Herbrand base, injected noise, a temperature, a tail tile, views
    that are only 4-byte aligned): same worlds, y_soft within 1e-5, adjoints within 1e-5 of the largest entry."""
    from vael_b200 import _lib
    from vael_b200.ops import gumbel_world
    gen = torch.Generator().manual_seed(int(tau * 10))
    B, W, H = 32 * 9 + 5, 100, 20
    p = torch.softmax(3 * torch.randn(B, W, generator=gen), 1)
    noise = -torch.empty(B, W).exponential_(generator=gen).log()
    up = torch.randn(B, H, generator=gen)
    hb = dense_path.herbrand_base(20).cuda()

    def run(offset):
        base = torch.empty(B * W + 1).cuda()
        sc = (base[1:] if offset else base[:-1]).view(B, W)
        sc.copy_(p)
        sc = sc.detach().requires_grad_(True)
        assert (sc.data_ptr() % 16 != 0) == bool(offset)
        wh, idx, ys = gumbel_world(sc, hb, tau=tau, is_log=False, noise=noise.cuda())
        (g,) = torch.autograd.grad((wh * up.cuda()).sum(), [sc])
        return wh.detach(), idx, ys, g

    monkeypatch.setenv("VAEL_GUMBEL_ROWS", "1")
    a, b = run(0), run(1)
    for x, y in zip(a, b):
        assert torch.equal(x, y)
    monkeypatch.setenv("VAEL_GUMBEL_ROWS", "0")
    o = run(0)
    assert torch.equal(a[1], o[1])
    np.testing.assert_allclose(a[2].cpu().numpy(), o[2].cpu().numpy(), rtol=1e-5, atol=1e-12)
    np.testing.assert_allclose(a[0].cpu().numpy(), o[0].cpu().numpy(), atol=2.4e-7)
    gm = o[3].abs().max().item()
    assert (a[3] - o[3]).abs().max().item() <= 1e-5 * gm
    # the default: the lane <-> row kernels from 16 Ki rows up (same worlds as the forced run at a large batch)
    monkeypatch.delenv("VAEL_GUMBEL_ROWS")
    big = p.repeat(64, 1).cuda()
    _, i_auto, _ = gumbel_world(big, hb, tau=tau, seed=3)
    monkeypatch.setenv("VAEL_GUMBEL_ROWS", "1")
    _, i_rows, _ = gumbel_world(big, hb, tau=tau, seed=3)
    assert torch.equal(i_auto, i_rows)


def test_gumbel_rows_philox_sampling_follows_world_distribution(cuda_device, monkeypatch):
    """The lane <-> row kernel's own Philox stream (W = 100, the training shape): deterministic per seed, worlds
    distributed like P(w), y_soft a distribution whose arg-max is the sampled world."""
    from vael_b200.ops import gumbel_world
    monkeypatch.setenv("VAEL_GUMBEL_ROWS", "1")
    W, B = 100, 600000
    p = torch.softmax(1.5 * torch.randn(1, W, generator=torch.Generator().manual_seed(5)), 1)
    scores = p.expand(B, W).contiguous().cuda()
    hb = dense_path.herbrand_base(20).cuda()
    wh, idx, ys = gumbel_world(scores, hb, seed=99)
    _, idx2, _ = gumbel_world(scores, hb, seed=99)
    _, idx3, _ = gumbel_world(scores, hb, seed=100)
    assert torch.equal(idx, idx2) and not torch.equal(idx, idx3)
    assert torch.equal(ys.argmax(1).int(), idx)
    assert float((ys.sum(1) - 1).abs().max()) < 1e-5
    assert torch.equal((wh > 0.5).float(), hb[idx.long()])
    counts = torch.bincount(idx.long(), minlength=W).double().cpu()
    expected = p[0].double() * B
    chi2 = ((counts - expected) ** 2 / expected).sum().item()
    assert chi2 < 190.0, chi2  # 99 dof: P(chi2 > 190) ~ 1e-7


def test_gumbel_world_philox_sampling_follows_world_distribution(cuda_device):
    """In-kernel Philox noise cannot match torch's generator bit for bit; check the statistics:
    world frequencies against P(w) with a chi-square bound, and determinism per seed."""
    from vael_b200.ops import gumbel_world
    W, B = 20, 400000
    p = torch.softmax(torch.randn(1, W, generator=torch.Generator().manual_seed(3)), 1)
    scores = p.expand(B, W).contiguous().cuda()
    hb = torch.eye(W).cuda()
    _, idx, _ = gumbel_world(scores, hb, seed=1234)
    _, idx2, _ = gumbel_world(scores, hb, seed=1234)
    _, idx3, _ = gumbel_world(scores, hb, seed=1235)
    assert torch.equal(idx, idx2) and not torch.equal(idx, idx3)
    counts = torch.bincount(idx.long(), minlength=W).double().cpu()
    expected = p[0].double() * B
    chi2 = ((counts - expected) ** 2 / expected).sum().item()
    assert chi2 < 60.0, chi2  # 19 dof: P(chi2 > 60) ~ 4e-6


def test_elbo_terms_match_reference_golden(cuda_device):
    from vael_b200.ops import elbo_terms
    g = load("elbo")
    rw, kw, qw = (float(v) for v in g["weights"])
    leaves = [cu(g[k]).requires_grad_(True) for k in ("recon", "mu", "logvar", "qprob")]
    loss, terms = elbo_terms(leaves[0], cu(g["x"]), leaves[1], leaves[2], leaves[3], rw, kw, qw)
    np.testing.assert_allclose(terms.cpu().numpy(), g["terms"], rtol=2e-6)
    assert abs(loss.item() - g["terms"][0]) <= 2e-6 * abs(g["terms"][0])
    grads = torch.autograd.grad(loss, leaves)
    for got, key in zip(grads, ("g_recon", "g_mu", "g_logvar", "g_q")):
        ref = g[key]
        assert rel_err(got.cpu().numpy(), ref, floor=1e-6 * np.abs(ref).max() + 1e-12) < 1e-4, key
    # determinism of the two-stage reduction
    loss2, terms2 = elbo_terms(leaves[0], cu(g["x"]), leaves[1], leaves[2], leaves[3], rw, kw, qw)
    assert torch.equal(terms, terms2)


@pytest.mark.parametrize("B,D,L", [(1, 7, 3), (300, 1568, 23), (4096, 30000 // 8, 48)])
def test_elbo_terms_match_oracle(cuda_device, B, D, L):
    from vael_b200.ops import elbo_terms
    gen = torch.Generator().manual_seed(B)
    x = torch.randn(B, 1, 1, D, generator=gen)
    recon = torch.sigmoid(torch.randn(B, 1, 1, D, generator=gen))
    mu, logvar = torch.randn(B, L, generator=gen), 0.5 * torch.randn(B, L, generator=gen)
    q = torch.rand(B, 1, generator=gen).clamp_min(1e-6)
    ref = dense_path.elbo_terms(recon, x, mu, logvar, q, 0.1, 1e-5, 1.0)
    loss, terms = elbo_terms(recon.cuda(), x.cuda(), mu.cuda(), logvar.cuda(), q.cuda(), 0.1, 1e-5, 1.0)
    np.testing.assert_allclose(terms.cpu().numpy(), [t.item() for t in ref], rtol=1e-5)
