"""GPU parity of the round-2 kernels: the fused training-mode logic kernel (logits -> facts -> worlds -> query ->
Gumbel arg-max -> Herbrand vector, and its adjoint), the clipped facts head of the Mario model, the MSE / BCE
reconstruction terms, the query-and-gradient host entry point, the per-row query-index check flag and Mario's
conditioned worlds on the evidence kernel — against the reference's goldens and the torch oracle."""
import ctypes as C
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


@pytest.fixture(scope="module")
def dc2(cuda_device):
    from vael_b200.mnist_task import compile_addition_family
    from vael_b200.ops import DeviceCircuit
    return DeviceCircuit(compile_addition_family(2, 10).circuit, cuda_device)


def unfused_chain(dc, logits, query, noise, want_grad_of=None):
    """The stage-by-stage kernels the fused one replaces."""
    from vael_b200.ops import circuit_forward, facts_from_logits, gumbel_world
    facts = facts_from_logits(logits, 2, 1e-5)
    worlds, q = circuit_forward(dc, facts, query)
    hb = dense_path.herbrand_base(20).cuda()
    world_h, idx, y_soft = gumbel_world(worlds, hb, tau=1.0, is_log=False, noise=noise)
    return facts, q, world_h, idx, y_soft


def test_fused_logic_matches_reference_golden(cuda_device, dc2):
    """The reference's own numbers (tests/golden/mnist_dense.npz + gumbel.npz: logits -> facts -> worlds/query for
    q = 9 -> gumbel_softmax(hard) with the captured noise -> herbrand), one kernel here."""
    from vael_b200.ops import logic_train
    g, gg = load("mnist_dense"), load("gumbel")
    logits = cu(g["logits"]).requires_grad_(True)
    facts, q, world_h, idx, _ = logic_train(dc2, logits, 9, noise=cu(gg["noise"]))
    assert rel_err(facts.detach().cpu().numpy(), g["facts"].reshape(32, 20)) < 1e-5
    assert rel_err(q.detach().cpu().numpy(), g["query_q9"]) < 1e-5
    np.testing.assert_array_equal(idx.cpu().numpy(), gg["world"].argmax(1))
    np.testing.assert_allclose(world_h.detach().cpu().numpy(), gg["world_h"], rtol=0, atol=2.4e-7)
    # gradients at the logits against autograd through the oracle's restatement of the same chain (CPU, fp64)
    gh, gq = torch.from_numpy(gg["gh"]), torch.from_numpy(g["gq"])
    (gl,) = torch.autograd.grad((world_h * gh.cuda()).sum() + (q * gq.cuda()).sum(), [logits])
    x64 = torch.from_numpy(g["logits"]).double().requires_grad_(True)
    f64 = dense_path.facts_probability(x64, n_groups=2)
    q64, w64 = dense_path.dense_inference(f64, dense_path.worlds_queries_matrix().double(), 9)
    world, _, _ = dense_path.gumbel_hard(torch.log(w64), torch.from_numpy(gg["noise"]).double(), tau=1.0)
    wh64 = world @ dense_path.herbrand_base(20).double()
    (ref,) = torch.autograd.grad((wh64 * gh.double()).sum() + (q64 * gq.double()).sum(), [x64])
    assert rel_err(gl.cpu().numpy(), ref.numpy(), floor=1e-3 * ref.abs().max().item()) < 1e-4


@pytest.mark.parametrize("B", [1, 33, 4096, 100003])
def test_fused_logic_equals_the_stage_kernels(cuda_device, dc2, B):
    """Same inputs, same injected noise: facts and query bit for bit (same arithmetic), the sampled world equal
    wherever the two largest perturbed logits are not within rounding of each other, y_soft within 2e-5, and the
    adjoint (upstream gradients on world_h, query AND facts) within 1e-4 of autograd through the three stages."""
    from vael_b200.ops import logic_train
    gen = torch.Generator().manual_seed(B)
    logits = (3 * torch.randn(B, 20, generator=gen)).cuda()
    query = torch.randint(0, 19, (B,), generator=gen).cuda()
    noise = (-torch.empty(B, 100).exponential_(generator=gen).log()).cuda()
    gh, gq, gf = (torch.randn(B, 20, generator=gen).cuda(), torch.randn(B, 1, generator=gen).cuda(),
                  torch.randn(B, 20, generator=gen).cuda())
    la = logits.clone().requires_grad_(True)
    fa, qa, wha, ia, ysa = logic_train(dc2, la, query, noise=noise, want_y_soft=True)
    lb = logits.clone().requires_grad_(True)
    fb, qb, whb, ib, ysb = unfused_chain(dc2, lb, query, noise)
    assert torch.equal(fa, fb) and torch.equal(qa, qb)
    same = ia == ib
    top2 = torch.topk(torch.log(ysb), 2, dim=1).values
    assert bool((same | ((top2[:, 0] - top2[:, 1]) < 1e-5)).all())
    assert same.float().mean().item() > 0.999
    np.testing.assert_allclose(ysa[same].cpu().numpy(), ysb[same].cpu().numpy(), rtol=2e-5, atol=1e-9)
    assert torch.allclose(wha[same], whb[same], atol=2.4e-7, rtol=0)
    # The adjoint: y_i (a_i - sum_j y_j a_j) cancels, so two fp32 evaluations differ by more than 1e-4 of small
    # entries; both are measured against fp64 autograd through the oracle's restatement of the reference chain
    # (the test_facts_head_matches_oracle protocol): ours within 1e-4 (floor 1e-2 * max) and no worse than twice
    # the stage kernels.  All three upstreams together, then each alone (null pointers on the others).
    x64 = logits.double().cpu().requires_grad_(True)
    f64 = dense_path.facts_probability(x64, n_groups=2)
    q64, w64 = dense_path.dense_inference_rows(f64, dense_path.worlds_queries_matrix().double(), query.cpu().long())
    world64, _, _ = dense_path.gumbel_hard(torch.log(w64), noise.double().cpu(), tau=1.0)
    wh64 = world64 @ dense_path.herbrand_base(20).double()
    outs64 = (f64.reshape(B, 20), q64, wh64)
    keep = same.cpu().numpy()
    for sel in ((0, 1, 2), (2,), (1,), (0,)):
        ups = (gf, gq, gh)
        (g64,) = torch.autograd.grad(sum((outs64[k] * ups[k].double().cpu()).sum() for k in sel), [x64], retain_graph=True)
        lc = logits.clone().requires_grad_(True)
        oc = logic_train(dc2, lc, query, noise=noise)
        (g1,) = torch.autograd.grad(sum((oc[k] * ups[k]).sum() for k in sel), [lc])
        ld = logits.clone().requires_grad_(True)
        od = unfused_chain(dc2, ld, query, noise)
        (g2,) = torch.autograd.grad(sum((od[k] * ups[k]).sum() for k in sel), [ld])
        g64n = g64.numpy()[keep]
        floor = 1e-2 * np.abs(g64n).max() + 1e-12
        ours, theirs = rel_err(g1.cpu().numpy()[keep], g64n, floor=floor), rel_err(g2.cpu().numpy()[keep], g64n, floor=floor)
        assert ours < 1e-4 and ours < 2 * theirs + 2e-5, (sel, ours, theirs)


@pytest.mark.parametrize("tau", [0.3, 2.0])
def test_fused_logic_temperature(cuda_device, dc2, tau):
    """tau != 1 (F.gumbel_softmax's temperature): against the stage kernels with the same injected noise."""
    from vael_b200.ops import circuit_forward, facts_from_logits, gumbel_world, logic_train
    gen = torch.Generator().manual_seed(3)
    B = 2000
    logits = (2 * torch.randn(B, 20, generator=gen)).cuda()
    noise = (-torch.empty(B, 100).exponential_(generator=gen).log()).cuda()
    gh = torch.randn(B, 20, generator=gen).cuda()
    la = logits.clone().requires_grad_(True)
    _, _, wha, ia, ysa = logic_train(dc2, la, 4, noise=noise, tau=tau, want_y_soft=True)
    lb = logits.clone().requires_grad_(True)
    worlds, _ = circuit_forward(dc2, facts_from_logits(lb, 2, 1e-5), 4)
    whb, ib, ysb = gumbel_world(worlds, dense_path.herbrand_base(20).cuda(), tau=tau, is_log=False, noise=noise)
    same = ia == ib
    assert same.float().mean().item() > 0.995
    np.testing.assert_allclose(ysa[same].cpu().numpy(), ysb[same].cpu().numpy(), rtol=5e-5, atol=1e-9)
    (ga,) = torch.autograd.grad((wha * gh).sum(), [la])
    (gb,) = torch.autograd.grad((whb * gh).sum(), [lb])
    ga, gb = ga[same].cpu().numpy(), gb[same].cpu().numpy()
    assert rel_err(ga, gb, floor=1e-2 * np.abs(gb).max()) < 2e-4


def test_fused_logic_scalar_query_and_unaligned_views(cuda_device, dc2):
    from vael_b200.ops import logic_train
    gen = torch.Generator().manual_seed(5)
    B = 257
    base = (3 * torch.randn(B * 20 + 1, generator=gen)).cuda()
    logits = base[1:].view(B, 20)                       # 4-byte aligned only: the scalar load path
    assert logits.data_ptr() % 16 != 0
    noise = (-torch.empty(B, 100).exponential_(generator=gen).log()).cuda()
    a = logic_train(dc2, logits.clone(), 7, noise=noise)
    b = logic_train(dc2, logits, 7, noise=noise)        # .contiguous() keeps the view's storage offset
    for x, y in zip(a[:4], b[:4]):
        assert torch.equal(x, y)
    f, q, wh, idx, _ = logic_train(dc2, logits, None, noise=noise)     # no query requested
    assert torch.equal(f, a[0]) and float(q.abs().max()) == 0.0 and torch.equal(idx, a[3])
    with pytest.raises(KeyError):
        logic_train(dc2, logits, 19, noise=noise)
    with pytest.raises(RuntimeError):
        logic_train(dc2, logits.cpu(), 7)
    e = logic_train(dc2, torch.empty(0, 20).cuda(), 3)
    assert e[0].shape == (0, 20) and e[3].shape == (0,)


def test_fused_logic_philox_mode(cuda_device, dc2):
    """In-kernel noise: deterministic per seed, distributed like P(w) (chi-square), and the adjoint — which RE-DRAWS
    the noise — agrees with the stage-wise Gumbel adjoint fed the y_soft the forward produced."""
    from vael_b200 import _lib
    from vael_b200.ops import _p, _stream_ptr, circuit_forward, facts_from_logits, logic_train
    gen = torch.Generator().manual_seed(11)
    row = 1.5 * torch.randn(1, 20, generator=gen)
    B = 400000
    logits = row.expand(B, 20).contiguous().cuda()
    _, _, _, i1, _ = logic_train(dc2, logits, 9, seed=77)
    _, _, _, i2, _ = logic_train(dc2, logits, 9, seed=77)
    _, _, _, i3, _ = logic_train(dc2, logits, 9, seed=78)
    assert torch.equal(i1, i2) and not torch.equal(i1, i3)
    sd = torch.tensor([77], dtype=torch.int64, device="cuda")
    _, _, _, i4, _ = logic_train(dc2, logits, 9, seed_dev=sd)         # the device word overrides the host seed
    assert torch.equal(i1, i4)
    # the production kernel (FAST: two uniforms pick the world, no per-world variates in the forward) and the two-pass
    # one (GENERAL, taken when y_soft is asked for) make the same draws: identical worlds for every tau the fast kernel
    # accepts, and the sampled world is the arg-max of the y_soft the GENERAL kernel builds around it
    for tau in (1.0, 0.5, 3.0):
        fa, qa, wha, ia, _ = logic_train(dc2, logits[:50000], 9, seed=5, tau=tau)
        fb, qb, whb, ib, ys = logic_train(dc2, logits[:50000], 9, seed=5, tau=tau, want_y_soft=True)
        assert torch.equal(ia, ib) and torch.equal(fa, fb) and torch.equal(qa, qb)
        assert torch.allclose(wha, whb, atol=2.4e-7, rtol=0)
        assert torch.allclose(ys.sum(1), torch.ones(50000, device="cuda"), atol=1e-5)
        assert torch.equal(ys.argmax(1).int(), ib)
    facts = facts_from_logits(logits[:1], 2, 1e-5)
    p = circuit_forward(dc2, facts, None)[0][0].double().cpu()
    counts = torch.bincount(i1.long(), minlength=100).double().cpu()
    expected = p * B
    keep = expected > 20                                              # pool the rare worlds
    chi2 = ((counts[keep] - expected[keep]) ** 2 / expected[keep]).sum().item()
    chi2 += (counts[~keep].sum() - expected[~keep].sum()).item() ** 2 / max(expected[~keep].sum().item(), 1.0)
    dof = int(keep.sum())
    assert chi2 < dof + 6 * (2 * dof) ** 0.5 + 10, (chi2, dof)
    # adjoint consistency
    Bs = 3001
    lg = (3 * torch.randn(Bs, 20, generator=gen)).cuda().requires_grad_(True)
    query = torch.randint(0, 19, (Bs,), generator=gen).cuda()
    f, q, wh, idx, ys = logic_train(dc2, lg, query, seed=4242, want_y_soft=True)
    gh, gq = torch.randn(Bs, 20, generator=gen).cuda(), torch.randn(Bs, 1, generator=gen).cuda()
    (g_fused,) = torch.autograd.grad((wh * gh).sum() + (q * gq).sum(), [lg])
    lg2 = lg.detach().clone().requires_grad_(True)
    f2 = facts_from_logits(lg2, 2, 1e-5)
    worlds, q2 = circuit_forward(dc2, f2, query)
    hb = dense_path.herbrand_base(20).cuda()
    gs = torch.empty_like(worlds)
    lib = _lib.load()
    _lib.check(lib.vael_gumbel_world_backward(_p(worlds.detach()), _p(ys), _p(gh), Bs, 100, 0, 1.0, _p(hb), 20, _p(gs),
                                              _stream_ptr()))
    (g_ref,) = torch.autograd.grad((worlds * gs).sum() + (q2 * gq).sum(), [lg2])
    assert rel_err(g_fused.cpu().numpy(), g_ref.cpu().numpy(), floor=1e-2 * g_ref.abs().max().item()) < 2e-4


@pytest.mark.parametrize("tau", [1.0, 0.6, 2.0])
def test_fused_philox_noise_has_the_law_of_gumbel_softmax(cuda_device, dc2, tau):
    """The Philox mode draws the race times top-down (winner by inverse CDF, the others as winner + Exp / p): the
    JOINT law of (sampled world, y_soft) must be that of F.gumbel_softmax.  Same row 300 000 times through the kernel's
    own noise and through injected torch Gumbel noise (the reference's computation): world frequencies, E[y_soft],
    E[y_soft^2], E[y_soft of the sampled world] and E[y_soft | world = w] for the likeliest worlds agree within
    sampling error."""
    from vael_b200.ops import logic_train
    gen = torch.Generator().manual_seed(23)
    row = 1.2 * torch.randn(1, 20, generator=gen)
    B = 300000
    logits = row.expand(B, 20).contiguous().cuda()
    _, _, _, ia, ya = logic_train(dc2, logits, 9, seed=1234, tau=tau, want_y_soft=True)
    noise = -torch.empty(B, 100, device="cuda").exponential_(generator=torch.Generator(device="cuda").manual_seed(77)).log()
    _, _, _, ib, yb = logic_train(dc2, logits, 9, noise=noise, tau=tau, want_y_soft=True)
    ya, yb = ya.double(), yb.double()
    se = 5.0 / B ** 0.5                                   # five standard errors of a [0,1]-valued mean (sd <= 1/2 ... 1)
    fa = torch.bincount(ia.long(), minlength=100).double() / B
    fb = torch.bincount(ib.long(), minlength=100).double() / B
    assert float((fa - fb).abs().max()) < 2 * se
    assert float((ya.mean(0) - yb.mean(0)).abs().max()) < 2 * se
    assert float(((ya ** 2).mean(0) - (yb ** 2).mean(0)).abs().max()) < 2 * se
    sa = ya.gather(1, ia.long()[:, None]).mean().item(); sb = yb.gather(1, ib.long()[:, None]).mean().item()
    assert abs(sa - sb) < 2 * se, (sa, sb)
    for w in torch.topk(fb, 3).indices.tolist():           # conditional on the sampled world (n ~ B f_w rows each)
        ca, cb = ya[ia == w].mean(0), yb[ib == w].mean(0)
        n = min(int((ia == w).sum()), int((ib == w).sum()))
        assert float((ca - cb).abs().max()) < 2 * 5.0 / n ** 0.5, (w, n)


def test_model_forward_fused_equals_stagewise(cuda_device):
    """MNISTPairsVAELModel.forward with fused_logic on / off: same outputs, same parameter gradients, the lazily
    materialised side-effect attributes (worlds_prob, world) agree."""
    from vael_b200.mnist_task import build_model_dict, build_worlds_queries_matrix
    from vael_b200.networks import MNISTPairsDecoder, MNISTPairsEncoder, MNISTPairsMLP
    from vael_b200.train import loss_function
    from vael_b200.vael import MNISTPairsVAELModel
    torch.manual_seed(0)
    model = MNISTPairsVAELModel(MNISTPairsEncoder(hidden_channels=64, latent_dim=23),
                                MNISTPairsDecoder(label_dim=20, hidden_channels=64, latent_dim=8),
                                MNISTPairsMLP(in_features=15, n_facts=20), (15, 8), build_model_dict(2, 10, device=cuda_device),
                                build_worlds_queries_matrix(2, 10).to(cuda_device), dropout=None, is_train=True,
                                device=cuda_device, query_per_row=True).to(cuda_device)
    model.eval()          # no dropout masks: the two passes must see the same network (reparametrize follows is_train)
    B = 64
    gen = torch.Generator().manual_seed(2)
    x = torch.randn(B, 1, 28, 56, generator=gen).cuda()
    d = torch.randint(0, 10, (B, 2), generator=gen)
    labels = torch.cat([d, d.sum(1, keepdim=True), -torch.ones(B, 1, dtype=torch.long)], 1).cuda()
    model.reparam_noise = torch.randn(B, 23, generator=gen).cuda()
    model.gumbel_noise = (-torch.empty(B, 100).exponential_(generator=gen).log()).cuda()
    res = {}
    for fused in (True, False):
        model.fused_logic = fused
        model.zero_grad()
        image, mu, logvar, qp = model(x, labels)
        loss, *_ = loss_function(image, x, mu, logvar, qp, recon_w=1e-1, kl_w=1e-5, query_w=1.0, sup_w=0.0,
                                 rec_loss="LAPLACE")
        loss.backward()
        res[fused] = (image.detach(), qp.detach(), model.facts_probs.detach(), model.worlds_prob.detach(),
                      model.world.detach(), model.world_idx.clone(),
                      {n: p.grad.clone() for n, p in model.named_parameters()})
    a, b = res[True], res[False]
    assert torch.equal(a[5], b[5]) and torch.equal(a[1], b[1]) and torch.equal(a[2], b[2]) and torch.equal(a[3], b[3])
    assert torch.allclose(a[0], b[0], atol=1e-6) and torch.allclose(a[4], b[4], atol=2.4e-7)
    for n in a[6]:
        ref = b[6][n]
        assert float((a[6][n] - ref).abs().max()) <= 1e-4 * float(ref.abs().max()) + 1e-9, n


@pytest.mark.parametrize("B,groups", [(1, 1), (77, 2), (5000, 2), (300, 1)])
def test_clipped_facts_head(cuda_device, B, groups):
    """MarioVAELModel's facts (models/vael.py:319-322): softmax + torch.clip(1e-5, 0.9999), and its gradient."""
    from vael_b200.ops import facts_from_logits_clipped
    gen = torch.Generator().manual_seed(B)
    x = 6 * torch.randn(B, groups * 9, generator=gen)          # wide logits: both clip bounds are hit
    up = torch.randn(B, groups * 9, generator=gen)
    xr = x.double().requires_grad_(True)
    ref = torch.cat([torch.clip(torch.softmax(c, 1), min=1e-5, max=0.9999) for c in torch.split(xr, 9, dim=1)], 1)
    (g_ref,) = torch.autograd.grad((ref * up.double()).sum(), [xr])
    xc = x.cuda().requires_grad_(True)
    out = facts_from_logits_clipped(xc, groups, 1e-5, 0.9999)
    (g,) = torch.autograd.grad((out * up.cuda()).sum(), [xc])
    o = out.detach().cpu().numpy()
    assert o.min() >= np.float32(1e-5) and o.max() <= np.float32(0.9999)
    assert rel_err(o, ref.detach().numpy()) < 1e-5
    # entries whose softmax sits within rounding of a clip bound may pass / block the gradient differently in fp32
    y = torch.cat([torch.softmax(c, 1) for c in torch.split(x.double(), 9, dim=1)], 1)
    near = ((y - 1e-5).abs() < 1e-9) | ((y - 0.9999).abs() < 1e-6)
    rows = ~near.any(1)
    gn, gr = g.cpu().numpy()[rows.numpy()], g_ref.numpy()[rows.numpy()]
    if gr.size:   # y_i (a_i - sum_j y_j a_j) cancels: floor as in test_facts_head_matches_oracle
        assert rel_err(gn, gr, floor=1e-2 * np.abs(gr).max() + 1e-12) < 1e-4
    # a width the tile kernels do not cover takes the generic one-thread-per-group kernel
    x7 = 4 * torch.randn(B, 14, generator=gen)
    o7 = facts_from_logits_clipped(x7.cuda(), 2, 1e-5, 0.9999).cpu()
    r7 = torch.cat([torch.clip(torch.softmax(c, 1), min=1e-5, max=0.9999) for c in torch.split(x7, 7, dim=1)], 1)
    assert rel_err(o7.numpy(), r7.numpy()) < 1e-5


@pytest.mark.parametrize("rec_loss", ["MSE", "BCE", "LAPLACE"])
def test_elbo_reconstruction_kinds_match_torch(cuda_device, rec_loss):
    """loss_function for each rec_loss of the reference (utils/mnist_utils/train.py:16-21) against the same torch
    expressions in fp64; CPU tensors are refused."""
    from vael_b200.train import loss_function
    gen = torch.Generator().manual_seed(1)
    B = 48
    x = torch.rand(B, 1, 28, 56, generator=gen)
    recon = torch.sigmoid(torch.randn(B, 1, 28, 56, generator=gen))
    mu, logvar = torch.randn(B, 23, generator=gen), 0.5 * torch.randn(B, 23, generator=gen)
    q = torch.rand(B, 1, generator=gen).clamp_min(1e-4)
    leaves = [t.double().requires_grad_(True) for t in (recon, mu, logvar, q)]
    r, m, lv, qq = leaves
    if rec_loss == "BCE":
        rl = torch.nn.BCELoss(reduction="mean")(torch.flatten(r), torch.flatten(x.double()))
    elif rec_loss == "MSE":
        rl = torch.nn.MSELoss(reduction="mean")(torch.flatten(r), torch.flatten(x.double()))
    else:
        rl = -torch.distributions.Laplace(r, torch.ones_like(r)).log_prob(x.double()).sum(dim=(1, 2, 3)).mean()
    kl = torch.mean(-0.5 * torch.sum(1. + lv - m.pow(2) - lv.exp(), dim=1))
    qce = torch.nn.BCELoss(reduction="mean")(torch.flatten(qq), torch.ones(B).double())
    ref_loss = 0.3 * rl + 0.01 * kl + 2.0 * qce
    ref_grads = torch.autograd.grad(ref_loss, leaves)
    cl = [t.cuda().requires_grad_(True) for t in (recon, mu, logvar, q)]
    loss, rl2, kl2, qce2, _ = loss_function(cl[0], x.cuda(), cl[1], cl[2], cl[3], query=True, recon_w=0.3, kl_w=0.01,
                                            query_w=2.0, sup_w=0.0, rec_loss=rec_loss)
    for got, want in ((loss, ref_loss), (rl2, rl), (kl2, kl), (qce2, qce)):
        assert abs(got.item() - want.item()) <= 1e-5 * abs(want.item()) + 1e-7
    grads = torch.autograd.grad(loss, cl)
    for got, want in zip(grads, ref_grads):
        assert rel_err(got.cpu().numpy(), want.numpy(), floor=1e-4 * want.abs().max().item() + 1e-12) < 1e-4
    with pytest.raises(RuntimeError):
        loss_function(recon, x, mu, logvar, q, rec_loss=rec_loss)
    with pytest.raises(ValueError):
        loss_function(cl[0], x.cuda(), cl[1], cl[2], cl[3], rec_loss="L1")


def test_query_grad_host_entry_point(cuda_device, dc2):
    """vael_circuit_query_grad_host: facts + query index (+ upstream) in, query + grad_facts out, nothing [B,W]
    over PCIe; equals the device entry points bit for bit; several host threads on ONE handle."""
    import threading
    from tests.util import synthetic_facts
    from vael_b200 import _lib
    from vael_b200.ops import _p, circuit_forward
    lib = _lib.load()
    n = 300001
    facts = torch.from_numpy(synthetic_facts(n, seed=4))
    qidx = torch.randint(0, 19, (n,), generator=torch.Generator().manual_seed(4), dtype=torch.int32)
    gq = torch.randn(n, generator=torch.Generator().manual_seed(5))
    fd = facts.cuda().requires_grad_(True)
    _, q = circuit_forward(dc2, fd, qidx.cuda(), want_worlds=False)
    (gd,) = torch.autograd.grad((q[:, 0] * gq.cuda()).sum(), [fd])

    def call():
        hq, hg = torch.empty(n), torch.empty(n, 20)
        rc = lib.vael_circuit_query_grad_host(dc2.handle, _p(facts), _p(qidx), -1, n, _p(gq), _p(hq), _p(hg), 0)
        return rc, hq, hg

    rc, hq, hg = call()
    assert rc == 0 and torch.equal(hq, q.detach()[:, 0].cpu()) and torch.equal(hg, gd.cpu())
    out = [None] * 4
    ts = [threading.Thread(target=lambda i=i: out.__setitem__(i, call())) for i in range(4)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    for rc, a, b in out:
        assert rc == 0 and torch.equal(a, hq) and torch.equal(b, hg)
    assert lib.vael_circuit_query_grad_host(dc2.handle, _p(facts), _p(qidx), -1, n, None, _p(hq), _p(hg), 0) == 1


def test_query_index_check_flag(cuda_device, dc2):
    """VAEL_FLAG_CHECK_QUERY: an out-of-range per-row index is an error instead of a silent 0; -1 stays the
    documented 'no query' sentinel."""
    from tests.util import synthetic_facts
    from vael_b200._lib import VaelError
    from vael_b200.ops import circuit_forward
    B = 5000
    f = torch.from_numpy(synthetic_facts(B, seed=9)).cuda()
    q = torch.randint(0, 19, (B,), generator=torch.Generator().manual_seed(9), dtype=torch.int32).cuda()
    _, ok = circuit_forward(dc2, f, q, check_query=True)
    q2 = q.clone(); q2[17] = -1
    _, sent = circuit_forward(dc2, f, q2, check_query=True)
    assert float(sent[17]) == 0.0 and torch.equal(sent[18:], ok[18:])
    for bad in (19, 1000, -2):
        q3 = q.clone(); q3[B - 1] = bad
        with pytest.raises(VaelError) as e:
            circuit_forward(dc2, f, q3, check_query=True)
        assert e.value.code == 1
        _, silent = circuit_forward(dc2, f, q3)                  # unchecked: that row's query is 0
        assert float(silent[B - 1]) == 0.0
    _, again = circuit_forward(dc2, f, q, check_query=True)      # the flag word was reset
    assert torch.equal(again, ok)


def test_mario_conditioned_worlds_on_the_evidence_kernel(cuda_device):
    """compute_conditioned_worlds_prob / compute_query_prob_cond (models/vael.py:401-430) through
    VAEL_FLAG_EVIDENCE on the admissible-world circuit, against the reference golden (cond_worlds: evidence 'down',
    first golden row) and the reference's formula for every move and a batch."""
    from vael_b200.mario_task import compile_mario_family, mario_tables
    from vael_b200.networks import MarioMLP
    from vael_b200.vael import MarioVAELModel
    from vael_b200 import _lib
    g = load("mario")
    tabs = [torch.from_numpy(t.astype("float32")).cuda() for t in mario_tables(compile_mario_family((3, 3)))]
    W_all, WQ_all, W_adm, WQ_adm = tabs
    model = MarioVAELModel(None, None, MarioMLP(18, 18, 32), (18, 30), None, WQ_all, W_all, WQ_adm, W_adm, (3, 3),
                           device=cuda_device)
    facts = cu(g["facts"])
    n0 = _lib.launch_count()
    cond = model.compute_conditioned_worlds_prob(facts[:1], torch.tensor([0., 1., 0., 0.]))
    assert _lib.launch_count() > n0
    np.testing.assert_allclose(cond.cpu().numpy(), g["cond_worlds"], rtol=1e-5, atol=1e-12)
    for m in range(4):
        ev = torch.eye(4)[m]
        got = model.compute_conditioned_worlds_prob(facts, ev)
        p_w = torch.exp(torch.log(facts.double()) @ W_adm.double().T)
        w_q = torch.all(WQ_adm.double() == ev.double().cuda(), dim=1)
        ref = (p_w * w_q + 1e-5) / (torch.sum(p_w * w_q) + 1e-5)
        np.testing.assert_allclose(got.cpu().numpy(), ref.cpu().numpy(), rtol=1e-5, atol=1e-12)
        pq = model.compute_query_prob_cond(m, got[0])
        ref_q = torch.sum(WQ_adm[:, m].double() * ref[0]) / (ref[0].sum() + 1e-5)
        assert abs(pq.item() - ref_q.item()) <= 1e-5 * abs(ref_q.item())
    ones = model.compute_conditioned_worlds_prob(facts[:2], torch.tensor([1., 1., 0., 0.]))   # matches no world
    assert torch.equal(ones, torch.ones(2, 24).cuda())


def test_epoch_without_a_host_synchronisation(cuda_device):
    """on-device loader + CUDA-graphed step.  A whole epoch — batches gathered and
    standardised in HBM (vael_b200/data.py), forward / ELBO / adjoint / Adam replayed from the graph — under
    torch.cuda.set_sync_debug_mode('error'): not one implicit device synchronisation; the reference's loop does two
    H2D copies and five `.cpu()` reads per step (utils/mnist_utils/train.py:115-144)."""
    from vael_b200.data import SyntheticPairsLoader
    from vael_b200.mnist_task import build_model_dict, build_worlds_queries_matrix
    from vael_b200.networks import MNISTPairsDecoder, MNISTPairsEncoder, MNISTPairsMLP
    from vael_b200.train import GraphedTrainStep
    from vael_b200.vael import MNISTPairsVAELModel
    torch.manual_seed(0)
    model = MNISTPairsVAELModel(MNISTPairsEncoder(hidden_channels=64, latent_dim=23),
                                MNISTPairsDecoder(label_dim=20, hidden_channels=64, latent_dim=8),
                                MNISTPairsMLP(in_features=15, n_facts=20), (15, 8), build_model_dict(2, 10, device=cuda_device),
                                build_worlds_queries_matrix(2, 10).to(cuda_device), dropout=0.5, is_train=True,
                                device=cuda_device, query_per_row=True).to(cuda_device)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, capturable=True)
    before = [p.detach().clone() for p in model.parameters()]
    loader = SyntheticPairsLoader(samples_per_world=32, batch_size=32, device=cuda_device, seed=3)
    step = GraphedTrainStep(model, opt, (32, 1, 28, 56), (32, 4), device=cuda_device)
    assert all(torch.equal(a, b) for a, b in zip(before, model.parameters())), "building the step changed the weights"
    assert all(float(v["step"]) == 0.0 for v in opt.state.values())
    losses = []
    torch.cuda.synchronize()
    torch.cuda.set_sync_debug_mode("error")
    try:
        for images, labels in loader:
            terms = step(images[:, None], labels)
            losses.append(terms[0].clone())
    finally:
        torch.cuda.set_sync_debug_mode("default")
    losses = torch.stack(losses).cpu()
    assert len(losses) == 100 and torch.isfinite(losses).all() and losses[-10:].mean() < losses[:10].mean()
