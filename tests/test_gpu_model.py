"""End-to-end parity of the model mirror against the reference-equivalent composition of
the torch oracle, with injected noise (SURVEY.md hard parts 3 and 7); the compiled-program
`evaluate` path against the graph-path golden; the Mario path against its golden."""
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


def build_mnist_model(device, query_per_row=False, dropout=None):
    from vael_b200.mnist_task import build_model_dict, build_worlds_queries_matrix
    from vael_b200.networks import MNISTPairsDecoder, MNISTPairsEncoder, MNISTPairsMLP
    from vael_b200.vael import MNISTPairsVAELModel
    torch.manual_seed(0)
    enc = MNISTPairsEncoder(hidden_channels=64, latent_dim=23, dropout=0.0)
    dec = MNISTPairsDecoder(label_dim=20, hidden_channels=64, latent_dim=8, dropout=0.0)
    mlp = MNISTPairsMLP(in_features=15, n_facts=20)
    model = MNISTPairsVAELModel(enc, dec, mlp, (15, 8), build_model_dict(2, 10), build_worlds_queries_matrix(2, 10).to(device),
                                dropout=dropout, is_train=True, device=device, query_per_row=query_per_row)
    return model.to(device)


def reference_forward(model, x, labels, eps, gumbel, per_row):
    """models/vael.py:39-70 composed from the oracle's dense restatement, on the CPU."""
    enc, dec, mlp = model.encoder, model.decoder, model.mlp
    mu, logvar = enc(x)
    z = mu + torch.exp(logvar / 2) * eps
    z_sub = z[:, 15:]
    logits = torch.cat(mlp(z[:, :15]), 1)
    facts = dense_path.facts_probability(logits)
    w_q = dense_path.worlds_queries_matrix()
    if per_row:
        qp, worlds = dense_path.dense_inference_rows(facts, w_q, labels[:, -2])
    else:
        qp, worlds = dense_path.dense_inference(facts, w_q, int(labels[0][-2]))
    world, _, _ = dense_path.gumbel_hard(torch.log(worlds), gumbel)
    world_h = world @ dense_path.herbrand_base(20)
    image = dec(torch.cat([z_sub, world_h], 1))
    return image, mu, logvar, qp, facts, worlds


@pytest.mark.parametrize("per_row", [False, True])
def test_mnist_model_forward_backward_matches_reference_composition(cuda_device, per_row):
    from vael_b200.train import loss_function
    # the CNNs run through cuDNN/cuBLAS here and through the CPU kernels in the reference arm:
    # keep them in true fp32 so that the comparison is about the logic path
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    B = 32
    model = build_mnist_model(cuda_device, query_per_row=per_row)
    model.eval()  # dropout off (its RNG stream cannot match); is_train=True still samples z
    gen = torch.Generator().manual_seed(7)
    x = torch.randn(B, 1, 28, 56, generator=gen)
    d = torch.randint(0, 10, (B, 2), generator=gen)
    if not per_row:
        d[:] = d[0]
    labels = torch.cat([d, d.sum(1, keepdim=True), -torch.ones(B, 1, dtype=torch.long)], 1)
    eps = torch.randn(B, 23, generator=gen)
    gumbel = -torch.empty(B, 100).exponential_(generator=gen).log()

    model.reparam_noise, model.gumbel_noise = eps.cuda(), gumbel.cuda()
    image, mu, logvar, qp = model(x.cuda(), labels.cuda())
    loss, rl, kl, qce, _ = loss_function(image, x.cuda(), mu, logvar, qp, labels=labels.cuda(), query=True,
                                         recon_w=1e-1, kl_w=1e-5, query_w=1.0, sup_w=0.0, rec_loss="LAPLACE")
    model.zero_grad()
    loss.backward()
    got = {n: p.grad.detach().cpu().clone() for n, p in model.named_parameters()}
    got_out = [t.detach().cpu() for t in (image, mu, logvar, qp, model.facts_probs, model.worlds_prob)]
    world_idx = model.world_idx.cpu()
    assert torch.equal(model.world.argmax(1).cpu(), world_idx.long())

    ref_model = build_mnist_model(torch.device("cpu"), query_per_row=per_row)
    ref_model.load_state_dict({k: v.cpu() for k, v in model.state_dict().items()})
    ref_model.eval()
    ref_model.zero_grad()
    r_image, r_mu, r_logvar, r_qp, r_facts, r_worlds = reference_forward(ref_model, x, labels, eps, gumbel, per_row)
    r_loss, *_ = dense_path.elbo_terms(r_image, x, r_mu, r_logvar, r_qp, 1e-1, 1e-5, 1.0)
    r_loss.backward()

    for a, b, name in zip(got_out, (r_image, r_mu, r_logvar, r_qp, r_facts, r_worlds),
                          ("image", "mu", "logvar", "query", "facts", "worlds")):
        b = b.detach().numpy()
        err = rel_err(a.numpy(), b, floor=1e-3 * float(np.abs(b).max()) + 1e-9)
        assert err < 2e-4, (name, err)   # conv/linear: cuDNN vs CPU fp32
    assert abs(loss.item() - r_loss.item()) < 1e-4 * abs(r_loss.item())
    for n, p in ref_model.named_parameters():
        ref = p.grad
        assert rel_err(got[n].numpy(), ref.numpy(), floor=1e-3 * ref.abs().max().item() + 1e-9) < 2e-3, n


def test_compiled_program_evaluate_matches_graph_path_golden(cuda_device):
    """model_dict[mode][label].evaluate(semiring) + extract_* against the golden produced by the
    reference's GraphSemiring / extract_* (query mode q=9, evidence mode e=7)."""
    from vael_b200.logic import Constant, Term
    g = load("graph_path")
    model = build_mnist_model(cuda_device)
    facts = cu(g["facts"])
    model.update_semiring_weights(facts)
    res = model.model_dict["query"][9].evaluate(semiring=model.semiring)
    keys = [Term("digits")(Constant(j), Constant(k)) for j in range(10) for k in range(10)]
    assert list(res.keys())[:100] == keys and str(list(res.keys())[-1]) == "addition(img,9)"
    stacked = torch.stack([res[k] for k in keys], 1).cpu().numpy()
    np.testing.assert_allclose(stacked, g["query_res_worlds"], rtol=1e-5, atol=1e-12)
    np.testing.assert_allclose(res[list(res.keys())[-1]].cpu().numpy(), g["query_res_last"], rtol=1e-5)
    np.testing.assert_allclose(model.extract_query_probability(res).cpu().numpy(), g["extract_query"], rtol=1e-5)
    np.testing.assert_allclose(model.extract_worlds_probability(res).cpu().numpy(), g["extract_worlds"], rtol=1e-5, atol=1e-12)
    wp = model.problog_inference_with_evidence(facts, 7)
    np.testing.assert_allclose(wp.cpu().numpy(), g["evidence_extract_worlds"], rtol=1e-5, atol=1e-12)
    # the generic (base-class) query path, made real
    from vael_b200.vael import VAELModel
    qp, worlds = VAELModel.problog_inference(model, facts, labels=None, query=9)
    np.testing.assert_allclose(qp.cpu().numpy(), g["extract_query"], rtol=1e-5)
    np.testing.assert_allclose(worlds.cpu().numpy(), g["extract_worlds"], rtol=1e-5, atol=1e-12)


def test_all_queries_matches_dense_mask_matmul(cuda_device):
    g = load("mnist_dense")
    model = build_mnist_model(cuda_device)
    facts = cu(g["facts"])
    allq = model.all_queries(facts).cpu().numpy()
    ref = g["worlds_q9"] @ g["w_q"][:, :19]
    np.testing.assert_allclose(allq, ref, rtol=1e-5)
    for q in (0, 9, 18):
        np.testing.assert_allclose(allq[:, q], g[f"query_q{q}"][:, 0], rtol=1e-5)


def test_mario_logic_path_matches_reference_golden(cuda_device):
    from vael_b200.networks import MarioMLP
    from vael_b200.vael import MarioVAELModel
    g = load("mario")
    tabs = {n: cu(g[n].astype(np.float32)) for n in ("W_all", "WQ_all", "W_adm", "WQ_adm")}
    m = MarioVAELModel(None, None, MarioMLP(18, 18, 32), (18, 30), None, tabs["WQ_all"], tabs["W_all"], tabs["WQ_adm"],
                       tabs["W_adm"], (3, 3), device=cuda_device)
    facts = cu(g["facts"]).requires_grad_(True)
    worlds = m.compute_worlds_prob_conj(facts)
    query = m.compute_query_prob(cu(g["moves"]), worlds)
    adm = m.admissible_worlds_logprob(facts)
    assert rel_err(worlds.detach().cpu().numpy(), g["worlds"]) < 1e-5
    assert rel_err(query.detach().cpu().numpy(), g["query"]) < 1e-5
    assert rel_err(adm.detach().cpu().numpy(), g["adm_logprob"]) < 1e-5
    (gf,) = torch.autograd.grad((worlds * cu(g["gw"])).sum() + (query * cu(g["gq"])).sum() + (adm * cu(g["ga"])).sum(), [facts])
    ref = g["grad_facts"]
    assert rel_err(gf.cpu().numpy(), ref, floor=1e-3 * np.abs(ref).max()) < 1e-4
    cond = m.compute_conditioned_worlds_prob(facts[:1].detach(), torch.tensor([0., 1., 0., 0.]))
    np.testing.assert_allclose(cond.cpu().numpy(), g["cond_worlds"], rtol=1e-5)


def test_cuda_graphed_train_step_runs_and_learns(cuda_device):
    """The whole step (forward, ELBO, adjoint kernels, Adam) replayed from one CUDA graph: the loss is
    finite and goes down on a fixed batch, parameters move, fresh randomness is drawn on every replay."""
    from vael_b200.train import GraphedTrainStep
    B = 32
    model = build_mnist_model(cuda_device, query_per_row=True, dropout=0.5)
    model.train()
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, capturable=True)
    gen = torch.Generator().manual_seed(3)
    x = torch.rand(B, 1, 28, 56, generator=gen).cuda()
    d = torch.randint(0, 10, (B, 2), generator=gen)
    labels = torch.cat([d, d.sum(1, keepdim=True), -torch.ones(B, 1, dtype=torch.long)], 1).cuda()
    step = GraphedTrainStep(model, opt, tuple(x.shape), tuple(labels.shape), device=cuda_device)
    before = [p.detach().clone() for p in model.parameters()]
    losses, worlds = [], []
    for _ in range(60):
        losses.append(step(x, labels)[0].item())
        worlds.append(model.world_idx.clone())
    assert all(np.isfinite(losses))
    assert np.mean(losses[-10:]) < np.mean(losses[:10])
    assert any(not torch.equal(a, b) for a, b in zip(before, model.parameters()))
    assert any(not torch.equal(worlds[0], w) for w in worlds[1:5]), "the graph replays the same random numbers"


def build_mario_model(device, g):
    from vael_b200.networks import MarioDecoder, MarioEncoder, MarioMLP
    from vael_b200.vael import MarioVAELModel
    torch.manual_seed(0)
    tabs = {n: torch.from_numpy(g[n].astype(np.float32)).to(device) for n in ("W_all", "WQ_all", "W_adm", "WQ_adm")}
    # shapes of the reference configuration (config.py:47-51): hidden 64/64/32, latent 18 symbolic + 30 sub-symbolic
    enc = MarioEncoder(hidden_channels=64, latent_dim=48, dropout=0.0)
    dec = MarioDecoder(hidden_channels=64, latent_dim_sub=30, label_dim=9, dropout=0.0)
    model = MarioVAELModel(enc, dec, MarioMLP(18, 18, 32), (18, 30), None, tabs["WQ_all"], tabs["W_all"], tabs["WQ_adm"],
                           tabs["W_adm"], (3, 3), dropout=None, is_train=True, device=device)
    return model.to(device), tabs


def test_mario_model_forward_backward_matches_reference_composition(cuda_device):
    """MarioVAELModel.forward + loss + backward (models/vael.py:295-346, utils/mario_utils/train.py:11-65)
    against the same composition built from the oracle's restatement of the matrix path, injected noise."""
    from vael_b200.train import mario_loss_function
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    g = load("mario")
    B = 8
    gen = torch.Generator().manual_seed(11)
    x = torch.rand(B, 3, 200, 100, generator=gen)
    moves = torch.eye(4, dtype=torch.long)[torch.randint(0, 4, (B,), generator=gen)]
    eps = torch.randn(B, 48, generator=gen)
    gumbel = -torch.empty(B, 24).exponential_(generator=gen).log()

    model, _ = build_mario_model(cuda_device, g)
    model.eval()
    model.reparam_noise, model.gumbel_noise = eps.cuda(), gumbel.cuda()
    outs = model(x.cuda(), moves.cuda())
    res = mario_loss_function(*outs, recon_w=1e-1, kl_w=1e-5, query_w=1.0, sup_w=0.0, rec_loss="LAPLACE")
    model.zero_grad()
    res["elbo"].backward()
    got = {n: p.grad.detach().cpu().clone() for n, p in model.named_parameters()}

    ref, tabs = build_mario_model(torch.device("cpu"), g)
    ref.load_state_dict({k: v.cpu() for k, v in model.state_dict().items()})
    ref.eval()
    x1, x2 = torch.split(x, 100, dim=2)
    (mu1, lv1), (mu2, lv2) = ref.encoder(x1), ref.encoder(x2)
    z1, z2 = mu1 + torch.exp(lv1 / 2) * eps, mu2 + torch.exp(lv2 / 2) * eps
    facts = dense_path.mario_facts(ref.mlp(z1[:, :18]), ref.mlp(z2[:, :18]))
    worlds = dense_path.mario_worlds(facts, tabs["W_all"])
    query = dense_path.mario_query_prob(moves, worlds, tabs["WQ_all"])
    adm = dense_path.mario_admissible_logprob(facts, tabs["W_adm"])
    world, _, _ = dense_path.gumbel_hard(adm, gumbel)
    wh = world @ tabs["W_adm"]
    r1, r2 = ref.decoder(torch.cat([z1[:, 18:], wh[:, :9]], 1)), ref.decoder(torch.cat([z2[:, 18:], wh[:, 9:]], 1))
    l1, *_ = dense_path.elbo_terms(r1, x1, mu1, lv1, query, 1e-1, 1e-5, 1.0)
    # frame 2 carries reconstruction + KL only (utils/mario_utils/train.py:49-52)
    l2 = 1e-1 * (-dense_path.laplace_loglik(r2, x2).mean()) + \
        1e-5 * torch.mean(-0.5 * torch.sum(1.0 + lv2 - mu2.pow(2) - lv2.exp(), dim=1))
    (l1 + l2).backward()

    assert rel_err(model.worlds_probs.detach().cpu().numpy(), worlds.detach().numpy(), floor=1e-9) < 2e-4
    assert rel_err(model.query_prob.detach().cpu().numpy(), query.detach().numpy(), floor=1e-9) < 2e-4
    assert rel_err(model.adm_worlds_logprob.detach().cpu().numpy(), adm.detach().numpy(), floor=1e-6) < 2e-4
    assert abs(res["elbo"].item() - (l1 + l2).item()) < 1e-4 * abs((l1 + l2).item())
    # gradients pass through four 5x5 (transposed) convolutions over 100x100 frames, computed by cuDNN here
    # (algorithm chosen by its heuristics, incl. transform-based ones) and by the CPU kernels in the reference
    # arm: measured agreement is ~5e-3 of a tensor's largest entry, so this is a wiring check of the whole step;
    # the logic path itself is held to 1e-4 in test_mario_logic_path_matches_reference_golden
    for n, p in ref.named_parameters():
        r = p.grad
        assert float((got[n] - r).abs().max()) < 2e-2 * float(r.abs().max()) + 1e-9, n


def test_reference_style_calls_through_the_problog_shim(cuda_device):
    """What the reference itself executes, line for line: programs compiled with
    SDD.create_from(LogicDAG.create_from(LogicFormula.create_from(text))) (mnist_task_VAEL.py:214-216),
    weights bound as update_semiring_weights does (models/vael.py:227-243) on a FOREIGN semiring object
    (the oracle's restatement of the reference class: it has no fast-path hooks), sdd.evaluate(semiring=...)
    (models/vael.py:109,130), extract_* — against the golden written by the reference's GraphSemiring."""
    from oracle import graph_path
    from oracle.semiring_port import SemiringPort
    from vael_b200 import compat
    from vael_b200.mnist_task import addition_rules, create_facts, define_ProbLog_model
    compat.install(force=True)
    from problog.formula import LogicDAG, LogicFormula
    from problog.logic import Constant, Term
    from problog.sdd_formula import SDD
    g = load("graph_path")
    facts = cu(g["facts"])                                   # [32,2,10]
    weights = {Term(f"p_{i}{j}"): facts[:, i - 1, j] for i in (1, 2) for j in range(10)}
    ads = create_facts(2, n_digits=10)
    keys = [Term("digits")(Constant(j), Constant(k)) for j in range(10) for k in range(10)]

    def run(mode, label):
        text = define_ProbLog_model(ads, addition_rules(2), label=label, digit_query="digits(X,Y)", mode=mode)
        sdd = SDD.create_from(LogicDAG.create_from(LogicFormula.create_from(text)))
        sr = SemiringPort(batch_size=32)
        sr.set_weights(weights)
        return sdd.evaluate(semiring=sr)

    res = run("query", 9)
    assert list(res.keys())[:100] == keys and str(list(res.keys())[-1]) == "addition(img,9)"
    np.testing.assert_allclose(torch.stack([res[k] for k in keys], 1).cpu().numpy(), g["query_res_worlds"], rtol=1e-5, atol=1e-12)
    np.testing.assert_allclose(graph_path.extract_query_probability(res).cpu().numpy(), g["extract_query"], rtol=1e-5)
    np.testing.assert_allclose(graph_path.extract_worlds_probability(res, keys).cpu().numpy(), g["extract_worlds"], rtol=1e-5, atol=1e-12)
    res_e = run("evidence", 7)
    np.testing.assert_allclose(torch.stack([res_e[k] for k in keys], 1).cpu().numpy(), g["evidence_res_worlds"], rtol=1e-5, atol=1e-12)


def test_cuda_graphed_mario_train_step_runs_and_learns(cuda_device):
    """The Mario step (two encoders / decoders, the 81-world circuit, the admissible-world literal-table kernel,
    Gumbel sampling over log-probabilities, two ELBO kernels, Adam) replayed from one CUDA graph."""
    from vael_b200.train import GraphedTrainStep, mario_train_step
    g = load("mario")
    model, _ = build_mario_model(cuda_device, g)
    model.train()
    opt = torch.optim.Adam(model.parameters(), lr=1e-4, capturable=True)
    B = 8
    gen = torch.Generator().manual_seed(5)
    x = torch.rand(B, 3, 200, 100, generator=gen).cuda()
    moves = torch.eye(4, dtype=torch.long)[torch.randint(0, 4, (B,), generator=gen)].cuda()
    step = GraphedTrainStep(model, opt, tuple(x.shape), tuple(moves.shape), device=cuda_device, step_fn=mario_train_step)
    before = [p.detach().clone() for p in model.parameters()]
    elbos, worlds = [], []
    for _ in range(40):
        elbos.append(step(x, moves).item())
        worlds.append(model.world_idx.clone())
    assert all(np.isfinite(elbos))
    assert np.mean(elbos[-8:]) < np.mean(elbos[:8])
    assert any(not torch.equal(a, b) for a, b in zip(before, model.parameters()))
    assert any(not torch.equal(worlds[0], w) for w in worlds[1:6]), "the graph replays the same random numbers"
