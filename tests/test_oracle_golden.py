"""The CPU oracle pinned against golden vectors written by the reference itself
(tests/golden/make_golden.py imports /root/reference; the .npz files travel)."""
import os

import numpy as np
import pytest
import torch

from oracle import dense_path, fp_interp, graph_path
from oracle.semiring_port import ElementwiseSemiring, SemiringPort

G = os.path.join(os.path.dirname(__file__), "golden")


def load(name):
    return np.load(os.path.join(G, name + ".npz"))


def T(a):
    return torch.from_numpy(np.asarray(a))


def test_semiring_port_matches_reference_ops_bit_exact():
    g = load("semiring_ops")
    sr = SemiringPort(batch_size=8)
    a = T(g["a"])
    cases = dict(pp=(a, T(g["b_plain"])), pz=(a, T(g["b_zero"])), zp=(T(g["a_zero"]), T(g["b_plain"])),
                 po=(a, T(g["b_one"])), op=(T(g["a_one"]), T(g["b_plain"])))
    for tag, (x, y) in cases.items():
        np.testing.assert_array_equal(sr.plus(x, y).numpy(), g["plus_" + tag])
        np.testing.assert_array_equal(sr.times(x, y).numpy(), g["times_" + tag])
    np.testing.assert_array_equal(sr.negate(a).numpy(), g["negate_a"])
    np.testing.assert_array_equal(sr.normalize(a, T(g["b_plain"])).numpy(), g["normalize"])
    np.testing.assert_array_equal(sr.one().numpy(), g["one"])
    np.testing.assert_array_equal(sr.zero().numpy(), g["zero"])
    assert [sr.is_zero(t) for t in (a, T(g["b_zero"]), T(g["a_zero"]))] == list(g["is_zero"])
    assert [sr.is_one(t) for t in (a, T(g["b_one"]), T(g["a_one"]))] == list(g["is_one"])
    sr.set_weights({"k": a})
    np.testing.assert_array_equal(sr.value("k").numpy(), g["value_hit"])
    assert sr.value(3.0) == float(g["value_miss"])


def test_reference_quirk_cross_row_corruption_is_reproduced():
    """utils/graph_semiring.py:27-49: one ~0 row in b makes plus() drop b for EVERY row.
    The port keeps that; the elementwise semiring (what the kernels do) does not."""
    g = load("semiring_ops")
    a, b = T(g["a"]), T(g["b_zero"])
    np.testing.assert_array_equal(SemiringPort(8).plus(a, b).numpy(), g["a"])           # b dropped everywhere
    assert not np.array_equal(ElementwiseSemiring(8).plus(a, b).numpy(), g["a"])


def test_dense_path_matches_reference():
    g = load("mnist_dense")
    logits = T(g["logits"]).requires_grad_(True)
    facts = dense_path.facts_probability(logits)
    np.testing.assert_array_equal(facts.detach().numpy(), g["facts"])
    np.testing.assert_array_equal(dense_path.worlds_queries_matrix().numpy(), g["w_q"])
    np.testing.assert_array_equal(dense_path.herbrand_base(20).numpy(), g["herbrand"])
    w_q, gw, gq = T(g["w_q"]), T(g["gw"]), T(g["gq"])
    for q in (0, 9, 18):
        qp, wp = dense_path.dense_inference(facts, w_q, q)
        np.testing.assert_array_equal(wp.detach().numpy(), g[f"worlds_q{q}"])
        np.testing.assert_array_equal(qp.detach().numpy(), g[f"query_q{q}"])
        gf, gl = torch.autograd.grad((wp * gw).sum() + (qp * gq).sum(), [facts, logits], retain_graph=True)
        np.testing.assert_allclose(gf.numpy(), g[f"grad_facts_q{q}"], rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(gl.numpy(), g[f"grad_logits_q{q}"], rtol=1e-5, atol=1e-7)
        qr, _ = dense_path.dense_inference_rows(facts, w_q, torch.full((32,), q))
        np.testing.assert_allclose(qr.detach().numpy(), g[f"query_q{q}"], rtol=1e-6)


def test_gumbel_hard_matches_reference():
    g = load("gumbel")
    worlds = T(g["worlds"]).requires_grad_(True)
    world, y_soft, idx = dense_path.gumbel_hard(torch.log(worlds), T(g["noise"]))
    np.testing.assert_array_equal(world.detach().numpy(), g["world"])
    wh = world @ dense_path.herbrand_base(20)
    np.testing.assert_array_equal(wh.detach().numpy(), g["world_h"])
    (gwd,) = torch.autograd.grad((wh * T(g["gh"])).sum(), [worlds])
    np.testing.assert_allclose(gwd.numpy(), g["grad_worlds"], rtol=1e-5, atol=1e-6)


def test_elbo_terms_match_reference():
    g = load("elbo")
    leaves = [T(g[k]).requires_grad_(True) for k in ("recon", "mu", "logvar", "qprob")]
    rw, kw, qw = (float(v) for v in g["weights"])
    loss, rl, kl, qb = dense_path.elbo_terms(leaves[0], T(g["x"]), leaves[1], leaves[2], leaves[3], rw, kw, qw)
    np.testing.assert_allclose([loss.item(), rl.item(), kl.item(), qb.item()], g["terms"], rtol=1e-6)
    grads = torch.autograd.grad(loss, leaves)
    for got, key in zip(grads, ("g_recon", "g_mu", "g_logvar", "g_q")):
        np.testing.assert_allclose(got.numpy(), g[key], rtol=1e-6, atol=1e-9)


def test_mario_path_matches_reference():
    g = load("mario")
    tabs = {n: T(g[n].astype(np.float32)) for n in ("W_all", "WQ_all", "W_adm", "WQ_adm")}
    facts = T(g["facts"]).requires_grad_(True)
    worlds = dense_path.mario_worlds(facts, tabs["W_all"])
    query = dense_path.mario_query_prob(T(g["moves"]), worlds, tabs["WQ_all"])
    adm = dense_path.mario_admissible_logprob(facts, tabs["W_adm"])
    np.testing.assert_array_equal(worlds.detach().numpy(), g["worlds"])
    np.testing.assert_allclose(query.detach().numpy(), g["query"], rtol=1e-6)
    np.testing.assert_array_equal(adm.detach().numpy(), g["adm_logprob"])
    (gf,) = torch.autograd.grad((worlds * T(g["gw"])).sum() + (query * T(g["gq"])).sum() + (adm * T(g["ga"])).sum(), [facts])
    np.testing.assert_allclose(gf.numpy(), g["grad_facts"], rtol=1e-5, atol=1e-5)
    cond = dense_path.mario_conditioned_worlds(facts[:1].detach(), tabs["W_adm"], tabs["WQ_adm"], [0., 1., 0., 0.])
    np.testing.assert_allclose(cond.numpy(), g["cond_worlds"], rtol=1e-6)
    np.testing.assert_array_equal(dense_path.herbrand_base(18).numpy(), g["herbrand"])


def test_graph_path_with_port_equals_reference_graphsemiring(addition_family):
    """The CSR interpreter driven by the SemiringPort reproduces, bit for bit, what the same
    interpreter produced when driven by the reference's own GraphSemiring + extract_*."""
    g = load("graph_path")
    circ = addition_family.circuit
    assert circ.sha256() == str(g["circuit_sha256"])
    facts = T(g["facts"])
    sr = SemiringPort(batch_size=facts.shape[0])
    sr.set_weights(graph_path.bind_weights(facts, circ.fact_labels))
    res = graph_path.evaluate(circ, sr, "query", 9)
    np.testing.assert_array_equal(torch.stack([res[k] for k in circ.world_atoms], 1).numpy(), g["query_res_worlds"])
    np.testing.assert_array_equal(res[next(reversed(res))].numpy(), g["query_res_last"])
    np.testing.assert_array_equal(graph_path.extract_query_probability(res).numpy(), g["extract_query"])
    np.testing.assert_array_equal(graph_path.extract_worlds_probability(res, circ.world_atoms).numpy(), g["extract_worlds"])
    res_e = graph_path.evaluate(circ, sr, "evidence", 7)
    np.testing.assert_array_equal(torch.stack([res_e[k] for k in circ.world_atoms], 1).numpy(), g["evidence_res_worlds"])
    np.testing.assert_array_equal(graph_path.extract_worlds_probability(res_e, circ.world_atoms).numpy(),
                                  g["evidence_extract_worlds"])


def test_graph_path_agrees_with_dense_closed_form_in_domain(addition_family):
    """(i) reference dense form == (ii) reference-semiring graph path == (iii) elementwise fp32/fp64
    interpreters, within the north-star tolerance, on in-domain facts."""
    g, gd = load("graph_path"), load("mnist_dense")
    circ = addition_family.circuit
    facts = g["facts"]
    w32, q32 = fp_interp.forward(circ, facts, query=9, normalize=True)
    w64, q64 = fp_interp.forward(circ, facts, query=9, normalize=True, dtype=np.float64)
    for w, q in ((g["query_res_worlds"], g["query_res_last"]), (w32, q32), (gd["worlds_q9"], gd["query_q9"][:, 0])):
        np.testing.assert_allclose(w, w64, rtol=1e-5, atol=1e-12)
        np.testing.assert_allclose(q, q64, rtol=1e-5)
    # raw (un-normalised) fp32 interpreter worlds are the dense worlds bit for bit
    w_raw, _ = fp_interp.forward(circ, facts, query=9)
    np.testing.assert_array_equal(w_raw, gd["worlds_q9"])
    # evidence mode: P(w|e) closed form
    we, _ = fp_interp.forward(circ, facts, query=7, evidence=True)
    ref = dense_path.evidence_worlds(T(facts).reshape(-1, 2, 10), T(gd["w_q"]), 7).numpy()
    np.testing.assert_allclose(we, ref, rtol=1e-5, atol=1e-12)
    np.testing.assert_allclose(g["evidence_res_worlds"], ref, rtol=1e-5, atol=1e-12)


def test_fp64_adjoint_matches_autograd_of_dense_form(addition_family):
    gd = load("mnist_dense")
    circ = addition_family.circuit
    facts = gd["facts"].reshape(32, 20)
    for q in (0, 9, 18):
        gref = gd[f"grad_facts_q{q}"].reshape(32, 20)
        got = fp_interp.backward(circ, facts, gd["gw"], gd["gq"][:, 0], q)
        np.testing.assert_allclose(got, gref, rtol=1e-4, atol=1e-5)
