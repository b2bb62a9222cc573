"""Generate the golden fixtures under tests/golden/ by running the REFERENCE itself.

Run in the build container only (`python tests/golden/make_golden.py`): it imports
/root/reference/{models/vael.py, utils/graph_semiring.py, utils/*/train.py} unmodified.
The reference's only missing imports are `problog.logic` / `problog.evaluator`
(third-party, not installable here); a throw-away stub package standing in for them is
written to a temp dir — the term classes come from vael_b200.logic, `Semiring` is an
empty base with problog's documented defaults.  /root/reference does not exist on the GPU
box, which is why the outputs are committed as small .npz files next to this script.
"""
import ast
import os
import sys
import tempfile
import textwrap

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("VAEL_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)


def install_problog_stub():
    d = tempfile.mkdtemp(prefix="problog_stub_")
    os.makedirs(os.path.join(d, "problog"))
    open(os.path.join(d, "problog", "__init__.py"), "w").write("")
    open(os.path.join(d, "problog", "logic.py"), "w").write(
        "from vael_b200.logic import Term, Constant, Var, AnnotatedDisjunction\n")
    open(os.path.join(d, "problog", "evaluator.py"), "w").write(textwrap.dedent("""
        class Semiring(object):
            def pos_value(self, a, key=None): return self.value(a)
            def neg_value(self, a, key=None): return self.negate(self.value(a))
            def ad_negate(self, pos, neg): return self.one()
            def ad_complement(self, ws, key=None):
                s = self.zero()
                for w in ws: s = self.plus(s, w)
                return self.negate(s)
    """))
    sys.path.insert(0, d)
    sys.path.insert(0, REF)


def reference_function(relpath, name, namespace):
    """exec ONE function of a reference file whose module-level imports cannot be satisfied
    here (matplotlib, problog.sdd_formula); the source is read from /root/reference at run time."""
    tree = ast.parse(open(os.path.join(REF, relpath)).read())
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == name)
    code = compile(ast.Module(body=[fn], type_ignores=[]), relpath, "exec")
    exec(code, namespace)
    return namespace[name]


def facts_like(B, groups, seed, scale=3.0):
    g = torch.Generator().manual_seed(seed)
    return scale * torch.randn(B, sum(groups), generator=g)


def main():
    install_problog_stub()
    from itertools import product
    from problog.logic import Constant, Term
    import models.vael as ref_vael
    from models.vael_networks import MNISTPairsMLP, MarioMLP
    from utils.graph_semiring import GraphSemiring
    import utils.mnist_utils.train as ref_mnist_train
    import utils.mario_utils.train as ref_mario_train
    from oracle import graph_path
    from vael_b200.mnist_task import compile_addition_family

    torch.manual_seed(0)
    out = {}

    # ---- G1: GraphSemiring ops, incl. the batch-wide shortcuts ------------------------------
    sr = GraphSemiring(batch_size=8)
    a = torch.tensor([.2, .3, .4, .5, .6, .7, .8, .9])
    b_plain = torch.tensor([.1, .2, .3, .4, .5, .6, .7, .8])
    b_zero = b_plain.clone(); b_zero[1] = 0.0          # one ~0 row  -> plus drops b for every row
    b_one = b_plain.clone(); b_one[6] = 1.0            # one ~1 row  -> times drops b for every row
    a_zero = a.clone(); a_zero[3] = 5e-13
    a_one = a.clone(); a_one[0] = 1.0 - 5e-13
    g1 = dict(a=a, b_plain=b_plain, b_zero=b_zero, b_one=b_one, a_zero=a_zero, a_one=a_one)
    for tag, (x, y) in dict(pp=(a, b_plain), pz=(a, b_zero), zp=(a_zero, b_plain), po=(a, b_one),
                            op=(a_one, b_plain)).items():
        g1["plus_" + tag] = sr.plus(x, y)
        g1["times_" + tag] = sr.times(x, y)
    g1["negate_a"] = sr.negate(a)
    g1["normalize"] = sr.normalize(a, b_plain)
    g1["one"], g1["zero"] = sr.one(), sr.zero()
    g1["is_zero"] = torch.tensor([bool(sr.is_zero(t)) for t in (a, b_zero, a_zero)])
    g1["is_one"] = torch.tensor([bool(sr.is_one(t)) for t in (a, b_one, a_one)])
    sr.set_weights({"k": a})
    g1["value_hit"], g1["value_miss"] = sr.value("k"), torch.tensor(float(sr.value(3.0)))
    np.savez(os.path.join(HERE, "semiring_ops.npz"), **{k: v.numpy() for k, v in g1.items()})

    # ---- G2/G3: MNIST dense closed form, tables ------------------------------------------------
    bwq = reference_function("utils/mnist_utils/mnist_task_VAEL.py", "build_worlds_queries_matrix",
                             {"torch": torch, "product": product})
    w_q = bwq(2, 10)
    mlp = MNISTPairsMLP(in_features=15, n_facts=20)
    model = ref_vael.MNISTPairsVAELModel(encoder=None, decoder=None, mlp=mlp, latent_dims=(15, 8), model_dict=None,
                                         w_q=w_q, dropout=None, is_train=True, device="cpu")
    B = 32
    logits = facts_like(B, (10, 10), seed=1).requires_grad_(True)

    class FixedMLP(torch.nn.Module):  # compute_facts_probability calls self.mlp(z) -> (z1, z2)
        n_facts = 20
        def forward(self, z):
            return torch.split(z, 10, dim=1)
    model.mlp = FixedMLP()
    facts = model.compute_facts_probability(logits)
    g = torch.Generator().manual_seed(2)
    gw, gq = torch.randn(B, 100, generator=g), torch.randn(B, 1, generator=g)
    g2 = dict(logits=logits.detach(), facts=facts.detach(), w_q=w_q, herbrand=model.herbrand_base, gw=gw, gq=gq)
    for q in (0, 9, 18):
        labels = torch.tensor([[0, 0, q, -1]] * B)
        qp, wp = model.problog_inference(facts, labels=labels)
        loss = (wp * gw).sum() + (qp * gq).sum()
        gf, gl = torch.autograd.grad(loss, [facts, logits], retain_graph=True)
        g2[f"worlds_q{q}"], g2[f"query_q{q}"] = wp.detach(), qp.detach()
        g2[f"grad_facts_q{q}"], g2[f"grad_logits_q{q}"] = gf, gl
    np.savez(os.path.join(HERE, "mnist_dense.npz"), **{k: v.numpy() for k, v in g2.items()})

    # ---- G4: gumbel-softmax(hard) + herbrand, noise captured --------------------------------------
    worlds = g2["worlds_q9"].clone().requires_grad_(True)
    torch.manual_seed(123)
    noise = -torch.empty(B, 100).exponential_().log()     # what F.gumbel_softmax draws first
    torch.manual_seed(123)
    world = torch.nn.functional.gumbel_softmax(torch.log(worlds), tau=1, hard=True)
    world_h = model.herbrand(world)
    gh = torch.randn(B, 20, generator=g)
    (gworlds,) = torch.autograd.grad((world_h * gh).sum(), [worlds])
    np.savez(os.path.join(HERE, "gumbel.npz"), worlds=worlds.detach().numpy(), noise=noise.numpy(),
             world=world.detach().numpy(), world_h=world_h.detach().numpy(), gh=gh.numpy(),
             grad_worlds=gworlds.numpy())

    # ---- G5: ELBO terms (LAPLACE) ----------------------------------------------------------------
    x = torch.randn(B, 1, 28, 56, generator=g)
    recon = torch.sigmoid(torch.randn(B, 1, 28, 56, generator=g)).requires_grad_(True)
    mu = torch.randn(B, 23, generator=g).requires_grad_(True)
    logvar = (0.3 * torch.randn(B, 23, generator=g)).requires_grad_(True)
    qprob = g2["query_q9"].clone().requires_grad_(True)
    loss, rl, kl, qb, _ = ref_mnist_train.loss_function(recon, x, mu, logvar, qprob, model=None, query=True,
                                                        recon_w=1e-1, kl_w=1e-5, query_w=1.0, sup_w=0.0, sup=False,
                                                        rec_loss="LAPLACE")
    grads = torch.autograd.grad(loss, [recon, mu, logvar, qprob])
    np.savez(os.path.join(HERE, "elbo.npz"), x=x.numpy(), recon=recon.detach().numpy(), mu=mu.detach().numpy(),
             logvar=logvar.detach().numpy(), qprob=qprob.detach().numpy(),
             terms=np.array([loss.item(), rl.item(), kl.item(), qb.item()], dtype=np.float32),
             weights=np.array([1e-1, 1e-5, 1.0], dtype=np.float32),
             g_recon=grads[0].numpy(), g_mu=grads[1].numpy(), g_logvar=grads[2].numpy(), g_q=grads[3].numpy())

    # ---- G6: Mario matrix path with the reference's own tables -------------------------------------
    tdir = os.path.join(REF, "utils/mario_utils/problog_matrices/3x3")
    tabs = {n: torch.as_tensor(torch.load(os.path.join(tdir, n + ".pt"), weights_only=False)).float()
            for n in ("W_all", "WQ_all", "W_adm", "WQ_adm")}
    mm = ref_vael.MarioVAELModel(encoder=None, decoder=None, mlp=MarioMLP(in_features=18, n_facts=18),
                                 latent_dims=(18, 30), model_dict=None, WQ_all=tabs["WQ_all"], W_all=tabs["W_all"],
                                 WQ_adm=tabs["WQ_adm"], W_adm=tabs["W_adm"], grid_dims=(3, 3), device="cpu")
    ml = facts_like(B, (9, 9), seed=7)
    f1 = torch.clip(torch.softmax(ml[:, :9], 1), min=1e-5, max=0.9999)
    f2 = torch.clip(torch.softmax(ml[:, 9:], 1), min=1e-5, max=0.9999)
    mfacts = torch.cat([f1, f2], 1).requires_grad_(True)
    moves = torch.nn.functional.one_hot(torch.randint(0, 4, (B,), generator=g), 4)
    mworlds = mm.compute_worlds_prob_conj(mfacts)
    mquery = mm.compute_query_prob(moves, mworlds)
    madm = mm.admissible_worlds_logprob(mfacts)
    gmw, gmq, gma = torch.randn(B, 81, generator=g), torch.randn(B, generator=g), torch.randn(B, 24, generator=g)
    (gmf,) = torch.autograd.grad((mworlds * gmw).sum() + (mquery * gmq).sum() + (madm * gma).sum(), [mfacts])
    cond = mm.compute_conditioned_worlds_prob(mfacts[:1].detach(), torch.tensor([0., 1., 0., 0.]))
    np.savez(os.path.join(HERE, "mario.npz"), facts=mfacts.detach().numpy(), moves=moves.numpy(),
             worlds=mworlds.detach().numpy(), query=mquery.detach().numpy(), adm_logprob=madm.detach().numpy(),
             gw=gmw.numpy(), gq=gmq.numpy(), ga=gma.numpy(), grad_facts=gmf.numpy(), cond_worlds=cond.numpy(),
             herbrand=mm.herbrand_base.numpy(),
             **{n: t.numpy().astype(np.int8) for n, t in tabs.items()})

    # ---- G7: the reference's GraphSemiring + extract_* driven over this repo's compiled circuit ----
    fam = compile_addition_family(2, 10)
    circ = fam.circuit
    model.mlp = mlp  # n_facts attribute back for extract_worlds_probability
    facts_d = facts.detach()
    model.update_semiring_weights(facts_d)               # reference leaf binding, Term('p_ij') keys
    by_label = {str(k): v for k, v in model.semiring.weights.items()}
    model.semiring.set_weights(by_label)                 # same tensors, keyed by label string
    keys = [Term("digits")(Constant(j), Constant(k)) for j in range(10) for k in range(10)]
    g7 = dict(facts=facts_d.numpy())
    res = graph_path.evaluate(circ, model.semiring, "query", 9, world_keys=keys, query_key=Term("addition")(Term("img"), Constant(9)))
    g7["query_res_worlds"] = torch.stack([res[k] for k in keys], 1).numpy()
    g7["query_res_last"] = res[next(reversed(res))].numpy()
    g7["extract_query"] = model.extract_query_probability(res).numpy()
    g7["extract_worlds"] = model.extract_worlds_probability(res).numpy()
    res_e = graph_path.evaluate(circ, model.semiring, "evidence", 7, world_keys=keys)
    g7["evidence_res_worlds"] = torch.stack([res_e[k] for k in keys], 1).numpy()
    g7["evidence_extract_worlds"] = model.extract_worlds_probability(res_e).numpy()
    g7["circuit_sha256"] = np.array(circ.sha256())
    np.savez(os.path.join(HERE, "graph_path.npz"), **g7)

    # ---- compiled-circuit goldens (this repo's compiler; bit-exact arrays) ---------------------------
    circ.save(os.path.join(HERE, "circuit_addition_2digit.npz"))
    print("wrote:", sorted(f for f in os.listdir(HERE) if f.endswith(".npz")))


if __name__ == "__main__":
    main()
