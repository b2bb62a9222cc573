"""Dense closed forms of the path, restated from the reference (torch, any device — the
oracle and the CPU baseline run it on the host).  Gradients come from autograd, as in the
reference (loss.backward(), utils/mnist_utils/train.py:147).
"""
import math

import torch
import torch.nn.functional as F


# ---- MNIST addition -----------------------------------------------------------------
def facts_probability(logits, n_groups=2, eps=1e-5):
    """models/vael.py:159-177 — per digit: softmax, + eps, divide by the row sum computed
    under no_grad; stacked to [B, n_groups, n]."""
    groups = []
    for g in torch.chunk(logits, n_groups, dim=1):
        p = torch.softmax(g, dim=1) + eps
        with torch.no_grad():
            z = p.sum(dim=-1, keepdim=True)
        groups.append(p / z)
    return torch.stack(groups, dim=1)


def worlds_queries_matrix(n_digits=10, sequence_len=2):
    """utils/mnist_utils/mnist_task_VAEL.py:222-234 — w_q[w, q] = [digit sum of world w == q],
    worlds in itertools.product order, n_digits*sequence_len columns."""
    idx = torch.cartesian_prod(*[torch.arange(n_digits)] * sequence_len).reshape(-1, sequence_len)
    sums = idx.sum(dim=1)
    w_q = torch.zeros(idx.shape[0], n_digits * sequence_len)
    w_q[torch.arange(idx.shape[0]), sums] = 1.0
    return w_q


def herbrand_base(n_facts=20):
    """models/vael.py:179-187 — row i*n+k = [onehot_n(i), onehot_n(k)]."""
    n = n_facts // 2
    eye = torch.eye(n)
    return torch.cat([eye.repeat_interleave(n, dim=0), eye.repeat(n, 1)], dim=1)


def dense_inference(facts_probs, w_q, query):
    """models/vael.py:200-225 — worlds = outer product of the two digit distributions
    flattened to d1*n+d2; query prob = sum_w w_q[w, query] * worlds[:, w]."""
    p1, p2 = facts_probs[:, 0], facts_probs[:, 1]
    worlds = (p1[:, :, None] * p2[:, None, :]).reshape(p1.shape[0], -1)
    q = torch.sum(w_q[:, query] * worlds, dim=1, keepdim=True)
    return q, worlds


def dense_inference_rows(facts_probs, w_q, query_rows):
    """Same with one query per row (mixed-label synthetic batches; the reference takes
    labels[0] for the whole batch, models/vael.py:212-213)."""
    p1, p2 = facts_probs[:, 0], facts_probs[:, 1]
    worlds = (p1[:, :, None] * p2[:, None, :]).reshape(p1.shape[0], -1)
    q = torch.sum(w_q[:, query_rows].T * worlds, dim=1, keepdim=True)
    return q, worlds


def evidence_worlds(facts_probs, w_q, evidence):
    """SURVEY.md 'Evidence-mode semantics': P(w|e) = [w |= e] P(w) / sum_{w' |= e} P(w')
    (what model_dict['evidence'][e].evaluate returns before extract_*; mirrors
    compute_conditioned_worlds_prob, models/vael.py:416-430)."""
    _, worlds = dense_inference(facts_probs, w_q, evidence)
    m = w_q[:, evidence]
    return worlds * m / torch.sum(worlds * m, dim=1, keepdim=True)


def gumbel_hard(logits, gumbels, tau=1.0):
    """torch.nn.functional.gumbel_softmax(hard=True) with the noise passed in
    (models/vael.py:61-62): y_soft = softmax((logits+g)/tau); straight-through one-hot."""
    y_soft = torch.softmax((logits + gumbels) / tau, dim=-1)
    index = y_soft.max(dim=-1, keepdim=True)[1]
    y_hard = torch.zeros_like(logits).scatter_(-1, index, 1.0)
    return y_hard - y_soft.detach() + y_soft, y_soft, index.squeeze(-1)


def laplace_loglik(recon, x):
    """utils/mnist_utils/train.py:54-55 — log Laplace(x | recon, b=1) summed over (C,H,W)."""
    return (-math.log(2.0) - (x - recon).abs()).sum(dim=(1, 2, 3))


def elbo_terms(recon, x, mu, logvar, query_prob, recon_w=1.0, kl_w=1.0, query_w=1.0):
    """utils/mnist_utils/train.py:13-51 with rec_loss='LAPLACE', query=True, sup=False."""
    recon_loss = -laplace_loglik(recon, x).mean()
    kl = torch.mean(-0.5 * torch.sum(1.0 + logvar - mu.pow(2) - logvar.exp(), dim=1))
    q = torch.flatten(query_prob)
    qbce = F.binary_cross_entropy(q, torch.ones_like(q), reduction="mean")
    return recon_w * recon_loss + kl_w * kl + query_w * qbce, recon_loss, kl, qbce


# ---- Mario ------------------------------------------------------------------------------
def mario_facts(logits1, logits2, eps=1e-5):
    """models/vael.py:319-323 — softmax, clip(eps, 0.9999), concatenate the two frames."""
    f1 = torch.clip(torch.softmax(logits1, dim=1), min=eps, max=0.9999)
    f2 = torch.clip(torch.softmax(logits2, dim=1), min=eps, max=0.9999)
    return torch.cat([f1, f2], dim=1)


def mario_worlds(facts, W_all):
    """models/vael.py:359-366 — exp(log p @ W_all^T)."""
    return torch.exp(torch.log(facts) @ W_all.T)


def mario_admissible_logprob(facts, W_adm):
    """models/vael.py:368-375 — log p @ W_adm^T + log(1-p) @ (1-W_adm)^T."""
    return torch.log(facts) @ W_adm.T + torch.log(1 - facts) @ (1 - W_adm).T


def mario_query_prob(moves, worlds, WQ_all, eps=1e-5):
    """models/vael.py:377-399 — per sample: mask = rows of WQ_all equal to (move one-hot, 1);
    p_q = sum(mask * P(w)) / (sum_w P(w) + eps).  Vectorised restatement of the loop."""
    target = torch.cat([moves.to(WQ_all.dtype), torch.ones(moves.shape[0], 1, dtype=WQ_all.dtype)], dim=1)
    mask = (WQ_all[None, :, :] == target[:, None, :]).all(dim=2).to(worlds.dtype)
    return (mask * worlds).sum(dim=1) / (worlds.sum(dim=1) + eps)


def mario_conditioned_worlds(facts, W_adm, WQ_adm, evidence, eps=1e-5):
    """models/vael.py:416-430 — P(w|e) over the admissible worlds, eps placement kept."""
    p_w = torch.exp(torch.log(facts) @ W_adm.T)
    m = (WQ_adm == torch.as_tensor(evidence, dtype=WQ_adm.dtype)).all(dim=1)
    return (p_w * m + eps) / (torch.sum(p_w * m) + eps)
