"""fp64 torch interpreter of a compiled circuit, both semirings, differentiable — the oracle for the
adjoint of arbitrary circuits (autograd through the node-by-node semantics of
utils/graph_semiring.py:13-60: negate = 1 - a, plus = a + b, times = a * b; log semiring: times = +,
plus = logsumexp as MarioVAELModel.admissible_worlds_logprob uses it, models/vael.py:368-375).
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py)."""
import torch

LEAF_POS, LEAF_NEG, CONST, PROD, SUM, COMPL = 0, 1, 2, 3, 4, 5


def node_values(circ, facts):
    """facts: [B,F] float64 tensor (may require grad) -> list of [B] node values."""
    log_sr = bool(circ.semiring)
    B = facts.shape[0]
    val = [None] * circ.n_nodes
    for n in range(circ.n_nodes):
        k = int(circ.node_kind[n])
        ch = [val[int(c)] for c in circ.children(n)]
        if k in (LEAF_POS, LEAF_NEG):
            v = facts[:, int(circ.node_arg[n])]
            v = 1.0 - v if k == LEAF_NEG else v
            val[n] = torch.log(v) if log_sr else v
        elif k == CONST:
            val[n] = torch.full((B,), float(circ.const_val[int(circ.node_arg[n])]), dtype=facts.dtype)
        elif k == PROD:
            if not ch:
                val[n] = torch.full((B,), 0.0 if log_sr else 1.0, dtype=facts.dtype)
            else:
                acc = ch[0]
                for v in ch[1:]:
                    acc = acc + v if log_sr else acc * v
                val[n] = acc
        else:  # SUM / COMPL
            if not ch:
                val[n] = torch.full((B,), (float("-inf") if log_sr else 0.0) if k == SUM else 1.0, dtype=facts.dtype)
            elif log_sr:
                val[n] = torch.logsumexp(torch.stack(ch, 0), 0)
            else:
                acc = ch[0]
                for v in ch[1:]:
                    acc = acc + v
                val[n] = 1.0 - acc if k == COMPL else acc
    return val


def forward(circ, facts, query=None, normalize=False):
    """-> (worlds [B,W], query [B] or None); Z is treated as a constant by the adjoint (detached)."""
    val = node_values(circ, facts)
    B = facts.shape[0]
    worlds = torch.stack([val[int(n)] for n in circ.world_node], 1) if circ.n_worlds else facts.new_zeros(B, 0)
    q = None
    if query is not None:
        qi = torch.as_tensor(query).expand(B).long() if not torch.is_tensor(query) or query.dim() == 0 else query.long()
        qv = torch.stack([val[int(n)] for n in circ.query_node], 1)
        q = qv[torch.arange(B), qi]
    if normalize and circ.norm_node >= 0:
        z = val[circ.norm_node].detach()
        worlds = worlds / z[:, None]
        q = q / z if q is not None else None
    return worlds, q
