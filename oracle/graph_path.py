"""Node-by-node evaluation of a compiled circuit through a semiring object — the oracle
for `sdd.evaluate(semiring=GraphSemiring)` and the extract_* glue around it.

Follows models/vael.py:98-135 (call), :227-243 (leaf binding), :245-266 (extraction, incl.
`+1e-7` and the GLOBAL detached sum) and problog's evaluation flow as recalled in
SURVEY.md Appendix A: leaves take `semiring.value(label)`; AD members have negative weight
one(); every AD gets a null-choice weight ad_complement(ws); a node is `plus` over ... of
`times(...)`; each labelled query is normalize(WMC(query), WMC(constraints)).
The traversal order over children is the CSR order (left fold), which is also what the
CUDA kernels do.
"""
from collections import OrderedDict

import torch

LEAF_POS, LEAF_NEG, CONST, PROD, SUM, COMPL = 0, 1, 2, 3, 4, 5


def bind_weights(facts_probs, labels):
    """{label: facts_probs[:, col]} — strided [B] views like update_semiring_weights builds
    (models/vael.py:236-240: weights[Term('p_ij')] = facts_probs[:, i-1, j])."""
    flat = facts_probs.reshape(facts_probs.shape[0], -1)
    return {lab: flat[:, c] for c, lab in enumerate(labels)}


def node_values(circ, semiring):
    """Bottom-up sweep; returns the list of node values ([B] tensors)."""
    val = [None] * circ.n_nodes
    for n in range(circ.n_nodes):
        k = int(circ.node_kind[n])
        ch = [val[int(c)] for c in circ.children(n)]
        if k == LEAF_POS:
            val[n] = semiring.value(circ.fact_labels[int(circ.node_arg[n])])
        elif k == LEAF_NEG:
            val[n] = semiring.negate(semiring.value(circ.fact_labels[int(circ.node_arg[n])]))
        elif k == CONST:
            c = float(circ.const_val[int(circ.node_arg[n])])
            val[n] = semiring.zero() if c == 0.0 else semiring.one() * c
        elif k == PROD:
            acc = ch[0]
            for v in ch[1:]:
                acc = semiring.times(acc, v)
            val[n] = acc
        elif k == SUM:
            acc = ch[0]
            for v in ch[1:]:
                acc = semiring.plus(acc, v)
            val[n] = acc
        elif k == COMPL:  # ad_complement: negate(plus-fold(ws))
            acc = ch[0]
            for v in ch[1:]:
                acc = semiring.plus(acc, v)
            val[n] = semiring.negate(acc)
        else:
            raise ValueError(k)
    return val


def evaluate(circ, semiring, mode="query", query_index=-1, world_keys=None, query_key=None):
    """OrderedDict {atom: [B] tensor} like sdd.evaluate: the world atoms in grounding order,
    then (query mode) the label atom LAST (extract_query_probability reads the last key,
    models/vael.py:245-248).  Every entry is normalize(wmc, Z)."""
    val = node_values(circ, semiring)
    z = val[circ.norm_node] if circ.norm_node >= 0 else None
    world_keys = world_keys if world_keys is not None else circ.world_atoms
    res = OrderedDict()
    if mode == "query":
        for key, n in zip(world_keys, circ.world_node):
            v = val[int(n)]
            res[key] = semiring.normalize(v, z) if z is not None else v
        if query_index >= 0:
            v = val[int(circ.query_node[query_index])]
            res[query_key if query_key is not None else circ.query_atoms[query_index]] = \
                semiring.normalize(v, z) if z is not None else v
    else:  # evidence: WMC(world and e) / WMC(e); worlds violating e evaluate to zero()
        qn = int(circ.query_node[query_index])
        members = set(int(c) for c in circ.children(qn)) if int(circ.node_kind[qn]) == SUM else {qn}
        pe = val[qn]
        for key, n in zip(world_keys, circ.world_node):
            res[key] = semiring.normalize(val[int(n)], pe) if int(n) in members else semiring.zero()
    return res


def extract_query_probability(res):
    """models/vael.py:245-248 — last entry, as a [B,1] column."""
    return res[next(reversed(res))][..., None]


def extract_worlds_probability(res, world_keys):
    """models/vael.py:250-266 — stack j-major, +1e-7, divide by the detached sum over the
    WHOLE [B,W] tensor."""
    p = torch.stack([res[k] for k in world_keys], dim=1) + 1e-7
    return p / p.detach().sum()
