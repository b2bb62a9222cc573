"""Elementwise numpy interpreter of a compiled circuit (no semiring object, no batch-wide
shortcuts): the fp32 variant performs the same IEEE operations in the same (CSR, left-fold)
order as the CUDA kernels' forward, so forward parity against it is BIT-exact; the fp64
variant separates "rounding" from "wrong".  Semantics per node kind as in
oracle/graph_path.py (utils/graph_semiring.py:13-60).
"""
import numpy as np

LEAF_POS, LEAF_NEG, CONST, PROD, SUM, COMPL = 0, 1, 2, 3, 4, 5


def node_values(circ, facts, dtype=np.float32, log_semiring=False):
    facts = np.asarray(facts).reshape(len(facts), -1).astype(dtype)
    one = dtype(1.0)
    val = [None] * circ.n_nodes
    with np.errstate(divide="ignore", invalid="ignore"):
        for n in range(circ.n_nodes):
            k = int(circ.node_kind[n])
            ch = [val[int(c)] for c in circ.children(n)]
            if k in (LEAF_POS, LEAF_NEG):
                v = facts[:, int(circ.node_arg[n])]
                v = (one - v) if k == LEAF_NEG else v
                val[n] = np.log(v).astype(dtype) if log_semiring else v
            elif k == CONST:
                val[n] = np.full(facts.shape[0], circ.const_val[int(circ.node_arg[n])], dtype=dtype)
            elif k == PROD:
                acc = ch[0]
                for v in ch[1:]:
                    acc = (acc + v) if log_semiring else (acc * v)
                val[n] = acc.astype(dtype)
            elif k in (SUM, COMPL):
                if log_semiring:
                    st = np.stack(ch, axis=0)
                    mx = st.max(axis=0)
                    val[n] = (mx + np.log(np.exp(st - mx).sum(axis=0))).astype(dtype)
                else:
                    acc = ch[0]
                    for v in ch[1:]:
                        acc = acc + v
                    val[n] = ((one - acc) if k == COMPL else acc).astype(dtype)
    return val


def forward(circ, facts, query=None, normalize=False, evidence=False, dtype=np.float32):
    """-> (worlds [B,W], query [B]) with the kernels' flag semantics (include/vael_b200.h)."""
    log_sr = bool(circ.semiring)
    val = node_values(circ, facts, dtype, log_sr)
    B = len(facts)
    worlds = np.stack([val[int(n)] for n in circ.world_node], axis=1) if circ.n_worlds else np.zeros((B, 0), dtype)
    z = val[circ.norm_node] if (normalize and circ.norm_node >= 0) else None
    q = None
    if query is not None:
        qi = np.broadcast_to(np.asarray(query), (B,)).astype(np.int64)
        qvals = np.stack([val[int(n)] for n in circ.query_node], axis=1)
        q = qvals[np.arange(B), qi]
    with np.errstate(divide="ignore", invalid="ignore"):
        if evidence:
            wq = circ.worlds_queries_matrix().astype(bool)  # [W,Q]
            member = wq[:, qi].T                             # [B,W]
            worlds = np.where(member, worlds / q[:, None], dtype(0)).astype(dtype)
        elif z is not None:
            worlds = (worlds / z[:, None]).astype(dtype)
        if q is not None and z is not None:
            q = (q / z).astype(dtype)
    return worlds, q


def queries_all(circ, facts, normalize=False, dtype=np.float32):
    val = node_values(circ, facts, dtype, bool(circ.semiring))
    q = np.stack([val[int(n)] for n in circ.query_node], axis=1)
    if normalize and circ.norm_node >= 0:
        q = (q / val[circ.norm_node][:, None]).astype(dtype)
    return q


def backward(circ, facts, grad_worlds=None, grad_query=None, query=None, normalize=False):
    """fp64 adjoint by explicit reverse sweep (prob semiring): d/d facts of
    sum(grad_worlds * worlds) + sum(grad_query * query).  Z is treated as a constant: its
    own derivative is identically zero (d/dp of s + (1 - s))."""
    assert not circ.semiring
    dtype = np.float64
    val = node_values(circ, facts, dtype, False)
    B = len(facts)
    adj = [np.zeros(B, dtype) for _ in range(circ.n_nodes)]
    z = val[circ.norm_node] if (normalize and circ.norm_node >= 0) else np.ones(B, dtype)
    if grad_worlds is not None:
        for w, n in enumerate(circ.world_node):
            adj[int(n)] = adj[int(n)] + np.asarray(grad_worlds, dtype)[:, w] / z
    if grad_query is not None and query is not None:
        qi = np.broadcast_to(np.asarray(query), (B,)).astype(np.int64)
        gq = np.asarray(grad_query, dtype).reshape(B) / z
        for q, n in enumerate(circ.query_node):
            adj[int(n)] = adj[int(n)] + np.where(qi == q, gq, 0.0)
    norm_cone = set()
    if circ.norm_node >= 0:  # nodes only reachable from Z carry no adjoint
        pass
    for n in range(circ.n_nodes - 1, -1, -1):
        k = int(circ.node_kind[n])
        ch = [int(c) for c in circ.children(n)]
        if k == PROD:
            for i, c in enumerate(ch):
                sib = np.ones(B, dtype)
                for j, o in enumerate(ch):
                    if j != i:
                        sib = sib * val[o]
                adj[c] = adj[c] + adj[n] * sib
        elif k == SUM:
            for c in ch:
                adj[c] = adj[c] + adj[n]
        elif k == COMPL:
            for c in ch:
                adj[c] = adj[c] - adj[n]
    g = np.zeros((B, circ.n_facts), dtype)
    for n in range(int(circ.level_ptr[1])):
        k = int(circ.node_kind[n])
        if k == LEAF_POS:
            g[:, int(circ.node_arg[n])] += adj[n]
        elif k == LEAF_NEG:
            g[:, int(circ.node_arg[n])] -= adj[n]
    return g
