"""Mario grid-navigation program family: text generation, compilation, tables.

Host-side replacement of utils/mario_utils/problog_model.py:3-62 (the move-rule program and
its 2x9 annotated-disjunction facts) and of utils/mario_utils/problog_matrices.py:14-126,
which enumerates the 81 worlds with problog/d-DNNF offline and stores four 0/1 tables
(utils/mario_utils/problog_matrices/3x3/{W_all,WQ_all,W_adm,WQ_adm}.pt).  Here the same
tables fall out of the compiled circuit and are checked bit for bit against goldens made
from the reference's .pt files (tests/golden/mario.npz).
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np

from .circuit import SEMIRING_LOG, SEMIRING_PROB, Circuit, CircuitBuilder
from .compiler import CompiledFamily, circuit_from_world_table, compile_family

MOVES = ("up", "down", "right", "left")          # label column order, utils/mario_utils/problog_model.py:53-56
QUERIES = tuple(f"action({m})" for m in MOVES) + ("constraint",)
_DELTA = {"up": (0, 1), "down": (0, -1), "right": (1, 0), "left": (-1, 0)}


def snake_positions(width: int = 3, height: int = 3) -> List[Tuple[int, int]]:
    """Fact order of one frame: rows bottom-up, alternating direction
    ((0,0),(1,0),(2,0),(2,1),(1,1),(0,1),(0,2),...: problog_model.py:31-51)."""
    out = []
    for y in range(height):
        xs = range(width) if y % 2 == 0 else range(width - 1, -1, -1)
        out.extend((x, y) for x in xs)
    return out


def fact_labels(grid=(3, 3)) -> List[str]:
    n = grid[0] * grid[1]
    return [f"a_{t}{i}" for t in range(2) for i in range(n)]


def mario_program(grid=(3, 3), queries=QUERIES) -> str:
    """Two annotated disjunctions (agent position at t=0 and t=1), one rule per move and the
    'exactly one step' constraint as a disjunction of the moves."""
    pos = snake_positions(*grid)
    labels = fact_labels(grid)
    lines = ["% Agent positions (X,Y,time)"]
    for t in range(2):
        heads = [f"{labels[t * len(pos) + i]}::agent({x},{y},{t})" for i, (x, y) in enumerate(pos)]
        lines.append("; ".join(heads) + ".")
    lines.append("% Possible actions")
    for m in MOVES:
        dx, dy = _DELTA[m]
        xs = "X1 is X" + (f"+{dx}" if dx > 0 else (f"-{-dx}" if dx < 0 else ""))
        ys = "Y1 is Y" + (f"+{dy}" if dy > 0 else (f"-{-dy}" if dy < 0 else ""))
        lines.append(f"action({m}) :- agent(X,Y,T), agent(X1,Y1,T1), {xs}, {ys}, T1 is T+1.")
    lines.append("constraint :- " + ";".join(f"action({m})" for m in MOVES) + ".")
    lines.extend(f"query({q})." for q in queries)
    return "\n".join(lines) + "\n"


def compile_mario_family(grid=(3, 3)) -> CompiledFamily:
    """81 worlds (product order, first frame most significant) and the five query nodes."""
    return compile_family({"base": mario_program(grid)}, world_query=None)


def mario_tables(fam: CompiledFamily):
    """(W_all [81,18], WQ_all [81,5], W_adm [24,18], WQ_adm [24,4]) as int8 arrays, from the circuit."""
    circ = fam.circuit
    n_tc = circ.n_worlds
    sizes = [9, 9] if circ.n_facts == 18 else None
    n = circ.n_facts // 2
    W_all = np.zeros((n_tc, circ.n_facts), dtype=np.int8)
    for w in range(n_tc):
        W_all[w, w // n] = 1
        W_all[w, n + w % n] = 1
    WQ_all = circ.worlds_queries_matrix().astype(np.int8)
    adm = WQ_all[:, -1] == 1
    return W_all, WQ_all, W_all[adm], WQ_all[adm][:, :4]


def admissible_circuit(W_adm: np.ndarray) -> Circuit:
    """Log-semiring circuit of admissible_worlds_logprob (models/vael.py:368-375)."""
    return circuit_from_world_table(W_adm, semiring=SEMIRING_LOG, with_negatives=True)


def admissible_evidence_circuit(W_adm: np.ndarray, WQ_adm: np.ndarray) -> Circuit:
    """Probability-semiring circuit over the ADMISSIBLE worlds only: world w = product of its positive literals
    (`exp(log p @ W_adm^T)`, models/vael.py:421-423), one query node per move = sum of the admissible worlds where
    that move holds (WQ_adm column).  Evaluated with VAEL_FLAG_EVIDENCE it yields P(w | move) over the 24 worlds:
    MarioVAELModel.compute_conditioned_worlds_prob / compute_query_prob_cond (models/vael.py:401-430)."""
    W_adm, WQ_adm = np.asarray(W_adm), np.asarray(WQ_adm)
    W, F = W_adm.shape
    b = CircuitBuilder(F, SEMIRING_PROB)
    pos = [b.leaf(f, True) for f in range(F)]
    worlds = [b.prod(tuple(pos[f] for f in range(F) if W_adm[w, f] != 0)) for w in range(W)]
    queries = []
    for m in range(WQ_adm.shape[1]):
        ws = [worlds[w] for w in range(W) if WQ_adm[w, m] != 0]
        queries.append(b.sum(tuple(ws)) if len(ws) > 1 else (ws[0] if ws else b.const(0.0)))
    return b.finish(worlds, queries, -1, fact_labels=fact_labels(), world_atoms=[f"adm{w}" for w in range(W)],
                    query_atoms=[f"action({m})" for m in MOVES[:WQ_adm.shape[1]]])
