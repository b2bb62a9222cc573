"""Flat, level-ordered CSR arithmetic circuit: the host-side product of the compiler
and the thing that is uploaded to HBM (include/vael_b200.h: vael_circuit_desc_t).

It takes the place of the SDD objects the reference keeps in `model_dict`
(utils/mnist_utils/mnist_task_VAEL.py:208-217, evaluated at models/vael.py:106-109,
:127-130).  Arrays are plain numpy so they can be hashed, diffed against the goldens
under tests/golden/ bit for bit, and handed to the C-ABI through ctypes.
"""
from __future__ import annotations

import hashlib
import io
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

# node kinds / semirings — keep in sync with include/vael_b200.h
LEAF_POS, LEAF_NEG, CONST, PROD, SUM, COMPL = 0, 1, 2, 3, 4, 5
SEMIRING_PROB, SEMIRING_LOG = 0, 1
FLAG_NORMALIZE, FLAG_EVIDENCE, FLAG_FORCE_GENERIC, FLAG_CHECK_QUERY, FLAG_FORCE_JIT = 1, 2, 4, 8, 16
KIND_NAMES = {LEAF_POS: "leaf+", LEAF_NEG: "leaf-", CONST: "const", PROD: "prod", SUM: "sum", COMPL: "compl"}


@dataclass
class Circuit:
    n_facts: int
    semiring: int
    node_kind: np.ndarray   # uint8  [n_nodes]
    node_arg: np.ndarray    # int32  [n_nodes]
    level_ptr: np.ndarray   # int32  [n_levels+1]
    child_ptr: np.ndarray   # int32  [n_nodes+1]
    child_idx: np.ndarray   # int32  [n_edges]
    const_val: np.ndarray   # float32[n_consts]
    world_node: np.ndarray  # int32  [n_worlds]
    query_node: np.ndarray  # int32  [n_queries]
    norm_node: int = -1
    # symbolic side tables (not uploaded)
    fact_labels: List[str] = field(default_factory=list)   # column -> weight label ('p_10', ...)
    world_atoms: List[str] = field(default_factory=list)   # world -> ground atom ('digits(0,0)', ...)
    query_atoms: List[str] = field(default_factory=list)   # query -> ground atom ('addition(img,0)', ...)

    def __post_init__(self):
        self.node_kind = np.ascontiguousarray(self.node_kind, dtype=np.uint8)
        for name in ("node_arg", "level_ptr", "child_ptr", "child_idx", "world_node", "query_node"):
            setattr(self, name, np.ascontiguousarray(getattr(self, name), dtype=np.int32))
        self.const_val = np.ascontiguousarray(self.const_val, dtype=np.float32)

    # ---- sizes ------------------------------------------------------------
    @property
    def n_nodes(self): return int(self.node_kind.shape[0])
    @property
    def n_levels(self): return int(self.level_ptr.shape[0] - 1)
    @property
    def n_edges(self): return int(self.child_idx.shape[0])
    @property
    def n_worlds(self): return int(self.world_node.shape[0])
    @property
    def n_queries(self): return int(self.query_node.shape[0])

    def children(self, n: int) -> np.ndarray:
        return self.child_idx[self.child_ptr[n]:self.child_ptr[n + 1]]

    def level_of(self) -> np.ndarray:
        lv = np.zeros(self.n_nodes, dtype=np.int32)
        for L in range(self.n_levels):
            lv[self.level_ptr[L]:self.level_ptr[L + 1]] = L
        return lv

    def validate(self) -> None:
        """Host-side twin of the checks vael_circuit_create makes; raises ValueError."""
        def need(ok, msg):
            if not ok:
                raise ValueError(f"invalid circuit: {msg}")
        need(self.level_ptr[0] == 0 and self.level_ptr[-1] == self.n_nodes, "level_ptr does not cover the nodes")
        need(self.child_ptr[0] == 0 and self.child_ptr[-1] == self.n_edges, "child_ptr does not cover the edges")
        lv = self.level_of()
        for n in range(self.n_nodes):
            k = int(self.node_kind[n])
            ch = self.children(n)
            if k in (LEAF_POS, LEAF_NEG):
                need(0 <= self.node_arg[n] < self.n_facts and len(ch) == 0 and lv[n] == 0, f"bad leaf {n}")
            elif k == CONST:
                need(0 <= self.node_arg[n] < len(self.const_val) and len(ch) == 0, f"bad constant {n}")
            else:
                need(k in (PROD, SUM, COMPL), f"unknown node kind {k}")
                need(all(0 <= c < self.n_nodes and lv[c] < lv[n] for c in ch), "children must be in strictly lower levels")
        need(all(0 <= w < self.n_nodes for w in self.world_node), "world node out of range")
        need(all(0 <= q < self.n_nodes for q in self.query_node), "query node out of range")

    # ---- identity -----------------------------------------------------------
    def serialize(self) -> bytes:
        """Canonical byte string of everything the device sees."""
        buf = io.BytesIO()
        hdr = np.array([self.n_facts, self.n_nodes, self.n_levels, self.n_edges, self.n_worlds, self.n_queries,
                        len(self.const_val), self.semiring, self.norm_node], dtype=np.int32)
        for a in (hdr, self.node_kind, self.node_arg, self.level_ptr, self.child_ptr, self.child_idx,
                  self.const_val, self.world_node, self.query_node):
            buf.write(np.ascontiguousarray(a).tobytes())
        return buf.getvalue()

    def sha256(self) -> str:
        return hashlib.sha256(self.serialize()).hexdigest()

    def fnv1a64(self) -> int:
        """Same hash vael_circuit_hash() computes over the uploaded arrays."""
        h = 1469598103934665603
        for b in self.serialize():
            h = ((h ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
        return h

    def to_npz_dict(self) -> dict:
        return dict(n_facts=np.int32(self.n_facts), semiring=np.int32(self.semiring), node_kind=self.node_kind,
                    node_arg=self.node_arg, level_ptr=self.level_ptr, child_ptr=self.child_ptr,
                    child_idx=self.child_idx, const_val=self.const_val, world_node=self.world_node,
                    query_node=self.query_node, norm_node=np.int32(self.norm_node),
                    fact_labels=np.array(self.fact_labels), world_atoms=np.array(self.world_atoms),
                    query_atoms=np.array(self.query_atoms))

    def save(self, path) -> None:
        np.savez(path, **self.to_npz_dict())

    @staticmethod
    def load(path) -> "Circuit":
        z = np.load(path, allow_pickle=False)
        return Circuit(n_facts=int(z["n_facts"]), semiring=int(z["semiring"]), node_kind=z["node_kind"],
                       node_arg=z["node_arg"], level_ptr=z["level_ptr"], child_ptr=z["child_ptr"],
                       child_idx=z["child_idx"], const_val=z["const_val"], world_node=z["world_node"],
                       query_node=z["query_node"], norm_node=int(z["norm_node"]),
                       fact_labels=[str(s) for s in z["fact_labels"]], world_atoms=[str(s) for s in z["world_atoms"]],
                       query_atoms=[str(s) for s in z["query_atoms"]])

    # ---- derived tables -------------------------------------------------------
    def worlds_queries_matrix(self) -> np.ndarray:
        """w_q[w, q] = 1 iff world w is a child of query node q — what the reference builds
        by hand in build_worlds_queries_matrix (utils/mnist_utils/mnist_task_VAEL.py:222-234)."""
        pos = {int(n): w for w, n in enumerate(self.world_node)}
        wq = np.zeros((self.n_worlds, self.n_queries), dtype=np.float32)
        for q, n in enumerate(self.query_node):
            if self.node_kind[n] == SUM:
                for c in self.children(int(n)):
                    if int(c) in pos:
                        wq[pos[int(c)], q] = 1.0
            elif int(n) in pos:
                wq[pos[int(n)], q] = 1.0
        return wq

    def describe(self) -> str:
        per_level = [int(self.level_ptr[i + 1] - self.level_ptr[i]) for i in range(self.n_levels)]
        return (f"Circuit(F={self.n_facts}, nodes={self.n_nodes}, edges={self.n_edges}, levels={per_level}, "
                f"W={self.n_worlds}, Q={self.n_queries}, norm={self.norm_node}, "
                f"semiring={'log' if self.semiring else 'prob'})")


class CircuitBuilder:
    """Hash-consing DAG builder; `finish` orders nodes level by level (stable in creation
    order inside a level) and emits the CSR arrays."""

    def __init__(self, n_facts: int, semiring: int = SEMIRING_PROB):
        self.n_facts, self.semiring = n_facts, semiring
        self.kind: List[int] = []
        self.arg: List[int] = []
        self.ch: List[tuple] = []
        self.consts: List[float] = []
        self._unique = {}

    def _node(self, kind, arg, ch) -> int:
        key = (kind, arg, tuple(ch))
        if key in self._unique:
            return self._unique[key]
        self.kind.append(kind); self.arg.append(arg); self.ch.append(tuple(ch))
        self._unique[key] = len(self.kind) - 1
        return len(self.kind) - 1

    def leaf(self, col: int, positive: bool = True) -> int:
        return self._node(LEAF_POS if positive else LEAF_NEG, col, ())

    def const(self, value: float) -> int:
        v = float(np.float32(value))
        if v not in self.consts:
            self.consts.append(v)
        return self._node(CONST, self.consts.index(v), ())

    def prod(self, ch) -> int: return self._node(PROD, 0, ch)
    def sum(self, ch) -> int: return self._node(SUM, 0, ch)
    def compl(self, ch) -> int: return self._node(COMPL, 0, ch)

    def finish(self, world_node, query_node, norm_node=-1, **tables) -> Circuit:
        n = len(self.kind)
        level = [0] * n
        for i in range(n):  # children are always created before their parents
            if self.ch[i]:
                level[i] = 1 + max(level[c] for c in self.ch[i])
        order = sorted(range(n), key=lambda i: (level[i], i))
        new_id = {old: new for new, old in enumerate(order)}
        n_levels = max(level) + 1 if n else 0
        level_ptr = [0] * (n_levels + 1)
        for i in range(n):
            level_ptr[level[i] + 1] += 1
        for L in range(n_levels):
            level_ptr[L + 1] += level_ptr[L]
        child_ptr, child_idx = [0], []
        for old in order:
            child_idx.extend(new_id[c] for c in self.ch[old])
            child_ptr.append(len(child_idx))
        circ = Circuit(
            n_facts=self.n_facts, semiring=self.semiring,
            node_kind=np.array([self.kind[o] for o in order], dtype=np.uint8),
            node_arg=np.array([self.arg[o] for o in order], dtype=np.int32),
            level_ptr=np.array(level_ptr, dtype=np.int32), child_ptr=np.array(child_ptr, dtype=np.int32),
            child_idx=np.array(child_idx, dtype=np.int32), const_val=np.array(self.consts, dtype=np.float32),
            world_node=np.array([new_id[w] for w in world_node], dtype=np.int32),
            query_node=np.array([new_id[q] for q in query_node], dtype=np.int32),
            norm_node=(new_id[norm_node] if norm_node >= 0 else -1), **tables)
        circ.validate()
        return circ
