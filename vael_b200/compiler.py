"""Host circuit compiler: ProbLog program text -> flat level-ordered CSR circuit.

Stands in for `LogicFormula.create_from -> LogicDAG.create_from -> SDD.create_from`
(utils/mnist_utils/mnist_task_VAEL.py:214-216) and for the offline world enumeration of
utils/mario_utils/problog_matrices.py:14-126.  Knowledge compilation is by enumeration of
the total choices of the annotated disjunctions (the programs VAEL uses have 81-1000 of
them): the result is a smooth, deterministic, decomposable circuit

    level 0  one positive literal per AD head      (fact column = order in the text)
    level 1+ nested products: one literal per AD   (world w = mixed-radix index, first AD
                                                   most significant: d1*10+d2, models/vael.py:218-221)
             a SUM node per ground query atom over the worlds where SLD proves it
             the ProbLog normaliser Z = prod_AD ( sum(p) + (1 - sum(p)) )   (null-choice atom)

Runs once per program family on the host; nothing here is on the per-batch path.
"""
from __future__ import annotations

import itertools
from dataclasses import dataclass, field
from typing import Dict, Hashable, List, Mapping, Optional, Sequence, Tuple

import numpy as np

from .circuit import SEMIRING_LOG, SEMIRING_PROB, Circuit, CircuitBuilder
from .logic import AnnotatedDisjunction, Engine, Program, Term, Var, parse_program, substitute


class CompileError(ValueError):
    pass


@dataclass
class ProgramView:
    """One program of a family, e.g. model_dict['evidence'][7]: which rows of the shared
    circuit it reads and in which order `evaluate` lists its results."""
    key: Hashable
    text: str
    mode: str                      # 'query' | 'evidence'
    query_index: int               # index into circuit.query_node (the addition/evidence atom), or -1
    result_atoms: List[Term] = field(default_factory=list)  # world atoms, then (query mode) the label atom


@dataclass
class CompiledFamily:
    circuit: Circuit
    views: Dict[Hashable, ProgramView]
    world_terms: List[Term]
    query_terms: List[Term]
    fact_terms: List[Term]         # probability labels, column order


def _strip(t: Term) -> Term:
    return t.with_probability(None)


def _all_heads_facts(ads: Sequence[AnnotatedDisjunction]):
    facts: Dict[Tuple[str, int], List[Tuple[Term, tuple]]] = {}
    for k, ad in enumerate(ads):
        for i, h in enumerate(ad.heads):
            facts.setdefault(h.signature, []).append((_strip(h), (k, i)))
    return facts


def _world_facts(ads, choice):
    facts: Dict[Tuple[str, int], List[Tuple[Term, tuple]]] = {}
    for k, i in enumerate(choice):
        h = ads[k].heads[i]
        facts.setdefault(h.signature, []).append((_strip(h), (k, i)))
    return facts


def _check_programs_share_logic(programs: Mapping[Hashable, Program]) -> Program:
    first = None
    for key, p in programs.items():
        if first is None:
            first = p
        elif p.logic_signature() != first.logic_signature():
            raise CompileError(f"program {key!r} does not share facts and rules with the rest of the family")
    if first is None:
        raise CompileError("empty program family")
    return first


def compile_family(texts: Mapping[Hashable, str], world_query: Optional[str] = None,
                   with_normaliser: bool = True) -> CompiledFamily:
    """Compile a family of programs that share facts and rules and differ only in their
    query / evidence directives (the 38 programs of build_model_dict) into ONE circuit.

    world_query: the query whose ground instances define the world outputs
                 ('digits(X,Y)' — utils/mnist_utils/mnist_task_VAEL.py:213).  If None, the
                 worlds are the total choices themselves (Mario, problog_matrices.py:38-47).
    """
    programs = {k: parse_program(t) for k, t in texts.items()}
    base = _check_programs_share_logic(programs)
    ads = base.ads
    if not ads:
        raise CompileError("no probabilistic facts")
    for ad in ads:
        if ad.body is not None:
            raise CompileError("annotated disjunctions with bodies are not supported")
    engine = Engine(base.clauses)

    # fact columns: one per AD head, in text order
    fact_terms: List[Term] = []
    col_of: Dict[Tuple[int, int], int] = {}
    for k, ad in enumerate(ads):
        for i, h in enumerate(ad.heads):
            col_of[(k, i)] = len(fact_terms)
            fact_terms.append(h.probability)
    F = len(fact_terms)

    # ---- ground the directives against "every head possible" (relevance grounding) -----
    possible = _all_heads_facts(ads)
    engine.trace = []

    def ground(atom: Term) -> List[Term]:
        return [atom] if atom.is_ground() else engine.solutions(atom, possible)

    wq_term = parse_program(f"query({world_query}).").queries[0] if world_query else None
    world_terms: List[Term] = ground(wq_term) if wq_term is not None else []
    query_terms: List[Term] = []
    views: Dict[Hashable, ProgramView] = {}
    for key, p in programs.items():
        label_atoms: List[Term] = []
        for q in p.queries:
            if wq_term is not None and str(q) == str(wq_term):
                continue
            label_atoms.extend(ground(q))
        ev_atoms = []
        for e, val in p.evidence:
            if not val:
                raise CompileError("negative evidence is not supported")
            ev_atoms.extend(ground(e))
        if ev_atoms and len(ev_atoms) != 1:
            raise CompileError("exactly one evidence atom per program is supported")
        for a in label_atoms + ev_atoms:
            if a not in query_terms:
                query_terms.append(a)
        if ev_atoms:
            views[key] = ProgramView(key, texts[key], "evidence", query_terms.index(ev_atoms[0]),
                                     list(world_terms) + label_atoms)
        else:
            qi = query_terms.index(label_atoms[-1]) if label_atoms else -1
            views[key] = ProgramView(key, texts[key], "query", qi, list(world_terms) + label_atoms)
    # every AD may be reached through one grounding only (non-ground ADs are one choice here)
    seen: Dict[tuple, Term] = {}
    for tag, inst in engine.trace:
        if tag in seen and seen[tag] != inst:
            raise CompileError(f"annotated disjunction reached through two groundings: {seen[tag]} / {inst}")
        seen[tag] = inst
    engine.trace = None

    # ---- enumerate total choices; which atoms hold where ----------------------------------
    sizes = [len(ad.heads) for ad in ads]
    choices = list(itertools.product(*[range(n) for n in sizes]))
    n_tc = len(choices)
    holds_world = {a: [] for a in world_terms}
    holds_query = {a: [] for a in query_terms}
    for w, choice in enumerate(choices):
        facts = _world_facts(ads, choice)
        if wq_term is not None:
            for a in engine.solutions(wq_term, facts):
                if a in holds_world:
                    holds_world[a].append(w)
        for a in query_terms:
            if engine.holds(a, facts):
                holds_query[a].append(w)

    # ---- build the DAG -------------------------------------------------------------------------
    b = CircuitBuilder(F, SEMIRING_PROB)
    leaves = [[b.leaf(col_of[(k, i)]) for i in range(sizes[k])] for k in range(len(ads))]
    prefix = {(i,): leaves[0][i] for i in range(sizes[0])}
    for k in range(1, len(ads)):
        nxt = {}
        for pre, node in prefix.items():
            for i in range(sizes[k]):
                nxt[pre + (i,)] = b.prod((node, leaves[k][i]))
        prefix = nxt
    tc_node = [prefix[c] for c in choices]

    def atom_node(ws: List[int]) -> int:
        if not ws:
            return b.const(0.0)
        if len(ws) == 1:
            return tc_node[ws[0]]
        return b.sum(tuple(tc_node[w] for w in ws))

    if wq_term is not None:
        world_nodes = [atom_node(holds_world[a]) for a in world_terms]
        world_names = [str(a) for a in world_terms]
    else:
        world_nodes = list(tc_node)
        world_names = ["world(" + ",".join(str(i) for i in c) + ")" for c in choices]
    query_nodes = [atom_node(holds_query[a]) for a in query_terms]

    norm = -1
    if with_normaliser:
        per_ad = []
        for k in range(len(ads)):
            c = b.compl(tuple(leaves[k]))
            per_ad.append(b.sum(tuple(leaves[k]) + (c,)))
        norm = per_ad[0] if len(per_ad) == 1 else b.prod(tuple(per_ad))

    circ = b.finish(world_nodes, query_nodes, norm, fact_labels=[str(t) for t in fact_terms],
                    world_atoms=world_names, query_atoms=[str(a) for a in query_terms])
    return CompiledFamily(circ, views, world_terms, query_terms, fact_terms)


def circuit_from_world_table(table: np.ndarray, semiring: int = SEMIRING_LOG, with_negatives: bool = True) -> Circuit:
    """Circuit whose world w is the product over ALL fact columns f of the literal
    (f if table[w,f] else not f) — MarioVAELModel.admissible_worlds_logprob
    (models/vael.py:368-375: log p @ W_adm^T + log(1-p) @ (1-W_adm)^T) when evaluated in the
    log semiring.  With with_negatives=False only the positive literals are multiplied
    (compute_worlds_prob_conj, models/vael.py:359-366)."""
    table = np.asarray(table)
    W, F = table.shape
    b = CircuitBuilder(F, semiring)
    pos = [b.leaf(f, True) for f in range(F)]
    neg = [b.leaf(f, False) for f in range(F)] if with_negatives else None
    worlds = []
    for w in range(W):
        lits = [pos[f] for f in range(F) if table[w, f] != 0]
        if with_negatives:  # positives first, then negatives: the order the reference adds the two matmuls
            lits += [neg[f] for f in range(F) if table[w, f] == 0]
        worlds.append(b.prod(tuple(lits)))
    return b.finish(worlds, [], -1, fact_labels=[f"f_{f}" for f in range(F)],
                    world_atoms=[f"row{w}" for w in range(W)], query_atoms=[])
