"""MNIST-addition program family: text generation and compilation.

Mirrors the reference's host-side compile step with the same function names and
arguments so its driver code runs unchanged against this package:
  create_facts / define_ProbLog_model   utils/mnist_utils/problog_model.py:5-24, :27-54
  build_model_dict                       utils/mnist_utils/mnist_task_VAEL.py:203-219
  build_worlds_queries_matrix            utils/mnist_utils/mnist_task_VAEL.py:222-234
"""
from __future__ import annotations

from typing import Dict, List

import numpy as np

from .compiler import CompiledFamily, compile_family
from .logic import AnnotatedDisjunction, Constant, Term, Var

WORLD_QUERY = {2: "digits(X,Y)", 3: "digits(X,Y,Z)"}


def create_facts(sequence_len: int, n_digits: int = 10) -> List[str]:
    """One annotated disjunction per digit position: p_{pos}{y}::digit(X,pos,y) for y < n_digits."""
    digit, image = Term("digit"), Var("X")
    out = []
    for pos in range(1, sequence_len + 1):
        heads = [digit(image, Constant(pos), Constant(y), p=f"p_{pos}{y}") for y in range(n_digits)]
        out.append(f"{AnnotatedDisjunction(heads, None)}.")
    return out


def addition_rules(sequence_len: int) -> str:
    """The two rules of the reference program (mnist_task_VAEL.py:206), generalised to k digits."""
    ks = range(1, sequence_len + 1)
    lits = ", ".join(f"digit(X,{k},N{k})" for k in ks)
    total = " + ".join(f"N{k}" for k in ks)
    names = ["X", "Y", "Z", "U", "V", "W"][:sequence_len]
    world = ", ".join(f"digit(img,{k},{v})" for k, v in zip(ks, names))
    return (f"addition(X,N) :- {lits}, N is {total}.\n"
            f"digits({','.join(names)}):-{world}.")


def define_ProbLog_model(facts, rules, label, digit_query=None, mode="query") -> str:
    """Program text = ADs, rules, the world query, then the addition query or evidence."""
    parts = []
    for i, ad in enumerate(facts, 1):
        parts.append(f"\n\n% Digit in position {i}\n\n{ad}")
    parts.append(f"\n\n% Rules\n{rules}")
    if digit_query:
        parts.append(f"\n\n% Digit Query\nquery({digit_query}).")
    if mode == "query":
        parts.append(f"\n\n% Addition Query\nquery(addition(img,{label})).")
    elif mode == "evidence":
        parts.append(f"\n\n% Addition Evidence\nevidence(addition(img,{label})).")
    return "".join(parts)


def compile_addition_family(sequence_len: int = 2, n_digits: int = 10) -> CompiledFamily:
    """All (mode, sum) programs of the task compiled into one shared circuit."""
    rules = addition_rules(sequence_len)
    facts = create_facts(sequence_len, n_digits=n_digits)
    wq = WORLD_QUERY[sequence_len]
    texts = {}
    for mode in ("query", "evidence"):
        for add in range(0, (n_digits - 1) * sequence_len + 1):
            texts[(mode, add)] = define_ProbLog_model(facts, rules, label=add, digit_query=wq, mode=mode)
    return compile_family(texts, world_query=wq)


def build_model_dict(sequence_len: int, n_digits: int, device=None) -> Dict[str, Dict[int, object]]:
    """{'query': {sum: program}, 'evidence': {sum: program}}; each value offers
    .evaluate(semiring=GraphSemiring) like the SDD objects it replaces."""
    from .program import CompiledProgramSet
    fam = compile_addition_family(sequence_len, n_digits)
    pset = CompiledProgramSet(fam, device=device)
    return {mode: {add: pset.view((mode, add)) for (m, add) in fam.views if m == mode} for mode in ("query", "evidence")}


def build_worlds_queries_matrix(sequence_len: int, n_digits: int):
    """w_q[w, q] = [digits of world w add up to q], worlds in product order, n_digits*sequence_len columns
    (one more than there are sums: the last column stays zero, as in the reference)."""
    import torch
    grids = np.indices((n_digits,) * sequence_len).reshape(sequence_len, -1)
    sums = grids.sum(axis=0)
    w_q = np.zeros((sums.shape[0], n_digits * sequence_len), dtype=np.float32)
    w_q[np.arange(sums.shape[0]), sums] = 1.0
    return torch.from_numpy(w_q)
