"""vael_b200 — B200-native probabilistic-logic hot path of VAEL (EleMisi/VAEL).

Compiled logic programs (flat level-ordered CSR circuits) evaluated by hand-written
sm_100a CUDA kernels behind the C-ABI of include/vael_b200.h; this package is the
host-side mirror of the reference's Python interface for that path.
"""
from .circuit import Circuit  # noqa: F401
from .graph_semiring import GraphSemiring  # noqa: F401
from .logic import AnnotatedDisjunction, Constant, Term, Var  # noqa: F401

__all__ = ["Circuit", "GraphSemiring", "Term", "Constant", "Var", "AnnotatedDisjunction"]
