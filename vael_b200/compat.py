"""Import shim that lets the REFERENCE's own code run on this path unchanged.

The reference builds its compiled programs with three ProbLog calls
(`SDD.create_from(LogicDAG.create_from(LogicFormula.create_from(program)))`,
utils/mnist_utils/mnist_task_VAEL.py:214-216) and evaluates them with
`sdd.evaluate(semiring=GraphSemiring(...))` (models/vael.py:109,130).  `install()` registers stand-ins for
the `problog` modules it imports — `problog.logic` (Term, Constant, Var, AnnotatedDisjunction),
`problog.evaluator.Semiring`, `problog.program.PrologString`, `problog.formula.{LogicFormula,LogicDAG}`,
`problog.sdd_formula.SDD`, `problog.ddnnf_formula.DDNNF` — whose `create_from` chain ends in this repo's
host compiler and whose result offers the same `.evaluate(semiring=...)`, backed by the CUDA circuit
kernels.  The semiring handed to `evaluate` may be the reference's own unmodified
utils/graph_semiring.py:GraphSemiring: only its `weights` / `value()` are read (once per call), none of
its per-node callbacks is invoked.

    import vael_b200.compat; vael_b200.compat.install()     # before importing the reference's modules
"""
from __future__ import annotations

import sys
import types

from . import logic as _logic


class Semiring(object):
    """problog.evaluator.Semiring defaults the reference's GraphSemiring relies on."""

    def pos_value(self, a, key=None):
        return self.value(a)

    def neg_value(self, a, key=None):
        return self.negate(self.value(a))

    def ad_negate(self, pos, neg):
        return self.one()

    def ad_complement(self, ws, key=None):
        s = self.zero()
        for w in ws:
            s = self.plus(s, w)
        return self.negate(s)

    def result(self, a, formula=None):
        return a

    def is_dsp(self):
        return False

    def is_nsp(self):
        return False

    def in_domain(self, a):
        return True


class PrologString(object):
    def __init__(self, string, *_, **__):
        self.text = str(string)

    def __str__(self):
        return self.text


def _text_of(x) -> str:
    return x.text if hasattr(x, "text") else str(x)


class _PassThrough(object):
    """LogicFormula / LogicDAG: grounding and cycle breaking happen inside the compiler; keep the text."""

    def __init__(self, text):
        self.text = text

    @classmethod
    def create_from(cls, program, **_):
        return cls(_text_of(program))


class LogicFormula(_PassThrough):
    pass


class LogicDAG(_PassThrough):
    pass


def compile_program(program):
    """ProbLog program text -> CompiledProgram (vael_b200.program) with `.evaluate(semiring=...)`.
    The first non-ground query (e.g. `query(digits(X,Y)).`) defines the world outputs."""
    from .compiler import compile_family
    from .program import CompiledProgramSet
    text = _text_of(program)
    parsed = _logic.parse_program(text)
    wq = next((str(q) for q in parsed.queries if not q.is_ground()), None)
    fam = compile_family({0: text}, world_query=wq)
    return CompiledProgramSet(fam).view(0)


class SDD(object):
    @classmethod
    def create_from(cls, formula, **_):
        return compile_program(formula)


class DDNNF(SDD):
    pass


def install(force: bool = False) -> None:
    """Register the stand-ins under the `problog.*` module names (no-op if a real ProbLog is importable,
    unless force=True)."""
    if not force:
        try:
            import problog  # noqa: F401
            if not getattr(problog, "__vael_b200_shim__", False):
                return
        except ImportError:
            pass
    pkg = types.ModuleType("problog")
    pkg.__vael_b200_shim__ = True
    pkg.__path__ = []
    mods = {
        "problog": pkg,
        "problog.logic": _logic,
        "problog.evaluator": types.ModuleType("problog.evaluator"),
        "problog.program": types.ModuleType("problog.program"),
        "problog.formula": types.ModuleType("problog.formula"),
        "problog.sdd_formula": types.ModuleType("problog.sdd_formula"),
        "problog.ddnnf_formula": types.ModuleType("problog.ddnnf_formula"),
    }
    mods["problog.evaluator"].Semiring = Semiring
    mods["problog.program"].PrologString = PrologString
    mods["problog.formula"].LogicFormula = LogicFormula
    mods["problog.formula"].LogicDAG = LogicDAG
    mods["problog.sdd_formula"].SDD = SDD
    mods["problog.ddnnf_formula"].DDNNF = DDNNF
    for name, m in mods.items():
        sys.modules[name] = m
        if "." in name:
            setattr(pkg, name.split(".", 1)[1], m)
