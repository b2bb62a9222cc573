"""A small ProbLog front end: terms, a parser for the program texts VAEL writes, and
an SLD engine that decides which query atoms hold in a total choice (world).

The reference hands its programs to the third-party `problog` package
(`from problog.logic import Term, Constant, Var, AnnotatedDisjunction`,
utils/mnist_utils/problog_model.py:1-2; `LogicFormula.create_from(...)`,
utils/mnist_utils/mnist_task_VAEL.py:214-216), which is neither vendored nor
installable here.  This module provides the same surface for the symbols the
reference touches (Term / Constant / Var / AnnotatedDisjunction with problog's
`str()` forms, hashable and callable) and enough of Prolog to ground the two
program families the reference ships:

* utils/mnist_utils/problog_model.py:5-54 + the rules at
  utils/mnist_utils/mnist_task_VAEL.py:206 (annotated disjunctions, `is/2`);
* utils/mario_utils/problog_model.py:3-62 (definite rules with arithmetic and a
  disjunctive body).
"""
from __future__ import annotations

import re
from dataclasses import dataclass, field
from typing import Dict, Iterator, List, Optional, Sequence, Tuple


# ---------------------------------------------------------------------------
# terms
# ---------------------------------------------------------------------------
class Term:
    """Compound term or atom.  `Term('digits')(Constant(3), Constant(4))` builds
    digits(3,4) exactly like problog.logic.Term.__call__ (models/vael.py:253-257)."""

    __slots__ = ("functor", "args", "probability", "_hash")

    def __init__(self, functor, *args, p=None):
        self.functor = functor
        self.args = tuple(args)
        self.probability = p if (p is None or isinstance(p, (Term, Var))) else _as_term(p)
        self._hash = None

    def __call__(self, *args, p=None):
        return Term(self.functor, *args, p=p if p is not None else self.probability)

    @property
    def arity(self):
        return len(self.args)

    @property
    def signature(self):
        return (self.functor, len(self.args))

    def with_probability(self, p=None):
        return Term(self.functor, *self.args, p=p)

    def _key(self):
        return (self.functor, self.args)

    def __eq__(self, other):
        return isinstance(other, Term) and not isinstance(other, Var) and self._key() == other._key()

    def __hash__(self):
        if self._hash is None:
            self._hash = hash(self._key())
        return self._hash

    def __repr__(self):
        body = str(self.functor)
        if self.args:
            body += "(" + ",".join(str(a) for a in self.args) + ")"
        if self.probability is not None:
            return f"{self.probability}::{body}"
        return body

    __str__ = __repr__

    def is_ground(self):
        return all(a.is_ground() for a in self.args)


class Constant(Term):
    """Number or quoted constant; str(Constant(3)) == '3'."""

    __slots__ = ()

    def __init__(self, value):
        super().__init__(value)

    @property
    def value(self):
        return self.functor

    def __int__(self):
        return int(self.functor)

    def __float__(self):
        return float(self.functor)


class Var(Term):
    __slots__ = ()

    def __init__(self, name):
        super().__init__(name)

    @property
    def name(self):
        return self.functor

    def __eq__(self, other):
        return isinstance(other, Var) and self.functor == other.functor

    def __hash__(self):
        return hash(("$var", self.functor))

    def is_ground(self):
        return False


def _as_term(v):
    if isinstance(v, Term):
        return v
    if isinstance(v, (int, float)):
        return Constant(v)
    s = str(v)
    if re.fullmatch(r"-?\d+", s):
        return Constant(int(s))
    if re.fullmatch(r"-?\d*\.\d+(e-?\d+)?", s):
        return Constant(float(s))
    return Term(s)


class AnnotatedDisjunction:
    """p1::h1; p2::h2; ... [:- body].  str() matches problog's rendering, which the
    reference concatenates into program text (utils/mnist_utils/problog_model.py:19-20)."""

    def __init__(self, heads: Sequence[Term], body: Optional[Term]):
        self.heads = list(heads)
        self.body = body

    def __str__(self):
        s = "; ".join(str(h) for h in self.heads)
        if self.body is not None:
            s += " :- " + str(self.body)
        return s

    __repr__ = __str__


@dataclass
class Clause:
    head: Term
    body: Optional[Term]  # None for facts


@dataclass
class Program:
    """Parsed program: annotated disjunctions (probabilistic facts are 1-head ADs),
    definite clauses, query and evidence directives, in source order."""

    ads: List[AnnotatedDisjunction] = field(default_factory=list)
    clauses: List[Clause] = field(default_factory=list)
    queries: List[Term] = field(default_factory=list)
    evidence: List[Tuple[Term, bool]] = field(default_factory=list)

    def logic_signature(self) -> str:
        """Text identity of the facts+rules part (what a family of programs shares)."""
        return "\n".join([str(a) for a in self.ads] + [f"{c.head} :- {c.body}" for c in self.clauses])


# ---------------------------------------------------------------------------
# parser (operator precedence, the Prolog subset the reference's programs use)
# ---------------------------------------------------------------------------
_TOKEN = re.compile(
    r"\s*(?:(%[^\n]*)|(\d+\.\d+(?:[eE][-+]?\d+)?|\d+)|([A-Z_][A-Za-z0-9_]*)|([a-z][A-Za-z0-9_]*)|"
    r"('(?:[^'\\]|\\.)*')|(:-|::|=:=|=\\=|=<|>=|\\=|\\\+|//|[-+*/<>=;,().]))"
)

_INFIX = {  # op -> (precedence, right-assoc)
    ":-": (1200, False), ";": (1100, True), ",": (1000, True),
    "is": (700, False), "<": (700, False), ">": (700, False), "=<": (700, False), ">=": (700, False),
    "=:=": (700, False), "=\\=": (700, False), "=": (700, False), "\\=": (700, False),
    "+": (500, False), "-": (500, False), "*": (400, False), "/": (400, False), "//": (400, False),
    "mod": (400, False),
}


class ParseError(ValueError):
    pass


def _tokenize(text: str) -> List[Tuple[str, str]]:
    toks, pos = [], 0
    text = text.rstrip()
    while pos < len(text):
        m = _TOKEN.match(text, pos)
        if not m or m.end() == pos:
            if text[pos:].strip() == "":
                break
            raise ParseError(f"cannot tokenize at: {text[pos:pos + 30]!r}")
        pos = m.end()
        if m.group(1):
            continue
        if m.group(2):
            toks.append(("num", m.group(2)))
        elif m.group(3):
            toks.append(("var", m.group(3)))
        elif m.group(4):
            toks.append(("atom", m.group(4)))
        elif m.group(5):
            toks.append(("atom", m.group(5)[1:-1]))
        else:
            toks.append(("punct", m.group(6)))
    return toks


class _Parser:
    def __init__(self, toks):
        self.toks, self.i = toks, 0

    def peek(self):
        return self.toks[self.i] if self.i < len(self.toks) else ("eof", "")

    def next(self):
        t = self.peek()
        self.i += 1
        return t

    def expect(self, val):
        t = self.next()
        if t[1] != val:
            raise ParseError(f"expected {val!r}, got {t[1]!r}")

    def _is_end(self):
        # a '.' that terminates a clause: followed by eof or not by a digit glued to it
        return self.peek() == ("punct", ".")

    def parse_expr(self, maxprec: int) -> Term:
        left = self.parse_primary()
        while True:
            kind, val = self.peek()
            op = val if (kind == "punct" or (kind == "atom" and val in ("is", "mod"))) else None
            if op == "::" and maxprec >= 1050:  # probability annotation: tighter than ';' / ':-', looser than ','
                self.next()
                atom = self.parse_expr(999)
                left = atom.with_probability(left)
                continue
            if op not in _INFIX:
                return left
            prec, right = _INFIX[op]
            if prec > maxprec:
                return left
            self.next()
            rhs = self.parse_expr(prec if right else prec - 1)
            left = Term(op, left, rhs)

    def parse_primary(self) -> Term:
        kind, val = self.next()
        if kind == "num":
            return Constant(float(val) if "." in val else int(val))
        if kind == "var":
            return Var(val)
        if kind == "atom":
            if self.peek() == ("punct", "("):
                self.next()
                args = [self.parse_expr(999)]
                while self.peek() == ("punct", ","):
                    self.next()
                    args.append(self.parse_expr(999))
                self.expect(")")
                return Term(val, *args)
            return Term(val)
        if (kind, val) == ("punct", "("):
            e = self.parse_expr(1200)
            self.expect(")")
            return e
        if (kind, val) == ("punct", "-"):
            e = self.parse_expr(200)
            if isinstance(e, Constant):
                return Constant(-e.value)
            return Term("-", Constant(0), e)
        if (kind, val) == ("punct", "\\+"):
            return Term("\\+", self.parse_expr(900))
        raise ParseError(f"unexpected token {val!r}")


def _flatten(op: str, t: Term) -> List[Term]:
    out = []
    while isinstance(t, Term) and t.functor == op and t.arity == 2:
        out.append(t.args[0])
        t = t.args[1]
    out.append(t)
    return out


def parse_program(text: str) -> Program:
    """Parse program text into ADs / clauses / query and evidence directives."""
    prog = Program()
    p = _Parser(_tokenize(text))
    while p.peek()[0] != "eof":
        stmt = p.parse_expr(1200)
        p.expect(".")
        body = None
        if stmt.functor == ":-" and stmt.arity == 2:
            stmt, body = stmt.args
        heads = _flatten(";", stmt)
        if any(h.probability is not None for h in heads):
            prog.ads.append(AnnotatedDisjunction(heads, body))
            continue
        if len(heads) != 1:
            raise ParseError("disjunctive heads need probability annotations")
        head = heads[0]
        if head.functor == "query" and head.arity == 1 and body is None:
            prog.queries.append(head.args[0])
        elif head.functor == "evidence" and head.arity in (1, 2) and body is None:
            val = True
            if head.arity == 2:
                val = str(head.args[1]) == "true"
            prog.evidence.append((head.args[0], val))
        else:
            prog.clauses.append(Clause(head, body))
    return prog


# ---------------------------------------------------------------------------
# SLD engine over a fixed set of true facts
# ---------------------------------------------------------------------------
Subst = Dict[str, Term]


def _walk(t: Term, s: Subst) -> Term:
    while isinstance(t, Var) and t.name in s:
        t = s[t.name]
    return t


def substitute(t: Term, s: Subst) -> Term:
    t = _walk(t, s)
    if isinstance(t, (Var, Constant)) or not t.args:
        return t
    return Term(t.functor, *[substitute(a, s) for a in t.args], p=t.probability)


def unify(a: Term, b: Term, s: Subst) -> Optional[Subst]:
    a, b = _walk(a, s), _walk(b, s)
    if isinstance(a, Var):
        if isinstance(b, Var) and a.name == b.name:
            return s
        s2 = dict(s)
        s2[a.name] = b
        return s2
    if isinstance(b, Var):
        s2 = dict(s)
        s2[b.name] = a
        return s2
    if a.functor != b.functor or a.arity != b.arity:
        return None
    if isinstance(a, Constant) != isinstance(b, Constant):
        return None
    for x, y in zip(a.args, b.args):
        s = unify(x, y, s)
        if s is None:
            return None
    return s


def _rename(t: Term, suffix: str) -> Term:
    if isinstance(t, Var):
        return Var(t.name + suffix)
    if isinstance(t, Constant) or not t.args:
        return t
    return Term(t.functor, *[_rename(a, suffix) for a in t.args], p=t.probability)


def _arith(t: Term, s: Subst):
    t = _walk(t, s)
    if isinstance(t, Constant):
        return t.value
    if isinstance(t, Var):
        raise ValueError("arithmetic on an unbound variable")
    if t.arity == 2:
        x, y = _arith(t.args[0], s), _arith(t.args[1], s)
        f = t.functor
        if f == "+": return x + y
        if f == "-": return x - y
        if f == "*": return x * y
        if f == "/": return x / y
        if f == "//": return x // y
        if f == "mod": return x % y
    raise ValueError(f"cannot evaluate {t}")


_COMPARE = {
    "<": lambda x, y: x < y, ">": lambda x, y: x > y, "=<": lambda x, y: x <= y, ">=": lambda x, y: x >= y,
    "=:=": lambda x, y: x == y, "=\\=": lambda x, y: x != y,
}


class Engine:
    """Depth-bounded SLD resolution.  `facts` maps a signature to the (atom, tag) pairs true
    in the current world (the heads selected by the total choice; tag identifies the
    probabilistic choice); `clauses` are the definite rules."""

    def __init__(self, clauses: Sequence[Clause], max_depth: int = 64):
        self.by_sig: Dict[Tuple[str, int], List[Clause]] = {}
        for c in clauses:
            self.by_sig.setdefault(c.head.signature, []).append(c)
        self.max_depth = max_depth
        self._fresh = 0
        self.trace = None  # set to a list to record (tag, ground fact) for every probabilistic fact used

    def solve(self, goal: Term, facts: Dict[Tuple[str, int], List[Term]], s: Optional[Subst] = None,
              depth: int = 0) -> Iterator[Subst]:
        s = {} if s is None else s
        if depth > self.max_depth:
            return
        goal = _walk(goal, s)
        f, n = goal.functor, goal.arity
        if f == "," and n == 2:
            for s1 in self.solve(goal.args[0], facts, s, depth):
                yield from self.solve(goal.args[1], facts, s1, depth)
            return
        if f == ";" and n == 2:
            yield from self.solve(goal.args[0], facts, s, depth)
            yield from self.solve(goal.args[1], facts, s, depth)
            return
        if f == "true" and n == 0:
            yield s
            return
        if f in ("fail", "false") and n == 0:
            return
        if f == "\\+" and n == 1:
            if next(self.solve(goal.args[0], facts, s, depth + 1), None) is None:
                yield s
            return
        if f == "is" and n == 2:
            s2 = unify(goal.args[0], Constant(_arith(goal.args[1], s)), s)
            if s2 is not None:
                yield s2
            return
        if f in _COMPARE and n == 2:
            if _COMPARE[f](_arith(goal.args[0], s), _arith(goal.args[1], s)):
                yield s
            return
        if f == "=" and n == 2:
            s2 = unify(goal.args[0], goal.args[1], s)
            if s2 is not None:
                yield s2
            return
        if f == "\\=" and n == 2:
            if unify(goal.args[0], goal.args[1], s) is None:
                yield s
            return
        for fact, tag in facts.get((f, n), ()):
            if not fact.is_ground():  # non-ground probabilistic fact: rename apart like a clause
                self._fresh += 1
                fact = _rename(fact, f"_{self._fresh}")
            s2 = unify(goal, fact, s)
            if s2 is not None:
                if self.trace is not None:
                    self.trace.append((tag, substitute(fact, s2)))
                yield s2
        for c in self.by_sig.get((f, n), ()):
            self._fresh += 1
            suffix = f"_{self._fresh}"
            s2 = unify(goal, _rename(c.head, suffix), s)
            if s2 is None:
                continue
            if c.body is None:
                yield s2
            else:
                yield from self.solve(_rename(c.body, suffix), facts, s2, depth + 1)

    def holds(self, goal: Term, facts) -> bool:
        return next(self.solve(goal, facts), None) is not None

    def solutions(self, goal: Term, facts) -> List[Term]:
        """Distinct ground instances of `goal`, in the order SLD finds them."""
        seen, out = set(), []
        for s in self.solve(goal, facts):
            g = substitute(goal, s)
            if g not in seen:
                seen.add(g)
                out.append(g)
        return out
