"""Minimal `problog.logic` terms: hashable, callable (`Term('digits')(Constant(3), Constant(4))`),
printing as ProbLog prints them (`p_10`, `digits(3,4)`)."""


class Term(object):
    def __init__(self, functor, *args, **_kw):
        self.functor, self.args = functor, tuple(args)

    def __call__(self, *args):
        return Term(self.functor, *args)

    def __str__(self):
        return str(self.functor) if not self.args else "%s(%s)" % (self.functor, ",".join(str(a) for a in self.args))

    __repr__ = __str__

    def __hash__(self):
        return hash((str(self.functor), self.args))

    def __eq__(self, other):
        return isinstance(other, Term) and str(self.functor) == str(other.functor) and self.args == other.args


class Constant(Term):
    def __init__(self, value):
        Term.__init__(self, value)
        self.value = value


class Var(Term):
    pass


class AnnotatedDisjunction(object):
    def __init__(self, heads, body):
        self.heads, self.body = heads, body
