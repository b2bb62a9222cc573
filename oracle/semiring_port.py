"""Port of the reference's batched semiring — utils/graph_semiring.py:5-60.

Elements are [B] float32 torch vectors.  The reference's identity tests reduce over the
WHOLE batch with `.any()` (:27-33): `plus` returns one operand unchanged for every row as
soon as any row of the other operand is ~0 (:35-41), `times` likewise for ~1 (:43-49).
That quirk is kept here on purpose: this class defines "the reference's graph_semiring
path"; vael_b200's kernels are elementwise ("Parity domain").
"""
import torch


class SemiringPort:
    EPS = 1e-12  # graph_semiring.py:9

    def __init__(self, batch_size=32, device="cpu"):
        self.batch_size, self.device, self.eps = batch_size, torch.device(device), self.EPS
        self.weights = {}

    # identities are rebuilt on every call, as in the reference (:17-25)
    def one(self):
        return torch.ones(self.batch_size).to(self.device)

    def zero(self):
        return torch.zeros(self.batch_size).to(self.device)

    def _near(self, a, c):  # (:27-33) closed interval test, reduced with any()
        return bool(((a >= c - self.eps) & (a <= c + self.eps)).any())

    def is_zero(self, a):
        return self._near(a, 0.0)

    def is_one(self, a):
        return self._near(a, 1.0)

    def plus(self, a, b):  # (:35-41) b is tested first
        if self.is_zero(b):
            return a
        return b if self.is_zero(a) else a + b

    def times(self, a, b):  # (:43-49) b is tested first
        if self.is_one(b):
            return a
        return b if self.is_one(a) else a * b

    def negate(self, a):  # (:13-15)
        return self.one() - a

    def normalize(self, a, z):  # (:54-55)
        return a / z

    def set_weights(self, weights):  # (:51-52)
        self.weights = weights

    def value(self, a):  # (:57-60) symbolic weight -> tensor, unknown keys pass through
        return self.weights.get(a, a)

    # problog.evaluator.Semiring defaults the reference inherits
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


class ElementwiseSemiring(SemiringPort):
    """The intended semantics: same ops without the batch-wide shortcuts (what the CUDA
    kernels implement).  dtype follows the inputs, so it doubles as the fp64 oracle."""

    def __init__(self, batch_size=32, device="cpu", dtype=torch.float32):
        super().__init__(batch_size, device)
        self.dtype = dtype

    def one(self):
        return torch.ones(self.batch_size, dtype=self.dtype, device=self.device)

    def zero(self):
        return torch.zeros(self.batch_size, dtype=self.dtype, device=self.device)

    def plus(self, a, b):
        return a + b

    def times(self, a, b):
        return a * b
