"""Host-side mirror of the reference's evaluator interface, utils/graph_semiring.py:5-60.

Callers keep writing `model.semiring = GraphSemiring(batch_size=..., device=model.device)`
(utils/mnist_utils/train.py:101,174; metrics.py:90) and the model keeps binding leaf weights
through `set_weights` (models/vael.py:242-243).  What changes is who consumes it: a compiled
program's `evaluate(semiring=...)` reads `semiring.weights` once, hands the [B,F] fact
matrix to the CUDA circuit kernel and never calls back per node.  The ten methods are kept
for code that uses the semiring directly; they act on whatever device the tensors are on.

`elementwise=False` (default) keeps the reference's batch-wide identity shortcuts
(`.any()`, :27-49) for those direct calls; the kernels always evaluate elementwise.
"""
from __future__ import annotations

import torch


class GraphSemiring:
    def __init__(self, batch_size=32, device=torch.device("cpu"), elementwise: bool = False):
        self.eps = 1e-12
        self.batch_size = batch_size
        self.device = torch.device(device) if not isinstance(device, torch.device) else device
        self.elementwise = elementwise
        self.weights = {}
        self._ident = {}

    def _identity(self, v: float) -> torch.Tensor:
        key = (v, self.batch_size)
        t = self._ident.get(key)
        if t is None:  # built on the device once, not on the host per call (reference :17-25)
            t = torch.full((self.batch_size,), v, device=self.device)
            self._ident[key] = t
        return t.clone()

    def one(self):
        return self._identity(1.0)

    def zero(self):
        return self._identity(0.0)

    def negate(self, a):
        return 1.0 - a

    def _around(self, a, c):
        return ((a >= c - self.eps) & (a <= c + self.eps)).any()

    def is_zero(self, a):
        return self._around(a, 0.0)

    def is_one(self, a):
        return self._around(a, 1.0)

    def plus(self, a, b):
        if not self.elementwise:
            if self.is_zero(b):
                return a
            if self.is_zero(a):
                return b
        return a + b

    def times(self, a, b):
        if not self.elementwise:
            if self.is_one(b):
                return a
            if self.is_one(a):
                return b
        return a * b

    def set_weights(self, weights):
        self.weights = weights

    def normalize(self, a, z):
        return a / z

    def value(self, a):
        return self.weights.get(a, a)

    # what a compiled program needs: the [B,F] fact matrix behind the bound weights
    def fact_matrix(self, fact_terms) -> torch.Tensor:
        cols = [self.value(t) for t in fact_terms]
        for t, c in zip(fact_terms, cols):
            if not isinstance(c, torch.Tensor):
                raise KeyError(f"no weight bound for {t}; call set_weights first (models/vael.py:242-243)")
        c0, F = cols[0], len(cols)
        # fast path: the strided views built by update_semiring_weights (facts_probs[:, i, j]) all
        # alias one contiguous [B,F] block -> hand that block to the kernel without a gather
        if all(c.dim() == 1 and c.stride(0) == F and c.shape == c0.shape and c.dtype == c0.dtype and
               c.storage_offset() == c0.storage_offset() + i and
               c.untyped_storage().data_ptr() == c0.untyped_storage().data_ptr() for i, c in enumerate(cols)):
            base = c0._base
            if base is not None and all(c._base is base for c in cols) and base.is_contiguous() and \
                    base.numel() == c0.shape[0] * F and base.storage_offset() == c0.storage_offset():
                return base.reshape(c0.shape[0], F)          # keeps the autograd link to facts_probs
            if not any(c.requires_grad for c in cols):
                return torch.as_strided(c0, (c0.shape[0], F), (F, 1), c0.storage_offset())
        return torch.stack(cols, dim=1)
