"""Shared synthetic inputs (the domain compute_facts_probability emits)."""
import numpy as np


def synthetic_facts(B, groups=(10, 10), seed=0, scale=3.0, eps=1e-5):
    rng = np.random.default_rng(seed)
    cols = []
    for n in groups:
        x = scale * rng.standard_normal((B, n)).astype(np.float32)
        x = x - x.max(axis=1, keepdims=True)
        p = np.exp(x) / np.exp(x).sum(axis=1, keepdims=True)
        p = (p + np.float32(eps)).astype(np.float32)
        cols.append((p / p.sum(axis=1, keepdims=True)).astype(np.float32))
    return np.concatenate(cols, axis=1)


def rel_err(a, b, floor=0.0):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.max(np.abs(a - b) / (np.abs(b) + floor)) if a.size else 0.0
