"""torch front end of the C-ABI: device circuits and autograd functions.

PyTorch is plumbing here (device memory, streams, autograd bookkeeping); every
computation below is a kernel of libvael_b200.so.  CPU tensors are rejected: there is
no fallback path.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Tuple, Union

import torch

from . import _lib
from .circuit import FLAG_CHECK_QUERY, FLAG_EVIDENCE, FLAG_FORCE_GENERIC, FLAG_FORCE_JIT, FLAG_NORMALIZE, Circuit


# NVTX ranges around every library call (tracing), for `ncu --nvtx --nvtx-include "vael/..."` and
# timeline tools.  Off by default (two host calls per op matter in the launch-bound B = 32 eager step);
# VAEL_NVTX=1 switches them on.
_NVTX = os.environ.get("VAEL_NVTX", "0") not in ("", "0")


class _nvtx:
    __slots__ = ("name",)

    def __init__(self, name: str):
        self.name = name

    def __enter__(self):
        if _NVTX:
            torch.cuda.nvtx.range_push("vael/" + self.name)

    def __exit__(self, *exc):
        if _NVTX:
            torch.cuda.nvtx.range_pop()
        return False


def _stream_ptr() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t: Optional[torch.Tensor]) -> C.c_void_p:
    return C.c_void_p(0 if t is None else t.data_ptr())


def _require_cuda_f32(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError(f"vael_b200: `{name}` must be a CUDA tensor (no CPU fallback)")
    if t.dtype != torch.float32:
        raise RuntimeError(f"vael_b200: `{name}` must be float32, got {t.dtype}")
    return t.contiguous()


class DeviceCircuit:
    """A Circuit uploaded to one GPU (vael_circuit_create).  Immutable afterwards."""

    def __init__(self, circuit: Circuit, device: Union[str, torch.device, int, None] = None):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise RuntimeError("vael_b200: no CUDA device (the circuit kernels have no CPU fallback)")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("vael_b200: circuits live in HBM; pass a cuda device")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.circuit = circuit
        handle = C.c_void_p()
        with torch.cuda.device(self.device):
            desc = _lib.make_desc(circuit)
            _lib.check(self.lib.vael_circuit_create(C.byref(desc), C.byref(handle)))
        self.handle = handle
        self.fast_kind = int(self.lib.vael_circuit_fast_kind(handle))
        self.hash = int(self.lib.vael_circuit_hash(handle))

    def jit_status(self):
        """(True, log) when the circuit-specialised kernels exist (compiling them now if need be), else (False, why)."""
        buf = C.create_string_buffer(1 << 16)
        with torch.cuda.device(self.device):
            ok = bool(self.lib.vael_circuit_jit_status(self.handle, buf, len(buf)))
        return ok, buf.value.decode(errors="replace")

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.vael_circuit_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    def __deepcopy__(self, memo):  # immutable and device-bound: share, never duplicate
        return self

    def __reduce__(self):
        raise TypeError("DeviceCircuit holds a device handle and cannot be pickled; pickle the Circuit")

    # sizes
    @property
    def n_facts(self): return self.circuit.n_facts
    @property
    def n_worlds(self): return self.circuit.n_worlds
    @property
    def n_queries(self): return self.circuit.n_queries


def _query_args(dc: DeviceCircuit, query, B: int):
    """-> (qidx int32 tensor or None, scalar int)"""
    if query is None:
        return None, -1
    if isinstance(query, torch.Tensor) and query.dim() > 0:
        q = query.reshape(-1)
        if q.shape[0] != B:
            raise RuntimeError("vael_b200: per-row query index must have one entry per batch row")
        if not q.is_cuda:
            raise RuntimeError("vael_b200: per-row query index must be a CUDA tensor")
        return q.to(torch.int32).contiguous(), -1
    qs = int(query)
    if not 0 <= qs < dc.n_queries:
        raise KeyError(f"query {qs} is not in the compiled program family (0..{dc.n_queries - 1})")
    return None, qs


class _CircuitFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, facts, dc: DeviceCircuit, qidx, qscalar: int, flags: int, want_worlds: bool, want_query: bool):
        B = facts.shape[0]
        if facts.numel() != B * dc.n_facts:
            raise RuntimeError(f"vael_b200: facts has {facts.numel() // max(B, 1)} columns, the circuit has {dc.n_facts}")
        f2 = _require_cuda_f32(facts, "facts").reshape(B, dc.n_facts)
        worlds = torch.empty((B, dc.n_worlds), device=f2.device, dtype=torch.float32) if want_worlds else None
        query = torch.empty((B, 1), device=f2.device, dtype=torch.float32) if want_query else None
        with torch.cuda.device(f2.device), _nvtx("circuit_forward"):
            _lib.check(dc.lib.vael_circuit_forward(dc.handle, _p(f2), _p(qidx), qscalar, B, _p(worlds), _p(query),
                                                   flags, _stream_ptr()))
        ctx.dc, ctx.qscalar, ctx.flags, ctx.shape = dc, qscalar, flags, facts.shape
        ctx.save_for_backward(f2, qidx)
        outs = (worlds if want_worlds else f2.new_empty(0), query if want_query else f2.new_empty(0))
        return outs

    @staticmethod
    def backward(ctx, g_worlds, g_query):
        f2, qidx = ctx.saved_tensors
        dc, B = ctx.dc, f2.shape[0]
        gw = g_worlds.contiguous().float() if (g_worlds is not None and g_worlds.numel()) else None
        gq = g_query.contiguous().float().reshape(-1) if (g_query is not None and g_query.numel()) else None
        gf = torch.empty_like(f2)
        with torch.cuda.device(f2.device), _nvtx("circuit_backward"):
            _lib.check(dc.lib.vael_circuit_backward(dc.handle, _p(f2), _p(qidx), ctx.qscalar, B, _p(gw), _p(gq),
                                                    _p(gf), ctx.flags & ~(FLAG_EVIDENCE | FLAG_CHECK_QUERY), _stream_ptr()))
        return gf.reshape(ctx.shape), None, None, None, None, None, None


def circuit_forward(dc: DeviceCircuit, facts: torch.Tensor, query=None, *, normalize: bool = False,
                    evidence: bool = False, force_generic: bool = False, want_worlds: bool = True,
                    check_query: bool = False, force_jit: bool = False
                    ) -> Tuple[Optional[torch.Tensor], Optional[torch.Tensor]]:
    """worlds [B,W], query [B,1] of the compiled program for the facts [B,F] (or [B,k,n]).
    `query`: None, an int (one query for the batch, the reference's behaviour,
    models/vael.py:212-213) or an integer tensor [B] (one per row; -1 = no query for that row, value 0).
    `check_query`: validate a per-row index tensor on the device first and raise on an out-of-range entry
    (synchronises the stream: not for CUDA-graph capture); unchecked, such a row's query is 0.
    `force_generic` / `force_jit`: run the level-by-level CSR interpreter / the circuit-specialised straight-line kernel
    (NVRTC) instead of whatever the library would pick for this circuit."""
    B = facts.shape[0]
    qidx, qs = _query_args(dc, query, B)
    flags = (FLAG_NORMALIZE if normalize else 0) | (FLAG_EVIDENCE if evidence else 0) | \
            (FLAG_FORCE_GENERIC if force_generic else 0) | (FLAG_CHECK_QUERY if check_query else 0) | \
            (FLAG_FORCE_JIT if force_jit else 0)
    if evidence and query is None:
        raise RuntimeError("vael_b200: evidence mode needs the evidence (query) index")
    want_query = query is not None
    if evidence and facts.requires_grad and torch.is_grad_enabled():
        raise RuntimeError("vael_b200: evidence mode is evaluation-only (the reference runs it under no_grad)")
    worlds, q = _CircuitFunction.apply(facts, dc, qidx, qs, flags, want_worlds, want_query)
    return (worlds if want_worlds else None), (q if want_query else None)


def circuit_queries_all(dc: DeviceCircuit, facts: torch.Tensor, *, normalize: bool = False,
                        force_generic: bool = False, force_jit: bool = False) -> torch.Tensor:
    """[B,Q] values of every query node (the worlds @ w_q contraction of
    utils/mnist_utils/metrics.py:71-75).  Evaluation-only."""
    f2 = _require_cuda_f32(facts.detach(), "facts").reshape(facts.shape[0], dc.n_facts)
    out = torch.empty((f2.shape[0], dc.n_queries), device=f2.device, dtype=torch.float32)
    with torch.cuda.device(f2.device), _nvtx("queries_all"):
        _lib.check(dc.lib.vael_circuit_queries_all(dc.handle, _p(f2), f2.shape[0], _p(out),
                                                   (FLAG_NORMALIZE if normalize else 0) |
                                                   (FLAG_FORCE_GENERIC if force_generic else 0) |
                                                   (FLAG_FORCE_JIT if force_jit else 0), _stream_ptr()))
    return out


# ---------------------------------------------------------------------------
class _FactsFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, n_groups: int, eps: float, clip_hi: float):
        x = _require_cuda_f32(logits, "logits")
        B = x.shape[0]
        group = x.shape[1] // n_groups
        out = torch.empty_like(x)
        lib = _lib.load()
        with torch.cuda.device(x.device), _nvtx("facts_forward"):
            if clip_hi > 0.0:
                _lib.check(lib.vael_facts_clip_forward(_p(x), B, n_groups, group, eps, clip_hi, _p(out), _stream_ptr()))
            else:
                _lib.check(lib.vael_facts_forward(_p(x), B, n_groups, group, eps, _p(out), _stream_ptr()))
        ctx.save_for_backward(x)
        ctx.n_groups, ctx.group, ctx.eps, ctx.clip_hi = n_groups, group, eps, clip_hi
        return out

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        g = g.contiguous().float()
        gx = torch.empty_like(x)
        lib = _lib.load()
        with torch.cuda.device(x.device), _nvtx("facts_backward"):
            if ctx.clip_hi > 0.0:
                _lib.check(lib.vael_facts_clip_backward(_p(x), _p(g), x.shape[0], ctx.n_groups, ctx.group, ctx.eps,
                                                        ctx.clip_hi, _p(gx), _stream_ptr()))
            else:
                _lib.check(lib.vael_facts_backward(_p(x), _p(g), x.shape[0], ctx.n_groups, ctx.group, ctx.eps, _p(gx),
                                                   _stream_ptr()))
        return gx, None, None, None


def facts_from_logits(logits: torch.Tensor, n_groups: int, eps: float = 1e-5) -> torch.Tensor:
    """[B, n_groups*n] logits -> per group softmax, +eps, / detached sum
    (MNISTPairsVAELModel.compute_facts_probability, models/vael.py:159-177)."""
    if logits.shape[1] % n_groups:
        raise RuntimeError("vael_b200: logits width is not a multiple of n_groups")
    return _FactsFunction.apply(logits, n_groups, float(eps), 0.0)


def facts_from_logits_clipped(logits: torch.Tensor, n_groups: int, lo: float = 1e-5, hi: float = 0.9999) -> torch.Tensor:
    """[B, n_groups*n] logits -> per group softmax, clamped to [lo, hi]
    (MarioVAELModel.forward, models/vael.py:319-322: Softmax + torch.clip)."""
    if logits.shape[1] % n_groups:
        raise RuntimeError("vael_b200: logits width is not a multiple of n_groups")
    if not (0.0 <= lo <= hi and hi > 0.0):
        raise RuntimeError("vael_b200: bad clip range")
    return _FactsFunction.apply(logits, n_groups, float(lo), float(hi))


# ---------------------------------------------------------------------------
class _GumbelFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, scores, herbrand, noise, seed: int, tau: float, is_log: bool):
        s = _require_cuda_f32(scores, "scores")
        hb = _require_cuda_f32(herbrand, "herbrand")
        B, W = s.shape
        if hb.shape[0] != W:
            raise RuntimeError("vael_b200: herbrand base must have one row per world")
        H = hb.shape[1]
        nz = _require_cuda_f32(noise, "noise") if noise is not None else None
        y_soft = torch.empty_like(s)
        idx = torch.empty((B,), device=s.device, dtype=torch.int32)
        world_h = torch.empty((B, H), device=s.device, dtype=torch.float32)
        lib = _lib.load()
        with torch.cuda.device(s.device), _nvtx("gumbel_forward"):
            _lib.check(lib.vael_gumbel_world_forward(_p(s), _p(nz), seed, B, W, int(is_log), tau, _p(hb), H,
                                                     _p(y_soft), _p(idx), _p(world_h), _stream_ptr()))
        ctx.save_for_backward(s, y_soft, hb)
        ctx.tau, ctx.is_log = tau, is_log
        ctx.mark_non_differentiable(idx, y_soft)
        return world_h, idx, y_soft

    @staticmethod
    def backward(ctx, g_world_h, _gi, _gy):
        s, y_soft, hb = ctx.saved_tensors
        g = g_world_h.contiguous().float()
        gs = torch.empty_like(s)
        lib = _lib.load()
        with torch.cuda.device(s.device), _nvtx("gumbel_backward"):
            _lib.check(lib.vael_gumbel_world_backward(_p(s), _p(y_soft), _p(g), s.shape[0], s.shape[1],
                                                      int(ctx.is_log), ctx.tau, _p(hb), hb.shape[1], _p(gs),
                                                      _stream_ptr()))
        return gs, None, None, None, None, None


def gumbel_world(scores: torch.Tensor, herbrand: torch.Tensor, *, tau: float = 1.0, is_log: bool = False,
                 noise: Optional[torch.Tensor] = None, seed: Optional[int] = None):
    """Sample one world per row with the Gumbel-max trick and return its Herbrand vector with
    the straight-through gradient of F.gumbel_softmax(hard=True) (models/vael.py:61-65).
    -> (world_h [B,H], world_idx [B] int32, y_soft [B,W])"""
    if noise is None and seed is None:
        seed = int(torch.randint(0, 2 ** 62, (1,)).item())  # drawn from torch's CPU generator: reproducible under manual_seed
    return _GumbelFunction.apply(scores, herbrand, noise, int(seed or 0), float(tau), bool(is_log))


# ---------------------------------------------------------------------------
class _FusedLogicFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, dc: DeviceCircuit, qidx, qscalar: int, eps: float, tau: float, noise, seed: int, seed_dev,
                want_y_soft: bool):
        x = _require_cuda_f32(logits, "logits")
        B, F = x.shape
        if F != dc.n_facts:
            raise RuntimeError(f"vael_b200: logits has {F} columns, the circuit has {dc.n_facts} facts")
        W = dc.n_worlds
        nz = _require_cuda_f32(noise, "noise") if noise is not None else None
        if nz is not None and tuple(nz.shape) != (B, W):
            raise RuntimeError("vael_b200: injected Gumbel noise must be [B, n_worlds]")
        facts = torch.empty_like(x)
        query = torch.empty((B, 1), device=x.device, dtype=torch.float32)
        idx = torch.empty((B,), device=x.device, dtype=torch.int32)
        world_h = torch.empty_like(x)
        y_soft = torch.empty((B, W), device=x.device, dtype=torch.float32) if want_y_soft else None
        has_q = qidx is not None or qscalar >= 0
        with torch.cuda.device(x.device), _nvtx("logic_train_forward"):
            _lib.check(dc.lib.vael_logic_train_forward(dc.handle, _p(x), _p(qidx), qscalar, B, eps, tau, _p(nz), seed,
                                                       _p(seed_dev), _p(facts), _p(query) if has_q else None, _p(idx),
                                                       _p(world_h), _p(y_soft), _stream_ptr()))
        if not has_q:
            query.zero_()
        ctx.dc, ctx.qscalar, ctx.eps, ctx.tau, ctx.seed, ctx.has_q = dc, qscalar, eps, tau, seed, has_q
        ctx.save_for_backward(x, qidx, nz, seed_dev)
        ctx.mark_non_differentiable(idx)
        if y_soft is None:
            y_soft = x.new_empty(0)
        ctx.mark_non_differentiable(y_soft)
        return facts, query, world_h, idx, y_soft

    @staticmethod
    def backward(ctx, g_facts, g_query, g_world_h, _gi, _gy):
        x, qidx, nz, seed_dev = ctx.saved_tensors
        dc = ctx.dc

        def c(g):
            return g.contiguous().float() if g is not None else None

        gf, gq, gh = c(g_facts), c(g_query), c(g_world_h)
        gq = gq.reshape(-1) if (gq is not None and ctx.has_q) else None
        gx = torch.empty_like(x)
        with torch.cuda.device(x.device), _nvtx("logic_train_backward"):
            _lib.check(dc.lib.vael_logic_train_backward(dc.handle, _p(x), _p(qidx), ctx.qscalar, x.shape[0], ctx.eps,
                                                        ctx.tau, _p(nz), ctx.seed, _p(seed_dev), _p(gh), _p(gq), _p(gf),
                                                        _p(gx), _stream_ptr()))
        return gx, None, None, None, None, None, None, None, None, None


def fused_logic_supported(dc: DeviceCircuit) -> bool:
    return bool(dc.lib.vael_logic_fused_supported(dc.handle))


def logic_train(dc: DeviceCircuit, logits: torch.Tensor, query=None, *, eps: float = 1e-5, tau: float = 1.0,
                noise: Optional[torch.Tensor] = None, seed: Optional[int] = None,
                seed_dev: Optional[torch.Tensor] = None, want_y_soft: bool = False):
    """The whole training-mode logic path of VAELModel.forward (models/vael.py:55-65) in one kernel, its adjoint
    in another: MLP logits [B,2N] -> (facts [B,2N], query [B,1], world_h [B,2N], world_idx [B] int32, y_soft [B,W]
    or empty).  `query` as in circuit_forward.  Gumbel noise: `noise` [B,W] injected (tests), else drawn in the
    kernel from Philox keyed by `seed` (host int; default: one draw from torch's CPU generator) or by `seed_dev`
    (a one-element int64 CUDA tensor — inside a CUDA graph fill it with a captured torch.randint so that every
    replay gets a new seed).  The adjoint re-draws the same noise; nothing [B,W]-shaped is kept."""
    B = logits.shape[0]
    qidx, qs = _query_args(dc, query, B)
    if seed_dev is not None:
        if not (seed_dev.is_cuda and seed_dev.dtype == torch.int64 and seed_dev.numel() == 1):
            raise RuntimeError("vael_b200: seed_dev must be a one-element int64 CUDA tensor")
    elif noise is None and seed is None:
        seed = int(torch.randint(0, 2 ** 62, (1,)).item())
    return _FusedLogicFunction.apply(logits, dc, qidx, qs, float(eps), float(tau), noise, int(seed or 0), seed_dev,
                                     bool(want_y_soft))


# ---------------------------------------------------------------------------
REC_LOSS = {"LAPLACE": 0, "MSE": 1, "BCE": 2}


class _ElboFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, recon, x, mu, logvar, query_prob, recon_w, kl_w, query_w, rec_loss):
        r = _require_cuda_f32(recon, "recon")
        xx = _require_cuda_f32(x, "x")
        m = _require_cuda_f32(mu, "mu")
        lv = _require_cuda_f32(logvar, "logvar")
        q = _require_cuda_f32(query_prob, "query_prob").reshape(-1) if query_prob is not None else None
        B = r.shape[0]
        D = r.numel() // B
        if xx.numel() != r.numel():
            raise RuntimeError("vael_b200: recon and x differ in size")
        lib = _lib.load()
        terms = torch.empty(4, device=r.device, dtype=torch.float32)
        g_r, g_m, g_lv = torch.empty_like(r), torch.empty_like(m), torch.empty_like(lv)
        g_q = torch.empty_like(q) if q is not None else None
        # the arrival counter + per-CTA partials belong to ONE call in flight: a workspace per call (49 KB from the
        # caching allocator, zeroed on this stream), so calls on different streams / models / graph replays
        # overlapping eager work never share one
        ws = torch.zeros(int(lib.vael_elbo_workspace_bytes()), dtype=torch.uint8, device=r.device)
        with torch.cuda.device(r.device), _nvtx("elbo_terms"):
            _lib.check(lib.vael_elbo_terms(_p(r), _p(xx), B, D, _p(m), _p(lv), m.shape[1], _p(q), recon_w, kl_w,
                                           query_w, rec_loss, _p(terms), _p(g_r), _p(g_m), _p(g_lv), _p(g_q),
                                           _p(ws), _stream_ptr()))
        ctx.save_for_backward(g_r, g_m, g_lv, g_q if g_q is not None else terms.new_empty(0))
        ctx.has_q = q is not None
        ctx.q_shape = query_prob.shape if query_prob is not None else None
        detached = terms.detach().clone()
        ctx.mark_non_differentiable(detached)
        return terms[0], detached

    @staticmethod
    def backward(ctx, g_loss, _g_terms):
        g_r, g_m, g_lv, g_q = ctx.saved_tensors
        gq = (g_q * g_loss).reshape(ctx.q_shape) if ctx.has_q else None
        return g_r * g_loss, None, g_m * g_loss, g_lv * g_loss, gq, None, None, None, None


def elbo_terms(recon, x, mu, logvar, query_prob=None, recon_w: float = 1.0, kl_w: float = 1.0, query_w: float = 1.0,
               rec_loss: str = "LAPLACE"):
    """(loss, terms[4] = {loss, recon, kl, qbce}) of loss_function (utils/mnist_utils/train.py:13-55) for
    rec_loss in {'LAPLACE', 'MSE', 'BCE'}; the gradients are produced by the same kernel pass."""
    if rec_loss not in REC_LOSS:
        raise ValueError("Unknown reconstruction loss! Valid input are 'BCE', 'LAPLACE', or 'MSE'")
    return _ElboFunction.apply(recon, x, mu, logvar, query_prob, float(recon_w), float(kl_w), float(query_w),
                               REC_LOSS[rec_loss])
