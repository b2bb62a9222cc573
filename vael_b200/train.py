"""Loss functions and the training step.

loss_function / mario_loss_function keep the reference signatures and return values
(utils/mnist_utils/train.py:13-51, utils/mario_utils/train.py:11-65); with the shipped
configuration (rec_loss='LAPLACE', config.py:17,45) as with 'MSE' / 'BCE' every term and its
gradient comes from one pass of the ELBO kernel; CPU tensors are rejected (no fallback).  `train_step` is the body of the
reference's batch loop (utils/mnist_utils/train.py:105-150) without its five per-step
`.cpu()` synchronisations; under torch.distributed the encoder/decoder gradients are
averaged by DistributedDataParallel (NCCL), the circuit has no parameters.
"""
from __future__ import annotations

import torch

from .ops import elbo_terms


def img_log_likelihood(recon, xs):
    """log Laplace(xs | recon, 1) summed over (C,H,W) (utils/mnist_utils/train.py:54-55)."""
    return (-0.6931471805599453 - (xs - recon).abs()).sum(dim=(1, 2, 3))


def _require_cuda(t, what):
    if not t.is_cuda:
        raise RuntimeError(f"vael_b200: {what} needs CUDA tensors — the ELBO terms are computed by a kernel of "
                           "libvael_b200.so and there is no CPU fallback")


def loss_function(x_recon, x, mu, logvar, add_prob, model=None, labels=None, query=True, recon_w=1, kl_w=1, query_w=1,
                  sup_w=1, sup=False, rec_loss="MSE"):
    """-> (loss, recon_loss, gauss_kl_div, query_cross_entropy, label_cross_entropy); every term and its gradient
    from one pass of the ELBO kernel, for each of the reference's rec_loss values."""
    if sup:
        raise NotImplementedError("digit supervision reads model.digits_probs, which no reference model defines "
                                  "(utils/mnist_utils/train.py:35-43; sup_w=0 in config.py:28)")
    _require_cuda(x_recon, "loss_function")
    zero = torch.zeros(size=(), device=x_recon.device)
    loss, terms = elbo_terms(x_recon, x, mu, logvar, add_prob if query else None, recon_w, kl_w,
                             query_w if query else 0.0, rec_loss)
    return loss, terms[1], terms[2], (terms[3] if query else zero), zero


def mario_loss_function(x1, x_recon1, mu1, logvar1, x2, x_recon2, mu2, logvar2, query_prob, recon_w=1, kl_w=1,
                        query_w=1, sup_w=1, sup=False, rec_loss="MSE"):
    """-> dict(elbo, true_elbo, recon_loss1/2, gauss_kl_div1/2, queryBCE, labelBCE)"""
    _require_cuda(x_recon1, "mario_loss_function")
    zero = torch.zeros(size=(), device=x_recon1.device)
    # frame 1 carries the query term, frame 2 only reconstruction + KL
    l1, t1 = elbo_terms(x_recon1, x1, mu1, logvar1, query_prob, recon_w, kl_w, query_w, rec_loss)
    l2, t2 = elbo_terms(x_recon2, x2, mu2, logvar2, None, recon_w, kl_w, 0.0, rec_loss)
    r1, k1, qb, r2, k2 = t1[1], t1[2], t1[3], t2[1], t2[2]
    elbo = l1 + l2
    true_elbo = (r1 + r2 + k1 + k2 + qb + zero).detach()
    return {"elbo": elbo, "true_elbo": true_elbo, "recon_loss1": r1, "recon_loss2": r2, "gauss_kl_div1": k1,
            "gauss_kl_div2": k2, "queryBCE": qb, "labelBCE": zero}


def forward_backward(model, data, labels, recon_w=1e-1, kl_w=1e-5, query_w=1.0, rec_loss="LAPLACE"):
    """Forward, ELBO and adjoint of one [B,1,28,56] batch (utils/mnist_utils/train.py:123-147); gradients
    accumulate into `p.grad`.  Returns the detached [4] term vector (loss, recon, kl, qbce) on the device."""
    recon, mu, logvar, add_prob = model(data, labels)
    loss, rl, kl, qce, _ = loss_function(recon, data, mu, logvar, add_prob, labels=labels, query=True,
                                         recon_w=recon_w, kl_w=kl_w, query_w=query_w, sup_w=0.0, sup=False,
                                         rec_loss=rec_loss)
    loss.backward()
    return torch.stack([loss.detach(), rl.detach(), kl.detach(), qce.detach()])


def train_step(model, optimizer, data, labels, **loss_kw):
    """One optimisation step on a [B,1,28,56] batch; returns the detached [4] term vector
    (loss, recon, kl, qbce) on the device — no host synchronisation."""
    optimizer.zero_grad(set_to_none=True)
    terms = forward_backward(model, data, labels, **loss_kw)
    optimizer.step()
    return terms


def mario_forward_backward(model, data, moves, recon_w=1e-1, kl_w=1e-5, query_w=1.0, rec_loss="LAPLACE"):
    x1, r1, mu1, lv1, x2, r2, mu2, lv2, q = model(data, moves)
    out = mario_loss_function(x1, r1, mu1, lv1, x2, r2, mu2, lv2, q, recon_w=recon_w, kl_w=kl_w, query_w=query_w,
                              sup_w=0.0, rec_loss=rec_loss)
    out["elbo"].backward()
    return out["elbo"].detach()


def mario_train_step(model, optimizer, data, moves, **loss_kw):
    """One optimisation step of the Mario model on a [B,3,200,100] batch of stacked frame pairs with one-hot
    moves [B,4] (the body of utils/mario_utils/train.py:105-150); returns the detached elbo on the device."""
    optimizer.zero_grad(set_to_none=True)
    elbo = mario_forward_backward(model, data, moves, **loss_kw)
    optimizer.step()
    return elbo


def as_channels_last(t: torch.Tensor) -> torch.Tensor:
    """`t.contiguous(memory_format=torch.channels_last)` that also works for ONE-channel images.  For C = 1 the NCHW
    and NHWC layouts are the same bytes, torch keeps the NCHW strides, and cuDNN's format heuristic then treats the
    tensor as NCHW: the first convolution writes an NCHW activation and its backward converts the whole
    [B,64,14,28] gradient back (≈0.9 ms of layout copies per B = 4096 step).  Giving the same bytes explicit NHWC
    strides (H*W, 1, W, 1) makes the whole network run channels_last end to end."""
    t = t.contiguous(memory_format=torch.channels_last)
    if t.dim() == 4 and t.size(1) == 1:
        _, _, H, W = t.shape
        t = t.as_strided(t.size(), (H * W, 1, W, 1))
    return t


class FlatGradBucket:
    """All parameter gradients of a model as views of ONE flat buffer, averaged over the ranks by a single
    all-reduce (7.4 MB for the MNIST model): what DistributedDataParallel's buckets do, without its autograd
    hooks — so forward/backward and the optimizer can each be a CUDA graph with the collective between them
    (the circuit has no parameters; this is the only collective in training).  Each view has
    its parameter's strides (channels_last weights included), so the fused Adam sees matching layouts."""

    def __init__(self, params, group=None):
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        dev = self.params[0].device
        total = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(total, dtype=self.params[0].dtype, device=dev)
        off = 0
        for p in self.params:
            if not (p.is_contiguous() or p.is_contiguous(memory_format=torch.channels_last)
                    or p.is_contiguous(memory_format=torch.channels_last_3d)):
                raise RuntimeError("vael_b200: FlatGradBucket needs dense parameters")
            p.grad = torch.as_strided(self.flat, p.size(), p.stride(), off)
            off += p.numel()

    def world(self) -> int:
        import torch.distributed as dist
        return dist.get_world_size(self.group) if dist.is_available() and dist.is_initialized() else 1

    def zero(self):
        self.flat.zero_()

    def broadcast_parameters(self, src=0):
        """Every rank starts from rank `src`'s weights, as DistributedDataParallel's constructor does."""
        import torch.distributed as dist
        if self.world() > 1:
            with torch.no_grad():
                for p in self.params:
                    dist.broadcast(p, src=src, group=self.group)

    def all_reduce_mean(self):
        import torch.distributed as dist
        n = self.world()
        if n == 1:
            return
        if dist.get_backend(self.group) == "nccl":
            dist.all_reduce(self.flat, op=dist.ReduceOp.AVG, group=self.group)
        else:   # gloo (the CPU tests) has no AVG
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
            self.flat.div_(n)


def _optimizer_snapshot(model, optimizer):
    params = [p.detach().clone() for p in model.parameters()]
    state = {}
    for group in optimizer.param_groups:
        for p in group["params"]:
            if p in optimizer.state:
                state[id(p)] = {k: (v.detach().clone() if torch.is_tensor(v) else v) for k, v in optimizer.state[p].items()}
    return params, state


def _optimizer_restore(model, optimizer, snap):
    """Put the weights and the optimizer state back IN PLACE (the tensors keep their identity: a CUDA graph
    captured afterwards updates the same memory); state created by the warm-up of a fresh optimizer is zeroed."""
    params, state = snap
    with torch.no_grad():
        for p, saved in zip(model.parameters(), params):
            p.copy_(saved)
        for group in optimizer.param_groups:
            for p in group["params"]:
                for k, v in optimizer.state.get(p, {}).items():
                    if torch.is_tensor(v):
                        old = state.get(id(p), {}).get(k)
                        v.copy_(old) if old is not None else v.zero_()


class GraphedTrainStep:
    """The whole optimisation step (forward, ELBO, adjoint, Adam) captured once in a CUDA graph and
    replayed per batch: the B=32 reference configuration is launch-bound (~150 kernels of a few
    microseconds each), so removing the per-kernel launch and Python cost is what makes the step fast
    (SURVEY.md §7 hard part 8, §8(f) rank 3).  Every vael_b200 kernel only enqueues work on the current
    stream, so the capture needs nothing special; random numbers come from torch's graph-safe Philox
    generator (`model.graph_safe_rng = True` makes the world sampler take its Gumbel noise from it
    instead of a host-drawn seed, which a graph would freeze).

    Data parallel (torch.distributed initialised with more than one rank, or `data_parallel=True`): the step
    is TWO graphs — forward + backward into a FlatGradBucket, then Adam — with one NCCL all-reduce (mean) of
    the flat gradient between them, issued on the same stream; parameters are broadcast from rank 0 first.

    Building the step leaves the model weights and the optimizer state as they were: the warm-up steps
    (allocator, lazy circuit upload, cuDNN plan selection, NCCL communicator) run on the zero-filled static
    batch and are undone before the capture.

    Usage: step = GraphedTrainStep(model, optimizer, batch_shape); terms = step(data, labels)
    The optimizer must be capturable (torch.optim.Adam(..., capturable=True)).  `step_fn` is the forward +
    backward to capture: `forward_backward` (MNIST, returns the [4] term vector) or `mario_forward_backward`
    (Mario: `labels` are the one-hot moves [B,4]; returns the elbo); the eager steps `train_step` /
    `mario_train_step` are accepted too and mapped to them.
    """

    def __init__(self, model, optimizer, data_shape, labels_shape=None, device=None, warmup=3,
                 memory_format=torch.contiguous_format, step_fn=None, data_parallel=None, group=None, **loss_kw):
        import torch.distributed as dist
        fb = {None: forward_backward, train_step: forward_backward, mario_train_step: mario_forward_backward}.get(step_fn, step_fn)
        device = torch.device(device) if device is not None else next(model.parameters()).device
        if device.type != "cuda":
            raise RuntimeError("vael_b200: CUDA graphs need a CUDA device")
        if any(not g.get("capturable", False) for g in optimizer.param_groups):
            raise RuntimeError("vael_b200: GraphedTrainStep needs a capturable optimizer (torch.optim.Adam(..., capturable=True))")
        if fb is forward_backward and hasattr(model, "query_per_row") and not model.query_per_row:
            raise RuntimeError("vael_b200: GraphedTrainStep needs query_per_row=True (the one-query-per-batch mode reads "
                               "labels[0] on the host, which a CUDA graph would freeze)")
        if data_parallel is None:
            data_parallel = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        self.model, self.optimizer, self.loss_kw = model, optimizer, loss_kw
        B = data_shape[0]
        self.data = torch.zeros(data_shape, device=device)
        if memory_format == torch.channels_last:
            self.data = as_channels_last(self.data)
        self.labels = torch.zeros(labels_shape or (B, 4), dtype=torch.long, device=device)
        rng_before = getattr(model, "graph_safe_rng", False)
        model.graph_safe_rng = True
        # The model keeps its last forward's tensors as attributes (facts_probs, worlds_prob, query_prob, ... — the
        # reference's side-effect attributes).  After eager steps they keep that autograd graph, and with it the
        # parameters' AccumulateGrad nodes created on the default stream, alive — and a backward captured on another
        # stream would then touch the legacy stream (cudaErrorStreamCaptureImplicit).  Detach them first.
        def _detached(v):
            if torch.is_tensor(v):
                return v.detach() if v.grad_fn is not None else v
            if isinstance(v, (tuple, list)):
                return type(v)(_detached(u) for u in v)
            return v

        for name, value in list(vars(model).items()):
            if torch.is_tensor(value) or isinstance(value, (tuple, list)):
                object.__setattr__(model, name, _detached(value))
        self.bucket = None
        try:
            if data_parallel:
                self.bucket = FlatGradBucket(model.parameters(), group)
                self.bucket.broadcast_parameters()
            snap = _optimizer_snapshot(model, optimizer)
            side = torch.cuda.Stream(device=device)
            side.wait_stream(torch.cuda.current_stream(device))
            with torch.cuda.stream(side):
                for _ in range(warmup):
                    self._eager(fb)
                _optimizer_restore(model, optimizer, snap)
            torch.cuda.current_stream(device).wait_stream(side)
            self.graph = torch.cuda.CUDAGraph()
            self.graph_opt = None
            if self.bucket is None:
                optimizer.zero_grad(set_to_none=True)
                with torch.cuda.graph(self.graph, stream=side):   # the stream the warm-up ran on: AccumulateGrad nodes match
                    self.terms = fb(model, self.data, self.labels, **loss_kw)
                    optimizer.step()
            else:
                with torch.cuda.graph(self.graph, stream=side):
                    self.bucket.zero()
                    self.terms = fb(model, self.data, self.labels, **loss_kw)
                self.graph_opt = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self.graph_opt, stream=side, pool=self.graph.pool()):
                    optimizer.step()
        except Exception:
            model.graph_safe_rng = rng_before
            raise

    def _eager(self, fb):
        if self.bucket is None:
            self.optimizer.zero_grad(set_to_none=True)
        else:
            self.bucket.zero()
        fb(self.model, self.data, self.labels, **self.loss_kw)
        if self.bucket is not None:
            self.bucket.all_reduce_mean()
        self.optimizer.step()

    def __call__(self, data, labels):
        """Copies the batch into the static buffers and replays the step; returns the [4] term vector
        (a static device tensor, overwritten by the next call; the local shard's means under data parallelism)."""
        self.data.copy_(data, non_blocking=True)
        self.labels.copy_(labels, non_blocking=True)
        self.graph.replay()
        if self.bucket is not None:
            self.bucket.all_reduce_mean()
            self.graph_opt.replay()
        return self.terms
