"""Loss functions and the training step.

loss_function / mario_loss_function keep the reference signatures and return values
(utils/mnist_utils/train.py:13-51, utils/mario_utils/train.py:11-65); with the shipped
configuration (rec_loss='LAPLACE', config.py:17,45) on CUDA tensors every term and its
gradient comes from one pass of the ELBO kernel.  `train_step` is the body of the
reference's batch loop (utils/mnist_utils/train.py:105-150) without its five per-step
`.cpu()` synchronisations; under torch.distributed the encoder/decoder gradients are
averaged by DistributedDataParallel (NCCL), the circuit has no parameters.
"""
from __future__ import annotations

import torch

from .ops import elbo_terms


def _recon_other(x_recon, x, rec_loss):
    if rec_loss == "BCE":
        return torch.nn.functional.binary_cross_entropy(torch.flatten(x_recon), torch.flatten(x), reduction="mean")
    if rec_loss == "MSE":
        return torch.nn.functional.mse_loss(torch.flatten(x_recon), torch.flatten(x), reduction="mean")
    raise ValueError("Unknown reconstruction loss! Valid input are 'BCE', 'LAPLACE', or 'MSE'")


def _kl(mu, logvar):
    return torch.mean(-0.5 * torch.sum(1.0 + logvar - mu.pow(2) - logvar.exp(), dim=1))


def _qbce(p):
    p = torch.flatten(p)
    return torch.nn.functional.binary_cross_entropy(p, torch.ones_like(p), reduction="mean")


def img_log_likelihood(recon, xs):
    """log Laplace(xs | recon, 1) summed over (C,H,W) (utils/mnist_utils/train.py:54-55)."""
    return (-0.6931471805599453 - (xs - recon).abs()).sum(dim=(1, 2, 3))


def loss_function(x_recon, x, mu, logvar, add_prob, model=None, labels=None, query=True, recon_w=1, kl_w=1, query_w=1,
                  sup_w=1, sup=False, rec_loss="MSE"):
    """-> (loss, recon_loss, gauss_kl_div, query_cross_entropy, label_cross_entropy)"""
    zero = torch.zeros(size=(), device=x_recon.device)
    if sup:
        raise NotImplementedError("digit supervision reads model.digits_probs, which no reference model defines "
                                  "(utils/mnist_utils/train.py:35-43; sup_w=0 in config.py:28)")
    if rec_loss == "LAPLACE" and x_recon.is_cuda:
        loss, terms = elbo_terms(x_recon, x, mu, logvar, add_prob if query else None, recon_w, kl_w,
                                 query_w if query else 0.0)
        return loss, terms[1], terms[2], (terms[3] if query else zero), zero
    recon_loss = -img_log_likelihood(x_recon, x).mean() if rec_loss == "LAPLACE" else _recon_other(x_recon, x, rec_loss)
    kl = _kl(mu, logvar)
    qce = _qbce(add_prob) if query else zero
    return recon_w * recon_loss + kl_w * kl + query_w * qce + sup_w * zero, recon_loss, kl, qce, zero


def mario_loss_function(x1, x_recon1, mu1, logvar1, x2, x_recon2, mu2, logvar2, query_prob, recon_w=1, kl_w=1,
                        query_w=1, sup_w=1, sup=False, rec_loss="MSE"):
    """-> dict(elbo, true_elbo, recon_loss1/2, gauss_kl_div1/2, queryBCE, labelBCE)"""
    zero = torch.zeros(size=(), device=x_recon1.device)
    if rec_loss == "LAPLACE" and x_recon1.is_cuda:
        # frame 1 carries the query term, frame 2 only reconstruction + KL
        l1, t1 = elbo_terms(x_recon1, x1, mu1, logvar1, query_prob, recon_w, kl_w, query_w)
        l2, t2 = elbo_terms(x_recon2, x2, mu2, logvar2, None, recon_w, kl_w, 0.0)
        r1, k1, qb, r2, k2 = t1[1], t1[2], t1[3], t2[1], t2[2]
        elbo = l1 + l2
    else:
        if rec_loss == "LAPLACE":
            r1, r2 = -img_log_likelihood(x_recon1, x1).mean(), -img_log_likelihood(x_recon2, x2).mean()
        else:
            r1, r2 = _recon_other(x_recon1, x1, rec_loss), _recon_other(x_recon2, x2, rec_loss)
        k1, k2, qb = _kl(mu1, logvar1), _kl(mu2, logvar2), _qbce(query_prob)
        elbo = recon_w * (r1 + r2) + kl_w * (k1 + k2) + query_w * qb + sup_w * zero
    true_elbo = (r1 + r2 + k1 + k2 + qb + zero).detach()
    return {"elbo": elbo, "true_elbo": true_elbo, "recon_loss1": r1, "recon_loss2": r2, "gauss_kl_div1": k1,
            "gauss_kl_div2": k2, "queryBCE": qb, "labelBCE": zero}


def train_step(model, optimizer, data, labels, recon_w=1e-1, kl_w=1e-5, query_w=1.0, rec_loss="LAPLACE"):
    """One optimisation step on a [B,1,28,56] batch; returns the detached [4] term vector
    (loss, recon, kl, qbce) on the device — no host synchronisation."""
    optimizer.zero_grad(set_to_none=True)
    recon, mu, logvar, add_prob = model(data, labels)
    loss, rl, kl, qce, _ = loss_function(recon, data, mu, logvar, add_prob, labels=labels, query=True,
                                         recon_w=recon_w, kl_w=kl_w, query_w=query_w, sup_w=0.0, sup=False,
                                         rec_loss=rec_loss)
    loss.backward()
    optimizer.step()
    return torch.stack([loss.detach(), rl.detach(), kl.detach(), qce.detach()])


def mario_train_step(model, optimizer, data, moves, recon_w=1e-1, kl_w=1e-5, query_w=1.0, rec_loss="LAPLACE"):
    """One optimisation step of the Mario model on a [B,3,200,100] batch of stacked frame pairs with one-hot
    moves [B,4] (the body of utils/mario_utils/train.py:105-150); returns the detached elbo on the device."""
    optimizer.zero_grad(set_to_none=True)
    x1, r1, mu1, lv1, x2, r2, mu2, lv2, q = model(data, moves)
    out = mario_loss_function(x1, r1, mu1, lv1, x2, r2, mu2, lv2, q, recon_w=recon_w, kl_w=kl_w, query_w=query_w,
                              sup_w=0.0, rec_loss=rec_loss)
    out["elbo"].backward()
    optimizer.step()
    return out["elbo"].detach()


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


class GraphedTrainStep:
    """The whole optimisation step (forward, ELBO, adjoint, Adam) captured once in a CUDA graph and
    replayed per batch: the B=32 reference configuration is launch-bound (~150 kernels of a few
    microseconds each), so removing the per-kernel launch and Python cost is what makes the step fast
    (SURVEY.md §7 hard part 8, §8(f) rank 3).  Every vael_b200 kernel only enqueues work on the current
    stream, so the capture needs nothing special; random numbers come from torch's graph-safe Philox
    generator (`model.graph_safe_rng = True` makes the world sampler take its Gumbel noise from it
    instead of a host-drawn seed, which a graph would freeze).

    Usage: step = GraphedTrainStep(model, optimizer, batch_shape); terms = step(data, labels)
    The optimizer must be capturable (torch.optim.Adam(..., capturable=True)).  `step_fn` is the eager step
    to capture: `train_step` (MNIST, returns the [4] term vector) or `mario_train_step` (Mario: `labels` are the
    one-hot moves [B,4]; returns the elbo).
    """

    def __init__(self, model, optimizer, data_shape, labels_shape=None, device=None, warmup=3,
                 memory_format=torch.contiguous_format, step_fn=None, **loss_kw):
        step_fn = step_fn or train_step
        device = torch.device(device) if device is not None else next(model.parameters()).device
        if device.type != "cuda":
            raise RuntimeError("vael_b200: CUDA graphs need a CUDA device")
        self.model, self.optimizer, self.loss_kw = model, optimizer, loss_kw
        B = data_shape[0]
        self.data = torch.zeros(data_shape, device=device)
        if memory_format == torch.channels_last:
            self.data = as_channels_last(self.data)
        self.labels = torch.zeros(labels_shape or (B, 4), dtype=torch.long, device=device)
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
        side = torch.cuda.Stream(device=device)
        side.wait_stream(torch.cuda.current_stream(device))
        with torch.cuda.stream(side):
            for _ in range(warmup):   # allocator warm-up, lazy circuit upload, cuDNN plan selection
                step_fn(model, optimizer, self.data, self.labels, **loss_kw)
        torch.cuda.current_stream(device).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        optimizer.zero_grad(set_to_none=True)
        with torch.cuda.graph(self.graph, stream=side):   # the stream the warm-up ran on: AccumulateGrad nodes match
            self.terms = step_fn(model, optimizer, self.data, self.labels, **loss_kw)

    def __call__(self, data, labels):
        """Copies the batch into the static buffers and replays the step; returns the [4] term vector
        (a static device tensor, overwritten by the next call)."""
        self.data.copy_(data, non_blocking=True)
        self.labels.copy_(labels, non_blocking=True)
        self.graph.replay()
        return self.terms
