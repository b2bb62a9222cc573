"""Out-of-bounds WRITE checks of our own (compute-sanitizer is not available on the GPU pool): every output of every
device entry point is carved out of a larger buffer filled with a sentinel bit pattern, the entry point is called
through the C-ABI at ragged batch sizes (tail tiles, plain-copy fallbacks) and at 4-byte-only alignment, and the
sentinels on both sides must be untouched afterwards — and every element inside must have been written."""
import math

import numpy as np
import pytest
import torch

from tests.util import synthetic_facts

pytestmark = pytest.mark.gpu

PAD = 2048                 # sentinel words on either side of an output
SENT = 0x7FC0DEAD          # a NaN payload no kernel produces
BATCHES = (1, 31, 33, 77, 1025)


class Guarded:
    """An output tensor with sentinel bands around it; `shift` = 1 makes its base 4-byte aligned only."""

    def __init__(self, shape, dtype=torch.float32, shift=0):
        n = int(math.prod(shape))
        self.n, self.lo = n, PAD + shift
        self.raw = torch.full((PAD + shift + n + PAD,), SENT, dtype=torch.int32, device="cuda")
        body = self.raw[self.lo:self.lo + n]
        self.t = (body.view(torch.float32) if dtype == torch.float32 else body).view(*shape)

    def check(self, what, written=True):
        raw = self.raw.cpu().numpy()
        assert (raw[:self.lo] == SENT).all(), f"{what}: wrote BEFORE the buffer"
        assert (raw[self.lo + self.n:] == SENT).all(), f"{what}: wrote PAST the buffer"
        if written and self.n:
            assert (raw[self.lo:self.lo + self.n] != SENT).all(), f"{what}: left part of the output unwritten"


def _api():
    from vael_b200 import _lib
    from vael_b200.ops import _p
    return _lib, _lib.load(), _p, torch.cuda.current_stream().cuda_stream


def _circuits():
    from vael_b200.mario_task import admissible_circuit, compile_mario_family, mario_tables
    from vael_b200.mnist_task import compile_addition_family
    from vael_b200.ops import DeviceCircuit
    dev = torch.device("cuda", 0)
    fam = compile_mario_family((3, 3))
    return {
        "2digit": (DeviceCircuit(compile_addition_family(2, 10).circuit, dev), (10, 10)),
        "3digit": (DeviceCircuit(compile_addition_family(3, 10).circuit, dev), (10, 10, 10)),
        "mario": (DeviceCircuit(fam.circuit, dev), (9, 9)),
        "admissible": (DeviceCircuit(admissible_circuit(mario_tables(fam)[2]), dev), (9, 9)),
    }


@pytest.fixture(scope="module")
def circuits(cuda_device):
    return _circuits()


@pytest.mark.parametrize("name", ["2digit", "3digit", "mario", "admissible"])
@pytest.mark.parametrize("path", ["default", "interpreter", "generated"])
def test_circuit_entry_points_stay_inside_their_outputs(circuits, name, path, monkeypatch):
    from vael_b200 import circuit as cflags
    _lib, lib, _p, sp = _api()
    dc, groups = circuits[name]
    flags = {"default": 0, "interpreter": cflags.FLAG_FORCE_GENERIC, "generated": cflags.FLAG_FORCE_JIT}[path]
    if path == "generated" and (not dc.jit_status()[0] or name == "admissible"):
        pytest.skip("no generated kernels for this circuit (log semiring) or no NVRTC on this box")
    F, W, Q = dc.n_facts, dc.n_worlds, dc.n_queries
    for B in BATCHES:
        for shift in (0, 1):
            if path == "generated" and name == "3digit" and shift:
                continue   # the column-chunked generated kernels need 16-byte aligned world tensors (else: interpreter)
            facts = torch.from_numpy(synthetic_facts(B, groups=groups, seed=B)).cuda().clamp(1e-5, 0.9999)
            qidx = (torch.arange(B, device="cuda", dtype=torch.int32) % max(Q, 1)) if Q else None
            worlds, query, gf = Guarded((B, W), shift=shift), Guarded((B,), shift=shift), Guarded((B, F), shift=shift)
            for fl in (flags, flags | cflags.FLAG_NORMALIZE):
                if name == "admissible" and fl & cflags.FLAG_NORMALIZE:
                    continue
                _lib.check(lib.vael_circuit_forward(dc.handle, _p(facts), _p(qidx), -1, B, _p(worlds.t),
                                                    _p(query.t) if Q else None, fl, sp))
                gw, gq = torch.randn(B, W, device="cuda"), torch.randn(B, device="cuda")
                _lib.check(lib.vael_circuit_backward(dc.handle, _p(facts), _p(qidx), -1, B, _p(gw), _p(gq) if Q else None,
                                                     _p(gf.t), fl, sp))
                torch.cuda.synchronize()
                tag = f"{name}/{path} B={B} shift={shift} flags={fl}"
                worlds.check(tag + " worlds"); gf.check(tag + " grad_facts")
                query.check(tag + " query", written=bool(Q))
            if Q and name != "admissible":
                for keep in ("0", "1"):   # Mario, default path: the generated kernel / the hand-written member-list kernel
                    monkeypatch.setenv("VAEL_QUERIES_MEMBERLIST", keep)
                    qa = Guarded((B, Q), shift=shift)
                    _lib.check(lib.vael_circuit_queries_all(dc.handle, _p(facts), B, _p(qa.t), flags, sp))
                    torch.cuda.synchronize()
                    qa.check(f"{name}/{path} B={B} shift={shift} queries_all memberlist={keep}")


@pytest.mark.parametrize("n_groups,group,clip", [(2, 10, False), (3, 10, False), (2, 9, True), (1, 9, True), (4, 7, False)])
def test_facts_head_stays_inside_its_outputs(cuda_device, n_groups, group, clip):
    _lib, lib, _p, sp = _api()
    F = n_groups * group
    for B in BATCHES:
        for shift in (0, 1):
            x, g = torch.randn(B, F, device="cuda"), torch.randn(B, F, device="cuda")
            out, gx = Guarded((B, F), shift=shift), Guarded((B, F), shift=shift)
            if clip:
                _lib.check(lib.vael_facts_clip_forward(_p(x), B, n_groups, group, 1e-5, 0.9999, _p(out.t), sp))
                _lib.check(lib.vael_facts_clip_backward(_p(x), _p(g), B, n_groups, group, 1e-5, 0.9999, _p(gx.t), sp))
            else:
                _lib.check(lib.vael_facts_forward(_p(x), B, n_groups, group, 1e-5, _p(out.t), sp))
                _lib.check(lib.vael_facts_backward(_p(x), _p(g), B, n_groups, group, 1e-5, _p(gx.t), sp))
            torch.cuda.synchronize()
            out.check(f"facts {n_groups}x{group} B={B} shift={shift}"); gx.check(f"facts adjoint {n_groups}x{group} B={B}")


# (100, 20), (24, 18), (16, 5): lane <-> row kernels; (81, 18), (1000, 30): one warp per row
@pytest.mark.parametrize("W,H,is_log", [(100, 20, False), (24, 18, True), (16, 5, False), (81, 18, False), (1000, 30, False)])
@pytest.mark.parametrize("inject", [False, True])
def test_gumbel_entry_points_stay_inside_their_outputs(cuda_device, W, H, is_log, inject, monkeypatch):
    monkeypatch.setenv("VAEL_GUMBEL_ROWS", "1")   # lane <-> row kernels wherever the shape allows, also at these batch sizes
    _lib, lib, _p, sp = _api()
    gen = torch.Generator().manual_seed(W)
    hb = (torch.rand(W, H, generator=gen) < (2.0 / H)).float().cuda()
    for B in BATCHES:
        for shift in (0, 1):
            p = torch.softmax(torch.randn(B, W, generator=gen), 1).cuda()
            scores = p.log() if is_log else p
            noise = (-torch.empty(B, W).exponential_(generator=gen).log()).cuda() if inject else None
            ys, idx = Guarded((B, W), shift=shift), Guarded((B,), dtype=torch.int32, shift=shift)
            wh, gs = Guarded((B, H), shift=shift), Guarded((B, W), shift=shift)
            _lib.check(lib.vael_gumbel_world_forward(_p(scores), _p(noise), 11, B, W, int(is_log), 1.0, _p(hb), H,
                                                     _p(ys.t), _p(idx.t), _p(wh.t), sp))
            gh = torch.randn(B, H, device="cuda")
            _lib.check(lib.vael_gumbel_world_backward(_p(scores), _p(ys.t), _p(gh), B, W, int(is_log), 1.0, _p(hb), H,
                                                      _p(gs.t), sp))
            torch.cuda.synchronize()
            tag = f"gumbel W={W} H={H} B={B} shift={shift} inject={inject}"
            ys.check(tag + " y_soft"); idx.check(tag + " world_idx"); wh.check(tag + " world_h"); gs.check(tag + " grad_scores")
            assert int(idx.t.min()) >= 0 and int(idx.t.max()) < W


@pytest.mark.parametrize("inject", [False, True])
def test_fused_logic_entry_points_stay_inside_their_outputs(circuits, inject):
    _lib, lib, _p, sp = _api()
    dc, _ = circuits["2digit"]
    gen = torch.Generator().manual_seed(3)
    for B in BATCHES + (20001,):   # 20001 rows: the tile-pipeline forward of streaming batches (ragged last tile)
        for shift in (0, 1):
            x = (2 * torch.randn(B, 20, generator=gen)).cuda()
            qidx = torch.arange(B, device="cuda", dtype=torch.int32) % 19
            noise = (-torch.empty(B, 100).exponential_(generator=gen).log()).cuda() if inject else None
            facts, query = Guarded((B, 20), shift=shift), Guarded((B,), shift=shift)
            idx, wh = Guarded((B,), dtype=torch.int32, shift=shift), Guarded((B, 20), shift=shift)
            ys, gl = Guarded((B, 100), shift=shift), Guarded((B, 20), shift=shift)
            for y_soft in (None, ys):
                _lib.check(lib.vael_logic_train_forward(dc.handle, _p(x), _p(qidx), -1, B, 1e-5, 1.0, _p(noise), 5, None,
                                                        _p(facts.t), _p(query.t), _p(idx.t), _p(wh.t),
                                                        _p(y_soft.t) if y_soft else None, sp))
            gh, gq, gf = torch.randn(B, 20, device="cuda"), torch.randn(B, device="cuda"), torch.randn(B, 20, device="cuda")
            _lib.check(lib.vael_logic_train_backward(dc.handle, _p(x), _p(qidx), -1, B, 1e-5, 1.0, _p(noise), 5, None,
                                                     _p(gh), _p(gq), _p(gf), _p(gl.t), sp))
            torch.cuda.synchronize()
            tag = f"fused B={B} shift={shift} inject={inject}"
            for g, what in ((facts, "facts"), (query, "query"), (idx, "world_idx"), (wh, "world_h"), (ys, "y_soft"),
                            (gl, "grad_logits")):
                g.check(f"{tag} {what}")


@pytest.mark.parametrize("rec", [0, 1, 2])
def test_elbo_terms_stay_inside_their_outputs(cuda_device, rec):
    _lib, lib, _p, sp = _api()
    for B, D, L in ((1, 7, 3), (33, 1568, 23), (130, 999, 5)):
        for shift in (0, 1):
            recon = torch.rand(B, D, device="cuda").clamp(1e-3, 1 - 1e-3); x = torch.rand(B, D, device="cuda")
            mu, lv = torch.randn(B, L, device="cuda"), torch.randn(B, L, device="cuda")
            q = torch.rand(B, device="cuda").clamp_min(1e-3)
            terms, gr = Guarded((4,), shift=shift), Guarded((B, D), shift=shift)
            gm, gv, gq = Guarded((B, L), shift=shift), Guarded((B, L), shift=shift), Guarded((B,), shift=shift)
            ws = torch.zeros(int(lib.vael_elbo_workspace_bytes()), dtype=torch.uint8, device="cuda")
            _lib.check(lib.vael_elbo_terms(_p(recon), _p(x), B, D, _p(mu), _p(lv), L, _p(q), 0.1, 1e-3, 1.0, rec, _p(terms.t),
                                           _p(gr.t), _p(gm.t), _p(gv.t), _p(gq.t), _p(ws), sp))
            torch.cuda.synchronize()
            for g, what in ((terms, "terms"), (gr, "grad_recon"), (gm, "grad_mu"), (gv, "grad_logvar"), (gq, "grad_query")):
                g.check(f"elbo rec={rec} B={B} D={D} shift={shift} {what}")
            assert np.isfinite(terms.t.cpu().numpy()).all()
