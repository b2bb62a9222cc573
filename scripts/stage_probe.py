"""The kernels around the circuit, each launched a few times at a fixed size through the C-ABI: CUDA-event timings
as JSON on stdout; the same command line is what `ncu -k regex:...` captures (scripts/gpu_r2*.sh).

    python scripts/stage_probe.py [fused|lt|gumbel|facts|elbo|generic|queries ...]     (default: all)
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from vael_b200 import _lib  # noqa: E402
from vael_b200.mario_task import admissible_circuit, compile_mario_family, mario_tables  # noqa: E402
from vael_b200.mnist_task import compile_addition_family  # noqa: E402
from vael_b200.ops import DeviceCircuit, _p  # noqa: E402

which = set(sys.argv[1:]) or {"fused", "lt", "gumbel", "facts", "elbo", "generic", "queries", "jit"}
dev = torch.device("cuda", 0)
lib = _lib.load()
sp = torch.cuda.current_stream().cuda_stream
g = torch.Generator(device=dev).manual_seed(0)
out = {}


WARM, ITERS = int(os.environ.get("PROBE_WARM", "3")), int(os.environ.get("PROBE_ITERS", "5"))


def timeit(name, fn, bytes_per_call, n=ITERS):
    for _ in range(WARM):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    out[name] = {"ms": round(ms, 5), "GBs": round(bytes_per_call / ms / 1e6, 1)}


ck = _lib.check
if "fused" in which:
    B = 1 << 20
    dc = DeviceCircuit(compile_addition_family(2, 10).circuit, dev)
    logits = 3.0 * torch.randn(B, 20, device=dev, generator=g)
    q = torch.randint(0, 19, (B,), device=dev, dtype=torch.int32, generator=g)
    gh, gq = torch.randn(B, 20, device=dev, generator=g), torch.randn(B, device=dev, generator=g)
    facts, query, wh = torch.empty(B, 20, device=dev), torch.empty(B, device=dev), torch.empty(B, 20, device=dev)
    idx, gl = torch.empty(B, device=dev, dtype=torch.int32), torch.empty(B, 20, device=dev)
    timeit("fused_forward", lambda: ck(lib.vael_logic_train_forward(dc.handle, _p(logits), _p(q), -1, B, 1e-5, 1.0, None, 7, None,
                                                                    _p(facts), _p(query), _p(idx), _p(wh), None, sp)), B * 252)
    timeit("fused_backward", lambda: ck(lib.vael_logic_train_backward(dc.handle, _p(logits), _p(q), -1, B, 1e-5, 1.0, None, 7, None,
                                                                      _p(gh), _p(gq), None, _p(gl), sp)), B * 248)
if "lt" in which:
    B = 1 << 21
    dca = DeviceCircuit(admissible_circuit(mario_tables(compile_mario_family((3, 3)))[2]), dev)
    pm = torch.softmax(torch.randn(B, 2, 9, device=dev, generator=g), -1).clamp(1e-5, 0.9999).reshape(B, 18).contiguous()
    la, gla, gfm = torch.empty(B, 24, device=dev), torch.randn(B, 24, device=dev, generator=g), torch.empty(B, 18, device=dev)
    timeit("lt_forward", lambda: ck(lib.vael_circuit_forward(dca.handle, _p(pm), None, -1, B, _p(la), None, 0, sp)), B * 168)
    timeit("lt_backward", lambda: ck(lib.vael_circuit_backward(dca.handle, _p(pm), None, -1, B, _p(gla), None, _p(gfm), 0, sp)), B * 240)
if "gumbel" in which:
    B = 1 << 20
    p = torch.softmax(torch.randn(B, 100, device=dev, generator=g), 1)
    hb = torch.cat([torch.eye(10).repeat_interleave(10, 0), torch.eye(10).repeat(10, 1)], 1).to(dev).contiguous()
    ys, idx, wh = torch.empty(B, 100, device=dev), torch.empty(B, device=dev, dtype=torch.int32), torch.empty(B, 20, device=dev)
    gh, gs = torch.randn(B, 20, device=dev, generator=g), torch.empty(B, 100, device=dev)
    timeit("gumbel_forward", lambda: ck(lib.vael_gumbel_world_forward(_p(p), None, 7, B, 100, 0, 1.0, _p(hb), 20, _p(ys), _p(idx),
                                                                      _p(wh), sp)), B * 884)
    timeit("gumbel_backward", lambda: ck(lib.vael_gumbel_world_backward(_p(p), _p(ys), _p(gh), B, 100, 0, 1.0, _p(hb), 20, _p(gs),
                                                                        sp)), B * 1280)
if "facts" in which:
    B = 1 << 21
    lg, fo = torch.randn(B, 20, device=dev, generator=g), torch.empty(B, 20, device=dev)
    gf = torch.randn(B, 20, device=dev, generator=g)
    timeit("facts_forward", lambda: ck(lib.vael_facts_forward(_p(lg), B, 2, 10, 1e-5, _p(fo), sp)), B * 160)
    timeit("facts_backward", lambda: ck(lib.vael_facts_backward(_p(lg), _p(gf), B, 2, 10, 1e-5, _p(fo), sp)), B * 240)
    lg9, fo9 = torch.randn(B, 18, device=dev, generator=g), torch.empty(B, 18, device=dev)
    timeit("facts_clip_forward", lambda: ck(lib.vael_facts_clip_forward(_p(lg9), B, 2, 9, 1e-5, 0.9999, _p(fo9), sp)), B * 144)
if "elbo" in which:
    B = 65536
    r, x = torch.rand(B, 1, 28, 56, device=dev, generator=g), torch.rand(B, 1, 28, 56, device=dev, generator=g)
    mu, lv, q = torch.randn(B, 23, device=dev, generator=g), torch.randn(B, 23, device=dev, generator=g), torch.rand(B, device=dev, generator=g)
    terms, gr, gm, glv, gq = (torch.empty(4, device=dev), torch.empty_like(r), torch.empty_like(mu), torch.empty_like(lv),
                              torch.empty_like(q))
    ws = torch.zeros(int(lib.vael_elbo_workspace_bytes()), dtype=torch.uint8, device=dev)
    timeit("elbo_terms", lambda: ck(lib.vael_elbo_terms(_p(r), _p(x), B, 1568, _p(mu), _p(lv), 23, _p(q), 0.1, 1e-5, 1.0, 0, _p(terms),
                                                        _p(gr), _p(gm), _p(glv), _p(gq), _p(ws), sp)), B * 1568 * 12)
if "generic" in which or "queries" in which:
    B = 1 << 21
    dc = DeviceCircuit(compile_addition_family(2, 10).circuit, dev)
    lgt = 3.0 * torch.randn(B, 2, 10, device=dev, generator=g)
    pr = torch.softmax(lgt, -1) + 1e-5
    facts = (pr / pr.sum(-1, keepdim=True)).reshape(B, 20).contiguous()
    q = torch.randint(0, 19, (B,), device=dev, dtype=torch.int32, generator=g)
    if "generic" in which:
        w, qo, gf = torch.empty(B, 100, device=dev), torch.empty(B, device=dev), torch.empty(B, 20, device=dev)
        gw, gq = torch.randn(B, 100, device=dev, generator=g), torch.randn(B, device=dev, generator=g)
        timeit("generic_forward", lambda: ck(lib.vael_circuit_forward(dc.handle, _p(facts), _p(q), -1, B, _p(w), _p(qo), 4, sp)), B * 488)
        timeit("generic_backward", lambda: ck(lib.vael_circuit_backward(dc.handle, _p(facts), _p(q), -1, B, _p(gw), _p(gq), _p(gf), 4, sp)), B * 568)
    if "queries" in which:
        fam = compile_mario_family((3, 3))
        dcm = DeviceCircuit(fam.circuit, dev)
        pm = torch.softmax(torch.randn(B, 2, 9, device=dev, generator=g), -1).clamp(1e-5, 0.9999).reshape(B, 18).contiguous()
        qa = torch.empty(B, 5, device=dev)
        timeit("mario_queries_all", lambda: ck(lib.vael_circuit_queries_all(dcm.handle, _p(pm), B, _p(qa), 0, sp)), B * (72 + 20))
        if dcm.jit_status()[0]:   # the same call on the circuit's generated kernel (VAEL_FLAG_FORCE_JIT = 16)
            timeit("mario_queries_all_generated", lambda: ck(lib.vael_circuit_queries_all(dcm.handle, _p(pm), B, _p(qa), 16, sp)), B * (72 + 20))
        dc2q = DeviceCircuit(compile_addition_family(2, 10).circuit, dev)
        qa2 = torch.empty(B, 19, device=dev)
        timeit("ad2_queries_diag", lambda: ck(lib.vael_circuit_queries_all(dc2q.handle, _p(facts), B, _p(qa2), 0, sp)), B * 156)
        if dc2q.jit_status()[0]:
            timeit("ad2_queries_generated", lambda: ck(lib.vael_circuit_queries_all(dc2q.handle, _p(facts), B, _p(qa2), 16, sp)), B * 156)
        dc3 = DeviceCircuit(compile_addition_family(3, 10).circuit, dev)
        B3 = 1 << 20
        p3 = torch.softmax(3.0 * torch.randn(B3, 3, 10, device=dev, generator=g), -1).reshape(B3, 30).contiguous()
        qa3 = torch.empty(B3, 28, device=dev)
        timeit("adw_queries_diag", lambda: ck(lib.vael_circuit_queries_all(dc3.handle, _p(p3), B3, _p(qa3), 0, sp)), B3 * 232)
if "jit" in which:
    B = 1 << 21
    for tag, circ, F, W, Q in (("2digit", compile_addition_family(2, 10).circuit, 20, 100, 19),
                               ("mario", compile_mario_family((3, 3)).circuit, 18, 81, 5)):
        dc = DeviceCircuit(circ, dev)
        ok, log = dc.jit_status()
        if not ok:
            out["jit_" + tag] = {"error": log[:300]}
            continue
        n = F // 2
        pr = torch.softmax(3.0 * torch.randn(B, 2, n, device=dev, generator=g), -1) + 1e-5
        facts = (pr / pr.sum(-1, keepdim=True)).reshape(B, F).contiguous()
        q = torch.randint(0, min(Q, 4), (B,), device=dev, dtype=torch.int32, generator=g)
        w, qo, gf = torch.empty(B, W, device=dev), torch.empty(B, device=dev), torch.empty(B, F, device=dev)
        gw, gq = torch.randn(B, W, device=dev, generator=g), torch.randn(B, device=dev, generator=g)
        fb, bb = F * 4 + 4 + W * 4 + 4, W * 4 + 4 + F * 4 + 4 + F * 4
        timeit(f"jit_forward_{tag}", lambda: ck(lib.vael_circuit_forward(dc.handle, _p(facts), _p(q), -1, B, _p(w), _p(qo), 16, sp)), B * fb)
        timeit(f"jit_backward_{tag}", lambda: ck(lib.vael_circuit_backward(dc.handle, _p(facts), _p(q), -1, B, _p(gw), _p(gq), _p(gf), 16, sp)), B * bb)
        timeit(f"structured_forward_{tag}", lambda: ck(lib.vael_circuit_forward(dc.handle, _p(facts), _p(q), -1, B, _p(w), _p(qo), 0, sp)), B * fb)
        timeit(f"structured_backward_{tag}", lambda: ck(lib.vael_circuit_backward(dc.handle, _p(facts), _p(q), -1, B, _p(gw), _p(gq), _p(gf), 0, sp)), B * bb)
if "jitwide" in which:
    B = 1 << 19
    dc = DeviceCircuit(compile_addition_family(3, 10).circuit, dev)
    ok, log = dc.jit_status()
    if ok:
        pr = torch.softmax(3.0 * torch.randn(B, 3, 10, device=dev, generator=g), -1)
        facts = pr.reshape(B, 30).contiguous()
        q = torch.randint(0, 28, (B,), device=dev, dtype=torch.int32, generator=g)
        w, qo, gf = torch.empty(B, 1000, device=dev), torch.empty(B, device=dev), torch.empty(B, 30, device=dev)
        gw, gq = torch.randn(B, 1000, device=dev, generator=g), torch.randn(B, device=dev, generator=g)
        for tag, flag in (("jit", 16), ("structured", 0), ("interpreter", 4)):
            timeit(f"{tag}_forward_3digit", lambda: ck(lib.vael_circuit_forward(dc.handle, _p(facts), _p(q), -1, B, _p(w), _p(qo), flag, sp)), B * 4128)
            timeit(f"{tag}_backward_3digit", lambda: ck(lib.vael_circuit_backward(dc.handle, _p(facts), _p(q), -1, B, _p(gw), _p(gq), _p(gf), flag, sp)), B * 4248)
    else:
        out["jitwide"] = {"error": log[:300]}
print(json.dumps(out))
