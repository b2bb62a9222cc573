"""Time the host-buffer entry points (vael_circuit_fwdbwd_host, vael_circuit_query_grad_host) on 1 Mi rows of the
2-digit circuit and the plain pinned-copy ceiling of the box beside them.
    [VAEL_HOST_CHUNK_ROWS=..] [VAEL_HOST_FIXED_CHUNKS=1] python scripts/e2e_probe.py"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vael_b200 import _lib
from vael_b200.mnist_task import compile_addition_family
from vael_b200.ops import DeviceCircuit, _p

lib = _lib.load()
dev = torch.device("cuda:0")
dc = DeviceCircuit(compile_addition_family(2, 10).circuit, dev)
B = 1 << 20
g = torch.Generator().manual_seed(0)
hf = torch.softmax(torch.randn(B, 2, 10, generator=g), -1).reshape(B, 20).contiguous().pin_memory()
hq = torch.randint(0, 19, (B,), dtype=torch.int32, generator=g).pin_memory()
hgw = torch.randn(B, 100, generator=g).pin_memory(); hgq = torch.randn(B, generator=g).pin_memory()
hw = torch.empty(B, 100).pin_memory(); hqo = torch.empty(B).pin_memory(); hgf = torch.empty(B, 20).pin_memory()

def full():
    _lib.check(lib.vael_circuit_fwdbwd_host(dc.handle, _p(hf), _p(hq), -1, B, _p(hgw), _p(hgq), _p(hw), _p(hqo), _p(hgf), 0))
def diet():
    _lib.check(lib.vael_circuit_query_grad_host(dc.handle, _p(hf), _p(hq), -1, B, _p(hgq), _p(hqo), _p(hgf), 0))
def timeit(fn, n=8):
    fn(); fn()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    return (time.perf_counter() - t0) / n * 1e3
out = {"chunk_env": os.environ.get("VAEL_HOST_CHUNK_ROWS"), "fixed": os.environ.get("VAEL_HOST_FIXED_CHUNKS")}
out["full_ms"] = timeit(full); out["diet_ms"] = timeit(diet)
s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()
d_up = torch.empty(B, 121, device=dev); d_dn = torch.empty(B, 121, device=dev)
h_up = torch.empty(B, 121).pin_memory(); h_dn = torch.empty(B, 121).pin_memory()
def link():
    with torch.cuda.stream(s_up):
        d_up.copy_(h_up, non_blocking=True)
    with torch.cuda.stream(s_dn):
        h_dn.copy_(d_dn, non_blocking=True)
    torch.cuda.synchronize()
out["link_same_bytes_ms"] = timeit(link)
out["full_frac_of_link"] = out["link_same_bytes_ms"] / out["full_ms"]
print(json.dumps(out))
