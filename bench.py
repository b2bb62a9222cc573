#!/usr/bin/env python
"""Benchmark of the VAEL logic hot path (BASELINE.json metric: circuit evals/s).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # the reference's own CPU path (baseline/_ref)

One "step" = one forward + one adjoint backward of the compiled 2-digit-addition circuit
over B batch rows per GPU (synthetic facts in the domain compute_facts_probability emits,
per-row random query, random upstream gradients).  One eval = one (row, labelled query)
output: 101 per row (100 digits(j,k) worlds + the addition query), SURVEY.md §8(d) M1.
Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# NCCL_DEBUG belongs to the caller (the driver reads the communicator lines): INFO / TRACE are honoured exactly as
# set, destination included.  Anything quieter (unset, or this image's default NCCL_DEBUG=VERSION, or WARN) is raised
# under torchrun to INFO for the INIT subsystem — the "rank r nranks N ... Init COMPLETE" lines and the transport — and
# sent to STDERR, so that stdout stays the one JSON line.  Done before torch (and with it NCCL) is loaded.
if int(os.environ.get("WORLD_SIZE", "1")) > 1 and os.environ.get("NCCL_DEBUG", "").upper() not in ("INFO", "TRACE"):
    os.environ["NCCL_DEBUG"] = "INFO"
    os.environ.setdefault("NCCL_DEBUG_SUBSYS", "INIT")
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

import torch  # noqa: E402

EVALS_PER_ROW = 101            # Q of query-mode evaluate: 100 worlds + 1 addition query
FWD_BYTES = 80 + 4 + 400 + 4   # read facts[20] + query idx, write worlds[100] + query   (fp32/int32)
BWD_BYTES = 400 + 4 + 80 + 4 + 80  # read grad_worlds, grad_query, facts, query idx; write grad_facts
WORKLOAD = ("2digit-MNIST-addition compiled circuit (F=20 facts, W=100 worlds, Q=19 sums), fwd+bwd, "
            "per-row query, fp32")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", type=int, default=1 << 23, help="batch rows per GPU per step (device-resident arm)")
    ap.add_argument("--e2e-rows", type=int, default=1 << 20, help="batch rows per GPU per step (host-buffer arm)")
    ap.add_argument("--cpu-rows", type=int, default=1 << 19, help="rows per step of the CPU baseline sample")
    ap.add_argument("--train-batch", type=int, default=4096, help="per-GPU batch of the VAEL train-step probe")
    ap.add_argument("--no-train", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-secondary", action="store_true")
    ap.add_argument("--wide-rows", type=int, default=1 << 20, help="rows per GPU of the 3-digit (W=1000) probe")
    return ap.parse_args()


def synthetic_facts(rows, device, seed):
    """softmax(3*N(0,1)) + 1e-5, renormalised: the a14 domain."""
    g = torch.Generator(device=device).manual_seed(seed)
    logits = 3.0 * torch.randn(rows, 2, 10, device=device, generator=g)
    p = torch.softmax(logits, dim=-1) + 1e-5
    return (p / p.sum(-1, keepdim=True)).reshape(rows, 20).contiguous(), g


# ---------------------------------------------------------------------------
# the reference's CPU path: oracle port of models/vael.py:208-225 + autograd
# ---------------------------------------------------------------------------
def host_threads():
    """Use every host core this process may run on (torchrun exports OMP_NUM_THREADS=1 by default)."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    if torch.get_num_threads() != n:
        torch.set_num_threads(n)
    return n


REF_DIR = os.path.join(ROOT, "baseline", "_ref")
STUB_DIR = os.path.join(ROOT, "baseline", "problog_stub")
TRAIN_WORKLOAD = ("2digit-MNIST addition VAEL train step (encoder/decoder CNNs, logic path, ELBO, Adam), "
                  "synthetic 28x56 images, random-init weights, fp32")


def workload_config(rows_per_gpu, train_batch, l2, parallelism, train=None):
    """`config` of the JSON line: the SAME keys from both arms (the driver compares them)."""
    t = train or {}
    return {"workload": WORKLOAD, "rows_per_gpu": rows_per_gpu, "evals_per_row": EVALS_PER_ROW, "l2": l2,
            "parallelism": parallelism, "train_workload": TRAIN_WORKLOAD, "train_batch_per_gpu": train_batch,
            "train_samples_per_sec": t.get("value"), "train_ms_per_step": t.get("ms_per_step"),
            "train_mode": t.get("mode")}


def reference_modules():
    """The UNMODIFIED reference, copied to baseline/_ref by scripts/install_ref.py (it has no setup.py to pip
    install).  Its only unsatisfiable imports, problog.logic / problog.evaluator, come from the inert stub under
    baseline/problog_stub (no vael_b200 code).  -> (models.vael, models.vael_networks, utils.mnist_utils.train)
    or None when baseline/_ref did not travel."""
    if not os.path.exists(os.path.join(REF_DIR, "models", "vael.py")):
        return None
    for d in (STUB_DIR, REF_DIR):
        if d not in sys.path:
            sys.path.insert(0, d)
    import importlib
    return tuple(importlib.import_module(m) for m in ("models.vael", "models.vael_networks", "utils.mnist_utils.train"))


def reference_w_q():
    """build_worlds_queries_matrix (utils/mnist_utils/mnist_task_VAEL.py:222-234): that module imports matplotlib
    and torchvision at the top, so the one function is exec'd from the reference's source file."""
    import ast
    from itertools import product
    rel = "utils/mnist_utils/mnist_task_VAEL.py"
    tree = ast.parse(open(os.path.join(REF_DIR, rel)).read())
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "build_worlds_queries_matrix")
    ns = {"torch": torch, "product": product}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), rel, "exec"), ns)
    return ns["build_worlds_queries_matrix"](2, 10)


def cpu_reference_rate(rows, steps, warmup, chunks=1):
    """evals/s of the reference's dense logic path (forward + autograd backward) on the host cores, torch CPU
    with all its threads: MNISTPairsVAELModel.problog_inference (models/vael.py:208-225) of the reference itself
    (baseline/_ref; `kind` "reference"), else the oracle's port of it (`kind` "port").  One step = `chunks`
    batches of `rows` rows (the batch is what bounds host memory: worlds + upstream gradient are 800 B per row).
    -> (rate, seconds per step, kind)"""
    facts, g = synthetic_facts(rows, "cpu", 1234)
    facts = facts.reshape(rows, 2, 10)
    gw = torch.randn(rows, 100, generator=g)
    gq = torch.randn(rows, 1, generator=g)
    q = int(torch.randint(0, 19, (1,), generator=g))
    ref = reference_modules()
    if ref is not None:
        vael, nets, _ = ref
        model = vael.MNISTPairsVAELModel(encoder=None, decoder=None, mlp=nets.MNISTPairsMLP(in_features=15, n_facts=20),
                                         latent_dims=(15, 8), model_dict=None, w_q=reference_w_q(), dropout=None,
                                         is_train=True, device="cpu")
        labels = torch.tensor([[0, 0, q, -1]]).expand(rows, 4)   # one query per batch: the reference's own mode (:212-213)
        infer, kind = (lambda f: model.problog_inference(f, labels=labels)), "reference"
    else:
        from oracle import dense_path
        w_q = dense_path.worlds_queries_matrix()
        infer, kind = (lambda f: dense_path.dense_inference(f, w_q, q)), "port"

    def step():
        for _ in range(chunks):
            f = facts.clone().requires_grad_(True)
            qp, worlds = infer(f)
            torch.autograd.backward([worlds, qp], [gw, gq])
        return f.grad

    for _ in range(warmup):
        step()
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
    dt = sum(times) / len(times)
    return rows * chunks * EVALS_PER_ROW / dt, dt, kind


def cpu_reference_train(batch, steps=3, warmup=1):
    """The reference's own train step (utils/mnist_utils/train.py:105-150: forward, loss_function LAPLACE,
    backward, Adam) on the host cores with the reference's classes from baseline/_ref; None without them."""
    ref = reference_modules()
    if ref is None:
        return None
    vael, nets, rtrain = ref
    torch.manual_seed(0)
    model = vael.MNISTPairsVAELModel(nets.MNISTPairsEncoder(hidden_channels=64, latent_dim=23, dropout=0.5),
                                     nets.MNISTPairsDecoder(hidden_channels=64, latent_dim=8, label_dim=20, dropout=0.5),
                                     nets.MNISTPairsMLP(in_features=15, n_facts=20), (15, 8), None, reference_w_q(),
                                     dropout=0.5, is_train=True, device="cpu")
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    g = torch.Generator().manual_seed(1234)
    x = torch.randn(batch, 1, 28, 56, generator=g)
    d = torch.randint(0, 10, (1, 2), generator=g).expand(batch, 2)      # the reference loader: one world per batch
    labels = torch.cat([d, d.sum(1, keepdim=True), -torch.ones(batch, 1, dtype=torch.long)], 1)

    def step():
        opt.zero_grad()
        recon, mu, logvar, add_prob = model(x, labels)
        loss, *_ = rtrain.loss_function(recon, x, mu, logvar, add_prob, model=model, labels=labels, query=True,
                                        recon_w=1e-1, kl_w=1e-5, query_w=1.0, sup_w=0.0, sup=False, rec_loss="LAPLACE")
        loss.backward()
        opt.step()

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return {"value": batch / dt, "unit": "samples/s", "ms_per_step": dt * 1e3, "mode": "reference eager, torch CPU"}


def cpu_extras():
    """The other CPU legs SURVEY.md §8(d) asks to be timed beside the GPU numbers (small, bounded):
    (ii) the reference's GRAPH path — the GraphSemiring port (batch-wide shortcuts, host-side identity
    vectors and all) driven node by node over the compiled circuit, one evaluate() = 101 queries;
    (iii) a reference-equivalent full VAEL train step (CNNs + dense logic + ELBO + Adam) at B = 32."""
    from oracle import dense_path, graph_path
    from oracle.semiring_port import SemiringPort
    from vael_b200.mnist_task import compile_addition_family
    from vael_b200.networks import MNISTPairsDecoder, MNISTPairsEncoder, MNISTPairsMLP
    out = {}
    circ = compile_addition_family(2, 10).circuit
    for B in (32, 4096):
        facts, _ = synthetic_facts(B, "cpu", 7)
        facts = facts.reshape(B, 2, 10)

        def evaluate():
            sr = SemiringPort(batch_size=B)
            sr.set_weights(graph_path.bind_weights(facts, circ.fact_labels))
            return graph_path.evaluate(circ, sr, mode="query", query_index=9)

        evaluate()
        n = 3
        t0 = time.perf_counter()
        for _ in range(n):
            evaluate()
        dt = (time.perf_counter() - t0) / n
        out[f"graph_path_B{B}"] = {"ms_per_evaluate": dt * 1e3, "evals_per_sec": B * EVALS_PER_ROW / dt,
                                   "what": "GraphSemiring port driven over the CSR circuit, forward only"}
    torch.manual_seed(0)
    enc, dec, mlp = (MNISTPairsEncoder(hidden_channels=64, latent_dim=23), MNISTPairsDecoder(label_dim=20, hidden_channels=64, latent_dim=8),
                     MNISTPairsMLP(in_features=15, n_facts=20))
    params = [p for m in (enc, dec, mlp) for p in m.parameters()]
    opt = torch.optim.Adam(params, lr=1e-3)
    B = 32
    x = torch.randn(B, 1, 28, 56)
    w_q, hb = dense_path.worlds_queries_matrix(), dense_path.herbrand_base(20)

    def step():
        opt.zero_grad()
        mu, logvar = enc(x)
        z = mu + torch.exp(logvar / 2) * torch.randn(B, 23)
        facts = dense_path.facts_probability(torch.cat(mlp(z[:, :15]), 1))
        qp, worlds = dense_path.dense_inference(facts, w_q, 9)
        world = torch.nn.functional.gumbel_softmax(torch.log(worlds), tau=1, hard=True)
        image = dec(torch.cat([z[:, 15:], world @ hb], 1))
        loss, *_ = dense_path.elbo_terms(image, x, mu, logvar, qp, 1e-1, 1e-5, 1.0)
        loss.backward()
        opt.step()

    for _ in range(2):
        step()
    t0 = time.perf_counter()
    for _ in range(5):
        step()
    dt = (time.perf_counter() - t0) / 5
    out["train_step_B32"] = {"ms_per_step": dt * 1e3, "samples_per_sec": B / dt,
                             "what": "reference-equivalent MNIST VAEL train step on the host cores (torch CPU)"}
    return out


def run_reference(args):
    """The reference arm: the reference's OWN dense logic path + autograd (baseline/_ref, stock torch CPU code
    path, none of this repo's kernels) on the box's host cores with every thread torch can use.  Under torchrun
    rank 0 alone runs; the other ranks exit 0."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, args.steps), max(1, args.warmup)
    host_threads()
    chunks = max(1, args.rows // args.cpu_rows)   # one step covers ONE GPU's rows per step (8 Mi by default)
    rate, dt, kind = cpu_reference_rate(args.cpu_rows, steps, warmup, chunks)
    cores = torch.get_num_threads()
    train = None
    if not args.no_train:
        try:
            train = cpu_reference_train(args.train_batch)
        except Exception as exc:  # noqa: BLE001 - the probe must not cost the line
            train = {"mode": "error: " + repr(exc)[:120]}
    what = ("the reference's own MNISTPairsVAELModel.problog_inference (baseline/_ref/models/vael.py:208-225) + autograd"
            if kind == "reference" else "oracle port of models/vael.py:208-225 + autograd (baseline/_ref absent)")
    line = {
        "impl": "reference", "metric": "circuit_evals_per_sec", "value": rate, "unit": "evals/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.rows, args.train_batch, "n/a (host arm: inputs in host DRAM)",
                                  "one host process, all torch CPU threads", train),
        "cpu_baseline": {"value": rate, "unit": "evals/s", "cores": cores, "kind": kind,
                         "sample": f"{chunks} batches x {args.cpu_rows} rows per step x {steps} steps "
                                   f"({dt:.2f} s/step), torch CPU fp32, {what}, os.cpu_count()={os.cpu_count()}"},
        "e2e": {"value": rate, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------
# clocks during the timed region
# ---------------------------------------------------------------------------
class ClockSampler:
    """SM clock + throttle reasons sampled from NVML every ~1 ms by a host thread while the
    main thread waits for the GPU (falls back to one nvidia-smi query if pynvml is missing)."""
    REASONS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, gpu_index):
        self.samples, self.stop_flag, self.thread, self.nv, self.h = [], False, None, None, None
        self.gpu_index = gpu_index
        try:
            import pynvml
            pynvml.nvmlInit()
            # CUDA_VISIBLE_DEVICES remaps torch's index; resolve through the PCI bus id
            bus = torch.cuda.get_device_properties(gpu_index).pci_bus_id if hasattr(
                torch.cuda.get_device_properties(gpu_index), "pci_bus_id") else None
            self.h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
            if bus is not None:
                for i in range(pynvml.nvmlDeviceGetCount()):
                    h = pynvml.nvmlDeviceGetHandleByIndex(i)
                    if pynvml.nvmlDeviceGetPciInfo(h).bus == bus:
                        self.h = h
            self.nv = pynvml
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.nv = None

    def _pump(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(
                    nv, "nvmlDeviceGetCurrentClocksEventReasons") else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((time.perf_counter(), mhz, rs))
            except Exception:
                pass
            time.sleep(0.0005)

    def stop(self, t0, t1):
        if self.nv is None:
            try:
                out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm", "--format=csv,noheader,nounits",
                                      "-i", str(self.gpu_index)], capture_output=True, text=True, timeout=10).stdout
                a, b = (float(v) for v in out.strip().split(","))
                return {"sm_mhz": a, "sm_max_mhz": b, "reasons": [], "samples": 1, "note": "single nvidia-smi query after the run"}
            except Exception:
                return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampling unavailable"]}
        self.stop_flag = True
        self.thread.join(timeout=1.0)
        inside = [s for s in self.samples if t0 <= s[0] <= t1] or self.samples
        mhz = [s[1] for s in inside]
        bits = 0
        for s in inside:
            bits |= int(s[2])
        reasons = sorted(n for n, b in self.REASONS.items() if bits & b)
        try:
            mx = float(self.nv.nvmlDeviceGetMaxClockInfo(self.h, self.nv.NVML_CLOCK_SM))
        except Exception:
            mx = None
        return {"sm_mhz": statistics.median(mhz) if mhz else None, "sm_max_mhz": mx, "reasons": reasons,
                "samples": len(mhz), "samples_whole_run": len(self.samples),
                "sm_mhz_whole_run": statistics.median([s[1] for s in self.samples]) if self.samples else None}


# ---------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist
    from vael_b200 import _lib
    from vael_b200.mnist_task import compile_addition_family
    from vael_b200.ops import DeviceCircuit, _p

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the vael_b200 path has no CPU fallback "
                         "(use --impl reference for the CPU baseline)")
    from vael_b200 import parallel
    world, rank, local = parallel.env_world()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    distributed = world > 1
    numa = parallel.bind_to_gpu_numa_node(local) if distributed and not os.environ.get("VAEL_NO_NUMA_BIND") else None
    if distributed:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        return parallel.max_over_ranks(x, dev)

    lib = _lib.load()
    fam = compile_addition_family(2, 10)
    dc = DeviceCircuit(fam.circuit, dev)
    B, K, W = args.rows, max(1, args.steps), max(3, args.warmup)
    facts, g = synthetic_facts(B, dev, parallel.rank_seed(1234, rank))
    qidx = torch.randint(0, 19, (B,), device=dev, dtype=torch.int32, generator=g)
    gw = torch.randn(B, 100, device=dev, generator=g)
    gq = torch.randn(B, device=dev, generator=g)
    worlds = torch.empty(B, 100, device=dev)
    query = torch.empty(B, device=dev)
    gf = torch.empty(B, 20, device=dev)
    stream = torch.cuda.current_stream()
    sp = torch.cuda.current_stream().cuda_stream

    def fwd():
        _lib.check(lib.vael_circuit_forward(dc.handle, _p(facts), _p(qidx), -1, B, _p(worlds), _p(query), 0, sp))

    def bwd():
        _lib.check(lib.vael_circuit_backward(dc.handle, _p(facts), _p(qidx), -1, B, _p(gw), _p(gq), _p(gf), 0, sp))

    sampler = ClockSampler(local) if rank == 0 else None   # started before warm-up: NVML's first calls are slow
    for _ in range(W):
        fwd(); bwd()
    barrier()
    # ---- timed region: exactly K steps, per-kernel events on the launching stream -------------
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(K)]
    launches0 = _lib.launch_count()
    barrier()
    t0 = time.perf_counter()
    for k in range(K):
        ev[k][0].record(stream); fwd()
        ev[k][1].record(stream); bwd()
        ev[k][2].record(stream)
    barrier()
    t1 = time.perf_counter()
    launches = _lib.launch_count() - launches0
    clocks = sampler.stop(t0, t1) if sampler else None
    total_ms = max_over_ranks(ev[0][0].elapsed_time(ev[K - 1][2]))
    fwd_ms = sum(e[0].elapsed_time(e[1]) for e in ev) / K
    bwd_ms = sum(e[1].elapsed_time(e[2]) for e in ev) / K
    ms_per_step = total_ms / K
    value = world * B * EVALS_PER_ROW / (ms_per_step * 1e-3)
    kernel_family = _lib.last_kernel()

    # ---- e2e: the C-ABI host-buffer call, pinned host memory, copies inside the timed region --------
    Be = args.e2e_rows
    e2e = None
    if not args.no_e2e:
        hf = facts[:Be].cpu().pin_memory(); hq = qidx[:Be].cpu().pin_memory()
        hgw = gw[:Be].cpu().pin_memory(); hgq = gq[:Be].cpu().pin_memory()
        hw = torch.empty(Be, 100).pin_memory(); hqo = torch.empty(Be).pin_memory(); hgf = torch.empty(Be, 20).pin_memory()

        def e2e_step():
            _lib.check(lib.vael_circuit_fwdbwd_host(dc.handle, _p(hf), _p(hq), -1, Be, _p(hgw), _p(hgq), _p(hw),
                                                    _p(hqo), _p(hgf), 0))

        for _ in range(2):
            e2e_step()
        barrier()
        ke = max(3, min(K, 10))
        te0 = time.perf_counter()
        for _ in range(ke):
            e2e_step()
        torch.cuda.synchronize()
        e2e_s = max_over_ranks((time.perf_counter() - te0) / ke)
        assert torch.equal(hw[:4096], worlds[:4096].cpu()), "e2e path and device path disagree"
        # what the link can do: the same bytes as plain pinned copies, both directions at once on two streams
        s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()
        d_up = torch.empty_like(hgw, device=dev); d_dn = torch.empty_like(hw, device=dev)
        torch.cuda.synchronize()
        tp0 = time.perf_counter()
        for _ in range(3):
            with torch.cuda.stream(s_up):
                d_up.copy_(hgw, non_blocking=True)
            with torch.cuda.stream(s_dn):
                hw.copy_(d_dn, non_blocking=True)
        torch.cuda.synchronize()
        link_gbs = Be * 400 / ((time.perf_counter() - tp0) / 3) / 1e9
        del d_up, d_dn
        # the byte diet: what the reference's training loop consumes (query in the BCE term, its gradient back to the
        # facts) without any [B,W] tensor over PCIe — vael_circuit_query_grad_host, 88 B up + 84 B down per row
        hgf2, hqo2 = torch.empty(Be, 20).pin_memory(), torch.empty(Be).pin_memory()

        def diet_step():
            _lib.check(lib.vael_circuit_query_grad_host(dc.handle, _p(hf), _p(hq), -1, Be, _p(hgq), _p(hqo2), _p(hgf2), 0))

        for _ in range(2):
            diet_step()
        barrier()
        td0 = time.perf_counter()
        for _ in range(ke):
            diet_step()
        torch.cuda.synchronize()
        diet_s = max_over_ranks((time.perf_counter() - td0) / ke)
        assert torch.equal(hqo2[:4096], query[:4096].cpu()), "query_grad host path and device path disagree"
        e2e = {"value": world * Be * EVALS_PER_ROW / e2e_s, "unit": "evals/s",
               "h2d_bytes_per_step": Be * (80 + 4 + 400 + 4), "d2h_bytes_per_step": Be * (400 + 4 + 80),
               "rows_per_gpu": Be, "ms_per_step": e2e_s * 1e3, "cpu_affinity": numa,
               "link": {"achieved_GBs_each_way": Be * 488 / e2e_s / 1e9, "pinned_copy_both_ways_GBs_each_way": link_gbs,
                        "frac": (Be * 488 / e2e_s / 1e9) / link_gbs,
                        "what": "PCIe-bound: bytes each way / step time against plain pinned cudaMemcpyAsync "
                                "running H2D and D2H concurrently on this box"},
               "api": "vael_circuit_fwdbwd_host (C-ABI, pinned host buffers, 256Ki-row chunks over 4 slots, uploads on their own streams)",
               "query_grad_rows_per_sec": world * Be / diet_s, "query_grad_ms_per_step": diet_s * 1e3,
               "query_grad_h2d_bytes_per_step": Be * (80 + 4 + 4), "query_grad_d2h_bytes_per_step": Be * (4 + 80),
               "query_grad_vs_full_rows_per_sec": e2e_s / diet_s,
               "query_grad_api": "vael_circuit_query_grad_host: facts + query index + grad_query in, query + grad_facts "
                                 "out (what utils/mnist_utils/train.py:123-147 consumes); no [B,W] tensor crosses PCIe"}

    # ---- secondary probe: VAEL train step (configs[1]), DDP when N > 1 --------------------------------
    train = None
    if not args.no_train:
        try:
            train = train_probe(args, dev, rank, world, distributed, barrier, max_over_ranks)
        except Exception as exc:  # the probe must never cost the headline line
            train = {"error": repr(exc)[:200]}

    # ---- secondary workloads (same kernels' siblings; informational, never the headline) ---------------------
    secondary = None
    if not args.no_secondary:
        try:
            secondary = secondary_probe(args, dev, rank, world, barrier, max_over_ranks, lib)
        except Exception as exc:
            secondary = {"error": repr(exc)[:200]}

    # ---- CPU baseline beside it (rank 0, N = 1 only) ------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        host_threads()
        cpu_chunks = max(1, args.rows // args.cpu_rows)      # one CPU step = this arm's rows per step
        rate, dt, kind = cpu_reference_rate(args.cpu_rows, 6, 1, cpu_chunks)
        cpu = {"value": rate, "unit": "evals/s", "cores": torch.get_num_threads(), "kind": kind,
               "sample": f"{cpu_chunks} batches x {args.cpu_rows} rows per step x 6 steps ({dt:.2f} s/step, "
                         f"{7 * dt:.0f} s of CPU work), torch CPU fp32, "
                         + ("the reference's own models/vael.py:208-225 (baseline/_ref) + autograd" if kind == "reference"
                            else "oracle port of models/vael.py:208-225 + autograd")
                         + f", os.cpu_count()={os.cpu_count()}"}
        try:
            cpu["also"] = cpu_extras()
        except Exception as exc:
            cpu["also"] = {"error": repr(exc)[:200]}

    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        kern = [{"name": "ad2_forward_kernel" if kernel_family == "ad2-tma" else "generic_forward_kernel",
                 "ms": fwd_ms, "bytes_per_row": FWD_BYTES, "GBs": B * FWD_BYTES / fwd_ms / 1e6},
                {"name": "ad2_backward_kernel" if kernel_family == "ad2-tma" else "generic_backward_kernel",
                 "ms": bwd_ms, "bytes_per_row": BWD_BYTES, "GBs": B * BWD_BYTES / bwd_ms / 1e6}]
        dom = max(kern, key=lambda k: k["ms"])
        # What this box streams in the same run (torch kernels over 2 GiB): the driver's MEASURED_PEAKS figure is a COPY
        # (read + write) bandwidth; pure stores run faster than that, and so do the non-persistent circuit kernels.
        box_probe = None
        try:
            pa = torch.empty(1 << 29, dtype=torch.float32, device=dev); pb = torch.empty_like(pa)

            def _gbs(fn, nbytes):
                for _ in range(2):
                    fn()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(5):
                    fn()
                e1.record(); torch.cuda.synchronize()
                return nbytes / (e0.elapsed_time(e1) / 5) / 1e6
            box_probe = {"fill_GBs": _gbs(lambda: pa.fill_(1.0), pa.numel() * 4),
                         "copy_GBs_read_plus_write": _gbs(lambda: pb.copy_(pa), 2 * pa.numel() * 4),
                         "what": "torch fill_ / copy_ over 2 GiB in this run"}
            del pa, pb
        except Exception as exc:   # noqa: BLE001
            box_probe = {"error": repr(exc)[:200]}
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            # per-launch dram__bytes_read.sum + dram__bytes_write.sum of the same kernel from the committed
            # `ncu --set full` capture (profiles/*_ncu.md); only quoted when that capture used this row count
            t = json.load(open(tpath)).get(dom["name"]) or {}
            traffic = t.get("bytes") if t.get("rows") == B else None
        line = {
            "metric": "circuit_evals_per_sec", "value": value, "unit": "evals/s", "n_gpus": world, "steps": K,
            "warmup": W, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(B, args.train_batch, "inputs larger than L2 (per-GPU step traffic "
                                      f"{B * (FWD_BYTES + BWD_BYTES) / 1e9:.2f} GB vs 126 MB L2)",
                                      f"rows sharded x{world}, circuit replicated, no collective on the data path; "
                                      "train step: DDP-style data parallel, one NCCL all-reduce of the flat gradient",
                                      train if isinstance(train, dict) and "value" in train else None),
            "kernel_family": kernel_family,
            "roofline": {"bound": "hbm", "kernel": dom["name"], "kernel_family": kernel_family, "achieved": dom["GBs"], "peak": peak, "unit": "GB/s",
                         "frac": dom["GBs"] / peak, "traffic": traffic, "peak_source": peak_src, "kernels": kern,
                         "algorithmic_bytes_per_launch": B * dom["bytes_per_row"],
                         "frac_of_nominal_7700": dom["GBs"] / 7700.0, "box_probe": box_probe,
                         "note": "peak = the driver's measured COPY bandwidth; frac > 1 means the kernel streams faster than "
                                 "that copy kernel does (non-persistent grid: scripts/micro/storebench.cu measures 7.5 TB/s "
                                 "for pure stores on this GPU; nominal HBM3e 7.7 TB/s)"},
            "e2e": e2e,
            "gpu_launches": int(launches), "clocks": clocks, "cpu_baseline": cpu, "train": train,
            "secondary": secondary,
        }
    if distributed:
        # the JSON line goes LAST: with NCCL_DEBUG=INFO on stdout the communicator teardown logs too
        dist.barrier()
        dist.destroy_process_group()
        if rank == 0:
            time.sleep(1.5)   # the other ranks' teardown lines are out
    if rank == 0:
        print(json.dumps(line), flush=True)


def _time_pair(fwd, bwd, n, barrier, max_over_ranks):
    for _ in range(3):
        fwd(); bwd()
    barrier()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * n + 1)]
    ev[0].record()
    for k in range(n):
        fwd(); ev[2 * k + 1].record()
        bwd(); ev[2 * k + 2].record()
    barrier()
    f = sum(ev[2 * k].elapsed_time(ev[2 * k + 1]) for k in range(n)) / n
    b = sum(ev[2 * k + 1].elapsed_time(ev[2 * k + 2]) for k in range(n)) / n
    return max_over_ranks(f), max_over_ranks(b)


def secondary_probe(args, dev, rank, world, barrier, max_over_ranks, lib):
    """configs[4] (3-digit addition, W = 1000: the wide-level TMA 2-D kernels), the generic level-by-level
    CSR kernel on the 2-digit circuit, and the fused all-queries kernel; per-GPU rows fixed (weak scaling)."""
    from vael_b200 import _lib, parallel
    from vael_b200.mnist_task import compile_addition_family
    from vael_b200.ops import DeviceCircuit, _p
    sp = torch.cuda.current_stream().cuda_stream
    peak_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak = float(json.load(open(peak_path))["hbm_gbs"]) if os.path.exists(peak_path) else 6650.0
    out = {}

    # -- configs[3]: the 2-digit circuit at batch 1e5 .. 1e7 (the headline runs 8.4e6).  At 1e5 rows the whole
    # working set (105 MB) sits in the 126 MB L2: that point is an L2-bandwidth number and is labelled so.
    sweep = {}
    dc2s = DeviceCircuit(compile_addition_family(2, 10).circuit, dev)
    for Bs in (100_000, 1_000_000, 10_000_000):
        fs, gs = synthetic_facts(Bs, dev, parallel.rank_seed(77, rank))
        qs = torch.randint(0, 19, (Bs,), device=dev, dtype=torch.int32, generator=gs)
        ws = torch.empty(Bs, 100, device=dev); qo = torch.empty(Bs, device=dev); gfs = torch.empty(Bs, 20, device=dev)
        gws = torch.randn(Bs, 100, device=dev, generator=gs); gqs = torch.randn(Bs, device=dev, generator=gs)
        f_ms, b_ms = _time_pair(
            lambda: _lib.check(lib.vael_circuit_forward(dc2s.handle, _p(fs), _p(qs), -1, Bs, _p(ws), _p(qo), 0, sp)),
            lambda: _lib.check(lib.vael_circuit_backward(dc2s.handle, _p(fs), _p(qs), -1, Bs, _p(gws), _p(gqs), _p(gfs), 0, sp)),
            10, barrier, max_over_ranks)
        sweep[f"B{Bs}"] = {"fwd_ms": f_ms, "bwd_ms": b_ms, "fwd_GBs": Bs * FWD_BYTES / f_ms / 1e6,
                           "bwd_GBs": Bs * BWD_BYTES / b_ms / 1e6,
                           "evals_per_sec": world * Bs * EVALS_PER_ROW / ((f_ms + b_ms) * 1e-3),
                           "memory": "L2-resident working set" if Bs * (FWD_BYTES + BWD_BYTES) < 120e6 else "HBM"}
        del fs, ws, gws
        torch.cuda.empty_cache()
    out["batch_sweep_2digit"] = sweep

    # -- 3-digit addition: F=30, W=1000, Q=28 -> fwd 4128 B/row, bwd 4248 B/row, 1001 evals/row
    B = args.wide_rows
    dc3 = DeviceCircuit(compile_addition_family(3, 10).circuit, dev)
    g = torch.Generator(device=dev).manual_seed(parallel.rank_seed(4321, rank))
    p = torch.softmax(3.0 * torch.randn(B, 3, 10, device=dev, generator=g), -1) + 1e-5
    facts = (p / p.sum(-1, keepdim=True)).reshape(B, 30).contiguous()
    qidx = torch.randint(0, 28, (B,), device=dev, dtype=torch.int32, generator=g)
    gw = torch.randn(B, 1000, device=dev, generator=g); gq = torch.randn(B, device=dev, generator=g)
    worlds = torch.empty(B, 1000, device=dev); query = torch.empty(B, device=dev); gf = torch.empty(B, 30, device=dev)
    f_ms, b_ms = _time_pair(
        lambda: _lib.check(lib.vael_circuit_forward(dc3.handle, _p(facts), _p(qidx), -1, B, _p(worlds), _p(query), 0, sp)),
        lambda: _lib.check(lib.vael_circuit_backward(dc3.handle, _p(facts), _p(qidx), -1, B, _p(gw), _p(gq), _p(gf), 0, sp)),
        10, barrier, max_over_ranks)
    out["addition_3digit"] = {
        "workload": "3digit-addition compiled circuit (F=30, W=1000, Q=28), fwd+bwd, per-row query, fp32",
        "rows_per_gpu": B, "kernel_family": _lib.last_kernel(), "evals_per_sec": world * B * 1001 / ((f_ms + b_ms) * 1e-3),
        "fwd": {"ms": f_ms, "bytes_per_row": 4128, "GBs": B * 4128 / f_ms / 1e6, "frac": B * 4128 / f_ms / 1e6 / peak},
        "bwd": {"ms": b_ms, "bytes_per_row": 4248, "GBs": B * 4248 / b_ms / 1e6, "frac": B * 4248 / b_ms / 1e6 / peak}}
    qa3 = torch.empty(B, 28, device=dev)
    a3_ms, _ = _time_pair(lambda: _lib.check(lib.vael_circuit_queries_all(dc3.handle, _p(facts), B, _p(qa3), 0, sp)),
                          lambda: None, 10, barrier, max_over_ranks)
    out["all_queries_3digit"] = {"workload": "all 28 sums per row of the 3-digit circuit (worlds not materialised)",
                                 "rows_per_gpu": B, "kernel_family": _lib.last_kernel(), "ms": a3_ms,
                                 "bytes_per_row": 120 + 112, "GBs": B * 232 / a3_ms / 1e6, "frac": B * 232 / a3_ms / 1e6 / peak,
                                 "rows_per_sec": world * B / (a3_ms * 1e-3)}
    del qa3
    Bg = B // 4   # the same circuit through the generic CSR kernels (1163 nodes: R = 1 tiles, chunked staging)
    f_ms, b_ms = _time_pair(
        lambda: _lib.check(lib.vael_circuit_forward(dc3.handle, _p(facts), _p(qidx), -1, Bg, _p(worlds), _p(query), 4, sp)),
        lambda: _lib.check(lib.vael_circuit_backward(dc3.handle, _p(facts), _p(qidx), -1, Bg, _p(gw), _p(gq), _p(gf), 4, sp)),
        5, barrier, max_over_ranks)
    out["generic_csr_3digit"] = {
        "workload": "3-digit circuit through the level-ordered CSR interpreter (VAEL_FLAG_FORCE_GENERIC)", "rows_per_gpu": Bg,
        "fwd": {"ms": f_ms, "GBs": Bg * 4128 / f_ms / 1e6, "frac": Bg * 4128 / f_ms / 1e6 / peak},
        "bwd": {"ms": b_ms, "GBs": Bg * 4248 / b_ms / 1e6, "frac": Bg * 4248 / b_ms / 1e6 / peak}}
    try:   # the same circuit through its circuit-specialised straight-line kernels (NVRTC, VAEL_FLAG_FORCE_JIT)
        f_ms, b_ms = _time_pair(
            lambda: _lib.check(lib.vael_circuit_forward(dc3.handle, _p(facts), _p(qidx), -1, Bg, _p(worlds), _p(query), 16, sp)),
            lambda: _lib.check(lib.vael_circuit_backward(dc3.handle, _p(facts), _p(qidx), -1, Bg, _p(gw), _p(gq), _p(gf), 16, sp)),
            5, barrier, max_over_ranks)
        out["jit_3digit"] = {
            "workload": "3-digit circuit through its generated straight-line kernels (column chunks of 100 worlds)",
            "rows_per_gpu": Bg, "kernel_family": _lib.last_kernel(),
            "fwd": {"ms": f_ms, "GBs": Bg * 4128 / f_ms / 1e6, "frac": Bg * 4128 / f_ms / 1e6 / peak},
            "bwd": {"ms": b_ms, "GBs": Bg * 4248 / b_ms / 1e6, "frac": Bg * 4248 / b_ms / 1e6 / peak}}
    except Exception as exc:  # noqa: BLE001
        out["jit_3digit"] = {"error": repr(exc)[:300]}
    del worlds, gw, facts, p
    torch.cuda.empty_cache()

    # -- generic level-by-level CSR kernel (VAEL_FLAG_FORCE_GENERIC) and all-queries on the 2-digit circuit
    B2 = 1 << 21
    dc2 = DeviceCircuit(compile_addition_family(2, 10).circuit, dev)
    facts2, g2 = synthetic_facts(B2, dev, parallel.rank_seed(99, rank))
    q2 = torch.randint(0, 19, (B2,), device=dev, dtype=torch.int32, generator=g2)
    gw2 = torch.randn(B2, 100, device=dev, generator=g2); gq2 = torch.randn(B2, device=dev, generator=g2)
    w2 = torch.empty(B2, 100, device=dev); qo2 = torch.empty(B2, device=dev); gf2 = torch.empty(B2, 20, device=dev)
    f_ms, b_ms = _time_pair(
        lambda: _lib.check(lib.vael_circuit_forward(dc2.handle, _p(facts2), _p(q2), -1, B2, _p(w2), _p(qo2), 4, sp)),
        lambda: _lib.check(lib.vael_circuit_backward(dc2.handle, _p(facts2), _p(q2), -1, B2, _p(gw2), _p(gq2), _p(gf2), 4, sp)),
        10, barrier, max_over_ranks)
    out["generic_csr_2digit"] = {
        "workload": "2-digit circuit through the level-ordered CSR interpreter (VAEL_FLAG_FORCE_GENERIC)", "rows_per_gpu": B2,
        "fwd": {"ms": f_ms, "GBs": B2 * FWD_BYTES / f_ms / 1e6, "frac": B2 * FWD_BYTES / f_ms / 1e6 / peak},
        "bwd": {"ms": b_ms, "GBs": B2 * BWD_BYTES / b_ms / 1e6, "frac": B2 * BWD_BYTES / b_ms / 1e6 / peak}}
    try:   # what a circuit WITHOUT a recognised shape runs: its generated straight-line kernels (here forced)
        f_ms, b_ms = _time_pair(
            lambda: _lib.check(lib.vael_circuit_forward(dc2.handle, _p(facts2), _p(q2), -1, B2, _p(w2), _p(qo2), 16, sp)),
            lambda: _lib.check(lib.vael_circuit_backward(dc2.handle, _p(facts2), _p(q2), -1, B2, _p(gw2), _p(gq2), _p(gf2), 16, sp)),
            10, barrier, max_over_ranks)
        out["jit_2digit"] = {
            "workload": "2-digit circuit through its generated straight-line kernels (NVRTC; the path of every circuit "
                        "without a recognised shape)", "rows_per_gpu": B2, "kernel_family": _lib.last_kernel(),
            "fwd": {"ms": f_ms, "GBs": B2 * FWD_BYTES / f_ms / 1e6, "frac": B2 * FWD_BYTES / f_ms / 1e6 / peak},
            "bwd": {"ms": b_ms, "GBs": B2 * BWD_BYTES / b_ms / 1e6, "frac": B2 * BWD_BYTES / b_ms / 1e6 / peak}}
    except Exception as exc:  # noqa: BLE001
        out["jit_2digit"] = {"error": repr(exc)[:300]}
    del w2, gw2, facts2
    torch.cuda.empty_cache()

    # -- Mario circuits (configs[2]): worlds / move query over 2 x 9 position facts (prob semiring, ad2<9,9>) and
    # the 24 admissible worlds' log-probabilities (log semiring: products of 18 positive / negative literals,
    # the literal-table kernel)
    try:
        from vael_b200.mario_task import admissible_circuit, compile_mario_family, mario_tables
        Bm = 1 << 21
        fam = compile_mario_family((3, 3))
        dcm = DeviceCircuit(fam.circuit, dev)
        dca = DeviceCircuit(admissible_circuit(mario_tables(fam)[2]), dev)
        gm = torch.Generator(device=dev).manual_seed(parallel.rank_seed(555, rank))
        pm = torch.softmax(torch.randn(Bm, 2, 9, device=dev, generator=gm), -1).clamp(1e-5, 0.9999).reshape(Bm, 18).contiguous()
        qm = torch.randint(0, 4, (Bm,), device=dev, dtype=torch.int32, generator=gm)
        wm = torch.empty(Bm, 81, device=dev); qom = torch.empty(Bm, device=dev); gfm = torch.empty(Bm, 18, device=dev)
        gwm = torch.randn(Bm, 81, device=dev, generator=gm); gqm = torch.randn(Bm, device=dev, generator=gm)
        f_ms, b_ms = _time_pair(
            lambda: _lib.check(lib.vael_circuit_forward(dcm.handle, _p(pm), _p(qm), -1, Bm, _p(wm), _p(qom), 0, sp)),
            lambda: _lib.check(lib.vael_circuit_backward(dcm.handle, _p(pm), _p(qm), -1, Bm, _p(gwm), _p(gqm), _p(gfm), 0, sp)),
            10, barrier, max_over_ranks)
        fam_m = _lib.last_kernel()
        fb, bb = 72 + 4 + 324 + 4, 324 + 4 + 72 + 4 + 72
        la = torch.empty(Bm, 24, device=dev); gla = torch.randn(Bm, 24, device=dev, generator=gm)
        fa_ms, ba_ms = _time_pair(
            lambda: _lib.check(lib.vael_circuit_forward(dca.handle, _p(pm), None, -1, Bm, _p(la), None, 0, sp)),
            lambda: _lib.check(lib.vael_circuit_backward(dca.handle, _p(pm), None, -1, Bm, _p(gla), None, _p(gfm), 0, sp)),
            10, barrier, max_over_ranks)
        fab, bab = 72 + 96, 96 + 72 + 72
        out["mario_circuits"] = {
            "workload": "Mario 3x3: worlds[81] + move query (prob semiring) and admissible-world log-probabilities[24] "
                        "(log semiring), fwd+bwd, fp32", "rows_per_gpu": Bm,
            "worlds_query": {"kernel_family": fam_m,
                             "fwd": {"ms": f_ms, "bytes_per_row": fb, "GBs": Bm * fb / f_ms / 1e6, "frac": Bm * fb / f_ms / 1e6 / peak},
                             "bwd": {"ms": b_ms, "bytes_per_row": bb, "GBs": Bm * bb / b_ms / 1e6, "frac": Bm * bb / b_ms / 1e6 / peak}},
            "admissible_logprob": {"kernel_family": _lib.last_kernel(),
                                   "fwd": {"ms": fa_ms, "bytes_per_row": fab, "GBs": Bm * fab / fa_ms / 1e6, "frac": Bm * fab / fa_ms / 1e6 / peak},
                                   "bwd": {"ms": ba_ms, "bytes_per_row": bab, "GBs": Bm * bab / ba_ms / 1e6, "frac": Bm * bab / ba_ms / 1e6 / peak}}}
        qam = torch.empty(Bm, 5, device=dev)
        qa_ms, _ = _time_pair(lambda: _lib.check(lib.vael_circuit_queries_all(dcm.handle, _p(pm), Bm, _p(qam), 0, sp)),
                              lambda: None, 10, barrier, max_over_ranks)
        out["mario_circuits"]["all_queries"] = {
            "workload": "the four move queries + constraint of every row (worlds not materialised): the circuit's generated "
                        "kernel where NVRTC is available, else the hand-written member-list kernel",
            "kernel_family": _lib.last_kernel(), "ms": qa_ms, "bytes_per_row": 72 + 20, "GBs": Bm * 92 / qa_ms / 1e6,
            "frac": Bm * 92 / qa_ms / 1e6 / peak}
        os.environ["VAEL_QUERIES_MEMBERLIST"] = "1"   # read per call: the hand-written kernel on the same rows
        try:
            ql_ms, _ = _time_pair(lambda: _lib.check(lib.vael_circuit_queries_all(dcm.handle, _p(pm), Bm, _p(qam), 0, sp)),
                                  lambda: None, 10, barrier, max_over_ranks)
            out["mario_circuits"]["all_queries"]["member_list_kernel"] = {
                "kernel_family": _lib.last_kernel(), "ms": ql_ms, "GBs": Bm * 92 / ql_ms / 1e6, "frac": Bm * 92 / ql_ms / 1e6 / peak}
        finally:
            del os.environ["VAEL_QUERIES_MEMBERLIST"]
        del wm, gwm, la, gla, pm, qam
        torch.cuda.empty_cache()
    except Exception as exc:  # noqa: BLE001 - a secondary probe must not take the headline down
        out["mario_circuits"] = {"error": repr(exc)[:300]}

    # -- the training-mode logic path (models/vael.py:55-65): ONE fused kernel per direction against the three-stage
    # chain it replaces (facts head -> circuit -> Gumbel sampling), both with in-kernel Philox noise
    try:
        out["stage_kernels"] = stage_kernels_probe(dev, rank, barrier, max_over_ranks, lib, peak)
    except Exception as exc:   # noqa: BLE001
        out["stage_kernels"] = {"error": repr(exc)[:300]}
    try:
        out["logic_train_path"] = logic_path_probe(dev, rank, barrier, max_over_ranks, lib, peak)
    except Exception as exc:  # noqa: BLE001
        out["logic_train_path"] = {"error": repr(exc)[:300]}

    facts2, g2 = synthetic_facts(B2, dev, parallel.rank_seed(99, rank))
    qa = torch.empty(B2, 19, device=dev)
    a_ms, _ = _time_pair(lambda: _lib.check(lib.vael_circuit_queries_all(dc2.handle, _p(facts2), B2, _p(qa), 0, sp)),
                         lambda: None, 10, barrier, max_over_ranks)
    out["all_queries_2digit"] = {"workload": "all 19 sums per row (worlds @ w_q fused, worlds not materialised)",
                                 "rows_per_gpu": B2, "kernel_family": _lib.last_kernel(), "ms": a_ms,
                                 "bytes_per_row": 80 + 76, "GBs": B2 * 156 / a_ms / 1e6, "frac": B2 * 156 / a_ms / 1e6 / peak,
                                 "rows_per_sec": world * B2 / (a_ms * 1e-3)}
    return out


def stage_kernels_probe(dev, rank, barrier, max_over_ranks, lib, peak):
    """The stand-alone kernels around the circuit (a14 facts head, a15/a16 Gumbel + Herbrand, a26 ELBO terms), each
    alone at a streaming size, against the copy peak: algorithmic bytes per row / launch time."""
    from vael_b200 import _lib, parallel
    from vael_b200.ops import _p
    sp = torch.cuda.current_stream().cuda_stream
    g = torch.Generator(device=dev).manual_seed(parallel.rank_seed(41, rank))
    out = {"workload": "facts head [B,2x10] (fwd, adjoint), Gumbel-softmax(hard) + Herbrand [B,100] -> [B,20] (fwd with "
                       "in-kernel Philox noise, adjoint), ELBO terms + gradients [B,1,28,56]; fp32, per-GPU rows fixed"}

    def entry(ms, rows, bpr):
        return {"ms": ms, "rows_per_gpu": rows, "bytes_per_row": bpr, "GBs": rows * bpr / ms / 1e6,
                "frac": rows * bpr / ms / 1e6 / peak}

    B = 1 << 21
    x, gfa = torch.randn(B, 20, device=dev, generator=g), torch.randn(B, 20, device=dev, generator=g)
    fo, gx = torch.empty(B, 20, device=dev), torch.empty(B, 20, device=dev)
    f_ms, b_ms = _time_pair(lambda: _lib.check(lib.vael_facts_forward(_p(x), B, 2, 10, 1e-5, _p(fo), sp)),
                            lambda: _lib.check(lib.vael_facts_backward(_p(x), _p(gfa), B, 2, 10, 1e-5, _p(gx), sp)),
                            10, barrier, max_over_ranks)
    out["facts_head"] = {"kernel": "facts_tile_kernel", "fwd": entry(f_ms, B, 160), "bwd": entry(b_ms, B, 240)}
    del x, gfa, fo, gx
    B = 1 << 20
    p = torch.softmax(torch.randn(B, 100, device=dev, generator=g), 1)
    hb = torch.cat([torch.eye(10).repeat_interleave(10, 0), torch.eye(10).repeat(10, 1)], 1).to(dev).contiguous()
    ys, idx, wh = torch.empty(B, 100, device=dev), torch.empty(B, device=dev, dtype=torch.int32), torch.empty(B, 20, device=dev)
    gh, gs = torch.randn(B, 20, device=dev, generator=g), torch.empty(B, 100, device=dev)
    f_ms, b_ms = _time_pair(
        lambda: _lib.check(lib.vael_gumbel_world_forward(_p(p), None, 7, B, 100, 0, 1.0, _p(hb), 20, _p(ys), _p(idx), _p(wh), sp)),
        lambda: _lib.check(lib.vael_gumbel_world_backward(_p(p), _p(ys), _p(gh), B, 100, 0, 1.0, _p(hb), 20, _p(gs), sp)),
        10, barrier, max_over_ranks)
    out["gumbel_herbrand"] = {"kernel": "gumbel_rows_forward_kernel / gumbel_rows_backward_kernel (lane <-> row; from 16 Ki rows up)",
                              "fwd": entry(f_ms, B, 884), "bwd": entry(b_ms, B, 1284),
                              "bound": "forward: instruction issue (Philox4x32-10 + variate transforms); adjoint: resident warps"}
    del p, ys, idx, wh, gh, gs
    B = 65536
    r, xi = torch.rand(B, 1, 28, 56, device=dev, generator=g), torch.rand(B, 1, 28, 56, device=dev, generator=g)
    mu, lv = torch.randn(B, 23, device=dev, generator=g), torch.randn(B, 23, device=dev, generator=g)
    q = torch.rand(B, device=dev, generator=g)
    terms, gr, gm, glv, gq = torch.empty(4, device=dev), torch.empty_like(r), torch.empty_like(mu), torch.empty_like(lv), torch.empty_like(q)
    ws = torch.zeros(int(lib.vael_elbo_workspace_bytes()), dtype=torch.uint8, device=dev)
    e_ms, _ = _time_pair(lambda: _lib.check(lib.vael_elbo_terms(_p(r), _p(xi), B, 1568, _p(mu), _p(lv), 23, _p(q), 0.1, 1e-5, 1.0, 0,
                                                                _p(terms), _p(gr), _p(gm), _p(glv), _p(gq), _p(ws), sp)),
                         lambda: None, 10, barrier, max_over_ranks)
    out["elbo_terms"] = {"kernel": "elbo_terms_kernel", "terms_and_grads": entry(e_ms, B, 1568 * 12)}
    return out


def logic_path_probe(dev, rank, barrier, max_over_ranks, lib, peak, Bl=1 << 20):
    from vael_b200 import _lib, parallel
    from vael_b200.mnist_task import compile_addition_family
    from vael_b200.ops import DeviceCircuit, _p
    sp = torch.cuda.current_stream().cuda_stream
    dc = DeviceCircuit(compile_addition_family(2, 10).circuit, dev)
    g = torch.Generator(device=dev).manual_seed(parallel.rank_seed(31, rank))
    logits = 3.0 * torch.randn(Bl, 20, device=dev, generator=g)
    qidx = torch.randint(0, 19, (Bl,), device=dev, dtype=torch.int32, generator=g)
    gh = torch.randn(Bl, 20, device=dev, generator=g); gq = torch.randn(Bl, device=dev, generator=g)
    facts = torch.empty(Bl, 20, device=dev); query = torch.empty(Bl, device=dev); wh = torch.empty(Bl, 20, device=dev)
    idx = torch.empty(Bl, device=dev, dtype=torch.int32); gl = torch.empty(Bl, 20, device=dev)
    f_ms, b_ms = _time_pair(
        lambda: _lib.check(lib.vael_logic_train_forward(dc.handle, _p(logits), _p(qidx), -1, Bl, 1e-5, 1.0, None, 1234, None,
                                                        _p(facts), _p(query), _p(idx), _p(wh), None, sp)),
        lambda: _lib.check(lib.vael_logic_train_backward(dc.handle, _p(logits), _p(qidx), -1, Bl, 1e-5, 1.0, None, 1234, None,
                                                         _p(gh), _p(gq), None, _p(gl), sp)),
        10, barrier, max_over_ranks)
    fb, bb = 80 + 4 + 80 + 4 + 4 + 80, 80 + 4 + 80 + 4 + 80
    # the stage kernels on the same rows
    worlds = torch.empty(Bl, 100, device=dev); ys = torch.empty(Bl, 100, device=dev); gs = torch.empty(Bl, 100, device=dev)
    gf = torch.empty(Bl, 20, device=dev)
    hb = torch.cat([torch.eye(10).repeat_interleave(10, 0), torch.eye(10).repeat(10, 1)], 1).to(dev).contiguous()

    def stage_fwd():
        _lib.check(lib.vael_facts_forward(_p(logits), Bl, 2, 10, 1e-5, _p(facts), sp))
        _lib.check(lib.vael_circuit_forward(dc.handle, _p(facts), _p(qidx), -1, Bl, _p(worlds), _p(query), 0, sp))
        _lib.check(lib.vael_gumbel_world_forward(_p(worlds), None, 1234, Bl, 100, 0, 1.0, _p(hb), 20, _p(ys), _p(idx), _p(wh), sp))

    def stage_bwd():
        _lib.check(lib.vael_gumbel_world_backward(_p(worlds), _p(ys), _p(gh), Bl, 100, 0, 1.0, _p(hb), 20, _p(gs), sp))
        _lib.check(lib.vael_circuit_backward(dc.handle, _p(facts), _p(qidx), -1, Bl, _p(gs), _p(gq), _p(gf), 0, sp))
        _lib.check(lib.vael_facts_backward(_p(logits), _p(gf), Bl, 2, 10, 1e-5, _p(gl), sp))

    sf_ms, sb_ms = _time_pair(stage_fwd, stage_bwd, 10, barrier, max_over_ranks)
    sfb = (80 + 80) + (80 + 4 + 400 + 4) + (400 + 400 + 4 + 80)              # facts head, circuit, gumbel
    sbb = (400 + 400 + 80 + 400) + (400 + 4 + 80 + 4 + 80) + (80 + 80 + 80)  # gumbel, circuit, facts head adjoints
    return {"workload": "2digit VAEL logic path in training mode: logits[20] -> facts -> worlds -> query -> Gumbel arg-max "
                        "-> Herbrand vector, and its adjoint; in-kernel Philox noise, per-row query", "rows_per_gpu": Bl,
            "fused": {"fwd": {"ms": f_ms, "bytes_per_row": fb, "GBs": Bl * fb / f_ms / 1e6, "frac_hbm": Bl * fb / f_ms / 1e6 / peak,
                              "rows_per_sec": Bl / (f_ms * 1e-3)},
                      "bwd": {"ms": b_ms, "bytes_per_row": bb, "GBs": Bl * bb / b_ms / 1e6, "frac_hbm": Bl * bb / b_ms / 1e6 / peak,
                              "rows_per_sec": Bl / (b_ms * 1e-3)},
                      "bound": "forward: HBM (the world is drawn top-down: two uniforms per row, ~550 instructions); adjoint: "
                               "FP32/SFU issue (25 Philox4x32-10 calls, 100 lg2 + 100 rcp per row)"},
            "three_stage_kernels": {"fwd": {"ms": sf_ms, "bytes_per_row": sfb, "GBs": Bl * sfb / sf_ms / 1e6},
                                    "bwd": {"ms": sb_ms, "bytes_per_row": sbb, "GBs": Bl * sbb / sb_ms / 1e6}},
            "speedup": {"fwd": sf_ms / f_ms, "bwd": sb_ms / b_ms}}


def train_probe(args, dev, rank, world, distributed, barrier, max_over_ranks):
    """VAEL train samples/s: BASELINE.json configs[1] (2digit-MNIST addition, batch 4096 per GPU, data parallel
    when N > 1) and configs[0] (the B = 32 reference configuration): encoder/decoder in PyTorch, the logic path
    on the fused vael_b200 kernel, Adam; the step is one CUDA graph at N = 1 and two graphs around one NCCL
    all-reduce of the flat gradient at N > 1 (eager DistributedDataParallel timed beside it)."""
    from vael_b200.mnist_task import build_model_dict, build_worlds_queries_matrix
    from vael_b200.networks import MNISTPairsDecoder, MNISTPairsEncoder, MNISTPairsMLP
    from vael_b200.train import GraphedTrainStep, as_channels_last, train_step
    from vael_b200.vael import MNISTPairsVAELModel
    from vael_b200 import parallel
    model_dict = build_model_dict(2, 10, device=dev)
    w_q = build_worlds_queries_matrix(2, 10).to(dev)

    def run(Bt, graphed, n):
        torch.manual_seed(0)
        model = MNISTPairsVAELModel(MNISTPairsEncoder(hidden_channels=64, latent_dim=23),
                                    MNISTPairsDecoder(label_dim=20, hidden_channels=64, latent_dim=8),
                                    MNISTPairsMLP(in_features=15, n_facts=20), (15, 8), model_dict, w_q, dropout=0.5,
                                    is_train=True, device=dev, query_per_row=True).to(dev)
        # plumbing around the stock-PyTorch CNNs: NHWC activations (no NCHW<->NHWC shuffles around cuDNN's
        # kernels) and the fused multi-tensor Adam: 14.9 -> 12.2 ms/step at B = 4096 (scripts/train_variants.py)
        model = model.to(memory_format=torch.channels_last)
        opt = torch.optim.Adam(model.parameters(), lr=1e-3, capturable=graphed, fused=True)
        g = torch.Generator(device=dev).manual_seed(parallel.rank_seed(1234, rank))
        x = as_channels_last(torch.randn(Bt, 1, 28, 56, device=dev, generator=g))   # explicit NHWC strides for C = 1
        d = torch.randint(0, 10, (Bt, 2), device=dev, generator=g)
        labels = torch.cat([d, d.sum(1, keepdim=True), -torch.ones(Bt, 1, dtype=torch.long, device=dev)], 1)
        if graphed:
            # N = 1: one CUDA graph.  N > 1: forward+backward graph, ONE NCCL all-reduce of the flat 7.4 MB gradient
            # (train.FlatGradBucket), Adam graph
            stepper = GraphedTrainStep(model, opt, tuple(x.shape), tuple(labels.shape), device=dev,
                                       memory_format=torch.channels_last)
            step = lambda: stepper(x, labels)
        else:
            step_model = torch.nn.parallel.DistributedDataParallel(model, device_ids=[dev.index]) if distributed else model
            step = lambda: train_step(step_model, opt, x, labels)
        for _ in range(5):
            step()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            terms = step()
        e1.record()
        barrier()
        ms = max_over_ranks(e0.elapsed_time(e1)) / n
        mode = ("cuda-graph" if not distributed else "2 cuda graphs + 1 nccl all-reduce of the flat gradient") if graphed \
            else ("ddp-nccl eager" if distributed else "eager")
        return {"value": world * Bt / (ms * 1e-3), "unit": "samples/s", "ms_per_step": ms, "global_batch": world * Bt,
                "loss": float(terms[0].item()), "mode": mode}

    out = {"metric": "vael_train_samples_per_sec",
           "config": "2digit-MNIST addition VAEL, synthetic 28x56 images, random-init weights, Adam, fp32 "
                     "(cuDNN TF32 convolutions as in stock PyTorch), channels_last activations, fused Adam"}
    try:
        main = run(args.train_batch, graphed=True, n=20)
    except Exception as exc:   # noqa: BLE001 - fall back to the eager step rather than lose the number
        if not distributed:
            raise
        main = run(args.train_batch, graphed=False, n=20)
        main["graphed_error"] = repr(exc)[:200]
    out.update(main)
    out["eager"] = run(args.train_batch, graphed=False, n=20)
    if not distributed:
        out["reference_config_B32"] = {"cuda_graph": run(32, True, 200), "eager": run(32, False, 50)}
        try:
            out["mario_B32"] = mario_probe(dev)
        except Exception as exc:
            out["mario_B32"] = {"error": repr(exc)[:200]}
    return out


def mario_probe(dev, B=32, n=30):
    """BASELINE.json configs[2]: Mario grid-navigation VAEL (two 100x100 RGB frames, 2 x 9 agent-position
    facts, the compiled move-rule program: 81 worlds, 24 admissible, 4 move queries), reference batch 32."""
    from vael_b200.mario_task import compile_mario_family, mario_tables
    from vael_b200.networks import MarioDecoder, MarioEncoder, MarioMLP
    from vael_b200.train import mario_train_step
    from vael_b200.vael import MarioVAELModel
    torch.manual_seed(0)
    tabs = [torch.from_numpy(t.astype("float32")).to(dev) for t in mario_tables(compile_mario_family((3, 3)))]
    W_all, WQ_all, W_adm, WQ_adm = tabs
    model = MarioVAELModel(MarioEncoder(hidden_channels=64, latent_dim=48, dropout=0.0),
                           MarioDecoder(hidden_channels=64, latent_dim_sub=30, label_dim=9, dropout=0.5),
                           MarioMLP(18, 18, 32), (18, 30), None, WQ_all, W_all, WQ_adm, W_adm, (3, 3), dropout=0.5,
                           is_train=True, device=dev).to(dev)
    # the same plumbing as the MNIST probe around the stock-PyTorch CNNs: NHWC activations + fused multi-tensor Adam
    # (scripts/mario_step_variants.py, graphed B = 32: 4.77 ms NCHW / foreach Adam -> 3.28 ms)
    model = model.to(memory_format=torch.channels_last)
    opt = torch.optim.Adam(model.parameters(), lr=1e-4, fused=True)
    x = torch.rand(B, 3, 200, 100, device=dev).contiguous(memory_format=torch.channels_last)
    moves = torch.eye(4, dtype=torch.long, device=dev)[torch.randint(0, 4, (B,), device=dev)]
    for _ in range(5):
        mario_train_step(model, opt, x, moves)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        elbo = mario_train_step(model, opt, x, moves)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    eager = {"value": B / (ms * 1e-3), "unit": "samples/s", "ms_per_step": ms, "elbo": float(elbo.item()), "mode": "eager"}
    # the same step captured once in a CUDA graph (B = 32 is launch-bound)
    from vael_b200.train import GraphedTrainStep
    opt_g = torch.optim.Adam(model.parameters(), lr=1e-4, capturable=True, fused=True)
    stepper = GraphedTrainStep(model, opt_g, tuple(x.shape), tuple(moves.shape), device=dev, step_fn=mario_train_step,
                               memory_format=torch.channels_last)
    for _ in range(3):
        stepper(x, moves)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        elbo = stepper(x, moves)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    return {"value": B / (ms * 1e-3), "unit": "samples/s", "ms_per_step": ms, "elbo": float(elbo.item()), "mode": "cuda-graph",
            "eager": eager, "config": "Mario VAEL, synthetic 2x(100x100x3) frames, random-init weights, Adam (fused), "
                                      "channels_last activations"}


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
