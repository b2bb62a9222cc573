"""world_size-2 gloo tests (CPU) of the N>1 host logic: the max-over-ranks timing reduction bench.py reports,
the flat gradient bucket of the graphed data-parallel step (one all-reduce == DistributedDataParallel's
global-batch gradient for batch-mean losses), and the reference arm printing from rank 0 only."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from vael_b200.parallel import max_over_ranks, rank_seed
        from vael_b200.train import FlatGradBucket
        res = {}
        # timing: the slow rank defines the step
        res["max"] = max_over_ranks(1.0 + rank)
        # data: distinct per-rank seeds
        g = torch.Generator().manual_seed(rank_seed(1234, rank))
        res["first"] = float(torch.randn(1, generator=g))

        # the flat gradient bucket of the graphed data-parallel step (train.FlatGradBucket): parameters broadcast
        # from rank 0, every .grad a view of one buffer, ONE all-reduce -> the gradient of a batch-mean loss on
        # equal shards equals the global-batch gradient, as with DistributedDataParallel
        torch.manual_seed(100 + rank)                      # ranks start from DIFFERENT weights
        net = torch.nn.Sequential(torch.nn.Conv2d(1, 4, 3, padding=1), torch.nn.Flatten(), torch.nn.Linear(4 * 6 * 6, 5),
                                  torch.nn.Tanh(), torch.nn.Linear(5, 3)).to(memory_format=torch.channels_last)
        torch.manual_seed(100)                             # = rank 0's initialisation
        ref = torch.nn.Sequential(torch.nn.Conv2d(1, 4, 3, padding=1), torch.nn.Flatten(), torch.nn.Linear(4 * 6 * 6, 5),
                                  torch.nn.Tanh(), torch.nn.Linear(5, 3))
        bucket = FlatGradBucket(net.parameters())
        bucket.broadcast_parameters()
        res["bcast_err"] = max(float((a - b).abs().max()) for a, b in zip(net.parameters(), ref.parameters()))
        res["views"] = all(p.grad.untyped_storage().data_ptr() == bucket.flat.untyped_storage().data_ptr() and
                           p.grad.stride() == p.stride() for p in net.parameters())
        x = torch.randn(8, 1, 6, 6, generator=torch.Generator().manual_seed(7))
        per = 8 // world
        for _ in range(2):                                 # twice: zero() really clears the accumulated views
            bucket.zero()
            net(x[rank * per:(rank + 1) * per]).pow(2).sum(1).mean().backward()
            bucket.all_reduce_mean()
        ref(x).pow(2).sum(1).mean().backward()
        res["ddp_err"] = max(float((a.grad - b.grad).abs().max()) for a, b in zip(net.parameters(), ref.parameters()))
        res["same_storage_after_backward"] = all(
            p.grad.untyped_storage().data_ptr() == bucket.flat.untyped_storage().data_ptr() for p in net.parameters())
        out[rank] = res
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_reductions_and_flat_bucket_equivalence():
    world, port = 2, free_port()
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
        r0, r1 = out[0], out[1]
    assert r0["max"] == r1["max"] == 2.0
    assert r0["first"] != r1["first"]
    for r in (r0, r1):
        assert r["bcast_err"] == 0.0 and r["views"] and r["same_storage_after_backward"]
        assert r["ddp_err"] < 1e-6


def test_reference_arm_prints_one_line_from_rank_zero_only():
    """bench.py --impl reference under torchrun: rank 0 runs and prints, the others exit 0 silently."""
    port = free_port()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
           "--steps", "1", "--warmup", "1", "--cpu-rows", "2048", "--rows", "8192", "--train-batch", "8"]
    env = dict(os.environ, OMP_NUM_THREADS="2")
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [json.loads(l) for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = lines[0]
    assert d["impl"] == "reference" and d["metric"] == "circuit_evals_per_sec" and d["n_gpus"] == 2
    from tests import refimport
    assert d["cpu_baseline"]["kind"] == ("reference" if refimport.available() else "port")
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["value"] > 0
    assert set(d["config"]) == {"workload", "rows_per_gpu", "evals_per_row", "l2", "parallelism", "train_workload",
                                "train_batch_per_gpu", "train_samples_per_sec", "train_ms_per_step", "train_mode"}
    if refimport.available():
        assert d["config"]["train_samples_per_sec"] > 0
