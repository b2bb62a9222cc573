"""world_size-2 gloo tests (CPU) of the N>1 host logic: row sharding, the max-over-ranks timing
reduction and whole-job rate bench.py reports, DDP gradient averaging == global-batch gradient
for batch-mean losses, and the reference arm printing from rank 0 only."""
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


def test_shard_rows_tiles_the_batch():
    from vael_b200.parallel import shard_rows
    for n in (0, 1, 7, 4096, 8388608 + 5):
        for world in (1, 2, 3, 8):
            spans = [shard_rows(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == n
            for (s0, c0), (s1, _) in zip(spans, spans[1:]):
                assert s0 + c0 == s1
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
    with pytest.raises(ValueError):
        shard_rows(10, 2, 2)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from vael_b200.parallel import job_rate, max_over_ranks, rank_seed, shard_rows
        res = {}
        # timing: the slow rank defines the step; rate = all rows / slowest time
        res["max"] = max_over_ranks(1.0 + rank)
        res["rate"] = job_rate(1000.0 * (rank + 1), 0.5 * (rank + 1))
        # data: distinct per-rank seeds, disjoint shards
        g = torch.Generator().manual_seed(rank_seed(1234, rank))
        res["first"] = float(torch.randn(1, generator=g))
        res["shard"] = shard_rows(101, world, rank)

        # DDP over gloo: gradient of a batch-mean loss on equal shards == global-batch gradient
        torch.manual_seed(0)
        net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3))
        ref = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3))
        ref.load_state_dict(net.state_dict())
        ddp = torch.nn.parallel.DistributedDataParallel(net)
        x = torch.randn(8, 6, generator=torch.Generator().manual_seed(7))
        s, c = shard_rows(8, world, rank)
        ddp(x[s:s + c]).pow(2).sum(1).mean().backward()
        ref(x).pow(2).sum(1).mean().backward()
        res["ddp_err"] = max(float((a.grad - b.grad).abs().max()) for a, b in zip(net.parameters(), ref.parameters()))
        out[rank] = res
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_reductions_and_ddp_equivalence():
    world, port = 2, free_port()
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
        r0, r1 = out[0], out[1]
    assert r0["max"] == r1["max"] == 2.0
    assert r0["rate"] == r1["rate"] == pytest.approx(3000.0 / 1.0)
    assert r0["first"] != r1["first"]
    assert r0["shard"] == (0, 51) and r1["shard"] == (51, 50)
    assert r0["ddp_err"] < 1e-6 and r1["ddp_err"] < 1e-6


def test_reference_arm_prints_one_line_from_rank_zero_only():
    """bench.py --impl reference under torchrun: rank 0 runs and prints, the others exit 0 silently."""
    port = free_port()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
           "--steps", "1", "--warmup", "1", "--cpu-rows", "2048", "--rows", "8192"]
    env = dict(os.environ, OMP_NUM_THREADS="2")
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [json.loads(l) for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = lines[0]
    assert d["impl"] == "reference" and d["metric"] == "circuit_evals_per_sec" and d["n_gpus"] == 2
    assert d["cpu_baseline"]["kind"] == "port" and d["e2e"]["h2d_bytes_per_step"] == 0 and d["value"] > 0
