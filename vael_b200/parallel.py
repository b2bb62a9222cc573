"""Batch-row data parallelism for the logic path: one process per GPU, rows sharded, the
compiled circuit replicated, NO collective on the data path (every output row depends only on
its own fact row — SURVEY.md §8(e)).  torch.distributed is plumbing: process group, barrier and
the max-over-ranks reduction of device times that bench.py reports.  The only collective in
training is the NCCL all-reduce of the encoder/decoder gradients (train.FlatGradBucket: one flat
7.4 MB buffer between the two CUDA graphs of the step); the circuit has no parameters, and
extract_worlds_probability's whole-batch normaliser (models/vael.py:263-266) stays shard-local.
"""
from __future__ import annotations

import os
from typing import Optional, Tuple

import torch
import torch.distributed as dist


def env_world() -> Tuple[int, int, int]:
    """(world_size, rank, local_rank) from the torchrun environment; (1, 0, 0) without it."""
    return (int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")),
            int(os.environ.get("LOCAL_RANK", "0")))


def rank_seed(base_seed: int, rank: int) -> int:
    """Per-rank data seed (1234 + rank)."""
    return int(base_seed) + int(rank)


def _active(group=None) -> bool:
    return dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1


def max_over_ranks(value: float, device: Optional[torch.device] = None, group=None) -> float:
    """Slowest rank's time: what a multi-GPU number is quoted on."""
    if not _active(group):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def bind_to_gpu_numa_node(device_index: int) -> Optional[str]:
    """Pin this process to the CPUs of the NUMA node the GPU hangs off (sysfs `local_cpulist` of its PCI
    function) BEFORE it allocates pinned host buffers: first-touch then places them next to the GPU's root
    port, which is what the host-buffer (e2e) path needs when eight ranks stream over PCIe at once.
    Returns the cpulist used, or None when the topology is not visible (containers may hide sysfs)."""
    try:
        p = torch.cuda.get_device_properties(device_index)
        addr = f"{getattr(p, 'pci_domain_id', 0):04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{addr}/local_cpulist") as fh:
            cpulist = fh.read().strip()
        cpus = set()
        for part in cpulist.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if not cpus or cpus == allowed:
            return None
        os.sched_setaffinity(0, cpus)
        return cpulist
    except Exception:
        return None
