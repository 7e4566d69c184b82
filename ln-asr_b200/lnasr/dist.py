"""Multi-GPU plumbing: one process per GPU (torch.distributed, NCCL over NVLink / NVSwitch).

The acoustic path shards by UTTERANCE: decoding needs no data-path collective; Baum-Welch
needs one all-reduce (SUM, float64) of the packed accumulator buffer per EM iteration, after which every rank
runs the same M-step kernel on identical numbers, so replicas stay bit-identical.
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np

__all__ = ["shard_utterances", "make_reduce_fn", "world", "bind_to_gpu_numa_node"]


def world():
    """(rank, world_size) of the default process group, (0, 1) when torch.distributed is not initialised."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_utterances(lengths: Sequence[int], world_size: int, rank: int) -> List[int]:
    """Indices of the utterances rank `rank` owns: longest-first greedy assignment to the least loaded rank, so
    frames per rank differ by at most one utterance.  Deterministic; every rank computes the same partition."""
    lengths = np.asarray(lengths, dtype=np.int64)
    order = np.argsort(-lengths, kind="stable")
    load = np.zeros(world_size, dtype=np.int64)
    owner = np.empty(len(lengths), dtype=np.int64)
    for i in order:
        r = int(np.argmin(load))
        owner[i] = r
        load[r] += lengths[i]
    return [int(i) for i in np.nonzero(owner == rank)[0]]


def make_reduce_fn(group=None):
    """In-place SUM all-reduce of a tensor over `group`; identity when there is a single rank."""
    import torch.distributed as dist

    def reduce_fn(t):
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        return t
    return reduce_fn


def bind_to_gpu_numa_node(device_index: int) -> dict:
    """Pin this process (and the pinned host buffers it allocates afterwards) to the CPUs of the NUMA node the GPU hangs off.

    With one process per GPU on a two-socket host, host->device copies from memory on the OTHER socket cross the
    inter-socket link and share it with the other ranks; the e2e path is PCIe-bound, so placement matters.  Best effort:
    reads /sys/bus/pci/devices/<bus id>/numa_node and /sys/devices/system/node/node<k>/cpulist; returns what it did."""
    import os
    info = {"device": device_index, "numa_node": None, "cpus": None, "bound": False}
    try:
        import torch
        p = torch.cuda.get_device_properties(device_index)
        bus = "%04x:%02x:%02x.0" % (getattr(p, "pci_domain_id", 0), p.pci_bus_id, p.pci_device_id)
        info["pci"] = bus
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            node = int(f.read().strip())
        info["numa_node"] = node
        if node < 0:
            return info
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        info["cpus"] = len(cpus)
        if cpus:
            os.sched_setaffinity(0, cpus)
            info["bound"] = True
    except Exception as e:                      # sysfs not visible in the container, no such attribute, ...
        info["error"] = repr(e)
    return info
