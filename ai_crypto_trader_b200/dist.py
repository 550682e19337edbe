"""Multi-GPU sharding of the population sweep: one process per GPU (torchrun),
individuals partitioned contiguously across ranks, market data replicated, ONE
all-gather of the fitness vector per generation.  NCCL on GPUs;
the same code runs over gloo on CPU tensors for the host-logic tests.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Sequence

import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(n: int, world: int, rank: int):
    """Contiguous shard [lo, hi) of n items for `rank`; every rank holds ceil(n/world) slots."""
    per = (n + world - 1) // world
    lo = min(rank * per, n)
    return lo, min(lo + per, n), per


class ShardedFitness:
    """Batched fitness callable: evaluates this rank's shard with `evaluate_local`
    (List[Dict] -> float64 array) and all-gathers the per-rank results.

    `fitness = ShardedFitness(sweep.evaluate)`; `GeneticAlgorithm(..., fitness_function=fitness)`
    then calls `fitness.batch(population)` once per generation on every rank (the GA
    operators run replicated from the same seed, so populations stay identical).
    """

    def __init__(self, evaluate_local: Callable[[List[Dict]], Sequence[float]], device=None, group=None):
        self.evaluate_local = evaluate_local
        self.group = group
        self.device = device
        self.batch = self.__call_batch

    def _world(self):
        if dist.is_available() and dist.is_initialized():
            return dist.get_world_size(self.group), dist.get_rank(self.group)
        return 1, 0

    def __call_batch(self, population: List[Dict]) -> List[float]:
        world, rank = self._world()
        n = len(population)
        lo, hi, per = shard_bounds(n, world, rank)
        local = np.asarray(self.evaluate_local(population[lo:hi]) if hi > lo else [], dtype=np.float64)
        if world == 1:
            return local.tolist()
        dev = self.device
        if dev is None:
            dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(self.group) == "nccl" else torch.device("cpu")
        send = torch.full((per,), float("nan"), dtype=torch.float64, device=dev)
        if hi > lo:
            send[:hi - lo] = torch.from_numpy(local).to(dev)
        recv = torch.empty(per * world, dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(recv, send, group=self.group)   # the one collective per generation
        full = recv.cpu().numpy()
        out = np.concatenate([full[r * per: r * per + (shard_bounds(n, world, r)[1] - shard_bounds(n, world, r)[0])]
                              for r in range(world)])
        return out.tolist()

    def __call__(self, individual: Dict) -> float:
        return float(self.evaluate_local([individual])[0])


def gather_paths(local_finals: torch.Tensor, local_maxdd: torch.Tensor, n_total: int, group=None):
    """All-gather the per-rank shards of a Monte-Carlo run (paths sharded contiguously by shard_bounds, the kernels
    keyed by the GLOBAL path index, so the gathered arrays equal a single-GPU run bit for bit).  One collective of
    8 bytes per path; every rank then holds all finals / drawdowns and computes the same order statistics."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local_finals, local_maxdd
    world = dist.get_world_size(group)
    per = -(-n_total // world)
    send = torch.full((2 * per,), float("nan"), dtype=torch.float32, device=local_finals.device)
    send[:local_finals.numel()] = local_finals
    send[per:per + local_maxdd.numel()] = local_maxdd
    recv = torch.empty((world * 2 * per,), dtype=torch.float32, device=local_finals.device)
    dist.all_gather_into_tensor(recv, send, group=group)
    recv = recv.view(world, 2, per)
    counts = [shard_bounds(n_total, world, r)[1] - shard_bounds(n_total, world, r)[0] for r in range(world)]
    finals = torch.cat([recv[r, 0, :counts[r]] for r in range(world)])
    maxdd = torch.cat([recv[r, 1, :counts[r]] for r in range(world)])
    return finals, maxdd
