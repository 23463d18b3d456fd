"""Multi-GPU harness: lane shards, one reduction per pass.

Lanes of a batch, scenario replicas and optimisation candidates never exchange state during
a rollout (the reference runs them sequentially: example/inverse/_inverse.py:113-170 trials,
:280-292 CMA population; example/control/trainer.py:148-152 episodes), so the path shards with
NO data-path collective: every rank owns a contiguous block of macro lanes and of micro lanes
(with their CSR vehicle ranges) and runs the fused rollouts on it.  The only exchange is one
``all_reduce(sum)`` per forward+backward pass over the scalar losses and over gradients of
parameters shared by all shards (none in the inverse problems; the signal controller in ITSCP).
One process per GPU (torchrun), NCCL over NVLink on GPUs; the same code runs on gloo for the
CPU tests.  Per-lane state gradients are never reduced: each lane belongs to one rank.
"""
from __future__ import annotations

from typing import Iterable, Sequence, Tuple

import torch
import torch.distributed as td


def world() -> Tuple[int, int]:
    """(rank, world_size); (0, 1) outside a process group."""
    if td.is_available() and td.is_initialized():
        return td.get_rank(), td.get_world_size()
    return 0, 1


def shard_range(n: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of n items for `rank`; the first n % world_size ranks get one extra."""
    base, rem = divmod(n, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_lanes(tensors: Sequence[torch.Tensor], rank: int, world_size: int):
    """Slice every [B, ...] tensor of a macro-lane batch to this rank's lanes."""
    lo, hi = shard_range(tensors[0].shape[0], rank, world_size)
    return [t[lo:hi] for t in tensors]


def shard_csr(lane_off: torch.Tensor, per_vehicle: Sequence[torch.Tensor], per_lane: Sequence[torch.Tensor], rank: int,
              world_size: int):
    """Micro lanes: this rank's lanes with re-based CSR offsets, the matching vehicle slices of the
    per-vehicle tensors (last axis = vehicles, e.g. p[V], params[6, V]) and lane slices of per-lane tensors."""
    L = lane_off.numel() - 1
    lo, hi = shard_range(L, rank, world_size)
    v0, v1 = int(lane_off[lo]), int(lane_off[hi])
    off = (lane_off[lo:hi + 1] - lane_off[lo]).to(lane_off.dtype)
    return off, [t[..., v0:v1] for t in per_vehicle], [t[lo:hi] for t in per_lane]


def reduce_losses(*losses: torch.Tensor, group=None) -> torch.Tensor:
    """Sum of each scalar loss over all shards: THE collective of a pass (one small all_reduce)."""
    t = torch.stack([l.detach().reshape(()) for l in losses])
    if world()[1] > 1:
        td.all_reduce(t, op=td.ReduceOp.SUM, group=group)
    return t


def reduce_shared_grads(params: Iterable[torch.Tensor], group=None) -> None:
    """Sum the gradients of parameters every shard shares (flattened into one bucket, one all_reduce)."""
    grads = [p.grad for p in params if p.grad is not None]
    if not grads or world()[1] == 1:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    td.all_reduce(flat, op=td.ReduceOp.SUM, group=group)
    o = 0
    for g in grads:
        g.copy_(flat[o:o + g.numel()].view_as(g))
        o += g.numel()


def gather_lanes(local: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """All ranks' per-lane rows in lane order (testing / result collection; not on the timed path)."""
    rank, ws = world()
    if ws == 1:
        return local
    sizes = [shard_range(n_total, r, ws) for r in range(ws)]
    width = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in range(ws)]
    td.all_gather(out, pad, group=group)
    return torch.cat([o[:hi - lo] for o, (lo, hi) in zip(out, sizes)])
