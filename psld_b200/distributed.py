"""Batch-sharded multi-GPU sampling: one process per GPU, no per-step communication.

The reference shards sampling across GPUs by running N independent Lightning processes, each
on its own slice of the latent dataset with seed ``seed + global_rank`` and no collective at
all (reference ``main/eval/sample.py:74-77,108-109``, ``main/models/wrapper.py:93-99``,
``main/callbacks.py:98``).  Here the same partition is kept (contiguous batch split, per-rank
Philox seed offset) and ONE all-gather of the final position half collects the samples
(BASELINE.json north_star); ``torch.distributed`` (NCCL over NVLink on GPUs, gloo in the CPU
tests) is plumbing only.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def env_rank():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), \
        int(os.environ.get("LOCAL_RANK", "0"))


def shard_bounds(n_total: int, rank: int, world: int):
    """Contiguous split; the first ``n_total % world`` ranks take one extra sample."""
    base, rem = divmod(int(n_total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard(t: torch.Tensor, rank: int, world: int):
    lo, hi = shard_bounds(t.shape[0], rank, world)
    return t[lo:hi]


def rank_seed(seed: int, rank: int) -> int:
    """Per-rank noise stream (reference: ``seed + global_rank``, wrapper.py:99)."""
    return int(seed) + int(rank)


def current_rank() -> int:
    """Global rank of this process AT CALL TIME: ``torch.distributed`` when initialised (Lightning's
    DDP launcher initialises it but exports only LOCAL_RANK / NODE_RANK / WORLD_SIZE to the children,
    not RANK), else RANK, else NODE_RANK x local world + LOCAL_RANK, else 0."""
    if dist.is_available() and dist.is_initialized():
        return int(dist.get_rank())
    if "RANK" in os.environ:
        return int(os.environ["RANK"])
    local = int(os.environ.get("LOCAL_RANK", "0"))
    node = int(os.environ.get("NODE_RANK", os.environ.get("GROUP_RANK", "0")))
    per_node = int(os.environ.get("LOCAL_WORLD_SIZE", "0")) or 1
    return node * per_node + local


_M64 = (1 << 64) - 1


def _splitmix64(x: int) -> int:
    x = (x + 0x9E3779B97F4A7C15) & _M64
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & _M64
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & _M64
    return x ^ (x >> 31)


def call_seed(seed: int, rank: int, call: int) -> int:
    """64-bit Philox key of the ``call``-th ``sample()`` of a sampler on ``rank``.

    The reference draws its noise from the global torch generator, seeded once per process with
    ``seed + global_rank`` (wrapper.py:93-99) and ADVANCING from batch to batch, so every
    ``predict_step`` sees fresh noise.  A counter-based generator has no state to advance: the key
    must change per call or every batch would reuse the same injected noise.  Deterministic in
    (seed, rank, call); call 0 of rank r keeps the reference's ``seed + r`` key."""
    base = rank_seed(seed, rank) & _M64
    if call == 0:
        return base
    return _splitmix64(base ^ _splitmix64(int(call) & _M64))


def gather_samples(u_local: torch.Tensor, n_total: int | None = None, group=None):
    """All-gathers the position half x of the local states ``[b, 2C, H, W]`` -> ``[B, C, H, W]``
    (the momentum half is dropped exactly like the reference's image writer does,
    callbacks.py:103-107).  Handles uneven shards by padding to the largest shard."""
    x = torch.chunk(u_local, 2, dim=1)[0].contiguous()
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return x
    world = dist.get_world_size(group)
    n_total = n_total if n_total is not None else None
    sizes = torch.tensor([x.shape[0]], device=x.device, dtype=torch.int64)
    all_sizes = [torch.zeros_like(sizes) for _ in range(world)]
    dist.all_gather(all_sizes, sizes, group=group)
    all_sizes = [int(s.item()) for s in all_sizes]
    mx = max(all_sizes)
    if x.shape[0] < mx:
        x = torch.cat([x, x.new_zeros(mx - x.shape[0], *x.shape[1:])], 0)
    out = torch.empty(world * mx, *x.shape[1:], dtype=x.dtype, device=x.device)
    dist.all_gather_into_tensor(out, x, group=group)
    parts = [out[r * mx: r * mx + all_sizes[r]] for r in range(world)]
    return torch.cat(parts, 0)


def max_over_ranks(value: float, device) -> float:
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
