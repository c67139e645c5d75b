"""One process per GPU; objects are sharded contiguously over ranks (every object's encoder pass, K
candidates, ranking and pooling are independent — SURVEY.md §8e), weights are replicated (8.8 MB), and the
ONLY exchange is one all-gather of the final poses (+ energies / pooled transforms) over NCCL (NVLink 5 /
NVSwitch).  ~115 KB per rank at 64x50x9 fp32 => latency-bound, so there is nothing to fuse.

Parity caveat: the PC sampler couples all candidates of a launch through the batch-mean
gradient norm (samplers.py:130) and the ODE sampler through one RK45 error norm (samplers.py:205).  A
sharded run therefore reproduces the reference executed on the same SHARD, not on the global batch."""
import os
from typing import Callable, Dict, List, Tuple

import torch
import torch.distributed as dist


def init_from_env(backend: str = None) -> Tuple[int, int, int]:
    """-> (rank, world_size, local_rank); initialises torch.distributed when launched under torchrun."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local_rank


def shard_bounds(n_objects: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous, balanced partition of object indices; the first (n % world) ranks get one more."""
    base, rem = divmod(n_objects, world)
    bounds, lo = [], 0
    for r in range(world):
        hi = lo + base + (1 if r < rem else 0)
        bounds.append((lo, hi))
        lo = hi
    return bounds


def all_gather_objects_dim0(local: torch.Tensor, n_objects: int, world: int, rank: int) -> torch.Tensor:
    """All-gather tensors whose dim 0 is this rank's object shard into the global [n_objects, ...] tensor
    (one collective; shards may differ by one object, so the exchange is padded to the largest shard)."""
    if world == 1:
        return local
    bounds = shard_bounds(n_objects, world)
    max_n = max(hi - lo for lo, hi in bounds)
    pad = torch.zeros((max_n,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    gathered = torch.empty((world * max_n,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(gathered, pad)
    if all(hi - lo == max_n for lo, hi in bounds):
        return gathered
    parts = [gathered[r * max_n: r * max_n + (hi - lo)] for r, (lo, hi) in enumerate(bounds)]
    return torch.cat(parts, dim=0)


def run_sharded(local_fn: Callable[[int, int], Dict[str, torch.Tensor]], n_objects: int, keys=("pred_pose",)) -> Dict[str, torch.Tensor]:
    """local_fn(lo, hi) computes this rank's objects [lo, hi) and returns tensors with dim 0 = hi - lo.
    The selected keys are packed into ONE buffer and all-gathered once."""
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    lo, hi = shard_bounds(n_objects, world)[rank]
    local = local_fn(lo, hi)
    flat = [local[k].reshape(hi - lo, -1).float() for k in keys]
    widths = [f.shape[1] for f in flat]
    packed = torch.cat(flat, dim=1).contiguous()
    full = all_gather_objects_dim0(packed, n_objects, world, rank)
    out, col = {}, 0
    for k, w in zip(keys, widths):
        out[k] = full[:, col: col + w].reshape((n_objects,) + tuple(local[k].shape[1:]))
        col += w
    return out
