"""One process per GPU; objects are sharded contiguously over ranks (every object's encoder pass, K
candidates, ranking and pooling are independent — SURVEY.md §8e), weights are replicated (8.8 MB), and the
ONLY exchange is one all-gather of the final poses (+ energies / pooled transforms) over NCCL (NVLink 5 /
NVSwitch).  ~115 KB per rank at 64x50x9 fp32 => latency-bound, so there is nothing to fuse.

Parity caveat: the PC sampler couples all candidates of a launch through the batch-mean
gradient norm (samplers.py:130) and the ODE sampler through one RK45 error norm (samplers.py:205).  A
sharded run therefore reproduces the reference executed on the same SHARD, not on the global batch."""
import math
import os
from typing import Callable, Dict, List, Tuple

import torch
import torch.distributed as dist


def init_from_env(backend: str = None) -> Tuple[int, int, int]:
    """-> (rank, world_size, local_rank); initialises torch.distributed when launched under torchrun."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if torch.cuda.is_available() and (backend or "nccl") == "nccl":
        # One process per GPU: the C ABI launches on the CUDA *current* device and PoseNet's cfg.device is plain 'cuda', so the
        # rank's device must be current before anything is allocated (and before NCCL picks its device).
        torch.cuda.set_device(local_rank % torch.cuda.device_count())
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
    # gloo has no device collectives: CUDA shards are staged through the host there (single-GPU test boxes, CPU-only ranks);
    # under NCCL the exchange stays on the devices (NVLink / NVSwitch)
    xdev = torch.device("cpu") if (dist.get_backend() == "gloo" and local.is_cuda) else local.device
    pad = torch.zeros((max_n,) + tuple(local.shape[1:]), dtype=local.dtype, device=xdev)
    pad[: local.shape[0]] = local
    gathered = torch.empty((world * max_n,) + tuple(local.shape[1:]), dtype=local.dtype, device=xdev)
    dist.all_gather_into_tensor(gathered, pad)
    gathered = gathered.to(local.device)
    if all(hi - lo == max_n for lo, hi in bounds):
        return gathered
    parts = [gathered[r * max_n: r * max_n + (hi - lo)] for r, (lo, hi) in enumerate(bounds)]
    return torch.cat(parts, dim=0)


def run_sharded(local_fn: Callable[[int, int], Dict[str, torch.Tensor]], n_objects: int, keys=("pred_pose",)) -> Dict[str, torch.Tensor]:
    """local_fn(lo, hi) computes this rank's objects [lo, hi) and returns tensors with dim 0 = hi - lo.
    The selected keys are packed into ONE buffer (in the widest dtype among them, so the ODE sampler's float64 poses
    survive) and all-gathered once.  A rank whose shard is empty (fewer objects than ranks: common in per-frame
    evaluation) does not call local_fn; it still joins the collective, with the shapes and dtypes the other ranks' results
    have, learnt from rank 0 through one small metadata broadcast (only taken when n_objects < world_size)."""
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    lo, hi = shard_bounds(n_objects, world)[rank]
    local = local_fn(lo, hi) if hi > lo else None
    meta = None if local is None else {k: (tuple(local[k].shape[1:]), local[k].dtype, local[k].device) for k in keys}
    if world > 1 and n_objects < world:
        # some shards are empty: their ranks learn shapes / dtypes from rank 0 (the first rank always owns an object)
        box = [None if meta is None else {k: (m[0], m[1]) for k, m in meta.items()}]
        dist.broadcast_object_list(box, src=0)
        if meta is None:
            on_gpu = dist.get_backend() == "nccl" or (torch.cuda.is_available() and torch.cuda.is_initialized())
            dev = torch.device("cuda", torch.cuda.current_device()) if on_gpu else torch.device("cpu")
            meta = {k: (tuple(box[0][k][0]), box[0][k][1], dev) for k in keys}
    if meta is None:
        raise ValueError("run_sharded: no objects at all")
    wide = torch.float64 if any(m[1] == torch.float64 for m in meta.values()) else torch.float32
    widths = [int(math.prod(meta[k][0])) for k in keys]
    device = next(iter(meta.values()))[2]
    if local is None:
        packed = torch.zeros(0, sum(widths), dtype=wide, device=device)
    else:
        packed = torch.cat([local[k].reshape(hi - lo, w).to(wide) for k, w in zip(keys, widths)], dim=1).contiguous()
    full = all_gather_objects_dim0(packed, n_objects, world, rank)
    out, col = {}, 0
    for k, w in zip(keys, widths):
        out[k] = full[:, col: col + w].reshape((n_objects,) + meta[k][0]).to(meta[k][1])
        col += w
    return out
