"""Energy ranking with the reference's signature (networks/reward.py:131-155), on the rank/pool kernel."""
import torch

from . import ops


def sort_poses_by_energy(poses: torch.Tensor, energy: torch.Tensor):
    """poses [bs, K, 9], energy [bs, K, 2] -> (sorted_poses, sorted_energy): descending per object; the
    rotation columns follow the rot-energy order and the translation columns the trans-energy order."""
    sp, se, _ = ops.rank_pool(poses.float().contiguous(), energy.float().contiguous(), want_pooled=False)
    return sp, se


def ranking_loss(energy):
    raise NotImplementedError("ranking_loss is a training loss (networks/reward.py:109-128); out of scope")
