"""Energy ranking with the reference's signature (networks/reward.py:131-155), on the rank/pool kernel."""
import torch

from . import ops


def sort_poses_by_energy(poses: torch.Tensor, energy: torch.Tensor):
    """poses [bs, K, 9], energy [bs, K, 2] -> (sorted_poses, sorted_energy): descending per object; the
    rotation columns follow the rot-energy order and the translation columns the trans-energy order."""
    if poses.dtype == torch.float32:
        sp, se, _ = ops.rank_pool(poses.contiguous(), energy.float().contiguous(), want_pooled=False)
        return sp, se
    # float64 poses (the ODE sampler's, samplers.py:206-207) keep their dtype like the reference's torch.gather (reward.py:145-152):
    # the kernel ranks, the poses are reordered by its indices
    _, se, _, order = ops.rank_pool(poses.float().contiguous(), energy.float().contiguous(), want_pooled=False, want_order=True)
    idx = order.long()
    rot = torch.gather(poses[:, :, :-3], 1, idx[:, :, 0:1].expand(-1, -1, poses.shape[-1] - 3))
    trans = torch.gather(poses[:, :, -3:], 1, idx[:, :, 1:2].expand(-1, -1, 3))
    return torch.cat([rot, trans], dim=-1), se


def ranking_loss(energy):
    raise NotImplementedError("ranking_loss is a training loss (networks/reward.py:109-128); out of scope")
