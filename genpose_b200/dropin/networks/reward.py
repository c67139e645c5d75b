"""`from networks.reward import sort_poses_by_energy, ranking_loss` (runners/evaluation_single.py:23)."""
from genpose_b200.reward import ranking_loss, sort_poses_by_energy  # noqa: F401
