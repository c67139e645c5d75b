"""`from networks.posenet_agent import PoseNet` -> the B200-native agent."""
from genpose_b200.posenet_agent import PoseNet  # noqa: F401
