"""`from networks.posenet import GFObjectPose` -> the B200-native model facade."""
from genpose_b200.posenet import GFObjectPose  # noqa: F401
from genpose_b200.sde import init_sde  # noqa: F401
