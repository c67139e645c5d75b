"""`from configs.config import get_config` -> same flag surface."""
from genpose_b200.config import get_config  # noqa: F401
