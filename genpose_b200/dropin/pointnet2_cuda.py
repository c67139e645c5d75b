"""`import pointnet2_cuda` -> genpose_b200's compat module."""
from genpose_b200.pointnet2_cuda import *  # noqa: F401,F403
from genpose_b200.pointnet2_cuda import (ball_query_wrapper, furthest_point_sampling_wrapper,  # noqa: F401
                                         gather_points_wrapper, group_points_wrapper)
