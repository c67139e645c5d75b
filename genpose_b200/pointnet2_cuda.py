"""Drop-in for the reference's pybind module `pointnet2_cuda` (pointnet2/src/pointnet2_api.cpp:10-24):
same function names, argument order and conventions (caller-allocated CUDA tensors written in place,
returns 1, runs on the current stream), backed by the C ABI's compat layer.  With this module on the path
the reference's own pointnet2_utils.py (`import pointnet2_cuda as pointnet2`, :8) runs unchanged on the
sm_100a kernels — see INTEGRATION.md §1.  Backward / interpolation ops are not on the forward hot path."""
import torch

from . import lib


def _p(t, dtype, name):
    if not t.is_cuda or t.dtype != dtype or not t.is_contiguous():
        raise lib.GenPoseB200Error(f"pointnet2_cuda.{name}: expected a contiguous CUDA {dtype} tensor")   # CHECK_INPUT, ball_query.cpp:12-19
    return t.data_ptr()


def _s():
    return torch.cuda.current_stream().cuda_stream


def furthest_point_sampling_wrapper(b, n, m, points_tensor, temp_tensor, idx_tensor):
    lib.check(lib.load().gpb_furthest_point_sampling(b, n, m, _p(points_tensor, torch.float32, "fps"),
                                                     _p(temp_tensor, torch.float32, "fps"), _p(idx_tensor, torch.int32, "fps"), _s()),
              "furthest_point_sampling_wrapper")
    return 1


def gather_points_wrapper(b, c, n, npoints, points_tensor, idx_tensor, out_tensor):
    lib.check(lib.load().gpb_gather_points(b, c, n, npoints, _p(points_tensor, torch.float32, "gather"),
                                           _p(idx_tensor, torch.int32, "gather"), _p(out_tensor, torch.float32, "gather"), _s()),
              "gather_points_wrapper")
    return 1


def ball_query_wrapper(b, n, m, radius, nsample, new_xyz_tensor, xyz_tensor, idx_tensor):
    lib.check(lib.load().gpb_ball_query(b, n, m, float(radius), nsample, _p(new_xyz_tensor, torch.float32, "ball_query"),
                                        _p(xyz_tensor, torch.float32, "ball_query"), _p(idx_tensor, torch.int32, "ball_query"), _s()),
              "ball_query_wrapper")
    return 1


def group_points_wrapper(b, c, n, npoints, nsample, points_tensor, idx_tensor, out_tensor):
    lib.check(lib.load().gpb_group_points(b, c, n, npoints, nsample, _p(points_tensor, torch.float32, "group"),
                                          _p(idx_tensor, torch.int32, "group"), _p(out_tensor, torch.float32, "group"), _s()),
              "group_points_wrapper")
    return 1


def _backward_only(*_a, **_k):
    raise NotImplementedError("genpose_b200 implements the forward inference ops only (SURVEY.md §2.2)")


gather_points_grad_wrapper = _backward_only
group_points_grad_wrapper = _backward_only
three_nn_wrapper = _backward_only
three_interpolate_wrapper = _backward_only
three_interpolate_grad_wrapper = _backward_only
