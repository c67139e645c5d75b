"""ORACLE — test infrastructure, NOT product code.

CPU stand-in for the reference's pybind module `pointnet2_cuda` (src/pointnet2_api.cpp:10-24):
same four forward entry points, same argument order, caller-allocated outputs written in place,
returns 1.  Backed by oracle/pointnet2_cpu.c through ctypes.  Installed as
`sys.modules['pointnet2_cuda']` by oracle/ref_loader.py so that the UNMODIFIED reference Python
(`pointnet2_utils.py:8 import pointnet2_cuda as pointnet2`) runs on the CPU in this container.
"""
import ctypes
import os
import subprocess

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libpointnet2_cpu.so")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "pointnet2_cpu.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "_build/libpointnet2_cpu.so"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        fp, ip, i, f = ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_float
        _lib.oracle_furthest_point_sampling.argtypes = [i, i, i, fp, fp, ip]
        _lib.oracle_gather_points.argtypes = [i, i, i, i, fp, ip, fp]
        _lib.oracle_ball_query.argtypes = [i, i, i, f, i, fp, fp, ip]
        _lib.oracle_group_points.argtypes = [i, i, i, i, i, fp, ip, fp]
    return _lib


def _chk(t, dtype):
    assert t.device.type == "cpu" and t.dtype == dtype and t.is_contiguous(), (t.device, t.dtype, t.is_contiguous())
    return t.data_ptr()


_pool = None


def _over_batch(b, fn):
    """Split the batch over host threads (ctypes drops the GIL during the foreign call; gcc in this
    image has no usable OpenMP runtime spec, so the parallel-for lives here)."""
    global _pool
    nthr = min(b, os.cpu_count() or 1)
    if nthr <= 1:
        fn(0, b)
        return 1
    if _pool is None:
        from concurrent.futures import ThreadPoolExecutor
        _pool = ThreadPoolExecutor(max_workers=os.cpu_count() or 1)
    bounds = [(b * t) // nthr for t in range(nthr + 1)]
    list(_pool.map(lambda t: fn(bounds[t], bounds[t + 1]), range(nthr)))
    return 1


def furthest_point_sampling_wrapper(b, n, m, xyz, temp, idx):
    px, pt, pi = _chk(xyz, torch.float32), _chk(temp, torch.float32), _chk(idx, torch.int32)
    L = lib()
    return _over_batch(b, lambda lo, hi: L.oracle_furthest_point_sampling(
        hi - lo, n, m, px + lo * n * 12, pt + lo * n * 4, pi + lo * m * 4))


def gather_points_wrapper(b, c, n, npoints, points, idx, out):
    pp, pi, po = _chk(points, torch.float32), _chk(idx, torch.int32), _chk(out, torch.float32)
    L = lib()
    return _over_batch(b, lambda lo, hi: L.oracle_gather_points(
        hi - lo, c, n, npoints, pp + lo * c * n * 4, pi + lo * npoints * 4, po + lo * c * npoints * 4))


def ball_query_wrapper(b, n, m, radius, nsample, new_xyz, xyz, idx):
    pn, px, pi = _chk(new_xyz, torch.float32), _chk(xyz, torch.float32), _chk(idx, torch.int32)
    L = lib()
    return _over_batch(b, lambda lo, hi: L.oracle_ball_query(
        hi - lo, n, m, float(radius), nsample, pn + lo * m * 12, px + lo * n * 12, pi + lo * m * nsample * 4))


def group_points_wrapper(b, c, n, npoints, nsample, points, idx, out):
    pp, pi, po = _chk(points, torch.float32), _chk(idx, torch.int32), _chk(out, torch.float32)
    L = lib()
    return _over_batch(b, lambda lo, hi: L.oracle_group_points(
        hi - lo, c, n, npoints, nsample, pp + lo * c * n * 4, pi + lo * npoints * nsample * 4,
        po + lo * c * npoints * nsample * 4))


def _not_on_path(*_a, **_k):  # backward / FP-module ops: out of scope
    raise NotImplementedError("oracle: op is not on the forward hot path")


gather_points_grad_wrapper = _not_on_path
group_points_grad_wrapper = _not_on_path
three_nn_wrapper = _not_on_path
three_interpolate_wrapper = _not_on_path
three_interpolate_grad_wrapper = _not_on_path
