"""ORACLE — test infrastructure, NOT product code.

Builds the REFERENCE'S OWN CUDA extension (`pointnet2_cuda`, pointnet2/setup.py:4-23) for sm_100a from
its sources where they lie under /root/reference — nothing is copied into this repo — with the output
only in oracle/_ref/ (git-ignored, but it travels to the GPU box with the snapshot).  On the B200 it is
the GPU-side oracle for bit-exact FPS / ball-query / gather / group parity: the actual reference
kernels, same flags (nvcc -O2, default --fmad=true), run on the same inputs as ours
(tests/test_gpu_reference_ext.py).  Recipe = the reference's own source list and flags, driven through
torch.utils.cpp_extension.load instead of its setup.py (which would write into the read-only tree).
"""
import importlib.util
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(_HERE, "_ref")
NAME = "pointnet2_cuda_ref"
SRC_ROOT = os.path.join(os.environ.get("GENPOSE_REFERENCE_ROOT", "/root/reference"),
                        "networks", "pts_encoder", "pointnet2_utils", "pointnet2", "src")
SOURCES = ["pointnet2_api.cpp", "ball_query.cpp", "ball_query_gpu.cu", "group_points.cpp", "group_points_gpu.cu",
           "interpolate.cpp", "interpolate_gpu.cu", "sampling.cpp", "sampling_gpu.cu"]   # setup.py:8-18


def so_path():
    return os.path.join(REF_DIR, NAME + ".so")


def build_if_possible(verbose: bool = False):
    """Compile when the reference tree is present and the .so is missing; no-op otherwise."""
    if os.path.exists(so_path()):
        return so_path()
    if not os.path.isdir(SRC_ROOT):
        return None
    from torch.utils.cpp_extension import load
    os.makedirs(REF_DIR, exist_ok=True)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    load(name=NAME, sources=[os.path.join(SRC_ROOT, s) for s in SOURCES], build_directory=REF_DIR,
         extra_cflags=["-g"], extra_cuda_cflags=["-O2", "-gencode", "arch=compute_100a,code=sm_100a"],
         verbose=verbose, is_python_module=False)
    return so_path() if os.path.exists(so_path()) else None


def load_ref():
    """Import the prebuilt extension (GPU box: no compiler step, no reference tree needed)."""
    path = so_path()
    if not os.path.exists(path):
        return None
    import torch  # noqa: F401  (libtorch must be loaded before the extension)
    spec = importlib.util.spec_from_file_location(NAME, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build_if_possible(verbose=True))
