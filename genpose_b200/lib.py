"""ctypes binding of the C ABI (include/genpose_b200.h).  There is NO fallback: if the shared library is
missing or a call fails, an exception is raised — the product never routes through the oracle or torch ops."""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
# GPB_LIB: an alternative build of the same library (experiment variants from `make -C csrc variant NAME=.. EXTRA=..`), in-tree
LIB_PATH = os.environ.get("GPB_LIB") or os.path.join(_HERE, "libgenpose_b200.so")
CSRC = os.path.join(_HERE, "csrc")

_vp, _i, _f, _d, _sz, _u64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_double, ctypes.c_size_t, ctypes.c_uint64

# name -> (restype, argtypes); mirrors include/genpose_b200.h one to one
SIGNATURES = {
    "gpb_abi_version": (_i, []),
    "gpb_last_error_string": (ctypes.c_char_p, []),
    "gpb_launch_count": (_u64, []),
    "gpb_furthest_point_sampling": (_i, [_i, _i, _i, _vp, _vp, _vp, _vp]),
    "gpb_gather_points": (_i, [_i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "gpb_ball_query": (_i, [_i, _i, _i, _f, _i, _vp, _vp, _vp, _vp]),
    "gpb_group_points": (_i, [_i, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "gpb_encoder_weights_floats": (_sz, []),
    "gpb_trunk_weights_floats": (_sz, []),
    "gpb_encode_workspace_bytes": (_sz, [_i]),
    "gpb_encode": (_i, [_vp, _i, _vp, _vp, _vp, _sz, _vp, _vp, _vp, _vp]),
    "gpb_object_bias": (_i, [_vp, _i, _vp, _vp, _vp]),
    "gpb_trunk_eval": (_i, [_vp, _i, _i, _f, _vp, _vp, _i, _vp, _vp]),
    "gpb_sampler_workspace_bytes": (_sz, [_i, _i]),
    "gpb_sample_pc": (_i, [_vp, _i, _i, _i, _f, _vp, _vp, _vp, _vp, _u64, _vp, _vp, _vp, _vp, _sz, _vp]),
    "gpb_trunk_tc_stream_bytes": (_sz, []),
    "gpb_encoder_tc_bytes": (_sz, []),
    "gpb_encode_tc": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _sz, _vp, _vp, _vp, _vp]),
    "gpb_sample_pc_tc": (_i, [_vp, _i, _i, _i, _f, _vp, _vp, _vp, _vp, _vp, _u64, _vp, _vp, _vp, _vp, _sz, _vp]),
    "gpb_sample_pc_tc_dbg": (_i, [_vp, _i, _i, _i, _f, _vp, _vp, _vp, _vp, _vp, _u64, _vp, _vp, _vp, _vp, _sz, _vp, _vp]),
    "gpb_sample_ode": (_i, [_vp, _i, _i, _d, _d, _d, _i, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _i, _vp, _sz, _vp]),
    "gpb_sampler_tc_max_rows": (_i, [_i]),
    "gpb_set_tc_team": (_i, [_i]),
    "gpb_sample_ode_tc": (_i, [_vp, _i, _i, _d, _d, _d, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _i, _vp, _sz, _vp]),
    "gpb_sample_ode_tc_dbg": (_i, [_vp, _i, _i, _d, _d, _d, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp, _i, _vp]),
    "gpb_sample_ode_tc16_dbg": (_i, [_vp, _i, _i, _d, _d, _d, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp, _i, _vp]),
    "gpb_trunk_tc16_stream_bytes": (_sz, []),
    "gpb_sample_pc_tc16": (_i, [_vp, _i, _i, _i, _f, _vp, _vp, _vp, _vp, _vp, _u64, _vp, _vp, _vp, _vp, _sz, _vp, _vp]),
    "gpb_sample_ode_tc16": (_i, [_vp, _i, _i, _d, _d, _d, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _i, _vp, _sz, _vp]),
    "gpb_energy": (_i, [_vp, _i, _i, _f, _vp, _vp, _vp, _vp, _vp]),
    "gpb_rank_pool": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "gpb_prepare_clouds": (_i, [_vp, _vp, ctypes.c_longlong, ctypes.c_longlong, _i, _i, _i, _vp, ctypes.POINTER(ctypes.c_float), _vp, _u64,
                                _vp, _vp, _vp]),
    "gpb_selftest_umma": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp, _vp]),
}


class GenPoseB200Error(RuntimeError):
    pass


def build(verbose: bool = False) -> str:
    """Compile the CUDA sources for sm_100a into genpose_b200/libgenpose_b200.so (nvcc cross-compiles
    without a GPU)."""
    cmd = ["make", "-C", CSRC, "-j", str(os.cpu_count() or 4)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise GenPoseB200Error("building libgenpose_b200.so failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stdout)
    return LIB_PATH


_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GenPoseB200Error(
            f"{LIB_PATH} not found: run `python -c 'import __graft_entry__ as g; g.build()'` (or "
            f"`make -C genpose_b200/csrc`). There is no CPU/PyTorch fallback for the hot path.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here == the .so does not export the declared symbol
        fn.restype = res
        fn.argtypes = args
    if lib.gpb_abi_version() != 1:
        raise GenPoseB200Error("libgenpose_b200.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc: int, what: str):
    if rc != 0:
        msg = load().gpb_last_error_string().decode("utf-8", "replace")
        raise GenPoseB200Error(f"{what} failed (code {rc}): {msg}")


def launch_count() -> int:
    return int(load().gpb_launch_count())
