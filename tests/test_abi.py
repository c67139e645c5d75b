"""CPU: the C-ABI shared library builds for sm_100a, loads, and exports every symbol that
include/genpose_b200.h declares (no compute calls — there is no GPU here)."""
import ctypes
import os
import re

import pytest

from genpose_b200 import lib, weights

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    lib.build()
    return lib.load()


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "genpose_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gpb_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(built):
    names = _declared_symbols()
    assert len(names) >= 18
    raw = ctypes.CDLL(lib.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), f"{n} declared in include/genpose_b200.h but not exported"
    # and the Python binding table covers exactly the header
    assert sorted(lib.SIGNATURES) == names


def test_abi_version_and_sizes(built):
    assert built.gpb_abi_version() == 1
    assert built.gpb_encoder_weights_floats() == weights.encoder_floats()
    assert built.gpb_trunk_weights_floats() == weights.trunk_floats()
    assert built.gpb_encode_workspace_bytes(64) > 0 and built.gpb_sampler_workspace_bytes(3200, 500) > 0
    assert built.gpb_trunk_tc_stream_bytes() == 2 * 65 * 16384 and built.gpb_trunk_tc16_stream_bytes() == 2 * 33 * 16384
    assert built.gpb_launch_count() == 0


def test_argument_validation_without_gpu(built):
    # invalid arguments are rejected before any CUDA call, with a message
    assert built.gpb_furthest_point_sampling(1, 4, 8, None, None, None, None) == -1
    assert b"m<=n" in built.gpb_last_error_string()
    assert built.gpb_rank_pool(None, None, 1, 500, 1, None, None, None, None, None) == -1
    assert built.gpb_encode(None, 0, None, None, None, 0, None, None, None, None) == 0      # empty batch is a no-op
    # the samplers: empty batch is a no-op, a too-small K is refused by the tensor-core entries (three- and two-product alike)
    for fn in (built.gpb_sample_pc_tc, built.gpb_sample_pc_tc16):
        extra = (None,) if fn is built.gpb_sample_pc_tc16 else ()
        assert fn(None, 0, 50, 500, 0.16, None, None, None, None, None, 0, None, None, None, None, 0, *extra, None) == 0
    assert built.gpb_sample_ode_tc16(None, 0, 50, 0.55, 1e-5, 1e-5, 1000, None, None, None, None, None, None, None, 0, None, 0, None, 0,
                                     None) == 0
    assert built.gpb_set_tc_team(3) == -1 and built.gpb_set_tc_team(0) == 0


def test_sass_is_sm100a_only():
    import subprocess
    out = subprocess.run(["cuobjdump", "--list-elf", lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs
