"""GPU: gpb_prepare_clouds (csrc/cloud_prep.cu) through the host mirror genpose_b200/cloud_prep.py against the golden
vectors produced by the reference's own functions, against the oracle on fresh frames, and its edge cases.
Bit-exact: the kernel restates integer index arithmetic and a fixed sequence of IEEE float32 operations."""
import os

import numpy as np
import pytest
import torch

from genpose_b200 import synth
from oracle import cloud_prep_oracle as P
from oracle import make_golden_prep

pytestmark = pytest.mark.gpu
GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def cp():
    from genpose_b200 import cloud_prep
    return cloud_prep


@pytest.mark.parametrize("name", sorted(make_golden_prep.CASES))
def test_prepare_frame_matches_reference_goldens_bit_exact(cp, name):
    g = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    depth, masks, rois = synth.make_frame(int(g["case_seed"]), int(g["case_n_inst"]))
    pts, valid_inst, n_valid = cp.prepare_frame(depth, masks, rois, synth.REAL_INTRINSICS, subset_ids=g["ids"])
    assert valid_inst == [i for i, v in enumerate(g["valid"]) if v]
    nv = n_valid.cpu().numpy()
    for i in range(masks.shape[2]):
        if g["valid"][i]:
            assert nv[i] == g["n_valid"][i]
            assert np.array_equal(pts[i].cpu().numpy(), g["points"][i]), f"instance {i}"
        else:
            assert nv[i] <= 1 and float(pts[i].abs().max()) == 0.0
    # the matrices the host mirror builds are the ones the reference built
    for i, r in enumerate(rois):
        assert np.array_equal(cp.crop_transform(r, *depth.shape), g["trans"][i])


@pytest.mark.parametrize("seed", [5, 6])
def test_prepare_frame_matches_oracle_on_fresh_frames(cp, seed):
    depth, masks, rois = synth.make_frame(seed, 18)
    rs = np.random.RandomState(seed)
    trans = np.stack([cp.crop_transform(r, *depth.shape) for r in rois])
    ids = np.zeros((len(rois), 1024), np.int32)
    ref = []
    for i in range(len(rois)):
        r, n = P.prepare_instance(depth, masks[:, :, i], trans[i], synth.REAL_INTRINSICS, ids_fn=lambda n: rs.permutation(n)[:1024])
        ref.append((r, n))
    # second pass to record the ids the oracle drew (same generator sequence)
    rs = np.random.RandomState(seed)
    for i, (r, n) in enumerate(ref):
        if n > 1024:
            ids[i] = rs.permutation(n)[:1024]
    pts, valid_inst, n_valid = cp.prepare_frame(depth, masks, rois, synth.REAL_INTRINSICS, subset_ids=ids)
    for i, (r, n) in enumerate(ref):
        assert int(n_valid[i]) == n
        if r is None:
            assert i not in valid_inst
        else:
            assert np.array_equal(pts[i].cpu().numpy(), r), f"instance {i}"


def test_throughput_mode_subset_is_the_keyed_permutation(cp):
    depth, masks, rois = synth.make_frame(7, 6)
    seed = 0x1234ABCD5678
    pts, valid_inst, n_valid = cp.prepare_frame(depth, masks, rois, synth.REAL_INTRINSICS, seed=seed)
    pts2, _, _ = cp.prepare_frame(depth, masks, rois, synth.REAL_INTRINSICS, seed=seed)
    assert torch.equal(pts, pts2)
    k0, k1 = seed & 0xFFFFFFFF, seed >> 32
    for i in valid_inst:
        n = int(n_valid[i])
        if n <= 1024:
            continue
        ids = P.feistel_permutation_prefix(n, 1024, (k0 + 0x85EBCA6B * i) & 0xFFFFFFFF, k1)
        trans = cp.crop_transform(rois[i], *depth.shape)
        ref, _ = P.prepare_instance(depth, masks[:, :, i], trans, synth.REAL_INTRINSICS, ids=ids)
        assert np.array_equal(pts[i].cpu().numpy(), ref)
        assert len(np.unique(ids)) == 1024                      # a subset without replacement, like np.random.permutation


def test_edge_cases(cp):
    from genpose_b200 import lib
    depth, masks, rois = synth.make_frame(8, 6)
    # no instances: a no-op
    pts, valid_inst, n_valid = cp.prepare_frame(depth, masks[:, :, :0], rois[:0], synth.REAL_INTRINSICS)
    assert pts.shape == (0, 1024, 3) and valid_inst == []
    # all-zero depth: every instance is skipped
    pts, valid_inst, n_valid = cp.prepare_frame(np.zeros_like(depth), masks, rois, synth.REAL_INTRINSICS)
    assert valid_inst == [] and int(n_valid.abs().max()) == 0 and float(pts.abs().max()) == 0.0
    # CPU tensors are refused (no fallback)
    with pytest.raises(lib.GenPoseB200Error):
        cp.prepare_clouds(torch.zeros(480, 640, dtype=torch.int16), torch.zeros(480, 640, 1, dtype=torch.bool),
                          torch.zeros(1, 6, dtype=torch.float64), synth.REAL_INTRINSICS)


def test_prepared_clouds_feed_the_pipeline(cp):
    """End to end: frame -> clouds -> PoseNet.pred_func, the hand-off of evaluation_single.py:394-403."""
    from genpose_b200.pipeline import PosePipeline
    depth, masks, rois = synth.make_frame(9, 6)
    pts, valid_inst, _ = cp.prepare_frame(depth, masks, rois, synth.REAL_INTRINSICS, seed=3)
    sd = synth.make_state_dict(0, kappa=-0.3)
    pipe = PosePipeline(sd, None, sampler="pc", sampling_steps=20)
    out = pipe.run(PosePipeline.make_batch(pts[valid_inst].contiguous()), repeat_num=50)
    assert out["pred_pose"].shape == (len(valid_inst), 50, 9) and torch.isfinite(out["pred_pose"]).all()
