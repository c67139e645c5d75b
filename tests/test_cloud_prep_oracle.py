"""CPU: the point-cloud preparation oracle (oracle/cloud_prep_oracle.py) against (a) the golden vectors produced by the
reference's own functions (oracle/make_golden_prep.py), (b) cv2.warpAffine itself where cv2 is importable, and (c) the
live reference where /root/reference exists."""
import os

import numpy as np
import pytest

from genpose_b200 import synth
from oracle import cloud_prep_oracle as P
from oracle import make_golden_prep, ref_loader

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(name):
    g = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    depth, masks, rois = synth.make_frame(int(g["case_seed"]), int(g["case_n_inst"]))
    assert float(depth.astype(np.float64).sum()) == float(g["cs_depth"]) and float(masks.sum()) == float(g["cs_masks"]), "input drift"
    return g, depth, masks, rois


@pytest.mark.parametrize("name", sorted(make_golden_prep.CASES))
def test_oracle_matches_reference_goldens_bit_exact(name):
    g, depth, masks, rois = _load(name)
    kinds = set()
    for i in range(masks.shape[2]):
        center, scale = P.crop_params(rois[i], *depth.shape)
        src, dst = P.affine_points(center, scale)
        trans = P.solve_affine(src, dst)
        np.testing.assert_allclose(trans, g["trans"][i], rtol=0, atol=1e-9)             # cv2.getAffineTransform (LU rounding: ~1 ulp)
        pts, n_valid = P.prepare_instance(depth, masks[:, :, i], g["trans"][i], synth.REAL_INTRINSICS, ids=g["ids"][i])
        assert (pts is not None) == bool(g["valid"][i])
        if pts is None:
            kinds.add("skipped")
            continue
        assert n_valid == int(g["n_valid"][i])
        assert np.array_equal(pts, g["points"][i]), f"instance {i}: points differ"
        kinds.add("tiled" if n_valid < P.NUM_POINTS else "subset")
    assert kinds == {"skipped", "tiled", "subset"}                                       # every branch of sample_points is exercised


def test_warp_index_map_matches_cv2_on_random_crops():
    cv2 = pytest.importorskip("cv2")
    rs = np.random.RandomState(3)
    H, W = 480, 640
    img = rs.randint(1, 60000, (H, W)).astype(np.uint16)
    xs = np.tile(np.arange(W, dtype=np.float32), (H, 1))
    ys = np.tile(np.arange(H, dtype=np.float32)[:, None], (1, W))
    for it in range(60):
        y1, x1 = rs.randint(-20, H - 10), rs.randint(-20, W - 10)
        roi = (max(y1, 0), max(x1, 0), min(y1 + rs.randint(2, 300), H), min(x1 + rs.randint(2, 300), W))
        center, scale = P.crop_params(roi, H, W)
        if it % 3 == 0:                                   # also non-integer centres / scales, outside what get_bbox produces
            center = center + rs.uniform(-3, 3, 2)
            scale = scale * rs.uniform(0.7, 1.3)
        src, dst = P.affine_points(center, scale)
        trans = cv2.getAffineTransform(np.float32(src), np.float32(dst))
        np.testing.assert_allclose(P.solve_affine(src, dst), trans, rtol=0, atol=1e-9)
        ref = cv2.warpAffine(img, trans, (256, 256), flags=cv2.INTER_NEAREST)
        assert np.array_equal(P.warp_nearest(img, trans), ref)
        X, Y = P.warp_source_index(trans)
        inb = (X >= 0) & (X < W) & (Y >= 0) & (Y < H)
        rx = cv2.warpAffine(xs, trans, (256, 256), flags=cv2.INTER_NEAREST)
        ry = cv2.warpAffine(ys, trans, (256, 256), flags=cv2.INTER_NEAREST)
        assert np.array_equal(np.where(inb, X, 0).astype(np.float32), rx) and np.array_equal(np.where(inb, Y, 0).astype(np.float32), ry)


def test_feistel_subset_is_a_permutation_prefix():
    for n in (1025, 1500, 4097, 50000):
        ids = P.feistel_permutation_prefix(n, 1024, 0x1234, 0xABCD)
        assert ids.min() >= 0 and ids.max() < n and len(np.unique(ids)) == 1024
    full = P.feistel_permutation_prefix(300, 300, 7, 9)
    assert np.array_equal(np.sort(full), np.arange(300))


@pytest.mark.skipif(not ref_loader.available(), reason="needs the reference checkout (build container only)")
def test_goldens_are_what_the_reference_produces_now():
    for name in make_golden_prep.CASES:
        rec = make_golden_prep.generate(name)
        old = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        for k in rec:
            assert np.array_equal(np.asarray(rec[k]), old[k]), (name, k)
