"""ORACLE — test infrastructure, NOT product code.

Golden vectors for the point-cloud preparation path, produced by EXECUTING THE REFERENCE'S OWN CODE on synthetic
frames (genpose_b200/synth.py::make_frame):
  * `crop_resize_by_warp_affine`, `get_2d_coord_np` imported from /root/reference/utils/datasets_utils.py (they call
    cv2.getAffineTransform / cv2.warpAffine of this container's OpenCV),
  * `get_bbox` and the nested functions `depth_to_pcl`, `sample_points` of `detect_mrcnn_genpose`
    (runners/evaluation_single.py:105-133), which cannot be imported (they are local to a function whose module needs
    CUDA at import): their source text is cut out of the unmodified file with `ast` and executed as is.
The loop glue below repeats evaluation_single.py:168-212 line by line around those calls.

    python -m oracle.make_golden_prep            # rewrites tests/golden/prep_*.npz
    python -m oracle.make_golden_prep --check
"""
import argparse
import ast
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from genpose_b200 import synth  # noqa: E402
from oracle import ref_loader  # noqa: E402

GOLDEN_DIR = os.path.join(_ROOT, "tests", "golden")
CASES = {"prep_frame0": dict(seed=0, n_inst=6), "prep_frame1": dict(seed=1, n_inst=12)}


def reference_functions():
    """-> dict with the reference's depth_to_pcl, sample_points, get_bbox, crop_resize_by_warp_affine, get_2d_coord_np."""
    import importlib.util
    ns = {"np": np}
    src = open(os.path.join(ref_loader.REFERENCE_ROOT, "runners", "evaluation_single.py")).read()
    tree = ast.parse(src)
    outer = next(n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef) and n.name == "detect_mrcnn_genpose")
    for n in outer.body:
        if isinstance(n, ast.FunctionDef) and n.name in ("depth_to_pcl", "sample_points"):
            exec(compile(ast.Module(body=[n], type_ignores=[]), "evaluation_single.py", "exec"), ns)
    sg = open(os.path.join(ref_loader.REFERENCE_ROOT, "utils", "sgpa_utils.py")).read()
    gb = next(n for n in ast.parse(sg).body if isinstance(n, ast.FunctionDef) and n.name == "get_bbox")
    exec(compile(ast.Module(body=[gb], type_ignores=[]), "sgpa_utils.py", "exec"), ns)
    spec = importlib.util.spec_from_file_location("_ref_datasets_utils", os.path.join(ref_loader.REFERENCE_ROOT, "utils", "datasets_utils.py"))
    du = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(du)
    ns["crop_resize_by_warp_affine"] = du.crop_resize_by_warp_affine
    ns["get_2d_coord_np"] = du.get_2d_coord_np
    ns["get_affine_transform"] = du.get_affine_transform
    return ns


def run_reference_frame(raw_depth, masks, rois, intrinsics, perm_seed):
    """evaluation_single.py:168-212 for every instance of one frame -> per-instance dict(valid, n_valid, points, ids, trans)."""
    import cv2
    R = reference_functions()
    im_H, im_W = raw_depth.shape
    img_size, num_points = 256, 1024                       # cfg.img_size, cfg.num_points (configs/config.py:78,24)
    out = []
    for i in range(masks.shape[2]):
        rmin, rmax, cmin, cmax = R["get_bbox"](rois[i])
        mask = np.logical_and(masks[:, :, i], raw_depth > 0)
        coord_2d = R["get_2d_coord_np"](im_W, im_H).transpose(1, 2, 0)
        bbox_xyxy = np.array([cmin, rmin, cmax, rmax])
        x1, y1, x2, y2 = bbox_xyxy
        cx = 0.5 * (x1 + x2)
        cy = 0.5 * (y1 + y2)
        bbox_center = np.array([cx, cy])
        scale = max(y2 - y1, x2 - x1)
        scale = min(scale, max(im_H, im_W)) * 1.0
        trans = R["get_affine_transform"](bbox_center, (scale, scale), 0, (img_size, img_size))
        roi_coord_2d = R["crop_resize_by_warp_affine"](coord_2d, bbox_center, scale, img_size, interpolation=cv2.INTER_NEAREST).transpose(2, 0, 1)
        mask_target = mask.copy().astype(np.float32)
        roi_mask = R["crop_resize_by_warp_affine"](mask_target, bbox_center, scale, img_size, interpolation=cv2.INTER_NEAREST)
        roi_mask = np.expand_dims(roi_mask, axis=0)
        roi_depth = R["crop_resize_by_warp_affine"](raw_depth, bbox_center, scale, img_size, interpolation=cv2.INTER_NEAREST)
        roi_depth = np.expand_dims(roi_depth, axis=0)
        depth_valid = roi_depth > 0
        rec = dict(trans=np.asarray(trans, dtype=np.float64), valid=False, n_valid=0, points=np.zeros((num_points, 3), np.float32),
                   ids=np.zeros(num_points, np.int32))
        if np.sum(depth_valid) > 1.0:
            roi_m_d_valid = roi_mask.astype(np.bool_) * depth_valid
            rec["n_valid"] = int(np.sum(roi_m_d_valid))
            if np.sum(roi_m_d_valid) > 1.0:
                pcl_in = R["depth_to_pcl"](roi_depth, intrinsics, roi_coord_2d, roi_mask) / 1000.0
                # sample_points draws np.random.permutation(total)[:n_pts] from the global generator when total > n_pts:
                # seed it per instance and record the ids it used so that parity mode can be given the same subset
                np.random.seed(perm_seed + i)
                if pcl_in.shape[0] > num_points:
                    state = np.random.get_state()
                    rec["ids"] = np.random.permutation(pcl_in.shape[0])[:num_points].astype(np.int32)
                    np.random.set_state(state)
                rec["points"] = R["sample_points"](pcl_in, num_points).astype(np.float32)
                rec["valid"] = True
        out.append(rec)
    return out


def generate(name):
    case = CASES[name]
    depth, masks, rois = synth.make_frame(case["seed"], case["n_inst"])
    recs = run_reference_frame(depth, masks, rois, synth.REAL_INTRINSICS, perm_seed=77 + case["seed"])
    rec = {"case_seed": np.array(case["seed"]), "case_n_inst": np.array(case["n_inst"]),
           "cs_depth": np.float64(depth.astype(np.float64).sum()), "cs_masks": np.float64(masks.sum()),
           "trans": np.stack([r["trans"] for r in recs]), "valid": np.array([r["valid"] for r in recs]),
           "n_valid": np.array([r["n_valid"] for r in recs], dtype=np.int32), "points": np.stack([r["points"] for r in recs]),
           "ids": np.stack([r["ids"] for r in recs])}
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--check", action="store_true")
    args = ap.parse_args()
    if not ref_loader.available():
        raise SystemExit("reference tree not found; goldens can only be generated in the build container")
    bad = 0
    for name in CASES:
        rec = generate(name)
        path = os.path.join(GOLDEN_DIR, name + ".npz")
        if args.check:
            old = np.load(path)
            for k in rec:
                if not np.array_equal(np.asarray(rec[k]), old[k]):
                    bad += 1
                    print(f"[MISMATCH] {name}:{k}")
            print(f"checked {name}")
        else:
            np.savez_compressed(path, **rec)
            print(f"wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB): valid {rec['valid'].tolist()} n_valid {rec['n_valid'].tolist()}")
    if bad:
        raise SystemExit(f"{bad} golden entries differ")


if __name__ == "__main__":
    main()
