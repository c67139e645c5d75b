"""Host-side mirror of the per-frame point-cloud preparation of the reference runner
(`detect_mrcnn_genpose`, runners/evaluation_single.py:156-216): from a depth image and the Mask-RCNN detections of a
frame to the `[n,1024,3]` clouds the pose pipeline consumes — crop windows and affine matrices on the host (a few
scalars per instance), everything per pixel in ONE launch of `gpb_prepare_clouds` (csrc/cloud_prep.cu).

    pts, valid_inst, n_valid = prepare_frame(raw_depth, masks, rois, intrinsics)
    data = PosePipeline.make_batch(pts[valid_inst])           # evaluation_single.py:394-403

Same quantities, names and skip rule as the reference: `valid_inst` lists the instances with more than one valid crop
pixel (:201-209, the others keep an identity pose there).  There is no CPU path: tensors must live on the GPU."""
from typing import Optional, Sequence, Tuple

import ctypes

import numpy as np
import torch

from . import lib

IMG_SIZE = 256        # cfg.img_size   (configs/config.py:78)
NUM_POINTS = 1024     # cfg.num_points (configs/config.py:24)


def get_bbox(bbox) -> Tuple[int, int, int, int]:
    """utils/sgpa_utils.py:214-242: square crop window (rmin, rmax, cmin, cmax) of a roi (y1, x1, y2, x2), 480 x 640 frames."""
    y1, x1, y2, x2 = (int(v) for v in bbox)
    img_width, img_length = 480, 640
    window_size = min((max(y2 - y1, x2 - x1) // 40 + 1) * 40, 440)
    center = [(y1 + y2) // 2, (x1 + x2) // 2]
    rmin, rmax = center[0] - int(window_size / 2), center[0] + int(window_size / 2)
    cmin, cmax = center[1] - int(window_size / 2), center[1] + int(window_size / 2)
    if rmin < 0:
        rmin, rmax = 0, rmax - rmin
    if cmin < 0:
        cmin, cmax = 0, cmax - cmin
    if rmax > img_width:
        rmin, rmax = rmin - (rmax - img_width), img_width
    if cmax > img_length:
        cmin, cmax = cmin - (cmax - img_length), img_length
    return rmin, rmax, cmin, cmax


def crop_transform(roi, im_H: int, im_W: int, out_size: int = IMG_SIZE) -> np.ndarray:
    """The forward 2x3 matrix `crop_resize_by_warp_affine` builds for a roi (evaluation_single.py:170-184 +
    utils/datasets_utils.py:97-138 with rot = 0): cv2.getAffineTransform on the same three float32 point pairs when
    cv2 is importable (bit-identical to the reference), else the same linear system solved with numpy (~1 ulp)."""
    rmin, rmax, cmin, cmax = get_bbox(roi)
    x1, y1, x2, y2 = cmin, rmin, cmax, rmax
    center = np.array([0.5 * (x1 + x2), 0.5 * (y1 + y2)])
    scale = min(max(y2 - y1, x2 - x1), max(im_H, im_W)) * 1.0
    src = np.zeros((3, 2), dtype=np.float32)
    dst = np.zeros((3, 2), dtype=np.float32)
    src[0, :] = center
    src[1, :] = center + np.array([0.0, scale * -0.5])
    dst[0, :] = [out_size * 0.5, out_size * 0.5]
    dst[1, :] = np.array([out_size * 0.5, out_size * 0.5], np.float32) + np.array([0, out_size * -0.5], np.float32)
    for pts in (src, dst):                                         # get_3rd_point
        d = pts[0] - pts[1]
        pts[2, :] = pts[1] + np.array([-d[1], d[0]], dtype=np.float32)
    try:
        import cv2
        return np.asarray(cv2.getAffineTransform(np.float32(src), np.float32(dst)), dtype=np.float64)
    except ImportError:
        A = np.zeros((6, 6))
        b = np.zeros(6)
        for i in range(3):
            A[i, 0:3] = [src[i, 0], src[i, 1], 1.0]
            A[i + 3, 3:6] = [src[i, 0], src[i, 1], 1.0]
            b[i], b[i + 3] = dst[i, 0], dst[i, 1]
        return np.linalg.solve(A, b).reshape(2, 3)


def prepare_clouds(depth: torch.Tensor, masks: torch.Tensor, trans: torch.Tensor, intrinsics,
                   subset_ids: Optional[torch.Tensor] = None, seed: int = 0):
    """gpb_prepare_clouds: depth [H,W] (int16/uint16 storage, CUDA), masks [H,W,n] bool/uint8 (CUDA, the detector's layout),
    trans [n,6] or [n,2,3] float64 (CUDA), intrinsics 3x3 (fx, fy on the diagonal, cx, cy in the last column) or
    (cx, cy, fx, fy).  -> pts [n,1024,3] float32, n_valid [n] int32."""
    if not (depth.is_cuda and masks.is_cuda and trans.is_cuda):
        raise lib.GenPoseB200Error("prepare_clouds: expected CUDA tensors (there is no CPU path)")
    if depth.dtype not in (torch.int16, torch.uint16) or not depth.is_contiguous() or depth.dim() != 2:
        raise lib.GenPoseB200Error("prepare_clouds: depth must be a contiguous [H,W] 16-bit tensor (millimetres)")
    if masks.dtype not in (torch.bool, torch.uint8) or masks.dim() != 3 or tuple(masks.shape[:2]) != tuple(depth.shape):
        raise lib.GenPoseB200Error("prepare_clouds: masks must be [H,W,n] bool / uint8")
    if masks.shape[2] and masks.stride(0) != masks.shape[1] * masks.stride(1):
        raise lib.GenPoseB200Error("prepare_clouds: masks rows must be dense (stride(0) == W * stride(1))")
    if trans.dtype != torch.float64 or not trans.is_contiguous():
        raise lib.GenPoseB200Error("prepare_clouds: trans must be contiguous float64")
    H, W = depth.shape
    n = masks.shape[2]
    if trans.numel() != n * 6:
        raise lib.GenPoseB200Error(f"prepare_clouds: trans has {trans.numel()} values, expected {n * 6}")
    K = np.asarray(intrinsics, dtype=np.float32)
    k4 = np.array([K[0, 2], K[1, 2], K[0, 0], K[1, 1]], dtype=np.float32) if K.shape == (3, 3) else K.reshape(4)
    k4c = (ctypes.c_float * 4)(*[float(v) for v in k4])
    if subset_ids is not None:
        if subset_ids.dtype != torch.int32 or not subset_ids.is_cuda or not subset_ids.is_contiguous() or tuple(subset_ids.shape) != (n, NUM_POINTS):
            raise lib.GenPoseB200Error("prepare_clouds: subset_ids must be a contiguous CUDA int32 [n,1024] tensor")
    pts = torch.empty(n, NUM_POINTS, 3, dtype=torch.float32, device=depth.device)
    n_valid = torch.empty(n, dtype=torch.int32, device=depth.device)
    lib.check(lib.load().gpb_prepare_clouds(depth.data_ptr(), masks.data_ptr(), masks.stride(1), masks.stride(2),
                                            H, W, n, trans.data_ptr(), k4c, 0 if subset_ids is None else subset_ids.data_ptr(),
                                            int(seed) & (2 ** 64 - 1), pts.data_ptr(), n_valid.data_ptr(),
                                            torch.cuda.current_stream().cuda_stream), "prepare_clouds")
    return pts, n_valid


def prepare_frame(raw_depth, masks, rois: Sequence, intrinsics, subset_ids=None, seed: int = 0, device="cuda"):
    """One frame of detect_mrcnn_genpose (evaluation_single.py:156-216).  raw_depth [H,W] uint16 (numpy or tensor),
    masks [H,W,n] bool, rois [n,4] (y1,x1,y2,x2) as the Mask-RCNN pickles hold them.
    -> (pts [n,1024,3] CUDA float32, valid_inst list[int], n_valid [n] CUDA int32)."""
    d = torch.as_tensor(np.ascontiguousarray(raw_depth).view(np.int16) if isinstance(raw_depth, np.ndarray) else raw_depth)
    m = torch.as_tensor(np.ascontiguousarray(masks) if isinstance(masks, np.ndarray) else masks)
    if m.stride(1) != m.shape[2] * m.stride(2) or m.stride(0) != m.shape[1] * m.stride(1):
        m = m.contiguous()
    H, W = d.shape
    trans = np.stack([crop_transform(r, H, W) for r in rois]) if len(rois) else np.zeros((0, 2, 3))
    ids = None if subset_ids is None else torch.as_tensor(np.ascontiguousarray(subset_ids, dtype=np.int32)).to(device)
    pts, n_valid = prepare_clouds(d.to(device), m.to(device), torch.from_numpy(np.ascontiguousarray(trans, dtype=np.float64)).to(device),
                                  intrinsics, subset_ids=ids, seed=seed)
    valid_inst = [i for i, v in enumerate(n_valid.cpu().tolist()) if v > 1]
    return pts, valid_inst, n_valid
