"""ORACLE — test infrastructure, NOT product code (only tests/, __graft_entry__.smoke() and bench.py's CPU arm
may import it).

CPU restatement (numpy) of the reference's per-instance point-cloud preparation, i.e. the body of the instance
loop of `detect_mrcnn_genpose` (runners/evaluation_single.py:168-216): square crop window -> affine crop/resize of
the pixel-coordinate map, the instance mask and the depth image to 256 x 256 with nearest-neighbour sampling ->
back-projection of the masked, valid-depth pixels -> resampling to exactly 1024 points.

Arithmetic that lives outside /root/reference:
  * cv2.warpAffine(flags=INTER_NEAREST), OpenCV (reference pins opencv-python 4.2.0.32, requirements.txt:1; this
    container has 4.13.0).  Restated from OpenCV's published algorithm (modules/imgproc/src/imgwarp.cpp,
    cv::warpAffine + WarpAffineInvoker): the forward matrix is inverted in double precision, destination -> source
    coordinates are evaluated in 10-bit fixed point with round-half-to-even conversions (cvRound) and a +0.5
    rounding offset, pixels that fall outside the source take the border value 0.
  * cv2.getAffineTransform: solved by the caller (host side), the oracle takes the 2x3 matrix as cv2 returns it.
Pinned by tests/test_cloud_prep_oracle.py: (i) against cv2.warpAffine itself on random crops (bit-exact source
indices), (ii) against the reference's own `crop_resize_by_warp_affine`, `get_2d_coord_np` (utils/datasets_utils.py)
and its nested `depth_to_pcl` / `sample_points` functions, extracted from the unmodified source file and executed
(oracle/make_golden_prep.py; committed vectors tests/golden/prep_*.npz).
"""
import numpy as np

AB_BITS = 10
AB_SCALE = 1 << AB_BITS
IMG_SIZE = 256          # cfg.img_size (configs/config.py:78)
NUM_POINTS = 1024       # cfg.num_points (configs/config.py:24)


def get_bbox(bbox):
    """utils/sgpa_utils.py:214-242 — square crop window of a Mask-RCNN roi (y1, x1, y2, x2) on a 480 x 640 image."""
    y1, x1, y2, x2 = (int(v) for v in bbox)
    img_width, img_length = 480, 640
    window_size = (max(y2 - y1, x2 - x1) // 40 + 1) * 40
    window_size = min(window_size, 440)
    center = [(y1 + y2) // 2, (x1 + x2) // 2]
    rmin = center[0] - int(window_size / 2)
    rmax = center[0] + int(window_size / 2)
    cmin = center[1] - int(window_size / 2)
    cmax = center[1] + int(window_size / 2)
    if rmin < 0:
        rmax += -rmin
        rmin = 0
    if cmin < 0:
        cmax += -cmin
        cmin = 0
    if rmax > img_width:
        rmin -= rmax - img_width
        rmax = img_width
    if cmax > img_length:
        cmin -= cmax - img_length
        cmax = img_length
    return rmin, rmax, cmin, cmax


def crop_params(roi, im_H, im_W):
    """evaluation_single.py:170,177-184 — (bbox_center [cx, cy], scale) of the square crop."""
    rmin, rmax, cmin, cmax = get_bbox(roi)
    x1, y1, x2, y2 = cmin, rmin, cmax, rmax
    cx, cy = 0.5 * (x1 + x2), 0.5 * (y1 + y2)
    scale = max(y2 - y1, x2 - x1)
    scale = min(scale, max(im_H, im_W)) * 1.0
    return np.array([cx, cy]), scale


def affine_points(center, scale, out_size=IMG_SIZE):
    """The three point pairs get_affine_transform (utils/datasets_utils.py:97-138, rot=0, shift=0) hands to
    cv2.getAffineTransform, in float32 as there."""
    center = np.asarray(center, dtype=np.float64)
    src = np.zeros((3, 2), dtype=np.float32)
    dst = np.zeros((3, 2), dtype=np.float32)
    src_dir = np.array([0.0, scale * -0.5])                     # get_dir([0, src_w * -0.5], 0)
    dst_dir = np.array([0, out_size * -0.5], np.float32)
    src[0, :] = center
    src[1, :] = center + src_dir
    dst[0, :] = [out_size * 0.5, out_size * 0.5]
    dst[1, :] = np.array([out_size * 0.5, out_size * 0.5], np.float32) + dst_dir
    for pts in (src, dst):                                      # get_3rd_point(a, b) = b + [-(a-b)[1], (a-b)[0]]
        d = pts[0] - pts[1]
        pts[2, :] = pts[1] + np.array([-d[1], d[0]], dtype=np.float32)
    return src, dst


def solve_affine(src, dst):
    """cv2.getAffineTransform restated (6 x 6 linear system in double precision).  The host mirror uses cv2's own
    solver when cv2 is importable; this fallback agrees with it to ~1 ulp (tested), which can in principle move a fixed-point rounding; parity claims are made on cv2's matrix."""
    A = np.zeros((6, 6))
    b = np.zeros(6)
    for i in range(3):
        A[i, 0:3] = [src[i, 0], src[i, 1], 1.0]
        A[i + 3, 3:6] = [src[i, 0], src[i, 1], 1.0]
        b[i], b[i + 3] = dst[i, 0], dst[i, 1]
    return np.linalg.solve(A, b).reshape(2, 3)


def warp_source_index(trans, out_w=IMG_SIZE, out_h=IMG_SIZE):
    """cv::warpAffine, INTER_NEAREST: for every destination pixel (x, y) the source pixel (X, Y) it copies
    (before the bounds test).  trans: forward 2x3 matrix (double).  Returns int32 arrays X, Y of shape [out_h, out_w]."""
    M = np.array(trans, dtype=np.float64).reshape(6).copy()
    D = M[0] * M[4] - M[1] * M[3]
    D = 1.0 / D if D != 0 else 0.0
    A11, A22 = M[4] * D, M[0] * D
    M[0] = A11
    M[1] *= -D
    M[3] *= -D
    M[4] = A22
    b1 = -M[0] * M[2] - M[1] * M[5]
    b2 = -M[3] * M[2] - M[4] * M[5]
    M[2], M[5] = b1, b2
    x = np.arange(out_w, dtype=np.float64)
    y = np.arange(out_h, dtype=np.float64)
    adelta = np.rint(M[0] * x * AB_SCALE).astype(np.int64)       # saturate_cast<int> = cvRound = round half to even
    bdelta = np.rint(M[3] * x * AB_SCALE).astype(np.int64)
    rd = AB_SCALE // 2
    X0 = np.rint((M[1] * y + M[2]) * AB_SCALE).astype(np.int64) + rd
    Y0 = np.rint((M[4] * y + M[5]) * AB_SCALE).astype(np.int64) + rd
    X = (X0[:, None] + adelta[None, :]) >> AB_BITS
    Y = (Y0[:, None] + bdelta[None, :]) >> AB_BITS
    X = np.clip(X, -32768, 32767)                                # saturate_cast<short>
    Y = np.clip(Y, -32768, 32767)
    return X.astype(np.int32), Y.astype(np.int32)


def warp_nearest(img, trans, out_w=IMG_SIZE, out_h=IMG_SIZE):
    """cv2.warpAffine(img, trans, (out_w, out_h), flags=cv2.INTER_NEAREST), border constant 0."""
    X, Y = warp_source_index(trans, out_w, out_h)
    H, W = img.shape[:2]
    inb = (X >= 0) & (X < W) & (Y >= 0) & (Y < H)
    out = np.zeros((out_h, out_w) + img.shape[2:], dtype=img.dtype)
    out[inb] = img[Y[inb], X[inb]]
    return out


def depth_to_pcl(roi_depth, intrinsics, roi_x, roi_y, roi_mask):
    """evaluation_single.py:107-119.  intrinsics is the float32 3x3 matrix of :50/:54, so every operation below is
    float32: real_x = ((x_map - cx) * depth) / fx, real_y likewise, z = depth."""
    K = np.asarray(intrinsics, dtype=np.float32).reshape(-1)
    cx, cy, fx, fy = K[2], K[5], K[0], K[4]
    depth = roi_depth.reshape(-1).astype(np.float32)
    valid = ((depth > 0) * roi_mask.reshape(-1)) > 0
    depth = depth[valid]
    x_map = roi_x.reshape(-1).astype(np.float32)[valid]
    y_map = roi_y.reshape(-1).astype(np.float32)[valid]
    real_x = (x_map - cx) * depth / fx
    real_y = (y_map - cy) * depth / fy
    return np.stack((real_x, real_y, depth), axis=-1).astype(np.float32)


def sample_points(pcl, n_pts, ids=None):
    """evaluation_single.py:121-133.  `ids` = np.random.permutation(total)[:n_pts] of the reference's host generator
    (required when total > n_pts)."""
    total = pcl.shape[0]
    if total < n_pts:
        pcl = np.concatenate([np.tile(pcl, (n_pts // total, 1)), pcl[:n_pts % total]], axis=0)
    elif total > n_pts:
        pcl = pcl[np.asarray(ids)[:n_pts]]
    return pcl


def prepare_instance(raw_depth, inst_mask, trans, intrinsics, ids=None, ids_fn=None):
    """One instance (evaluation_single.py:171-212) -> (points [1024,3] float32 or None when the instance is skipped,
    n_valid).  inst_mask: [H,W] bool (Mask-RCNN mask of the instance); raw_depth: [H,W] uint16 (mm)."""
    H, W = raw_depth.shape
    mask = np.logical_and(inst_mask, raw_depth > 0)
    X, Y = warp_source_index(trans)
    inb = (X >= 0) & (X < W) & (Y >= 0) & (Y < H)
    Xc, Yc = np.where(inb, X, 0), np.where(inb, Y, 0)
    roi_depth = np.where(inb, raw_depth[Yc, Xc], 0).astype(raw_depth.dtype)
    roi_mask = np.where(inb, mask[Yc, Xc], False).astype(np.float32)
    roi_x = np.where(inb, Xc, 0).astype(np.float32)              # warp of the coordinate map: the source pixel's own coordinates
    roi_y = np.where(inb, Yc, 0).astype(np.float32)
    if np.sum(roi_depth > 0) <= 1.0:                             # :201-204
        return None, int(np.sum((roi_mask > 0) & (roi_depth > 0)))
    n_valid = int(np.sum(roi_mask.astype(np.bool_) * (roi_depth > 0)))
    if n_valid <= 1.0:                                           # :206-209
        return None, n_valid
    pcl = depth_to_pcl(roi_depth, intrinsics, roi_x, roi_y, roi_mask) / 1000.0
    if ids is None and ids_fn is not None and n_valid > NUM_POINTS:
        ids = ids_fn(n_valid)
    return sample_points(pcl, NUM_POINTS, ids).astype(np.float32), n_valid


# ---- throughput-mode subset: the keyed permutation the CUDA kernel uses when no ids are supplied ------------------
def _mix32(x):
    x = np.uint32(x)
    with np.errstate(over="ignore"):
        x ^= x >> np.uint32(16)
        x = np.uint32(x * np.uint32(0x7FEB352D))
        x ^= x >> np.uint32(15)
        x = np.uint32(x * np.uint32(0x846CA68B))
        x ^= x >> np.uint32(16)
    return x


def feistel_permutation_prefix(n, count, key0, key1):
    """First `count` values of a keyed pseudo-random permutation of range(n): 4-round balanced Feistel network on
    ceil(log2 n) bits (rounded up to even) with cycle walking.  Replaces np.random.permutation(n)[:count] of
    sample_points in throughput mode (same distribution family: a uniform-looking subset without replacement in
    random order; not the same stream as NumPy's Mersenne Twister, which parity mode takes as explicit ids)."""
    bits = max(2, int(np.ceil(np.log2(max(n, 2)))))
    bits += bits & 1
    half = bits // 2
    mask = (1 << half) - 1
    out = np.empty(count, dtype=np.int32)
    for j in range(count):
        v = j
        while True:
            lo, hi = v & mask, v >> half
            for rnd in range(4):
                f = int(_mix32(np.uint32(lo) ^ np.uint32((key0 + 0x9E3779B9 * rnd) & 0xFFFFFFFF)) ^ np.uint32(key1)) & mask
                lo, hi = hi ^ f, lo
            v = (hi << half) | lo
            if v < n:
                break
        out[j] = v
    return out
