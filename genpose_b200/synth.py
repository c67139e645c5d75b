"""Synthetic inputs and synthetic checkpoints (there is no REAL275 data or released checkpoint here).

Shared by bench.py, the tests and the golden-vector generator so that every arm sees identical
inputs.  Everything is drawn from numpy's legacy MT19937 `RandomState`, whose stream is stable
across numpy versions and machines, so the same seed gives bit-identical fp32 tensors on the
build container and on the GPU box (goldens store only outputs + checksums of these inputs).

Recipe follows SURVEY.md §8(d):
  * clouds   : n_valid ~ U{200..3000}; p = c + 0.05*randn; c = (U(-.2,.2), U(-.2,.2), U(.5,1.5)) metres,
               then the reference's resampling rule (runners/evaluation_single.py:120-133: tile when
               fewer than 1024 points, random subset when more) — so duplicated-point clouds occur.
  * weights  : reference key schema.  Kaiming-style random weights, NON-trivial BN
               running statistics (so the BN fold is exercised), non-zero output layers (the
               reference zero-initialises them, scorenet.py:156), plus a weak restoring component
               `out ~= -kappa*(x - mu)` routed through dedicated ReLU pairs so that the sampler's
               discrete dynamics are contractive like a trained score model's (a purely random
               score field is chaotic over 500 steps and would make ANY two fp32 implementations
               diverge — that would test conditioning, not correctness).
"""
from collections import OrderedDict
from typing import Dict

import numpy as np
import torch

from . import arch


def make_clouds(batch: int, seed: int = 0, n_pts: int = arch.NUM_POINTS) -> np.ndarray:
    """[batch, n_pts, 3] float32 camera-frame clouds (metres)."""
    rs = np.random.RandomState(1000 + seed)
    out = np.empty((batch, n_pts, 3), dtype=np.float32)
    for b in range(batch):
        n_valid = int(rs.randint(200, 3001))
        c = np.array([rs.uniform(-0.2, 0.2), rs.uniform(-0.2, 0.2), rs.uniform(0.5, 1.5)])
        # anisotropic blob so that objects are not rotationally symmetric
        scale = np.array([0.05, 0.035, 0.02]) * rs.uniform(0.8, 1.6)
        pcl = (c + rs.standard_normal((n_valid, 3)) * scale).astype(np.float32)
        if n_valid < n_pts:      # evaluation_single.py:128-129
            pcl = np.concatenate([np.tile(pcl, (n_pts // n_valid, 1)), pcl[: n_pts % n_valid]], axis=0)
        elif n_valid > n_pts:    # evaluation_single.py:130-132
            ids = rs.permutation(n_valid)[:n_pts]
            pcl = pcl[ids]
        out[b] = pcl
    return out


def batch_from_clouds(pts: np.ndarray, device="cpu") -> Dict[str, torch.Tensor]:
    """The `data` dict the runner builds (runners/evaluation_single.py:394-403)."""
    t = torch.from_numpy(np.ascontiguousarray(pts)).to(device)
    center = torch.mean(t[:, :, :3], dim=1)
    return {"pts": t, "zero_mean_pts": t - center.unsqueeze(1), "pts_center": center}


def _linear(rs, out_f, in_f, scale=1.0):
    bound = 1.0 / np.sqrt(in_f)
    w = rs.uniform(-bound, bound, (out_f, in_f)) * scale
    b = rs.uniform(-bound, bound, (out_f,)) * scale
    return w.astype(np.float32), b.astype(np.float32)


def make_state_dict(seed: int = 0, kappa: float = 1.0, alpha: float = 0.05,
                    out_scale: float = 0.02) -> "OrderedDict[str, torch.Tensor]":
    """A synthetic `model_state_dict` with the reference's exact keys/shapes (score and energy nets
    share the schema).  Deterministic in `seed`."""
    rs = np.random.RandomState(2000 + seed)
    sd = OrderedDict()

    # ---- encoder: SA_modules.{l}.mlps.{s}.layer{j}.{conv,bn.bn} ----
    for l, lv in enumerate(arch.SA_LEVELS):
        for s in range(2):
            spec = lv.mlps[s]
            for j in range(len(spec) - 1):
                cin, cout = spec[j], spec[j + 1]
                p = f"pts_encoder.SA_modules.{l}.mlps.{s}.layer{j}"
                w = rs.standard_normal((cout, cin, 1, 1)) * np.sqrt(2.0 / cin)
                if j == 0:
                    # xyz channels carry metres (relative offsets <= radius, absolute at GroupAll):
                    # scale them so that they matter as much as the O(1) feature channels
                    w[:, :3] = rs.standard_normal((cout, 3, 1, 1)) * (25.0, 12.0, 6.0, 1.0)[l]
                sd[f"{p}.conv.weight"] = torch.from_numpy(w.astype(np.float32))
                sd[f"{p}.bn.bn.weight"] = torch.from_numpy(rs.uniform(0.8, 1.2, cout).astype(np.float32))
                sd[f"{p}.bn.bn.bias"] = torch.from_numpy((rs.standard_normal(cout) * 0.1).astype(np.float32))
                sd[f"{p}.bn.bn.running_mean"] = torch.from_numpy((rs.standard_normal(cout) * 0.1).astype(np.float32))
                sd[f"{p}.bn.bn.running_var"] = torch.from_numpy(rs.uniform(0.5, 1.5, cout).astype(np.float32))
                sd[f"{p}.bn.bn.num_batches_tracked"] = torch.tensor(1000, dtype=torch.long)

    # ---- pose encoder (scorenet.py:104-109) ----
    w0, b0 = _linear(rs, 256, 9, scale=0.1)
    w2, b2 = _linear(rs, 256, 256)
    # restoring pathway: rows 0..8 = +alpha*x_i, rows 9..17 = -alpha*x_i, passed through unchanged
    w0[:18] = 0.0
    b0[:18] = 0.0
    for i in range(9):
        w0[i, i] = alpha
        w0[9 + i, i] = -alpha
    w2[:18] = 0.0
    b2[:18] = 0.0
    w2[:, :18] *= 0.1
    for i in range(18):
        w2[i, i] = 1.0
    sd["pose_score_net.pose_encoder.0.weight"] = torch.from_numpy(w0)
    sd["pose_score_net.pose_encoder.0.bias"] = torch.from_numpy(b0)
    sd["pose_score_net.pose_encoder.2.weight"] = torch.from_numpy(w2)
    sd["pose_score_net.pose_encoder.2.bias"] = torch.from_numpy(b2)

    # ---- time encoder (scorenet.py:112-117; W ~ N(0, 30^2) :61) ----
    sd["pose_score_net.t_encoder.0.W"] = torch.from_numpy((rs.standard_normal(64) * 30.0).astype(np.float32))
    wt, bt = _linear(rs, 128, 128)
    sd["pose_score_net.t_encoder.1.weight"] = torch.from_numpy(wt)
    sd["pose_score_net.t_encoder.1.bias"] = torch.from_numpy(bt)

    # ---- three heads (scorenet.py:149-170) ----
    mu = rs.standard_normal(9)
    mu[0:3] /= np.linalg.norm(mu[0:3])
    mu[3:6] /= np.linalg.norm(mu[3:6])
    mu[6:9] *= 0.02
    for k, name in enumerate(arch.HEADS):
        a, ab = _linear(rs, 256, arch.FUSED_IN)
        a[:, arch.PTS_FEAT_DIM + arch.T_EMBED_DIM:] *= 0.5
        a[:18] = 0.0
        ab[:18] = 0.0
        for i in range(18):
            a[i, arch.PTS_FEAT_DIM + arch.T_EMBED_DIM + i] = 1.0
        o = (rs.standard_normal((3, 256)) * out_scale).astype(np.float32)
        ob = np.zeros(3, dtype=np.float32)
        o[:, :18] = 0.0
        for c in range(3):
            comp = 3 * k + c
            o[c, comp] = -kappa / alpha
            o[c, 9 + comp] = +kappa / alpha
            ob[c] = kappa * mu[comp]
        sd[f"pose_score_net.fusion_tail_{name}.0.weight"] = torch.from_numpy(a)
        sd[f"pose_score_net.fusion_tail_{name}.0.bias"] = torch.from_numpy(ab)
        sd[f"pose_score_net.fusion_tail_{name}.2.weight"] = torch.from_numpy(o)
        sd[f"pose_score_net.fusion_tail_{name}.2.bias"] = torch.from_numpy(ob)
    return sd


def stable_kappa(num_steps: int) -> float:
    """Largest restoring strength (negative sign: the PC predictor as written moves AGAINST the score,
    samplers.py:147-148) whose per-step gain 17*sigma_max*|kappa|*dt stays below 0.5, capped at 0.3.
    Above gain ~2 the synthetic dynamics diverge and rounding differences are amplified without bound."""
    return -min(0.3, 0.5 * (num_steps - 1) / (17.0 * arch.SIGMA_MAX))


def make_prior_noise(rows: int, seed: int = 0, sigma: float = arch.SIGMA_MAX) -> np.ndarray:
    """x0 = sigma * randn [rows, 9] (ve_prior, sde.py:26-28); numpy stream for reproducibility."""
    rs = np.random.RandomState(3000 + seed)
    return (rs.standard_normal((rows, arch.POSE_DIM)) * sigma).astype(np.float32)


def make_step_noise(steps: int, rows: int, seed: int = 0) -> np.ndarray:
    """[steps, 2, rows, 9]: z1 (Langevin) and z2 (Euler-Maruyama) of cond_pc_sampler
    (samplers.py:131,149) in the order the reference draws them."""
    rs = np.random.RandomState(4000 + seed)
    return rs.standard_normal((steps, 2, rows, arch.POSE_DIM)).astype(np.float32)


def checksum(arr) -> float:
    a = np.asarray(arr, dtype=np.float64).ravel()
    w = np.cos(np.arange(a.size, dtype=np.float64) * 0.7853981633974483 + 0.3)
    return float(np.dot(a, w))


# ---- synthetic RGB-D frames for the point-cloud preparation path (runners/evaluation_single.py:168-216) -------------
def make_frame(seed: int, n_inst: int = 6, H: int = 480, W: int = 640):
    """A REAL275-shaped frame: depth [H,W] uint16 in millimetres (a tilted background plane with dropout holes and one
    raised blob per instance), Mask-RCNN style instance masks [H,W,n_inst] bool and rois [n_inst,4] (y1,x1,y2,x2).
    The instances cover the cases the reference distinguishes: large (> 1024 valid pixels: random subset), small
    (< 1024: tiled), touching the image border (crop window clipped / out-of-bounds ROI pixels), holes inside the
    mask (depth == 0), a mask with no valid depth at all (instance skipped) and a one-pixel mask (skipped)."""
    rs = np.random.RandomState(1000 + seed)
    yy, xx = np.mgrid[0:H, 0:W]
    depth = 900.0 + 0.35 * xx + 0.2 * yy + rs.normal(0, 2.0, (H, W))
    masks = np.zeros((H, W, n_inst), dtype=bool)
    rois = np.zeros((n_inst, 4), dtype=np.int32)
    for i in range(n_inst):
        kind = i % 6
        if kind == 0:      # large object in the middle
            cy, cx, ry, rx = rs.randint(150, 330), rs.randint(200, 440), rs.randint(50, 90), rs.randint(60, 110)
        elif kind == 1:    # small object (< 1024 pixels)
            cy, cx, ry, rx = rs.randint(60, 420), rs.randint(60, 580), rs.randint(8, 14), rs.randint(8, 14)
        elif kind == 2:    # at the top-left border
            cy, cx, ry, rx = rs.randint(5, 30), rs.randint(5, 40), rs.randint(30, 60), rs.randint(30, 60)
        elif kind == 3:    # at the bottom-right border
            cy, cx, ry, rx = H - rs.randint(5, 30), W - rs.randint(5, 40), rs.randint(40, 80), rs.randint(40, 80)
        elif kind == 4:    # no valid depth under the mask
            cy, cx, ry, rx = rs.randint(100, 380), rs.randint(100, 540), rs.randint(10, 20), rs.randint(10, 20)
        else:              # single pixel
            cy, cx, ry, rx = rs.randint(100, 380), rs.randint(100, 540), 0, 0
        m = ((yy - cy) / max(ry, 0.5)) ** 2 + ((xx - cx) / max(rx, 0.5)) ** 2 <= 1.0
        if kind == 5:
            m = (yy == cy) & (xx == cx)
        masks[:, :, i] = m
        depth = np.where(m, depth - 150.0 * np.exp(-(((yy - cy) / (ry + 1.0)) ** 2 + ((xx - cx) / (rx + 1.0)) ** 2)), depth)
        ys, xs = np.nonzero(m)
        rois[i] = [ys.min(), xs.min(), ys.max() + 1, xs.max() + 1]
    depth = np.clip(depth, 1, 60000).astype(np.uint16)
    holes = rs.rand(H, W) < 0.08
    depth[holes] = 0
    for i in range(n_inst):
        if i % 6 == 4:
            depth[masks[:, :, i]] = 0
    return depth, masks, rois


REAL_INTRINSICS = np.array([[591.0125, 0, 322.525], [0, 590.16775, 244.11084], [0, 0, 1]], dtype=np.float32)   # evaluation_single.py:54
