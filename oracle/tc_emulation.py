"""TEST INFRASTRUCTURE ONLY (like everything under oracle/: imported by tests/ alone): a CPU emulation of the ARITHMETIC of the tensor-core samplers
(genpose_b200/csrc/tc_sampler.cu) on top of the oracle port, used to decide numeric design questions without a GPU and to
pin DESIGN.md §5's precision claim in the CPU suite.  It restates what the kernel computes, not how:

  * the three dense layers that run on tcgen05 (pose encoder P1, P2 and the pose block of the stacked heads) as the
    error-compensated bf16 split  A.B ~= Ahi.Bhi + Alo.Bhi + Ahi.Blo  with fp32 accumulation;
  * layer 0 from THREE bf16 pieces of the pose row (x1 + x2 + x3) against P1 hi | lo (five products, publish_x);
  * everything the kernel hoists out of the row loop in fp32: object bias A_pts.pts_feat + a, time bias A_t.relu(L_t.fourier(t)).

`terms` selects how many products of the split are kept (3 = "bf16x3", 1 = plain bf16) so that a test can show why the
split is needed; terms = "x2" is the two-product mode "f16x2" (tc_sampler.cu, TcStream<true, .>): activations fp16 hi + lo
(hi = truncation, lo = rounded residual, both through a ReLU; tc_common.cuh relu_split_f16x2), the weights of layer 1 and of the
heads ONE fp16 value each, products Ahi.W + Alo.W (layer 0 keeps its five bf16 products).  The score network it emulates is PoseScoreNet.forward (networks/gf_algorithms/scorenet.py:178-222)."""
from contextlib import contextmanager

import numpy as np
import torch
import torch.nn.functional as F

from . import genpose_oracle as O


def split_bf16(t: torch.Tensor):
    hi = t.float().to(torch.bfloat16).float()
    lo = (t.float() - hi).to(torch.bfloat16).float()
    return hi, lo


FP16_MAX = 65504.0


def split_act_f16(a: torch.Tensor):
    """relu_split_f16x2 (tc_common.cuh): hi = fp16 TRUNCATION of max(a, 0) saturated at 65504 (cvt.rz.relu.satfinite), lo =
    rn_fp16(max(a - hi, 0)) (cvt.rn.relu.satfinite)."""
    a = torch.clamp(a.float(), min=0.0)
    normal = (a.view(torch.int32) & ~0x1FFF).view(torch.float32)             # keep 10 mantissa bits (fp16 normal range)
    sub = torch.floor(a * 2.0 ** 24) * 2.0 ** -24                              # below 2^-14: multiples of the fp16 subnormal step
    hi = torch.clamp(torch.where(a >= 2.0 ** -14, normal, sub), max=FP16_MAX)
    lo = torch.clamp(a - hi, min=0.0, max=FP16_MAX).to(torch.float16).float()
    return hi, lo


def mm_split(a: torch.Tensor, w: torch.Tensor, terms=3) -> torch.Tensor:
    """a [R,K] . w [N,K]^T with both operands split into bf16 hi + lo; products are exact in fp32, sums are fp32.
    a is a post-ReLU activation, w a weight block.  terms = "x2": fp16 hi/lo activations against ONE fp16 weight."""
    if terms == "x2":
        ah, al = split_act_f16(a)
        w16 = w.float().to(torch.float16).float()
        return ah @ w16.t() + al @ w16.t()
    ah, al = split_bf16(a)
    wh, wl = split_bf16(w)
    out = ah @ wh.t()
    if terms >= 2:
        out = out + al @ wh.t()
    if terms >= 3:
        out = out + ah @ wl.t()
    return out


def trunk_tc(sd, pts_feat, pose, t, terms=3, prefix: str = "pose_score_net") -> torch.Tensor:
    g = lambda k: sd[f"{prefix}.{k}"].float()
    tt = t.float().squeeze(1)
    x_proj = tt[:, None] * g("t_encoder.0.W")[None, :] * 2 * np.pi                      # scorenet.py:63
    emb = torch.cat([torch.sin(x_proj), torch.cos(x_proj)], dim=-1)
    t_feat = torch.relu(F.linear(emb, g("t_encoder.1.weight"), g("t_encoder.1.bias")))
    x = pose.float()
    x1 = x.to(torch.bfloat16).float()
    x2 = (x - x1).to(torch.bfloat16).float()
    x3 = (x - x1 - x2).to(torch.bfloat16).float()
    ph, pl = split_bf16(g("pose_encoder.0.weight"))
    l0_terms = 3 if terms == "x2" else terms
    acc = x1 @ ph.t()
    if l0_terms >= 2:
        acc = acc + x2 @ ph.t() + x3 @ ph.t()
    if l0_terms >= 3:
        acc = acc + x1 @ pl.t() + x2 @ pl.t()
    h1 = torch.relu(acc + g("pose_encoder.0.bias"))
    pf = torch.relu(mm_split(h1, g("pose_encoder.2.weight"), terms) + g("pose_encoder.2.bias"))
    outs = []
    n_pts, n_t = pts_feat.shape[1], t_feat.shape[1]
    for name in O.HEADS:
        w = g(f"fusion_tail_{name}.0.weight")                                          # [256, 1408] = [pts | t | pose] (scorenet.py:204)
        obj_bias = F.linear(pts_feat.float(), w[:, :n_pts], g(f"fusion_tail_{name}.0.bias"))
        t_bias = F.linear(t_feat, w[:, n_pts:n_pts + n_t])
        hk = torch.relu(mm_split(pf, w[:, n_pts + n_t:], terms) + (obj_bias + t_bias))
        outs.append(F.linear(hk, g(f"fusion_tail_{name}.2.weight"), g(f"fusion_tail_{name}.2.bias")))
    return torch.cat(outs, dim=-1)


@contextmanager
def emulated_score(terms=3):
    """Inside the block every sampler of the oracle evaluates the score network with the tensor-core arithmetic."""
    orig = O.score

    def score_tc(sd, pts_feat, pose, t, dtype=torch.float32):
        f = trunk_tc(sd, pts_feat, pose, t, terms)
        return f / (O.sigma_of_t(t.float()) + 1e-7)

    O.score = score_tc
    try:
        yield
    finally:
        O.score = orig
