"""Packs a reference `model_state_dict` (checkpoint contract: networks/posenet_agent.py:117-173, key schema
SURVEY.md §8b) into the two flat fp32 blobs the C ABI consumes (layouts: DESIGN.md §3, and
genpose_b200/csrc/{encoder.cu: enc_spec/off_*, common.cuh: TrunkLayout}).

  encoder blob : for level l in 0..3, scale s in 0..1 (in that order), with channels padded to multiples of 8
                 (196 -> 200, zero filled):
                     Wx [3][c1]  Wf [cin_f][c1]  b1 [c1]  W2 [c1][c2]  b2 [c2]  W3 [c2][c3]  b3 [c3]
                 every matrix K-major ([in][out]); eval-mode BatchNorm2d folded in
                 (pytorch_utils.py:20-32, eps 1e-5):  W' = W * g / sqrt(var + eps),  b' = beta - mean * g / sqrt(var + eps)
  trunk blob   : fourier W [64] | L_t^T [128][128] | b_t | P1^T [9][256] | b | P2^T [256][256] | b |
                 A_pts [1024][768] | A_t [128][768] | A_pose [256][768] | a [768] | O [9][256] | o_b [9] (+3 pad)
                 where the three heads (rot_x, rot_y, trans; scorenet.py:149-170) are stacked along the 768
                 axis and the 1408 input columns are split as [pts_feat | t_feat | pose_feat] (scorenet.py:204).
"""
from typing import Dict

import numpy as np
import torch

from . import arch


def _pad8(c: int) -> int:
    return (c + 7) // 8 * 8


def _fold(sd: Dict[str, torch.Tensor], prefix: str):
    """conv (no bias) + eval BatchNorm -> (W' [cout, cin], b' [cout]) in float64."""
    w = sd[f"{prefix}.conv.weight"].double()[:, :, 0, 0]
    gamma = sd[f"{prefix}.bn.bn.weight"].double()
    beta = sd[f"{prefix}.bn.bn.bias"].double()
    mean = sd[f"{prefix}.bn.bn.running_mean"].double()
    var = sd[f"{prefix}.bn.bn.running_var"].double()
    scale = gamma / torch.sqrt(var + arch.BN_EPS)
    return w * scale[:, None], beta - mean * scale


def encoder_spec(l: int, s: int):
    lv = arch.SA_LEVELS[l]
    spec = lv.mlps[s]
    return dict(cin_f=lv.c_in, c1=_pad8(spec[1]), c2=_pad8(spec[2]), c3=_pad8(spec[3]),
                c1o=spec[1], c2o=spec[2], c3o=spec[3])


def pack_encoder(sd: Dict[str, torch.Tensor], prefix: str = "pts_encoder") -> torch.Tensor:
    chunks = []
    for l in range(4):
        for s in range(2):
            m = encoder_spec(l, s)
            base = f"{prefix}.SA_modules.{l}.mlps.{s}"
            w1, b1 = _fold(sd, f"{base}.layer0")
            w2, b2 = _fold(sd, f"{base}.layer1")
            w3, b3 = _fold(sd, f"{base}.layer2")
            assert w1.shape == (m["c1o"], 3 + m["cin_f"]), (w1.shape, m)
            wx = torch.zeros(3, m["c1"], dtype=torch.float64)
            wx[:, :m["c1o"]] = w1[:, :3].t()
            wf = torch.zeros(m["cin_f"], m["c1"], dtype=torch.float64)
            wf[:, :m["c1o"]] = w1[:, 3:].t()
            pb1 = torch.zeros(m["c1"], dtype=torch.float64)
            pb1[:m["c1o"]] = b1
            pw2 = torch.zeros(m["c1"], m["c2"], dtype=torch.float64)
            pw2[:m["c1o"], :m["c2o"]] = w2.t()
            pb2 = torch.zeros(m["c2"], dtype=torch.float64)
            pb2[:m["c2o"]] = b2
            pw3 = torch.zeros(m["c2"], m["c3"], dtype=torch.float64)
            pw3[:m["c2o"], :m["c3o"]] = w3.t()
            pb3 = torch.zeros(m["c3"], dtype=torch.float64)
            pb3[:m["c3o"]] = b3
            chunks += [wx.reshape(-1), wf.reshape(-1), pb1, pw2.reshape(-1), pb2, pw3.reshape(-1), pb3]
    return torch.cat(chunks).float().contiguous()


def pack_trunk(sd: Dict[str, torch.Tensor], prefix: str = "pose_score_net") -> torch.Tensor:
    g = lambda k: sd[f"{prefix}.{k}"].float()
    heads_w = [g(f"fusion_tail_{h}.0.weight") for h in arch.HEADS]      # [256, 1408] each
    heads_b = [g(f"fusion_tail_{h}.0.bias") for h in arch.HEADS]
    out_w = [g(f"fusion_tail_{h}.2.weight") for h in arch.HEADS]        # [3, 256] each
    out_b = [g(f"fusion_tail_{h}.2.bias") for h in arch.HEADS]
    a = torch.cat(heads_w, dim=0)                                        # [768, 1408], n = head*256 + j
    P, T = arch.PTS_FEAT_DIM, arch.T_EMBED_DIM
    chunks = [
        g("t_encoder.0.W").reshape(-1),                                  # fourier_w [64]
        g("t_encoder.1.weight").t().contiguous().reshape(-1),            # t_w [128 in][128 out]
        g("t_encoder.1.bias"),
        g("pose_encoder.0.weight").t().contiguous().reshape(-1),         # p1_w [9][256]
        g("pose_encoder.0.bias"),
        g("pose_encoder.2.weight").t().contiguous().reshape(-1),         # p2_w [256][256]
        g("pose_encoder.2.bias"),
        a[:, :P].t().contiguous().reshape(-1),                           # a_pts [1024][768]
        a[:, P:P + T].t().contiguous().reshape(-1),                      # a_t   [128][768]
        a[:, P + T:].t().contiguous().reshape(-1),                       # a_pose [256][768]
        torch.cat(heads_b),                                              # a_b [768]
        torch.cat(out_w, dim=0).contiguous().reshape(-1),                # o_w [9][256]
        torch.cat(out_b), torch.zeros(3),                                # o_b [9] + pad
    ]
    return torch.cat([c.float() for c in chunks]).contiguous()


def encoder_floats() -> int:
    n = 0
    for l in range(4):
        for s in range(2):
            m = encoder_spec(l, s)
            n += 3 * m["c1"] + m["cin_f"] * m["c1"] + m["c1"] + m["c1"] * m["c2"] + m["c2"] + m["c2"] * m["c3"] + m["c3"]
    return n


def trunk_floats() -> int:
    return 64 + 128 * 128 + 128 + 9 * 256 + 256 + 256 * 256 + 256 + 1024 * 768 + 128 * 768 + 256 * 768 + 768 + 9 * 256 + 12


# ---------------------------------------------------------------------------------------------------------
# tcgen05 operand images (genpose_b200/csrc/tc_common.cuh, tc_sampler.cu)
# ---------------------------------------------------------------------------------------------------------
def split_bf16(w: torch.Tensor):
    """w (fp32) -> (hi, lo) bf16 with hi = rn(w), lo = rn(w - hi)."""
    hi = w.float().to(torch.bfloat16)
    lo = (w.float() - hi.float()).to(torch.bfloat16)
    return hi, lo


def umma_image(w_bf16: torch.Tensor, variant: int = 0) -> torch.Tensor:
    """[rows, K] bf16 -> canonical K-major no-swizzle operand image as int16 words.
    variant 0 (used by the kernels): off(r,k) = (k/8)*(rows/8*128) + (r/8)*128 + (r%8)*16 + (k%8)*2  == [K/8][rows][8]
    variant 1 (self-test only)     : off(r,k) = (r/8)*(K/8*128) + (k/8)*128 + (r%8)*16 + (k%8)*2     == [rows/8][K/8][8][8]"""
    rows, K = w_bf16.shape
    assert rows % 8 == 0 and K % 8 == 0
    x = w_bf16.view(torch.int16).reshape(rows, K // 8, 8)
    if variant == 0:
        return x.permute(1, 0, 2).contiguous().reshape(-1)
    return x.reshape(rows // 8, 8, K // 8, 8).permute(0, 2, 1, 3).contiguous().reshape(-1)


def pack_encoder_tc(sd: Dict[str, torch.Tensor], prefix: str = "pts_encoder") -> torch.Tensor:
    """Tensor-core operand image of set-abstraction level 3 (csrc/sa_tc.cu), uint8:
       [8192 B: per scale  wx[3][128] | b1[128] | b2[224] | b3[256]  fp32]
       then per scale 22 slots of 16 KiB: 8 x (W2 K-step: hi [2][224][8] bf16 | lo)  and  14 x (W3 K-step: hi [2][256][8] | lo)
    with hi = rn_bf16(w), lo = rn_bf16(w - hi); 196 -> 224 zero padded (rows of W2 / columns of W3);
    followed by the level-2 block (see below)."""
    consts = torch.zeros(2048, dtype=torch.float32)
    streams = []
    for s in range(2):
        base = f"{prefix}.SA_modules.2.mlps.{s}"
        w1, b1 = _fold(sd, f"{base}.layer0")
        w2, b2 = _fold(sd, f"{base}.layer1")
        w3, b3 = _fold(sd, f"{base}.layer2")
        assert w1.shape == (128, 259) and w2.shape == (196, 128) and w3.shape == (256, 196)
        c = consts[s * 992:(s + 1) * 992]
        c[0:384] = w1[:, :3].t().reshape(-1).float()
        c[384:512] = b1.float()
        c[512:512 + 196] = b2.float()
        c[736:992] = b3.float()
        pw2 = torch.zeros(224, 128, dtype=torch.float32)
        pw2[:196] = w2.float()
        pw3 = torch.zeros(256, 224, dtype=torch.float32)
        pw3[:, :196] = w3.float()
        slots = torch.zeros(22, 8192, dtype=torch.int16)
        h2, l2 = split_bf16(pw2)
        for k in range(8):
            slots[k, 0:3584] = umma_image(h2[:, 16 * k:16 * k + 16].contiguous())
            slots[k, 3584:7168] = umma_image(l2[:, 16 * k:16 * k + 16].contiguous())
        h3, l3 = split_bf16(pw3)
        for k in range(14):
            slots[8 + k, 0:4096] = umma_image(h3[:, 16 * k:16 * k + 16].contiguous())
            slots[8 + k, 4096:8192] = umma_image(l3[:, 16 * k:16 * k + 16].contiguous())
        streams.append(slots.reshape(-1).view(torch.uint8))
    # level 2 (resident image of csrc/sa_tc.cu::sa2_tc_kernel): per scale
    #   [512 fp32: wx[3][64] | b1[64] | b2[C2] | b3[128]] [W2 hi [8][C2][8] | W2 lo | W3 hi [C2/8][128][8] | W3 lo]
    for s, c2 in ((0, 64), (1, 96)):
        base = f"{prefix}.SA_modules.1.mlps.{s}"
        w1, b1 = _fold(sd, f"{base}.layer0")
        w2, b2 = _fold(sd, f"{base}.layer1")
        w3, b3 = _fold(sd, f"{base}.layer2")
        assert w1.shape == (64, 99) and w2.shape == (c2, 64) and w3.shape == (128, c2)
        c = torch.zeros(512, dtype=torch.float32)
        c[0:192] = w1[:, :3].t().reshape(-1).float()
        c[192:256] = b1.float()
        c[256:256 + c2] = b2.float()
        c[256 + c2:384 + c2] = b3.float()
        h2, l2 = split_bf16(w2.float())
        h3, l3 = split_bf16(w3.float())
        img = torch.cat([umma_image(h2), umma_image(l2), umma_image(h3), umma_image(l3)])
        streams += [c.view(torch.uint8), img.view(torch.uint8)]
    # GroupAll (csrc/ga_tc.cu): [16 KiB fp32: per scale wx[3][256] | b1[256] | b2[384] | b3[512]] then 8 KiB slots
    # (one 128-row tile of W, one K=16 step: hi [2][128][8] | lo), ordered layer-major, scale, tile, K-step.
    gconst = torch.zeros(4096, dtype=torch.float32)
    folded = []
    for s, c2 in ((0, 256), (1, 384)):
        base = f"{prefix}.SA_modules.3.mlps.{s}"
        w1, b1 = _fold(sd, f"{base}.layer0")
        w2, b2 = _fold(sd, f"{base}.layer1")
        w3, b3 = _fold(sd, f"{base}.layer2")
        assert w1.shape == (256, 515) and w2.shape == (c2, 256) and w3.shape == (512, c2)
        c = gconst[s * 1920:(s + 1) * 1920]
        c[0:768] = w1[:, :3].t().reshape(-1).float()
        c[768:1024] = b1.float()
        c[1024:1024 + c2] = b2.float()
        c[1408:1920] = b3.float()
        folded.append((w1[:, 3:].float().contiguous(), w2.float(), w3.float()))

    def tile_slots(w):                      # w [N, K] -> per 128-row tile, per K-step: hi | lo  (int16 words)
        out = []
        hi, lo = split_bf16(w)
        for n0 in range(0, w.shape[0], 128):
            for k in range(0, w.shape[1], 16):
                out += [umma_image(hi[n0:n0 + 128, k:k + 16].contiguous()), umma_image(lo[n0:n0 + 128, k:k + 16].contiguous())]
        return out

    gslots = []
    for layer in range(3):
        for s in range(2):
            gslots += tile_slots(folded[s][layer])
    streams += [gconst.view(torch.uint8), torch.cat(gslots).view(torch.uint8)]
    # level 1, scale 1 (MLP [3,32,32,64]; same resident-image kernel as level 2): [512 fp32: wx[3][32] | b1 | b2 | b3][W2 hi|lo|W3 hi|lo]
    base = f"{prefix}.SA_modules.0.mlps.1"
    w1, b1 = _fold(sd, f"{base}.layer0")
    w2, b2 = _fold(sd, f"{base}.layer1")
    w3, b3 = _fold(sd, f"{base}.layer2")
    assert w1.shape == (32, 3) and w2.shape == (32, 32) and w3.shape == (64, 32)
    c = torch.zeros(512, dtype=torch.float32)
    c[0:96] = w1.t().reshape(-1).float()
    c[96:128] = b1.float()
    c[128:160] = b2.float()
    c[160:224] = b3.float()
    h2, l2 = split_bf16(w2.float())
    h3, l3 = split_bf16(w3.float())
    streams += [c.view(torch.uint8), torch.cat([umma_image(h2), umma_image(l2), umma_image(h3), umma_image(l3)]).view(torch.uint8)]
    return torch.cat([consts.view(torch.uint8)] + streams).contiguous()


def _trunk_tc_layouts(sd, prefix, common_slots, unit_slots):
    """Both head layouts of the tensor-core weight stream (tc_sampler.cu TcStream): A = team of 4 (per rank r the stacked head
    units [192r, 192r+128) as a 128-row unit, then [192r+128, 192r+192) as a 64-row unit), then B = teams of 2 and 1 (six 128-row
    units in column order; rank r of a team of 2 streams units 3r..3r+2).  Each layout starts with the common slots (P1, P2)."""
    g = lambda k: sd[f"{prefix}.{k}"].float()
    off = arch.PTS_FEAT_DIM + arch.T_EMBED_DIM
    stacked = torch.cat([g(f"fusion_tail_{h}.0.weight")[:, off:] for h in arch.HEADS], dim=0)      # [768, 256] (scorenet.py:204)
    lay_a, lay_b = list(common_slots), list(common_slots)
    for r in range(4):
        lay_a += unit_slots(stacked[192 * r: 192 * r + 128], False)
        lay_a += unit_slots(stacked[192 * r + 128: 192 * r + 192], True)
    for u in range(6):
        lay_b += unit_slots(stacked[128 * u: 128 * u + 128], False)
    assert len(lay_a) == len(lay_b) and all(sl.numel() * 2 == 16384 for sl in lay_a + lay_b), (len(lay_a), len(lay_b))
    return torch.cat(lay_a + lay_b).contiguous()


def pack_trunk_tc(sd: Dict[str, torch.Tensor], prefix: str = "pose_score_net") -> torch.Tensor:
    """The weight stream of the three-product tensor-core samplers: 2 layouts x 65 slots of 16 KiB (int16 words), each slot a pair
    `hi image | lo image` of K-major no-swizzle operand blocks [K/8][rows][8]:
         slot 0         : P1 (K padded 9 -> 16), output neurons [0,128): hi (4 KiB) | lo (4 KiB), then [128,256): hi | lo
         slots 1..16    : P2: for unit in ([0,128), [128,256)): for K-chunk kc in 0..7 (32 inputs): 128-row hi (8 KiB) | lo (8 KiB)
         slots 17..64   : the pose block of the three stacked heads (rows h*256 + j of fusion_tail_{rot_x,rot_y,trans}.0.weight[:, 1152:1408]):
                          128-row units, K-chunk kc: hi (8 KiB) | lo (8 KiB), 8 slots each; 64-row units (layout A), two K-chunks
                          per slot: hi (4 KiB) | lo (4 KiB), twice, 4 slots each; order per layout: _trunk_tc_layouts.
    Every CTA streams slots 0..16 plus its rank's head slots each step."""
    g = lambda k: sd[f"{prefix}.{k}"].float()
    p1 = torch.zeros(256, 16)
    p1[:, :9] = g("pose_encoder.0.weight")
    parts = []
    for unit in range(2):
        hi, lo = split_bf16(p1[128 * unit: 128 * unit + 128])
        parts += [umma_image(hi), umma_image(lo)]
    common = [torch.cat(parts)]

    def unit_slots(w_rows, small):
        chunks_per_slot = 2 if small else 1
        rows = w_rows.shape[0]
        hi, lo = split_bf16(w_rows)
        ih, il = umma_image(hi).reshape(32, rows * 8), umma_image(lo).reshape(32, rows * 8)
        out = []
        for s0 in range(0, 8, chunks_per_slot):
            pieces = []
            for kc in range(s0, s0 + chunks_per_slot):
                pieces += [ih[4 * kc: 4 * kc + 4].reshape(-1), il[4 * kc: 4 * kc + 4].reshape(-1)]
            out.append(torch.cat(pieces))
        return out

    p2 = g("pose_encoder.2.weight")
    for unit in range(2):
        common += unit_slots(p2[128 * unit: 128 * unit + 128], False)
    out = _trunk_tc_layouts(sd, prefix, common, unit_slots)
    assert out.numel() * 2 == 2 * 65 * 16384
    return out


FP16_MAX = 65504.0


def pack_trunk_tc16(sd: Dict[str, torch.Tensor], prefix: str = "pose_score_net") -> torch.Tensor:
    """Weight stream of the two-product tensor-core samplers (tc_sampler.cu, TcStream<true, .>): 2 layouts x 33 slots of 16 KiB
    (int16 words), K-major no-swizzle operand blocks [K/8][rows][8] like pack_trunk_tc, but the weights of layer 1 and of the heads
    as ONE fp16 image each (no lo image):
         slot 0          : P1 exactly as in pack_trunk_tc (bf16 hi | lo per unit; layer 0 keeps its five products)
         slots 1..8      : P2: for unit in ([0,128), [128,256)): 4 slots, slot q = inputs [64q, 64q+64) of the 128 rows (fp16)
         slots 9..32     : head units: 128-row units 4 slots (K = 64 each), 64-row units (layout A) 2 slots (K = 128 each)
    Weights beyond the fp16 range are refused (use the three-product stream)."""
    g = lambda k: sd[f"{prefix}.{k}"].float()
    common = [pack_trunk_tc(sd, prefix)[:8192]]                                   # slot 0: 16 KiB = 8192 int16 words

    def unit_slots16(w_rows, small):
        k_per_slot = 128 if small else 64
        if float(w_rows.abs().max()) >= FP16_MAX:
            raise ValueError("pack_trunk_tc16: a weight exceeds the fp16 range; use pack_trunk_tc (bf16x3)")
        rows = w_rows.shape[0]
        img = umma_image(w_rows.to(torch.float16).view(torch.int16).view(rows, -1).view(torch.bfloat16)).reshape(32, rows * 8)
        return [img[k0 // 8: (k0 + k_per_slot) // 8].reshape(-1) for k0 in range(0, 256, k_per_slot)]

    p2 = g("pose_encoder.2.weight")
    for unit in range(2):
        common += unit_slots16(p2[128 * unit: 128 * unit + 128], False)
    out = _trunk_tc_layouts(sd, prefix, common, unit_slots16)
    assert out.numel() * 2 == 2 * 33 * 16384
    return out
