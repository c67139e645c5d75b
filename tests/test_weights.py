"""CPU: host-side weight packing (BN fold, head stacking, W1 column split) against the oracle."""
import numpy as np
import pytest
import torch

from genpose_b200 import arch, synth, weights
from oracle import genpose_oracle as O


def test_pack_sizes_and_padding():
    sd = synth.make_state_dict(0)
    enc = weights.pack_encoder(sd)
    trunk = weights.pack_trunk(sd)
    assert enc.numel() == weights.encoder_floats() and trunk.numel() == weights.trunk_floats()
    assert enc.dtype == torch.float32 and trunk.dtype == torch.float32


def test_bn_fold_reproduces_shared_mlp():
    """relu(W'.x + b') == relu(bn(conv(x))) for every folded layer (pytorch_utils.py:20-32)."""
    sd = synth.make_state_dict(3)
    rs = np.random.RandomState(0)
    for l in range(4):
        for s in range(2):
            spec = arch.SA_LEVELS[l].mlps[s]
            x = torch.from_numpy(rs.standard_normal((50, spec[0])).astype(np.float32))
            ref = O._shared_mlp(sd, f"pts_encoder.SA_modules.{l}.mlps.{s}", x, 3, torch.float32)
            h = x.double()
            for j in range(3):
                w, b = weights._fold(sd, f"pts_encoder.SA_modules.{l}.mlps.{s}.layer{j}")
                h = torch.relu(h @ w.t() + b)
            np.testing.assert_allclose(h.float().numpy(), ref.numpy(), rtol=2e-5, atol=2e-5)


def test_two_product_stream_layout():
    """pack_trunk_tc16: every slot is the K-major no-swizzle image [K/8][rows][8] of the fp16-rounded weight block the kernel's
    descriptors expect (tc_sampler.cu, TcStream<true, .>); slot 0 is the three-product stream's P1 slot; two layouts back to back:
    A (team of 4: per rank a 128-row and a 64-row head unit) and B (teams of 2 and 1: six 128-row head units)."""
    sd = synth.make_state_dict(3, kappa=0.3)
    st = weights.pack_trunk_tc16(sd)
    assert st.dtype == torch.int16 and st.numel() * 2 == 2 * 33 * 16384
    assert torch.equal(st[:8192], weights.pack_trunk_tc(sd)[:8192])
    slot = lambda i: st[i * 8192:(i + 1) * 8192].view(torch.float16)
    r16 = lambda w: w.to(torch.float16).float()
    p2 = sd["pose_score_net.pose_encoder.2.weight"].float()
    for unit in range(2):
        for q in range(4):
            w = slot(1 + unit * 4 + q).reshape(8, 128, 8).permute(1, 0, 2).reshape(128, 64).float()
            assert torch.equal(w, r16(p2[128 * unit:128 * unit + 128, 64 * q:64 * q + 64]))
    stacked = torch.cat([sd[f"pose_score_net.fusion_tail_{h}.0.weight"].float()[:, 1152:] for h in ("rot_x", "rot_y", "trans")], 0)
    for r in range(4):
        for q in range(4):
            w = slot(9 + 6 * r + q).reshape(8, 128, 8).permute(1, 0, 2).reshape(128, 64).float()
            assert torch.equal(w, r16(stacked[192 * r:192 * r + 128, 64 * q:64 * q + 64]))
        for q in range(2):
            w = slot(13 + 6 * r + q).reshape(16, 64, 8).permute(1, 0, 2).reshape(64, 128).float()
            assert torch.equal(w, r16(stacked[192 * r + 128:192 * r + 192, 128 * q:128 * q + 128]))
    assert torch.equal(st[33 * 8192: 42 * 8192], st[: 9 * 8192])              # layout B starts with the same common slots
    for u in range(6):
        for q in range(4):
            w = slot(33 + 9 + 4 * u + q).reshape(8, 128, 8).permute(1, 0, 2).reshape(128, 64).float()
            assert torch.equal(w, r16(stacked[128 * u:128 * u + 128, 64 * q:64 * q + 64]))
    sd_big = dict(sd)
    sd_big["pose_score_net.pose_encoder.2.weight"] = sd["pose_score_net.pose_encoder.2.weight"] * 1e6
    with pytest.raises(ValueError):
        weights.pack_trunk_tc16(sd_big)


def test_three_product_stream_layouts():
    """pack_trunk_tc: layout A then layout B, 65 slots each; a head slot = 128-row hi image (8 KiB) | lo image of one 32-input K-chunk."""
    sd = synth.make_state_dict(4, kappa=0.3)
    st = weights.pack_trunk_tc(sd)
    assert st.dtype == torch.int16 and st.numel() * 2 == 2 * 65 * 16384
    assert torch.equal(st[65 * 8192: 82 * 8192], st[: 17 * 8192])
    stacked = torch.cat([sd[f"pose_score_net.fusion_tail_{h}.0.weight"].float()[:, 1152:] for h in ("rot_x", "rot_y", "trans")], 0)
    slot = lambda i: st[i * 8192:(i + 1) * 8192].view(torch.bfloat16)
    for u in range(6):
        for kc in range(8):
            sl = slot(65 + 17 + 8 * u + kc).float()
            hi = sl[:4096].reshape(4, 128, 8).permute(1, 0, 2).reshape(128, 32)
            lo = sl[4096:].reshape(4, 128, 8).permute(1, 0, 2).reshape(128, 32)
            w = stacked[128 * u:128 * u + 128, 32 * kc:32 * kc + 32]
            wh, wl = weights.split_bf16(w)
            assert torch.equal(hi, wh.float()) and torch.equal(lo, wl.float())
    # layout A, rank 1, 64-row unit, slot 2: K-chunks 4 and 5
    sl = slot(17 + 12 * 1 + 8 + 2).float()
    for c in range(2):
        hi = sl[c * 4096: c * 4096 + 2048].reshape(4, 64, 8).permute(1, 0, 2).reshape(64, 32)
        wh, _ = weights.split_bf16(stacked[192 + 128:192 + 192, 32 * (4 + c):32 * (5 + c)])
        assert torch.equal(hi, wh.float())


def test_trunk_pack_reproduces_score():
    """Evaluate the hoisted form  O.relu(A_pose.pf + (A_pts.feat + a) + A_t.tf)  from the packed blob
    with numpy and compare with the oracle's literal PoseScoreNet.forward."""
    sd = synth.make_state_dict(5)
    blob = weights.pack_trunk(sd).double().numpy()
    off = 0

    def take(n):
        nonlocal off
        out = blob[off:off + n]
        off += n
        return out
    fw = take(64); tw = take(128 * 128).reshape(128, 128); tb = take(128)
    p1 = take(9 * 256).reshape(9, 256); p1b = take(256); p2 = take(256 * 256).reshape(256, 256); p2b = take(256)
    apts = take(1024 * 768).reshape(1024, 768); at = take(128 * 768).reshape(128, 768)
    apose = take(256 * 768).reshape(256, 768); ab = take(768); ow = take(9 * 256).reshape(9, 256); ob = take(9)
    rs = np.random.RandomState(1)
    R = 6
    feat = rs.standard_normal((R, 1024)); pose = rs.standard_normal((R, 9)); t = 0.37
    xp = t * fw * 2 * np.pi
    tf = np.maximum(np.concatenate([np.sin(xp), np.cos(xp)]) @ tw + tb, 0)
    pf = np.maximum(np.maximum(pose @ p1 + p1b, 0) @ p2 + p2b, 0)
    h = np.maximum(pf @ apose + (feat @ apts + ab) + tf @ at, 0)
    f = np.stack([(h[:, (c // 3) * 256:(c // 3 + 1) * 256] * ow[c]).sum(1) + ob[c] for c in range(9)], 1)
    ref = O.score(sd, torch.from_numpy(feat), torch.from_numpy(pose), torch.ones(R, 1, dtype=torch.float64) * t,
                  dtype=torch.float64)
    std = 0.01 * 5000.0 ** t
    np.testing.assert_allclose(f / (std + 1e-7), ref.numpy(), rtol=1e-5, atol=1e-6)
