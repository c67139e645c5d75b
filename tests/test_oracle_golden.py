"""CPU: the portable oracle (oracle/genpose_oracle.py) against the golden vectors that
oracle/make_golden.py produced by executing the unmodified reference."""
import numpy as np
import pytest
import torch

from genpose_b200 import synth
from oracle import genpose_oracle as O
from tests import _cases


@pytest.mark.parametrize("name", _cases.golden_names())
def test_oracle_matches_reference_golden(name):
    case, g, inp = _cases.load(name)
    sd, esd = inp["sd"], inp["esd"]
    B, K = case["B"], case["K"]
    data = synth.batch_from_clouds(inp["clouds"])
    trace = O.encoder_levels(sd, data["pts"])

    # indices: bit-exact
    for l in range(3):
        assert np.array_equal(trace["fps_idx"][l].numpy(), g[f"fps_idx_l{l}"]), f"fps l{l}"
        for s in range(2):
            assert np.array_equal(trace["ball_idx"][l][s].numpy(), g[f"ball_idx_l{l}_s{s}"]), f"ball l{l}s{s}"

    feat = trace["pts_feat"]
    np.testing.assert_allclose(feat.numpy(), g["ref_pts_feat"], rtol=1e-4, atol=1e-4)

    rep = feat.unsqueeze(1).repeat(1, K, 1).view(B * K, -1)
    probe = O.score(sd, rep, _cases.t(inp["x0"]) * 0.02, torch.ones(B * K, 1) * 0.7)
    np.testing.assert_allclose(probe.numpy(), g["ref_score_probe"], rtol=1e-4, atol=1e-5)

    cen = data["pts_center"].unsqueeze(1).repeat(1, K, 1).view(B * K, -1)
    if case["sampler"] == "pc":
        pose = O.pc_sampler(sd, rep, cen, _cases.t(inp["x0"]), _cases.t(inp["step_noise"]), case["T"])
    else:
        pose = O.ode_sampler(sd, rep, cen, _cases.t(inp["x0"]), T0=case["T0"])
        assert pose.dtype == torch.float64 and g["ref_pred_pose"].dtype == np.float64   # samplers.py:206-207
    # tolerance: north_star's 1e-3 on sampled poses (the oracle in practice lands ~1e-5 from the reference)
    np.testing.assert_allclose(pose.numpy().reshape(B, K, 9), g["ref_pred_pose"], rtol=0, atol=1e-3)

    if case["energy"]:
        ref_pose = torch.from_numpy(g["ref_pred_pose"]).float()
        en = O.get_energy(esd, data, ref_pose)
        np.testing.assert_allclose(en.numpy(), g["ref_energy"], rtol=2e-4, atol=1e-2)
        sp, se, RT = O.rank_and_pool(ref_pose, torch.from_numpy(g["ref_energy"]))
        assert np.array_equal(sp.numpy(), g["ref_sorted_pose"])
        assert np.array_equal(se.numpy(), g["ref_sorted_energy"])
        np.testing.assert_allclose(RT.numpy(), g["ref_pooled_RT"], rtol=0, atol=1e-5)


def _brev10(i):
    return int(format(i, "010b")[::-1], 2)


def test_fps_tie_rules_on_duplicated_clouds():
    """Clouds with < 1024 valid points are tiled (evaluation_single.py:128-129), so exact ties are
    routine.  The block tree (__update, sampling_gpu.cu:86-91) keeps the LOWER SLOT on ties at every
    level (strides 512,...,1), i.e. among equal maxima the winner is the index whose bit-reversal is
    smallest — NOT the lowest index.  Once every temp is 0 the winner is index 0 (0 > -1, besti = 0)."""
    rs = np.random.RandomState(0)
    base = rs.standard_normal((100, 3)).astype(np.float32)
    cloud = np.concatenate([np.tile(base, (10, 1)), base[:24]], 0)[None]
    idx = O.furthest_point_sample(torch.from_numpy(cloud), 512).numpy()[0]
    assert idx[0] == 0
    assert len(set((idx[:100] % 100).tolist())) == 100            # 100 distinct points first
    for j in range(1, 100):                                        # each the min-bit-reversal duplicate
        dups = [k for k in range(1024) if k % 100 == idx[j] % 100]
        assert idx[j] == min(dups, key=_brev10), (j, idx[j])
    assert (idx[100:] == 0).all()


def test_fps_non_power_of_two_block():
    """opt_n_threads (cuda_utils.h:10-14): N=1000 -> 512 threads; brute-force argmax must agree on
    generic data (ties have measure zero)."""
    rs = np.random.RandomState(1)
    pts = rs.standard_normal((2, 1000, 3)).astype(np.float32)
    idx = O.furthest_point_sample(torch.from_numpy(pts), 64).numpy()
    for b in range(2):
        d = np.full(1000, 1e10, np.float32)
        cur = 0
        for j in range(1, 64):
            diff = pts[b] - pts[b, cur]
            dd = (diff[:, 0] * diff[:, 0] + diff[:, 1] * diff[:, 1] + diff[:, 2] * diff[:, 2]).astype(np.float32)
            d = np.minimum(d, dd)
            cur = int(np.argmax(d))
            assert abs(d[cur] - d[idx[b, j]]) <= 1e-6 * d[cur]


def test_ball_query_padding_and_empty():
    xyz = torch.tensor([[[0., 0, 0], [0.01, 0, 0], [1, 1, 1], [0.015, 0, 0]]])
    new_xyz = torch.tensor([[[0., 0, 0], [5, 5, 5]]])
    idx = O.ball_query(0.02, 4, xyz, new_xyz).numpy()[0]
    assert idx[0].tolist() == [0, 1, 3, 0]          # first hit pads the tail (ball_query_gpu.cu:35-39)
    assert idx[1].tolist() == [0, 0, 0, 0]          # no hit: stays at the caller's zeros (pointnet2_utils.py:219)
