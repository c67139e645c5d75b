"""GPU parity tests proper: the CUDA path (through the C ABI, genpose_b200/ops.py) against
(a) the golden vectors produced by executing the unmodified reference and (b) the portable oracle on
fresh seeds.  Indices are compared bit-exactly; floating point with the tolerance stated per check."""
import numpy as np
import pytest
import torch

from genpose_b200 import arch, synth
from oracle import genpose_oracle as O
from tests import _cases

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from genpose_b200 import ops as _ops
    return _ops


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


# ------------------------------------------------------------------------------------------------------
# compat layer: bit-exact indices
# ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,m", [(1024, 512), (512, 256), (256, 128), (1000, 64), (37, 37), (2, 2), (1, 1), (4096, 100)])
def test_fps_bit_exact(ops, n, m):
    rs = np.random.RandomState(n * 7 + m)
    xyz = rs.standard_normal((5, n, 3)).astype(np.float32)
    if n >= 64:   # tiled duplicates (evaluation_single.py:128-129): exact ties everywhere
        k = n // 5
        xyz[1] = np.concatenate([np.tile(xyz[1, :k], (n // k, 1)), xyz[1, : n % k]], 0)
        xyz[2] = xyz[2, 0]   # all points identical
    ref = O.furthest_point_sample(torch.from_numpy(xyz), m).numpy()
    got = ops.furthest_point_sample(_dev(xyz), m).cpu().numpy()
    assert np.array_equal(got, ref)


@pytest.mark.parametrize("n,m,r,ns", [(1024, 512, 0.02, 16), (1024, 512, 0.04, 32), (512, 256, 0.08, 32),
                                      (256, 128, 0.16, 32), (300, 77, 0.5, 5), (64, 64, 1e-6, 8), (20000, 33, 0.3, 64)])
def test_ball_query_bit_exact(ops, n, m, r, ns):
    rs = np.random.RandomState(n + m)
    scale = 0.05 if r < 0.2 else 1.0
    xyz = (rs.standard_normal((3, n, 3)) * scale).astype(np.float32)
    xyz[1, n // 2:] = xyz[1, : n - n // 2]          # duplicates
    new_xyz = xyz[:, rs.permutation(n)[:m]].copy()
    new_xyz[2, 0] = 100.0                            # a centre with an empty ball
    ref = O.ball_query(r, ns, torch.from_numpy(xyz), torch.from_numpy(new_xyz)).numpy()
    got = ops.ball_query(r, ns, _dev(xyz), _dev(new_xyz)).cpu().numpy()
    assert np.array_equal(got, ref)


def test_gather_and_group(ops):
    rs = np.random.RandomState(3)
    pts = rs.standard_normal((2, 7, 100)).astype(np.float32)
    idx = rs.randint(0, 100, (2, 40)).astype(np.int32)
    got = ops.gather_points(_dev(pts), _dev(idx)).cpu().numpy()
    assert np.array_equal(got, np.take_along_axis(pts, idx[:, None, :].astype(np.int64).repeat(7, 1), axis=2))
    gidx = rs.randint(0, 100, (2, 10, 6)).astype(np.int32)
    got = ops.group_points(_dev(pts), _dev(gidx)).cpu().numpy()
    ref = np.stack([pts[b][:, gidx[b]] for b in range(2)])
    assert np.array_equal(got, ref)


# ------------------------------------------------------------------------------------------------------
# fused path against the reference goldens
# ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", [n for n in _cases.golden_names() if "kappa0" not in n])      # kappa0: tests/test_gpu_tc_teams.py::test_undamped_dynamics_golden
def test_fused_path_matches_reference_golden(ops, name):
    case, g, inp = _cases.load(name)
    B, K = case["B"], case["K"]
    eng = ops.Engine(inp["sd"])
    data = synth.batch_from_clouds(inp["clouds"], device="cuda")

    feat, fps = eng.encode(data["pts"], return_fps=True)
    for l in range(3):
        assert np.array_equal(fps[l].cpu().numpy(), g[f"fps_idx_l{l}"]), f"fps level {l}"
    # tolerance 1e-4 relative to the feature scale
    ref_feat = g["ref_pts_feat"]
    np.testing.assert_allclose(feat.cpu().numpy(), ref_feat, rtol=1e-4, atol=1e-4 * max(1.0, np.abs(ref_feat).max()))

    ob = eng.object_bias(feat)
    x0 = _dev(inp["x0"])
    probe = eng.trunk_eval(ob, (x0 * 0.02).contiguous(), K, 0.7, divide_mode=1)
    np.testing.assert_allclose(probe.cpu().numpy(), g["ref_score_probe"], rtol=1e-4, atol=1e-5)

    if case["sampler"] == "pc":
        pose = eng.sample_pc(ob, data["pts_center"], x0, K, case["T"], step_noise=_dev(inp["step_noise"]))
    else:
        pose, stats = eng.sample_ode(ob, data["pts_center"], x0, K, T0=case["T0"])
        assert pose.dtype == torch.float64 and int(stats[3]) == 0
    # north_star: 1e-3 on sampled SE(3) poses under fixed seed.  The adaptive RK45 controller (rtol=atol=1e-5,
    # one error norm for the whole batch) is itself only reproducible to ~2e-4 RELATIVE across rounding
    # differences (the oracle port differs from the live reference by that much, tests/test_oracle_vs_reference.py),
    # so ODE translations of magnitude ~16 m get a matching relative term.
    rtol = 2e-4 if case["sampler"] == "ode" else 0
    np.testing.assert_allclose(pose.cpu().numpy().reshape(B, K, 9), g["ref_pred_pose"], rtol=rtol, atol=1e-3)

    if case["energy"]:
        eeng = ops.Engine(inp["esd"])
        efeat = eeng.encode(data["pts"])
        eob = eeng.object_bias(efeat)
        ref_pose = _dev(g["ref_pred_pose"].astype(np.float32).reshape(B * K, 9))
        en = eeng.energy(eob, data["pts_center"], ref_pose, K, 1e-5).reshape(B, K, 2)
        np.testing.assert_allclose(en.cpu().numpy(), g["ref_energy"], rtol=2e-4, atol=1e-2)
        sp, se, rt = ops.rank_pool(ref_pose.reshape(B, K, 9), _dev(g["ref_energy"]))
        assert np.array_equal(sp.cpu().numpy(), g["ref_sorted_pose"])
        assert np.array_equal(se.cpu().numpy(), g["ref_sorted_energy"])
        np.testing.assert_allclose(rt.cpu().numpy(), g["ref_pooled_RT"], rtol=0, atol=1e-5)


# ------------------------------------------------------------------------------------------------------
# fused path against the oracle on fresh seeds / awkward sizes
# ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("precision", ["fp32", "bf16x3"])
def test_encoder_levels_against_oracle(ops, precision):
    seed, B = 21, 5
    sd = synth.make_state_dict(seed, kappa=-0.3)
    clouds = synth.make_clouds(B, seed)
    clouds[1] = clouds[1, 0]                              # degenerate: every point identical
    trace = O.encoder_levels(sd, torch.from_numpy(clouds))
    eng = ops.Engine(sd)
    feat, fps = eng.encode(_dev(clouds), return_fps=True, precision=precision)
    for l in range(3):
        assert np.array_equal(fps[l].cpu().numpy(), trace["fps_idx"][l].numpy())
    ref = trace["pts_feat"].numpy()
    np.testing.assert_allclose(feat.cpu().numpy(), ref, rtol=1e-4, atol=1e-4 * max(1.0, np.abs(ref).max()))
    if precision == "bf16x3":      # tensor-core level 3 against the FFMA level 3: the bf16x3 split keeps ~fp32 accuracy
        f32 = eng.encode(_dev(clouds), precision="fp32")
        err = (feat - f32).abs().max().item() / max(1.0, f32.abs().max().item())
        assert err < 5e-5, err


@pytest.mark.parametrize("B,K,T", [(1, 1, 10), (3, 7, 40), (7, 50, 30), (2, 30, 500)])
def test_pc_sampler_against_oracle(ops, B, K, T):
    seed = 30 + B
    # the predictor's per-step gain is 17*sigma*|kappa|*dt: keep it < 1 at sigma = 50 (T = 10 needs a weaker field),
    # otherwise the synthetic dynamics themselves diverge and no two fp32 implementations agree
    sd = synth.make_state_dict(seed, kappa=synth.stable_kappa(T))
    clouds = synth.make_clouds(B, seed)
    x0 = synth.make_prior_noise(B * K, seed)
    sn = synth.make_step_noise(T, B * K, seed)
    data = synth.batch_from_clouds(clouds)
    eng = ops.Engine(sd)
    # This test isolates the fp32 sampler: FFMA encoder, and the oracle samples from ITS features.  (At T = 10 the first
    # predictor step multiplies a score perturbation by g^2*dt = 4.7e3, so an encoder's 1e-5 feature difference — checked on
    # its own in test_encoder_levels_against_oracle — alone moves a translation by 1.5e-3; the end-to-end anchor is
    # test_fused_path_matches_reference_golden.)
    feat = eng.encode(_dev(clouds), precision="fp32")
    ref_pose, _ = O.pred_func_pc(sd, data, K, T, torch.from_numpy(x0), torch.from_numpy(sn), pts_feat=feat.cpu())
    ob = eng.object_bias(feat)
    pose, proc = eng.sample_pc(ob, data["pts_center"].cuda(), _dev(x0), K, T, step_noise=_dev(sn), return_process=True)
    # 1e-3 absolute on poses of natural scale; the short chains (T = 10, 30) stop with synthetic translations of O(100),
    # where 1e-3 would be ~1 fp32 ulp, so a relative term (5e-5, ~6 ulp) covers them
    np.testing.assert_allclose(pose.cpu().numpy().reshape(B, K, 9), ref_pose.numpy(), rtol=5e-5, atol=1e-3)
    assert torch.isfinite(proc).all()


def test_pc_sampler_philox_mode_statistics(ops):
    """Throughput mode draws z1, z2 in-kernel (Philox4x32-10 + Box-Muller).  It cannot match the reference's
    generator bit for bit; check it is deterministic in the seed, changes with it, and keeps the rotation
    part orthonormal."""
    seed, B, K, T = 5, 4, 25, 50
    sd = synth.make_state_dict(seed, kappa=synth.stable_kappa(T))
    eng = ops.Engine(sd)
    clouds = synth.make_clouds(B, seed)
    data = synth.batch_from_clouds(clouds, device="cuda")
    ob = eng.object_bias(eng.encode(data["pts"]))
    x0 = _dev(synth.make_prior_noise(B * K, seed))
    a = eng.sample_pc(ob, data["pts_center"], x0, K, T, seed=123)
    b = eng.sample_pc(ob, data["pts_center"], x0, K, T, seed=123)
    c = eng.sample_pc(ob, data["pts_center"], x0, K, T, seed=124)
    assert torch.equal(a, b) and not torch.equal(a, c)
    r1, r2 = a[:, 0:3], a[:, 3:6]
    assert torch.allclose(r1.norm(dim=1), torch.ones_like(r1[:, 0]), atol=1e-5)
    assert torch.allclose((r1 * r2).sum(1), torch.zeros_like(r1[:, 0]), atol=1e-5)


@pytest.mark.parametrize("B,K,T0", [(2, 5, 0.55), (3, 50, 0.55), (1, 1, 0.15)])
def test_ode_sampler_against_oracle(ops, B, K, T0):
    seed = 40 + B
    sd = synth.make_state_dict(seed, kappa=0.3)
    clouds = synth.make_clouds(B, seed)
    sig = float(O.sigma_of_t(torch.tensor(T0)))
    x0 = synth.make_prior_noise(B * K, seed, sigma=sig)
    data = synth.batch_from_clouds(clouds)
    feat = O.encode(sd, data["pts"])
    rep = feat.unsqueeze(1).repeat(1, K, 1).view(B * K, -1)
    cen = data["pts_center"].unsqueeze(1).repeat(1, K, 1).view(B * K, -1)
    ref, st = O.ode_sampler(sd, rep, cen, torch.from_numpy(x0), T0=T0, return_stats=True)
    eng = ops.Engine(sd)
    ob = eng.object_bias(eng.encode(_dev(clouds)))
    pose, stats = eng.sample_ode(ob, data["pts_center"].cuda(), _dev(x0), K, T0=T0)
    stats = stats.cpu().numpy()
    assert stats[3] == 0
    # the controller is discrete: the same accept/reject sequence is expected on well-conditioned cases
    assert abs(int(stats[0]) - st["nfev"]) <= 12, (stats, st)
    np.testing.assert_allclose(pose.cpu().numpy(), ref.numpy(), rtol=0, atol=1e-3)


def test_rank_pool_against_oracle(ops):
    rs = np.random.RandomState(9)
    B, K = 6, 50
    pose = rs.standard_normal((B, K, 9)).astype(np.float32)
    energy = rs.standard_normal((B, K, 2)).astype(np.float32)
    sp_ref, se_ref, rt_ref = O.rank_and_pool(torch.from_numpy(pose), torch.from_numpy(energy))
    sp, se, rt = ops.rank_pool(_dev(pose), _dev(energy))
    assert np.array_equal(sp.cpu().numpy(), sp_ref.numpy())
    assert np.array_equal(se.cpu().numpy(), se_ref.numpy())
    np.testing.assert_allclose(rt.cpu().numpy(), rt_ref.numpy(), rtol=0, atol=2e-5)


def test_no_cpu_path(ops):
    from genpose_b200 import lib
    with pytest.raises(lib.GenPoseB200Error):
        ops.furthest_point_sample(torch.zeros(1, 8, 3), 4)


def test_tracking_warm_start_matches_oracle(ops):
    """The tracking hand-off (runners/evaluation_tracking.py:302-316): init_x from the previous frame's pose, ODE sampler
    from T0 = 0.15.  The agent draws the T0-prior noise from torch's CPU generator like the reference (sde.py:26-28), so a
    seeded run can be replayed by the oracle."""
    from genpose_b200.pipeline import PosePipeline
    from genpose_b200.sde import init_sde
    ve_prior = init_sde("ve")[0]               # sigma_max = 50 (sde.py:90-97)
    B, K, T0, seed = 3, 50, 0.15, 21
    sd = synth.make_state_dict(seed, kappa=0.3)
    esd = synth.make_state_dict(seed + 100, kappa=0.3)
    clouds = synth.make_clouds(B, seed)
    data_cpu = synth.batch_from_clouds(clouds)
    rs = np.random.RandomState(seed)
    init_sRT = torch.eye(4).repeat(B, 1, 1)
    q, _ = np.linalg.qr(rs.randn(B, 3, 3))
    init_sRT[:, :3, :3] = torch.from_numpy(q).float()
    init_sRT[:, :3, 3] = data_cpu["pts_center"] + torch.from_numpy(rs.randn(B, 3).astype(np.float32)) * 0.02
    for precision in ("fp32", "bf16x3"):
        pipe = PosePipeline(sd, esd, sampler="ode", sampling_steps=None, precision=precision)
        data = synth.batch_from_clouds(clouds, device="cuda")
        torch.manual_seed(5)
        out = pipe.track_step(data, init_sRT.cuda(), repeat_num=K, T0=T0)
        # oracle replay: same initial pose, same prior draw
        init_pose = init_sRT[:, :3, [0, 1, 3]].permute(0, 2, 1).reshape(B, -1).clone()
        init_pose[:, -3:] -= data_cpu["pts_center"]
        torch.manual_seed(5)
        prior = ve_prior((B * K, 9), T=T0)
        x0 = init_pose.unsqueeze(1).repeat(1, K, 1).view(B * K, -1) + prior
        feat = O.encode(sd, data_cpu["pts"])
        rep = feat.unsqueeze(1).repeat(1, K, 1).view(B * K, -1)
        cen = data_cpu["pts_center"].unsqueeze(1).repeat(1, K, 1).view(B * K, -1)
        ref = O.ode_sampler(sd, rep, cen, x0, T0=T0).reshape(B, K, 9)
        np.testing.assert_allclose(out["pred_pose"].cpu().numpy(), ref.numpy(), rtol=2e-4, atol=1e-3)
        en = O.get_energy(esd, data_cpu, ref.float())
        _, _, rt_ref = O.rank_and_pool(ref.float(), en)
        np.testing.assert_allclose(out["pooled_RT"].cpu().numpy(), np.asarray(rt_ref), rtol=0, atol=2e-3)
        assert out["pooled_RT"].shape == (B, 4, 4)


def test_run_stream_equals_run(ops):
    """PosePipeline.run_stream (next batch's encoder beside the current sampler) must give exactly what run() gives."""
    from genpose_b200.pipeline import PosePipeline
    sd = synth.make_state_dict(3, kappa=synth.stable_kappa(30))
    pipe = PosePipeline(sd, None, sampler="pc", sampling_steps=30)
    batches = [synth.batch_from_clouds(synth.make_clouds(4, 60 + i), device="cuda") for i in range(3)]
    torch.manual_seed(1)
    seq = [pipe.run(dict(b), repeat_num=50)["pred_pose"].clone() for b in batches]
    torch.manual_seed(1)
    stream = [p.clone() for p in pipe.run_stream([dict(b) for b in batches], repeat_num=50)]
    torch.cuda.synchronize()
    assert len(stream) == 3
    for a, b in zip(seq, stream):
        assert torch.equal(a, b)
