"""GPU: the tensor-core samplers beyond round 1's single configuration (one 4-CTA team per tile, three products):
  (1) the two-product arithmetic 'f16x2' (fp16 hi/lo activations x ONE fp16 weight image): the instruction itself against a
      device-independent product, then both samplers against the oracle / the fp32 FFMA kernel / the three-product kernel;
  (2) every tile-team size (4, 2, 1 CTAs per 128-row tile) in both arithmetics: same answer as the FFMA kernel,
      bitwise reproducible;
  (3) the reference's own evaluation batch (scripts/eval_single.sh:7: --batch_size 256, i.e. 12,800 rows at K = 50), which round 1
      could not hold on the tensor cores: PC and ODE against the oracle with explicit noise."""
import numpy as np
import pytest
import torch

from genpose_b200 import lib, synth, weights
from oracle import genpose_oracle as O
from oracle import tc_emulation as E

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("K,N,a_tmem", [(16, 128, 0), (64, 128, 1), (128, 256, 1), (256, 64, 1)])
def test_f16_two_product_mma(K, N, a_tmem):
    g = torch.Generator().manual_seed(K + N)
    A = torch.randn(128, K, generator=g).abs() * 3.0            # post-ReLU activations are non-negative (relu_split_f16x2)
    B = torch.randn(N, K, generator=g)
    b16 = B.to(torch.float16)
    img = weights.umma_image(b16.view(torch.int16).view(N, K).view(torch.bfloat16)).cuda()
    D = torch.zeros(128, N, device="cuda")
    Ad = A.cuda().contiguous()
    lib.check(lib.load().gpb_selftest_umma(Ad.data_ptr(), img.data_ptr(), img.data_ptr(), D.data_ptr(), K, N, 0, 4, 2, a_tmem, 1, 0,
                                           torch.cuda.current_stream().cuda_stream), "selftest_umma")
    torch.cuda.synchronize()
    ahi, alo = E.split_act_f16(A)
    ref = (ahi.double() + alo.double()) @ b16.double().t()
    scale = float((A.abs().double() @ B.abs().double().t()).max())
    assert float((D.cpu().double() - ref).abs().max()) <= 1e-6 * scale


def _pc_case(B, K, T, kappa=None):
    from genpose_b200 import ops
    seed = 50 + B
    sd = synth.make_state_dict(seed, kappa=synth.stable_kappa(T) if kappa is None else kappa)
    clouds = synth.make_clouds(B, seed)
    data = synth.batch_from_clouds(clouds)
    eng = ops.Engine(sd)
    feat = eng.encode(torch.from_numpy(clouds).cuda())
    x0, sn = synth.make_prior_noise(B * K, seed), synth.make_step_noise(T, B * K, seed)
    return sd, data, eng, feat, eng.object_bias(feat), data["pts_center"].cuda(), x0, sn


@pytest.mark.parametrize("B,K,T", [(2, 50, 30), (3, 64, 100), (5, 50, 500), (64, 50, 20)])
def test_two_product_pc_sampler(B, K, T):
    sd, data, eng, feat, ob, cen, x0, sn = _pc_case(B, K, T)
    args = (ob, cen, torch.from_numpy(x0).cuda(), K, T)
    noise = torch.from_numpy(sn).cuda()
    p2, proc = eng.sample_pc(*args, step_noise=noise, precision="f16x2", return_process=True)
    p3 = eng.sample_pc(*args, step_noise=noise, precision="bf16x3")
    p32 = eng.sample_pc(*args, step_noise=noise, precision="fp32")
    again = eng.sample_pc(*args, step_noise=noise, precision="f16x2")
    torch.cuda.synchronize()
    assert torch.isfinite(p2).all() and torch.isfinite(proc).all()
    assert torch.equal(p2, again)
    tol = 1e-3 + 5e-5 * p32.abs()
    print(f"f16x2 vs fp32 kernel: {float(((p2 - p32).abs() / tol).max()):.3f} of the bound; bf16x3 vs fp32: {float(((p3 - p32).abs() / tol).max()):.3f}")
    assert bool(((p2 - p32).abs() <= tol).all())
    if B * K <= 400:
        ref, _ = O.pred_func_pc(sd, data, K, T, torch.from_numpy(x0), torch.from_numpy(sn), pts_feat=feat.cpu())
        np.testing.assert_allclose(p2.cpu().numpy().reshape(B, K, 9), ref.numpy(), rtol=5e-5, atol=1e-3)


def _ode_case(B, K, T0, seed):
    from genpose_b200 import ops
    sd = synth.make_state_dict(seed, kappa=0.3)
    clouds = synth.make_clouds(B, seed)
    x0 = torch.from_numpy(synth.make_prior_noise(B * K, seed, sigma=float(O.sigma_of_t(torch.tensor(T0))))).cuda()
    data = synth.batch_from_clouds(clouds)
    eng = ops.Engine(sd)
    feat = eng.encode(torch.from_numpy(clouds).cuda())
    return sd, data, eng, feat, eng.object_bias(feat), data["pts_center"].cuda(), x0


@pytest.mark.parametrize("B,K,T0", [(3, 50, 0.55), (2, 64, 0.15), (64, 50, 0.55)])
def test_two_product_ode_sampler(B, K, T0):
    sd, data, eng, feat, ob, cen, x0 = _ode_case(B, K, T0, 70 + B)
    p2, s2 = eng.sample_ode(ob, cen, x0, K, T0=T0, precision="f16x2")
    p32, s32 = eng.sample_ode(ob, cen, x0, K, T0=T0, precision="fp32")
    again, _ = eng.sample_ode(ob, cen, x0, K, T0=T0, precision="f16x2")
    torch.cuda.synchronize()
    s2, s32 = s2.cpu().numpy(), s32.cpu().numpy()
    assert s2[3] == 0 and torch.isfinite(p2).all() and torch.equal(p2, again)
    assert abs(int(s2[0]) - int(s32[0])) <= 12, (s2, s32)
    np.testing.assert_allclose(p2.cpu().numpy(), p32.cpu().numpy(), rtol=2e-4, atol=1e-3)


@pytest.mark.parametrize("precision", ["bf16x3", "f16x2"])
@pytest.mark.parametrize("team", [1, 2, 4])
def test_every_team_size_pc(team, precision):
    """5 tiles (one of them partial, tiles spanning 2-3 objects) through a team of `team` CTAs per tile."""
    B, K, T = 10, 60, 40
    sd, data, eng, feat, ob, cen, x0, sn = _pc_case(B, K, T)
    args = (ob, cen, torch.from_numpy(x0).cuda(), K, T)
    noise = torch.from_numpy(sn).cuda()
    p, proc = eng.sample_pc(*args, step_noise=noise, precision=precision, team=team, return_process=True)
    again = eng.sample_pc(*args, step_noise=noise, precision=precision, team=team)
    p32, proc32 = eng.sample_pc(*args, step_noise=noise, precision="fp32", return_process=True)
    torch.cuda.synchronize()
    assert torch.isfinite(p).all() and torch.equal(p, again)
    np.testing.assert_allclose(p.cpu().numpy(), p32.cpu().numpy(), rtol=5e-5, atol=1e-3)
    np.testing.assert_allclose(proc.cpu().numpy(), proc32.cpu().numpy(), rtol=5e-5, atol=1e-3)
    ref, _ = O.pred_func_pc(sd, data, K, T, torch.from_numpy(x0), torch.from_numpy(sn), pts_feat=feat.cpu())
    np.testing.assert_allclose(p.cpu().numpy().reshape(B, K, 9), ref.numpy(), rtol=5e-5, atol=1e-3)


@pytest.mark.parametrize("precision", ["bf16x3", "f16x2"])
@pytest.mark.parametrize("team", [1, 2, 4])
def test_every_team_size_ode(team, precision):
    B, K, T0 = 10, 60, 0.55
    sd, data, eng, feat, ob, cen, x0 = _ode_case(B, K, T0, 81)
    p, s = eng.sample_ode(ob, cen, x0, K, T0=T0, precision=precision, team=team)
    again, _ = eng.sample_ode(ob, cen, x0, K, T0=T0, precision=precision, team=team)
    p32, s32 = eng.sample_ode(ob, cen, x0, K, T0=T0, precision="fp32")
    torch.cuda.synchronize()
    s, s32 = s.cpu().numpy(), s32.cpu().numpy()
    assert s[3] == 0 and torch.isfinite(p).all() and torch.equal(p, again)
    assert abs(int(s[0]) - int(s32[0])) <= 12, (s, s32)
    np.testing.assert_allclose(p.cpu().numpy(), p32.cpu().numpy(), rtol=2e-4, atol=1e-3)


def test_team_knob_is_validated_and_restored():
    L = lib.load()
    assert L.gpb_set_tc_team(3) != 0
    assert L.gpb_set_tc_team(2) == 0 and L.gpb_sampler_tc_max_rows(50) >= 48 * 128
    assert L.gpb_set_tc_team(4) == 0 and 3200 <= L.gpb_sampler_tc_max_rows(50) < 64 * 128
    assert L.gpb_set_tc_team(0) == 0 and L.gpb_sampler_tc_max_rows(50) >= 12800


@pytest.mark.parametrize("precision", ["bf16x3", "f16x2"])
def test_reference_eval_batch_256_objects_pc(precision):
    """scripts/eval_single.sh:7 (--batch_size 256) x K = 50 = 12,800 rows = 100 tiles: one tile per SM (team of 1).  The whole
    batch is ONE launch (the batch-mean gradient norm, samplers.py:130, couples every row); against the oracle with explicit noise."""
    B, K, T = 256, 50, 24
    sd, data, eng, feat, ob, cen, x0, sn = _pc_case(B, K, T)
    assert eng.tc_supported(B * K, K)
    args = (ob, cen, torch.from_numpy(x0).cuda(), K, T)
    noise = torch.from_numpy(sn).cuda()
    p = eng.sample_pc(*args, step_noise=noise, precision=precision)
    again = eng.sample_pc(*args, step_noise=noise, precision=precision)
    torch.cuda.synchronize()
    assert torch.isfinite(p).all() and torch.equal(p, again)
    ref, _ = O.pred_func_pc(sd, data, K, T, torch.from_numpy(x0), torch.from_numpy(sn), pts_feat=feat.cpu())
    np.testing.assert_allclose(p.cpu().numpy().reshape(B, K, 9), ref.numpy(), rtol=5e-5, atol=1e-3)


def test_reference_eval_batch_256_objects_ode():
    """The shipped recipe at its shipped batch (eval_single.sh:5-7: ode, T0 = 0.55, 256 objects x 50): tensor cores, team of 1,
    against the oracle's SciPy-controller port — same accept / reject sequence, poses within the ODE bound."""
    B, K, T0 = 256, 50, 0.55
    sd, data, eng, feat, ob, cen, x0 = _ode_case(B, K, T0, 91)
    assert eng.tc_supported(B * K, K)
    p, s = eng.sample_ode(ob, cen, x0, K, T0=T0, precision="auto")
    torch.cuda.synchronize()
    s = s.cpu().numpy()
    rep = feat.cpu().unsqueeze(1).repeat(1, K, 1).view(B * K, -1)
    cen_rep = data["pts_center"].unsqueeze(1).repeat(1, K, 1).view(B * K, -1)
    ref, st = O.ode_sampler(sd, rep, cen_rep, x0.cpu(), T0=T0, return_stats=True)
    assert s[3] == 0 and int(s[0]) == st["nfev"], (s, st)
    np.testing.assert_allclose(p.cpu().numpy(), ref.numpy(), rtol=2e-4, atol=1e-3)
