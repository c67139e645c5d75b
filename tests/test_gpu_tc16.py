"""GPU, EXPERIMENTAL (skipped unless GPB_EXPERIMENTAL=1): the two-product tensor-core samplers (precision='bf16x2': bf16 hi/lo
activations x ONE fp16 weight image, include/genpose_b200.h).  Their arithmetic is pinned on the CPU (tests/test_tc_emulation.py,
oracle/tc_emulation.py); these tests are the hardware side, to be run before the mode may become what 'auto' selects:
(1) the mixed-format instruction itself (A = bf16, B = fp16 in one kind::f16 tcgen05.mma) against a device-independent product,
(2) PC and ODE samplers against the oracle, the fp32 FFMA kernel and the shipped three-product kernel, (3) run-to-run identity."""
import os

import numpy as np
import pytest
import torch

from genpose_b200 import lib, synth, weights
from oracle import genpose_oracle as O

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("GPB_EXPERIMENTAL") != "1",
                                 reason="two-product tensor-core samplers are not validated on hardware yet (GPB_EXPERIMENTAL=1 runs them)")]


@pytest.mark.parametrize("K,N,a_tmem", [(16, 128, 0), (64, 128, 1), (128, 256, 1), (256, 64, 1)])
def test_mixed_format_mma(K, N, a_tmem):
    g = torch.Generator().manual_seed(K + N)
    A = torch.randn(128, K, generator=g)
    B = torch.randn(N, K, generator=g)
    b16 = B.to(torch.float16)
    img = weights.umma_image(b16.view(torch.int16).view(N, K).view(torch.bfloat16)).cuda()
    D = torch.zeros(128, N, device="cuda")
    Ad = A.cuda().contiguous()
    lib.check(lib.load().gpb_selftest_umma(Ad.data_ptr(), img.data_ptr(), img.data_ptr(), D.data_ptr(), K, N, 0, 4, 2, a_tmem, 1, 0,
                                           torch.cuda.current_stream().cuda_stream), "selftest_umma")
    torch.cuda.synchronize()
    ahi, alo = weights.split_bf16(A)
    ref = (ahi.double() + alo.double()) @ b16.double().t()
    scale = float((A.abs().double() @ B.abs().double().t()).max())
    assert float((D.cpu().double() - ref).abs().max()) <= 1e-6 * scale


def _pc_case(B, K, T):
    from genpose_b200 import ops
    seed = 50 + B
    sd = synth.make_state_dict(seed, kappa=synth.stable_kappa(T))
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
    p2, proc = eng.sample_pc(*args, step_noise=noise, precision="bf16x2", return_process=True)
    p3 = eng.sample_pc(*args, step_noise=noise, precision="bf16x3")
    p32 = eng.sample_pc(*args, step_noise=noise, precision="fp32")
    again = eng.sample_pc(*args, step_noise=noise, precision="bf16x2")
    torch.cuda.synchronize()
    assert torch.isfinite(p2).all() and torch.isfinite(proc).all()
    assert torch.equal(p2, again)
    tol = 1e-3 + 5e-5 * p32.abs()
    print(f"bf16x2 vs fp32 kernel: {float(((p2 - p32).abs() / tol).max()):.3f} of the bound; bf16x3 vs fp32: {float(((p3 - p32).abs() / tol).max()):.3f}")
    assert bool(((p2 - p32).abs() <= tol).all())
    if B * K <= 400:
        ref, _ = O.pred_func_pc(sd, data, K, T, torch.from_numpy(x0), torch.from_numpy(sn), pts_feat=feat.cpu())
        np.testing.assert_allclose(p2.cpu().numpy().reshape(B, K, 9), ref.numpy(), rtol=5e-5, atol=1e-3)


@pytest.mark.parametrize("B,K,T0", [(3, 50, 0.55), (2, 64, 0.15), (64, 50, 0.55)])
def test_two_product_ode_sampler(B, K, T0):
    from genpose_b200 import ops
    seed = 70 + B
    sd = synth.make_state_dict(seed, kappa=0.3)
    clouds = synth.make_clouds(B, seed)
    x0 = torch.from_numpy(synth.make_prior_noise(B * K, seed, sigma=float(O.sigma_of_t(torch.tensor(T0))))).cuda()
    data = synth.batch_from_clouds(clouds)
    eng = ops.Engine(sd)
    ob = eng.object_bias(eng.encode(torch.from_numpy(clouds).cuda()))
    cen = data["pts_center"].cuda()
    p2, s2 = eng.sample_ode(ob, cen, x0, K, T0=T0, precision="bf16x2")
    p32, s32 = eng.sample_ode(ob, cen, x0, K, T0=T0, precision="fp32")
    again, _ = eng.sample_ode(ob, cen, x0, K, T0=T0, precision="bf16x2")
    torch.cuda.synchronize()
    s2, s32 = s2.cpu().numpy(), s32.cpu().numpy()
    assert s2[3] == 0 and torch.isfinite(p2).all() and torch.equal(p2, again)
    assert abs(int(s2[0]) - int(s32[0])) <= 12, (s2, s32)
    np.testing.assert_allclose(p2.cpu().numpy(), p32.cpu().numpy(), rtol=2e-4, atol=1e-3)
