"""CPU: the arithmetic of the tensor-core samplers (bf16 hi/lo split, fp32 accumulation), emulated on the oracle port, keeps
the sampled poses within the parity bound — and plain bf16 does not.  No GPU, no product code path."""
import numpy as np
import pytest
import torch

from genpose_b200 import synth
from oracle import genpose_oracle as O
from oracle import tc_emulation as E


def _case(B, K, T, seed):
    sd = synth.make_state_dict(seed, kappa=synth.stable_kappa(T))
    data = synth.batch_from_clouds(synth.make_clouds(B, seed))
    x0 = torch.from_numpy(synth.make_prior_noise(B * K, seed))
    sn = torch.from_numpy(synth.make_step_noise(T, B * K, seed))
    feat = O.encode(sd, data["pts"])
    return sd, data, x0, sn, feat


@pytest.mark.parametrize("B,K,T", [(2, 50, 30), (1, 16, 100)])
def test_bf16x3_split_keeps_the_parity_bound(B, K, T):
    sd, data, x0, sn, feat = _case(B, K, T, 50 + B)
    ref, _ = O.pred_func_pc(sd, data, K, T, x0, sn, pts_feat=feat)
    with E.emulated_score(terms=3):
        tc, _ = O.pred_func_pc(sd, data, K, T, x0, sn, pts_feat=feat)
    tol = 1e-3 + 5e-5 * ref.abs()                       # the bound of tests/test_gpu_tc.py
    frac = float(((tc - ref).abs() / tol).max())
    assert frac < 0.5, frac                             # measured 0.09: an order of magnitude of head room
    with E.emulated_score(terms=1):
        bf, _ = O.pred_func_pc(sd, data, K, T, x0, sn, pts_feat=feat)
    assert float(((bf - ref).abs() / tol).max()) > 3 * frac     # single bf16 is several times worse


@pytest.mark.parametrize("B,K,T", [(2, 50, 30), (3, 50, 20), (1, 16, 100)])
def test_two_product_mode_keeps_the_parity_bound(B, K, T):
    """The 'f16x2' samplers (fp16 hi/lo activations x ONE fp16 weight): what the dropped Ahi.Blo product carried is the weights'
    bits beyond the first image, and an fp16 weight keeps three more of them than a bf16 one.  Shortest chains are the worst case
    (measured ~0.3-0.5 of the bound at T = 20..30, 0.05 at T = 100, 0.03 at T = 500); with bf16 weights the same two products
    are 3.7x over."""
    sd, data, x0, sn, feat = _case(B, K, T, 50 + B)
    ref, _ = O.pred_func_pc(sd, data, K, T, x0, sn, pts_feat=feat)
    with E.emulated_score(terms="x2"):
        tc, _ = O.pred_func_pc(sd, data, K, T, x0, sn, pts_feat=feat)
    tol = 1e-3 + 5e-5 * ref.abs()
    assert float(((tc - ref).abs() / tol).max()) < 0.7


def test_two_product_mode_ode():
    B, K, T0, seed = 3, 50, 0.55, 73
    sd = synth.make_state_dict(seed, kappa=0.3)
    data = synth.batch_from_clouds(synth.make_clouds(B, seed))
    x0 = torch.from_numpy(synth.make_prior_noise(B * K, seed, sigma=float(O.sigma_of_t(torch.tensor(T0)))))
    feat = O.encode(sd, data["pts"])
    rep = feat.unsqueeze(1).repeat(1, K, 1).view(B * K, -1)
    cen = data["pts_center"].unsqueeze(1).repeat(1, K, 1).view(B * K, -1)
    ref, st = O.ode_sampler(sd, rep, cen, x0, T0=T0, return_stats=True)
    with E.emulated_score(terms="x2"):
        out, st2 = O.ode_sampler(sd, rep, cen, x0, T0=T0, return_stats=True)
    assert st2["nfev"] == st["nfev"]
    assert float(((out - ref).abs() / (1e-3 + 2e-4 * ref.abs())).max()) < 0.1


def test_split_product_is_exact_to_2_pow_minus_16():
    g = torch.Generator().manual_seed(0)
    a, w = torch.randn(64, 256, generator=g), torch.randn(96, 256, generator=g)
    exact = a.double() @ w.double().t()
    scale = float((a.abs().double() @ w.abs().double().t()).max())
    assert float((E.mm_split(a, w, 3).double() - exact).abs().max()) <= 2.0 ** -15 * scale
    assert float((E.mm_split(a, w, 1).double() - exact).abs().max()) > 2.0 ** -12 * scale * 0.1


def test_f16_activation_split_matches_its_definition():
    """split_act_f16 = truncation to fp16 + rounded residual: hi is an fp16 value <= a, hi + lo carries >= 20 mantissa bits."""
    g = torch.Generator().manual_seed(1)
    a = torch.rand(4096, generator=g) * torch.tensor([1e-6, 1e-3, 1.0, 300.0]).repeat(1024)
    hi, lo = E.split_act_f16(a)
    assert torch.equal(hi, hi.to(torch.float16).float()) and bool((hi <= a).all()) and bool((lo >= 0).all())
    big = a >= 1e-3
    assert float(((hi + lo - a).abs()[big] / a[big]).max()) <= 2.0 ** -20
    assert float((hi + lo - a).abs().max()) <= 2.0 ** -20 * 300 + 2.0 ** -25
    assert float(E.split_act_f16(torch.tensor([1e6, -3.0]))[0][0]) == E.FP16_MAX and float(E.split_act_f16(torch.tensor([-3.0]))[0][0]) == 0.0


def test_undamped_golden_separates_the_two_arithmetics():
    """pc_B3_K50_T12_kappa0 (a vector made by the UNMODIFIED reference, no stabilising field, T = 12): sampler fed the reference's own
    features.  Three products: 0.05 of the parity bound.  Two products (fp16 weights): 1.24x the bound (3.2e-3 on ~30 m translations)
    — the number the B200 kernel reproduces to the digit (tests/test_gpu_tc_teams.py::test_undamped_dynamics_golden), which is why
    `precision='bf16x3'` stays available as the high-fidelity choice."""
    from tests import _cases
    case, g, inp = _cases.load("pc_B3_K50_T12_kappa0")
    data = synth.batch_from_clouds(inp["clouds"])
    feat = torch.from_numpy(g["ref_pts_feat"])
    ref = g["ref_pred_pose"]
    frac = {}
    for terms in (3, "x2"):
        with E.emulated_score(terms=terms):
            tc, _ = O.pred_func_pc(inp["sd"], data, case["K"], case["T"], torch.from_numpy(inp["x0"]), torch.from_numpy(inp["step_noise"]),
                                   pts_feat=feat)
        frac[terms] = float((np.abs(tc.numpy() - ref) / (1e-3 + 5e-5 * np.abs(ref))).max())
    assert frac[3] < 0.2, frac
    assert 0.5 < frac["x2"] < 2.0, frac
