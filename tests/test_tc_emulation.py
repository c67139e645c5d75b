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


def test_split_product_is_exact_to_2_pow_minus_16():
    g = torch.Generator().manual_seed(0)
    a, w = torch.randn(64, 256, generator=g), torch.randn(96, 256, generator=g)
    exact = a.double() @ w.double().t()
    scale = float((a.abs().double() @ w.abs().double().t()).max())
    assert float((E.mm_split(a, w, 3).double() - exact).abs().max()) <= 2.0 ** -15 * scale
    assert float((E.mm_split(a, w, 1).double() - exact).abs().max()) > 2.0 ** -12 * scale * 0.1
