"""ORACLE — test infrastructure, NOT product code.

Portable CPU restatement (torch fp32 on the host + oracle/pointnet2_cpu.c for the index ops) of the
reference's per-object inference hot path, SURVEY.md §8(a) rows a1-a13.  It exists because the
reference itself (/root/reference, Python + a CUDA-only extension) cannot travel to the GPU box.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may
import this module; the product (genpose_b200/) never does and fails loudly without its CUDA library.

PINNING: tests/test_oracle_vs_reference.py runs this file against the UNMODIFIED reference
(imported by oracle/ref_loader.py) on the same seeded inputs whenever /root/reference exists, and
tests/test_oracle_golden.py checks it everywhere against tests/golden/*.npz, which
oracle/make_golden.py produced by executing the reference.  The reference ships no tests or golden
vectors of its own, so those generated vectors are the pin.

Every function cites the reference file:line it follows (paths relative to /root/reference).
`dtype=torch.float64` re-runs the same algorithm in double precision; tests use it only to measure
how well-conditioned a configuration is (fp32-vs-fp64 drift), never as the pass/fail oracle.
"""
import math
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

from oracle import pointnet2_cpu

# ---- static configuration (restated from the reference; see the cited lines) -------------------
# networks/pts_encoder/pointnet2.py:57-66 ClsMSG_CFG_Light, use_xyz=True (pointnet2_modules.py:89-90)
NPOINTS = [512, 256, 128, None]
RADIUS = [[0.02, 0.04], [0.04, 0.08], [0.08, 0.16], [None, None]]
NSAMPLE = [[16, 32], [16, 32], [16, 32], [None, None]]
MLPS = [[[3, 16, 16, 32], [3, 32, 32, 64]],
        [[99, 64, 64, 128], [99, 64, 96, 128]],
        [[259, 128, 196, 256], [259, 128, 196, 256]],
        [[515, 256, 256, 512], [515, 256, 384, 512]]]
SIGMA_MIN, SIGMA_MAX, EPS = 0.01, 50.0, 1e-5       # networks/gf_algorithms/sde.py:90-97
HEADS = ("rot_x", "rot_y", "trans")
BN_EPS = 1e-5                                       # torch BatchNorm2d default (pytorch_utils.py:117)
# sde.py:23: torch.sqrt(torch.tensor(2*(np.log(sigma_max)-np.log(sigma_min)))) — a 0-dim FLOAT64 tensor
# (numpy float64 scalar in), which multiplies fp32 sigma as a scalar (result stays fp32).
G_COEF = torch.sqrt(torch.tensor(2 * (np.log(SIGMA_MAX) - np.log(SIGMA_MIN))))


# ------------------------------------------------------------------------------------------------
# a1-a4: index ops (exact), via the C restatement
# ------------------------------------------------------------------------------------------------
def furthest_point_sample(xyz: torch.Tensor, npoint: int) -> torch.Tensor:
    """pointnet2_utils.py:13-30 FurthestPointSampling.forward (temp=1e10, idx int32)."""
    xyz = xyz.float().contiguous()
    B, N, _ = xyz.shape
    idx = torch.zeros(B, npoint, dtype=torch.int32)
    temp = torch.full((B, N), 1e10, dtype=torch.float32)
    pointnet2_cpu.furthest_point_sampling_wrapper(B, N, npoint, xyz, temp, idx)
    return idx


def ball_query(radius: float, nsample: int, xyz: torch.Tensor, new_xyz: torch.Tensor) -> torch.Tensor:
    """pointnet2_utils.py:204-222 BallQuery.forward (idx pre-zeroed :219)."""
    xyz = xyz.float().contiguous()
    new_xyz = new_xyz.float().contiguous()
    B, N, _ = xyz.shape
    npoint = new_xyz.shape[1]
    idx = torch.zeros(B, npoint, nsample, dtype=torch.int32)
    pointnet2_cpu.ball_query_wrapper(B, N, npoint, radius, nsample, new_xyz, xyz, idx)
    return idx


def _gather_rows(x: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """x [B,N,C], idx [B,...] int -> [B,...,C] (GatherOperation / GroupingOperation as pure indexing)."""
    B = x.shape[0]
    flat = idx.reshape(B, -1).long()
    out = torch.gather(x, 1, flat.unsqueeze(-1).expand(-1, -1, x.shape[-1]))
    return out.reshape(*idx.shape, x.shape[-1])


# ------------------------------------------------------------------------------------------------
# a5-a7: encoder
# ------------------------------------------------------------------------------------------------
def _shared_mlp(sd: Dict[str, torch.Tensor], prefix: str, x: torch.Tensor, n_layers: int, dtype) -> torch.Tensor:
    """pytorch_utils.py:5-32 SharedMLP = n x [Conv2d 1x1 (no bias, :57) -> BatchNorm2d(eval) -> ReLU].
    x: [..., Cin] channel-last; evaluated exactly as conv -> bn -> relu (no folding) in `dtype`."""
    for j in range(n_layers):
        p = f"{prefix}.layer{j}"
        w = sd[f"{p}.conv.weight"].to(dtype)[:, :, 0, 0]                       # [Cout, Cin]
        x = x @ w.t()
        mean = sd[f"{p}.bn.bn.running_mean"].to(dtype)
        var = sd[f"{p}.bn.bn.running_var"].to(dtype)
        gamma = sd[f"{p}.bn.bn.weight"].to(dtype)
        beta = sd[f"{p}.bn.bn.bias"].to(dtype)
        x = (x - mean) / torch.sqrt(var + BN_EPS) * gamma + beta               # F.batch_norm eval form
        x = torch.relu(x)
    return x


def encoder_levels(sd: Dict[str, torch.Tensor], pts: torch.Tensor, dtype=torch.float32,
                   prefix: str = "pts_encoder") -> Dict[str, object]:
    """pointnet2.py:203-211 Pointnet2ClsMSG.forward + pointnet2_modules.py:19-56 per level.
    Returns every intermediate (indices are exact int32) so kernels can be checked level by level.
    The input is data['pts'] — raw camera-frame xyz, NOT zero-centred (posenet.py:79)."""
    xyz = pts[..., 0:3].float().contiguous()
    feats: Optional[torch.Tensor] = None                     # [B, N, C] channel-last
    trace = {"fps_idx": [], "new_xyz": [], "ball_idx": [], "feats": []}
    for l in range(4):
        npoint = NPOINTS[l]
        outs = []
        if npoint is not None:
            fps_idx = furthest_point_sample(xyz, npoint)                        # pointnet2_modules.py:33
            new_xyz = _gather_rows(xyz, fps_idx)                                # :31-35
            trace["fps_idx"].append(fps_idx)
            trace["new_xyz"].append(new_xyz)
            level_ball = []
            for s in range(2):
                idx = ball_query(RADIUS[l][s], NSAMPLE[l][s], xyz, new_xyz)     # pointnet2_utils.py:250
                level_ball.append(idx)
                g_xyz = _gather_rows(xyz, idx) - new_xyz.unsqueeze(2)           # :251-253
                if feats is not None:
                    g = torch.cat([g_xyz.to(dtype), _gather_rows(feats, idx)], dim=-1)   # xyz channels FIRST :258
                else:
                    g = g_xyz.to(dtype)
                h = _shared_mlp(sd, f"{prefix}.SA_modules.{l}.mlps.{s}", g, 3, dtype)
                outs.append(h.max(dim=2).values)                                # max_pool2d over nsample, modules.py:43
            trace["ball_idx"].append(level_ball)
            xyz_next = new_xyz
        else:
            # GroupAll: absolute xyz || feats over all remaining points (pointnet2_utils.py:281-289)
            g = torch.cat([xyz.to(dtype), feats], dim=-1).unsqueeze(1)          # [B,1,N,3+C]
            for s in range(2):
                h = _shared_mlp(sd, f"{prefix}.SA_modules.{l}.mlps.{s}", g, 3, dtype)
                outs.append(h.max(dim=2).values)                                # [B,1,C]
            xyz_next = None
        feats = torch.cat(outs, dim=-1)                                         # scale-0 channels first, modules.py:56
        trace["feats"].append(feats)
        xyz = xyz_next
    trace["pts_feat"] = feats.squeeze(1)                                        # [B,1024] pointnet2.py:211
    return trace


def encode(sd, pts, dtype=torch.float32) -> torch.Tensor:
    return encoder_levels(sd, pts, dtype)["pts_feat"]


# ------------------------------------------------------------------------------------------------
# a8: score network
# ------------------------------------------------------------------------------------------------
def sigma_of_t(t):
    """ve_marginal_prob std (sde.py:15-18): sigma_min * (sigma_max / sigma_min) ** t."""
    return SIGMA_MIN * (SIGMA_MAX / SIGMA_MIN) ** t


def _trunk(sd: Dict[str, torch.Tensor], pts_feat, pose, t, dtype, prefix="pose_score_net") -> torch.Tensor:
    """scorenet.py:195-216 / energynet.py:144-163: f_theta = cat of the three heads (before the /std)."""
    W = sd[f"{prefix}.t_encoder.0.W"].to(dtype)
    tt = t.to(dtype).squeeze(1)
    x_proj = tt[:, None] * W[None, :] * 2 * np.pi                               # scorenet.py:63
    emb = torch.cat([torch.sin(x_proj), torch.cos(x_proj)], dim=-1)            # :64 [sin | cos]
    t_feat = torch.relu(F.linear(emb, sd[f"{prefix}.t_encoder.1.weight"].to(dtype),
                                 sd[f"{prefix}.t_encoder.1.bias"].to(dtype)))
    h = torch.relu(F.linear(pose.to(dtype), sd[f"{prefix}.pose_encoder.0.weight"].to(dtype),
                            sd[f"{prefix}.pose_encoder.0.bias"].to(dtype)))
    pose_feat = torch.relu(F.linear(h, sd[f"{prefix}.pose_encoder.2.weight"].to(dtype),
                                    sd[f"{prefix}.pose_encoder.2.bias"].to(dtype)))
    total = torch.cat([pts_feat.to(dtype), t_feat, pose_feat], dim=-1)          # :204 [pts | t | pose]
    outs = []
    for name in HEADS:
        hk = torch.relu(F.linear(total, sd[f"{prefix}.fusion_tail_{name}.0.weight"].to(dtype),
                                 sd[f"{prefix}.fusion_tail_{name}.0.bias"].to(dtype)))
        outs.append(F.linear(hk, sd[f"{prefix}.fusion_tail_{name}.2.weight"].to(dtype),
                             sd[f"{prefix}.fusion_tail_{name}.2.bias"].to(dtype)))
    return torch.cat(outs, dim=-1)                                              # [rot_x | rot_y | trans] :217


def score(sd, pts_feat, pose, t, dtype=torch.float32) -> torch.Tensor:
    """PoseScoreNet.forward scorenet.py:178-222 (Rx_Ry_and_T): f / (std + 1e-7)."""
    f = _trunk(sd, pts_feat, pose, t, dtype)
    std = sigma_of_t(t.to(dtype))
    return f / (std + 1e-7)


# ------------------------------------------------------------------------------------------------
# rotation helpers (utils/misc.py + pytorch3d v0.7.2 published definitions)
# ------------------------------------------------------------------------------------------------
def rot6d_columns(r6: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """pytorch3d rotation_6d_to_matrix (F.normalize eps 1e-12) as used by get_rot_matrix
    (utils/misc.py:136, result transposed => b1,b2,b3 are the matrix COLUMNS)."""
    a1, a2 = r6[..., :3], r6[..., 3:6]
    b1 = F.normalize(a1, dim=-1)
    b2 = a2 - (b1 * a2).sum(-1, keepdim=True) * b1
    b2 = F.normalize(b2, dim=-1)
    b3 = torch.cross(b1, b2, dim=-1)
    return b1, b2, b3


def normalize_rotation(r6: torch.Tensor) -> torch.Tensor:
    """utils/misc.py:259-265 normalize_rotation('rot_matrix'): Gram-Schmidt of the two axes."""
    b1, b2, _ = rot6d_columns(r6)
    return torch.cat([b1, b2], dim=-1)


def get_rot_matrix(r6: torch.Tensor) -> torch.Tensor:
    b1, b2, b3 = rot6d_columns(r6)
    return torch.stack([b1, b2, b3], dim=-1)                                    # columns


# ------------------------------------------------------------------------------------------------
# a9: predictor-corrector sampler
# ------------------------------------------------------------------------------------------------
def pc_sampler(sd, pts_feat, pts_center, x0, step_noise, num_steps: int, snr: float = 0.16,
               dtype=torch.float32, return_process: bool = False):
    """cond_pc_sampler, networks/gf_algorithms/samplers.py:102-160, pose_mode 'rot_matrix'.
    pts_feat [R,1024], pts_center [R,3], x0 [R,9] (= prior sample, sde.py:26-28),
    step_noise [T,2,R,9] replaces the two torch.randn_like draws per step (:131, :149) in order.
    Returns mean_x (the reference's `res`) and optionally the recorded xs [R,T,9]."""
    R = x0.shape[0]
    x = x0.to(dtype).clone()
    time_steps = torch.linspace(1.0, EPS, num_steps).to(dtype) if dtype == torch.float32 else \
        torch.linspace(1.0, EPS, num_steps, dtype=torch.float32).to(dtype)      # linspace is fp32 in the reference (:118)
    time_steps = time_steps.to(x.device)                                        # (device-agnostic: bench.py's GPU-torch context baseline)
    step_size = time_steps[0] - time_steps[1]                                   # :119
    noise_norm = np.sqrt(9)                                                     # :120
    pf = pts_feat.to(dtype)
    poses = []
    mean_x = None
    for i in range(num_steps):
        ts = time_steps[i]
        bt = torch.ones(R, 1, dtype=dtype, device=x.device) * ts                # :125
        grad = score(sd, pf, x, bt, dtype)                                      # :129
        grad_norm = torch.norm(grad.reshape(R, -1), dim=-1).mean()              # :130  ONE scalar per batch
        ls = 2 * (snr * noise_norm / grad_norm) ** 2                            # :131
        x = x + ls * grad + torch.sqrt(2 * ls) * step_noise[i, 0].to(dtype)     # :132
        x[:, :3] /= torch.norm(x[:, :3], dim=-1, keepdim=True)                  # :142
        x[:, 3:6] /= torch.norm(x[:, 3:6], dim=-1, keepdim=True)                # :143
        sigma = sigma_of_t(bt)                                                  # ve_sde sde.py:20-24
        diffusion = sigma * G_COEF.to(dtype) if dtype != torch.float32 else sigma * G_COEF
        drift = 0 - diffusion ** 2 * grad                                       # :147 (same grad, sign as written)
        mean_x = x + drift * step_size                                          # :148
        x = mean_x + diffusion * torch.sqrt(step_size) * step_noise[i, 1].to(dtype)   # :149
        x[:, :-3] = normalize_rotation(x[:, :-3])                               # :152
        if return_process:
            poses.append(x.clone().unsqueeze(0))
    mean_x = mean_x.clone()
    mean_x[:, -3:] += pts_center.to(dtype)                                      # :157
    mean_x[:, :-3] = normalize_rotation(mean_x[:, :-3])                         # :158
    if return_process:
        xs = torch.cat(poses, dim=0)
        xs[:, :, -3:] += pts_center.to(dtype).unsqueeze(0)                      # :156
        return mean_x, xs.permute(1, 0, 2)
    return mean_x


# ------------------------------------------------------------------------------------------------
# a10: probability-flow ODE sampler (SciPy RK45 restated: scipy/integrate/_ivp/{rk,common}.py)
# ------------------------------------------------------------------------------------------------
_RK45_C = np.array([0, 1 / 5, 3 / 10, 4 / 5, 8 / 9, 1])
_RK45_A = np.array([
    [0, 0, 0, 0, 0],
    [1 / 5, 0, 0, 0, 0],
    [3 / 40, 9 / 40, 0, 0, 0],
    [44 / 45, -56 / 15, 32 / 9, 0, 0],
    [19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729, 0],
    [9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656]])
_RK45_B = np.array([35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84])
_RK45_E = np.array([-71 / 57600, 0, 71 / 16695, -71 / 1920, 17253 / 339200, -22 / 525, 1 / 40])


def _rms(v: np.ndarray) -> float:
    return float(np.linalg.norm(v) / v.size ** 0.5)


def rk45_integrate(fun, t0: float, t_bound: float, y0: np.ndarray, rtol: float, atol: float):
    """scipy.integrate.solve_ivp(method='RK45') without dense output / events, as called at
    samplers.py:205.  Follows scipy 1.x `RK45`/`RungeKutta._step_impl`, `select_initial_step`
    (SAFETY 0.9, MIN_FACTOR 0.2, MAX_FACTOR 10, error exponent -1/5, RMS norm over ALL components).
    Returns (y_final, nfev, n_accepted, n_rejected)."""
    SAFETY, MIN_FACTOR, MAX_FACTOR = 0.9, 0.2, 10.0
    err_exp = -1.0 / 5.0
    y = np.array(y0, dtype=np.float64)
    t = float(t0)
    direction = np.sign(t_bound - t0) if t_bound != t0 else 1.0
    nfev = 0

    def f_eval(tt, yy):
        nonlocal nfev
        nfev += 1
        return np.asarray(fun(tt, yy), dtype=np.float64)

    f = f_eval(t, y)
    # select_initial_step (scipy/integrate/_ivp/common.py)
    interval_length = abs(t_bound - t0)
    if y.size == 0 or interval_length == 0.0:
        h_abs = 0.0 if interval_length == 0.0 else np.inf
    else:
        scale = atol + np.abs(y) * rtol
        d0 = _rms(y / scale)
        d1 = _rms(f / scale)
        h0 = 1e-6 if (d0 < 1e-5 or d1 < 1e-5) else 0.01 * d0 / d1
        h0 = min(h0, interval_length)
        y1 = y + h0 * direction * f
        f1 = f_eval(t + h0 * direction, y1)
        d2 = _rms((f1 - f) / scale) / h0
        if d1 <= 1e-15 and d2 <= 1e-15:
            h1 = max(1e-6, h0 * 1e-3)
        else:
            h1 = (0.01 / max(d1, d2)) ** (1.0 / 5.0)
        h_abs = min(100 * h0, h1, interval_length)
    n_acc = n_rej = 0
    K = np.empty((7, y.size), dtype=np.float64)
    while direction * (t - t_bound) < 0:
        min_step = 10 * np.abs(np.nextafter(t, direction * np.inf) - t)
        if h_abs < min_step:
            h_abs = min_step
        step_accepted = False
        step_rejected = False
        while not step_accepted:
            if h_abs < min_step:
                raise RuntimeError("rk45: step size underflow")
            h = h_abs * direction
            t_new = t + h
            if direction * (t_new - t_bound) > 0:
                t_new = t_bound
            h = t_new - t
            h_abs = np.abs(h)
            K[0] = f
            for s in range(1, 6):
                dy = np.dot(K[:s].T, _RK45_A[s, :s]) * h
                K[s] = f_eval(t + _RK45_C[s] * h, y + dy)
            y_new = y + h * np.dot(K[:6].T, _RK45_B)
            f_new = f_eval(t + h, y_new)
            K[6] = f_new
            scale = atol + np.maximum(np.abs(y), np.abs(y_new)) * rtol
            error_norm = _rms(np.dot(K.T, _RK45_E) * h / scale)
            if error_norm < 1:
                factor = MAX_FACTOR if error_norm == 0 else min(MAX_FACTOR, SAFETY * error_norm ** err_exp)
                if step_rejected:
                    factor = min(1, factor)
                h_abs *= factor
                step_accepted = True
                n_acc += 1
            else:
                h_abs *= max(MIN_FACTOR, SAFETY * error_norm ** err_exp)
                step_rejected = True
                n_rej += 1
        t, y, f = t_new, y_new, f_new
    return y, nfev, n_acc, n_rej


def ode_sampler(sd, pts_feat, pts_center, x0, T0: float = 1.0, rtol: float = 1e-5, atol: float = 1e-5,
                num_steps=None, denoise: bool = True, return_stats: bool = False, return_process: bool = False):
    """cond_ode_sampler, samplers.py:163-227 (T=T0, eps=1e-5).  x0 [R,9] is the already-noised
    start (prior sample at T0, plus init_x when tracking, :180).  State float64 on the host,
    score in fp32 (:191), drift-free VE: dx/dt = -0.5 g(t)^2 score.  NumPy-1.23 value-based
    promotion made `0.5*diffusion**2*score` float32 before SciPy upcast it; that is
    restated explicitly.  Output float64 like the reference (:206-207)."""
    R = x0.shape[0]
    pf = pts_feat.float()

    def ode_func(t, x):
        xt = torch.tensor(x.reshape(-1, 9), dtype=torch.float32)                # :191
        ts = torch.ones(R, 1) * t                                               # :192 (fp32)
        sigma = sigma_of_t(torch.tensor(t))                                     # sde_coeff(torch.tensor(t)): 0-dim fp32
        diffusion = sigma * G_COEF                                              # 0-dim fp32 * 0-dim fp64 -> fp64
        s = score(sd, pf, xt, ts).numpy().reshape(-1)                           # fp32
        # numpy<2 value-based casting: python/0-d float64 scalars do not upcast the fp32 array
        coef = np.float32(0.5 * float(diffusion.numpy()) ** 2)
        return (np.float32(0.0) - coef * s).astype(np.float32)

    y0 = x0.reshape(-1).double().numpy()
    y, nfev, n_acc, n_rej = rk45_integrate(ode_func, T0, EPS, y0, rtol, atol)
    x = torch.tensor(y).reshape(R, 9)                                           # float64 :207
    if denoise:                                                                 # :209-218
        vec_eps = torch.ones(R, 1) * EPS
        sigma = sigma_of_t(vec_eps)
        diffusion = sigma * G_COEF
        grad = score(sd, pf, x.float(), vec_eps)
        drift = 0 - diffusion ** 2 * grad
        x = x + drift * ((1 - EPS) / (1000 if num_steps is None else num_steps))
        nfev += 1
    x = x.clone()
    x[:, :-3] = normalize_rotation(x[:, :-3])                                   # :225
    x[:, -3:] += pts_center.to(x.dtype)                                         # :226
    if return_process:
        # the reference's `xs` (samplers.py:201-206, :220-224) through SciPy itself: accepted states, or the dense output at
        # t_eval = np.linspace(T, eps, num_steps) when num_steps is given; rotations normalised, pts_center added
        from scipy import integrate
        t_eval = None if num_steps is None else np.linspace(T0, EPS, num_steps)
        res = integrate.solve_ivp(ode_func, (T0, EPS), y0, rtol=rtol, atol=atol, method="RK45", t_eval=t_eval)
        xs = torch.tensor(res.y).T.reshape(-1, R, 9).clone()                    # [n, R, 9] float64
        n = xs.shape[0]
        flat = xs.reshape(n * R, 9)
        flat[:, :-3] = normalize_rotation(flat[:, :-3])
        xs = flat.reshape(n, R, 9)
        xs[:, :, -3:] += pts_center.to(xs.dtype).unsqueeze(0)
        process = xs.permute(1, 0, 2)                                           # [R, n, 9]
        if return_stats:
            return x, dict(nfev=nfev, accepted=n_acc, rejected=n_rej), process
        return x, process
    if return_stats:
        return x, dict(nfev=nfev, accepted=n_acc, rejected=n_rej)
    return x


# ------------------------------------------------------------------------------------------------
# a11-a13: agent-level functions
# ------------------------------------------------------------------------------------------------
def pred_func_pc(sd, data, repeat_num: int, num_steps: int, x0, step_noise, dtype=torch.float32, pts_feat=None):
    """PoseNet.pred_func, posenet_agent.py:416-439, sampler 'pc': encoder once, repeat K x, sample.
    pts_feat given: skip the encoder and sample from THESE features.  A sampler parity test passes the features of the
    encoder under test: a feature difference is the same bias error at every step, so a short chain amplifies it coherently
    (1e-5 of the feature scale moves a T = 20..30 pose by several 1e-3) and would otherwise decide the comparison; the
    encoder has its own tolerance (1e-4 of the scale) and the end-to-end anchor is the committed golden vectors."""
    pts_feat = encode(sd, data["pts"], dtype) if pts_feat is None else pts_feat.to(dtype)
    B = pts_feat.shape[0]
    rep_feat = pts_feat.unsqueeze(1).repeat(1, repeat_num, 1).view(B * repeat_num, -1)
    rep_center = data["pts_center"].unsqueeze(1).repeat(1, repeat_num, 1).view(B * repeat_num, -1)
    res = pc_sampler(sd, rep_feat, rep_center, x0, step_noise, num_steps, dtype=dtype)
    return res.reshape(B, repeat_num, -1), pts_feat


def get_energy(sd, data, pose_samples: torch.Tensor, T: float = 1e-5, pts_feat=None, dtype=torch.float32):
    """PoseNet.get_energy posenet_agent.py:471-527 (mode 'test', T given) ->
    PoseEnergyNet.get_energy energynet.py:143-198 with energy_mode 'IP', s_theta_mode 'score',
    norm_energy 'identical' (configs/config.py:40-42): [B,K,2] = (rot, trans) inner products."""
    B, K, _ = pose_samples.shape
    if pts_feat is None:
        pts_feat = encode(sd, data["pts"], dtype)
    rep_feat = pts_feat.to(dtype).unsqueeze(1).repeat(1, K, 1).view(B * K, -1)
    pose = pose_samples.clone().view(B * K, -1).to(dtype)
    t = torch.ones(B * K, 1, dtype=dtype) * T
    rep_center = data["pts_center"].to(dtype).unsqueeze(1).repeat(1, K, 1).view(B * K, -1)
    pose[:, -3:] -= rep_center                                                  # :516
    f = _trunk(sd, rep_feat, pose, t, dtype)
    s_theta = f / sigma_of_t(t)                                                 # energynet.py:166-167 (no +1e-7)
    e_rot = torch.sum(pose[:, :-3] * s_theta[:, :-3], dim=-1)                   # :181
    e_trans = torch.sum(pose[:, -3:] * s_theta[:, -3:], dim=-1)                 # :182
    return torch.stack([e_rot, e_trans], dim=-1).reshape(B, K, 2)


def sort_poses_by_energy(poses: torch.Tensor, energy: torch.Tensor):
    """networks/reward.py:131-155: descending sort per object; rotation part follows the rot-energy
    order, translation part the trans-energy order, independently."""
    sorted_energy, order = torch.sort(energy, descending=True, dim=1)
    rot = torch.gather(poses[:, :, :-3], 1, order[:, :, 0:1].expand(-1, -1, poses.shape[-1] - 3))
    trans = torch.gather(poses[:, :, -3:], 1, order[:, :, 1:2].expand(-1, -1, 3))
    return torch.cat([rot, trans], dim=-1), sorted_energy


def matrix_to_quaternion(m: torch.Tensor) -> torch.Tensor:
    """pytorch3d v0.7.2 matrix_to_quaternion (published definition), wxyz."""
    m00, m01, m02 = m[..., 0, 0], m[..., 0, 1], m[..., 0, 2]
    m10, m11, m12 = m[..., 1, 0], m[..., 1, 1], m[..., 1, 2]
    m20, m21, m22 = m[..., 2, 0], m[..., 2, 1], m[..., 2, 2]
    q_abs = torch.sqrt(torch.clamp(torch.stack([1.0 + m00 + m11 + m22, 1.0 + m00 - m11 - m22,
                                                1.0 - m00 + m11 - m22, 1.0 - m00 - m11 + m22], dim=-1), min=0))
    cand = torch.stack([
        torch.stack([q_abs[..., 0] ** 2, m21 - m12, m02 - m20, m10 - m01], dim=-1),
        torch.stack([m21 - m12, q_abs[..., 1] ** 2, m10 + m01, m02 + m20], dim=-1),
        torch.stack([m02 - m20, m10 + m01, q_abs[..., 2] ** 2, m12 + m21], dim=-1),
        torch.stack([m10 - m01, m20 + m02, m21 + m12, q_abs[..., 3] ** 2], dim=-1)], dim=-2)
    cand = cand / (2.0 * q_abs[..., None].clamp(min=0.1))
    best = q_abs.argmax(dim=-1)
    return torch.gather(cand, -2, best[..., None, None].expand(*best.shape, 1, 4)).squeeze(-2)


def quaternion_to_matrix(q: torch.Tensor) -> torch.Tensor:
    r, i, j, k = torch.unbind(q, -1)
    two_s = 2.0 / (q * q).sum(-1)
    o = torch.stack((1 - two_s * (j * j + k * k), two_s * (i * j - k * r), two_s * (i * k + j * r),
                     two_s * (i * j + k * r), 1 - two_s * (i * i + k * k), two_s * (j * k - i * r),
                     two_s * (i * k - j * r), two_s * (j * k + i * r), 1 - two_s * (i * i + j * j)), -1)
    return o.reshape(q.shape[:-1] + (3, 3))


def average_quaternion_batch(Q: torch.Tensor) -> torch.Tensor:
    """utils/misc.py:227-249 (uniform weights): sign-align w>0, mean outer product, top eigenvector."""
    oriented = ((Q[:, :, 0:1] > 0).to(Q.dtype) - 0.5) * 2 * Q
    A = torch.einsum("abi,abk->aik", oriented, oriented) / Q.shape[1]
    q_avg = torch.linalg.eigh(A)[1][:, :, -1]
    return ((q_avg[:, 0:1] > 0).to(Q.dtype) - 0.5) * 2 * q_avg


def rank_and_pool(poses: torch.Tensor, energy: torch.Tensor, ratio: float = 0.6):
    """pred_energy_batch (runners/evaluation_single.py:337-353) + sort_sRT_by_energy(ratio, 'average')
    (utils/sgpa_utils.py:897-954): returns sorted poses/energy and the pooled [B,4,4] transform.
    NOTE the second sort inside compute_mAP (sort_sRT) re-sorts already-sorted energies and is the
    identity permutation on distinct values."""
    sorted_poses, sorted_energy = sort_poses_by_energy(poses, energy)
    B, K, _ = poses.shape
    keep = max(1, int(K * ratio))                                               # sgpa_utils.py:912
    sel = sorted_poses[:, :keep]
    R = get_rot_matrix(sel[..., :6].reshape(B * keep, 6))
    q = matrix_to_quaternion(R).reshape(B, keep, 4)
    q_avg = average_quaternion_batch(q)
    t_avg = sel[..., 6:9].mean(dim=1)
    RT = torch.eye(4, dtype=poses.dtype).unsqueeze(0).repeat(B, 1, 1)
    RT[:, :3, :3] = quaternion_to_matrix(q_avg)
    RT[:, :3, 3] = t_avg
    return sorted_poses, sorted_energy, RT
