"""Host-side mirror of `GFObjectPose` (networks/posenet.py:18-179): same constructor, same mode-dispatching
forward, same `data` dict contract — but no nn.Module arithmetic: every mode calls the
sm_100a kernels through genpose_b200.ops.Engine.  Weights arrive through load_state_dict() with the
reference's exact key schema."""
from collections import OrderedDict
from typing import Dict, Optional

import numpy as np
import torch

from . import arch, lib, ops


class GFObjectPose:
    def __init__(self, cfg, prior_fn, marginal_prob_fn, sde_fn, sampling_eps, T):
        self.cfg = cfg
        self.device = cfg.device
        self.is_testing = False
        self.prior_fn = prior_fn
        self.marginal_prob_fn = marginal_prob_fn
        self.sde_fn = sde_fn
        self.sampling_eps = sampling_eps
        self.T = T
        if getattr(cfg, "pts_encoder", "pointnet2") != "pointnet2":
            raise NotImplementedError("only --pts_encoder pointnet2 (configs/config.py:39) is implemented")
        if getattr(cfg, "regression_head", "Rx_Ry_and_T") != "Rx_Ry_and_T" or getattr(cfg, "pose_mode", "rot_matrix") != "rot_matrix":
            raise NotImplementedError("only regression_head Rx_Ry_and_T / pose_mode rot_matrix are implemented")
        if getattr(cfg, "posenet_mode", "score") == "energy":
            if (getattr(cfg, "energy_mode", "IP"), getattr(cfg, "s_theta_mode", "score"), getattr(cfg, "norm_energy", "identical")) \
                    != ("IP", "score", "identical"):
                raise NotImplementedError("energy net: only energy_mode IP / s_theta_mode score / norm_energy identical")
        self.is_energy_net = getattr(cfg, "posenet_mode", "score") == "energy"
        self._state: Optional[Dict[str, torch.Tensor]] = None
        self.engine: Optional[ops.Engine] = None
        # noise for the PC sampler: 'philox' (in-kernel, throughput) or 'torch' (torch.randn_like draws in the
        # reference's order on the CUDA generator -> same stream as the reference under torch.manual_seed)
        self.noise_mode = getattr(cfg, "noise_mode", "philox")
        # engine of the dense layers inside the PC sampler: 'auto' (tcgen05 bf16x3 when the shape allows),
        # 'bf16x3' (force tensor cores) or 'fp32' (FFMA parity kernel)
        self.precision = getattr(cfg, "precision", "auto")
        self._philox_calls = 0

    # ---- nn.Module-like surface used by PoseNet --------------------------------------------------------
    def to(self, device):
        self.device = device
        return self

    def cuda(self):
        return self.to("cuda")

    def eval(self):
        self.is_testing = True
        return self

    def train(self, mode=True):
        return self

    def parameters(self):
        return [] if self._state is None else [v for v in self._state.values() if v.dtype.is_floating_point]

    def state_dict(self):
        if self._state is None:
            raise lib.GenPoseB200Error("no weights loaded")
        return OrderedDict((k, v.clone()) for k, v in self._state.items())

    def expected_keys(self):
        keys = []
        for l, lv in enumerate(arch.SA_LEVELS):
            for s in range(2):
                for j in range(3):
                    p = f"pts_encoder.SA_modules.{l}.mlps.{s}.layer{j}"
                    keys += [f"{p}.conv.weight"] + [f"{p}.bn.bn.{n}" for n in
                                                    ("weight", "bias", "running_mean", "running_var", "num_batches_tracked")]
        keys += [f"pose_score_net.pose_encoder.{i}.{n}" for i in (0, 2) for n in ("weight", "bias")]
        keys += ["pose_score_net.t_encoder.0.W", "pose_score_net.t_encoder.1.weight", "pose_score_net.t_encoder.1.bias"]
        keys += [f"pose_score_net.fusion_tail_{h}.{i}.{n}" for h in arch.HEADS for i in (0, 2) for n in ("weight", "bias")]
        return keys

    def load_state_dict(self, state_dict, strict=True):
        """Strict like the reference (posenet_agent.py:166-168): unknown or missing keys raise."""
        want, got = set(self.expected_keys()), set(state_dict.keys())
        if strict and want != got:
            raise RuntimeError(f"Error(s) in loading state_dict for GFObjectPose: missing {sorted(want - got)[:5]} "
                               f"unexpected {sorted(got - want)[:5]}")
        self._state = OrderedDict((k, v.detach().cpu().clone()) for k, v in state_dict.items())
        self.engine = ops.Engine(self._state, device="cuda" if str(self.device) == "cuda" else self.device)
        return self

    def _require_score_net(self, what: str):
        """An energy network's score is the autograd gradient of its energy (PoseEnergyNet.forward(return_item='score'),
        energynet.py:187-198): f/std + x . df/dx / std, not the trunk output f/std a score network returns.  Only the energy itself
        (inference: PoseNet.get_energy, evaluation_single.py:339-343) is implemented for --posenet_mode energy."""
        if self.is_energy_net:
            raise NotImplementedError(f"{what} on an energy network (--posenet_mode energy) needs the gradient of the energy; "
                                      "only mode='energy' / PoseNet.get_energy is implemented for it")

    def _eng(self) -> ops.Engine:
        if self.engine is None:
            raise lib.GenPoseB200Error("GFObjectPose has no weights: call load_state_dict()/PoseNet.load_ckpt() first")
        return self.engine

    # ---- modes ------------------------------------------------------------------------------------------
    def extract_pts_feature(self, data):
        """posenet.py:71-91 — uses data['pts'] (raw camera frame)."""
        return self._eng().encode(data["pts"].float().contiguous(), precision=self.precision)

    def _step_noise(self, num_steps, rows, device):
        if self.noise_mode == "torch":
            noise = torch.empty(num_steps, 2, rows, arch.POSE_DIM, device=device)
            like = noise[0, 0]
            for i in range(num_steps):                      # the reference's draw order (samplers.py:131,149)
                noise[i, 0] = torch.randn_like(like)
                noise[i, 1] = torch.randn_like(like)
            return noise, 0
        # philox: derive a fresh key from torch's CPU generator so torch.manual_seed() controls it
        seed = int(torch.randint(0, 2 ** 62, (1,)).item())
        return None, seed

    def sample_candidates(self, pts_feat, pts_center, repeat_num, sampler, init_x=None, T0=None, return_process=False,
                          step_noise=None):
        """Fast path used by PoseNet.pred_func: K candidates per object WITHOUT repeating the features.
        pts_feat [B,1024], pts_center [B,3] -> res [B*K,9] (+ process)."""
        self._require_score_net("sample_candidates")
        eng = self._eng()
        B = pts_feat.shape[0]
        R = B * repeat_num
        ob = eng.object_bias(pts_feat.float().contiguous())
        center = pts_center.float().contiguous()
        if sampler == "pc":
            num_steps = self.cfg.sampling_steps
            if num_steps is None:
                raise ValueError("--sampling_steps is required for the pc sampler (samplers.py:118 linspace)")
            x0 = self.prior_fn((R, arch.POSE_DIM)).to(pts_feat.device) if init_x is None else init_x      # samplers.py:117
            seed = 0
            if step_noise is None:
                step_noise, seed = self._step_noise(num_steps, R, pts_feat.device)
            out = eng.sample_pc(ob, center, x0.float().contiguous(), repeat_num, num_steps, step_noise=step_noise,
                                seed=seed, snr=0.16, return_process=return_process, precision=self.precision)
            return (out[1], out[0]) if return_process else (None, out)          # (in_process_sample [R,T,9], res [R,9]) like samplers.py:160
        elif sampler == "ode":
            T0 = self.T if T0 is None else T0
            prior = self.prior_fn((R, arch.POSE_DIM), T=T0).to(pts_feat.device)                            # samplers.py:180
            x0 = prior if init_x is None else init_x + prior
            num_steps = self.cfg.sampling_steps
            kw = dict(T0=T0, rtol=1e-5, atol=1e-5, denoise_steps=1000 if num_steps is None else num_steps, precision=self.precision)
            if not return_process:
                pose, stats = eng.sample_ode(ob, center, x0.float().contiguous(), repeat_num, **kw)
                self.last_ode_stats = stats
                return None, pose
            # in_process_sample (samplers.py:201-206, :220-224): SciPy's accepted states, or its dense output at
            # t_eval = np.linspace(T, eps, num_steps) when --sampling_steps is set; [n, R, 9] -> [R, n, 9] like xs.permute(1, 0, 2)
            t_eval = None if num_steps is None else np.linspace(float(T0), float(self.sampling_eps), int(num_steps))
            pose, stats, process = eng.sample_ode(ob, center, x0.float().contiguous(), repeat_num, return_process=True, t_eval=t_eval, **kw)
            self.last_ode_stats = stats
            return process.permute(1, 0, 2), pose
        raise NotImplementedError(sampler)

    def forward(self, data, mode="score", init_x=None, T0=None):
        """posenet.py:150-179.  In the sample modes `data['pts_feat']` may be per-row (the reference repeats
        it K times, posenet_agent.py:427-435); every row is then treated as its own object."""
        eng = self._eng()
        if mode == "pts_feature":
            return self.extract_pts_feature(data)
        if mode in ("score", "energy"):
            feat = data["pts_feat"].float().contiguous()
            pose = data["sampled_pose"].float().contiguous()
            t = data["t"]
            t0 = float(t.reshape(-1)[0])
            if not bool((t == t.reshape(-1)[0]).all()):
                raise NotImplementedError("per-row time values: both samplers and get_energy(T=...) use a batch-constant t")
            if mode == "score":
                self._require_score_net("forward(mode='score')")
            ob = eng.object_bias(feat)
            if mode == "score":
                return eng.trunk_eval(ob, pose, 1, t0, divide_mode=1)
            zero_center = torch.zeros(feat.shape[0], 3, device=feat.device)
            return eng.energy(ob, zero_center, pose, 1, t0)
        if mode in ("pc_sample", "ode_sample"):
            feat = data["pts_feat"]
            center = data["pts_center"]
            proc, res = self.sample_candidates(feat, center, 1, mode.split("_")[0], init_x=init_x, T0=T0, return_process=True)
            return proc, res
        raise NotImplementedError(mode)

    __call__ = forward
