"""Host-side (PyTorch) wrappers over the C ABI.  PyTorch is used for device memory and streams only;
every arithmetic operation below runs in the hand-written sm_100a kernels of libgenpose_b200.so.
All tensors must be CUDA, contiguous, fp32 / int32; violations raise (no silent copies to other devices,
no CPU fallback)."""
from typing import Dict, Optional, Tuple

import torch

from . import arch, lib, weights

# what precision="auto" means for the samplers when the tensor-core kernels accept the shape
AUTO_TC_PRECISION = "f16x2"


def _chk(t: torch.Tensor, dtype, name: str) -> int:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise lib.GenPoseB200Error(f"{name}: expected a CUDA tensor (there is no CPU path)")
    if t.dtype != dtype or not t.is_contiguous():
        raise lib.GenPoseB200Error(f"{name}: expected contiguous {dtype}, got {t.dtype} contiguous={t.is_contiguous()}")
    return t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


# -------------------------------------------------------------------------------------------------
# compat ops (reference pointnet2_utils.py semantics)
# -------------------------------------------------------------------------------------------------
def furthest_point_sample(xyz: torch.Tensor, npoint: int) -> torch.Tensor:
    """pointnet2_utils.py:13-30: xyz [B,N,3] -> idx [B,npoint] int32."""
    B, N, _ = xyz.shape
    idx = torch.empty(B, npoint, dtype=torch.int32, device=xyz.device)
    temp = torch.empty(B, N, dtype=torch.float32, device=xyz.device)
    lib.check(lib.load().gpb_furthest_point_sampling(B, N, npoint, _chk(xyz, torch.float32, "xyz"), temp.data_ptr(),
                                                     idx.data_ptr(), _stream()), "furthest_point_sampling")
    return idx


def gather_points(points: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """pointnet2_utils.py:43-61: points [B,C,N], idx [B,npoint] -> [B,C,npoint]."""
    B, C, N = points.shape
    npoint = idx.shape[1]
    out = torch.empty(B, C, npoint, dtype=torch.float32, device=points.device)
    lib.check(lib.load().gpb_gather_points(B, C, N, npoint, _chk(points, torch.float32, "points"),
                                           _chk(idx, torch.int32, "idx"), out.data_ptr(), _stream()), "gather_points")
    return out


def ball_query(radius: float, nsample: int, xyz: torch.Tensor, new_xyz: torch.Tensor) -> torch.Tensor:
    """pointnet2_utils.py:204-222: -> idx [B,npoint,nsample] int32."""
    B, N, _ = xyz.shape
    npoint = new_xyz.shape[1]
    idx = torch.empty(B, npoint, nsample, dtype=torch.int32, device=xyz.device)
    lib.check(lib.load().gpb_ball_query(B, N, npoint, float(radius), nsample, _chk(new_xyz, torch.float32, "new_xyz"),
                                        _chk(xyz, torch.float32, "xyz"), idx.data_ptr(), _stream()), "ball_query")
    return idx


def group_points(points: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """pointnet2_utils.py:160-178: points [B,C,N], idx [B,npoint,nsample] -> [B,C,npoint,nsample]."""
    B, C, N = points.shape
    _, npoint, nsample = idx.shape
    out = torch.empty(B, C, npoint, nsample, dtype=torch.float32, device=points.device)
    lib.check(lib.load().gpb_group_points(B, C, N, npoint, nsample, _chk(points, torch.float32, "points"),
                                          _chk(idx, torch.int32, "idx"), out.data_ptr(), _stream()), "group_points")
    return out


# -------------------------------------------------------------------------------------------------
# fused path
# -------------------------------------------------------------------------------------------------
def time_grid(num_steps: int, device) -> torch.Tensor:
    """torch.linspace(1., eps, num_steps) exactly as samplers.py:118 builds it (fp32)."""
    return torch.linspace(1.0, arch.SAMPLING_EPS, num_steps, device="cpu").to(device)


class Engine:
    """Device-resident packed weights of ONE network (encoder + score-or-energy trunk) plus scratch.
    Built from a reference `model_state_dict`; mirrors what GFObjectPose owns (networks/posenet.py:18-66)."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], device="cuda"):
        L = lib.load()
        self.device = torch.device(device)
        enc = weights.pack_encoder(state_dict)
        trunk = weights.pack_trunk(state_dict)
        if enc.numel() != L.gpb_encoder_weights_floats() or trunk.numel() != L.gpb_trunk_weights_floats():
            raise lib.GenPoseB200Error("packed weight sizes disagree with the library "
                                       f"({enc.numel()} vs {L.gpb_encoder_weights_floats()}, "
                                       f"{trunk.numel()} vs {L.gpb_trunk_weights_floats()})")
        self.enc_w = enc.to(self.device)
        self.trunk_w = trunk.to(self.device)
        tc = weights.pack_trunk_tc(state_dict)
        if tc.numel() * 2 != L.gpb_trunk_tc_stream_bytes():
            raise lib.GenPoseB200Error("tensor-core weight stream size disagrees with the library")
        self.trunk_tc = tc.to(self.device)
        etc = weights.pack_encoder_tc(state_dict)
        if etc.numel() != L.gpb_encoder_tc_bytes():
            raise lib.GenPoseB200Error("tensor-core encoder image size disagrees with the library")
        self.enc_tc = etc.to(self.device)
        self._ws: Dict[Tuple[str, int], torch.Tensor] = {}
        self._state_for_tc16 = state_dict
        self._time_grids: Dict[int, torch.Tensor] = {}
        self._trunk_tc16: Optional[torch.Tensor] = None

    def trunk_tc16(self) -> torch.Tensor:
        """Two-product stream (weights.pack_trunk_tc16), packed on first use of precision='f16x2'."""
        if self._trunk_tc16 is None:
            tc = weights.pack_trunk_tc16(self._state_for_tc16)
            if tc.numel() * 2 != lib.load().gpb_trunk_tc16_stream_bytes():
                raise lib.GenPoseB200Error("two-product weight stream size disagrees with the library")
            self._trunk_tc16 = tc.to(self.device)
        return self._trunk_tc16

    def _workspace(self, kind: str, nbytes: int) -> torch.Tensor:
        ws = self._ws.get(kind)
        if ws is None or ws.numel() < nbytes:
            ws = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=self.device)
            self._ws[kind] = ws
        return ws

    # ---- a7: encoder ------------------------------------------------------------------------------
    def encode(self, pts: torch.Tensor, return_fps: bool = False, precision: str = "auto"):
        """Pointnet2ClsMSG.forward: pts [B,1024,3] (raw camera frame) -> pts_feat [B,1024].
        precision 'fp32' = every level on FFMA; 'bf16x3' / 'auto' = set-abstraction level 3 on tcgen05 (bf16x3 split)."""
        if precision not in ("auto", "bf16x3", "f16x2", "fp32"):      # 'f16x2' concerns the samplers only: tensor-core encoder as 'bf16x3'
            raise lib.GenPoseB200Error(f"encode: unknown precision {precision!r}")
        B, N, C = pts.shape
        if N != arch.NUM_POINTS or C != 3:
            raise lib.GenPoseB200Error(f"encode: expected [B,1024,3], got {tuple(pts.shape)}")
        L = lib.load()
        ws = self._workspace("enc", L.gpb_encode_workspace_bytes(B))
        feat = torch.empty(B, arch.PTS_FEAT_DIM, dtype=torch.float32, device=self.device)
        fps = [None, None, None]
        if return_fps:
            fps = [torch.empty(B, n, dtype=torch.int32, device=self.device) for n in (512, 256, 128)]
        tail = (feat.data_ptr(), ws.data_ptr(), ws.numel(), *[0 if f is None else f.data_ptr() for f in fps], _stream())
        if precision == "fp32":
            lib.check(L.gpb_encode(_chk(pts, torch.float32, "pts"), B, self.enc_w.data_ptr(), *tail), "encode")
        else:
            lib.check(L.gpb_encode_tc(_chk(pts, torch.float32, "pts"), B, self.enc_w.data_ptr(), self.enc_tc.data_ptr(), *tail),
                      "encode_tc")
        return (feat, fps) if return_fps else feat

    def object_bias(self, pts_feat: torch.Tensor) -> torch.Tensor:
        B = pts_feat.shape[0]
        out = torch.empty(B, 768, dtype=torch.float32, device=self.device)
        lib.check(lib.load().gpb_object_bias(_chk(pts_feat, torch.float32, "pts_feat"), B, self.trunk_w.data_ptr(),
                                             out.data_ptr(), _stream()), "object_bias")
        return out

    # ---- a8: one score evaluation --------------------------------------------------------------------
    def trunk_eval(self, obj_bias: torch.Tensor, pose: torch.Tensor, K: int, t: float, divide_mode: int = 1) -> torch.Tensor:
        R = pose.shape[0]
        out = torch.empty(R, 9, dtype=torch.float32, device=self.device)
        lib.check(lib.load().gpb_trunk_eval(_chk(pose, torch.float32, "pose"), R, K, float(t),
                                            _chk(obj_bias, torch.float32, "obj_bias"), self.trunk_w.data_ptr(),
                                            divide_mode, out.data_ptr(), _stream()), "trunk_eval")
        return out

    # ---- a9: PC sampler -------------------------------------------------------------------------------
    @staticmethod
    def tc_supported(R: int, K: int) -> bool:
        """tcgen05 sampler constraints (asked of the library): a 128-row tile spans <= 8 objects (K >= 19) and every tile
        needs one co-resident team of CTAs — 4 per tile up to 33 tiles, 2 up to 66, 1 up to one tile per SM (148 tiles =
        18,944 rows = 378 objects x 50 candidates on a B200)."""
        return 0 < R <= lib.load().gpb_sampler_tc_max_rows(int(K))

    def _resolve_precision(self, what: str, precision: str, R: int, K: int) -> str:
        if precision == "auto":
            precision = AUTO_TC_PRECISION if self.tc_supported(R, K) else "fp32"
        if precision not in ("fp32", "bf16x3", "f16x2"):
            raise lib.GenPoseB200Error(f"{what}: unknown precision {precision!r}")
        return precision

    def sample_pc(self, obj_bias: torch.Tensor, pts_center: torch.Tensor, x0: torch.Tensor, K: int, num_steps: int,
                  step_noise: Optional[torch.Tensor] = None, seed: int = 0, snr: float = arch.SNR,
                  return_process: bool = False, precision: str = "fp32", team: int = 0):
        """precision 'fp32' = FFMA parity kernel; 'bf16x3' = tcgen05 kernel, three products per K-step (bf16 hi/lo operands);
        'f16x2' = tcgen05 kernel, two products (fp16 hi/lo activations x one fp16 weight image); 'auto' = tensor cores
        (AUTO_TC_PRECISION) when the shape allows.  team: 0 = tile-team size chosen from the row count, 1 / 2 / 4 force it."""
        R = x0.shape[0]
        precision = self._resolve_precision("sample_pc", precision, R, K)
        L = lib.load()
        ws = self._workspace("samp", L.gpb_sampler_workspace_bytes(R, num_steps))
        ts = self._time_grids.get(num_steps)         # built once per T: a pageable host->device copy per call would make the host
        if ts is None:                               # wait for the encoder in front of it before it can enqueue the sampler
            ts = self._time_grids[num_steps] = time_grid(num_steps, self.device)
        mean_x = torch.empty(R, 9, dtype=torch.float32, device=self.device)
        process = torch.empty(R, num_steps, 9, dtype=torch.float32, device=self.device) if return_process else None
        if step_noise is not None and tuple(step_noise.shape) != (num_steps, 2, R, 9):
            raise lib.GenPoseB200Error(f"sample_pc: step_noise must be [T,2,R,9], got {tuple(step_noise.shape)}")
        common = (0 if step_noise is None else _chk(step_noise, torch.float32, "step_noise"), int(seed) & (2 ** 64 - 1),
                  ts.data_ptr(), mean_x.data_ptr(), 0 if process is None else process.data_ptr(), ws.data_ptr(), ws.numel(),
                  _stream())
        head = (_chk(x0, torch.float32, "x0"), R, K, num_steps, float(snr), _chk(obj_bias, torch.float32, "obj_bias"),
                self.trunk_w.data_ptr())
        if precision != "fp32" and team:
            lib.check(L.gpb_set_tc_team(int(team)), "set_tc_team")          # a forced team size holds for this launch only
        try:
            if precision == "bf16x3":
                lib.check(L.gpb_sample_pc_tc(*head, self.trunk_tc.data_ptr(), _chk(pts_center, torch.float32, "pts_center"), *common),
                          "sample_pc_tc")
            elif precision == "f16x2":
                lib.check(L.gpb_sample_pc_tc16(*head, self.trunk_tc16().data_ptr(), _chk(pts_center, torch.float32, "pts_center"),
                                               *common[:-1], 0, common[-1]), "sample_pc_tc16")
            else:
                lib.check(L.gpb_sample_pc(*head, _chk(pts_center, torch.float32, "pts_center"), *common), "sample_pc")
        finally:
            if precision != "fp32" and team:
                L.gpb_set_tc_team(0)
        return (mean_x, process) if return_process else mean_x

    # ---- a10: ODE sampler --------------------------------------------------------------------------------
    def sample_ode(self, obj_bias: torch.Tensor, pts_center: torch.Tensor, x0: torch.Tensor, K: int, T0: float = 1.0,
                   rtol: float = 1e-5, atol: float = 1e-5, denoise_steps: int = 1000, precision: str = "fp32", team: int = 0,
                   return_process: bool = False, t_eval=None, process_cap: int = 192):
        """cond_ode_sampler (samplers.py:163-227): RK45 with SciPy's controller on device -> (pose [R,9] float64, stats [4]).
        precision and team as in sample_pc.  return_process: additionally the trajectory `xs` the reference returns
        (samplers.py:206, :220-224) as float64 [n, R, 9] — the solver's accepted states (n = accepted + 1) or, with
        t_eval (float64 times, decreasing from T0 to eps: np.linspace(T0, eps, num_steps), :203), RK45's dense output there."""
        R = x0.shape[0]
        precision = self._resolve_precision("sample_ode", precision, R, K)
        L = lib.load()
        ws = self._workspace("samp", L.gpb_sampler_workspace_bytes(R, 1))
        pose = torch.empty(R, 9, dtype=torch.float64, device=self.device)
        stats = torch.zeros(4, dtype=torch.int32, device=self.device)
        process, te, n_te = None, None, 0
        if return_process:
            if t_eval is not None:
                te = torch.as_tensor(t_eval, dtype=torch.float64).to(self.device).contiguous()
                n_te = int(te.numel())
                if n_te < 1 or (n_te > 1 and not bool((te[1:] < te[:-1]).all())):
                    raise lib.GenPoseB200Error("sample_ode: t_eval must be strictly decreasing (np.linspace(T0, eps, num_steps))")
                process = torch.empty(n_te, R, 9, dtype=torch.float64, device=self.device)
            else:
                process = torch.empty(int(process_cap), R, 9, dtype=torch.float64, device=self.device)
        head = (_chk(x0, torch.float32, "x0"), R, K, float(T0), float(rtol), float(atol), int(denoise_steps),
                _chk(obj_bias, torch.float32, "obj_bias"), self.trunk_w.data_ptr())
        tail = (_chk(pts_center, torch.float32, "pts_center"), pose.data_ptr(), stats.data_ptr(),
                0 if process is None else process.data_ptr(), int(process_cap), 0 if te is None else te.data_ptr(), n_te,
                ws.data_ptr(), ws.numel(), _stream())
        if precision != "fp32" and team:
            lib.check(L.gpb_set_tc_team(int(team)), "set_tc_team")          # a forced team size holds for this launch only
        try:
            if precision == "bf16x3":
                lib.check(L.gpb_sample_ode_tc(*head, self.trunk_tc.data_ptr(), *tail), "sample_ode_tc")
            elif precision == "f16x2":
                lib.check(L.gpb_sample_ode_tc16(*head, self.trunk_tc16().data_ptr(), *tail), "sample_ode_tc16")
            else:
                lib.check(L.gpb_sample_ode(*head, *tail), "sample_ode")
        finally:
            if precision != "fp32" and team:
                L.gpb_set_tc_team(0)
        if not return_process:
            return pose, stats
        if te is None:
            n = int(stats[1]) + 1                              # start + accepted steps (synchronises: the trajectory is a visualisation path)
            if n > process_cap:
                raise lib.GenPoseB200Error(f"sample_ode: {n} trajectory states exceed process_cap={process_cap}; raise it")
            process = process[:n]
        return pose, stats, process

    # ---- a12: energy ------------------------------------------------------------------------------------------
    def energy(self, obj_bias: torch.Tensor, pts_center: torch.Tensor, pose: torch.Tensor, K: int, t: float = 1e-5) -> torch.Tensor:
        R = pose.shape[0]
        out = torch.empty(R, 2, dtype=torch.float32, device=self.device)
        lib.check(lib.load().gpb_energy(_chk(pose, torch.float32, "pose"), R, K, float(t),
                                        _chk(obj_bias, torch.float32, "obj_bias"), self.trunk_w.data_ptr(),
                                        _chk(pts_center, torch.float32, "pts_center"), out.data_ptr(), _stream()), "energy")
        return out


def rank_pool(pose: torch.Tensor, energy: torch.Tensor, ratio: float = 0.6, want_pooled: bool = True, want_order: bool = False):
    """sort_poses_by_energy (reward.py:131-155) + sort_sRT_by_energy(ratio,'average') pooling
    (sgpa_utils.py:897-954).  pose [B,K,9], energy [B,K,2] -> sorted_pose, sorted_energy, pooled_RT [B,4,4]
    (+ order [B,K,2] int32 = the sort's indices per energy column when want_order)."""
    B, K, _ = pose.shape
    keep = max(1, int(K * ratio))                                   # sgpa_utils.py:912
    sp = torch.empty_like(pose)
    se = torch.empty_like(energy)
    rt = torch.empty(B, 4, 4, dtype=torch.float32, device=pose.device) if want_pooled else None
    order = torch.empty(B, K, 2, dtype=torch.int32, device=pose.device) if want_order else None
    lib.check(lib.load().gpb_rank_pool(_chk(pose, torch.float32, "pose"), _chk(energy, torch.float32, "energy"), B, K, keep,
                                       sp.data_ptr(), se.data_ptr(), 0 if rt is None else rt.data_ptr(),
                                       0 if order is None else order.data_ptr(), _stream()), "rank_pool")
    return (sp, se, rt, order) if want_order else (sp, se, rt)
