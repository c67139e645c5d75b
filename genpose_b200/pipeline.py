"""The per-object inference pipeline as one call: encoder -> K-candidate sampler -> (energy -> rank -> pool).
This is `inference_pose` + `inference_energy` of runners/evaluation_single.py:356-489 without the pickle
round-trip between the two agents.  It is built on the two agents' public methods
(PoseNet.pred_func / get_energy), so it exercises exactly the reference-facing surface."""
from types import SimpleNamespace
from typing import Dict, Optional

import torch

from . import ops
from .config import get_config
from .posenet_agent import PoseNet


def default_cfg(sampler="pc", sampling_steps=500, posenet_mode="score", noise_mode="philox", precision="auto", extra=()):
    argv = ["--sampler_mode", sampler, "--posenet_mode", posenet_mode, "--noise_mode", noise_mode, "--precision", precision]
    if sampling_steps is not None:
        argv += ["--sampling_steps", str(sampling_steps)]
    return get_config(argv + list(extra))


def poses_to_RTs(pred_pose: torch.Tensor):
    """[B,K,9] poses (rx, ry, t) -> float64 numpy [B,K,4,4], the `RTs_all` that pred_pose_batch / pred_energy_batch build
    (runners/evaluation_single.py:325-332, :346-353) with one get_rot_matrix call and two `.cpu().numpy()` copies PER CANDIDATE.
    Here: all B*K candidates at once on the device the poses live on and ONE device->host copy.  Result formatting after the hot
    path (the runner pickles it for compute_mAP); the columns are get_rot_matrix's (utils/misc.py:136: Gram-Schmidt b1, b2 and
    b1 x b2 as matrix COLUMNS, F.normalize eps 1e-12)."""
    import numpy as np
    p = pred_pose.detach()
    b1 = torch.nn.functional.normalize(p[..., 0:3], dim=-1)
    a2 = p[..., 3:6]
    b2 = torch.nn.functional.normalize(a2 - (b1 * a2).sum(-1, keepdim=True) * b1, dim=-1)
    b3 = torch.cross(b1, b2, dim=-1)
    rt = torch.zeros(*p.shape[:-1], 4, 4, dtype=p.dtype, device=p.device)
    rt[..., :3, 0], rt[..., :3, 1], rt[..., :3, 2], rt[..., :3, 3] = b1, b2, b3, p[..., 6:9]
    rt[..., 3, 3] = 1.0
    return rt.cpu().numpy().astype(np.float64)


class PosePipeline:
    def __init__(self, score_state_dict, energy_state_dict=None, sampler="pc", sampling_steps=500, noise_mode="philox",
                 precision="auto"):
        self.score_agent = PoseNet(default_cfg(sampler, sampling_steps, "score", noise_mode, precision))
        self.score_agent.net.load_state_dict(score_state_dict)
        self.energy_agent = None
        self._side = None
        if energy_state_dict is not None:
            self.energy_agent = PoseNet(default_cfg(sampler, sampling_steps, "energy", noise_mode, precision))
            self.energy_agent.net.load_state_dict(energy_state_dict)
            # The energy network has its own encoder (evaluation_single.py:431) whose input is the same cloud, and the
            # tensor-core sampler holds only 100 of the 148 SMs at 64 objects: the second encoder pass runs on a side stream, on
            # the SMs the sampler leaves free, instead of after it.  It is ordered BEHIND the score network's encoder (an event
            # recorded between that encoder and the sampler), so the two encoder passes do not slow each other on the critical path.
            self._side = torch.cuda.Stream()
            self.score_agent.on_features_ready = self._mark_score_encoder_done

    def _mark_clouds_ready(self):
        """Call on entry, before the score path is enqueued: whatever produced the clouds on the main stream is ordered
        before this event."""
        if self._side is not None:
            self._clouds_ready = torch.cuda.Event()
            self._clouds_ready.record(torch.cuda.current_stream())

    def _mark_score_encoder_done(self, data):
        self._clouds_ready = torch.cuda.Event()
        self._clouds_ready.record(torch.cuda.current_stream())

    def _energy_features_async(self, data):
        """Start the energy net's encoder on the side stream -> pts_feat (join with _energy_rank_pool).  It waits for the
        clouds only (the event of _mark_clouds_ready), NOT for the main stream: it is launched after the sampler has been
        enqueued and must run beside it, not behind it."""
        main = torch.cuda.current_stream()
        self._side.wait_event(self._clouds_ready)
        with torch.cuda.stream(self._side):
            feat = self.energy_agent.net(data, mode="pts_feature")
        feat.record_stream(main)
        return feat

    def _energy_rank_pool(self, data, efeat, pred_pose, ratio):
        pose = pred_pose.float().contiguous()
        torch.cuda.current_stream().wait_stream(self._side)
        d2 = dict(data)
        d2["pts_feat"] = efeat                           # posenet_agent.py:484: extract_pts_feature=False reads data['pts_feat']
        energy = self.energy_agent.get_energy(data=d2, pose_samples=pose, T=1e-5, extract_pts_feature=False)   # evaluation_single.py:339-343
        sp, se, rt = ops.rank_pool(pose, energy.contiguous(), ratio=ratio)                                     # :344 + sgpa_utils.py:897
        return dict(energy=energy, sorted_pose=sp, sorted_energy=se, pooled_RT=rt)

    @staticmethod
    def make_batch(pts: torch.Tensor) -> Dict[str, torch.Tensor]:
        """The `data` dict of runners/evaluation_single.py:394-403 from device clouds [B,1024,3]."""
        center = torch.mean(pts[:, :, :3], dim=1)
        return {"pts": pts, "zero_mean_pts": pts - center.unsqueeze(1), "pts_center": center}

    def run(self, data: Dict[str, torch.Tensor], repeat_num: int = 50, T0: Optional[float] = None, ratio: float = 0.6):
        """-> dict(pred_pose [B,K,9], and with an energy net: energy [B,K,2], sorted_pose, sorted_energy, pooled_RT [B,4,4])."""
        self._mark_clouds_ready()
        out = {"pred_pose": self.score_agent.pred_func(data, repeat_num=repeat_num, save_path=None, T0=T0)}
        if self.energy_agent is not None:
            efeat = self._energy_features_async(data)     # launched while the sampler (already in flight) runs
            out.update(self._energy_rank_pool(data, efeat, out["pred_pose"], ratio))
        return out

    def track_step(self, data: Dict[str, torch.Tensor], initial_sRT: torch.Tensor, repeat_num: int = 50, T0: float = 0.15,
                   ratio: float = 0.6):
        """One frame of the tracking runner (runners/evaluation_tracking.py:302-316): warm-start every object from the
        previous frame's pooled pose `initial_sRT [B,4,4]` — initial_pose = [R[:, 0] | R[:, 1] | t - pts_center] (:309-310) —
        sample K candidates from T0 (`cond_ode_sampler` adds the T0-prior noise to init_x, samplers.py:180), rank by
        energy, pool the best `ratio` (cal_average_sRT, :58-75).  -> run()'s dict; out['pooled_RT'] feeds the next frame."""
        initial_pose = initial_sRT[:, :3, [0, 1, 3]].permute(0, 2, 1).reshape(initial_sRT.shape[0], -1).float().clone()
        initial_pose[:, -3:] -= data["pts_center"]
        self._mark_clouds_ready()
        out = {"pred_pose": self.score_agent.pred_func(data, repeat_num=repeat_num, save_path=None, init_x=initial_pose, T0=T0)}
        if self.energy_agent is not None:
            efeat = self._energy_features_async(data)
            out.update(self._energy_rank_pool(data, efeat, out["pred_pose"], ratio))
        return out

    def run_stream(self, batches, repeat_num: int = 50):
        """Throughput mode for a stream of independent batches (PC sampler): yields pred_pose [B,K,9] per batch.  Batch
        i+1's encoder is launched on a side stream as soon as batch i's sampler is in flight and runs beside it on the SMs the
        sampler leaves free, so a batch costs the sampler's time only (bench.py `pipelined`: 8.4 -> 7.4 ms per 64 objects).
        Same kernels and results as run(); only the launch order differs.  `batches` yields `data` dicts (make_batch)."""
        net = self.score_agent.net
        if self.score_agent.cfg.sampler_mode[0] != "pc":
            raise NotImplementedError("run_stream pipelines the PC sampler; use run() for the ODE sampler")
        eng = net._eng()
        side = self._side if self._side is not None else torch.cuda.Stream()
        main = torch.cuda.current_stream()

        def clouds_ready():
            ev = torch.cuda.Event()
            ev.record(main)
            return ev

        def encode(data, ready):
            """Encoder + object bias of `data` on the side stream; it waits only for `ready` (recorded on the main stream when the
            batch was pulled, i.e. BEFORE the previous batch's sampler was enqueued), never for that sampler."""
            side.wait_event(ready)
            with torch.cuda.stream(side):
                ob = eng.object_bias(eng.encode(data["pts"].float().contiguous(), precision=net.precision))
            ob.record_stream(main)
            return ob

        it = iter(batches)
        cur = next(it, None)
        ob = encode(cur, clouds_ready()) if cur is not None else None
        while cur is not None:
            nxt = next(it, None)                      # pull the next batch (its producer may enqueue copies on the main stream) ...
            nxt_ready = clouds_ready() if nxt is not None else None      # ... and mark it ready before this batch's sampler is enqueued
            main.wait_stream(side)
            B = cur["pts"].shape[0]
            R = B * repeat_num
            x0 = net.prior_fn((R, 9)).to(cur["pts"].device)                                      # samplers.py:117
            noise, seed = net._step_noise(net.cfg.sampling_steps, R, cur["pts"].device)          # (None, key) in philox mode
            pose = eng.sample_pc(ob, cur["pts_center"].float().contiguous(), x0.float().contiguous(), repeat_num,
                                 net.cfg.sampling_steps, step_noise=noise, seed=seed, snr=0.16, precision=net.precision)
            if nxt is not None:
                ob = encode(nxt, nxt_ready)
            yield pose.reshape(B, repeat_num, 9)
            cur = nxt
