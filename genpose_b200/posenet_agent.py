"""Host-side mirror of the reference agent `PoseNet` (networks/posenet_agent.py:46-527), inference surface
only: __init__, load_ckpt, pred_func, get_energy (+ the tiny helpers they need).  Same signatures, same
return shapes/dtypes, same `data` dict side effects (pred_func adds data['pts_feat'], :422).  Training
methods are out of scope and raise."""
import os

import numpy as np
import torch

from . import arch, lib, ops
from .posenet import GFObjectPose
from .sde import init_sde


class PoseNet:
    def __init__(self, cfg):
        self.cfg = cfg
        self.is_testing = False
        self.pts_feature = False
        if getattr(cfg, "is_train", False):
            raise NotImplementedError("genpose_b200 covers the inference hot path only (no training)")
        if getattr(cfg, "parallel", False):
            raise NotImplementedError("nn.DataParallel is replaced by one process per GPU (genpose_b200.distributed)")
        self.model_dir = f"./results/ckpts/{getattr(cfg, 'log_dir', 'debug')}"          # posenet_agent.py:36
        self.prior_fn, self.marginal_prob_fn, self.sde_fn, self.sampling_eps, self.T = init_sde(cfg.sde_mode)
        self.net = self.build_net()
        # optional callable(data) invoked inside pred_func once data['pts_feat'] has been enqueued (before the sampler is): lets a
        # pipeline order side-stream work behind the encoder instead of beside it (genpose_b200.pipeline.PosePipeline)
        self.on_features_ready = None

    def get_network(self, name):
        if name == "GFObjectPose":
            return GFObjectPose(self.cfg, self.prior_fn, self.marginal_prob_fn, self.sde_fn, self.sampling_eps, self.T)
        raise NotImplementedError(f"Got name '{name}'")

    def build_net(self):
        return self.get_network("GFObjectPose").to(self.cfg.device)

    def eval(self):
        self.net.eval()
        return self

    # ---- checkpoint contract (posenet_agent.py:143-173) ---------------------------------------------------
    def load_ckpt(self, name=None, model_dir=None, model_path=False, load_model_only=False):
        if not model_path:
            if name not in ("latest", "best"):
                name = "ckpt_epoch{}".format(name)
            load_path = os.path.join(self.model_dir if model_dir is None else model_dir, "{}.pth".format(name))
        else:
            load_path = model_dir
        if not os.path.exists(load_path):
            raise ValueError("Checkpoint {} not exists.".format(load_path))
        checkpoint = torch.load(load_path, map_location="cpu")
        print("Loading checkpoint from {} ...".format(load_path))
        self.net.load_state_dict(checkpoint["model_state_dict"])
        # optimizer / scheduler / clock state exist only for training; ignored (load_model_only semantics)

    # ---- K-candidate inference (posenet_agent.py:416-468) ----------------------------------------------------
    def pred_func(self, data, repeat_num, save_path="./visualization_results", return_average_res=False, init_x=None,
                  T0=None, return_process=False):
        self.is_testing = True
        self.net.eval()
        with torch.no_grad():
            data["pts_feat"] = self.net(data, mode="pts_feature")                       # :422 (side effect kept)
            if self.on_features_ready is not None:
                self.on_features_ready(data)
            bs = data["pts"].shape[0]
            self.pts_feature = True
            repeated_init_x = None if init_x is None else init_x.unsqueeze(1).repeat(1, repeat_num, 1).view(bs * repeat_num, -1)
            sampler = self.cfg.sampler_mode[0]
            in_process_sample, res = self.net.sample_candidates(
                data["pts_feat"], data["pts_center"], repeat_num, sampler, init_x=repeated_init_x, T0=T0,
                return_process=return_process)
            pred_pose = res.reshape(bs, repeat_num, -1)
            if return_process:
                in_process_sample = in_process_sample.reshape(bs, repeat_num, in_process_sample.shape[1], -1)
            self.pts_feature = False
            if return_average_res:
                pose_f = pred_pose.float().contiguous()
                zeros = torch.zeros(bs, repeat_num, 2, device=pose_f.device)           # equal energies: stable order = identity
                _, _, rt = ops.rank_pool(pose_f, zeros, ratio=1.0)
                pred_pose_q_wxyz = _poses_to_quat(res.float()).reshape(bs, repeat_num, -1)
                q_avg = _matrix_to_quat(rt[:, :3, :3])
                q_avg = ((q_avg[:, 0:1] > 0).float() - 0.5) * 2 * q_avg                 # oriented w > 0 (utils/misc.py:247)
                average = torch.cat([q_avg, rt[:, :3, 3]], dim=-1)
                if return_process:
                    return pred_pose, pred_pose_q_wxyz, average, in_process_sample
                return pred_pose, pred_pose_q_wxyz, average
            if return_process:
                return [pred_pose, in_process_sample]
            return pred_pose

    # ---- energies of given candidates (posenet_agent.py:471-527) -----------------------------------------------
    def get_energy(self, data, pose_samples, T=None, mode="test", extract_pts_feature=True):
        if mode != "test":
            raise NotImplementedError("get_energy(mode='train') needs autograd; inference path only")
        self.is_testing = True
        self.net.eval()
        bs, repeat_num = pose_samples.shape[0], pose_samples.shape[1]
        with torch.no_grad():
            pts_feat = data["pts_feat"] if extract_pts_feature is False else self.net(data, mode="pts_feature")
            self.pts_feature = True
            if T is None:
                raise NotImplementedError("get_energy(T=None) draws per-object random T for training-time ranking; pass T")
            eng = self.net._eng()
            ob = eng.object_bias(pts_feat.float().contiguous())
            pose = pose_samples.reshape(bs * repeat_num, -1).to(pts_feat.dtype).contiguous()
            energy = eng.energy(ob, data["pts_center"].float().contiguous(), pose, repeat_num, float(T))
            return energy.reshape(bs, repeat_num, -1)

    # ---- training surface: out of scope -----------------------------------------------------------------------
    def _no_training(self, *a, **k):
        raise NotImplementedError("training is out of scope for genpose_b200 (SURVEY.md §2.1)")

    train_score_func = train_energy_func = eval_score_func = eval_energy_func = save_ckpt = update_network = _no_training


def _matrix_to_quat(m: torch.Tensor) -> torch.Tensor:
    """pytorch3d matrix_to_quaternion (published definition) for small host-side conversions."""
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


def _poses_to_quat(res: torch.Tensor) -> torch.Tensor:
    """[R,9] (rx, ry, t) -> [R,7] (quat wxyz, t): get_rot_matrix (utils/misc.py:136) + matrix_to_quaternion."""
    b1 = torch.nn.functional.normalize(res[:, 0:3], dim=-1)
    a2 = res[:, 3:6]
    b2 = torch.nn.functional.normalize(a2 - (b1 * a2).sum(-1, keepdim=True) * b1, dim=-1)
    b3 = torch.cross(b1, b2, dim=-1)
    R = torch.stack([b1, b2, b3], dim=-1)
    return torch.cat([_matrix_to_quat(R), res[:, 6:9]], dim=-1)
