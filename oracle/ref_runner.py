"""ORACLE — test infrastructure, NOT product code.

Drives the UNMODIFIED reference (via oracle/ref_loader.py) on synthetic inputs with the random
draws injected, so that its outputs can be stored as golden vectors (oracle/make_golden.py) and
compared with the portable restatement (tests/test_oracle_vs_reference.py).

Noise injection touches only the SOURCES of randomness, never arithmetic:
  * prior draw  : `GFObjectPose.prior_fn` (networks/posenet.py:27) is replaced by a callable that
                  returns the supplied x0 (already scaled by sigma(T0)), instead of
                  `torch.randn(*shape) * sigma` (sde.py:26-28) on the CPU generator;
  * step noise  : `torch.randn_like` is patched for the duration of the call to pop the supplied
                  z1, z2 tensors in the order cond_pc_sampler draws them (samplers.py:131,149).
"""
import contextlib
from typing import Dict, Optional

import numpy as np
import torch

from oracle import ref_loader


@contextlib.contextmanager
def _inject_randn_like(step_noise: Optional[torch.Tensor]):
    if step_noise is None:
        yield
        return
    queue = [step_noise[i, j] for i in range(step_noise.shape[0]) for j in range(2)]
    queue.reverse()
    orig = torch.randn_like

    def fake(x, *a, **k):
        z = queue.pop()
        assert z.shape == x.shape, (z.shape, x.shape)
        return z.to(x.dtype).clone()

    torch.randn_like = fake
    try:
        yield
    finally:
        torch.randn_like = orig
    assert not queue, f"{len(queue)} injected noise tensors were not consumed"


def run_reference(score_sd, clouds: np.ndarray, repeat_num: int, sampler: str, *,
                  num_steps: Optional[int] = None, T0: Optional[float] = None,
                  x0: Optional[np.ndarray] = None, step_noise: Optional[np.ndarray] = None,
                  energy_sd=None, want_score_probe: bool = True) -> Dict[str, np.ndarray]:
    """One full pass of the reference agent on CPU: pts_feature -> pred_func -> (get_energy ->
    sort_poses_by_energy).  Returns numpy arrays."""
    from genpose_b200 import synth

    out: Dict[str, np.ndarray] = {}
    argv = ref_loader.default_argv(sampler, num_steps, "score")
    with ref_loader.reference_env(argv):
        agent, cfg = ref_loader.build_agent(score_sd, argv)
        data = synth.batch_from_clouds(clouds)
        B = clouds.shape[0]
        with torch.no_grad():
            pts_feat = agent.net(data, mode="pts_feature")
        out["pts_feat"] = pts_feat.numpy().copy()
        if want_score_probe and x0 is not None:
            # one bare score evaluation at t = 0.7 (GFObjectPose.forward mode 'score', posenet.py:160)
            rows = x0.shape[0]
            probe = {"pts_feat": pts_feat.unsqueeze(1).repeat(1, repeat_num, 1).view(rows, -1),
                     "sampled_pose": torch.from_numpy(x0).clone() * 0.02,
                     "t": torch.ones(rows, 1) * 0.7}
            with torch.no_grad():
                out["score_probe"] = agent.net(probe, mode="score").numpy().copy()
        if x0 is not None:
            x0_t = torch.from_numpy(x0).clone()
            agent.net.prior_fn = lambda shape, T=1.0: x0_t.clone()
        sn = None if step_noise is None else torch.from_numpy(step_noise)
        with _inject_randn_like(sn):
            res = agent.pred_func(dict(data), repeat_num=repeat_num, save_path=None, T0=T0, return_process=True)
        pred_pose, process = res
        out["pred_pose"] = pred_pose.numpy().copy()
        out["process_last"] = process[:, :, -1].numpy().copy()
    if energy_sd is not None:
        argv = ref_loader.default_argv(sampler, num_steps, "energy")
        with ref_loader.reference_env(argv):
            agent, cfg = ref_loader.build_agent(energy_sd, argv)
            from networks.reward import sort_poses_by_energy
            data = synth.batch_from_clouds(clouds)
            pose_t = torch.from_numpy(out["pred_pose"]).float()
            energy = agent.get_energy(data=data, pose_samples=pose_t, T=1e-5)
            out["energy"] = energy.numpy().copy()
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                sp, se = sort_poses_by_energy(pose_t, energy)
            out["sorted_pose"] = sp.numpy().copy()
            out["sorted_energy"] = se.numpy().copy()
            # pooled pose: sort_sRT_by_energy(ratio=0.6,'average') needs .cuda() (sgpa_utils.py:939); its
            # arithmetic is average_quaternion_batch + pytorch3d conversions, executed here on CPU.
            from utils.misc import average_quaternion_batch, get_rot_matrix
            import pytorch3d.transforms as p3d
            K = pose_t.shape[1]
            keep = max(1, int(K * 0.6))
            sel = sp[:, :keep]
            Bn = sel.shape[0]
            R = get_rot_matrix(sel[..., :6].reshape(Bn * keep, 6), "rot_matrix")
            q = p3d.matrix_to_quaternion(R).reshape(Bn, keep, 4)
            q_avg = average_quaternion_batch(q)
            RT = np.identity(4)[np.newaxis, ...].repeat(Bn, 0)
            RT[:, :3, :3] = p3d.quaternion_to_matrix(q_avg).numpy()
            RT[:, :3, 3] = torch.mean(sel[..., 6:9], dim=1).numpy()
            out["pooled_RT"] = RT.astype(np.float32)
    return out
