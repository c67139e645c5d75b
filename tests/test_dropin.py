"""The drop-in boundary: (CPU) with genpose_b200/dropin ahead of a reference checkout, the runner's imports resolve to
our modules for exactly the replaced names and to the reference for the rest; (GPU) the `pointnet2_cuda` twin obeys the
reference extension's calling convention (caller-allocated tensors written in place, returns 1)."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("GENPOSE_REFERENCE_ROOT", "/root/reference")


@pytest.mark.reference
@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "networks")), reason="reference tree not present")
def test_namespace_resolution_against_reference_checkout():
    # Mimics `python runners/evaluation_single.py`: sys.path[0] is the script's directory (runners/), the checkout root is
    # APPENDED by the runner itself (evaluation_single.py:16), so PYTHONPATH entries are searched first.
    code = r"""
import sys
sys.argv = ['evaluation_single.py', '--sampler_mode', 'ode', '--T0', '0.55']
sys.path.append(%r)
""" % REF + r"""
import networks.posenet_agent, networks.reward, configs.config, networks.posenet, pointnet2_cuda
import utils.genpose_utils                      # NOT replaced: must come from the reference checkout
print(networks.posenet_agent.__file__); print(networks.reward.__file__); print(configs.config.__file__)
print(networks.posenet.__file__); print(pointnet2_cuda.__file__); print(utils.genpose_utils.__file__)
from networks.posenet_agent import PoseNet
from networks.reward import sort_poses_by_energy, ranking_loss
cfg = configs.config.get_config()
assert cfg.sampler_mode == ['ode'] and cfg.T0 == 0.55 and cfg.eval_repeat_num == 50 and cfg.batch_size == 192
agent = PoseNet(cfg)
assert agent.net.__class__.__name__ == 'GFObjectPose' and agent.T == 1.0
"""
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([os.path.join(ROOT, "genpose_b200", "dropin"), ROOT, os.path.join(ROOT, "oracle", "shims")])
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, cwd=os.path.join(REF, "runners"))
    assert out.returncode == 0, out.stderr[-2000:]
    lines = out.stdout.strip().splitlines()
    assert all("genpose_b200" in l for l in lines[:5]), lines
    assert lines[5].startswith(REF), lines


def test_agent_contract_without_gpu():
    """Signature / checkpoint-schema level checks that need no device."""
    import inspect
    from genpose_b200 import synth
    from genpose_b200.config import get_config
    from genpose_b200.posenet import GFObjectPose
    from genpose_b200.posenet_agent import PoseNet
    from genpose_b200.sde import init_sde

    assert list(inspect.signature(PoseNet.pred_func).parameters)[1:] == [
        "data", "repeat_num", "save_path", "return_average_res", "init_x", "T0", "return_process"]
    assert list(inspect.signature(PoseNet.get_energy).parameters)[1:] == ["data", "pose_samples", "T", "mode", "extract_pts_feature"]
    assert list(inspect.signature(PoseNet.load_ckpt).parameters)[1:] == ["name", "model_dir", "model_path", "load_model_only"]
    assert list(inspect.signature(GFObjectPose.forward).parameters)[1:] == ["data", "mode", "init_x", "T0"]
    cfg = get_config(["--sampler_mode", "pc", "--sampling_steps", "10"])
    net = GFObjectPose(cfg, *init_sde("ve"))
    sd = synth.make_state_dict(0)
    assert set(net.expected_keys()) == set(sd.keys())               # the reference's state_dict schema, key for key
    with pytest.raises(RuntimeError):
        net.load_state_dict({k: v for k, v in list(sd.items())[:-1]})   # strict like posenet_agent.py:166-168
    prior, marg, sde, eps, T = init_sde("ve")
    torch.manual_seed(0)
    a = prior((4, 9), T=0.55)
    torch.manual_seed(0)
    assert torch.equal(a, torch.randn(4, 9) * (0.01 * (50 / 0.01) ** 0.55)) and eps == 1e-5 and T == 1.0


@pytest.mark.gpu
def test_pointnet2_cuda_twin_convention():
    from genpose_b200 import pointnet2_cuda as pc
    from oracle import genpose_oracle as O
    g = torch.Generator().manual_seed(0)
    xyz = torch.randn(2, 256, 3, generator=g) * 0.1
    d = xyz.cuda()
    idx = torch.cuda.IntTensor(2, 64)                        # pointnet2_utils.py:26-27
    temp = torch.cuda.FloatTensor(2, 256).fill_(1e10)
    assert pc.furthest_point_sampling_wrapper(2, 256, 64, d, temp, idx) == 1
    assert torch.equal(idx.cpu(), O.furthest_point_sample(xyz, 64))
    new_xyz = torch.gather(d, 1, idx.long().unsqueeze(-1).expand(-1, -1, 3)).contiguous()
    bidx = torch.cuda.IntTensor(2, 64, 16).zero_()           # pointnet2_utils.py:219
    assert pc.ball_query_wrapper(2, 256, 64, 0.08, 16, new_xyz, d, bidx) == 1
    assert torch.equal(bidx.cpu(), O.ball_query(0.08, 16, xyz, new_xyz.cpu()))
    feats = torch.randn(2, 5, 256, generator=g).cuda()
    out = torch.cuda.FloatTensor(2, 5, 64, 16)
    assert pc.group_points_wrapper(2, 5, 256, 64, 16, feats, bidx, out) == 1
    ref = torch.stack([feats[b][:, bidx[b].long()] for b in range(2)])
    assert torch.equal(out, ref)
    with pytest.raises(NotImplementedError):
        pc.three_nn_wrapper()


@pytest.mark.gpu
def test_agent_end_to_end_on_gpu(tmp_path):
    """PoseNet(cfg) -> load_ckpt(file) -> pred_func -> get_energy -> sort_poses_by_energy, as evaluation_single.py drives it."""
    from genpose_b200 import synth
    from genpose_b200.config import get_config
    from genpose_b200.posenet_agent import PoseNet
    from genpose_b200.reward import sort_poses_by_energy
    from oracle import genpose_oracle as O
    T, B, K = 40, 2, 50
    sd = synth.make_state_dict(3, kappa=synth.stable_kappa(T))
    esd = synth.make_state_dict(103, kappa=synth.stable_kappa(T))
    ck = tmp_path / "ckpt_genpose.pth"
    torch.save({"model_state_dict": sd, "clock": {}, "optimizer_state_dict": {}, "scheduler_state_dict": {}}, ck)
    cfg = get_config(["--sampler_mode", "pc", "--sampling_steps", str(T), "--noise_mode", "torch", "--precision", "fp32"])
    agent = PoseNet(cfg)
    agent.load_ckpt(model_dir=str(ck), model_path=True, load_model_only=True)       # evaluation_single.py:362
    data = synth.batch_from_clouds(synth.make_clouds(B, 3), device="cuda")
    torch.manual_seed(5)
    pose = agent.pred_func(data=data, repeat_num=K, save_path=None, T0=1.0)
    assert pose.shape == (B, K, 9) and "pts_feat" in data
    # same seeds through the oracle: prior from the CPU generator, z1/z2 from the CUDA generator in the reference's order
    torch.manual_seed(5)
    x0 = torch.randn(B * K, 9) * 50.0
    like = torch.empty(B * K, 9, device="cuda")
    sn = torch.stack([torch.stack([torch.randn_like(like), torch.randn_like(like)]) for _ in range(T)]).cpu()
    # the oracle samples from the features the agent left in data['pts_feat'] (posenet_agent.py:422); see O.pred_func_pc
    ref, _ = O.pred_func_pc(sd, synth.batch_from_clouds(synth.make_clouds(B, 3)), K, T, x0, sn, pts_feat=data["pts_feat"].cpu())
    assert float((pose.cpu() - ref).abs().max()) < 1e-3
    cfg_e = get_config(["--sampler_mode", "pc", "--sampling_steps", str(T), "--posenet_mode", "energy"])
    eagent = PoseNet(cfg_e)
    eagent.net.load_state_dict(esd)
    en = eagent.get_energy(data=data, pose_samples=pose, T=1e-5)
    en_ref = O.get_energy(esd, synth.batch_from_clouds(synth.make_clouds(B, 3)), pose.cpu())
    assert torch.allclose(en.cpu(), en_ref, rtol=2e-4, atol=1e-2)
    sp, se = sort_poses_by_energy(pose, en)
    sp_ref, se_ref = O.sort_poses_by_energy(pose.cpu(), en.cpu())
    assert torch.equal(sp.cpu(), sp_ref) and torch.equal(se.cpu(), se_ref)


@pytest.mark.gpu
def test_pred_func_average_branch_on_gpu():
    """pred_func(return_average_res=True) (posenet_agent.py:449-461): per-candidate quaternions (get_rot_matrix + pytorch3d
    matrix_to_quaternion), the eigen-mean quaternion of all K candidates oriented to w > 0 (average_quaternion_batch,
    utils/misc.py:227-249) and the mean translation — against the oracle's restatement of those functions on the SAME poses."""
    from genpose_b200 import synth
    from genpose_b200.config import get_config
    from genpose_b200.posenet_agent import PoseNet
    from oracle import genpose_oracle as O
    T, B, K = 30, 3, 50
    sd = synth.make_state_dict(9, kappa=synth.stable_kappa(T))
    agent = PoseNet(get_config(["--sampler_mode", "pc", "--sampling_steps", str(T)]))
    agent.net.load_state_dict(sd)
    data = synth.batch_from_clouds(synth.make_clouds(B, 9), device="cuda")
    torch.manual_seed(2)
    pred_pose, pred_q, average = agent.pred_func(data=data, repeat_num=K, save_path=None, return_average_res=True)
    torch.manual_seed(2)
    four = agent.pred_func(data=data, repeat_num=K, save_path=None, return_average_res=True, return_process=True)
    assert len(four) == 4 and torch.equal(four[0], pred_pose) and four[3].shape[:2] == (B, K)
    assert pred_pose.shape == (B, K, 9) and pred_q.shape == (B, K, 7) and average.shape == (B, 7)
    res = pred_pose.reshape(B * K, 9).cpu()
    q_ref = O.matrix_to_quaternion(O.get_rot_matrix(res[:, :6]))
    ref_q = torch.cat([q_ref, res[:, 6:]], dim=-1).reshape(B, K, 7)
    assert float((pred_q.cpu() - ref_q).abs().max()) <= 1e-5
    avg_q = O.average_quaternion_batch(ref_q[:, :, :4])
    assert bool((average[:, 0] > 0).all())
    assert float((average[:, :4].cpu() - avg_q).abs().max()) <= 2e-5
    assert float((average[:, 4:].cpu() - ref_q[:, :, 4:].mean(dim=1)).abs().max()) <= 1e-5 * max(1.0, float(ref_q[:, :, 4:].abs().max()))


@pytest.mark.gpu
def test_sort_poses_by_energy_keeps_float64():
    """reward.sort_poses_by_energy on the ODE sampler's float64 poses: the reference gathers them without a cast
    (reward.py:145-152), so must we — bit-identical rows, float64 out, same order as torch.sort(descending) on distinct energies."""
    from genpose_b200.reward import sort_poses_by_energy
    from oracle import genpose_oracle as O
    g = torch.Generator().manual_seed(3)
    poses = torch.randn(4, 50, 9, generator=g, dtype=torch.float64)
    energy = torch.randn(4, 50, 2, generator=g)
    sp, se = sort_poses_by_energy(poses.cuda(), energy.cuda())
    sp_ref, se_ref = O.sort_poses_by_energy(poses, energy)
    assert sp.dtype == torch.float64 and torch.equal(sp.cpu(), sp_ref) and torch.equal(se.cpu(), se_ref)
    sp32, _ = sort_poses_by_energy(poses.float().cuda(), energy.cuda())
    assert sp32.dtype == torch.float32 and torch.equal(sp32.cpu(), sp_ref.float())


def test_energy_agent_refuses_score_and_sampling():
    """An energy network's score is the gradient of its energy (energynet.py:187-198), not the trunk output: the score / sampling
    entry points raise instead of silently returning f / std."""
    from genpose_b200.config import get_config
    from genpose_b200.posenet import GFObjectPose
    from genpose_b200.sde import init_sde
    cfg = get_config(["--sampler_mode", "pc", "--sampling_steps", "10", "--posenet_mode", "energy"])
    net = GFObjectPose(cfg, *init_sde(cfg.sde_mode))
    for call in (lambda: net.sample_candidates(None, None, 5, "pc"), lambda: net._require_score_net("forward(mode='score')")):
        with pytest.raises(NotImplementedError):
            call()


def test_poses_to_RTs_is_the_runners_loop():
    """pipeline.poses_to_RTs against the literal per-candidate loop of pred_pose_batch (runners/evaluation_single.py:325-332)."""
    import numpy as np
    from genpose_b200.pipeline import poses_to_RTs
    from oracle import genpose_oracle as O
    g = torch.Generator().manual_seed(4)
    pred_pose = torch.randn(3, 7, 9, generator=g)
    RTs_all = np.ones((3, 7, 4, 4))
    for i in range(pred_pose.shape[1]):
        R = O.get_rot_matrix(pred_pose[:, i, :-3])
        T = pred_pose[:, i, -3:]
        RTs = np.identity(4, dtype=float)[np.newaxis, ...].repeat(R.shape[0], 0)
        RTs[:, :3, :3] = R.cpu().numpy()
        RTs[:, :3, 3] = T.cpu().numpy()
        RTs_all[:, i, :, :] = RTs
    got = poses_to_RTs(pred_pose)
    assert got.dtype == np.float64 and got.shape == (3, 7, 4, 4)
    np.testing.assert_allclose(got, RTs_all, rtol=0, atol=1e-6)
