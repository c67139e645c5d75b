"""Executed in a subprocess by tests/test_runner_flow_reference.py (build container only: needs the reference checkout).

Runs the UNMODIFIED reference runner — runners/evaluation_single.py: unpack_data, inference_pose (pred_pose_batch, :311-334, :356-424)
and inference_energy (pred_energy_batch, :337-353, :427-489) — over a synthetic detection pickle, with genpose_b200's agent dropped in by
PYTHONPATH exactly as INTEGRATION.md describes.  There is no GPU in the build container, so the C library is replaced by a stand-in
AT THE C ABI: every gpb_* compute entry point the agent calls is a small numpy function with a known answer over the raw pointers it
receives (sizes and packing queries go to the real libgenpose_b200.so).  Everything above the ABI is the shipped code: dropin modules,
PoseNet.load_ckpt / pred_func / get_energy, GFObjectPose, ops.Engine (weight packing included), reward.sort_poses_by_energy.
What this proves: the runner's control flow (category batching, the pickle hand-off between the two agents, RT assembly) runs unchanged on
our surface and every tensor it gets has the shape / dtype / ordering it expects.  Arithmetic parity is the GPU tests' job."""
import ctypes
import os
import pickle
import sys

import numpy as np
import torch

REF, ROOT, WORK = sys.argv[1], sys.argv[2], sys.argv[3]
K, BATCH = 50, 3
os.chdir(WORK)
sys.argv = ["evaluation_single.py", "--score_model_dir", "score.pth", "--energy_model_dir", "energy.pth", "--data_path", WORK,
            "--sampler_mode", "ode", "--eval_repeat_num", str(K), "--batch_size", str(BATCH), "--T0", "0.55",
            "--result_dir", os.path.join(WORK, "results"), "--device", "cpu", "--test_source", "real_test", "--pooling_mode", "average"]
sys.path[:0] = [os.path.join(ROOT, "genpose_b200", "dropin"), ROOT, os.path.join(ROOT, "oracle", "shims")]
torch.cuda.FloatTensor = torch.FloatTensor          # the runner allocates through these (evaluation_single.py:386, :458)
torch.cuda.IntTensor = torch.IntTensor

from genpose_b200 import lib, ops, synth  # noqa: E402

real = lib.load()
calls = {"encode": 0, "sample": 0, "energy": 0, "rank_pool": 0, "poses": []}
_CT = {np.float32: ctypes.c_float, np.float64: ctypes.c_double, np.int32: ctypes.c_int}


def arr(ptr, shape, dtype=np.float32):
    n = int(np.prod(shape))
    return np.ctypeslib.as_array((_CT[dtype] * n).from_address(ptr)).reshape(shape)


def gram_schmidt(v):
    b1 = v[:, 0:3] / np.maximum(np.linalg.norm(v[:, 0:3], axis=1, keepdims=True), 1e-12)
    c = v[:, 3:6] - (b1 * v[:, 3:6]).sum(1, keepdims=True) * b1
    return np.concatenate([b1, c / np.maximum(np.linalg.norm(c, axis=1, keepdims=True), 1e-12)], axis=1)


class StandIn:
    """gpb_* compute entry points with known answers; everything else is the real library."""

    def __getattr__(self, name):
        return getattr(real, name)

    def gpb_sampler_tc_max_rows(self, K_):
        return 0                                       # no device: 'auto' resolves to the plain entry points

    def gpb_set_tc_team(self, team):
        return 0

    def _encode(self, pts, B, feat):
        p = arr(pts, (B, 1024, 3))
        arr(feat, (B, 1024))[:] = np.tile(p.mean(axis=1), (1, 342))[:, :1024]
        calls["encode"] += 1
        return 0

    def gpb_encode(self, pts, B, enc_w, feat, ws, wsb, f1, f2, f3, stream):
        return self._encode(pts, B, feat)

    def gpb_encode_tc(self, pts, B, enc_w, enc_tc, feat, ws, wsb, f1, f2, f3, stream):
        return self._encode(pts, B, feat)

    def gpb_object_bias(self, feat, B, W, out, stream):
        arr(out, (B, 768))[:] = arr(feat, (B, 1024))[:, :768]
        return 0

    def gpb_sample_ode(self, x0, R, K_, T0, rtol, atol, den, ob, W, center, pose, stats, process, cap, t_eval, n_eval, ws, wsb, stream):
        assert abs(T0 - 0.55) < 1e-12 and rtol == 1e-5 and atol == 1e-5 and den == 1000 and K_ == K, (T0, rtol, atol, den, K_)
        x = arr(x0, (R, 9)).astype(np.float64)
        c = np.repeat(arr(center, (R // K_, 3)).astype(np.float64), K_, axis=0)
        out = np.concatenate([gram_schmidt(x), x[:, 6:9] + c], axis=1)
        arr(pose, (R, 9), np.float64)[:] = out
        arr(stats, (4,), np.int32)[:] = [99, 16, 0, 0]
        if process:
            assert cap >= 17 and not t_eval
            arr(process, (cap, R, 9), np.float64)[:17] = out[None]
        calls["sample"] += 1
        calls["poses"].append(out.reshape(R // K_, K_, 9).copy())
        return 0

    def gpb_energy(self, pose, R, K_, t, ob, W, center, energy, stream):
        assert abs(t - 1e-5) < 1e-12
        p = arr(pose, (R, 9))
        arr(energy, (R, 2))[:] = np.stack([p[:, 0] + 2 * p[:, 4], p[:, 8] - p[:, 6]], axis=1)
        calls["energy"] += 1
        return 0

    def gpb_rank_pool(self, pose, energy, B, K_, keep, sp, se, rt, order, stream):
        p, e = arr(pose, (B, K_, 9)), arr(energy, (B, K_, 2))
        for b in range(B):
            o_r = np.argsort(-e[b, :, 0], kind="stable")
            o_t = np.argsort(-e[b, :, 1], kind="stable")
            if sp:
                arr(sp, (B, K_, 9))[b] = np.concatenate([p[b, o_r, :6], p[b, o_t, 6:]], axis=1)
            if se:
                arr(se, (B, K_, 2))[b] = np.stack([e[b, o_r, 0], e[b, o_t, 1]], axis=1)
        assert not rt and not order
        calls["rank_pool"] += 1
        return 0


lib._lib = StandIn()


def _chk_cpu(t, dtype, name):
    assert isinstance(t, torch.Tensor) and t.dtype == dtype and t.is_contiguous(), (name, t.dtype, dtype)
    return t.data_ptr()


ops._chk = _chk_cpu
ops._stream = lambda: 0
sys.path.append(REF)

# ---- checkpoints where the runner looks for them (evaluation_single.py:32-33) and the detection pickle it unpacks (:263-303) ----
os.makedirs(os.path.join(WORK, "results", "ckpts"), exist_ok=True)
torch.save({"model_state_dict": synth.make_state_dict(1, kappa=0.3)}, os.path.join(WORK, "results", "ckpts", "score.pth"))
torch.save({"model_state_dict": synth.make_state_dict(101, kappa=0.3)}, os.path.join(WORK, "results", "ckpts", "energy.pth"))

import runners.evaluation_single as ES  # noqa: E402  (module-level code parses argv and creates the result directories)
import networks.posenet_agent  # noqa: E402

assert networks.posenet_agent.__file__.startswith(os.path.join(ROOT, "genpose_b200", "dropin")), networks.posenet_agent.__file__
assert ES.PoseNet is networks.posenet_agent.PoseNet

clouds = synth.make_clouds(7, 5)
# three frames, 7 valid instances: category 0 (bottle) gets exactly BATCH of them (index[-1] == num edge, :376-377), category 5 (mug) four
frames = {"f0": [(0, 0), (2, 5)], "f1": [(0, 0), (1, 5), (3, 5)], "f2": [(1, 0), (0, 5)]}      # (instance slot, category id)
detect, k = {}, 0
for name, insts in frames.items():
    n_det = max(i for i, _ in insts) + 1
    detect[name] = {"result": {"pred_RTs": np.identity(4)[None].repeat(n_det, 0), "pred_class_ids": np.zeros(n_det, dtype=np.int32)},
                    "valid_pts": [], "valid_rgb": None, "cat_id": [], "valid_inst": []}
    for slot, cat in insts:
        detect[name]["valid_pts"].append(clouds[k])
        detect[name]["cat_id"].append(cat)
        detect[name]["valid_inst"].append(slot)
        k += 1
with open(ES.segmentation_results_path, "wb") as f:
    pickle.dump(detect, f)

torch.manual_seed(0)
ES.inference_pose(ES.segmentation_results_path, ES.inference_res_dir, ES.cfg.pose_mode, record_process=False)
assert calls["encode"] == 3 and calls["sample"] == 3, calls            # bottle: one batch of 3; mug: batches of 3 + 1
ES.inference_energy(ES.inference_res_dir, ES.cfg.pose_mode)
assert calls["energy"] == 3 and calls["rank_pool"] == 3 and calls["encode"] == 6, calls   # the energy agent has its own encoder pass (:431)

with open(os.path.join(ES.inference_res_dir, "results_with_energy.pkl"), "rb") as f:
    res = pickle.load(f)
flat = np.concatenate(calls["poses"], axis=0)                                      # [7, K, 9] in category order: bottle x3, mug x4
order = [("f0", 0), ("f1", 0), ("f2", 1), ("f0", 2), ("f1", 1), ("f1", 3), ("f2", 0)]
for (name, slot), pose in zip(order, flat):
    r = res[name]["result"]
    RT, en = r["multi_hypothesis_pred_RTs"][slot], r["energy"][slot]
    assert RT.shape == (K, 4, 4) and RT.dtype == np.float64 and en.shape == (K, 2)
    assert np.all(np.diff(en[:, 0]) <= 0) and np.all(np.diff(en[:, 1]) <= 0)       # sorted descending, column-wise (reward.py:145-152)
    p32 = pose.astype(np.float32)                                                  # the pickle hand-off stores float32 (:458-459)
    e = np.stack([p32[:, 0] + 2 * p32[:, 4], p32[:, 8] - p32[:, 6]], axis=1)
    o_r, o_t = np.argsort(-e[:, 0], kind="stable"), np.argsort(-e[:, 1], kind="stable")
    assert np.allclose(RT[:, :3, 0], p32[o_r, 0:3], atol=1e-6) and np.allclose(RT[:, :3, 1], p32[o_r, 3:6], atol=1e-6)   # columns b1, b2
    assert np.allclose(RT[:, :3, 2], np.cross(p32[o_r, 0:3], p32[o_r, 3:6]), atol=1e-5)
    assert np.allclose(RT[:, :3, 3], p32[o_t, 6:9], atol=1e-6) and np.allclose(RT[:, 3], [0, 0, 0, 1])
    assert np.allclose(en, np.stack([e[o_r, 0], e[o_t, 1]], axis=1), atol=1e-5)
untouched = res["f1"]["result"]["multi_hypothesis_pred_RTs"][2]                    # a detection without a valid cloud keeps identity
assert np.allclose(untouched, np.identity(4)[None].repeat(K, 0))

# the trajectory surface (--save_video / record_process=True): [pred_pose, in_process_sample [bs, K, n, 9]] (posenet_agent.py:436-466)
ES.inference_pose(ES.segmentation_results_path, ES.inference_res_dir, ES.cfg.pose_mode, record_process=True)
with open(os.path.join(ES.inference_res_dir, "cls_data.pkl"), "rb") as f:
    cls = pickle.load(f)
assert len(cls["mug"]["pred_pose_process"]) == 4 and cls["mug"]["pred_pose_process"][0].shape == (K, 17, 9)
print("RUNNER_FLOW_OK", calls["encode"], calls["sample"], calls["energy"], calls["rank_pool"])
