"""Short, deterministic workloads for ncu (never used for bench numbers).
    python tools/profile_target.py <what> [T] [precision] [objects] [shape.json]
what: sampler (FFMA PC sampler) | tc_sampler (tensor-core PC sampler) | tc_ode (tensor-core ODE sampler) | encoder |
      energy_rank_pool (energy pass + rank + pool of config 3).
Two passes of the chosen workload at the bench shape by default (64 objects x 50 candidates, T = 500).  When a path is given as the
fifth argument the shape of the profiled launch is written there for tools/summarize_ncu.py (-> profiles/*.meta.json)."""
import json
import sys

import torch

sys.path.insert(0, ".")
from genpose_b200 import ops, synth  # noqa: E402

what = sys.argv[1]
T = int(sys.argv[2]) if len(sys.argv) > 2 else 500
PRECISION = sys.argv[3] if len(sys.argv) > 3 else ops.AUTO_TC_PRECISION
B = int(sys.argv[4]) if len(sys.argv) > 4 else 64
K = 50
sd = synth.make_state_dict(0, kappa=-0.3)
eng = ops.Engine(sd)
eeng = ops.Engine(synth.make_state_dict(100, kappa=-0.3)) if what == "energy_rank_pool" else None
pts = torch.from_numpy(synth.make_clouds(B, 100)).cuda()
center = pts.mean(dim=1).contiguous()
x0 = torch.from_numpy(synth.make_prior_noise(B * K, 100)).cuda()
for i in range(2):
    feat = eng.encode(pts)
    ob = eng.object_bias(feat)
    if what == "sampler":
        eng.sample_pc(ob, center, x0, K, T, seed=i, precision="fp32")
    if what == "tc_sampler":
        eng.sample_pc(ob, center, x0, K, T, seed=i, precision=PRECISION)
    if what == "tc_ode":
        from genpose_b200.sde import init_sde
        ve_prior = init_sde("ve")[0]               # sigma_max = 50 (sde.py:90-97)
        torch.manual_seed(0)
        eng.sample_ode(ob, center, ve_prior((B * K, 9), T=0.55).cuda().contiguous(), K, T0=0.55, precision=PRECISION)
    if what == "energy_rank_pool":
        pose = eng.sample_pc(ob, center, x0, K, 4, seed=i, precision="fp32")
        eob = eeng.object_bias(eeng.encode(pts))
        en = eeng.energy(eob, center, pose, K, 1e-5)
        ops.rank_pool(pose.view(B, K, 9), en.view(B, K, 2))
torch.cuda.synchronize()
if len(sys.argv) > 5:
    json.dump({"what": what, "precision": PRECISION, "objects": B, "rows": B * K, "steps": T}, open(sys.argv[5], "w"))
print("done", what)
