"""Short, deterministic workloads for ncu (never used for bench numbers): `sampler` = 2 PC-sampler launches of
the bench shape (3200 rows, T=50 to keep replays short... T=500 for the real shape), `encoder` = 2 encoder passes."""
import sys

import torch

sys.path.insert(0, ".")
from genpose_b200 import ops, synth  # noqa: E402

what = sys.argv[1]
T = int(sys.argv[2]) if len(sys.argv) > 2 else 500
B, K = 64, 50
sd = synth.make_state_dict(0, kappa=-0.3)
eng = ops.Engine(sd)
pts = torch.from_numpy(synth.make_clouds(B, 100)).cuda()
center = pts.mean(dim=1).contiguous()
x0 = torch.from_numpy(synth.make_prior_noise(B * K, 100)).cuda()
for i in range(2):
    feat = eng.encode(pts)
    ob = eng.object_bias(feat)
    if what == "sampler":
        eng.sample_pc(ob, center, x0, K, T, seed=i)
    if what == "tc_sampler":
        eng.sample_pc(ob, center, x0, K, T, seed=i, precision="bf16x3")
    if what == "tc_ode":
        from genpose_b200.sde import init_sde
        ve_prior = init_sde("ve")[0]               # sigma_max = 50 (sde.py:90-97)
        torch.manual_seed(0)
        eng.sample_ode(ob, center, ve_prior((B * K, 9), T=0.55).cuda().contiguous(), K, T0=0.55, precision="bf16x3")
torch.cuda.synchronize()
print("done", what)
