"""Time the ODE sampler (shipped recipe: T0=0.55, K=50, rtol=atol=1e-5) at the bench shape and print nfev.
    python tools/ode_timing.py [precision] [B] [K]"""
import sys

import torch

sys.path.insert(0, ".")
from genpose_b200 import ops, synth  # noqa: E402
from genpose_b200.sde import init_sde  # noqa: E402

ve_prior = init_sde("ve")[0]               # sigma_max = 50 (sde.py:90-97)

precision = sys.argv[1] if len(sys.argv) > 1 else "fp32"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
K = int(sys.argv[3]) if len(sys.argv) > 3 else 50
T0 = 0.55
sd = synth.make_state_dict(0, kappa=-0.3)
eng = ops.Engine(sd)
pts = torch.from_numpy(synth.make_clouds(B, 100)).cuda()
center = pts.mean(dim=1).contiguous()
R = B * K
torch.manual_seed(0)
x0 = ve_prior((R, 9), T=T0).cuda().contiguous()
ob = eng.object_bias(eng.encode(pts))
kw = {} if precision == "fp32" else {"precision": precision}
for _ in range(2):
    pose, stats = eng.sample_ode(ob, center, x0, K, T0=T0, **kw)
torch.cuda.synchronize()
ms = []
for _ in range(5):
    a, b = torch.cuda.Event(True), torch.cuda.Event(True)
    a.record()
    pose, stats = eng.sample_ode(ob, center, x0, K, T0=T0, **kw)
    b.record()
    torch.cuda.synchronize()
    ms.append(a.elapsed_time(b))
st = stats.cpu().tolist()
print(f"ode sampler [{precision}] R={R} T0={T0}: {min(ms):.3f} ms (median {sorted(ms)[2]:.3f}); stats nfev/accepted/rejected/status = {st}; "
      f"{1e3 * min(ms) / max(1, st[0]):.1f} us per evaluation; pose[0] = {pose[0].cpu().numpy().round(5).tolist()}")
