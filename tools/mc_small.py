"""compute-sanitizer workload: a short PC and ODE launch per tile-team size (python tools/mc_small.py [teams] [precision])."""
import sys
import torch
sys.path.insert(0, ".")
from genpose_b200 import ops, synth
from genpose_b200.sde import init_sde
teams = [int(t) for t in sys.argv[1].split(",")] if len(sys.argv) > 1 else [1, 2, 4]
prec = sys.argv[2] if len(sys.argv) > 2 else "f16x2"
eng = ops.Engine(synth.make_state_dict(1, kappa=0.3))
B, K = 5, 50
pts = torch.from_numpy(synth.make_clouds(B, 1)).cuda()
cen = pts.mean(dim=1).contiguous()
ob = eng.object_bias(eng.encode(pts))
x0 = torch.from_numpy(synth.make_prior_noise(B * K, 1)).cuda()
torch.manual_seed(0)
x0o = init_sde("ve")[0]((B * K, 9), T=0.15).cuda().contiguous()
for team in teams:
    eng.sample_pc(ob, cen, x0, K, 6, seed=1, precision=prec, team=team)
    torch.cuda.synchronize()
    print("pc team", team, "ok", flush=True)
    eng.sample_ode(ob, cen, x0o, K, T0=0.15, precision=prec, team=team)
    torch.cuda.synchronize()
    print("ode team", team, "ok", flush=True)
