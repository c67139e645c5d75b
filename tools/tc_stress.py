"""Stress the tensor-core PC sampler at small grids (1-3 tile teams), where a step's tail is shortest and inter-warp skew largest:
many repetitions, tc vs fp32 FFMA kernel on the same explicit noise, allocator memory dirtied with NaN bit patterns in between.
    python tools/tc_stress.py [reps] [precision] [team]   exits non-zero on the first disagreement; run under `timeout` (a hang is a failure)
team = 0 (chosen from the row count: 4 at these sizes), 1, 2 or 4: the tile-team size to force."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from genpose_b200 import ops, synth  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 50
PRECISION = sys.argv[2] if len(sys.argv) > 2 else "f16x2"
TEAM = int(sys.argv[3]) if len(sys.argv) > 3 else 0
SHAPES = [(3, 64, 100, True), (2, 50, 30, True), (5, 50, 60, False), (7, 50, 40, False), (1, 128, 25, True)]
t0 = time.time()
worst = 0.0
for (B, K, T, proc) in SHAPES:
    seed = 50 + B
    sd = synth.make_state_dict(seed, kappa=synth.stable_kappa(T))
    eng = ops.Engine(sd)
    clouds = torch.from_numpy(synth.make_clouds(B, seed)).cuda()
    ob = eng.object_bias(eng.encode(clouds))
    cen = clouds.mean(dim=1).contiguous()
    x0 = torch.from_numpy(synth.make_prior_noise(B * K, seed)).cuda()
    sn = torch.from_numpy(synth.make_step_noise(T, B * K, seed)).cuda()
    ref = eng.sample_pc(ob, cen, x0, K, T, step_noise=sn, precision="fp32")
    torch.cuda.synchronize()
    first = None
    for i in range(reps):
        junk = torch.full((8 << 20,), float("nan"), device="cuda")     # dirty what the caching allocator hands out next
        del junk
        eng._ws.clear()                                                   # fresh (dirty) workspace every repetition
        out = eng.sample_pc(ob, cen, x0, K, T, step_noise=sn, precision=PRECISION, return_process=proc, team=TEAM)
        pose = out[0] if proc else out
        torch.cuda.synchronize()
        d = float((pose - ref).abs().max())
        worst = max(worst, d)
        if not (d <= 1e-3) or not bool(torch.isfinite(pose).all()):
            print(f"FAIL shape {(B, K, T)} rep {i}: max|tc - fp32| = {d:.3e}")
            sys.exit(1)
        if first is None:
            first = pose.clone()
        elif not torch.equal(first, pose):
            print(f"FAIL shape {(B, K, T)} rep {i}: run-to-run difference {float((first - pose).abs().max()):.3e}")
            sys.exit(1)
    print(f"shape B={B} K={K} T={T} process={proc}: {reps} repetitions bit-identical, max|tc - fp32| {worst:.3e}", flush=True)
# the ODE sampler at the same small grids: bit-identical repetitions, same accept / reject sequence as the FFMA kernel
from genpose_b200.sde import init_sde  # noqa: E402
ve_prior = init_sde("ve")[0]
for (B, K) in [(3, 64), (1, 128), (5, 50)]:
    sd = synth.make_state_dict(70 + B, kappa=0.3)
    eng = ops.Engine(sd)
    clouds = torch.from_numpy(synth.make_clouds(B, 70 + B)).cuda()
    ob = eng.object_bias(eng.encode(clouds))
    cen = clouds.mean(dim=1).contiguous()
    torch.manual_seed(B)
    x0 = ve_prior((B * K, 9), T=0.55).cuda().contiguous()
    ref, sref = eng.sample_ode(ob, cen, x0, K, T0=0.55, precision="fp32")
    first = None
    for i in range(max(2, reps // 4)):
        eng._ws.clear()
        pose, st = eng.sample_ode(ob, cen, x0, K, T0=0.55, precision=PRECISION, team=TEAM)
        torch.cuda.synchronize()
        if first is None:
            first = pose.clone()
        if not torch.equal(first, pose) or int(st[3]) != 0 or abs(int(st[0]) - int(sref[0])) > 12 or not bool(((pose - ref).abs() <= 1e-3 + 2e-4 * ref.abs()).all()):
            print(f"FAIL ode shape {(B, K)} rep {i}: stats {st.tolist()} vs {sref.tolist()}, max diff {float((pose - ref).abs().max()):.3e}")
            sys.exit(1)
    print(f"ode shape B={B} K={K}: {max(2, reps // 4)} repetitions bit-identical, nfev {int(st[0])} (fp32 kernel {int(sref[0])})", flush=True)
print(f"ok ({time.time() - t0:.1f} s)")
