"""Sampler launch time against batch size, arithmetic and tile-team size (the tensor-core PC and ODE samplers alone, CUDA events).
    python tools/tc_batch_sweep.py [T] [objects,objects,...] [precisions] [teams]
Prints one line per configuration: ms per launch, cycles per step at 1.965 GHz, candidates/s, algorithmic TFLOP/s and its fraction of
the measured sustained bf16 peak (MEASURED_PEAKS.json)."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from genpose_b200 import ops, synth  # noqa: E402
from genpose_b200.sde import init_sde  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 100
OBJECTS = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [64, 128, 256, 378]
PRECISIONS = sys.argv[3].split(",") if len(sys.argv) > 3 else ["bf16x3", "f16x2"]
TEAMS = [int(v) for v in sys.argv[4].split(",")] if len(sys.argv) > 4 else [0]
K = 50
peak = 1358.5
if os.path.exists("MEASURED_PEAKS.json"):
    peak = json.load(open("MEASURED_PEAKS.json")).get("bf16_tflops_sustained", peak)
sd = synth.make_state_dict(0, kappa=-0.3)
eng = ops.Engine(sd)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ve_prior = init_sde("ve")[0]


def timed(fn, n=5, warm=2):
    ms = []
    for i in range(warm + n):
        flush.fill_(i)
        a, b = torch.cuda.Event(True), torch.cuda.Event(True)
        torch.cuda.synchronize()
        a.record()
        fn(i)
        b.record()
        torch.cuda.synchronize()
        if i >= warm:
            ms.append(a.elapsed_time(b))
    return float(np.median(ms))


for B in OBJECTS:
    R = B * K
    pts = torch.from_numpy(synth.make_clouds(B, 100)).cuda()
    cen = pts.mean(dim=1).contiguous()
    ob = eng.object_bias(eng.encode(pts))
    x0 = torch.from_numpy(synth.make_prior_noise(R, 100)).cuda()
    torch.manual_seed(0)
    x0o = ve_prior((R, 9), T=0.55).cuda().contiguous()
    for prec in PRECISIONS:
        for team in TEAMS:
            try:
                ms = timed(lambda i: eng.sample_pc(ob, cen, x0, K, T, seed=i, precision=prec, team=team))
                tf = R * T * 2 * 266752 / (ms / 1e3) / 1e12
                print(f"pc  B={B:4d} R={R:6d} tiles={(R + 127) // 128:4d} {prec:7s} team={team}: {ms:8.3f} ms  {ms / T * 1.965e6:8.0f} cyc/step  "
                      f"{R / (ms / 1e3) * (T / 500):10.0f} cand/s@T=500  {tf:7.1f} TFLOP/s = {tf / peak:.3f} of sustained peak", flush=True)
                st_box = {}

                def ode(i):
                    st_box["s"] = eng.sample_ode(ob, cen, x0o, K, T0=0.55, precision=prec, team=team)[1]
                ms = timed(ode)
                nfev = int(st_box["s"][0])
                tf = R * nfev * 2 * 266752 / (ms / 1e3) / 1e12
                print(f"ode B={B:4d} R={R:6d} tiles={(R + 127) // 128:4d} {prec:7s} team={team}: {ms:8.3f} ms  nfev {nfev}  {ms / nfev * 1.965e6:8.0f} cyc/eval  "
                      f"{R / (ms / 1e3):10.0f} cand/s  {tf:7.1f} TFLOP/s = {tf / peak:.3f}", flush=True)
            except Exception as e:  # noqa: BLE001
                print(f"B={B} {prec} team={team}: {str(e)[:160]}", flush=True)
