"""Cycle stamps of gpb_sample_ode_tc_dbg (CTA 0, row thread 0): per-phase durations of one evaluation, averaged
separately over the evaluations inside an RK45 attempt and the ones that end an evaluation group (controller decision +
time-bias refresh), plus the MMA warp's side.   python tools/tc_ode_phase_times.py [f16x2|bf16x3]   (bench shape: 64 objects x 50
candidates, T0 = 0.55)"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from genpose_b200 import lib, ops, synth  # noqa: E402
from genpose_b200.sde import init_sde  # noqa: E402

ve_prior = init_sde("ve")[0]               # sigma_max = 50 (sde.py:90-97)

W16 = not (len(sys.argv) > 1 and sys.argv[1] == "bf16x3")
B, K, T0, NE = 64, 50, 0.55, 128
sd = synth.make_state_dict(0, kappa=-0.3)
eng = ops.Engine(sd)
pts = torch.from_numpy(synth.make_clouds(B, 100)).cuda()
center = pts.mean(dim=1).contiguous()
R = B * K
torch.manual_seed(0)
x0 = ve_prior((R, 9), T=T0).cuda().contiguous()
ob = eng.object_bias(eng.encode(pts))
L = lib.load()
ws = torch.empty(L.gpb_sampler_workspace_bytes(R, 1), dtype=torch.uint8, device="cuda")
pose = torch.empty(R, 9, dtype=torch.float64, device="cuda")
stats = torch.zeros(4, dtype=torch.int32, device="cuda")
dbg = torch.zeros(2 * NE * 16 + NE * 8, dtype=torch.int64, device="cuda")   # + the trace block of experiment builds
for _ in range(2):
    fn, stream_w = (L.gpb_sample_ode_tc16_dbg, eng.trunk_tc16()) if W16 else (L.gpb_sample_ode_tc_dbg, eng.trunk_tc)
    lib.check(fn(x0.data_ptr(), R, K, T0, 1e-5, 1e-5, 1000, ob.data_ptr(), eng.trunk_w.data_ptr(),
                                      stream_w.data_ptr(), center.data_ptr(), pose.data_ptr(), stats.data_ptr(), ws.data_ptr(),
                                      ws.numel(), dbg.data_ptr(), NE, torch.cuda.current_stream().cuda_stream), "dbg")
torch.cuda.synchronize()
n = int(stats[0].item())
raw = dbg.cpu().numpy().astype(np.float64)
m = raw[:2 * NE * 16].reshape(2, NE, 16)[1, :min(n, NE)]
d = raw[:2 * NE * 16].reshape(2, NE, 16)[0, :min(n, NE)]
trc = raw[2 * NE * 16:].reshape(NE, 8)[:min(n, NE)]
step = d[1:, 0] - d[:-1, 0]
bnd = d[:-1, 14] > 0
seq = [("wait layer-0 accumulator", 0, 1), ("epilogue layer 0", 1, 2), ("wait layer-1 accumulator", 2, 3), ("epilogue layer 1", 3, 4),
       ("wait head 128", 4, 5), ("epilogue head 128", 5, 6), ("wait head 64", 6, 7), ("epilogue head 64", 7, 8), ("column-half sync", 8, 9),
       ("send partials", 9, 11), ("wait peers' partials", 11, 10), ("RK45 stage math (+ controller)", 10, 12), ("publish x", 12, 13)]
print(f"nfev {n}; evaluations recorded {len(d)}; mean cycles per evaluation {step.mean():.0f} (inside a group {step[~bnd].mean():.0f}, "
      f"group-ending {step[bnd].mean():.0f}; {bnd.sum()} of {len(bnd)} end a group)")
for name, a, b in seq:
    v = d[:-1, b] - d[:-1, a]
    print(f"  {name:36s} inside {v[~bnd].mean():8.0f}   group-ending {v[bnd].mean():8.0f}")
v = d[:-1]
print(f"  {'wait decision barrier':36s} inside {0:8.0f}   group-ending {(v[bnd, 14] - v[bnd, 13]).mean():8.0f}")
print(f"  {'time biases of the next group':36s} inside {0:8.0f}   group-ending {(v[bnd, 15] - v[bnd, 14]).mean():8.0f}")
ins = d[:-1][~bnd]
sub = ins[ins[:, 15] < 0]
print(f"  inside a group, stage math split: f -> k {(-sub[:, 14] - sub[:, 10]).mean():.0f} | K loads + combination {(-sub[:, 15] + sub[:, 14]).mean():.0f} | "
      f"rest (x = fp32(y_stage)) {(sub[:, 12] + sub[:, 15]).mean():.0f}")
mi = m[:-1][~bnd]
print(f"MMA warp, inside a group ({'f16x2' if W16 else 'bf16x3'}): wait x_ready {(mi[:, 1] - mi[:, 0]).mean():.0f} | layer 0 issue -> first h1 quarter in {(mi[:, 3] - mi[:, 1]).mean():.0f} | "
      f"layer 1 span {(mi[:, 4] - mi[:, 3]).mean():.0f} | head slice span {(mi[:, 7] - mi[:, 4]).mean():.0f} | in issue groups {mi[:, 11].mean():.0f}, "
      f"waiting on weight slots {mi[:, 8].mean():.0f}, on the A operand {mi[:, 9].mean():.0f}")
if trc.any():
    tr = trc[:-1][~bnd]
    base = d[:-1][~bnd][:, 0]
    names = ["half 0: step top", "half 0: before pre", "half 0: after pre", "half 0: L1a acc seen", "half 1: step top", "half 1: before pre", "half 1: after pre", "half 1: L1a acc seen"]
    print("trace (cycles after row thread 0's step start, inside a group):")
    for i, nm in enumerate(names):
        print(f"  {nm:24s} {(tr[:, i] - base).mean():9.0f}")
