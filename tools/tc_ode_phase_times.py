"""Cycle stamps of gpb_sample_ode_tc_dbg (CTA 0, row thread 0): per-phase durations of one evaluation, averaged
separately over the evaluations inside an RK45 attempt and the ones that end an evaluation group (controller decision +
time-bias refresh).   python tools/tc_ode_phase_times.py   (bench shape: 64 objects x 50 candidates, T0 = 0.55)"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from genpose_b200 import lib, ops, synth  # noqa: E402
from genpose_b200.sde import init_sde  # noqa: E402

ve_prior = init_sde("ve")[0]               # sigma_max = 50 (sde.py:90-97)

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
dbg = torch.zeros(2, NE, 16, dtype=torch.int64, device="cuda")
for _ in range(2):
    lib.check(L.gpb_sample_ode_tc_dbg(x0.data_ptr(), R, K, T0, 1e-5, 1e-5, 1000, ob.data_ptr(), eng.trunk_w.data_ptr(),
                                      eng.trunk_tc.data_ptr(), center.data_ptr(), pose.data_ptr(), stats.data_ptr(), ws.data_ptr(),
                                      ws.numel(), dbg.data_ptr(), NE, torch.cuda.current_stream().cuda_stream), "dbg")
torch.cuda.synchronize()
n = int(stats[0].item())
d = dbg.cpu().numpy().astype(np.float64)[0, :min(n, NE)]
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
