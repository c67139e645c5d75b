"""Decode the cycle stamps of gpb_sample_pc_tc_dbg: per-phase durations (cycles) of CTA 0, averaged over steps.
    python tools/tc_phase_times.py [T]     (bench shape: 64 objects x 50 candidates)"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from genpose_b200 import lib, ops, synth  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 100
B, K = 64, 50
sd = synth.make_state_dict(0, kappa=synth.stable_kappa(T))
eng = ops.Engine(sd)
pts = torch.from_numpy(synth.make_clouds(B, 100)).cuda()
center = pts.mean(dim=1).contiguous()
R = B * K
x0 = torch.from_numpy(synth.make_prior_noise(R, 100)).cuda()
ob = eng.object_bias(eng.encode(pts))
L = lib.load()
ws = torch.empty(L.gpb_sampler_workspace_bytes(R, T), dtype=torch.uint8, device="cuda")
ts = ops.time_grid(T, "cuda")
out = torch.empty(R, 9, device="cuda")
dbg = torch.zeros(2, T, 16, dtype=torch.int64, device="cuda")
for _ in range(2):
    lib.check(L.gpb_sample_pc_tc_dbg(x0.data_ptr(), R, K, T, 0.16, ob.data_ptr(), eng.trunk_w.data_ptr(), eng.trunk_tc.data_ptr(),
                                     center.data_ptr(), 0, 1, ts.data_ptr(), out.data_ptr(), 0, ws.data_ptr(), ws.numel(), dbg.data_ptr(),
                                     torch.cuda.current_stream().cuda_stream), "dbg")
torch.cuda.synchronize()
d = dbg.cpu().numpy().astype(np.float64)
row, mma = d[0, 5:-1], d[1, 5:-1]
names = ["wait acc0", "epi0", "wait acc1", "epi1", "wait h0", "epi h0", "wait h1", "epi h1", "wait h2", "epi h2", "halves sync",
         "norm+grid barrier", "update+x pieces"]
idx = [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13]
print("row thread 0 (cycles, mean over steps):")
for n, a, b in zip(names, idx[:-1], idx[1:]):
    print(f"  {n:22s} {np.mean(row[:, b] - row[:, a]):9.0f}")
print(f"  step total             {np.mean(row[1:, 0] - row[:-1, 0]):9.0f}")
print(f"  (of the barrier phase: score+norm+named barrier {np.mean(row[:, 14] - row[:, 11]):7.0f}, publish+poll {np.mean(row[:, 15] - row[:, 14]):7.0f}, release named barrier {np.mean(row[:, 12] - row[:, 15]):7.0f})")
print("MMA thread:")
print(f"  wait x_ready           {np.mean(mma[:, 1] - mma[:, 0]):9.0f}")
for l in range(1, 5):
    nxt = mma[:, 3 + l] if l < 4 else mma[:, 7]
    print(f"  layer {l} issue span     {np.mean(nxt - mma[:, 2 + l]):9.0f}")
print(f"  waiting on weights     {np.mean(mma[:, 8]):9.0f}   (sum over the step)")
print(f"  waiting on A operand   {np.mean(mma[:, 9]):9.0f}")
print(f"  waiting on acc buffers {np.mean(mma[:, 10]):9.0f}")
print(f"  inside issue groups    {np.mean(mma[:, 11]):9.0f}   (32 groups of 12 MMAs per step)")
print(f"  tcgen05 fences         {np.mean(mma[:, 12]):9.0f}")
print(f"  step total             {np.mean(mma[1:, 0] - mma[:-1, 0]):9.0f}")
