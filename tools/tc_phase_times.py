"""Decode the cycle stamps of gpb_sample_pc_tc_dbg: per-phase durations (cycles) of CTA 0, averaged over steps.
    python tools/tc_phase_times.py [T] [cta,cta,...] [bf16x3|f16x2] [team 0|1|2|4] [objects]     (default: the bench shape, 64 objects x 50 candidates)"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from genpose_b200 import lib, ops, synth  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 100
CTAS = [int(c) for c in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0]
W16 = len(sys.argv) > 3 and sys.argv[3] == "f16x2"
TEAM = int(sys.argv[4]) if len(sys.argv) > 4 else 0
B, K = (int(sys.argv[5]) if len(sys.argv) > 5 else 64), 50
sd = synth.make_state_dict(0, kappa=synth.stable_kappa(T))
eng = ops.Engine(sd)
pts = torch.from_numpy(synth.make_clouds(B, 100)).cuda()
center = pts.mean(dim=1).contiguous()
R = B * K
x0 = torch.from_numpy(synth.make_prior_noise(R, 100)).cuda()
ob = eng.object_bias(eng.encode(pts))
L = lib.load()
lib.check(L.gpb_set_tc_team(TEAM), "set_tc_team")
ws = torch.empty(L.gpb_sampler_workspace_bytes(R, T), dtype=torch.uint8, device="cuda")
ts = ops.time_grid(T, "cuda")
out = torch.empty(R, 9, device="cuda")
dbg = torch.zeros(2, T, 16, dtype=torch.int64, device="cuda")


def record(cta):
    for _ in range(2):
        fn, stream_w = (L.gpb_sample_pc_tc16, eng.trunk_tc16()) if W16 else (L.gpb_sample_pc_tc_dbg, eng.trunk_tc)
        lib.check(fn(x0.data_ptr(), R, K, T, 0.16, ob.data_ptr(), eng.trunk_w.data_ptr(), stream_w.data_ptr(),
                                         center.data_ptr(), 0, (cta << 56) | 1, ts.data_ptr(), out.data_ptr(), 0, ws.data_ptr(), ws.numel(),
                                         dbg.data_ptr(), torch.cuda.current_stream().cuda_stream), "dbg")
    torch.cuda.synchronize()
    return dbg.cpu().numpy().astype(np.float64)


if len(CTAS) > 1:
    print("per-CTA comparison (cycles, mean over steps): cta | gather wait | grid wait (publish->release) | heads epi | L1 epi | step")
    for c in CTAS:
        dd = record(c)[0, 5:-1]
        # non-leaders have no ds[14]: use ds[10]->ds[12] for them
        lead = c % 4 == 0
        grid = np.mean(dd[:, 12] - (dd[:, 14] if lead else dd[:, 10]))
        print(f"  cta {c:3d} (rank {c % 4}): send {np.mean(dd[:, 11] - dd[:, 9]):6.0f} wait {np.mean(dd[:, 10] - dd[:, 11]):6.0f}  {'grid' if lead else 'score+grid'} {grid:7.0f}  "
              f"heads {np.mean(dd[:, 8] - dd[:, 4]):7.0f}  L1 {np.mean(dd[:, 4] - dd[:, 2]):7.0f}  L0 {np.mean(dd[:, 2] - dd[:, 0]):7.0f}  "
              f"tail {np.mean(dd[:, 13] - dd[:, 12]):7.0f}  step {np.mean(dd[1:, 0] - dd[:-1, 0]):7.0f}")
    sys.exit(0)
d = record(CTAS[0])
row, mma = d[0, 5:-1], d[1, 5:-1]
# stamps of CTA 0 (team leader of tile 0), row thread 0 — see tc_sampler.cu
seq = [("wait layer-0 accumulator", 0, 1), ("epilogue layer 0 (2 units)", 1, 2), ("wait layer-1 accumulator", 2, 3),
       ("epilogue layer 1 (2 units)", 3, 4), ("wait head 128-col unit", 4, 5), ("epilogue head 128-col unit", 5, 6),
       ("wait head 64-col unit", 6, 7), ("epilogue head 64-col unit", 7, 8), ("column-half sync", 8, 9),
       ("wait peers' partials", 9, 10), ("score + norm + CTA barrier", 10, 14), ("publish + poll grid barrier", 14, 15),
       ("release CTA barrier", 15, 12), ("sum partials, update, mail x, publish x", 12, 13)]
print("leader row thread 0 (cycles, mean over steps):")
for n, a, b in seq:
    print(f"  {n:42s} {np.mean(row[:, b] - row[:, a]):9.0f}")
print(f"  {'step total':42s} {np.mean(row[1:, 0] - row[:-1, 0]):9.0f}")
print("MMA warp:")
print(f"  {'wait x_ready':42s} {np.mean(mma[:, 1] - mma[:, 0]):9.0f}")
print(f"  {'layer 0 issue':42s} {np.mean(mma[:, 3] - mma[:, 1]):9.0f}")
print(f"  {'layer 1 (2 units) span':42s} {np.mean(mma[:, 4] - mma[:, 3]):9.0f}")
print(f"  {'head slice (128 + 64 columns) span':42s} {np.mean(mma[:, 7] - mma[:, 4]):9.0f}")
print(f"  {'  of which: inside issue groups':42s} {np.mean(mma[:, 11]):9.0f}")
print(f"  {'            waiting on weight slots':42s} {np.mean(mma[:, 8]):9.0f}")
print(f"  {'            waiting on the A operand':42s} {np.mean(mma[:, 9]):9.0f}")
print(f"  {'            waiting on accumulator slots':42s} {np.mean(mma[:, 10]):9.0f}")
print(f"  {'step total':42s} {np.mean(mma[1:, 0] - mma[:-1, 0]):9.0f}")
# merged timeline of one step (same SM => same clock): offsets from the row thread's step start, mean over steps
ev = [("row: step start", row[:, 0]), ("row: L0 acc ready", row[:, 1]), ("row: L0 epilogue done (h1 published)", row[:, 2]),
      ("row: L1 unit a acc ready", row[:, 3]), ("row: L1 epilogue done (pf published)", row[:, 4]), ("row: head128 acc ready", row[:, 5]),
      ("row: head128 epilogue done", row[:, 6]), ("row: head64 acc ready", row[:, 7]), ("row: head64 epilogue done", row[:, 8]),
      ("row: partials sent", row[:, 11]), ("row: peers' partials in", row[:, 10]), ("row: norm published", row[:, 14]),
      ("row: grid word complete", row[:, 15]), ("row: x published", row[:, 13]),
      ("mma: x_ready seen", mma[:, 1]), ("mma: L1 unit a first half A ready", mma[:, 3]), ("mma: head128 first half A ready", mma[:, 4]),
      ("mma: all issued", mma[:, 7]), ("mma: L1 unit a all issued", mma[:, 12]), ("mma: L1 unit b all issued", mma[:, 13]),
      ("mma: L1 unit b issue starts", mma[:, 14])]
base = row[:, 0]
print("timeline (cycles after the row thread's step start; x_ready / x published belong to the step boundary):")
for name, v in sorted(ev, key=lambda e: np.mean((e[1] - base) % 1e9)):
    if np.all(v == 0):
        continue                       # stamp not recorded by this path (f16x2 issues layer 1 as fused groups)
    print(f"  {np.mean(v - base):9.0f}  {name}")
