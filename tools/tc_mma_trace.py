"""One step of the PC sampler's MMA warp, boundary by boundary (experiment build only: `make variant NAME=trace EXTRA=-DGPB_DBG_TRACE=1`,
run with GPB_LIB=genpose_b200/libgenpose_b200_trace.so).  The trace block sits behind the [2][T][16] phase stamps.
The boundary stamps live in the issue LOOP, which only the three-product arithmetic still runs (the f16x2 issuer is the straight-line
code this trace led to: profiles/r2x_mma_trace.txt was taken on the f16x2 loop of the commit before it), so bf16x3 is the default here.
    python tools/tc_mma_trace.py [T] [bf16x3] [team] [objects]"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from genpose_b200 import lib, ops, synth  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 100
W16 = len(sys.argv) > 2 and sys.argv[2] == "f16x2"   # (no boundary stamps in the f16x2 issuer: the trace prints only the row side)
TEAM = int(sys.argv[3]) if len(sys.argv) > 3 else 0
B, K = (int(sys.argv[4]) if len(sys.argv) > 4 else 64), 50
eng = ops.Engine(synth.make_state_dict(0, kappa=synth.stable_kappa(T)))
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
dbg = torch.zeros(2 * T * 16 + 64, dtype=torch.int64, device="cuda")
fn, stream_w = (L.gpb_sample_pc_tc16, eng.trunk_tc16()) if W16 else (L.gpb_sample_pc_tc_dbg, eng.trunk_tc)
for _ in range(2):
    lib.check(fn(x0.data_ptr(), R, K, T, 0.16, ob.data_ptr(), eng.trunk_w.data_ptr(), stream_w.data_ptr(), center.data_ptr(), 0, 1,
                 ts.data_ptr(), out.data_ptr(), 0, ws.data_ptr(), ws.numel(), dbg.data_ptr(), torch.cuda.current_stream().cuda_stream), "dbg")
torch.cuda.synchronize()
d = dbg.cpu().numpy().astype(np.float64)
tr = d[2 * T * 16:]
row = d[:T * 16].reshape(T, 16)[T // 2]
names = {0: "step start", 1: "L0 slot in", 2: "x ready", 3: "L0 half 0: accumulator free", 4: "L0 half 0: issued", 5: "L0 half 1: accumulator free",
         6: "L0 half 1: issued", 7: "L1a: slots in", 20: "L1b: slots in", 21: "L1b: accumulator free", 22: "L1b: issued",
         31: "head64: slots in", 32: "head64: accumulator free", 33: "head64: issued"}
for g in range(4):
    names[8 + 3 * g] = f"L1a q{g}: h1 quarter in"
    names[9 + 3 * g] = f"L1a q{g}: accumulator free / fence"
    names[10 + 3 * g] = f"L1a q{g}: issued"
for g in range(2):
    names[23 + 4 * g] = f"head128 g{g}: slots in"
    names[24 + 4 * g] = f"head128 g{g}: pf half in"
    names[25 + 4 * g] = f"head128 g{g}: accumulator free / fence"
    names[26 + 4 * g] = f"head128 g{g}: issued"
rnames = {1: "L0 acc ready", 2: "L0 epilogue done", 3: "L1a acc ready", 4: "L1 epilogue done", 5: "head128 acc ready", 6: "head128 epilogue done",
          7: "head64 acc ready", 8: "head64 epilogue done", 11: "partials sent", 10: "peers' partials in", 14: "norm published",
          15: "grid word complete", 13: "x published"}
ev = [(tr[i] - tr[0], "mma: " + n) for i, n in names.items() if tr[i] > 0] + [(row[i] - tr[0], "    row: " + n) for i, n in rnames.items()]
prev = 0.0
for t, n in sorted(ev):
    print(f"{t:8.0f}  (+{t - prev:5.0f})  {n}")
    prev = t
