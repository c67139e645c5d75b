"""Encoder timing (CUDA events, L2 flushed between passes): FFMA level 3 vs tcgen05 level 3, plus object_bias."""
import sys

import torch

sys.path.insert(0, ".")
from genpose_b200 import ops, synth  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
sd = synth.make_state_dict(0, kappa=-0.3)
eng = ops.Engine(sd)
pts = torch.from_numpy(synth.make_clouds(B, 100)).cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for prec in ("fp32", "bf16x3"):
    for _ in range(3):
        feat = eng.encode(pts, precision=prec)
    ts = []
    for _ in range(10):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        feat = eng.encode(pts, precision=prec)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    print(f"encode B={B} {prec}: median {sorted(ts)[5]:.3f} ms  min {min(ts):.3f} ms")
f32 = eng.encode(pts, precision="fp32")
ftc = eng.encode(pts, precision="bf16x3")
print("max |tc - fp32| / max|fp32| =", ((ftc - f32).abs().max() / f32.abs().max()).item())
ts = []
for _ in range(10):
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    ob = eng.object_bias(f32)
    b.record()
    torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
print(f"object_bias B={B}: median {sorted(ts)[5]:.3f} ms")
