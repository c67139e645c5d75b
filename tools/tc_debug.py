import sys
import torch
sys.path.insert(0, ".")
from tests.test_gpu_tc import _selftest
combos = [(0, 0)] if len(sys.argv) < 2 else [tuple(int(c) for c in a.split(",")) for a in sys.argv[1:]]
for variant, swap in combos:
    d, ref, exact = _selftest(64, 256, variant, swap, 1)
    print("variant", variant, "swap", swap, "max|err| vs same-products ref:", float((d - ref).abs().max()), "ref scale", float(ref.abs().max()))
    print(" D[0,:4]", d[0, :4].tolist(), " ref[0,:4]", ref[0, :4].tolist())
    print(" D[37,100:104]", d[37, 100:104].tolist(), " ref", ref[37, 100:104].tolist())
