"""Per-kernel SASS opcode histogram of the built library (the evidence B200_PROFILING.md asks for: UTC*MMA = tcgen05.mma,
LDTM / STTM = tcgen05.ld / st, UBLKCP = cp.async.bulk, FFMA2 / FADD2 = packed fp32 pairs, HMMA = legacy tensor path: must be absent).
    python tools/sass_histogram.py > profiles/<tag>_sass_opcode_histogram.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "genpose_b200", "libgenpose_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
KEYS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UBLKCP", "UTMALDG", "SYNCS", "UTCBAR", "FFMA2", "FADD2", "FFMA", "HMMA", "REDUX", "RED", "ATOM"]
hist, cur, n_inst = {}, None, collections.Counter()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", cur)[:90]
        hist[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        op = m.group(1)
        n_inst[cur] += 1
        for k in KEYS:
            if op == k or op.startswith(k + "."):
                hist[cur][k] += 1
                break
archs = sorted(set(re.findall(r"arch = (sm_\w+)", out)))
print(f"# {os.path.basename(lib)}: SASS architectures {archs}; per kernel: total instructions, then opcode counts (non-zero only)")
for name in sorted(hist, key=lambda n: -n_inst[n]):
    h = hist[name]
    print(f"{name:92s} {n_inst[name]:7d}  " + "  ".join(f"{k}={h[k]}" for k in KEYS if h[k]))
tot = collections.Counter()
for h in hist.values():
    tot.update(h)
print("# library totals: " + "  ".join(f"{k}={tot[k]}" for k in KEYS))
