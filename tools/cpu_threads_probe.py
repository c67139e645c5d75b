"""How many torch intra-op threads make the CPU oracle fastest on this host (tiny matmuls do not scale to 128 threads)."""
import sys, time, os
sys.path.insert(0, ".")
import torch
import bench
for thr in (8, 16, 32, 64, os.cpu_count()):
    t0 = time.time()
    v, secs, _ = bench.cpu_oracle_rate(4, 50, 100, 2, threads=thr)
    print(f"threads {thr:4d}: {v:8.1f} cand/s at T=100 ({secs:.1f} s)")
