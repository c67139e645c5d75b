"""Tensor-pipe micro-benchmark with the self-test kernel: cycles per tcgen05.mma (M = 128, K = 16 per instruction) for SS / TS
operands over N, and for M = 64 instructions (SS form) — is a 64-row tile half the tensor time of a 128-row one?"""
import sys
import torch
sys.path.insert(0, ".")
from genpose_b200 import lib, weights
L = lib.load()
CASES = [(128, 256, 0, 128), (128, 256, 1, 128), (128, 128, 0, 128), (128, 128, 1, 128), (128, 64, 1, 128),
         (128, 192, 1, 128), (128, 160, 1, 128), (128, 96, 1, 128), (128, 32, 1, 128), (128, 16, 1, 128),
         (128, 256, 0, 64), (128, 128, 0, 64), (128, 64, 0, 64)]
for (K, N, a_tmem, M) in CASES:
    A = torch.randn(128, K).cuda()
    B = torch.randn(N, K)
    bhi, blo = weights.split_bf16(B)
    ih, il = weights.umma_image(bhi).cuda(), weights.umma_image(blo).cuda()
    D = torch.zeros(128, N, device="cuda")
    cyc = torch.zeros(2, dtype=torch.int64, device="cuda")
    for rep in (1, 16):
        lib.check(L.gpb_selftest_umma(A.data_ptr(), ih.data_ptr(), il.data_ptr(), D.data_ptr(), K, N, 0, 2 if M == 64 else 0, 3, a_tmem,
                                      rep, cyc.data_ptr(), torch.cuda.current_stream().cuda_stream), "selftest")
        torch.cuda.synchronize()
        n_mma = rep * 3 * K // 16
        c = cyc.cpu().tolist()
        print(f"M={M} K={K} N={N} A_in_{'TMEM' if a_tmem else 'SMEM'} mmas={n_mma:4d}: issue {c[0] / n_mma:7.1f} cyc/mma, "
              f"complete {c[1] / n_mma:7.1f} cyc/mma", flush=True)
