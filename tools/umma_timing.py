"""Tensor-pipe micro-benchmark with the self-test kernel: cycles per tcgen05.mma for SS / TS operands and N = 128 / 256."""
import sys
import torch
sys.path.insert(0, ".")
from genpose_b200 import lib, weights
L = lib.load()
for (K, N, a_tmem) in [(128, 256, 0), (128, 256, 1), (128, 128, 0), (128, 128, 1), (128, 64, 1)]:
    A = torch.randn(128, K).cuda()
    B = torch.randn(N, K)
    bhi, blo = weights.split_bf16(B)
    ih, il = weights.umma_image(bhi).cuda(), weights.umma_image(blo).cuda()
    D = torch.zeros(128, N, device="cuda")
    cyc = torch.zeros(2, dtype=torch.int64, device="cuda")
    for rep in (1, 16):
        lib.check(L.gpb_selftest_umma(A.data_ptr(), ih.data_ptr(), il.data_ptr(), D.data_ptr(), K, N, 0, 0, 3, a_tmem, rep, cyc.data_ptr(),
                                      torch.cuda.current_stream().cuda_stream), "selftest")
        torch.cuda.synchronize()
        n_mma = rep * 3 * K // 16
        c = cyc.cpu().tolist()
        print(f"K={K} N={N} A_in_{'TMEM' if a_tmem else 'SMEM'} mmas={n_mma:4d}: issue {c[0] / n_mma:7.1f} cyc/mma, complete {c[1] / n_mma:7.1f} cyc/mma")
