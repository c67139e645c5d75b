// Self-test of the tcgen05 building block used by the tensor-core sampler: one CTA computes
//     D[128 x N] = A[128 x K] . B[N x K]^T     with the bf16x3 split  (Ahi.Bhi + Ahi.Blo + Alo.Bhi, fp32 accumulate in TMEM)
// A arrives as fp32 rows and is split + laid out by the row threads (exactly what the sampler's epilogues do);
// B arrives as pre-tiled bf16 operand images (what genpose_b200/weights.py produces) streamed with cp.async.bulk.
// Exposed through the C ABI as gpb_selftest_umma so that tests/test_gpu_tc.py can check the descriptor encodings
// against a plain matmul on the device.
#include "common.cuh"
#include "tc_common.cuh"

namespace gpb {
using namespace tc;

__global__ void __launch_bounds__(192, 1)
umma_selftest_kernel(const float *__restrict__ A, const uint16_t *__restrict__ Bhi, const uint16_t *__restrict__ Blo,
                     float *__restrict__ D, int K, int N, int variant, int swap_fields, int n_terms, int a_tmem, int repeat,
                     unsigned long long *__restrict__ cycles_out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t a_bytes = 128u * K * 2u, b_bytes = (uint32_t)N * K * 2u;
    uint8_t *sAhi = smem, *sAlo = sAhi + a_bytes, *sBhi = sAlo + a_bytes, *sBlo = sBhi + b_bytes;
    __shared__ __align__(8) uint64_t bar_b, bar_mma;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // operand geometry
    const uint32_t a_lbo = variant == 0 ? (128u / 8u) * 128u : 128u;
    const uint32_t a_sbo = variant == 0 ? 128u : (uint32_t)(K / 8) * 128u;
    const uint32_t b_lbo = variant == 0 ? (uint32_t)(N / 8) * 128u : 128u;
    const uint32_t b_sbo = variant == 0 ? 128u : (uint32_t)(K / 8) * 128u;

    if (tid == 0) {
        mbar_init(&bar_b, 1);
        mbar_init(&bar_mma, 1);
        fence_mbar_init();
    }
    if (warp == 4) tmem_alloc(&tmem_base_s, 512);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = tmem_base_s;

    if (warp == 5 && lane == 0) {   // producer: stream the two B images
        mbar_arrive_expect_tx(&bar_b, 2 * b_bytes);
        for (uint32_t off = 0; off < b_bytes; off += 32768u) {
            const uint32_t n = (b_bytes - off) < 32768u ? (b_bytes - off) : 32768u;
            bulk_g2s(sBhi + off, reinterpret_cast<const uint8_t *>(Bhi) + off, n, &bar_b);
            bulk_g2s(sBlo + off, reinterpret_cast<const uint8_t *>(Blo) + off, n, &bar_b);
        }
    }
    const bool a_f16 = (swap_fields & 4) != 0;    // two-product mode: A as fp16 hi/lo (relu_split_f16x2: A must be >= 0)
    if (warp < 4) {   // row threads: split A row `tid` into bf16 hi/lo and write the operand images
        const int r = tid;
        for (int k8 = 0; k8 < K / 8; ++k8) {
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (a_f16) {
                    relu_split_f16x2(A[(size_t)r * K + k8 * 8 + 2 * j], A[(size_t)r * K + k8 * 8 + 2 * j + 1], hi[j], lo[j]);
                    continue;
                }
                __nv_bfloat16 h0, l0, h1, l1;
                split_bf16(A[(size_t)r * K + k8 * 8 + 2 * j], h0, l0);
                split_bf16(A[(size_t)r * K + k8 * 8 + 2 * j + 1], h1, l1);
                hi[j] = pack_bf16(h0, h1);
                lo[j] = pack_bf16(l0, l1);
            }
            const uint32_t off = (uint32_t)k8 * a_lbo + (uint32_t)(r / 8) * a_sbo + (uint32_t)(r % 8) * 16u;
            *reinterpret_cast<uint4 *>(sAhi + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<uint4 *>(sAlo + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
        fence_proxy_async_smem();   // generic-proxy stores -> async proxy (tensor core reads)
        if (a_tmem) {   // the same rows straight into tensor memory: hi at columns [256, 256+K/2), lo at [384, 384+K/2)
            for (int c0 = 0; c0 < K / 2; c0 += 8) {
                uint32_t hi[8], lo[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (a_f16) {
                        relu_split_f16x2(A[(size_t)r * K + 2 * (c0 + j)], A[(size_t)r * K + 2 * (c0 + j) + 1], hi[j], lo[j]);
                        continue;
                    }
                    __nv_bfloat16 h0, l0, h1, l1;
                    split_bf16(A[(size_t)r * K + 2 * (c0 + j)], h0, l0);
                    split_bf16(A[(size_t)r * K + 2 * (c0 + j) + 1], h1, l1);
                    hi[j] = pack_bf16(h0, h1);
                    lo[j] = pack_bf16(l0, l1);
                }
                const uint32_t ta = tmem_base + ((uint32_t)(warp * 32) << 16);
                tmem_st8(ta + 256u + (uint32_t)c0, hi);
                tmem_st8(ta + 384u + (uint32_t)c0, lo);
            }
            tmem_st_wait();
            tc_fence_before_sync();
        }
    }
    __syncthreads();

    if (warp == 4) {   // MMA issuer: the whole warp runs the loop, one elected lane issues
        mbar_wait(&bar_b, 0);
        tc_fence_after_sync();
        // swap_fields bit 1 (timing aid only, SS form): issue M = 64 instructions (half the rows; D is then NOT the product the
        // caller expects) to measure whether a 64-row tile costs half the tensor time of a 128-row one
        // swap_fields bit 2: the B images hold fp16 (not bf16) values, A is split into fp16 hi/lo and only two products are formed,
        // Ahi.B (term 0) and Alo.B (term 1) — the arithmetic of the two-product samplers.  (swap_fields bit 3 additionally keeps A
        // in bf16: the mixed-format kind::f16 instruction, which this part refuses with an illegal-instruction fault.)
        const bool w16 = (swap_fields & 4) != 0;
        const uint32_t idesc = w16 ? make_idesc_f16kind_f32((swap_fields & 2) ? 64 : 128, N, 0, 0) : make_idesc_bf16_f32((swap_fields & 2) ? 64 : 128, N);
        const uint32_t ahi = smem_u32(sAhi), alo = smem_u32(sAlo), bhi = smem_u32(sBhi), blo = smem_u32(sBlo);
        bool acc = false;
        const unsigned long long t_start = clock64();
        if (repeat > 1) {
            // timing mode: ONE elected lane issues the whole stream back to back (as the sampler kernels do)
            if (elect_one_sync()) {
                for (int rep = 0; rep < repeat; ++rep)
                    for (int term = 0; term < n_terms; ++term) {
                        const bool a_lo = w16 ? term == 1 : term == 2;
                        const uint32_t a0 = a_lo ? alo : ahi, b0 = (!w16 && term == 1) ? blo : bhi;
                        for (int k16 = 0; k16 < K / 16; ++k16) {
                            const uint64_t ad = make_smem_desc(a0 + k16 * 2 * a_lbo, a_lbo, a_sbo);
                            const uint64_t bd = make_smem_desc(b0 + k16 * 2 * b_lbo, b_lbo, b_sbo);
                            if (a_tmem) umma_bf16_ts(tmem_base, tmem_base + (a_lo ? 384u : 256u) + (uint32_t)k16 * 8u, bd, idesc, acc);
                            else umma_bf16(tmem_base, ad, bd, idesc, acc);
                            acc = true;
                        }
                    }
            }
            __syncwarp();
        } else
        for (int rep = 0; rep < repeat; ++rep)
        for (int term = 0; term < n_terms; ++term) {   // 0: Ahi.Bhi  1: Ahi.Blo  2: Alo.Bhi   (w16: 0: Ahi.B  1: Alo.B)
            const bool a_lo = w16 ? term == 1 : term == 2;
            const uint32_t a0 = a_lo ? alo : ahi, b0 = (!w16 && term == 1) ? blo : bhi;
            for (int k16 = 0; k16 < K / 16; ++k16) {
                const uint64_t ad = (swap_fields & 1) ? make_smem_desc(a0 + k16 * 2 * a_lbo, a_sbo, a_lbo) : make_smem_desc(a0 + k16 * 2 * a_lbo, a_lbo, a_sbo);
                const uint64_t bd = (swap_fields & 1) ? make_smem_desc(b0 + k16 * 2 * b_lbo, b_sbo, b_lbo) : make_smem_desc(b0 + k16 * 2 * b_lbo, b_lbo, b_sbo);
                if (elect_one_sync()) {
                    if (a_tmem) umma_bf16_ts(tmem_base, tmem_base + (a_lo ? 384u : 256u) + (uint32_t)k16 * 8u, bd, idesc, acc);
                    else umma_bf16(tmem_base, ad, bd, idesc, acc);
                }
                __syncwarp();
                acc = true;
            }
        }
        if (elect_one_sync()) umma_commit(&bar_mma);
        __syncwarp();
        const unsigned long long t_issued = clock64();
        mbar_wait(&bar_mma, 0);
        if (cycles_out && lane == 0) {
            cycles_out[0] = t_issued - t_start;      // issue time of repeat * n_terms * K/16 MMAs
            cycles_out[1] = clock64() - t_start;     // until the last one has completed
        }
    }
    if (warp < 4) {   // epilogue: TMEM -> global
        mbar_wait(&bar_mma, 0);
        tc_fence_after_sync();
        const int r = tid;
        for (int c0 = 0; c0 < N; c0 += 32) {
            uint32_t v[32];
            tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) D[(size_t)r * N + c0 + j] = __uint_as_float(v[j]);
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 4) tmem_dealloc(tmem_base, 512);
}

}  // namespace gpb

using namespace gpb;

extern "C" int gpb_selftest_umma(const float *A, const uint16_t *Bhi, const uint16_t *Blo, float *D, int K, int N, int variant,
                                 int swap_fields, int n_terms, int a_tmem, int repeat, unsigned long long *cycles_out, void *stream) {
    GPB_REQUIRE(A && Bhi && Blo && D, "selftest_umma: NULL buffer");
    GPB_REQUIRE(K % 16 == 0 && K >= 16 && N % 16 == 0 && N >= 16 && N <= 256, "selftest_umma: K %% 16, N %% 16, N <= 256");
    const size_t smem = (size_t)2 * 128 * K * 2 + (size_t)2 * N * K * 2;
    GPB_REQUIRE(smem <= 200 * 1024, "selftest_umma: operands do not fit shared memory");
    GPB_CUDA(cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    umma_selftest_kernel<<<1, 192, smem, (cudaStream_t)stream>>>(A, Bhi, Blo, D, K, N, variant, swap_fields, n_terms, a_tmem, repeat < 1 ? 1 : repeat, cycles_out);
    GPB_LAUNCHED();
    return GPB_OK;
}
