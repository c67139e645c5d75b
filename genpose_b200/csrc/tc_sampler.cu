// Predictor-corrector sampler on the 5th-generation tensor cores (tcgen05 + TMEM), throughput mode.
//
// Same algorithm, launch contract and update code as pc_sampler_kernel (scorenet.cu); only the score network's
// dense layers change engine: every layer is a 128-row x 256-column x K tcgen05.mma (kind::f16, bf16 operands,
// fp32 accumulation in TMEM) evaluated as the error-compensated split
//        A.B ~= Ahi.Bhi + Alo.Bhi + Ahi.Blo          (A = Ahi + Alo, B = Bhi + Blo, all bf16)
// whose products are exact in fp32, so the result carries ~2^-17 relative error per operand instead of bf16's
// 2^-9.  Measured in the oracle: final poses move by 2e-5 (single bf16: 9e-3, tf32: 8e-4) against
// the 1e-3 parity bound.
//
// One CTA owns a 128-row tile of candidates for all T steps.  Roles (warp-specialised, 320 threads):
//   warps 0-7  row warps.  Warp w reads TMEM lanes 32*(w%4).. (its rows) and columns [128*(w/4), +128).  They
//              turn accumulators into the next layer's A operand (bias + ReLU + bf16 split, written straight
//              into the canonical K-major operand image with conflict-free 16-byte stores), fold the three
//              head outputs into the 9 score components, and (warps 0-3, one thread per row, pose state in
//              registers) run the grid-wide gradient-norm reduction and the Langevin / Euler-Maruyama update.
//   warp 8     one elected thread issues every tcgen05.mma and the tcgen05.commit's that publish "accumulator
//              ready" / "weight stage free" on mbarriers; the warp also owns the TMEM allocation (512 columns =
//              two 256-column accumulators, so a head's epilogue overlaps the next head's MMAs).
//   warp 9     one elected thread streams the pre-tiled bf16 weight images (1,040 KiB per step, L2-resident) with
//              cp.async.bulk into a 4 x 16 KiB ring, completing on mbarriers.
// Per step and tile: P1 (K=16, x split in three bf16 pieces) -> P2 -> three heads, i.e. 5 + 4*48 MMAs.
#include "common.cuh"
#include "sampler_common.cuh"
#include "tc_common.cuh"

namespace gpb {
using namespace tc;

constexpr int kTcRows = 128;
constexpr int kTcRowWarps = 8;
constexpr int kTcThreads = (kTcRowWarps + 2) * 32;
constexpr uint32_t kStageBytes = 16384;
constexpr int kStages = 4;
constexpr int kStagesPerStep = 1 + 4 * 16;      // P1 (hi|lo in one stage) + 4 layers x 8 K-chunks x (hi, lo)
constexpr uint32_t kLboA = 2048, kLboB = 4096, kSbo = 128;
constexpr int kMaxObjPerTile = 4;
using TL = TrunkLayout;

// dynamic shared memory map (bytes)
constexpr uint32_t kOffAhi = 0;
constexpr uint32_t kOffAlo = kOffAhi + 65536;
constexpr uint32_t kOffX = kOffAlo + 65536;                       // 3 pieces x 4 KiB (rows x 16 K)
constexpr uint32_t kOffRing = kOffX + 3 * 4096;
constexpr uint32_t kOffOb = kOffRing + kStages * kStageBytes;     // [4][768] fp32 object biases of this tile
constexpr uint32_t kOffFpart = kOffOb + kMaxObjPerTile * 768 * 4; // [128][12] fp32 partial scores of the upper column half
constexpr uint32_t kTcSmemBytes = kOffFpart + 128 * 12 * 4;
static_assert(kTcSmemBytes <= 227 * 1024 - 2048, "tc sampler shared memory budget");

struct TcPcParams {
    PcParams pc;
    const uint8_t *wstream;   // kStagesPerStep x 16 KiB of bf16 operand images (genpose_b200/weights.py::pack_trunk_tc)
};

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// write one row's pose as three bf16 pieces (x = x1 + x2 + x3 to 24 bits) into the K=16 operand images
__device__ __forceinline__ void write_x_pieces(uint8_t *sX, int r, const float *x) {
    __nv_bfloat16 p[3][16];
#pragma unroll
    for (int c = 0; c < 16; ++c) {
        const float v = c < 9 ? x[c] : 0.f;
        p[0][c] = __float2bfloat16_rn(v);
        const float r1 = v - __bfloat162float(p[0][c]);
        p[1][c] = __float2bfloat16_rn(r1);
        p[2][c] = __float2bfloat16_rn(r1 - __bfloat162float(p[1][c]));
    }
#pragma unroll
    for (int pc = 0; pc < 3; ++pc)
#pragma unroll
        for (int k8 = 0; k8 < 2; ++k8) {
            const uint32_t off = (uint32_t)pc * 4096u + (uint32_t)k8 * kLboA + (uint32_t)(r >> 3) * kSbo + (uint32_t)(r & 7) * 16u;
            *reinterpret_cast<uint4 *>(sX + off) =
                make_uint4(pack_bf16(p[pc][8 * k8 + 0], p[pc][8 * k8 + 1]), pack_bf16(p[pc][8 * k8 + 2], p[pc][8 * k8 + 3]),
                           pack_bf16(p[pc][8 * k8 + 4], p[pc][8 * k8 + 5]), pack_bf16(p[pc][8 * k8 + 6], p[pc][8 * k8 + 7]));
        }
}

__global__ void __launch_bounds__(kTcThreads, 1)
tc_pc_sampler_kernel(TcPcParams tp) {
    const PcParams &p = tp.pc;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *sAhi = smem + kOffAhi, *sAlo = smem + kOffAlo, *sX = smem + kOffX, *sRing = smem + kOffRing;
    float *sOb = reinterpret_cast<float *>(smem + kOffOb);
    float *sFpart = reinterpret_cast<float *>(smem + kOffFpart);
    __shared__ __align__(8) uint64_t bar_full[kStages], bar_empty[kStages], bar_acc_full[2], bar_acc_empty[2], bar_x_ready, bar_a_ready;
    __shared__ uint32_t s_tmem_base;
    __shared__ float s_red[4];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row0 = blockIdx.x * kTcRows;
    const int obj_lo = row0 / p.K;

    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&bar_full[s], 1);
            mbar_init(&bar_empty[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&bar_acc_full[b], 1);
            mbar_init(&bar_acc_empty[b], kTcRowWarps);
        }
        mbar_init(&bar_x_ready, 4);
        mbar_init(&bar_a_ready, kTcRowWarps);
        fence_mbar_init();
    }
    if (warp == kTcRowWarps) tmem_alloc(&s_tmem_base, 512);
    // object biases of the (at most 4) objects this tile touches
    {
        const int last_row = min(row0 + kTcRows, p.R) - 1;
        const int n_obj = last_row / p.K - obj_lo + 1;
        for (int i = tid; i < n_obj * 768; i += kTcThreads) sOb[i] = p.obj_bias[(size_t)obj_lo * 768 + i];
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = s_tmem_base;
    const uint32_t idesc = make_idesc_bf16_f32(128, 256);

    if (warp == kTcRowWarps + 1) {
        // =============================== weight producer ===============================
        if (lane == 0) {
            const uint32_t total = (uint32_t)p.T * kStagesPerStep;
            for (uint32_t it = 0; it < total; ++it) {
                const uint32_t s = it % kStages;
                mbar_wait(&bar_empty[s], ((it / kStages) & 1u) ^ 1u);
                mbar_arrive_expect_tx(&bar_full[s], kStageBytes);
                bulk_g2s(sRing + s * kStageBytes, tp.wstream + (size_t)(it % kStagesPerStep) * kStageBytes, kStageBytes, &bar_full[s]);
            }
        }
    } else if (warp == kTcRowWarps) {
        // =============================== MMA issuer ===============================
        if (lane == 0) {
            const uint32_t ahi = smem_u32(sAhi), alo = smem_u32(sAlo), xb = smem_u32(sX), ring = smem_u32(sRing);
            uint32_t g = 0, it = 0, xr = 0, ar = 0;
            for (int step = 0; step < p.T; ++step) {
                // ---- layer 0: h1_pre = x . P1^T   (K = 16; x1,x2,x3 pieces against P1 hi|lo)
                mbar_wait(&bar_x_ready, xr & 1u);
                ++xr;
                {
                    const uint32_t b = g & 1u, n = g >> 1;
                    mbar_wait(&bar_acc_empty[b], (n & 1u) ^ 1u);
                    const uint32_t s = it % kStages;
                    mbar_wait(&bar_full[s], (it / kStages) & 1u);
                    tc_fence_after_sync();
                    const uint32_t d = tmem_base + b * 256u;
                    const uint64_t bhi = make_smem_desc(ring + s * kStageBytes, kLboB, kSbo);
                    const uint64_t blo = make_smem_desc(ring + s * kStageBytes + 8192u, kLboB, kSbo);
                    const uint64_t x1 = make_smem_desc(xb, kLboA, kSbo), x2 = make_smem_desc(xb + 4096u, kLboA, kSbo),
                                   x3 = make_smem_desc(xb + 8192u, kLboA, kSbo);
                    umma_bf16(d, x1, bhi, idesc, false);
                    umma_bf16(d, x2, bhi, idesc, true);
                    umma_bf16(d, x3, bhi, idesc, true);
                    umma_bf16(d, x1, blo, idesc, true);
                    umma_bf16(d, x2, blo, idesc, true);
                    umma_commit(&bar_empty[s]);
                    ++it;
                    umma_commit(&bar_acc_full[b]);
                    ++g;
                }
                // ---- layers 1..4: P2, head rot_x, head rot_y, head trans   (K = 256 each)
                for (int layer = 1; layer <= 4; ++layer) {
                    if (layer <= 2) {
                        mbar_wait(&bar_a_ready, ar & 1u);
                        ++ar;
                    }
                    const uint32_t b = g & 1u, n = g >> 1;
                    mbar_wait(&bar_acc_empty[b], (n & 1u) ^ 1u);
                    tc_fence_after_sync();
                    const uint32_t d = tmem_base + b * 256u;
                    bool acc = false;
                    for (int kc = 0; kc < 8; ++kc) {
                        const uint32_t s0 = it % kStages, s1 = (it + 1) % kStages;
                        mbar_wait(&bar_full[s0], (it / kStages) & 1u);
                        mbar_wait(&bar_full[s1], ((it + 1) / kStages) & 1u);
                        tc_fence_after_sync();
#pragma unroll
                        for (int j = 0; j < 2; ++j) {
                            const uint32_t k8 = (uint32_t)kc * 4u + 2u * j;
                            const uint64_t a_hi = make_smem_desc(ahi + k8 * kLboA, kLboA, kSbo);
                            const uint64_t a_lo = make_smem_desc(alo + k8 * kLboA, kLboA, kSbo);
                            const uint64_t b_hi = make_smem_desc(ring + s0 * kStageBytes + 2u * j * kLboB, kLboB, kSbo);
                            const uint64_t b_lo = make_smem_desc(ring + s1 * kStageBytes + 2u * j * kLboB, kLboB, kSbo);
                            umma_bf16(d, a_hi, b_hi, idesc, acc);
                            acc = true;
                            umma_bf16(d, a_lo, b_hi, idesc, true);
                            umma_bf16(d, a_hi, b_lo, idesc, true);
                        }
                        umma_commit(&bar_empty[s0]);
                        umma_commit(&bar_empty[s1]);
                        it += 2;
                    }
                    umma_commit(&bar_acc_full[b]);
                    ++g;
                }
            }
        }
    } else {
        // =============================== row warps ===============================
        const int q = warp & 3, half = warp >> 2;
        const int r = q * 32 + lane;              // row of the tile == TMEM lane
        const int row = row0 + r;                 // global candidate row
        const bool valid = row < p.R;
        const int cb = half * 128;                // this warp's column half
        const uint32_t tm_lane = (uint32_t)(q * 32) << 16;
        const uint32_t a_row_off = (uint32_t)(r >> 3) * kSbo + (uint32_t)(r & 7) * 16u;
        const float *ob_row = sOb + (size_t)((valid ? row : p.R - 1) / p.K - obj_lo) * 768;
        const float *W = p.W;

        float x[9];
#pragma unroll
        for (int c = 0; c < 9; ++c) x[c] = (valid && half == 0) ? p.x0[(size_t)row * 9 + c] : 0.f;
        if (half == 0) {
            write_x_pieces(sX, r, x);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_x_ready);
        }
        const float step_size = p.ts[0] - p.ts[1];
        const float sqrt_step = sqrtf(step_size);
        const float snr_norm = (float)((double)p.snr * 3.0);
        uint32_t g = 0;
        unsigned bar_target = 0;

        for (int step = 0; step < p.T; ++step) {
            // ---- epilogues of layers 0 and 1: accumulator -> bias + ReLU -> bf16 hi/lo -> A operand image ----
            for (int layer = 0; layer < 2; ++layer) {
                const uint32_t b = g & 1u, n = g >> 1;
                mbar_wait(&bar_acc_full[b], n & 1u);
                tc_fence_after_sync();
                const float *bias = W + (layer == 0 ? TL::p1_b : TL::p2_b);
#pragma unroll 1
                for (int c0 = cb; c0 < cb + 128; c0 += 32) {
                    uint32_t v[32];
                    tmem_ld32(tmem_base + tm_lane + b * 256u + (uint32_t)c0, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int j8 = 0; j8 < 4; ++j8) {
                        uint32_t hi[4], lo[4];
                        const float4 b0 = __ldg(reinterpret_cast<const float4 *>(bias + c0 + 8 * j8));
                        const float4 b1 = __ldg(reinterpret_cast<const float4 *>(bias + c0 + 8 * j8 + 4));
                        const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            __nv_bfloat16 h0, l0, h1, l1;
                            split_bf16(fmaxf(__uint_as_float(v[8 * j8 + 2 * e]) + bb[2 * e], 0.f), h0, l0);
                            split_bf16(fmaxf(__uint_as_float(v[8 * j8 + 2 * e + 1]) + bb[2 * e + 1], 0.f), h1, l1);
                            hi[e] = pack_bf16(h0, h1);
                            lo[e] = pack_bf16(l0, l1);
                        }
                        const uint32_t off = (uint32_t)((c0 >> 3) + j8) * kLboA + a_row_off;
                        *reinterpret_cast<uint4 *>(sAhi + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                        *reinterpret_cast<uint4 *>(sAlo + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                    }
                }
                fence_proxy_async_smem();
                tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(&bar_acc_empty[b]);
                    mbar_arrive(&bar_a_ready);
                }
                ++g;
            }
            // ---- epilogues of the three heads: relu(acc + obj_bias + t_bias) . O  -> 3 score components each ----
            const float *tb = p.tb_table + (size_t)step * 768;
            float f[9];
#pragma unroll 1
            for (int h = 0; h < 3; ++h) {
                const uint32_t b = g & 1u, n = g >> 1;
                mbar_wait(&bar_acc_full[b], n & 1u);
                tc_fence_after_sync();
                float o0 = 0.f, o1 = 0.f, o2 = 0.f;
                const float *w0 = W + TL::o_w + (size_t)(3 * h) * 256;
#pragma unroll 1
                for (int c0 = cb; c0 < cb + 128; c0 += 32) {
                    uint32_t v[32];
                    tmem_ld32(tmem_base + tm_lane + b * 256u + (uint32_t)c0, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int j4 = 0; j4 < 8; ++j4) {
                        const int c = c0 + 4 * j4;
                        const float4 ob = *reinterpret_cast<const float4 *>(ob_row + h * 256 + c);
                        const float4 t4 = __ldg(reinterpret_cast<const float4 *>(tb + h * 256 + c));
                        const float4 wa = __ldg(reinterpret_cast<const float4 *>(w0 + c));
                        const float4 wb = __ldg(reinterpret_cast<const float4 *>(w0 + 256 + c));
                        const float4 wc = __ldg(reinterpret_cast<const float4 *>(w0 + 512 + c));
                        const float h0 = fmaxf(__uint_as_float(v[4 * j4 + 0]) + ob.x + t4.x, 0.f);
                        const float h1 = fmaxf(__uint_as_float(v[4 * j4 + 1]) + ob.y + t4.y, 0.f);
                        const float h2 = fmaxf(__uint_as_float(v[4 * j4 + 2]) + ob.z + t4.z, 0.f);
                        const float h3 = fmaxf(__uint_as_float(v[4 * j4 + 3]) + ob.w + t4.w, 0.f);
                        o0 = fmaf(h3, wa.w, fmaf(h2, wa.z, fmaf(h1, wa.y, fmaf(h0, wa.x, o0))));
                        o1 = fmaf(h3, wb.w, fmaf(h2, wb.z, fmaf(h1, wb.y, fmaf(h0, wb.x, o1))));
                        o2 = fmaf(h3, wc.w, fmaf(h2, wc.z, fmaf(h1, wc.y, fmaf(h0, wc.x, o2))));
                    }
                }
                tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_acc_empty[b]);
                if (half == 1) {
                    sFpart[r * 12 + 3 * h + 0] = o0;
                    sFpart[r * 12 + 3 * h + 1] = o1;
                    sFpart[r * 12 + 3 * h + 2] = o2;
                } else {
                    f[3 * h + 0] = o0;
                    f[3 * h + 1] = o1;
                    f[3 * h + 2] = o2;
                }
                ++g;
            }
            named_bar_sync(1, kTcRowWarps * 32);      // upper-half partials are in sFpart
            if (half == 1) continue;                  // warps 4-7 go straight to the next step's epilogues

            // ---- score, batch-mean gradient norm, update (warps 0-3: one thread per row) ----
            const float t = p.ts[step];
            const float sigma = sigma_of_t(t);
            const float stdv = sigma + 1e-7f;
            float gr[9], n2 = 0.f;
#pragma unroll
            for (int c = 0; c < 9; ++c) {
                gr[c] = ((f[c] + sFpart[r * 12 + c]) + __ldg(W + TL::o_b + c)) / stdv;
                n2 = fmaf(gr[c], gr[c], n2);
            }
            const float wsum = warp_sum(valid ? sqrtf(n2) : 0.f);
            if (lane == 0) s_red[q] = wsum;
            named_bar_sync(2, 128);
            if (tid == 0) {
                p.partial[(step & 1) * gridDim.x + blockIdx.x] = (s_red[0] + s_red[1]) + (s_red[2] + s_red[3]);
                __threadfence();
                bar_target += gridDim.x;
                atomicAdd(p.barrier, 1u);
                while (ld_acquire_u32(p.barrier) < bar_target) {
                }
                __threadfence();
            } else {
                bar_target += gridDim.x;
            }
            named_bar_sync(2, 128);
            float tot = 0.f;
            for (int i = 0; i < (int)gridDim.x; ++i) tot += __ldcg(p.partial + (step & 1) * gridDim.x + i);
            const float grad_norm = tot / (float)p.R;
            const PcStepConsts sc = pc_step_consts(grad_norm, snr_norm, sigma, step_size, sqrt_step);
            float m[9];
            if (valid) {
                pc_row_update(p, sc, step, row, x, gr, m);
                const float *ctr = p.pts_center + (size_t)(row / p.K) * 3;
                if (p.process) {
                    float *dst = p.process + ((size_t)row * p.T + step) * 9;
#pragma unroll
                    for (int c = 0; c < 9; ++c) dst[c] = x[c] + (c >= 6 ? ctr[c - 6] : 0.f);
                }
                if (step == p.T - 1) {
#pragma unroll
                    for (int c = 6; c < 9; ++c) m[c] += ctr[c - 6];
                    gram_schmidt6(m);
#pragma unroll
                    for (int c = 0; c < 9; ++c) p.mean_x[(size_t)row * 9 + c] = m[c];
                }
            }
            write_x_pieces(sX, r, x);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_x_ready);
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == kTcRowWarps) tmem_dealloc(tmem_base, 512);
}

}  // namespace gpb

using namespace gpb;

extern "C" size_t gpb_trunk_tc_stream_bytes(void) { return (size_t)kStagesPerStep * kStageBytes; }

extern "C" int gpb_sample_pc_tc(const float *x0, int R, int K, int num_steps, float snr, const float *obj_bias, const float *W,
                                const void *tc_stream, const float *pts_center, const float *step_noise, uint64_t seed,
                                const float *time_grid, float *mean_x, float *process, void *workspace, size_t workspace_bytes,
                                void *stream) {
    GPB_REQUIRE(R >= 0 && K >= 1 && num_steps >= 2, "sample_pc_tc: need R >= 0, K >= 1, num_steps >= 2");
    if (R == 0) return GPB_OK;
    GPB_REQUIRE(x0 && obj_bias && W && tc_stream && pts_center && time_grid && mean_x && workspace, "sample_pc_tc: NULL buffer");
    GPB_REQUIRE(127 / K + 2 <= kMaxObjPerTile, "sample_pc_tc: K=%d too small (a 128-row tile may span at most %d objects); "
                "use gpb_sample_pc", K, kMaxObjPerTile);
    GPB_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0 && (reinterpret_cast<uintptr_t>(tc_stream) & 15) == 0,
                "sample_pc_tc: workspace must be 256-byte and the weight stream 16-byte aligned");
    const size_t need = carve_sampler(nullptr, R, num_steps).bytes;
    if (workspace_bytes < need) {
        set_error("sample_pc_tc: workspace %zu < required %zu bytes", workspace_bytes, need);
        return GPB_EWORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    int dev = 0, sms = 0;
    GPB_CUDA(cudaGetDevice(&dev));
    GPB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int grid = (R + kTcRows - 1) / kTcRows;
    GPB_REQUIRE(grid <= sms, "sample_pc_tc: R=%d needs %d co-resident CTAs but the device has %d SMs; split the batch", R, grid, sms);

    SamplerWs w = carve_sampler(workspace, R, num_steps);
    GPB_CUDA(cudaMemsetAsync(w.barrier, 0, 256, st));
    int rc = launch_time_bias_table(time_grid, num_steps, W, w.tb_table, st);
    if (rc) return rc;

    TcPcParams tp{};
    PcParams &p = tp.pc;
    p.x0 = x0; p.R = R; p.K = K; p.T = num_steps; p.snr = snr;
    p.obj_bias = obj_bias; p.W = W; p.pts_center = pts_center; p.noise = step_noise; p.seed = seed;
    p.ts = time_grid; p.tb_table = w.tb_table; p.partial = w.partial; p.barrier = w.barrier;
    p.mean_x = mean_x; p.process = process; p.tiles_per_cta = 1;
    tp.wstream = reinterpret_cast<const uint8_t *>(tc_stream);
    GPB_CUDA(cudaFuncSetAttribute(tc_pc_sampler_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmemBytes));
    void *args[] = {&tp};
    GPB_CUDA(cudaLaunchCooperativeKernel((void *)tc_pc_sampler_kernel, dim3(grid), dim3(kTcThreads), args, kTcSmemBytes, st));
    g_launches.fetch_add(1);
    return GPB_OK;
}
