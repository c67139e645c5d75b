// GroupAll (SA level 4: the 128 level-3 centres of an object -> MLP [515,256,C2,512] -> max) on the tensor cores.
//
// The 128 points of one object are exactly one M = 128 MMA tile, so every layer is a plain [128 x K] . [K x N] product
// per (object, scale).  Activations do not fit tensor memory next to the accumulators at these widths (K up to 512), so
// the three layers are three launches of ONE generic kernel; activations travel between them through L2 as bf16 hi/lo
// operand images in the canonical K-major UMMA layout  [K/8][128 rows][8]  (written coalesced by the epilogue, 16 bytes
// per thread), and both operands stream through a 6 x 16 KiB cp.async.bulk ring:
//     slot = A_hi K-step (4 KiB) | A_lo (4 KiB) | W_hi (128 rows, 4 KiB) | W_lo (4 KiB);   3 SS-form MMAs per slot
// One CTA = (object, scale, 128 output columns); two CTAs per SM (128 TMEM columns each).
//   layer 1: A = level-3 features (image written by sa3_tc_kernel), + Wx . xyz in the epilogue      4 tiles / object
//   layer 2: K = 256, N = 256 | 384                                                                5 tiles / object
//   layer 3: K = 256 | 384, N = 512, epilogue = relu + max over the 128 rows -> pts_feat            8 tiles / object
#include "common.cuh"
#include "tc_common.cuh"

namespace gpb {
using namespace tc;

constexpr int kGaEpiWarps = 4;
constexpr int kGaThreads = (kGaEpiWarps + 2) * 32;
constexpr int kGaSlots = 6;
constexpr uint32_t kGaSlotBytes = 16384;
constexpr uint32_t kGaSmemBytes = kGaSlots * kGaSlotBytes + 4 * 128 * 4;

struct GaTile {
    uint32_t w_off;   // bytes from GaParams::w to this tile's first slot image (k16 x 8 KiB)
    int scale, n0, k16;
};
struct GaParams {
    const uint8_t *a_hi[2], *a_lo[2];   // per scale: A operand images, object stride a_stride bytes
    uint32_t a_stride[2];
    const uint8_t *w;
    const float *bias[2];               // per scale, indexed by output column
    const float *wx[2];                 // layer 1: [3][256] per scale, else nullptr
    const float *xyz;                   // [B,128,3]
    uint8_t *o_hi[2], *o_lo[2];         // EPI_ACT: next layer's A images
    uint32_t o_stride[2];
    float *pts_feat;                    // EPI_MAX: [B,1024]
    GaTile tile[8];
};

template <bool kMax>
__global__ void __maxnreg__(128) ga_gemm_kernel(const __grid_constant__ GaParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    int *sMax = reinterpret_cast<int *>(smem + kGaSlots * kGaSlotBytes);   // [4 warps][128]
    __shared__ __align__(8) uint64_t bar_full[kGaSlots], bar_empty[kGaSlots], bar_acc_full;
    __shared__ uint32_t s_tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const GaTile t = p.tile[blockIdx.x];
    const int b = blockIdx.y, s = t.scale;

    if (tid == 0) {
        for (int i = 0; i < kGaSlots; ++i) {
            mbar_init(&bar_full[i], 1);
            mbar_init(&bar_empty[i], 1);
        }
        mbar_init(&bar_acc_full, 1);
        fence_mbar_init();
    }
    if (warp == kGaEpiWarps) tmem_alloc(&s_tmem_base, 128);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = s_tmem_base;

    if (warp == kGaEpiWarps + 1) {
        if (lane == 0) {
            const uint8_t *a_hi = p.a_hi[s] + (size_t)b * p.a_stride[s], *a_lo = p.a_lo[s] + (size_t)b * p.a_stride[s];
            const uint8_t *w = p.w + t.w_off;
            for (int k = 0; k < t.k16; ++k) {
                const int sl = k % kGaSlots;
                mbar_wait(&bar_empty[sl], (((uint32_t)k / kGaSlots) & 1u) ^ 1u);
                mbar_arrive_expect_tx(&bar_full[sl], kGaSlotBytes);
                uint8_t *dst = smem + sl * kGaSlotBytes;
                bulk_g2s(dst, a_hi + (size_t)k * 4096, 4096, &bar_full[sl]);
                bulk_g2s(dst + 4096, a_lo + (size_t)k * 4096, 4096, &bar_full[sl]);
                bulk_g2s(dst + 8192, w + (size_t)k * 8192, 8192, &bar_full[sl]);
            }
        }
    } else if (warp == kGaEpiWarps) {
        const uint32_t ring = smem_u32(smem);
        const uint32_t idesc = make_idesc_bf16_f32(128, 128);
        for (int k = 0; k < t.k16; ++k) {
            const int sl = k % kGaSlots;
            mbar_wait(&bar_full[sl], ((uint32_t)k / kGaSlots) & 1u);
            tc_fence_after_sync();
            if (elect_one_sync()) {
                const uint32_t sb = ring + sl * kGaSlotBytes;
                const uint64_t a_hi = make_smem_desc(sb, 2048, 128), a_lo = make_smem_desc(sb + 4096, 2048, 128);
                const uint64_t b_hi = make_smem_desc(sb + 8192, 2048, 128), b_lo = make_smem_desc(sb + 12288, 2048, 128);
                umma_bf16(tmem_base, a_hi, b_hi, idesc, k != 0);
                umma_bf16(tmem_base, a_lo, b_hi, idesc, true);
                umma_bf16(tmem_base, a_hi, b_lo, idesc, true);
                umma_commit(&bar_empty[sl]);
                if (k == t.k16 - 1) umma_commit(&bar_acc_full);
            }
            __syncwarp();
        }
    } else {
        const int r = warp * 32 + lane;
        const uint32_t tm_row = tmem_base + ((uint32_t)(warp * 32) << 16);
        const float *bias = p.bias[s] + t.n0;
        float x = 0.f, y = 0.f, z = 0.f;
        const float *wx = p.wx[s];
        if (!kMax && wx) {
            const float *q = p.xyz + ((size_t)b * 128 + r) * 3;
            x = q[0], y = q[1], z = q[2];
            wx += t.n0;
        }
        mbar_wait(&bar_acc_full, 0);
        tc_fence_after_sync();
#pragma unroll 1
        for (int blk = 0; blk < 4; ++blk) {
            uint32_t v[32];
            tmem_ld32(tm_row + (uint32_t)(blk * 32), v);
            tmem_ld_wait();
            if constexpr (kMax) {
                int keep = 0;
#pragma unroll
                for (int jj = 0; jj < 32; ++jj) {
                    const float h = fmaxf(__uint_as_float(v[jj]) + __ldg(bias + blk * 32 + jj), 0.f);
                    const int m = __reduce_max_sync(0xffffffffu, __float_as_int(h));   // h >= 0: integer order == float order
                    keep = lane == jj ? m : keep;
                }
                sMax[warp * 128 + blk * 32 + lane] = keep;
            } else {
                uint8_t *o_hi = p.o_hi[s] + (size_t)b * p.o_stride[s], *o_lo = p.o_lo[s] + (size_t)b * p.o_stride[s];
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    uint32_t hi[4], lo[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int c = blk * 32 + g * 8 + e * 2;
                        float h0 = __uint_as_float(v[g * 8 + e * 2]) + __ldg(bias + c), h1 = __uint_as_float(v[g * 8 + e * 2 + 1]) + __ldg(bias + c + 1);
                        if (wx) {
                            h0 = fmaf(z, __ldg(wx + 512 + c), fmaf(y, __ldg(wx + 256 + c), fmaf(x, __ldg(wx + c), h0)));
                            h1 = fmaf(z, __ldg(wx + 512 + c + 1), fmaf(y, __ldg(wx + 256 + c + 1), fmaf(x, __ldg(wx + c + 1), h1)));
                        }
                        split_bf16x2(fmaxf(h0, 0.f), fmaxf(h1, 0.f), hi[e], lo[e]);
                    }
                    const size_t off = (size_t)((t.n0 + blk * 32 + g * 8) >> 3) * 2048 + (size_t)r * 16;
                    *reinterpret_cast<uint4 *>(o_hi + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                    *reinterpret_cast<uint4 *>(o_lo + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                }
            }
        }
        if constexpr (kMax) {
            asm volatile("bar.sync 1, 128;" ::: "memory");
            const int m = max(max(sMax[r], sMax[128 + r]), max(sMax[256 + r], sMax[384 + r]));
            p.pts_feat[(size_t)b * 1024 + s * 512 + t.n0 + r] = __int_as_float(m);
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == kGaEpiWarps) tmem_dealloc(tmem_base, 128);
}

// Weight blob of the GroupAll block (weights.py::pack_encoder_tc), bytes:
//   per scale s: fp32 [wx 3x256 | b1 256 | b2 384 (C2 used) | b3 512] = 1920 floats -> 7680 B, both scales first (15360 B, padded to 16384)
//   then slots: s0 L1 (2 tiles x 32) | s1 L1 (2 x 32) | s0 L2 (2 x 16) | s1 L2 (3 x 16) | s0 L3 (4 x 16) | s1 L3 (4 x 24), 8 KiB each
constexpr size_t kGaConstBytes = 16384;
constexpr int kGaConstFloats = 1920;
constexpr size_t kGaL1Slots = 2 * 2 * 32, kGaL2Slots = 2 * 16 + 3 * 16, kGaL3Slots = 4 * 16 + 4 * 24;
constexpr size_t kGaBlobBytes = kGaConstBytes + (kGaL1Slots + kGaL2Slots + kGaL3Slots) * 8192;

size_t ga_tc_blob_bytes() { return kGaBlobBytes; }
// scratch: A0 hi|lo (128 KiB each / object), h1 hi|lo per scale (64 KiB), h2 hi|lo per scale (96 KiB)
size_t ga_tc_scratch_bytes(int B) { return (size_t)B * (2 * 131072 + 4 * 65536 + 4 * 98304); }

int launch_groupall_tc(const uint8_t *blob, uint8_t *scratch, const float *xyz3, float *pts_feat, int B, cudaStream_t st) {
    const float *consts = reinterpret_cast<const float *>(blob);
    const uint8_t *w = blob + kGaConstBytes;
    uint8_t *a0_hi = scratch, *a0_lo = a0_hi + (size_t)B * 131072;
    uint8_t *h1 = a0_lo + (size_t)B * 131072;        // [hi s0 | lo s0 | hi s1 | lo s1] each B x 64 KiB
    uint8_t *h2 = h1 + (size_t)B * 4 * 65536;        // same with 96 KiB
    GPB_CUDA(cudaFuncSetAttribute(ga_gemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGaSmemBytes));
    GPB_CUDA(cudaFuncSetAttribute(ga_gemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGaSmemBytes));
    const int c2[2] = {256, 384};
    // ---- layer 1
    {
        GaParams p{};
        for (int s = 0; s < 2; ++s) {
            p.a_hi[s] = a0_hi, p.a_lo[s] = a0_lo, p.a_stride[s] = 131072;
            p.bias[s] = consts + s * kGaConstFloats + 768;
            p.wx[s] = consts + s * kGaConstFloats;
            p.o_hi[s] = h1 + (size_t)B * 65536 * (2 * s), p.o_lo[s] = h1 + (size_t)B * 65536 * (2 * s + 1), p.o_stride[s] = 65536;
        }
        p.w = w, p.xyz = xyz3;
        for (int i = 0; i < 4; ++i) p.tile[i] = GaTile{(uint32_t)(i * 32 * 8192), i >> 1, (i & 1) * 128, 32};
        ga_gemm_kernel<false><<<dim3(4, B), kGaThreads, kGaSmemBytes, st>>>(p);
        GPB_LAUNCHED();
    }
    // ---- layer 2
    {
        GaParams p{};
        for (int s = 0; s < 2; ++s) {
            p.a_hi[s] = h1 + (size_t)B * 65536 * (2 * s), p.a_lo[s] = h1 + (size_t)B * 65536 * (2 * s + 1), p.a_stride[s] = 65536;
            p.bias[s] = consts + s * kGaConstFloats + 1024;
            p.wx[s] = nullptr;
            p.o_hi[s] = h2 + (size_t)B * 98304 * (2 * s), p.o_lo[s] = h2 + (size_t)B * 98304 * (2 * s + 1), p.o_stride[s] = 98304;
        }
        p.w = w + kGaL1Slots * 8192;
        int nt = 0;
        uint32_t off = 0;
        for (int s = 0; s < 2; ++s)
            for (int n0 = 0; n0 < c2[s]; n0 += 128, off += 16 * 8192) p.tile[nt++] = GaTile{off, s, n0, 16};
        ga_gemm_kernel<false><<<dim3(nt, B), kGaThreads, kGaSmemBytes, st>>>(p);
        GPB_LAUNCHED();
    }
    // ---- layer 3 + max
    {
        GaParams p{};
        for (int s = 0; s < 2; ++s) {
            p.a_hi[s] = h2 + (size_t)B * 98304 * (2 * s), p.a_lo[s] = h2 + (size_t)B * 98304 * (2 * s + 1), p.a_stride[s] = 98304;
            p.bias[s] = consts + s * kGaConstFloats + 1408;
        }
        p.w = w + (kGaL1Slots + kGaL2Slots) * 8192;
        p.pts_feat = pts_feat;
        int nt = 0;
        uint32_t off = 0;
        for (int s = 0; s < 2; ++s)
            for (int n0 = 0; n0 < 512; n0 += 128, off += (uint32_t)(c2[s] / 16) * 8192) p.tile[nt++] = GaTile{off, s, n0, c2[s] / 16};
        ga_gemm_kernel<true><<<dim3(nt, B), kGaThreads, kGaSmemBytes, st>>>(p);
        GPB_LAUNCHED();
    }
    return GPB_OK;
}

}  // namespace gpb
