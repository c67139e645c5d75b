// Set-abstraction level 3 (256 -> 128 centres, MLP [259,128,196,256], both radii) on the tensor cores.
//
// Level 3 is 61 % of the encoder's MACs.  One CTA = one tile of 128 (centre, neighbour) rows of one
// object and scale (8 centres x 16 neighbours or 4 x 32).  Layer 1 keeps the hoisted form of encoder.cu
// (h1 = relu(U[nbr] + Wx.(x_nbr - c) + b1), U = W1_feat . f from point_gemm_kernel) and is produced by the row threads
// straight into TENSOR MEMORY as the bf16 hi/lo A operand; layers 2 and 3 are tcgen05.mma (kind::f16, bf16x3 split,
// fp32 accumulation in TMEM, TS form) with the weights streamed as pre-tiled operand images through a 9 x 16 KiB
// cp.async.bulk ring (same machinery as tc_sampler.cu); the max over the neighbourhood is a lane reduction
// (redux.sync on the non-negative float bit patterns) on the accumulator rows.
//   TMEM: D [0,256) | A_hi [256,368) | A_lo [384,496)     (K up to 224 = 196 padded to a multiple of 32 columns)
//   stream per tile: 8 slots W2 (N = 224 rows, K = 128: one K=16 step per slot, hi 7 KiB | lo 7 KiB)
//                    14 slots W3 (N = 256, K = 224: hi 8 KiB | lo 8 KiB)                  = 22 slots of 16 KiB
#include "common.cuh"
#include "tc_common.cuh"

namespace gpb {
using namespace tc;

constexpr int kSaRows = 128;
constexpr int kSaRowWarps = 8;
constexpr int kSaThreads = (kSaRowWarps + 2) * 32;
constexpr int kSaC1 = 128, kSaC2 = 224, kSaC3 = 256;      // padded widths (196 -> 224)
constexpr uint32_t kSaSlotBytes = 16384;
constexpr int kSaSlots = 9;
constexpr int kSaW2Slots = kSaC1 / 16, kSaW3Slots = kSaC2 / 16;   // 8, 14
constexpr int kSaSlotsPerTile = kSaW2Slots + kSaW3Slots;
constexpr uint32_t kSaLboW2 = (kSaC2 / 8) * 128, kSaLboW3 = (kSaC3 / 8) * 128, kSaSbo = 128;
constexpr uint32_t kSaColD = 0, kSaColAhi = 256, kSaColAlo = 384;
constexpr int kSaNIn = 256, kSaNPoint = 128, kSaCTotal = 512;

// fp32 side constants of one (level 3, scale): [wx 3x128 | b1 128 | b2 224 | b3 256]
constexpr int kSaConstFloats = 3 * 128 + 128 + 224 + 256;

constexpr uint32_t kSaOffRing = 0;
constexpr uint32_t kSaOffConst = kSaOffRing + kSaSlots * kSaSlotBytes;
constexpr uint32_t kSaOffXyz = kSaOffConst + kSaConstFloats * 4;
constexpr uint32_t kSaOffNbr = kSaOffXyz + kSaNIn * 3 * 4;
constexpr uint32_t kSaOffCtr = kSaOffNbr + kSaRows * 4;
constexpr uint32_t kSaOffOut = kSaOffCtr + 32 * 4;                  // [8 centres][256] fp32
constexpr uint32_t kSaSmemBytes = kSaOffOut + 8 * 256 * 4;

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t *r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
                 "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
                 "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
                 : "memory");
}

template <int NS>
__global__ void __launch_bounds__(kSaThreads, 1)
sa3_tc_kernel(const float *__restrict__ xyz_in,    // [B,256,3]   level-2 centres
              const float *__restrict__ new_xyz,   // [B,128,3]   level-3 centres
              const float *__restrict__ U,         // [B,256,128] W1_feat . f
              const float *__restrict__ consts,    // kSaConstFloats
              const uint8_t *__restrict__ wstream, // kSaSlotsPerTile x 16 KiB
              float radius, int ch_off, float *__restrict__ feat_out /* [B,128,512] */,
              uint8_t *__restrict__ a0_hi, uint8_t *__restrict__ a0_lo /* GroupAll A-operand images [B][64][128][8] bf16, or null */) {
    constexpr int TC = kSaRows / NS;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *sRing = smem + kSaOffRing;
    float *sConst = reinterpret_cast<float *>(smem + kSaOffConst);
    float *sXyz = reinterpret_cast<float *>(smem + kSaOffXyz);
    int *sNbr = reinterpret_cast<int *>(smem + kSaOffNbr);
    float *sCtr = reinterpret_cast<float *>(smem + kSaOffCtr);
    int *sOut = reinterpret_cast<int *>(smem + kSaOffOut);
    __shared__ __align__(8) uint64_t bar_full[kSaSlots], bar_empty[kSaSlots], bar_acc_full, bar_a_ready, bar_in;
    __shared__ uint32_t s_tmem_base;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.y, c_base = blockIdx.x * TC;
    const float *sWx = sConst, *sB1 = sConst + 384, *sB2 = sConst + 512, *sB3 = sConst + 736;

    if (tid == 0) {
        for (int s = 0; s < kSaSlots; ++s) {
            mbar_init(&bar_full[s], 1);
            mbar_init(&bar_empty[s], 1);
        }
        mbar_init(&bar_acc_full, 1);
        mbar_init(&bar_a_ready, kSaRowWarps);
        mbar_init(&bar_in, 1);
        fence_mbar_init();
        // side inputs by bulk copy: fp32 constants, the object's level-2 centres, this tile's level-3 centres
        mbar_arrive_expect_tx(&bar_in, (uint32_t)(kSaConstFloats * 4 + kSaNIn * 12 + TC * 12));
        bulk_g2s(sConst, consts, kSaConstFloats * 4, &bar_in);
        bulk_g2s(sXyz, xyz_in + (size_t)b * kSaNIn * 3, kSaNIn * 12, &bar_in);
        bulk_g2s(sCtr, new_xyz + ((size_t)b * kSaNPoint + c_base) * 3, TC * 12, &bar_in);
    }
    if (warp == kSaRowWarps) tmem_alloc(&s_tmem_base, 512);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = s_tmem_base;

    if (warp == kSaRowWarps + 1) {
        // =============================== weight producer ===============================
        if (lane == 0) {
            for (uint32_t it = 0; it < (uint32_t)kSaSlotsPerTile; ++it) {
                const uint32_t s = it % kSaSlots;
                mbar_wait(&bar_empty[s], ((it / kSaSlots) & 1u) ^ 1u);
                mbar_arrive_expect_tx(&bar_full[s], kSaSlotBytes);
                bulk_g2s(sRing + s * kSaSlotBytes, wstream + (size_t)it * kSaSlotBytes, kSaSlotBytes, &bar_full[s]);
            }
        }
    } else if (warp == kSaRowWarps) {
        // =============================== MMA issuer (whole warp, elected issue groups) ===============================
        const uint32_t ring = smem_u32(sRing);
        const uint32_t t_ahi = tmem_base + kSaColAhi, t_alo = tmem_base + kSaColAlo, d = tmem_base + kSaColD;
        const uint32_t idesc2 = make_idesc_bf16_f32(128, kSaC2), idesc3 = make_idesc_bf16_f32(128, kSaC3);
        uint32_t it = 0;
        auto wait_slots = [&](int n) {
            if (lane < n) {
                const uint32_t itl = it + (uint32_t)lane;
                mbar_wait(&bar_full[itl % kSaSlots], (itl / kSaSlots) & 1u);
            }
            __syncwarp();
            tc_fence_after_sync();
        };
        // ---- layer 2: D[128 x 224] = h1[128 x 128] . W2^T
        mbar_wait(&bar_a_ready, 0);
        wait_slots(kSaW2Slots);
        {
            const uint32_t s_first = it % kSaSlots;
            if (elect_one_sync()) {
#pragma unroll
                for (int k = 0; k < kSaW2Slots; ++k) {
                    uint32_t s = s_first + (uint32_t)k;
                    s = s >= (uint32_t)kSaSlots ? s - (uint32_t)kSaSlots : s;
                    const uint32_t sb = ring + s * kSaSlotBytes;
                    const uint64_t b_hi = make_smem_desc(sb, kSaLboW2, kSaSbo);
                    const uint64_t b_lo = make_smem_desc(sb + 2u * kSaLboW2, kSaLboW2, kSaSbo);
                    umma_bf16_ts(d, t_ahi + 8u * k, b_hi, idesc2, k != 0);
                    umma_bf16_ts(d, t_alo + 8u * k, b_hi, idesc2, true);
                    umma_bf16_ts(d, t_ahi + 8u * k, b_lo, idesc2, true);
                    umma_commit(&bar_empty[s]);
                }
                umma_commit(&bar_acc_full);
            }
            __syncwarp();
            it += kSaW2Slots;
        }
        // ---- layer 3: D[128 x 256] = h2[128 x 224] . W3^T   (14 slots > ring: two issue groups)
        mbar_wait(&bar_a_ready, 1);
        for (int grp = 0; grp < 2; ++grp) {
            const int k0 = grp * 7;
            wait_slots(7);
            const uint32_t s_first = it % kSaSlots;
            if (elect_one_sync()) {
#pragma unroll
                for (int k = 0; k < 7; ++k) {
                    uint32_t s = s_first + (uint32_t)k;
                    s = s >= (uint32_t)kSaSlots ? s - (uint32_t)kSaSlots : s;
                    const uint32_t sb = ring + s * kSaSlotBytes;
                    const uint64_t b_hi = make_smem_desc(sb, kSaLboW3, kSaSbo);
                    const uint64_t b_lo = make_smem_desc(sb + 8192u, kSaLboW3, kSaSbo);
                    const uint32_t ac = 8u * (uint32_t)(k0 + k);
                    umma_bf16_ts(d, t_ahi + ac, b_hi, idesc3, (k0 + k) != 0);
                    umma_bf16_ts(d, t_alo + ac, b_hi, idesc3, true);
                    umma_bf16_ts(d, t_ahi + ac, b_lo, idesc3, true);
                    umma_commit(&bar_empty[s]);
                }
                if (grp == 1) umma_commit(&bar_acc_full);
            }
            __syncwarp();
            it += 7;
        }
    } else {
        // =============================== row warps ===============================
        const int q = warp & 3, cs = warp >> 2;
        const int r = q * 32 + lane;
        const uint32_t tm_row = tmem_base + ((uint32_t)(q * 32) << 16);
        // ---- ball query (one warp per centre; ascending-k order; pad with the first hit) — identical to encoder.cu ----
        mbar_wait(&bar_in, 0);
        {
            const float r2 = radius * radius;
            for (int tc = warp; tc < TC; tc += kSaRowWarps) {
                const float cx = sCtr[tc * 3 + 0], cy = sCtr[tc * 3 + 1], cz = sCtr[tc * 3 + 2];
                int cnt = 0, first = 0;
                for (int base = 0; base < kSaNIn && cnt < NS; base += 32) {
                    const int k = base + lane;
                    const bool hit = dist2_ref(cx, cy, cz, sXyz[k * 3], sXyz[k * 3 + 1], sXyz[k * 3 + 2]) < r2;
                    const unsigned mask = __ballot_sync(0xffffffffu, hit);
                    if (mask) {
                        if (cnt == 0) first = base + __ffs(mask) - 1;
                        const int slot = cnt + __popc(mask & ((1u << lane) - 1u));
                        if (hit && slot < NS) sNbr[tc * NS + slot] = k;
                        cnt += __popc(mask);
                    }
                }
                cnt = cnt < NS ? cnt : NS;
                for (int s = cnt + lane; s < NS; s += 32) sNbr[tc * NS + s] = first;
            }
            for (int i = tid; i < TC * 256; i += kSaRowWarps * 32) sOut[i] = 0;
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        // ---- layer 1 -> A operand: this thread = row r, channels [cs*64, +64) ----
        {
            const int j = sNbr[r], tc = r / NS;
            const float dx = sXyz[j * 3 + 0] - sCtr[tc * 3 + 0];
            const float dy = sXyz[j * 3 + 1] - sCtr[tc * 3 + 1];
            const float dz = sXyz[j * 3 + 2] - sCtr[tc * 3 + 2];
            const float *urow = U + ((size_t)b * kSaNIn + j) * kSaC1 + cs * 64;
#pragma unroll
            for (int blk = 0; blk < 2; ++blk) {
                uint32_t hi[16], lo[16];
#pragma unroll
                for (int c4 = 0; c4 < 8; ++c4) {
                    const int c = cs * 64 + blk * 32 + c4 * 4;
                    float4 v = __ldg(reinterpret_cast<const float4 *>(urow + blk * 32 + c4 * 4));
                    const float4 w0 = *reinterpret_cast<const float4 *>(sWx + c), w1 = *reinterpret_cast<const float4 *>(sWx + 128 + c),
                                 w2 = *reinterpret_cast<const float4 *>(sWx + 256 + c), bb = *reinterpret_cast<const float4 *>(sB1 + c);
                    v.x = fmaxf(fmaf(dz, w2.x, fmaf(dy, w1.x, fmaf(dx, w0.x, v.x + bb.x))), 0.f);
                    v.y = fmaxf(fmaf(dz, w2.y, fmaf(dy, w1.y, fmaf(dx, w0.y, v.y + bb.y))), 0.f);
                    v.z = fmaxf(fmaf(dz, w2.z, fmaf(dy, w1.z, fmaf(dx, w0.z, v.z + bb.z))), 0.f);
                    v.w = fmaxf(fmaf(dz, w2.w, fmaf(dy, w1.w, fmaf(dx, w0.w, v.w + bb.w))), 0.f);
                    split_bf16x2(v.x, v.y, hi[2 * c4], lo[2 * c4]);
                    split_bf16x2(v.z, v.w, hi[2 * c4 + 1], lo[2 * c4 + 1]);
                }
                tmem_st16(tm_row + kSaColAhi + (uint32_t)(cs * 32 + blk * 16), hi);
                tmem_st16(tm_row + kSaColAlo + (uint32_t)(cs * 32 + blk * 16), lo);
            }
            tmem_st_wait();
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_a_ready);
        }
        // ---- epilogue layer 2: relu(acc + b2) -> A operand (K = 224): cs 0: columns [0,128), cs 1: [128,224) ----
        {
            mbar_wait(&bar_acc_full, 0);
            tc_fence_after_sync();
            const int nblk = cs == 0 ? 4 : 3;
#pragma unroll 1
            for (int blk = 0; blk < nblk; ++blk) {
                const int c0 = cs * 128 + blk * 32;
                uint32_t v[32], hi[16], lo[16];
                tmem_ld32(tm_row + kSaColD + (uint32_t)c0, v);
                tmem_ld_wait();
#pragma unroll
                for (int jj = 0; jj < 16; ++jj) {
                    const float2 bb = *reinterpret_cast<const float2 *>(sB2 + c0 + 2 * jj);
                    split_bf16x2(fmaxf(__uint_as_float(v[2 * jj]) + bb.x, 0.f), fmaxf(__uint_as_float(v[2 * jj + 1]) + bb.y, 0.f), hi[jj], lo[jj]);
                }
                tmem_st16(tm_row + kSaColAhi + (uint32_t)(c0 / 2), hi);     // every layer-2 MMA has completed (bar_acc_full): A is free
                tmem_st16(tm_row + kSaColAlo + (uint32_t)(c0 / 2), lo);
            }
            tmem_st_wait();
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_a_ready);
        }
        // ---- epilogue layer 3: relu(acc + b3), max over the NS rows of a centre (lane reduction), columns [cs*128, +128) ----
        {
            mbar_wait(&bar_acc_full, 1);
            tc_fence_after_sync();
            // NS = 16: two centres share a warp.  A lane-dependent member mask would make the compiler emit two divergent
            // copies of the loop, so both halves are reduced with full-warp redux on values masked to the identity (0).
            const bool upper = lane >= 16;
#pragma unroll 1
            for (int blk = 0; blk < 4; ++blk) {
                const int c0 = cs * 128 + blk * 32;
                uint32_t v[32];
                tmem_ld32(tm_row + kSaColD + (uint32_t)c0, v);
                tmem_ld_wait();
                if constexpr (NS == 32) {
                    int keep = 0;
#pragma unroll
                    for (int jj = 0; jj < 32; ++jj) {
                        const float h = fmaxf(__uint_as_float(v[jj]) + sB3[c0 + jj], 0.f);
                        const int m = __reduce_max_sync(0xffffffffu, __float_as_int(h));   // h >= 0: integer order == float order
                        keep = lane == jj ? m : keep;                                      // lane jj keeps column jj
                    }
                    sOut[q * 256 + c0 + lane] = keep;
                } else {
                    int keep0 = 0, keep1 = 0;
#pragma unroll
                    for (int jj = 0; jj < 32; ++jj) {
                        const int h = __float_as_int(fmaxf(__uint_as_float(v[jj]) + sB3[c0 + jj], 0.f));
                        const int m0 = __reduce_max_sync(0xffffffffu, upper ? 0 : h);
                        const int m1 = __reduce_max_sync(0xffffffffu, upper ? h : 0);
                        keep0 = lane == jj ? m0 : keep0;
                        keep1 = lane == jj ? m1 : keep1;
                    }
                    sOut[(q * 2) * 256 + c0 + lane] = keep0;
                    sOut[(q * 2 + 1) * 256 + c0 + lane] = keep1;
                }
            }
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        for (int i = tid; i < TC * 256; i += kSaRowWarps * 32) {
            const int tc = i >> 8, c = i & 255;
            feat_out[((size_t)b * kSaNPoint + c_base + tc) * kSaCTotal + ch_off + c] = __int_as_float(sOut[i]);
        }
        if (a0_hi) {   // the same values as the bf16 hi/lo A operand of GroupAll layer 1 (ga_tc.cu): [K/8][128 rows][8]
            for (int i = tid; i < TC * 32; i += kSaRowWarps * 32) {
                const int tc = i >> 5, g = i & 31;
                const float4 v0 = *reinterpret_cast<const float4 *>(sOut + tc * 256 + g * 8), v1 = *reinterpret_cast<const float4 *>(sOut + tc * 256 + g * 8 + 4);
                uint32_t hi[4], lo[4];
                split_bf16x2(v0.x, v0.y, hi[0], lo[0]);
                split_bf16x2(v0.z, v0.w, hi[1], lo[1]);
                split_bf16x2(v1.x, v1.y, hi[2], lo[2]);
                split_bf16x2(v1.z, v1.w, hi[3], lo[3]);
                const size_t off = (size_t)b * 131072 + (size_t)((ch_off >> 3) + g) * 2048 + (size_t)(c_base + tc) * 16;
                *reinterpret_cast<uint4 *>(a0_hi + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                *reinterpret_cast<uint4 *>(a0_lo + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == kSaRowWarps) tmem_dealloc(tmem_base, 512);
}

// ===================================================================================================================
// Small-width set abstraction (level 1 scale 1: MLP [3,32,32,64]; level 2: [99,64,64|96,128]) on the tensor cores,
// PERSISTENT: the whole bf16 hi/lo operand image of W2 and W3 (<= 72 KiB) stays in shared memory, two CTAs per SM
// (256 TMEM columns each) loop over 128-row tiles so that one CTA's row phases overlap the other's MMAs.
//   TMEM (per CTA): D [0,128) | A_hi [128,176) | A_lo [192,240)
//   weight image:   W2 hi [C1/8][C2][8] | W2 lo | W3 hi [C2/8][C3][8] | W3 lo      (bf16, canonical K-major)
//   fp32 constants: wx[3][C1] | b1[C1] | b2[C2] | b3[C3]   (padded to 512 floats)
// ===================================================================================================================
constexpr int kS2Threads = (kSaRowWarps + 1) * 32;
constexpr int kS2ConstFloats = 512;
constexpr uint32_t kS2ColD = 0, kS2ColAhi = 128, kS2ColAlo = 192;

template <int NS_, int C1_, int C2_, int C3_, int NIN_, int NPOINT_, int CTOTAL_, bool HAS_U_>
struct SmallCfg {
    static constexpr int NS = NS_, C1 = C1_, C2 = C2_, C3 = C3_, NIN = NIN_, NPOINT = NPOINT_, CTOTAL = CTOTAL_;
    static constexpr bool HAS_U = HAS_U_;
    static constexpr int TC = kSaRows / NS;
    static constexpr int H = kSaRowWarps / TC;        // warps sharing one centre's ball query (1 or 2)
    static constexpr uint32_t w_bytes = (uint32_t)(C1 * C2 + C2 * C3) * 4u;
    static constexpr uint32_t off_w = 0;
    static constexpr uint32_t off_const = off_w + w_bytes;
    static constexpr uint32_t off_xyz = off_const + kS2ConstFloats * 4;          // [2][NIN*3]
    static constexpr uint32_t off_ctr = off_xyz + 2 * NIN * 12;                  // [2][32]
    static constexpr uint32_t off_hit = off_ctr + 2 * 32 * 4;                    // [TC][H][NS] candidate lists
    static constexpr uint32_t off_cnt = off_hit + kSaRows * H * 4;               // [TC][H]
    static constexpr uint32_t off_out = off_cnt + 16 * 4;                        // [TC][C3]
    static constexpr uint32_t bytes = off_out + TC * C3 * 4;
    static_assert(H == 1 || H == 2, "ball query split");
    static_assert(C1 % 32 == 0 && C2 % 32 == 0 && C3 % 64 == 0 && C1 <= 96 && C2 <= 96 && C3 <= 128, "widths");
};

template <int NCOLS>
__device__ __forceinline__ void tmem_st_cols(uint32_t taddr, const uint32_t *r) {
    static_assert(NCOLS == 8 || NCOLS == 16, "store width");
    if constexpr (NCOLS == 16) {
        tmem_st16(taddr, r);
    } else {
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]),
                     "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                     : "memory");
    }
}

template <class C>
__global__ void __maxnreg__(80)
sa_small_tc_kernel(const float *__restrict__ xyz_in,    // [B,NIN,3]      source points
                   const float *__restrict__ new_xyz,   // [B,NPOINT,3]   centres
                   const float *__restrict__ U,         // [B,NIN,C1]     W1_feat . f   (HAS_U) or nullptr
                   const float *__restrict__ consts,    // kS2ConstFloats
                   const uint8_t *__restrict__ wimg,    // C::w_bytes
                   float radius, int ch_off, int n_tiles, float *__restrict__ feat_out /* [B,NPOINT,CTOTAL] */) {
    constexpr int NS = C::NS, TC = C::TC, H = C::H, C1 = C::C1, C2 = C::C2, C3 = C::C3, NIN = C::NIN;
    constexpr int kTilesPerObj = C::NPOINT / TC;
    constexpr uint32_t kLboW2 = C2 * 16, kLboW3 = C3 * 16;
    constexpr uint32_t kW2Bytes = C1 * C2 * 2, kW3Bytes = C2 * C3 * 2;
    extern __shared__ __align__(1024) uint8_t smem[];
    float *sConst = reinterpret_cast<float *>(smem + C::off_const);
    float *sXyz = reinterpret_cast<float *>(smem + C::off_xyz);
    float *sCtr = reinterpret_cast<float *>(smem + C::off_ctr);
    int *sHit = reinterpret_cast<int *>(smem + C::off_hit);
    int *sCnt = reinterpret_cast<int *>(smem + C::off_cnt);
    int *sOut = reinterpret_cast<int *>(smem + C::off_out);
    __shared__ __align__(8) uint64_t bar_w, bar_in[2], bar_acc_full, bar_a_ready;
    __shared__ uint32_t s_tmem_base;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const float *sWx = sConst, *sB1 = sConst + 3 * C1, *sB2 = sConst + 4 * C1, *sB3 = sConst + 4 * C1 + C2;

    auto prefetch = [&](int tile, int buf) {   // one thread: the object's source points and this tile's centres
        const int b = tile / kTilesPerObj, c_base = (tile % kTilesPerObj) * TC;
        mbar_arrive_expect_tx(&bar_in[buf], (uint32_t)(NIN * 12 + TC * 12));
        bulk_g2s(sXyz + buf * NIN * 3, xyz_in + (size_t)b * NIN * 3, NIN * 12, &bar_in[buf]);
        bulk_g2s(sCtr + buf * 32, new_xyz + ((size_t)b * C::NPOINT + c_base) * 3, TC * 12, &bar_in[buf]);
    };

    if (tid == 0) {
        mbar_init(&bar_w, 1);
        mbar_init(&bar_in[0], 1);
        mbar_init(&bar_in[1], 1);
        mbar_init(&bar_acc_full, 1);
        mbar_init(&bar_a_ready, kSaRowWarps);
        fence_mbar_init();
        mbar_arrive_expect_tx(&bar_w, C::w_bytes + kS2ConstFloats * 4);
        bulk_g2s(smem + C::off_w, wimg, C::w_bytes, &bar_w);
        bulk_g2s(sConst, consts, kS2ConstFloats * 4, &bar_w);
        if ((int)blockIdx.x < n_tiles) prefetch(blockIdx.x, 0);
    }
    if (warp == kSaRowWarps) tmem_alloc(&s_tmem_base, 256);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = s_tmem_base;

    if (warp == kSaRowWarps) {
        // =============================== MMA issuer ===============================
        const uint32_t wbase = smem_u32(smem + C::off_w);
        const uint32_t w2_hi = wbase, w2_lo = wbase + kW2Bytes, w3_hi = wbase + 2 * kW2Bytes, w3_lo = w3_hi + kW3Bytes;
        const uint32_t t_ahi = tmem_base + kS2ColAhi, t_alo = tmem_base + kS2ColAlo, d = tmem_base + kS2ColD;
        const uint32_t idesc2 = make_idesc_bf16_f32(128, C2), idesc3 = make_idesc_bf16_f32(128, C3);
        mbar_wait(&bar_w, 0);
        int it = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            mbar_wait(&bar_a_ready, 0);
            tc_fence_after_sync();
            if (elect_one_sync()) {
                // the other side-input buffer was last read in the previous tile, which every row warp has left
                if (tile + (int)gridDim.x < n_tiles) prefetch(tile + gridDim.x, (it + 1) & 1);
#pragma unroll
                for (int k = 0; k < C1 / 16; ++k) {
                    const uint64_t b_hi = make_smem_desc(w2_hi + (uint32_t)k * 2u * kLboW2, kLboW2, kSaSbo);
                    const uint64_t b_lo = make_smem_desc(w2_lo + (uint32_t)k * 2u * kLboW2, kLboW2, kSaSbo);
                    umma_bf16_ts(d, t_ahi + 8u * k, b_hi, idesc2, k != 0);
                    umma_bf16_ts(d, t_alo + 8u * k, b_hi, idesc2, true);
                    umma_bf16_ts(d, t_ahi + 8u * k, b_lo, idesc2, true);
                }
                umma_commit(&bar_acc_full);
            }
            __syncwarp();
            mbar_wait(&bar_a_ready, 1);
            tc_fence_after_sync();
            if (elect_one_sync()) {
#pragma unroll
                for (int k = 0; k < C2 / 16; ++k) {
                    const uint64_t b_hi = make_smem_desc(w3_hi + (uint32_t)k * 2u * kLboW3, kLboW3, kSaSbo);
                    const uint64_t b_lo = make_smem_desc(w3_lo + (uint32_t)k * 2u * kLboW3, kLboW3, kSaSbo);
                    umma_bf16_ts(d, t_ahi + 8u * k, b_hi, idesc3, k != 0);
                    umma_bf16_ts(d, t_alo + 8u * k, b_hi, idesc3, true);
                    umma_bf16_ts(d, t_ahi + 8u * k, b_lo, idesc3, true);
                }
                umma_commit(&bar_acc_full);
            }
            __syncwarp();
        }
    } else {
        // =============================== row warps ===============================
        const int q = warp & 3, cs = warp >> 2;
        const int r = q * 32 + lane;
        const uint32_t tm_row = tmem_base + ((uint32_t)(q * 32) << 16);
        const float r2 = radius * radius;
        const bool upper = lane >= 16;
        mbar_wait(&bar_w, 0);                            // fp32 constants
        int it = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const int b = tile / kTilesPerObj, c_base = (tile % kTilesPerObj) * TC;
            const int buf = it & 1;
            const float *xyz = sXyz + buf * NIN * 3, *ctr = sCtr + buf * 32;
            mbar_wait(&bar_in[buf], (uint32_t)(it >> 1) & 1u);
            // ---- ball query: H warps per centre, each over its ascending slice of the source points; the row threads below
            //      concatenate the slices in order (== the first NS hits in ascending index) and pad with the first hit ----
            {
                const int tc = warp / H, h = warp % H;
                const float cx = ctr[tc * 3 + 0], cy = ctr[tc * 3 + 1], cz = ctr[tc * 3 + 2];
                int *hit = sHit + (tc * H + h) * NS;
                int cnt = 0;
                for (int base = h * (NIN / H); base < (h + 1) * (NIN / H) && cnt < NS; base += 32) {
                    const int k = base + lane;
                    const bool in = dist2_ref(cx, cy, cz, xyz[k * 3], xyz[k * 3 + 1], xyz[k * 3 + 2]) < r2;
                    const unsigned mask = __ballot_sync(0xffffffffu, in);
                    const int slot = cnt + __popc(mask & ((1u << lane) - 1u));
                    if (in && slot < NS) hit[slot] = k;
                    cnt += __popc(mask);
                }
                if (lane == 0) sCnt[tc * H + h] = cnt < NS ? cnt : NS;
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            // ---- layer 1 -> A operand: row r, channels [cs*C1/2, +C1/2) ----
            {
                constexpr int CH = C1 / 2;
                const int tc = r / NS, sl = r % NS;
                int j;
                {
                    const int c0 = sCnt[tc * H];
                    const int *h0 = sHit + (tc * H) * NS;
                    if constexpr (H == 1) {
                        j = sl < c0 ? h0[sl] : h0[0];          // the centre itself is a source point: c0 >= 1
                    } else {
                        const int c1 = sCnt[tc * H + 1];
                        const int *h1 = h0 + NS;
                        const int first = c0 > 0 ? h0[0] : h1[0];
                        j = sl < c0 ? h0[sl] : (sl - c0 < c1 ? h1[sl - c0] : first);
                    }
                }
                const float dx = xyz[j * 3 + 0] - ctr[tc * 3 + 0];
                const float dy = xyz[j * 3 + 1] - ctr[tc * 3 + 1];
                const float dz = xyz[j * 3 + 2] - ctr[tc * 3 + 2];
                float4 u[CH / 4];
                if constexpr (C::HAS_U) {
                    const float4 *urow = reinterpret_cast<const float4 *>(U + ((size_t)b * NIN + j) * C1 + cs * CH);
#pragma unroll
                    for (int c4 = 0; c4 < CH / 4; ++c4) u[c4] = __ldg(urow + c4);
                } else {
#pragma unroll
                    for (int c4 = 0; c4 < CH / 4; ++c4) u[c4] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
                uint32_t hi[CH / 2], lo[CH / 2];
#pragma unroll
                for (int c4 = 0; c4 < CH / 4; ++c4) {
                    const int c = cs * CH + c4 * 4;
                    float4 v = u[c4];
                    const float4 w0 = *reinterpret_cast<const float4 *>(sWx + c), w1 = *reinterpret_cast<const float4 *>(sWx + C1 + c),
                                 w2 = *reinterpret_cast<const float4 *>(sWx + 2 * C1 + c), bb = *reinterpret_cast<const float4 *>(sB1 + c);
                    v.x = fmaxf(fmaf(dz, w2.x, fmaf(dy, w1.x, fmaf(dx, w0.x, v.x + bb.x))), 0.f);
                    v.y = fmaxf(fmaf(dz, w2.y, fmaf(dy, w1.y, fmaf(dx, w0.y, v.y + bb.y))), 0.f);
                    v.z = fmaxf(fmaf(dz, w2.z, fmaf(dy, w1.z, fmaf(dx, w0.z, v.z + bb.z))), 0.f);
                    v.w = fmaxf(fmaf(dz, w2.w, fmaf(dy, w1.w, fmaf(dx, w0.w, v.w + bb.w))), 0.f);
                    split_bf16x2(v.x, v.y, hi[2 * c4], lo[2 * c4]);
                    split_bf16x2(v.z, v.w, hi[2 * c4 + 1], lo[2 * c4 + 1]);
                }
                tmem_st_cols<CH / 2>(tm_row + kS2ColAhi + (uint32_t)(cs * (CH / 2)), hi);
                tmem_st_cols<CH / 2>(tm_row + kS2ColAlo + (uint32_t)(cs * (CH / 2)), lo);
                tmem_st_wait();
                tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_a_ready);
            }
            // ---- epilogue layer 2: relu(acc + b2) -> A operand (K = C2); 32-column blocks dealt to the two column warps ----
            {
                mbar_wait(&bar_acc_full, 0);
                tc_fence_after_sync();
#pragma unroll 1
                for (int blk = cs; blk < C2 / 32; blk += 2) {
                    const int c0 = blk * 32;
                    uint32_t v[32], hi[16], lo[16];
                    tmem_ld32(tm_row + kS2ColD + (uint32_t)c0, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int jj = 0; jj < 16; ++jj) {
                        const float2 bb = *reinterpret_cast<const float2 *>(sB2 + c0 + 2 * jj);
                        split_bf16x2(fmaxf(__uint_as_float(v[2 * jj]) + bb.x, 0.f), fmaxf(__uint_as_float(v[2 * jj + 1]) + bb.y, 0.f), hi[jj],
                                     lo[jj]);
                    }
                    tmem_st16(tm_row + kS2ColAhi + (uint32_t)(c0 / 2), hi);
                    tmem_st16(tm_row + kS2ColAlo + (uint32_t)(c0 / 2), lo);
                }
                tmem_st_wait();
                tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_a_ready);
            }
            // ---- epilogue layer 3: relu(acc + b3), max over the NS rows of a centre; columns [cs*C3/2, +C3/2) ----
            {
                mbar_wait(&bar_acc_full, 1);
                tc_fence_after_sync();
#pragma unroll 1
                for (int blk = 0; blk < C3 / 64; ++blk) {
                    const int c0 = cs * (C3 / 2) + blk * 32;
                    uint32_t v[32];
                    tmem_ld32(tm_row + kS2ColD + (uint32_t)c0, v);
                    tmem_ld_wait();
                    if constexpr (NS == 32) {
                        int keep = 0;
#pragma unroll
                        for (int jj = 0; jj < 32; ++jj) {
                            const float h = fmaxf(__uint_as_float(v[jj]) + sB3[c0 + jj], 0.f);
                            const int m = __reduce_max_sync(0xffffffffu, __float_as_int(h));
                            keep = lane == jj ? m : keep;
                        }
                        sOut[q * C3 + c0 + lane] = keep;
                    } else {
                        int keep0 = 0, keep1 = 0;
#pragma unroll
                        for (int jj = 0; jj < 32; ++jj) {
                            const int h = __float_as_int(fmaxf(__uint_as_float(v[jj]) + sB3[c0 + jj], 0.f));
                            const int m0 = __reduce_max_sync(0xffffffffu, upper ? 0 : h);
                            const int m1 = __reduce_max_sync(0xffffffffu, upper ? h : 0);
                            keep0 = lane == jj ? m0 : keep0;
                            keep1 = lane == jj ? m1 : keep1;
                        }
                        sOut[(q * 2) * C3 + c0 + lane] = keep0;
                        sOut[(q * 2 + 1) * C3 + c0 + lane] = keep1;
                    }
                }
                tc_fence_before_sync();
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            for (int i = tid; i < TC * C3; i += kSaRowWarps * 32) {
                const int tc = i / C3, c = i % C3;
                feat_out[((size_t)b * C::NPOINT + c_base + tc) * C::CTOTAL + ch_off + c] = __int_as_float(sOut[i]);
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");     // sOut / sHit are reused by the next tile
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == kSaRowWarps) tmem_dealloc(tmem_base, 256);
}

template <class C>
static int launch_sa_small(const float *xyz_in, const float *new_xyz, const float *U, const float *consts, const void *wimg, float radius,
                           int ch_off, float *feat_out, int B, cudaStream_t st) {
    int dev = 0, sms = 0;
    GPB_CUDA(cudaGetDevice(&dev));
    GPB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    GPB_CUDA(cudaFuncSetAttribute(sa_small_tc_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::bytes));
    const int n_tiles = B * (C::NPOINT / C::TC);
    const int grid = n_tiles < 2 * sms ? n_tiles : 2 * sms;
    sa_small_tc_kernel<C><<<grid, kS2Threads, C::bytes, st>>>(xyz_in, new_xyz, U, consts, reinterpret_cast<const uint8_t *>(wimg), radius, ch_off,
                                                             n_tiles, feat_out);
    GPB_LAUNCHED();
    return GPB_OK;
}

using CfgL1S1 = SmallCfg<32, 32, 32, 64, 1024, 512, 96, false>;
using CfgL2S0 = SmallCfg<16, 64, 64, 128, 512, 256, 256, true>;
using CfgL2S1 = SmallCfg<32, 64, 96, 128, 512, 256, 256, true>;

int launch_sa1_tc(const float *pts, const float *new_xyz, const float *consts, const void *wimg, float *feat_out, int B, cudaStream_t st) {
    return launch_sa_small<CfgL1S1>(pts, new_xyz, nullptr, consts, wimg, 0.04f, 32, feat_out, B, st);
}

int launch_sa2_tc(const float *xyz_in, const float *new_xyz, const float *U, const float *consts, const void *wimg, int scale,
                  float *feat_out, int B, cudaStream_t st) {
    return scale == 0 ? launch_sa_small<CfgL2S0>(xyz_in, new_xyz, U, consts, wimg, 0.04f, 0, feat_out, B, st)
                      : launch_sa_small<CfgL2S1>(xyz_in, new_xyz, U, consts, wimg, 0.08f, 128, feat_out, B, st);
}

int launch_sa3_tc(const float *xyz_in, const float *new_xyz, const float *U, const float *consts, const void *wstream, int scale,
                  float *feat_out, uint8_t *a0_hi, uint8_t *a0_lo, int B, cudaStream_t st) {
    if (scale == 0) {
        GPB_CUDA(cudaFuncSetAttribute(sa3_tc_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSaSmemBytes));
        sa3_tc_kernel<16><<<dim3(kSaNPoint / 8, B), kSaThreads, kSaSmemBytes, st>>>(xyz_in, new_xyz, U, consts,
                                                                                   reinterpret_cast<const uint8_t *>(wstream), 0.08f, 0, feat_out,
                                                                                   a0_hi, a0_lo);
    } else {
        GPB_CUDA(cudaFuncSetAttribute(sa3_tc_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSaSmemBytes));
        sa3_tc_kernel<32><<<dim3(kSaNPoint / 4, B), kSaThreads, kSaSmemBytes, st>>>(xyz_in, new_xyz, U, consts,
                                                                                   reinterpret_cast<const uint8_t *>(wstream), 0.16f, 256, feat_out,
                                                                                   a0_hi, a0_lo);
    }
    GPB_LAUNCHED();
    return GPB_OK;
}

}  // namespace gpb

