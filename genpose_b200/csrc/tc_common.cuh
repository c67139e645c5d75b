// Blackwell (sm_100a) tensor-core plumbing used by the tcgen05 sampler: raw PTX wrappers for mbarrier,
// bulk async copies (UBLKCP), TMEM allocation, tcgen05.mma / commit / ld, and the shared-memory matrix
// descriptors for the canonical K-major, non-swizzled ("interleave") operand layout.
//
// Operand layout (both A [rows x K] and B [cols x K], bf16, K-major), in bytes from the operand base:
//     off(r, k) = (k / 8) * LBO  +  (r / 8) * SBO  +  (r % 8) * 16  +  (k % 8) * 2
// i.e. 8x8 "core matrices" of 128 contiguous bytes; SBO = 128 walks 8-row groups, LBO = (rows/8)*128 walks
// 8-element K chunks (CuTe: Layout_K_INTER_Atom, ((8,n),2):((1,SBO),LBO) in 16-byte units,
// cute/atom/mma_traits_sm100.hpp).  One tcgen05.mma (kind::f16) consumes K = 16 = two K chunks, so stepping
// K by 16 advances the descriptor start address by 2*LBO.  Consecutive rows of one K chunk are consecutive
// 16-byte words => row-per-thread epilogues write the A operand with conflict-free STS.128.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

#include <type_traits>

namespace gpb {
namespace tc {

// compile-time loop: f(std::integral_constant<int, I>{}) for I in [I0, N)
template <int I, int N, class F>
__device__ __forceinline__ void static_for(F &&f) {
    if constexpr (I < N) {
        f(std::integral_constant<int, I>{});
        static_for<I + 1, N>(f);
    }
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier -------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// The same with the retry loop inside the asm block: the C++ loop above is compiled as a jump to an out-of-line retry block at the
// end of the kernel and a jump back — two taken branches into cold instruction-cache lines even when the phase is already complete.
// For the MMA issuer, where every cycle between two issue groups idles the tensor pipe.
__device__ __forceinline__ void mbar_wait_inline(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t}"
        ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}

// ---- thread-block cluster: distributed shared memory + remote mbarrier arrival ------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t mapa_shared(uint32_t local_smem_addr, uint32_t cta_rank) {   // same offset in CTA `cta_rank`
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(cta_rank));
    return r;
}
__device__ __forceinline__ void st_cluster_f4(uint32_t cluster_addr, float a, float b, float c, float d) {
    asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(cluster_addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void st_cluster_f2(uint32_t cluster_addr, float a, float b) {
    asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(cluster_addr), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_bar_addr) {   // release at cluster scope: prior DSMEM stores visible
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar_addr) : "memory");
}
// asynchronous DSMEM store that completes `bytes` on the DESTINATION CTA's mbarrier (transaction count): producer/consumer
// hand-off without any cluster-scope fence
__device__ __forceinline__ void st_async_f4(uint32_t cluster_addr, float a, float b, float c, float d, uint32_t cluster_bar_addr) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(cluster_addr),
                 "f"(a), "f"(b), "f"(c), "f"(d), "r"(cluster_bar_addr)
                 : "memory");
}
__device__ __forceinline__ void st_async_f2(uint32_t cluster_addr, float a, float b, uint32_t cluster_bar_addr) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f32 [%0], {%1, %2}, [%3];" ::"r"(cluster_addr), "f"(a),
                 "f"(b), "r"(cluster_bar_addr)
                 : "memory");
}
__device__ __forceinline__ void fence_acq_rel_cluster() { asm volatile("fence.acq_rel.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_cluster_relaxed(uint32_t cluster_bar_addr) {   // pair with ONE fence_acq_rel_cluster() before a batch
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t *bar, uint32_t parity) {   // acquire at cluster scope
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.test_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"   // spin: try_wait's suspend is slow to notice REMOTE arrivals
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ void cluster_sync_all() {   // every thread of every CTA of the cluster
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// ---- bulk async copy global -> shared (contiguous bytes, completes on an mbarrier) -----------------------
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// generic-proxy smem writes -> visible to the async proxy (tcgen05.mma operand reads, bulk copies)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- TMEM ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t *smem_result, uint32_t ncols) {   // one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {       // same warp that allocated
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// 32 lanes x 32 consecutive fp32 columns: thread t of warp w reads TMEM lane 32*(w%4)+t, columns [col, col+32)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
#if defined(GPB_DBG_SKIP_LDTM) && GPB_DBG_SKIP_LDTM
    // TIMING EXPERIMENT ONLY (wrong results): no accumulator read-back, to see what the tensor-memory read path costs a step
#pragma unroll
    for (int i = 0; i < 32; ++i) r[i] = taddr + i;
    return;
#endif
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// registers -> TMEM: thread t of warp w writes lane 32*(w%4)+t, 32 (or 8) consecutive 32-bit columns
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
          "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]),
          "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]),
          "r"(r[30]), "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
                 "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
                 "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
                 : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]),
                 "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- descriptors ----------------------------------------------------------------------------------------------
// shared-memory matrix descriptor (cute/arch/mma_sm100_desc.hpp SmemDescriptor): start>>4 [0,14), LBO>>4 [16,30),
// SBO>>4 [32,46), version=1 [46,48), base_offset 0, lbo_mode 0, layout_type SWIZZLE_NONE=0 [61,64)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
// instruction descriptor, kind::f16: D=f32 (c_format 1 @[4,6)), A=B=bf16 (1 @[7,10), 1 @[10,13)), both K-major
// (0 @15, 0 @16), N>>3 @[17,23), M>>4 @[24,29)   (cute/arch/mma_sm100_desc.hpp InstrDescriptor)
__host__ __device__ constexpr uint32_t make_idesc_bf16_f32(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// the same with the operand formats chosen separately (kind::f16: 0 = fp16, 1 = bf16): e.g. bf16 activations against fp16 weights
__host__ __device__ constexpr uint32_t make_idesc_f16kind_f32(int M, int N, uint32_t a_fmt, uint32_t b_fmt) {
    return (1u << 4) | (a_fmt << 7) | (b_fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// D[tmem] (+)= A[smem] . B[smem]^T   — issued by ONE thread on behalf of the CTA
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
        : "memory");
}
// D[tmem] (+)= A[tmem] . B[smem]^T : A operand read from tensor memory (lane = row, 32-bit column c holds K elements 2c, 2c+1;
// one instruction consumes K = 16 = 8 columns).  cute SM100_MMA_F16BF16_TS.
__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
        : "memory");
}
// arrive on `bar` once every tcgen05.mma issued so far by this thread has completed (implies fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// one lane of a fully converged warp (CUTLASS elect_one_sync): keeps the surrounding control flow warp-uniform so that
// descriptors / TMEM addresses stay in uniform registers and UTCHMMA needs no per-instruction uniformisation loop
__device__ __forceinline__ bool elect_one_sync() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}

// ---- packed fp32 pairs (Blackwell FADD2 / FFMA2: one issue slot for two lanes of fp32 math) ------------------------------------------
// The row warps' epilogues are issue-bound (two warps per scheduler, ~600-700 instructions per thread and layer), so halving the
// instruction count of their adds and multiply-adds shortens the step's critical path.  A pair lives in one 64-bit register: x = low.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack_f32x2(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack_f32x2(f32x2 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 add_f32x2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 sub_f32x2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 fma_f32x2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

// ---- bf16 hi/lo splitting ----------------------------------------------------------------------------------------
// x = hi + lo + O(2^-17 |x|), both bf16: hi = rn(x), lo = rn(x - hi)
__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16 &hi, __nv_bfloat16 &lo) {
    hi = __float2bfloat16_rn(x);
    lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}
__device__ __forceinline__ uint32_t pack_bf16(__nv_bfloat16 a, __nv_bfloat16 b) {   // a = low half (lower k)
    return (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(b) << 16);
}
// two values at once through the packed converter (F2FP.BF16.F32.PACK_AB): hi = (rn(a) | rn(b) << 16), lo likewise for the residuals
__device__ __forceinline__ void split_bf16x2(float a, float b, uint32_t &hi, uint32_t &lo) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);            // .x = a (low half), .y = b
    hi = *reinterpret_cast<const uint32_t *>(&h);
    float ra, rb;
    unpack_f32x2(sub_f32x2(pack_f32x2(a, b), pack_f32x2(__uint_as_float(hi << 16), __uint_as_float(hi & 0xffff0000u))), ra, rb);
    const __nv_bfloat162 l = __floats2bfloat162_rn(ra, rb);
    lo = *reinterpret_cast<const uint32_t *>(&l);
}

// fp16 halves of a ReLU'd activation for the two-product layers (A = hi + lo, both fp16, against ONE fp16 weight image):
// hi = fp16 TRUNCATION of max(x, 0) (cvt.rz.relu: for x >= 0 the residual x - hi is then >= 0; for x < 0 hi = 0 and the residual
// x < 0 is removed by the second relu), lo = rn_fp16(max(residual, 0)) (cvt.rn.relu); hi + lo carries 21 mantissa bits.
// .satfinite clamps an activation beyond the fp16 range (65504) instead of producing inf.
__device__ __forceinline__ void relu_split_f16x2(float a, float b, uint32_t &hi, uint32_t &lo) {
    asm("cvt.rz.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(b), "f"(a));      // d = {upper half: first source, lower half: second}
    const float2 h = __half22float2(*reinterpret_cast<const __half2 *>(&hi));
    float ra, rb;
    unpack_f32x2(sub_f32x2(pack_f32x2(a, b), pack_f32x2(h.x, h.y)), ra, rb);
    asm("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(rb), "f"(ra));
}

}  // namespace tc
}  // namespace gpb
