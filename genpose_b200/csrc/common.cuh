// Shared device/host helpers for the genpose_b200 kernels (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../include/genpose_b200.h"

namespace gpb {

// ---- error plumbing -----------------------------------------------------------------------------
void set_error(const char *fmt, ...);
extern std::atomic<uint64_t> g_launches;

#define GPB_REQUIRE(cond, ...)                \
    do {                                      \
        if (!(cond)) {                        \
            ::gpb::set_error(__VA_ARGS__);    \
            return GPB_EINVAL;                \
        }                                     \
    } while (0)

#define GPB_CUDA(expr)                                                                   \
    do {                                                                                 \
        cudaError_t _e = (expr);                                                         \
        if (_e != cudaSuccess) {                                                         \
            ::gpb::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return GPB_ECUDA;                                                            \
        }                                                                                \
    } while (0)

// call after every <<<>>> launch
#define GPB_LAUNCHED()                         \
    do {                                       \
        ::gpb::g_launches.fetch_add(1);        \
        GPB_CUDA(cudaGetLastError());          \
    } while (0)

// ---- network constants (restated from the reference; see genpose_b200/arch.py for citations) ------
constexpr int kPoseDim = 9;
constexpr int kPtsFeat = 1024;
constexpr int kTEmbed = 128;
constexpr int kPoseFeat = 256;
constexpr int kHeadHidden = 256;
constexpr int kHeads = 3;
constexpr int kHeadCols = kHeads * kHeadHidden;   // 768 stacked hidden units
constexpr float kSigmaMin = 0.01f;
constexpr float kSigmaRatio = 5000.0f;            // sigma_max / sigma_min = 50 / 0.01 (sde.py:90-97)
constexpr float kSamplingEps = 1e-5f;
// sqrt(2 * (ln 50 - ln 0.01)) rounded to fp32 (sde.py:23 — a float64 0-dim tensor multiplying fp32 sigma)
constexpr float kGCoef = 4.12727348049926f;

// ---- packed trunk weights (score / energy net) — float offsets -------------------------------------
// All matrices are stored K-MAJOR ([in][out], out contiguous) so that consecutive threads read
// consecutive output columns.  See DESIGN.md §3.
struct TrunkLayout {
    static constexpr size_t fourier_w = 0;                                   // [64]
    static constexpr size_t t_w = fourier_w + 64;                            // [128 in][128 out]
    static constexpr size_t t_b = t_w + 128 * 128;                           // [128]
    static constexpr size_t p1_w = t_b + 128;                                // [9 in][256 out]
    static constexpr size_t p1_b = p1_w + 9 * 256;                           // [256]
    static constexpr size_t p2_w = p1_b + 256;                               // [256 in][256 out]
    static constexpr size_t p2_b = p2_w + 256 * 256;                         // [256]
    static constexpr size_t a_pts = p2_b + 256;                              // [1024 in][768 out]
    static constexpr size_t a_t = a_pts + 1024 * 768;                        // [128 in][768 out]
    static constexpr size_t a_pose = a_t + 128 * 768;                        // [256 in][768 out]
    static constexpr size_t a_b = a_pose + 256 * 768;                        // [768]
    static constexpr size_t o_w = a_b + 768;                                 // [9 out][256 in]  (row c uses head c/3)
    static constexpr size_t o_b = o_w + 9 * 256;                             // [9] (+3 pad)
    static constexpr size_t total = o_b + 12;
};

// ---- small device helpers ---------------------------------------------------------------------------
__device__ __forceinline__ float sigma_of_t(float t) {
    // ve_marginal_prob (sde.py:15-18): sigma_min * (sigma_max/sigma_min) ** t, fp32 powf like torch.pow
    return kSigmaMin * powf(kSigmaRatio, t);
}

// squared distance exactly as the reference kernels compile: mul, fma, fma on rounded
// fp32 differences, x then y then z.
__device__ __forceinline__ float dist2_ref(float ax, float ay, float az, float bx, float by, float bz) {
    const float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by), dz = __fsub_rn(az, bz);
    return __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ void st_relaxed_gpu_u64(unsigned long long *p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_gpu_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// Monotonic-counter grid barrier.  All CTAs of the (cooperatively launched, hence co-resident) grid
// call it the same number of times; `target` = (#barriers so far + 1) * gridDim.x.
__device__ __forceinline__ void grid_barrier(unsigned *counter, unsigned target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(counter, 1u);
        while (ld_acquire_u32(counter) < target) {
        }
        __threadfence();
    }
    __syncthreads();
}

// cp.async (LDGSTS) 16-byte copy global -> shared
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// ---- Philox4x32-10 + Box-Muller (throughput-mode noise; parity mode reads explicit noise) -----------
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
    constexpr unsigned M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const unsigned hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
        const unsigned hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += W0;
        key.y += W1;
    }
    return ctr;
}

__device__ __forceinline__ float2 box_muller(unsigned a, unsigned b) {
    // u1 in (0,1], u2 in [0,1)
    const float u1 = ((float)(a >> 8) + 1.0f) * (1.0f / 16777216.0f);
    const float u2 = (float)(b >> 8) * (1.0f / 16777216.0f);
    const float r = sqrtf(-2.0f * logf(u1));
    float s, c;
    sincospif(2.0f * u2, &s, &c);
    return make_float2(r * c, r * s);
}

}  // namespace gpb
