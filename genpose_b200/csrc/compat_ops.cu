// Compat layer: the four forward ops of the reference's `pointnet2_cuda` extension behind the C ABI
// (include/genpose_b200.h §1).  Same tensor layouts as the reference so its pointnet2_utils.py can run
// on these unchanged; kernels are re-designed (whole cloud staged in shared memory, one key-max per
// FPS round instead of a 10-level __syncthreads tree, warp-ballot ordered compaction for ball query).
//
// Index parity (bit-exact, including ties):
//   FPS  — reference sampling_gpu.cu:93-209.  Thread tid of a block of bs = opt_n_threads(n)
//          (cuda_utils.h:10-14) owns k = tid, tid+bs, ...; strict '>' keeps the lowest k inside a
//          thread, and the block tree (__update :86-91, strides bs/2..1, lower slot wins ties) keeps,
//          among equal maxima, the tid whose log2(bs)-bit reversal is smallest.  We reduce the 64-bit
//          key  (float_bits(dist) << 32) | ~((bitrev(tid) << 21) | (k / bs))  with a plain max, which
//          is order-independent and selects the same element.
//   ball — reference ball_query_gpu.cu:9-45: first `nsample` k (ascending) with d2 < r2 (strict),
//          remaining slots = first hit; no hit -> zeros (pointnet2_utils.py:219 pre-zeroes).
#include "common.cuh"

namespace gpb {

static inline int opt_n_threads_host(int work_size) {
    int pow_2 = 0;
    while ((1 << (pow_2 + 1)) <= work_size) ++pow_2;   // == floor(log2(work_size)) for work_size >= 1
    int t = 1 << pow_2;
    if (t > 1024) t = 1024;
    if (t < 1) t = 1;
    return t;
}

__device__ __forceinline__ unsigned long long u64_max(unsigned long long a, unsigned long long b) {
    return a > b ? a : b;
}

__device__ __forceinline__ unsigned long long warp_max_u64(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = u64_max(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

constexpr int kFpsThreads = 512;

// One CTA per cloud.  xyz (SoA) and temp live in shared memory for the whole run.
__global__ void __launch_bounds__(kFpsThreads)
fps_compat_kernel(int n, int m, int bs, int log2bs, const float *__restrict__ xyz, float *__restrict__ temp_out,
                  int *__restrict__ idx_out) {
    extern __shared__ float smem[];
    float *sx = smem, *sy = sx + n, *sz = sy + n, *st = sz + n;
    __shared__ unsigned long long warp_best[2][kFpsThreads / 32];

    const int tid = threadIdx.x;
    const float *p = xyz + (size_t)blockIdx.x * n * 3;
    for (int k = tid; k < n; k += kFpsThreads) {
        sx[k] = p[k * 3 + 0];
        sy[k] = p[k * 3 + 1];
        sz[k] = p[k * 3 + 2];
        st[k] = 1e10f;   // pointnet2_utils.py:27
    }
    int *idx = idx_out + (size_t)blockIdx.x * m;
    if (tid == 0) idx[0] = 0;
    __syncthreads();

    int old = 0;
    for (int j = 1; j < m; ++j) {
        const float cx = sx[old], cy = sy[old], cz = sz[old];
        unsigned long long best = 0ull;
        for (int k = tid; k < n; k += kFpsThreads) {
            const float d = dist2_ref(sx[k], sy[k], sz[k], cx, cy, cz);
            const float t = fminf(d, st[k]);
            st[k] = t;
            const unsigned rtid = __brev((unsigned)(k & (bs - 1))) >> (32 - log2bs);   // log2bs >= 1 here
            const unsigned prio = (rtid << 21) | (unsigned)(k >> log2bs);
            const unsigned long long key =
                ((unsigned long long)__float_as_uint(t) << 32) | (unsigned long long)(0xffffffffu - prio);
            best = u64_max(best, key);
        }
        best = warp_max_u64(best);
        if ((tid & 31) == 0) warp_best[j & 1][tid >> 5] = best;
        __syncthreads();
        unsigned long long w = warp_best[j & 1][0];
#pragma unroll
        for (int i = 1; i < kFpsThreads / 32; ++i) w = u64_max(w, warp_best[j & 1][i]);
        const unsigned prio = 0xffffffffu - (unsigned)(w & 0xffffffffull);
        const unsigned rtid = prio >> 21;
        const unsigned wtid = __brev(rtid) >> (32 - log2bs);
        old = (int)(((prio & 0x1fffffu) << log2bs) | wtid);
        if (tid == 0) idx[j] = old;
    }
    if (temp_out != nullptr) {
        __syncthreads();
        float *t = temp_out + (size_t)blockIdx.x * n;
        for (int k = tid; k < n; k += kFpsThreads) t[k] = st[k];
    }
}

// out[b,c,j] = points[b,c,idx[b,j]]
__global__ void gather_points_kernel(int c, int n, int npoints, const float *__restrict__ points,
                                     const int *__restrict__ idx, float *__restrict__ out) {
    const int b = blockIdx.z, ci = blockIdx.y;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= npoints) return;
    out[((size_t)b * c + ci) * npoints + j] = points[((size_t)b * c + ci) * n + idx[(size_t)b * npoints + j]];
}

// One warp per centre; the cloud is staged in shared memory once per CTA; the scan proceeds 32 points
// at a time and a ballot + popc prefix gives each hit its slot, preserving ascending-k order.
constexpr int kBqWarps = 8;

__global__ void __launch_bounds__(kBqWarps * 32)
ball_query_kernel(int n, int m, float radius2, int nsample, const float *__restrict__ new_xyz,
                  const float *__restrict__ xyz, int *__restrict__ idx, int n_smem) {
    extern __shared__ float smem[];
    const int b = blockIdx.y;
    const float *p = xyz + (size_t)b * n * 3;
    // stage as much of the cloud as fits (n_smem points), AoS copy is contiguous -> coalesced
    for (int i = threadIdx.x; i < n_smem * 3; i += blockDim.x) smem[i] = p[i];
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int pi = blockIdx.x * kBqWarps + warp;
    if (pi >= m) return;
    const float *c = new_xyz + ((size_t)b * m + pi) * 3;
    const float cx = c[0], cy = c[1], cz = c[2];
    int *o = idx + ((size_t)b * m + pi) * nsample;

    int cnt = 0, first = 0;
    for (int base = 0; base < n && cnt < nsample; base += 32) {
        const int k = base + lane;
        bool hit = false;
        if (k < n) {
            const float *q = (k < n_smem) ? (smem + k * 3) : (p + k * 3);
            hit = dist2_ref(cx, cy, cz, q[0], q[1], q[2]) < radius2;
        }
        const unsigned mask = __ballot_sync(0xffffffffu, hit);
        if (mask) {
            if (cnt == 0) first = base + __ffs(mask) - 1;
            const int slot = cnt + __popc(mask & ((1u << lane) - 1u));
            if (hit && slot < nsample) o[slot] = k;
            cnt += __popc(mask);
        }
    }
    if (cnt > nsample) cnt = nsample;
    const int fill = (cnt == 0) ? 0 : first;   // empty ball: zeros, as left by the reference's caller
    for (int s = cnt + lane; s < nsample; s += 32) o[s] = fill;
}

// out[b,c,i,s] = points[b,c,idx[b,i,s]]
__global__ void group_points_kernel(int c, int n, int total /* npoints*nsample */, const float *__restrict__ points,
                                    const int *__restrict__ idx, float *__restrict__ out) {
    const int b = blockIdx.z, ci = blockIdx.y;
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= total) return;
    out[((size_t)b * c + ci) * total + q] = points[((size_t)b * c + ci) * n + idx[(size_t)b * total + q]];
}

}  // namespace gpb

using namespace gpb;

extern "C" int gpb_furthest_point_sampling(int b, int n, int m, const float *xyz, float *temp, int *idx,
                                           void *stream) {
    GPB_REQUIRE(b >= 0 && n >= 1 && m >= 0 && m <= n, "fps: need b>=0, 1<=n, 0<=m<=n (b=%d n=%d m=%d)", b, n, m);
    if (b == 0 || m == 0) return GPB_OK;
    GPB_REQUIRE(xyz && idx, "fps: NULL buffer");
    const size_t smem = (size_t)n * 4 * sizeof(float);
    GPB_REQUIRE(smem <= 200 * 1024, "fps: n=%d exceeds the shared-memory resident limit (12800 points)", n);
    const int bs = opt_n_threads_host(n);
    int log2bs = 0;
    while ((1 << log2bs) < bs) ++log2bs;
    if (n == 1) {   // bs == 1: every round re-selects index 0
        GPB_CUDA(cudaMemsetAsync(idx, 0, (size_t)b * m * sizeof(int), (cudaStream_t)stream));
        return GPB_OK;
    }
    GPB_CUDA(cudaFuncSetAttribute(fps_compat_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    fps_compat_kernel<<<b, kFpsThreads, smem, (cudaStream_t)stream>>>(n, m, bs, log2bs, xyz, temp, idx);
    GPB_LAUNCHED();
    return GPB_OK;
}

extern "C" int gpb_gather_points(int b, int c, int n, int npoints, const float *points, const int *idx, float *out,
                                 void *stream) {
    GPB_REQUIRE(b >= 0 && c >= 0 && n >= 1 && npoints >= 0, "gather_points: bad shape");
    if (b == 0 || c == 0 || npoints == 0) return GPB_OK;
    GPB_REQUIRE(points && idx && out, "gather_points: NULL buffer");
    GPB_REQUIRE(c <= 65535 && b <= 65535, "gather_points: c and b must be <= 65535");
    dim3 grid((npoints + 255) / 256, c, b);
    gather_points_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(c, n, npoints, points, idx, out);
    GPB_LAUNCHED();
    return GPB_OK;
}

extern "C" int gpb_ball_query(int b, int n, int m, float radius, int nsample, const float *new_xyz, const float *xyz,
                              int *idx, void *stream) {
    GPB_REQUIRE(b >= 0 && n >= 1 && m >= 0 && nsample >= 1, "ball_query: bad shape");
    if (b == 0 || m == 0) return GPB_OK;
    GPB_REQUIRE(new_xyz && xyz && idx, "ball_query: NULL buffer");
    GPB_REQUIRE(b <= 65535, "ball_query: b must be <= 65535");
    const float radius2 = radius * radius;   // fp32, ball_query_gpu.cu:23
    const int n_smem = n < 16384 ? n : 16384;
    const size_t smem = (size_t)n_smem * 3 * sizeof(float);
    GPB_CUDA(cudaFuncSetAttribute(ball_query_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((m + kBqWarps - 1) / kBqWarps, b);
    ball_query_kernel<<<grid, kBqWarps * 32, smem, (cudaStream_t)stream>>>(n, m, radius2, nsample, new_xyz, xyz, idx,
                                                                         n_smem);
    GPB_LAUNCHED();
    return GPB_OK;
}

extern "C" int gpb_group_points(int b, int c, int n, int npoints, int nsample, const float *points, const int *idx,
                                float *out, void *stream) {
    GPB_REQUIRE(b >= 0 && c >= 0 && n >= 1 && npoints >= 0 && nsample >= 0, "group_points: bad shape");
    const int total = npoints * nsample;
    if (b == 0 || c == 0 || total == 0) return GPB_OK;
    GPB_REQUIRE(points && idx && out, "group_points: NULL buffer");
    GPB_REQUIRE(c <= 65535 && b <= 65535, "group_points: c and b must be <= 65535");
    dim3 grid((total + 255) / 256, c, b);
    group_points_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(c, n, total, points, idx, out);
    GPB_LAUNCHED();
    return GPB_OK;
}
