// PointNet++ MSG encoder of GenPose (Pointnet2ClsMSG, networks/pts_encoder/pointnet2.py:166-211 with
// ClsMSG_CFG_Light :57-66) as four fused stages; nothing of shape (B, C, npoint, nsample) is ever
// materialised (the reference writes ~1 GB of grouped tensors per 64-object batch).
//
//   fps3_kernel        one CTA per object runs the three chained furthest-point samplings
//                      (1024->512->256->128) out of shared memory / registers and emits new_xyz per level.
//   point_gemm_kernel  per SOURCE point: U = W1_feat . f_j  (layer 1 hoisted out of the (centre,
//                      neighbour) pairs: W1.[x_j - c_i ; f_j] = W1_x.(x_j - c_i) + W1_feat.f_j).
//   sa_kernel          per tile of centres: warp-ballot ball query -> neighbour list in smem ->
//                      h1 = relu(U[nbr] + W1_x.(x_nbr - c) + b1) staged K-major in smem -> layer 2 ->
//                      layer 3 -> max over the neighbourhood -> feats [B, npoint, C] (point-major, so
//                      the next level's gathers are contiguous float4 rows).
//   groupall_kernel    SA4 (GroupAll, pointnet2_utils.py:273-291): 3-layer MLP over the 128 L3 points with
//                      ABSOLUTE xyz, max over points (atomicMax on non-negative floats).
//
// BatchNorm (eval) is folded into W and b on the host (genpose_b200/weights.py); all math is fp32 FFMA.
#include <cstdlib>
#include <mutex>

#include "common.cuh"

namespace gpb {

// ---------------------------------------------------------------------------------------------------
// packed encoder weights
// ---------------------------------------------------------------------------------------------------
struct MlpSpec {
    int cin_f, c1, c2, c3;   // channels, padded to multiples of 8 (196 -> 200)
};
__host__ __device__ constexpr MlpSpec enc_spec(int l, int s) {
    return l == 0 ? (s == 0 ? MlpSpec{0, 16, 16, 32} : MlpSpec{0, 32, 32, 64})
         : l == 1 ? (s == 0 ? MlpSpec{96, 64, 64, 128} : MlpSpec{96, 64, 96, 128})
         : l == 2 ? MlpSpec{256, 128, 200, 256}
                  : (s == 0 ? MlpSpec{512, 256, 256, 512} : MlpSpec{512, 256, 384, 512});
}
__host__ __device__ constexpr size_t spec_floats(MlpSpec m) {
    return (size_t)3 * m.c1 + (size_t)m.cin_f * m.c1 + m.c1 + (size_t)m.c1 * m.c2 + m.c2 + (size_t)m.c2 * m.c3 + m.c3;
}
__host__ __device__ constexpr size_t spec_offset(int l, int s) {
    size_t off = 0;
    for (int ll = 0; ll < 4; ++ll)
        for (int ss = 0; ss < 2; ++ss) {
            if (ll == l && ss == s) return off;
            off += spec_floats(enc_spec(ll, ss));
        }
    return off;
}
constexpr size_t kEncoderFloats = spec_offset(3, 1) + spec_floats(enc_spec(3, 1));
// sub-offsets inside one (level, scale) block; every matrix is K-major [in][out]
__host__ __device__ constexpr size_t off_wx(MlpSpec) { return 0; }
__host__ __device__ constexpr size_t off_wf(MlpSpec m) { return (size_t)3 * m.c1; }
__host__ __device__ constexpr size_t off_b1(MlpSpec m) { return off_wf(m) + (size_t)m.cin_f * m.c1; }
__host__ __device__ constexpr size_t off_w2(MlpSpec m) { return off_b1(m) + m.c1; }
__host__ __device__ constexpr size_t off_b2(MlpSpec m) { return off_w2(m) + (size_t)m.c1 * m.c2; }
__host__ __device__ constexpr size_t off_w3(MlpSpec m) { return off_b2(m) + m.c2; }
__host__ __device__ constexpr size_t off_b3(MlpSpec m) { return off_w3(m) + (size_t)m.c2 * m.c3; }

// ---------------------------------------------------------------------------------------------------
// fps3_kernel
// ---------------------------------------------------------------------------------------------------
constexpr int kFps3Threads = 256;

// One furthest-point-sampling level over N points held P-per-thread in registers (k = tid + 256*p).
// Tie rule = reference block tree: among equal maxima the smallest bit-reversed index wins (see
// compat_ops.cu header).  One __syncthreads per round; the winner's coordinates are read from smem.
template <int N, int M>
__device__ __forceinline__ void fps_level(const float *sx, const float *sy, const float *sz,   // smem SoA, N points
                                          float *ox, float *oy, float *oz,                      // smem SoA out, M points
                                          int *idx_out /* global or nullptr */, unsigned long long (*warp_best)[8]) {
    constexpr int P = N / kFps3Threads;
    static_assert(P >= 1, "fps_level: N must be >= 256");
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float px[P], py[P], pz[P], pt[P];
    unsigned pk[P];
#pragma unroll
    for (int p = 0; p < P; ++p) {
        const int k = tid + kFps3Threads * p;
        px[p] = sx[k];
        py[p] = sy[k];
        pz[p] = sz[k];
        pt[p] = 1e10f;
        pk[p] = ~__brev((unsigned)k);   // larger = higher priority on ties
    }
    if (tid == 0) {
        ox[0] = sx[0];
        oy[0] = sy[0];
        oz[0] = sz[0];
        if (idx_out) idx_out[0] = 0;
    }
    float cx = sx[0], cy = sy[0], cz = sz[0];
    for (int j = 1; j < M; ++j) {
        unsigned bd = 0u, bk = 0u;
#pragma unroll
        for (int p = 0; p < P; ++p) {
            const float t = fminf(dist2_ref(px[p], py[p], pz[p], cx, cy, cz), pt[p]);
            pt[p] = t;
            const unsigned d = __float_as_uint(t);   // t >= 0: uint order == float order
            const bool better = (d > bd) || (d == bd && pk[p] > bk);
            bd = better ? d : bd;
            bk = better ? pk[p] : bk;
        }
        const unsigned wd = __reduce_max_sync(0xffffffffu, bd);
        const unsigned wk = __reduce_max_sync(0xffffffffu, bd == wd ? bk : 0u);
        if (lane == 0) warp_best[j & 1][warp] = ((unsigned long long)wd << 32) | wk;
        __syncthreads();
        unsigned long long w = warp_best[j & 1][0];
#pragma unroll
        for (int i = 1; i < kFps3Threads / 32; ++i) {
            const unsigned long long o = warp_best[j & 1][i];
            w = o > w ? o : w;
        }
        const int win = (int)__brev(~(unsigned)(w & 0xffffffffull));
        cx = sx[win];
        cy = sy[win];
        cz = sz[win];
        if (tid == 0) {
            ox[j] = cx;
            oy[j] = cy;
            oz[j] = cz;
            if (idx_out) idx_out[j] = win;
        }
    }
    __syncthreads();
}

// One furthest-point-sampling level with ONE point per thread (threads >= N idle but take part in the barrier): the per-round
// dependent chain is one distance + one min instead of P of them, and the cross-warp step is a second redux over the (<= 32) warp
// winners read one per lane instead of every thread scanning all of them.  Same keys, same tie rule, same winner as fps_level.
template <int N, int M, int NT>
__device__ __forceinline__ void fps_level_1pt(const float *sx, const float *sy, const float *sz, float *ox, float *oy, float *oz,
                                              int *idx_out, unsigned long long (*warp_best)[32]) {
    static_assert(N <= NT && NT <= 1024 && NT % 32 == 0, "fps_level_1pt: one point per thread");
    constexpr int NW = NT / 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool mine = tid < N;
    const float px = mine ? sx[tid] : 0.f, py = mine ? sy[tid] : 0.f, pz = mine ? sz[tid] : 0.f;
    float pt = 1e10f;
    const unsigned pk = mine ? ~__brev((unsigned)tid) : 0u;   // larger = higher priority on ties; an idle thread's key (0, 0) never wins
    if (tid == 0) {
        ox[0] = sx[0];
        oy[0] = sy[0];
        oz[0] = sz[0];
        if (idx_out) idx_out[0] = 0;
    }
    float cx = sx[0], cy = sy[0], cz = sz[0];
    for (int j = 1; j < M; ++j) {
        pt = fminf(dist2_ref(px, py, pz, cx, cy, cz), pt);
        const unsigned d = mine ? __float_as_uint(pt) : 0u;   // pt >= 0: uint order == float order
        const unsigned wd = __reduce_max_sync(0xffffffffu, d);
        const unsigned wk = __reduce_max_sync(0xffffffffu, d == wd ? pk : 0u);
        if (lane == 0) warp_best[j & 1][warp] = ((unsigned long long)wd << 32) | wk;
        __syncthreads();
        const unsigned long long v = lane < NW ? warp_best[j & 1][lane] : 0ull;
        const unsigned hi = (unsigned)(v >> 32), lo = (unsigned)(v & 0xffffffffull);
        const unsigned gd = __reduce_max_sync(0xffffffffu, hi);
        const unsigned gk = __reduce_max_sync(0xffffffffu, hi == gd ? lo : 0u);
        const int win = (int)__brev(~gk);
        cx = sx[win];
        cy = sy[win];
        cz = sz[win];
        if (tid == 0) {
            ox[j] = cx;
            oy[j] = cy;
            oz[j] = cz;
            if (idx_out) idx_out[j] = win;
        }
    }
    __syncthreads();
}

// The same chain as two kernels, so that levels 2 and 3 (384 of the 896 dependent rounds) can run on a side stream BESIDE level 1's
// set abstraction, which needs new_xyz1 only: fps_l1_kernel = level 1 (1024 threads, one point each), fps_l23_kernel = levels 2 and 3
// from new_xyz1 (512 threads).
__global__ void __launch_bounds__(1024)
fps_l1_kernel(const float *__restrict__ pts /* [B,1024,3] */, float *__restrict__ nx1, int *__restrict__ idx1) {
    __shared__ float s0[3][1024], s1[3][512];
    __shared__ unsigned long long warp_best[2][32];
    const int b = blockIdx.x, tid = threadIdx.x;
    const float *p = pts + (size_t)b * 1024 * 3;
    for (int i = tid; i < 1024 * 3; i += 1024) s0[i % 3][i / 3] = p[i];
    __syncthreads();
    fps_level_1pt<1024, 512, 1024>(s0[0], s0[1], s0[2], s1[0], s1[1], s1[2], idx1 ? idx1 + (size_t)b * 512 : nullptr, warp_best);
    for (int i = tid; i < 512 * 3; i += 1024) nx1[(size_t)b * 512 * 3 + i] = s1[i % 3][i / 3];
}
__global__ void __launch_bounds__(512)
fps_l23_kernel(const float *__restrict__ nx1 /* [B,512,3] */, float *__restrict__ nx2, float *__restrict__ nx3, int *__restrict__ idx2,
               int *__restrict__ idx3) {
    __shared__ float s1[3][512], s2[3][256], s3[3][128];
    __shared__ unsigned long long warp_best[2][32];
    const int b = blockIdx.x, tid = threadIdx.x;
    const float *p = nx1 + (size_t)b * 512 * 3;
    for (int i = tid; i < 512 * 3; i += 512) s1[i % 3][i / 3] = p[i];
    __syncthreads();
    fps_level_1pt<512, 256, 512>(s1[0], s1[1], s1[2], s2[0], s2[1], s2[2], idx2 ? idx2 + (size_t)b * 256 : nullptr, warp_best);
    fps_level_1pt<256, 128, 512>(s2[0], s2[1], s2[2], s3[0], s3[1], s3[2], idx3 ? idx3 + (size_t)b * 128 : nullptr, warp_best);
    for (int i = tid; i < 256 * 3; i += 512) nx2[(size_t)b * 256 * 3 + i] = s2[i % 3][i / 3];
    for (int i = tid; i < 128 * 3; i += 512) nx3[(size_t)b * 128 * 3 + i] = s3[i % 3][i / 3];
}

__global__ void __launch_bounds__(kFps3Threads)
fps3_kernel(const float *__restrict__ pts /* [B,1024,3] */, float *__restrict__ nx1, float *__restrict__ nx2,
            float *__restrict__ nx3, int *__restrict__ idx1, int *__restrict__ idx2, int *__restrict__ idx3) {
    __shared__ float s0[3][1024], s1[3][512], s2[3][256], s3[3][128];
    __shared__ unsigned long long warp_best[2][8];
    const int b = blockIdx.x, tid = threadIdx.x;
    const float *p = pts + (size_t)b * 1024 * 3;
    for (int i = tid; i < 1024 * 3; i += kFps3Threads) s0[i % 3][i / 3] = p[i];
    __syncthreads();
    fps_level<1024, 512>(s0[0], s0[1], s0[2], s1[0], s1[1], s1[2], idx1 ? idx1 + (size_t)b * 512 : nullptr, warp_best);
    fps_level<512, 256>(s1[0], s1[1], s1[2], s2[0], s2[1], s2[2], idx2 ? idx2 + (size_t)b * 256 : nullptr, warp_best);
    fps_level<256, 128>(s2[0], s2[1], s2[2], s3[0], s3[1], s3[2], idx3 ? idx3 + (size_t)b * 128 : nullptr, warp_best);
    for (int i = tid; i < 512 * 3; i += kFps3Threads) nx1[(size_t)b * 512 * 3 + i] = s1[i % 3][i / 3];
    for (int i = tid; i < 256 * 3; i += kFps3Threads) nx2[(size_t)b * 256 * 3 + i] = s2[i % 3][i / 3];
    for (int i = tid; i < 128 * 3; i += kFps3Threads) nx3[(size_t)b * 128 * 3 + i] = s3[i % 3][i / 3];
}

// ---------------------------------------------------------------------------------------------------
// dense_layer: acc[ROWS x N] = At^T[ROWS x K] . W[K x N], everything but W resident in shared memory.
//   At      smem, K-major: element (row r, k) at At[k * PITCH + r]  (PITCH % 4 == 0)
//   Wg      global, K-major [K][N]; streamed through `wstage` (2 x KC x N floats) with cp.async
//   thread  (rg = tid % RG, cg = tid / RG) owns rows rg*TM.. and columns cg*TN..; lanes of a warp run
//           over rg first, so A fragments are contiguous across lanes and W fragments broadcast.
//   epi(r0, c0, acc) consumes the TM x TN accumulator tile.
// All NT threads of the CTA must call it (it contains __syncthreads).
// ---------------------------------------------------------------------------------------------------
constexpr int KC = 8;

template <int ROWS, int PITCH, int K, int N, int TM, int TN, int NT, class Epi>
__device__ __forceinline__ void dense_layer(const float *At, const float *__restrict__ Wg, float *wstage, Epi epi) {
    static_assert(ROWS % TM == 0 && N % TN == 0 && K % KC == 0 && TM % 4 == 0 && TN % 4 == 0 && PITCH % 4 == 0, "tile");
    constexpr int RG = ROWS / TM, CG = N / TN;
    static_assert(RG * CG <= NT, "not enough threads for this layer");
    constexpr int CHUNK_F4 = KC * N / 4;
    const int tid = threadIdx.x;
    const bool active = tid < RG * CG;
    const int rg = tid % RG, cg = tid / RG;
    const int r0 = rg * TM, c0 = cg * TN;

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    auto load_chunk = [&](int chunk, int buf) {
        const float4 *src = reinterpret_cast<const float4 *>(Wg + (size_t)chunk * KC * N);
        float4 *dst = reinterpret_cast<float4 *>(wstage + buf * KC * N);
        for (int i = tid; i < CHUNK_F4; i += NT) cp_async16(dst + i, src + i);
        cp_async_commit();
    };

    constexpr int NCHUNK = K / KC;
    load_chunk(0, 0);
    for (int ch = 0; ch < NCHUNK; ++ch) {
        if (ch + 1 < NCHUNK) {
            load_chunk(ch + 1, (ch + 1) & 1);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        if (active) {
            const float *wbuf = wstage + (ch & 1) * KC * N + c0;
            const float *abuf = At + (size_t)ch * KC * PITCH + r0;
#pragma unroll
            for (int kk = 0; kk < KC; ++kk) {
                float a[TM], w[TN];
#pragma unroll
                for (int i = 0; i < TM; i += 4) {
                    const float4 v = *reinterpret_cast<const float4 *>(abuf + kk * PITCH + i);
                    a[i] = v.x; a[i + 1] = v.y; a[i + 2] = v.z; a[i + 3] = v.w;
                }
#pragma unroll
                for (int j = 0; j < TN; j += 4) {
                    const float4 v = *reinterpret_cast<const float4 *>(wbuf + kk * N + j);
                    w[j] = v.x; w[j + 1] = v.y; w[j + 2] = v.z; w[j + 3] = v.w;
                }
#pragma unroll
                for (int i = 0; i < TM; ++i)
#pragma unroll
                    for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
            }
        }
        __syncthreads();   // the buffer (ch & 1) is refilled two iterations later by load_chunk(ch + 2)
    }
    if (active) epi(r0, c0, acc);
}

// epilogue helper: out_T[c][r] = relu(acc + bias[c]) written K-major for the next layer
template <int PITCH, int TM, int TN>
struct EpiReluToSmem {
    float *out;
    const float *__restrict__ bias;
    __device__ __forceinline__ void operator()(int r0, int c0, float (&acc)[TM][TN]) const {
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const float bj = __ldg(bias + c0 + j);
#pragma unroll
            for (int i = 0; i < TM; i += 4) {
                float4 v;
                v.x = fmaxf(acc[i][j] + bj, 0.f);
                v.y = fmaxf(acc[i + 1][j] + bj, 0.f);
                v.z = fmaxf(acc[i + 2][j] + bj, 0.f);
                v.w = fmaxf(acc[i + 3][j] + bj, 0.f);
                *reinterpret_cast<float4 *>(out + (size_t)(c0 + j) * PITCH + r0 + i) = v;
            }
        }
    }
};

// epilogue helper: relu(acc + bias) then max over the thread's TM rows, merged into smax[group][c] with
// an integer atomicMax (valid because every value is >= +0).  group = r0 / NS.
template <int NS, int N, int TM, int TN>
struct EpiReluMaxToSmem {
    int *smax;   // [ROWS / NS][N] as int bit patterns, pre-zeroed
    const float *__restrict__ bias;
    __device__ __forceinline__ void operator()(int r0, int c0, float (&acc)[TM][TN]) const {
        static_assert(NS % TM == 0, "a thread's rows must stay inside one neighbourhood");
        const int g = r0 / NS;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const float bj = __ldg(bias + c0 + j);
            float m = 0.f;
#pragma unroll
            for (int i = 0; i < TM; ++i) m = fmaxf(m, acc[i][j] + bj);
            atomicMax(smax + g * N + c0 + j, __float_as_int(m));
        }
    }
};

// ---------------------------------------------------------------------------------------------------
// point_gemm_kernel: U[row, 0:N] = F[row, 0:K] . Wf[K][N]   (no bias, no activation)
// ---------------------------------------------------------------------------------------------------
template <int K, int N, int TM, int TN, int NT>
__global__ void __launch_bounds__(NT)
point_gemm_kernel(const float *__restrict__ F, int ldf, const float *__restrict__ Wf, float *__restrict__ U) {
    constexpr int ROWS = 64, PITCH = ROWS + 4;
    extern __shared__ __align__(16) float smem[];
    float *At = smem;                     // [K][PITCH]
    float *wstage = At + (size_t)K * PITCH;   // [2][KC][N]
    const int tid = threadIdx.x;
    const size_t row0 = (size_t)blockIdx.x * ROWS;
    // transpose-load the 64 x K tile: lane = (row_local 0..15, k-quad 0..1) keeps the smem stores conflict-free
    for (int it = tid; it < ROWS * (K / 4); it += NT) {
        const int rl = it & 15, kq = (it >> 4) & 1, rest = it >> 5;
        const int rblk = rest % (ROWS / 16), kblk = rest / (ROWS / 16);
        const int r = rblk * 16 + rl, k = (kblk * 2 + kq) * 4;
        const float4 v = *reinterpret_cast<const float4 *>(F + (row0 + r) * ldf + k);
        At[(k + 0) * PITCH + r] = v.x;
        At[(k + 1) * PITCH + r] = v.y;
        At[(k + 2) * PITCH + r] = v.z;
        At[(k + 3) * PITCH + r] = v.w;
    }
    __syncthreads();
    auto epi = [&](int r0, int c0, float (&acc)[TM][TN]) {
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
            for (int j = 0; j < TN; j += 4)
                *reinterpret_cast<float4 *>(U + (row0 + r0 + i) * N + c0 + j) =
                    make_float4(acc[i][j], acc[i][j + 1], acc[i][j + 2], acc[i][j + 3]);
    };
    dense_layer<ROWS, PITCH, K, N, TM, TN, NT>(At, Wf, wstage, epi);
}

// ---------------------------------------------------------------------------------------------------
// sa_kernel: one set-abstraction scale for a tile of TC centres of one object
// ---------------------------------------------------------------------------------------------------
template <int LEVEL, int SCALE>
struct SaCfg;
// level 0: 1024 -> 512, no input features
template <> struct SaCfg<0, 0> { static constexpr int N_IN = 1024, NPOINT = 512, NS = 16, TC = 16, NT = 256, TM2 = 8, TN2 = 4, TM3 = 8, TN3 = 4, C_TOTAL = 96, CH_OFF = 0; static constexpr float RADIUS = 0.02f; };
template <> struct SaCfg<0, 1> { static constexpr int N_IN = 1024, NPOINT = 512, NS = 32, TC = 8,  NT = 256, TM2 = 8, TN2 = 4, TM3 = 8, TN3 = 8, C_TOTAL = 96, CH_OFF = 32; static constexpr float RADIUS = 0.04f; };
// level 1: 512 -> 256, 96 input features
template <> struct SaCfg<1, 0> { static constexpr int N_IN = 512, NPOINT = 256, NS = 16, TC = 8, NT = 256, TM2 = 4, TN2 = 8, TM3 = 8, TN3 = 8, C_TOTAL = 256, CH_OFF = 0; static constexpr float RADIUS = 0.04f; };
template <> struct SaCfg<1, 1> { static constexpr int N_IN = 512, NPOINT = 256, NS = 32, TC = 4, NT = 256, TM2 = 8, TN2 = 8, TM3 = 8, TN3 = 8, C_TOTAL = 256, CH_OFF = 128; static constexpr float RADIUS = 0.08f; };
// level 2: 256 -> 128, 256 input features
template <> struct SaCfg<2, 0> { static constexpr int N_IN = 256, NPOINT = 128, NS = 16, TC = 4, NT = 256, TM2 = 8, TN2 = 8, TM3 = 8, TN3 = 8, C_TOTAL = 512, CH_OFF = 0; static constexpr float RADIUS = 0.08f; };
template <> struct SaCfg<2, 1> { static constexpr int N_IN = 256, NPOINT = 128, NS = 32, TC = 2, NT = 256, TM2 = 8, TN2 = 8, TM3 = 8, TN3 = 8, C_TOTAL = 512, CH_OFF = 256; static constexpr float RADIUS = 0.16f; };

template <int LEVEL, int SCALE>
struct SaSmem {
    using C = SaCfg<LEVEL, SCALE>;
    static constexpr MlpSpec M = enc_spec(LEVEL, SCALE);
    static constexpr int ROWS = C::TC * C::NS, PITCH = ROWS + 4;
    static constexpr size_t h1 = 0;                                         // [c1][PITCH]
    static constexpr size_t h2 = h1 + (size_t)M.c1 * PITCH;                 // [c2][PITCH]
    static constexpr size_t wstage = h2 + (size_t)M.c2 * PITCH;             // [2][KC][max(c2,c3)]
    static constexpr size_t xyz = wstage + (size_t)2 * KC * (M.c2 > M.c3 ? M.c2 : M.c3);   // [N_IN*3]
    static constexpr size_t nbr = xyz + (size_t)C::N_IN * 3;                // int [ROWS]
    static constexpr size_t smax = nbr + ROWS;                              // int [TC][c3]
    static constexpr size_t ctr = smax + (size_t)C::TC * M.c3;              // [TC*3] (+pad)
    static constexpr size_t total_floats = ctr + ((C::TC * 3 + 3) / 4) * 4;
    static constexpr size_t bytes = total_floats * sizeof(float);
};

template <int LEVEL, int SCALE>
__global__ void __launch_bounds__(SaCfg<LEVEL, SCALE>::NT)
sa_kernel(const float *__restrict__ xyz_in,    // [B, N_IN, 3]
          const float *__restrict__ new_xyz,   // [B, NPOINT, 3]
          const float *__restrict__ U,         // [B, N_IN, c1]  W1_feat . f  (unused at level 0)
          const float *__restrict__ W,         // packed weights of this (level, scale)
          float *__restrict__ feat_out)        // [B, NPOINT, C_TOTAL]
{
    using C = SaCfg<LEVEL, SCALE>;
    using S = SaSmem<LEVEL, SCALE>;
    constexpr MlpSpec M = enc_spec(LEVEL, SCALE);
    constexpr int ROWS = S::ROWS, PITCH = S::PITCH, NT = C::NT, NS = C::NS, TC = C::TC;
    extern __shared__ __align__(16) float smem[];
    float *h1 = smem + S::h1, *h2 = smem + S::h2, *wstage = smem + S::wstage, *sxyz = smem + S::xyz;
    int *nbr = reinterpret_cast<int *>(smem + S::nbr);
    int *smax = reinterpret_cast<int *>(smem + S::smax);
    float *ctr = smem + S::ctr;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.y, c_base = blockIdx.x * TC;

    // ---- stage the object's source cloud (AoS, contiguous) and this tile's centres ------------------
    const float *p = xyz_in + (size_t)b * C::N_IN * 3;
    for (int i = tid; i < C::N_IN * 3 / 4; i += NT)
        reinterpret_cast<float4 *>(sxyz)[i] = reinterpret_cast<const float4 *>(p)[i];
    if (tid < TC * 3) ctr[tid] = new_xyz[((size_t)b * C::NPOINT + c_base) * 3 + tid];
    for (int i = tid; i < TC * M.c3; i += NT) smax[i] = 0;
    __syncthreads();

    // ---- ball query: one warp per centre, 32 candidates per ballot, ascending-k order preserved -------
    {
        const float r2 = C::RADIUS * C::RADIUS;   // fp32 product, ball_query_gpu.cu:23
        for (int tc = warp; tc < TC; tc += NT / 32) {
            const float cx = ctr[tc * 3 + 0], cy = ctr[tc * 3 + 1], cz = ctr[tc * 3 + 2];
            int cnt = 0, first = 0;
            for (int base = 0; base < C::N_IN && cnt < NS; base += 32) {
                const int k = base + lane;
                const bool hit = dist2_ref(cx, cy, cz, sxyz[k * 3], sxyz[k * 3 + 1], sxyz[k * 3 + 2]) < r2;
                const unsigned mask = __ballot_sync(0xffffffffu, hit);
                if (mask) {
                    if (cnt == 0) first = base + __ffs(mask) - 1;
                    const int slot = cnt + __popc(mask & ((1u << lane) - 1u));
                    if (hit && slot < NS) nbr[tc * NS + slot] = k;
                    cnt += __popc(mask);
                }
            }
            cnt = cnt < NS ? cnt : NS;
            for (int s = cnt + lane; s < NS; s += 32) nbr[tc * NS + s] = first;   // pad with the first hit
        }
    }
    __syncthreads();

    // ---- layer 1 for every (centre, neighbour) row, written K-major into h1 --------------------------
    {
        const float *wx = W + off_wx(M), *b1 = W + off_b1(M);
        // lane = (row_local 0..15, channel-quad 0..1): conflict-free transposed stores, 32 B gathers
        for (int it = tid; it < ROWS * (M.c1 / 4); it += NT) {
            const int rl = it & 15, cq = (it >> 4) & 1, rest = it >> 5;
            const int rblk = rest % (ROWS / 16), cblk = rest / (ROWS / 16);
            const int r = rblk * 16 + rl, c = (cblk * 2 + cq) * 4;
            const int j = nbr[r], tc = r / NS;
            const float dx = sxyz[j * 3 + 0] - ctr[tc * 3 + 0];   // grouped_xyz -= new_xyz (pointnet2_utils.py:253)
            const float dy = sxyz[j * 3 + 1] - ctr[tc * 3 + 1];
            const float dz = sxyz[j * 3 + 2] - ctr[tc * 3 + 2];
            float4 v = __ldg(reinterpret_cast<const float4 *>(b1 + c));
            if (LEVEL > 0) {
                const float4 u = __ldg(reinterpret_cast<const float4 *>(U + ((size_t)b * C::N_IN + j) * M.c1 + c));
                v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w;
            }
            const float4 w0 = __ldg(reinterpret_cast<const float4 *>(wx + 0 * M.c1 + c));
            const float4 w1 = __ldg(reinterpret_cast<const float4 *>(wx + 1 * M.c1 + c));
            const float4 w2 = __ldg(reinterpret_cast<const float4 *>(wx + 2 * M.c1 + c));
            v.x = fmaf(dz, w2.x, fmaf(dy, w1.x, fmaf(dx, w0.x, v.x)));
            v.y = fmaf(dz, w2.y, fmaf(dy, w1.y, fmaf(dx, w0.y, v.y)));
            v.z = fmaf(dz, w2.z, fmaf(dy, w1.z, fmaf(dx, w0.z, v.z)));
            v.w = fmaf(dz, w2.w, fmaf(dy, w1.w, fmaf(dx, w0.w, v.w)));
            h1[(c + 0) * PITCH + r] = fmaxf(v.x, 0.f);
            h1[(c + 1) * PITCH + r] = fmaxf(v.y, 0.f);
            h1[(c + 2) * PITCH + r] = fmaxf(v.z, 0.f);
            h1[(c + 3) * PITCH + r] = fmaxf(v.w, 0.f);
        }
    }
    __syncthreads();

    // ---- layer 2 -> h2 (K-major), layer 3 -> max over the neighbourhood -------------------------------
    dense_layer<ROWS, PITCH, M.c1, M.c2, C::TM2, C::TN2, NT>(
        h1, W + off_w2(M), wstage, EpiReluToSmem<PITCH, C::TM2, C::TN2>{h2, W + off_b2(M)});
    __syncthreads();
    dense_layer<ROWS, PITCH, M.c2, M.c3, C::TM3, C::TN3, NT>(
        h2, W + off_w3(M), wstage, EpiReluMaxToSmem<NS, M.c3, C::TM3, C::TN3>{smax, W + off_b3(M)});
    __syncthreads();

    for (int i = tid; i < TC * M.c3; i += NT) {
        const int tc = i / M.c3, c = i % M.c3;
        feat_out[((size_t)b * C::NPOINT + c_base + tc) * C::C_TOTAL + C::CH_OFF + c] = __int_as_float(smax[i]);
    }
}

// ---------------------------------------------------------------------------------------------------
// groupall_kernel: SA4.  grid (4 row blocks of the 128 points, B) per scale; 256 threads.
// ---------------------------------------------------------------------------------------------------
constexpr int kGaThreads = 256;
template <int SCALE>
struct GaSmem {
    static constexpr MlpSpec M = enc_spec(3, SCALE);
    static constexpr int ROWS = 32, PITCH = ROWS + 4;
    static constexpr size_t a0 = 0;                                      // [512][PITCH] features, later h2 [c2][PITCH]
    static constexpr size_t h1 = a0 + (size_t)512 * PITCH;               // [256][PITCH]
    static constexpr size_t wstage = h1 + (size_t)M.c1 * PITCH;          // [2][KC][512]
    static constexpr size_t xyz = wstage + (size_t)2 * KC * 512;         // [ROWS*3]
    static constexpr size_t total_floats = xyz + ROWS * 3;
    static constexpr size_t bytes = total_floats * sizeof(float);
};

template <int SCALE>
__global__ void __launch_bounds__(kGaThreads)
groupall_kernel(const float *__restrict__ xyz3,    // [B,128,3]  absolute coordinates
                const float *__restrict__ feat3,   // [B,128,512]
                const float *__restrict__ W,       // packed weights of (level 3, SCALE)
                float *__restrict__ pts_feat)      // [B,1024], pre-zeroed; this scale writes [SCALE*512, +512)
{
    using S = GaSmem<SCALE>;
    constexpr MlpSpec M = enc_spec(3, SCALE);
    constexpr int ROWS = S::ROWS, PITCH = S::PITCH, NT = kGaThreads;
    extern __shared__ __align__(16) float smem[];
    float *a0 = smem + S::a0, *h1 = smem + S::h1, *wstage = smem + S::wstage, *sxyz = smem + S::xyz;
    const int tid = threadIdx.x, b = blockIdx.y, row0 = blockIdx.x * ROWS;

    const float *F = feat3 + ((size_t)b * 128 + row0) * 512;
    for (int it = tid; it < ROWS * (512 / 4); it += NT) {
        const int rl = it & 15, kq = (it >> 4) & 1, rest = it >> 5;
        const int rblk = rest % (ROWS / 16), kblk = rest / (ROWS / 16);
        const int r = rblk * 16 + rl, k = (kblk * 2 + kq) * 4;
        const float4 v = __ldg(reinterpret_cast<const float4 *>(F + (size_t)r * 512 + k));
        a0[(k + 0) * PITCH + r] = v.x;
        a0[(k + 1) * PITCH + r] = v.y;
        a0[(k + 2) * PITCH + r] = v.z;
        a0[(k + 3) * PITCH + r] = v.w;
    }
    if (tid < ROWS * 3) sxyz[tid] = xyz3[((size_t)b * 128 + row0) * 3 + tid];
    __syncthreads();

    // layer 1: features through the tiled GEMM, xyz (absolute: GroupAll, pointnet2_utils.py:281-289) in the epilogue
    {
        constexpr int TM = 8, TN = 8;
        const float *wx = W + off_wx(M), *b1 = W + off_b1(M);
        auto epi = [&](int r0, int c0, float (&acc)[TM][TN]) {
#pragma unroll
            for (int j = 0; j < TN; ++j) {
                const float bj = __ldg(b1 + c0 + j);
                const float w0 = __ldg(wx + 0 * M.c1 + c0 + j), w1 = __ldg(wx + 1 * M.c1 + c0 + j),
                            w2 = __ldg(wx + 2 * M.c1 + c0 + j);
#pragma unroll
                for (int i = 0; i < TM; ++i) {
                    const int r = r0 + i;
                    float v = acc[i][j] + bj;
                    v = fmaf(sxyz[r * 3 + 2], w2, fmaf(sxyz[r * 3 + 1], w1, fmaf(sxyz[r * 3 + 0], w0, v)));
                    h1[(size_t)(c0 + j) * PITCH + r] = fmaxf(v, 0.f);
                }
            }
        };
        dense_layer<ROWS, PITCH, 512, M.c1, TM, TN, NT>(a0, W + off_wf(M), wstage, epi);
    }
    __syncthreads();
    float *h2 = a0;   // features are dead now
    dense_layer<ROWS, PITCH, M.c1, M.c2, 8, 8, NT>(h1, W + off_w2(M), wstage,
                                                   EpiReluToSmem<PITCH, 8, 8>{h2, W + off_b2(M)});
    __syncthreads();
    {
        constexpr int TM = 8, TN = 8;
        const float *b3 = W + off_b3(M);
        int *out = reinterpret_cast<int *>(pts_feat + (size_t)b * 1024 + SCALE * 512);
        // per-thread max over its 8 rows, then across the 4 row groups via shuffles (lanes run over rg first)
        auto epi = [&](int r0, int c0, float (&acc)[TM][TN]) {
#pragma unroll
            for (int j = 0; j < TN; ++j) {
                const float bj = __ldg(b3 + c0 + j);
                float m = 0.f;
#pragma unroll
                for (int i = 0; i < TM; ++i) m = fmaxf(m, acc[i][j] + bj);
                m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
                m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
                if ((threadIdx.x & 3) == 0) atomicMax(out + c0 + j, __float_as_int(m));
            }
        };
        dense_layer<ROWS, PITCH, M.c2, 512, TM, TN, NT>(h2, W + off_w3(M), wstage, epi);
    }
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
size_t ga_tc_scratch_bytes(int B);   // ga_tc.cu

struct EncWorkspace {
    float *nx1, *nx2, *nx3;   // [B,512,3] [B,256,3] [B,128,3]
    float *feat1, *feat2, *feat3;   // [B,512,96] [B,256,256] [B,128,512]
    float *u[2];   // scratch for U (max over levels): [B,512,64] / [B,256,128]
    uint8_t *ga;   // tensor-core GroupAll operand images (ga_tc.cu)
    size_t bytes;
};

static EncWorkspace carve(void *base, int B) {
    EncWorkspace w{};
    size_t off = 0;
    auto take = [&](size_t floats) {
        float *p = base ? reinterpret_cast<float *>(reinterpret_cast<char *>(base) + off) : nullptr;
        off += ((floats * sizeof(float) + 255) / 256) * 256;
        return p;
    };
    w.nx1 = take((size_t)B * 512 * 3);
    w.nx2 = take((size_t)B * 256 * 3);
    w.nx3 = take((size_t)B * 128 * 3);
    w.feat1 = take((size_t)B * 512 * 96);
    w.feat2 = take((size_t)B * 256 * 256);
    w.feat3 = take((size_t)B * 128 * 512);
    w.u[0] = take((size_t)B * 512 * 64);   // >= B*256*128
    w.u[1] = take((size_t)B * 512 * 64);
    w.ga = reinterpret_cast<uint8_t *>(take(ga_tc_scratch_bytes(B) / sizeof(float)));
    w.bytes = off;
    return w;
}

template <int LEVEL, int SCALE>
static int launch_sa(const float *xyz_in, const float *new_xyz, const float *U, const float *enc_w, float *feat_out,
                     int B, cudaStream_t st) {
    using C = SaCfg<LEVEL, SCALE>;
    using S = SaSmem<LEVEL, SCALE>;
    static_assert(C::NPOINT % C::TC == 0, "tile");
    GPB_CUDA(cudaFuncSetAttribute(sa_kernel<LEVEL, SCALE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::bytes));
    dim3 grid(C::NPOINT / C::TC, B);
    sa_kernel<LEVEL, SCALE><<<grid, C::NT, S::bytes, st>>>(xyz_in, new_xyz, U, enc_w + spec_offset(LEVEL, SCALE), feat_out);
    GPB_LAUNCHED();
    return GPB_OK;
}

template <int K, int N, int TM, int TN, int NT>
static int launch_point_gemm(const float *F, int ldf, const float *Wf, float *U, size_t rows, cudaStream_t st) {
    constexpr size_t smem = ((size_t)K * 68 + (size_t)2 * KC * N) * sizeof(float);
    GPB_CUDA(cudaFuncSetAttribute(point_gemm_kernel<K, N, TM, TN, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    point_gemm_kernel<K, N, TM, TN, NT><<<(unsigned)(rows / 64), NT, smem, st>>>(F, ldf, Wf, U);
    GPB_LAUNCHED();
    return GPB_OK;
}

}  // namespace gpb

using namespace gpb;

extern "C" size_t gpb_encoder_weights_floats(void) { return kEncoderFloats; }

extern "C" size_t gpb_encode_workspace_bytes(int B) { return B > 0 ? carve(nullptr, B).bytes : 0; }

namespace gpb {
int launch_sa3_tc(const float *xyz_in, const float *new_xyz, const float *U, const float *consts, const void *wstream, int scale,
                  float *feat_out, uint8_t *a0_hi, uint8_t *a0_lo, int B, cudaStream_t st);   // sa_tc.cu
size_t ga_tc_blob_bytes();                                                                    // ga_tc.cu
size_t ga_tc_scratch_bytes(int B);
int launch_sa1_tc(const float *pts, const float *new_xyz, const float *consts, const void *wimg, float *feat_out, int B, cudaStream_t st);
constexpr size_t kEncTcL1Bytes = 2048 + (32 * 32 + 32 * 64) * 4;   // level 1 scale 1: [consts 2 KiB | image 12 KiB], after the GroupAll block
int launch_groupall_tc(const uint8_t *blob, uint8_t *scratch, const float *xyz3, float *pts_feat, int B, cudaStream_t st);
constexpr size_t kEncTcConstBytes = 8192;                     // 2 x 992 floats, padded
constexpr size_t kEncTcScaleBytes = (size_t)22 * 16384;
int launch_sa2_tc(const float *xyz_in, const float *new_xyz, const float *U, const float *consts, const void *wimg, int scale,
                  float *feat_out, int B, cudaStream_t st);   // sa_tc.cu
// level-2 block, appended: [consts s0 2 KiB | image s0 48 KiB | consts s1 2 KiB | image s1 72 KiB]
constexpr size_t kEncTcL2Off = kEncTcConstBytes + 2 * kEncTcScaleBytes;
constexpr size_t kEncTcL2Const = 2048, kEncTcL2Img0 = (64 * 64 + 64 * 128) * 4, kEncTcL2Img1 = (64 * 96 + 96 * 128) * 4;
constexpr size_t kEncTcGaOff = kEncTcL2Off + 2 * kEncTcL2Const + kEncTcL2Img0 + kEncTcL2Img1;   // GroupAll block (ga_tc.cu) last
}  // namespace gpb

// Two side streams per device for the work that may run beside the caller's stream inside one encoder pass: [0] (highest priority)
// furthest-point sampling of levels 2 and 3, [1] the narrow scale of every set-abstraction level.  Fork / join use events created per
// call, so concurrent callers on different streams share nothing beyond the side streams' own ordering.
static bool encoder_side_streams(cudaStream_t (&out)[2]) {
    static std::mutex mu;
    static cudaStream_t side[64][2] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return false;
    std::lock_guard<std::mutex> lock(mu);
    if (!side[dev][0]) {
        // FPS: a few small CTAs (B x 256 threads) that must be placed ahead of the set-abstraction grids they run beside
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        if (cudaStreamCreateWithPriority(&side[dev][0], cudaStreamNonBlocking, hi) != cudaSuccess ||
            cudaStreamCreateWithFlags(&side[dev][1], cudaStreamNonBlocking) != cudaSuccess) {
            cudaGetLastError();
            side[dev][0] = side[dev][1] = nullptr;
            return false;
        }
    }
    out[0] = side[dev][0];
    out[1] = side[dev][1];
    return true;
}

// `to` waits for everything enqueued on `from` so far (one throw-away event; the runtime releases it when the work has completed)
static int stream_follows(cudaStream_t to, cudaStream_t from) {
    cudaEvent_t ev;
    GPB_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    cudaError_t e = cudaEventRecord(ev, from);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(to, ev, 0);
    cudaEventDestroy(ev);
    GPB_CUDA(e);
    return GPB_OK;
}

static int encode_impl(const float *pts, int B, const float *enc_w, const uint8_t *enc_tc, float *pts_feat, void *workspace,
                       size_t workspace_bytes, int *fps_idx1, int *fps_idx2, int *fps_idx3, void *stream) {
    GPB_REQUIRE(B >= 0, "encode: B < 0");
    if (B == 0) return GPB_OK;
    GPB_REQUIRE(pts && enc_w && pts_feat && workspace, "encode: NULL buffer");
    GPB_REQUIRE(B <= 65535, "encode: B must be <= 65535");
    EncWorkspace w = carve(workspace, B);
    if (workspace_bytes < w.bytes) {
        set_error("encode: workspace %zu < required %zu bytes", workspace_bytes, w.bytes);
        return GPB_EWORKSPACE;
    }
    GPB_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "encode: workspace must be 256-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    int rc;
    // GPB_ENC_SERIAL=1: everything on the caller's stream in program order (the round-1 schedule), for A/B timing
    static const bool serial = [] { const char *v = getenv("GPB_ENC_SERIAL"); return v && v[0] == '1'; }();
    cudaStream_t side[2] = {nullptr, nullptr};
    const bool forked = !serial && encoder_side_streams(side);
    // the two scales of a level are independent kernels writing disjoint channel ranges of the level's feature tensor: the narrow
    // scale (index 0) runs on side stream 1 beside the wide one, which stays on the caller's stream
    cudaStream_t s0 = forked ? side[1] : st;
    auto fork = [&]() -> int { return forked ? stream_follows(s0, st) : GPB_OK; };     // scale 0 may start: its inputs are complete on `st`
    auto join = [&]() -> int { return forked ? stream_follows(st, s0) : GPB_OK; };     // the level's output is complete on `st`

    // Furthest-point sampling is 896 dependent arg-max rounds on B CTAs: level 1 (512 rounds) first, then levels 2 and 3 (384 rounds)
    // on side stream 0 beside level 1's set abstraction, which only needs new_xyz1; joined in front of level 2.
    if (forked) {
        fps_l1_kernel<<<B, 1024, 0, st>>>(pts, w.nx1, fps_idx1);
        GPB_LAUNCHED();
        if ((rc = stream_follows(side[0], st))) return rc;
        fps_l23_kernel<<<B, 512, 0, side[0]>>>(w.nx1, w.nx2, w.nx3, fps_idx2, fps_idx3);
        GPB_LAUNCHED();
    } else {
        fps3_kernel<<<B, kFps3Threads, 0, st>>>(pts, w.nx1, w.nx2, w.nx3, fps_idx1, fps_idx2, fps_idx3);
        GPB_LAUNCHED();
    }

    // level 1 (no input features)
    if ((rc = fork())) return rc;
    if ((rc = launch_sa<0, 0>(pts, w.nx1, nullptr, enc_w, w.feat1, B, s0))) return rc;
    if (enc_tc) {
        const uint8_t *l1 = enc_tc + kEncTcGaOff + ga_tc_blob_bytes();
        if ((rc = launch_sa1_tc(pts, w.nx1, reinterpret_cast<const float *>(l1), l1 + 2048, w.feat1, B, st))) return rc;
    } else {
        if ((rc = launch_sa<0, 1>(pts, w.nx1, nullptr, enc_w, w.feat1, B, st))) return rc;
    }
    if ((rc = join())) return rc;
    if (forked && (rc = stream_follows(st, side[0]))) return rc;   // new_xyz2 / new_xyz3 (and fps_idx2 / fps_idx3) are complete from here on

    // level 2: U = W1_feat . feat1 per source point, then the grouped part
    {
        constexpr MlpSpec m0 = enc_spec(1, 0), m1 = enc_spec(1, 1);
        if ((rc = fork())) return rc;
        if ((rc = launch_point_gemm<96, 64, 4, 8, 128>(w.feat1, 96, enc_w + spec_offset(1, 0) + off_wf(m0), w.u[0], (size_t)B * 512, s0))) return rc;
        if ((rc = launch_point_gemm<96, 64, 4, 8, 128>(w.feat1, 96, enc_w + spec_offset(1, 1) + off_wf(m1), w.u[1], (size_t)B * 512, st))) return rc;
        if (enc_tc) {
            const uint8_t *l2 = enc_tc + kEncTcL2Off;
            if ((rc = launch_sa2_tc(w.nx1, w.nx2, w.u[0], reinterpret_cast<const float *>(l2), l2 + kEncTcL2Const, 0, w.feat2, B, s0))) return rc;
            l2 += kEncTcL2Const + kEncTcL2Img0;
            if ((rc = launch_sa2_tc(w.nx1, w.nx2, w.u[1], reinterpret_cast<const float *>(l2), l2 + kEncTcL2Const, 1, w.feat2, B, st))) return rc;
        } else {
            if ((rc = launch_sa<1, 0>(w.nx1, w.nx2, w.u[0], enc_w, w.feat2, B, s0))) return rc;
            if ((rc = launch_sa<1, 1>(w.nx1, w.nx2, w.u[1], enc_w, w.feat2, B, st))) return rc;
        }
        if ((rc = join())) return rc;
    }
    // level 3
    {
        constexpr MlpSpec m0 = enc_spec(2, 0), m1 = enc_spec(2, 1);
        if ((rc = fork())) return rc;
        if ((rc = launch_point_gemm<256, 128, 8, 4, 256>(w.feat2, 256, enc_w + spec_offset(2, 0) + off_wf(m0), w.u[0], (size_t)B * 256, s0))) return rc;
        if ((rc = launch_point_gemm<256, 128, 8, 4, 256>(w.feat2, 256, enc_w + spec_offset(2, 1) + off_wf(m1), w.u[1], (size_t)B * 256, st))) return rc;
        if (enc_tc) {
            const float *consts = reinterpret_cast<const float *>(enc_tc);
            uint8_t *a0_hi = w.ga, *a0_lo = w.ga + (size_t)B * 131072;
            if ((rc = launch_sa3_tc(w.nx2, w.nx3, w.u[0], consts, enc_tc + kEncTcConstBytes, 0, w.feat3, a0_hi, a0_lo, B, s0))) return rc;
            if ((rc = launch_sa3_tc(w.nx2, w.nx3, w.u[1], consts + 992, enc_tc + kEncTcConstBytes + kEncTcScaleBytes, 1, w.feat3, a0_hi, a0_lo, B,
                                    st)))
                return rc;
        } else {
            if ((rc = launch_sa<2, 0>(w.nx2, w.nx3, w.u[0], enc_w, w.feat3, B, s0))) return rc;
            if ((rc = launch_sa<2, 1>(w.nx2, w.nx3, w.u[1], enc_w, w.feat3, B, st))) return rc;
        }
        if ((rc = join())) return rc;
    }
    // level 4 (GroupAll)
    if (enc_tc) return launch_groupall_tc(enc_tc + kEncTcGaOff, w.ga, w.nx3, pts_feat, B, st);
    GPB_CUDA(cudaMemsetAsync(pts_feat, 0, (size_t)B * 1024 * sizeof(float), st));
    GPB_CUDA(cudaFuncSetAttribute(groupall_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GaSmem<0>::bytes));
    GPB_CUDA(cudaFuncSetAttribute(groupall_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GaSmem<1>::bytes));
    groupall_kernel<0><<<dim3(4, B), kGaThreads, GaSmem<0>::bytes, st>>>(w.nx3, w.feat3, enc_w + spec_offset(3, 0), pts_feat);
    GPB_LAUNCHED();
    groupall_kernel<1><<<dim3(4, B), kGaThreads, GaSmem<1>::bytes, st>>>(w.nx3, w.feat3, enc_w + spec_offset(3, 1), pts_feat);
    GPB_LAUNCHED();
    return GPB_OK;
}

extern "C" int gpb_encode(const float *pts, int B, const float *enc_w, float *pts_feat, void *workspace,
                          size_t workspace_bytes, int *fps_idx1, int *fps_idx2, int *fps_idx3, void *stream) {
    return encode_impl(pts, B, enc_w, nullptr, pts_feat, workspace, workspace_bytes, fps_idx1, fps_idx2, fps_idx3, stream);
}

extern "C" size_t gpb_encoder_tc_bytes(void) { return kEncTcGaOff + ga_tc_blob_bytes() + kEncTcL1Bytes; }

extern "C" int gpb_encode_tc(const float *pts, int B, const float *enc_w, const void *enc_tc, float *pts_feat, void *workspace,
                             size_t workspace_bytes, int *fps_idx1, int *fps_idx2, int *fps_idx3, void *stream) {
    GPB_REQUIRE(enc_tc != nullptr, "encode_tc: NULL tensor-core weight image");
    GPB_REQUIRE((reinterpret_cast<uintptr_t>(enc_tc) & 127) == 0, "encode_tc: weight image must be 128-byte aligned");
    return encode_impl(pts, B, enc_w, reinterpret_cast<const uint8_t *>(enc_tc), pts_feat, workspace, workspace_bytes, fps_idx1, fps_idx2,
                       fps_idx3, stream);
}
