// Score / energy trunk and the samplers.
//
// One evaluation of PoseScoreNet.forward (networks/gf_algorithms/scorenet.py:178-222, heads
// 'Rx_Ry_and_T') is restructured as
//     f(x, t, obj) = O . relu( A_pose . pf(x) + obj_bias[obj] + t_bias(t) ),   pf = relu(P2 . relu(P1 . x))
// where obj_bias = A_pts . pts_feat + a  is computed ONCE per object (gpb_object_bias; 67 % of the FLOPs
// of the reference form) and t_bias = A_t . relu(L_t . fourier(t) + b_t) once per time value (10 %).
// What remains per (candidate, step) is 266,752 MAC.
//
//   pc_sampler_kernel    persistent, cooperatively launched: every CTA owns tiles of 24 candidate rows whose
//                        pose state stays in shared memory for all T steps; per step it runs the trunk on
//                        its tiles (weights stream from L2 through registers, activations stay in smem),
//                        contributes to the batch-mean gradient norm (samplers.py:130 — one grid barrier
//                        per step, deterministic fixed-order sum) and applies the Langevin + Euler-Maruyama
//                        update with Gram-Schmidt re-normalisation (samplers.py:132-152).
//   ode_sampler_kernel   persistent: SciPy's RK45 (Dormand-Prince 5(4), scipy/integrate/_ivp/rk.py) with
//                        float64 state and ONE error norm over the whole batch (samplers.py:205), on device.
//   trunk_eval_kernel    a single evaluation (mode 'score' / energy f_theta).
//
// All arithmetic is fp32 FFMA (parity mode).  Thread layout of the two GEMMs: a thread owns two adjacent
// output columns, keeps a 16-deep slice of their weights in registers (prefetching the next slice from
// L2) and sweeps the tile's rows, whose activations are broadcast float4 reads from shared memory.
#include "common.cuh"
#include "sampler_common.cuh"

namespace gpb {

constexpr int kRT = 24;          // rows per tile
constexpr int kNT = 384;         // threads per CTA
constexpr int kHP = 256 + 4;     // pitch of the 256-wide activation tiles
constexpr int kHhP = 768 + 4;    // pitch of the stacked head activations
using TL = TrunkLayout;

struct TrunkSmem {
    float *x;     // [kRT][12]  pose rows of the current tile (fp32)
    float *H1;    // [kRT][kHP]
    float *H2;    // [kRT][kHP]
    float *Hh;    // [kRT][kHhP]
    float *f;     // [kRT][12]  trunk output (before the division by sigma)
    float *wp1;   // [9][256] + [256] bias
    float *ow;    // [9][256] + [12] bias
    float *tb;    // [768] time bias of the current evaluation
    float *tf;    // [128] + [128] scratch for the time embedding
};
constexpr size_t kTrunkSmemFloats = (size_t)kRT * 12 * 2 + (size_t)kRT * kHP * 2 + (size_t)kRT * kHhP + (9 * 256 + 256) +
                                    (9 * 256 + 12) + 768 + 256;

__device__ __forceinline__ TrunkSmem carve_trunk_smem(float *base) {
    TrunkSmem s;
    s.x = base;
    s.f = s.x + kRT * 12;
    s.H1 = s.f + kRT * 12;
    s.H2 = s.H1 + kRT * kHP;
    s.Hh = s.H2 + kRT * kHP;
    s.wp1 = s.Hh + kRT * kHhP;
    s.ow = s.wp1 + 9 * 256 + 256;
    s.tb = s.ow + 9 * 256 + 12;
    s.tf = s.tb + 768;
    return s;
}

__device__ __forceinline__ void load_trunk_constants(const TrunkSmem &s, const float *__restrict__ W) {
    for (int i = threadIdx.x; i < 9 * 256 + 256; i += blockDim.x) s.wp1[i] = W[TL::p1_w + i];   // p1_w then p1_b are adjacent
    for (int i = threadIdx.x; i < 9 * 256 + 12; i += blockDim.x) s.ow[i] = W[TL::o_w + i];     // o_w then o_b are adjacent
}

// t_bias(t)[0:768] into s.tb.  scorenet.py:63-64 (x_proj = ((t*W)*2)*pi in fp32, [sin | cos]) and :195.
__device__ __forceinline__ void compute_time_bias(const TrunkSmem &s, float t, const float *__restrict__ W) {
    const int tid = threadIdx.x;
    if (tid < 64) {
        const float xp = ((t * W[TL::fourier_w + tid]) * 2.0f) * 3.14159265358979323846f;
        s.tf[tid] = sinf(xp);
        s.tf[64 + tid] = cosf(xp);
    }
    __syncthreads();
    if (tid < 128) {
        float acc = W[TL::t_b + tid];
        const float *w = W + TL::t_w + tid;
#pragma unroll 8
        for (int k = 0; k < 128; ++k) acc = fmaf(s.tf[k], __ldg(w + k * 128), acc);
        s.tf[128 + tid] = fmaxf(acc, 0.f);
    }
    __syncthreads();
    for (int n = tid; n < 768; n += blockDim.x) {
        float acc = 0.f;
        const float *w = W + TL::a_t + n;
#pragma unroll 8
        for (int k = 0; k < 128; ++k) acc = fmaf(s.tf[128 + k], __ldg(w + k * 768), acc);
        s.tb[n] = acc;
    }
    __syncthreads();
}

// The trunk for one tile: s.x (rows) -> s.f (f_theta rows).  `tb` is the time bias ([768], smem or global),
// `row0` the global index of the tile's first row, rows >= R are computed on zeros and ignored by callers.
__device__ __forceinline__ void trunk_tile(const TrunkSmem &s, const float *__restrict__ tb, int row0, int R, int K,
                                           const float *__restrict__ obj_bias, const float *__restrict__ W) {
    const int tid = threadIdx.x;
    // ---- P0: h1 = relu(P1 . x + b)            (scorenet.py:105-106)
    if (tid < 256) {
        float w[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) w[i] = s.wp1[i * 256 + tid];
        const float b = s.wp1[9 * 256 + tid];
#pragma unroll 4
        for (int r = 0; r < kRT; ++r) {
            float acc = b;
#pragma unroll
            for (int i = 0; i < 9; ++i) acc = fmaf(s.x[r * 12 + i], w[i], acc);
            s.H1[r * kHP + tid] = fmaxf(acc, 0.f);
        }
    }
    __syncthreads();
    // ---- P1: pf = relu(P2 . h1 + b)           (scorenet.py:107-108)   256 -> 256
    {
        const int cp = tid & 127, r0 = (tid >> 7) * 8;
        const float2 *Wp = reinterpret_cast<const float2 *>(W + TL::p2_w) + cp;
        float2 acc[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) acc[r] = make_float2(0.f, 0.f);
        float2 wc[16], wn[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) wc[i] = __ldg(Wp + i * 128);
        for (int k0 = 0; k0 < 256; k0 += 16) {
            if (k0 + 16 < 256) {
#pragma unroll
                for (int i = 0; i < 16; ++i) wn[i] = __ldg(Wp + (k0 + 16 + i) * 128);
            }
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const float4 *a = reinterpret_cast<const float4 *>(s.H1 + (r0 + r) * kHP + k0);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 av = a[q];
                    acc[r].x = fmaf(av.x, wc[4 * q + 0].x, acc[r].x); acc[r].y = fmaf(av.x, wc[4 * q + 0].y, acc[r].y);
                    acc[r].x = fmaf(av.y, wc[4 * q + 1].x, acc[r].x); acc[r].y = fmaf(av.y, wc[4 * q + 1].y, acc[r].y);
                    acc[r].x = fmaf(av.z, wc[4 * q + 2].x, acc[r].x); acc[r].y = fmaf(av.z, wc[4 * q + 2].y, acc[r].y);
                    acc[r].x = fmaf(av.w, wc[4 * q + 3].x, acc[r].x); acc[r].y = fmaf(av.w, wc[4 * q + 3].y, acc[r].y);
                }
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) wc[i] = wn[i];
        }
        const float2 b = __ldg(reinterpret_cast<const float2 *>(W + TL::p2_b) + cp);
#pragma unroll
        for (int r = 0; r < 8; ++r)
            *reinterpret_cast<float2 *>(s.H2 + (r0 + r) * kHP + 2 * cp) =
                make_float2(fmaxf(acc[r].x + b.x, 0.f), fmaxf(acc[r].y + b.y, 0.f));
    }
    __syncthreads();
    // ---- P2: h = relu(A_pose . pf + obj_bias + t_bias)   (scorenet.py:204-216, three heads stacked)  256 -> 768
    {
        const int cp = tid;   // 384 column pairs
        const float2 *Wp = reinterpret_cast<const float2 *>(W + TL::a_pose) + cp;
        float2 acc[kRT];
#pragma unroll
        for (int r = 0; r < kRT; ++r) acc[r] = make_float2(0.f, 0.f);
        float2 wc[16], wn[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) wc[i] = __ldg(Wp + i * 384);
        for (int k0 = 0; k0 < 256; k0 += 16) {
            if (k0 + 16 < 256) {
#pragma unroll
                for (int i = 0; i < 16; ++i) wn[i] = __ldg(Wp + (k0 + 16 + i) * 384);
            }
#pragma unroll
            for (int r = 0; r < kRT; ++r) {
                const float4 *a = reinterpret_cast<const float4 *>(s.H2 + r * kHP + k0);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 av = a[q];
                    acc[r].x = fmaf(av.x, wc[4 * q + 0].x, acc[r].x); acc[r].y = fmaf(av.x, wc[4 * q + 0].y, acc[r].y);
                    acc[r].x = fmaf(av.y, wc[4 * q + 1].x, acc[r].x); acc[r].y = fmaf(av.y, wc[4 * q + 1].y, acc[r].y);
                    acc[r].x = fmaf(av.z, wc[4 * q + 2].x, acc[r].x); acc[r].y = fmaf(av.z, wc[4 * q + 2].y, acc[r].y);
                    acc[r].x = fmaf(av.w, wc[4 * q + 3].x, acc[r].x); acc[r].y = fmaf(av.w, wc[4 * q + 3].y, acc[r].y);
                }
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) wc[i] = wn[i];
        }
        const float2 t2 = *reinterpret_cast<const float2 *>(tb + 2 * cp);
#pragma unroll
        for (int r = 0; r < kRT; ++r) {
            int row = row0 + r;
            row = row < R ? row : R - 1;
            const float2 ob = __ldg(reinterpret_cast<const float2 *>(obj_bias + (size_t)(row / K) * 768) + cp);
            *reinterpret_cast<float2 *>(s.Hh + r * kHhP + 2 * cp) =
                make_float2(fmaxf(acc[r].x + ob.x + t2.x, 0.f), fmaxf(acc[r].y + ob.y + t2.y, 0.f));
        }
    }
    __syncthreads();
    // ---- P3: f[r][c] = O_c . h_head(c) + o_b   (3 outputs per head, block-diagonal)   768 -> 9
    if (tid < kRT * 9) {
        const int r = tid / 9, c = tid % 9;
        const float4 *h = reinterpret_cast<const float4 *>(s.Hh + r * kHhP + (c / 3) * 256);
        const float4 *w = reinterpret_cast<const float4 *>(s.ow + c * 256);
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 8
        for (int j = 0; j < 64; ++j) {
            const float4 hv = h[j], wv = w[j];
            a0 = fmaf(hv.x, wv.x, a0);
            a1 = fmaf(hv.y, wv.y, a1);
            a2 = fmaf(hv.z, wv.z, a2);
            a3 = fmaf(hv.w, wv.w, a3);
        }
        s.f[r * 12 + c] = ((a0 + a1) + (a2 + a3)) + s.ow[9 * 256 + c];
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------------
// object bias: obj_bias[b, n] = sum_k pts_feat[b,k] * A_pts[k][n] + a_b[n]; 8 objects per CTA
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
object_bias_kernel(const float *__restrict__ pts_feat, int B, const float *__restrict__ W, float *__restrict__ out) {
    // grid (ceil(B/8), 3): 8 objects x 256 of the 768 columns per CTA; the K = 1024 reduction is split over 4 thread groups
    // and combined in fixed order (deterministic)
    __shared__ float sf[8][1024];
    float (*sred)[8][256] = reinterpret_cast<float (*)[8][256]>(&sf[0][0]);      // aliases sf after the main loop
    const int b0 = blockIdx.x * 8, tid = threadIdx.x, col = tid & 255, kg = tid >> 8;
    for (int i = tid; i < 8 * 1024; i += 1024) {
        const int bi = b0 + (i >> 10);
        sf[i >> 10][i & 1023] = bi < B ? pts_feat[(size_t)bi * 1024 + (i & 1023)] : 0.f;
    }
    __syncthreads();
    float acc[8];
#pragma unroll
    for (int o = 0; o < 8; ++o) acc[o] = 0.f;
    const int n = blockIdx.y * 256 + col;
    const float *w = W + TL::a_pts + n;
#pragma unroll 8
    for (int k = kg * 256; k < kg * 256 + 256; ++k) {
        const float wv = __ldg(w + (size_t)k * 768);
#pragma unroll
        for (int o = 0; o < 8; ++o) acc[o] = fmaf(sf[o][k], wv, acc[o]);
    }
    __syncthreads();
    if (kg > 0) {
#pragma unroll
        for (int o = 0; o < 8; ++o) sred[kg - 1][o][col] = acc[o];
    }
    __syncthreads();
    if (kg == 0) {
        const float bias = W[TL::a_b + n];
#pragma unroll
        for (int o = 0; o < 8; ++o)
            if (b0 + o < B) out[(size_t)(b0 + o) * 768 + n] = ((acc[o] + sred[0][o][col]) + sred[1][o][col]) + sred[2][o][col] + bias;
    }
}

// time-bias table for the PC sampler: tb[i, 0:768] = t_bias(ts[i])
__global__ void __launch_bounds__(kNT)
time_bias_table_kernel(const float *__restrict__ ts, const float *__restrict__ W, float *__restrict__ table) {
    __shared__ float stf[256];
    __shared__ float stb[768];
    TrunkSmem s{};
    s.tf = stf;
    s.tb = stb;
    compute_time_bias(s, ts[blockIdx.x], W);
    for (int n = threadIdx.x; n < 768; n += blockDim.x) table[(size_t)blockIdx.x * 768 + n] = stb[n];
}

// ---------------------------------------------------------------------------------------------------
// single evaluation
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kNT, 1)
trunk_eval_kernel(const float *__restrict__ pose, int R, int K, float t, const float *__restrict__ obj_bias,
                  const float *__restrict__ W, int divide_mode, const float *__restrict__ center_sub /* [B,3] or null */,
                  float *__restrict__ out /* [R,9] or null */, float *__restrict__ energy /* [R,2] or null */) {
    extern __shared__ __align__(16) float smem[];
    TrunkSmem s = carve_trunk_smem(smem);
    load_trunk_constants(s, W);
    compute_time_bias(s, t, W);   // contains the barriers that publish the constants
    const float sigma = sigma_of_t(t);
    const float div = divide_mode == 1 ? (sigma + 1e-7f) : (divide_mode == 2 ? sigma : 1.0f);
    const int n_tiles = (R + kRT - 1) / kRT;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int row0 = tile * kRT;
        for (int i = threadIdx.x; i < kRT * 9; i += blockDim.x) {
            const int r = i / 9, c = i % 9, row = row0 + r;
            float v = 0.f;
            if (row < R) {
                v = pose[(size_t)row * 9 + c];
                if (center_sub && c >= 6) v -= center_sub[(size_t)(row / K) * 3 + (c - 6)];   // posenet_agent.py:516
            }
            s.x[r * 12 + c] = v;
        }
        __syncthreads();
        trunk_tile(s, s.tb, row0, R, K, obj_bias, W);
        if (out) {
            for (int i = threadIdx.x; i < kRT * 9; i += blockDim.x) {
                const int r = i / 9, c = i % 9, row = row0 + r;
                if (row < R) out[(size_t)row * 9 + c] = s.f[r * 12 + c] / div;
            }
        }
        if (energy && threadIdx.x < kRT) {
            const int r = threadIdx.x, row = row0 + r;
            if (row < R) {   // energynet.py:180-185 'IP', decoupled rot / trans
                float er = 0.f, et = 0.f;
#pragma unroll
                for (int c = 0; c < 6; ++c) er += s.x[r * 12 + c] * (s.f[r * 12 + c] / div);
#pragma unroll
                for (int c = 6; c < 9; ++c) et += s.x[r * 12 + c] * (s.f[r * 12 + c] / div);
                energy[(size_t)row * 2 + 0] = er;
                energy[(size_t)row * 2 + 1] = et;
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------
// predictor-corrector sampler (cond_pc_sampler, samplers.py:102-160)
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kNT, 1)
pc_sampler_kernel(PcParams p) {
    extern __shared__ __align__(16) float smem[];
    TrunkSmem s = carve_trunk_smem(smem);
    float *xs = smem + kTrunkSmemFloats;                          // [tiles_per_cta*kRT][12] persistent pose state
    float *sc = xs + (size_t)p.tiles_per_cta * kRT * 12;          // [tiles_per_cta*kRT][12] scores of this step
    __shared__ float red[kNT / 32];
    __shared__ float s_gnorm;

    const int tid = threadIdx.x;
    const int n_tiles = (p.R + kRT - 1) / kRT;
    load_trunk_constants(s, p.W);
    // initial state: x = x0 (prior sample, samplers.py:116-121)
    for (int lt = 0; lt < p.tiles_per_cta; ++lt) {
        const int tile = blockIdx.x + lt * gridDim.x;
        for (int i = tid; i < kRT * 9; i += kNT) {
            const int r = i / 9, c = i % 9, row = tile * kRT + r;
            xs[(lt * kRT + r) * 12 + c] = (tile < n_tiles && row < p.R) ? p.x0[(size_t)row * 9 + c] : 0.f;
        }
    }
    __syncthreads();

    const float step_size = p.ts[0] - p.ts[1];                    // samplers.py:119
    const float sqrt_step = sqrtf(step_size);
    const float snr_norm = (float)((double)p.snr * 3.0);          // snr * sqrt(pose_dim) evaluated in double (:120,:131)
    unsigned bar_target = 0;

    for (int step = 0; step < p.T; ++step) {
        const float t = p.ts[step];
        const float sigma = sigma_of_t(t);
        const float stdv = sigma + 1e-7f;                         // scorenet.py:217
        const float *tb = p.tb_table + (size_t)step * 768;
        // ---- score for every owned tile ------------------------------------------------------------
        float local = 0.f;
        for (int lt = 0; lt < p.tiles_per_cta; ++lt) {
            const int tile = blockIdx.x + lt * gridDim.x;
            if (tile >= n_tiles) break;                            // uniform across the CTA
            for (int i = tid; i < kRT * 12; i += kNT) s.x[i] = xs[lt * kRT * 12 + i];
            __syncthreads();
            trunk_tile(s, tb, tile * kRT, p.R, p.K, p.obj_bias, p.W);
            if (tid < kRT) {
                const int row = tile * kRT + tid;
                float n2 = 0.f;
#pragma unroll
                for (int c = 0; c < 9; ++c) {
                    const float g = s.f[tid * 12 + c] / stdv;
                    sc[(lt * kRT + tid) * 12 + c] = g;
                    n2 = fmaf(g, g, n2);
                }
                if (row < p.R) local += sqrtf(n2);                 // torch.norm(grad, dim=-1)  (:130)
            }
            __syncthreads();
        }
        // ---- batch-mean gradient norm: CTA partial -> global -> grid barrier -> fixed-order sum ---------
        if (tid < 32) {
            const float v = warp_sum(tid < kRT ? local : 0.f);
            if (tid == 0) p.partial[(step & 1) * gridDim.x + blockIdx.x] = v;
        }
        bar_target += gridDim.x;
        grid_barrier(p.barrier, bar_target);
        if (tid < 32) {
            float v = 0.f;
            for (int i = tid; i < (int)gridDim.x; i += 32) v += __ldcg(p.partial + (step & 1) * gridDim.x + i);
            v = warp_sum(v);
            if (tid == 0) s_gnorm = v / (float)p.R;
        }
        __syncthreads();
        const float grad_norm = s_gnorm;
        const float q = snr_norm / grad_norm;
        const float ls = 2.0f * (q * q);                          // langevin_step_size (:131)
        const float sq2ls = sqrtf(2.0f * ls);
        const float g = sigma * kGCoef;                           // ve_sde diffusion (sde.py:20-24)
        const float g2 = g * g;
        const bool last = step == p.T - 1;
        // ---- update: one thread per owned row --------------------------------------------------------
        for (int lr = tid; lr < p.tiles_per_cta * kRT; lr += kNT) {
            const int lt = lr / kRT, r = lr % kRT;
            const int tile = blockIdx.x + lt * gridDim.x;
            const int row = tile * kRT + r;
            if (tile >= n_tiles || row >= p.R) continue;
            float x[9], gr[9], z[9];
#pragma unroll
            for (int c = 0; c < 9; ++c) {
                x[c] = xs[lr * 12 + c];
                gr[c] = sc[lr * 12 + c];
            }
            row_noise(p, step, 0, row, z);
#pragma unroll
            for (int c = 0; c < 9; ++c) x[c] = (x[c] + ls * gr[c]) + sq2ls * z[c];          // corrector (:132)
            {
                const float n1 = sqrtf(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);           // (:142-143), no eps
                const float n2 = sqrtf(x[3] * x[3] + x[4] * x[4] + x[5] * x[5]);
                x[0] /= n1; x[1] /= n1; x[2] /= n1;
                x[3] /= n2; x[4] /= n2; x[5] /= n2;
            }
            float m[9];
#pragma unroll
            for (int c = 0; c < 9; ++c) m[c] = x[c] + (0.0f - g2 * gr[c]) * step_size;     // predictor mean (:147-148), sign as written
            row_noise(p, step, 1, row, z);
#pragma unroll
            for (int c = 0; c < 9; ++c) x[c] = m[c] + (g * sqrt_step) * z[c];              // (:149)
            gram_schmidt6(x);                                                               // (:152)
#pragma unroll
            for (int c = 0; c < 9; ++c) xs[lr * 12 + c] = x[c];
            const float *ctr = p.pts_center + (size_t)(row / p.K) * 3;
            if (p.process) {
                float *dst = p.process + ((size_t)row * p.T + step) * 9;
#pragma unroll
                for (int c = 0; c < 9; ++c) dst[c] = x[c] + (c >= 6 ? ctr[c - 6] : 0.f);    // (:156)
            }
            if (last) {
#pragma unroll
                for (int c = 6; c < 9; ++c) m[c] += ctr[c - 6];                             // (:157)
                gram_schmidt6(m);                                                           // (:158)
#pragma unroll
                for (int c = 0; c < 9; ++c) p.mean_x[(size_t)row * 9 + c] = m[c];
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------
// probability-flow ODE sampler: SciPy RK45 on device (cond_ode_sampler, samplers.py:163-227)
// ---------------------------------------------------------------------------------------------------
struct OdeParams {
    const float *x0;   // [R,9]
    int R, K;
    double T0, rtol, atol;     // Python floats in the reference (samplers.py:178, posenet.py:94): float64 through the ABI
    int denoise_steps;         // 1000 when num_steps is None (samplers.py:217); 0 = no denoise
    const float *obj_bias, *W, *pts_center;
    double *y;                 // [R,9]   workspace
    double *ynew;              // [R,9]
    double *Kst;               // [7][R,9] stage derivatives
    double *partial;           // [4][gridDim] reduction slots (rotating)
    unsigned *barrier;
    double *pose;              // [R,9] out (float64, :206-207)
    int *stats;                // [4] nfev, accepted, rejected, status
    int tiles_per_cta;
    OdeProcess proc;           // optional trajectory output (the reference's in_process_sample)
};

// f(t, Y) for every owned row: Y (double, global [R,9]) -> Kout (double, global [R,9]).
// ode_func (samplers.py:189-198): x -> fp32, t -> fp32, score in fp32, f = 0 - fp32(0.5 g^2) * score (fp32,
// NumPy-1.23 value-based casting), g = double(sigma_fp32(t32)) * sqrt(2 ln(5000)) (float64).
__device__ __forceinline__ void ode_rhs(const OdeParams &p, const TrunkSmem &s, double t, const double *Y, double *Kout) {
    const int tid = threadIdx.x;
    const int n_tiles = (p.R + kRT - 1) / kRT;
    const float t32 = (float)t;
    compute_time_bias(s, t32, p.W);
    const float sigma = sigma_of_t(t32);
    const double gd = (double)sigma * 4.12727348049926;
    const float coef = (float)(0.5 * gd * gd);
    const float stdv = sigma + 1e-7f;
    for (int lt = 0; lt < p.tiles_per_cta; ++lt) {
        const int tile = blockIdx.x + lt * gridDim.x;
        if (tile >= n_tiles) break;
        for (int i = tid; i < kRT * 9; i += kNT) {
            const int r = i / 9, c = i % 9, row = tile * kRT + r;
            s.x[r * 12 + c] = row < p.R ? (float)__ldcg(Y + (size_t)row * 9 + c) : 0.f;
        }
        __syncthreads();
        trunk_tile(s, s.tb, tile * kRT, p.R, p.K, p.obj_bias, p.W);
        for (int i = tid; i < kRT * 9; i += kNT) {
            const int r = i / 9, c = i % 9, row = tile * kRT + r;
            if (row < p.R) Kout[(size_t)row * 9 + c] = (double)(0.0f - coef * (s.f[r * 12 + c] / stdv));
        }
        __syncthreads();
    }
}

// grid-wide sum of a per-thread double; every CTA gets the identical (fixed-order) result
__device__ __forceinline__ double grid_sum(const OdeParams &p, double v, unsigned &bar_target, int &slot, double *sred) {
    const int tid = threadIdx.x;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((tid & 31) == 0) sred[tid >> 5] = v;
    __syncthreads();
    if (tid == 0) {
        double t = 0.0;
        for (int i = 0; i < kNT / 32; ++i) t += sred[i];
        p.partial[(size_t)slot * gridDim.x + blockIdx.x] = t;
    }
    bar_target += gridDim.x;
    grid_barrier(p.barrier, bar_target);
    if (tid == 0) {
        double t = 0.0;
        for (int i = 0; i < (int)gridDim.x; ++i) t += __ldcg(p.partial + (size_t)slot * gridDim.x + i);
        sred[kNT / 32] = t;
    }
    __syncthreads();
    const double total = sred[kNT / 32];
    __syncthreads();
    slot = (slot + 1) & 3;
    return total;
}

__global__ void __launch_bounds__(kNT, 1)
ode_sampler_kernel(OdeParams p) {
    extern __shared__ __align__(16) float smem[];
    TrunkSmem s = carve_trunk_smem(smem);
    __shared__ double sred[kNT / 32 + 1];
    const int tid = threadIdx.x;
    const int n_tiles = (p.R + kRT - 1) / kRT;
    const size_t NE = (size_t)p.R * 9;
    const double n_total = (double)NE;
    unsigned bar_target = 0;
    int slot = 0;
    load_trunk_constants(s, p.W);
    __syncthreads();

    // owned element iteration helper: elements (row, c) of the CTA's tiles
    auto for_owned = [&](auto fn) {
        for (int lt = 0; lt < p.tiles_per_cta; ++lt) {
            const int tile = blockIdx.x + lt * gridDim.x;
            if (tile >= n_tiles) break;
            for (int i = tid; i < kRT * 9; i += kNT) {
                const int row = tile * kRT + i / 9;
                if (row < p.R) fn((size_t)row * 9 + i % 9);
            }
        }
    };

    for_owned([&](size_t e) { p.y[e] = (double)p.x0[e]; });          // y0 = float64(init_x) (:205)
    __syncthreads();
    // trajectory output (the reference's `xs`, samplers.py:206, :220-224): one thread per owned row
    int te_next = 0;
    auto emit_rows = [&](int n_acc_now, double t_old, double t_new, double h_step) {
        for (int lt = 0; lt < p.tiles_per_cta; ++lt) {
            const int tile = blockIdx.x + lt * gridDim.x;
            if (tile >= n_tiles) break;
            const int row = tile * kRT + tid;
            if (tid < kRT && row < p.R) {
                int te = te_next;
                const double *kb = p.Kst + (size_t)row * 9;
                ode_emit_step(p.proc, p.R, row, p.pts_center + (size_t)(row / p.K) * 3, true, n_acc_now, te, t_old, t_new, h_step,
                              p.y + (size_t)row * 9, p.ynew + (size_t)row * 9,
                              [&](int j, int c) -> double { return __ldcg(kb + (size_t)j * NE + c); });
            }
        }
        ode_emit_step(p.proc, p.R, 0, p.pts_center, false, n_acc_now, te_next, t_old, t_new, h_step, p.y, p.ynew,
                      [&](int, int) -> double { return 0.0; });          // every thread advances the t_eval cursor identically
    };
    if (p.proc.out && p.proc.t_eval == nullptr && p.proc.cap > 0) {       // state 0 = the start
        for (int lt = 0; lt < p.tiles_per_cta; ++lt) {
            const int tile = blockIdx.x + lt * gridDim.x;
            if (tile >= n_tiles) break;
            const int row = tile * kRT + tid;
            if (tid < kRT && row < p.R) ode_write_state(p.proc.out + (size_t)row * 9, p.y + (size_t)row * 9, p.pts_center + (size_t)(row / p.K) * 3);
        }
    }

    const double t0 = (double)p.T0, t_bound = (double)kSamplingEps;   // eps = 1e-5 as a Python float
    const double tb_exact = 1e-5;
    (void)t_bound;
    const double direction = tb_exact < t0 ? -1.0 : 1.0;
    const double rtol = (double)p.rtol, atol = (double)p.atol;
    double *K0 = p.Kst;
    int nfev = 0, n_acc = 0, n_rej = 0, status = 0;
    double t = t0;

    ode_rhs(p, s, t, p.y, K0);   // f0
    ++nfev;
    // ---- select_initial_step (scipy/integrate/_ivp/common.py) -------------------------------------------
    double h_abs;
    {
        const double interval = fabs(tb_exact - t0);
        double a0 = 0.0, a1 = 0.0;
        for_owned([&](size_t e) {
            const double sc = atol + fabs(p.y[e]) * rtol;
            const double u = p.y[e] / sc, v = K0[e] / sc;
            a0 += u * u;
            a1 += v * v;
        });
        const double d0 = sqrt(grid_sum(p, a0, bar_target, slot, sred) / n_total);
        const double d1 = sqrt(grid_sum(p, a1, bar_target, slot, sred) / n_total);
        double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
        h0 = fmin(h0, interval);
        for_owned([&](size_t e) { p.ynew[e] = p.y[e] + h0 * direction * K0[e]; });
        __syncthreads();
        ode_rhs(p, s, t + h0 * direction, p.ynew, p.Kst + NE);   // f1 into stage slot 1 (scratch)
        ++nfev;
        double a2 = 0.0;
        for_owned([&](size_t e) {
            const double sc = atol + fabs(p.y[e]) * rtol;
            const double w = (p.Kst[NE + e] - K0[e]) / sc;
            a2 += w * w;
        });
        const double d2 = sqrt(grid_sum(p, a2, bar_target, slot, sred) / n_total) / h0;
        double h1;
        if (d1 <= 1e-15 && d2 <= 1e-15) h1 = fmax(1e-6, h0 * 1e-3);
        else h1 = pow(0.01 / fmax(d1, d2), 1.0 / 5.0);
        h_abs = fmin(fmin(100.0 * h0, h1), interval);
    }
    // ---- RungeKutta._step_impl loop (scipy/integrate/_ivp/rk.py) ------------------------------------------
    const double SAFETY = 0.9, MIN_FACTOR = 0.2, MAX_FACTOR = 10.0, ERR_EXP = -1.0 / 5.0;
    while (direction * (t - tb_exact) < 0 && status == 0) {
        const double min_step = 10.0 * fabs(nextafter(t, direction * INFINITY) - t);
        if (h_abs < min_step) h_abs = min_step;
        bool accepted = false, rejected = false;
        double t_new = t, h = 0.0;
        while (!accepted) {
            if (h_abs < min_step) { status = -1; break; }
            h = h_abs * direction;
            t_new = t + h;
            if (direction * (t_new - tb_exact) > 0) t_new = tb_exact;
            h = t_new - t;
            h_abs = fabs(h);
            for (int st = 1; st < 6; ++st) {
                for_owned([&](size_t e) {
                    double dy = 0.0;
                    for (int j = 0; j < st; ++j) dy += p.Kst[(size_t)j * NE + e] * kRkA[st][j];   // np.dot(K[:s].T, a[:s])
                    p.ynew[e] = p.y[e] + dy * h;
                });
                __syncthreads();
                ode_rhs(p, s, t + kRkC[st] * h, p.ynew, p.Kst + (size_t)st * NE);
                ++nfev;
            }
            for_owned([&](size_t e) {
                double dy = 0.0;
                for (int j = 0; j < 6; ++j) dy += p.Kst[(size_t)j * NE + e] * kRkB[j];
                p.ynew[e] = p.y[e] + h * dy;
            });
            __syncthreads();
            ode_rhs(p, s, t + h, p.ynew, p.Kst + (size_t)6 * NE);   // f_new = K[6]
            ++nfev;
            double ae = 0.0;
            for_owned([&](size_t e) {
                double err = 0.0;
                for (int j = 0; j < 7; ++j) err += p.Kst[(size_t)j * NE + e] * kRkE[j];
                const double sc = atol + fmax(fabs(p.y[e]), fabs(p.ynew[e])) * rtol;
                const double w = err * h / sc;
                ae += w * w;
            });
            const double error_norm = sqrt(grid_sum(p, ae, bar_target, slot, sred) / n_total);
            if (error_norm < 1.0) {
                double factor = error_norm == 0.0 ? MAX_FACTOR : fmin(MAX_FACTOR, SAFETY * pow(error_norm, ERR_EXP));
                if (rejected) factor = fmin(1.0, factor);
                h_abs *= factor;
                accepted = true;
                ++n_acc;
            } else {
                h_abs *= fmax(MIN_FACTOR, SAFETY * pow(error_norm, ERR_EXP));
                rejected = true;
                ++n_rej;
            }
        }
        if (status != 0) break;
        if (p.proc.out) {
            __syncthreads();
            emit_rows(n_acc, t, t_new, h);
            __syncthreads();
        }
        // accept: y <- y_new, f <- f_new (FSAL)
        for_owned([&](size_t e) {
            p.y[e] = p.ynew[e];
            K0[e] = p.Kst[(size_t)6 * NE + e];
        });
        __syncthreads();
        t = t_new;
    }
    // ---- denoise + normalise + centre (samplers.py:209-226) ------------------------------------------------
    if (p.denoise_steps > 0) {
        // grad = score(float32(x), eps); x = x + (0 - g^2 grad) * ((1 - eps) / denoise_steps)
        const float eps32 = kSamplingEps;
        compute_time_bias(s, eps32, p.W);
        const float sigma = sigma_of_t(eps32);
        const float g = sigma * kGCoef, g2 = g * g, stdv = sigma + 1e-7f;
        const double dt = (1.0 - 1e-5) / (double)p.denoise_steps;
        for (int lt = 0; lt < p.tiles_per_cta; ++lt) {
            const int tile = blockIdx.x + lt * gridDim.x;
            if (tile >= n_tiles) break;
            for (int i = tid; i < kRT * 9; i += kNT) {
                const int r = i / 9, c = i % 9, row = tile * kRT + r;
                s.x[r * 12 + c] = row < p.R ? (float)p.y[(size_t)row * 9 + c] : 0.f;
            }
            __syncthreads();
            trunk_tile(s, s.tb, tile * kRT, p.R, p.K, p.obj_bias, p.W);
            for (int i = tid; i < kRT * 9; i += kNT) {
                const int r = i / 9, c = i % 9, row = tile * kRT + r;
                if (row < p.R) {
                    const float drift = 0.0f - g2 * (s.f[r * 12 + c] / stdv);       // fp32 (:215)
                    p.y[(size_t)row * 9 + c] = p.y[(size_t)row * 9 + c] + (double)drift * dt;   // float64 + fp32*python float
                }
            }
            __syncthreads();
        }
        ++nfev;
    }
    for (int lt = 0; lt < p.tiles_per_cta; ++lt) {
        const int tile = blockIdx.x + lt * gridDim.x;
        if (tile >= n_tiles) break;
        if (tid < kRT) {
            const int row = tile * kRT + tid;
            if (row < p.R) {
                double v[9];
                for (int c = 0; c < 9; ++c) v[c] = p.y[(size_t)row * 9 + c];
                // normalize_rotation in float64 (x is float64 at this point, :225)
                const double n1 = fmax(sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]), 1e-12);
                const double b0 = v[0] / n1, b1 = v[1] / n1, b2 = v[2] / n1;
                const double d = b0 * v[3] + b1 * v[4] + b2 * v[5];
                const double c0 = v[3] - d * b0, c1 = v[4] - d * b1, c2 = v[5] - d * b2;
                const double n2 = fmax(sqrt(c0 * c0 + c1 * c1 + c2 * c2), 1e-12);
                double *o = p.pose + (size_t)row * 9;
                o[0] = b0; o[1] = b1; o[2] = b2;
                o[3] = c0 / n2; o[4] = c1 / n2; o[5] = c2 / n2;
                const float *ctr = p.pts_center + (size_t)(row / p.K) * 3;
                for (int c = 6; c < 9; ++c) o[c] = v[c] + (double)ctr[c - 6];       // (:226)
            }
        }
    }
    if (p.stats && blockIdx.x == 0 && tid == 0) {
        p.stats[0] = nfev;
        p.stats[1] = n_acc;
        p.stats[2] = n_rej;
        p.stats[3] = status;
    }
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
static int device_sm_count(int *out) {
    int dev = 0;
    GPB_CUDA(cudaGetDevice(&dev));
    GPB_CUDA(cudaDeviceGetAttribute(out, cudaDevAttrMultiProcessorCount, dev));
    return GPB_OK;
}

int launch_time_bias_table(const float *ts, int T, const float *W, float *table, cudaStream_t st) {
    time_bias_table_kernel<<<T, kNT, 0, st>>>(ts, W, table);
    GPB_LAUNCHED();
    return GPB_OK;
}

}  // namespace gpb

using namespace gpb;

extern "C" size_t gpb_trunk_weights_floats(void) { return TL::total; }

extern "C" size_t gpb_sampler_workspace_bytes(int R, int num_steps) {
    return R > 0 ? carve_sampler(nullptr, R, num_steps).bytes : 0;
}

extern "C" int gpb_object_bias(const float *pts_feat, int B, const float *W, float *obj_bias, void *stream) {
    GPB_REQUIRE(B >= 0, "object_bias: B < 0");
    if (B == 0) return GPB_OK;
    GPB_REQUIRE(pts_feat && W && obj_bias, "object_bias: NULL buffer");
    object_bias_kernel<<<dim3((B + 7) / 8, 3), 1024, 0, (cudaStream_t)stream>>>(pts_feat, B, W, obj_bias);
    GPB_LAUNCHED();
    return GPB_OK;
}

static int launch_trunk_eval(const float *pose, int R, int K, float t, const float *obj_bias, const float *W, int divide_mode,
                             const float *center_sub, float *out, float *energy, cudaStream_t st) {
    const size_t smem = kTrunkSmemFloats * sizeof(float);
    GPB_CUDA(cudaFuncSetAttribute(trunk_eval_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int sms = 0, rc;
    if ((rc = device_sm_count(&sms))) return rc;
    const int n_tiles = (R + kRT - 1) / kRT;
    const int grid = n_tiles < sms ? n_tiles : sms;
    trunk_eval_kernel<<<grid, kNT, smem, st>>>(pose, R, K, t, obj_bias, W, divide_mode, center_sub, out, energy);
    GPB_LAUNCHED();
    return GPB_OK;
}

extern "C" int gpb_trunk_eval(const float *pose, int R, int K, float t, const float *obj_bias, const float *W,
                              int divide_mode, float *out, void *stream) {
    GPB_REQUIRE(R >= 0 && K >= 1, "trunk_eval: need R >= 0, K >= 1");
    if (R == 0) return GPB_OK;
    GPB_REQUIRE(pose && obj_bias && W && out, "trunk_eval: NULL buffer");
    GPB_REQUIRE(divide_mode >= 0 && divide_mode <= 2, "trunk_eval: divide_mode must be 0, 1 or 2");
    return launch_trunk_eval(pose, R, K, t, obj_bias, W, divide_mode, nullptr, out, nullptr, (cudaStream_t)stream);
}

extern "C" int gpb_energy(const float *pose, int R, int K, float t, const float *obj_bias, const float *W,
                          const float *pts_center, float *energy, void *stream) {
    GPB_REQUIRE(R >= 0 && K >= 1, "energy: need R >= 0, K >= 1");
    if (R == 0) return GPB_OK;
    GPB_REQUIRE(pose && obj_bias && W && pts_center && energy, "energy: NULL buffer");
    return launch_trunk_eval(pose, R, K, t, obj_bias, W, 2, pts_center, nullptr, energy, (cudaStream_t)stream);
}

extern "C" int gpb_sample_pc(const float *x0, int R, int K, int num_steps, float snr, const float *obj_bias,
                             const float *W, const float *pts_center, const float *step_noise, uint64_t seed,
                             const float *time_grid, float *mean_x, float *process, void *workspace,
                             size_t workspace_bytes, void *stream) {
    GPB_REQUIRE(R >= 0 && K >= 1 && num_steps >= 2, "sample_pc: need R >= 0, K >= 1, num_steps >= 2");
    if (R == 0) return GPB_OK;
    GPB_REQUIRE(x0 && obj_bias && W && pts_center && time_grid && mean_x && workspace, "sample_pc: NULL buffer");
    GPB_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "sample_pc: workspace must be 256-byte aligned");
    SamplerWs w = carve_sampler(workspace, R, num_steps);
    if (workspace_bytes < w.bytes) {
        set_error("sample_pc: workspace %zu < required %zu bytes", workspace_bytes, w.bytes);
        return GPB_EWORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    int sms = 0, rc;
    if ((rc = device_sm_count(&sms))) return rc;
    const int n_tiles = (R + kRT - 1) / kRT;
    const int grid = n_tiles < sms ? n_tiles : sms;
    GPB_REQUIRE(grid <= 1024, "sample_pc: more than 1024 SMs?");
    const int tiles_per_cta = (n_tiles + grid - 1) / grid;
    const size_t smem = (kTrunkSmemFloats + (size_t)2 * tiles_per_cta * kRT * 12) * sizeof(float);
    GPB_REQUIRE(smem <= 227 * 1024, "sample_pc: R=%d needs %zu B of shared memory per CTA; split the batch", R, smem);

    GPB_CUDA(cudaMemsetAsync(w.barrier, 0, 256, st));
    if ((rc = launch_time_bias_table(time_grid, num_steps, W, w.tb_table, st))) return rc;

    PcParams p{};
    p.x0 = x0; p.R = R; p.K = K; p.T = num_steps; p.snr = snr;
    p.obj_bias = obj_bias; p.W = W; p.pts_center = pts_center; p.noise = step_noise; p.seed = seed;
    p.ts = time_grid; p.tb_table = w.tb_table; p.partial = w.partial; p.barrier = w.barrier;
    p.mean_x = mean_x; p.process = process; p.tiles_per_cta = tiles_per_cta;
    GPB_CUDA(cudaFuncSetAttribute(pc_sampler_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    void *args[] = {&p};
    GPB_CUDA(cudaLaunchCooperativeKernel((void *)pc_sampler_kernel, dim3(grid), dim3(kNT), args, smem, st));
    g_launches.fetch_add(1);
    return GPB_OK;
}

extern "C" int gpb_sample_ode(const float *x0, int R, int K, double T0, double rtol, double atol, int denoise_steps,
                              const float *obj_bias, const float *W, const float *pts_center, double *pose, int *stats,
                              double *process, int process_cap, const double *t_eval, int n_t_eval,
                              void *workspace, size_t workspace_bytes, void *stream) {
    GPB_REQUIRE(R >= 0 && K >= 1, "sample_ode: need R >= 0, K >= 1");
    if (R == 0) return GPB_OK;
    GPB_REQUIRE(x0 && obj_bias && W && pts_center && pose && workspace, "sample_ode: NULL buffer");
    GPB_REQUIRE(T0 > 1e-5 && rtol > 0 && atol > 0, "sample_ode: need T0 > eps and positive tolerances");
    GPB_REQUIRE(!process || (t_eval ? n_t_eval > 0 : process_cap > 0), "sample_ode: process output needs process_cap > 0 or t_eval / n_t_eval");
    GPB_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "sample_ode: workspace must be 256-byte aligned");
    SamplerWs w = carve_sampler(workspace, R, 1);
    if (workspace_bytes < w.bytes) {
        set_error("sample_ode: workspace %zu < required %zu bytes", workspace_bytes, w.bytes);
        return GPB_EWORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    int sms = 0, rc;
    if ((rc = device_sm_count(&sms))) return rc;
    const int n_tiles = (R + kRT - 1) / kRT;
    const int grid = n_tiles < sms ? n_tiles : sms;
    GPB_REQUIRE(grid <= 1024, "sample_ode: more than 1024 SMs?");
    const size_t smem = kTrunkSmemFloats * sizeof(float);
    GPB_CUDA(cudaMemsetAsync(w.barrier, 0, 256, st));

    OdeParams p{};
    p.x0 = x0; p.R = R; p.K = K; p.T0 = T0; p.rtol = rtol; p.atol = atol; p.denoise_steps = denoise_steps;
    p.obj_bias = obj_bias; p.W = W; p.pts_center = pts_center;
    p.y = w.y; p.ynew = w.ynew; p.Kst = w.Kst; p.partial = reinterpret_cast<double *>(w.partial); p.barrier = w.barrier;
    p.pose = pose; p.stats = stats; p.tiles_per_cta = (n_tiles + grid - 1) / grid;
    p.proc = OdeProcess{process, process ? t_eval : nullptr, process_cap, n_t_eval};
    GPB_CUDA(cudaFuncSetAttribute(ode_sampler_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    void *args[] = {&p};
    GPB_CUDA(cudaLaunchCooperativeKernel((void *)ode_sampler_kernel, dim3(grid), dim3(kNT), args, smem, st));
    g_launches.fetch_add(1);
    return GPB_OK;
}
