// Pieces shared by the two predictor-corrector sampler kernels (fp32 FFMA parity kernel in scorenet.cu and the
// tcgen05 kernel in tc_sampler.cu): launch parameters, per-row noise, Gram-Schmidt, the per-row update.
#pragma once
#include "common.cuh"

namespace gpb {

// F.normalize(v, eps=1e-12) pieces of pytorch3d rotation_6d_to_matrix as used by normalize_rotation
// (utils/misc.py:259-265): b1 = a1/|a1|, b2 = normalize(a2 - (b1.a2) b1).
__device__ __forceinline__ void gram_schmidt6(float *v) {
    const float n1 = fmaxf(sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]), 1e-12f);
    const float b0 = v[0] / n1, b1 = v[1] / n1, b2 = v[2] / n1;
    const float d = b0 * v[3] + b1 * v[4] + b2 * v[5];
    const float c0 = v[3] - d * b0, c1 = v[4] - d * b1, c2 = v[5] - d * b2;
    const float n2 = fmaxf(sqrtf(c0 * c0 + c1 * c1 + c2 * c2), 1e-12f);
    v[0] = b0; v[1] = b1; v[2] = b2;
    v[3] = c0 / n2; v[4] = c1 / n2; v[5] = c2 / n2;
}

// the same with MUFU.RSQ reciprocals instead of IEEE sqrt + divisions (<= 2 ulp per factor); min(., 1e12) is the eps clamp
// (x / max(n, 1e-12)), including n = 0 -> 0.  Used on the tcgen05 sampler's serial update path (tc_sampler.cu).
__device__ __forceinline__ void gram_schmidt6_rsq(float *v) {
    const float i1 = fminf(rsqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]), 1e12f);
    const float b0 = v[0] * i1, b1 = v[1] * i1, b2 = v[2] * i1;
    const float d = b0 * v[3] + b1 * v[4] + b2 * v[5];
    const float c0 = v[3] - d * b0, c1 = v[4] - d * b1, c2 = v[5] - d * b2;
    const float i2 = fminf(rsqrtf(c0 * c0 + c1 * c1 + c2 * c2), 1e12f);
    v[0] = b0; v[1] = b1; v[2] = b2;
    v[3] = c0 * i2; v[4] = c1 * i2; v[5] = c2 * i2;
}

struct PcParams {
    const float *x0;          // [R,9]
    int R, K, T;
    float snr;
    const float *obj_bias;    // [B,768]
    const float *W;           // trunk weights
    const float *pts_center;  // [B,3]
    const float *noise;       // [T,2,R,9] or null
    uint64_t seed;
    const float *ts;          // [T] time grid (torch.linspace(1, eps, T), computed on the host in fp32)
    const float *tb_table;    // [T,768]
    float *partial;           // [2][gridDim] per-CTA sums of row norms
    unsigned *barrier;        // monotonic arrival counter (zeroed before launch)
    unsigned long long *acc;  // [T] zeroed before launch; tcgen05 sampler only (see tc_sampler.cu, "batch-mean gradient norm")
    float *mean_x;            // [R,9] out
    float *process;           // [R,T,9] out or null
    int tiles_per_cta;
    unsigned long long *dbg;  // optional [2][T][16] cycle stamps of CTA `dbg_cta` (profiling aid; NULL in production)
    int dbg_cta;
};

__device__ __forceinline__ void row_noise(const PcParams &p, int step, int which, int row, float *z) {
    if (p.noise) {
        const float *src = p.noise + (((size_t)step * 2 + which) * p.R + row) * 9;
#pragma unroll
        for (int c = 0; c < 9; ++c) z[c] = __ldg(src + c);
    } else {
        const uint2 key = make_uint2((unsigned)p.seed, (unsigned)(p.seed >> 32));
        float buf[12];
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            const uint4 r = philox4x32_10(make_uint4((unsigned)row, (unsigned)step, (unsigned)(which * 3 + q), 0x47504232u), key);
            const float2 a = box_muller(r.x, r.y), b = box_muller(r.z, r.w);
            buf[4 * q + 0] = a.x; buf[4 * q + 1] = a.y; buf[4 * q + 2] = b.x; buf[4 * q + 3] = b.y;
        }
#pragma unroll
        for (int c = 0; c < 9; ++c) z[c] = buf[c];
    }
}


// One predictor-corrector update of a pose row (cond_pc_sampler, samplers.py:130-158), shared verbatim by both
// kernels so that they differ only in how the score is produced.  x: in/out state, gr: score, returns the
// predictor mean in m (the reference's `mean_x`).
struct PcStepConsts {
    float ls, sq2ls, g, g2, step_size, sqrt_step;
};
__device__ __forceinline__ PcStepConsts pc_step_consts(float grad_norm, float snr_norm, float sigma, float step_size, float sqrt_step) {
    PcStepConsts c;
    const float q = snr_norm / grad_norm;
    c.ls = 2.0f * (q * q);                 // langevin_step_size (:131)
    c.sq2ls = sqrtf(2.0f * c.ls);
    c.g = sigma * kGCoef;                  // ve_sde diffusion (sde.py:20-24)
    c.g2 = c.g * c.g;
    c.step_size = step_size;
    c.sqrt_step = sqrt_step;
    return c;
}
__device__ __forceinline__ void pc_row_update(const PcParams &p, const PcStepConsts &c, int step, int row, float *x, const float *gr, float *m) {
    float z[9];
    row_noise(p, step, 0, row, z);
#pragma unroll
    for (int i = 0; i < 9; ++i) x[i] = (x[i] + c.ls * gr[i]) + c.sq2ls * z[i];          // corrector (:132)
    {
        const float n1 = sqrtf(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);               // (:142-143), no eps
        const float n2 = sqrtf(x[3] * x[3] + x[4] * x[4] + x[5] * x[5]);
        x[0] /= n1; x[1] /= n1; x[2] /= n1;
        x[3] /= n2; x[4] /= n2; x[5] /= n2;
    }
#pragma unroll
    for (int i = 0; i < 9; ++i) m[i] = x[i] + (0.0f - c.g2 * gr[i]) * c.step_size;     // predictor mean (:147-148), sign as written
    row_noise(p, step, 1, row, z);
#pragma unroll
    for (int i = 0; i < 9; ++i) x[i] = m[i] + (c.g * c.sqrt_step) * z[i];              // (:149)
    gram_schmidt6(x);                                                                   // (:152)
}

// Dormand-Prince 5(4) tableau exactly as scipy/integrate/_ivp/rk.py (class RK45) states it
static __constant__ double kRkC[6] = {0.0, 1.0 / 5, 3.0 / 10, 4.0 / 5, 8.0 / 9, 1.0};
static __constant__ double kRkA[6][5] = {
    {0, 0, 0, 0, 0},
    {1.0 / 5, 0, 0, 0, 0},
    {3.0 / 40, 9.0 / 40, 0, 0, 0},
    {44.0 / 45, -56.0 / 15, 32.0 / 9, 0, 0},
    {19372.0 / 6561, -25360.0 / 2187, 64448.0 / 6561, -212.0 / 729, 0},
    {9017.0 / 3168, -355.0 / 33, 46732.0 / 5247, 49.0 / 176, -5103.0 / 18656}};
static __constant__ double kRkB[6] = {35.0 / 384, 0, 500.0 / 1113, 125.0 / 192, -2187.0 / 6784, 11.0 / 84};
static __constant__ double kRkE[7] = {-71.0 / 57600, 0, 71.0 / 16695, -71.0 / 1920, 17253.0 / 339200, -22.0 / 525, 1.0 / 40};

// workspace shared by the samplers (gpb_sampler_workspace_bytes)
struct SamplerWs {
    float *tb_table;     // [T,768]
    float *ts;           // [T]
    float *partial;      // [2*1024] floats (PC)  /  doubles [4*1024] (ODE) share the slot
    unsigned *barrier;   // [64] (256 B)
    double *y, *ynew, *Kst;   // Kst: [4 copies][7][R,9] (the tcgen05 ODE kernel keeps one copy per tile-team rank; the FFMA kernel uses copy 0)
    float *tb_cta;             // [160 CTAs][6][768] per-CTA time-bias scratch of the tcgen05 ODE kernel
    unsigned long long *acc;   // [T] per-step (arrival count | fixed-point norm sum) words of the tcgen05 sampler's grid reduction
    size_t bytes;
};
inline SamplerWs carve_sampler(void *base, int R, int T) {
    SamplerWs w{};
    size_t off = 0;
    auto take = [&](size_t nbytes) {
        char *p = base ? reinterpret_cast<char *>(base) + off : nullptr;
        off += ((nbytes + 255) / 256) * 256;
        return p;
    };
    w.barrier = reinterpret_cast<unsigned *>(take(256));
    w.partial = reinterpret_cast<float *>(take(4 * 1024 * sizeof(double)));
    w.ts = reinterpret_cast<float *>(take((size_t)(T > 0 ? T : 1) * sizeof(float)));
    w.tb_table = reinterpret_cast<float *>(take((size_t)(T > 0 ? T : 1) * 768 * sizeof(float)));
    w.y = reinterpret_cast<double *>(take((size_t)R * 9 * sizeof(double)));
    w.ynew = reinterpret_cast<double *>(take((size_t)R * 9 * sizeof(double)));
    w.Kst = reinterpret_cast<double *>(take((size_t)4 * 7 * R * 9 * sizeof(double)));
    w.tb_cta = reinterpret_cast<float *>(take((size_t)160 * 6 * 768 * sizeof(float)));
    w.acc = reinterpret_cast<unsigned long long *>(take((size_t)(T > 0 ? T : 1) * sizeof(unsigned long long)));
    w.bytes = off;
    return w;
}


// scorenet.cu: tb[i, 0:768] = t_bias(ts[i]) for a whole time grid
int launch_time_bias_table(const float *ts, int T, const float *W, float *table, cudaStream_t st);

}  // namespace gpb
