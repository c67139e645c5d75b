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

// RK45's dense-output polynomial (scipy/integrate/_ivp/rk.py, RK45.P; RkDenseOutput._call_impl): y(t) = y_old + h * K^T . P . (x, x^2, x^3, x^4),
// x = (t - t_old) / h, K = the seven stage derivatives of the accepted step
static __constant__ double kRkP[7][4] = {
    {1.0, -8048581381.0 / 2820520608.0, 8663915743.0 / 2820520608.0, -12715105075.0 / 11282082432.0},
    {0.0, 0.0, 0.0, 0.0},
    {0.0, 131558114200.0 / 32700410799.0, -68118460800.0 / 10900136933.0, 87487479700.0 / 32700410799.0},
    {0.0, -1754552775.0 / 470086768.0, 14199869525.0 / 1410260304.0, -10690763975.0 / 1880347072.0},
    {0.0, 127303824393.0 / 49829197408.0, -318862633887.0 / 49829197408.0, 701980252875.0 / 199316789632.0},
    {0.0, -282668133.0 / 205662961.0, 2019193451.0 / 616988883.0, -1453857185.0 / 822651844.0},
    {0.0, 40617522.0 / 29380423.0, -110615467.0 / 29380423.0, 69997945.0 / 29380423.0}};

// Trajectory output of the ODE samplers — the reference's `xs` (samplers.py:206, :220-224).  All zero when not requested.
//   t_eval == NULL : out [cap][R][9], state 0 = the start, state i = the solver's i-th ACCEPTED step (solve_ivp's res.y when
//                    t_eval is None); states beyond cap are dropped (stats[1] tells how many exist)
//   t_eval != NULL : out [n_eval][R][9], the dense output at t_eval (np.linspace(T0, eps, num_steps), strictly decreasing float64)
// Every state is written as the reference returns it: rotation part Gram-Schmidt-normalised in float64, translation + pts_center.
struct OdeProcess {
    double *out;
    const double *t_eval;
    int cap, n_eval;
};

__device__ __forceinline__ void ode_write_state(double *o, const double *v, const float *ctr) {
    const double n1 = fmax(sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]), 1e-12);
    const double b0 = v[0] / n1, b1 = v[1] / n1, b2 = v[2] / n1;
    const double d = b0 * v[3] + b1 * v[4] + b2 * v[5];
    const double c0 = v[3] - d * b0, c1 = v[4] - d * b1, c2 = v[5] - d * b2;
    const double n2 = fmax(sqrt(c0 * c0 + c1 * c1 + c2 * c2), 1e-12);
    o[0] = b0; o[1] = b1; o[2] = b2;
    o[3] = c0 / n2; o[4] = c1 / n2; o[5] = c2 / n2;
#pragma unroll
    for (int c = 6; c < 9; ++c) o[c] = v[c] + (double)ctr[c - 6];
}

// One row's trajectory output for an accepted step from (t_old, y_old) to (t_new, y_new) with stage derivatives K_0..K_6
// (K(j, c) returns component c of stage j).  `n_acc` = number of accepted steps including this one; `te_next` = first t_eval
// index not yet emitted (advanced identically by every caller: pass write = false for rows that only keep the count).
template <typename KFn>
__device__ __noinline__ void ode_emit_step(const OdeProcess &pr, int R, int row, const float *ctr, bool write, int n_acc, int &te_next,
                                           double t_old, double t_new, double h, const double *y_old, const double *y_new, KFn K) {
    if (pr.t_eval == nullptr) {
        if (write && n_acc < pr.cap) ode_write_state(pr.out + ((size_t)n_acc * R + row) * 9, y_new, ctr);
        return;
    }
    while (te_next < pr.n_eval) {
        const double te = __ldg(pr.t_eval + te_next);
        if (!(te >= t_new)) break;                         // direction < 0: this step covers t_eval values in [t_new, t_old]
        if (write) {
            const double x = (te - t_old) / h;
            const double p1 = x, p2 = x * x, p3 = p2 * x, p4 = p3 * x;
            double v[9];
#pragma unroll
            for (int c = 0; c < 9; ++c) v[c] = 0.0;
#pragma unroll
            for (int j = 0; j < 7; ++j) {
                const double w = kRkP[j][0] * p1 + kRkP[j][1] * p2 + kRkP[j][2] * p3 + kRkP[j][3] * p4;
#pragma unroll
                for (int c = 0; c < 9; ++c) v[c] += K(j, c) * w;
            }
#pragma unroll
            for (int c = 0; c < 9; ++c) v[c] = y_old[c] + h * v[c];
            ode_write_state(pr.out + ((size_t)te_next * R + row) * 9, v, ctr);
        }
        ++te_next;
    }
}

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
