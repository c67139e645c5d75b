// The two samplers on the 5th-generation tensor cores (tcgen05 + TMEM): predictor-corrector (tc_pc_sampler_kernel, the
// BASELINE metric) and the RK45 probability-flow ODE (tc_ode_sampler_kernel, the reference's shipped recipe).  ONE body
// (tc_sampler_body<kOde, kW16, kTeam>) evaluates the score network for both; they differ in what the row warps do with the score and
// in how the number of evaluations is known.
//
// Same algorithm, launch contract and update code as pc_sampler_kernel / ode_sampler_kernel (scorenet.cu); only the score network's
// dense layers change engine: every layer is evaluated by tcgen05.mma (kind::f16, fp32 accumulation in TMEM) in one of two
// error-compensated arithmetics:
//   kW16 = false ("bf16x3")  A.B ~= Ahi.Bhi + Alo.Bhi + Ahi.Blo   (A = Ahi + Alo, B = Bhi + Blo, all bf16; ~2^-17 per operand)
//   kW16 = true  ("f16x2")   A.B ~= Ahi.W + Alo.W                 (A = Ahi + Alo fp16, W one fp16 image; layer 0 stays bf16x3)
// Measured against the oracle: bf16x3 moves final poses by 2e-5 (single bf16: 9e-3, tf32: 8e-4) against the 1e-3
// parity bound; f16x2 spends more of that bound (0.04-0.9 of it on the damped synthetic checkpoints) for 2/3 of the MMAs and half
// the weight stream.
//
// A TEAM of kTeam CTAs (4, 2 or 1; the team is a thread-block cluster) shares a 128-row tile of candidates for all T steps: every
// rank runs layers 0 and 1 (P1, P2) redundantly — they are the sequential prefix — and then only ITS 768 / kTeam columns of the
// 768 stacked head units (team 4: N = 128 + N = 64 MMAs), so the head phase costs 1 / kTeam of the tensor time.  The ranks'
// partial score components are exchanged all-to-all through DISTRIBUTED SHARED MEMORY (st.async into every rank's mailbox, bytes
// counted by the receiver's mbarrier) and summed in a fixed rank order; every rank then applies the update redundantly and
// bit-identically (same noise stream, same batch-mean norm, which rank 0 of every tile publishes to the grid-wide reduction), so
// the new pose never has to be sent back.  A team of 1 has no exchange and no redundancy: one tile per SM, the throughput
// configuration.  Inside a CTA nothing of the per-step state leaves the SM:
//   * activations never touch shared memory: the A operand of every layer lives in TENSOR MEMORY (written by the
//     epilogue with tcgen05.st, lane = row, two 16-bit values per 32-bit column; consumed by the TS form of tcgen05.mma),
//     TMEM map: D0 [0,128) D1 [128,256) accumulators (N = 128 "units", ping-pong), A_hi [256,384), A_lo [384,512);
//   * the whole 227 KB of shared memory is therefore free for the weight stream: 9-12 slots x 16 KiB of pre-tiled operand images
//     (L2-resident) fetched with cp.async.bulk on mbarriers, up to a ring ahead;
//   * pose state, noise and score of a row live in the registers of "its" thread.
// Roles (warp-specialised, 320 threads):
//   warps 0-7  row warps: warp w owns TMEM lanes 32*(w%4).. (rows) and the column sub-half w/4 of every unit.
//              Layer 0: accumulator -> bias + ReLU -> hi/lo -> A operand in QUARTERS (layer 1 starts on the first 32 converted
//              columns per thread).  Layer 1: the same in halves (unit a is held in registers until the layer's last MMA has
//              consumed the old A).  Heads: relu(acc + obj_bias + t_bias) . O on packed fp32 pairs (FADD2 / FFMA2), overlapped
//              with the next unit's MMAs.  Warps 0-3 then run the grid-wide gradient-norm reduction and the Langevin /
//              Euler-Maruyama update (noise for the step is generated while the tensor core works); warps 4-7 meanwhile refresh
//              the (object bias + time bias) table for the next step.
//   warp 8     one elected thread issues every tcgen05.mma / tcgen05.commit; the warp owns the TMEM allocation.  The tensor pipe
//              accepts an MMA only about when it starts it, so whatever this warp does between two issue groups idles the pipe:
//              straight-line issue code over a slot schedule computed under the wait for x, few large groups (f16x2: layer 1 =
//              q0 | q1 | q2 + q3 + unit b, team-of-4 heads = first half | second half + 64-column unit; three products: one slot
//              wait per unit, the ring holds no more), an accumulator wait only where no operand hand-off already implies its release.
//   warp 9     one elected thread runs the weight producer.
// Per step and CTA (team 4, f16x2): layer 0 (10 MMAs), layer 1 (2 x 32 MMAs of 128x128x16), head slice (32 of 128x128x16 + 32 of 128x64x16).
// History (3200 rows x 500 steps, one B200): A in shared memory + 2-deep weight ring 18.2 ms ... round 1 7.05 ms ... this version
// 4.46 ms: DESIGN.md §5.
#include <cstdlib>
#include <initializer_list>
#include <type_traits>

#include "common.cuh"
#include "sampler_common.cuh"
#include "tc_common.cuh"

namespace gpb {
using namespace tc;

constexpr int kTcRows = 128;
constexpr int kTcRowWarps = 8;
constexpr int kTcThreads = (kTcRowWarps + 2) * 32;
constexpr uint32_t kSlotBytes = 16384;
// Weight stream geometry.  A 128-row tile of candidates is owned by a TEAM of kTeam CTAs (4, 2 or 1; the team is a thread-block
// cluster) for all steps: every rank runs layers 0 and 1 and then ITS kCols = 768 / kTeam of the 768 stacked head units.  The team
// size is chosen from the row count (launch_tc_sampler): 4 while one 4-CTA cluster per tile fits the device (<= 33 tiles on a
// B200: short dependent chain per step), 2 up to 66 tiles, 1 beyond (one tile per SM, no redundant layer 1, no exchange: the
// throughput configuration, 148 tiles = 18,944 rows).
// kW16 = false ("bf16x3"): every weight block as bf16 hi image | lo image, three products per K-step.  kW16 = true ("f16x2"): the
// weights of layer 1 and of the heads as ONE fp16 image (11-bit mantissa) against fp16 hi/lo activations, two products per K-step
// (Ahi.W + Alo.W), so a step streams 33 slots instead of 65 and issues 2/3 of the MMAs.  (kind::f16 takes ONE format for both
// operands on this part: bf16 activations against fp16 weights raise an illegal-instruction fault, measured in round 2.)
// Layer 0 (K = 16, ten small MMAs) keeps the bf16 hi | lo form in both.
template <bool kW16, int kTeam> struct TcStream {
    static_assert(kTeam == 1 || kTeam == 2 || kTeam == 4, "team size");
    static constexpr int kCols = 768 / kTeam;                   // head columns of one rank
    static constexpr int kFullUnits = kCols / 128;              // N = 128 accumulator units of the head slice (1, 3, 6)
    static constexpr bool kSmallUnit = (kCols % 128) != 0;      // team 4: one more unit of N = 64
    static constexpr int kHeadUnits = kFullUnits + (kSmallUnit ? 1 : 0);
    static constexpr int kCommonSlots = 1 + (kW16 ? 8 : 16);   // P1 (both units) + P2 (2 units x 8 K-chunks of 32; kW16: x 4 slots of K = 64)
    static constexpr int kHeadSlots = kFullUnits * (kW16 ? 4 : 8) + (kSmallUnit ? (kW16 ? 2 : 4) : 0);
    static constexpr int kSlotsPerCtaStep = kCommonSlots + kHeadSlots;        // team 4: 29 slots = 464 KiB per CTA and step (kW16: 15)
    static constexpr int kSlotsPerStep = kCommonSlots + kTeam * kHeadSlots;   // 65 slots in the global stream (kW16: 33), any team
    // the stream buffer holds two layouts back to back: A (team 4: per rank a 128-unit and a 64-unit) then B (teams 2 and 1: six
    // 128-units in column order)  (genpose_b200/weights.py::pack_trunk_tc / pack_trunk_tc16)
    static constexpr int kLayoutSlot0 = kTeam == 4 ? 0 : kSlotsPerStep;
};
constexpr uint32_t kLboB64 = 1024;                 // 64-row operand images
constexpr uint32_t kLboB = 2048, kSbo = 128;       // 128-row operand images
constexpr int kMaxObjPerTile = 8;                 // objects a 128-row tile may span: K >= 19 candidates per object
constexpr uint32_t kColD = 0, kColAhi = 256, kColAlo = 384;
// x (three bf16 pieces, 8 columns each) lives in the LAST 24 columns of A_lo: layer 0's epilogue overwrites them only with its third
// h1 quarter, which needs the second layer-0 accumulator anyway — so the first two quarters go out as soon as the FIRST accumulator
// is complete, while the tensor pipe is still on the second (x at the front of A_hi made quarter 0 wait for every layer-0 MMA)
constexpr uint32_t kColX = kColAlo + 104;
// per-step grid reduction word of the PC kernel: [63:56] arrived tiles, [55:48] poisoned tiles, [47:0] sum of the tile sums as
// 22.18 fixed point of (tile sum x min(sigma(t), 1)) x up to 255 tiles (a tile sum is 128 row norms |f| / sigma: the limit 2^22 is
// 32,768 per row on average at sigma >= 1 and 3.3 M per row at sigma = 0.01)
constexpr float kNormSumLimit = 4194304.f, kNormSumScale = 262144.f;
using TL = TrunkLayout;

// dynamic shared memory map (bytes): [8 objects][kCols] fp32 obj_bias + t_bias(step) of the rank's head columns | [9][256] fp32 output
// layer + [16] bias | p1_b [256] | p2_b [256] | [128][12] fp32 scratch (partial scores of column sub-half 1, ODE time-bias scratch) |
// PC: [18][128] fp32 noise of the step |
// the team's mailboxes [2 parities][team][128][8] fp32 (DSMEM; teams of 2 and 4) | ODE: float64 state y [9][128] | y_new [9][128] | the
// weight ring, which takes what is left (team 4: 10 slots PC / 9 ODE; team 1: 11 / 10)
template <bool kOde, int kTeam> struct TcSmem {
    static constexpr uint32_t kCols = 768u / kTeam;
    static constexpr uint32_t kOffObt = 0;
    static constexpr uint32_t kOffOw = kOffObt + kMaxObjPerTile * kCols * 4u;
    static constexpr uint32_t kOffBias = kOffOw + (9u * 256u + 16u) * 4u;
    static constexpr uint32_t kOffFpart = kOffBias + 512u * 4u;
    static constexpr uint32_t kOffNoise = kOffFpart + 128u * 12u * 4u;           // PC: [18][128] fp32 z1 | z2 of the step (column half 1 -> half 0)
    static constexpr bool kNoiseByHalf1 = !kOde && kTeam > 1;   // (a team of 1 has no idle tail for half 1 and would pay a ring slot)
    static constexpr uint32_t kOffMail = kOffNoise + (kNoiseByHalf1 ? 18u * 128u * 4u : 0u);
    static constexpr uint32_t kMailBytes = kTeam > 1 ? 2u * kTeam * 128u * 8u * 4u : 0u;
    static constexpr uint32_t kOffOdeY = kOffMail + kMailBytes;
    static constexpr uint32_t kOffRing = ((kOffOdeY + (kOde ? 2u * 9u * 128u * 8u : 0u)) + 1023u) & ~1023u;
    static constexpr int kSlots = (int)((227u * 1024u - 1280u - kOffRing) / kSlotBytes);
    static constexpr uint32_t kBytes = kOffRing + kSlots * kSlotBytes;
};
static_assert(TcSmem<false, 4>::kSlots == 10 && TcSmem<true, 4>::kSlots == 9 && TcSmem<false, 1>::kSlots == 11 && TcSmem<true, 1>::kSlots == 10,
              "tc sampler shared memory budget");

// probability-flow ODE mode (cond_ode_sampler, samplers.py:163-227): everything PcParams does not already carry
struct TcOdeParams {
    double T0, rtol, atol;
    int denoise_steps;
    double *Kst;        // [4 ranks][7 stages][R,9] float64 stage derivatives, one private copy per tile-team rank
    double *partial;    // [4 slots][n_tiles][2] per-tile partial sums of the step controller's norms
    float *tb_cta;      // [gridDim][6][768] time biases of the evaluation group in flight (private per CTA)
    double *pose;       // [R,9] float64 out (samplers.py:206-207)
    int *stats;         // [4] nfev, accepted, rejected, status (optional)
    OdeProcess proc;    // optional trajectory output (the reference's in_process_sample)
};

struct TcPcParams {
    PcParams pc;
    const uint8_t *wstream;   // TcStream::kSlotsPerStep x 16 KiB of operand images (genpose_b200/weights.py::pack_trunk_tc / pack_trunk_tc16)
    TcOdeParams ode;          // used by tc_ode_sampler_kernel only
};

__device__ __forceinline__ int ld_volatile_shared(const int *p) {
    int v;
    asm volatile("ld.volatile.shared.s32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
    return v;
}
__device__ __forceinline__ void st_volatile_shared(int *p, int v) {
    asm volatile("st.volatile.shared.s32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}
__device__ __forceinline__ void red_release_add_u32(unsigned *p, unsigned v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void red_relaxed_add_u64(unsigned long long *p, unsigned long long v) {
    asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}


// bias + ReLU + hi/lo split of 32 accumulator columns -> 16 + 16 packed words (column pairs); kF16: fp16 halves (the A operand
// of the two-product layers), else bf16 halves
template <bool kF16>
__device__ __forceinline__ void relu_split32(const uint32_t (&v)[32], const float *bias, uint32_t *hi, uint32_t *lo) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const float2 bb = *reinterpret_cast<const float2 *>(bias + 2 * j);
        float a, b;
        unpack_f32x2(add_f32x2(pack_f32x2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1])), pack_f32x2(bb.x, bb.y)), a, b);   // one FADD2
        if constexpr (kF16) relu_split_f16x2(a, b, hi[j], lo[j]);
        else split_bf16x2(fmaxf(a, 0.f), fmaxf(b, 0.f), hi[j], lo[j]);
    }
}

// One body, two kernels: kOde = false is the predictor-corrector sampler (tc_pc_sampler_kernel), kOde = true the RK45
// probability-flow ODE sampler (tc_ode_sampler_kernel).  They share the score-network evaluation (weight producer, MMA issuer,
// epilogues, team exchange) and differ in what the row warps do with the score and in how the trip count is known:
// PC runs exactly T evaluations; the ODE solver's count depends on its step controller, so the row warps publish how many
// evaluations are known to exist (s_allowed, always at least one ahead of every decision point) and when the last one is (s_final).
template <bool kOde, bool kW16, int kTeam>
__device__ __forceinline__ void tc_sampler_body(const TcPcParams &tp) {
    using TS = TcStream<kW16, kTeam>;
    using SM = TcSmem<kOde, kTeam>;
    constexpr int kCols = TS::kCols;
    const PcParams &p = tp.pc;
    extern __shared__ __align__(1024) uint8_t smem[];
    constexpr int kSlots = SM::kSlots;
    uint8_t *sRing = smem + SM::kOffRing;
    float *sObt = reinterpret_cast<float *>(smem + SM::kOffObt);
    float *sOw = reinterpret_cast<float *>(smem + SM::kOffOw);
    float *sBias = reinterpret_cast<float *>(smem + SM::kOffBias);
    float *sFpart = reinterpret_cast<float *>(smem + SM::kOffFpart);
    float *sMail = reinterpret_cast<float *>(smem + SM::kOffMail);
    float *sNoise = reinterpret_cast<float *>(smem + SM::kOffNoise);       // PC only
    __shared__ __align__(8) uint64_t bar_full[kSlots], bar_empty[kSlots], bar_acc_full[2], bar_acc_empty[2], bar_x_ready, bar_a_ready[2],
        bar_h1_ready[4], bar_mail[2];
    __shared__ uint32_t s_tmem_base;
    __shared__ float s_red[8];
    // ODE mode: evaluation bookkeeping shared with the producer / MMA warps, the evaluation group in flight, fp64 reduction scratch
    __shared__ int s_allowed, s_final, s_gn;
    __shared__ float s_times[8];
    __shared__ double s_redd[12];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tile = blockIdx.x / kTeam, rank = kTeam > 1 ? (int)cluster_ctarank() : 0;   // cluster = tile team (launch: cluster dims kTeam x 1 x 1); rank 0 leads
    const int n_tiles = gridDim.x / kTeam;
    const int row0 = tile * kTcRows;
    const int obj_lo = row0 / p.K;
    const int n_obj = (min(row0 + kTcRows, p.R) - 1) / p.K - obj_lo + 1;
    const float *W = p.W;
    const int n_lo = rank * kCols;                // this rank's slice [n_lo, n_lo + kCols) of the 768 stacked head units
    const int hA = n_lo / 256;                    // first head the slice touches (teams 2 and 4: at most two heads, hA and hA + 1)

    if (tid == 0) {
        for (int s = 0; s < kSlots; ++s) {
            mbar_init(&bar_full[s], 1);
            mbar_init(&bar_empty[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&bar_acc_full[b], 1);
            mbar_init(&bar_acc_empty[b], kTcRowWarps);
        }
        mbar_init(&bar_x_ready, 4);
        // one barrier per HALF of a freshly written A operand: a warp's first-half and second-half arrivals must not be
        // interchangeable (four fast warps arriving twice would otherwise complete the first-half phase for all eight)
        mbar_init(&bar_a_ready[0], kTcRowWarps);
        mbar_init(&bar_a_ready[1], kTcRowWarps);
        for (int qd = 0; qd < 4; ++qd) mbar_init(&bar_h1_ready[qd], kTcRowWarps);   // layer 0 -> layer 1: the A operand in QUARTERS
        mbar_init(&bar_mail[0], 1);                // one local arrive.expect_tx per use; the peers' st.async complete the bytes
        mbar_init(&bar_mail[1], 1);
        fence_mbar_init();
        if constexpr (kOde) {
            s_allowed = 2 + 6 + 1;     // f0, f1 (select_initial_step), the first RK45 attempt, and whatever follows it
            s_final = 0;
            s_gn = 1;
            s_times[0] = (float)tp.ode.T0;
        }
    }
    if (warp == kTcRowWarps) tmem_alloc(&s_tmem_base, 512);
    if constexpr (!kOde) {
        for (int i = tid; i < n_obj * kCols; i += kTcThreads)                                                                        // step 0
            sObt[i] = p.obj_bias[(size_t)(obj_lo + i / kCols) * 768 + n_lo + i % kCols] + p.tb_table[n_lo + i % kCols];
    }
    for (int i = tid; i < 9 * 256 + 12; i += kTcThreads) sOw[i] = W[TL::o_w + i];   // o_w then o_b are adjacent
    for (int i = tid; i < 256; i += kTcThreads) {
        sBias[i] = W[TL::p1_b + i];
        sBias[256 + i] = W[TL::p2_b + i];
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    if constexpr (kTeam > 1) cluster_sync_all();          // the team's mailbox barriers are initialised before anyone arrives remotely
    const uint32_t tmem_base = s_tmem_base;
    const uint32_t idesc128 = make_idesc_bf16_f32(128, 128), idesc64 = make_idesc_bf16_f32(128, 64);
    // kW16: A = fp16 hi / lo (tensor memory), B = fp16 (shared memory)
    const uint32_t idesc128w = make_idesc_f16kind_f32(128, 128, 0, 0), idesc64w = make_idesc_f16kind_f32(128, 64, 0, 0);
    const bool dbg_cta = p.dbg != nullptr && (int)blockIdx.x == p.dbg_cta;

    if (warp == kTcRowWarps + 1) {
        // =============================== weight producer ===============================
        if (lane == 0) {
            auto produce = [&](uint32_t it) {
                const uint32_t s = it % kSlots;
                const uint32_t idx = it % TS::kSlotsPerCtaStep;
                const uint32_t src = (uint32_t)TS::kLayoutSlot0 + (idx < (uint32_t)TS::kCommonSlots ? idx : idx + (uint32_t)(rank * TS::kHeadSlots));
                mbar_wait(&bar_empty[s], ((it / kSlots) & 1u) ^ 1u);
                mbar_arrive_expect_tx(&bar_full[s], kSlotBytes);
                bulk_g2s(sRing + s * kSlotBytes, tp.wstream + (size_t)src * kSlotBytes, kSlotBytes, &bar_full[s]);
            };
            if constexpr (!kOde) {
                const uint32_t total = (uint32_t)p.T * TS::kSlotsPerCtaStep;
                for (uint32_t it = 0; it < total; ++it) produce(it);
            } else {
                // stream the weights of evaluation e only once it is known to exist; the row warps keep s_allowed one evaluation
                // ahead of every decision point, so this loop only ever waits at the very end (no copy is left in flight at exit)
                bool more = true;
                for (int e = 0; more; ++e) {
                    while (e >= ld_volatile_shared(&s_allowed)) {
                        if (ld_volatile_shared(&s_final)) {
                            more = false;
                            break;
                        }
                    }
                    if (!more) break;
                    for (uint32_t k = 0; k < (uint32_t)TS::kSlotsPerCtaStep; ++k) produce((uint32_t)e * TS::kSlotsPerCtaStep + k);
                }
            }
        }
    } else if (warp == kTcRowWarps) {
        // =============================== MMA issuer ===============================
        // The WHOLE warp runs this loop (uniform control flow => descriptors and TMEM addresses live in uniform registers);
        // one elected lane issues each group of tcgen05.mma / tcgen05.commit.
        const uint32_t ring = smem_u32(sRing);
        const uint32_t t_ahi = tmem_base + kColAhi, t_alo = tmem_base + kColAlo, t_x = tmem_base + kColX;
        uint32_t u = 0, it = 0, xr = 0, ar = 0, hr = 0;
        // wait for n_slots consecutive ring slots starting at stream position `it`: lane l polls slot l (one wait latency)
        auto wait_slots = [&](int n_slots, int first = 0) {
            if (lane < n_slots) {
                const uint32_t itl = it + (uint32_t)(first + lane);
                mbar_wait(&bar_full[itl % kSlots], (itl / kSlots) & 1u);
            }
            __syncwarp();
            tc_fence_after_sync();
        };
        // ---- straight-line issue code ----
        // The tensor pipe accepts an MMA only about when it starts it (issue time = execution time), so everything this thread does
        // between two issue groups idles the pipe.  Hence: (1) the step's slot schedule — ring position, descriptor address field and
        // release barrier of each of its slots — is computed at the top of the step, under the wait for x, and every loop below is
        // unrolled over it (no index arithmetic behind a wait); (2) few, large groups: with f16x2 (one weight image: layer 1 is 8 slots,
        // a team-of-4 rank's head slice 6) layer 1 is q0 | q1 | q2 + q3 + unit b behind ONE slot wait and the team-of-4 head slice
        // first half | second half + 64-column unit behind one more; with three products a unit is 8 slots (layer 1 alone is more
        // than the ring holds), so every unit has its own slot wait; (3) no accumulator waits where an operand hand-off implies the
        // release (see the row warps).
        constexpr int NS = TS::kSlotsPerCtaStep;
        constexpr int kUS = kW16 ? 4 : 8;                                                       // slots of a 128-column unit (K = 256)
        constexpr int kKpS = 16 / kUS;                                                          // K-steps (16 inputs) per slot
        constexpr uint32_t kDescHi = ((kSbo >> 4) & 0x3FFFu) | (1u << 14);                      // SBO | descriptor version (bit 46)
        constexpr uint32_t kLf128 = ((kLboB >> 4) & 0x3FFFu) << 16, kLf64 = ((kLboB64 >> 4) & 0x3FFFu) << 16;
        constexpr bool kFuseL1 = kW16, kFuseHeads = kW16 && TS::kSmallUnit;
        static_assert(kUS <= kSlots && (!kFuseL1 || 2 * kUS <= kSlots), "slot waits must fit the ring");
        static_assert(!kFuseHeads || (TS::kHeadSlots <= kSlots && TS::kHeadUnits == 2), "a team-of-4 head slice waits for all its slots at once");
        static_assert(!TS::kSmallUnit || TS::kFullUnits == 1, "the 64-column unit exists in teams of 4 only (head unit 1)");
        auto desc = [&](uint32_t lo) { return ((uint64_t)kDescHi << 32) | (uint64_t)lo; };
        const uint32_t empty0 = smem_u32(&bar_empty[0]);
        uint32_t sring = 0;                                                                     // ring position of the step's first slot
        for (int step = 0; kOde || step < p.T; ++step) {
            unsigned long long *ds = (dbg_cta && lane == 0 && step < p.T) ? p.dbg + ((size_t)p.T + step) * 16 : nullptr;   // ODE: T = recorded evaluations
            unsigned long long w_full = 0, w_a = 0, w_issue = 0, tq = 0;
            if (ds) ds[0] = clock64();
            uint32_t dl[NS], eb[NS];
            {
                uint32_t s = sring;
#pragma unroll
                for (int i = 0; i < NS; ++i) {
                    dl[i] = ((ring + s * kSlotBytes) >> 4) & 0x3FFFu;   // (the mask matters: in a cluster the shared::cta window address carries the CTA's rank above bit 18)
                    eb[i] = empty0 + 8u * s;
                    s = s + 1u == (uint32_t)kSlots ? 0u : s + 1u;
                }
                sring = s;
            }
            auto commit_slot = [&](int i) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(eb[i]) : "memory"); };
            // K-step ks (16 inputs) of a 128-column, K = 256 unit whose first slot is `first`.  f16x2: four K-steps per slot, one fp16 image,
            // products Ahi.W + Alo.W; three products: two K-steps per slot, [hi 8 KiB | lo 8 KiB], Ahi.Bhi + Alo.Bhi + Ahi.Blo
            auto kstep128 = [&](uint32_t d, int first, int ks, bool acc) {
                const uint32_t ac = (uint32_t)ks * 8u;                                          // K index / 2
                if constexpr (kW16) {
                    const uint64_t b_w = desc(dl[first + (ks >> 2)] + (uint32_t)((2u * (uint32_t)(ks & 3) * kLboB) >> 4) + kLf128);
                    umma_bf16_ts(d, t_ahi + ac, b_w, idesc128w, acc);
                    umma_bf16_ts(d, t_alo + ac, b_w, idesc128w, true);
                } else {
                    const uint32_t lo = dl[first + (ks >> 1)] + (uint32_t)((2u * (uint32_t)(ks & 1) * kLboB) >> 4) + kLf128;
                    const uint64_t b_hi = desc(lo), b_lo = desc(lo + (8192u >> 4));
                    umma_bf16_ts(d, t_ahi + ac, b_hi, idesc128, acc);
                    umma_bf16_ts(d, t_alo + ac, b_hi, idesc128, true);
                    umma_bf16_ts(d, t_ahi + ac, b_lo, idesc128, true);
                }
            };
            auto unit_ksteps = [&](uint32_t d, int first, int k0, int k1) {                     // K-steps [k0, k1) in order, slots released as they end
#pragma unroll
                for (int ks = k0; ks < k1; ++ks) {
                    kstep128(d, first, ks, ks != 0);
                    if (ks % kKpS == kKpS - 1) commit_slot(first + ks / kKpS);
                }
            };
            // ---- layer 0: h1_pre = x . P1^T   (K = 16; x pieces x1,x2,x3 at kColX + 0, 8, 16; P1 hi|lo per unit) ----
            wait_slots(1);
            mbar_wait_inline(&bar_x_ready, xr & 1u);          // (implies both accumulators free: x follows every head epilogue)
            ++xr;
            tc_fence_after_sync();
            if (ds) { ds[1] = clock64(); tq = clock64(); }
            if (elect_one_sync()) {
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    const uint32_t b = (u + (uint32_t)half) & 1u;
                    const uint32_t d = tmem_base + kColD + b * 128u;
                    const uint64_t bhi = desc(dl[0] + (uint32_t)half * 512u + kLf128), blo = desc(dl[0] + (uint32_t)half * 512u + 256u + kLf128);
                    umma_bf16_ts(d, t_x + 0u, bhi, idesc128, false);
                    umma_bf16_ts(d, t_x + 8u, bhi, idesc128, true);
                    umma_bf16_ts(d, t_x + 16u, bhi, idesc128, true);
                    umma_bf16_ts(d, t_x + 0u, blo, idesc128, true);
                    umma_bf16_ts(d, t_x + 8u, blo, idesc128, true);
                    if (half == 1) commit_slot(0);
                    umma_commit(&bar_acc_full[b]);
                }
            }
            __syncwarp();
            if (ds) { w_issue += clock64() - tq; tq = clock64(); }
            u += 2;
            it += 1;
            // ---- layer 1: unit a on the h1 quarters (quarter g = K-steps {base, base + 1, base + 4, base + 5}, base = 8 (g / 2) + 2 (g % 2):
            //      row thread (q, cs) converts columns [64 cs, 64 cs + 64) of either unit, 32 at a time), unit b behind it ----
            // (three products: the ring holds 10 of a step's 29 slots, so the stream runs just in time — a unit's slots are awaited in two
            // halves, or its first MMAs would wait for slots that can only be refilled once the previous unit is nearly through)
            wait_slots(kFuseL1 ? 2 * kUS : kUS / 2);
            if (ds) { w_full += clock64() - tq; tq = clock64(); }
            {
                const uint32_t da = tmem_base + kColD + (u & 1u) * 128u, db = tmem_base + kColD + ((u + 1u) & 1u) * 128u;
#pragma unroll
                for (int g = 0; g < 3; ++g) {
                    if (!kFuseL1 && g == 2) wait_slots(kUS / 2, kUS / 2);
                    mbar_wait_inline(&bar_h1_ready[g == 2 ? 3 : g], hr & 1u);      // quarters 2 and 3 are both there once quarter 1 is issued
                    tc_fence_after_sync();
                    if (ds) {
                        w_a += clock64() - tq;
                        if (g == 0) ds[3] = clock64();
                        tq = clock64();
                    }
                    if (elect_one_sync()) {
#pragma unroll
                        for (int gg = g; gg < (g == 2 ? 4 : g + 1); ++gg) {
                            const int base = 8 * (gg >> 1) + 2 * (gg & 1);
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const int ks = base + (e & 1) + 4 * (e >> 1);
                                kstep128(da, 1, ks, (gg | e) != 0);
                            }
                            if constexpr (kW16) {
                                if (gg & 1) {   // quarters 2 (gg / 2), 2 (gg / 2) + 1 cover K-steps [8 (gg / 2), +8): two slots done
                                    commit_slot(1 + (gg >> 1) * 2);
                                    commit_slot(2 + (gg >> 1) * 2);
                                }
                            } else {            // two K-steps per slot: the quarter's K-step pairs are slots base / 2 and base / 2 + 2
                                commit_slot(1 + base / 2);
                                commit_slot(1 + base / 2 + 2);
                            }
                        }
                        if (g == 2) {
                            umma_commit(&bar_acc_full[u & 1u]);
                            if constexpr (kFuseL1) {   // unit b: its accumulator was read before h1 quarter 2 was published
                                unit_ksteps(db, 1 + kUS, 0, 16);
                                umma_commit(&bar_acc_full[(u + 1u) & 1u]);
                            }
                        }
                    }
                    __syncwarp();
                    if (ds) { w_issue += clock64() - tq; tq = clock64(); }
                }
                ++hr;
                it += (uint32_t)kUS;
                if constexpr (!kFuseL1) {       // three products: unit b behind its own slot waits (accumulator: see above)
#pragma unroll
                    for (int hf = 0; hf < 2; ++hf) {
                        wait_slots(kUS / 2, hf * (kUS / 2));
                        if (ds) { w_full += clock64() - tq; tq = clock64(); }
                        if (elect_one_sync()) {
                            unit_ksteps(db, 1 + kUS, 8 * hf, 8 * hf + 8);
                            if (hf == 1) umma_commit(&bar_acc_full[(u + 1u) & 1u]);
                        }
                        __syncwarp();
                        if (ds) { w_issue += clock64() - tq; tq = clock64(); }
                    }
                }
                it += (uint32_t)kUS;
                u += 2;
            }
            // ---- this rank's head slice: pf arrives in two halves (K columns [0,128) and [128,256)) ----
            if (ds) tq = clock64();
            static_for<0, TS::kFullUnits>([&](auto HU) {   // (a compile-time loop: `#pragma unroll` leaves six units partly rolled, which
                constexpr int hu = decltype(HU)::value;     //  turns the slot schedule into a local-memory array)
                constexpr int hs = TS::kCommonSlots + kUS * hu;           // the unit's first slot
                const uint32_t b = u & 1u, n = u >> 1;
                const uint32_t d = tmem_base + kColD + b * 128u;
                constexpr int kGroups = (hu == 0 || !kW16) ? 2 : 1;       // K halves: pf arrives in halves (unit 0), three-product slots do
                if (kW16 && (hu == 0 || !kFuseHeads)) wait_slots(hu == 0 && kFuseHeads ? TS::kHeadSlots : kUS);
                if (ds) { w_full += clock64() - tq; tq = clock64(); }
#pragma unroll
                for (int g = 0; g < kGroups; ++g) {
                    if constexpr (!kW16) {
                        wait_slots(kUS / 2, g * (kUS / 2));
                        if (ds) { w_full += clock64() - tq; tq = clock64(); }
                    }
                    if (hu == 0) {
                        mbar_wait_inline(&bar_a_ready[g], ar & 1u);        // (implies the accumulator of head unit g is free)
                        if (g == 1) ++ar;
                        tc_fence_after_sync();
                    } else if (hu >= 2 && g == 0) {
                        mbar_wait_inline(&bar_acc_empty[b], (n & 1u) ^ 1u);   // released by the epilogue of head unit hu - 2
                        tc_fence_after_sync();
                    }
                    if (ds) {
                        w_a += clock64() - tq;
                        if (hu == 0 && g == 0) ds[4] = clock64();
                        tq = clock64();
                    }
                    if (elect_one_sync()) {
                        const int k0 = kGroups == 2 ? 8 * g : 0, k1 = kGroups == 2 ? 8 * g + 8 : 16;
                        unit_ksteps(d, hs, k0, k1);
                        if (k1 == 16) umma_commit(&bar_acc_full[b]);
                        if constexpr (kFuseHeads) {
                            if (g == 1) {   // the 64-column unit (head unit 1: free once pf is complete): one fp16 image of K = 128 per slot
                                const uint32_t d2 = tmem_base + kColD + ((u + 1u) & 1u) * 128u;
#pragma unroll
                                for (int ks = 0; ks < 16; ++ks) {
                                    const int slot = TS::kCommonSlots + 4 + (ks >> 3);
                                    const uint64_t b_w = desc(dl[slot] + (uint32_t)((2u * (uint32_t)(ks & 7) * kLboB64) >> 4) + kLf64);
                                    umma_bf16_ts(d2, t_ahi + (uint32_t)ks * 8u, b_w, idesc64w, ks != 0);
                                    umma_bf16_ts(d2, t_alo + (uint32_t)ks * 8u, b_w, idesc64w, true);
                                    if ((ks & 7) == 7) commit_slot(slot);
                                }
                                umma_commit(&bar_acc_full[(u + 1u) & 1u]);
                            }
                        }
                    }
                    __syncwarp();
                    if (ds) { w_issue += clock64() - tq; tq = clock64(); }
                }
                ++u;
                it += (uint32_t)kUS;
            });
            if constexpr (TS::kSmallUnit) {
                if constexpr (!kFuseHeads) {
                    // three products, the 64-column unit (head unit 1: free once pf is complete): four slots of two K-chunks of 32 inputs,
                    // each [hi 4 KiB | lo 4 KiB] at c * 8 KiB
                    constexpr int first = TS::kCommonSlots + kUS;
                    const uint32_t d2 = tmem_base + kColD + (u & 1u) * 128u;
                    wait_slots(4);
                    if (ds) { w_full += clock64() - tq; tq = clock64(); }
                    if (elect_one_sync()) {
#pragma unroll
                        for (int sl = 0; sl < 4; ++sl) {
#pragma unroll
                            for (int c = 0; c < 2; ++c)
#pragma unroll
                                for (int j = 0; j < 2; ++j) {
                                    const uint32_t lo = dl[first + sl] + (uint32_t)((c * 8192u + 2u * (uint32_t)j * kLboB64) >> 4) + kLf64;
                                    const uint64_t b_hi = desc(lo), b_lo = desc(lo + (4096u >> 4));
                                    const uint32_t ac = (uint32_t)(2 * sl + c) * 16u + 8u * (uint32_t)j;
                                    umma_bf16_ts(d2, t_ahi + ac, b_hi, idesc64, (sl | c | j) != 0);
                                    umma_bf16_ts(d2, t_alo + ac, b_hi, idesc64, true);
                                    umma_bf16_ts(d2, t_ahi + ac, b_lo, idesc64, true);
                                }
                            commit_slot(first + sl);
                        }
                        umma_commit(&bar_acc_full[u & 1u]);
                    }
                    __syncwarp();
                    if (ds) { w_issue += clock64() - tq; tq = clock64(); }
                }
                ++u;
                it += (uint32_t)(kW16 ? 2 : 4);
            }
            if (ds) {
                ds[7] = clock64();
                ds[8] = w_full;
                ds[9] = w_a;
                ds[10] = 0;
                ds[11] = w_issue;
            }
            if constexpr (kOde) {   // the flags were written before the x_ready arrival that released this evaluation
                if (ld_volatile_shared(&s_final) && step + 1 == ld_volatile_shared(&s_allowed)) break;
            }
        }
    } else {
        // =============================== row warps ===============================
        const int q = warp & 3, cs = warp >> 2;
        const int r = q * 32 + lane;              // row of the tile == TMEM lane
        const int row = row0 + r;                 // global candidate row
        const bool valid = row < p.R;
        const bool leader = rank == 0;
        const uint32_t tm_row = tmem_base + ((uint32_t)(q * 32) << 16);
        const float *obt_row = sObt + (size_t)((valid ? row : p.R - 1) / p.K - obj_lo) * kCols - n_lo;   // indexed by the stacked unit n in [n_lo, n_lo + kCols)
        const bool dbg = dbg_cta && tid == 0;

        float x[9];
#pragma unroll
        for (int c = 0; c < 9; ++c) x[c] = (valid && cs == 0) ? p.x0[(size_t)row * 9 + c] : 0.f;   // every rank keeps its own (identical) copy
        auto publish_x = [&]() {   // three bf16 pieces of the pose row -> A_hi[0,24)
            uint32_t pc[3][8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                __nv_bfloat16 b1[2], b2[2], b3[2];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int c = 2 * j + e;
                    const float v = c < 9 ? x[c < 9 ? c : 0] : 0.f;
                    b1[e] = __float2bfloat16_rn(v);
                    const float r1 = v - __bfloat162float(b1[e]);
                    b2[e] = __float2bfloat16_rn(r1);
                    b3[e] = __float2bfloat16_rn(r1 - __bfloat162float(b2[e]));
                }
                pc[0][j] = pack_bf16(b1[0], b1[1]);
                pc[1][j] = pack_bf16(b2[0], b2[1]);
                pc[2][j] = pack_bf16(b3[0], b3[1]);
            }
            tmem_st8(tm_row + kColX + 0u, pc[0]);
            tmem_st8(tm_row + kColX + 8u, pc[1]);
            tmem_st8(tm_row + kColX + 16u, pc[2]);
            tmem_st_wait();
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_x_ready);
        };
        if (cs == 0) publish_x();

        const float step_size = kOde ? 0.f : p.ts[0] - p.ts[1];
        const float sqrt_step = sqrtf(step_size);
        const float snr_norm = (float)((double)p.snr * 3.0);
        const float snr_R = snr_norm * (float)p.R;   // q = snr |z| / (sum / R)
        uint32_t u = 0;

        // ================= ODE mode: helpers and solver state (dead code in the PC instantiation) =================
        const TcOdeParams &od = tp.ode;
        const int rt = tid;                                                   // 0..255 over the eight row warps
        float *tb_mine = od.tb_cta + (size_t)blockIdx.x * 6 * kCols;         // this CTA's [6][kCols] time biases (its head columns)
        double *sY = reinterpret_cast<double *>(smem + SM::kOffOdeY);             // [9][128]
        double *sYn = sY + 9 * 128;                                           // [9][128]
        // t_bias(t_j) for the gn times in s_times (scorenet.py:63-64, :195; same operation order as compute_time_bias in scorenet.cu),
        // restricted to THIS rank's kCols head columns [n_lo, n_lo + kCols) — the only ones its head epilogue reads.  All eight row
        // warps; scratch = sFpart (free between the team exchange and the next head epilogue).  The weights come from L2: every
        // thread keeps a batch of 32 loads in flight while it works on the previous batch.
        auto time_biases = [&](auto GN) {
            constexpr int gn = decltype(GN)::value;
            float *tf = sFpart, *te = sFpart + 768;
            for (int i = rt; i < gn * 64; i += 256) {
                const int j = i >> 6, k = i & 63;
                const float xp = ((s_times[j] * W[TL::fourier_w + k]) * 2.0f) * 3.14159265358979323846f;
                tf[j * 128 + k] = sinf(xp);
                tf[j * 128 + 64 + k] = cosf(xp);
            }
            named_bar_sync(5, 256);
            {   // te[j][n] = relu(L_t . [sin | cos] + b): column n = rt & 127, times [j0, j0 + per)
                constexpr int per = (gn + 1) / 2;
                const int n = rt & 127, j0 = (rt >> 7) * per;
                const bool act = j0 < gn;
                float acc[per];
                const float b = W[TL::t_b + n];
#pragma unroll
                for (int jj = 0; jj < per; ++jj) acc[jj] = b;
                const float *w = W + TL::t_w + n;
                const float *tfj = tf + (act ? j0 : 0) * 128;
                float wa[32], wb[32];
#pragma unroll
                for (int k = 0; k < 32; ++k) wa[k] = __ldg(w + k * 128);
#pragma unroll
                for (int half = 0; half < 2; ++half) {
#pragma unroll
                    for (int k = 0; k < 32; ++k) wb[k] = __ldg(w + (64 * half + 32 + k) * 128);
#pragma unroll
                    for (int k = 0; k < 32; ++k)
#pragma unroll
                        for (int jj = 0; jj < per; ++jj) acc[jj] = fmaf(tfj[(j0 + jj < gn ? jj : 0) * 128 + 64 * half + k], wa[k], acc[jj]);
                    if (half == 0) {
#pragma unroll
                        for (int k = 0; k < 32; ++k) wa[k] = __ldg(w + (64 + k) * 128);
                    }
#pragma unroll
                    for (int k = 0; k < 32; ++k)
#pragma unroll
                        for (int jj = 0; jj < per; ++jj) acc[jj] = fmaf(tfj[(j0 + jj < gn ? jj : 0) * 128 + 64 * half + 32 + k], wb[k], acc[jj]);
                }
#pragma unroll
                for (int jj = 0; jj < per; ++jj)
                    if (j0 + jj < gn) te[(j0 + jj) * 128 + n] = fmaxf(acc[jj], 0.f);
            }
            named_bar_sync(5, 256);
            {   // tb[j][n_lo + c] = A_t[:, n_lo + c] . te[j]: thread rt owns columns c = rt + 256 i (i < kCpt) of the rank's kCols
                constexpr int kCpt = (kCols + 255) / 256;
                float acc[kCpt][gn];
#pragma unroll
                for (int i = 0; i < kCpt; ++i)
#pragma unroll
                    for (int j = 0; j < gn; ++j) acc[i][j] = 0.f;
                const float *w = W + TL::a_t + n_lo + (rt < kCols ? rt : 0);
                constexpr int kBatch = kCpt == 1 ? 32 : 16;                   // k values per batch of loads in flight
                float wa[kCpt][kBatch], wb[kCpt][kBatch];
                auto load = [&](float (&dst)[kCpt][kBatch], int k0) {
#pragma unroll
                    for (int i = 0; i < kCpt; ++i) {
                        const int off = (256 * i + rt < kCols) ? 256 * i : 0;   // out-of-slice columns re-read a valid one; discarded below
#pragma unroll
                        for (int k = 0; k < kBatch; ++k) dst[i][k] = __ldg(w + off + (size_t)(k0 + k) * 768);
                    }
                };
                auto mac = [&](const float (&src)[kCpt][kBatch], int k0) {
#pragma unroll
                    for (int k4 = 0; k4 < kBatch; k4 += 4)
#pragma unroll
                        for (int j = 0; j < gn; ++j) {
                            const float4 e = *reinterpret_cast<const float4 *>(te + j * 128 + k0 + k4);
#pragma unroll
                            for (int i = 0; i < kCpt; ++i)
                                acc[i][j] = fmaf(e.w, src[i][k4 + 3], fmaf(e.z, src[i][k4 + 2], fmaf(e.y, src[i][k4 + 1], fmaf(e.x, src[i][k4], acc[i][j]))));
                        }
                };
                load(wa, 0);
#pragma unroll 1
                for (int k0 = 0; k0 < 128; k0 += 2 * kBatch) {
                    load(wb, k0 + kBatch);
                    mac(wa, k0);
                    if (k0 + 2 * kBatch < 128) load(wa, k0 + 2 * kBatch);
                    mac(wb, k0 + kBatch);
                }
#pragma unroll
                for (int i = 0; i < kCpt; ++i)
                    if (256 * i + rt < kCols) {
#pragma unroll
                        for (int j = 0; j < gn; ++j) __stcg(tb_mine + j * kCols + 256 * i + rt, acc[i][j]);
                    }
            }
            named_bar_sync(5, 256);
        };
        // (object bias + time bias of evaluation `gi` of the group) -> this rank's column slice of sObt, by `nthr` threads from `t0`
        // Column c0 + nt j of every object: the time biases once, then the object biases of up to four objects as ONE batch of independent
        // loads (a tile of K = 50 candidates spans three or four objects).  (As a plain loop over (object, column) this was one L2 round
        // trip per element and thread — 6 in a team of 4, 24 in a team of 1 — and column half 1 came late to the next layer-0 epilogue.)
        auto fill_obt = [&](int gi, int t0, auto NT) {
            constexpr int nt = decltype(NT)::value, kPerT = ((int)kCols + nt - 1) / nt;
            const int c0 = tid - t0;
            const float *tbn = tb_mine + gi * kCols;
            const float *obn = p.obj_bias + (size_t)obj_lo * 768 + n_lo;
            float tbv[kPerT];
#pragma unroll
            for (int j = 0; j < kPerT; ++j) tbv[j] = c0 + nt * j < (int)kCols ? __ldcg(tbn + c0 + nt * j) : 0.f;
            for (int o0 = 0; o0 < n_obj; o0 += 4) {
                float v[4][kPerT];
#pragma unroll
                for (int k = 0; k < 4; ++k)
#pragma unroll
                    for (int j = 0; j < kPerT; ++j)
                        v[k][j] = (o0 + k < n_obj && c0 + nt * j < (int)kCols) ? __ldg(obn + (size_t)(o0 + k) * 768 + c0 + nt * j) : 0.f;
#pragma unroll
                for (int k = 0; k < 4; ++k)
#pragma unroll
                    for (int j = 0; j < kPerT; ++j)
                        if (o0 + k < n_obj && c0 + nt * j < (int)kCols) sObt[(o0 + k) * kCols + c0 + nt * j] = v[k][j] + tbv[j];
            }
        };
        // solver state, replicated bit-identically in every row thread of every rank (scipy/integrate/_ivp/rk.py, common.py)
        enum { kPhF0 = 0, kPhF1 = 1, kPhAttempt = 2, kPhDenoise = 3 };
        int phase = kPhF0, st = 0, gi = 0, gn = 1, slot = 0, nfev = 0, n_acc = 0, n_rej = 0, status = 0;
        unsigned bar_target = 0;
        bool rejected = false, tb_pending = kOde;                             // time biases of the current group still to be computed
        // Part of the stage combination that does not depend on the evaluation in flight: sum_{j < st} coef_j K_j (float64), built
        // from the stored stages UNDER that evaluation's layer-1 MMAs, so that only one multiply-add per component is left between the
        // score and the next input (the combination used to cost ~4.4 k cycles per evaluation on the critical path)
        double pre[9] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
        const double t_end = 1e-5, direction = -1.0;                          // eps as a Python float; T0 > eps is checked by the host
        double t_cur = (double)od.T0, t_next = 0.0, h = 0.0, h_abs = 0.0, h0 = 0.0, d1 = 0.0, min_step = 0.0;
        const double rtol = (double)od.rtol, atol = (double)od.atol, n_total = (double)p.R * 9.0;
        // stage derivatives K_j of this row: the reference's f is fp32 converted to float64 (samplers.py:198), so fp32 storage is
        // exact; [rank][stage][row][12] floats = three 16-byte words per row and stage, private to this thread, L2-resident
        float *Kmine = reinterpret_cast<float *>(od.Kst) + ((size_t)rank * 7 * p.R + (valid ? row : 0)) * 12;
        const size_t kst = (size_t)p.R * 12;
        // P (below): [4 ranks][R][10] float64 behind the 4 x 7 x R x 12 floats of the stages (the buffer is sized for float64 stages)
        double *Pmine = reinterpret_cast<double *>(reinterpret_cast<float *>(od.Kst) + (size_t)4 * 7 * p.R * 12) + ((size_t)rank * p.R + (valid ? row : 0)) * 10;
        auto store_k = [&](int stage, const double (&k)[9]) {
            float4 *dst = reinterpret_cast<float4 *>(Kmine + (size_t)stage * kst);
            __stcg(dst, make_float4((float)k[0], (float)k[1], (float)k[2], (float)k[3]));
            __stcg(dst + 1, make_float4((float)k[4], (float)k[5], (float)k[6], (float)k[7]));
            __stcg(dst + 2, make_float4((float)k[8], 0.f, 0.f, 0.f));
        };
        // all requested stages are fetched with volatile loads issued back to back (one L2 round trip for the lot; a plain load
        // would be sunk into the select that consumes it and the stages would be fetched one after the other)
        auto load_k = [&](int stage, float (&k)[12]) {
            const float *src = Kmine + (size_t)stage * kst;
#pragma unroll
            for (int w4 = 0; w4 < 3; ++w4)
                asm volatile("ld.global.cg.v4.f32 {%0, %1, %2, %3}, [%4];"
                             : "=f"(k[4 * w4]), "=f"(k[4 * w4 + 1]), "=f"(k[4 * w4 + 2]), "=f"(k[4 * w4 + 3])
                             : "l"(src + 4 * w4)
                             : "memory");
        };
        // RK45 attempt, column half 1: P = sum_{j < stn - 1} coef_j K_j of the NEXT stage stn >= 2 for this row (see barrier 3); those
        // stages were stored before the current evaluation's barriers, K_{stn-1} is what half 0 produces in this evaluation's tail.
        // Teams of 2 and 4 only (in the tail, where half 1 idles); a team of 1 has no exchange, a short tail and a tensor-pipe-bound
        // evaluation: there half 0 accumulates all of `pre` under the layer-1 MMAs (measured: 2.07 vs 2.14 ms at 256 objects).
        auto accumulate_P = [&](int stn) {
            float kf[5][12];
#pragma unroll
            for (int j = 0; j < 5; ++j) load_k(j, kf[j]);
            double P[10];
#pragma unroll
            for (int c = 0; c < 10; ++c) P[c] = 0.0;
#pragma unroll
            for (int j = 0; j < 5; ++j) {
                const double a = stn < 5 ? kRkA[stn + 1][j] : (stn == 5 ? kRkB[j] : kRkE[j]);
#pragma unroll
                for (int c = 0; c < 9; ++c) P[c] += (j < stn - 1 && valid) ? (double)kf[j][c] * a : 0.0;
            }
            if (valid) {                  // (an invalid thread's pointers alias row 0)
                double2 *pd2 = reinterpret_cast<double2 *>(Pmine);
#pragma unroll
                for (int i = 0; i < 5; ++i) __stcg(pd2 + i, make_double2(P[2 * i], P[2 * i + 1]));
            }
        };
        // two grid-wide float64 sums with ONE barrier; every CTA obtains the identical fixed-order totals
        auto grid_sum2 = [&](double a, double b, double &A, double &B) {
            a = warp_sum_f64(valid ? a : 0.0);
            b = warp_sum_f64(valid ? b : 0.0);
            if (lane == 0) {
                s_redd[2 * q] = a;
                s_redd[2 * q + 1] = b;
            }
            named_bar_sync(2, 128);
            bar_target += (unsigned)n_tiles;
            if (warp == 0) {
                if (lane == 0) {
                    if (leader) {
                        double *dst = od.partial + ((size_t)slot * n_tiles + tile) * 2;
                        __stcg(dst, (s_redd[0] + s_redd[2]) + (s_redd[4] + s_redd[6]));
                        __stcg(dst + 1, (s_redd[1] + s_redd[3]) + (s_redd[5] + s_redd[7]));
                        red_release_add_u32(p.barrier, 1u);
                    }
                    while (ld_acquire_u32(p.barrier) < bar_target) {
                    }
                }
                __syncwarp();
                // lanes fetch the per-tile partials in parallel (n_tiles < 64); fixed-shape shuffle tree => identical in every CTA
                double ta = 0.0, tb = 0.0;
                for (int i = lane; i < n_tiles; i += 32) {
                    ta += __ldcg(od.partial + ((size_t)slot * n_tiles + i) * 2);
                    tb += __ldcg(od.partial + ((size_t)slot * n_tiles + i) * 2 + 1);
                }
                ta = warp_sum_f64(ta);
                tb = warp_sum_f64(tb);
                if (lane == 0) {
                    s_redd[8] = ta;
                    s_redd[9] = tb;
                }
            }
            named_bar_sync(2, 128);
            A = s_redd[8];
            B = s_redd[9];
            slot = (slot + 1) & 3;
        };
        int te_next = 0;                                                      // trajectory output: first t_eval index not yet written
        if constexpr (kOde) {
            if (cs == 0) {
#pragma unroll
                for (int c = 0; c < 9; ++c) sY[c * 128 + r] = (double)x[c];   // y0 = float64(init_x) (samplers.py:205)
                if (od.proc.out && od.proc.t_eval == nullptr && leader && valid && od.proc.cap > 0) {   // state 0 = the start
                    double y0[9];
#pragma unroll
                    for (int c = 0; c < 9; ++c) y0[c] = (double)x[c];
                    ode_write_state(od.proc.out + (size_t)row * 9, y0, p.pts_center + (size_t)(row / p.K) * 3);
                }
            }
            // the time biases of the first group are computed inside the first evaluation (tb_pending, below)
        }

        // relu(acc + obj_bias + t_bias) . O over one 32-column block whose first stacked hidden unit is n (one head per block)
        // Packed fp32 pairs (FADD2 / FFMA2): even and odd columns accumulate in the two halves of one 64-bit register per output
        // component and are added at the end — 12 instead of 20 issue slots per four columns.
        auto head_block = [&](const uint32_t (&v)[32], int n, float &o0, float &o1, float &o2) {
            const float *ob = obt_row + n;
            const float *w0 = sOw + (size_t)(3 * (n >> 8)) * 256 + (n & 255);
            f32x2 a0 = pack_f32x2(o0, 0.f), a1 = pack_f32x2(o1, 0.f), a2 = pack_f32x2(o2, 0.f);
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
                const float4 o4 = *reinterpret_cast<const float4 *>(ob + 4 * j4);
                const float4 wa = *reinterpret_cast<const float4 *>(w0 + 4 * j4);
                const float4 wb = *reinterpret_cast<const float4 *>(w0 + 256 + 4 * j4);
                const float4 wc = *reinterpret_cast<const float4 *>(w0 + 512 + 4 * j4);
                float h0, h1, h2, h3;
                unpack_f32x2(add_f32x2(pack_f32x2(__uint_as_float(v[4 * j4 + 0]), __uint_as_float(v[4 * j4 + 1])), pack_f32x2(o4.x, o4.y)), h0, h1);
                unpack_f32x2(add_f32x2(pack_f32x2(__uint_as_float(v[4 * j4 + 2]), __uint_as_float(v[4 * j4 + 3])), pack_f32x2(o4.z, o4.w)), h2, h3);
                const f32x2 p01 = pack_f32x2(fmaxf(h0, 0.f), fmaxf(h1, 0.f)), p23 = pack_f32x2(fmaxf(h2, 0.f), fmaxf(h3, 0.f));
                a0 = fma_f32x2(p23, pack_f32x2(wa.z, wa.w), fma_f32x2(p01, pack_f32x2(wa.x, wa.y), a0));
                a1 = fma_f32x2(p23, pack_f32x2(wb.z, wb.w), fma_f32x2(p01, pack_f32x2(wb.x, wb.y), a1));
                a2 = fma_f32x2(p23, pack_f32x2(wc.z, wc.w), fma_f32x2(p01, pack_f32x2(wc.x, wc.y), a2));
            }
            float e, od;
            unpack_f32x2(a0, e, od); o0 = e + od;
            unpack_f32x2(a1, e, od); o1 = e + od;
            unpack_f32x2(a2, e, od); o2 = e + od;
        };

        for (int step = 0; kOde || step < p.T; ++step) {
            unsigned long long *ds = (dbg && step < p.T) ? p.dbg + (size_t)step * 16 : nullptr;
            if (ds) ds[0] = clock64();
#ifdef GPB_DBG_TRACE
            // experiment builds only: eight extra stamps per step of row thread 0 (slots 0..3) and of thread 128, column half 1 (4..7)
            unsigned long long *rtr = (dbg_cta && (tid == 0 || tid == 128) && step < p.T) ? p.dbg + (size_t)2 * p.T * 16 + (size_t)step * 8 + (tid >> 5) : nullptr;
#define RTR(i) do { if (rtr) rtr[i] = clock64(); } while (0)
#else
#define RTR(i) do { } while (0)
#endif
            RTR(0);
            if constexpr (SM::kNoiseByHalf1) {
                // The two noise draws of this step (18 normals per row: six Philox blocks + Box-Muller, or 18 loads) are produced by
                // COLUMN HALF 1, whose warps reach this point during the previous step's tail (exchange, norm, grid word, update
                // belong to half 0) and would otherwise idle; half 0 picks them up after barrier 3.  They used to be generated by
                // half 0 between the layer-1 and the head epilogues, i.e. in front of barrier 3 on the step's critical path.
                if (cs == 1 && valid) {
                    float za[9], zb[9];
                    row_noise(p, step, 0, row, za);
                    row_noise(p, step, 1, row, zb);
#pragma unroll
                    for (int c = 0; c < 9; ++c) {
                        sNoise[c * 128 + r] = za[c];
                        sNoise[(9 + c) * 128 + r] = zb[c];
                    }
                }
            }
            const float t = kOde ? s_times[gi] : p.ts[step];
            const float sigma = sigma_of_t(t);
            const float stdv = sigma + 1e-7f;
            float ode_coef = 0.f;                       // fp32(0.5 g^2), g = float64(sigma_fp32) * sqrt(2 ln 5000): off the critical path here
            if constexpr (kOde) {
                const double gd = (double)sigma * 4.12727348049926;
                ode_coef = (float)(0.5 * gd * gd);
            }
            // ---- layers 0 and 1: accumulator -> bias + ReLU -> bf16 hi/lo -> A operand in tensor memory ----
#pragma unroll 1
            for (int layer = 0; layer < 2; ++layer) {
                const float *bias = sBias + layer * 256 + cs * 64;
                if (layer == 0) {
                    // Layer 0 -> layer 1 in QUARTERS.  Nothing but layer 0's own MMAs reads the A region now (the previous step's
                    // head MMAs completed before x was published), and those need only x at the END of A_lo (kColX), which quarters
                    // 2 and 3 overwrite: quarters 0 and 1 are stored as soon as the FIRST accumulator is converted, the other two
                    // once the second one is complete.  The MMA warp starts layer 1 on the first quarter;
                    // quarter g = K-steps {base, base + 1, base + 4, base + 5}, base = 8 (g / 2) + 2 (g % 2).
                    const uint32_t b0 = u & 1u, n0 = u >> 1, b1 = (u + 1u) & 1u, n1 = (u + 1u) >> 1;
                    uint32_t v[32], h16[16], l16[16];
                    mbar_wait(&bar_acc_full[b0], n0 & 1u);
                    if (ds) ds[1] = clock64();
                    tc_fence_after_sync();
                    tmem_ld32(tm_row + kColD + b0 * 128u + (uint32_t)cs * 64u, v);
                    tmem_ld_wait();
                    relu_split32<kW16>(v, bias, h16, l16);
#pragma unroll
                    for (int qd = 0; qd < 4; ++qd) {
                        const uint32_t bq = qd < 2 ? b0 : b1;
                        const uint32_t a_col = (uint32_t)(qd >> 1) * 64u + (uint32_t)cs * 32u + (uint32_t)(qd & 1) * 16u;
                        tmem_st16(tm_row + kColAhi + a_col, h16);
                        tmem_st16(tm_row + kColAlo + a_col, l16);
                        if (qd < 3) {   // the next 32 accumulator columns travel while the stores drain
                            if (qd == 1) {   // second accumulator: complete = every layer-0 MMA has read x, which quarters 2 and 3 overwrite
                                mbar_wait(&bar_acc_full[b1], n1 & 1u);
                                tc_fence_after_sync();
                            }
                            const uint32_t bn = qd + 1 < 2 ? b0 : b1;
                            tmem_ld32(tm_row + kColD + bn * 128u + (uint32_t)cs * 64u + (uint32_t)((qd + 1) & 1) * 32u, v);
                            tmem_ld_wait();
                        }
                        tmem_st_wait();
                        tc_fence_before_sync();
                        __syncwarp();
                        // accumulator hand-back: every operand announcement below (h1 quarters, pf halves, x) is made AFTER this warp's
                        // last read of the accumulator the announced MMAs will overwrite, so the MMA warp waits for bar_acc_empty only
                        // where no operand hand-off stands in between (head units >= 2); the arrivals keep the phase count uniform
                        if (lane == 0) {
                            mbar_arrive(&bar_h1_ready[qd]);
                            if (qd == 0) mbar_arrive(&bar_acc_empty[b0]);      // (its second 32 columns are in registers by now)
                            if (qd == 2) mbar_arrive(&bar_acc_empty[b1]);
                        }
                        if (qd < 3) relu_split32<kW16>(v, bias + (qd + 1 < 2 ? 0 : 128) + ((qd + 1) & 1) * 32, h16, l16);
                    }
                    u += 2;
                    if (ds) ds[2] = clock64();
                } else {
                uint32_t hi[32], lo[32];
                {   // unit a: columns [0,128) of the layer; this thread: [cs*64, +64)
                    const uint32_t b = u & 1u, n = u >> 1;
                    mbar_wait(&bar_acc_full[b], n & 1u);
                    RTR(3);
                    if (ds) ds[1 + 2 * layer] = clock64();
                    tc_fence_after_sync();
                    {
                        uint32_t v[32];
                        tmem_ld32(tm_row + kColD + b * 128u + (uint32_t)cs * 64u, v);
                        tmem_ld_wait();
                        relu_split32<kW16>(v, bias, hi, lo);
                        tmem_ld32(tm_row + kColD + b * 128u + (uint32_t)cs * 64u + 32u, v);
                        tmem_ld_wait();
                        tc_fence_before_sync();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&bar_acc_empty[b]);
                        relu_split32<kW16>(v, bias + 32, hi + 16, lo + 16);
                    }
                    ++u;
                }
                {   // unit b: once its accumulator is complete every MMA of the layer has consumed the old A
                    const uint32_t b = u & 1u, n = u >> 1;
                    mbar_wait(&bar_acc_full[b], n & 1u);
                    tc_fence_after_sync();
                    tmem_st32(tm_row + kColAhi + (uint32_t)cs * 32u, hi);
                    tmem_st32(tm_row + kColAlo + (uint32_t)cs * 32u, lo);
                    // first half of the new A operand (K columns [0,128)) is in tensor memory: let the MMA warp start on it at once
                    // (the heads' first eight K-steps then cover the conversion of the second half)
                    tmem_st_wait();
                    tc_fence_before_sync();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bar_a_ready[0]);
                    {
                        uint32_t v[32];
                        tmem_ld32(tm_row + kColD + b * 128u + (uint32_t)cs * 64u, v);
                        tmem_ld_wait();
                        relu_split32<kW16>(v, bias + 128, hi, lo);
                        tmem_ld32(tm_row + kColD + b * 128u + (uint32_t)cs * 64u + 32u, v);
                        tmem_ld_wait();
                        tc_fence_before_sync();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&bar_acc_empty[b]);
                        relu_split32<kW16>(v, bias + 128 + 32, hi + 16, lo + 16);
                    }
                    tmem_st32(tm_row + kColAhi + 64u + (uint32_t)cs * 32u, hi);
                    tmem_st32(tm_row + kColAlo + 64u + (uint32_t)cs * 32u, lo);
                    tmem_st_wait();
                    tc_fence_before_sync();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bar_a_ready[1]);
                    ++u;
                }
                if (ds) ds[2 + 2 * layer] = clock64();
                }
                if constexpr (kOde) {
                    // First evaluation of a group: its time biases are only needed by the head epilogue, so they are computed
                    // here, after h1 has been handed to the MMA warp — under the layer-1 MMAs instead of in front of layer 0.
                    if (layer == 0 && tb_pending) {
                        if (gn == 6) time_biases(std::integral_constant<int, 6>{});
                        else time_biases(std::integral_constant<int, 1>{});
                        fill_obt(0, 0, std::integral_constant<int, 256>{});
                        tb_pending = false;
                    }
                    if constexpr (kTeam == 1) {
                        if (layer == 0 && cs == 0 && phase == kPhAttempt) {
                            // a team of 1: all of pre = sum_{j < st} coef_j K_j here, under the layer-1 MMAs (see accumulate_P / barrier 3)
                            float kf[6][12];
#pragma unroll
                            for (int j = 0; j < 6; ++j) load_k(j, kf[j]);
                            const int nj = st < 5 ? st : (st == 5 ? 5 : 6);
#pragma unroll
                            for (int c = 0; c < 9; ++c) pre[c] = 0.0;
#pragma unroll
                            for (int j = 0; j < 6; ++j) {
                                const double a = st < 5 ? kRkA[st + 1][j < 5 ? j : 0] : (st == 5 ? kRkB[j] : kRkE[j]);
#pragma unroll
                                for (int c = 0; c < 9; ++c) pre[c] += (j < nj && valid) ? (double)kf[j][c] * a : 0.0;
                            }
                        }
                    }
                }
            }
            float z1[9], z2[9];                       // noise of this step (PC): written by column half 1 at the top of the step,
            if constexpr (!kOde && !SM::kNoiseByHalf1) {   // or (team of 1) generated here, under the first head unit's MMAs
                if (cs == 0 && valid) {
                    row_noise(p, step, 0, row, z1);
                    row_noise(p, step, 1, row, z2);
                }
            }
            if constexpr (!kOde) {
                // warps 4-7: (object bias + time bias) table of THIS step, while the tensor core runs the head slice (everyone left
                // the previous step's table at its barrier 1).  Not in the previous step's tail: a warp that is still here when
                // x is published would join the layer-0 epilogue late and stall the MMA warp behind it.  Column c0 + 128 j of every
                // object: the time biases once, then one batch of six independent loads per object.
                if (cs == 1 && step > 0) {
                    const float *tbn = p.tb_table + (size_t)step * 768 + n_lo;
                    const float *obn = p.obj_bias + (size_t)obj_lo * 768 + n_lo;
                    constexpr int kPer = (kCols + 127) / 128;                  // columns per thread: c0 + 128 j
                    const int c0 = tid - 128;
                    float tbv[kPer];
#pragma unroll
                    for (int j = 0; j < kPer; ++j) tbv[j] = c0 + 128 * j < kCols ? __ldg(tbn + c0 + 128 * j) : 0.f;
                    for (int o = 0; o < n_obj; ++o) {
                        float v[kPer];
#pragma unroll
                        for (int j = 0; j < kPer; ++j) v[j] = c0 + 128 * j < kCols ? __ldg(obn + (size_t)o * 768 + c0 + 128 * j) : 0.f;
#pragma unroll
                        for (int j = 0; j < kPer; ++j)
                            if (c0 + 128 * j < kCols) sObt[o * kCols + c0 + 128 * j] = v[j] + tbv[j];
                    }
                }
            }
            named_bar_sync(3, kTcRowWarps * 32);      // sObt holds obj_bias + t_bias of THIS step (written by warps 4-7)
            if constexpr (kOde && kTeam > 1) {
                if (cs == 0 && phase == kPhAttempt) {
                    // Stage `st` is being evaluated.  The part of the next stage combination that does not depend on this evaluation is
                    //     pre = sum_{j < st} coef_j K_j,   coef = a[st+1] (st < 5), b (st == 5 -> y_new), E (st == 6 -> error),
                    // summed j ascending like the in-line form.  Column half 1 accumulated j < st - 1 during the previous evaluation's
                    // tail, where it idles (P, below); what is left here is ONE multiply-add per component with the newest stage, issued
                    // behind barrier 3 and complete long before the first head accumulator.  (All of it used to run here in half 0,
                    // between the layer-0 and layer-1 epilogues: ~4 k cycles — an out-of-line block of ~400 instructions fetched
                    // from L2 at every evaluation — of which ~2.9 k delayed pf and with it the head MMAs.)
                    float kl[12];
                    load_k(st - 1, kl);
                    double P[10];
                    if (st >= 2) {
                        const double2 *ps = reinterpret_cast<const double2 *>(Pmine);
#pragma unroll
                        for (int i = 0; i < 5; ++i) {
                            const double2 v = __ldcg(ps + i);
                            P[2 * i] = v.x;
                            P[2 * i + 1] = v.y;
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < 10; ++i) P[i] = 0.0;
                    }
                    const double a = st < 5 ? kRkA[st + 1][st - 1] : (st == 5 ? kRkB[4] : kRkE[5]);
#pragma unroll
                    for (int c = 0; c < 9; ++c) pre[c] = valid ? P[c] + (double)kl[c] * a : 0.0;
                }
            }
            if (SM::kNoiseByHalf1 && cs == 0 && valid) {   // (before barrier 1: half 1 overwrites the buffer at the top of the next step)
#pragma unroll
                for (int c = 0; c < 9; ++c) {
                    z1[c] = sNoise[c * 128 + r];
                    z2[c] = sNoise[(9 + c) * 128 + r];
                }
            }
            // ---- head slice: relu(acc + obj_bias + t_bias) . O, partial sums per touched head: o[h - hA] (teams 2 and 4 touch two
            //      heads, team 1 all three) ----
            float o[3][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};
            auto head_cols = [&](const uint32_t (&v)[32], int nn) {          // 32 columns starting at stacked unit nn (one head per block)
                const int sel = (nn >> 8) - hA;                              // team 1: a compile-time constant after unrolling
                if (sel == 0) head_block(v, nn, o[0][0], o[0][1], o[0][2]);
                else if (kTeam > 1 || sel == 1) head_block(v, nn, o[1][0], o[1][1], o[1][2]);
                else head_block(v, nn, o[2][0], o[2][1], o[2][2]);
            };
#pragma unroll
            for (int hu = 0; hu < TS::kHeadUnits; ++hu) {
                const uint32_t b = u & 1u, n = u >> 1;
                mbar_wait(&bar_acc_full[b], n & 1u);
                if (ds && hu == 0) ds[5] = clock64();
                if (ds && hu == TS::kHeadUnits - 1) ds[7] = clock64();
                tc_fence_after_sync();
                if (!(TS::kSmallUnit && hu == TS::kHeadUnits - 1)) {
                    // 128-column unit: this thread's columns [cs*64, +64) of the unit
#pragma unroll 1
                    for (int blk = 0; blk < 2; ++blk) {
                        uint32_t v[32];
                        tmem_ld32(tm_row + kColD + b * 128u + (uint32_t)(cs * 64 + blk * 32), v);
                        tmem_ld_wait();
                        head_cols(v, n_lo + hu * 128 + cs * 64 + blk * 32);
                    }
                    tc_fence_before_sync();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bar_acc_empty[b]);
                } else {
                    // 64-column unit (team 4): this thread's columns [cs*32, +32) of the unit
                    uint32_t v[32];
                    tmem_ld32(tm_row + kColD + b * 128u + (uint32_t)(cs * 32), v);
                    tmem_ld_wait();
                    tc_fence_before_sync();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bar_acc_empty[b]);
                    head_cols(v, n_lo + hu * 128 + cs * 32);
                }
                ++u;
                if (ds && hu == 0) ds[6] = clock64();
                if (ds && hu == TS::kHeadUnits - 1) ds[8] = clock64();
            }
            // column sub-half 1 hands its partial sums to sub-half 0
            constexpr int kTouched = kTeam == 1 ? 3 : 2;
            if (cs == 1) {
#pragma unroll
                for (int hh = 0; hh < kTouched; ++hh)
#pragma unroll
                    for (int c = 0; c < 3; ++c) sFpart[r * 12 + 3 * hh + c] = o[hh][c];
            }
            named_bar_sync(1, kTcRowWarps * 32);      // sub-half 1 partials are in sFpart; everyone is done with sObt
            if (ds) ds[9] = clock64();
            if (cs == 1) {
                // warps 4-7: table of (object bias + time bias) for the NEXT step, while warps 0-3 exchange / reduce / update
                if constexpr (!kOde) {
                    continue;
                } else {
                    if (gi + 1 < gn) {                    // inside an evaluation group: the next time bias is already in tb_mine
                        ++gi;
                        fill_obt(gi, 128, std::integral_constant<int, 128>{});
                        if (kTeam > 1 && gn == 6) accumulate_P(gi + 1);
                        continue;
                    }
                    named_bar_sync(4, 256);               // group boundary: warps 0-3 have decided what comes next
                    gn = s_gn;
                    gi = 0;
                    if (gn == 0) break;                   // solved
                    tb_pending = true;
                    continue;
                }
            }
#pragma unroll
            for (int hh = 0; hh < kTouched; ++hh)
#pragma unroll
                for (int c = 0; c < 3; ++c) o[hh][c] += sFpart[r * 12 + 3 * hh + c];
            // step constants of the update, computed before the team's partials arrive (they depend on t only)
            const float inv_std = 1.0f / stdv;   // one IEEE division per row and step instead of nine (<= 1.5 ulp from f / std, scorenet.py:217)
            const float g_sde = sigma * kGCoef, g2_sde = g_sde * g_sde, g_sqrt_step = g_sde * sqrt_step;   // ve_sde diffusion (sde.py:20-24)
            const float nsum_scale = kNormSumScale * fminf(sigma, 1.0f);   // fixed-point scale of the reduction word: norms grow like 1 / sigma(t)
            float f[9];
            if constexpr (kTeam == 1) {
#pragma unroll
                for (int c = 0; c < 9; ++c) f[c] = o[c / 3][c % 3];
                if (ds) ds[11] = ds[10] = clock64();
            } else {
                // ---- all-to-all inside the tile team through distributed shared memory: 6 floats per row and rank (its two heads),
                //      summed by every rank in the same fixed rank order ----
                const uint32_t par = (uint32_t)step & 1u;
                float *my_slot = sMail + (((size_t)par * kTeam + rank) * 128 + r) * 8;
                const uint32_t mail_off = smem_u32(my_slot);
                const uint32_t bar_off = smem_u32(&bar_mail[par]);
                if (tid == 0) mbar_arrive_expect_tx(&bar_mail[par], (uint32_t)(kTeam - 1) * 128u * 24u);   // peers x 128 rows x 24 B
                *reinterpret_cast<float4 *>(my_slot) = make_float4(o[0][0], o[0][1], o[0][2], o[1][0]);   // own slot: read back by this same thread
                *reinterpret_cast<float2 *>(my_slot + 4) = make_float2(o[1][1], o[1][2]);
#pragma unroll
                for (uint32_t d = 1; d < (uint32_t)kTeam; ++d) {
                    const uint32_t dst = ((uint32_t)rank + d) & (uint32_t)(kTeam - 1);
                    const uint32_t ra = mapa_shared(mail_off, dst), rb = mapa_shared(bar_off, dst);
                    st_async_f4(ra, o[0][0], o[0][1], o[0][2], o[1][0], rb);
                    st_async_f2(ra + 16u, o[1][1], o[1][2], rb);
                }
                if (ds) ds[11] = clock64();
                mbar_wait_cluster(&bar_mail[par], ((uint32_t)step >> 1) & 1u);
                if (ds) ds[10] = clock64();
                const float *mb = sMail + ((size_t)(par * kTeam) * 128 + r) * 8;
#pragma unroll
                for (int c = 0; c < 9; ++c) f[c] = 0.f;
#pragma unroll
                for (int pr = 0; pr < kTeam; ++pr) {     // rank pr covers stacked units [kCols pr, kCols pr + kCols): heads ha and (if it crosses a boundary) ha + 1
                    const float4 a = *reinterpret_cast<const float4 *>(mb + (size_t)pr * 128 * 8);
                    const float2 b = *reinterpret_cast<const float2 *>(mb + (size_t)pr * 128 * 8 + 4);
                    const int ha = (kCols * pr) / 256, hb = (kCols * pr + kCols - 1) / 256;
                    f[3 * ha + 0] += a.x; f[3 * ha + 1] += a.y; f[3 * ha + 2] += a.z;
                    if (hb != ha) { f[3 * hb + 0] += a.w; f[3 * hb + 1] += b.x; f[3 * hb + 2] += b.y; }
                }
            }
            if constexpr (kOde) {
                // ======================= RK45 with SciPy's controller (samplers.py:178-227, scipy/integrate/_ivp) =======================
                // k = f(t, x) as ode_func returns it (samplers.py:189-198): score in fp32, f = 0 - fp32(0.5 g^2) * score in fp32
                // (NumPy-1.23 value-based casting), g = float64(sigma_fp32) * sqrt(2 ln 5000)
                double kc[9];
#pragma unroll
                for (int c = 0; c < 9; ++c) kc[c] = valid ? (double)(0.0f - ode_coef * ((f[c] + sOw[9 * 256 + c]) / stdv)) : 0.0;
                ++nfev;
                long long dbg_a = 0, dbg_b = 0;                   // profiling: after f -> k, after the stage combination
                if (ds) dbg_a = -(long long)clock64();
                bool boundary = false, finished = false;
                double xn[9];                                     // input of the next evaluation (float64, rounded to fp32 on publication)
                // start one attempt of a step from (t_cur, y, K0): RungeKutta._step_impl, rk.py; returns false when the step size underflows
                auto begin_attempt = [&](const double *k0_known) -> bool {   // k0_known: K_0 of the step when it is still in registers (FSAL)
                    if (h_abs < min_step) {
                        status = -1;
                        return false;
                    }
                    h = h_abs * direction;
                    t_next = t_cur + h;
                    if (direction * (t_next - t_end) > 0) t_next = t_end;
                    h = t_next - t_cur;
                    h_abs = fabs(h);
                    double k0v[9];
                    if (k0_known) {
#pragma unroll
                        for (int c = 0; c < 9; ++c) k0v[c] = k0_known[c];
                    } else {
                        float k0[12];
                        load_k(0, k0);
#pragma unroll
                        for (int c = 0; c < 9; ++c) k0v[c] = valid ? (double)k0[c] : 0.0;
                    }
#pragma unroll
                    for (int c = 0; c < 9; ++c) xn[c] = sY[c * 128 + r] + (k0v[c] * kRkA[1][0]) * h;
                    if (tid == 0) {
                        for (int j = 1; j < 6; ++j) s_times[j - 1] = (float)(t_cur + kRkC[j] * h);
                        s_times[5] = (float)(t_cur + h);
                        s_gn = 6;
                    }
                    phase = kPhAttempt;
                    st = 1;
                    return true;
                };
                auto begin_step = [&]() {                        // the outer `while` of solve_ivp / RK45.step
                    min_step = 10.0 * fabs(nextafter(t_cur, direction * INFINITY) - t_cur);
                    if (h_abs < min_step) h_abs = min_step;
                    rejected = false;
                };
                auto begin_denoise = [&]() {                     // samplers.py:209-218: one more score evaluation at t = eps
#pragma unroll
                    for (int c = 0; c < 9; ++c) xn[c] = sY[c * 128 + r];
                    if (tid == 0) {
                        s_times[0] = kSamplingEps;
                        s_gn = 1;
                        st_volatile_shared(&s_allowed, step + 2);     // this evaluation and the denoise one (already so after an attempt)
                        st_volatile_shared(&s_final, 1);
                    }
                    phase = kPhDenoise;
                };
                if (phase == kPhF0) {
                    // ---- select_initial_step (scipy/integrate/_ivp/common.py), first half
                    double a0 = 0.0, a1 = 0.0;
#pragma unroll
                    for (int c = 0; c < 9; ++c) {
                        const double yv = sY[c * 128 + r], sc = atol + fabs(yv) * rtol;
                        const double uu = yv / sc, vv = kc[c] / sc;
                        a0 += uu * uu;
                        a1 += vv * vv;
                    }
                    if (valid) store_k(0, kc);                                        // K0 = f0
                    double A0, A1;
                    grid_sum2(a0, a1, A0, A1);
                    const double d0 = sqrt(A0 / n_total);
                    d1 = sqrt(A1 / n_total);
                    h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
                    h0 = fmin(h0, fabs(t_end - t_cur));
#pragma unroll
                    for (int c = 0; c < 9; ++c) xn[c] = sY[c * 128 + r] + h0 * direction * kc[c];
                    if (tid == 0) {
                        s_times[0] = (float)(t_cur + h0 * direction);
                        s_gn = 1;
                    }
                    phase = kPhF1;
                    boundary = true;
                } else if (phase == kPhF1) {
                    // ---- select_initial_step, second half: d2 from f1 - f0, then the first attempt
                    double a2 = 0.0;
                    float k0f[12];
                    load_k(0, k0f);
#pragma unroll
                    for (int c = 0; c < 9; ++c) {
                        const double yv = sY[c * 128 + r], sc = atol + fabs(yv) * rtol;
                        const double k0 = valid ? (double)k0f[c] : 0.0;
                        const double ww = (kc[c] - k0) / sc;
                        a2 += ww * ww;
                    }
                    double A2, unused;
                    grid_sum2(a2, 0.0, A2, unused);
                    const double d2 = sqrt(A2 / n_total) / h0;
                    double h1;
                    if (d1 <= 1e-15 && d2 <= 1e-15) h1 = fmax(1e-6, h0 * 1e-3);
                    else h1 = pow(0.01 / fmax(d1, d2), 1.0 / 5.0);
                    h_abs = fmin(fmin(100.0 * h0, h1), fabs(t_end - t_cur));
                    begin_step();
                    if (!begin_attempt(nullptr)) begin_denoise();
                    boundary = true;
                } else if (phase == kPhAttempt) {
                    if (valid && st < 6) store_k(st, kc);                           // K[st] (K[6] = f_new stays in registers)
                    if (st < 6) {
                        // next stage input: y + h * sum_j a[st+1][j] K_j for st < 5, the 5th-order solution y_new (b weights) for st == 5
                        // (the earlier stages come back from L2; every load is issued unconditionally so that all of them are in
                        // flight together, and a stage that does not take part is discarded by the select, not by a branch)
                        // (the earlier stages' part of the sum was pre-accumulated under this evaluation's MMAs: `pre`)
                        double dy[9];
#pragma unroll
                        for (int c = 0; c < 9; ++c) dy[c] = pre[c];
                        {
                            const double a = st < 5 ? kRkA[st + 1][st] : kRkB[st];
#pragma unroll
                            for (int c = 0; c < 9; ++c) {
                                dy[c] += kc[c] * a;
                                xn[c] = sY[c * 128 + r] + dy[c] * h;
                                if (st == 5) sYn[c * 128 + r] = xn[c];
                            }
                        }
                        if (ds) dbg_b = -(long long)clock64();
                        ++st;
                    } else {
                        // ---- error estimate over the WHOLE batch (one controller, samplers.py:205 / rk.py _estimate_error_norm)
                        double ae = 0.0;
                        double errv[9];
#pragma unroll
                        for (int c = 0; c < 9; ++c) errv[c] = pre[c];                     // sum_{j < 6} E_j K_j, pre-accumulated
#pragma unroll
                        for (int c = 0; c < 9; ++c) {
                            const double err = errv[c] + kc[c] * kRkE[6];
                            const double sc = atol + fmax(fabs(sY[c * 128 + r]), fabs(sYn[c * 128 + r])) * rtol;
                            const double ww = err * h / sc;
                            ae += ww * ww;
                        }
                        double AE, unused;
                        grid_sum2(ae, 0.0, AE, unused);
                        const double error_norm = sqrt(AE / n_total);
                        const double SAFETY = 0.9, MIN_FACTOR = 0.2, MAX_FACTOR = 10.0, ERR_EXP = -1.0 / 5.0;
                        bool go_on, k0_in_regs = false;
                        if (error_norm < 1.0) {
                            double factor = error_norm == 0.0 ? MAX_FACTOR : fmin(MAX_FACTOR, SAFETY * pow(error_norm, ERR_EXP));
                            if (rejected) factor = fmin(1.0, factor);
                            h_abs *= factor;
                            ++n_acc;
                            if (od.proc.out) {   // trajectory output (rare: --save_video / return_process); K_0..K_5 from L2, K_6 = f_new
                                double yo[9], yn[9], k6[9];
#pragma unroll
                                for (int c = 0; c < 9; ++c) {
                                    yo[c] = sY[c * 128 + r];
                                    yn[c] = sYn[c * 128 + r];
                                    k6[c] = kc[c];
                                }
                                const float *kbase = Kmine;
                                const size_t kstride = kst;
                                ode_emit_step(od.proc, p.R, valid ? row : 0, p.pts_center + (size_t)((valid ? row : 0) / p.K) * 3, leader && valid,
                                              n_acc, te_next, t_cur, t_next, h, yo, yn, [&](int j, int c) -> double {
                                                  return j < 6 ? (double)__ldcg(kbase + (size_t)j * kstride + c) : k6[c];
                                              });
                            }
                            // accept: y <- y_new, K0 <- K6 = f_new (FSAL), t <- t + h clipped to the bound
#pragma unroll
                            for (int c = 0; c < 9; ++c) sY[c * 128 + r] = sYn[c * 128 + r];
                            if (valid) store_k(0, kc);
                            k0_in_regs = true;
                            t_cur = t_next;
                            go_on = direction * (t_cur - t_end) < 0;
                            if (go_on) begin_step();
                        } else {
                            h_abs *= fmax(MIN_FACTOR, SAFETY * pow(error_norm, ERR_EXP));
                            rejected = true;
                            ++n_rej;
                            go_on = true;
                        }
                        if (go_on && begin_attempt(k0_in_regs ? kc : nullptr)) {
                            if (tid == 0) st_volatile_shared(&s_allowed, ld_volatile_shared(&s_allowed) + 6);
                        } else {
                            begin_denoise();
                        }
                        boundary = true;
                    }
                } else {
                    // ---- denoise (samplers.py:209-218), normalize_rotation in float64 (:225), + pts_center (:226)
                    double v[9];
                    {
                        const float g = sigma * kGCoef, g2 = g * g;
                        const double dt = od.denoise_steps > 0 ? (1.0 - 1e-5) / (double)od.denoise_steps : 0.0;
#pragma unroll
                        for (int c = 0; c < 9; ++c) {
                            const float drift = 0.0f - g2 * ((f[c] + sOw[9 * 256 + c]) / stdv);
                            v[c] = sY[c * 128 + r] + (double)drift * dt;
                        }
                    }
                    if (leader && valid) {
                        const double n1 = fmax(sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]), 1e-12);
                        const double b0 = v[0] / n1, b1 = v[1] / n1, b2 = v[2] / n1;
                        const double d = b0 * v[3] + b1 * v[4] + b2 * v[5];
                        const double c0 = v[3] - d * b0, c1 = v[4] - d * b1, c2 = v[5] - d * b2;
                        const double n2 = fmax(sqrt(c0 * c0 + c1 * c1 + c2 * c2), 1e-12);
                        double *o = od.pose + (size_t)row * 9;
                        o[0] = b0; o[1] = b1; o[2] = b2;
                        o[3] = c0 / n2; o[4] = c1 / n2; o[5] = c2 / n2;
                        const float *ctr = p.pts_center + (size_t)(row / p.K) * 3;
#pragma unroll
                        for (int c = 6; c < 9; ++c) o[c] = v[c] + (double)ctr[c - 6];
                    }
                    if (od.denoise_steps <= 0) --nfev;            // the reference skips this evaluation; we run it and drop its result
                    if (od.stats && blockIdx.x == 0 && tid == 0) {
                        od.stats[0] = nfev;
                        od.stats[1] = n_acc;
                        od.stats[2] = n_rej;
                        od.stats[3] = status;
                    }
                    if (tid == 0) s_gn = 0;
                    boundary = true;
                    finished = true;
                }
                if (ds) ds[12] = clock64();
                if (!finished) {
#pragma unroll
                    for (int c = 0; c < 9; ++c) x[c] = (float)xn[c];
                    publish_x();                                  // the MMA warp starts on layer 0 while the time biases are refreshed
                }
                if (ds) {
                    ds[13] = clock64();
                    ds[14] = (unsigned long long)dbg_a;           // negative inside a group; overwritten with positive stamps at a group end
                    ds[15] = (unsigned long long)dbg_b;
                }
                if (!boundary) {
                    ++gi;
                    continue;
                }
                named_bar_sync(4, 256);
                if (ds) ds[14] = clock64();
                gn = s_gn;
                gi = 0;
                if (gn == 0) break;
                tb_pending = true;
                if (ds) ds[15] = clock64();
                continue;
            } else {
                // ---- every rank: score, batch-mean gradient norm (published by the leaders), update — redundantly, bit-identically ----
                float gr[9], n2 = 0.f;
#pragma unroll
                for (int c = 0; c < 9; ++c) {
                    gr[c] = (f[c] + sOw[9 * 256 + c]) * inv_std;
                    n2 = fmaf(gr[c], gr[c], n2);
                }
                // ---- batch-mean gradient norm = ONE 64-bit word per step: the leader of every tile adds
                //          (1 << 56 | poisoned << 48 | tile sum as 22.18 fixed point)      [up to 255 tiles]
                //      with a single relaxed RED; thread 0 of every CTA polls the word until the arrival count reaches n_tiles and then
                //      holds the count AND the sum.  Integer addition is associative, so the total is independent of arrival order
                //      (bitwise reproducible, identical in every CTA) and exact to 2^-19 per tile; no payload travels beside the word, so
                //      no release/acquire pair and no second round trip for the partials.  (Measured predecessors: partial array +
                //      RED.release counter + acquire poll + __ldcg of the partials, 2.8 k + 0.9 k cycles per step; per-warp tagged words
                //      polled by every warp 8.2 k; tagged tile sums polled by one warp per CTA 4.4 k.)  A tile whose sum is NaN or
                //      >= 2^22 marks the word poisoned and the step's norm becomes NaN, as it would be (or diverge) in the reference.
                if (leader) {
                    const float wsum = warp_sum(valid ? sqrtf(n2) : 0.f);
                    if (lane == 0) s_red[q] = wsum;
                    named_bar_sync(2, 128);
                    if (tid == 0) {
                        if (ds) ds[14] = clock64();
                        const float tile_sum = (s_red[0] + s_red[1]) + (s_red[2] + s_red[3]);
                        const float scaled = tile_sum * nsum_scale;
                        const bool ok = scaled >= 0.f && scaled < kNormSumLimit * kNormSumScale;   // false for NaN
                        const unsigned long long fx = ok ? __float2ull_rn(scaled) : 0ull;
                        red_relaxed_add_u64(p.acc + step, (1ull << 56) | (ok ? 0ull : (1ull << 48)) | fx);
                    }
                }
                // the predictor drift (0 - g^2 s) dt does not depend on the batch norm: computed while the word is in flight
                float pd[9];
#pragma unroll
                for (int c = 0; c < 9; ++c) pd[c] = (0.0f - g2_sde * gr[c]) * step_size;
                if (tid == 0) {
                    unsigned long long v;
                    do {
                        v = ld_relaxed_gpu_u64(p.acc + step);
                    } while ((unsigned)(v >> 56) < (unsigned)n_tiles);
                    const bool poisoned = ((v >> 48) & 255ull) != 0ull;
                    s_red[4] = poisoned ? __int_as_float(0x7fc00000) : (float)((double)(v & ((1ull << 48) - 1ull)) / (double)nsum_scale);
                    if (ds) ds[15] = clock64();
                }
                named_bar_sync(2, 128);
                if (ds) ds[12] = clock64();
                // Everything between the grid word and the next x is ONE warp per scheduler running a dependent chain with the tensor
                // core idle, so this update spends latency, not throughput: the reference's divisions by a norm are multiplications by
                // MUFU.RSQ reciprocals (<= 2 ulp, far below the bf16x3 score's 2^-17), q = snr.R / sum is one fast division and
                // sqrt(2 ls) = sqrt(4 q^2) = 2 q.  The FFMA parity kernel keeps the IEEE forms (pc_row_update).
                const float tot = s_red[4];
                const float qq = __fdividef(snr_R, tot);                                                     // snr |z| / mean|s| (:130-131)
                const float ls = 2.0f * (qq * qq), sq2ls = 2.0f * qq;
                if (valid) {
                    float m[9];
#pragma unroll
                    for (int c = 0; c < 9; ++c) x[c] = (x[c] + ls * gr[c]) + sq2ls * z1[c];                 // corrector (samplers.py:132)
                    {
                        const float i1 = rsqrtf(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);                   // (:142-143), no eps
                        const float i3 = rsqrtf(x[3] * x[3] + x[4] * x[4] + x[5] * x[5]);
                        x[0] *= i1; x[1] *= i1; x[2] *= i1;
                        x[3] *= i3; x[4] *= i3; x[5] *= i3;
                    }
#pragma unroll
                    for (int c = 0; c < 9; ++c) m[c] = x[c] + pd[c];                                         // predictor mean (:147-148), sign as written
#pragma unroll
                    for (int c = 0; c < 9; ++c) x[c] = m[c] + g_sqrt_step * z2[c];                           // (:149)
                    gram_schmidt6_rsq(x);                                                                    // (:152)
                    if (leader) {
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
                }
                publish_x();
                if (ds) ds[13] = clock64();
            }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if constexpr (kTeam > 1) cluster_sync_all();          // nobody leaves while a team mate may still write into its mailbox
    if (warp == kTcRowWarps) tmem_dealloc(tmem_base, 512);
}

template <bool kW16, int kTeam>
__global__ void __launch_bounds__(kTcThreads, 1)   // 10 warps are allocated as 12 (granularity 4): <= 168 registers per thread
tc_pc_sampler_kernel(TcPcParams tp) {
    tc_sampler_body<false, kW16, kTeam>(tp);
}
template <bool kW16, int kTeam>
__global__ void __launch_bounds__(kTcThreads, 1)
tc_ode_sampler_kernel(TcPcParams tp) {
    tc_sampler_body<true, kW16, kTeam>(tp);
}

// ---- the 12 instantiations (sampler x arithmetic x team size) and how one is chosen ----
using TcKernel = void (*)(TcPcParams);
struct TcVariant {
    TcKernel fn;
    int team;
    uint32_t smem;
    int max_tiles;     // co-resident tiles on the current device model (-1 = not asked yet); per process
};
template <bool kOde, bool kW16, int kTeam> static TcVariant make_variant() {
    TcKernel fn;
    if constexpr (kOde) fn = tc_ode_sampler_kernel<kW16, kTeam>;
    else fn = tc_pc_sampler_kernel<kW16, kTeam>;
    return TcVariant{fn, kTeam, TcSmem<kOde, kTeam>::kBytes, -1};
}
static TcVariant g_variants[2][2][3] = {
    {{make_variant<false, false, 1>(), make_variant<false, false, 2>(), make_variant<false, false, 4>()},
     {make_variant<false, true, 1>(), make_variant<false, true, 2>(), make_variant<false, true, 4>()}},
    {{make_variant<true, false, 1>(), make_variant<true, false, 2>(), make_variant<true, false, 4>()},
     {make_variant<true, true, 1>(), make_variant<true, true, 2>(), make_variant<true, true, 4>()}}};
static std::atomic<int> g_forced_team{0};

// Nsight Compute cannot launch a kernel that is both clustered and cooperative (every replay mode ends in LaunchFailed,
// profiles/README): under a profiler (its injection environment, or GPB_PROFILE_NO_COOP=1) the cooperative attribute is dropped.
// Co-residency, which the grid barrier needs, is still established by the occupancy check (1 CTA per SM, grid <= what fits, and a
// profiler serialises kernels, so the device is idle).
static bool profiler_attached() {
    static const bool attached = [] {
        const char *no_coop = getenv("GPB_PROFILE_NO_COOP");
        if (no_coop) return no_coop[0] == '1';
        for (const char *name : {"NV_NSIGHT_INJECTION_TRANSPORT_TYPE", "NV_NSIGHT_INJECTION_PORT_BASE", "NV_COMPUTE_PROFILER_PERFWORKS_DIR",
                                 "NV_TPS_LAUNCH_TOKEN", "CUDA_INJECTION64_PATH", "NVTX_INJECTION64_PATH"})
            if (const char *v = getenv(name); v && v[0]) return true;
        return false;
    }();
    return attached;
}

static void fill_launch_config(const TcVariant &v, int n_tiles, cudaStream_t st, bool cooperative, cudaLaunchConfig_t &cfg,
                               cudaLaunchAttribute (&attrs)[2]) {
    cfg = cudaLaunchConfig_t{};
    cfg.gridDim = dim3(n_tiles * v.team);
    cfg.blockDim = dim3(kTcThreads);
    cfg.dynamicSmemBytes = v.smem;
    cfg.stream = st;
    int n = 0;
    if (v.team > 1) {
        attrs[n].id = cudaLaunchAttributeClusterDimension;
        attrs[n].val.clusterDim.x = v.team;
        attrs[n].val.clusterDim.y = 1;
        attrs[n].val.clusterDim.z = 1;
        ++n;
    }
    if (cooperative) {
        attrs[n].id = cudaLaunchAttributeCooperative;
        attrs[n].val.cooperative = 1;
        ++n;
    }
    cfg.attrs = attrs;
    cfg.numAttrs = n;
}

// how many 128-row tiles this variant can hold co-resident (one team per tile, one CTA per SM); 0 on error
static int variant_max_tiles(TcVariant &v) {
    if (v.max_tiles >= 0) return v.max_tiles;
    int n = 0;
    if (cudaFuncSetAttribute(v.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)v.smem) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    if (v.team > 1) {
        cudaLaunchConfig_t cfg;
        cudaLaunchAttribute attrs[2];
        fill_launch_config(v, 64, nullptr, false, cfg, attrs);
        if (cudaOccupancyMaxActiveClusters(&n, v.fn, &cfg) != cudaSuccess) {
            cudaGetLastError();
            return 0;
        }
    } else {
        int dev = 0, sms = 0, per_sm = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, v.fn, kTcThreads, v.smem) != cudaSuccess) {
            cudaGetLastError();
            return 0;
        }
        n = sms * per_sm;
    }
    v.max_tiles = n < 255 ? n : 255;               // the PC kernel's reduction word counts at most 255 tiles
    return v.max_tiles;
}

// team size for n_tiles tiles: the largest team (shortest per-step dependent chain) whose clusters are all co-resident
static TcVariant *pick_variant(bool ode, bool w16, int n_tiles) {
    const int forced = g_forced_team.load();
    for (int ti = 2; ti >= 0; --ti) {
        TcVariant &v = g_variants[ode][w16][ti];
        if (forced && v.team != forced) continue;
        if (n_tiles <= variant_max_tiles(v)) return &v;
    }
    return nullptr;
}

// cluster (the tile team, DSMEM) + cooperative (grid barrier => all CTAs must be co-resident) launch
static int launch_tc_sampler(TcVariant &v, const char *what, const TcPcParams &tp, int n_tiles, cudaStream_t st) {
    GPB_CUDA(cudaFuncSetAttribute(v.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)v.smem));
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attrs[2];
    fill_launch_config(v, n_tiles, st, !profiler_attached(), cfg, attrs);
    GPB_CUDA(cudaLaunchKernelEx(&cfg, v.fn, tp));
    g_launches.fetch_add(1);
    return GPB_OK;
}

}  // namespace gpb

using namespace gpb;

extern "C" size_t gpb_trunk_tc_stream_bytes(void) { return (size_t)2 * TcStream<false, 4>::kSlotsPerStep * kSlotBytes; }
extern "C" size_t gpb_trunk_tc16_stream_bytes(void) { return (size_t)2 * TcStream<true, 4>::kSlotsPerStep * kSlotBytes; }

// Tuning / test knob: force the tile-team size of the tensor-core samplers (1, 2 or 4; 0 = chosen from the row count).
extern "C" int gpb_set_tc_team(int team) {
    GPB_REQUIRE(team == 0 || team == 1 || team == 2 || team == 4, "set_tc_team: team must be 0 (auto), 1, 2 or 4");
    g_forced_team.store(team);
    return GPB_OK;
}

// Largest row count R the tensor-core samplers accept on the current device for K candidates per object (0 = not at all):
// every 128-row tile needs one co-resident CTA (team of 1: one tile per SM), and a tile may span at most kMaxObjPerTile objects.
extern "C" int gpb_sampler_tc_max_rows(int K) {
    if (K < 1 || 127 / K + 2 > kMaxObjPerTile) return 0;      // K >= 19
    const int forced = g_forced_team.load();
    int best = 0;
    for (int ti = 0; ti < 3; ++ti) {
        if (forced && g_variants[1][0][ti].team != forced) continue;
        const int a = variant_max_tiles(g_variants[1][0][ti]), b = variant_max_tiles(g_variants[1][1][ti]);   // the ODE kernels: the larger footprint
        const int n = a < b ? a : b;
        best = n > best ? n : best;
    }
    return best * kTcRows;
}

static int sample_pc_tc_impl(bool w16, const float *x0, int R, int K, int num_steps, float snr, const float *obj_bias, const float *W,
                             const void *tc_stream, const float *pts_center, const float *step_noise, uint64_t seed,
                             const float *time_grid, float *mean_x, float *process, void *workspace, size_t workspace_bytes,
                             unsigned long long *dbg, void *stream) {
    const int dbg_cta_sel = (int)(seed >> 56);   // profiling aid: the top byte of the seed selects the recording CTA when dbg != NULL
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
    const int n_tiles = (R + kTcRows - 1) / kTcRows;
    TcVariant *v = pick_variant(false, w16, n_tiles);
    GPB_REQUIRE(v != nullptr, "sample_pc_tc: R=%d is %d tiles of 128 rows; more than fit co-resident on this device (%d); use gpb_sample_pc",
                R, n_tiles, variant_max_tiles(g_variants[0][w16][0]));

    SamplerWs w = carve_sampler(workspace, R, num_steps);
    GPB_CUDA(cudaMemsetAsync(w.barrier, 0, 256, st));
    GPB_CUDA(cudaMemsetAsync(w.acc, 0, (size_t)num_steps * sizeof(unsigned long long), st));
    int rc = launch_time_bias_table(time_grid, num_steps, W, w.tb_table, st);
    if (rc) return rc;

    TcPcParams tp{};
    PcParams &p = tp.pc;
    p.x0 = x0; p.R = R; p.K = K; p.T = num_steps; p.snr = snr;
    p.obj_bias = obj_bias; p.W = W; p.pts_center = pts_center; p.noise = step_noise; p.seed = seed;
    p.ts = time_grid; p.tb_table = w.tb_table; p.partial = w.partial; p.barrier = w.barrier; p.acc = w.acc;
    p.mean_x = mean_x; p.process = process; p.tiles_per_cta = 1; p.dbg = dbg; p.dbg_cta = dbg ? dbg_cta_sel : 0;
    tp.wstream = reinterpret_cast<const uint8_t *>(tc_stream);
    return launch_tc_sampler(*v, "sample_pc_tc", tp, n_tiles, st);
}

extern "C" int gpb_sample_pc_tc_dbg(const float *x0, int R, int K, int num_steps, float snr, const float *obj_bias, const float *W,
                                    const void *tc_stream, const float *pts_center, const float *step_noise, uint64_t seed,
                                    const float *time_grid, float *mean_x, float *process, void *workspace, size_t workspace_bytes,
                                    unsigned long long *dbg, void *stream) {
    return sample_pc_tc_impl(false, x0, R, K, num_steps, snr, obj_bias, W, tc_stream, pts_center, step_noise, seed, time_grid, mean_x,
                             process, workspace, workspace_bytes, dbg, stream);
}

extern "C" int gpb_sample_pc_tc16(const float *x0, int R, int K, int num_steps, float snr, const float *obj_bias, const float *W,
                                  const void *tc16_stream, const float *pts_center, const float *step_noise, uint64_t seed,
                                  const float *time_grid, float *mean_x, float *process, void *workspace, size_t workspace_bytes,
                                  unsigned long long *dbg, void *stream) {
    return sample_pc_tc_impl(true, x0, R, K, num_steps, snr, obj_bias, W, tc16_stream, pts_center, step_noise, seed, time_grid, mean_x,
                             process, workspace, workspace_bytes, dbg, stream);
}

static int sample_ode_tc_impl(bool w16, const float *x0, int R, int K, double T0, double rtol, double atol, int denoise_steps,
                              const float *obj_bias, const float *W, const void *tc_stream, const float *pts_center, double *pose,
                              int *stats, double *process, int process_cap, const double *t_eval, int n_t_eval,
                              void *workspace, size_t workspace_bytes, unsigned long long *dbg, int dbg_evals,
                              void *stream) {
    GPB_REQUIRE(R >= 0 && K >= 1, "sample_ode_tc: need R >= 0, K >= 1");
    GPB_REQUIRE(!process || (t_eval ? n_t_eval > 0 : process_cap > 0), "sample_ode_tc: process output needs process_cap > 0 or t_eval / n_t_eval");
    if (R == 0) return GPB_OK;
    GPB_REQUIRE(x0 && obj_bias && W && tc_stream && pts_center && pose && workspace, "sample_ode_tc: NULL buffer");
    GPB_REQUIRE(T0 > 1e-5 && rtol > 0 && atol > 0, "sample_ode_tc: need T0 > eps and positive tolerances");
    GPB_REQUIRE(127 / K + 2 <= kMaxObjPerTile, "sample_ode_tc: K=%d too small (a 128-row tile may span at most %d objects); "
                "use gpb_sample_ode", K, kMaxObjPerTile);
    GPB_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0 && (reinterpret_cast<uintptr_t>(tc_stream) & 15) == 0,
                "sample_ode_tc: workspace must be 256-byte and the weight stream 16-byte aligned");
    SamplerWs w = carve_sampler(workspace, R, 1);
    if (workspace_bytes < w.bytes) {
        set_error("sample_ode_tc: workspace %zu < required %zu bytes", workspace_bytes, w.bytes);
        return GPB_EWORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int n_tiles = (R + kTcRows - 1) / kTcRows;
    TcVariant *v = pick_variant(true, w16, n_tiles);
    GPB_REQUIRE(v != nullptr && n_tiles * v->team <= 160, "sample_ode_tc: R=%d is %d tiles of 128 rows; more than fit co-resident on this "
                "device (%d); use gpb_sample_ode", R, n_tiles, variant_max_tiles(g_variants[1][w16][0]));
    GPB_CUDA(cudaMemsetAsync(w.barrier, 0, 256, st));

    TcPcParams tp{};
    PcParams &p = tp.pc;
    p.x0 = x0; p.R = R; p.K = K; p.T = dbg ? dbg_evals : 0;       // ODE: T only bounds the profiling record [2][T][16]
    p.obj_bias = obj_bias; p.W = W; p.pts_center = pts_center; p.barrier = w.barrier; p.tiles_per_cta = 1;
    p.dbg = dbg; p.dbg_cta = 0;
    tp.wstream = reinterpret_cast<const uint8_t *>(tc_stream);
    tp.ode.T0 = T0; tp.ode.rtol = rtol; tp.ode.atol = atol; tp.ode.denoise_steps = denoise_steps;
    tp.ode.Kst = w.Kst; tp.ode.partial = reinterpret_cast<double *>(w.partial); tp.ode.tb_cta = w.tb_cta;
    tp.ode.pose = pose; tp.ode.stats = stats;
    tp.ode.proc = OdeProcess{process, process ? t_eval : nullptr, process_cap, n_t_eval};
    return launch_tc_sampler(*v, "sample_ode_tc", tp, n_tiles, st);
}

extern "C" int gpb_sample_ode_tc_dbg(const float *x0, int R, int K, double T0, double rtol, double atol, int denoise_steps,
                                     const float *obj_bias, const float *W, const void *tc_stream, const float *pts_center, double *pose,
                                     int *stats, void *workspace, size_t workspace_bytes, unsigned long long *dbg, int dbg_evals,
                                     void *stream) {
    return sample_ode_tc_impl(false, x0, R, K, T0, rtol, atol, denoise_steps, obj_bias, W, tc_stream, pts_center, pose, stats, nullptr, 0,
                              nullptr, 0, workspace, workspace_bytes, dbg, dbg_evals, stream);
}

extern "C" int gpb_sample_ode_tc16(const float *x0, int R, int K, double T0, double rtol, double atol, int denoise_steps,
                                   const float *obj_bias, const float *W, const void *tc16_stream, const float *pts_center, double *pose,
                                   int *stats, double *process, int process_cap, const double *t_eval, int n_t_eval,
                                   void *workspace, size_t workspace_bytes, void *stream) {
    return sample_ode_tc_impl(true, x0, R, K, T0, rtol, atol, denoise_steps, obj_bias, W, tc16_stream, pts_center, pose, stats, process,
                              process_cap, t_eval, n_t_eval, workspace, workspace_bytes, nullptr, 0, stream);
}

extern "C" int gpb_sample_ode_tc16_dbg(const float *x0, int R, int K, double T0, double rtol, double atol, int denoise_steps,
                                       const float *obj_bias, const float *W, const void *tc16_stream, const float *pts_center, double *pose,
                                       int *stats, void *workspace, size_t workspace_bytes, unsigned long long *dbg, int dbg_evals,
                                       void *stream) {
    return sample_ode_tc_impl(true, x0, R, K, T0, rtol, atol, denoise_steps, obj_bias, W, tc16_stream, pts_center, pose, stats, nullptr, 0,
                              nullptr, 0, workspace, workspace_bytes, dbg, dbg_evals, stream);
}

extern "C" int gpb_sample_pc_tc(const float *x0, int R, int K, int num_steps, float snr, const float *obj_bias, const float *W,
                                const void *tc_stream, const float *pts_center, const float *step_noise, uint64_t seed,
                                const float *time_grid, float *mean_x, float *process, void *workspace, size_t workspace_bytes,
                                void *stream) {
    return gpb_sample_pc_tc_dbg(x0, R, K, num_steps, snr, obj_bias, W, tc_stream, pts_center, step_noise, seed, time_grid, mean_x,
                                process, workspace, workspace_bytes, nullptr, stream);
}

extern "C" int gpb_sample_ode_tc(const float *x0, int R, int K, double T0, double rtol, double atol, int denoise_steps,
                                 const float *obj_bias, const float *W, const void *tc_stream, const float *pts_center, double *pose,
                                 int *stats, double *process, int process_cap, const double *t_eval, int n_t_eval,
                                 void *workspace, size_t workspace_bytes, void *stream) {
    return sample_ode_tc_impl(false, x0, R, K, T0, rtol, atol, denoise_steps, obj_bias, W, tc_stream, pts_center, pose, stats, process,
                              process_cap, t_eval, n_t_eval, workspace, workspace_bytes, nullptr, 0, stream);
}
