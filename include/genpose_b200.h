/*
 * genpose_b200 — C ABI of the B200-native (sm_100a) GenPose per-object inference hot path.
 *
 * Boundary rules (all entry points):
 *   - plain C, `extern "C"`; raw DEVICE pointers unless a parameter is documented as host;
 *   - caller owns every buffer (no hidden allocation, no internal state, re-entrant, thread-safe);
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*; NULL = legacy default stream) and
 *     the call returns without synchronising;
 *   - returns GPB_OK (0) or a negative GPB_E* code; never calls exit() (the reference's launchers
 *     do: src/sampling_gpu.cu:39-43).  gpb_last_error_string() describes the last failure of the
 *     calling thread.
 *
 * Reference paths below are relative to the GenPose checkout; "pointnet2/" abbreviates
 * networks/pts_encoder/pointnet2_utils/pointnet2/.
 */
#ifndef GENPOSE_B200_H
#define GENPOSE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GPB_OK 0
#define GPB_EINVAL (-1)   /* bad argument (NULL pointer, unsupported shape) */
#define GPB_ECUDA (-2)    /* a CUDA runtime call or launch failed */
#define GPB_EWORKSPACE (-3) /* workspace too small */

#define GPB_ABI_VERSION 1

/* exported from libgenpose_b200.so (everything else is built with -fvisibility=hidden) */
#if defined(__GNUC__)
#define GPB_API __attribute__((visibility("default")))
#else
#define GPB_API
#endif

GPB_API int gpb_abi_version(void);
GPB_API const char *gpb_last_error_string(void);

/* ------------------------------------------------------------------------------------------------
 * (1) COMPAT LAYER — the four forward ops of the reference's pybind module `pointnet2_cuda`
 *     (pointnet2/src/pointnet2_api.cpp:10-24), same argument order and buffer conventions, so the
 *     reference's own pointnet2_utils.py can run on them unchanged.
 *     Indices are bit-exact with the reference kernels, including tie-breaking.
 * ---------------------------------------------------------------------------------------------- */

/* replaces furthest_point_sampling_wrapper (pointnet2/src/sampling.cpp:40-51; kernel
 * sampling_gpu.cu:93-209).  xyz [b,n,3] f32, temp [b,n] f32 scratch (the reference requires it
 * pre-filled with 1e10, pointnet2_utils.py:27; this implementation ignores its contents but writes
 * the final min-distances back for parity), idx [b,m] i32 out.  m <= n. */
GPB_API int gpb_furthest_point_sampling(int b, int n, int m, const float *xyz, float *temp, int *idx, void *stream);

/* replaces gather_points_wrapper (sampling.cpp:13-23; sampling_gpu.cu:8-24):
 * out[b,c,j] = points[b,c,idx[b,j]];  points [b,c,n], idx [b,npoints], out [b,c,npoints]. */
GPB_API int gpb_gather_points(int b, int c, int n, int npoints, const float *points, const int *idx, float *out,
                      void *stream);

/* replaces ball_query_wrapper (pointnet2/src/ball_query.cpp:16-28; ball_query_gpu.cu:9-45).
 * new_xyz [b,m,3], xyz [b,n,3], idx [b,m,nsample] i32 out.  Unlike the reference the caller need not
 * pre-zero idx (pointnet2_utils.py:219): empty balls are written as zeros. */
GPB_API int gpb_ball_query(int b, int n, int m, float radius, int nsample, const float *new_xyz, const float *xyz,
                   int *idx, void *stream);

/* replaces group_points_wrapper (pointnet2/src/group_points.cpp:26-37; group_points_gpu.cu:47-66):
 * out[b,c,i,s] = points[b,c,idx[b,i,s]];  points [b,c,n], idx [b,npoints,nsample]. */
GPB_API int gpb_group_points(int b, int c, int n, int npoints, int nsample, const float *points, const int *idx,
                     float *out, void *stream);

/* ------------------------------------------------------------------------------------------------
 * (2) FUSED PATH — what networks/posenet_agent.py::PoseNet.pred_func / get_energy call.
 *     Weight blobs are produced on the host by genpose_b200/weights.py from a reference
 *     `model_state_dict` (BatchNorm folded, heads stacked, W1 split into [pts | t | pose] blocks);
 *     their layouts are fixed by the gpb_*_weights_floats() sizes and documented in DESIGN.md §3.
 * ---------------------------------------------------------------------------------------------- */

/* number of floats in the packed encoder / score-trunk weight blobs */
GPB_API size_t gpb_encoder_weights_floats(void);
GPB_API size_t gpb_trunk_weights_floats(void);

/* scratch bytes gpb_encode needs for a batch of B objects */
GPB_API size_t gpb_encode_workspace_bytes(int B);

/* replaces Pointnet2ClsMSG.forward (networks/pts_encoder/pointnet2.py:203-211) == GFObjectPose
 * mode 'pts_feature' (networks/posenet.py:71-91,169-171).
 * pts [B,1024,3] f32 (camera frame, metres, NOT centred) -> pts_feat [B,1024] f32.
 * Optional debug outputs (may be NULL): fps_idx1 [B,512], fps_idx2 [B,256], fps_idx3 [B,128] i32. */
GPB_API int gpb_encode(const float *pts, int B, const float *enc_weights, float *pts_feat, void *workspace,
               size_t workspace_bytes, int *fps_idx1, int *fps_idx2, int *fps_idx3, void *stream);

/* Same function with the MLPs of set-abstraction level 1 (wide scale), 2, 3 and 4 (GroupAll) on the tensor cores
 * (tcgen05, bf16x3 error-compensated split, fp32 accumulation; pts_feat within 1e-5 relative of gpb_encode).
 * enc_tc = the operand image made by genpose_b200/weights.py::pack_encoder_tc, gpb_encoder_tc_bytes() long,
 * 128-byte aligned.  FPS, ball query (same index semantics) and the narrow scale of level 1 stay on the CUDA cores. */
GPB_API size_t gpb_encoder_tc_bytes(void);
GPB_API int gpb_encode_tc(const float *pts, int B, const float *enc_weights, const void *enc_tc, float *pts_feat,
                  void *workspace, size_t workspace_bytes, int *fps_idx1, int *fps_idx2, int *fps_idx3, void *stream);

/* Per-object hoisted head bias: obj_bias[b, 0:768] = A_pts . pts_feat[b] + a   (the constant 67 % of
 * PoseScoreNet.forward, networks/gf_algorithms/scorenet.py:204-216).  obj_bias [B,768]. */
GPB_API int gpb_object_bias(const float *pts_feat, int B, const float *trunk_weights, float *obj_bias, void *stream);

/* replaces GFObjectPose.forward(mode='score') (networks/posenet.py:160-162 ->
 * PoseScoreNet.forward scorenet.py:178-222) for a batch-constant time t:
 *   out[r, 0:9] = f_theta(pose[r], t, object row_object(r)) / (sigma(t) + 1e-7)     if divide_mode == 1
 *               = f_theta / sigma(t)                                               if divide_mode == 2 (energynet.py:166-167)
 *               = f_theta                                                          if divide_mode == 0
 * rows R = B*K, row r belongs to object r / K.  pose [R,9], out [R,9]. */
GPB_API int gpb_trunk_eval(const float *pose, int R, int K, float t, const float *obj_bias, const float *trunk_weights,
                   int divide_mode, float *out, void *stream);

/* scratch bytes the samplers need */
GPB_API size_t gpb_sampler_workspace_bytes(int R, int num_steps);

/* replaces cond_pc_sampler (networks/gf_algorithms/samplers.py:102-160), pose_mode 'rot_matrix', VE SDE.
 *   x0 [R,9]            prior sample (sde.py:26-28: 50*randn), consumed read-only
 *   step_noise          [T,2,R,9] explicit z1,z2 (parity mode) or NULL -> in-kernel Philox4x32-10 keyed by `seed`
 *   pts_center [B,3]    added to the translation at the end (samplers.py:157)
 *   time_grid [T]       torch.linspace(1, 1e-5, T) in fp32 (samplers.py:118), computed by the host so that
 *                       the grid is bit-identical to the reference's
 *   mean_x [R,9] out    the reference's `res` (last predictor mean, Gram-Schmidt'd, centre added)
 *   process [R,T,9] out optional (NULL to skip): the recorded noisy iterates `xs` (samplers.py:153-156)
 * One launch runs all T steps; the batch-mean gradient norm (samplers.py:130) is a grid-wide reduction. */
GPB_API int gpb_sample_pc(const float *x0, int R, int K, int num_steps, float snr, const float *obj_bias,
                  const float *trunk_weights, const float *pts_center, const float *step_noise,
                  uint64_t seed, const float *time_grid, float *mean_x, float *process, void *workspace,
                  size_t workspace_bytes, void *stream);

/* Tensor-core (tcgen05, bf16x3 split, fp32 accumulate in TMEM) variant of gpb_sample_pc: same contract plus
 * `tc_stream`, the bf16 operand-image stream of the trunk (gpb_trunk_tc_stream_bytes() bytes, produced by
 * genpose_b200/weights.py::pack_trunk_tc).  Requires K >= 19 (a 128-row tile may span at most 8 objects) and
 * ceil(R/128) <= #SMs; otherwise use gpb_sample_pc.  Poses agree with the fp32 path to ~2e-5. */
GPB_API size_t gpb_trunk_tc_stream_bytes(void);
GPB_API int gpb_sample_pc_tc(const float *x0, int R, int K, int num_steps, float snr, const float *obj_bias,
                             const float *trunk_weights, const void *tc_stream, const float *pts_center,
                             const float *step_noise, uint64_t seed, const float *time_grid, float *mean_x,
                             float *process, void *workspace, size_t workspace_bytes, void *stream);

/* Same, with an optional cycle-stamp buffer (device, [2][T][16] u64; NULL = off): CTA 0's row thread 0 and MMA thread
 * record clock64() at the phase boundaries of every step (tools/tc_phase_times.py decodes them). */
GPB_API int gpb_sample_pc_tc_dbg(const float *x0, int R, int K, int num_steps, float snr, const float *obj_bias,
                                 const float *trunk_weights, const void *tc_stream, const float *pts_center,
                                 const float *step_noise, uint64_t seed, const float *time_grid, float *mean_x,
                                 float *process, void *workspace, size_t workspace_bytes, unsigned long long *dbg,
                                 void *stream);

/* replaces cond_ode_sampler (samplers.py:163-227) + scipy.integrate.solve_ivp(RK45) (samplers.py:205):
 * Dormand-Prince 5(4) with SciPy's step controller, float64 state, fp32 score, one error norm over the
 * whole [R*9] state, followed by the reference's Euler "denoise" step (:209-218).
 *   x0 [R,9] f32        already-noised start (sigma(T0)*randn, plus init_x when tracking, :180)
 *   T0, rtol, atol      float64, as the Python floats the reference hands to solve_ivp (samplers.py:178, :205)
 *   pose [R,9] f64 out  (the reference returns float64, :206-207)
 *   stats [4] i32 out   optional: nfev, accepted, rejected, status
 *   process             optional trajectory, the reference's in_process_sample (`xs`, samplers.py:206, :220-224), float64,
 *                       every state with its rotation part normalised and pts_center added like the reference's:
 *                         t_eval == NULL: [process_cap][R][9] — state 0 = the start, state i = the i-th ACCEPTED RK45
 *                           step (solve_ivp's res.y with t_eval=None); stats[1] + 1 states exist, those beyond
 *                           process_cap are dropped
 *                         t_eval != NULL: [n_t_eval][R][9] — RK45's dense output at t_eval[0..n_t_eval), device float64,
 *                           strictly decreasing from T0 to eps (np.linspace(T0, eps, num_steps), samplers.py:203)
 *                       NULL = no trajectory (then process_cap, t_eval, n_t_eval are ignored) */
GPB_API int gpb_sample_ode(const float *x0, int R, int K, double T0, double rtol, double atol, int denoise_steps,
                   const float *obj_bias, const float *trunk_weights, const float *pts_center, double *pose,
                   int *stats, double *process, int process_cap, const double *t_eval, int n_t_eval,
                   void *workspace, size_t workspace_bytes, void *stream);

/* largest R gpb_sample_pc_tc / gpb_sample_ode_tc accept on the current device for K candidates per object (0 = K too
 * small or no device): every 128-row tile needs one co-resident team of CTAs, and a team may be a single CTA, so
 * one tile per SM (148 tiles = 18,944 rows on a B200). */
GPB_API int gpb_sampler_tc_max_rows(int K);

/* Tuning / test knob: the tensor-core samplers give every 128-row tile to a TEAM of 4, 2 or 1 CTAs (a thread-block
 * cluster; the head layer is split over the ranks).  By default the largest team whose clusters are all co-resident
 * is used (4 up to 33 tiles, 2 up to 66, 1 up to 148 on a B200); team = 1, 2 or 4 forces one, 0 restores the default.
 * Process-wide; results of different team sizes agree to the summation order of the head partials (~1e-6). */
GPB_API int gpb_set_tc_team(int team);

/* gpb_sample_ode on the tensor cores: the same solver (same controller, float64 state, one error norm over the whole
 * batch) with the score network's dense layers evaluated by tcgen05.mma (bf16x3 split, fp32 accumulation in tensor
 * memory), one team of CTAs per 128-row tile as in gpb_sample_pc_tc.  tc_stream as there.
 * Constraints: K >= 19, R <= gpb_sampler_tc_max_rows(K); otherwise use gpb_sample_ode. */
GPB_API int gpb_sample_ode_tc(const float *x0, int R, int K, double T0, double rtol, double atol, int denoise_steps,
                      const float *obj_bias, const float *trunk_weights, const void *tc_stream, const float *pts_center,
                      double *pose, int *stats, double *process, int process_cap, const double *t_eval, int n_t_eval,
                      void *workspace, size_t workspace_bytes, void *stream);

/* profiling aid: as gpb_sample_ode_tc, additionally records clock64 stamps of CTA 0 for the first dbg_evals evaluations
 * into dbg [2][dbg_evals][16] (row thread 0 | MMA warp); see tools/tc_ode_phase_times.py. */
GPB_API int gpb_sample_ode_tc_dbg(const float *x0, int R, int K, double T0, double rtol, double atol, int denoise_steps,
                          const float *obj_bias, const float *trunk_weights, const void *tc_stream, const float *pts_center,
                          double *pose, int *stats, void *workspace, size_t workspace_bytes, unsigned long long *dbg,
                          int dbg_evals, void *stream);
/* the same for the two-product arithmetic (tc16_stream as gpb_sample_ode_tc16, below) */
GPB_API int gpb_sample_ode_tc16_dbg(const float *x0, int R, int K, double T0, double rtol, double atol, int denoise_steps,
                            const float *obj_bias, const float *trunk_weights, const void *tc16_stream, const float *pts_center,
                            double *pose, int *stats, void *workspace, size_t workspace_bytes, unsigned long long *dbg,
                            int dbg_evals, void *stream);

/* gpb_sample_pc_tc / gpb_sample_ode_tc with TWO tensor-core products per K-step instead of three ("f16x2"): the activations
 * of layer 1 and of the heads are split into fp16 hi + lo (tensor memory, 21 mantissa bits), their weights are ONE fp16 image
 * (11-bit mantissa) — the dropped Ahi.Blo product only carried the weights' bits beyond the first image.  (kind::f16 wants
 * one format for both operands: bf16 activations against fp16 weights fault.)  Weights beyond the fp16 range are refused by the
 * packer; activations beyond it saturate.  tc16_stream = gpb_trunk_tc16_stream_bytes() bytes from
 * genpose_b200/weights.py::pack_trunk_tc16 (2 layouts x 33 slots of 16 KiB).  Same constraints, workspace and semantics as
 * the functions they mirror; dbg as in gpb_sample_pc_tc_dbg (may be NULL). */
GPB_API size_t gpb_trunk_tc16_stream_bytes(void);
GPB_API int gpb_sample_pc_tc16(const float *x0, int R, int K, int num_steps, float snr, const float *obj_bias,
                       const float *trunk_weights, const void *tc16_stream, const float *pts_center,
                       const float *step_noise, uint64_t seed, const float *time_grid, float *mean_x, float *process,
                       void *workspace, size_t workspace_bytes, unsigned long long *dbg, void *stream);
GPB_API int gpb_sample_ode_tc16(const float *x0, int R, int K, double T0, double rtol, double atol, int denoise_steps,
                        const float *obj_bias, const float *trunk_weights, const void *tc16_stream, const float *pts_center,
                        double *pose, int *stats, double *process, int process_cap, const double *t_eval, int n_t_eval,
                        void *workspace, size_t workspace_bytes, void *stream);

/* replaces PoseNet.get_energy's arithmetic after the encoder (networks/posenet_agent.py:508-523 ->
 * PoseEnergyNet.get_energy energynet.py:143-198, 'IP' decoupled): energy [B,K,2] = (rot, trans). */
GPB_API int gpb_energy(const float *pose, int R, int K, float t, const float *obj_bias, const float *trunk_weights,
               const float *pts_center, float *energy, void *stream);

/* replaces sort_poses_by_energy (networks/reward.py:131-155) + sort_sRT_by_energy(ratio,'average')
 * (utils/sgpa_utils.py:897-954) + average_quaternion_batch (utils/misc.py:227-249).
 *   pose [B,K,9], energy [B,K,2] -> sorted_pose [B,K,9], sorted_energy [B,K,2], pooled_RT [B,4,4].
 * K <= 128; keep = max(1, int(K * ratio)) is evaluated by the caller (sgpa_utils.py:912).  order [B,K,2] i32: the
 * candidate index at every rank by rotation (…,0) and translation (…,1) energy = torch.sort(energy, dim=1, descending=True)'s
 * indices (reward.py:145), for callers that reorder their own (e.g. float64) copies of the poses.  sorted_pose, sorted_energy,
 * pooled_RT and order may each be NULL to skip that output. */
GPB_API int gpb_rank_pool(const float *pose, const float *energy, int B, int K, int keep, float *sorted_pose,
                  float *sorted_energy, float *pooled_RT, int *order, void *stream);

/* ------------------------------------------------------------------------------------------------
 * (3) BOOK-KEEPING
 * ---------------------------------------------------------------------------------------------- */
/* Kernel launch counter (for bench.py's gpu_launches claim): number of kernels this library has
 * launched since load, across all threads. */
GPB_API uint64_t gpb_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * (4) SELF-TEST of the tcgen05 building block (one CTA: D[128,N] = A[128,K] . B[N,K]^T, bf16 split, fp32
 *     accumulate in TMEM).  A fp32 row-major; Bhi/Blo = bf16 operand images in the canonical K-major
 *     no-swizzle layout (genpose_b200/weights.py::umma_image).  variant/swap_fields select layout
 *     conventions under test (swap_fields bit 0 = LBO/SBO fields exchanged; bit 1 = issue M = 64 instructions, a timing aid
 *     whose D is not the product; bit 2 = Bhi holds fp16 values and the products are Ahi.B, Alo.B — the mixed-format
 *     instruction of the two-product samplers, n_terms <= 2); n_terms 1..3 = how many of the bf16x3 products are accumulated; a_tmem = 1 feeds the
 *     A operand from tensor memory (tcgen05.st + the TS form of tcgen05.mma) instead of shared memory; repeat > 1 re-issues
 *     the whole K loop (timing aid: cycles_out[0] = issue cycles, cycles_out[1] = cycles until completion; may be NULL).
 * ---------------------------------------------------------------------------------------------- */
GPB_API int gpb_selftest_umma(const float *A, const uint16_t *Bhi, const uint16_t *Blo, float *D, int K, int N,
                              int variant, int swap_fields, int n_terms, int a_tmem, int repeat,
                              unsigned long long *cycles_out, void *stream);

/* ------------------------------------------------------------------------------------------------
 * 3. point-cloud preparation (the step before the path)
 * ------------------------------------------------------------------------------------------------
 * replaces, per detected instance, the body of the instance loop of detect_mrcnn_genpose
 * (runners/evaluation_single.py:168-216): three cv2.warpAffine(INTER_NEAREST) crops to 256 x 256
 * (utils/datasets_utils.py:82-95: coordinate map, mask, depth), depth_to_pcl (:107-119), "/ 1000.0" (:211) and
 * sample_points (:121-133).  One launch per frame, all instances.
 *   depth  [H,W] u16 (device)       millimetres, 0 = no measurement (load_depth, utils/sgpa_utils.py:194-211)
 *   masks  u8 (device)              Mask-RCNN masks; instance i at pixel p: masks[p*mask_pixel_stride + i*mask_inst_stride]
 *                                   ([H,W,n] bool as the detector pickles store it: strides (n, 1))
 *   trans  [n_inst,6] f64 (device)  forward 2x3 crop matrices exactly as get_affine_transform returns them
 *                                   (utils/datasets_utils.py:97-138, cv2.getAffineTransform)
 *   intrinsics [4] f32 (HOST)       cx, cy, fx, fy of the float32 camera matrix (evaluation_single.py:50,54)
 *   subset_ids [n_inst,1024] i32 (device) or NULL   the reference's np.random.permutation(n)[:1024] per instance with
 *                                   n > 1024 valid pixels (parity mode); NULL = keyed in-kernel permutation of `seed`
 *   pts    [n_inst,1024,3] f32 out  camera-frame metres; all zero for an instance the reference skips
 *   n_valid [n_inst] i32 out        valid crop pixels; <= 1 means "skip the instance" (:201-209) */
GPB_API int gpb_prepare_clouds(const unsigned short *depth, const unsigned char *masks, long long mask_pixel_stride,
                       long long mask_inst_stride, int H, int W, int n_inst, const double *trans, const float *intrinsics,
                       const int *subset_ids, uint64_t seed, float *pts, int *n_valid, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* GENPOSE_B200_H */
