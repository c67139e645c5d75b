// Energy ranking and pose pooling: one CTA per object, K <= 128 candidates.
//   * sort_poses_by_energy (networks/reward.py:131-155): descending sort of the rot- and trans-energies,
//     rotation part of the pose follows the rot order, translation part the trans order, independently;
//   * sort_sRT_by_energy(ratio, 'average') (utils/sgpa_utils.py:897-954): keep the first `keep` sorted poses,
//     6D -> matrix (utils/misc.py:136: Gram-Schmidt, b1,b2,b1xb2 as COLUMNS) -> quaternion (pytorch3d
//     matrix_to_quaternion) -> average_quaternion_batch (utils/misc.py:227-249: sign-align w>0, mean outer
//     product, eigenvector of the largest eigenvalue) -> matrix; translation = arithmetic mean.
#include "common.cuh"

namespace gpb {

constexpr int kMaxK = 128;

__device__ __forceinline__ void rot6d_to_quat(const float *v, float *q) {
    // rotation_6d_to_matrix (F.normalize eps 1e-12) then columns b1,b2,b3  => m[r][c] = b_c[r]
    const float n1 = fmaxf(sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]), 1e-12f);
    const float b1[3] = {v[0] / n1, v[1] / n1, v[2] / n1};
    const float d = b1[0] * v[3] + b1[1] * v[4] + b1[2] * v[5];
    float b2[3] = {v[3] - d * b1[0], v[4] - d * b1[1], v[5] - d * b1[2]};
    const float n2 = fmaxf(sqrtf(b2[0] * b2[0] + b2[1] * b2[1] + b2[2] * b2[2]), 1e-12f);
    b2[0] /= n2; b2[1] /= n2; b2[2] /= n2;
    const float b3[3] = {b1[1] * b2[2] - b1[2] * b2[1], b1[2] * b2[0] - b1[0] * b2[2], b1[0] * b2[1] - b1[1] * b2[0]};
    const float m00 = b1[0], m01 = b2[0], m02 = b3[0];
    const float m10 = b1[1], m11 = b2[1], m12 = b3[1];
    const float m20 = b1[2], m21 = b2[2], m22 = b3[2];
    // pytorch3d v0.7.2 matrix_to_quaternion: q_abs = sqrt(max(0, 1 +- m00 +- m11 +- m22)), best-conditioned candidate
    float qa[4] = {1.0f + m00 + m11 + m22, 1.0f + m00 - m11 - m22, 1.0f - m00 + m11 - m22, 1.0f - m00 - m11 + m22};
#pragma unroll
    for (int i = 0; i < 4; ++i) qa[i] = qa[i] > 0.f ? sqrtf(qa[i]) : 0.f;
    int best = 0;
#pragma unroll
    for (int i = 1; i < 4; ++i)
        if (qa[i] > qa[best]) best = i;   // argmax returns the first maximum
    float c[4];
    if (best == 0) { c[0] = qa[0] * qa[0]; c[1] = m21 - m12; c[2] = m02 - m20; c[3] = m10 - m01; }
    else if (best == 1) { c[0] = m21 - m12; c[1] = qa[1] * qa[1]; c[2] = m10 + m01; c[3] = m02 + m20; }
    else if (best == 2) { c[0] = m02 - m20; c[1] = m10 + m01; c[2] = qa[2] * qa[2]; c[3] = m12 + m21; }
    else { c[0] = m10 - m01; c[1] = m20 + m02; c[2] = m21 + m12; c[3] = qa[3] * qa[3]; }
    const float den = 2.0f * fmaxf(qa[best], 0.1f);
#pragma unroll
    for (int i = 0; i < 4; ++i) q[i] = c[i] / den;
}

// eigenvector of the largest eigenvalue of a symmetric 4x4 matrix (cyclic Jacobi, double precision)
__device__ void top_eigenvector4(double A[4][4], double *vec) {
    double V[4][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}};
    for (int sweep = 0; sweep < 32; ++sweep) {
        double off = 0.0;
        for (int i = 0; i < 4; ++i)
            for (int j = i + 1; j < 4; ++j) off += A[i][j] * A[i][j];
        if (off < 1e-30) break;
        for (int pp = 0; pp < 4; ++pp)
            for (int qq = pp + 1; qq < 4; ++qq) {
                if (fabs(A[pp][qq]) < 1e-300) continue;
                const double theta = (A[qq][qq] - A[pp][pp]) / (2.0 * A[pp][qq]);
                const double tt = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double cs = 1.0 / sqrt(tt * tt + 1.0), sn = tt * cs;
                for (int k = 0; k < 4; ++k) {
                    const double akp = A[k][pp], akq = A[k][qq];
                    A[k][pp] = cs * akp - sn * akq;
                    A[k][qq] = sn * akp + cs * akq;
                }
                for (int k = 0; k < 4; ++k) {
                    const double apk = A[pp][k], aqk = A[qq][k];
                    A[pp][k] = cs * apk - sn * aqk;
                    A[qq][k] = sn * apk + cs * aqk;
                }
                for (int k = 0; k < 4; ++k) {
                    const double vkp = V[k][pp], vkq = V[k][qq];
                    V[k][pp] = cs * vkp - sn * vkq;
                    V[k][qq] = sn * vkp + cs * vkq;
                }
            }
    }
    int best = 0;
    for (int i = 1; i < 4; ++i)
        if (A[i][i] > A[best][best]) best = i;
    for (int k = 0; k < 4; ++k) vec[k] = V[k][best];
}

__global__ void __launch_bounds__(kMaxK)
rank_pool_kernel(const float *__restrict__ pose, const float *__restrict__ energy, int K, int keep,
                 float *__restrict__ sorted_pose, float *__restrict__ sorted_energy, float *__restrict__ pooled, int *__restrict__ order) {
    __shared__ float sp[kMaxK][9];
    __shared__ float se[kMaxK][2];
    __shared__ float srt[kMaxK][9];   // sorted pose
    __shared__ float sq[kMaxK][4];    // aligned quaternions of the kept poses
    __shared__ double sA[4][4];
    const int b = blockIdx.x, i = threadIdx.x;
    if (i < K) {
#pragma unroll
        for (int c = 0; c < 9; ++c) sp[i][c] = pose[((size_t)b * K + i) * 9 + c];
        se[i][0] = energy[((size_t)b * K + i) * 2 + 0];
        se[i][1] = energy[((size_t)b * K + i) * 2 + 1];
    }
    __syncthreads();
    if (i < K) {
        int rr = 0, rt = 0;   // stable descending ranks
        const float er = se[i][0], et = se[i][1];
        for (int j = 0; j < K; ++j) {
            rr += (se[j][0] > er) || (se[j][0] == er && j < i);
            rt += (se[j][1] > et) || (se[j][1] == et && j < i);
        }
#pragma unroll
        for (int c = 0; c < 6; ++c) srt[rr][c] = sp[i][c];
#pragma unroll
        for (int c = 6; c < 9; ++c) srt[rt][c] = sp[i][c];
        if (sorted_energy) {
            sorted_energy[((size_t)b * K + rr) * 2 + 0] = er;
            sorted_energy[((size_t)b * K + rt) * 2 + 1] = et;
        }
        if (order) {   // order[b, rank, 0 / 1] = index of the candidate at that rank by rotation / translation energy
            order[((size_t)b * K + rr) * 2 + 0] = i;
            order[((size_t)b * K + rt) * 2 + 1] = i;
        }
    }
    __syncthreads();
    if (i < K && sorted_pose) {
#pragma unroll
        for (int c = 0; c < 9; ++c) sorted_pose[((size_t)b * K + i) * 9 + c] = srt[i][c];
    }
    if (pooled == nullptr) return;
    if (i < keep) {
        float q[4];
        rot6d_to_quat(srt[i], q);
        const float sgn = q[0] > 0.f ? 1.0f : -1.0f;   // ((w > 0) - 0.5) * 2, utils/misc.py:243
#pragma unroll
        for (int c = 0; c < 4; ++c) sq[i][c] = sgn * q[c];
    }
    __syncthreads();
    if (i < 16) {
        const int r = i >> 2, c = i & 3;
        float acc = 0.f;
        const float wgt = 1.0f / (float)keep;   // weights = 1/num (misc.py:239), weight_sum = 1
        for (int k = 0; k < keep; ++k) acc += (sq[k][r] * sq[k][c]) * wgt;
        sA[r][c] = (double)acc;
    }
    __syncthreads();
    if (i == 0) {
        double A[4][4], v[4];
        for (int r = 0; r < 4; ++r)
            for (int c = 0; c < 4; ++c) A[r][c] = 0.5 * (sA[r][c] + sA[c][r]);
        top_eigenvector4(A, v);
        const double sgn = v[0] > 0 ? 1.0 : -1.0;
        const float qr = (float)(sgn * v[0]), qi = (float)(sgn * v[1]), qj = (float)(sgn * v[2]), qk = (float)(sgn * v[3]);
        const float two_s = 2.0f / (qr * qr + qi * qi + qj * qj + qk * qk);   // pytorch3d quaternion_to_matrix
        float t[3] = {0.f, 0.f, 0.f};
        for (int k = 0; k < keep; ++k) {
            t[0] += srt[k][6];
            t[1] += srt[k][7];
            t[2] += srt[k][8];
        }
        float *o = pooled + (size_t)b * 16;
        o[0] = 1 - two_s * (qj * qj + qk * qk); o[1] = two_s * (qi * qj - qk * qr); o[2] = two_s * (qi * qk + qj * qr); o[3] = t[0] / (float)keep;
        o[4] = two_s * (qi * qj + qk * qr); o[5] = 1 - two_s * (qi * qi + qk * qk); o[6] = two_s * (qj * qk - qi * qr); o[7] = t[1] / (float)keep;
        o[8] = two_s * (qi * qk - qj * qr); o[9] = two_s * (qj * qk + qi * qr); o[10] = 1 - two_s * (qi * qi + qj * qj); o[11] = t[2] / (float)keep;
        o[12] = 0.f; o[13] = 0.f; o[14] = 0.f; o[15] = 1.f;
    }
}

}  // namespace gpb

using namespace gpb;

extern "C" int gpb_rank_pool(const float *pose, const float *energy, int B, int K, int keep, float *sorted_pose,
                             float *sorted_energy, float *pooled_RT, int *order, void *stream) {
    GPB_REQUIRE(B >= 0 && K >= 1 && K <= kMaxK, "rank_pool: need B >= 0 and 1 <= K <= %d", kMaxK);
    GPB_REQUIRE(keep >= 1 && keep <= K, "rank_pool: need 1 <= keep <= K");
    if (B == 0) return GPB_OK;
    GPB_REQUIRE(pose && energy, "rank_pool: NULL buffer");
    rank_pool_kernel<<<B, kMaxK, 0, (cudaStream_t)stream>>>(pose, energy, K, keep, sorted_pose, sorted_energy, pooled_RT, order);
    GPB_LAUNCHED();
    return GPB_OK;
}
