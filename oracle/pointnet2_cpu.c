/*
 * ORACLE — test infrastructure, NOT product code.
 *
 * Plain-C restatement of the four forward native ops of the reference's `pointnet2_cuda`
 * extension (the only native boundary on the hot path).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this.
 *
 * Each function cites the reference file:line (under
 * networks/pts_encoder/pointnet2_utils/pointnet2/) it follows.  Distances are evaluated in the
 * exact form the reference compiles to on sm_100a (FMUL, FFMA, FFMA):
 *     d2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx))
 * so indices are bit-exact.  Build with -ffp-contract=off (see oracle/Makefile) so that the
 * compiler does not re-associate the explicit fmaf chain.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* src/cuda_utils.h:10-14  opt_n_threads */
static int opt_n_threads(int work_size) {
    const int pow_2 = (int)(log((double)work_size) / log(2.0));
    int t = 1 << pow_2;
    if (t > 1024) t = 1024;
    if (t < 1) t = 1;
    return t;
}

static inline float dist2(float ax, float ay, float az, float bx, float by, float bz) {
    /* (a - b) differences, then mul, fma, fma — order x, y, z (sampling_gpu.cu:133, ball_query_gpu.cu:33) */
    const float dx = ax - bx, dy = ay - by, dz = az - bz;
    return fmaf(dz, dz, fmaf(dy, dy, dx * dx));
}

/*
 * src/sampling_gpu.cu:93-209 furthest_point_sampling_kernel<block_size>, launched with
 * block_size = opt_n_threads(n) (:219).  Caller pre-fills temp with 1e10 and allocates idx
 * (pointnet2_utils.py:26-27).  Thread `tid` owns k = tid, tid+block, ...; per-thread running
 * (best, besti) start at (-1, 0) with strict '>' (:135-136); the block tree (__update :86-91)
 * takes the upper half's candidate only if strictly greater.
 */
int oracle_furthest_point_sampling(int b, int n, int m, const float *xyz, float *temp, int *idx) {
    if (m <= 0) return 1;
    const int bs = opt_n_threads(n);
    for (int bi = 0; bi < b; ++bi) {
        const float *p = xyz + (size_t)bi * n * 3;
        float *t = temp + (size_t)bi * n;
        int *out = idx + (size_t)bi * m;
        float *dists = (float *)malloc(sizeof(float) * bs);
        int *dists_i = (int *)malloc(sizeof(int) * bs);
        int old = 0;
        out[0] = old;
        for (int j = 1; j < m; ++j) {
            const float x1 = p[old * 3 + 0], y1 = p[old * 3 + 1], z1 = p[old * 3 + 2];
            for (int tid = 0; tid < bs; ++tid) {
                int besti = 0;
                float best = -1.0f;
                for (int k = tid; k < n; k += bs) {
                    const float d = dist2(p[k * 3 + 0], p[k * 3 + 1], p[k * 3 + 2], x1, y1, z1);
                    const float d2 = fminf(d, t[k]);
                    t[k] = d2;
                    besti = d2 > best ? k : besti;
                    best = d2 > best ? d2 : best;
                }
                dists[tid] = best;
                dists_i[tid] = besti;
            }
            for (int stride = bs / 2; stride >= 1; stride >>= 1) {
                for (int tid = 0; tid < stride; ++tid) {
                    const float v1 = dists[tid], v2 = dists[tid + stride];
                    const int i1 = dists_i[tid], i2 = dists_i[tid + stride];
                    dists[tid] = v1 > v2 ? v1 : v2; /* max(v1, v2) */
                    dists_i[tid] = v2 > v1 ? i2 : i1;
                }
            }
            old = dists_i[0];
            out[j] = old;
        }
        free(dists);
        free(dists_i);
    }
    return 1;
}

/* src/sampling_gpu.cu:8-24 gather_points_kernel_fast: out[b,c,j] = points[b,c,idx[b,j]] */
int oracle_gather_points(int b, int c, int n, int npoints, const float *points, const int *idx, float *out) {
    for (int bi = 0; bi < b; ++bi)
        for (int ci = 0; ci < c; ++ci)
            for (int j = 0; j < npoints; ++j)
                out[((size_t)bi * c + ci) * npoints + j] =
                    points[((size_t)bi * c + ci) * n + idx[(size_t)bi * npoints + j]];
    return 1;
}

/*
 * src/ball_query_gpu.cu:9-45 ball_query_kernel_fast.  idx is pre-zeroed by the caller
 * (pointnet2_utils.py:219); the first hit fills all nsample slots (:35-39); strict d2 < r2 (:34);
 * radius2 = radius*radius in fp32 (:23).  Centre minus point (:33).
 */
int oracle_ball_query(int b, int n, int m, float radius, int nsample, const float *new_xyz,
                      const float *xyz, int *idx) {
    const float radius2 = radius * radius;
    for (int bi = 0; bi < b; ++bi) {
        for (int pi = 0; pi < m; ++pi) {
            const float *c = new_xyz + ((size_t)bi * m + pi) * 3;
            const float *p = xyz + (size_t)bi * n * 3;
            int *o = idx + ((size_t)bi * m + pi) * nsample;
            int cnt = 0;
            for (int k = 0; k < n; ++k) {
                const float d2 = dist2(c[0], c[1], c[2], p[k * 3 + 0], p[k * 3 + 1], p[k * 3 + 2]);
                if (d2 < radius2) {
                    if (cnt == 0)
                        for (int l = 0; l < nsample; ++l) o[l] = k;
                    o[cnt] = k;
                    ++cnt;
                    if (cnt >= nsample) break;
                }
            }
        }
    }
    return 1;
}

/* src/group_points_gpu.cu:47-66 group_points_kernel_fast: out[b,c,i,s] = points[b,c,idx[b,i,s]] */
int oracle_group_points(int b, int c, int n, int npoints, int nsample, const float *points,
                        const int *idx, float *out) {
    for (int bi = 0; bi < b; ++bi)
        for (int ci = 0; ci < c; ++ci) {
            const float *src = points + ((size_t)bi * c + ci) * n;
            float *dst = out + ((size_t)bi * c + ci) * npoints * nsample;
            const int *ii = idx + (size_t)bi * npoints * nsample;
            for (int q = 0; q < npoints * nsample; ++q) dst[q] = src[ii[q]];
        }
    return 1;
}
