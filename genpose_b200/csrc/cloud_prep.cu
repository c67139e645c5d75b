// Point-cloud preparation of one RGB-D frame on the device: for every detected instance the
// body of the instance loop of detect_mrcnn_genpose (runners/evaluation_single.py:168-216) —
//     crop_resize_by_warp_affine(coord_2d | mask | depth, INTER_NEAREST)  ->  depth_to_pcl  ->  / 1000  ->  sample_points
// — without ever materialising the three 256 x 256 crops: a destination pixel's source pixel is a pure function of the
// crop's affine matrix (OpenCV's fixed-point nearest-neighbour rule, restated below), so the crop of the coordinate map
// IS that source pixel, and mask / depth are gathered from the full-resolution frame on the fly.
//
// One CTA per instance, 1024 threads, three passes over the 65,536 crop pixels in raster order (the order
// depth_to_pcl's boolean indexing keeps):
//   1. validity of every crop pixel (mask & depth > 0 at the source pixel, source inside the image): warp ballots -> a
//      bitmap in shared memory + the count per crop row;
//   2. exclusive scan of the 256 row counts; a sweep over the bitmap writes the crop index (16 bits) of every valid
//      pixel, in raster order, into SHARED memory (128 KiB): the compacted list never touches HBM;
//   3. the 1024 output points: k-th valid pixel for k = j mod n (n < 1024: np.tile + remainder), k = j (n == 1024) or
//      k = ids[j] (n > 1024: the reference's np.random.permutation(n)[:1024], passed in for parity, or a keyed
//      Feistel permutation evaluated in place in throughput mode), back-projected in the reference's float32
//      operation order  x = ((u - cx) * d) / fx,  y = ((v - cy) * d) / fy,  z = d,  then / 1000.
// HBM traffic per instance: the gathered depth (2 B) and mask (1 B) source pixels of the crop window, once, + 1024 depth
// re-reads + 12 KiB of points out: gather-latency bound integer work (one CTA of 32 warps per instance), no tensor cores
// (profiles/r1i_cloud_prep.txt).
//
// OpenCV arithmetic restated (modules/imgproc/src/imgwarp.cpp, cv::warpAffine / WarpAffineInvoker, INTER_NEAREST,
// BORDER_CONSTANT 0; the reference pins opencv-python 4.2.0.32, the oracle is pinned against 4.13.0):
//   M <- inverse of the forward 2x3 matrix, in double, without fused multiply-adds;
//   adelta[x] = cvRound(M0 * x * 1024), bdelta[x] = cvRound(M3 * x * 1024)         (cvRound = round half to even)
//   X0[y] = cvRound((M1 * y + M2) * 1024) + 512, Y0[y] = cvRound((M4 * y + M5) * 1024) + 512
//   X = (X0[y] + adelta[x]) >> 10, Y = (Y0[y] + bdelta[x]) >> 10   (saturated to int16), border value 0 outside the image.
#include "common.cuh"

namespace gpb {

constexpr int kPrepRoi = 256;            // cfg.img_size (configs/config.py:78)
constexpr int kPrepPoints = 1024;        // cfg.num_points (configs/config.py:24)
constexpr int kPrepThreads = 1024;       // 32 warps: the sweeps are gather-latency bound, so as many loads in flight as one CTA can hold
constexpr size_t kPrepSmemBytes = (size_t)kPrepRoi * kPrepRoi * sizeof(unsigned short);

struct PrepParams {
    const unsigned short *depth;   // [H,W] millimetres
    const unsigned char *masks;    // instance i, pixel p: masks[p * mask_pixel_stride + i * mask_inst_stride], non-zero = inside
    long long mask_pixel_stride, mask_inst_stride;
    int H, W, n_inst;
    const double *trans;           // [n_inst][6] forward affine matrices (what the reference passes to cv2.warpAffine)
    float cx, cy, fx, fy;          // camera intrinsics as float32 (evaluation_single.py:50,54)
    const int *subset_ids;         // [n_inst][1024] or null
    unsigned key0, key1;
    float *pts;                    // [n_inst][1024][3]
    int *n_valid;                  // [n_inst]
};

__device__ __forceinline__ unsigned prep_mix32(unsigned x) {
    x ^= x >> 16;
    x *= 0x7FEB352Du;
    x ^= x >> 15;
    x *= 0x846CA68Bu;
    x ^= x >> 16;
    return x;
}

// j-th value of a keyed pseudo-random permutation of [0, n): 4-round balanced Feistel network + cycle walking
// (oracle/cloud_prep_oracle.py::feistel_permutation_prefix is the same function)
__device__ __forceinline__ int prep_feistel(int j, int n, unsigned key0, unsigned key1) {
    int bits = 32 - __clz(max(n, 2) - 1);
    bits = max(bits, 2);
    bits += bits & 1;
    const int half = bits >> 1;
    const unsigned mask = (1u << half) - 1u;
    unsigned v = (unsigned)j;
    do {
        unsigned lo = v & mask, hi = v >> half;
#pragma unroll
        for (int rnd = 0; rnd < 4; ++rnd) {
            const unsigned f = (prep_mix32(lo ^ (key0 + 0x9E3779B9u * (unsigned)rnd)) ^ key1) & mask;
            const unsigned nlo = hi ^ f;
            hi = lo;
            lo = nlo;
        }
        v = (hi << half) | lo;
    } while (v >= (unsigned)n);
    return (int)v;
}

__device__ __forceinline__ int prep_round_to_int(double v) {   // cv::saturate_cast<int>(double) = cvRound, saturated
    v = fmin(fmax(v, -2147483648.0), 2147483647.0);
    return (int)__double2ll_rn(v);
}

__global__ void __launch_bounds__(kPrepThreads)
cloud_prep_kernel(PrepParams p) {
    extern __shared__ __align__(16) unsigned short s_ids[];          // crop index (y << 8 | x) of the k-th valid pixel
    __shared__ int s_adelta[kPrepRoi], s_bdelta[kPrepRoi], s_x0[kPrepRoi], s_y0[kPrepRoi];
    __shared__ int s_row[kPrepRoi + 1];
    __shared__ unsigned s_bits[kPrepRoi][kPrepRoi / 32];             // validity of every crop pixel, one bit each (8 KiB)
    const int inst = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    if (tid < kPrepRoi) {   // inverse affine + fixed-point tables: thread t owns column x = t and row y = t (no FMA contraction: OpenCV's scalar code)
        const double *T = p.trans + (size_t)inst * 6;
        double m0 = T[0], m1 = T[1], m2 = T[2], m3 = T[3], m4 = T[4], m5 = T[5];
        double D = __dsub_rn(__dmul_rn(m0, m4), __dmul_rn(m1, m3));
        D = D != 0.0 ? __ddiv_rn(1.0, D) : 0.0;
        const double a11 = __dmul_rn(m4, D), a22 = __dmul_rn(m0, D);
        m0 = a11;
        m1 = __dmul_rn(m1, -D);
        m3 = __dmul_rn(m3, -D);
        m4 = a22;
        const double b1 = __dsub_rn(__dmul_rn(-m0, m2), __dmul_rn(m1, m5));
        const double b2 = __dsub_rn(__dmul_rn(-m3, m2), __dmul_rn(m4, m5));
        m2 = b1;
        m5 = b2;
        const double t = (double)tid;
        s_adelta[tid] = prep_round_to_int(__dmul_rn(__dmul_rn(m0, t), 1024.0));
        s_bdelta[tid] = prep_round_to_int(__dmul_rn(__dmul_rn(m3, t), 1024.0));
        s_x0[tid] = prep_round_to_int(__dmul_rn(__dadd_rn(__dmul_rn(m1, t), m2), 1024.0)) + 512;
        s_y0[tid] = prep_round_to_int(__dmul_rn(__dadd_rn(__dmul_rn(m4, t), m5), 1024.0)) + 512;
    }
    __syncthreads();

    const unsigned char *mask = p.masks + (size_t)inst * p.mask_inst_stride;
    // source pixel of crop pixel (x, y): linear index into the frame, or -1 outside (border value 0)
    auto source = [&](int x, int y) -> int {
        int X = (s_x0[y] + s_adelta[x]) >> 10, Y = (s_y0[y] + s_bdelta[x]) >> 10;
        X = min(max(X, -32768), 32767);
        Y = min(max(Y, -32768), 32767);
        return ((unsigned)X < (unsigned)p.W && (unsigned)Y < (unsigned)p.H) ? Y * p.W + X : -1;
    };
    auto is_valid = [&](int src) -> bool {   // both gathers are issued before either is tested (one latency, not two)
        const int s0 = max(src, 0);
        const unsigned short d = __ldg(p.depth + s0);
        const unsigned char m = __ldg(mask + (size_t)s0 * p.mask_pixel_stride);
        return src >= 0 && d > 0 && m != 0;
    };

    // ---- pass 1: validity bitmap (kept in shared memory for pass 2) and valid pixels per crop row (warp w: rows w, w + 32, ...) ----
    for (int y = warp; y < kPrepRoi; y += kPrepThreads / 32) {
        int cnt = 0;
#pragma unroll
        for (int seg = 0; seg < kPrepRoi / 32; ++seg) {
            const unsigned bal = __ballot_sync(0xffffffffu, is_valid(source(seg * 32 + lane, y)));
            if (lane == seg) s_bits[y][seg] = bal;
            cnt += __popc(bal);
        }
        if (lane == 0) s_row[y] = cnt;
    }
    __syncthreads();
    if (warp == 0) {   // exclusive scan of the 256 row counts: 8 consecutive rows per lane
        int loc[8], sum = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            loc[i] = sum;
            sum += s_row[lane * 8 + i];
        }
        int incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        const int base = incl - sum;
#pragma unroll
        for (int i = 0; i < 8; ++i) s_row[lane * 8 + i] = base + loc[i];
        if (lane == 31) s_row[kPrepRoi] = incl;
    }
    __syncthreads();
    const int n = s_row[kPrepRoi];

    // ---- pass 2: compacted list of valid crop pixels, raster order, in shared memory ----
    for (int y = warp; y < kPrepRoi; y += kPrepThreads / 32) {
        int pos = s_row[y];
#pragma unroll
        for (int seg = 0; seg < kPrepRoi / 32; ++seg) {
            const int x = seg * 32 + lane;
            const unsigned bal = s_bits[y][seg];
            if ((bal >> lane) & 1u) s_ids[pos + __popc(bal & ((1u << lane) - 1u))] = (unsigned short)((y << 8) | x);
            pos += __popc(bal);
        }
    }
    __syncthreads();

    // ---- pass 3: resample to 1024 points and back-project (evaluation_single.py:107-133, :211) ----
    if (tid == 0) p.n_valid[inst] = n;
    float *out = p.pts + (size_t)inst * kPrepPoints * 3;
    for (int j = tid; j < kPrepPoints; j += kPrepThreads) {
        float X = 0.f, Y = 0.f, Z = 0.f;
        if (n > 1) {                                           // n <= 1: the reference skips the instance (:201-209)
            int k;
            if (n < kPrepPoints) k = j % n;                    // np.tile(pcl, (1024 // n, 1)) ++ pcl[:1024 % n]
            else if (n == kPrepPoints) k = j;
            else if (p.subset_ids) k = min(max(p.subset_ids[(size_t)inst * kPrepPoints + j], 0), n - 1);
            else k = prep_feistel(j, n, p.key0 + 0x85EBCA6Bu * (unsigned)inst, p.key1);
            const int pix = s_ids[k];
            const int src = source(pix & 255, pix >> 8);
            const int sy = src / p.W, sx = src - sy * p.W;
            const float d = (float)__ldg(p.depth + src);
            X = __fdiv_rn(__fdiv_rn(__fmul_rn(__fsub_rn((float)sx, p.cx), d), p.fx), 1000.0f);
            Y = __fdiv_rn(__fdiv_rn(__fmul_rn(__fsub_rn((float)sy, p.cy), d), p.fy), 1000.0f);
            Z = __fdiv_rn(d, 1000.0f);
        }
        out[j * 3 + 0] = X;
        out[j * 3 + 1] = Y;
        out[j * 3 + 2] = Z;
    }
}

}  // namespace gpb

using namespace gpb;

extern "C" int gpb_prepare_clouds(const unsigned short *depth, const unsigned char *masks, long long mask_pixel_stride,
                                  long long mask_inst_stride, int H, int W, int n_inst, const double *trans, const float *intrinsics,
                                  const int *subset_ids, uint64_t seed, float *pts, int *n_valid, void *stream) {
    GPB_REQUIRE(n_inst >= 0 && H > 0 && W > 0 && H <= 32767 && W <= 32767 && (long long)H * W < (1ll << 31),
                "prepare_clouds: need n_inst >= 0 and 0 < H, W <= 32767");
    if (n_inst == 0) return GPB_OK;
    GPB_REQUIRE(depth && masks && trans && intrinsics && pts && n_valid, "prepare_clouds: NULL buffer");
    GPB_REQUIRE(mask_pixel_stride > 0 && mask_inst_stride > 0, "prepare_clouds: mask strides must be positive");
    PrepParams p{};
    p.depth = depth; p.masks = masks; p.mask_pixel_stride = mask_pixel_stride; p.mask_inst_stride = mask_inst_stride;
    p.H = H; p.W = W; p.n_inst = n_inst; p.trans = trans;
    p.cx = intrinsics[0]; p.cy = intrinsics[1]; p.fx = intrinsics[2]; p.fy = intrinsics[3];
    p.subset_ids = subset_ids; p.key0 = (unsigned)seed; p.key1 = (unsigned)(seed >> 32);
    p.pts = pts; p.n_valid = n_valid;
    GPB_CUDA(cudaFuncSetAttribute(cloud_prep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPrepSmemBytes));
    cloud_prep_kernel<<<n_inst, kPrepThreads, kPrepSmemBytes, (cudaStream_t)stream>>>(p);
    GPB_LAUNCHED();
    return GPB_OK;
}
