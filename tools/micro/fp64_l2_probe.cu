// Micro-probe (experiment, not part of the library): what does the ODE sampler's pre-accumulation block cost on this GPU?
//   (a) 12 x ld.global.cg.v4 issued back to back from L2-resident data   (b) 30 F2F.F64.F32 + 30 dependent-by-5 DFMA
//   (c) 6 dynamically indexed __constant__ double loads                   — each timed with clock64 by 4 warps of a 320-thread CTA on 100 CTAs
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_l2_probe fp64_l2_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
__constant__ double kTab[7][6];
__global__ void probe(const float *__restrict__ K, size_t stride, int st, long long *out, double *sink) {
    const int tid = threadIdx.x;
    if (tid >= 128) return;
    const float *src = K + ((size_t)blockIdx.x * 128 + tid) * 12;
    float kf[6][8];
    long long t0 = clock64();
#pragma unroll
    for (int j = 0; j < 6; ++j)
#pragma unroll
        for (int w = 0; w < 2; ++w)
            asm volatile("ld.global.cg.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(kf[j][4 * w]), "=f"(kf[j][4 * w + 1]), "=f"(kf[j][4 * w + 2]), "=f"(kf[j][4 * w + 3]) : "l"(src + j * stride + 4 * w) : "memory");
#pragma unroll
    for (int j = 0; j < 6; ++j)
        asm volatile("" : "+f"(kf[j][0]), "+f"(kf[j][1]), "+f"(kf[j][2]), "+f"(kf[j][3]), "+f"(kf[j][4]), "+f"(kf[j][5]), "+f"(kf[j][6]), "+f"(kf[j][7]));
    long long t1 = clock64();
    double a[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) a[j] = st < 5 ? kTab[st + 1][j < 5 ? j : 0] : kTab[0][j];
    asm volatile("" : "+d"(a[0]), "+d"(a[1]), "+d"(a[2]), "+d"(a[3]), "+d"(a[4]), "+d"(a[5]));
    long long t2 = clock64();
    double acc[5] = {0, 0, 0, 0, 0};
#pragma unroll
    for (int j = 0; j < 6; ++j)
#pragma unroll
        for (int i = 0; i < 5; ++i) acc[i] += (j < st) ? (double)kf[j][i] * a[j] : 0.0;
    asm volatile("" : "+d"(acc[0]), "+d"(acc[1]), "+d"(acc[2]), "+d"(acc[3]), "+d"(acc[4]));
    long long t3 = clock64();
    // fp32 reference of the same shape
    float accf[5] = {0, 0, 0, 0, 0};
#pragma unroll
    for (int j = 0; j < 6; ++j)
#pragma unroll
        for (int i = 0; i < 5; ++i) accf[i] += (j < st) ? kf[j][i + 3] * (float)j : 0.f;
    asm volatile("" : "+f"(accf[0]), "+f"(accf[1]), "+f"(accf[2]), "+f"(accf[3]), "+f"(accf[4]));
    long long t4 = clock64();
    if (tid == 0 && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t1; out[2] = t3 - t2; out[3] = t4 - t3; }
    sink[(size_t)blockIdx.x * 128 + tid] = acc[0] + acc[1] + acc[2] + acc[3] + acc[4] + accf[0] + accf[1] + accf[2] + accf[3] + accf[4];
}
int main() {
    const int R = 12800; float *K; double *sink; long long *out;
    cudaMalloc(&K, (size_t)7 * R * 12 * 4); cudaMemset(K, 0, (size_t)7 * R * 12 * 4);
    cudaMalloc(&sink, R * 8); cudaMalloc(&out, 64);
    double tab[7][6]; for (int i = 0; i < 42; ++i) (&tab[0][0])[i] = 0.1 * i; cudaMemcpyToSymbol(kTab, tab, sizeof(tab));
    for (int rep = 0; rep < 3; ++rep) {
        probe<<<100, 320>>>(K, (size_t)R * 12, 4, out, sink);
        cudaDeviceSynchronize();
        long long h[4]; cudaMemcpy(h, out, 32, cudaMemcpyDeviceToHost);
        printf("rep %d: 12 x LDG.128 (L2) %lld cycles | 6 indexed constant loads %lld | 30 F2F + 30 DFMA %lld | 30 FFMA %lld   (%s)\n", rep, h[0], h[1], h[2], h[3], cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
