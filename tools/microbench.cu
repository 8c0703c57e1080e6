// FP64 microbenchmarks for the design decisions of DESIGN.md §3 (sm_100a):
//   DFMA peak / dependent latency, DMMA (mma.sync f64) throughput per shape and per number of
//   independent accumulators, broadcast LDS.64 / LDS.128 feeding DFMAs.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench tools/microbench.cu
#include <cstdio>
#include <cuda_runtime.h>

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

template <int CH>
__global__ void k_dfma(double *out, int iters, double seed) {
    double a[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) a[i] = seed + i;
    const double m = 1.0000001, c = 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 16; ++u)
#pragma unroll
            for (int i = 0; i < CH; ++i) a[i] = fma(a[i], m, c);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += a[i];
    out[(long long)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1688(double *c, const double *a, const double *b) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void dmma16816(double *c, const double *a, const double *b) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                   "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

template <int CH>
__global__ void k_dmma884(double *out, int iters, double seed) {
    double c0[CH], c1[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) { c0[i] = seed + i; c1[i] = seed - i; }
    const double a = 1e-3 * (threadIdx.x & 3), b = 1e-3 * (threadIdx.x >> 2);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int i = 0; i < CH; ++i) dmma884(c0[i], c1[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += c0[i] + c1[i];
    out[(long long)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int CH>
__global__ void k_dmma1688(double *out, int iters, double seed) {
    double c[CH][4];
#pragma unroll
    for (int i = 0; i < CH; ++i) { c[i][0] = seed + i; c[i][1] = seed - i; c[i][2] = seed; c[i][3] = -seed; }
    double a[4], b[2];
    for (int i = 0; i < 4; ++i) a[i] = 1e-3 * ((threadIdx.x + i) & 3);
    for (int i = 0; i < 2; ++i) b[i] = 1e-3 * ((threadIdx.x >> 2) + i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int i = 0; i < CH; ++i) dmma1688(c[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[(long long)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int CH>
__global__ void k_dmma16816(double *out, int iters, double seed) {
    double c[CH][4];
#pragma unroll
    for (int i = 0; i < CH; ++i) { c[i][0] = seed + i; c[i][1] = seed - i; c[i][2] = seed; c[i][3] = -seed; }
    double a[8], b[4];
    for (int i = 0; i < 8; ++i) a[i] = 1e-3 * ((threadIdx.x + i) & 3);
    for (int i = 0; i < 4; ++i) b[i] = 1e-3 * ((threadIdx.x >> 2) + i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int i = 0; i < CH; ++i) dmma16816(c[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[(long long)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// broadcast LDS feeding DFMAs: W = 8 (LDS.64) or 16 (LDS.128) bytes per load, F DFMAs per loaded double
template <int VEC, int F>
__global__ void k_lds_dfma(double *out, int iters, double seed) {
    __shared__ __align__(16) double tab[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) tab[i] = 1.0 + 1e-9 * i;
    __syncthreads();
    double acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = seed + i + threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 64; ++u) {
            if (VEC == 2) {
                const double2 p = reinterpret_cast<const double2 *>(tab)[(it + u) & 511];
#pragma unroll
                for (int f = 0; f < F; ++f) { acc[(2 * f) & 7] = fma(p.x, acc[(2 * f) & 7], 1e-9); acc[(2 * f + 1) & 7] = fma(p.y, acc[(2 * f + 1) & 7], 1e-9); }
            } else {
                const double p = tab[(it + u) & 1023];
#pragma unroll
                for (int f = 0; f < F; ++f) acc[f & 7] = fma(p, acc[f & 7], 1e-9);
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += acc[i];
    out[(long long)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float time_ms(F launch) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    launch(); launch();
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(a);
        launch();
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp prop;
    CHECK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    printf("device %s, %d SMs\n", prop.name, sms);
    double *out;
    CHECK(cudaMalloc(&out, sizeof(double) * sms * 16 * 1024));
    const int iters = 2048;
#define RUN_DFMA(CH, WARPS) { const int blocks = sms * (WARPS / 4 > 0 ? WARPS / 4 : 1), tpb = WARPS >= 4 ? 128 : 32 * WARPS; \
        float ms = time_ms([&] { k_dfma<CH><<<blocks, tpb>>>(out, iters, 1.0); }); \
        double fl = 2.0 * CH * 16.0 * iters * blocks * tpb; \
        printf("dfma chains=%d warps/SM=%2d: %8.3f ms  %7.2f TFLOP/s  (%.2f cycles/dfma/warp @1.965GHz per SMSP)\n", CH, WARPS, ms, fl / ms * 1e-9, \
               ms * 1e-3 * 1.965e9 / (CH * 16.0 * iters * (WARPS >= 4 ? WARPS / 4.0 : 1.0))); }
    RUN_DFMA(1, 4) RUN_DFMA(2, 4) RUN_DFMA(4, 4) RUN_DFMA(8, 4)
    RUN_DFMA(1, 8) RUN_DFMA(2, 8) RUN_DFMA(4, 8) RUN_DFMA(8, 8)
    RUN_DFMA(1, 16) RUN_DFMA(2, 16) RUN_DFMA(4, 16) RUN_DFMA(8, 16) RUN_DFMA(8, 32)
#define RUN_MMA(KERN, NAME, FMAS, CH, WARPS) { const int blocks = sms * (WARPS / 4), tpb = 128; \
        float ms = time_ms([&] { KERN<CH><<<blocks, tpb>>>(out, iters, 1.0); }); \
        double fl = 2.0 * FMAS * CH * 8.0 * iters * blocks * (tpb / 32); \
        printf("%s acc=%d warps/SM=%2d: %8.3f ms  %7.2f TFLOP/s\n", NAME, CH, WARPS, ms, fl / ms * 1e-9); }
    RUN_MMA(k_dmma884, "dmma m8n8k4  ", 256.0, 1, 4) RUN_MMA(k_dmma884, "dmma m8n8k4  ", 256.0, 2, 4) RUN_MMA(k_dmma884, "dmma m8n8k4  ", 256.0, 4, 4) RUN_MMA(k_dmma884, "dmma m8n8k4  ", 256.0, 8, 4)
    RUN_MMA(k_dmma884, "dmma m8n8k4  ", 256.0, 4, 8) RUN_MMA(k_dmma884, "dmma m8n8k4  ", 256.0, 8, 8) RUN_MMA(k_dmma884, "dmma m8n8k4  ", 256.0, 8, 16)
    RUN_MMA(k_dmma1688, "dmma m16n8k8 ", 1024.0, 1, 4) RUN_MMA(k_dmma1688, "dmma m16n8k8 ", 1024.0, 2, 4) RUN_MMA(k_dmma1688, "dmma m16n8k8 ", 1024.0, 4, 4) RUN_MMA(k_dmma1688, "dmma m16n8k8 ", 1024.0, 4, 8) RUN_MMA(k_dmma1688, "dmma m16n8k8 ", 1024.0, 4, 16)
    RUN_MMA(k_dmma16816, "dmma m16n8k16", 2048.0, 1, 4) RUN_MMA(k_dmma16816, "dmma m16n8k16", 2048.0, 2, 4) RUN_MMA(k_dmma16816, "dmma m16n8k16", 2048.0, 4, 4) RUN_MMA(k_dmma16816, "dmma m16n8k16", 2048.0, 4, 8)
#define RUN_LDS(VEC, F, WARPS) { const int blocks = sms * (WARPS / 4), tpb = 128; \
        float ms = time_ms([&] { k_lds_dfma<VEC, F><<<blocks, tpb>>>(out, iters / 4, 1.0); }); \
        double fl = 2.0 * VEC * F * 64.0 * (iters / 4) * blocks * tpb; \
        printf("lds.%d broadcast + %d dfma per double, warps/SM=%2d: %8.3f ms  %7.2f TFLOP/s\n", VEC * 64, F, WARPS, ms, fl / ms * 1e-9); }
    RUN_LDS(1, 1, 8) RUN_LDS(1, 2, 8) RUN_LDS(1, 4, 8) RUN_LDS(1, 8, 8)
    RUN_LDS(2, 1, 8) RUN_LDS(2, 2, 8) RUN_LDS(2, 4, 8) RUN_LDS(2, 8, 8)
    RUN_LDS(1, 2, 16) RUN_LDS(2, 2, 16) RUN_LDS(2, 4, 16)
    cudaFree(out);
    return 0;
}
