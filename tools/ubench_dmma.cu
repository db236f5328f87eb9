// tools/ubench_dmma.cu -- development aid: fp64 tensor-core (DMMA.8x8x4) and DFMA throughput of the device; the measured DMMA
// figure is the roofline denominator of k_anneal_dense.  nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1688(double (&c)[4], const double (&a)[4], const double (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void dmma16816(double (&c)[4], const double (&a)[8], const double (&b)[4]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}
template <int SHAPE, int TILES>
__global__ void k(double *out, int iters, double a0, double b0) {
    double c[TILES][4];
    for (int t = 0; t < TILES; ++t) for (int i = 0; i < 4; ++i) c[t][i] = 0;
    double a[8], b[4];
    for (int i = 0; i < 8; ++i) a[i] = a0 + threadIdx.x * 1e-3 + i;
    for (int i = 0; i < 4; ++i) b[i] = b0 + threadIdx.x * 1e-4 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int t = 0; t < TILES; ++t) {
            if (SHAPE == 0) dmma884(c[t][0], c[t][1], a[0], b[0]);
            else if (SHAPE == 1) { double aa[4] = {a[0], a[1], a[2], a[3]}; double bb[2] = {b[0], b[1]}; dmma1688(c[t], aa, bb); }
            else if (SHAPE == 2) dmma16816(c[t], a, b);
            else { c[t][0] = fma(a[0], b[0], c[t][0]); c[t][1] = fma(a[1], b[1], c[t][1]); c[t][2] = fma(a[2], b[2], c[t][2]); c[t][3] = fma(a[3], b[3], c[t][3]); }
        }
    }
    double s = 0;
    for (int t = 0; t < TILES; ++t) for (int i = 0; i < 4; ++i) s += c[t][i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int SHAPE>
void run(const char *name, double macs_per_inst_per_warp, int warps) {
    double *out; cudaMalloc(&out, 148 * 4 * 1024 * 8);
    const int iters = 20000, TILES = 8;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        k<SHAPE, TILES><<<148 * 2, warps * 32>>>(out, iters, 1.0, 2.0);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
    }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double macs = (double)148 * 2 * warps * iters * TILES * macs_per_inst_per_warp;
    printf("%-10s warps/CTA %2d: %.3f ms  %.2f TFLOP/s  (%s)\n", name, warps, ms, 2 * macs / ms / 1e9, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
}
int main() {
    for (int w : {4, 8, 16}) {
        run<0>("m8n8k4", 256, w); run<1>("m16n8k8", 1024, w); run<2>("m16n8k16", 2048, w); run<3>("dfma", 128, w);
    }
    return 0;
}
