// Micro-benchmark (development aid, not product code): how fast can a B200 apply the annealing kernels' neighbour updates?
// Access pattern of the lockstep kernels: a warp owns a private matrix f[rows][32 lanes] of fp64 (256 B per row) and adds a
// value to a pseudo-random row, only on the lanes that "accepted" (probability a per lane).
//   mode 0: predicated red.global.add.f64 (fire and forget, performed at L2)
//   mode 1: predicated ld.global.cg / add / st.global.cg, 8 rows in flight per lane
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_update ubench_update.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cuda_runtime.h>

__device__ __forceinline__ void red_if(double *p, double v, bool pred) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\t@q red.global.add.f64 [%0], %1;\n\t}" ::"l"(p), "d"(v), "r"((int)pred) : "memory");
}

template <int MODE>
__global__ void __launch_bounds__(256) k_update(double *f, long long rows, int iters, unsigned thresh, unsigned long long *count) {
    const int lane = threadIdx.x & 31;
    const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    double *base = f + warp * rows * 32 + lane;
    unsigned long long s = 0x9E3779B97F4A7C15ull * (warp + 1);           // warp-uniform row stream
    unsigned long long t = 0xD1B54A32D192ED03ull * (warp * 32 + lane + 1);  // per-lane accept stream
    unsigned long long done = 0;
    for (int it = 0; it < iters; ++it) {
        long long j[8];
        bool p[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            s = s * 6364136223846793005ull + 1442695040888963407ull;
            j[q] = (long long)((s >> 33) & (unsigned long long)(rows - 1));  // rows is a power of two
            t = t * 6364136223846793005ull + 1442695040888963407ull;
            p[q] = (unsigned)(t >> 32) < thresh;
            done += p[q];
        }
        if (MODE == 0) {
#pragma unroll
            for (int q = 0; q < 8; ++q) red_if(base + j[q] * 32, 1.0, p[q]);
        } else {
            double x[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) x[q] = p[q] ? __ldcg(base + j[q] * 32) : 0.0;
#pragma unroll
            for (int q = 0; q < 8; ++q)
                if (p[q]) __stcg(base + j[q] * 32, x[q] + 1.0);
        }
    }
    for (int off = 16; off; off >>= 1) done += __shfl_xor_sync(0xffffffffu, done, off);
    if (lane == 0) atomicAdd(count, done);
}

int main() {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    const long long rows = 32768;  // 8 MB per warp: the whole set is far larger than L2
    const int max_wps = 48;
    const size_t bytes = (size_t)sms * max_wps * rows * 32 * 8;
    double *f;
    if (cudaMalloc(&f, bytes) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    cudaMemset(f, 0, bytes);
    unsigned long long *cnt;
    cudaMalloc(&cnt, 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    printf("# %s, %d SMs, %.1f GB of fields\n", prop.name, sms, bytes / 1e9);
    printf("mode warps/SM accept lane_updates/s sectors/s(est) GB/s(64B per touched sector)\n");
    const double accs[] = {1.0, 0.5, 0.2, 0.05};
    for (int mode = 0; mode < 2; ++mode)
        for (int wps : {8, 16, 32, 48})
            for (double a : accs) {
                const unsigned thresh = a >= 1.0 ? 0xffffffffu : (unsigned)(a * 4294967296.0);
                const int iters = 2000;
                const int blocks = sms * wps / 8;
                cudaMemset(cnt, 0, 8);
                for (int rep = 0; rep < 2; ++rep) {
                    if (rep == 1) { cudaMemset(cnt, 0, 8); cudaEventRecord(e0); }
                    if (mode == 0) k_update<0><<<blocks, 256>>>(f, rows, iters, thresh, cnt);
                    else k_update<1><<<blocks, 256>>>(f, rows, iters, thresh, cnt);
                }
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                float ms;
                cudaEventElapsedTime(&ms, e0, e1);
                unsigned long long c;
                cudaMemcpy(&c, cnt, 8, cudaMemcpyDeviceToHost);
                const double rows_total = (double)blocks * 8 * iters * 8;
                const double psec = 1.0 - pow(1.0 - a, 4.0);   // a 32 B sector holds 4 lanes
                const double sectors = rows_total * 8 * psec;
                printf("%s %2d %.2f %.3e %.3e %.0f\n", mode == 0 ? "red " : "ldst", wps, a, c / (ms * 1e-3), sectors / (ms * 1e-3),
                       sectors * 64 / (ms * 1e-3) / 1e9);
                fflush(stdout);
            }
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) { printf("error %s\n", cudaGetErrorString(err)); return 1; }
    return 0;
}
