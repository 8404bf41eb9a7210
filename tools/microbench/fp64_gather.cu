// Microbenchmarks that size the LBA kernel: (1) DFMA issue rate per SM, (2) throughput of per-lane gathers of 32-byte
// records (32 distinct 128-byte lines per warp instruction) from an L2-resident array, as LDG.128 x2 vs LDG.256.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void dfma(double* o, int iters, double a, double b) {
    double x[8];
    for (int i = 0; i < 8; ++i) x[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = fma(x[i], a, b);
    double s = 0; for (int i = 0; i < 8; ++i) s += x[i];
    o[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE>
__global__ void gather(const double* __restrict__ E, const int* __restrict__ idx, double* o, int nPerThread, int stride) {
    double s = 0;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    for (int k = 0; k < nPerThread; k += 4) {
        int e[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) e[u] = idx[(size_t)(k + u) * stride + t];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const double* p = E + 4 * (size_t)e[u];
            if (MODE == 0) { const double2 a = ((const double2*)p)[0], b = ((const double2*)p)[1]; s += a.x + a.y + b.x + b.y; }
            else { double a, b, c, d; asm volatile("ld.global.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p)); s += a + b + c + d; }
        }
    }
    o[t] = s;
}
int main() {
    int nsm = 148; cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1); float ms;
    double* o; cudaMalloc(&o, 8 * 148 * 1024 * 4);
    for (int nt : {128, 256, 512, 1024}) {
        const int iters = 20000;
        dfma<<<nsm, nt>>>(o, iters, 1.0000001, 1e-9); cudaEventRecord(e0); dfma<<<nsm, nt>>>(o, iters, 1.0000001, 1e-9); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        printf("dfma threads/SM %4d: %.1f DFMA/clk/SM (at 1.965 GHz)\n", nt, (double)nt * iters * 8 / (ms * 1e-3 * 1.965e9));
    }
    const int nRec = 1 << 16;   // 2 MB of 32-byte records: L2 resident, larger than L1
    double* E; cudaMalloc(&E, 32 * nRec); cudaMemset(E, 0, 32 * nRec);
    const int nt = 512, per = 256, total = nsm * nt;
    int* h = new int[(size_t)per * total]; unsigned r = 12345;
    for (size_t i = 0; i < (size_t)per * total; ++i) { r = r * 1664525u + 1013904223u; h[i] = (r >> 8) % nRec; }
    int* idx; cudaMalloc(&idx, 4 * (size_t)per * total); cudaMemcpy(idx, h, 4 * (size_t)per * total, cudaMemcpyHostToDevice);
    for (int mode = 0; mode < 2; ++mode) for (int blocks : {8, 148}) {
        auto run = [&]() { if (mode == 0) gather<0><<<blocks, nt>>>(E, idx, o, per, total); else gather<1><<<blocks, nt>>>(E, idx, o, per, total); };
        run(); cudaEventRecord(e0); run(); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        const double clk = ms * 1e-3 * 1.965e9;
        printf("gather %s, %3d SMs x 512 thr: %.1f clk per warp-record-gather per SM, %.2f records/clk/SM, %.1f GB/s total\n", mode ? "LDG.256  " : "2xLDG.128", blocks,
               clk / (16.0 * per), 512.0 * per / clk, blocks * 512.0 * per * 32 / (ms * 1e-3) / 1e9);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
