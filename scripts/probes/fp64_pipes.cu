// Probe: FP64 throughput of one B200 SM.  DFMA alone, DMMA (m8n8k4) alone, and both mixed in one warp,
// for 1..4 warps per scheduler.  Prints FMA per clock per SM.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>   // 0 = DFMA, 1 = DMMA, 2 = mixed (1 DMMA : 8 DFMA, equal FMA counts... 256 vs 8*32)
__global__ void k(double *out, long long *cyc, int iters) {
    double a = threadIdx.x * 1e-3, b = 1.0000001;
    double c[8][2], f[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { c[i][0] = i; c[i][1] = -i; f[i] = i * 0.5; }
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 1 || MODE == 2)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
            if (MODE == 0 || MODE == 2) {
#pragma unroll
                for (int j = 0; j < 8; ++j) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(f[j]) : "d"(a), "d"(b));
            }
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1] + f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
    double *out; long long *cyc, h[148];
    cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 148 * 8);
    const int iters = 2000;
    for (int mode = 0; mode < 3; ++mode)
        for (int wps = 1; wps <= 4; ++wps) {
            const int threads = 128 * wps;
            for (int rep = 0; rep < 2; ++rep) {
                if (mode == 0) k<0><<<148, threads>>>(out, cyc, iters);
                if (mode == 1) k<1><<<148, threads>>>(out, cyc, iters);
                if (mode == 2) k<2><<<148, threads>>>(out, cyc, iters);
                cudaDeviceSynchronize();
            }
            cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
            const double warps = threads / 32.0;
            const double fma_dfma = (mode != 1) ? warps * iters * 64.0 * 32 : 0, fma_dmma = (mode != 0) ? warps * iters * 8.0 * 256 : 0;
            printf("mode %d (%s) warps/scheduler %d: %lld cycles, DFMA %.1f + DMMA %.1f FMA/clk/SM\n", mode,
                   mode == 0 ? "DFMA" : mode == 1 ? "DMMA" : "mixed", wps, h[0], fma_dfma / h[0], fma_dmma / h[0]);
        }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
