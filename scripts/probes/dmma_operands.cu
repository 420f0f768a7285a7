// Probe: does the DMMA (m8n8k4 f64) rate depend on operand reuse?  mode 0: every DMMA reads the same A and B registers
// (as in fp64_pipes.cu); mode 1: 8 accumulators x distinct A and B registers; mode 2: distinct A and B and the
// accumulators chained through 4 k-steps with different operands (what the closed-loop kernels issue).
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(double *out, long long *cyc, int iters) {
    double a[8], b[8], c[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = threadIdx.x * 1e-3 + i; b[i] = 1.0000001 + 0.1 * i; c[i][0] = i; c[i][1] = -i; }
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int ia = MODE == 0 ? 0 : (MODE == 1 ? i : (i + ks) & 7), ib = MODE == 0 ? 0 : (MODE == 1 ? i : (i + 2 * ks + 1) & 7);
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a[ia]), "d"(b[ib]));
            }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
    double *out; long long *cyc, h[148];
    cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 148 * 8);
    const int iters = 1000;
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
            printf("mode %d warps/scheduler %d: %.1f FMA/clk/SM  (%.1f cycles per DMMA per scheduler)\n", mode, wps,
                   (threads / 32.0) * iters * 32.0 * 256 / h[0], h[0] / (wps * iters * 32.0));
        }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
