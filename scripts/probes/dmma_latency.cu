// Probe: DMMA (m8n8k4 f64) issue interval as a function of the number of independent accumulator chains per warp (NA) and of
// the warps per scheduler.  NA = 1 gives the dependent-issue latency; the table says how many chains x warps a kernel must
// keep in flight to reach the 16 cycles per DMMA per scheduler (64 FMA/clk/SM) of the FP64 units.
#include <cstdio>
#include <cuda_runtime.h>
template <int NA>
__global__ void k(double *out, long long *cyc, int iters) {
    double a = threadIdx.x * 1e-3 + 1.0, b = 1.0000001, c[NA][2];
#pragma unroll
    for (int i = 0; i < NA; ++i) { c[i][0] = i; c[i][1] = -i; }
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8 / NA; ++r)
#pragma unroll
            for (int i = 0; i < NA; ++i)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < NA; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int NA>
void run(double *out, long long *cyc) {
    long long h[148];
    const int iters = 2000;
    printf("chains per warp %d:", NA);
    for (int wps = 1; wps <= 8; wps *= 2) {
        const int threads = 128 * wps;
        for (int rep = 0; rep < 2; ++rep) { k<NA><<<148, threads>>>(out, cyc, iters); cudaDeviceSynchronize(); }
        cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        printf("   %d warps/scheduler: %5.1f cycles per DMMA per warp, %5.1f per scheduler", wps, h[0] / (iters * 8.0), h[0] / (iters * 8.0 * wps));
    }
    printf("\n");
}
int main() {
    double *out; long long *cyc;
    cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 148 * 8);
    run<1>(out, cyc); run<2>(out, cyc); run<4>(out, cyc); run<8>(out, cyc);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
