// Probe: do FP64 DMMAs and the operand loads that feed them overlap on a B200 SM?
// Per "k-step" a warp issues one DMMA (m8n8k4 f64) and, depending on the mode,
//   mode 0: nothing else (operands stay in registers)
//   mode 1: one LDS.64 per lane (the B fragment, 256 B per warp, conflict-free)
//   mode 2: one LDS.64 and one LDG.64 (A fragment from an L1-resident array, 256 B per warp, coalesced)
//   mode 3: the two loads without the DMMA
//   mode 4: one LDG.64 only with the DMMA
// 8 independent accumulators per warp; the loaded values feed the DMMAs of the NEXT trip (software pipelined), so no
// DMMA waits for a load of its own trip.  Reports cycles per k-step per scheduler for 1..4 warps per scheduler.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(const double *__restrict__ ag, double *out, long long *cyc, int iters) {
    __shared__ double sb[4][32 * 33];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 4 * 32 * 33; i += blockDim.x) (&sb[0][0])[i] = 1.0 + 1e-9 * i;
    double a[8], b[8], c[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = threadIdx.x * 1e-3 + i; b[i] = 1.0000001 + 0.1 * i; c[i][0] = i; c[i][1] = -i; }
    __syncthreads();
    const double *bp = &sb[warp & 3][lane];
    const double *ap = ag + (size_t)(warp & 15) * 8 * 32 + lane;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        double an[8], bn[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 1 || MODE == 2 || MODE == 3) bn[i] = *(volatile const double *)(bp + 32 * ((i + it) & 31));
            else bn[i] = b[i];
            if (MODE == 2 || MODE == 3 || MODE == 4) asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(an[i]) : "l"(ap + 32 * i));
            else an[i] = a[i];
            if (MODE != 3)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a[i]), "d"(b[i]));
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) { a[i] = an[i]; b[i] = bn[i]; }
        if (MODE == 3) {
#pragma unroll
            for (int i = 0; i < 8; ++i) { c[i][0] += a[i]; c[i][1] += b[i]; }
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
    double *out, *ag; long long *cyc, h[148];
    cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 148 * 8); cudaMalloc(&ag, 16 * 8 * 32 * 8);
    cudaMemset(ag, 0, 16 * 8 * 32 * 8);
    const int iters = 2000;
    const char *names[5] = {"DMMA only", "DMMA + LDS.64", "DMMA + LDS.64 + LDG.64", "LDS.64 + LDG.64 only", "DMMA + LDG.64"};
    for (int mode = 0; mode < 5; ++mode)
        for (int wps = 1; wps <= 4; ++wps) {
            const int threads = 128 * wps;
            for (int rep = 0; rep < 2; ++rep) {
                if (mode == 0) k<0><<<148, threads>>>(ag, out, cyc, iters);
                if (mode == 1) k<1><<<148, threads>>>(ag, out, cyc, iters);
                if (mode == 2) k<2><<<148, threads>>>(ag, out, cyc, iters);
                if (mode == 3) k<3><<<148, threads>>>(ag, out, cyc, iters);
                if (mode == 4) k<4><<<148, threads>>>(ag, out, cyc, iters);
                cudaDeviceSynchronize();
            }
            cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
            printf("%-24s warps/scheduler %d: %6.2f cycles per k-step per scheduler  (%5.1f FMA/clk/SM if a DMMA each)\n", names[mode], wps,
                   h[0] / (wps * iters * 8.0), (threads / 32.0) * iters * 8.0 * 256 / h[0]);
        }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
