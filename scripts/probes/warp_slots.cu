// Probe: which hardware warp slots (%warpid) and SMs do the warps of 64-thread CTAs with 29 KB of shared memory get?
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(64, 7) k(int *out, long spin) {
    __shared__ double pad[3600];
    unsigned smid, wid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    asm volatile("mov.u32 %0, %%warpid;" : "=r"(wid));
    pad[threadIdx.x] = smid;
    long t0 = clock64();
    while (clock64() - t0 < spin) {}
    if ((threadIdx.x & 31) == 0) {
        out[(blockIdx.x * 2 + (threadIdx.x >> 5)) * 2 + 0] = smid;
        out[(blockIdx.x * 2 + (threadIdx.x >> 5)) * 2 + 1] = wid + (pad[threadIdx.x] < 0 ? 1 : 0);
    }
}
int main() {
    const int G = 1024;
    int *d, h[G * 4];
    cudaMalloc(&d, sizeof(h));
    cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    k<<<G, 64>>>(d, 2000000);
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
    for (int sm = 0; sm < 3; ++sm) {
        printf("SM %d:", sm);
        for (int b = 0; b < G; ++b)
            if (h[b * 4] == sm) printf("  cta%d:(%d,%d)", b, h[b * 4 + 1], h[b * 4 + 3]);
        printf("\n");
    }
    int hist[4][2] = {};
    for (int b = 0; b < G; ++b) { hist[h[b * 4 + 1] & 3][0]++; hist[h[b * 4 + 3] & 3][1]++; }
    for (int s = 0; s < 4; ++s) printf("smsp %d: warp0 x%d warp1 x%d\n", s, hist[s][0], hist[s][1]);
    return 0;
}
