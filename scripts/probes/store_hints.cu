// Probe: does a cache hint on the 32-byte trajectory stores (pattern A of store_pattern.cu) change the write floor?
#include <cstdio>
#include <cuda_runtime.h>
constexpr int NS = 401;
template <int MODE>
__device__ __forceinline__ void st32(double *p, double v, unsigned long long pol) {
    if (MODE == 0) asm volatile("st.global.v4.f64 [%0], {%1, %1, %1, %1};" ::"l"(p), "d"(v) : "memory");
    if (MODE == 1) asm volatile("st.global.cs.v4.f64 [%0], {%1, %1, %1, %1};" ::"l"(p), "d"(v) : "memory");
    if (MODE == 2) asm volatile("st.global.wt.v4.f64 [%0], {%1, %1, %1, %1};" ::"l"(p), "d"(v) : "memory");
    if (MODE == 3) asm volatile("st.global.L2::cache_hint.v4.f64 [%0], {%1, %1, %1, %1}, %2;" ::"l"(p), "d"(v), "l"(pol) : "memory");
    if (MODE == 4) asm volatile("st.global.cg.v4.f64 [%0], {%1, %1, %1, %1};" ::"l"(p), "d"(v) : "memory");
}
template <int MODE>
__global__ void kA(double *u, double *y, int B) {
    const int b = blockIdx.x * 64 + 2 * (threadIdx.x & 31) + (threadIdx.x >> 5);
    if (b >= B) return;
    unsigned long long pol = 0;
    if (MODE == 3) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    const size_t byte0 = (size_t)b * NS * 16, byte1 = byte0 + (size_t)NS * 16;
    double v = b;
    for (size_t a = byte0 & ~(size_t)31; a < byte1; a += 32) {
        v = v * 1.0000001 + 1.0;
        if (a >= byte0 && a + 32 <= byte1) {
            st32<MODE>((double *)((char *)u + a), v, pol);
            st32<MODE>((double *)((char *)y + a), v, pol);
        }
    }
}
int main() {
    const int B = 65536;
    double *u, *y;
    const size_t bytes = (size_t)B * NS * 16;
    cudaMalloc(&u, bytes + 256); cudaMalloc(&y, bytes + 256);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const char *names[5] = {"default", ".cs", ".wt", "L2::evict_first", ".cg"};
    for (int mode = 0; mode < 5; ++mode) {
        float best = 1e9;
        for (int rep = 0; rep < 6; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) kA<0><<<(B + 63) / 64, 64>>>(u, y, B);
            if (mode == 1) kA<1><<<(B + 63) / 64, 64>>>(u, y, B);
            if (mode == 2) kA<2><<<(B + 63) / 64, 64>>>(u, y, B);
            if (mode == 3) kA<3><<<(B + 63) / 64, 64>>>(u, y, B);
            if (mode == 4) kA<4><<<(B + 63) / 64, 64>>>(u, y, B);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (rep > 0 && ms < best) best = ms;
        }
        printf("%-18s %.4f ms -> %.0f GB/s   (%s)\n", names[mode], best, 2.0 * bytes / best * 1e-6, cudaGetErrorString(cudaGetLastError()));
    }
}
