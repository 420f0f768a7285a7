// Probe: how fast can a B200 absorb 841 MB of trajectory stores in the reference layout (B, n_steps, 2) x 2 arrays?
//   A: the fused kernels' pattern - a thread owns a loop and writes one 32-B sector (2 steps) at a time: a warp
//      store touches 32 sectors that are 6416 B apart.
//   B: 4 lanes share a loop and write one full 128-B line (8 steps) per instruction: a warp store = 8 lines.
//   C: fully coalesced (a warp store = 1 KB contiguous), same bytes.  Upper bound of the store path.
#include <cstdio>
#include <cuda_runtime.h>
constexpr int NS = 401;
__global__ void kA(double *u, double *y, int B) {
    const int b = blockIdx.x * 64 + 2 * (threadIdx.x & 31) + (threadIdx.x >> 5);
    if (b >= B) return;
    const size_t f0 = (size_t)b * NS;
    double v = b;
    for (int k = 0; k < NS; ++k) {
        const size_t f = f0 + k;
        v = v * 1.0000001 + 1.0;
        if (f & 1) {
            if (k == 0) { *reinterpret_cast<double2 *>(u + f * 2) = make_double2(v, v); *reinterpret_cast<double2 *>(y + f * 2) = make_double2(v, v); }
            else {
                asm volatile("st.global.v4.f64 [%0], {%1, %1, %1, %1};" ::"l"(u + (f - 1) * 2), "d"(v) : "memory");
                asm volatile("st.global.v4.f64 [%0], {%1, %1, %1, %1};" ::"l"(y + (f - 1) * 2), "d"(v) : "memory");
            }
        }
    }
    const size_t fl = f0 + NS - 1;
    if ((fl & 1) == 0) { *reinterpret_cast<double2 *>(u + fl * 2) = make_double2(v, v); *reinterpret_cast<double2 *>(y + fl * 2) = make_double2(v, v); }
}
// B: warp handles 8 loops x {u}, then {y}; lane = 4 * loop_in_group + sector; iterates over 128-B lines of the loop's region
template <int LPL>   // lanes per loop: each instruction writes LPL * 32 contiguous bytes per loop
__global__ void kB(double *u, double *y, int B) {
    const int lane = threadIdx.x & 31, wg = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int b = wg * (32 / LPL) + (lane / LPL), sec = lane % LPL;
    if (b >= B) return;
    const size_t byte0 = (size_t)b * NS * 16, byte1 = byte0 + (size_t)NS * 16;
    const size_t line0 = byte0 & ~(size_t)(32 * LPL - 1);
    double v = b;
    for (size_t a = line0 + 32 * sec; a < byte1; a += 32 * LPL) {
        v = v * 1.0000001 + 1.0;
        if (a >= byte0 && a + 32 <= byte1) {
            asm volatile("st.global.v4.f64 [%0], {%1, %1, %1, %1};" ::"l"((char *)u + a), "d"(v) : "memory");
            asm volatile("st.global.v4.f64 [%0], {%1, %1, %1, %1};" ::"l"((char *)y + a), "d"(v) : "memory");
        }
    }
}
// A2: owner thread, two consecutive sectors (64 B) back to back every 4 steps
__global__ void kA2(double *u, double *y, int B) {
    const int b = blockIdx.x * 64 + 2 * (threadIdx.x & 31) + (threadIdx.x >> 5);
    if (b >= B) return;
    const size_t byte0 = (size_t)b * NS * 16, byte1 = byte0 + (size_t)NS * 16;
    double v = b;
    for (size_t a = byte0 & ~(size_t)63; a < byte1; a += 64) {
        v = v * 1.0000001 + 1.0;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const size_t aa = a + 32 * h;
            if (aa >= byte0 && aa + 32 <= byte1) {
                asm volatile("st.global.v4.f64 [%0], {%1, %1, %1, %1};" ::"l"((char *)u + aa), "d"(v) : "memory");
                asm volatile("st.global.v4.f64 [%0], {%1, %1, %1, %1};" ::"l"((char *)y + aa), "d"(v) : "memory");
            }
        }
    }
}
// A1: pattern A with consecutive lanes on consecutive loops
__global__ void kA1(double *u, double *y, int B) {
    const int b = blockIdx.x * 64 + threadIdx.x;
    if (b >= B) return;
    const size_t byte0 = (size_t)b * NS * 16, byte1 = byte0 + (size_t)NS * 16;
    double v = b;
    for (size_t a = byte0 & ~(size_t)31; a < byte1; a += 32) {
        v = v * 1.0000001 + 1.0;
        if (a >= byte0 && a + 32 <= byte1) {
            asm volatile("st.global.v4.f64 [%0], {%1, %1, %1, %1};" ::"l"((char *)u + a), "d"(v) : "memory");
            asm volatile("st.global.v4.f64 [%0], {%1, %1, %1, %1};" ::"l"((char *)y + a), "d"(v) : "memory");
        }
    }
}
__global__ void kC(double *u, double *y, size_t n32) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n32; i += stride) {
        const double v = (double)i;
        asm volatile("st.global.v4.f64 [%0], {%1, %1, %1, %1};" ::"l"((char *)u + i * 32), "d"(v) : "memory");
        asm volatile("st.global.v4.f64 [%0], {%1, %1, %1, %1};" ::"l"((char *)y + i * 32), "d"(v) : "memory");
    }
}
int main() {
    for (int B : {65536, 262144}) {
        double *u, *y;
        const size_t bytes = (size_t)B * NS * 16;
        cudaMalloc(&u, bytes + 256); cudaMalloc(&y, bytes + 256);
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        for (int mode = 0; mode < 8; ++mode) {
            float best = 1e9;
            for (int rep = 0; rep < 6; ++rep) {
                cudaEventRecord(e0);
                if (mode == 0) kA<<<(B + 63) / 64, 64>>>(u, y, B);
                if (mode == 1) kB<4><<<(B / 8 * 32 + 127) / 128, 128>>>(u, y, B);
                if (mode == 3) kB<2><<<(B / 16 * 32 + 127) / 128, 128>>>(u, y, B);
                if (mode == 4) kB<8><<<(B / 4 * 32 + 127) / 128, 128>>>(u, y, B);
                if (mode == 5) kB<16><<<(B / 2 * 32 + 127) / 128, 128>>>(u, y, B);
                if (mode == 6) kA2<<<(B + 63) / 64, 64>>>(u, y, B);
                if (mode == 7) kA1<<<(B + 63) / 64, 64>>>(u, y, B);
                if (mode == 2) kC<<<148 * 8, 256>>>(u, y, bytes / 32);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                if (rep > 0 && ms < best) best = ms;
            }
            printf("B=%d mode %d(%c): %.4f ms  -> %.0f GB/s\n", B, mode, "ABCbdeaa"[mode], best, 2.0 * bytes / best * 1e-6);
        }
        printf("%s\n", cudaGetErrorString(cudaGetLastError()));
        cudaFree(u); cudaFree(y);
    }
}
