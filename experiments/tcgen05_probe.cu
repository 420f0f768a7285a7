// Stage-1 probe for the tcgen05 route: D (128 x N, FP32 in TMEM) = A (128 x K, TF32, IN TMEM)
// x B^T (N x K, TF32, shared memory, K-major, no swizzle), K = 8 per tcgen05.mma.  Checks the descriptor encodings, the
// TMEM operand layout and the ld / st shapes against an exact host reference (inputs are multiples of 1/8, so every
// product and partial sum is exact in FP32).   nvcc -arch=sm_100a -o tcgen05_probe tcgen05_probe.cu && ./tcgen05_probe
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

constexpr int M = 128, N = 80, K = 96, KS = K / 8, COL_A = 0, COL_D = 256;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, no swizzle ("interleave"): 16-byte chunk c (4 tf32 along K) of row r at byte c * R * 16 + r * 16.
// Canonical layout ((8, n), 2) : ((1, SBO), LBO) in 16-byte units: SBO = 8 (next 8-row group), LBO = R (next K chunk).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, int rows) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);
    d |= (uint64_t)((rows * 16) >> 4 & 0x3fff) << 16;     // leading dimension byte offset: between the two K chunks
    d |= (uint64_t)((8 * 16) >> 4 & 0x3fff) << 32;        // stride dimension byte offset: between 8-row groups
    d |= (uint64_t)1 << 46;                               // descriptor version (Blackwell)
    return d;                                             // layout type 0 = no swizzle, base offset 0
}

__global__ void __launch_bounds__(128) k_probe(const float *A, const float *B, float *D) {
    __shared__ __align__(16) float Bs[(K / 4) * N * 4];   // [chunk][row][4]
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(8) uint64_t mbar;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int e = tid; e < N * K; e += 128) {
        const int r = e / K, k = e % K;
        Bs[((k >> 2) * N + r) * 4 + (k & 3)] = B[e];
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tb = tmem_base_s;
    // A: thread t owns row t = TMEM lane t; element k in column COL_A + k
    const uint32_t lane_addr = tb + ((uint32_t)(warp * 32) << 16);
    for (int k0 = 0; k0 < K; k0 += 8) {
        uint32_t v[8];
        for (int j = 0; j < 8; ++j) v[j] = __float_as_uint(A[tid * K + k0 + j]);
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(lane_addr + COL_A + k0),
                     "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]));
    }
    asm volatile("tcgen05.wait::st.sync.aligned;");
    asm volatile("fence.proxy.async.shared::cta;");       // Bs was written with ordinary stores: visible to the async proxy
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    if (tid == 0) {
        // instruction descriptor: D = F32, A = B = TF32, both K-major, N >> 3 at bit 17, M >> 4 at bit 24
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        const uint32_t sB = smem_u32(Bs);
        for (int ks = 0; ks < KS; ++ks) {
            const uint64_t db = make_desc(sB + ks * 2 * N * 16, N);
            const uint32_t acc = ks > 0;
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                         "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                         ::"r"(tb + COL_D), "r"(tb + COL_A + ks * 8), "l"(db), "r"(idesc), "r"(acc));
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
    }
    {   // wait for the MMAs
        uint32_t ok = 0;
        while (!ok)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(smem_u32(&mbar)), "r"(0u) : "memory");
    }
    asm volatile("tcgen05.fence::after_thread_sync;");
    for (int n0 = 0; n0 < N; n0 += 8) {
        uint32_t v[8];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                     : "r"(lane_addr + COL_D + n0));
        asm volatile("tcgen05.wait::ld.sync.aligned;");
        for (int j = 0; j < 8; ++j) D[tid * N + n0 + j] = __uint_as_float(v[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512));
}

int main() {
    std::vector<float> A(M * K), B(N * K), D(M * N, -1.f);
    for (int i = 0; i < M; ++i)
        for (int k = 0; k < K; ++k) A[i * K + k] = ((i * 3 + k * 5) % 17 - 8) / 8.0f;
    for (int n = 0; n < N; ++n)
        for (int k = 0; k < K; ++k) B[n * K + k] = ((n * 7 + k * 11) % 13 - 6) / 8.0f;
    float *dA, *dB, *dD;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0xff, D.size() * 4);
    k_probe<<<1, 128>>>(dA, dB, dD);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    double worst = 0;
    int bad = 0;
    for (int i = 0; i < M; ++i)
        for (int n = 0; n < N; ++n) {
            double ref = 0;
            for (int k = 0; k < K; ++k) ref += (double)A[i * K + k] * B[n * K + k];
            const double err = fabs(ref - D[i * N + n]);
            if (err > worst) worst = err;
            if (err > 1e-6 && bad < 8) { printf("  D[%d][%d] = %g, expected %g\n", i, n, D[i * N + n], ref); ++bad; }
        }
    printf("max |D - ref| = %g over %d x %d  ->  %s\n", worst, M, N, worst < 1e-6 ? "PASS" : "FAIL");
    return worst < 1e-6 ? 0 : 1;
}
