// Roofline denominators measured on the device the library runs on (bench.py calls these in the same run as the timed
// kernels, so the fractions it prints are same-box, same-clocks numbers and not constants from an earlier probe):
//
//   ddmpc_probe_fp64_tflops   sustained FP64 rate of the whole chip, issued as DFMA or as DMMA (mma.sync m8n8k4.f64, the
//                             only FP64 tensor instruction of sm_100a: tcgen05.mma has no f64 kind).  On B200 both reach
//                             the same 64 FMA/clk/SM (profiles/r1_probes.txt item 1), which is why moving work between the
//                             two pipes never added throughput to the closed-loop kernels.
//   ddmpc_probe_store_ms      time to write the (B, n_steps, 2) x 2 trajectory arrays of the closed-loop kernels with NO
//                             compute, either in the kernels' own pattern (a thread owns a loop and writes one 32-byte
//                             sector at a time: the reference layout makes a warp store touch 32 sectors 16*n_steps bytes
//                             apart) or fully coalesced.  The first is the floor of any kernel that produces the reference
//                             layout one block of steps at a time; the second is the HBM write peak for the same bytes.
//
// Standalone versions with more modes: scripts/probes/fp64_pipes.cu, scripts/probes/store_pattern.cu.
#include "common.cuh"

namespace ddmpc {

template <int MODE>   // 0 = DFMA, 1 = DMMA
__global__ void __launch_bounds__(512)
k_probe_fp64(double *out, int iters) {
    double a = threadIdx.x * 1e-3, b = 1.0000001;
    double c[8][2], f[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { c[i][0] = i; c[i][1] = -i; f[i] = i * 0.5; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 1)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
            else {
#pragma unroll
                for (int j = 0; j < 8; ++j) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(f[j]) : "d"(a), "d"(b));
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1] + f[i];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// the closed-loop kernels' store pattern: thread = loop, one 32-byte sector (two 16-byte steps) per store and array
__global__ void k_probe_store_owner(double *u, double *y, int B, int ns) {
    const int b = blockIdx.x * 64 + 2 * (threadIdx.x & 31) + (threadIdx.x >> 5);
    if (b >= B) return;
    const size_t f0 = (size_t)b * ns;
    double v = b;
    for (int k = 0; k < ns; ++k) {
        const size_t f = f0 + k;
        v = v * 1.0000001 + 1.0;
        if (f & 1) {
            if (k == 0) {
                *reinterpret_cast<double2 *>(u + f * 2) = make_double2(v, v);
                *reinterpret_cast<double2 *>(y + f * 2) = make_double2(v, v);
            } else {
                asm volatile("st.global.v4.f64 [%0], {%1, %1, %1, %1};" ::"l"(u + (f - 1) * 2), "d"(v) : "memory");
                asm volatile("st.global.v4.f64 [%0], {%1, %1, %1, %1};" ::"l"(y + (f - 1) * 2), "d"(v) : "memory");
            }
        }
    }
    const size_t fl = f0 + ns - 1;
    if ((fl & 1) == 0) {
        *reinterpret_cast<double2 *>(u + fl * 2) = make_double2(v, v);
        *reinterpret_cast<double2 *>(y + fl * 2) = make_double2(v, v);
    }
}

__global__ void k_probe_store_coalesced(double *u, double *y, size_t n32) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n32; i += stride) {
        const double v = (double)i;
        asm volatile("st.global.v4.f64 [%0], {%1, %1, %1, %1};" ::"l"((char *)u + i * 32), "d"(v) : "memory");
        asm volatile("st.global.v4.f64 [%0], {%1, %1, %1, %1};" ::"l"((char *)y + i * 32), "d"(v) : "memory");
    }
}

struct EventPair {
    cudaEvent_t a = nullptr, b = nullptr;
    ~EventPair() {
        if (a) cudaEventDestroy(a);
        if (b) cudaEventDestroy(b);
    }
};

}  // namespace ddmpc

using namespace ddmpc;

extern "C" {

int ddmpc_probe_fp64_tflops(int use_dmma, double *tflops, void *stream) {
    if (!tflops) return fail(DDMPC_ERR_INVALID_ARG, "probe: null argument");
    cudaStream_t st = (cudaStream_t)stream;
    int dev = 0, sms = 0;
    DDMPC_CUDA(cudaGetDevice(&dev));
    DDMPC_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    ScratchStreamScope scope(st);
    DevBuf out;
    const int threads = 512, ctas = sms * 2;              // 8 warps per scheduler: the FP64 units never starve
    DDMPC_CUDA(out.alloc(sizeof(double) * (size_t)ctas * threads));
    EventPair ev;
    DDMPC_CUDA(cudaEventCreate(&ev.a));
    DDMPC_CUDA(cudaEventCreate(&ev.b));
    const int iters = 4000;
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        DDMPC_CUDA(cudaEventRecord(ev.a, st));
        if (use_dmma) k_probe_fp64<1><<<ctas, threads, 0, st>>>(out.d(), iters);
        else k_probe_fp64<0><<<ctas, threads, 0, st>>>(out.d(), iters);
        DDMPC_LAUNCH_CHECK();
        DDMPC_CUDA(cudaEventRecord(ev.b, st));
        DDMPC_CUDA(cudaEventSynchronize(ev.b));
        float ms = 0.f;
        DDMPC_CUDA(cudaEventElapsedTime(&ms, ev.a, ev.b));
        if (rep > 0 && ms < best) best = ms;
    }
    // per warp and iteration: 8 DMMA x 256 FMA, or 64 DFMA x 32 lanes - 2048 FMA either way
    const double fma = (double)ctas * (threads / 32) * (double)iters * 2048.0;
    *tflops = 2.0 * fma / (best * 1e-3) / 1e12;
    return DDMPC_OK;
}

int ddmpc_probe_store_ms(int B, int n_steps, int coalesced, double *u, double *y, double *ms_out, void *stream) {
    if (!u || !y || !ms_out || B <= 0 || n_steps <= 0) return fail(DDMPC_ERR_INVALID_ARG, "probe: bad argument");
    if ((reinterpret_cast<uintptr_t>(u) & 31) || (reinterpret_cast<uintptr_t>(y) & 31))
        return fail(DDMPC_ERR_INVALID_ARG, "probe: buffers must be 32-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    int dev = 0, sms = 0;
    DDMPC_CUDA(cudaGetDevice(&dev));
    DDMPC_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    EventPair ev;
    DDMPC_CUDA(cudaEventCreate(&ev.a));
    DDMPC_CUDA(cudaEventCreate(&ev.b));
    const size_t bytes = (size_t)B * n_steps * 16;
    float best = 1e30f;
    for (int rep = 0; rep < 6; ++rep) {
        DDMPC_CUDA(cudaEventRecord(ev.a, st));
        if (coalesced) k_probe_store_coalesced<<<sms * 8, 256, 0, st>>>(u, y, bytes / 32);
        else k_probe_store_owner<<<ceil_div(B, 64), 64, 0, st>>>(u, y, B, n_steps);
        DDMPC_LAUNCH_CHECK();
        DDMPC_CUDA(cudaEventRecord(ev.b, st));
        DDMPC_CUDA(cudaEventSynchronize(ev.b));
        float ms = 0.f;
        DDMPC_CUDA(cudaEventElapsedTime(&ms, ev.a, ev.b));
        if (rep > 0 && ms < best) best = ms;
    }
    *ms_out = best;
    return DDMPC_OK;
}

}  // extern "C"
