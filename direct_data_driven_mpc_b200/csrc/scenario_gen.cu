// On-device scenario generation with NumPy-compatible random streams.
//
// Replaces, for S seeds at once, stages 1-3 of the reference's example script
// (examples/direct_data_driven_mpc_example.py:263-300):
//   randomize_initial_system_state       utilities/controller/controller_operation.py:59-75
//   generate_initial_input_output_data   utilities/controller/controller_operation.py:126-133
// and the later noise draws (controller_operation.py:193, 263).  Thread s owns
// np.random.default_rng(seeds[s]): SeedSequence entropy mixing -> PCG64 (XSL-RR 128/64) ->
// Generator.uniform, restated from NumPy's published algorithm (numpy/random/bit_generator.pyx,
// _pcg64.pyx, src/pcg64/pcg64.h), so every draw is bit-identical to what the reference would
// draw for `--seed seeds[s]`, in the same order.
#include <vector>

#include "common.cuh"
#include "plan.cuh"

namespace ddmpc {

struct U128 {
    uint64_t hi, lo;
};
__device__ __forceinline__ U128 u128_add(U128 a, U128 b) {
    U128 r;
    r.lo = a.lo + b.lo;
    r.hi = a.hi + b.hi + (r.lo < a.lo ? 1ull : 0ull);
    return r;
}
__device__ __forceinline__ U128 u128_mul(U128 a, U128 b) {   // low 128 bits of a*b
    U128 r;
    r.lo = a.lo * b.lo;
    r.hi = __umul64hi(a.lo, b.lo) + a.hi * b.lo + a.lo * b.hi;
    return r;
}

struct Pcg64 {
    U128 state, inc;
    __device__ __forceinline__ void step() {
        const U128 mult = {0x2360ED051FC65DA4ull, 0x4385DF649FCCF645ull};
        state = u128_add(u128_mul(state, mult), inc);
    }
    __device__ __forceinline__ uint64_t next64() {     // pcg_setseq_128_xsl_rr_64_random_r
        step();
        const uint64_t x = state.hi ^ state.lo;
        const unsigned r = (unsigned)(state.hi >> 58);
        return (x >> r) | (x << ((64u - r) & 63u));
    }
    __device__ __forceinline__ double next_double() { return (double)(next64() >> 11) * (1.0 / 9007199254740992.0); }
    // Generator.uniform(low, high): low + (high - low) * next_double, no FMA contraction
    __device__ __forceinline__ double uniform(double low, double range) {
        return __dadd_rn(low, __dmul_rn(range, next_double()));
    }
};

// np.random.default_rng(seed): SeedSequence(seed) -> generate_state(4 x uint64) -> pcg64_set_seed
__device__ Pcg64 pcg64_from_seed(uint64_t seed) {
    const uint32_t INIT_A = 0x43b0d7e5u, MULT_A = 0x931e8875u, INIT_B = 0x8b51f9ddu, MULT_B = 0x58f38dedu;
    const uint32_t MIX_L = 0xca01f9ddu, MIX_R = 0x4973f715u;
    uint32_t ent[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    const int n_ent = ent[1] ? 2 : 1;
    uint32_t hc = INIT_A, pool[4];
    auto hashmix = [&](uint32_t v) {
        v ^= hc;
        hc *= MULT_A;
        v *= hc;
        v ^= v >> 16;
        return v;
    };
    auto mix = [&](uint32_t x, uint32_t y) {
        uint32_t r = MIX_L * x - MIX_R * y;
        r ^= r >> 16;
        return r;
    };
    for (int i = 0; i < 4; ++i) pool[i] = hashmix(i < n_ent ? ent[i] : 0u);
    for (int i_src = 0; i_src < 4; ++i_src)
        for (int i_dst = 0; i_dst < 4; ++i_dst)
            if (i_src != i_dst) pool[i_dst] = mix(pool[i_dst], hashmix(pool[i_src]));
    uint32_t w[8];
    hc = INIT_B;
    for (int i = 0; i < 8; ++i) {
        uint32_t v = pool[i & 3];
        v ^= hc;
        hc *= MULT_B;
        v *= hc;
        v ^= v >> 16;
        w[i] = v;
    }
    uint64_t v64[4];
    for (int i = 0; i < 4; ++i) v64[i] = (uint64_t)w[2 * i] | ((uint64_t)w[2 * i + 1] << 32);
    Pcg64 g;
    const U128 initstate = {v64[0], v64[1]}, initseq = {v64[2], v64[3]};
    g.inc.hi = (initseq.hi << 1) | (initseq.lo >> 63);
    g.inc.lo = (initseq.lo << 1) | 1ull;
    g.state = {0ull, 0ull};
    g.step();
    g.state = u128_add(g.state, initstate);
    g.step();
    return g;
}

struct GenArgs {
    int S, N, n_x, m, p;
    const double *A, *B, *C, *D;        // device, row-major
    const double *pinvOt, *Tt;          // (n_x, p*n_x), (p*n_x, m*n_x)
    const unsigned long long *seeds;    // device
    double u_lo, u_range, eps;
    double *x0, *u_d, *y_d, *x_end, *scratch;   // scratch: S * (n_x*(m+p) + 2*n_x) doubles
    unsigned long long *rng_state;      // (S, 4): state.hi, state.lo, inc.hi, inc.lo
};

__device__ void plant_step(const GenArgs &a, double *x, double *xn, const double *u, const double *w, double *y) {
    for (int i = 0; i < a.p; ++i) {     // y = C x + D u + w   (model_simulation.py:94)
        double acc = 0.0, acd = 0.0;
        for (int j = 0; j < a.n_x; ++j) acc = fma(a.C[i * a.n_x + j], x[j], acc);
        for (int j = 0; j < a.m; ++j) acd = fma(a.D[i * a.m + j], u[j], acd);
        y[i] = (acc + acd) + w[i];
    }
    for (int i = 0; i < a.n_x; ++i) {   // x <- A x + B u      (model_simulation.py:96)
        double acc = 0.0, acb = 0.0;
        for (int j = 0; j < a.n_x; ++j) acc = fma(a.A[i * a.n_x + j], x[j], acc);
        for (int j = 0; j < a.m; ++j) acb = fma(a.B[i * a.m + j], u[j], acb);
        xn[i] = acc + acb;
    }
    for (int i = 0; i < a.n_x; ++i) x[i] = xn[i];
}

__global__ void k_generate_example_data(GenArgs a) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= a.S) return;
    const int nx = a.n_x, m = a.m, p = a.p;
    Pcg64 g = pcg64_from_seed(a.seeds[s]);
    double *sc = a.scratch + (size_t)s * (nx * (m + p) + 2 * nx);
    double *ui = sc, *yi = ui + nx * m, *x = yi + nx * p, *xn = x + nx;
    // (1) x_i0 ~ U(-1, 1)^n_x   (2) u_i ~ U(u_range)^(n_x, m)   (3) w_i = eps U(-1, 1)^(n_x, p)
    for (int i = 0; i < nx; ++i) x[i] = g.uniform(-1.0, 2.0);
    for (int i = 0; i < nx * m; ++i) ui[i] = g.uniform(a.u_lo, a.u_range);
    for (int i = 0; i < nx * p; ++i) yi[i] = a.eps * g.uniform(-1.0, 2.0);          // holds w_i, then y_i
    for (int k = 0; k < nx; ++k) plant_step(a, x, xn, ui + k * m, yi + k * p, yi + k * p);
    // x_0 = pinv(Ot) (Y - Tt U)   (initial_state_estimation.py:131)
    double *x0 = a.x0 + (size_t)s * nx;
    for (int i = 0; i < nx; ++i) {
        double acc = 0.0;
        for (int r = 0; r < p * nx; ++r) {
            double t = 0.0;
            for (int c = 0; c < m * nx; ++c) t = fma(a.Tt[r * (m * nx) + c], ui[c], t);
            acc = fma(a.pinvOt[i * (p * nx) + r], yi[r] - t, acc);
        }
        x0[i] = acc;
    }
    for (int i = 0; i < nx; ++i) x[i] = x0[i];
    // (4) u_d ~ U(u_range)^(N, m), drawn completely before (5) w_d = eps U(-1, 1)^(N, p)
    double *ud = a.u_d + (size_t)s * a.N * m, *yd = a.y_d + (size_t)s * a.N * p;
    for (int i = 0; i < a.N * m; ++i) ud[i] = g.uniform(a.u_lo, a.u_range);
    for (int k = 0; k < a.N; ++k) {
        double *yk = yd + (size_t)k * p;
        for (int i = 0; i < p; ++i) yk[i] = a.eps * g.uniform(-1.0, 2.0);
        plant_step(a, x, xn, ud + (size_t)k * m, yk, yk);
    }
    for (int i = 0; i < nx; ++i) a.x_end[(size_t)s * nx + i] = x[i];
    unsigned long long *rs = a.rng_state + (size_t)s * 4;
    rs[0] = g.state.hi; rs[1] = g.state.lo; rs[2] = g.inc.hi; rs[3] = g.inc.lo;
}

// out[s][i] = scale * uniform(lo, lo + range), i < count, continuing stream s
__global__ void k_pcg64_uniform(int S, int count, double lo, double range, double scale,
                                unsigned long long *__restrict__ rng_state, double *__restrict__ out) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    unsigned long long *rs = rng_state + (size_t)s * 4;
    Pcg64 g;
    g.state = {rs[0], rs[1]};
    g.inc = {rs[2], rs[3]};
    double *o = out + (size_t)s * count;
    for (int i = 0; i < count; ++i) o[i] = scale * g.uniform(lo, range);
    rs[0] = g.state.hi; rs[1] = g.state.lo;
}

}  // namespace ddmpc

using namespace ddmpc;

extern "C" {

int ddmpc_generate_example_data(const ddmpc_plant *plant, const double *pinv_Ot, const double *Tt, int S,
                                const uint64_t *seeds, int N, double u_lo, double u_hi, double eps, double *x0,
                                double *u_d, double *y_d, double *x_end, uint64_t *rng_state, void *stream) {
    if (!plant || !pinv_Ot || !Tt || !seeds || S <= 0 || N <= 0 || !x0 || !u_d || !y_d || !x_end || !rng_state)
        return fail(DDMPC_ERR_INVALID_ARG, "generate_example_data: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    const int nx = plant->n_x, m = plant->m, p = plant->p;
    const size_t nA = (size_t)nx * nx, nB = (size_t)nx * m, nC = (size_t)p * nx, nD = (size_t)p * m;
    const size_t nP = (size_t)nx * p * nx, nT = (size_t)p * nx * m * nx;
    std::vector<double> h(nA + nB + nC + nD + nP + nT);
    size_t o = 0;
    auto put = [&](const double *src, size_t n) { std::copy(src, src + n, h.begin() + o); o += n; };
    put(plant->A, nA); put(plant->B, nB); put(plant->C, nC); put(plant->D, nD); put(pinv_Ot, nP); put(Tt, nT);
    DevBuf dm, ds, dscr;
    DDMPC_CUDA(dm.alloc(sizeof(double) * h.size()));
    DDMPC_CUDA(ds.alloc(sizeof(uint64_t) * S));
    DDMPC_CUDA(dscr.alloc(sizeof(double) * (size_t)S * (nx * (m + p) + 2 * nx)));
    DDMPC_CUDA(cudaMemcpyAsync(dm.p, h.data(), sizeof(double) * h.size(), cudaMemcpyHostToDevice, st));
    DDMPC_CUDA(cudaMemcpyAsync(ds.p, seeds, sizeof(uint64_t) * S, cudaMemcpyHostToDevice, st));
    GenArgs a{};
    a.S = S; a.N = N; a.n_x = nx; a.m = m; a.p = p;
    a.A = dm.d(); a.B = a.A + nA; a.C = a.B + nB; a.D = a.C + nC; a.pinvOt = a.D + nD; a.Tt = a.pinvOt + nP;
    a.seeds = (const unsigned long long *)ds.p;
    a.u_lo = u_lo; a.u_range = u_hi - u_lo; a.eps = eps;
    a.x0 = x0; a.u_d = u_d; a.y_d = y_d; a.x_end = x_end; a.scratch = dscr.d();
    a.rng_state = (unsigned long long *)rng_state;
    k_generate_example_data<<<ceil_div(S, 64), 64, 0, st>>>(a);
    DDMPC_LAUNCH_CHECK();
    DDMPC_CUDA(cudaStreamSynchronize(st));   // staging buffers die here
    return DDMPC_OK;
}

int ddmpc_pcg64_uniform(uint64_t *rng_state, int S, int count, double lo, double hi, double scale, double *out,
                        void *stream) {
    if (!rng_state || !out || S <= 0 || count < 0) return fail(DDMPC_ERR_INVALID_ARG, "pcg64_uniform: bad argument");
    if (count == 0) return DDMPC_OK;
    k_pcg64_uniform<<<ceil_div(S, 64), 64, 0, (cudaStream_t)stream>>>(S, count, lo, hi - lo, scale,
                                                                       (unsigned long long *)rng_state, out);
    DDMPC_LAUNCH_CHECK();
    return DDMPC_OK;
}

}  // extern "C"
