// Per-generation reduction of the mesh tally to k, flux and fission source,
// kept on the device so generations can be queued back to back without a host
// round trip.  Replaces the serial G*N loop of src/mc_code.rs:340-363 and
// reproduces its binary32 arithmetic operation for operation (the previous k
// cancels algebraically but not in rounding, so it is carried on the device).
//
// One block.  Phase 1 is elementwise over bins (parallel).  Phase 2 is the
// running sum `k_new += k*dx*fission_source`, which the reference performs in
// (group-major, cell-minor) order; a single thread replays that order from
// shared memory so the result is bit-identical (G*N <= ~16k adds).
#include "mc_device.cuh"
#include "mc_internal.h"

namespace nraps {

namespace {

constexpr int kFinalizeThreads = 1024;
constexpr int kStage = 4096;

__global__ void __launch_bounds__(kFinalizeThreads) finalize_kernel(const FinalizeParams P)
{
    __shared__ float stage[kStage];
    __shared__ float k_acc;
    const int tid = threadIdx.x;
    const int N = (int)P.N, G = (int)P.G, M = (int)P.M, GN = G * N;
    const float k = *P.k_cur;
    const bool accumulate = P.gen >= P.skip;
    // (3565e6 * k * 36.2) / (200e6 * 1.602176634e-19 * nut[0 + M*1] * mesh_right[N-1])
    const float conversion = fdiv(fmul(fmul(3565e6f, k), 36.2f),
                                  fmul(fmul(fmul(200e6f, 1.602176634e-19f), P.nut_m1), P.length));
    const float kh = fmul(k, P.histories_f32);

    // phase 1a: per cell, groups in order (fission_source[i] accumulates over g in loop order)
    for (int i = tid; i < N; i += kFinalizeThreads) {
        const float dx = P.dx[i];
        const int mat = P.matid[i];
        float fis_acc = P.res_fission[i];
        for (int g = 0; g < G; ++g) {
            const int bin = g * N + i;
            const float tally = (float)((double)P.tally[bin] * kTallyInvScale);
            const float flux = fdiv(tally, fmul(kh, dx));
            const float fs = fmul(fmul(P.nusigf_nut[mat + M * g], P.sigf[mat + M * g]), flux);
            P.terms[bin] = fmul(fmul(k, dx), fs);
            if (accumulate) {
                const float term = fmul(flux, conversion);
                P.res_flux[bin] = fadd(P.res_flux[bin], fmul(term, P.fund));
                P.res_moments[bin] += (double)term; // extension outputs in f64: term^2 ~ 1e38 overflows f32
                P.res_moments[GN + bin] += (double)term * (double)term;
                fis_acc = fadd(fis_acc, fmul(fs, P.fund));
            }
        }
        if (accumulate) P.res_fission[i] = fis_acc;
    }
    if (tid == 0) k_acc = 0.0f;
    __syncthreads();

    // phase 2: ordered sum
    for (int base = 0; base < GN; base += kStage) {
        const int n = (GN - base < kStage) ? GN - base : kStage;
        for (int i = tid; i < n; i += kFinalizeThreads) stage[i] = P.terms[base + i];
        __syncthreads();
        if (tid == 0) {
            float acc = k_acc;
            for (int i = 0; i < n; ++i) acc = fadd(acc, stage[i]);
            k_acc = acc;
        }
        __syncthreads();
    }
    if (tid == 0) {
        P.k_hist[P.gen] = k_acc;
        *P.k_cur = k_acc;
    }
    if (tid < NRAPS_CT_WORDS && P.counters) P.counters_total[tid] += P.counters[tid];
}

__global__ void probe_logf_kernel(const float *x, float *out, uint32_t n)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[i] = mc_logf(x[i]);
}

__global__ void probe_div_kernel(const float *t, const float *mu, float *out_fast, float *out_ieee, uint32_t n)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        out_fast[i] = fast_div(t[i], make_recip(mu[i]));
        out_ieee[i] = fdiv(t[i], mu[i]);
    }
}

__global__ void probe_pcg_kernel(uint64_t state, uint64_t inc, uint32_t n, uint32_t *out_u32, float *out_unit)
{
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        for (uint32_t i = 0; i < n; ++i) {
            const uint32_t u = pcg32_next(state, inc);
            out_u32[i] = u;
            out_unit[i] = unit_from_u32(u);
        }
    }
}

} // namespace

cudaError_t launch_finalize(const FinalizeParams &p, cudaStream_t s)
{
    finalize_kernel<<<1, kFinalizeThreads, 0, s>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_probe_logf(const float *x, float *out, uint32_t n, cudaStream_t s)
{
    probe_logf_kernel<<<148, 256, 0, s>>>(x, out, n);
    return cudaGetLastError();
}

cudaError_t launch_probe_div(const float *t, const float *mu, float *out_fast, float *out_ieee, uint32_t n, cudaStream_t s)
{
    probe_div_kernel<<<148, 256, 0, s>>>(t, mu, out_fast, out_ieee, n);
    return cudaGetLastError();
}

cudaError_t launch_probe_pcg(uint64_t state, uint64_t inc, uint32_t n, uint32_t *out_u32, float *out_unit, cudaStream_t s)
{
    probe_pcg_kernel<<<1, 32, 0, s>>>(state, inc, n, out_u32, out_unit);
    return cudaGetLastError();
}

} // namespace nraps
