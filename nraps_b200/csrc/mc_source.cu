// Source kernel: every history of the shard is born here, with all 32 lanes of a warp busy, instead of inside the
// persistent transport kernels where a neutron dies on one lane at a time (ncu: the single-lane spawn path was 6 %
// of the issue slots of the surface kernel and 16 % of the Woodcock kernel, profiles/r1d, r1g).
//
// Replaces spawn_neutron + energy (src/mc_code.rs:7-53, 228-230); draw order cell, position, mu, chi (:46-51), or
// site index, mu, chi in fission_bank mode.  Each thread takes kRun histories 32 apart so that only the first
// needs the full PCG32 jump; the next ones are one affine map (32 * stride draws) further.  Output: one 32-byte
// record per history {x, mu, cell | g << 16, first tally row of its generation, rng state, -}, read back by the lane that adopts the history.
#include "mc_lane.cuh"

namespace nraps {

namespace {

constexpr int kRun = 8; // histories per thread

template <int TG, bool BANK>
__global__ void __launch_bounds__(256) source_kernel(const TransportParams P, uint4 *out)
{
    __shared__ ulonglong2 s_jump[64];
    __shared__ unsigned long long s_first[kMaxPeers + 1]; // first global site index of each rank's bank; [n_peers] = bank size
    for (int i = threadIdx.x; i < 64; i += blockDim.x) s_jump[i] = P.jump[i];
    if (BANK && threadIdx.x == 0) {
        // every rank's site count sits in word 0 of its bank buffer: read them where they are (NVLink for the peers)
        unsigned long long acc = 0ull;
        for (uint32_t r = 0; r < P.n_peers; ++r) {
            s_first[r] = acc;
            acc += *reinterpret_cast<const volatile unsigned long long *>(P.peer_bank[r]);
        }
        s_first[P.n_peers] = acc;
    }
    __syncthreads();
    const int G = TG ? TG : (int)P.G, MG = (int)P.M * G;
    const float *chi = P.xs + 2 * MG;
    const uint64_t n = (uint64_t)(P.rows / P.G) * P.hist_shard;
    const unsigned long long src_count = (BANK && P.n_peers) ? s_first[P.n_peers] : 0ull;
    // A warp takes 32 * kRun consecutive histories and lane l the ones at l, l + 32, l + 64, ...: the 32-byte records of
    // one store instruction are consecutive in memory (a thread that owned kRun consecutive histories wrote 256 bytes
    // apart from its neighbour: 32 sectors per store).  The thread's next history is 32 further: one affine map, jump[5].
    const uint64_t tid_global = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t first = (tid_global >> 5) * (32u * kRun) + (tid_global & 31u);
    if (first >= n) return;
    // record i belongs to generation i / hist_shard of the launch and history hist_begin + i % hist_shard of it;
    // its stream starts (generation * hist_total + history) * stride draws into the master stream
    uint32_t gen_local = (uint32_t)(first / P.hist_shard);
    uint64_t in_gen = first % P.hist_shard;
    uint64_t base = jump_ahead(P.rng_state, (uint64_t)gen_local * P.hist_total + P.hist_begin + in_gen, s_jump);
    const ulonglong2 J32 = s_jump[5];
    if (BANK && src_count) {
        // Bank source (never batched: one generation per launch).  No gathered copy of the bank exists: a site index is
        // resolved to (rank, offset) and the 8-byte site is loaded from the rank that banked it -- over NVLink for 7 of
        // 8 histories on a full box.  The loads of the thread's kRun histories are issued together, before anything is
        // done with them: a remote load takes microseconds, and the kernel is bound by how many are in flight.
        unsigned long long site[kRun];
        uint64_t b = base;
#pragma unroll
        for (int j = 0; j < kRun; ++j) {
            uint64_t rng = b;
            b = J32.x * b + J32.y;
            const uint32_t u = pcg32_next(rng, P.rng_inc);
            const unsigned long long idx = ((unsigned long long)u * src_count) >> 32;
            uint32_t r = 0;
            while (r + 1 < P.n_peers && idx >= s_first[r + 1]) ++r;
            site[j] = first + 32u * j < n ? __ldg(P.peer_bank[r] + kBankHeader + (idx - s_first[r])) : 0ull;
        }
#pragma unroll
        for (int j = 0; j < kRun; ++j) {
            const uint64_t i = first + 32u * j;
            if (i >= n) break;
            uint64_t rng = base;
            base = J32.x * base + J32.y;
            (void)pcg32_next(rng, P.rng_inc); // the site draw, already used above
            const int cell = (int)(site[j] >> 32);
            const float x = __uint_as_float((uint32_t)site[j]);
            const float mu = fsub(fmul(2.0f, pcg32_unit(rng, P.rng_inc)), 1.0f);
            const int g = search_cdf_global<TG>(chi + __ldg(P.matid + cell) * G, G, pcg32_unit(rng, P.rng_inc));
            out[2 * i] = make_uint4(__float_as_uint(x), __float_as_uint(mu), (uint32_t)cell | ((uint32_t)g << 16), gen_local * P.G);
            out[2 * i + 1] = make_uint4((uint32_t)rng, (uint32_t)(rng >> 32), 0u, 0u);
        }
        return;
    }
#pragma unroll 1
    for (uint64_t i = first; i < n; ) {
        uint64_t rng = base;
        const uint32_t u = pcg32_next(rng, P.rng_inc);
        const int cell = __ldg(P.fuel + __umulhi(u, P.NF));
        const float xi_pos = pcg32_unit(rng, P.rng_inc);
        const float mu = fsub(fmul(2.0f, pcg32_unit(rng, P.rng_inc)), 1.0f);
        const float x = fadd(__ldg(P.edges + cell), fmul(xi_pos, P.dx_fuel));
        const int g = search_cdf_global<TG>(chi + __ldg(P.matid + cell) * G, G, pcg32_unit(rng, P.rng_inc));
        out[2 * i] = make_uint4(__float_as_uint(x), __float_as_uint(mu), (uint32_t)cell | ((uint32_t)g << 16), gen_local * P.G);
        out[2 * i + 1] = make_uint4((uint32_t)rng, (uint32_t)(rng >> 32), 0u, 0u);
        i += 32u;
        if (i >= first + 32u * kRun) break;
        in_gen += 32u;
        if (in_gen >= P.hist_shard) { // the run crosses into a later generation of the batch
            gen_local = (uint32_t)(i / P.hist_shard);
            in_gen = i % P.hist_shard;
            base = jump_ahead(P.rng_state, (uint64_t)gen_local * P.hist_total + P.hist_begin + in_gen, s_jump);
        } else {
            base = J32.x * base + J32.y; // stream of the history 32 further
        }
    }
}

template <int TG> cudaError_t launch_g(const TransportParams &p, bool bank, uint4 *out, cudaStream_t s)
{
    const uint64_t n = (uint64_t)(p.rows / p.G) * p.hist_shard;
    const unsigned blocks = (unsigned)((n + 256ull * kRun - 1) / (256ull * kRun));
    if (!blocks) return cudaSuccess;
    if (bank) source_kernel<TG, true><<<blocks, 256, 0, s>>>(p, out);
    else source_kernel<TG, false><<<blocks, 256, 0, s>>>(p, out);
    return cudaGetLastError();
}

} // namespace

cudaError_t launch_source(const TransportParams &p, bool bank, uint4 *out, cudaStream_t s)
{
    switch (p.G) {
    case 2: return launch_g<2>(p, bank, out, s);
    case 4: return launch_g<4>(p, bank, out, s);
    default: return launch_g<0>(p, bank, out, s);
    }
}

} // namespace nraps
