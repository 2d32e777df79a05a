// Event-based pipeline over a structure-of-arrays particle bank in HBM: the
// layout the north star prescribes (source -> loop { advance, collide, compact }),
// built as an alternative to the fused persistent kernels so the choice between
// them rests on measurements (profiles/).  Woodcock physics
// only: one "event" = one tentative collision, so every thread of a launch does
// the same work.  Each record carries its own PCG32 state and the tallies are
// integer sums, hence the result is bit-identical to woodcock_kernel's.
//
//   ev_source   : history y -> record (replaces spawn_neutron, src/mc_code.rs:40-53)
//   ev_advance  : flight + walls + cell lookup + collision-estimator score + accept draw
//   ev_collide  : real collisions (src/mc_code.rs:183-209); absorbed records die
//   ev_compact  : survivors -> the other bank half (warp ballot + one atomic per warp)
#include <algorithm>

#include "mc_lane.cuh"

namespace nraps {

namespace {

// pack: cell[0:16) | g[16:20) | xsg[20:24) | left[24] | accepted[25] | alive[26]
__device__ __forceinline__ uint32_t pack_state(int cell, int g, int xsg, bool left, bool accepted, bool alive)
{
    return (uint32_t)cell | ((uint32_t)g << 16) | ((uint32_t)xsg << 20) | ((uint32_t)left << 24) | ((uint32_t)accepted << 25) |
           ((uint32_t)alive << 26);
}

struct Tables {
    const float *edges, *inv_sigtr, *p_abs, *chi, *nusigf, *sigtr, *scat, *inv_maj;
    const uint32_t *runb;
    const ulonglong2 *jump;
    const uint16_t *fuel, *bucket;
    const uint8_t *matid;
};

__device__ __forceinline__ Tables tables_of(const SmemView &S, int MG, int G)
{
    Tables T;
    T.edges = S.edges; T.runb = S.runb; T.jump = S.jump; T.fuel = S.fuel; T.bucket = S.bucket; T.matid = S.matid;
    T.inv_sigtr = S.xs; T.p_abs = S.xs + MG; T.chi = S.xs + 2 * MG; T.nusigf = S.xs + 3 * MG; T.sigtr = S.xs + 4 * MG;
    T.scat = S.xs + 5 * MG; T.inv_maj = S.xs + 5 * MG + MG * G * G;
    return T;
}

// per-launch death bookkeeping: warp-reduced into the global counters
__device__ __forceinline__ void count_deaths(const TransportParams &P, uint32_t hist, uint32_t coll, uint32_t flight, uint32_t refl,
                                             uint32_t leak, uint32_t trunc)
{
    unsigned long long *ct = P.tally + (size_t)P.rows * P.N;
    const uint32_t vals[8] = {hist, coll, 0u, flight, refl, leak, trunc, 0u};
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        unsigned long long v = vals[c];
#pragma unroll
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
        if ((threadIdx.x & 31) == 0 && v) atomicAdd(&ct[c], v);
    }
}

template <int TG>
__global__ void __launch_bounds__(512) ev_source(const TransportParams P, const EventHalf A)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int G = TG ? TG : (int)P.G, MG = (int)P.M * G;
    const SmemView S = load_block_tables(smem_raw, P, make_layout(P.M, P.G, P.N, P.NF, P.NB, 0, P.rows));
    const Tables T = tables_of(S, MG, G);
    const uint64_t n = P.hist_end - P.hist_begin;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t rng = jump_ahead(P.rng_state, P.hist_begin + i, T.jump);
        const uint32_t u = pcg32_next(rng, P.rng_inc);
        const int cell = T.fuel[__umulhi(u, P.NF)];
        const float xi_pos = pcg32_unit(rng, P.rng_inc);
        const float mu = fsub(fmul(2.0f, pcg32_unit(rng, P.rng_inc)), 1.0f);
        const int g = search_cdf<TG>(T.chi + T.matid[cell] * G, G, pcg32_unit(rng, P.rng_inc));
        A.x[i] = fadd(T.edges[cell], fmul(xi_pos, P.dx_fuel));
        A.mu[i] = mu;
        A.pack[i] = pack_state(cell, g, g, false, false, true);
        A.rng[i] = rng;
        A.cnt[i] = 0u; // flights[0:20) | reflections[20:32); collisions live in ccnt
        A.ccnt[i] = 0u;
    }
}

template <int TG>
__global__ void __launch_bounds__(512) ev_advance(const TransportParams P, const EventHalf A, const unsigned long long *n_alive)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int G = TG ? TG : (int)P.G, M = (int)P.M, N = (int)P.N, NB = (int)P.NB, MG = M * G;
    const SmemLayout L = make_layout(P.M, P.G, P.N, P.NF, P.NB, 0, P.rows);
    const SmemView S = load_block_tables(smem_raw, P, L);
    const Tables T = tables_of(S, MG, G);
    const uint32_t lo_base = (uint32_t)__cvta_generic_to_shared(S.lo), hi_off = L.tally_hi - L.tally_lo;
    const float len = T.edges[N];
    const uint64_t n = *n_alive;
    uint32_t d_hist = 0, d_coll = 0, d_flight = 0, d_refl = 0, d_leak = 0, d_trunc = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        float x = A.x[i], mu = A.mu[i];
        const uint32_t pk = A.pack[i];
        uint64_t rng = A.rng[i];
        uint32_t cnt = A.cnt[i];
        int cell = (int)(pk & 0xffffu);
        const int g = (int)((pk >> 16) & 15u), xsg = (int)((pk >> 20) & 15u);
        bool left = (pk >> 24) & 1u;
        uint32_t fate = 0;
        bool accepted = false;
        if ((cnt & 0xfffffu) >= P.max_flights) {
            fate = NRAPS_FATE_TRUNCATED;
        } else {
            const uint32_t rb = T.runb[cell]; // while !left the neutron is inside its home run
            const int home_lo = (int)(rb & 0xffffu), home_hi = (int)(rb >> 16);
            const float inv_maj = T.inv_maj[xsg * G + g];
            float xn = fadd(x, fmul(fmul(mu, -mc_logf(pcg32_unit(rng, P.rng_inc))), inv_maj));
            cnt += 1u;
            while (xn < 0.0f || xn > len) {
                const bool lo_wall = xn < 0.0f;
                const float wall = lo_wall ? 0.0f : len, b = lo_wall ? P.boundl : P.boundr;
                if (!(b > 0.0f)) { fate = NRAPS_FATE_LEAKED; break; }
                const float rem = fsub(xn, wall);
                mu = fmul(mu, -b);
                xn = fadd(wall, fmul(rem, -b));
                if (!left && (lo_wall ? (home_lo != 0) : (home_hi != N))) left = true;
                cnt += 1u << 20;
            }
            if (!fate) {
                int c = __float2int_rz(fmul(xn, P.inv_h));
                c = T.bucket[c < NB - 1 ? c : NB - 1];
                while (c < N - 1 && T.edges[c + 1] <= xn) ++c;
                while (c > 0 && T.edges[c] > xn) --c;
                left = left || c < home_lo || c >= home_hi;
                cell = c;
                x = xn;
                score<false>(lo_base + 4u * (uint32_t)(g * N + cell), hi_off, inv_maj, nullptr);
                const int g_eff = left ? g : xsg;
                accepted = pcg32_unit(rng, P.rng_inc) < fmul(T.sigtr[T.matid[cell] + M * g_eff], inv_maj);
            }
        }
        if (fate) {
            ++d_hist; d_coll += A.ccnt[i]; d_flight += cnt & 0xfffffu; d_refl += cnt >> 20;
            d_leak += fate == NRAPS_FATE_LEAKED; d_trunc += fate == NRAPS_FATE_TRUNCATED;
        }
        A.x[i] = x;
        A.mu[i] = mu;
        A.pack[i] = pack_state(cell, g, xsg, left, accepted, fate == 0);
        A.rng[i] = rng;
        A.cnt[i] = cnt;
    }
    const uint32_t zero[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    flush_block(S, P, zero);
    count_deaths(P, d_hist, d_coll, d_flight, d_refl, d_leak, d_trunc);
}

template <int TG>
__global__ void __launch_bounds__(512) ev_collide(const TransportParams P, const EventHalf A, const unsigned long long *n_alive)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int G = TG ? TG : (int)P.G, M = (int)P.M, MG = M * G;
    const SmemView S = load_block_tables(smem_raw, P, make_layout(P.M, P.G, P.N, P.NF, P.NB, 0, P.rows));
    const Tables T = tables_of(S, MG, G);
    const uint64_t n = *n_alive;
    uint32_t d_hist = 0, d_coll = 0, d_flight = 0, d_refl = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t pk = A.pack[i];
        if (!((pk >> 25) & 1u)) continue; // virtual collision or dead
        uint64_t rng = A.rng[i];
        const int cell = (int)(pk & 0xffffu), g = (int)((pk >> 16) & 15u), xsg = (int)((pk >> 20) & 15u);
        const bool left = (pk >> 24) & 1u;
        const int mat = T.matid[cell], g_eff = left ? g : xsg, xs = mat + M * g_eff;
        const uint32_t ncoll = A.ccnt[i] + 1u;
        const float xi_int = pcg32_unit(rng, P.rng_inc);
        const float mu_new = fsub(fmul(2.0f, pcg32_unit(rng, P.rng_inc)), 1.0f);
        const int g_new = sample_group<TG>(T.scat + ((mat * G + g) * G + g_eff) * G, G, P.scatter_mode, rng, P.rng_inc);
        A.rng[i] = rng;
        A.ccnt[i] = ncoll;
        if (xi_int < T.p_abs[xs]) {
            const uint32_t cnt = A.cnt[i];
            ++d_hist; d_coll += ncoll; d_flight += cnt & 0xfffffu; d_refl += cnt >> 20;
            A.pack[i] = pack_state(cell, g, xsg, left, false, false);
        } else {
            A.mu[i] = mu_new;
            A.pack[i] = pack_state(cell, g_new, P.stale_xs ? g_eff : g_new, false, false, true);
        }
    }
    count_deaths(P, d_hist, d_coll, d_flight, d_refl, 0u, 0u);
}

__global__ void __launch_bounds__(512) ev_compact(const EventHalf A, const EventHalf Z, const unsigned long long *n_alive,
                                                  unsigned long long *n_next)
{
    const uint64_t n = *n_alive;
    const unsigned lane = threadIdx.x & 31;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t base = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x - lane; base < n; base += stride) {
        const uint64_t i = base + lane;
        const bool keep = i < n && ((A.pack[i] >> 26) & 1u);
        const unsigned m = __ballot_sync(kFull, keep);
        if (!m) continue;
        unsigned long long dst = 0;
        if (lane == 0) dst = atomicAdd(n_next, (unsigned long long)__popc(m));
        dst = __shfl_sync(kFull, dst, 0) + __popc(m & ((1u << lane) - 1u));
        if (keep) {
            Z.x[dst] = A.x[i];
            Z.mu[dst] = A.mu[i];
            Z.pack[dst] = A.pack[i];
            Z.rng[dst] = A.rng[i];
            Z.cnt[dst] = A.cnt[i];
            Z.ccnt[dst] = A.ccnt[i];
        }
    }
}

template <int TG> cudaError_t run_g(const TransportParams &P, const EventBank &B, uint32_t smem, int sm_count, cudaStream_t s, uint32_t *iters)
{
    cudaError_t e;
    const auto attr = cudaFuncAttributeMaxDynamicSharedMemorySize;
    if ((e = cudaFuncSetAttribute(ev_source<TG>, attr, (int)smem)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(ev_advance<TG>, attr, (int)smem)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(ev_collide<TG>, attr, (int)smem)) != cudaSuccess) return e;
    unsigned long long n = P.hist_end - P.hist_begin;
    auto grid_for = [&](unsigned long long m) { return (unsigned)std::max<unsigned long long>(1, std::min<unsigned long long>((m + 2047) / 2048, (unsigned long long)sm_count * 4)); };
    if ((e = cudaMemcpyAsync(B.n_alive, &n, sizeof(n), cudaMemcpyHostToDevice, s)) != cudaSuccess) return e;
    ev_source<TG><<<grid_for(n), 512, smem, s>>>(P, B.half[0]);
    int cur = 0;
    uint32_t it = 0;
    while (n) {
        const unsigned g = grid_for(n);
        ev_advance<TG><<<g, 512, smem, s>>>(P, B.half[cur], B.n_alive);
        ev_collide<TG><<<g, 512, smem, s>>>(P, B.half[cur], B.n_alive);
        if ((e = cudaMemsetAsync(B.n_next, 0, sizeof(unsigned long long), s)) != cudaSuccess) return e;
        ev_compact<<<g, 512, 0, s>>>(B.half[cur], B.half[cur ^ 1], B.n_alive, B.n_next);
        if ((e = cudaMemcpyAsync(B.n_alive, B.n_next, sizeof(n), cudaMemcpyDeviceToDevice, s)) != cudaSuccess) return e;
        if ((e = cudaMemcpyAsync(&n, B.n_next, sizeof(n), cudaMemcpyDeviceToHost, s)) != cudaSuccess) return e;
        if ((e = cudaStreamSynchronize(s)) != cudaSuccess) return e; // the host sizes the next launches
        cur ^= 1;
        ++it;
    }
    *iters = it;
    return cudaGetLastError();
}

} // namespace

cudaError_t run_event_generation(const TransportParams &P, const EventBank &B, uint32_t smem, int sm_count, cudaStream_t s, uint32_t *iters)
{
    switch (P.G) {
    case 2: return run_g<2>(P, B, smem, sm_count, s, iters);
    case 4: return run_g<4>(P, B, smem, sm_count, s, iters);
    default: return run_g<0>(P, B, smem, sm_count, s, iters);
    }
}

} // namespace nraps
