// Fused history kernel: one persistent lane per in-flight neutron, refilled
// from a per-warp chunk of history indices the moment its neutron dies.
//
// Replaces, per history (reference src/mc_code.rs):
//   spawn_neutron + energy ........ :7-53, 228-230   (stage SPAWN)
//   particle_travel flight draw ... :147-148, 209    (stage FLIGHT)
//   boundary / cross_mesh loop .... :151-181, 56-79  (stage WALK)
//   scat_mat_calc + interaction ... :82-132, 183-208 (stage COLLIDE)
//   particle_lifetime tally ....... :224, 163/173/184 (shared-memory fixed-point bins)
//
// Warp structure: every trip of the outer loop each live lane draws one flight,
// walks cell crossings until its flight ends (collision, material change, leak),
// then the lanes that collided run the collision stage together.  The walk is
// bounded by the length of a material run (8 fuel / 4 water cells in the
// shipped decks), so lanes of a warp stay within a small factor of each other.
//
// Tallies: score -> (u64)(score * 2^28), added to a 64-bit bin kept as two u32
// words in shared memory (ATOMS.ADD is native for u32 only; f32 and u64 shared
// atomics compile to CAS loops on sm_100a).  Integer sums are associative, so
// the result is independent of lane / block / GPU scheduling and bit-identical
// to the oracle's.
#include "mc_device.cuh"
#include "mc_internal.h"

namespace nraps {

namespace {

constexpr unsigned kFull = 0xffffffffu;

struct Smem {
    uint32_t *lo, *hi;
    const float *edges;
    const uint32_t *runb;
    const ulonglong2 *jump;
    const float *inv_sigtr, *p_abs, *chi_cdf, *scat_cdf;
    const uint16_t *fuel;
    const uint8_t *matid;
};

__device__ __forceinline__ void score(const Smem &S, int bin, float v)
{
    const unsigned long long fx = __float2ull_rz(fmul(v, kTallyScale));
    const uint32_t l = (uint32_t)fx;
    uint32_t h = (uint32_t)(fx >> 32);
    const uint32_t old = atomicAdd(&S.lo[bin], l);
    h += (uint32_t)(old + l < old);
    if (h) atomicAdd(&S.hi[bin], h);
}

template <int TG>
__device__ __forceinline__ int sample_group(const float *cdf, int G, int mode, uint64_t &rng, uint64_t inc)
{
    const int n = TG ? TG : G;
    if (mode == NRAPS_SCATTER_SINGLE_XI) return lower_bound_clamped<TG>(cdf, G, pcg32_unit(rng, inc));
    if (mode == NRAPS_SCATTER_RUST_PRE182) { // a fresh draw per probe, pre-1.82 probe order
        int size = n, left = 0, right = n;
        while (left < right) {
            const int mid = left + size / 2;
            if (cdf[mid] < pcg32_unit(rng, inc)) left = mid + 1;
            else right = mid;
            size = right - left;
        }
        return left < n - 1 ? left : n - 1;
    }
    int size = n, base = 0; // rustc >= 1.82 probe order
    while (size > 1) {
        const int half = size / 2, mid = base + half;
        if (cdf[mid] < pcg32_unit(rng, inc)) base = mid;
        size -= half;
    }
    const int res = base + (cdf[base] < pcg32_unit(rng, inc) ? 1 : 0);
    return res < n - 1 ? res : n - 1;
}

enum { EV_NONE = 0, EV_COLLIDE = 1, EV_MATCHANGE = 2 };

template <int TG, bool TRACE>
__global__ void __launch_bounds__(1024, 1) transport_kernel(const TransportParams P)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int G = TG ? TG : (int)P.G;
    const int M = (int)P.M, N = (int)P.N;
    const SmemLayout L = make_layout(P.M, P.G, P.N, P.NF);

    Smem S;
    S.lo = reinterpret_cast<uint32_t *>(smem_raw + L.tally_lo);
    S.hi = reinterpret_cast<uint32_t *>(smem_raw + L.tally_hi);
    float *w_edges = reinterpret_cast<float *>(smem_raw + L.edges);
    uint32_t *w_runb = reinterpret_cast<uint32_t *>(smem_raw + L.runb);
    ulonglong2 *w_jump = reinterpret_cast<ulonglong2 *>(smem_raw + L.jump);
    float *w_xs = reinterpret_cast<float *>(smem_raw + L.xs);
    uint16_t *w_fuel = reinterpret_cast<uint16_t *>(smem_raw + L.fuel);
    uint8_t *w_matid = smem_raw + L.matid;

    const int tid = threadIdx.x, nthr = blockDim.x;
    const int GN = G * N, MG = M * G;
    for (int i = tid; i < GN; i += nthr) { S.lo[i] = 0u; S.hi[i] = 0u; }
    for (int i = tid; i <= N; i += nthr) w_edges[i] = P.edges[i];
    for (int i = tid; i < N; i += nthr) { w_runb[i] = P.runb[i]; w_matid[i] = P.matid[i]; }
    for (int i = tid; i < (int)P.NF; i += nthr) w_fuel[i] = P.fuel[i];
    for (int i = tid; i < 3 * MG + MG * G * G; i += nthr) w_xs[i] = P.xs[i];
    for (int i = tid; i < 64; i += nthr) w_jump[i] = P.jump[i];
    __syncthreads();
    S.edges = w_edges; S.runb = w_runb; S.jump = w_jump; S.fuel = w_fuel; S.matid = w_matid;
    S.inv_sigtr = w_xs; S.p_abs = w_xs + MG; S.chi_cdf = w_xs + 2 * MG; S.scat_cdf = w_xs + 3 * MG;

    const unsigned lane = tid & 31;
    const uint64_t inc = P.rng_inc;

    // warp-uniform cursor over the chunk of history indices this warp owns
    uint64_t w_next = 0, w_end = 0;
    bool exhausted = false;

    // lane state: one neutron
    bool alive = false;
    uint64_t rng = 0, y = 0;
    float x = 0.f, mu = 1.f, ds = 0.f;
    int cell = 0, g = 0, xsg = 0, mat = 0, run_lo = 0, run_hi = 0;
    uint32_t h_coll = 0, h_cross = 0, h_flight = 0, h_refl = 0; // this history
    uint32_t c_hist = 0, c_coll = 0, c_cross = 0, c_flight = 0, c_refl = 0, c_leak = 0, c_trunc = 0;

    for (;;) {
        // ---------------- SPAWN: hand fresh history indices to dead lanes
        const unsigned need = __ballot_sync(kFull, !alive);
        if (need) {
            if (w_next == w_end && !exhausted) {
                unsigned long long base = 0;
                if (lane == 0) base = atomicAdd(P.work, (unsigned long long)P.chunk);
                base = __shfl_sync(kFull, base, 0);
                const uint64_t b = P.hist_begin + base;
                if (b >= P.hist_end) exhausted = true;
                else { w_next = b; w_end = (b + P.chunk < P.hist_end) ? b + P.chunk : P.hist_end; }
            }
            const uint32_t avail = (uint32_t)(w_end - w_next);
            if (avail) {
                const uint32_t rank = __popc(need & ((1u << lane) - 1u));
                if (!alive && rank < avail) {
                    y = w_next + rank;
                    // per-history stream: master advanced by y*stride draws (jump maps commute)
                    rng = P.rng_state;
                    for (uint64_t h = y; h;) {
                        const int b = __ffsll((long long)h) - 1;
                        h &= h - 1;
                        const ulonglong2 J = S.jump[b];
                        rng = J.x * rng + J.y;
                    }
                    // draw order cell, position, mu, chi (src/mc_code.rs:46-51)
                    const uint32_t u = pcg32_next(rng, inc);
                    cell = S.fuel[__umulhi(u, P.NF)];
                    const float xi_pos = pcg32_unit(rng, inc);
                    mu = fsub(fmul(2.0f, pcg32_unit(rng, inc)), 1.0f);
                    const float xi_chi = pcg32_unit(rng, inc);
                    mat = S.matid[cell];
                    g = lower_bound_clamped<TG>(S.chi_cdf + mat * G, G, xi_chi);
                    xsg = g;
                    x = fadd(S.edges[cell], fmul(xi_pos, P.dx_fuel));
                    const uint32_t rb = S.runb[cell];
                    run_lo = (int)(rb & 0xffffu);
                    run_hi = (int)(rb >> 16);
                    h_coll = h_cross = h_flight = h_refl = 0;
                    alive = true;
                }
                const uint32_t want = __popc(need);
                w_next += want < avail ? want : avail;
            } else if (need == kFull) {
                break; // no work left anywhere and every lane is dead
            }
        }

        if (alive) {
            uint32_t fate = 0;
            int ev = EV_NONE;
            float end = 0.f;
            if (h_flight >= P.max_flights) {
                fate = NRAPS_FATE_TRUNCATED;
            } else {
                // ------------ FLIGHT: signed x-displacement to the next collision
                ds = fmul(fmul(mu, -mc_logf(pcg32_unit(rng, inc))), S.inv_sigtr[mat + M * xsg]);
                ++h_flight;
                // ------------ WALK: cell by cell inside one material run
                bool fwd = mu >= 0.0f;
                for (;;) {
                    end = fadd(x, ds);
                    const float edge = S.edges[cell + (fwd ? 1 : 0)];
                    const float t = fsub(x, edge);
                    const bool at_wall = fwd ? (cell == N - 1 && end > edge) : (cell == 0 && edge > end);
                    if (at_wall) {
                        score(S, g * N + cell, fabsf(fdiv(t, mu)));
                        const float b = fwd ? P.boundr : P.boundl;
                        if (b > 0.0f) { // hit_boundary
                            mu = fmul(mu, -b);
                            ds = fmul(fadd(ds, t), -b);
                            x = edge;
                            fwd = mu >= 0.0f;
                            if (TRACE) ++h_refl;
                            continue;
                        }
                        fate = NRAPS_FATE_LEAKED;
                        break;
                    }
                    if (fabsf(fsub(end, x)) > fabsf(t)) { // cross_mesh; |edge - x| == |x - edge|
                        score(S, g * N + cell, fabsf(fdiv(t, mu)));
                        ds = fadd(ds, t);
                        x = edge;
                        cell += fwd ? 1 : -1;
                        if (TRACE) ++h_cross;
                        if (cell < run_lo || cell >= run_hi) { ev = EV_MATCHANGE; break; }
                    } else {
                        ev = EV_COLLIDE;
                        break;
                    }
                }
            }

            // ---------------- COLLIDE / material change
            if (ev == EV_COLLIDE) {
                score(S, g * N + cell, fabsf(fdiv(fsub(x, end), mu)));
                ++h_coll;
                const int xs = mat + M * xsg; // stale group index, src/mc_code.rs:147
                const float xi_int = pcg32_unit(rng, inc);
                const float mu_new = fsub(fmul(2.0f, pcg32_unit(rng, inc)), 1.0f);
                const int g_new = sample_group<TG>(S.scat_cdf + ((mat * G + g) * G + xsg) * G, G, P.scatter_mode, rng, inc);
                if (xi_int < S.p_abs[xs]) {
                    fate = NRAPS_FATE_ABSORBED;
                } else {
                    x = end;
                    g = g_new;
                    mu = mu_new;
                    if (!P.stale_xs) xsg = g;
                }
            } else if (ev == EV_MATCHANGE) {
                if ((unsigned)cell >= (unsigned)N) {
                    fate = NRAPS_FATE_TRUNCATED; // unreachable for validated input
                    cell = cell < 0 ? 0 : N - 1;
                } else {
                    mat = S.matid[cell];
                    xsg = g;
                    const uint32_t rb = S.runb[cell];
                    run_lo = (int)(rb & 0xffffu);
                    run_hi = (int)(rb >> 16);
                }
            }

            if (fate) {
                alive = false;
                ++c_hist;
                c_coll += h_coll;
                c_flight += h_flight;
                c_leak += (fate == NRAPS_FATE_LEAKED);
                c_trunc += (fate == NRAPS_FATE_TRUNCATED);
                if (TRACE) {
                    c_cross += h_cross; c_refl += h_refl;
                    if (P.trace) {
                        uint32_t *t = P.trace + (y - P.hist_begin) * NRAPS_TR_WORDS;
                        t[NRAPS_TR_COLLISIONS] = h_coll;
                        t[NRAPS_TR_CROSSINGS] = h_cross;
                        t[NRAPS_TR_FLIGHTS] = h_flight;
                        t[NRAPS_TR_REFLECTIONS] = h_refl;
                        t[NRAPS_TR_RNG_LO] = (uint32_t)rng;
                        t[NRAPS_TR_RNG_HI] = (uint32_t)(rng >> 32);
                        t[NRAPS_TR_CELL] = (uint32_t)cell;
                        t[NRAPS_TR_XBITS] = __float_as_uint(x);
                        t[NRAPS_TR_FATE] = fate;
                        t[NRAPS_TR_GROUP] = (uint32_t)g;
                    }
                }
            }
        }
    }

    // ---------------- flush: block bins -> global 64-bit bins, lane counters -> global
    __syncthreads();
    for (int i = tid; i < GN; i += nthr) {
        const unsigned long long v = ((unsigned long long)S.hi[i] << 32) + S.lo[i];
        if (v) atomicAdd(&P.tally[i], v);
    }
    unsigned long long *ct = P.tally + GN;
    uint32_t vals[7] = {c_hist, c_coll, c_cross, c_flight, c_refl, c_leak, c_trunc};
#pragma unroll
    for (int c = 0; c < 7; ++c) {
        unsigned long long v = vals[c];
#pragma unroll
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
        if (lane == 0 && v) atomicAdd(&ct[c], v);
    }
}

template <int TG>
cudaError_t launch_g(const TransportParams &p, bool trace, dim3 grid, dim3 block, uint32_t smem, cudaStream_t s)
{
    if (trace) transport_kernel<TG, true><<<grid, block, smem, s>>>(p);
    else transport_kernel<TG, false><<<grid, block, smem, s>>>(p);
    return cudaGetLastError();
}

template <int TG, bool TRACE> cudaError_t set_smem(uint32_t bytes)
{
    return cudaFuncSetAttribute(transport_kernel<TG, TRACE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

} // namespace

cudaError_t prepare_transport(uint32_t smem_bytes)
{
    cudaError_t e;
    if ((e = set_smem<2, false>(smem_bytes)) != cudaSuccess) return e;
    if ((e = set_smem<2, true>(smem_bytes)) != cudaSuccess) return e;
    if ((e = set_smem<4, false>(smem_bytes)) != cudaSuccess) return e;
    if ((e = set_smem<4, true>(smem_bytes)) != cudaSuccess) return e;
    if ((e = set_smem<0, false>(smem_bytes)) != cudaSuccess) return e;
    return set_smem<0, true>(smem_bytes);
}

cudaError_t launch_transport(const TransportParams &p, bool trace, dim3 grid, dim3 block, uint32_t smem, cudaStream_t s)
{
    switch (p.G) {
    case 2: return launch_g<2>(p, trace, grid, block, smem, s);
    case 4: return launch_g<4>(p, trace, grid, block, smem, s);
    default: return launch_g<0>(p, trace, grid, block, smem, s);
    }
}

} // namespace nraps
