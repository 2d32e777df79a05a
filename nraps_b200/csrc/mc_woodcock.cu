// Woodcock delta tracking with a collision-estimator tally: the tracking the
// north star names.  No surface crossings: a flight is sampled against the
// majorant of the two energy groups in play, every tentative collision scores
// 1/Sigma_maj into the cell it lands in (unbiased track-length estimate) and is
// accepted as a real collision with probability sigtr/Sigma_maj.  Cost per
// history is independent of the mesh (36 tentative collisions on deck C against
// 364 cell crossings; 3642 on the fine mesh) and every lane of a warp executes
// the same short loop body, so SIMT efficiency is high by construction.
//
// Same source, collision physics, draw order inside a collision, stale
// cross-section group (the group index set when a material run is
// entered is kept until the neutron leaves that run) and bank rules as the
// surface kernel; statistically equivalent to the reference, bit-identical to
// the oracle's Woodcock mode (oracle/oracle_mc.c: run_history_woodcock).
#include "mc_lane.cuh"

namespace nraps {

namespace {

// Register budget per instantiation (see RegCap in mc_transport.cu): blocks of 1024 threads must be launchable (64);
// the production instantiations are asked for NRAPS_WREG registers so that two blocks of 640 threads share an SM.
#ifndef NRAPS_WREG
#define NRAPS_WREG 48
#endif
template <bool TRACE, bool BIG> struct WoodcockRegCap { static constexpr int k = (!TRACE && !BIG) ? NRAPS_WREG : 64; };

template <int TG, bool TRACE, bool BANK, bool BIG>
__global__ void __maxnreg__((WoodcockRegCap<TRACE, BIG>::k)) woodcock_kernel(const TransportParams P)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int G = TG ? TG : (int)P.G;
    const int M = (int)P.M, N = (int)P.N, NB = (int)P.NB;
    const SmemLayout L = make_layout(P.M, P.G, P.N, P.NF, P.NB, BIG, P.rows);
    const SmemView S = load_block_tables(smem_raw, P, L);
    const float *s_edges = S.edges, *s_xs = S.xs;
    const MeshRef<BIG> mesh(S);
    const int tid = threadIdx.x, MG = M * G;
    const float *s_inv_sigtr = s_xs, *s_p_abs = s_xs + MG, *s_nusigf = s_xs + 3 * MG,
                *s_sigtr = s_xs + 4 * MG, *s_scat = s_xs + 5 * MG, *s_inv_maj = s_xs + 5 * MG + MG * G * G;
    const float inv_k = BANK ? fdiv(1.0f, *P.k_cur) : 1.0f;
    const uint32_t lo_base = BIG ? 0u : (uint32_t)__cvta_generic_to_shared(S.lo);
    const uint32_t hi_off = L.tally_hi - L.tally_lo;
    const float len = mesh.edge(N);

    const unsigned lane = tid & 31;
    const uint64_t inc = P.rng_inc;
    uint64_t w_next = 0, w_end = 0;
    bool exhausted = false;

    bool alive = false, left = false;
    uint64_t rng = 0, y = 0;
    float x = 0.f, mu = 1.f;
    int cell = 0, g = 0, xsg = 0, home_lo = 0, home_hi = 0, row0 = 0;
    uint32_t h_coll = 0, h_flight = 0, h_refl = 0, h_bank = 0;
    uint32_t c_hist = 0, c_coll = 0, c_flight = 0, c_refl = 0, c_leak = 0, c_trunc = 0, c_bank = 0;

    for (;;) {
        __syncwarp();
        // ---------------- SPAWN (identical to the surface kernel)
        const unsigned need = __ballot_sync(kFull, !alive);
        if (need && ((uint32_t)__popc(need) >= P.spawn_batch || need == kFull)) {
            if (w_next == w_end && !exhausted) {
                unsigned long long base = 0;
                if (lane == 0) base = atomicAdd(P.work, (unsigned long long)P.chunk);
                base = __shfl_sync(kFull, base, 0);
                const uint64_t b = P.hist_begin + base;
                if (b >= P.hist_end) exhausted = true;
                else {
                    w_next = b;
                    w_end = (b + P.chunk < P.hist_end) ? b + P.chunk : P.hist_end;
                }
            }
            const uint32_t avail = (uint32_t)(w_end - w_next);
            if (avail) {
                const uint32_t rank = __popc(need & ((1u << lane) - 1u));
                if (!alive && rank < avail) {
                    y = w_next + rank;
                    // adopt the neutron source_kernel gave birth to (mc_source.cu)
                    const uint4 *rec = P.source + 2 * (y - P.hist_begin);
                    const uint4 r0 = __ldg(rec), r1 = __ldg(rec + 1);
                    x = __uint_as_float(r0.x);
                    mu = __uint_as_float(r0.y);
                    cell = (int)(r0.z & 0xffffu);
                    g = (int)(r0.z >> 16);
                    row0 = (int)r0.w; // first tally row of this history's generation (0 unless generations are batched)
                    rng = (uint64_t)r1.x | ((uint64_t)r1.y << 32);
                    xsg = g;
                    const uint32_t rb = mesh.run_bounds(cell);
                    home_lo = (int)(rb & 0xffffu);
                    home_hi = (int)(rb >> 16);
                    left = false;
                    h_coll = h_flight = h_refl = h_bank = 0;
                    alive = true;
                }
                const uint32_t want = __popc(need);
                const uint32_t took = want < avail ? want : avail;
                w_next += took;
            } else if (need == kFull) {
                break;
            }
        }
        __syncwarp();

        // ---------------- FLIGHT to the next tentative collision
        uint32_t fate = 0;
        bool accepted = false, moved = false;
        int mat = 0, g_eff = 0;
        float xn = 0.f, inv_maj = 0.f;
        if (alive) {
            if (h_flight >= P.max_flights) {
                fate = NRAPS_FATE_TRUNCATED;
            } else {
                inv_maj = s_inv_maj[xsg * G + g];
                xn = fadd(x, fmul(fmul(mu, -mc_logf(pcg32_unit(rng, inc))), inv_maj));
                ++h_flight;
                moved = true;
                while (xn < 0.0f || xn > len) { // albedo walls; ~0.9 per history
                    const bool lo_wall = xn < 0.0f;
                    const float wall = lo_wall ? 0.0f : len, b = lo_wall ? P.boundl : P.boundr;
                    if (!(b > 0.0f)) { fate = NRAPS_FATE_LEAKED; moved = false; break; }
                    const float rem = fsub(xn, wall);
                    mu = fmul(mu, -b);
                    xn = fadd(wall, fmul(rem, -b));
                    if (lo_wall ? (home_lo != 0) : (home_hi != N)) left = true;
                    ++h_refl;
                }
            }
        }
        __syncwarp(); // the lane that bounced off a wall rejoins before the common part (profiles/r1d: it ran twice)
        if (moved) {
            // cell containing xn: bucket guess (a bucket is no wider than a cell, so at most one step right),
            // then an exact check against the edges with a rarely-taken repair path
            int c = __float2int_rz(fmul(xn, P.inv_h));
            c = mesh.bucket_cell(c < NB - 1 ? c : NB - 1);
            c += (c < N - 1 && mesh.edge(c + 1) <= xn) ? 1 : 0;
            if (xn < mesh.edge(c) || (xn >= mesh.edge(c + 1) && c < N - 1)) {
                while (c < N - 1 && mesh.edge(c + 1) <= xn) ++c;
                while (c > 0 && mesh.edge(c) > xn) --c;
            }
            cell = c;
            x = xn;
            left = left || cell < home_lo || cell >= home_hi;
            mat = mesh.material(cell);
            g_eff = left ? g : xsg;
            score<BIG>(tally_ref<BIG>(lo_base, (row0 + g) * N + cell), hi_off, inv_maj, P.tally);
            accepted = pcg32_unit(rng, inc) < fmul(s_sigtr[mat + M * g_eff], inv_maj);
        }
        __syncwarp();

        // ---------------- COLLIDE: real collisions only (src/mc_code.rs:183-209)
        if (accepted) {
            ++h_coll;
            const int xs = mat + M * g_eff;
            const float xi_int = pcg32_unit(rng, inc);
            const float mu_new = fsub(fmul(2.0f, pcg32_unit(rng, inc)), 1.0f);
            const int g_new = sample_group<TG>(s_scat + ((mat * G + g) * G + g_eff) * G, G, P.scatter_mode, rng, inc);
            if (BANK) {
                const float nusigf = s_nusigf[mat + M * g];
                if (nusigf > 0.0f) {
                    const float wgt = fmul(fmul(nusigf, s_inv_sigtr[xs]), inv_k);
                    const uint32_t n = (uint32_t)__float2int_rz(fadd(wgt, pcg32_unit(rng, inc)));
                    const unsigned long long site = ((unsigned long long)(uint32_t)cell << 32) | __float_as_uint(x);
                    if (n) { // a history's sites fill its slot row in order; what does not fit is counted, not kept
                        const uint32_t have = h_bank < P.bank_cap ? h_bank : P.bank_cap;
                        uint32_t fit = P.bank_cap - have;
                        fit = n < fit ? n : fit;
                        unsigned long long *dst = P.slots + ((y - P.hist_begin) * P.bank_cap + have);
#pragma unroll 1
                        for (; fit; --fit) *dst++ = site;
                        h_bank += n;
                    }
                }
            }
            if (xi_int < s_p_abs[xs]) {
                fate = NRAPS_FATE_ABSORBED;
            } else {
                g = g_new;
                mu = mu_new;
                xsg = P.stale_xs ? g_eff : g;
                const uint32_t rb = mesh.run_bounds(cell);
                home_lo = (int)(rb & 0xffffu);
                home_hi = (int)(rb >> 16);
                left = false;
            }
        }

        if (fate) {
            alive = false;
            ++c_hist;
            c_coll += h_coll;
            c_flight += h_flight;
            c_refl += h_refl;
            c_leak += (fate == NRAPS_FATE_LEAKED);
            c_trunc += (fate == NRAPS_FATE_TRUNCATED);
            if (BANK) {
                const uint32_t kept = h_bank < P.bank_cap ? h_bank : P.bank_cap;
                P.counts[y - P.hist_begin] = (uint8_t)kept;
                c_bank += kept;
            }
            if (TRACE && P.trace) {
                uint32_t *t = P.trace + (y - P.hist_begin) * NRAPS_TR_WORDS;
                t[NRAPS_TR_COLLISIONS] = h_coll;
                t[NRAPS_TR_CROSSINGS] = 0u;
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

    const uint32_t vals[8] = {c_hist, c_coll, 0u, c_flight, c_refl, c_leak, c_trunc, c_bank};
    flush_block(S, P, vals);
}

template <int TG, bool BIG>
cudaError_t launch_gb(const TransportParams &p, bool trace, bool bank, dim3 grid, dim3 block, uint32_t smem, cudaStream_t s)
{
    if (bank) {
        if (trace) woodcock_kernel<TG, true, true, BIG><<<grid, block, smem, s>>>(p);
        else woodcock_kernel<TG, false, true, BIG><<<grid, block, smem, s>>>(p);
    } else {
        if (trace) woodcock_kernel<TG, true, false, BIG><<<grid, block, smem, s>>>(p);
        else woodcock_kernel<TG, false, false, BIG><<<grid, block, smem, s>>>(p);
    }
    return cudaGetLastError();
}

template <int TG>
cudaError_t launch_g(const TransportParams &p, bool trace, bool bank, dim3 grid, dim3 block, uint32_t smem, cudaStream_t s)
{
    return p.big ? launch_gb<TG, true>(p, trace, bank, grid, block, smem, s) : launch_gb<TG, false>(p, trace, bank, grid, block, smem, s);
}

template <int TG> cudaError_t set_smem(uint32_t bytes, bool trace, bool bank)
{
    const auto attr = cudaFuncAttributeMaxDynamicSharedMemorySize;
    if (bank) return trace ? cudaFuncSetAttribute(woodcock_kernel<TG, true, true, false>, attr, (int)bytes)
                           : cudaFuncSetAttribute(woodcock_kernel<TG, false, true, false>, attr, (int)bytes);
    return trace ? cudaFuncSetAttribute(woodcock_kernel<TG, true, false, false>, attr, (int)bytes)
                 : cudaFuncSetAttribute(woodcock_kernel<TG, false, false, false>, attr, (int)bytes);
}

} // namespace

cudaError_t prepare_woodcock(uint32_t smem_bytes, uint32_t G, bool trace, bool bank)
{
    switch (G) {
    case 2: return set_smem<2>(smem_bytes, trace, bank);
    case 4: return set_smem<4>(smem_bytes, trace, bank);
    default: return set_smem<0>(smem_bytes, trace, bank);
    }
}

cudaError_t launch_woodcock(const TransportParams &p, bool trace, bool bank, dim3 grid, dim3 block, uint32_t smem, cudaStream_t s)
{
    switch (p.G) {
    case 2: return launch_g<2>(p, trace, bank, grid, block, smem, s);
    case 4: return launch_g<4>(p, trace, bank, grid, block, smem, s);
    default: return launch_g<0>(p, trace, bank, grid, block, smem, s);
    }
}


namespace {
template <int TG, bool BIG> int occ_gb(bool trace, bool bank, int block, uint32_t smem)
{
    int n = 0;
    cudaError_t e;
    if (bank) e = trace ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, woodcock_kernel<TG, true, true, BIG>, block, smem)
                        : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, woodcock_kernel<TG, false, true, BIG>, block, smem);
    else e = trace ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, woodcock_kernel<TG, true, false, BIG>, block, smem)
                   : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, woodcock_kernel<TG, false, false, BIG>, block, smem);
    return e == cudaSuccess ? n : 0;
}
template <int TG> int occ_g(bool big, bool trace, bool bank, int block, uint32_t smem)
{
    return big ? occ_gb<TG, true>(trace, bank, block, smem) : occ_gb<TG, false>(trace, bank, block, smem);
}
} // namespace

// resident blocks per SM of the instantiation that would be launched (registers and shared memory both count)
int occupancy_woodcock(uint32_t G, bool big, bool trace, bool bank, int block, uint32_t smem)
{
    switch (G) {
    case 2: return occ_g<2>(big, trace, bank, block, smem);
    case 4: return occ_g<4>(big, trace, bank, block, smem);
    default: return occ_g<0>(big, trace, bank, block, smem);
    }
}

} // namespace nraps
