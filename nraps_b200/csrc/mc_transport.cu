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
// then the warp reconverges (__syncwarp) and the lanes that collided run the
// collision stage together.  The walk is bounded by the length of a material
// run (8 fuel / 4 water cells in the shipped decks), so lanes of a warp stay
// within a small factor of each other.  profiles/r1a_* is the ncu evidence that
// drove this shape: without the explicit reconvergence points the compiler let
// lanes fall out of the walk loop one by one (7.8 of 32 lanes active).
//
// Tallies: score -> (u64)(score * 2^28), added to a 64-bit bin kept as two u32
// words in shared memory (ATOMS.ADD is native for u32 only; f32 and u64 shared
// atomics compile to CAS loops on sm_100a).  Integer sums are associative, so
// the result is independent of lane / block / GPU scheduling and bit-identical
// to the oracle's.
#include "mc_lane.cuh"

namespace nraps {

namespace {

enum { EV_NONE = 0, EV_COLLIDE = 1, EV_MATCHANGE = 2 };

template <int TG, bool TRACE, bool BANK, bool BIG>
__global__ void __launch_bounds__(1024, 1) transport_kernel(const TransportParams P)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int G = TG ? TG : (int)P.G;
    const int M = (int)P.M, N = (int)P.N;
    const SmemLayout L = make_layout(P.M, P.G, P.N, P.NF, P.NB, BIG, P.rows);

    const SmemView S = load_block_tables(smem_raw, P, L);
    uint32_t *s_lo = S.lo;
    const float *s_edges = S.edges, *s_xs = S.xs;
    const MeshRef<BIG> mesh(S);
    const int tid = threadIdx.x;
    const int MG = M * G;
    const float *s_inv_sigtr = s_xs, *s_p_abs = s_xs + MG, *s_nusigf = s_xs + 3 * MG,
                *s_scat = s_xs + 5 * MG;
    // fission_bank mode: sites are banked with weight nu*Sigma_f * inv_sigtr / k_prev; an empty bank => uniform source
    const float inv_k = BANK ? fdiv(1.0f, *P.k_cur) : 1.0f;
    const uint32_t lo_base = BIG ? 0u : (uint32_t)__cvta_generic_to_shared(s_lo);
    const uint32_t hi_off = L.tally_hi - L.tally_lo;
    // edge reference: shared byte address (running pointer of the walk) or, in BIG mode, the edge index
    const uint32_t edges_base = BIG ? 0u : (uint32_t)__cvta_generic_to_shared(s_edges);
    constexpr int kStep = BIG ? 1 : 4;

    const unsigned lane = tid & 31;
    const uint64_t inc = P.rng_inc;

    // warp-uniform cursor over the chunk of history indices this warp owns, and
    // the master stream positioned at history w_next
    uint64_t w_next = 0, w_end = 0;
    bool exhausted = false;

    // lane state: one neutron
    bool alive = false, pending = false;
    uint64_t rng = 0, y = 0;
    float x = 0.f, mu = 1.f, ds = 0.f;
    int cell = 0, g = 0, xsg = 0, mat = 0, run_lo = 0, run_hi = 0, row0 = 0;
    uint32_t h_coll = 0, h_cross = 0, h_flight = 0, h_refl = 0, h_bank = 0; // this history
    uint32_t c_hist = 0, c_coll = 0, c_cross = 0, c_flight = 0, c_refl = 0, c_leak = 0, c_trunc = 0, c_bank = 0;

    for (;;) {
        __syncwarp();
        // ---------------- SPAWN: hand fresh history indices to dead lanes
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
                    mat = mesh.material(cell);
                    xsg = g;
                    h_bank = 0;
                    const uint32_t rb = mesh.run_bounds(cell);
                    run_lo = (int)(rb & 0xffffu);
                    run_hi = (int)(rb >> 16);
                    h_coll = h_cross = h_flight = h_refl = 0;
                    alive = true;
                }
                const uint32_t want = __popc(need);
                const uint32_t took = want < avail ? want : avail;
                w_next += took;
            } else if (need == kFull) {
                break; // no work left anywhere and every lane is dead
            }
        }
        __syncwarp();

        uint32_t fate = 0;
        int ev = EV_NONE;
        float end = 0.f;
        Recip rc{1.f, 1.f};
        if (alive) {
            if (!pending) {
                if (h_flight >= P.max_flights) {
                    fate = NRAPS_FATE_TRUNCATED;
                } else {
                    // ------------ FLIGHT: signed x-displacement to the next collision
                    ds = fmul(fmul(mu, -mc_logf(pcg32_unit(rng, inc))), s_inv_sigtr[mat + M * xsg]);
                    ++h_flight;
                }
            }
            if (!fate) {
                // ------------ WALK: cell by cell inside one material run.  Single-exit loop with
                // running shared addresses: ~30 SASS instructions per crossing (profiles/r1c_*).
                rc = make_recip(mu);
                int fwd = mu >= 0.0f ? 1 : 0;
                int dir = 2 * fwd - 1;
                int wall = fwd ? N - 1 : 0;
                int run_exit = fwd ? run_hi : run_lo - 1;
                uint32_t e_addr = edges_base + (uint32_t)(kStep * (cell + fwd)); // edge ahead of the neutron
                uint32_t t_addr = tally_ref<BIG>(lo_base, (row0 + g) * N + cell);         // tally[g][cell]
                // The hot loop below has two ways out and no wall logic.  The domain-boundary cell in the direction
                // of travel is handled here, before the loop (src/mc_code.rs:159-170): a lane that reaches it
                // inside the loop stops there (it is its stop_cell) and comes back through this block next trip.
                bool in_loop = true;
                pending = false;
                if (cell == wall) {
                    end = fadd(x, ds);
                    const float edge = BIG ? __ldg(P.edges + e_addr) : lds_f32(e_addr);
                    const float t = fsub(x, edge);
                    const bool beyond = fwd ? (end > edge) : (edge > end);
                    if (beyond) {
                        score<BIG>(t_addr, hi_off, fabsf(fdiv(t, mu)), P.tally);
                        const float b = fwd ? P.boundr : P.boundl;
                        if (!(b > 0.0f)) {
                            fate = NRAPS_FATE_LEAKED;
                            in_loop = false;
                        } else { // hit_boundary; the flight goes on in the other direction
                            mu = fmul(mu, -b);
                            ds = fmul(fadd(ds, t), -b);
                            x = edge;
                            rc = make_recip(mu);
                            fwd = mu >= 0.0f ? 1 : 0;
                            dir = 2 * fwd - 1;
                            wall = fwd ? N - 1 : 0;
                            run_exit = fwd ? run_hi : run_lo - 1;
                            e_addr = edges_base + (uint32_t)(kStep * (cell + fwd));
                            if (TRACE) ++h_refl;
                            // N == 1, or albedo <= 0 after the flip: the other wall is this very cell -> next trip
                            if (cell == wall) { in_loop = false; pending = true; }
                        }
                    } else if (!(fabsf(fsub(end, x)) > fabsf(t))) {
                        ev = EV_COLLIDE; // collision inside the boundary cell
                        in_loop = false;
                    } else {
                        // unreachable for finite positive inv_sigtr (a crossing test that succeeds where the wall
                        // test failed); keep the reference's order of tests and treat it as a collision at `end`
                        ev = EV_COLLIDE;
                        in_loop = false;
                    }
                }
                if (in_loop) {
                    // a walk longer than `walk_cap` crossings is suspended (pending) and resumed on the next trip, so
                    // the lanes that finished early are not kept waiting for the longest flight of the warp; the cap
                    // and the boundary cell are folded into the loop's one exit compare (`stop_cell`)
                    int steps = min((run_exit - cell) * dir, (int)P.walk_cap);
                    const int to_wall = (wall - cell) * dir; // cells until the boundary cell, if it lies in this run
                    if (to_wall < steps) steps = to_wall;
                    const int stride = kStep * dir;
                    const uint32_t e_first = e_addr;
                    // the loop tests the tally address, which nothing reads afterwards: testing e_addr lets the compiler
                    // substitute the stop value on that exit and brings the per-crossing moves back
                    const uint32_t t_stop = t_addr + (uint32_t)(stride * steps);
                    // The loop carries the two running addresses, ds and the position `xc` only.  x and cell are NOT
                    // kept up to date inside it: an unrolled loop with an exit per crossing otherwise pays four
                    // register moves per crossing to hold them in place for those exits (r1o SASS); both are rebuilt
                    // from the edge address afterwards.
                    float xc = x;
                    // one crossing; false = the walk is over (collision, or the stop edge reached)
                    auto step = [&]() -> bool {
                        end = fadd(xc, ds);
                        const float edge = BIG ? __ldg(P.edges + e_addr) : lds_f32(e_addr);
                        const float t = fsub(xc, edge);
                        if (!(fabsf(fsub(end, xc)) > fabsf(t))) return false; // collision at `end` (|edge - x| == |x - edge| exactly)
                        // cross_mesh, src/mc_code.rs:171-181
                        score<BIG>(t_addr, hi_off, fabsf(fast_div(t, rc)), P.tally);
                        ds = fadd(ds, t);
                        xc = edge;
                        e_addr += stride;
                        t_addr += stride;
                        return t_addr != t_stop; // left the material run, reached the boundary cell, or time to regroup
                    };
                    while (step() && step() && step() && step()) {} // unrolled by four: one back branch per four crossings
                    const int moved = (int)(e_addr - e_first) / kStep; // signed cells travelled
                    if (moved) {
                        // x after a crossing is the edge just crossed, bit for bit (src/mc_code.rs:72,77)
                        const uint32_t behind = e_addr - (uint32_t)stride;
                        x = BIG ? __ldg(P.edges + behind) : lds_f32(behind);
                        cell += moved;
                        if (TRACE) h_cross += (uint32_t)(moved * dir);
                    }
                    if (moved != dir * steps) ev = EV_COLLIDE;
                    else if (cell == run_exit) ev = EV_MATCHANGE;
                    else pending = true;
                }
            }
        }
        __syncwarp();

        // ---------------- COLLIDE / material change
        if (ev == EV_COLLIDE) {
            score<BIG>(tally_ref<BIG>(lo_base, (row0 + g) * N + cell), hi_off, fabsf(fast_div(fsub(x, end), rc)), P.tally);
            ++h_coll;
            const int xs = mat + M * xsg; // stale group index, src/mc_code.rs:147
            const float xi_int = pcg32_unit(rng, inc);
            const float mu_new = fsub(fmul(2.0f, pcg32_unit(rng, inc)), 1.0f);
            const int g_new = sample_group<TG>(s_scat + ((mat * G + g) * G + xsg) * G, G, P.scatter_mode, rng, inc);
            if (BANK) {
                const float nusigf = s_nusigf[mat + M * g];
                if (nusigf > 0.0f) {
                    const float wgt = fmul(fmul(nusigf, s_inv_sigtr[xs]), inv_k);
                    const uint32_t n = (uint32_t)__float2int_rz(fadd(wgt, pcg32_unit(rng, inc)));
                    const unsigned long long site = ((unsigned long long)(uint32_t)cell << 32) | __float_as_uint(end);
                    for (uint32_t j = 0; j < n; ++j) {
                        if (h_bank < P.bank_cap) P.slots[(y - P.hist_begin) * P.bank_cap + h_bank] = site;
                        ++h_bank;
                    }
                }
            }
            if (xi_int < s_p_abs[xs]) {
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
                mat = mesh.material(cell);
                xsg = g;
                const uint32_t rb = mesh.run_bounds(cell);
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
            if (BANK) {
                const uint32_t kept = h_bank < P.bank_cap ? h_bank : P.bank_cap;
                P.counts[y - P.hist_begin] = (uint8_t)kept;
                c_bank += kept;
            }
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

    const uint32_t vals[8] = {c_hist, c_coll, c_cross, c_flight, c_refl, c_leak, c_trunc, c_bank};
    flush_block(S, P, vals);
}

template <int TG, bool BIG>
cudaError_t launch_gb(const TransportParams &p, bool trace, bool bank, dim3 grid, dim3 block, uint32_t smem, cudaStream_t s)
{
    if (bank) {
        if (trace) transport_kernel<TG, true, true, BIG><<<grid, block, smem, s>>>(p);
        else transport_kernel<TG, false, true, BIG><<<grid, block, smem, s>>>(p);
    } else {
        if (trace) transport_kernel<TG, true, false, BIG><<<grid, block, smem, s>>>(p);
        else transport_kernel<TG, false, false, BIG><<<grid, block, smem, s>>>(p);
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
    if (bank) return trace ? cudaFuncSetAttribute(transport_kernel<TG, true, true, false>, attr, (int)bytes)
                           : cudaFuncSetAttribute(transport_kernel<TG, false, true, false>, attr, (int)bytes);
    return trace ? cudaFuncSetAttribute(transport_kernel<TG, true, false, false>, attr, (int)bytes)
                 : cudaFuncSetAttribute(transport_kernel<TG, false, false, false>, attr, (int)bytes);
}

} // namespace

// opt in to > 48 KB dynamic shared memory for the one instantiation about to be launched (BIG mode needs < 48 KB)
cudaError_t prepare_transport(uint32_t smem_bytes, uint32_t G, bool trace, bool bank)
{
    switch (G) {
    case 2: return set_smem<2>(smem_bytes, trace, bank);
    case 4: return set_smem<4>(smem_bytes, trace, bank);
    default: return set_smem<0>(smem_bytes, trace, bank);
    }
}

cudaError_t launch_transport(const TransportParams &p, bool trace, bool bank, dim3 grid, dim3 block, uint32_t smem, cudaStream_t s)
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
    if (bank) e = trace ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, transport_kernel<TG, true, true, BIG>, block, smem)
                        : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, transport_kernel<TG, false, true, BIG>, block, smem);
    else e = trace ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, transport_kernel<TG, true, false, BIG>, block, smem)
                   : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, transport_kernel<TG, false, false, BIG>, block, smem);
    return e == cudaSuccess ? n : 0;
}
template <int TG> int occ_g(bool big, bool trace, bool bank, int block, uint32_t smem)
{
    return big ? occ_gb<TG, true>(trace, bank, block, smem) : occ_gb<TG, false>(trace, bank, block, smem);
}
} // namespace

// resident blocks per SM of the instantiation that would be launched (registers and shared memory both count)
int occupancy_transport(uint32_t G, bool big, bool trace, bool bank, int block, uint32_t smem)
{
    switch (G) {
    case 2: return occ_g<2>(big, trace, bank, block, smem);
    case 4: return occ_g<4>(big, trace, bank, block, smem);
    default: return occ_g<0>(big, trace, bank, block, smem);
    }
}

} // namespace nraps
