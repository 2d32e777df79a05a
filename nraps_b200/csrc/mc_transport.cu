// Fused history kernel: one persistent lane per in-flight neutron, refilled
// from a per-warp chunk of history indices the moment its neutron dies.
//
// Replaces, per history (reference src/mc_code.rs):
//   spawn_neutron + energy ........ :7-53, 228-230   (source_kernel, mc_source.cu; the lane adopts the record)
//   particle_travel flight draw ... :147-148, 209    (stage FLIGHT)
//   boundary / cross_mesh loop .... :151-181, 56-79  (stage WALK)
//   scat_mat_calc + interaction ... :82-132, 183-208 (stage COLLIDE)
//   particle_lifetime tally ....... :224, 163/173/184 (fixed-point bins, see below)
//
// Warp structure: every trip of the outer loop each live lane draws one flight,
// walks cell crossings until its flight ends (collision, segment end, leak),
// then the warp reconverges (__syncwarp) and the lanes that collided run the
// collision stage together.  profiles/r1a_* is the ncu evidence that drove this
// shape: without the explicit reconvergence points the compiler let lanes fall
// out of the walk loop one by one (7.8 of 32 lanes active).
//
// Tallies (round 2: range updates).  The reference scores |dx / mu| into every
// cell a flight crosses (:163,173,184).  Round 1 did exactly that, one division
// + one shared atomic per crossing: 23 SASS instructions per crossing, 44 % of
// the kernel (profiles/r1r_*).  But a flight that crosses a cell completely
// scores |(e[i] - e[i+1]) / mu|, and inside a *segment* -- consecutive cells of
// one material whose widths e[i+1] - e[i] are the same binary32 number, i.e.
// practically a whole material run -- that is one and the same value v for every
// cell.  So the walk loop only replays what decides the trajectory (the two
// additions and the compare of the reference's collision test and the running
// ds += t, all in the reference's operation order, 10 instructions per
// crossing), and the scores of the n cells crossed completely are booked as ONE
// range update of a difference array: diff[first] += fx(v), diff[last+1] -=
// fx(v).  fx() is the same float -> 2^-28 fixed-point conversion as before and
// the bins are exact 64-bit integers (two u32 words, ATOMS.ADD is native for u32
// only), so tally = direct + prefix_sum(diff) (tally_prefix_kernel) equals the
// cell-by-cell sums of the oracle bit for bit, for any scheduling.  Only the
// partial cells at the two ends of a flight are still scored one by one.
// While |ds| is so large that the reference's test cannot fail (|ds| >= w + 2^-21 (L + |ds|)) a crossing changes
// nothing but ds <- fl(ds -+ w) and the edge address: those cells go through a loop of five instructions, the
// position is read once afterwards, and the ten-instruction step is left with the cell the flight ends in.
// On fine meshes most of the surely-crossed cells of a segment are not even walked: skip_cells() takes them in
// closed-form strides (see there).
#include "mc_lane.cuh"

namespace nraps {

namespace {

enum { EV_NONE = 0, EV_COLLIDE = 1, EV_SEGEXIT = 2 };

// 64-bit bin (low word at shared address `ref`, high word hi_off bytes above) += / -= fx.  The high word is touched
// only when the low word wraps or fx itself needs it (a score >= 16 cm): both rare, one predicate pair guards the RED.
// `wide` = the score itself needs the high word (fx >= 2^32, i.e. >= 16 cm), known from one compare on the scaled float.
static __device__ __forceinline__ void bin_add(uint32_t ref, uint32_t hi_off, unsigned long long fx, bool wide)
{
    const uint32_t l = (uint32_t)fx, h = (uint32_t)(fx >> 32);
    const uint32_t old = atoms_add(ref, l);
    const bool carry = old > ~l;
    if (carry | wide) reds_add(ref + hi_off, h + (carry ? 1u : 0u));
}
static __device__ __forceinline__ void bin_sub(uint32_t ref, uint32_t hi_off, unsigned long long fx, bool wide)
{
    const uint32_t l = (uint32_t)fx, h = (uint32_t)(fx >> 32);
    const uint32_t old = atoms_add(ref, 0u - l);
    const bool borrow = old < l;
    if (borrow | wide) reds_add(ref + hi_off, 0u - (h + (borrow ? 1u : 0u)));
}

// Cells of a segment that can be taken in one stride.  The walk changes ds by fl(ds -+ w) per cell; with a = |ds| in
// the binade [2^e, 2^(e+1)), grid U = 2^(e-23), and w not a rounding tie on that grid, every such step subtracts the
// same multiple q U from a (w rounded to the grid) as long as the exact difference stays inside the binade, so after j
// steps the mantissa is m - j q -- bit for bit what j sequential subtractions give.  j is limited to the cells whose
// |ds| stays >= lim (their collision test cannot fail), to the binade and to nmax; it is deliberately rounded down
// (a shorter stride is still exact).  Checked against the cell-by-cell loop on 3.6e7 random walks over six meshes
// before it went into the kernel (tools/closed_form_walk.c).
static __device__ __forceinline__ uint32_t skip_cells(float a, uint32_t mw, int ewb, float lim, uint32_t nmax, float &out)
{
    const uint32_t ia = __float_as_uint(a);
    const int eb = (int)(ia >> 23);
    const uint32_t m = (ia & 0x7fffffu) | 0x800000u;
    const int sh = eb - ewb;
    out = a;
    if (sh < 0 || sh > 24 || m <= 0x800000u) return 0u;
    uint32_t q = mw >> sh;
    const uint32_t rem = mw & ((1u << sh) - 1u), half = (1u << sh) >> 1;
    if (sh && rem == half) {
        // tie: round to even.  Once the mantissa is even it stays even and every step subtracts the even one of
        // {q, q + 1}; an odd mantissa is left to the caller's one exact subtraction, which makes it even
        if (m & 1u) return 0u;
        q += q & 1u;
    } else if (rem > half) q += 1u;
    const float ls = fmul(lim, __uint_as_float((uint32_t)(277 - eb) << 23)); // lim / U
    if (!(ls < 16777216.0f)) return 0u;
    const uint32_t mt = (uint32_t)__float2uint_ru(ls);
    if (m < mt) return 0u;
    float rq; // ~1/q, biased low by far more than the error of the approximate reciprocal
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rq) : "f"(__uint2float_rn(q)));
    rq = fmul(rq, 0.99999f);
    const uint32_t j1 = __float2uint_rz(fmul(__uint2float_rn(m - mt), rq)) + 1u;
    const uint32_t j2 = __float2uint_rz(fmul(__uint2float_rn(m - 0x800001u), rq));
    uint32_t j = j1 < j2 ? j1 : j2;
    if (j > nmax) j = nmax;
    const uint32_t mr = m - j * q;
    out = __uint_as_float(((uint32_t)eb << 23) | (mr & 0x7fffffu));
    return j;
}

// Where the tallies of this block live (see SurfLayout, mc_internal.h).
template <int MODE> struct Tally {
    uint32_t diff_base, direct_base, diff_hi_off, direct_hi_off; // shared byte addresses / offsets
    uint32_t N;
    unsigned long long *g_direct, *g_diff; // SURF_GLOBAL
    // score of one partially crossed cell (the cell a flight starts or ends in, a wall cell)
    __device__ __forceinline__ void direct(int row, int cell, float v) const
    {
        const float vs = fmul(v, kTallyScale);
        const unsigned long long fx = __float2ull_rz(vs);
        const bool wide = vs >= 4294967296.0f;
        if (MODE == SURF_SPLIT) {
            bin_add(direct_base + 4u * (uint32_t)(row * (int)N + cell), direct_hi_off, fx, wide);
        } else if (MODE == SURF_UNIFIED) {
            const uint32_t ref = diff_base + 4u * (uint32_t)(row * (int)(N + 1u) + cell);
            bin_add(ref, diff_hi_off, fx, wide);
            bin_sub(ref + 4u, diff_hi_off, fx, wide);
        } else {
            atomicAdd(g_direct + (size_t)row * N + cell, fx);
        }
    }
    // the same score v for every cell of [lo, lo + n): one range update of the difference array
    __device__ __forceinline__ void range(int row, int lo, int n, float v) const
    {
        const float vs = fmul(v, kTallyScale);
        const unsigned long long fx = __float2ull_rz(vs);
        const bool wide = vs >= 4294967296.0f;
        if (MODE != SURF_GLOBAL) {
            const uint32_t ref = diff_base + 4u * (uint32_t)(row * (int)(N + 1u) + lo);
            bin_add(ref, diff_hi_off, fx, wide);
            bin_sub(ref + 4u * (uint32_t)n, diff_hi_off, fx, wide);
        } else {
            unsigned long long *d = g_diff + (size_t)row * N + lo;
            atomicAdd(d, fx);
            if ((uint32_t)(lo + n) < N) atomicAdd(d + n, 0ull - fx); // the entry one past the last cell feeds no prefix
        }
    }
};

// Register budget per instantiation.  Blocks of up to 1024 threads must be launchable (64 registers); the instantiations
// that share an SM between two blocks (SURF_SPLIT) are asked for 56, which lets 2 x 576 threads reside -- left to itself
// ptxas gives the fission-bank one anything from 55 to 64 depending on unrelated edits.
template <bool TRACE, int MODE> struct RegCap { static constexpr int k = (MODE == SURF_SPLIT && !TRACE) ? 56 : 64; };

template <int TG, bool TRACE, bool BANK, int MODE>
__global__ void __maxnreg__((RegCap<TRACE, MODE>::k)) transport_kernel(const TransportParams P)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr bool BIG = (MODE == SURF_GLOBAL);
    constexpr bool kSkip = (MODE != SURF_SPLIT); // meshes fine enough to leave SURF_SPLIT have segments worth striding over
    const int G = TG ? TG : (int)P.G;
    const int M = (int)P.M, N = (int)P.N;
    const SurfLayout L = make_surface_layout(P.M, P.G, P.N, MODE, P.rows);
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int MG = M * G;

    // ---------------- block prologue: stage the tables, clear the bins
    float *s_xs = reinterpret_cast<float *>(smem_raw + L.xs);
    for (int i = tid; i < (int)xs_floats(P.M, P.G); i += nthr) s_xs[i] = P.xs[i];
    if (!BIG) {
        uint32_t *bins = reinterpret_cast<uint32_t *>(smem_raw + L.diff_lo);
        const int n_words = (int)(L.edges - L.diff_lo) / 4; // diff lo/hi (+ direct lo/hi) are contiguous
        for (int i = tid; i < n_words; i += nthr) bins[i] = 0u;
        float *e = reinterpret_cast<float *>(smem_raw + L.edges);
        for (int i = tid; i <= N; i += nthr) e[i] = P.edges[i];
        uint2 *sw = reinterpret_cast<uint2 *>(smem_raw + L.segw);
        for (int i = tid; i < N; i += nthr) { sw[i] = P.segw[i]; smem_raw[L.matid + i] = P.matid[i]; }
    }
    __syncthreads();

    const uint32_t smem_base = (uint32_t)__cvta_generic_to_shared(smem_raw);
    Tally<MODE> T;
    T.diff_base = smem_base + L.diff_lo; T.diff_hi_off = P.diff_hi_off;
    T.direct_base = smem_base + L.direct_lo; T.direct_hi_off = P.direct_hi_off;
    T.N = P.N; T.g_direct = P.tally; T.g_diff = P.diff;
    uint32_t edges_base = BIG ? 0u : smem_base + L.edges; // edge reference: shared byte address, or the edge index
    uint32_t segw_base = smem_base + L.segw, matid_base = smem_base + L.matid;
#ifndef NRAPS_NO_PIN
    asm volatile("" : "+r"(edges_base), "+r"(segw_base), "+r"(matid_base));
#endif
    constexpr int kStep = BIG ? 1 : 4;
    // keep the table addresses in registers: left to itself the compiler rebuilds them from SR_CgaCtaId on every trip
#ifndef NRAPS_NO_PIN
    asm volatile("" : "+r"(T.diff_base), "+r"(T.direct_base));
#endif
    auto ld_edge = [&](uint32_t ref) -> float { return BIG ? __ldg(P.edges + ref) : lds_f32(ref); };
    auto ld_mat = [&](int c) -> int { return BIG ? (int)__ldg(P.matid + c) : (int)lds_u8(matid_base + (uint32_t)c); };

    const float *s_inv_sigtr = s_xs, *s_p_abs = s_xs + MG, *s_nusigf = s_xs + 3 * MG, *s_scat = s_xs + 5 * MG;
    // fission_bank mode: sites are banked with weight nu*Sigma_f * inv_sigtr / k_prev; an empty bank => uniform source
    const float inv_k = BANK ? fdiv(1.0f, *P.k_cur) : 1.0f;

    const unsigned lane = tid & 31;
    const uint64_t inc = P.rng_inc;

    // warp-uniform cursor over the chunk of histories this warp owns, relative to hist_begin (a launch never covers
    // 2^32 histories: the host hands larger shards over in sub-shards)
    uint32_t w_next = 0, w_end = 0;
    const uint32_t span = (uint32_t)(P.hist_end - P.hist_begin);
    bool exhausted = false;

    // lane state: one neutron
    bool alive = false, pending = false;
    uint64_t rng = 0;
    uint32_t y = 0; // history index relative to hist_begin
    float x = 0.f, mu = 1.f, ds = 0.f;
    int cell = 0, g = 0, xsg = 0, mat = 0, row0 = 0;
    uint32_t h_coll = 0, h_cross = 0, h_flight = 0, h_refl = 0, h_bank = 0; // this history
    uint32_t c_hist = 0, c_coll = 0, c_cross = 0, c_flight = 0, c_refl = 0, c_leak = 0, c_trunc = 0, c_bank = 0;

    // set by the walk of a trip, read by its collision stage only: declared out here so that they are not re-initialised
    // every trip (4 instructions of ~410; measured +0.6 %)
    float end = 0.f;
    Recip rc{1.f, 1.f};
    for (;;) {
        // ---------------- SPAWN: hand fresh history indices to dead lanes (the ballot is the warp's convergence point)
        const unsigned need = __ballot_sync(kFull, !alive);
        if (need && ((uint32_t)__popc(need) >= P.spawn_batch || need == kFull)) {
            if (w_next == w_end && !exhausted) {
                unsigned long long base = 0;
                if (lane == 0) base = atomicAdd(P.work, (unsigned long long)P.chunk);
                base = __shfl_sync(kFull, base, 0);
                if (base >= span) exhausted = true;
                else {
                    w_next = (uint32_t)base;
                    w_end = (base + P.chunk < span) ? (uint32_t)base + P.chunk : span;
                }
            }
            const uint32_t avail = w_end - w_next;
            if (avail) {
                const uint32_t rank = __popc(need & ((1u << lane) - 1u));
                if (!alive && rank < avail) {
                    y = w_next + rank;
                    // adopt the neutron source_kernel gave birth to (mc_source.cu)
                    const uint4 *rec = P.source + 2 * (size_t)y;
                    const uint4 r0 = __ldg(rec);
                    const uint2 r1 = __ldg(reinterpret_cast<const uint2 *>(rec + 1));
                    x = __uint_as_float(r0.x);
                    mu = __uint_as_float(r0.y);
                    cell = (int)(r0.z & 0xffffu);
                    g = (int)(r0.z >> 16);
                    row0 = (int)r0.w; // first tally row of this history's generation (0 unless generations are batched)
                    rng = (uint64_t)r1.x | ((uint64_t)r1.y << 32);
                    mat = ld_mat(cell);
                    xsg = g;
                    h_bank = 0;
                    h_coll = h_cross = h_flight = h_refl = 0;
                    alive = true;
                    pending = false;
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
                // ------------ WALK: cell by cell inside one segment
                rc = make_recip(mu);
                // {stop edge going left | stop edge going right << 16, width bits} of the segment `cell` lies in
                const uint2 sw = BIG ? __ldg(P.segw + cell) : make_uint2(lds_u32(segw_base + 8u * (uint32_t)cell), lds_u32(segw_base + 8u * (uint32_t)cell + 4u));
                int fwd = mu >= 0.0f ? 1 : 0;
                int dir = 2 * fwd - 1;
                int wall = fwd ? N - 1 : 0;
                uint32_t e_addr = edges_base + (uint32_t)(kStep * (cell + fwd)); // edge ahead of the neutron
                // The loop below has no wall logic.  The domain-boundary cell in the direction of travel is handled
                // here, before it (src/mc_code.rs:159-170): a lane that reaches that cell inside the loop stops there
                // and comes back through this block on its next trip.
                bool in_loop = true;
                pending = false;
                if (cell == wall) {
                    end = fadd(x, ds);
                    const float edge = ld_edge(e_addr);
                    const float t = fsub(x, edge);
                    const bool beyond = fwd ? (end > edge) : (edge > end);
                    if (beyond) {
                        T.direct(row0 + g, cell, fabsf(fdiv(t, mu)));
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
                            e_addr = edges_base + (uint32_t)(kStep * (cell + fwd));
                            if (TRACE) ++h_refl;
                            // N == 1, or albedo <= 0 after the flip: the other wall is this very cell -> next trip
                            if (cell == wall) { in_loop = false; pending = true; }
                        }
                    } else {
                        // collision inside the boundary cell.  (A crossing test that succeeds where the wall test
                        // failed is unreachable for finite positive inv_sigtr; the reference's order of tests is kept
                        // and that case too is a collision at `end`.)
                        ev = EV_COLLIDE;
                        in_loop = false;
                    }
                }
                if (in_loop) {
                    // the cell the flight starts in: an arbitrary part of it is crossed (src/mc_code.rs:171-181)
                    end = fadd(x, ds);
                    const float edge0 = ld_edge(e_addr);
                    const float t0 = fsub(x, edge0);
                    if (!(fabsf(fsub(end, x)) > fabsf(t0))) {
                        ev = EV_COLLIDE; // collision at `end`, before the first edge
                    } else {
                        T.direct(row0 + g, cell, fabsf(fast_div(t0, rc)));
                        ds = fadd(ds, t0);
                        const int stride = kStep * dir;
                        // where this walk stops if no collision does: the edge reference once the neutron has entered
                        // the first cell that is not its segment's -- or the boundary cell, which the block above handles
                        // on the next trip (both folded into one per-cell table entry by the host)
                        const uint32_t e_stop = edges_base + (uint32_t)kStep * (fwd ? (sw.x >> 16) : (sw.x & 0xffffu));
                        const int cell0 = cell;
                        float xc = edge0;
                        e_addr += (uint32_t)stride;
                        if (e_addr != e_stop) {
                            // Cells crossed completely: x - edge is -+w for every one of them, w the segment's width,
                            // so the loop carries ds, the position and the edge address only.  One step = the
                            // reference's test `|end - x| > |edge - x|` and `ds += x - edge` in its operation order.
                            const float w = __uint_as_float(sw.y);
                            const float tn = fwd ? -w : w;
                            auto step = [&]() -> bool {
                                end = fadd(xc, ds);
                                if (!(fabsf(fsub(end, xc)) > w)) return false; // collision at `end` inside this cell
                                ds = fadd(ds, tn);
                                xc = ld_edge(e_addr);
                                e_addr += (uint32_t)stride;
                                return e_addr != e_stop;
                            };
                            // |ds| >= lim: the reference's test cannot fail in the cell ahead (the roundings of x + ds and
                            // of the difference are bounded by 2^-22 (L + |ds|)).  |ds| only shrinks along a walk, so one
                            // bound computed from the |ds| at its start serves the whole walk.
                            const float lim = __fmaf_rn(fadd(P.length, fabsf(ds)), 4.76837158203125e-07f, w); // w + 2^-21 (L + |ds|)
                            if (kSkip && P.skip_walk) {
                                // Fine meshes: a flight crosses tens of cells of one segment.  While |ds| is surely large
                                // enough for the collision test to pass, the only thing a crossing changes is
                                // ds <- fl(ds -+ w), and inside one binade of |ds| that is an exact integer recurrence on
                                // the mantissa (skip_cells above): those cells are taken in one stride, in registers.  One
                                // real subtraction carries |ds| over what the recurrence does not cover (the change of
                                // binade, an odd mantissa under a rounding tie), then the next stride.
                                const uint32_t mw = (sw.y & 0x7fffffu) | 0x800000u;
                                const int ewb = (int)(sw.y >> 23);
                                float a = fabsf(ds);
                                // strides stop at the first power of two >= stride_min * w: the last stride then ends with
                                // its binade instead of a few cells into the next one (a stride costs a dozen crossings)
                                const float w4 = __uint_as_float((__float_as_uint(fmul(w, P.stride_min)) + 0x7fffffu) & 0xff800000u);
                                const uint32_t total = (uint32_t)((int)(e_stop - e_addr) * dir) / (uint32_t)kStep; // cells to the stop
                                uint32_t done = 0u;
#pragma unroll 1
                                for (int round = 0; round < 8; ++round) {
                                    const uint32_t rem = total - done;
                                    if (rem < 4u || !(a > w4)) break; // a handful of cells left: the loops below are cheaper
                                    float an;
                                    done += skip_cells(a, mw, ewb, lim, rem, an);
                                    a = an;
                                    if (done == total || !(a >= lim)) break;
                                    a = fsub(a, w); // |fl(ds -+ w)|: this cell too is surely crossed
                                    done += 1u;
                                }
                                if (done) {
                                    ds = copysignf(a, ds);
                                    e_addr += (uint32_t)(stride * (int)done);
                                }
                            }
#ifndef NRAPS_NO_SURE_LOOP
                            // Surely crossed cells, one by one: nothing but ds <- fl(ds -+ w) and the edge address, five
                            // instructions per crossing against the ten of the exact step; the position is read once,
                            // afterwards.  What is left for the exact loop is the cell the flight ends in.
                            if (e_addr != e_stop && fabsf(ds) >= lim) {
#pragma unroll 1
                                do {
                                    ds = fadd(ds, tn);
                                    e_addr += (uint32_t)stride;
                                } while ((e_addr != e_stop) & (fabsf(ds) >= lim));
                                asm volatile("" : "+r"(e_addr)); // or the compiler carries e_addr - stride through the loop
                            }
#endif
                            xc = ld_edge(e_addr - (uint32_t)stride); // the edge crossed last
                            if (e_addr != e_stop) {
                                // the reference's own step, for the cell the flight ends in (one iteration in 99 % of the
                                // trips).  Not unrolled: measured the same unrolled by four when it still did all the cells.
#pragma unroll 1
                                while (step()) {}
                            }
                        }
                        // the cell the neutron is in now: the edge ahead of it is e_addr
                        cell = (int)((e_addr - edges_base) / (uint32_t)kStep) - fwd;
                        const int n_full = (cell - cell0) * dir - 1; // cells crossed completely
                        if (n_full > 0) // [cell0+1, cell-1] going right, [cell+1, cell0-1] going left
                            T.range(row0 + g, fwd ? cell0 + 1 : cell + 1, n_full, fabsf(fast_div(fwd ? -__uint_as_float(sw.y) : __uint_as_float(sw.y), rc)));
                        ev = (e_addr != e_stop) ? EV_COLLIDE : EV_SEGEXIT; // the loop leaves early only on a collision
                        x = xc; // x after a crossing is the edge just crossed (src/mc_code.rs:72,77)
                        if (TRACE) h_cross += (uint32_t)(n_full + 1);
                    }
                }
            }
            if (ev == EV_SEGEXIT) { // ------------ end of the segment
                // (the stop-edge table keeps the walk inside [0, N-1]: the boundary cells are never left through the loop)
                ev = EV_NONE;
                const int m2 = ld_mat(cell);
                if (m2 != mat) { // material change: the flight ends here, alive (src/mc_code.rs:175-181)
                    mat = m2;
                    xsg = g;
                } else {
                    // same material: the next segment of the run (the cell width changed by an ulp), or the
                    // boundary cell, which the walk never enters on its own: the flight goes on next trip
                    pending = true;
                }
            }
        }
        __syncwarp();

        // ---------------- COLLIDE
        if (ev == EV_COLLIDE) {
            T.direct(row0 + g, cell, fabsf(fast_div(fsub(x, end), rc)));
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
                    if (n) { // a history's sites fill its slot row in order; what does not fit is counted, not kept
                        const uint32_t have = h_bank < P.bank_cap ? h_bank : P.bank_cap;
                        uint32_t fit = P.bank_cap - have;
                        fit = n < fit ? n : fit;
                        unsigned long long *dst = P.slots + ((size_t)y * P.bank_cap + have);
#pragma unroll 1
                        for (; fit; --fit) *dst++ = site;
                        h_bank += n;
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
                P.counts[y] = (uint8_t)kept;
                c_bank += kept;
            }
            if (TRACE) {
                c_cross += h_cross; c_refl += h_refl;
                if (P.trace) {
                    uint32_t *t = P.trace + (size_t)y * NRAPS_TR_WORDS;
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

    // ---------------- block epilogue: shared bins -> global 64-bit bins, lane counters -> global counters
    __syncthreads();
    const int rows = (int)P.rows;
    if (!BIG) {
        const uint32_t *d_lo = reinterpret_cast<const uint32_t *>(smem_raw + L.diff_lo), *d_hi = reinterpret_cast<const uint32_t *>(smem_raw + L.diff_hi);
        const uint32_t *t_lo = reinterpret_cast<const uint32_t *>(smem_raw + L.direct_lo), *t_hi = reinterpret_cast<const uint32_t *>(smem_raw + L.direct_hi);
        for (int i = tid; i < rows * N; i += nthr) {
            const int r = i / N, c = i - r * N, j = r * (N + 1) + c; // the entry one past the last cell of a row feeds no prefix
            const unsigned long long d = ((unsigned long long)d_hi[j] << 32) + d_lo[j];
            if (d) atomicAdd(&P.diff[i], d);
            if (MODE == SURF_SPLIT) {
                const unsigned long long v = ((unsigned long long)t_hi[i] << 32) + t_lo[i];
                if (v) atomicAdd(&P.tally[i], v);
            }
        }
    }
    unsigned long long *ct = P.tally + (size_t)rows * N;
    const uint32_t vals[8] = {c_hist, c_coll, c_cross, c_flight, c_refl, c_leak, c_trunc, c_bank};
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        unsigned long long v = vals[c];
#pragma unroll
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
        if ((tid & 31) == 0 && v) atomicAdd(&ct[c], v);
    }
}

// tally[r][i] += sum_{j <= i} diff[r][j]: one block per tally row, warp-shuffle scan, 64-bit wrapping sums
__global__ void __launch_bounds__(1024) tally_prefix_kernel(const unsigned long long *diff, unsigned long long *tally, const uint32_t N)
{
    __shared__ unsigned long long warp_tot[32];
    const unsigned long long *d = diff + (size_t)blockIdx.x * N;
    unsigned long long *t = tally + (size_t)blockIdx.x * N;
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long carry = 0ull;
    for (uint32_t base = 0; base < N; base += 1024) {
        const uint32_t i = base + threadIdx.x;
        unsigned long long v = i < N ? d[i] : 0ull;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long up = __shfl_up_sync(kFull, v, o);
            if (lane >= (unsigned)o) v += up;
        }
        if (lane == 31) warp_tot[warp] = v;
        __syncthreads();
        if (warp == 0) {
            unsigned long long w = warp_tot[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned long long up = __shfl_up_sync(kFull, w, o);
                if (lane >= (unsigned)o) w += up;
            }
            warp_tot[lane] = w; // inclusive totals of warps 0..lane
        }
        __syncthreads();
        const unsigned long long incl = carry + v + (warp ? warp_tot[warp - 1] : 0ull);
        if (i < N && incl) t[i] += incl;
        carry += warp_tot[31];
        __syncthreads();
    }
}

template <int TG, int MODE>
cudaError_t launch_gm(const TransportParams &p, bool trace, bool bank, dim3 grid, dim3 block, uint32_t smem, cudaStream_t s)
{
    if (bank) {
        if (trace) transport_kernel<TG, true, true, MODE><<<grid, block, smem, s>>>(p);
        else transport_kernel<TG, false, true, MODE><<<grid, block, smem, s>>>(p);
    } else {
        if (trace) transport_kernel<TG, true, false, MODE><<<grid, block, smem, s>>>(p);
        else transport_kernel<TG, false, false, MODE><<<grid, block, smem, s>>>(p);
    }
    return cudaGetLastError();
}

template <int TG>
cudaError_t launch_g(const TransportParams &p, bool trace, bool bank, dim3 grid, dim3 block, uint32_t smem, cudaStream_t s)
{
    switch (p.surf_mode) {
    case SURF_SPLIT: return launch_gm<TG, SURF_SPLIT>(p, trace, bank, grid, block, smem, s);
    case SURF_UNIFIED: return launch_gm<TG, SURF_UNIFIED>(p, trace, bank, grid, block, smem, s);
    default: return launch_gm<TG, SURF_GLOBAL>(p, trace, bank, grid, block, smem, s);
    }
}

template <typename F> cudaError_t set_smem_one(F *kernel, uint32_t bytes)
{
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

template <int TG, int MODE> cudaError_t set_smem(uint32_t bytes, bool trace, bool bank)
{
    if (bank) return trace ? set_smem_one(transport_kernel<TG, true, true, MODE>, bytes) : set_smem_one(transport_kernel<TG, false, true, MODE>, bytes);
    return trace ? set_smem_one(transport_kernel<TG, true, false, MODE>, bytes) : set_smem_one(transport_kernel<TG, false, false, MODE>, bytes);
}

template <int TG> cudaError_t set_smem_g(uint32_t bytes, uint32_t mode, bool trace, bool bank)
{
    return mode == SURF_SPLIT ? set_smem<TG, SURF_SPLIT>(bytes, trace, bank) : set_smem<TG, SURF_UNIFIED>(bytes, trace, bank);
}

} // namespace

// opt in to > 48 KB dynamic shared memory for the one instantiation about to be launched (SURF_GLOBAL needs < 48 KB)
cudaError_t prepare_transport(uint32_t smem_bytes, uint32_t G, uint32_t surf_mode, bool trace, bool bank)
{
    if (surf_mode == SURF_GLOBAL) return cudaSuccess;
    switch (G) {
    case 2: return set_smem_g<2>(smem_bytes, surf_mode, trace, bank);
    case 4: return set_smem_g<4>(smem_bytes, surf_mode, trace, bank);
    default: return set_smem_g<0>(smem_bytes, surf_mode, trace, bank);
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

cudaError_t launch_tally_prefix(const unsigned long long *diff, unsigned long long *tally, uint32_t rows, uint32_t N, cudaStream_t s)
{
    if (!rows || !N) return cudaSuccess;
    tally_prefix_kernel<<<rows, 1024, 0, s>>>(diff, tally, N);
    return cudaGetLastError();
}

namespace {
template <typename F> int occ_one(F *kernel, int block, uint32_t smem)
{
    int n = 0;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, block, smem) == cudaSuccess ? n : 0;
}
template <int TG, int MODE> int occ_gm(bool trace, bool bank, int block, uint32_t smem)
{
    if (bank) return trace ? occ_one(transport_kernel<TG, true, true, MODE>, block, smem) : occ_one(transport_kernel<TG, false, true, MODE>, block, smem);
    return trace ? occ_one(transport_kernel<TG, true, false, MODE>, block, smem) : occ_one(transport_kernel<TG, false, false, MODE>, block, smem);
}
template <int TG> int occ_g(uint32_t mode, bool trace, bool bank, int block, uint32_t smem)
{
    switch (mode) {
    case SURF_SPLIT: return occ_gm<TG, SURF_SPLIT>(trace, bank, block, smem);
    case SURF_UNIFIED: return occ_gm<TG, SURF_UNIFIED>(trace, bank, block, smem);
    default: return occ_gm<TG, SURF_GLOBAL>(trace, bank, block, smem);
    }
}
} // namespace

// resident blocks per SM of the instantiation that would be launched (registers and shared memory both count)
int occupancy_transport(uint32_t G, uint32_t surf_mode, bool trace, bool bank, int block, uint32_t smem)
{
    switch (G) {
    case 2: return occ_g<2>(surf_mode, trace, bank, block, smem);
    case 4: return occ_g<4>(surf_mode, trace, bank, block, smem);
    default: return occ_g<0>(surf_mode, trace, bank, block, smem);
    }
}

} // namespace nraps
