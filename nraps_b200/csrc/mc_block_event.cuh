// Block-level event pipeline for surface tracking (kernel_variant = NRAPS_KERNEL_BLOCK_EVENT): EXPERIMENTAL.
//
// Same physics, same per-history RNG streams and the same integer tallies as transport_kernel (mc_transport.cu), so
// the results are bit-identical; what changes is who executes which event.  In the lane kernel a neutron lives in
// the registers of one lane for its whole life and a warp trip runs flight / walk / collide for whatever mix of
// neutrons its lanes hold (ncu, profiles/r1r: 18 of 32 lanes busy; the collide stage runs with 10 lanes, a walk lasts
// as long as the longest of the warp, 8 cells in a fuel pin against 4 in a water gap).  Here the neutrons of a block
// are 28-byte records in shared memory (structure of arrays, several per thread) and a record MOVES to the list of
// the event it needs next, so a warp always reads and writes 32 consecutive records (no bank conflicts):
//
//   arena A   `coll` grows up from record 0, `go` grows down from record S-1
//   arena B   `walk[0]` (short material runs) grows up, `walk[1]` (long runs) grows down
//
//   phase AB  reads arena A: coll -> collide (src/mc_code.rs:182-210), then the survivor's flight draw (:147-148);
//                            go   -> flight draw (the neutron entered another material run, :179-181), or nothing
//                                    for a walk that was only suspended (flag `pending`);
//             and adopts histories born by source_kernel (src/mc_code.rs:40-53) into the free capacity;
//             writes arena B by the length class of the material run the neutron sits in.
//   phase C   reads arena B: boundary + cross_mesh loop (:151-181, 56-79); writes arena A: coll / go; leaked and
//             truncated histories simply are not written back.
//
// A list that grows up and one that grows down share an arena of S records exactly, whatever their split.  That is the
// north star's event-based pipeline over structure-of-arrays banks, kept on chip: the HBM version of it (mc_event.cu)
// lost 8x to bank round trips and launch tails (profiles/r1e); 227 KB of shared memory per SM hold the bank instead.
//
// The per-thread body below is written against a small context type C (thread index, barrier, shared / global
// atomics, warp-aggregated list claims, shared-space loads of the mesh tables and the tally score), so that the very
// same code runs as a CUDA block (mc_block_event.cu) and, for tests only, on CPU threads (tests/emul) where it is
// bit-compared with the oracle without a GPU.  Restrictions of this variant: uniform source, no trace records, no
// generation batching, mesh image in shared memory (no BIG mode); everything else is refused with NRAPS_ERR_OPTION.
#pragma once
#include "mc_lane.cuh"

namespace nraps {
namespace bev {

enum { OUT_COLLIDE = 1, OUT_MATCHANGE = 2, OUT_PENDING = 3, OUT_LEAK = 4, OUT_TRUNC = 5 };

// counters in shared memory (u32 words).  The two lists of an arena share ONE word -- length of the list growing up in
// the low 16 bits, of the list growing down in the high 16 bits (S <= 65535) -- so that a warp claims its positions in
// both with a single shared atomic; the words exist once per round parity.
enum {
    K_A0 = 0, K_A1, // arena A, [parity]: coll | go << 16
    K_B0, K_B1,     // arena B, [parity]: walk[short] | walk[long] << 16
    K_SRC_NEXT, K_SRC_END, K_EXHAUSTED, K_DONE, K_SPLIT, K_ADOPT,
    K_WORDS = 16
};

// one arena of S neutron records, structure of arrays
struct Arena {
    float *x, *mu, *ds;            // position, direction cosine, signed distance left to the collision site (or the site itself)
    uint32_t *pk, *rlo, *rhi, *hf; // cell | g<<16 | xsg<<19 | mat<<22 | pending<<28, rng state, flights so far
};

struct Bank {
    Arena A, B;
    uint32_t *k; // [K_WORDS]
    uint32_t S;
};

__host__ __device__ inline uint32_t bank_bytes(uint32_t S) { return 2 * S * 7 * 4 + K_WORDS * 4; }

__host__ __device__ inline Arena make_arena(uint32_t *raw, uint32_t S)
{
    Arena a;
    a.x = reinterpret_cast<float *>(raw);
    a.mu = a.x + S;
    a.ds = a.mu + S;
    a.pk = raw + 3 * S;
    a.rlo = a.pk + S;
    a.rhi = a.rlo + S;
    a.hf = a.rhi + S;
    return a;
}

// carve a Bank out of `raw` (4-byte aligned)
__host__ __device__ inline Bank make_bank(unsigned char *raw, uint32_t S)
{
    Bank b;
    b.S = S;
    uint32_t *w = reinterpret_cast<uint32_t *>(raw);
    b.k = w;
    b.A = make_arena(w + K_WORDS, S);
    b.B = make_arena(w + K_WORDS + 7 * S, S);
    return b;
}

// Arithmetic of a warp-aggregated claim on the packed length word of an arena (low half: the list growing up, high half:
// the list growing down).  m_up / m_down are the ballots of the lanes pushing to either list; the warp's leader adds
// claim2_increment() to the word once and every lane derives its own position from the value the word had before.
// Kept apart from the CUDA intrinsics so that the CPU tests can check it lane by lane (tests/emul).
NRAPS_HD uint32_t popc32(uint32_t v)
{
#if defined(__CUDA_ARCH__) || !defined(NRAPS_EMUL)
    return (uint32_t)__popc(v);
#else
    return (uint32_t)__builtin_popcount(v);
#endif
}
NRAPS_HD uint32_t claim2_increment(uint32_t m_up, uint32_t m_down) { return popc32(m_up) | (popc32(m_down) << 16); }
NRAPS_HD uint32_t claim2_position(uint32_t m_up, uint32_t m_down, uint32_t lane, uint32_t old, bool down)
{
    const uint32_t below = (1u << lane) - 1u;
    return down ? (old >> 16) + popc32(m_down & below) : (old & 0xffffu) + popc32(m_up & below);
}

struct Neutron {
    float x, mu, ds;
    int cell, g, xsg, mat;
    bool pending;
    uint64_t rng;
    uint32_t hf;
};

NRAPS_HD Neutron load_neutron(const Arena &a, uint32_t s)
{
    Neutron n;
    n.x = a.x[s]; n.mu = a.mu[s]; n.ds = a.ds[s];
    const uint32_t pk = a.pk[s];
    n.cell = (int)(pk & 0xffffu); n.g = (int)((pk >> 16) & 7u); n.xsg = (int)((pk >> 19) & 7u); n.mat = (int)((pk >> 22) & 63u);
    n.pending = ((pk >> 28) & 1u) != 0;
    n.rng = (uint64_t)a.rlo[s] | ((uint64_t)a.rhi[s] << 32);
    n.hf = a.hf[s];
    return n;
}

NRAPS_HD void store_neutron(const Arena &a, uint32_t s, const Neutron &n)
{
    a.x[s] = n.x; a.mu[s] = n.mu; a.ds[s] = n.ds;
    a.pk[s] = (uint32_t)n.cell | ((uint32_t)n.g << 16) | ((uint32_t)n.xsg << 19) | ((uint32_t)n.mat << 22) | ((n.pending ? 1u : 0u) << 28);
    a.rlo[s] = (uint32_t)n.rng; a.rhi[s] = (uint32_t)(n.rng >> 32);
    a.hf[s] = n.hf;
}

struct Counts {
    uint32_t hist = 0, coll = 0, flight = 0, leak = 0, trunc = 0;
};

// src/mc_code.rs:147-148 (and :209): one flight draw.  false = the flight cap truncated the history.
template <class C> NRAPS_HD bool flight(C &c, const TransportParams &P, Neutron &n, Counts &ct)
{
    if (n.pending) { // a walk that was only suspended (walk cap, wall cell) goes on with the flight it has
        n.pending = false;
        return true;
    }
    if (n.hf >= P.max_flights) return false;
    n.ds = fmul(fmul(n.mu, -mc_logf(pcg32_unit(n.rng, P.rng_inc))), c.inv_sigtr(n.mat + (int)P.M * n.xsg));
    ++n.hf;
    ++ct.flight;
    n.pending = false;
    return true;
}

// boundary / cross_mesh loop inside one material run (src/mc_code.rs:151-181, 56-79): the walk of transport_kernel,
// statement for statement.  On OUT_COLLIDE n.ds holds the collision site (`end`), n.x the point the last segment began.
template <class C> NRAPS_HD int walk(C &c, const TransportParams &P, Neutron &n)
{
    const int N = (int)P.N;
    Recip rc = make_recip(n.mu);
    int fwd = n.mu >= 0.0f ? 1 : 0;
    int dir = 2 * fwd - 1;
    int wall = fwd ? N - 1 : 0;
    const uint32_t rb = c.run_bounds(n.cell);
    const int run_lo = (int)(rb & 0xffffu), run_hi = (int)(rb >> 16);
    int run_exit = fwd ? run_hi : run_lo - 1;
    uint32_t e_ref = c.edge_ref(n.cell + fwd);
    uint32_t t_ref = c.tally_ref(n.g * N + n.cell);
    float end = 0.0f;
    n.pending = false;
    if (n.cell == wall) { // the domain-boundary cell in the direction of travel, src/mc_code.rs:159-170
        end = fadd(n.x, n.ds);
        const float edge = c.edge(e_ref);
        const float t = fsub(n.x, edge);
        const bool beyond = fwd ? (end > edge) : (edge > end);
        if (!beyond) {
            n.ds = end;
            return OUT_COLLIDE;
        }
        c.score(t_ref, fabsf(fdiv(t, n.mu)));
        const float b = fwd ? P.boundr : P.boundl;
        if (!(b > 0.0f)) return OUT_LEAK;
        n.mu = fmul(n.mu, -b); // hit_boundary
        n.ds = fmul(fadd(n.ds, t), -b);
        n.x = edge;
        rc = make_recip(n.mu);
        fwd = n.mu >= 0.0f ? 1 : 0;
        dir = 2 * fwd - 1;
        wall = fwd ? N - 1 : 0;
        run_exit = fwd ? run_hi : run_lo - 1;
        e_ref = c.edge_ref(n.cell + fwd);
        if (n.cell == wall) { n.pending = true; return OUT_PENDING; }
    }
    int steps = (run_exit - n.cell) * dir;
    if ((int)P.walk_cap < steps) steps = (int)P.walk_cap;
    const int to_wall = (wall - n.cell) * dir;
    if (to_wall < steps) steps = to_wall;
    const int stride = 4 * dir;
    const uint32_t e_first = e_ref;
    const uint32_t t_stop = t_ref + (uint32_t)(stride * steps);
    float xc = n.x;
    // one crossing; false = the walk is over (collision, or the stop edge reached).  Like the lane kernel's loop it
    // carries neither x nor the cell (both are rebuilt from the edge reference afterwards) and is unrolled by four.
    auto step = [&]() -> bool {
        end = fadd(xc, n.ds);
        const float edge = c.edge(e_ref);
        const float t = fsub(xc, edge);
        if (!(fabsf(fsub(end, xc)) > fabsf(t))) return false; // collision at `end`
        c.score(t_ref, fabsf(fast_div(t, rc)));                // cross_mesh, src/mc_code.rs:171-181
        n.ds = fadd(n.ds, t);
        xc = edge;
        e_ref += (uint32_t)stride;
        t_ref += (uint32_t)stride;
        return t_ref != t_stop;
    };
    while (step() && step() && step() && step()) {}
    const int moved = (int)(e_ref - e_first) / 4; // signed cells travelled
    if (moved) {
        n.x = c.edge(e_ref - (uint32_t)stride);   // x after a crossing is the edge just crossed (src/mc_code.rs:72,77)
        n.cell += moved;
    }
    if (moved != dir * steps) {
        n.ds = end;
        return OUT_COLLIDE;
    }
    if (n.cell == run_exit) return OUT_MATCHANGE;
    n.pending = true;
    return OUT_PENDING;
}

// scat_mat_calc + interaction (src/mc_code.rs:82-132, 183-208).  false = absorbed.
template <int TG, class C> NRAPS_HD bool collide(C &c, const TransportParams &P, Neutron &n, Counts &ct)
{
    const int G = TG ? TG : (int)P.G, M = (int)P.M, N = (int)P.N;
    const float end = n.ds;
    const Recip rc = make_recip(n.mu);
    c.score(c.tally_ref(n.g * N + n.cell), fabsf(fast_div(fsub(n.x, end), rc)));
    ++ct.coll;
    const int xs = n.mat + M * n.xsg; // stale group index, src/mc_code.rs:147
    const float xi_int = pcg32_unit(n.rng, P.rng_inc);
    const float mu_new = fsub(fmul(2.0f, pcg32_unit(n.rng, P.rng_inc)), 1.0f);
    const int g_new = sample_group<TG>(c.scat_cdf(((n.mat * G + n.g) * G + n.xsg) * G), G, P.scatter_mode, n.rng, P.rng_inc);
    if (xi_int < c.p_abs(xs)) return false;
    n.x = end;
    n.g = g_new;
    n.mu = mu_new;
    if (!P.stale_xs) n.xsg = n.g;
    return true;
}

// Which of the two walk lists a neutron about to walk belongs to.  P.spawn_batch == 0 (default): by the length of the
// material run it sits in (> split cells = "long").  P.spawn_batch = T > 0: by a PREDICTION of how many cells its flight
// will cross in this run (>= T = "long"), from the distance to the edge ahead and the cell width; the prediction only
// sorts, so its rounding cannot change a result.
template <class C> NRAPS_HD uint32_t walk_class(C &c, const TransportParams &P, const Neutron &n, uint32_t split)
{
    const uint32_t rb = c.run_bounds(n.cell);
    const int run_lo = (int)(rb & 0xffffu), run_hi = (int)(rb >> 16);
    if (P.spawn_batch == 0u) return (uint32_t)(run_hi - run_lo) > split ? 1u : 0u;
    const int fwd = n.mu >= 0.0f ? 1 : 0;
    const int room = fwd ? run_hi - n.cell : n.cell - run_lo + 1; // cells up to the end of the run, this one included
    const float lo = c.edge(c.edge_ref(n.cell)), hi = c.edge(c.edge_ref(n.cell + 1));
    const float d0 = fwd ? hi - n.x : n.x - lo, a = fabsf(n.ds);
    if (!(a > d0)) return 0u;
    float k = 1.0f + (a - d0) / (hi - lo);
    const float cap = (float)(room < (int)P.walk_cap ? room : (int)P.walk_cap);
    k = k < cap ? k : cap;
    return k >= (float)P.spawn_batch ? 1u : 0u;
}

// the next history of the block's source range comes to life; false = none left right now
template <class C> NRAPS_HD bool adopt(C &c, const TransportParams &P, const Bank &b, Neutron &n)
{
    if (c.load_shared(&b.k[K_EXHAUSTED])) return false;
    const uint32_t idx = c.atomic_add_shared(&b.k[K_SRC_NEXT], 1u);
    if (idx >= c.load_shared(&b.k[K_SRC_END])) return false;
    uint32_t r0[4], r1[4];
    c.load_record(P.source + 2 * (uint64_t)idx, r0, r1);
    n.x = bits2f(r0[0]);
    n.mu = bits2f(r0[1]);
    n.cell = (int)(r0[2] & 0xffffu);
    n.g = (int)(r0[2] >> 16);
    n.rng = (uint64_t)r1[0] | ((uint64_t)r1[1] << 32);
    n.mat = c.material(n.cell);
    n.xsg = n.g;
    n.hf = 0;
    n.pending = false;
    n.ds = 0.0f;
    return true;
}

// The body every thread of the block runs.  `b` and the tables behind `c` are already set up; the tally image is zeroed.
template <int TG, class C> NRAPS_HD void block_event_thread(C &c, const TransportParams &P, const Bank &b, Counts &ct)
{
    const uint32_t tid = c.tid(), nthr = c.nthreads(), S = b.S;
    const uint32_t n_total = (uint32_t)(P.hist_end - P.hist_begin);
    // ---- set-up: the bank starts empty; material runs longer than half the longest run are class 1 ("long")
    if (tid == 0) {
        for (int i = 0; i < K_WORDS; ++i) b.k[i] = 0u;
        uint32_t longest = 1;
        for (uint32_t i = 0; i < P.N;) {
            const uint32_t rb = c.run_bounds((int)i);
            const uint32_t len = (rb >> 16) - (rb & 0xffffu);
            longest = len > longest ? len : longest;
            i = rb >> 16;
        }
        b.k[K_SPLIT] = longest / 2;
    }
    uint32_t p = 0; // parity of the lists phase AB reads
    for (;;) {
        c.sync(); // B1: every record phase C of the previous round wrote is visible
        if (tid == 0) {
            b.k[K_B0 + (p ^ 1)] = 0u; // the walk lists phase C of the previous round consumed
            const uint32_t live = (b.k[K_A0 + p] & 0xffffu) + (b.k[K_A0 + p] >> 16);
            if (!b.k[K_EXHAUSTED] && b.k[K_SRC_NEXT] >= b.k[K_SRC_END]) { // next chunk of histories for this block
                const unsigned long long base = c.atomic_add_global(P.work, (unsigned long long)P.chunk);
                if (base >= n_total) {
                    b.k[K_EXHAUSTED] = 1u;
                    b.k[K_SRC_NEXT] = b.k[K_SRC_END] = 0u;
                } else {
                    b.k[K_SRC_NEXT] = (uint32_t)base;
                    b.k[K_SRC_END] = base + P.chunk < n_total ? (uint32_t)(base + P.chunk) : n_total;
                }
            }
            // births this round: as many as the free capacity of the bank and the block's source range allow
            const uint32_t avail = b.k[K_SRC_NEXT] < b.k[K_SRC_END] ? b.k[K_SRC_END] - b.k[K_SRC_NEXT] : 0u, room = S - live;
            b.k[K_ADOPT] = avail < room ? avail : room;
            b.k[K_DONE] = (b.k[K_EXHAUSTED] && live == 0u) ? 1u : 0u;
            if (c.expired()) { // watchdog (device: ~10 s of SM clock): give up, the records still live count as truncated
                b.k[K_DONE] = 1u;
                ct.hist += live;
                ct.trunc += live;
            }
        }
        c.sync(); // B2
        if (c.load_shared(&b.k[K_DONE])) break;
        const uint32_t split = c.load_shared(&b.k[K_SPLIT]);
        const uint32_t k_a = c.load_shared(&b.k[K_A0 + p]);
        const uint32_t n_coll = k_a & 0xffffu, n_go = k_a >> 16, n_adopt = c.load_shared(&b.k[K_ADOPT]);
        // ---- phase AB: arena A -> arena B.  coll, then go, then births; every survivor lands in walk[class].
        const uint32_t n_ab = n_coll + n_go + n_adopt;
        if (C::kStats && tid == 0) c.note_round(n_coll, n_adopt, n_go);
        for (uint32_t base = 0; base < n_ab; base += nthr) {
            const uint32_t i = base + tid;
            bool alive = false;
            Neutron n{};
            if (i < n_ab) {
                if (i < n_coll) {
                    n = load_neutron(b.A, i);
                    alive = collide<TG>(c, P, n, ct);
                    if (!alive) {
                        ++ct.hist;
                        alive = adopt(c, P, b, n); // the freed record takes a new history at once, if there is one
                    }
                } else if (i < n_coll + n_go) {
                    n = load_neutron(b.A, S - 1u - (i - n_coll));
                    alive = true;
                } else {
                    alive = adopt(c, P, b, n);
                }
                if (alive && !flight(c, P, n, ct)) { // flight cap: the history ends here
                    ++ct.hist;
                    ++ct.trunc;
                    alive = false;
                }
            }
            uint32_t cls = 0;
            if (alive) cls = walk_class(c, P, n, split);
            c.converge();
            const uint32_t w = c.claim2(&b.k[K_B0 + p], alive && cls == 0, alive && cls == 1);
            if (alive) store_neutron(b.B, cls ? S - 1u - w : w, n);
        }
        c.sync(); // B3: arena B is complete, arena A is consumed
        if (tid == 0) b.k[K_A0 + p] = 0u; // phase C fills the lists of parity p^1
        // ---- phase C: arena B -> arena A.  One index space over both walk lists, long runs first (the longest work
        // starts first and every warp gets its share of both lists: no warp is left with long walks only)
        const uint32_t k_b = c.load_shared(&b.k[K_B0 + p]);
        const uint32_t n_long = k_b >> 16, n_walk = n_long + (k_b & 0xffffu);
        for (uint32_t base = 0; base < n_walk; base += nthr) {
            const uint32_t i = base + tid;
            int out = 0;
            Neutron n{};
            if (i < n_walk) {
                const uint32_t cls = i < n_long ? 1u : 0u, j = cls ? i : i - n_long;
                n = load_neutron(b.B, cls ? S - 1u - j : j);
                const int cell0 = n.cell;
                out = walk(c, P, n);
                if (C::kStats) c.note_walk(i, n.cell > cell0 ? n.cell - cell0 : cell0 - n.cell); // emulation only
                if (out == OUT_MATCHANGE) {
                    if ((unsigned)n.cell >= (unsigned)P.N) {
                        out = OUT_TRUNC; // unreachable for validated input
                    } else {
                        n.mat = c.material(n.cell);
                        n.xsg = n.g;
                    }
                }
                if (out == OUT_LEAK) { ++ct.hist; ++ct.leak; }
                else if (out == OUT_TRUNC) { ++ct.hist; ++ct.trunc; }
            }
            c.converge();
            const bool to_go = out == OUT_MATCHANGE || out == OUT_PENDING;
            const uint32_t q = c.claim2(&b.k[K_A0 + (p ^ 1)], out == OUT_COLLIDE, to_go);
            if (out == OUT_COLLIDE) store_neutron(b.A, q, n);
            else if (to_go) store_neutron(b.A, S - 1u - q, n);
        }
        p ^= 1;
    }
}

} // namespace bev
} // namespace nraps
