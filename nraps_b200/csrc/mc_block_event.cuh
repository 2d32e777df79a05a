// Block-level event pipeline for surface tracking (kernel_variant = NRAPS_KERNEL_BLOCK_EVENT): EXPERIMENTAL.
//
// Same physics, same per-history RNG streams and the same integer tallies as transport_kernel (mc_transport.cu), so
// the results are bit-identical; what changes is who executes which event.  In the lane kernel a neutron lives in
// the registers of one lane for its whole life and a warp trip runs flight / walk / collide for whatever mix of
// neutrons its lanes hold (ncu, profiles/r1r: 18 of 32 lanes busy; the collide stage runs with 10 lanes, a walk lasts
// as long as the longest of the warp, 8 cells in a fuel pin against 4 in a water gap).  Here the neutrons of a block
// live in shared memory (28 bytes each, structure of arrays, several per thread) and every round sorts them by what
// they need next:
//
//   phase AB   list `coll`  : collide (src/mc_code.rs:182-210), then the survivor's next flight draw (:147-148)
//              list `dead`  : adopt a history born by source_kernel (src/mc_code.rs:40-53), then its first flight
//              list `fly`   : flight draw only (the neutron entered another material run, :179-181)
//   phase C    list `walk[0]`, `walk[1]` (short / long material runs): boundary + cross_mesh loop (:151-181, 56-79)
//
// so that the lanes of a warp run the same stage on neutrons with walks of similar length.  That is the north star's
// event-based pipeline over structure-of-arrays banks, kept on chip: the HBM version of it (mc_event.cu) lost 8x to
// bank round trips and launch tails (profiles/r1e); 227 KB of shared memory per SM hold the bank instead.
//
// The per-thread body below is written against a small context type C (thread index, barrier, shared / global
// atomics, warp-aggregated list claims, shared-space loads of the mesh tables and the tally score), so that the very
// same code runs as a CUDA block (mc_block_event.cu) and, for tests only, on CPU threads (tests/emul) where it is
// bit-compared with the oracle without a GPU.  Restrictions of this variant: uniform source, no trace records, no
// generation batching, mesh image in shared memory (no BIG mode); everything else falls back to NRAPS_ERR_OPTION.
#pragma once
#include "mc_lane.cuh"

namespace nraps {
namespace bev {

enum { OUT_COLLIDE = 1, OUT_MATCHANGE = 2, OUT_PENDING = 3, OUT_LEAK = 4, OUT_TRUNC = 5 };

// counters in shared memory (u32 words)
enum {
    K_COLL0 = 0, K_COLL1, K_FLY0, K_FLY1, K_DEAD0, K_DEAD1, // list lengths, one per parity
    K_WALK00, K_WALK01, K_WALK10, K_WALK11,                // walk[parity][class]
    K_SRC_NEXT, K_SRC_END, K_EXHAUSTED, K_DONE, K_SPLIT,
    K_WORDS = 16
};

// the block's neutron bank and event lists (pointers into shared memory, or into a heap image on the host)
struct Bank {
    float *x, *mu, *ds;          // [S] position, direction cosine, signed distance left to the collision site (or the site itself)
    uint32_t *pk, *rlo, *rhi, *hf; // [S] cell | g<<16 | xsg<<19 | mat<<22 | pending<<28, rng state, flights so far
    uint16_t *lists;             // ten slot lists of [S] each, addressed arithmetically (a pointer table indexed by the
                                 // parity would live in local memory): coll[2] | fly[2] | dead[2] | walk[2][2]
    uint32_t *k;                 // [K_WORDS]
    uint32_t S;
    NRAPS_HD uint16_t *coll(uint32_t p) const { return lists + (0u + p) * S; }
    NRAPS_HD uint16_t *fly(uint32_t p) const { return lists + (2u + p) * S; }
    NRAPS_HD uint16_t *dead(uint32_t p) const { return lists + (4u + p) * S; }
    NRAPS_HD uint16_t *walk(uint32_t p, uint32_t cls) const { return lists + (6u + 2u * p + cls) * S; }
};

__host__ __device__ inline uint32_t bank_bytes(uint32_t S) { return S * (7 * 4 + 10 * 2) + K_WORDS * 4; }

// carve a Bank out of `raw` (16-byte aligned)
__host__ __device__ inline Bank make_bank(unsigned char *raw, uint32_t S)
{
    Bank b;
    b.S = S;
    b.x = reinterpret_cast<float *>(raw);
    b.mu = b.x + S;
    b.ds = b.mu + S;
    b.pk = reinterpret_cast<uint32_t *>(b.ds + S);
    b.rlo = b.pk + S;
    b.rhi = b.rlo + S;
    b.hf = b.rhi + S;
    b.k = b.hf + S;
    b.lists = reinterpret_cast<uint16_t *>(b.k + K_WORDS);
    return b;
}

struct Neutron {
    float x, mu, ds;
    int cell, g, xsg, mat;
    bool pending;
    uint64_t rng;
    uint32_t hf;
};

NRAPS_HD Neutron load_neutron(const Bank &b, uint32_t s)
{
    Neutron n;
    n.x = b.x[s]; n.mu = b.mu[s]; n.ds = b.ds[s];
    const uint32_t pk = b.pk[s];
    n.cell = (int)(pk & 0xffffu); n.g = (int)((pk >> 16) & 7u); n.xsg = (int)((pk >> 19) & 7u); n.mat = (int)((pk >> 22) & 63u);
    n.pending = ((pk >> 28) & 1u) != 0;
    n.rng = (uint64_t)b.rlo[s] | ((uint64_t)b.rhi[s] << 32);
    n.hf = b.hf[s];
    return n;
}

NRAPS_HD void store_neutron(const Bank &b, uint32_t s, const Neutron &n)
{
    b.x[s] = n.x; b.mu[s] = n.mu; b.ds[s] = n.ds;
    b.pk[s] = (uint32_t)n.cell | ((uint32_t)n.g << 16) | ((uint32_t)n.xsg << 19) | ((uint32_t)n.mat << 22) | ((n.pending ? 1u : 0u) << 28);
    b.rlo[s] = (uint32_t)n.rng; b.rhi[s] = (uint32_t)(n.rng >> 32);
    b.hf[s] = n.hf;
}

struct Counts {
    uint32_t hist = 0, coll = 0, flight = 0, leak = 0, trunc = 0;
};

// src/mc_code.rs:147-148 (and :209): one flight draw.  false = the flight cap truncated the history.
template <class C> NRAPS_HD bool flight(C &c, const TransportParams &P, Neutron &n, Counts &ct)
{
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

// a dead slot takes the next history of the block's source range; false = none left right now
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
    // ---- set-up: every slot starts dead; material runs longer than half the longest run are class 1 ("long")
    for (uint32_t i = tid; i < S; i += nthr) b.dead(0)[i] = (uint16_t)i;
    if (tid == 0) {
        for (int i = 0; i < K_WORDS; ++i) b.k[i] = 0u;
        b.k[K_DEAD0] = S;
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
        c.sync(); // B1: every push of the previous round is visible
        if (tid == 0) {
            b.k[K_WALK00 + 2 * (p ^ 1)] = 0u; // the walk lists phase C of the previous round consumed
            b.k[K_WALK01 + 2 * (p ^ 1)] = 0u;
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
            b.k[K_DONE] = (b.k[K_EXHAUSTED] && b.k[K_DEAD0 + p] == S) ? 1u : 0u;
        }
        c.sync(); // B2
        if (c.load_shared(&b.k[K_DONE])) break;
        const uint32_t split = c.load_shared(&b.k[K_SPLIT]);
        const uint32_t n_coll = c.load_shared(&b.k[K_COLL0 + p]), n_fly = c.load_shared(&b.k[K_FLY0 + p]),
                       n_dead = c.load_shared(&b.k[K_DEAD0 + p]);
        // ---- phase AB: the three lists one after another; every entry ends in walk[p][class] or in dead[p^1]
        const uint32_t n_ab = n_coll + n_dead + n_fly;
        if (C::kStats && tid == 0) c.note_round(n_coll, n_dead, n_fly);
        for (uint32_t base = 0; base < n_ab; base += nthr) {
            const uint32_t i = base + tid;
            const bool active = i < n_ab;
            bool alive = false;
            uint32_t slot = 0;
            Neutron n{};
            if (active) {
                if (i < n_coll) {
                    slot = b.coll(p)[i];
                    n = load_neutron(b, slot);
                    alive = collide<TG>(c, P, n, ct);
                    if (!alive) {
                        ++ct.hist;
                        alive = adopt(c, P, b, n);
                    }
                } else if (i < n_coll + n_dead) {
                    slot = b.dead(p)[i - n_coll];
                    alive = adopt(c, P, b, n);
                } else {
                    slot = b.fly(p)[i - n_coll - n_dead];
                    n = load_neutron(b, slot);
                    alive = true;
                }
                if (alive && !flight(c, P, n, ct)) { // flight cap: the history ends here
                    ++ct.hist;
                    ++ct.trunc;
                    alive = false;
                }
            }
            uint32_t cls = 0;
            if (alive) {
                const uint32_t rb = c.run_bounds(n.cell);
                cls = ((rb >> 16) - (rb & 0xffffu)) > split ? 1u : 0u;
                store_neutron(b, slot, n);
            }
            c.converge();
            const uint32_t w0 = c.claim(&b.k[K_WALK00 + 2 * p], alive && cls == 0);
            const uint32_t w1 = c.claim(&b.k[K_WALK01 + 2 * p], alive && cls == 1);
            const uint32_t d = c.claim(&b.k[K_DEAD0 + (p ^ 1)], active && !alive);
            if (alive) b.walk(p, cls)[cls ? w1 : w0] = (uint16_t)slot;
            else if (active) b.dead(p ^ 1)[d] = (uint16_t)slot;
        }
        c.sync(); // B3: the walk lists are complete
        if (tid == 0) b.k[K_COLL0 + p] = b.k[K_FLY0 + p] = b.k[K_DEAD0 + p] = 0u; // consumed; phase C writes parity p^1 only
        // ---- phase C: walks, short runs first, long runs second; outcomes go to the lists of parity p^1
        for (uint32_t cls = 0; cls < 2; ++cls) {
            const uint32_t n_walk = c.load_shared(&b.k[K_WALK00 + 2 * p + cls]);
            for (uint32_t base = 0; base < n_walk; base += nthr) {
                const uint32_t i = base + tid;
                const bool active = i < n_walk;
                int out = 0;
                uint32_t slot = 0;
                if (active) {
                    slot = b.walk(p, cls)[i];
                    Neutron n = load_neutron(b, slot);
                    const int cell0 = n.cell;
                    out = walk(c, P, n);
                    if (C::kStats) c.note_walk(cls, i, n.cell > cell0 ? n.cell - cell0 : cell0 - n.cell); // emulation only
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
                    else store_neutron(b, slot, n);
                }
                c.converge();
                const uint32_t qc = c.claim(&b.k[K_COLL0 + (p ^ 1)], out == OUT_COLLIDE);
                const uint32_t qf = c.claim(&b.k[K_FLY0 + (p ^ 1)], out == OUT_MATCHANGE);
                const uint32_t qw = c.claim(&b.k[K_WALK00 + 2 * (p ^ 1) + cls], out == OUT_PENDING);
                const uint32_t qd = c.claim(&b.k[K_DEAD0 + (p ^ 1)], out == OUT_LEAK || out == OUT_TRUNC);
                if (out == OUT_COLLIDE) b.coll(p ^ 1)[qc] = (uint16_t)slot;
                else if (out == OUT_MATCHANGE) b.fly(p ^ 1)[qf] = (uint16_t)slot;
                else if (out == OUT_PENDING) b.walk(p ^ 1, cls)[qw] = (uint16_t)slot;
                else if (out == OUT_LEAK || out == OUT_TRUNC) b.dead(p ^ 1)[qd] = (uint16_t)slot;
            }
        }
        p ^= 1;
    }
}

} // namespace bev
} // namespace nraps
