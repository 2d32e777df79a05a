// TEST INFRASTRUCTURE ONLY: runs the per-thread body of the block-level event pipeline
// (nraps_b200/csrc/mc_block_event.cuh, the code the CUDA kernel block_event_kernel executes) on CPU threads, one
// pthread per CUDA thread of a block, blocks one after another, so that its logic -- event lists, double buffering,
// adoption of source records, the restated walk / collide / flight stages -- can be bit-compared with the oracle on a
// machine without a GPU.  Nothing under nraps_b200/ links or loads this; it is built by tests/emul/Makefile.
//
// What it does NOT exercise: the CUDA-only context (warp-aggregated claims, shared-space PTX loads, the split 64-bit
// shared bins) -- the latter two are the lane kernels' own, GPU-verified helpers -- and source_kernel, whose births are
// restated here on the host (spawn_neutron + energy, src/mc_code.rs:7-53).
#include <pthread.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../nraps_b200/csrc/mc_block_event.cuh"

using namespace nraps;

namespace {

struct Image { // what a block keeps in shared memory, on the heap
    std::vector<unsigned long long> bins; // [G*N] 64-bit fixed-point bins (the device splits them in two u32 words)
    const float *edges;
    const uint32_t *runb;
    const uint8_t *matid;
    const float *xs;
    int MG;
};

// schedule statistics of one emulated block (what warps of 32 consecutive threads would see)
struct Stats {
    unsigned long long rounds = 0, coll = 0, dead = 0, fly = 0;     // entries phase AB consumed
    unsigned long long walk_entries = 0, crossings = 0;            // phase C
    unsigned long long warp_chunks = 0, warp_crossing_slots = 0;   // chunks of 32 entries of the walk index space; sum of the longest walk of each
    std::vector<int> cur;                                          // crossings by position in the walk index space, this round
};

struct HostCtx {
    static constexpr bool kStats = true;
    uint32_t tid_, nthr_;
    pthread_barrier_t *bar;
    Image *im;
    Stats *st;
    void note_walk(uint32_t i, int crossings) const
    {   // distinct positions are written by distinct threads; the vector is reset by thread 0 in note_round
        st->cur[i] = crossings;
    }
    void note_round(uint32_t n_coll, uint32_t n_dead, uint32_t n_fly) const
    {   // thread 0, after barrier B2: fold the walk phase of the previous round, then make room for this one
        std::vector<int> &v = st->cur;
        size_t n = v.size();
        while (n && v[n - 1] < 0) --n; // positions never written: fewer walks than records
        st->walk_entries += n;
        for (size_t b = 0; b < n; b += 32) {
            int mx = 0;
            for (size_t j = b; j < std::min(n, b + 32); ++j) { st->crossings += (unsigned long long)v[j]; mx = std::max(mx, v[j]); }
            st->warp_chunks += 1;
            st->warp_crossing_slots += (unsigned long long)std::max(mx, 1); // a walk executes the loop body at least once
        }
        std::fill(v.begin(), v.end(), -1);
        st->rounds += 1; st->coll += n_coll; st->dead += n_dead; st->fly += n_fly;
    }
    bool expired() const { return false; } // the device context has a clock watchdog here
    uint32_t tid() const { return tid_; }
    uint32_t nthreads() const { return nthr_; }
    // Mutation hook for the race detector run (tools/tsan_block_event.sh): BEV_EMUL_DROP_SYNC=k makes every thread skip
    // the k-th of the three barriers of a round (0 = B1, 1 = B2, 2 = B3), which ThreadSanitizer must then report.
    mutable unsigned long long n_sync = 0;
    int drop = -1;
    void sync() const
    {
        const unsigned long long k = n_sync++;
        if (drop >= 0 && (int)(k % 3ull) == drop) return;
        pthread_barrier_wait(bar);
    }
    void converge() const {}
    uint32_t atomic_add_shared(uint32_t *p, uint32_t v) const { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
    uint32_t load_shared(const uint32_t *p) const { return __atomic_load_n(p, __ATOMIC_SEQ_CST); }
    unsigned long long atomic_add_global(unsigned long long *p, unsigned long long v) const { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
    uint32_t claim2(uint32_t *word, bool up, bool down) const
    {   // the packed length word of an arena: low half = list growing up, high half = list growing down
        if (up) return atomic_add_shared(word, 1u) & 0xffffu;
        if (down) return atomic_add_shared(word, 0x10000u) >> 16;
        return 0u;
    }
    uint32_t run_bounds(int i) const { return im->runb[i]; }
    int material(int i) const { return (int)im->matid[i]; }
    uint32_t edge_ref(int i) const { return 4u * (uint32_t)i; }
    float edge(uint32_t ref) const { return im->edges[ref / 4u]; }
    uint32_t tally_ref(int bin) const { return 4u * (uint32_t)bin; }
    void score(uint32_t ref, float v) const
    {
        const float vs = v * kTallyScale;
        __atomic_fetch_add(&im->bins[ref / 4u], (unsigned long long)vs, __ATOMIC_RELAXED); // = __float2ull_rz for finite vs >= 0
    }
    float inv_sigtr(int i) const { return im->xs[i]; }
    float p_abs(int i) const { return im->xs[im->MG + i]; }
    const float *scat_cdf(int off) const { return im->xs + 5 * im->MG + off; }
    void load_record(const uint4 *rec, uint32_t (&r0)[4], uint32_t (&r1)[4]) const
    {
        r0[0] = rec[0].x; r0[1] = rec[0].y; r0[2] = rec[0].z; r0[3] = rec[0].w;
        r1[0] = rec[1].x; r1[1] = rec[1].y; r1[2] = rec[1].z; r1[3] = rec[1].w;
    }
};

struct ThreadArg {
    HostCtx ctx;
    const TransportParams *P;
    const bev::Bank *bank;
    bev::Counts counts;
    int G;
};

void *thread_main(void *arg)
{
    ThreadArg *a = static_cast<ThreadArg *>(arg);
    switch (a->G) {
    case 2: bev::block_event_thread<2>(a->ctx, *a->P, *a->bank, a->counts); break;
    case 4: bev::block_event_thread<4>(a->ctx, *a->P, *a->bank, a->counts); break;
    default: bev::block_event_thread<0>(a->ctx, *a->P, *a->bank, a->counts); break;
    }
    return nullptr;
}

// PCG32 master stream (src/rand.rs:49-85) and its jump-ahead, as mc_api.cu sets them up
struct Pcg { uint64_t state, inc; };
void step(Pcg &r) { r.state = r.state * NRAPS_PCG_MULT + r.inc; }
Pcg seed_pcg(uint64_t seed, uint64_t seq)
{
    Pcg r{0u, (seq << 1) | 1u};
    step(r);
    r.state += seed;
    step(r);
    return r;
}
uint64_t advance(uint64_t state, uint64_t inc, uint64_t delta)
{
    uint64_t cm = NRAPS_PCG_MULT, cp = inc, am = 1u, ap = 0u;
    while (delta) {
        if (delta & 1u) { am *= cm; ap = ap * cm + cp; }
        cp = (cm + 1u) * cp;
        cm *= cm;
        delta >>= 1;
    }
    return am * state + ap;
}

} // namespace

// The arithmetic of DevCtx::claim2 (mc_block_event.cu), lane by lane: given the two ballots and the value the packed word
// had, what the leader adds and where each of the 32 lanes lands.  pos_out[lane] is only meaningful for pushing lanes.
extern "C" uint32_t bev_emul_claim2(uint32_t m_up, uint32_t m_down, uint32_t old, uint32_t *pos_out)
{
    for (uint32_t lane = 0; lane < 32; ++lane)
        pos_out[lane] = bev::claim2_position(m_up, m_down, lane, old, ((m_down >> lane) & 1u) != 0);
    return old + bev::claim2_increment(m_up, m_down);
}

// One generation of [hist_begin, hist_begin + hist_count) through `n_blocks` emulated blocks of `n_threads` threads
// with `slots` neutron slots each.  tally_out: u64[G*N]; counters_out: u64[8] in NRAPS_CT order.
extern "C" int bev_emul_generation(const nraps_problem *p, uint64_t gen, uint64_t hist_begin, uint64_t hist_count, uint64_t seed,
                                   uint64_t stream, uint64_t stride, int32_t scatter_mode, int32_t stale_xs, uint32_t walk_cap,
                                   uint32_t max_flights, uint32_t n_blocks, uint32_t n_threads, uint32_t slots, uint32_t chunk,
                                   unsigned long long *tally_out, unsigned long long *counters_out, unsigned long long *stats_out)
{
    if (stats_out) std::fill(stats_out, stats_out + 8, 0ull);
    const uint32_t M = p->M, G = p->G, N = p->N, NF = p->NF, MG = M * G;
    // ---- derived tables, the expressions of mc_api.cu::create_ctx (binary32, reference order)
    std::vector<float> edges(N + 1);
    std::vector<uint32_t> runb(N);
    for (uint32_t i = 0; i < N; ++i) edges[i] = p->left[i];
    edges[N] = p->right[N - 1];
    for (uint32_t i = 0; i < N;) {
        uint32_t j = i;
        while (j < N && p->matid[j] == p->matid[i]) ++j;
        for (uint32_t q = i; q < j; ++q) runb[q] = i | (j << 16);
        i = j;
    }
    float *xs = static_cast<float *>(aligned_alloc(16, (xs_floats(M, G) * 4 + 15) / 16 * 16));
    std::fill(xs, xs + xs_floats(M, G), 0.0f);
    float *inv_sigtr = xs, *p_abs = xs + MG, *chi_cdf = p_abs + MG, *nusigf = chi_cdf + MG, *scat_cdf = xs + 5 * MG;
    for (uint32_t i = 0; i < MG; ++i) {
        inv_sigtr[i] = p->inv_sigtr[i];
        p_abs[i] = p->siga[i] / p->sigt[i];
        nusigf[i] = p->nut[i] * p->sigf[i];
    }
    for (uint32_t m = 0; m < M; ++m) {
        float cum = 0.0f;
        for (uint32_t g = 0; g < G; ++g) { cum = cum + p->chit[m + M * g]; chi_cdf[m * G + g] = cum; }
        for (uint32_t g = 0; g < G; ++g)
            for (uint32_t xg = 0; xg < G; ++xg) {
                const float inv_sigs = 1.0f / p->sigs[m + M * xg];
                float c2 = 0.0f;
                for (uint32_t j = 0; j < G; ++j) {
                    c2 = c2 + p->scat[G * G * m + G * g + j];
                    scat_cdf[((m * G + g) * G + xg) * G + j] = c2 * inv_sigs;
                }
            }
    }
    // ---- births of the shard (source_kernel restated: draw order cell, position, mu, chi)
    const Pcg master = seed_pcg(seed, stream);
    std::vector<uint4> source(2 * hist_count);
    for (uint64_t i = 0; i < hist_count; ++i) {
        uint64_t rng = advance(master.state, master.inc, (gen * p->histories + hist_begin + i) * stride);
        const uint32_t u = pcg32_next(rng, master.inc);
        const int cell = (int)p->fuel_indices[((uint64_t)u * NF) >> 32];
        const float xi_pos = pcg32_unit(rng, master.inc);
        const float mu = fsub(fmul(2.0f, pcg32_unit(rng, master.inc)), 1.0f);
        const float x = fadd(edges[cell], fmul(xi_pos, p->dx_fuel));
        const int g = lower_bound_clamped<0>(chi_cdf + p->matid[cell] * G, (int)G, pcg32_unit(rng, master.inc));
        source[2 * i] = make_uint4(f2bits(x), f2bits(mu), (uint32_t)cell | ((uint32_t)g << 16), 0u);
        source[2 * i + 1] = make_uint4((uint32_t)rng, (uint32_t)(rng >> 32), 0u, 0u);
    }
    unsigned long long work = 0;
    TransportParams P{};
    P.source = source.data();
    P.M = M; P.G = G; P.N = N; P.NF = NF; P.rows = G;
    P.boundl = p->boundl; P.boundr = p->boundr; P.dx_fuel = p->dx_fuel;
    P.rng_inc = master.inc;
    P.hist_begin = hist_begin; P.hist_end = hist_begin + hist_count;
    P.work = &work;
    P.chunk = chunk; P.max_flights = max_flights; P.walk_cap = walk_cap;
    P.scatter_mode = scatter_mode; P.stale_xs = stale_xs;
    if (const char *t = getenv("BEV_EMUL_CLASS_T")) P.spawn_batch = (uint32_t)atoi(t); // walk-class threshold (schedule studies)

    std::fill(tally_out, tally_out + (size_t)G * N, 0ull);
    std::fill(counters_out, counters_out + 8, 0ull);
    for (uint32_t blk = 0; blk < n_blocks; ++blk) { // blocks share nothing but the work counter: one after another
        Image im;
        im.bins.assign((size_t)G * N, 0ull);
        im.edges = edges.data(); im.runb = runb.data(); im.matid = p->matid; im.xs = xs; im.MG = (int)MG;
        std::vector<unsigned char> raw(bev::bank_bytes(slots) + 16);
        unsigned char *base = raw.data() + ((16 - ((uintptr_t)raw.data() & 15u)) & 15u);
        const bev::Bank bank = bev::make_bank(base, slots);
        Stats stats;
        stats.cur.assign(slots, -1);
        pthread_barrier_t bar;
        pthread_barrier_init(&bar, nullptr, n_threads);
        std::vector<ThreadArg> args(n_threads);
        std::vector<pthread_t> th(n_threads);
        for (uint32_t t = 0; t < n_threads; ++t) {
            args[t].ctx = HostCtx{t, n_threads, &bar, &im, &stats};
            if (const char *d = getenv("BEV_EMUL_DROP_SYNC")) args[t].ctx.drop = atoi(d);
            args[t].P = &P; args[t].bank = &bank; args[t].G = (int)G;
            if (pthread_create(&th[t], nullptr, thread_main, &args[t]) != 0) return -1;
        }
        for (uint32_t t = 0; t < n_threads; ++t) pthread_join(th[t], nullptr);
        pthread_barrier_destroy(&bar);
        for (size_t i = 0; i < (size_t)G * N; ++i) tally_out[i] += im.bins[i];
        if (stats_out) { // [rounds, coll, births, go, walk entries, crossings, warp chunks, warp crossing slots]
            const unsigned long long v[8] = {stats.rounds, stats.coll, stats.dead, stats.fly, stats.walk_entries, stats.crossings,
                                             stats.warp_chunks, stats.warp_crossing_slots};
            for (int i = 0; i < 8; ++i) stats_out[i] += v[i];
        }
        for (uint32_t t = 0; t < n_threads; ++t) {
            const bev::Counts &c = args[t].counts;
            counters_out[NRAPS_CT_HISTORIES] += c.hist; counters_out[NRAPS_CT_COLLISIONS] += c.coll;
            counters_out[NRAPS_CT_FLIGHTS] += c.flight; counters_out[NRAPS_CT_LEAKS] += c.leak;
            counters_out[NRAPS_CT_TRUNCATED] += c.trunc;
        }
    }
    free(xs);
    return 0;
}
