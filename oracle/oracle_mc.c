/*
 * oracle_mc.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Line-by-line CPU restatement of the reference Monte Carlo k-eigenvalue loop,
 * /root/reference/src/mc_code.rs:7-380, with every result-changing quirk of
 * SURVEY.md section 9 behind an explicit switch.  All transport arithmetic is
 * IEEE binary32 in the reference's operation order (compile with
 * -ffp-contract=off); the only fused operations are the explicit fmaf() calls
 * inside oracle_logf().
 *
 * Deviations from HEAD that are forced, not chosen:
 *   - RNG: per-history PCG32 streams (src/rand.rs algorithm + advance) instead
 *     of the unseedable thread_rng (mc_code.rs:46-51,122,125,148,195,209).
 *   - uniforms never hit 0 / 0.5 / 1 (Q2), so no NaN tallies.
 *   - source cell: fuel[(u32 * NF) >> 32] instead of rand's gen_range.
 * "parity unpinned": nothing in the reference fixes k or flux numerically.
 */
#define _GNU_SOURCE
#include "oracle.h"
#include "oracle_math.h"

#include <pthread.h>
#include <stdlib.h>
#include <time.h>
#include <unistd.h>

#define FIXED_SCALE 268435456.0f /* 2^ORACLE_TALLY_FRAC_BITS */

#if defined(__x86_64__) && defined(__GNUC__) && !defined(ORACLE_NO_CLONES)
#define ORACLE_HOT __attribute__((target_clones("arch=x86-64-v3", "default")))
#else
#define ORACLE_HOT
#endif

typedef struct {
    const oracle_problem *p;
    const oracle_options *o;
    oracle_pcg32 master;
    uint64_t gen;
    uint64_t max_flights;
    /* fission_bank source mode (new capability, no reference counterpart) */
    int bank_mode;
    float inv_k;               /* 1 / k of the previous generation: keeps the bank near H sites */
    const uint64_t *src_bank;  /* dense bank feeding this generation, NULL => uniform_fuel source */
    uint64_t src_count;
    uint64_t *slots;           /* [hist_count][bank_cap] sites of this generation, by history */
    uint8_t *counts;           /* [hist_count] */
    uint32_t bank_cap;
    uint64_t slot_base;        /* history index of slot row 0 */
    /* Woodcock delta tracking (new capability named by the north star) */
    const float *sigtr;        /* [M*G]  sigt - mu*sigs, the collision density flights are sampled with */
    const float *inv_maj;      /* [G*G]  1 / max over present materials of max(sigtr[m][a], sigtr[m][b]) */
    const uint32_t *run_lo, *run_hi; /* [N] cell bounds of each cell's material run */
} run_shared;

typedef struct {
    const run_shared *sh;
    uint64_t start, end; /* [start, end) within the generation */
    uint64_t *tally_fixed; /* [G][N] */
    float *tally_f32;      /* [G][N] */
    uint64_t counters[ORACLE_CT_WORDS];
    uint32_t *trace;       /* base of the generation's trace array or NULL */
    uint64_t trace_base;   /* history index of trace row 0 */
} worker;

/* optional diagnostics for kernel design (single-threaded runs only): crossings-per-flight histogram by group */
static int g_diag = 0;
static uint64_t g_diag_hist[8][130];
static uint64_t g_diag_coll[8];
void oracle_diag_enable(int on) { g_diag = on; memset(g_diag_hist, 0, sizeof g_diag_hist); memset(g_diag_coll, 0, sizeof g_diag_coll); }
void oracle_diag_read(uint64_t *hist /*[8][130]*/, uint64_t *coll /*[8]*/) { memcpy(hist, g_diag_hist, sizeof g_diag_hist); memcpy(coll, g_diag_coll, sizeof g_diag_coll); }

/* ---- reference helpers, one function per reference function ------------- */

/* src/mc_code.rs:56-62 */
void oracle_hit_boundary(float mu, float start_x, float delta_s, float bound, float mesh_end, float out[3])
{
    out[0] = mu * (-bound);
    out[1] = (delta_s + (start_x - mesh_end)) * (-bound);
    out[2] = mesh_end;
}

/* src/mc_code.rs:65-79 */
void oracle_cross_mesh(uint64_t mesh_index, float mu, float start_x, float mesh_end, float delta_s,
                       float out_ds_x[2], uint64_t *out_index)
{
    *out_index = (mu >= 0.0f) ? mesh_index + 1 : mesh_index - 1;
    out_ds_x[0] = delta_s + (start_x - mesh_end);
    out_ds_x[1] = mesh_end;
}

float oracle_direction_f(float xi) { return oracle_direction(xi); }

/* src/mc_code.rs:82-111 (index arithmetic widened from u8, Q13) */
void oracle_scat_mat_calc(uint32_t G, uint32_t matid, uint32_t g, float inv_sigs, const float *scat, float *out)
{
    uint32_t base = G * G * matid + G * g;
    float cumulative = 0.0f;
    for (uint32_t j = 0; j < G; ++j) {
        cumulative += scat[base + j];
        out[j] = cumulative * inv_sigs;
    }
}

uint32_t oracle_energy_search(const float *cumulative, uint32_t n, float chi)
{
    return oracle_lower_bound_clamped(cumulative, n, chi);
}

/* src/mc_code.rs:7-32 */
uint32_t oracle_energy(const oracle_problem *p, float chi, uint64_t cell)
{
    float cum[256];
    float c = 0.0f;
    uint32_t mat = p->matid[cell];
    for (uint32_t g = 0; g < p->G; ++g) {
        c += p->chit[mat + p->M * g];
        cum[g] = c;
    }
    return oracle_lower_bound_clamped(cum, p->G, chi);
}

void oracle_pcg32_demo(uint64_t seed, uint64_t seq, uint32_t n, uint32_t *out)
{
    oracle_pcg32 r;
    oracle_pcg32_seed(&r, seed, seq);
    for (uint32_t i = 0; i < n; ++i) out[i] = oracle_pcg32_next(&r);
}

void oracle_pcg32_state(uint64_t seed, uint64_t seq, uint64_t delta, uint64_t out_state_inc[2])
{
    oracle_pcg32 r;
    oracle_pcg32_seed(&r, seed, seq);
    oracle_pcg32_advance(&r, delta);
    out_state_inc[0] = r.state;
    out_state_inc[1] = r.inc;
}

float oracle_logf_f(float x) { return oracle_logf(x); }
float oracle_unit_f(uint32_t u) { return oracle_u32_to_unit(u); }

/* max |oracle_logf - log| in ulps over the uniforms k = first .. first+count-1 */
double oracle_logf_max_ulp(uint32_t first, uint32_t count)
{
    double worst = 0.0;
    for (uint32_t i = 0; i < count; ++i) {
        uint32_t k = first + i;
        float x = ((float)k + 0.5f) * 1.1920928955078125e-07f;
        float got = oracle_logf(x);
        double want = log((double)x);
        float wf = (float)want;
        float up = nextafterf(fabsf(wf), INFINITY);
        double ulp = (double)up - (double)fabsf(wf);
        double err = fabs((double)got - want) / ulp;
        if (err > worst) worst = err;
    }
    return worst;
}

/* group-transfer sampling, src/mc_code.rs:124-126, three readings of Q3 */
static inline __attribute__((always_inline)) uint32_t sample_group(const float *cdf, uint32_t G, int mode, oracle_pcg32 *rng)
{
    if (mode == ORACLE_SCATTER_SINGLE_XI) {
        float xi = oracle_uniform(rng);
        return oracle_lower_bound_clamped(cdf, G, xi);
    }
    if (mode == ORACLE_SCATTER_RUST_PRE182) {
        /* core::slice::binary_search_by before rustc 1.82 */
        uint32_t size = G, left = 0, right = G;
        while (left < right) {
            uint32_t mid = left + size / 2;
            float xi = oracle_uniform(rng);
            if (cdf[mid] < xi)
                left = mid + 1;
            else
                right = mid;
            size = right - left;
        }
        return left < G - 1 ? left : G - 1;
    }
    /* rustc >= 1.82 branchless variant */
    uint32_t size = G, base = 0;
    while (size > 1) {
        uint32_t half = size / 2, mid = base + half;
        float xi = oracle_uniform(rng);
        int less = cdf[mid] < xi; /* pred true -> Less ; false -> Greater */
        base = less ? mid : base;
        size -= half;
    }
    float xi = oracle_uniform(rng);
    uint32_t res = base + ((cdf[base] < xi) ? 1u : 0u);
    return res < G - 1 ? res : G - 1;
}

static inline __attribute__((always_inline)) void score(worker *w, uint32_t g, uint64_t cell, float v)
{
    const oracle_problem *p = w->sh->p;
    if (w->tally_fixed)
        w->tally_fixed[(uint64_t)g * p->N + cell] += (uint64_t)(v * FIXED_SCALE);
    else
        w->tally_f32[(uint64_t)g * p->N + cell] += v;
}

/*
 * One history: spawn_neutron (mc_code.rs:40-53, 228-230) followed by
 * particle_travel calls until death (mc_code.rs:232-254, 134-213).
 */
static inline __attribute__((always_inline)) void run_history(worker *w, uint64_t y)
{
    const run_shared *sh = w->sh;
    const oracle_problem *p = sh->p;
    const oracle_options *o = sh->o;
    const uint32_t G = p->G, M = p->M;
    const uint64_t N = p->N;

    oracle_pcg32 rng = sh->master;
    uint64_t hid = sh->gen * p->histories + y;
    oracle_pcg32_advance(&rng, hid * o->stride);

    uint32_t u = oracle_pcg32_next(&rng);
    uint64_t cell;
    float x, mu;
    uint32_t g;
    if (sh->src_bank) {
        /* fission_bank source: site index, mu, chi (no position draw) */
        uint64_t site = sh->src_bank[((uint64_t)u * sh->src_count) >> 32];
        uint32_t xb = (uint32_t)site;
        cell = site >> 32;
        memcpy(&x, &xb, 4);
        mu = oracle_direction(oracle_uniform(&rng));
        g = oracle_energy(p, oracle_uniform(&rng), cell);
    } else {
        /* draw order: cell, position, mu, chi */
        cell = p->fuel_indices[((uint64_t)u * (uint64_t)p->NF) >> 32];
        float xi_pos = oracle_uniform(&rng);
        mu = oracle_direction(oracle_uniform(&rng));
        g = oracle_energy(p, oracle_uniform(&rng), cell);
        x = p->left[cell] + (xi_pos * p->dx_fuel);
    }
    uint32_t n_bank = 0;

    uint32_t n_coll = 0, n_cross = 0, n_flight = 0, n_refl = 0, fate = 0;
    const uint64_t max_flights = sh->max_flights;
    float cdf[256];

    int alive = 1;
    while (alive) {
        /* ---- particle_travel, mc_code.rs:147-148 */
        if (n_flight >= max_flights) { fate = ORACLE_FATE_TRUNCATED; break; } /* safety cap, not in the reference */
        uint32_t mat = p->matid[cell];
        uint32_t xs = mat + M * g;
        float ds = mu * -oracle_logf(oracle_uniform(&rng)) * p->inv_sigtr[xs];
        ++n_flight;
        uint32_t diag_k = 0;
        for (;;) {
            float end_x = x + ds;
            float mesh_end = (mu >= 0.0f) ? p->right[cell] : p->left[cell];
            if ((mu < 0.0f && mesh_end > end_x && cell == 0) ||
                (mu >= 0.0f && end_x > mesh_end && cell == N - 1)) {
                /* mc_code.rs:162-170 */
                score(w, g, cell, fabsf((x - mesh_end) / mu));
                float bound = (mu >= 0.0f) ? p->boundr : p->boundl;
                if (bound > 0.0f) {
                    float r[3];
                    oracle_hit_boundary(mu, x, ds, bound, mesh_end, r);
                    mu = r[0]; ds = r[1]; x = r[2];
                    ++n_refl;
                } else {
                    alive = 0; fate = ORACLE_FATE_LEAKED;
                    break;
                }
            } else if (fabsf(end_x - x) > fabsf(mesh_end - x)) {
                /* mc_code.rs:172-181 */
                score(w, g, cell, fabsf((x - mesh_end) / mu));
                uint32_t prev_mat = p->matid[cell];
                float dsx[2];
                oracle_cross_mesh(cell, mu, x, mesh_end, ds, dsx, &cell);
                ds = dsx[0]; x = dsx[1];
                ++n_cross; ++diag_k;
                if (prev_mat != p->matid[cell]) { if (g_diag) g_diag_hist[g & 7][diag_k > 129 ? 129 : diag_k]++; break; } /* alive, new flight */
            } else {
                /* mc_code.rs:183-209 */
                score(w, g, cell, fabsf((x - end_x) / mu));
                ++n_coll;
                if (g_diag) { g_diag_hist[g & 7][diag_k > 129 ? 129 : diag_k]++; g_diag_coll[g & 7]++; diag_k = 0; }
                oracle_scat_mat_calc(G, p->matid[cell], g, 1.0f / p->sigs[xs], p->scat, cdf);
                float xi_int = oracle_uniform(&rng);
                float absorption = p->siga[xs] / p->sigt[xs];
                float mu_new = 2.0f * oracle_uniform(&rng) - 1.0f;
                uint32_t g_new = sample_group(cdf, G, o->scatter_mode, &rng);
                if (sh->bank_mode) {
                    /* collision estimator of the fission source: nu*Sigma_f(mat, g) per unit track length over the
                     * collision density 1/inv_sigtr[xs] the flight was sampled with, normalised by the last k */
                    float nusigf = p->nut[p->matid[cell] + M * g] * p->sigf[p->matid[cell] + M * g];
                    if (nusigf > 0.0f) {
                        float wgt = nusigf * p->inv_sigtr[xs] * sh->inv_k;
                        uint32_t n = (uint32_t)(int32_t)(wgt + oracle_uniform(&rng));
                        uint32_t xb;
                        memcpy(&xb, &end_x, 4);
                        for (uint32_t j = 0; j < n; ++j) {
                            if (n_bank < sh->bank_cap)
                                sh->slots[(y - sh->slot_base) * sh->bank_cap + n_bank] = ((uint64_t)cell << 32) | xb;
                            ++n_bank;
                        }
                    }
                }
                if (xi_int < absorption) {
                    alive = 0; fate = ORACLE_FATE_ABSORBED;
                    break;
                }
                x = end_x;
                g = g_new;
                mu = mu_new;
                if (!o->stale_xs) xs = p->matid[cell] + M * g; /* Q1 "fixed" */
                if (n_flight >= max_flights) { alive = 0; fate = ORACLE_FATE_TRUNCATED; break; }
                ds = mu * -oracle_logf(oracle_uniform(&rng)) * p->inv_sigtr[xs];
                ++n_flight;
            }
        }
    }

    w->counters[ORACLE_CT_HISTORIES] += 1;
    w->counters[ORACLE_CT_COLLISIONS] += n_coll;
    w->counters[ORACLE_CT_CROSSINGS] += n_cross;
    w->counters[ORACLE_CT_FLIGHTS] += n_flight;
    w->counters[ORACLE_CT_REFLECTIONS] += n_refl;
    w->counters[ORACLE_CT_LEAKS] += (fate == ORACLE_FATE_LEAKED);
    w->counters[ORACLE_CT_TRUNCATED] += (fate == ORACLE_FATE_TRUNCATED);
    if (sh->bank_mode) {
        uint32_t kept = n_bank < sh->bank_cap ? n_bank : sh->bank_cap;
        sh->counts[y - sh->slot_base] = (uint8_t)kept;
        w->counters[ORACLE_CT_BANKED] += kept;
    }
    if (w->trace) {
        uint32_t *t = w->trace + (y - w->trace_base) * ORACLE_TR_WORDS;
        uint32_t xb;
        memcpy(&xb, &x, 4);
        t[ORACLE_TR_COLLISIONS] = n_coll;
        t[ORACLE_TR_CROSSINGS] = n_cross;
        t[ORACLE_TR_FLIGHTS] = n_flight;
        t[ORACLE_TR_REFLECTIONS] = n_refl;
        t[ORACLE_TR_RNG_LO] = (uint32_t)rng.state;
        t[ORACLE_TR_RNG_HI] = (uint32_t)(rng.state >> 32);
        t[ORACLE_TR_CELL] = (uint32_t)cell;
        t[ORACLE_TR_XBITS] = xb;
        t[ORACLE_TR_FATE] = fate;
        t[ORACLE_TR_GROUP] = g;
    }
}

/* cell containing x: number of interior edges <= x */
static inline uint64_t locate_cell(const oracle_problem *p, float x)
{
    uint64_t lo = 0, hi = p->N - 1; /* interior edges are right[0 .. N-2] */
    while (lo < hi) {
        uint64_t mid = lo + (hi - lo) / 2;
        if (p->right[mid] <= x) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

/*
 * One history under Woodcock delta tracking with a collision-estimator tally.
 * Statistically equivalent to run_history (same source, same collision physics
 * including the stale cross-section group of SURVEY 9-Q1: the group index set on
 * entering a material run is kept until the neutron leaves that run), but no
 * surface crossings: flights are sampled against the majorant of the two groups
 * in play, every tentative collision scores 1/Sigma_maj (an unbiased track-length
 * estimate), and it is accepted as real with probability sigtr/Sigma_maj.
 */
static inline __attribute__((always_inline)) void run_history_woodcock(worker *w, uint64_t y)
{
    const run_shared *sh = w->sh;
    const oracle_problem *p = sh->p;
    const oracle_options *o = sh->o;
    const uint32_t G = p->G, M = p->M;
    const uint64_t N = p->N;
    const float L = p->right[N - 1];

    oracle_pcg32 rng = sh->master;
    oracle_pcg32_advance(&rng, (sh->gen * p->histories + y) * o->stride);
    uint32_t u = oracle_pcg32_next(&rng);
    uint64_t cell;
    float x, mu;
    uint32_t g;
    if (sh->src_bank) {
        uint64_t site = sh->src_bank[((uint64_t)u * sh->src_count) >> 32];
        uint32_t xb = (uint32_t)site;
        cell = site >> 32;
        memcpy(&x, &xb, 4);
        mu = oracle_direction(oracle_uniform(&rng));
        g = oracle_energy(p, oracle_uniform(&rng), cell);
    } else {
        cell = p->fuel_indices[((uint64_t)u * (uint64_t)p->NF) >> 32];
        float xi_pos = oracle_uniform(&rng);
        mu = oracle_direction(oracle_uniform(&rng));
        g = oracle_energy(p, oracle_uniform(&rng), cell);
        x = p->left[cell] + (xi_pos * p->dx_fuel);
    }
    uint32_t xsg = g, home_lo = sh->run_lo[cell], home_hi = sh->run_hi[cell];
    int left = 0;
    uint32_t n_coll = 0, n_flight = 0, n_refl = 0, n_bank = 0, fate = 0;
    float cdf[256];

    for (;;) {
        if (n_flight >= sh->max_flights) { fate = ORACLE_FATE_TRUNCATED; break; }
        const float inv_maj = sh->inv_maj[xsg * G + g];
        float ds = mu * -oracle_logf(oracle_uniform(&rng)) * inv_maj;
        float xn = x + ds;
        ++n_flight;
        while (xn < 0.0f || xn > L) { /* albedo walls, SURVEY 9-Q8 */
            const int lo_wall = xn < 0.0f;
            const float wall = lo_wall ? 0.0f : L, b = lo_wall ? p->boundl : p->boundr;
            if (!(b > 0.0f)) { fate = ORACLE_FATE_LEAKED; break; }
            const float rem = xn - wall;
            mu = mu * (-b);
            xn = wall + rem * (-b);
            if (lo_wall ? (home_lo != 0) : (home_hi != N)) left = 1;
            ++n_refl;
        }
        if (fate) break;
        cell = locate_cell(p, xn);
        x = xn;
        if (cell < home_lo || cell >= home_hi) left = 1;
        const uint32_t mat = p->matid[cell];
        const uint32_t g_eff = left ? g : xsg;
        const uint32_t xs = mat + M * g_eff;
        score(w, g, cell, inv_maj);
        if (oracle_uniform(&rng) < sh->sigtr[xs] * inv_maj) {
            /* real collision: same physics and draw order as mc_code.rs:183-209 */
            ++n_coll;
            oracle_scat_mat_calc(G, mat, g, 1.0f / p->sigs[xs], p->scat, cdf);
            float xi_int = oracle_uniform(&rng);
            float absorption = p->siga[xs] / p->sigt[xs];
            float mu_new = 2.0f * oracle_uniform(&rng) - 1.0f;
            uint32_t g_new = sample_group(cdf, G, o->scatter_mode, &rng);
            if (sh->bank_mode) {
                float nusigf = p->nut[mat + M * g] * p->sigf[mat + M * g];
                if (nusigf > 0.0f) {
                    float wgt = nusigf * p->inv_sigtr[xs] * sh->inv_k;
                    uint32_t n = (uint32_t)(int32_t)(wgt + oracle_uniform(&rng));
                    uint32_t xb;
                    memcpy(&xb, &x, 4);
                    for (uint32_t j = 0; j < n; ++j) {
                        if (n_bank < sh->bank_cap)
                            sh->slots[(y - sh->slot_base) * sh->bank_cap + n_bank] = ((uint64_t)cell << 32) | xb;
                        ++n_bank;
                    }
                }
            }
            if (xi_int < absorption) { fate = ORACLE_FATE_ABSORBED; break; }
            g = g_new;
            mu = mu_new;
            xsg = o->stale_xs ? g_eff : g;
            home_lo = sh->run_lo[cell];
            home_hi = sh->run_hi[cell];
            left = 0;
        }
    }

    w->counters[ORACLE_CT_HISTORIES] += 1;
    w->counters[ORACLE_CT_COLLISIONS] += n_coll;
    w->counters[ORACLE_CT_FLIGHTS] += n_flight;
    w->counters[ORACLE_CT_REFLECTIONS] += n_refl;
    w->counters[ORACLE_CT_LEAKS] += (fate == ORACLE_FATE_LEAKED);
    w->counters[ORACLE_CT_TRUNCATED] += (fate == ORACLE_FATE_TRUNCATED);
    if (sh->bank_mode) {
        uint32_t kept = n_bank < sh->bank_cap ? n_bank : sh->bank_cap;
        sh->counts[y - sh->slot_base] = (uint8_t)kept;
        w->counters[ORACLE_CT_BANKED] += kept;
    }
    if (w->trace) {
        uint32_t *t = w->trace + (y - w->trace_base) * ORACLE_TR_WORDS;
        uint32_t xb;
        memcpy(&xb, &x, 4);
        t[ORACLE_TR_COLLISIONS] = n_coll;
        t[ORACLE_TR_CROSSINGS] = 0;
        t[ORACLE_TR_FLIGHTS] = n_flight;
        t[ORACLE_TR_REFLECTIONS] = n_refl;
        t[ORACLE_TR_RNG_LO] = (uint32_t)rng.state;
        t[ORACLE_TR_RNG_HI] = (uint32_t)(rng.state >> 32);
        t[ORACLE_TR_CELL] = (uint32_t)cell;
        t[ORACLE_TR_XBITS] = xb;
        t[ORACLE_TR_FATE] = fate;
        t[ORACLE_TR_GROUP] = g;
    }
}

/* particle_lifetime, mc_code.rs:215-257 */
ORACLE_HOT static void run_range(worker *w)
{
    if (w->sh->o->tracking_mode == ORACLE_TRACK_WOODCOCK)
        for (uint64_t y = w->start; y < w->end; ++y) run_history_woodcock(w, y);
    else
        for (uint64_t y = w->start; y < w->end; ++y) run_history(w, y);
}

static void *worker_main(void *arg)
{
    run_range((worker *)arg);
    return NULL;
}

/* mc_code.rs:259-274 */
void oracle_average_assembly(const float *flux, uint32_t G, uint32_t N, uint32_t numass, float *out)
{
    uint32_t mesh_assembly = N / numass;
    for (uint32_t g = 0; g < G; ++g) {
        for (uint32_t a = 1; a <= numass; ++a) {
            float s = 0.0f;
            for (uint32_t i = (a - 1) * mesh_assembly; i < a * mesh_assembly; ++i) s += flux[(uint64_t)g * N + i];
            float avg = s / (float)mesh_assembly;
            for (uint32_t i = (a - 1) * mesh_assembly; i < a * mesh_assembly; ++i) out[(uint64_t)g * N + i] = avg;
        }
    }
}

/* mc_code.rs:368-376 */
void oracle_k_fund(const float *k, uint64_t gens, uint64_t skip, float *out)
{
    for (uint64_t i = 0; i < gens; ++i) out[i] = 0.0f;
    if (skip >= gens) return; /* the reference panics here */
    out[skip] = k[skip];
    for (uint64_t n = skip + 1; n < gens; ++n) {
        float s = 0.0f;
        for (uint64_t x = skip; x <= n; ++x) s += k[x];
        out[n] = s / (float)(n - (skip - 1));
    }
}

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* monte_carlo, mc_code.rs:276-380 */
int oracle_monte_carlo(const oracle_problem *p, const oracle_options *o, oracle_results *r)
{
    if (!p || !o || !r) return -1;
    if (p->G < 2 /* mc_code.rs:356 indexes nut[M*1] */ || p->G > 256 || p->M == 0 || p->N == 0 || p->NF == 0 || p->numass == 0) return -2;
    if (o->source_mode < 0 || o->source_mode > ORACLE_SOURCE_FISSION_BANK || o->tracking_mode < 0 || o->tracking_mode > ORACLE_TRACK_WOODCOCK) return -3;
    const uint32_t G = p->G, M = p->M;
    const uint64_t N = p->N, GN = (uint64_t)G * N;

    int T = o->threads;
    if (T <= 0) {
        long hw = sysconf(_SC_NPROCESSORS_ONLN);
        T = (int)(hw > 1 ? hw - 1 : 1); /* mc_code.rs:302 (1-core hosts divide by zero there) */
    }
    uint64_t h0 = o->hist_begin, hc = o->hist_count ? o->hist_count : p->histories;

    run_shared sh;
    sh.p = p;
    sh.o = o;
    oracle_pcg32_seed(&sh.master, o->seed, o->seq);
    sh.max_flights = o->max_flights ? o->max_flights : (1ull << 24);

    worker *ws = (worker *)calloc((size_t)T, sizeof(worker));
    pthread_t *th = (pthread_t *)calloc((size_t)T, sizeof(pthread_t));
    uint64_t *tally_fixed = (uint64_t *)calloc(GN, sizeof(uint64_t));
    float *tally = (float *)calloc(GN, sizeof(float));
    for (int t = 0; t < T; ++t) {
        ws[t].sh = &sh;
        if (o->tally_mode == ORACLE_TALLY_FIXED64)
            ws[t].tally_fixed = (uint64_t *)calloc(GN, sizeof(uint64_t));
        else
            ws[t].tally_f32 = (float *)calloc(GN, sizeof(float));
    }

    for (uint64_t i = 0; i < GN; ++i) { r->flux[i] = 0.0f; r->assembly_average[i] = 0.0f; }
    for (uint64_t i = 0; i < N; ++i) r->fission_source[i] = 0.0f;
    for (uint64_t i = 0; i < p->generations; ++i) { r->k[i] = 0.0f; r->k_fund[i] = 0.0f; }
    memset(r->counters, 0, sizeof(r->counters));
    r->seconds_transport = 0.0;

    /* Woodcock tables */
    float *sigtr = (float *)malloc((size_t)M * G * sizeof(float));
    float *inv_maj = (float *)malloc((size_t)G * G * sizeof(float));
    uint32_t *run_lo = (uint32_t *)malloc(N * sizeof(uint32_t)), *run_hi = (uint32_t *)malloc(N * sizeof(uint32_t));
    {
        int present[256] = {0};
        for (uint64_t i = 0; i < N; ++i) present[p->matid[i]] = 1;
        for (uint32_t i = 0; i < M * G; ++i) sigtr[i] = p->sigt[i] - p->mu[i] * p->sigs[i];
        for (uint32_t a = 0; a < G; ++a)
            for (uint32_t b = 0; b < G; ++b) {
                float mx = 0.0f;
                for (uint32_t m = 0; m < M; ++m) {
                    if (!present[m]) continue;
                    if (sigtr[m + M * a] > mx) mx = sigtr[m + M * a];
                    if (sigtr[m + M * b] > mx) mx = sigtr[m + M * b];
                }
                inv_maj[a * G + b] = 1.0f / mx;
            }
        for (uint64_t i = 0; i < N;) {
            uint64_t j = i;
            while (j < N && p->matid[j] == p->matid[i]) ++j;
            for (uint64_t q = i; q < j; ++q) { run_lo[q] = (uint32_t)i; run_hi[q] = (uint32_t)j; }
            i = j;
        }
    }
    sh.sigtr = sigtr; sh.inv_maj = inv_maj; sh.run_lo = run_lo; sh.run_hi = run_hi;

    sh.bank_mode = (o->source_mode == ORACLE_SOURCE_FISSION_BANK);
    sh.bank_cap = o->bank_cap > 0 ? (uint32_t)o->bank_cap : 8u;
    if (sh.bank_cap > 255) sh.bank_cap = 255;
    sh.src_bank = NULL; sh.src_count = 0; sh.slots = NULL; sh.counts = NULL; sh.slot_base = h0;
    uint64_t *bank_cur = NULL, *bank_next = NULL, *cell_hist = NULL;
    if (sh.bank_mode) {
        sh.slots = (uint64_t *)malloc(hc * sh.bank_cap * sizeof(uint64_t));
        sh.counts = (uint8_t *)calloc(hc, 1);
        bank_next = (uint64_t *)malloc(hc * sh.bank_cap * sizeof(uint64_t));
        bank_cur = (uint64_t *)malloc(hc * sh.bank_cap * sizeof(uint64_t));
        cell_hist = (uint64_t *)calloc(N, sizeof(uint64_t));
    }

    float k_new = p->k0;
    for (uint64_t x = 0; x < p->generations; ++x) {
        float k = k_new;
        k_new = 0.0f;
        sh.gen = x;
        sh.inv_k = 1.0f / k;

        /* static contiguous ranges, mc_code.rs:303-307 */
        uint64_t per = hc / (uint64_t)T;
        double t0 = now_s();
        for (int t = 0; t < T; ++t) {
            worker *w = &ws[t];
            w->start = h0 + (uint64_t)t * per;
            w->end = (t == T - 1) ? h0 + hc : h0 + (uint64_t)(t + 1) * per;
            if (o->inclusive_ranges) w->end += 1; /* Q4 */
            if (w->tally_fixed) memset(w->tally_fixed, 0, GN * sizeof(uint64_t));
            if (w->tally_f32) memset(w->tally_f32, 0, GN * sizeof(float));
            w->trace = (r->trace && r->trace_gen == x && !o->inclusive_ranges) ? r->trace : NULL;
            w->trace_base = h0;
            if (T == 1) worker_main(w);
            else pthread_create(&th[t], NULL, worker_main, w);
        }
        /* join + reduce in worker order, mc_code.rs:331-338 */
        for (uint64_t i = 0; i < GN; ++i) { tally[i] = 0.0f; tally_fixed[i] = 0; }
        for (int t = 0; t < T; ++t) {
            if (T != 1) pthread_join(th[t], NULL);
            worker *w = &ws[t];
            if (w->tally_fixed)
                for (uint64_t i = 0; i < GN; ++i) tally_fixed[i] += w->tally_fixed[i];
            else
                for (uint64_t i = 0; i < GN; ++i) tally[i] += w->tally_f32[i];
        }
        r->seconds_transport += now_s() - t0;
        if (sh.bank_mode) {
            /* compaction in canonical (history, site) order, then the bank feeds the next generation */
            uint64_t n_sites = 0;
            memset(cell_hist, 0, N * sizeof(uint64_t));
            for (uint64_t y = 0; y < hc; ++y)
                for (uint32_t j = 0; j < sh.counts[y]; ++j) {
                    uint64_t site = sh.slots[y * sh.bank_cap + j];
                    bank_next[n_sites++] = site;
                    cell_hist[site >> 32]++;
                }
            if (r->bank_sizes) r->bank_sizes[x] = n_sites;
            if (r->entropy) {
                double e = 0.0;
                for (uint64_t i = 0; i < N; ++i)
                    if (cell_hist[i]) { double pr = (double)cell_hist[i] / (double)n_sites; e -= pr * log2(pr); }
                r->entropy[x] = e;
            }
            if (r->bank_sites && r->bank_gen == x)
                memcpy(r->bank_sites, bank_next, (n_sites < r->bank_sites_cap ? n_sites : r->bank_sites_cap) * sizeof(uint64_t));
            uint64_t *tmp = bank_cur; bank_cur = bank_next; bank_next = tmp;
            sh.src_bank = n_sites ? bank_cur : NULL; /* an empty bank falls back to the uniform source */
            sh.src_count = n_sites;
        }
        if (o->tally_mode == ORACLE_TALLY_FIXED64) {
            for (uint64_t i = 0; i < GN; ++i) tally[i] = (float)((double)tally_fixed[i] * (1.0 / 268435456.0));
            if (r->tally_fixed) memcpy(r->tally_fixed + x * GN, tally_fixed, GN * sizeof(uint64_t));
        }

        /* mc_code.rs:340-363 ; usize arithmetic wraps in the release profile */
        float fund = 1.0f / (float)(uint64_t)(p->generations - (p->skip - 1));
        for (uint32_t g = 0; g < G; ++g) {
            for (uint64_t i = 0; i < N; ++i) {
                float delta_x = p->dx[i];
                uint32_t matid = p->matid[i];
                float flux = tally[(uint64_t)g * N + i] / (k * (float)p->histories * delta_x);
                float fission_source = p->nut[matid + M * g] * p->sigf[matid + M * g] * flux;
                k_new += k * delta_x * fission_source;
                if (x >= p->skip) {
                    float conversion = (3565e6f * k * 36.2f) /
                                       (200e6f * 1.602176634e-19f * p->nut[0 + M * 1] * p->right[N - 1]);
                    r->flux[(uint64_t)g * N + i] += flux * conversion * fund;
                    r->fission_source[i] += fission_source * fund;
                }
            }
        }
        r->k[x] = k_new;
    }

    oracle_average_assembly(r->flux, G, (uint32_t)N, p->numass, r->assembly_average);
    oracle_k_fund(r->k, p->generations, p->skip, r->k_fund);

    for (int t = 0; t < T; ++t) {
        for (int c = 0; c < ORACLE_CT_WORDS; ++c) r->counters[c] += ws[t].counters[c];
        free(ws[t].tally_fixed);
        free(ws[t].tally_f32);
    }
    free(ws); free(th); free(tally_fixed); free(tally);
    free(sh.slots); free(sh.counts); free(bank_cur); free(bank_next); free(cell_hist);
    free(sigtr); free(inv_maj); free(run_lo); free(run_hi);
    return 0;
}
