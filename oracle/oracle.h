/*
 * oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement of the NRAPS Monte Carlo hot path (reference
 * src/mc_code.rs:7-380).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library; the product
 * (nraps_b200/) never does.
 *
 * PARITY STATUS: the reference cannot be built here (no Rust toolchain) and is
 * unseedable at HEAD, so end-to-end k / flux are "parity unpinned"; the helper
 * functions are pinned by the reference's own unit tests
 * (src/mc_code.rs:392-556) and PCG32 by the upstream demo vector.
 */
#ifndef NRAPS_ORACLE_H
#define NRAPS_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Borrowed views of the reference's Variables / XSData / Mesh / fuel_indices
 * (src/main.rs:22-75).  Tables are indexed mat + M*g, scat G*G*mat + G*g + g'. */
typedef struct {
    uint32_t M, G, N, NF, numass;
    uint64_t generations, histories, skip;
    float boundl, boundr, dx_fuel, dx_water, k0;
    const float *sigt, *sigs, *mu, *siga, *sigf, *nut, *chit, *inv_sigtr, *scat;
    const uint8_t *matid;
    const float *dx, *left, *right;
    const uint64_t *fuel_indices;
} oracle_problem;

enum { ORACLE_SCATTER_SINGLE_XI = 0, ORACLE_SCATTER_RUST_PRE182 = 1, ORACLE_SCATTER_RUST_182 = 2 };
enum { ORACLE_TALLY_FIXED64 = 0, ORACLE_TALLY_F32_PER_WORKER = 1 };
enum { ORACLE_SOURCE_UNIFORM_FUEL = 0, ORACLE_SOURCE_FISSION_BANK = 1 };
enum { ORACLE_TRACK_SURFACE = 0, ORACLE_TRACK_WOODCOCK = 1 };

typedef struct {
    uint64_t seed, seq, stride; /* PCG32 master stream and per-history jump  */
    int32_t scatter_mode;       /* Q3                                         */
    int32_t stale_xs;           /* Q1: 1 = faithful                           */
    int32_t tally_mode;         /* Q15                                        */
    int32_t inclusive_ranges;   /* Q4: 1 = reference's start..=end            */
    int32_t threads;            /* 0 => hardware_concurrency-1 (mc_code.rs:302) */
    int32_t source_mode;
    int32_t tracking_mode;
    int32_t bank_cap;           /* fission_bank: max sites kept per history (0 => 8) */
    uint64_t hist_begin, hist_count; /* sub-range of each generation; count 0 => all */
    uint64_t max_flights;            /* per-history safety cap, 0 => 1<<24    */
} oracle_options;

/* One record per history of generation `trace_gen` (10 x u32). */
enum {
    ORACLE_TR_COLLISIONS = 0, ORACLE_TR_CROSSINGS, ORACLE_TR_FLIGHTS, ORACLE_TR_REFLECTIONS,
    ORACLE_TR_RNG_LO, ORACLE_TR_RNG_HI, ORACLE_TR_CELL, ORACLE_TR_XBITS, ORACLE_TR_FATE, ORACLE_TR_GROUP,
    ORACLE_TR_WORDS
};
enum { ORACLE_FATE_ABSORBED = 1, ORACLE_FATE_LEAKED = 2, ORACLE_FATE_TRUNCATED = 3 };

enum {
    ORACLE_CT_HISTORIES = 0, ORACLE_CT_COLLISIONS, ORACLE_CT_CROSSINGS, ORACLE_CT_FLIGHTS,
    ORACLE_CT_REFLECTIONS, ORACLE_CT_LEAKS, ORACLE_CT_TRUNCATED, ORACLE_CT_BANKED,
    ORACLE_CT_WORDS
};

#define ORACLE_TALLY_FRAC_BITS 28

typedef struct {
    float *flux, *assembly_average; /* [G][N] */
    float *fission_source;          /* [N]    */
    float *k, *k_fund;              /* [gens] */
    uint64_t *tally_fixed;          /* optional [gens][G][N], 2^-28 units  */
    uint32_t *trace;                /* optional [hist_count][ORACLE_TR_WORDS] */
    uint64_t trace_gen;
    uint64_t counters[ORACLE_CT_WORDS]; /* summed over all generations      */
    uint64_t *bank_sizes;           /* optional [gens], fission_bank mode: sites banked by each generation */
    uint64_t *bank_sites;           /* optional: dense bank produced by generation `bank_gen`, (cell << 32 | x bits) */
    uint64_t bank_sites_cap, bank_gen;
    double *entropy;                /* optional [gens]: Shannon entropy (bits) of the banked sites over cells */
    double seconds_transport;       /* wall time inside the history loops   */
} oracle_results;

int oracle_monte_carlo(const oracle_problem *p, const oracle_options *o, oracle_results *r);

/* scalar helpers, exported for the golden-vector tests */
void oracle_hit_boundary(float mu, float start_x, float delta_s, float bound, float mesh_end, float out[3]);
void oracle_cross_mesh(uint64_t mesh_index, float mu, float start_x, float mesh_end, float delta_s,
                       float out_ds_x[2], uint64_t *out_index);
float oracle_direction_f(float xi);
void oracle_scat_mat_calc(uint32_t G, uint32_t matid, uint32_t g, float inv_sigs, const float *scat, float *out);
uint32_t oracle_energy_search(const float *cumulative, uint32_t n, float chi);
uint32_t oracle_energy(const oracle_problem *p, float chi, uint64_t cell);
void oracle_pcg32_demo(uint64_t seed, uint64_t seq, uint32_t n, uint32_t *out);
void oracle_pcg32_state(uint64_t seed, uint64_t seq, uint64_t delta, uint64_t out_state_inc[2]);
float oracle_logf_f(float x);
float oracle_unit_f(uint32_t u);
double oracle_logf_max_ulp(uint32_t first, uint32_t count); /* exhaustive check helper */
void oracle_average_assembly(const float *flux, uint32_t G, uint32_t N, uint32_t numass, float *out);
void oracle_k_fund(const float *k, uint64_t gens, uint64_t skip, float *out);

#ifdef __cplusplus
}
#endif
#endif
