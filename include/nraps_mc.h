/*
 * nraps_mc.h -- C ABI of the B200 Monte Carlo k-eigenvalue transport path.
 *
 * Drop-in boundary for the reference's
 *
 *     pub fn monte_carlo(variables: &Variables, xsdata: &XSData, delta_x: &DeltaX,
 *                        meshid: &Vec<Mesh>, fuel_indices: &Vec<usize>, k_new: f32)
 *                        -> SolutionResults                (src/mc_code.rs:276-283)
 *
 * The reference has no FFI of its own; this header is what a Rust `extern "C"`
 * block (rust/) binds.  Plain pointers and sizes only, no
 * allocation crosses the boundary, every entry point returns an int status
 * (the reference panics instead: src/mc_code.rs:302,332).
 *
 * All entry points require a CUDA device (sm_100a).  There is NO CPU fallback:
 * without a usable device they return NRAPS_ERR_CUDA.
 */
#ifndef NRAPS_MC_H
#define NRAPS_MC_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NRAPS_ABI_VERSION 4
#define NRAPS_TALLY_FRAC_BITS 28 /* tallies are exact integers in 2^-28 cm */

enum {
    NRAPS_OK = 0,
    NRAPS_ERR_NULL = 1,        /* a required pointer is NULL                   */
    NRAPS_ERR_SHAPE = 2,       /* M/G/N/NF/numass/generations/skip out of range */
    NRAPS_ERR_MESH = 3,        /* right[i] != left[i+1], matid >= M, fuel index >= N */
    NRAPS_ERR_XS = 4,          /* inv_sigtr not finite-positive for a used material */
    NRAPS_ERR_TOO_LARGE = 5,   /* tables do not fit the 227 KB shared-memory budget */
    NRAPS_ERR_CUDA = 6,        /* CUDA runtime error (nraps_last_cuda_error())  */
    NRAPS_ERR_OPTION = 7,      /* unknown mode in nraps_options                 */
    NRAPS_ERR_STATE = 8,       /* call order violated                           */
    NRAPS_ERR_IO = 9           /* host-side file error (nraps_host.h)           */
};

/*
 * Flat views of the reference's in-memory contract.
 *   Variables ......... src/main.rs:22-39      XSData .. src/main.rs:46-56
 *   Mesh (AoS there) .. src/main.rs:68-75      DeltaX .. src/main.rs:41-44
 * XS tables are indexed [mat + M*g]; scat is [G*G*mat + G*g_from + g_to]
 * (src/mc_code.rs:89-98).  Inputs are borrowed for the duration of the call.
 */
typedef struct nraps_problem {
    uint32_t M, G, N, NF, numass;
    uint64_t generations, histories, skip;
    float boundl, boundr, dx_fuel, dx_water, k0;
    const float *sigt, *sigs, *mu, *siga, *sigf, *nut, *chit, *inv_sigtr; /* [M*G]   */
    const float *scat;                                                     /* [M*G*G] */
    const uint8_t *matid;                                                  /* [N]     */
    const float *dx, *left, *right;                                        /* [N]     */
    const uint64_t *fuel_indices;                                          /* [NF]    */
} nraps_problem;

enum { NRAPS_SCATTER_SINGLE_XI = 0, NRAPS_SCATTER_RUST_PRE182 = 1, NRAPS_SCATTER_RUST_182 = 2 };
enum { NRAPS_SOURCE_UNIFORM_FUEL = 0, NRAPS_SOURCE_FISSION_BANK = 1 };
enum { NRAPS_TRACK_SURFACE = 0, NRAPS_TRACK_WOODCOCK = 1 };
/* FUSED: one persistent lane per neutron (default).  EVENT: the structure-of-arrays bank pipeline in HBM (Woodcock only).
 * BLOCK_EVENT: surface tracking with the neutrons of a block banked in shared memory and sorted by their next event
 * every round.  Measured on a B200 (bit-exact, 0.72x the fused kernel) and therefore NOT compiled into the product
 * library: the value is rejected with NRAPS_ERR_OPTION unless the library was built with `make BLOCK_EVENT=1`. */
enum { NRAPS_KERNEL_FUSED = 0, NRAPS_KERNEL_EVENT = 1, NRAPS_KERNEL_BLOCK_EVENT = 2 };

typedef struct nraps_options {
    uint64_t seed, stream, stride; /* PCG32 master (seed, sequence) and per-history jump; 0,0,0 => 42,54,152917 */
    int32_t device;                /* CUDA ordinal                                    */
    int32_t scatter_mode;          /* SURVEY 9-Q3; default single_xi                  */
    int32_t stale_xs;              /* SURVEY 9-Q1; 1 = faithful to src/mc_code.rs:147 */
    int32_t source_mode;
    int32_t tracking_mode;
    int32_t kernel_variant;
    int32_t threads_per_block;     /* 0 = auto */
    int32_t blocks_per_sm;         /* 0 = auto */
    int32_t chunk;                 /* histories a warp claims per global atomic; 0 = auto */
    int32_t quiet;                 /* 0 = print "running MC code" (src/mc_code.rs:292) */
    int32_t bank_cap;              /* fission_bank: sites kept per history, 1..255; 0 = 8 */
    int32_t spawn_batch;           /* refill a warp's dead lanes only once this many are dead (0 = auto); block_event variant:
                                    * walks predicted to cross >= this many cells form the "long" list (0 = split by run length) */
    int32_t walk_cap;              /* tuning knob of the surface kernel's walk.  Round 1: crossings before a warp regroups (gone:
                                    * with range-update tallies a lane walks to the end of its segment).  Round 2, fine meshes:
                                    * > 0 = closed-form strides only while |ds| > the first power of two >= walk_cap cell
                                    * widths (0 = auto, 6);
                                    * -2 = never stride (cell-by-cell loop only).  No value changes a result bit. */
    int32_t slots_per_thread;      /* block_event variant: neutrons banked per thread of a block; 0 = 3 (was reserved1) */
    uint64_t max_flights;          /* per-history safety cap; 0 = 1<<24               */
    int32_t profile_phases;        /* 1 = time the phases of every generation with CUDA events (nraps_mc_phase_ms); the host
                                    * then waits for generation g's kernels before it launches g+1 */
    int32_t reserved0;
} nraps_options;

/* phases timed by profile_phases (milliseconds, summed over the generations since creation / nraps_mc_reset) */
enum { NRAPS_PH_SOURCE = 0, NRAPS_PH_TRANSPORT, NRAPS_PH_PREFIX, NRAPS_PH_COMPACT, NRAPS_PH_FINALIZE, NRAPS_PH_WORDS };

enum {
    NRAPS_CT_HISTORIES = 0, NRAPS_CT_COLLISIONS, NRAPS_CT_CROSSINGS, NRAPS_CT_FLIGHTS,
    NRAPS_CT_REFLECTIONS, NRAPS_CT_LEAKS, NRAPS_CT_TRUNCATED, NRAPS_CT_BANKED,
    NRAPS_CT_WORDS
};

/* per-history replay record written by nraps_mc_trace (10 x u32) */
enum {
    NRAPS_TR_COLLISIONS = 0, NRAPS_TR_CROSSINGS, NRAPS_TR_FLIGHTS, NRAPS_TR_REFLECTIONS,
    NRAPS_TR_RNG_LO, NRAPS_TR_RNG_HI, NRAPS_TR_CELL, NRAPS_TR_XBITS, NRAPS_TR_FATE, NRAPS_TR_GROUP,
    NRAPS_TR_WORDS
};
enum { NRAPS_FATE_ABSORBED = 1, NRAPS_FATE_LEAKED = 2, NRAPS_FATE_TRUNCATED = 3 };

/* SolutionResults, src/main.rs:77-83; row-major [g][cell]; caller-allocated. */
typedef struct nraps_results {
    float *flux, *assembly_average; /* [G*N]  */
    float *fission_source;          /* [N]    */
    float *k, *k_fund;              /* [generations] */
    uint64_t *tally_fixed;          /* optional [generations][G][N], 2^-28 units (forces a sync per generation) */
    uint64_t counters[NRAPS_CT_WORDS]; /* summed over generations; crossings/flights/reflections only in trace runs */
    double seconds_device;          /* CUDA-event time of the generation loop */
    uint64_t *bank_sizes;           /* optional [generations]: sites banked by each generation (fission_bank mode) */
    double *entropy;                /* optional [generations]: Shannon entropy (bits) of that bank over mesh cells */
    double *flux_moments;           /* optional [2][G][N] extension: over the generations >= skip, the sum and the sum of
                                     * squares of the per-generation term flux * conversion that src/mc_code.rs:358
                                     * accumulates (flux[] = fund * sum); gives the per-bin variance between generations */
} nraps_results;

typedef struct nraps_mc_ctx nraps_mc_ctx;

/* The parity defaults, i.e. the switch set every front end of this repository uses: everything
 * zero except stale_xs = 1 (faithful to src/mc_code.rs:147) and quiet = 0.  A zero-initialised nraps_options differs
 * from it in stale_xs -- C callers should start from this initialiser rather than from memset. */
void nraps_options_default(nraps_options *o);

/* Whole job on one GPU: the monte_carlo() replacement (src/mc_code.rs:276-380).  With the uniform source, generations
 * are independent, so one launch carries several small generations (up to ~2^23 histories, each generation scoring
 * into its own tally rows) and they are folded in order afterwards: results are bit-identical to one launch per
 * generation, which is what the generation-level API below always does.  Generations large enough for a launch of their
 * own are pipelined instead: consecutive launches alternate between two streams (see nraps_mc_select_lane), the
 * finalizes stay in order -- same bits again.  Environment: NRAPS_PIPELINE=0 keeps every launch on one stream,
 * NRAPS_TAIL_BATCH=n lets a launch carry n generations up to 2^25 histories, NRAPS_TIMING=1 prints the host-side
 * split of the call on stderr. */
int nraps_mc_run(const nraps_problem *p, const nraps_options *o, nraps_results *r);

/*
 * Generation-level API (one context per GPU / process).  A multi-GPU host
 * shards [0, histories) across ranks, calls transport on its shard, sums the
 * integer tally buffer across ranks (NCCL all-reduce on int64), then finalizes;
 * every rank then holds identical k / flux.  `stream` is a cudaStream_t (NULL
 * = default stream); calls are asynchronous unless stated.
 */
int nraps_mc_create(const nraps_problem *p, const nraps_options *o, nraps_mc_ctx **out);
int nraps_mc_destroy(nraps_mc_ctx *ctx);
/* Device buffers come from a library-owned memory pool per device that keeps up to 2 GiB mapped after
 * nraps_mc_destroy / nraps_mc_run, so that the next context does not pay the driver's allocation cost again
 * (50-300 ms per context, measured).  nraps_mc_trim hands that memory back to the driver. */
int nraps_mc_trim(int32_t device);
int nraps_mc_reset(nraps_mc_ctx *ctx, float k0, void *stream);
int nraps_mc_transport(nraps_mc_ctx *ctx, uint64_t gen, uint64_t hist_begin, uint64_t hist_count, void *stream);
int nraps_mc_finalize_generation(nraps_mc_ctx *ctx, uint64_t gen, void *stream);
/* device buffer of G*N tally words followed by NRAPS_CT_WORDS counter words (uint64); in fission_bank mode N more words
 * follow (cell histogram of the generation's bank).  n_words is what a multi-GPU host all-reduces. */
int nraps_mc_tally_buffer(nraps_mc_ctx *ctx, void **device_ptr, uint64_t *n_words);
int nraps_mc_set_tally_buffer(nraps_mc_ctx *ctx, void *device_ptr); /* caller-owned, same size */
/* A transport launch owns a set of scratch buffers (birth records, chunk cursor, difference array) until its kernels are
 * done; a context has two such sets.  nraps_mc_select_lane(ctx, 0 | 1) chooses the set the following nraps_mc_transport
 * calls use (default 0), so that -- uniform source only: generations independent -- generation g+1 can be launched on a
 * second stream, into a second tally buffer, while the tail of generation g still runs.  The caller orders the reuse of
 * a lane and of a tally buffer (same stream, or events); finalize calls stay in generation order. */
int nraps_mc_select_lane(nraps_mc_ctx *ctx, int32_t lane);
int nraps_mc_read_tally(nraps_mc_ctx *ctx, uint64_t *host_words, void *stream);      /* synchronous */
int nraps_mc_fetch(nraps_mc_ctx *ctx, nraps_results *r, void *stream);               /* synchronous */
/* replay: run [hist_begin, hist_begin+hist_count) of `gen` and return one record per history (synchronous) */
int nraps_mc_trace(nraps_mc_ctx *ctx, uint64_t gen, uint64_t hist_begin, uint64_t hist_count,
                   uint32_t *host_records, void *stream);
/*
 * fission_bank source mode (power iteration; no reference counterpart).  Per generation g, on every rank:
 *   nraps_mc_transport(g)            births from the bank of g-1 (uniform source for g = 0), then transport
 *   nraps_mc_bank_compact(g)         this rank's sites of g in canonical (history, site) order, (cell << 32 | x bits)
 *                                    each, into one of its two bank buffers; the bank's cell histogram is added to
 *                                    the N words behind the counters of the tally buffer
 *   [all-reduce of the tally buffer] multi-GPU only; it also orders every rank's compaction before any rank's next births
 *   nraps_mc_finalize_generation(g)
 *   nraps_mc_bank_advance(g)         records size and entropy of the bank of g; generation g+1 samples from it
 * (finalize may also come before compact: it does not touch the bank.)
 *
 * Several GPUs: no gathered copy of the bank exists.  Each rank keeps its bank where it compacted it and the source
 * kernel of every rank resolves a site index to (rank, offset) from the ranks' site counts -- word 0 of each bank buffer
 * -- and loads the 8-byte site from the rank that banked it, over NVLink.  Set up once, before generation 0:
 *   nraps_mc_bank_reserve   allocate the two buffers mappable by the other ranks, for shards up to `shard_histories`
 *   one process per GPU:    reserve(.., NULL) -> nraps_mc_bank_export (2 tickets) -> exchange them -> nraps_mc_bank_import
 *                           (cuMemCreate allocations shared as file descriptors, 2 MB pages on both sides; the legacy
 *                           cudaIpc mappings of round 2's first version thrashed the TLB under random reads)
 *   one process, N devices: reserve(.., buffers) -> enable peer access -> every rank's two pointers to nraps_mc_bank_peers
 */
#define NRAPS_IPC_HANDLE_BYTES 64 /* one exported bank buffer: {magic, pid, fd, size}, padded */
#define NRAPS_MAX_RANKS 8         /* GPUs of one NVSwitch box */
int nraps_mc_bank_compact(nraps_mc_ctx *ctx, uint64_t gen, void *stream);
int nraps_mc_bank_advance(nraps_mc_ctx *ctx, uint64_t gen, void *stream);
/* device pointer + count of the dense local bank of the last compaction (synchronous; tests) */
int nraps_mc_bank_local(nraps_mc_ctx *ctx, void **device_sites, uint64_t *count, void *stream);
int nraps_mc_bank_reserve(nraps_mc_ctx *ctx, uint64_t shard_histories, void **device_buffers /* [2] out; NULL = for other processes */);
int nraps_mc_bank_export(nraps_mc_ctx *ctx, void *handles /* [2][NRAPS_IPC_HANDLE_BYTES] out; valid while the context lives */);
int nraps_mc_bank_import(nraps_mc_ctx *ctx, int32_t world, int32_t rank, const void *handles /* [world][2][NRAPS_IPC_HANDLE_BYTES] */);
int nraps_mc_bank_peers(nraps_mc_ctx *ctx, int32_t world, int32_t rank, const void *const *device_buffers /* [world][2] */);
/* profile_phases = 1: milliseconds spent in {births, transport kernel, tally prefix sum, bank compaction + histogram,
 * finalize} since creation / reset (synchronizes the device) */
int nraps_mc_phase_ms(nraps_mc_ctx *ctx, double out[NRAPS_PH_WORDS]);
/* launch geometry chosen for this context: {grid, block, dynamic smem bytes, blocks/SM, SM count, chunk} */
int nraps_mc_launch_info(nraps_mc_ctx *ctx, uint32_t out[6]);

/* device-side unit probes for the golden-vector tests (each runs a 1-block kernel) */
int nraps_dev_logf(const float *x, float *out, uint32_t n, int32_t device);
/* walk-loop division (hoisted reciprocal) next to the device's IEEE division, for the exactness test */
int nraps_dev_div(const float *t, const float *mu, float *out_fast, float *out_ieee, uint32_t n, int32_t device);
int nraps_dev_pcg32(uint64_t seed, uint64_t stream, uint64_t stride, uint64_t hid, uint32_t n,
                    uint32_t *out_u32, float *out_unit, int32_t device);

const char *nraps_strerror(int code);
const char *nraps_last_cuda_error(void);
int nraps_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif
