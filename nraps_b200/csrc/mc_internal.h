// Internal contract between the C-ABI layer (mc_api.cu) and the kernels.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/nraps_mc.h"

namespace nraps {

constexpr float kTallyScale = 268435456.0f; // 2^NRAPS_TALLY_FRAC_BITS
constexpr double kTallyInvScale = 1.0 / 268435456.0;

// Byte offsets of the per-block shared-memory image.  Everything a history
// touches between birth and death lives here; HBM is only read at block start
// (tables, ~N*11 bytes) and written at block end (tally flush).
struct SmemLayout {
    uint32_t tally_lo, tally_hi; // u32[G*N] each: 64-bit fixed-point bins split in two words
    uint32_t edges;              // f32[N+1]   cell edges; left[i]=edges[i], right[i]=edges[i+1]
    uint32_t runb;               // u32[N]     material-run bounds of each cell: lo | hi<<16
    uint32_t jump;               // u64x2[64]  PCG32 jump table (A_b, C_b) for stride*2^b draws
    uint32_t xs;                 // f32[...]   inv_sigtr[MG] | p_abs[MG] | chi_cdf[MG] | nusigf[MG] | scat_cdf[M*G*G*G]
    uint32_t fuel;               // u16[NF]    fuel cell indices
    uint32_t matid;              // u8[N]
    uint32_t bucket;             // u16[NB]    Woodcock: first-guess cell of each position bucket
    uint32_t total;
};

__host__ __device__ inline uint32_t align_up(uint32_t v, uint32_t a) { return (v + a - 1) / a * a; }

// floats in the XS block: inv_sigtr | p_abs | chi_cdf | nusigf | sigtr [MG each] | scat_cdf [M*G^3] | inv_maj [G*G]
__host__ __device__ inline uint32_t xs_floats(uint32_t M, uint32_t G) { return 5 * M * G + M * G * G * G + G * G; }

// big = 1: the mesh does not fit one SM's shared memory; only the jump table and the XS block are staged, the mesh
// tables are read through L1/L2 and the tally goes straight to the global 64-bit bins (slower, any N <= 65535)
// rows = tally rows: G, or batch*G when several small generations share one launch (0 = G)
__host__ __device__ inline SmemLayout make_layout(uint32_t M, uint32_t G, uint32_t N, uint32_t NF, uint32_t NB, uint32_t big = 0,
                                                  uint32_t rows = 0)
{
    if (big) { N = 0; NF = 0; NB = 0; }
    if (!rows) rows = G;
    SmemLayout L;
    uint32_t off = 0;
    L.jump = off;     off += 64 * 16;
    L.tally_lo = off; off += rows * N * 4;
    L.tally_hi = off; off += rows * N * 4;
    L.edges = off;    off += (N + 1) * 4;
    L.runb = off;     off += N * 4;
    off = align_up(off, 16); // float4 rows of the CDF tables
    L.xs = off;       off += xs_floats(M, G) * 4;
    L.fuel = off;     off += align_up(NF * 2, 4);
    L.matid = off;    off += align_up(N, 4);
    L.bucket = off;   off += align_up(NB * 2, 4);
    L.total = align_up(off, 16);
    return L;
}

// Shared-memory image of the surface-tracking kernel (mc_transport.cu).  A flight scores the same value into every
// cell it crosses completely inside one *segment* (cells of one material run whose widths are the same binary32
// number), so those scores are kept as a difference array `diff` (+v at the first cell, -v one past the last) and
// only the partial cells at the two ends of a flight are scored one by one (`direct`); tally = direct + prefix(diff).
// All bins are 64-bit integers in 2^-28 cm split in two u32 words, sums are exact mod 2^64, so the prefix sum at the
// end reproduces the cell-by-cell tally of src/mc_code.rs:163,173,184 bit for bit.
//   SURF_SPLIT   : direct[rows][N] and diff[rows][N+1] in shared memory (meshes that leave room for two blocks per SM)
//   SURF_UNIFIED : diff only; a direct score is the point update +s at the cell, -s at the next (half the memory)
//   SURF_GLOBAL  : mesh tables through L1/L2, both arrays in global memory (mesh too large for one SM)
enum { SURF_SPLIT = 0, SURF_UNIFIED = 1, SURF_GLOBAL = 2 };
struct SurfLayout {
    uint32_t diff_lo, diff_hi;     // u32[rows*(N+1)] each
    uint32_t direct_lo, direct_hi; // u32[rows*N] each (SURF_SPLIT only)
    uint32_t edges;                // f32[N+1]
    uint32_t segw;                 // u32x2[N]: segment bounds of the cell (lo | hi<<16), its width bits
    uint32_t xs;                   // f32[xs_floats]
    uint32_t matid;                // u8[N]
    uint32_t total;
};
__host__ __device__ inline SurfLayout make_surface_layout(uint32_t M, uint32_t G, uint32_t N, uint32_t mode, uint32_t rows = 0)
{
    if (!rows) rows = G;
    if (mode == SURF_GLOBAL) N = 0;
    SurfLayout L;
    uint32_t off = 0;
    L.diff_lo = off;   off += (mode == SURF_GLOBAL ? 0u : rows * (N + 1) * 4u);
    L.diff_hi = off;   off += (mode == SURF_GLOBAL ? 0u : rows * (N + 1) * 4u);
    L.direct_lo = off; off += (mode == SURF_SPLIT ? rows * N * 4u : 0u);
    L.direct_hi = off; off += (mode == SURF_SPLIT ? rows * N * 4u : 0u);
    L.edges = off;     off += (mode == SURF_GLOBAL ? 0u : (N + 1) * 4u);
    off = align_up(off, 8);
    L.segw = off;      off += N * 8u;
    off = align_up(off, 16); // float4 rows of the CDF tables
    L.xs = off;        off += xs_floats(M, G) * 4u;
    L.matid = off;     off += align_up(N, 4);
    L.total = align_up(off, 16);
    return L;
}

constexpr int kMaxPeers = 8;        // GPUs of one NVSwitch box
constexpr uint32_t kBankHeader = 16; // 64-bit words in front of the sites of a bank buffer (word 0: site count)

struct TransportParams {
    // read-only tables in global memory (device pointers)
    const float *edges;
    const uint32_t *runb;
    const uint8_t *matid;
    const uint16_t *fuel;
    const float *xs;
    const ulonglong2 *jump;
    const uint16_t *bucket;  // [NB] Woodcock position buckets (NB = 0 in surface mode)
    const uint4 *source;     // [hist_end-hist_begin][2] born neutrons written by source_kernel
    const uint2 *segw;       // [N] surface kernel: {segment lo | hi << 16, width bits} of each cell
    unsigned long long *diff; // [rows*N] surface kernel: difference array of the full-cell scores (global, summed over blocks)
    uint32_t surf_mode;      // SURF_SPLIT / SURF_UNIFIED / SURF_GLOBAL
    uint32_t skip_walk;      // surface kernel: stride over the surely-crossed cells of a segment in closed form (fine meshes)
    float length;            // right edge of the slab (bounds the rounding of x + ds)
    float stride_min;        // closed-form strides only while |ds| > the first power of two >= stride_min * w (then cell by cell)
    uint32_t diff_hi_off, direct_hi_off; // surface kernel: byte distance from the low to the high words of a shared-memory bin array
                                         // (host-computed from SurfLayout: the kernel reads them in the rare carry path only)
    uint32_t M, G, N, NF, NB, big;
    uint32_t rows;       // tally rows of this launch = batch * G: generations gen .. gen+batch-1 share the launch
    uint64_t hist_shard; // histories of one generation in this launch (hist_end - hist_begin = batch * hist_shard)
    uint64_t hist_total; // histories per generation (stream position of generation j starts at j * hist_total)
    float boundl, boundr, dx_fuel, inv_h;
    uint64_t rng_state; // master stream advanced to history 0 of this generation
    uint64_t rng_inc;
    uint64_t hist_begin, hist_end; // this launch covers y in [hist_begin, hist_end)
    unsigned long long *work;      // chunk cursor, zeroed before launch
    unsigned long long *tally;     // [G*N] + counters[NRAPS_CT_WORDS]
    uint32_t *trace;               // [hist_end-hist_begin][NRAPS_TR_WORDS] or nullptr
    uint32_t chunk;
    uint32_t max_flights;
    uint32_t walk_cap;    // surface kernel: crossings a lane walks before the warp regroups (0xffffffff = never)
    uint32_t spawn_batch; // dead lanes a warp waits for before it refills (amortises the divergent spawn path)
    int32_t scatter_mode, stale_xs;
    // fission_bank source mode: the bank of the previous generation, one dense buffer per rank (this rank's own and,
    // over NVLink, the peers'); word 0 of a buffer is its site count, the sites (cell << 32 | x bits) start at word
    // kBankHeader.  Canonical order = rank order, history order inside a rank.  n_peers = 0: no bank yet (uniform source)
    const unsigned long long *peer_bank[kMaxPeers];
    uint32_t n_peers;
    unsigned long long *slots;               // [hist_end-hist_begin][bank_cap] sites produced, by history
    uint8_t *counts;                         // [hist_end-hist_begin]
    const float *k_cur;                      // device scalar: k of the previous generation
    uint32_t bank_cap;
};

// SoA particle bank of the event-based pipeline, two halves for ping-pong compaction (28 B per record)
struct EventHalf {
    float *x, *mu;
    uint32_t *pack, *cnt, *ccnt;
    unsigned long long *rng;
};
struct EventBank {
    EventHalf half[2];
    unsigned long long *n_alive, *n_next; // device scalars
};

struct BankParams {
    const uint8_t *counts;            // [n_hist padded to kBankTile]
    const unsigned long long *slots;  // [n_hist][cap]
    unsigned long long *dense;        // [dense_cap]
    unsigned long long *block_sums;   // [n_tiles]
    unsigned long long *count_out;    // device scalar
    uint64_t n_hist, dense_cap;
    uint32_t cap, n_tiles;
};
constexpr uint32_t kBankTile = 16384; // histories per block of the compaction kernels (1024 threads x 16)

struct FinalizeParams {
    const unsigned long long *tally; // [G*N] of this generation
    const unsigned long long *counters; // [NRAPS_CT_WORDS] of the launch, or nullptr if already accounted
    double *res_moments;                // [2][G*N] running sum and sum of squares of flux * conversion over the accumulated generations
    const float *dx;                 // [N]
    const uint8_t *matid;            // [N]
    const float *nusigf_nut;         // nut[MG]
    const float *sigf;               // sigf[MG]
    float *terms;                    // scratch [G*N]
    float *res_flux;                 // [G*N]
    float *res_fission;              // [N]
    float *k_hist;                   // [generations]
    float *k_cur;                    // [1]
    unsigned long long *counters_total; // [NRAPS_CT_WORDS], summed over generations
    uint32_t M, G, N;
    float histories_f32, length, nut_m1, fund;
    uint64_t gen, skip;
};

cudaError_t launch_transport(const TransportParams &p, bool trace, bool bank, dim3 grid, dim3 block, uint32_t smem, cudaStream_t s);
int occupancy_transport(uint32_t G, uint32_t surf_mode, bool trace, bool bank, int block, uint32_t smem);
// tally[r][i] += sum_{j <= i} diff[r][j] for every tally row of a launch (one block per row)
cudaError_t launch_tally_prefix(const unsigned long long *diff, unsigned long long *tally, uint32_t rows, uint32_t N, cudaStream_t s);
int occupancy_woodcock(uint32_t G, bool big, bool trace, bool bank, int block, uint32_t smem);
cudaError_t launch_source(const TransportParams &p, bool bank, uint4 *out, cudaStream_t s);
cudaError_t launch_woodcock(const TransportParams &p, bool trace, bool bank, dim3 grid, dim3 block, uint32_t smem, cudaStream_t s);
cudaError_t prepare_woodcock(uint32_t smem_bytes, uint32_t G, bool trace, bool bank);
cudaError_t run_event_generation(const TransportParams &p, const EventBank &b, uint32_t smem, int sm_count, cudaStream_t s, uint32_t *iters);
uint32_t block_event_smem(const TransportParams &p, uint32_t slots);
cudaError_t launch_block_event(const TransportParams &p, dim3 grid, dim3 block, uint32_t smem, uint32_t slots, cudaStream_t s);
cudaError_t launch_bank_compact(const BankParams &p, cudaStream_t s);
cudaError_t launch_bank_histogram(const unsigned long long *bank, const unsigned long long *count_ptr, unsigned long long *hist, uint32_t N,
                                  cudaStream_t s);
cudaError_t launch_bank_entropy(const unsigned long long *hist, uint32_t N, double *entropy_out, unsigned long long *size_out, cudaStream_t s);
cudaError_t prepare_transport(uint32_t smem_bytes, uint32_t G, uint32_t surf_mode, bool trace, bool bank);
cudaError_t launch_finalize(const FinalizeParams &p, cudaStream_t s);
cudaError_t launch_probe_logf(const float *x, float *out, uint32_t n, cudaStream_t s);
cudaError_t launch_probe_div(const float *t, const float *mu, float *out_fast, float *out_ieee, uint32_t n, cudaStream_t s);
cudaError_t launch_probe_pcg(uint64_t state, uint64_t inc, uint32_t n, uint32_t *out_u32, float *out_unit, cudaStream_t s);

} // namespace nraps
