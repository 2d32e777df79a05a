// Fission-bank compaction and source-convergence diagnostic (new capability;
// the reference never updates its source from fission sites).
//
// The transport kernel leaves, per history, up to `cap` sites in a fixed slot
// row plus a count byte.  Compaction turns that into a dense bank in canonical
// (history, site) order -- the order is what makes the next generation's
// sampling independent of GPU scheduling and GPU count:
//   1. tile_sums:   1024 threads x 16 counts per block, warp-shuffle reduction
//   2. scan_sums:   one block, warp-shuffle exclusive scan over the tile sums
//   3. scatter:     per-thread offsets from a warp-shuffle scan inside the
//                   tile, then copy slots -> dense
// Shannon entropy of the bank over mesh cells is the usual source-convergence
// diagnostic for a power iteration.
#include "mc_internal.h"

namespace nraps {

namespace {

constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ uint32_t sum16(const uint4 &v)
{
    // 16 count bytes -> their sum (each byte <= 255)
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) s += (w[i] & 0xffu) + ((w[i] >> 8) & 0xffu) + ((w[i] >> 16) & 0xffu) + (w[i] >> 24);
    return s;
}

// inclusive scan across the block (1024 threads); returns this thread's inclusive value, total in *block_total
__device__ __forceinline__ uint32_t block_inclusive_scan(uint32_t v, uint32_t *warp_tot /*[32] smem*/, uint32_t *block_total)
{
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t up = __shfl_up_sync(kFull, v, o);
        if (lane >= (unsigned)o) v += up;
    }
    if (lane == 31) warp_tot[warp] = v;
    __syncthreads();
    if (warp == 0) {
        uint32_t t = warp_tot[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t up = __shfl_up_sync(kFull, t, o);
            if (lane >= (unsigned)o) t += up;
        }
        warp_tot[lane] = t; // inclusive totals of warps 0..lane
    }
    __syncthreads();
    const uint32_t before = warp ? warp_tot[warp - 1] : 0u;
    *block_total = warp_tot[31];
    return v + before;
}

__global__ void __launch_bounds__(1024) bank_tile_sums(const BankParams P)
{
    __shared__ uint32_t warp_tot[32];
    const uint64_t base = (uint64_t)blockIdx.x * kBankTile + (uint64_t)threadIdx.x * 16u;
    const uint4 c = *reinterpret_cast<const uint4 *>(P.counts + base); // counts are padded with zeros to a tile
    uint32_t total;
    block_inclusive_scan(sum16(c), warp_tot, &total);
    if (threadIdx.x == 0) P.block_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024) bank_scan_sums(const BankParams P)
{
    __shared__ unsigned long long warp_tot[32];
    __shared__ unsigned long long carry;
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0ull;
    __syncthreads();
    for (uint32_t base = 0; base < P.n_tiles; base += 1024) {
        const uint32_t i = base + threadIdx.x;
        const unsigned long long v = i < P.n_tiles ? P.block_sums[i] : 0ull;
        unsigned long long incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long up = __shfl_up_sync(kFull, incl, o);
            if (lane >= (unsigned)o) incl += up;
        }
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            unsigned long long t = warp_tot[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned long long up = __shfl_up_sync(kFull, t, o);
                if (lane >= (unsigned)o) t += up;
            }
            warp_tot[lane] = t;
        }
        __syncthreads();
        const unsigned long long c = carry;
        incl += warp ? warp_tot[warp - 1] : 0ull;
        if (i < P.n_tiles) P.block_sums[i] = c + incl - v; // exclusive offset of tile i
        const unsigned long long total = warp_tot[31];
        __syncthreads();
        if (threadIdx.x == 0) carry = c + total;
        __syncthreads();
    }
    if (threadIdx.x == 0) *P.count_out = carry < P.dense_cap ? carry : P.dense_cap;
}

__global__ void __launch_bounds__(1024) bank_scatter(const BankParams P)
{
    __shared__ uint32_t warp_tot[32];
    const uint64_t y0 = (uint64_t)blockIdx.x * kBankTile + (uint64_t)threadIdx.x * 16u;
    const uint4 c = *reinterpret_cast<const uint4 *>(P.counts + y0);
    const uint32_t mine = sum16(c);
    uint32_t total;
    const uint32_t incl = block_inclusive_scan(mine, warp_tot, &total);
    if (!mine) return;
    unsigned long long off = P.block_sums[blockIdx.x] + incl - mine;
    const uint32_t w[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const uint32_t n = (w[i >> 2] >> (8 * (i & 3))) & 0xffu;
        const unsigned long long *row = P.slots + (y0 + i) * P.cap;
        for (uint32_t j = 0; j < n; ++j, ++off)
            if (off < P.dense_cap) P.dense[off] = row[j];
    }
}

// Cell histogram of this rank's dense bank, added to the N words behind the tally and the counters of the generation's
// tally buffer: the buffer is summed across ranks once per generation anyway, so the histogram of the whole bank --
// what the entropy diagnostic needs -- comes with it.  Bins are privatised in shared memory when they fit (sites
// cluster on a few hundred fuel cells, so global atomics would serialise: 180 ms for 1e9 sites before).
__global__ void __launch_bounds__(1024) bank_histogram(const unsigned long long *bank, const unsigned long long *count_ptr,
                                                       unsigned long long *hist, uint32_t N, int use_smem)
{
    extern __shared__ uint32_t s_hist[];
    const unsigned long long n = *count_ptr;
    if (use_smem) {
        for (uint32_t i = threadIdx.x; i < N; i += blockDim.x) s_hist[i] = 0u;
        __syncthreads();
    }
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x) {
        const uint32_t cell = (uint32_t)(bank[i] >> 32);
        if (use_smem) atomicAdd(&s_hist[cell], 1u);
        else atomicAdd(&hist[cell], 1ull);
    }
    if (use_smem) {
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < N; i += blockDim.x)
            if (s_hist[i]) atomicAdd(&hist[i], (unsigned long long)s_hist[i]);
    }
}

// Shannon entropy (bits) of the bank over mesh cells and the bank's size, from the summed histogram
__global__ void __launch_bounds__(1024) bank_entropy(const unsigned long long *hist, uint32_t N, double *entropy_out, unsigned long long *size_out)
{
    __shared__ double part[32];
    __shared__ unsigned long long tot[32];
    unsigned long long cnt = 0ull;
    for (uint32_t i = threadIdx.x; i < N; i += blockDim.x) cnt += hist[i];
#pragma unroll
    for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(kFull, cnt, o);
    if ((threadIdx.x & 31) == 0) tot[threadIdx.x >> 5] = cnt;
    __syncthreads();
    unsigned long long total = 0ull;
    for (int w = 0; w < 32; ++w) total += tot[w];
    const double n = (double)total;
    double e = 0.0;
    for (uint32_t i = threadIdx.x; i < N; i += blockDim.x) {
        const unsigned long long h = hist[i];
        if (h) {
            const double p = (double)h / n;
            e -= p * log2(p);
        }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) e += __shfl_xor_sync(kFull, e, o);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = e;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 32; ++w) t += part[w];
        *entropy_out = t;
        *size_out = total;
    }
}

} // namespace

cudaError_t launch_bank_compact(const BankParams &p, cudaStream_t s)
{
    if (p.n_tiles == 0) return cudaMemsetAsync(p.count_out, 0, sizeof(unsigned long long), s);
    bank_tile_sums<<<p.n_tiles, 1024, 0, s>>>(p);
    bank_scan_sums<<<1, 1024, 0, s>>>(p);
    bank_scatter<<<p.n_tiles, 1024, 0, s>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_bank_histogram(const unsigned long long *bank, const unsigned long long *count_ptr, unsigned long long *hist, uint32_t N,
                                  cudaStream_t s)
{
    const int use_smem = N * sizeof(uint32_t) <= 96 * 1024;
    const size_t smem = use_smem ? N * sizeof(uint32_t) : 0;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(bank_histogram, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    bank_histogram<<<148 * 2, 1024, smem, s>>>(bank, count_ptr, hist, N, use_smem);
    return cudaGetLastError();
}

cudaError_t launch_bank_entropy(const unsigned long long *hist, uint32_t N, double *entropy_out, unsigned long long *size_out, cudaStream_t s)
{
    bank_entropy<<<1, 1024, 0, s>>>(hist, N, entropy_out, size_out);
    return cudaGetLastError();
}

} // namespace nraps
