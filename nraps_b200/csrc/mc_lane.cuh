// Lane-level building blocks shared by the transport kernels (surface tracking
// in mc_transport.cu, Woodcock delta tracking in mc_woodcock.cu).
#pragma once
#include "mc_device.cuh"
#include "mc_internal.h"

namespace nraps {

static constexpr unsigned kFull = 0xffffffffu;

// shared-space atomics on 32-bit shared addresses (the generic-pointer forms
// drag a cluster-window address computation into the inner loop)
static __device__ __forceinline__ uint32_t atoms_add(uint32_t saddr, uint32_t v)
{
    uint32_t old;
    asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(saddr), "r"(v) : "memory");
    return old;
}
static __device__ __forceinline__ void reds_add(uint32_t saddr, uint32_t v)
{
    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(saddr), "r"(v) : "memory");
}

static __device__ __forceinline__ float lds_f32(uint32_t saddr)
{
    float v;
    asm("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(saddr));
    return v;
}

static __device__ __forceinline__ uint32_t lds_u32(uint32_t saddr)
{
    uint32_t v;
    asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(saddr));
    return v;
}

static __device__ __forceinline__ uint32_t lds_u16(uint32_t saddr)
{
    uint32_t v;
    asm("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(saddr));
    return v;
}

static __device__ __forceinline__ uint32_t lds_u8(uint32_t saddr)
{
    uint32_t v;
    asm("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(saddr));
    return v;
}

// 64-bit fixed-point bin += score.  Shared-memory mode: two u32 words with an explicit carry, `ref` is the shared
// address of the low word and the high word sits hi_off bytes above.  BIG mode (mesh too large for shared
// memory): `ref` is the bin index and the add goes to the global 64-bit bin directly.
template <bool BIG>
static __device__ __forceinline__ void score(uint32_t ref, uint32_t hi_off, float v, unsigned long long *global_bins)
{
    const float vs = fmul(v, kTallyScale);
    const unsigned long long fx = __float2ull_rz(vs);
    if (BIG) {
        atomicAdd(global_bins + ref, fx);
        return;
    }
    const uint32_t l = (uint32_t)fx, h = (uint32_t)(fx >> 32);
    const uint32_t old = atoms_add(ref, l);
    // high word += h + carry; both are rare (a score >= 16 cm, or the low word wrapping), so one test guards the pair
    const bool carry = old > ~l;
    if (carry | (vs >= 4294967296.0f)) reds_add(ref + hi_off, h + (carry ? 1u : 0u)); // vs >= 2^32 <=> h != 0, one FSETP
}
// tally reference of bin `bin`: shared byte address or plain index
template <bool BIG> static __device__ __forceinline__ uint32_t tally_ref(uint32_t lo_base, int bin)
{
    return BIG ? (uint32_t)bin : lo_base + 4u * (uint32_t)bin;
}

template <int TG> static NRAPS_HD int search_cdf(const float *cdf, int G, float v)
{
    if (TG == 4) { // partition_point on 4 entries, probes 2 then 3 or 1 then 0
        const float4 c = *reinterpret_cast<const float4 *>(cdf);
        return (c.z < v) ? 3 : ((c.y < v) ? 2 : ((c.x < v) ? 1 : 0));
    }
    if (TG == 2) {
        const float2 c = *reinterpret_cast<const float2 *>(cdf);
        return ((c.y < v) || (c.x < v)) ? 1 : 0;
    }
    return lower_bound_clamped<TG>(cdf, G, v);
}

template <int TG> static __device__ __forceinline__ int search_cdf_global(const float *cdf, int G, float v)
{
    float c[8];
    const int n = TG ? TG : G;
#pragma unroll
    for (int j = 0; j < 8; ++j) c[j] = j < n ? __ldg(cdf + j) : 0.0f;
    int lo = 0, hi = n; // partition_point(|x| x < v).min(n - 1), same probes as lower_bound_clamped
    while (lo < hi) {
        const int mid = lo + ((hi - lo) >> 1);
        float cm = c[0];
#pragma unroll
        for (int j = 1; j < 8; ++j) cm = (j == mid) ? c[j] : cm;
        if (cm < v) lo = mid + 1;
        else hi = mid;
    }
    return lo < n - 1 ? lo : n - 1;
}

template <int TG>
static NRAPS_HD int sample_group(const float *cdf, int G, int mode, uint64_t &rng, uint64_t inc)
{
    const int n = TG ? TG : G;
    if (mode == NRAPS_SCATTER_SINGLE_XI) return search_cdf<TG>(cdf, G, pcg32_unit(rng, inc));
    if (mode == NRAPS_SCATTER_RUST_PRE182) { // a fresh draw per probe, pre-1.82 probe order
        int size = n, left = 0, right = n;
        while (left < right) {
            const int mid = left + size / 2;
            if (cdf[mid] < pcg32_unit(rng, inc)) left = mid + 1;
            else right = mid;
            size = right - left;
        }
        return left < n - 1 ? left : n - 1;
    }
    int size = n, base = 0; // rustc >= 1.82 probe order
    while (size > 1) {
        const int half = size / 2, mid = base + half;
        if (cdf[mid] < pcg32_unit(rng, inc)) base = mid;
        size -= half;
    }
    const int res = base + (cdf[base] < pcg32_unit(rng, inc) ? 1 : 0);
    return res < n - 1 ? res : n - 1;
}

// apply the jump maps selected by the set bits of `steps` (each map = stride * 2^b draws)
static __device__ __forceinline__ uint64_t jump_ahead(uint64_t state, uint64_t steps, const ulonglong2 *jump)
{
    while (steps) {
        const int b = __ffsll((long long)steps) - 1;
        steps &= steps - 1;
        const ulonglong2 J = jump[b];
        state = J.x * state + J.y;
    }
    return state;
}


// Block prologue shared by both kernels: carve the shared-memory image and fill it from global memory.
struct SmemView {
    uint32_t *lo, *hi;
    float *edges;
    uint32_t *runb;
    ulonglong2 *jump;
    float *xs;
    uint16_t *fuel;
    uint8_t *matid;
    uint16_t *bucket;
};

// Read-only mesh tables as the kernels look them up.  They sit in shared memory, or in BIG mode in global memory;
// the pointers of SmemView are generic because of that choice, and a lookup through them costs 64-bit address
// arithmetic plus a generic load.  With BIG known at compile time the shared case becomes a 32-bit address and an LDS.
template <bool BIG> struct MeshRef {
    const float *edges;
    const uint32_t *runb;
    const uint8_t *matid;
    const uint16_t *bucket;
    uint32_t s_edges, s_runb, s_matid, s_bucket; // shared byte addresses (!BIG)
    __device__ __forceinline__ explicit MeshRef(const SmemView &S) : edges(S.edges), runb(S.runb), matid(S.matid), bucket(S.bucket)
    {
        s_edges = BIG ? 0u : (uint32_t)__cvta_generic_to_shared(S.edges);
        s_runb = BIG ? 0u : (uint32_t)__cvta_generic_to_shared(S.runb);
        s_matid = BIG ? 0u : (uint32_t)__cvta_generic_to_shared(S.matid);
        s_bucket = BIG ? 0u : (uint32_t)__cvta_generic_to_shared(S.bucket);
    }
    __device__ __forceinline__ float edge(int i) const { return BIG ? __ldg(edges + i) : lds_f32(s_edges + 4u * (uint32_t)i); }
    __device__ __forceinline__ uint32_t run_bounds(int i) const { return BIG ? __ldg(runb + i) : lds_u32(s_runb + 4u * (uint32_t)i); }
    __device__ __forceinline__ int material(int i) const { return BIG ? (int)__ldg(matid + i) : (int)lds_u8(s_matid + (uint32_t)i); }
    __device__ __forceinline__ int bucket_cell(int i) const { return BIG ? (int)__ldg(bucket + i) : (int)lds_u16(s_bucket + 2u * (uint32_t)i); }
};

static __device__ __forceinline__ SmemView load_block_tables(unsigned char *smem_raw, const TransportParams &P, const SmemLayout &L)
{
    SmemView S;
    if (P.big) { // mesh tables stay in global memory (read-only, L1/L2 cached); stage only the small tables
        S.lo = S.hi = nullptr;
        S.edges = const_cast<float *>(P.edges); S.runb = const_cast<uint32_t *>(P.runb);
        S.fuel = const_cast<uint16_t *>(P.fuel); S.matid = const_cast<uint8_t *>(P.matid);
        S.bucket = const_cast<uint16_t *>(P.bucket);
        S.jump = reinterpret_cast<ulonglong2 *>(smem_raw + L.jump);
        S.xs = reinterpret_cast<float *>(smem_raw + L.xs);
        for (int i = threadIdx.x; i < (int)xs_floats(P.M, P.G); i += blockDim.x) S.xs[i] = P.xs[i];
        for (int i = threadIdx.x; i < 64; i += blockDim.x) S.jump[i] = P.jump[i];
        __syncthreads();
        return S;
    }
    S.lo = reinterpret_cast<uint32_t *>(smem_raw + L.tally_lo);
    S.hi = reinterpret_cast<uint32_t *>(smem_raw + L.tally_hi);
    S.edges = reinterpret_cast<float *>(smem_raw + L.edges);
    S.runb = reinterpret_cast<uint32_t *>(smem_raw + L.runb);
    S.jump = reinterpret_cast<ulonglong2 *>(smem_raw + L.jump);
    S.xs = reinterpret_cast<float *>(smem_raw + L.xs);
    S.fuel = reinterpret_cast<uint16_t *>(smem_raw + L.fuel);
    S.matid = smem_raw + L.matid;
    S.bucket = reinterpret_cast<uint16_t *>(smem_raw + L.bucket);
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int G = (int)P.G, M = (int)P.M, N = (int)P.N, GN = (int)P.rows * N, MG = M * G;
    for (int i = tid; i < GN; i += nthr) { S.lo[i] = 0u; S.hi[i] = 0u; }
    for (int i = tid; i <= N; i += nthr) S.edges[i] = P.edges[i];
    for (int i = tid; i < N; i += nthr) { S.runb[i] = P.runb[i]; S.matid[i] = P.matid[i]; }
    for (int i = tid; i < (int)P.NF; i += nthr) S.fuel[i] = P.fuel[i];
    for (int i = tid; i < (int)xs_floats(P.M, P.G); i += nthr) S.xs[i] = P.xs[i];
    for (int i = tid; i < 64; i += nthr) S.jump[i] = P.jump[i];
    for (int i = tid; i < (int)P.NB; i += nthr) S.bucket[i] = P.bucket[i];
    __syncthreads();
    return S;
}

// Block epilogue: shared bins -> global 64-bit bins, lane counters -> global counters.
static __device__ __forceinline__ void flush_block(const SmemView &S, const TransportParams &P, const uint32_t (&vals)[8])
{
    const int tid = threadIdx.x, nthr = blockDim.x, GN = (int)(P.rows * P.N);
    __syncthreads();
    for (int i = tid; i < GN && !P.big; i += nthr) {
        const unsigned long long v = ((unsigned long long)S.hi[i] << 32) + S.lo[i];
        if (v) atomicAdd(&P.tally[i], v);
    }
    unsigned long long *ct = P.tally + GN;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        unsigned long long v = vals[c];
#pragma unroll
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
        if ((tid & 31) == 0 && v) atomicAdd(&ct[c], v);
    }
}

} // namespace nraps
