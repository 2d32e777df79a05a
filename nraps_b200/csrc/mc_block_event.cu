// CUDA side of the block-level event pipeline (see mc_block_event.cuh): the context type that maps the shared
// per-thread body onto a thread block of sm_100a, the kernel and its launcher.  EXPERIMENTAL, opt-in
// (nraps_options.kernel_variant = NRAPS_KERNEL_BLOCK_EVENT); transport_kernel stays the default.
#include "mc_block_event.cuh"

namespace nraps {

namespace {

struct DevCtx {
    long long t0; // clock64() at block start, for the watchdog
    // A block that is still running after 2e10 SM cycles (~10 s; a generation of 1.25e8 histories takes 0.2 s) stops
    // and reports its live records as truncated histories: an experimental kernel must not be able to hang a GPU.
    __device__ __forceinline__ bool expired() const { return clock64() - t0 > 20000000000ll; }
    static constexpr bool kStats = false; // the schedule statistics hooks exist for the CPU emulation only
    __device__ __forceinline__ void note_walk(uint32_t, int) const {}
    __device__ __forceinline__ void note_round(uint32_t, uint32_t, uint32_t) const {}
    uint32_t s_edges, s_runb, s_matid, s_lo, hi_off;
    const float *xs;
    int MG;
    __device__ __forceinline__ uint32_t tid() const { return threadIdx.x; }
    __device__ __forceinline__ uint32_t nthreads() const { return blockDim.x; }
    __device__ __forceinline__ void sync() const { __syncthreads(); }
    __device__ __forceinline__ void converge() const { __syncwarp(); }
    __device__ __forceinline__ uint32_t atomic_add_shared(uint32_t *p, uint32_t v) const { return atomicAdd(p, v); }
    __device__ __forceinline__ uint32_t load_shared(const uint32_t *p) const { return *reinterpret_cast<const volatile uint32_t *>(p); }
    __device__ __forceinline__ unsigned long long atomic_add_global(unsigned long long *p, unsigned long long v) const { return atomicAdd(p, v); }
    // Positions in the two lists of an arena for the lanes that push (at most one of the predicates holds per lane):
    // ONE shared atomic per warp on the packed length word (low half: list growing up, high half: list growing down).
    __device__ __forceinline__ uint32_t claim2(uint32_t *word, bool up, bool down) const
    {
        const unsigned m_up = __ballot_sync(kFull, up), m_down = __ballot_sync(kFull, down), m = m_up | m_down;
        if (!m) return 0u;
        const unsigned lane = threadIdx.x & 31u;
        const int leader = __ffs((int)m) - 1;
        uint32_t old = 0u;
        if ((int)lane == leader) old = atomicAdd(word, bev::claim2_increment(m_up, m_down));
        old = __shfl_sync(kFull, old, leader);
        return bev::claim2_position(m_up, m_down, lane, old, down);
    }
    __device__ __forceinline__ uint32_t run_bounds(int i) const { return lds_u32(s_runb + 4u * (uint32_t)i); }
    __device__ __forceinline__ int material(int i) const { return (int)lds_u8(s_matid + (uint32_t)i); }
    __device__ __forceinline__ uint32_t edge_ref(int i) const { return s_edges + 4u * (uint32_t)i; }
    __device__ __forceinline__ float edge(uint32_t ref) const { return lds_f32(ref); }
    __device__ __forceinline__ uint32_t tally_ref(int bin) const { return s_lo + 4u * (uint32_t)bin; }
    __device__ __forceinline__ void score(uint32_t ref, float v) const { nraps::score<false>(ref, hi_off, v, nullptr); }
    __device__ __forceinline__ float inv_sigtr(int i) const { return xs[i]; }
    __device__ __forceinline__ float p_abs(int i) const { return xs[MG + i]; }
    __device__ __forceinline__ const float *scat_cdf(int off) const { return xs + 5 * MG + off; }
    __device__ __forceinline__ void load_record(const uint4 *rec, uint32_t (&r0)[4], uint32_t (&r1)[4]) const
    {
        const uint4 a = __ldg(rec), b = __ldg(rec + 1);
        r0[0] = a.x; r0[1] = a.y; r0[2] = a.z; r0[3] = a.w;
        r1[0] = b.x; r1[1] = b.y; r1[2] = b.z; r1[3] = b.w;
    }
};

template <int TG> __global__ void __launch_bounds__(1024, 1) block_event_kernel(const TransportParams P, const uint32_t S)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const SmemLayout L = make_layout(P.M, P.G, P.N, P.NF, P.NB, 0, P.rows);
    const SmemView V = load_block_tables(smem_raw, P, L); // zeroes the tally image, stages the tables, __syncthreads
    const bev::Bank b = bev::make_bank(smem_raw + L.total, S);
    DevCtx c;
    c.t0 = clock64();
    c.s_edges = (uint32_t)__cvta_generic_to_shared(V.edges);
    c.s_runb = (uint32_t)__cvta_generic_to_shared(V.runb);
    c.s_matid = (uint32_t)__cvta_generic_to_shared(V.matid);
    c.s_lo = (uint32_t)__cvta_generic_to_shared(V.lo);
    c.hi_off = L.tally_hi - L.tally_lo;
    c.xs = V.xs;
    c.MG = (int)(P.M * P.G);
    bev::Counts ct;
    bev::block_event_thread<TG>(c, P, b, ct);
    const uint32_t vals[8] = {ct.hist, ct.coll, 0u, ct.flight, 0u, ct.leak, ct.trunc, 0u};
    flush_block(V, P, vals);
}

template <int TG> cudaError_t launch_g(const TransportParams &p, dim3 grid, dim3 block, uint32_t smem, uint32_t S, cudaStream_t s)
{
    cudaError_t e = cudaFuncSetAttribute(block_event_kernel<TG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    block_event_kernel<TG><<<grid, block, smem, s>>>(p, S);
    return cudaGetLastError();
}

} // namespace

// dynamic shared memory of one block: the mesh / tally image of the lane kernels followed by the neutron bank
uint32_t block_event_smem(const TransportParams &p, uint32_t S)
{
    return make_layout(p.M, p.G, p.N, p.NF, p.NB, 0, p.rows).total + align_up(bev::bank_bytes(S), 16);
}

cudaError_t launch_block_event(const TransportParams &p, dim3 grid, dim3 block, uint32_t smem, uint32_t S, cudaStream_t s)
{
    switch (p.G) {
    case 2: return launch_g<2>(p, grid, block, smem, S, s);
    case 4: return launch_g<4>(p, grid, block, smem, S, s);
    default: return launch_g<0>(p, grid, block, smem, S, s);
    }
}

} // namespace nraps
