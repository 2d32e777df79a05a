// C-ABI layer: validation, derived tables, device residency, launch order.
//
// What is precomputed on the host (once per context) and why it is bit-safe:
// every derived table is a pure function of the input tables evaluated with the
// reference's own binary32 operations in the reference's order, so looking it
// up on the device equals recomputing it per collision as the reference does.
//   p_abs[xs]   = siga[xs] / sigt[xs] ...................... src/mc_code.rs:121
//   scat_cdf[mat][g][xs_g][j] = cumsum_j(scat row (mat,g)) * (1/sigs[mat+M*xs_g])
//                                                            src/mc_code.rs:89-100,190
//   chi_cdf[mat][j] = cumsum_j chit[mat + M*j] .............. src/mc_code.rs:19-30
//   run bounds per cell: first/last+1 cell of its material run (replaces the
//   per-crossing matid compare of src/mc_code.rs:175-181)
//   PCG32 jump table (A_b, C_b): the affine map of stride*2^b LCG steps
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include <cuda.h>
#include <sys/syscall.h>
#include <unistd.h>

#include "../../include/nraps_host.h"
#include "mc_internal.h"

using namespace nraps;

namespace {

thread_local std::string g_cuda_error;

int cuda_fail(cudaError_t e, const char *what)
{
    g_cuda_error = std::string(what) + ": " + cudaGetErrorString(e);
    return NRAPS_ERR_CUDA;
}

#define CU(call)                                             \
    do {                                                     \
        cudaError_t e_ = (call);                             \
        if (e_ != cudaSuccess) return cuda_fail(e_, #call);  \
    } while (0)

constexpr uint64_t kPcgMult = 6364136223846793005ULL;
constexpr uint32_t kMaxSmem = 232448; // 227 KB opt-in limit per block on sm_100

struct Pcg { uint64_t state, inc; };

uint32_t pcg_step(Pcg &r)
{
    const uint64_t old = r.state;
    r.state = old * kPcgMult + r.inc;
    const uint32_t xs = (uint32_t)(((old >> 18) ^ old) >> 27);
    const uint32_t rot = (uint32_t)(old >> 59);
    return (xs >> rot) | (xs << ((32u - rot) & 31u));
}

Pcg pcg_seed(uint64_t seed, uint64_t seq) // src/rand.rs:49-71 with an explicit seed
{
    Pcg r{0u, (seq << 1) | 1u};
    pcg_step(r);
    r.state += seed;
    pcg_step(r);
    return r;
}

// affine map of `delta` LCG steps: state' = mult*state + plus
void pcg_jump_coeffs(uint64_t inc, uint64_t delta, uint64_t *mult, uint64_t *plus)
{
    uint64_t cm = kPcgMult, cp = inc, am = 1u, ap = 0u;
    while (delta) {
        if (delta & 1u) { am *= cm; ap = ap * cm + cp; }
        cp = (cm + 1u) * cp;
        cm *= cm;
        delta >>= 1;
    }
    *mult = am;
    *plus = ap;
}

// Device buffers come from a library-owned stream-ordered pool per device instead of cudaMalloc / cudaFree: driver
// allocation calls cost 50-300 ms per context on the gpurun boxes (a 320 MB birth-record buffer mapped and unmapped
// per monte_carlo() call), which made the end-to-end time of a 0.3 s run vary by 2x.  The pool keeps up to 2 GiB
// mapped between contexts (nraps_mc_trim releases it); both wrappers keep cudaMalloc's / cudaFree's synchronous
// contract, so callers need no stream ordering.
std::mutex g_pool_mutex;
cudaMemPool_t g_pools[64] = {};

cudaError_t device_pool(cudaMemPool_t *out)
{
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
    std::lock_guard<std::mutex> lock(g_pool_mutex);
    if (!g_pools[dev]) {
        cudaMemPoolProps props{};
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = dev;
        cudaMemPool_t pool = nullptr;
        if ((e = cudaMemPoolCreate(&pool, &props)) != cudaSuccess) return e;
        uint64_t keep = 2ull << 30;
        if ((e = cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep)) != cudaSuccess) {
            cudaMemPoolDestroy(pool);
            return e;
        }
        g_pools[dev] = pool;
    }
    *out = g_pools[dev];
    return cudaSuccess;
}

cudaError_t dev_malloc(void **p, size_t bytes)
{
    cudaMemPool_t pool = nullptr;
    cudaError_t e = device_pool(&pool);
    if (e != cudaSuccess) return e;
    if ((e = cudaMallocFromPoolAsync(p, bytes, pool, nullptr)) != cudaSuccess) return e;
    return cudaStreamSynchronize(nullptr); // from here on the buffer may be used on any stream
}

cudaError_t dev_free(void *p)
{
    if (!p) return cudaSuccess;
    cudaError_t e = cudaDeviceSynchronize(); // like cudaFree: nothing in flight may still use the buffer
    if (e != cudaSuccess) return e;
    return cudaFreeAsync(p, nullptr);
}

// ---- Bank buffers other PROCESSES map (one process per GPU).  Legacy CUDA IPC (cudaIpcOpenMemHandle) maps the peer's
// memory with small pages: random 8-byte reads over a few GB of it thrash the TLB -- measured on 8 B200s, births of
// 1.25e8 histories per GPU took 1.3 s with IPC mappings against 10 ms through plain peer pointers in one process
// (profiles/r2_bank8_ipc_vs_vmm.txt).  So the buffers are created with the virtual memory management API (cuMemCreate,
// 2 MB granularity on both sides), exported as POSIX file descriptors and duplicated into the importing process with
// pidfd_getfd.  The driver API is reached through cudaGetDriverEntryPoint: the library keeps no link-time dependency on
// libcuda and still loads on a host without a driver.
struct DriverApi {
    decltype(&cuMemCreate) memCreate = nullptr;
    decltype(&cuMemRelease) memRelease = nullptr;
    decltype(&cuMemAddressReserve) addressReserve = nullptr;
    decltype(&cuMemAddressFree) addressFree = nullptr;
    decltype(&cuMemMap) memMap = nullptr;
    decltype(&cuMemUnmap) memUnmap = nullptr;
    decltype(&cuMemSetAccess) setAccess = nullptr;
    decltype(&cuMemGetAllocationGranularity) granularity = nullptr;
    decltype(&cuMemExportToShareableHandle) exportHandle = nullptr;
    decltype(&cuMemImportFromShareableHandle) importHandle = nullptr;
    bool ok = false;
};

const DriverApi &driver_api()
{
    static DriverApi api = [] {
        DriverApi a;
        auto get = [](const char *name, void **fn) {
            cudaDriverEntryPointQueryResult q;
            return cudaGetDriverEntryPoint(name, fn, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess && *fn;
        };
        a.ok = get("cuMemCreate", (void **)&a.memCreate) && get("cuMemRelease", (void **)&a.memRelease) &&
               get("cuMemAddressReserve", (void **)&a.addressReserve) && get("cuMemAddressFree", (void **)&a.addressFree) &&
               get("cuMemMap", (void **)&a.memMap) && get("cuMemUnmap", (void **)&a.memUnmap) && get("cuMemSetAccess", (void **)&a.setAccess) &&
               get("cuMemGetAllocationGranularity", (void **)&a.granularity) && get("cuMemExportToShareableHandle", (void **)&a.exportHandle) &&
               get("cuMemImportFromShareableHandle", (void **)&a.importHandle);
        return a;
    }();
    return api;
}

struct VmmBuffer { // one mapping of one physical allocation in this process
    CUmemGenericAllocationHandle handle = 0;
    CUdeviceptr ptr = 0;
    size_t size = 0;
    bool mapped = false;
};

int vmm_fail(CUresult r, const char *what)
{
    g_cuda_error = std::string(what) + ": CUresult " + std::to_string((int)r);
    return NRAPS_ERR_CUDA;
}

void vmm_release(VmmBuffer &b)
{
    const DriverApi &d = driver_api();
    if (!d.ok) return;
    if (b.mapped) d.memUnmap(b.ptr, b.size);
    if (b.ptr) d.addressFree(b.ptr, b.size);
    if (b.handle) d.memRelease(b.handle);
    b = VmmBuffer{};
}

// map `b.handle` (created here or imported) read-write for `device`
int vmm_map(VmmBuffer &b, int device)
{
    const DriverApi &d = driver_api();
    CUresult r;
    if ((r = d.addressReserve(&b.ptr, b.size, 0, 0, 0)) != CUDA_SUCCESS) return vmm_fail(r, "cuMemAddressReserve");
    if ((r = d.memMap(b.ptr, b.size, 0, b.handle, 0)) != CUDA_SUCCESS) return vmm_fail(r, "cuMemMap");
    b.mapped = true;
    CUmemAccessDesc acc{};
    acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    acc.location.id = device;
    acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    if ((r = d.setAccess(b.ptr, b.size, &acc, 1)) != CUDA_SUCCESS) return vmm_fail(r, "cuMemSetAccess");
    return NRAPS_OK;
}

int vmm_create(VmmBuffer &b, size_t bytes, int device)
{
    const DriverApi &d = driver_api();
    if (!d.ok) return cuda_fail(cudaErrorNotSupported, "CUDA virtual memory management API unavailable");
    CUmemAllocationProp prop{};
    prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    prop.location.id = device;
    prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    size_t gran = 0;
    CUresult r;
    if ((r = d.granularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED)) != CUDA_SUCCESS) return vmm_fail(r, "cuMemGetAllocationGranularity");
    b.size = (bytes + gran - 1) / gran * gran;
    if ((r = d.memCreate(&b.handle, b.size, &prop, 0)) != CUDA_SUCCESS) return vmm_fail(r, "cuMemCreate");
    int rc = vmm_map(b, device);
    if (rc != NRAPS_OK) vmm_release(b);
    return rc;
}

// what one rank publishes per bank buffer (NRAPS_IPC_HANDLE_BYTES each): enough for another process to duplicate the fd
struct BankTicket {
    uint32_t magic;
    int32_t pid, fd;
    uint32_t pad;
    uint64_t size;
};
static_assert(sizeof(BankTicket) <= NRAPS_IPC_HANDLE_BYTES, "ticket fits the handle slot");
constexpr uint32_t kTicketMagic = 0x4b4e4142u;

template <typename T> cudaError_t upload(T **dst, const std::vector<T> &src)
{
    cudaError_t e = dev_malloc(reinterpret_cast<void **>(dst), std::max<size_t>(1, src.size()) * sizeof(T));
    if (e != cudaSuccess) return e;
    if (src.empty()) return cudaSuccess;
    return cudaMemcpy(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice);
}

} // namespace

struct nraps_mc_ctx {
    nraps_options opt{};
    uint32_t M = 0, G = 0, N = 0, NF = 0, numass = 0;
    uint64_t generations = 0, histories = 0, skip = 0;
    float boundl = 0, boundr = 0, dx_fuel = 0, length = 0, nut_m1 = 0, k0 = 1.0f;
    int device = 0, sm_count = 0;
    Pcg master{};
    uint64_t stride = 0;
    SmemLayout layout{};
    uint32_t smem_total = 0;       // dynamic shared memory of the transport launch
    uint32_t surf_mode = SURF_SPLIT; // tally placement of the surface kernel (mc_internal.h)
    uint32_t skip_walk = 0;          // closed-form strides through long segments (set when the mean segment is >= 16 cells)
    uint32_t grid = 0, block = 0, blocks_per_sm = 0, chunk = 0, max_flights = 0;
    uint32_t geo_grid[2] = {0, 0}, geo_block[2] = {0, 0}; // launch geometry of the plain / trace instantiation

    float *d_edges = nullptr, *d_xs = nullptr, *d_dx = nullptr, *d_nut = nullptr, *d_sigf = nullptr;
    uint32_t *d_runb = nullptr;
    uint2 *d_segw = nullptr;                 // surface kernel: segment bounds + width bits per cell
    // Per-launch scratch that a transport launch owns until its kernels are done: the difference array of the surface
    // kernel's full-cell scores [batch*G*N], the chunk cursor, the birth records of the shard.  Two sets ("lanes",
    // nraps_mc_select_lane): with the uniform source generation g+1 may be launched on a second stream while the tail
    // of generation g still runs -- its blocks move in as blocks of g retire.
    struct LaneBuffers {
        unsigned long long *d_diff = nullptr, *d_work = nullptr;
        uint4 *d_source = nullptr; // born neutrons of the current shard (source_kernel -> transport kernels)
        uint64_t source_cap = 0;
    } lanes[2];
    int lane = 0;
    uint64_t diff_words = 0; // size of a d_diff
    uint8_t *d_matid = nullptr;
    uint16_t *d_fuel = nullptr, *d_bucket = nullptr;
    uint32_t NB = 0;
    float inv_h = 0.0f;
    bool woodcock = false;
    uint32_t prepared = 0, big = 0;
    uint32_t batch = 1; // generations one launch may carry (small generations do not fill the GPU on their own)
    ulonglong2 *d_jump = nullptr;
    unsigned long long *d_tally_own = nullptr, *d_tally = nullptr, *d_counters_total = nullptr;
    double *d_res_moments = nullptr;
    float *d_terms = nullptr, *d_res_flux = nullptr, *d_res_fission = nullptr, *d_k_hist = nullptr, *d_k_cur = nullptr;
    uint32_t *d_trace = nullptr;
    uint64_t trace_cap = 0;

    // fission_bank source mode
    bool bank_mode = false;
    uint32_t bank_cap = 8;
    uint64_t bank_hist_cap = 0, dense_cap = 0; // histories / sites the buffers below are sized for
    unsigned long long *d_slots = nullptr, *d_block_sums = nullptr;
    // the two bank buffers of this rank ([kBankHeader words: site count first][sites]), written alternately
    unsigned long long *d_bank[2] = {nullptr, nullptr};
    int bank_kind = 0;                           // 0: pool memory; 1: cudaMalloc (peer access inside one process); 2: cuMemCreate
                                                 // (mapped by other processes through exported file descriptors)
    VmmBuffer bank_vmm[2];                       // kind 2: this rank's allocations
    VmmBuffer peer_vmm[2][kMaxPeers];            // kind 2: the peers' allocations mapped here
    int bank_fd[2] = {-1, -1};                   // kind 2: exported file descriptors (peers duplicate them)
    int bank_world = 1, bank_rank = 0;           // ranks whose banks together feed a generation
    const unsigned long long *peer_bank[2][kMaxPeers] = {}; // [buffer][rank]; own entries point at d_bank
    unsigned long long *d_bank_sizes = nullptr;  // [generations]
    double *d_entropy = nullptr;                 // [generations]
    uint8_t *d_counts = nullptr;
    int bank_which = 0;                          // buffer the next compaction writes
    int bank_src = -1;                           // buffer generation g+1 samples from (-1: no bank yet, uniform source)
    int bank_last = -1;                          // buffer of the last compaction (nraps_mc_bank_local)
    uint64_t last_shard = 0;

    // profile_phases: event pairs around each phase of the generation in flight, folded into phase_ms when the next
    // generation (or a query) comes along
    cudaEvent_t ph_ev[2 * NRAPS_PH_WORDS] = {};
    bool ph_armed[NRAPS_PH_WORDS] = {};
    double phase_ms[NRAPS_PH_WORDS] = {};

    // event-based pipeline (kernel_variant = NRAPS_KERNEL_EVENT)
    EventBank ev{};
    uint64_t ev_cap = 0;
    uint32_t ev_iterations = 0; // advance/collide/compact rounds of the last generation

    // block-level event pipeline (kernel_variant = NRAPS_KERNEL_BLOCK_EVENT): launch geometry chosen at creation
    uint32_t bev_block = 0, bev_grid = 0, bev_slots = 0, bev_smem = 0;
};

namespace {

int validate(const nraps_problem *p, const nraps_options *o)
{
    if (!p || !o) return NRAPS_ERR_NULL;
    if (!p->sigt || !p->sigs || !p->mu || !p->siga || !p->sigf || !p->nut || !p->chit || !p->inv_sigtr || !p->scat ||
        !p->matid || !p->dx || !p->left || !p->right || !p->fuel_indices)
        return NRAPS_ERR_NULL;
    // G >= 2: the flux conversion reads nut[M*1] (src/mc_code.rs:356); skip < generations: k_fund[skip] (:368)
    if (p->M == 0 || p->M > 64 || p->G < 2 || p->G > 8 || p->N == 0 || p->N > 65535 || p->NF == 0 || p->numass == 0 ||
        p->numass > p->N || p->generations == 0 || p->skip >= p->generations || p->histories == 0)
        return NRAPS_ERR_SHAPE;
    if ((uint64_t)p->M * p->G * p->G * p->G > 8192) return NRAPS_ERR_TOO_LARGE;
    if (o->scatter_mode < 0 || o->scatter_mode > NRAPS_SCATTER_RUST_182) return NRAPS_ERR_OPTION;
    if (o->source_mode < 0 || o->source_mode > NRAPS_SOURCE_FISSION_BANK || o->tracking_mode < 0 || o->tracking_mode > NRAPS_TRACK_WOODCOCK ||
        o->kernel_variant < 0 || o->kernel_variant > NRAPS_KERNEL_BLOCK_EVENT || o->bank_cap < 0 || o->bank_cap > 255 ||
        o->slots_per_thread < 0 || o->slots_per_thread > 64)
        return NRAPS_ERR_OPTION;
    // The block-level event pipeline was measured on a B200 in round 2 (bit-exact, 0.72x the fused kernel: DESIGN.md
    // section 5) and is compiled out of the product library; `make BLOCK_EVENT=1` builds it back in for experiments.
#ifndef NRAPS_WITH_BLOCK_EVENT
    if (o->kernel_variant == NRAPS_KERNEL_BLOCK_EVENT) return NRAPS_ERR_OPTION;
#endif
    if (o->kernel_variant == NRAPS_KERNEL_BLOCK_EVENT &&
        (o->tracking_mode != NRAPS_TRACK_SURFACE || o->source_mode != NRAPS_SOURCE_UNIFORM_FUEL))
        return NRAPS_ERR_OPTION;
    // an albedo outside [0, 1] makes the reflected direction cosine leave [-1, 1] (mu' = -mu * b, src/mc_code.rs:56-62)
    // and lets the Woodcock wall loop diverge; NaN fails both comparisons
    if (!(p->boundl >= 0.0f && p->boundl <= 1.0f) || !(p->boundr >= 0.0f && p->boundr <= 1.0f)) return NRAPS_ERR_SHAPE;
    // the event pipeline is built for Woodcock tracking with the uniform source (one event = one tentative collision)
    if (o->kernel_variant == NRAPS_KERNEL_EVENT &&
        (o->tracking_mode != NRAPS_TRACK_WOODCOCK || o->source_mode != NRAPS_SOURCE_UNIFORM_FUEL || o->max_flights > 0xfffffull))
        return NRAPS_ERR_OPTION;
    for (uint32_t i = 0; i < p->N; ++i) {
        if (p->matid[i] >= p->M) return NRAPS_ERR_MESH;
        if (i + 1 < p->N && std::memcmp(&p->right[i], &p->left[i + 1], sizeof(float)) != 0) return NRAPS_ERR_MESH;
        if (!(p->right[i] > p->left[i])) return NRAPS_ERR_MESH;
        for (uint32_t g = 0; g < p->G; ++g) {
            const float v = p->inv_sigtr[p->matid[i] + p->M * g];
            if (!(v > 0.0f) || !std::isfinite(v)) return NRAPS_ERR_XS;
        }
    }
    for (uint32_t j = 0; j < p->NF; ++j)
        if (p->fuel_indices[j] >= p->N) return NRAPS_ERR_MESH;
    if (!(p->k0 > 0.0f) || !std::isfinite(p->k0)) return NRAPS_ERR_SHAPE;
    return NRAPS_OK;
}

// phase timing (nraps_options.profile_phases): begin/end record an event pair on the stream, fold() waits for the
// pairs recorded so far and adds their elapsed times
void phase_fold(nraps_mc_ctx *c)
{
    for (int p = 0; p < NRAPS_PH_WORDS; ++p) {
        if (!c->ph_armed[p]) continue;
        float ms = 0.0f;
        if (cudaEventSynchronize(c->ph_ev[2 * p + 1]) == cudaSuccess && cudaEventElapsedTime(&ms, c->ph_ev[2 * p], c->ph_ev[2 * p + 1]) == cudaSuccess)
            c->phase_ms[p] += (double)ms;
        c->ph_armed[p] = false;
    }
}
void phase_begin(nraps_mc_ctx *c, int p, cudaStream_t s)
{
    if (!c->opt.profile_phases) return;
    if (c->ph_armed[p]) phase_fold(c);
    if (!c->ph_ev[2 * p]) { cudaEventCreate(&c->ph_ev[2 * p]); cudaEventCreate(&c->ph_ev[2 * p + 1]); }
    cudaEventRecord(c->ph_ev[2 * p], s);
}
void phase_end(nraps_mc_ctx *c, int p, cudaStream_t s)
{
    if (!c->opt.profile_phases) return;
    cudaEventRecord(c->ph_ev[2 * p + 1], s);
    c->ph_armed[p] = true;
}

void free_bank_buffers(nraps_mc_ctx *c)
{
    if (c->bank_kind != 0) cudaDeviceSynchronize();
    for (int w = 0; w < 2; ++w) {
        for (int r = 0; r < kMaxPeers; ++r) {
            vmm_release(c->peer_vmm[w][r]);
            c->peer_bank[w][r] = nullptr;
        }
        if (c->bank_kind == 2) vmm_release(c->bank_vmm[w]);
        else if (c->bank_kind == 1) cudaFree(c->d_bank[w]);
        else dev_free(c->d_bank[w]);
        if (c->bank_fd[w] >= 0) close(c->bank_fd[w]);
        c->bank_fd[w] = -1;
        c->d_bank[w] = nullptr;
    }
    c->bank_kind = 0;
}

void free_ctx(nraps_mc_ctx *c)
{
    if (!c) return;
    dev_free(c->d_edges); dev_free(c->d_xs); dev_free(c->d_dx); dev_free(c->d_nut); dev_free(c->d_sigf);
    dev_free(c->d_segw);
    for (auto &ln : c->lanes) { dev_free(ln.d_diff); dev_free(ln.d_work); dev_free(ln.d_source); }
    dev_free(c->d_runb); dev_free(c->d_matid); dev_free(c->d_fuel); dev_free(c->d_jump); dev_free(c->d_bucket);
    dev_free(c->d_tally_own); dev_free(c->d_counters_total);
    dev_free(c->d_res_moments); dev_free(c->d_terms); dev_free(c->d_res_flux); dev_free(c->d_res_fission); dev_free(c->d_k_hist); dev_free(c->d_k_cur);
    dev_free(c->d_trace);
    dev_free(c->d_slots); dev_free(c->d_block_sums);
    free_bank_buffers(c);
    for (cudaEvent_t e : c->ph_ev)
        if (e) cudaEventDestroy(e);
    for (EventHalf &h : c->ev.half) { dev_free(h.x); dev_free(h.mu); dev_free(h.pack); dev_free(h.cnt); dev_free(h.ccnt); dev_free(h.rng); }
    dev_free(c->ev.n_alive);
    dev_free(c->d_bank_sizes); dev_free(c->d_entropy); dev_free(c->d_counts);
    delete c;
}

int ensure_event_bank(nraps_mc_ctx *c, uint64_t count)
{
    if (count <= c->ev_cap) return NRAPS_OK;
    for (EventHalf &h : c->ev.half) {
        dev_free(h.x); dev_free(h.mu); dev_free(h.pack); dev_free(h.cnt); dev_free(h.ccnt); dev_free(h.rng);
        h = EventHalf{};
        CU(dev_malloc((void **)&h.x, count * sizeof(float)));
        CU(dev_malloc((void **)&h.mu, count * sizeof(float)));
        CU(dev_malloc((void **)&h.pack, count * sizeof(uint32_t)));
        CU(dev_malloc((void **)&h.cnt, count * sizeof(uint32_t)));
        CU(dev_malloc((void **)&h.ccnt, count * sizeof(uint32_t)));
        CU(dev_malloc((void **)&h.rng, count * sizeof(unsigned long long)));
    }
    if (!c->ev.n_alive) {
        CU(dev_malloc((void **)&c->ev.n_alive, 2 * sizeof(unsigned long long)));
        c->ev.n_next = c->ev.n_alive + 1;
    }
    c->ev_cap = count;
    return NRAPS_OK;
}

// Size the per-history slot rows and the two bank buffers for a shard of `count` histories.  kind: 0 = pool memory
// (one GPU), 1 = cudaMalloc (peer access from the other devices of this process), 2 = cuMemCreate (other processes).
int alloc_bank(nraps_mc_ctx *c, uint64_t count, int kind, cudaStream_t s)
{
    const uint64_t padded = (count + kBankTile - 1) / kBankTile * kBankTile;
    // every history keeps at most bank_cap sites, so this bound is exact: no generation can overflow the dense bank
    // (a tighter guess of 3 sites per history truncated the first generations of problems with k / k0 > 3, found by
    // the GPU fuzz in round 2)
    const uint64_t dense_cap = count * c->bank_cap + 1024;
    const size_t bytes = (kBankHeader + dense_cap) * sizeof(unsigned long long);
    unsigned long long *fresh[2] = {nullptr, nullptr};
    VmmBuffer fresh_vmm[2];
    for (int w = 0; w < 2; ++w) {
        if (kind == 2) {
            int rc = vmm_create(fresh_vmm[w], bytes, c->device);
            if (rc != NRAPS_OK) { vmm_release(fresh_vmm[0]); return rc; }
            fresh[w] = reinterpret_cast<unsigned long long *>(fresh_vmm[w].ptr);
        } else if (kind == 1) CU(cudaMalloc((void **)&fresh[w], bytes));
        else CU(dev_malloc((void **)&fresh[w], bytes));
        CU(cudaMemsetAsync(fresh[w], 0, kBankHeader * sizeof(unsigned long long), s));
    }
    // the bank the next generation samples from may live in the buffers about to be replaced: carry it over
    for (int w = 0; w < 2; ++w)
        if (c->d_bank[w]) CU(cudaMemcpyAsync(fresh[w], c->d_bank[w], (kBankHeader + c->dense_cap) * sizeof(unsigned long long), cudaMemcpyDeviceToDevice, s));
    CU(cudaStreamSynchronize(s));
    dev_free(c->d_slots); dev_free(c->d_counts); dev_free(c->d_block_sums);
    c->d_slots = c->d_block_sums = nullptr;
    c->d_counts = nullptr;
    free_bank_buffers(c);
    for (int w = 0; w < 2; ++w) {
        c->d_bank[w] = fresh[w];
        c->bank_vmm[w] = fresh_vmm[w];
        c->peer_bank[w][c->bank_rank] = fresh[w];
    }
    c->bank_kind = kind;
    c->dense_cap = dense_cap;
    c->bank_hist_cap = 0;
    CU(dev_malloc((void **)&c->d_slots, padded * c->bank_cap * sizeof(unsigned long long)));
    CU(dev_malloc((void **)&c->d_counts, padded));
    CU(dev_malloc((void **)&c->d_block_sums, (padded / kBankTile) * sizeof(unsigned long long)));
    c->bank_hist_cap = count;
    return NRAPS_OK;
}

int ensure_bank(nraps_mc_ctx *c, uint64_t count, cudaStream_t s)
{
    if (count <= c->bank_hist_cap) return NRAPS_OK;
    if (c->bank_world > 1) return NRAPS_ERR_STATE; // peers hold mappings of the current buffers: reserve for the largest shard first
    return alloc_bank(c, count, c->bank_kind, s);
}

// Transport generations gen .. gen+nb-1 (shard [begin, begin+count) of each) in one launch.  nb > 1 only for the
// uniform source, where generations are independent; each generation scores into its own G tally rows.
int run_transport(nraps_mc_ctx *c, uint64_t gen, uint64_t begin, uint64_t count, bool trace, cudaStream_t s, uint32_t nb = 1)
{
    if (begin > c->histories || count > c->histories - begin) return NRAPS_ERR_SHAPE;
    if (nb < 1 || nb > c->batch || gen + nb > c->generations || (nb > 1 && (trace || c->d_tally != c->d_tally_own))) return NRAPS_ERR_STATE;
    c->last_shard = count;
    auto &LN = c->lanes[c->lane]; // the scratch set of this launch (nraps_mc_select_lane)
    if (c->bank_mode) {
        int rc = ensure_bank(c, count, s);
        if (rc != NRAPS_OK) return rc;
        const uint64_t padded = (count + kBankTile - 1) / kBankTile * kBankTile;
        if (padded) CU(cudaMemsetAsync(c->d_counts, 0, padded, s));
    }
    // bank mode: the cell histogram of the generation's bank rides behind the counters (one all-reduce for all of it)
    const uint64_t words = (uint64_t)nb * c->G * c->N + NRAPS_CT_WORDS + (c->bank_mode ? c->N : 0u);
    CU(cudaMemsetAsync(c->d_tally, 0, words * sizeof(unsigned long long), s));
    CU(cudaMemsetAsync(LN.d_work, 0, sizeof(unsigned long long), s));
    if (count == 0) return NRAPS_OK;
    const bool surface_fused = !c->woodcock && c->opt.kernel_variant == NRAPS_KERNEL_FUSED;
    if (surface_fused) CU(cudaMemsetAsync(LN.d_diff, 0, (uint64_t)nb * c->G * c->N * sizeof(unsigned long long), s));

    TransportParams P{};
    P.segw = c->d_segw; P.diff = LN.d_diff; P.surf_mode = c->surf_mode; P.skip_walk = c->skip_walk; P.length = c->length;
    // closed-form strides while |ds| exceeds the first power of two >= stride_min cell widths; swept on config 4 with that
    // rounding: 3 -> 4.73e8, 4 -> 4.69e8, 6 -> 4.88e8, 8 / 10 / 12 -> 4.77e8 (without it 6 -> 4.76e8, 12 -> 4.89e8: luck of
    // where 6 or 12 widths of this mesh fall inside a binade)
    P.stride_min = c->opt.walk_cap > 0 ? (float)c->opt.walk_cap : 6.0f;
    P.edges = c->d_edges; P.runb = c->d_runb; P.matid = c->d_matid; P.fuel = c->d_fuel; P.xs = c->d_xs; P.jump = c->d_jump;
    P.M = c->M; P.G = c->G; P.N = c->N; P.NF = c->NF; P.NB = c->NB; P.bucket = c->d_bucket; P.inv_h = c->inv_h; P.big = c->big;
    P.boundl = c->boundl; P.boundr = c->boundr; P.dx_fuel = c->dx_fuel;
    // history (gen, y) owns the stream position (gen*H + y)*stride; the kernel adds the y part
    uint64_t jm, jp;
    pcg_jump_coeffs(c->master.inc, gen * c->histories * c->stride, &jm, &jp);
    P.rng_state = jm * c->master.state + jp;
    P.rng_inc = c->master.inc;
    P.hist_begin = begin; P.hist_end = begin + (uint64_t)nb * count;
    P.rows = nb * c->G; P.hist_shard = count; P.hist_total = c->histories;
    if (surface_fused) {
        const SurfLayout SL = make_surface_layout(c->M, c->G, c->N, c->surf_mode, P.rows);
        P.diff_hi_off = SL.diff_hi - SL.diff_lo; P.direct_hi_off = SL.direct_hi - SL.direct_lo;
    }
    P.work = LN.d_work; P.tally = c->d_tally;
    P.trace = trace ? c->d_trace : nullptr;
    P.chunk = c->chunk; P.max_flights = c->max_flights;
    P.spawn_batch = c->opt.spawn_batch > 0 ? (uint32_t)std::min(c->opt.spawn_batch, 32) : (c->woodcock ? 2u : 1u);
    P.walk_cap = 0x7fffffffu; // round 1's regrouping cap; the surface kernel walks to the end of the segment now
    P.scatter_mode = c->opt.scatter_mode; P.stale_xs = c->opt.stale_xs;
    P.n_peers = (c->bank_mode && c->bank_src >= 0) ? (uint32_t)c->bank_world : 0u;
    for (uint32_t r = 0; r < P.n_peers; ++r) P.peer_bank[r] = c->peer_bank[c->bank_src][r];
    P.slots = c->d_slots; P.counts = c->d_counts; P.k_cur = c->d_k_cur; P.bank_cap = c->bank_cap;
    // The fused kernels take a shard in sub-shards of at most `sub` histories: births of a sub-shard (32 bytes per
    // history) -> its transport -> the next one, all into the same tallies (integer sums) with the streams still keyed by
    // the global history index, so the split changes nothing but the size of the birth-record buffer (4 GiB at most
    // instead of 32 bytes x histories; NRAPS_SUBSHARD overrides the size for tests).
    uint64_t sub = 1ull << 27;
    if (const char *e = std::getenv("NRAPS_SUBSHARD")) sub = std::max<uint64_t>(1, std::strtoull(e, nullptr, 10));
    if (nb > 1 || c->opt.kernel_variant != NRAPS_KERNEL_FUSED) sub = (uint64_t)nb * count; // batched generations / other variants: one piece
    if (sub > 0xffffffffull) { // the fused kernels index a launch's histories with 32 bits
        if (nb > 1 || c->opt.kernel_variant != NRAPS_KERNEL_FUSED) return NRAPS_ERR_TOO_LARGE;
        sub = 1ull << 31;
    }
    if (c->opt.kernel_variant != NRAPS_KERNEL_EVENT) { // births first, every lane busy; the transport lanes adopt them
        const uint64_t births = std::min<uint64_t>((uint64_t)nb * count, sub);
        if (births > LN.source_cap) {
            CU(dev_free(LN.d_source));
            LN.d_source = nullptr; LN.source_cap = 0;
            CU(dev_malloc((void **)&LN.d_source, births * 2 * sizeof(uint4)));
            LN.source_cap = births;
        }
        P.source = LN.d_source;
    }
    auto births_of = [&](TransportParams &Q) -> int {
        phase_begin(c, NRAPS_PH_SOURCE, s);
        CU(launch_source(Q, c->bank_mode, LN.d_source, s));
        phase_end(c, NRAPS_PH_SOURCE, s);
        return NRAPS_OK;
    };
    if (c->opt.kernel_variant == NRAPS_KERNEL_EVENT) {
        if (trace) return NRAPS_ERR_OPTION;
        int rc = ensure_event_bank(c, count);
        if (rc != NRAPS_OK) return rc;
        CU(run_event_generation(P, c->ev, c->smem_total, c->sm_count, s, &c->ev_iterations)); // synchronous: host-driven loop
        return NRAPS_OK;
    }
#ifdef NRAPS_WITH_BLOCK_EVENT
    if (c->opt.kernel_variant == NRAPS_KERNEL_BLOCK_EVENT) {
        if (trace || nb != 1) return NRAPS_ERR_OPTION;
        P.spawn_batch = c->opt.spawn_batch > 0 ? (uint32_t)c->opt.spawn_batch : 0u; // walk-class threshold, 0 = by run length
        if (count >> 32) return NRAPS_ERR_TOO_LARGE; // the block's source cursor is 32-bit
        { int rc = births_of(P); if (rc != NRAPS_OK) return rc; }
        CU(launch_block_event(P, dim3(c->bev_grid), dim3(c->bev_block), c->bev_smem, c->bev_slots, s));
        return NRAPS_OK;
    }
#endif
    const int ti = trace ? 1 : 0;
    if (!(c->prepared & (1u << ti))) { // once per (kernel, trace) instantiation: shared-memory opt-in and launch geometry
        if (!c->big)
            // the attribute belongs to the kernel instantiation, not to this context: always opt in to the sm_100
            // maximum, so that a second live context with a smaller image cannot lower it under this one
            CU(c->woodcock ? prepare_woodcock(kMaxSmem, c->G, trace, c->bank_mode)
                           : prepare_transport(kMaxSmem, c->G, c->surf_mode, trace, c->bank_mode));
        // auto geometry: 2 x 640 threads per SM (40 warps) when the instantiation's registers allow it, else 2 x 576 or 2 x 512;
        // ncu: the kernels are issue bound and the extra warps buy ~3 % (gpurun sweep, profiles/r1_sweeps.txt)
        uint32_t block = c->block, bps = c->blocks_per_sm;
        if (c->opt.threads_per_block <= 0 && c->opt.blocks_per_sm <= 0 && bps == 2) {
            auto occ = [&](int b) {
                return c->woodcock ? occupancy_woodcock(c->G, c->big, trace, c->bank_mode, b, c->smem_total)
                                   : occupancy_transport(c->G, c->surf_mode, trace, c->bank_mode, b, c->smem_total);
            };
            block = occ(640) >= 2 ? 640u : (occ(576) >= 2 ? 576u : 512u);
        }
        c->geo_block[ti] = block;
        c->geo_grid[ti] = (uint32_t)c->sm_count * bps;
        c->prepared |= 1u << ti;
    }
    const uint64_t total = (uint64_t)nb * count;
    for (uint64_t off = 0; off < total; off += sub) {
        TransportParams Q = P;
        const uint64_t n = std::min<uint64_t>(sub, total - off);
        if (sub < total) { // a sub-shard of one generation (nb == 1): everything indexed by history moves along
            Q.hist_begin = begin + off; Q.hist_end = begin + off + n; Q.hist_shard = n;
            Q.slots = P.slots ? P.slots + off * c->bank_cap : nullptr;
            Q.counts = P.counts ? P.counts + off : nullptr;
            Q.trace = P.trace ? P.trace + off * NRAPS_TR_WORDS : nullptr;
            if (off) CU(cudaMemsetAsync(LN.d_work, 0, sizeof(unsigned long long), s));
        }
        { int rc = births_of(Q); if (rc != NRAPS_OK) return rc; }
        phase_begin(c, NRAPS_PH_TRANSPORT, s);
        if (c->woodcock) CU(launch_woodcock(Q, trace, c->bank_mode, dim3(c->geo_grid[ti]), dim3(c->geo_block[ti]), c->smem_total, s));
        else CU(launch_transport(Q, trace, c->bank_mode, dim3(c->geo_grid[ti]), dim3(c->geo_block[ti]), c->smem_total, s));
        phase_end(c, NRAPS_PH_TRANSPORT, s);
    }
    if (!c->woodcock) {
        // the cells a flight crossed completely were booked as range updates: fold their prefix sums into the tally
        phase_begin(c, NRAPS_PH_PREFIX, s);
        CU(launch_tally_prefix(LN.d_diff, c->d_tally, nb * c->G, c->N, s));
        phase_end(c, NRAPS_PH_PREFIX, s);
    }
    return NRAPS_OK;
}

} // namespace

extern "C" int nraps_mc_trim(int32_t device)
{
    CU(cudaSetDevice(device));
    cudaMemPool_t pool = nullptr;
    CU(device_pool(&pool));
    CU(cudaDeviceSynchronize());
    CU(cudaMemPoolTrimTo(pool, 0));
    return NRAPS_OK;
}

extern "C" int nraps_abi_version(void) { return NRAPS_ABI_VERSION; }

extern "C" void nraps_options_default(nraps_options *o)
{
    if (!o) return;
    std::memset(o, 0, sizeof(*o));
    o->stale_xs = 1;
}
extern "C" const char *nraps_last_cuda_error(void) { return g_cuda_error.c_str(); }

extern "C" const char *nraps_strerror(int code)
{
    switch (code) {
    case NRAPS_OK: return "ok";
    case NRAPS_ERR_NULL: return "required pointer is NULL";
    case NRAPS_ERR_SHAPE: return "problem dimensions out of range (need M in 1..64, G in 2..8, N in 1..65535, skip < generations)";
    case NRAPS_ERR_MESH: return "mesh inconsistent (right[i] != left[i+1], matid >= M or fuel index >= N)";
    case NRAPS_ERR_XS: return "inv_sigtr must be finite and positive for every material present in the mesh";
    case NRAPS_ERR_TOO_LARGE: return "tables exceed the 227 KB shared-memory budget of one SM";
    case NRAPS_ERR_CUDA: return "CUDA error (see nraps_last_cuda_error)";
    case NRAPS_ERR_OPTION: return "unsupported value in nraps_options";
    case NRAPS_ERR_STATE: return "call order violated";
    case NRAPS_ERR_IO: return "file error";
    default: return "unknown error";
    }
}

namespace {

// allow_batch: the caller runs whole generations back to back (nraps_mc_run) and lets one launch carry several
int create_ctx(const nraps_problem *p, const nraps_options *o, nraps_mc_ctx **out, bool allow_batch)
{
    if (!out) return NRAPS_ERR_NULL;
    *out = nullptr;
    int rc = validate(p, o);
    if (rc != NRAPS_OK) return rc;

    const uint32_t M = p->M, G = p->G, N = p->N, NF = p->NF, MG = M * G;
    // Woodcock: position buckets no wider than the narrowest cell, so a bucket overlaps at most two cells
    const bool woodcock = (o->tracking_mode == NRAPS_TRACK_WOODCOCK);
    uint32_t NB = 0;
    if (woodcock) {
        float min_dx = p->right[0] - p->left[0];
        for (uint32_t i = 0; i < N; ++i) min_dx = std::min(min_dx, p->right[i] - p->left[i]);
        NB = (uint32_t)std::min(16384.0, std::max(1.0, std::ceil((double)p->right[N - 1] / (double)min_dx)));
    }
    // Shared-memory image and where the tallies live.  Surface tracking with the fused kernel: SurfLayout (direct bins +
    // difference array in shared memory when two blocks still fit an SM, the difference array alone when not, both in
    // global memory when even that exceeds one SM).  Woodcock and the event variants: SmemLayout; a mesh too large for
    // one SM runs in BIG mode there (tables through L1/L2, tally in global memory).
    const bool surface_fused = !woodcock && o->kernel_variant == NRAPS_KERNEL_FUSED;
    SmemLayout L = make_layout(M, G, N, NF, NB, 0);
    uint32_t big = 0, surf_mode = SURF_SPLIT, smem_total = 0;
    auto surf_total = [&](uint32_t mode, uint32_t rows) { return make_surface_layout(M, G, N, mode, rows).total; };
    if (surface_fused) {
        if (2ull * (surf_total(SURF_SPLIT, G) + 1024) <= 233472ull) surf_mode = SURF_SPLIT;
        else if (surf_total(SURF_UNIFIED, G) <= kMaxSmem) surf_mode = SURF_UNIFIED;
        else surf_mode = SURF_GLOBAL;
        big = surf_mode == SURF_GLOBAL;
        smem_total = surf_total(surf_mode, G);
        if (smem_total > kMaxSmem) return NRAPS_ERR_TOO_LARGE;
    } else {
        big = L.total > kMaxSmem ? 1u : 0u;
        if (big) L = make_layout(M, G, N, NF, NB, 1);
        if (L.total > kMaxSmem || (big && o->kernel_variant != NRAPS_KERNEL_FUSED)) return NRAPS_ERR_TOO_LARGE;
        smem_total = L.total;
    }
    // Generations of the uniform source are independent, and one of the shipped decks' 1e5..1e6 histories leaves most
    // of the 148 SMs idle: let a launch carry enough generations for ~2^23 histories, each scoring into its own G
    // tally rows, as long as the block still fits twice on an SM.
    uint32_t batch = 1;
    if (allow_batch && !big && o->source_mode == NRAPS_SOURCE_UNIFORM_FUEL && o->kernel_variant == NRAPS_KERNEL_FUSED) {
        uint64_t want = (1ull << 23) / p->histories;
        // Larger generations fill the GPU, but every launch of the persistent kernel ends in a tail -- the last histories
        // to start finish alone, ~0.6 ms on config 3 whatever the size of the launch (profiles/r3_size_scan.txt).  A launch
        // can carry NRAPS_TAIL_BATCH generations up to 2^25 histories so that the tail is paid once for them (measured:
        // three per launch save 2 % on config 3 but need a birth buffer three times the size); by default the launches
        // of large generations are pipelined on two streams instead (nraps_mc_run).
        uint64_t tail_batch = 1; // (3 was the default until the launches were pipelined instead, see nraps_mc_run: no 1 GB birth buffer)
        if (const char *e = std::getenv("NRAPS_TAIL_BATCH")) tail_batch = std::max<uint64_t>(1, std::strtoull(e, nullptr, 10));
        if (want < tail_batch && p->histories * tail_batch <= (1ull << 25)) want = tail_batch;
        want = std::min<uint64_t>(std::min<uint64_t>(want, p->generations), 64);
        const uint32_t budget = 100u * 1024u;
        if (surface_fused) {
            // the split image doubles the bins per generation: keep it if the whole batch still fits, else the unified one
            uint32_t bm = surf_mode;
            if (want > 1 && bm == SURF_SPLIT && surf_total(SURF_SPLIT, (uint32_t)want * G) > budget) bm = SURF_UNIFIED;
            while (want > 1 && surf_total(bm, (uint32_t)want * G) > budget) --want;
            if (want > 1) { batch = (uint32_t)want; surf_mode = bm; smem_total = surf_total(bm, batch * G); }
        } else {
            while (want > 1 && make_layout(M, G, N, NF, NB, 0, (uint32_t)want * G).total > budget) --want;
            if (want > 1) {
                batch = (uint32_t)want;
                L = make_layout(M, G, N, NF, NB, 0, batch * G);
                smem_total = L.total;
            }
        }
    }

    int ndev = 0;
    CU(cudaGetDeviceCount(&ndev));
    if (o->device < 0 || o->device >= ndev) return cuda_fail(cudaErrorInvalidDevice, "nraps_options.device");
    CU(cudaSetDevice(o->device));
    int cc_major = 0, sm_count = 0; // two attribute reads, not cudaGetDeviceProperties (tens of ms per call)
    CU(cudaDeviceGetAttribute(&cc_major, cudaDevAttrComputeCapabilityMajor, o->device));
    CU(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, o->device));
    if (cc_major != 10) return cuda_fail(cudaErrorNoKernelImageForDevice, "this library is built for sm_100a only");

    nraps_mc_ctx *c = new nraps_mc_ctx();
    c->opt = *o;
    c->M = M; c->G = G; c->N = N; c->NF = NF; c->numass = p->numass;
    c->generations = p->generations; c->histories = p->histories; c->skip = p->skip;
    c->boundl = p->boundl; c->boundr = p->boundr; c->dx_fuel = p->dx_fuel;
    c->length = p->right[N - 1]; c->nut_m1 = p->nut[0 + M * 1]; c->k0 = p->k0;
    c->device = o->device; c->sm_count = sm_count;
    const bool dflt = (o->seed == 0 && o->stream == 0 && o->stride == 0);
    c->master = pcg_seed(dflt ? 42u : o->seed, dflt ? 54u : o->stream);
    c->stride = dflt ? 152917u : o->stride;
    c->layout = L; c->batch = batch; c->smem_total = smem_total; c->surf_mode = surf_mode;
    c->max_flights = (uint32_t)std::min<uint64_t>(o->max_flights ? o->max_flights : (1ull << 24), 0xffffffffull);
    // the event pipeline packs the flight count of a record into 20 bits (mc_event.cu): a cap it cannot count to would
    // never fire and the host-driven round loop would spin on a runaway history
    if (o->kernel_variant == NRAPS_KERNEL_EVENT) c->max_flights = std::min<uint32_t>(c->max_flights, 0xfffffu); // validate(): <= 2^20-1 if given
    c->chunk = o->chunk > 0 ? (uint32_t)o->chunk : 32u; // histories a warp claims per global atomic (swept on config 3: 16-32 best, 64 -0.4 %, 256 -3 %)
    c->bank_mode = (o->source_mode == NRAPS_SOURCE_FISSION_BANK);
    c->bank_cap = o->bank_cap > 0 ? (uint32_t)o->bank_cap : 8u;

    // launch geometry: persistent grid, a multiple of the SM count
    uint32_t bps = o->blocks_per_sm > 0 ? (uint32_t)o->blocks_per_sm : 2u;
    while (bps > 1 && (uint64_t)bps * (smem_total + 1024) > 233472ull) --bps;
    uint32_t threads = o->threads_per_block > 0 ? (uint32_t)o->threads_per_block : 1024u / bps;
    threads = std::max(32u, std::min(1024u, threads / 32u * 32u));
    c->blocks_per_sm = bps; c->block = threads; c->grid = (uint32_t)c->sm_count * bps;
#ifdef NRAPS_WITH_BLOCK_EVENT
    if (o->kernel_variant == NRAPS_KERNEL_BLOCK_EVENT) {
        // 2 x 512 threads per SM with 3 neutrons banked per thread unless told otherwise; shrink the bank, then the
        // residency, until the mesh image + the bank of every resident block fit the 227 KB of one SM
        uint32_t bb = o->threads_per_block > 0 ? threads : 512u, bbps = o->blocks_per_sm > 0 ? bps : 2u;
        uint32_t spt = o->slots_per_thread > 0 ? (uint32_t)o->slots_per_thread : 3u;
        TransportParams q{};
        q.M = M; q.G = G; q.N = N; q.NF = NF; q.NB = NB; q.rows = G;
        while (bb * spt > 65535u) --spt; // list lengths are the 16-bit halves of a packed word
        while ((uint64_t)bbps * (block_event_smem(q, bb * spt) + 1024) > 233472ull) {
            if (spt > 1) --spt;
            else if (bbps > 1) { --bbps; spt = o->slots_per_thread > 0 ? (uint32_t)o->slots_per_thread : 3u; }
            else { delete c; return NRAPS_ERR_TOO_LARGE; }
        }
        c->bev_block = bb; c->bev_grid = (uint32_t)c->sm_count * bbps; c->bev_slots = bb * spt;
        c->bev_smem = block_event_smem(q, c->bev_slots);
        if (o->chunk <= 0) c->chunk = 256u; // histories a block claims per global atomic
    }
#endif

    // ---- derived tables
    std::vector<float> edges(N + 1);
    std::vector<uint32_t> runb(N);
    std::vector<uint8_t> matid(p->matid, p->matid + N);
    std::vector<uint16_t> fuel(NF);
    for (uint32_t i = 0; i < N; ++i) edges[i] = p->left[i];
    edges[N] = p->right[N - 1];
    for (uint32_t i = 0; i < N;) {
        uint32_t j = i;
        while (j < N && p->matid[j] == p->matid[i]) ++j;
        for (uint32_t q = i; q < j; ++q) runb[q] = i | (j << 16);
        i = j;
    }
    for (uint32_t j = 0; j < NF; ++j) fuel[j] = (uint16_t)p->fuel_indices[j];
    // Segments of the surface kernel: maximal ranges of cells of one material run whose widths e[i+1] - e[i] are the same
    // binary32 number.  Edges accumulate in f32 upstream (src/main.rs:119-140), so the width of a fuel or water cell
    // changes by an ulp where the position crosses a power of two: a material run is one segment, or two around such a
    // point.  Inside a segment x - edge is the same number for every cell crossed completely (mc_transport.cu).
    std::vector<uint2> segw(N);
    uint32_t n_segments = 0;
    {
        std::vector<uint32_t> stops(N), wbits(N);
        nraps_walk_segments(p->matid, p->left, p->right, N, stops.data(), wbits.data(), &n_segments); // host_mesh.cpp
        for (uint32_t i = 0; i < N; ++i) segw[i] = make_uint2(stops[i], wbits[i]);
    }
    c->skip_walk = (o->walk_cap != -2 && N >= 16u * n_segments) ? 1u : 0u; // walk_cap = -2: never stride (tests, comparisons)
    std::vector<float> xs(xs_floats(M, G));
    float *inv_sigtr = xs.data(), *p_abs = inv_sigtr + MG, *chi_cdf = p_abs + MG, *nusigf = chi_cdf + MG, *sigtr = nusigf + MG,
          *scat_cdf = sigtr + MG, *inv_maj = scat_cdf + MG * G * G;
    for (uint32_t i = 0; i < MG; ++i) {
        inv_sigtr[i] = p->inv_sigtr[i];
        p_abs[i] = p->siga[i] / p->sigt[i];
        nusigf[i] = p->nut[i] * p->sigf[i];
        const float prod = p->mu[i] * p->sigs[i];
        sigtr[i] = p->sigt[i] - prod; // the expression inside inv_sigtr, src/process_input.rs:152-156
    }
    {   // majorant of every (stale group, current group) pair over the materials present in the mesh
        std::vector<char> present(M, 0);
        for (uint32_t i = 0; i < N; ++i) present[p->matid[i]] = 1;
        for (uint32_t ga = 0; ga < G; ++ga)
            for (uint32_t gb = 0; gb < G; ++gb) {
                float mx = 0.0f;
                for (uint32_t m = 0; m < M; ++m) {
                    if (!present[m]) continue;
                    mx = std::max(mx, std::max(sigtr[m + M * ga], sigtr[m + M * gb]));
                }
                inv_maj[ga * G + gb] = 1.0f / mx;
            }
    }
    std::vector<uint16_t> bucket(NB);
    for (uint32_t b = 0; b < NB; ++b) {
        const double xb = (double)b * (double)p->right[N - 1] / (double)NB;
        uint32_t cidx = 0;
        while (cidx + 1 < N && (double)p->right[cidx] <= xb) ++cidx;
        bucket[b] = (uint16_t)cidx;
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
    std::vector<ulonglong2> jump(64);
    {
        uint64_t a, cc;
        pcg_jump_coeffs(c->master.inc, c->stride, &a, &cc);
        for (int b = 0; b < 64; ++b) {
            jump[b].x = a; jump[b].y = cc;
            cc = (a + 1u) * cc;
            a = a * a;
        }
    }
    std::vector<float> dx(p->dx, p->dx + N), nut(p->nut, p->nut + MG), sigf(p->sigf, p->sigf + MG);

    const uint64_t GN = (uint64_t)G * N;
    cudaError_t e = cudaSuccess;
    auto ok = [&](cudaError_t r) { if (e == cudaSuccess) e = r; return r == cudaSuccess; };
    ok(upload(&c->d_edges, edges)); ok(upload(&c->d_runb, runb)); ok(upload(&c->d_matid, matid)); ok(upload(&c->d_segw, segw));
    c->diff_words = batch * GN;
    ok(dev_malloc((void **)&c->lanes[0].d_diff, c->diff_words * sizeof(unsigned long long)));
    ok(upload(&c->d_fuel, fuel)); ok(upload(&c->d_xs, xs)); ok(upload(&c->d_jump, jump)); ok(upload(&c->d_bucket, bucket));
    ok(upload(&c->d_dx, dx)); ok(upload(&c->d_nut, nut)); ok(upload(&c->d_sigf, sigf));
    ok(dev_malloc((void **)&c->d_tally_own, (batch * GN + NRAPS_CT_WORDS + N) * sizeof(unsigned long long)));
    ok(dev_malloc((void **)&c->lanes[0].d_work, sizeof(unsigned long long)));
    ok(dev_malloc((void **)&c->d_counters_total, NRAPS_CT_WORDS * sizeof(unsigned long long)));
    ok(dev_malloc((void **)&c->d_terms, GN * sizeof(float)));
    ok(dev_malloc((void **)&c->d_res_flux, GN * sizeof(float)));
    ok(dev_malloc((void **)&c->d_res_moments, 2 * GN * sizeof(double)));
    ok(dev_malloc((void **)&c->d_res_fission, N * sizeof(float)));
    ok(dev_malloc((void **)&c->d_k_hist, c->generations * sizeof(float)));
    ok(dev_malloc((void **)&c->d_k_cur, sizeof(float)));
    ok(dev_malloc((void **)&c->d_bank_sizes, c->generations * sizeof(unsigned long long)));
    ok(dev_malloc((void **)&c->d_entropy, c->generations * sizeof(double)));
    c->NB = NB; c->woodcock = woodcock; c->big = big;
    c->inv_h = NB ? (float)((double)NB / (double)p->right[N - 1]) : 0.0f;
    if (e != cudaSuccess) {
        free_ctx(c);
        return cuda_fail(e, "nraps_mc_create");
    }
    c->d_tally = c->d_tally_own;
    rc = nraps_mc_reset(c, c->k0, nullptr);
    if (rc != NRAPS_OK) { free_ctx(c); return rc; }
    if ((e = cudaDeviceSynchronize()) != cudaSuccess) {
        free_ctx(c);
        return cuda_fail(e, "nraps_mc_create");
    }
    *out = c;
    return NRAPS_OK;
}

// fold generation `gen`, whose tally is rows [slice*G, slice*G+G) of the launch that carried nb generations
int finalize_slice(nraps_mc_ctx *c, uint64_t gen, uint32_t slice, uint32_t nb, cudaStream_t s)
{
    const uint64_t GN = (uint64_t)c->G * c->N;
    FinalizeParams F{};
    F.tally = c->d_tally + slice * GN;
    F.counters = slice == 0 ? c->d_tally + nb * GN : nullptr; // the launch's counters are accounted once
    F.dx = c->d_dx; F.matid = c->d_matid; F.nusigf_nut = c->d_nut; F.sigf = c->d_sigf;
    F.res_moments = c->d_res_moments;
    F.terms = c->d_terms; F.res_flux = c->d_res_flux; F.res_fission = c->d_res_fission;
    F.k_hist = c->d_k_hist; F.k_cur = c->d_k_cur; F.counters_total = c->d_counters_total;
    F.M = c->M; F.G = c->G; F.N = c->N;
    F.histories_f32 = (float)c->histories;
    F.length = c->length; F.nut_m1 = c->nut_m1;
    // 1 / (generations - (skip - 1)) in wrapping usize arithmetic, src/mc_code.rs:340
    F.fund = 1.0f / (float)(uint64_t)(c->generations - (c->skip - 1));
    F.gen = gen; F.skip = c->skip;
    phase_begin(c, NRAPS_PH_FINALIZE, s);
    CU(launch_finalize(F, s));
    phase_end(c, NRAPS_PH_FINALIZE, s);
    return NRAPS_OK;
}

} // namespace

extern "C" int nraps_mc_create(const nraps_problem *p, const nraps_options *o, nraps_mc_ctx **out)
{
    return create_ctx(p, o, out, false);
}

extern "C" int nraps_mc_destroy(nraps_mc_ctx *ctx)
{
    if (!ctx) return NRAPS_ERR_NULL;
    cudaSetDevice(ctx->device);
    free_ctx(ctx);
    return NRAPS_OK;
}

extern "C" int nraps_mc_reset(nraps_mc_ctx *c, float k0, void *stream)
{
    if (!c) return NRAPS_ERR_NULL;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    CU(cudaSetDevice(c->device));
    const uint64_t GN = (uint64_t)c->G * c->N;
    CU(cudaMemsetAsync(c->d_res_flux, 0, GN * sizeof(float), s));
    CU(cudaMemsetAsync(c->d_res_moments, 0, 2 * GN * sizeof(double), s));
    CU(cudaMemsetAsync(c->d_res_fission, 0, c->N * sizeof(float), s));
    CU(cudaMemsetAsync(c->d_k_hist, 0, c->generations * sizeof(float), s));
    CU(cudaMemsetAsync(c->d_counters_total, 0, NRAPS_CT_WORDS * sizeof(unsigned long long), s));
    CU(cudaMemsetAsync(c->d_bank_sizes, 0, c->generations * sizeof(unsigned long long), s));
    CU(cudaMemsetAsync(c->d_entropy, 0, c->generations * sizeof(double), s));
    c->bank_src = -1; c->bank_last = -1; c->bank_which = 0;
    phase_fold(c);
    for (double &v : c->phase_ms) v = 0.0;
    CU(cudaMemcpyAsync(c->d_k_cur, &k0, sizeof(float), cudaMemcpyHostToDevice, s));
    CU(cudaStreamSynchronize(s)); // k0 lives on the caller's stack
    return NRAPS_OK;
}

extern "C" int nraps_mc_transport(nraps_mc_ctx *c, uint64_t gen, uint64_t hist_begin, uint64_t hist_count, void *stream)
{
    if (!c) return NRAPS_ERR_NULL;
    CU(cudaSetDevice(c->device));
    return run_transport(c, gen, hist_begin, hist_count, false, static_cast<cudaStream_t>(stream));
}

extern "C" int nraps_mc_finalize_generation(nraps_mc_ctx *c, uint64_t gen, void *stream)
{
    if (!c) return NRAPS_ERR_NULL;
    if (gen >= c->generations) return NRAPS_ERR_SHAPE;
    CU(cudaSetDevice(c->device));
    return finalize_slice(c, gen, 0, 1, static_cast<cudaStream_t>(stream));
}

extern "C" int nraps_mc_tally_buffer(nraps_mc_ctx *c, void **device_ptr, uint64_t *n_words)
{
    if (!c || !device_ptr || !n_words) return NRAPS_ERR_NULL;
    *device_ptr = c->d_tally;
    *n_words = (uint64_t)c->G * c->N + NRAPS_CT_WORDS + (c->bank_mode ? c->N : 0u);
    return NRAPS_OK;
}

extern "C" int nraps_mc_set_tally_buffer(nraps_mc_ctx *c, void *device_ptr)
{
    if (!c) return NRAPS_ERR_NULL;
    c->d_tally = device_ptr ? static_cast<unsigned long long *>(device_ptr) : c->d_tally_own;
    return NRAPS_OK;
}

extern "C" int nraps_mc_select_lane(nraps_mc_ctx *c, int32_t lane)
{
    if (!c) return NRAPS_ERR_NULL;
    if (lane < 0 || lane > 1) return NRAPS_ERR_OPTION;
    CU(cudaSetDevice(c->device));
    auto &ln = c->lanes[lane];
    if (!ln.d_diff) CU(dev_malloc((void **)&ln.d_diff, std::max<uint64_t>(1, c->diff_words) * sizeof(unsigned long long)));
    if (!ln.d_work) CU(dev_malloc((void **)&ln.d_work, sizeof(unsigned long long)));
    c->lane = lane;
    return NRAPS_OK;
}

extern "C" int nraps_mc_read_tally(nraps_mc_ctx *c, uint64_t *host_words, void *stream)
{
    if (!c || !host_words) return NRAPS_ERR_NULL;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    CU(cudaSetDevice(c->device));
    const uint64_t words = (uint64_t)c->G * c->N + NRAPS_CT_WORDS;
    CU(cudaMemcpyAsync(host_words, c->d_tally, words * sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    return NRAPS_OK;
}

extern "C" int nraps_mc_fetch(nraps_mc_ctx *c, nraps_results *r, void *stream)
{
    if (!c || !r || !r->flux || !r->assembly_average || !r->fission_source || !r->k || !r->k_fund) return NRAPS_ERR_NULL;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    CU(cudaSetDevice(c->device));
    const uint64_t GN = (uint64_t)c->G * c->N;
    CU(cudaMemcpyAsync(r->flux, c->d_res_flux, GN * sizeof(float), cudaMemcpyDeviceToHost, s));
    if (r->flux_moments) CU(cudaMemcpyAsync(r->flux_moments, c->d_res_moments, 2 * GN * sizeof(double), cudaMemcpyDeviceToHost, s));
    CU(cudaMemcpyAsync(r->fission_source, c->d_res_fission, c->N * sizeof(float), cudaMemcpyDeviceToHost, s));
    CU(cudaMemcpyAsync(r->k, c->d_k_hist, c->generations * sizeof(float), cudaMemcpyDeviceToHost, s));
    CU(cudaMemcpyAsync(r->counters, c->d_counters_total, NRAPS_CT_WORDS * sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
    if (r->bank_sizes) CU(cudaMemcpyAsync(r->bank_sizes, c->d_bank_sizes, c->generations * sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
    if (r->entropy) CU(cudaMemcpyAsync(r->entropy, c->d_entropy, c->generations * sizeof(double), cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    // average_assembly (src/mc_code.rs:259-274) and k_fund (:368-376): O(G*N), O(gens^2) host work
    const uint32_t span = c->N / c->numass;
    for (uint32_t g = 0; g < c->G; ++g) {
        const float *row = r->flux + (size_t)g * c->N;
        float *dst = r->assembly_average + (size_t)g * c->N;
        for (uint32_t i = 0; i < c->N; ++i) dst[i] = 0.0f;
        for (uint32_t a = 0; a < c->numass; ++a) {
            float acc = 0.0f;
            for (uint32_t i = a * span; i < (a + 1) * span; ++i) acc = acc + row[i];
            const float mean = acc / (float)span;
            for (uint32_t i = a * span; i < (a + 1) * span; ++i) dst[i] = mean;
        }
    }
    for (uint64_t n = 0; n < c->generations; ++n) r->k_fund[n] = 0.0f;
    r->k_fund[c->skip] = r->k[c->skip];
    for (uint64_t n = c->skip + 1; n < c->generations; ++n) {
        float acc = 0.0f;
        for (uint64_t j = c->skip; j <= n; ++j) acc = acc + r->k[j];
        r->k_fund[n] = acc / (float)(uint64_t)(n - (c->skip - 1));
    }
    return NRAPS_OK;
}

extern "C" int nraps_mc_trace(nraps_mc_ctx *c, uint64_t gen, uint64_t hist_begin, uint64_t hist_count,
                              uint32_t *host_records, void *stream)
{
    if (!c || !host_records) return NRAPS_ERR_NULL;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    CU(cudaSetDevice(c->device));
    if (hist_count > c->trace_cap) {
        CU(dev_free(c->d_trace));
        c->d_trace = nullptr; c->trace_cap = 0;
        CU(dev_malloc((void **)&c->d_trace, hist_count * NRAPS_TR_WORDS * sizeof(uint32_t)));
        c->trace_cap = hist_count;
    }
    if (hist_count) CU(cudaMemsetAsync(c->d_trace, 0, hist_count * NRAPS_TR_WORDS * sizeof(uint32_t), s));
    int rc = run_transport(c, gen, hist_begin, hist_count, true, s);
    if (rc != NRAPS_OK) return rc;
    if (hist_count)
        CU(cudaMemcpyAsync(host_records, c->d_trace, hist_count * NRAPS_TR_WORDS * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    return NRAPS_OK;
}

extern "C" int nraps_mc_bank_compact(nraps_mc_ctx *c, uint64_t gen, void *stream)
{
    if (!c) return NRAPS_ERR_NULL;
    if (!c->bank_mode || gen >= c->generations || !c->d_bank[c->bank_which]) return NRAPS_ERR_STATE;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    CU(cudaSetDevice(c->device));
    BankParams B{};
    const uint64_t padded = (c->last_shard + kBankTile - 1) / kBankTile * kBankTile;
    unsigned long long *buf = c->d_bank[c->bank_which];
    B.counts = c->d_counts; B.slots = c->d_slots; B.dense = buf + kBankHeader;
    B.block_sums = c->d_block_sums; B.count_out = buf; // word 0 of the buffer: where the peers read the count
    B.n_hist = c->last_shard; B.dense_cap = c->dense_cap; B.cap = c->bank_cap; B.n_tiles = (uint32_t)(padded / kBankTile);
    phase_begin(c, NRAPS_PH_COMPACT, s);
    CU(launch_bank_compact(B, s));
    // this rank's share of the bank's cell histogram, into the words behind the counters of the tally buffer
    CU(launch_bank_histogram(buf + kBankHeader, buf, c->d_tally + (uint64_t)c->G * c->N + NRAPS_CT_WORDS, c->N, s));
    phase_end(c, NRAPS_PH_COMPACT, s);
    c->bank_last = c->bank_which;
    return NRAPS_OK;
}

extern "C" int nraps_mc_bank_local(nraps_mc_ctx *c, void **device_sites, uint64_t *count, void *stream)
{
    if (!c || !device_sites || !count) return NRAPS_ERR_NULL;
    if (!c->bank_mode || c->bank_last < 0) return NRAPS_ERR_STATE;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    CU(cudaSetDevice(c->device));
    unsigned long long n = 0;
    CU(cudaMemcpyAsync(&n, c->d_bank[c->bank_last], sizeof(n), cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    *device_sites = c->d_bank[c->bank_last] + kBankHeader;
    *count = n;
    return NRAPS_OK;
}

extern "C" int nraps_mc_bank_advance(nraps_mc_ctx *c, uint64_t gen, void *stream)
{
    if (!c) return NRAPS_ERR_NULL;
    if (!c->bank_mode || gen >= c->generations || c->bank_last != c->bank_which) return NRAPS_ERR_STATE;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    CU(cudaSetDevice(c->device));
    // size and entropy of the whole bank from the histogram words (summed across ranks by the caller's all-reduce)
    CU(launch_bank_entropy(c->d_tally + (uint64_t)c->G * c->N + NRAPS_CT_WORDS, c->N, c->d_entropy + gen, c->d_bank_sizes + gen, s));
    c->bank_src = c->bank_which;
    c->bank_which ^= 1;
    return NRAPS_OK;
}

extern "C" int nraps_mc_bank_reserve(nraps_mc_ctx *c, uint64_t shard_histories, void **device_buffers)
{
    if (!c) return NRAPS_ERR_NULL;
    if (!c->bank_mode || c->bank_world > 1 || shard_histories == 0) return NRAPS_ERR_STATE;
    CU(cudaSetDevice(c->device));
    // device_buffers given: the caller will hand the pointers to the other devices of this process (peer access);
    // NULL: other processes will map the buffers through nraps_mc_bank_export / import
    const int kind = device_buffers ? 1 : 2;
    if (c->bank_kind != kind || shard_histories > c->bank_hist_cap) {
        int rc = alloc_bank(c, std::max<uint64_t>(shard_histories, c->bank_hist_cap), kind, nullptr);
        if (rc != NRAPS_OK) return rc;
    }
    if (device_buffers) { device_buffers[0] = c->d_bank[0]; device_buffers[1] = c->d_bank[1]; }
    return NRAPS_OK;
}

extern "C" int nraps_mc_bank_export(nraps_mc_ctx *c, void *handles)
{
    if (!c || !handles) return NRAPS_ERR_NULL;
    if (!c->bank_mode || c->bank_kind != 2) return NRAPS_ERR_STATE;
    CU(cudaSetDevice(c->device));
    const DriverApi &d = driver_api();
    std::memset(handles, 0, 2 * NRAPS_IPC_HANDLE_BYTES);
    for (int w = 0; w < 2; ++w) {
        if (c->bank_fd[w] < 0) {
            int fd = -1;
            const CUresult r = d.exportHandle(&fd, c->bank_vmm[w].handle, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0);
            if (r != CUDA_SUCCESS) return vmm_fail(r, "cuMemExportToShareableHandle");
            c->bank_fd[w] = fd; // stays open until the buffers are freed: the peers duplicate it from this process
        }
        BankTicket t{kTicketMagic, (int32_t)getpid(), c->bank_fd[w], 0u, (uint64_t)c->bank_vmm[w].size};
        std::memcpy(static_cast<unsigned char *>(handles) + w * NRAPS_IPC_HANDLE_BYTES, &t, sizeof(t));
    }
    return NRAPS_OK;
}

namespace {
int set_peers(nraps_mc_ctx *c, int32_t world, int32_t rank, int kind)
{
    if (world < 1 || world > kMaxPeers || rank < 0 || rank >= world) return NRAPS_ERR_OPTION;
    if (!c->bank_mode || c->bank_kind != kind || c->bank_src >= 0) return NRAPS_ERR_STATE; // before the first bank is advanced to
    return NRAPS_OK;
}
} // namespace

extern "C" int nraps_mc_bank_import(nraps_mc_ctx *c, int32_t world, int32_t rank, const void *handles)
{
    if (!c || !handles) return NRAPS_ERR_NULL;
    int rc = set_peers(c, world, rank, 2);
    if (rc != NRAPS_OK) return rc;
    CU(cudaSetDevice(c->device));
    const DriverApi &d = driver_api();
    for (int r = 0; r < world; ++r)
        for (int w = 0; w < 2; ++w) {
            if (r == rank) { c->peer_bank[w][r] = c->d_bank[w]; continue; }
            BankTicket t;
            std::memcpy(&t, static_cast<const unsigned char *>(handles) + ((size_t)r * 2 + w) * NRAPS_IPC_HANDLE_BYTES, sizeof(t));
            if (t.magic != kTicketMagic) return NRAPS_ERR_OPTION;
            // duplicate the exporter's file descriptor into this process (Linux >= 5.6; same user, as under torchrun)
            const int pidfd = (int)syscall(SYS_pidfd_open, (pid_t)t.pid, 0);
            if (pidfd < 0) return cuda_fail(cudaErrorOperatingSystem, "pidfd_open on the exporting rank");
            const int fd = (int)syscall(SYS_pidfd_getfd, pidfd, t.fd, 0);
            close(pidfd);
            if (fd < 0) return cuda_fail(cudaErrorOperatingSystem, "pidfd_getfd of the exported bank buffer");
            VmmBuffer &b = c->peer_vmm[w][r];
            b.size = (size_t)t.size;
            const CUresult cr = d.importHandle(&b.handle, (void *)(uintptr_t)fd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR);
            close(fd);
            if (cr != CUDA_SUCCESS) { b = VmmBuffer{}; return vmm_fail(cr, "cuMemImportFromShareableHandle"); }
            if ((rc = vmm_map(b, c->device)) != NRAPS_OK) { vmm_release(b); return rc; }
            c->peer_bank[w][r] = reinterpret_cast<const unsigned long long *>(b.ptr);
        }
    c->bank_world = world; c->bank_rank = rank;
    return NRAPS_OK;
}

extern "C" int nraps_mc_bank_peers(nraps_mc_ctx *c, int32_t world, int32_t rank, const void *const *device_buffers)
{
    if (!c || !device_buffers) return NRAPS_ERR_NULL;
    int rc = set_peers(c, world, rank, 1);
    if (rc != NRAPS_OK) return rc;
    for (int r = 0; r < world; ++r)
        for (int w = 0; w < 2; ++w)
            c->peer_bank[w][r] = r == rank ? c->d_bank[w] : static_cast<const unsigned long long *>(device_buffers[(size_t)r * 2 + w]);
    c->bank_world = world; c->bank_rank = rank;
    return NRAPS_OK;
}

extern "C" int nraps_mc_phase_ms(nraps_mc_ctx *c, double out[NRAPS_PH_WORDS])
{
    if (!c || !out) return NRAPS_ERR_NULL;
    CU(cudaSetDevice(c->device));
    phase_fold(c);
    for (int p = 0; p < NRAPS_PH_WORDS; ++p) out[p] = c->phase_ms[p];
    return NRAPS_OK;
}

extern "C" int nraps_mc_launch_info(nraps_mc_ctx *c, uint32_t out[6])
{
    if (!c || !out) return NRAPS_ERR_NULL;
    out[0] = c->geo_grid[0] ? c->geo_grid[0] : c->grid; out[1] = c->geo_block[0] ? c->geo_block[0] : c->block; out[2] = c->smem_total; out[3] = c->blocks_per_sm;
    if (c->opt.kernel_variant == NRAPS_KERNEL_BLOCK_EVENT) {
        out[0] = c->bev_grid; out[1] = c->bev_block; out[2] = c->bev_smem; out[3] = c->bev_grid / (uint32_t)c->sm_count;
    }
    out[4] = (uint32_t)c->sm_count; out[5] = c->opt.kernel_variant == NRAPS_KERNEL_EVENT ? c->ev_iterations : c->chunk;
    return NRAPS_OK;
}

extern "C" int nraps_mc_run(const nraps_problem *p, const nraps_options *o, nraps_results *r)
{
    if (!p || !o || !r) return NRAPS_ERR_NULL;
    if (!r->flux || !r->assembly_average || !r->fission_source || !r->k || !r->k_fund) return NRAPS_ERR_NULL;
    // NRAPS_TIMING=1: host wall-clock split of this call on stderr (where a cold process spends its time)
    const bool timing = std::getenv("NRAPS_TIMING") != nullptr;
    const auto w0 = std::chrono::steady_clock::now();
    auto since = [&](std::chrono::steady_clock::time_point t) { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t).count(); };
    if (timing) cudaFree(nullptr); // bring the CUDA context up on its own line of the split
    const double ms_context = since(w0);
    nraps_mc_ctx *c = nullptr;
    int rc = create_ctx(p, o, &c, true);
    if (rc != NRAPS_OK) return rc;
    const double ms_create = since(w0) - ms_context;
    if (!o->quiet) { std::printf("running MC code\n"); std::fflush(stdout); } // src/mc_code.rs:292

    cudaEvent_t e0 = nullptr, e1 = nullptr, fin[2] = {nullptr, nullptr};
    cudaStream_t s = nullptr, s2 = nullptr;
    unsigned long long *tally2 = nullptr;
    auto bail = [&](int code) {
        if (s2) cudaStreamSynchronize(s2);
        if (e0) cudaEventDestroy(e0);
        if (e1) cudaEventDestroy(e1);
        for (cudaEvent_t f : fin) if (f) cudaEventDestroy(f);
        if (s) cudaStreamDestroy(s);
        if (s2) cudaStreamDestroy(s2);
        c->d_tally = c->d_tally_own; // tally2 is not the context's to free
        dev_free(tally2);
        nraps_mc_destroy(c);
        return code;
    };
    if (cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) != cudaSuccess || cudaEventCreate(&e0) != cudaSuccess ||
        cudaEventCreate(&e1) != cudaSuccess)
        return bail(cuda_fail(cudaGetLastError(), "stream/event create"));
    cudaEventRecord(e0, s);
    const uint64_t GN = (uint64_t)p->G * p->N;
    // Uniform source, one generation per launch: the generations are independent, so their launches alternate between two
    // streams and the context's two scratch lanes (into two tally buffers).  The launch of g+1 is queued on the device
    // while the tail of g runs -- the last neutrons of a persistent launch finish one by one -- and its blocks move in as
    // blocks of g retire.  Finalizes stay in generation order (an event chain): k and the running
    // flux sums are sequential f32 accumulations.  NRAPS_PIPELINE=0: everything on one stream.
    const char *pipe_env = std::getenv("NRAPS_PIPELINE");
    const bool pipelined = c->batch == 1 && !c->bank_mode && !r->tally_fixed && !o->profile_phases && p->generations > 1 &&
                           c->opt.kernel_variant == NRAPS_KERNEL_FUSED && !(pipe_env && pipe_env[0] == '0');
    if (pipelined) {
        if (cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&fin[0], cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&fin[1], cudaEventDisableTiming) != cudaSuccess)
            return bail(cuda_fail(cudaGetLastError(), "stream/event create"));
        if (dev_malloc((void **)&tally2, (GN + NRAPS_CT_WORDS + p->N) * sizeof(unsigned long long)) != cudaSuccess)
            return bail(cuda_fail(cudaGetLastError(), "second tally buffer"));
        for (uint64_t gen = 0; gen < p->generations; ++gen) {
            const int lane = (int)(gen & 1u);
            cudaStream_t st = lane ? s2 : s;
            if ((rc = nraps_mc_select_lane(c, lane)) != NRAPS_OK) return bail(rc);
            c->d_tally = lane ? tally2 : c->d_tally_own;
            if ((rc = run_transport(c, gen, 0, p->histories, false, st, 1)) != NRAPS_OK) return bail(rc);
            if (gen && cudaStreamWaitEvent(st, fin[lane ^ 1], 0) != cudaSuccess) return bail(cuda_fail(cudaGetLastError(), "finalize order"));
            if ((rc = finalize_slice(c, gen, 0, 1, st)) != NRAPS_OK) return bail(rc);
            if (cudaEventRecord(fin[lane], st) != cudaSuccess) return bail(cuda_fail(cudaGetLastError(), "finalize order"));
        }
        // the last finalize is after everything before it on both streams (each follows its transport and the finalize before it)
        if (cudaStreamWaitEvent(s, fin[(p->generations - 1) & 1u], 0) != cudaSuccess) return bail(cuda_fail(cudaGetLastError(), "join"));
        c->lane = 0;
        c->d_tally = c->d_tally_own;
    }
    for (uint64_t gen = pipelined ? p->generations : 0; gen < p->generations;) {
        const uint32_t nb = (uint32_t)std::min<uint64_t>(c->batch, p->generations - gen);
        if ((rc = run_transport(c, gen, 0, p->histories, false, s, nb)) != NRAPS_OK) return bail(rc);
        if (r->tally_fixed) {
            if (cudaMemcpyAsync(r->tally_fixed + gen * GN, c->d_tally, nb * GN * sizeof(uint64_t), cudaMemcpyDeviceToHost, s) != cudaSuccess ||
                cudaStreamSynchronize(s) != cudaSuccess)
                return bail(cuda_fail(cudaGetLastError(), "tally read-back"));
        }
        // generations fold in order: k and the running flux sums are sequential f32 accumulations
        for (uint32_t j = 0; j < nb; ++j)
            if ((rc = finalize_slice(c, gen + j, j, nb, s)) != NRAPS_OK) return bail(rc);
        gen += nb;
        if (c->bank_mode) {
            if ((rc = nraps_mc_bank_compact(c, gen - 1, s)) != NRAPS_OK) return bail(rc);
            if ((rc = nraps_mc_bank_advance(c, gen - 1, s)) != NRAPS_OK) return bail(rc);
        }
    }
    cudaEventRecord(e1, s);
    const double ms_enqueue = since(w0) - ms_context - ms_create;
    if ((rc = nraps_mc_fetch(c, r, s)) != NRAPS_OK) return bail(rc);
    const double ms_fetch = since(w0) - ms_context - ms_create - ms_enqueue;
    float ms = 0.0f;
    cudaEventElapsedTime(&ms, e0, e1);
    r->seconds_device = 1e-3 * (double)ms;
    const auto w1 = std::chrono::steady_clock::now();
    const uint32_t batch_used = c->batch; // generations per launch (the context is gone after bail)
    rc = bail(NRAPS_OK);
    if (timing)
        std::fprintf(stderr, "{\"nraps_mc_run_ms\": {\"cuda_context\": %.1f, \"create_tables_buffers\": %.1f, \"enqueue_generations\": %.1f, "
                             "\"wait_and_fetch\": %.1f, \"destroy\": %.1f, \"device_generations\": %.1f, \"batch\": %u}}\n",
                     ms_context, ms_create, ms_enqueue, ms_fetch, since(w1), (double)ms, batch_used);
    return rc;
}

// device scratch of the unit probes below: released on every return path
namespace {
struct Scratch {
    void *p = nullptr;
    ~Scratch() { dev_free(p); }
    cudaError_t alloc(size_t bytes) { return dev_malloc(&p, std::max<size_t>(1, bytes)); }
    template <typename T> T *as() const { return static_cast<T *>(p); }
};
} // namespace

extern "C" int nraps_dev_logf(const float *x, float *out, uint32_t n, int32_t device)
{
    if (!x || !out) return NRAPS_ERR_NULL;
    CU(cudaSetDevice(device));
    Scratch dx, dout;
    CU(dx.alloc(n * sizeof(float)));
    CU(dout.alloc(n * sizeof(float)));
    CU(cudaMemcpy(dx.p, x, n * sizeof(float), cudaMemcpyHostToDevice));
    CU(launch_probe_logf(dx.as<float>(), dout.as<float>(), n, nullptr));
    CU(cudaMemcpy(out, dout.p, n * sizeof(float), cudaMemcpyDeviceToHost));
    return NRAPS_OK;
}

extern "C" int nraps_dev_div(const float *t, const float *mu, float *out_fast, float *out_ieee, uint32_t n, int32_t device)
{
    if (!t || !mu || !out_fast || !out_ieee) return NRAPS_ERR_NULL;
    CU(cudaSetDevice(device));
    Scratch d[4];
    for (Scratch &b : d) CU(b.alloc(n * sizeof(float)));
    CU(cudaMemcpy(d[0].p, t, n * sizeof(float), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(d[1].p, mu, n * sizeof(float), cudaMemcpyHostToDevice));
    CU(launch_probe_div(d[0].as<float>(), d[1].as<float>(), d[2].as<float>(), d[3].as<float>(), n, nullptr));
    CU(cudaMemcpy(out_fast, d[2].p, n * sizeof(float), cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(out_ieee, d[3].p, n * sizeof(float), cudaMemcpyDeviceToHost));
    return NRAPS_OK;
}

extern "C" int nraps_dev_pcg32(uint64_t seed, uint64_t stream, uint64_t stride, uint64_t hid, uint32_t n,
                               uint32_t *out_u32, float *out_unit, int32_t device)
{
    if (!out_u32 || !out_unit) return NRAPS_ERR_NULL;
    CU(cudaSetDevice(device));
    Pcg m = pcg_seed(seed, stream);
    uint64_t jm, jp;
    pcg_jump_coeffs(m.inc, hid * stride, &jm, &jp);
    Scratch du, df;
    CU(du.alloc(n * sizeof(uint32_t)));
    CU(df.alloc(n * sizeof(float)));
    CU(launch_probe_pcg(jm * m.state + jp, m.inc, n, du.as<uint32_t>(), df.as<float>(), nullptr));
    CU(cudaMemcpy(out_u32, du.p, n * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(out_unit, df.p, n * sizeof(float), cudaMemcpyDeviceToHost));
    return NRAPS_OK;
}
