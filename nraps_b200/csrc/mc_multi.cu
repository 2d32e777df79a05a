// Single-process multi-GPU driver over the generation-level C ABI (include/nraps_multi.h).
// One host thread per device; NCCL over NVLink for the per-generation exchange:
//   tally:  ncclAllReduce(sum) on uint64[G*N + 8]   (mirrors the ordered thread join, src/mc_code.rs:331-338)
//   bank :  ncclAllGather of counts, then of the padded local banks, compacted in rank order
#include <nccl.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/nraps_multi.h"

namespace {

struct Rank {
    int device = 0;
    nraps_mc_ctx *ctx = nullptr;
    ncclComm_t comm = nullptr;
    int rc = NRAPS_OK;
};

// A rank that fails must not leave the others waiting inside a collective for ever: the first failure aborts every
// communicator of the job (ncclCommAbort unblocks operations in flight), after which no rank issues NCCL calls.
struct Job {
    std::vector<ncclComm_t> comms;
    std::mutex lock;
    std::atomic<bool> aborted{false};
    void abort_all()
    {
        std::lock_guard<std::mutex> g(lock);
        if (aborted.exchange(true)) return;
        for (ncclComm_t &c : comms)
            if (c) { ncclCommAbort(c); c = nullptr; }
    }
};

// device scratch and the stream of one rank, released on every return path
struct RankScratch {
    cudaStream_t s = nullptr;
    unsigned long long *d_counts = nullptr, *d_stage = nullptr, *d_gathered = nullptr, *d_global[2] = {nullptr, nullptr};
    ~RankScratch()
    {
        cudaFree(d_counts); cudaFree(d_stage); cudaFree(d_gathered); cudaFree(d_global[0]); cudaFree(d_global[1]);
        if (s) cudaStreamDestroy(s);
    }
};

#define FAIL(code)                 \
    do {                           \
        me.rc = (code);            \
        job.abort_all();           \
        return;                    \
    } while (0)
#define RK(call)                                  \
    do {                                          \
        const int rc_ = (call);                   \
        if (rc_ != NRAPS_OK) FAIL(rc_);           \
    } while (0)
#define RCU(call)                                         \
    do {                                                  \
        if ((call) != cudaSuccess) FAIL(NRAPS_ERR_CUDA);  \
    } while (0)
#define RNC(call)                                                           \
    do {                                                                    \
        if (job.aborted.load()) { me.rc = NRAPS_ERR_STATE; return; }        \
        if ((call) != ncclSuccess) FAIL(NRAPS_ERR_CUDA);                    \
    } while (0)

void rank_main(Rank &me, Job &job, int rank, int world, const nraps_problem *p, const nraps_options *o, nraps_results *r)
{
    RankScratch sc;
    RCU(cudaSetDevice(me.device));
    RCU(cudaStreamCreateWithFlags(&sc.s, cudaStreamNonBlocking));
    cudaStream_t s = sc.s;
    void *tally = nullptr;
    uint64_t words = 0;
    RK(nraps_mc_tally_buffer(me.ctx, &tally, &words));
    const uint64_t H = p->histories;
    const uint64_t begin = H * (uint64_t)rank / (uint64_t)world, end = H * (uint64_t)(rank + 1) / (uint64_t)world;
    const bool bank = o->source_mode == NRAPS_SOURCE_FISSION_BANK;

    uint64_t stage_cap = 0, global_cap[2] = {0, 0};
    std::vector<unsigned long long> counts((size_t)world);
    if (bank) RCU(cudaMalloc((void **)&sc.d_counts, (size_t)(world + 1) * sizeof(unsigned long long)));

    for (uint64_t gen = 0; gen < p->generations; ++gen) {
        RK(nraps_mc_transport(me.ctx, gen, begin, end - begin, s));
        RNC(ncclAllReduce(tally, tally, words, ncclUint64, ncclSum, me.comm, s));
        RK(nraps_mc_finalize_generation(me.ctx, gen, s));
        if (!bank) continue;
        RK(nraps_mc_bank_compact(me.ctx, gen, s));
        void *local = nullptr;
        uint64_t n_local = 0;
        RK(nraps_mc_bank_local(me.ctx, &local, &n_local, s));
        RCU(cudaMemcpyAsync(sc.d_counts + world, &n_local, sizeof(n_local), cudaMemcpyHostToDevice, s));
        RNC(ncclAllGather(sc.d_counts + world, sc.d_counts, 1, ncclUint64, me.comm, s));
        RCU(cudaMemcpyAsync(counts.data(), sc.d_counts, (size_t)world * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
        RCU(cudaStreamSynchronize(s));
        unsigned long long max_n = 0, total = 0;
        for (unsigned long long c : counts) { max_n = std::max(max_n, c); total += c; }
        if (total == 0) { RK(nraps_mc_bank_set_source(me.ctx, gen, nullptr, 0, s)); continue; }
        if (max_n > stage_cap) { // padded staging: NCCL all-gather wants equal contributions
            cudaFree(sc.d_stage); cudaFree(sc.d_gathered);
            sc.d_stage = sc.d_gathered = nullptr;
            stage_cap = 0;
            RCU(cudaMalloc((void **)&sc.d_stage, max_n * sizeof(unsigned long long)));
            RCU(cudaMalloc((void **)&sc.d_gathered, max_n * (size_t)world * sizeof(unsigned long long)));
            stage_cap = max_n;
        }
        const int w = (int)(gen & 1u); // the bank read by generation gen+1 must outlive the next gather
        if (total > global_cap[w]) {
            cudaFree(sc.d_global[w]);
            sc.d_global[w] = nullptr;
            global_cap[w] = 0;
            RCU(cudaMalloc((void **)&sc.d_global[w], total * sizeof(unsigned long long)));
            global_cap[w] = total;
        }
        if (n_local) RCU(cudaMemcpyAsync(sc.d_stage, local, n_local * sizeof(unsigned long long), cudaMemcpyDeviceToDevice, s));
        RNC(ncclAllGather(sc.d_stage, sc.d_gathered, max_n, ncclUint64, me.comm, s));
        unsigned long long off = 0;
        for (int q = 0; q < world; ++q) { // rank order == canonical history order
            if (counts[(size_t)q])
                RCU(cudaMemcpyAsync(sc.d_global[w] + off, sc.d_gathered + (size_t)q * max_n, counts[(size_t)q] * sizeof(unsigned long long),
                                    cudaMemcpyDeviceToDevice, s));
            off += counts[(size_t)q];
        }
        RK(nraps_mc_bank_set_source(me.ctx, gen, sc.d_global[w], total, s));
    }
    if (rank == 0) RK(nraps_mc_fetch(me.ctx, r, s));
    RCU(cudaStreamSynchronize(s));
}

} // namespace

extern "C" int nraps_mc_run_multi(const nraps_problem *p, const nraps_options *o, nraps_results *r, int32_t num_gpus,
                                  const int32_t *devices)
{
    if (!p || !o || !r) return NRAPS_ERR_NULL;
    if (num_gpus < 1) return NRAPS_ERR_OPTION;
    if (num_gpus == 1 && !devices) return nraps_mc_run(p, o, r);
    if (r->tally_fixed) return NRAPS_ERR_OPTION; // per-generation tally dumps are a single-GPU debugging aid

    std::vector<Rank> ranks((size_t)num_gpus);
    std::vector<int> devs((size_t)num_gpus);
    for (int i = 0; i < num_gpus; ++i) devs[(size_t)i] = devices ? devices[i] : i;
    int rc = NRAPS_OK;
    for (int i = 0; i < num_gpus && rc == NRAPS_OK; ++i) {
        nraps_options oo = *o;
        oo.device = devs[(size_t)i];
        oo.quiet = 1;
        ranks[(size_t)i].device = devs[(size_t)i];
        rc = nraps_mc_create(p, &oo, &ranks[(size_t)i].ctx);
    }
    Job job;
    job.comms.assign((size_t)num_gpus, nullptr);
    if (rc == NRAPS_OK && ncclCommInitAll(job.comms.data(), num_gpus, devs.data()) != ncclSuccess) rc = NRAPS_ERR_CUDA;
    if (rc == NRAPS_OK) {
        if (!o->quiet) { std::printf("running MC code\n"); std::fflush(stdout); }
        for (int i = 0; i < num_gpus; ++i) ranks[(size_t)i].comm = job.comms[(size_t)i];
        std::vector<std::thread> threads;
        for (int i = 0; i < num_gpus; ++i)
            threads.emplace_back(rank_main, std::ref(ranks[(size_t)i]), std::ref(job), i, num_gpus, p, o, r);
        for (std::thread &t : threads) t.join();
        // report the failure that started it, not the NRAPS_ERR_STATE of the ranks that were told to stop
        for (const Rank &k : ranks)
            if (k.rc != NRAPS_OK && (rc == NRAPS_OK || rc == NRAPS_ERR_STATE)) rc = k.rc;
    }
    for (int i = 0; i < num_gpus; ++i) {
        if (job.comms[(size_t)i]) ncclCommDestroy(job.comms[(size_t)i]); // aborted communicators are already gone
        if (ranks[(size_t)i].ctx) nraps_mc_destroy(ranks[(size_t)i].ctx);
    }
    return rc;
}
