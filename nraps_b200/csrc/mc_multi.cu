// Single-process multi-GPU driver over the generation-level C ABI (include/nraps_multi.h).
// One host thread per device; NCCL over NVLink for the per-generation exchange:
//   tally:  ncclAllReduce(sum) on the uint64 tally buffer (mirrors the ordered thread join, src/mc_code.rs:331-338)
//   bank :  nothing is exchanged.  Every device keeps the bank it compacted, peer access is enabled between all pairs,
//           and the source kernel of the next generation loads each site from the device that banked it (NVLink);
//           the all-reduce, issued after the local compaction, is the barrier in between and carries the bank's
//           cell histogram along.
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
    ~RankScratch()
    {
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

    for (uint64_t gen = 0; gen < p->generations; ++gen) {
        RK(nraps_mc_transport(me.ctx, gen, begin, end - begin, s));
        if (bank) RK(nraps_mc_bank_compact(me.ctx, gen, s)); // before the all-reduce: it orders every bank before any reader
        RNC(ncclAllReduce(tally, tally, words, ncclUint64, ncclSum, me.comm, s));
        RK(nraps_mc_finalize_generation(me.ctx, gen, s));
        if (bank) RK(nraps_mc_bank_advance(me.ctx, gen, s));
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
    if (rc == NRAPS_OK && o->source_mode == NRAPS_SOURCE_FISSION_BANK) {
        // every device keeps its own bank; the others read it in place
        const uint64_t shard_max = (p->histories + (uint64_t)num_gpus - 1) / (uint64_t)num_gpus;
        std::vector<void *> bufs((size_t)num_gpus * 2, nullptr);
        for (int i = 0; i < num_gpus && rc == NRAPS_OK; ++i) rc = nraps_mc_bank_reserve(ranks[(size_t)i].ctx, shard_max, &bufs[(size_t)i * 2]);
        for (int i = 0; i < num_gpus && rc == NRAPS_OK; ++i) {
            if (cudaSetDevice(devs[(size_t)i]) != cudaSuccess) { rc = NRAPS_ERR_CUDA; break; }
            for (int j = 0; j < num_gpus; ++j) {
                if (j == i) continue;
                const cudaError_t e = cudaDeviceEnablePeerAccess(devs[(size_t)j], 0);
                if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
                else if (e != cudaSuccess) { rc = NRAPS_ERR_CUDA; break; }
            }
        }
        for (int i = 0; i < num_gpus && rc == NRAPS_OK; ++i) rc = nraps_mc_bank_peers(ranks[(size_t)i].ctx, num_gpus, i, bufs.data());
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
