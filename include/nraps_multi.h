/*
 * nraps_multi.h -- optional single-process multi-GPU driver of the Monte Carlo path (library
 * libnraps_b200_nccl.so, links NCCL).  One host thread and one nraps_mc_ctx per device; histories shard
 * across devices (the reference's thread fork / ordered join, src/mc_code.rs:302-338, at GPU granularity);
 * per generation one ncclAllReduce(sum, uint64) of the tally buffer -- nothing else: in fission_bank mode every
 * device keeps the bank it compacted and the births of the next generation load each site from the device that
 * banked it (peer access over NVLink; nraps_mc_bank_reserve / nraps_mc_bank_peers of nraps_mc.h).  Results are
 * bit-identical to nraps_mc_run on one GPU.
 *
 * The one-process-per-GPU route (torchrun + nraps_b200.monte_carlo_distributed) needs none of this; this entry
 * point exists so that a host without Python (the `nraps` driver, the Rust shim) gets all GPUs of a box
 * through the same C ABI.
 */
#ifndef NRAPS_MULTI_H
#define NRAPS_MULTI_H

#include "nraps_mc.h"

#ifdef __cplusplus
extern "C" {
#endif

/* devices == NULL => CUDA ordinals 0 .. num_gpus-1; nraps_options.device is ignored */
int nraps_mc_run_multi(const nraps_problem *p, const nraps_options *o, nraps_results *r, int32_t num_gpus,
                       const int32_t *devices);

#ifdef __cplusplus
}
#endif
#endif
