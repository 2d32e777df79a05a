/*
 * oracle_math.h -- TEST INFRASTRUCTURE ONLY (see oracle/README.md).
 *
 * Scalar arithmetic shared by every function of the CPU oracle: the PCG32
 * generator, the uniform mapping, the natural log and the CDF search.  Plain
 * C11, no dependency on the product code under nraps_b200/.
 *
 * Reference lines each function follows (paths relative to /root/reference):
 *   pcg32 step / output ..... src/rand.rs:74-85
 *   pcg32 seeding ........... src/rand.rs:49-71 (== pcg-c-basic pcg32_srandom_r)
 *   advance ................. pcg-c `pcg_advance_lcg_64` (published algorithm;
 *                             rand.rs has no jump-ahead -- new capability
 *                             named by BASELINE.json north_star)
 *   uniform ................. replaces src/rand.rs:95-100 / crate rand 0.8.5
 *                             `random::<f32>()`; see SURVEY.md section 9 Q2
 *   logf .................... replaces `f32::ln` at src/mc_code.rs:148,209
 *   lower_bound ............. `partition_point(|&x| x < v).min(len-1)` at
 *                             src/mc_code.rs:31,124-126
 *
 * PARITY STATUS: the reference pins none of RNG stream, ln() or end-to-end k
 * (no seedable RNG at HEAD, no golden k): "parity unpinned" for those.  What
 * IS pinned: PCG32 against the upstream pcg-c-basic demo vector, and the four
 * reference unit tests (tests/test_oracle_golden.py).
 *
 * Build with -ffp-contract=off: every rounding below is intentional.
 */
#ifndef NRAPS_ORACLE_MATH_H
#define NRAPS_ORACLE_MATH_H

#include <math.h>
#include <stdint.h>
#include <string.h>

#define ORACLE_PCG_MULT 6364136223846793005ULL

typedef struct {
    uint64_t state;
    uint64_t inc;
} oracle_pcg32;

/* src/rand.rs:74-85 */
static inline uint32_t oracle_pcg32_next(oracle_pcg32 *r)
{
    uint64_t old = r->state;
    r->state = old * ORACLE_PCG_MULT + r->inc;
    uint32_t xorshifted = (uint32_t)(((old >> 18) ^ old) >> 27);
    uint32_t rot = (uint32_t)(old >> 59);
    return (xorshifted >> rot) | (xorshifted << ((32u - rot) & 31u));
}

/* src/rand.rs:49-71 with the wall-clock seed replaced by an explicit one. */
static inline void oracle_pcg32_seed(oracle_pcg32 *r, uint64_t seed, uint64_t seq)
{
    r->state = 0u;
    r->inc = (seq << 1) | 1u;
    (void)oracle_pcg32_next(r);
    r->state += seed; /* rand.rs:65 writes seed+inc over a state that equals inc */
    (void)oracle_pcg32_next(r);
}

/* Jump the stream `delta` draws ahead in O(log delta). */
static inline void oracle_pcg32_advance(oracle_pcg32 *r, uint64_t delta)
{
    uint64_t cur_mult = ORACLE_PCG_MULT, cur_plus = r->inc;
    uint64_t acc_mult = 1u, acc_plus = 0u;
    while (delta > 0) {
        if (delta & 1u) {
            acc_mult *= cur_mult;
            acc_plus = acc_plus * cur_mult + cur_plus;
        }
        cur_plus = (cur_mult + 1u) * cur_plus;
        cur_mult *= cur_mult;
        delta >>= 1;
    }
    r->state = acc_mult * r->state + acc_plus;
}

/* xi = ((u >> 9) + 0.5) * 2^-23 : exact in f32, never 0, 0.5 or 1 (Q2). */
static inline float oracle_u32_to_unit(uint32_t u)
{
    return ((float)(u >> 9) + 0.5f) * 1.1920928955078125e-07f;
}

static inline float oracle_uniform(oracle_pcg32 *r)
{
    return oracle_u32_to_unit(oracle_pcg32_next(r));
}

/* src/mc_code.rs:35-37 */
static inline float oracle_direction(float xi)
{
    return 2.0f * xi - 1.0f;
}

/*
 * Natural log for normal positive x.  Cephes-style: x = m * 2^e with m in
 * (sqrt(1/2), sqrt(2)], degree-8 polynomial in f = m - 1 evaluated with fused
 * multiply-adds in a fixed order so the GPU kernel can reproduce every bit.
 */
static inline float oracle_logf(float x)
{
    uint32_t ix;
    memcpy(&ix, &x, 4);
    int e = (int)(ix >> 23) - 127;
    uint32_t im = (ix & 0x007fffffu) | 0x3f800000u;
    float m;
    memcpy(&m, &im, 4);
    if (m > 1.41421356f) {
        m = m * 0.5f;
        e += 1;
    }
    float f = m - 1.0f;
    float z = f * f;
    float p = 7.0376836292e-2f;
    p = fmaf(p, f, -1.1514610310e-1f);
    p = fmaf(p, f, 1.1676998740e-1f);
    p = fmaf(p, f, -1.2420140846e-1f);
    p = fmaf(p, f, 1.4249322787e-1f);
    p = fmaf(p, f, -1.6668057665e-1f);
    p = fmaf(p, f, 2.0000714765e-1f);
    p = fmaf(p, f, -2.4999993993e-1f);
    p = fmaf(p, f, 3.3333331174e-1f);
    float y = (f * z) * p;
    float fe = (float)e;
    y = fmaf(fe, -2.12194440e-4f, y);
    y = fmaf(-0.5f, z, y);
    float r = f + y;
    r = fmaf(fe, 0.693359375f, r);
    return r;
}

/* partition_point(|&x| x < v).min(n-1) for a non-decreasing (or all-NaN) cdf. */
static inline uint32_t oracle_lower_bound_clamped(const float *cdf, uint32_t n, float v)
{
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
        uint32_t mid = lo + (hi - lo) / 2;
        if (cdf[mid] < v)
            lo = mid + 1;
        else
            hi = mid;
    }
    return lo < n - 1 ? lo : n - 1;
}

#endif
