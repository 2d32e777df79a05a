// Prototype of the closed-form walk through one segment against the cell-by-cell loop (bit-exact or bust).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <math.h>
static inline uint32_t f2u(float f){uint32_t u;memcpy(&u,&f,4);return u;}
static inline float u2f(uint32_t u){float f;memcpy(&f,&u,4);return f;}

// cells that can be skipped at once: every skipped cell k has a_k >= lim (so its collision test surely passes) and
// every intermediate ds equals the sequential f32 result.  Returns j (0 = none) and the new |ds|.
static uint32_t jump(float a, uint32_t mw, int ewb, float lim, uint32_t nmax, float *out)
{
    const uint32_t ia = f2u(a);
    const int eb = (int)(ia >> 23);
    const uint32_t m = (ia & 0x7fffffu) | 0x800000u;
    const int sh = eb - ewb;
    *out = a;
    if (sh < 0 || sh > 24 || m <= 0x800000u) return 0;
    uint32_t q = mw >> sh;
    const uint32_t rem = mw & ((1u << sh) - 1u), half = (1u << sh) >> 1;
    if (sh && rem == half) {                // tie: round to even.  Once the mantissa is even it stays even and every step
        if (m & 1u) return 0;               // subtracts the even one of {q, q+1}; an odd mantissa is fixed by one exact step
        q += q & 1u;
    } else if (rem > half) q += 1;          // (sh == 0: rem = half = 0)
    // lim on the grid of a: mt = ceil(lim / U), U = 2^(eb-150)
    const float scale = u2f((uint32_t)(277 - eb) << 23);   // 2^(150-eb)
    const float ls = lim * scale;
    if (!(ls < 16777216.0f)) return 0;
    const uint32_t mt = (uint32_t)ceilf(ls);
    if (m < mt) return 0;
    const float rq = (1.0f / (float)q) * 0.99999f;           // biased low: an underestimate of j is still exact
    uint32_t j1 = (uint32_t)((float)(m - mt) * rq) + 1u;
    uint32_t j2 = (uint32_t)((float)(m - 0x800001u) * rq);
    uint32_t j = j1 < j2 ? j1 : j2;
    if (j > nmax) j = nmax;
    const uint32_t mr = m - j * q;
    *out = u2f(((uint32_t)eb << 23) | (mr & 0x7fffffu));
    return j;
}

typedef struct { int cell; float ds, x, end; int collided; } Res;

// sequential: neutron at edge position e[idx] (idx = edge index behind... ) moving dir, in cell c; walk while cell != stop
static Res walk_seq(const float *e, int c, int dir, float ds, float w, int stop)
{
    Res r; int fwd = dir > 0; float xc = e[c + (fwd ? 0 : 1)]; const float tn = fwd ? -w : w;
    r.collided = 0; r.end = 0;
    while (c != stop) {
        r.end = xc + ds;
        float d = r.end - xc;
        if (!(fabsf(d) > w)) { r.collided = 1; break; }
        ds = ds + tn;
        xc = e[c + (fwd ? 1 : 0)];
        c += dir;
    }
    r.cell = c; r.ds = ds; r.x = xc; return r;
}

static long g_skipped = 0, g_steps = 0, g_rounds = 0, g_sure = 0;
// the walk as mc_transport.cu does it: closed-form strides while |ds| exceeds the first power of two >= 6 widths, then
// the surely-crossed cells one by one without the test (ds <- fl(ds -+ w) only), then the reference's own step
static Res walk_cf(const float *e, int c, int dir, float ds, float w, int stop, float L)
{
    Res r; int fwd = dir > 0; float xc = e[c + (fwd ? 0 : 1)]; const float tn = fwd ? -w : w;
    const uint32_t iw = f2u(w); const uint32_t mw = (iw & 0x7fffffu) | 0x800000u; const int ewb = (int)(iw >> 23);
    r.collided = 0; r.end = 0;
    const float lim = fmaf(L + fabsf(ds), 4.76837158203125e-07f, w); // w + 2^-21 (L + |ds|)
    const float w4 = u2f((f2u(w * 6.0f) + 0x7fffffu) & 0xff800000u); // first power of two >= 6 w
    const uint32_t total = (uint32_t)((stop - c) * dir);
    uint32_t done = 0; float a = fabsf(ds);
    for (int round = 0; round < 8; ++round) {
        const uint32_t rem = total - done;
        if (rem < 4 || !(a > w4)) break;
        ++g_rounds;
        float an; const uint32_t j = jump(a, mw, ewb, lim, rem, &an);
        a = an; done += j;
        if (done == total || !(a >= lim)) break;
        a = a - w; done += 1;            // one exact decrement: binade transition / parity fix; the cell is surely crossed
    }
    if (done) { ds = ds < 0 ? -a : a; c += (int)done * dir; g_skipped += done; }
    while (c != stop && fabsf(ds) >= lim) { ds = ds + tn; c += dir; ++g_sure; }   // the sure loop: no test, no position
    xc = e[c + (fwd ? 0 : 1)];                                                    // the edge crossed last
    while (c != stop) {
        r.end = xc + ds; float d = r.end - xc; ++g_steps;
        if (!(fabsf(d) > w)) { r.collided = 1; break; }
        ds = ds + tn; xc = e[c + (fwd ? 1 : 0)]; c += dir;
    }
    r.cell = c; r.ds = ds; r.x = xc; return r;
}

int main(int argc, char **argv)
{
    long trials = argc > 1 ? atol(argv[1]) : 1000000;
    srand48(argc > 2 ? atol(argv[2]) : 1);
    // edges: alternate fuel (dxf) / water (dxw) runs accumulated in f32 like mesh_gen; segments = equal-width ranges inside a run
    enum { NMAX = 70000 };
    static float e[NMAX + 1]; static int segid[NMAX];
    long bad = 0;
    for (int mesh = 0; mesh < 6; ++mesh) {
        int mpfr = (int[]){8, 80, 160, 320, 17, 640}[mesh], mpwr = mpfr / 2 ? mpfr / 2 : 1;
        float dxf = 0.94f / (float)mpfr, dxw = (1.262f - 0.94f) / (float)mpwr;
        int N = 0; float left = 0.f; int run = 0;
        while (N + mpfr + mpwr < NMAX && run < 34) {
            for (int i = 0; i < mpwr && run; ++i) { e[N] = left; left = left + dxw; segid[N++] = 2 * run; }
            for (int i = 0; i < mpfr; ++i) { e[N] = left; left = left + dxf; segid[N++] = 2 * run + 1; }
            ++run;
        }
        e[N] = left;
        const float L = e[N];
        for (long t = 0; t < trials; ++t) {
            int c = (int)(drand48() * N); int dir = lrand48() & 1 ? 1 : -1;
            // segment = cells of the same run with the same width bits around c
            float w = e[c + 1] - e[c];
            int lo = c, hi = c + 1;
            while (lo > 0 && segid[lo - 1] == segid[c] && f2u(e[lo] - e[lo - 1]) == f2u(w)) --lo;
            while (hi < N && segid[hi] == segid[c] && f2u(e[hi + 1] - e[hi]) == f2u(w)) ++hi;
            int stop = dir > 0 ? hi : lo - 1;
            double mag = lrand48() % 3 == 0 ? drand48() * 40.0 : (lrand48() % 2 ? drand48() * 3.0 : drand48() * (hi - lo) * w * 1.2);
            if (lrand48() % 4 == 0) { // adversarial: |ds| lands within +-10 % of the sure-crossing margin, or within an ulp or two of a whole number of widths, after k cells
                const double k = (double)(lrand48() % (hi - lo + 2));
                mag = lrand48() % 2 ? k * w + w * 0.5 + (L + k * w) * 4.76837158203125e-07 * (0.9 + 0.2 * drand48())
                                    : k * w + w * 0.5 + (drand48() - 0.5) * 4e-6;
            }
            float ds = (float)(dir * (w * 0.5 + mag));
            Res a = walk_seq(e, c, dir, ds, w, stop), b = walk_cf(e, c, dir, ds, w, stop, L);
            if (a.cell != b.cell || f2u(a.ds) != f2u(b.ds) || a.collided != b.collided || f2u(a.x) != f2u(b.x) || (a.collided && f2u(a.end) != f2u(b.end))) {
                if (bad < 10) printf("BAD mesh %d c=%d dir=%d ds=%.9g w=%.9g stop=%d: seq cell %d ds %.9g col %d | cf cell %d ds %.9g col %d\n", mpfr, c, dir, ds, w, stop, a.cell, a.ds, a.collided, b.cell, b.ds, b.collided);
                ++bad;
            }
        }
        printf("mesh mpfr=%d N=%d L=%.4f: bad so far %ld, skipped %ld sure steps %ld exact steps %ld rounds %ld\n", mpfr, N, L, bad, g_skipped, g_sure, g_steps, g_rounds);
    }
    return bad != 0;
}
