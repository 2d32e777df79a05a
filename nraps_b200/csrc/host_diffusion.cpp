// Multigroup finite-difference diffusion eigenvalue solver: the reference's second solver, kept as an independent
// physics cross-check of the Monte Carlo path.  Host code; a 408-cell tridiagonal system is
// not GPU work.
//
//   matrix .......... src/discrete.rs:4-107  (matrix_gen: per-group tridiagonal, albedo rows, the last row reading
//                     d_next / d_nextcurr of the FIRST two cells because the loop's bindings are out of scope)
//   fission source .. src/discrete.rs:109-134 (q_gen)
//   in-scatter ...... src/discrete.rs:136-160 (scat_calc)
//   iteration ....... src/discrete.rs:181-290 (Gauss-Seidel over groups, k update, the convergence test as written)
//   normalisation ... src/discrete.rs:292-343 (power constant with chunks(G) of the group-major flattening)
//
// Where the reference multiplies by a dense f32 inverse (nalgebra LU, :196-204), this file solves the same f32
// tridiagonal system by elimination in f64 and rounds the solution to f32: same answer to the rounding noise of the
// reference's inverse (~1e-4 relative), O(N) instead of O(N^3).
#include "nraps_host.h"

#include <cmath>
#include <cstdio>
#include <vector>

namespace {

struct Tri {
    std::vector<float> lo, di, up; // a[x][x-1], a[x][x], a[x][x+1]
};

float beta_of(float bound, float d_next, float d_curr)
{
    if (bound == 1.0f) return 1.0f;
    if (bound == 0.0f) return 0.25f;
    const float r = (1.0f - bound) / (1.0f + bound);
    return (1.0f - 0.25f * (r * (1.0f / d_next))) / (1.0f + 0.25f * (r * (1.0f / d_curr)));
}

Tri matrix_gen(const nraps_problem *p, uint32_t g)
{
    const uint32_t n = p->N, M = p->M, G = p->G;
    const float third = 1.0f / 3.0f;
    auto d_mul = [&](uint32_t i) { return (third * p->inv_sigtr[p->matid[i] + M * g]) * (1.0f / p->dx[i]); };
    auto removal = [&](uint32_t i) {
        const uint32_t m = p->matid[i];
        return p->dx[i] * (p->sigt[m + M * g] - p->scat[((G + 1) * g + G * G * m) & 0xffu]); // u8 index math, :40-42
    };
    Tri t;
    t.lo.assign(n, 0.0f); t.di.assign(n, 0.0f); t.up.assign(n, 0.0f);
    const float d_curr = d_mul(0), d_next = d_mul(1);
    const float d_nextcurr0 = (2.0f * d_curr * d_next) * (1.0f / (d_curr + d_next));
    const float beta_l = beta_of(p->boundl, d_next, d_curr);
    t.di[0] = 2.0f * d_curr * (1.0f - beta_l) + removal(0) + d_nextcurr0;
    t.up[0] = -d_nextcurr0;
    for (uint32_t x = 1; x + 1 < n; ++x) {
        const float dc = d_mul(x), dp = d_mul(x - 1), dn = d_mul(x + 1);
        const float d_prevcurr = (2.0f * dc * dp) * (1.0f / (dc + dp));
        const float d_nextcurr = (2.0f * dc * dn) * (1.0f / (dc + dn));
        t.lo[x] = -d_prevcurr;
        t.di[x] = d_prevcurr + removal(x) + d_nextcurr;
        t.up[x] = -d_nextcurr;
    }
    const float dce = (third * p->inv_sigtr[p->matid[n - 1] + M * g]) / p->dx[n - 1];
    const float dpe = (third * p->inv_sigtr[p->matid[n - 2] + M * g]) / p->dx[n - 2];
    const float d_prevcurr_e = (2.0f * dce * dpe) / (dce + dpe);
    const float beta_r = beta_of(p->boundr, d_next, dce);
    t.lo[n - 1] = -d_prevcurr_e;
    t.di[n - 1] = 2.0f * dce * (1.0f - beta_r) + removal(n - 1) + d_nextcurr0;
    return t;
}

// forward elimination factors of one group's matrix, reused every iteration
struct Factor {
    std::vector<double> cp, inv; // modified super-diagonal, reciprocal pivots
    std::vector<double> lo;
    bool ok = true;
};

Factor factor(const Tri &t)
{
    const size_t n = t.di.size();
    Factor f;
    f.cp.resize(n); f.inv.resize(n); f.lo.assign(t.lo.begin(), t.lo.end());
    double piv = t.di[0];
    for (size_t i = 0; i < n; ++i) {
        if (i) piv = (double)t.di[i] - f.lo[i] * f.cp[i - 1];
        if (!(std::fabs(piv) > 0.0) || !std::isfinite(piv)) { f.ok = false; return f; }
        f.inv[i] = 1.0 / piv;
        f.cp[i] = (double)t.up[i] * f.inv[i];
    }
    return f;
}

void solve(const Factor &f, const std::vector<float> &rhs, std::vector<double> &work, std::vector<float> &x)
{
    const size_t n = rhs.size();
    work[0] = (double)rhs[0] * f.inv[0];
    for (size_t i = 1; i < n; ++i) work[i] = ((double)rhs[i] - f.lo[i] * work[i - 1]) * f.inv[i];
    x[n - 1] = (float)work[n - 1];
    for (size_t i = n - 1; i-- > 0;) {
        work[i] -= f.cp[i] * work[i + 1];
        x[i] = (float)work[i];
    }
}

void q_gen(const nraps_problem *p, const std::vector<float> &flux, std::vector<float> &q)
{
    const uint32_t n = p->N, M = p->M, G = p->G;
    for (uint32_t i = 0; i < n; ++i) {
        const uint32_t m = p->matid[i];
        float prod = 0.0f;
        for (uint32_t x = 0; x < G; ++x) prod += (p->nut[m + M * x] * p->sigf[m + M * x]) * flux[(size_t)x * n + i];
        for (uint32_t g = 0; g < G; ++g) q[(size_t)g * n + i] = (prod * p->dx[i]) * p->chit[m + M * g];
    }
}

float seq_sum(const std::vector<float> &v)
{
    float s = 0.0f;
    for (float x : v) s += x;
    return s;
}

} // namespace

extern "C" int nraps_diffusion_run(const nraps_problem *p, nraps_results *r, uint64_t max_iterations, uint64_t *iterations)
{
    if (!p || !r || !r->flux || !r->assembly_average || !r->k) return NRAPS_ERR_NULL;
    if (!p->sigt || !p->sigf || !p->nut || !p->chit || !p->inv_sigtr || !p->scat || !p->matid || !p->dx) return NRAPS_ERR_NULL;
    const uint32_t n = p->N, M = p->M, G = p->G;
    if (n < 3 || G < 1 || G > 8 || M < 1 || M > 64 || p->numass < 1 || p->numass > n) return NRAPS_ERR_SHAPE;
    for (uint32_t i = 0; i < n; ++i) {
        if (p->matid[i] >= M) return NRAPS_ERR_MESH;
        if (!(p->dx[i] > 0.0f)) return NRAPS_ERR_MESH;
    }
    if (!max_iterations) max_iterations = 100000;

    std::vector<Factor> fac;
    for (uint32_t g = 0; g < G; ++g) {
        fac.push_back(factor(matrix_gen(p, g)));
        if (!fac.back().ok) return NRAPS_ERR_XS; // the reference panics in try_inverse().unwrap(), :201
    }
    std::vector<float> flux((size_t)G * n, 1.0f), q((size_t)G * n), temp_q, rhs(n), fresh(n);
    std::vector<double> work(n);
    q_gen(p, flux, q);
    float k = 1.0f, delta_flux = 1.0f, delta_k = 1.0f;
    uint64_t it = 0;
    while (delta_flux >= 1e-5f && delta_k >= 1e-6f && it < max_iterations) {
        ++it;
        temp_q = q;
        const float inv_k = 1.0f / k; // k.powi(-1)
        for (uint32_t g = 0; g < G; ++g) {
            float *fg = flux.data() + (size_t)g * n;
            for (uint32_t i = 0; i < n; ++i) {
                const uint32_t m = p->matid[i];
                float scat = 0.0f;
                for (uint32_t e = 0; e < G; ++e)
                    if (e != g) scat += (p->scat[G * G * m + G * e + g] * flux[(size_t)e * n + i]) * p->dx[i];
                rhs[i] = (q[(size_t)g * n + i] * inv_k) + scat;
            }
            solve(fac[g], rhs, work, fresh);
            // :236-288: index 0 and n-1 keep a running max, the indices between overwrite it, and every index but
            // the first divides by the already-replaced flux[g][0]
            delta_flux = std::fmax(std::fabs((fg[0] - fresh[0]) / fg[0]), delta_flux);
            delta_flux = std::fabs((fg[n - 2] - fresh[n - 2]) / fresh[0]);
            delta_flux = std::fmax(std::fabs((fg[n - 1] - fresh[n - 1]) / fresh[0]), delta_flux);
            for (uint32_t i = 0; i < n; ++i) fg[i] = fresh[i];
        }
        q_gen(p, flux, q);
        const float temp_k = k;
        k = temp_k * (seq_sum(q) / seq_sum(temp_q));
        delta_k = std::fabs((k - temp_k) / temp_k);
    }

    // power normalisation, :292-343 -- chunks(G) runs over the group-major flattening: G consecutive cells of a group
    std::vector<float> temp((size_t)G * n);
    for (uint32_t g = 0; g < G; ++g)
        for (uint32_t i = 0; i < n; ++i) {
            const uint32_t m = p->matid[i];
            temp[(size_t)g * n + i] = (flux[(size_t)g * n + i] * p->nut[m + M * g]) * p->sigf[m + M * g];
        }
    float s = 0.0f;
    for (uint32_t j = 0; j < n; ++j) {
        float chunk = 0.0f;
        for (uint32_t c = 0; c < G; ++c) chunk += temp[(size_t)j * G + c];
        s += chunk * p->dx[j];
    }
    const float power_constant = 3565e6f / (1.6022e-13f * 200.0f * s);
    for (size_t i = 0; i < flux.size(); ++i) r->flux[i] = flux[i] * power_constant;
    nraps_average_assembly(r->flux, G, n, p->numass, r->assembly_average);
    r->k[0] = k;
    if (iterations) *iterations = it;
    return NRAPS_OK;
}
