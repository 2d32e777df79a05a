// Mesh construction, average_assembly and k_fund: the O(N) host pieces either
// side of the transport path, bug-for-bug with the reference.
//
//   nraps_mesh_gen ........... src/main.rs:85-143  (Q9 centre trim, Q10 CR
//                              regions meshed at water width, Q11 f32 edge accumulation)
//   nraps_average_assembly ... src/mc_code.rs:259-274 (Q12)
//   nraps_k_fund ............. src/mc_code.rs:368-376
#include "nraps_host.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

static inline bool is_fuel(uint8_t m) { return m == 0 || m == 1; }

extern "C" int nraps_mesh_gen(const uint8_t *pins, uint32_t n_pins, uint64_t mpfr, uint64_t mpwr, uint32_t numass,
                              float dx_fuel, float dx_water, nraps_mesh *out)
{
    if (!pins || !out) return NRAPS_ERR_NULL;
    std::memset(out, 0, sizeof(*out));
    if (numass == 0) return NRAPS_ERR_SHAPE;

    // one entry per cell: fuel pins get mpfr cells, everything else mpwr cells
    std::vector<uint8_t> cells;
    for (uint32_t i = 0; i < n_pins; ++i) cells.insert(cells.end(), is_fuel(pins[i]) ? mpfr : mpwr, pins[i]);

    // drop the doubled water gap between assemblies; the cut position is
    // recomputed from the shrinking length after every removal (src/main.rs:99-103)
    for (uint32_t a = 1; a < numass; ++a)
        for (uint64_t r = 0; r < mpwr; ++r) {
            const size_t cut = ((size_t)a * cells.size()) / numass;
            if (cut >= cells.size()) return NRAPS_ERR_SHAPE; // Vec::remove would panic
            cells.erase(cells.begin() + (std::ptrdiff_t)cut);
        }
    // half a gap off each end (src/main.rs:107-108)
    const size_t half = (size_t)(mpwr / 2);
    if (cells.size() < 2 * half || cells.size() == 2 * half) return NRAPS_ERR_SHAPE;
    cells.erase(cells.begin(), cells.begin() + (std::ptrdiff_t)half);
    cells.resize(cells.size() - half);

    const size_t N = cells.size();
    out->N = (uint32_t)N;
    out->matid = static_cast<uint8_t *>(std::malloc(N));
    out->dx = static_cast<float *>(std::malloc(N * sizeof(float)));
    out->left = static_cast<float *>(std::malloc(N * sizeof(float)));
    out->right = static_cast<float *>(std::malloc(N * sizeof(float)));
    std::vector<uint64_t> fuel;
    float edge = 0.0f; // accumulated in f32 exactly like `mesh_left += dx`
    for (size_t i = 0; i < N; ++i) {
        const float w = is_fuel(cells[i]) ? dx_fuel : dx_water;
        out->matid[i] = cells[i];
        out->dx[i] = w;
        out->left[i] = edge;
        edge = edge + w;
        out->right[i] = edge;
        if (is_fuel(cells[i])) fuel.push_back(i);
    }
    out->NF = (uint32_t)fuel.size();
    out->fuel_indices = static_cast<uint64_t *>(std::malloc((fuel.size() ? fuel.size() : 1) * sizeof(uint64_t)));
    if (!fuel.empty()) std::memcpy(out->fuel_indices, fuel.data(), fuel.size() * sizeof(uint64_t));
    return NRAPS_OK;
}

extern "C" void nraps_mesh_free(nraps_mesh *m)
{
    if (!m) return;
    std::free(m->matid); std::free(m->dx); std::free(m->left); std::free(m->right); std::free(m->fuel_indices);
    std::memset(m, 0, sizeof(*m));
}

extern "C" int nraps_problem_from(const nraps_deck *d, const nraps_mesh *m, float k0, nraps_problem *p)
{
    if (!d || !m || !p) return NRAPS_ERR_NULL;
    std::memset(p, 0, sizeof(*p));
    p->M = d->mattypes; p->G = d->energygroups; p->N = m->N; p->NF = m->NF; p->numass = d->numass;
    p->generations = d->generations; p->histories = d->histories; p->skip = d->skip;
    p->boundl = d->boundl; p->boundr = d->boundr; p->dx_fuel = d->dx_fuel; p->dx_water = d->dx_water; p->k0 = k0;
    p->sigt = d->sigt; p->sigs = d->sigs; p->mu = d->mu; p->siga = d->siga; p->sigf = d->sigf; p->nut = d->nut;
    p->chit = d->chit; p->inv_sigtr = d->inv_sigtr; p->scat = d->scat;
    p->matid = m->matid; p->dx = m->dx; p->left = m->left; p->right = m->right; p->fuel_indices = m->fuel_indices;
    const uint64_t need = (uint64_t)p->M * p->G;
    if (d->n_xs < need || d->n_scat < need * p->G) return NRAPS_ERR_SHAPE;
    return NRAPS_OK;
}

extern "C" void nraps_average_assembly(const float *flux, uint32_t G, uint32_t N, uint32_t numass, float *out)
{
    const uint32_t span = N / numass; // integer split; trailing cells (if any) keep 0
    for (uint32_t g = 0; g < G; ++g) {
        const float *row = flux + (size_t)g * N;
        float *dst = out + (size_t)g * N;
        for (uint32_t i = 0; i < N; ++i) dst[i] = 0.0f;
        for (uint32_t a = 0; a < numass; ++a) {
            float acc = 0.0f;
            for (uint32_t i = a * span; i < (a + 1) * span; ++i) acc += row[i]; // sequential f32 sum
            const float mean = acc / (float)span;
            for (uint32_t i = a * span; i < (a + 1) * span; ++i) dst[i] = mean;
        }
    }
}

extern "C" void nraps_k_fund(const float *k, uint64_t gens, uint64_t skip, float *out)
{
    for (uint64_t n = 0; n < gens; ++n) out[n] = 0.0f;
    if (skip >= gens) return; // the reference indexes out of bounds and panics here
    out[skip] = k[skip];
    for (uint64_t n = skip + 1; n < gens; ++n) {
        float acc = 0.0f;
        for (uint64_t j = skip; j <= n; ++j) acc += k[j];
        out[n] = acc / (float)(uint64_t)(n - (skip - 1)); // usize arithmetic wraps for skip == 0
    }
}

extern "C" int nraps_walk_segments(const uint8_t *matid, const float *left, const float *right, uint32_t N, uint32_t *stops,
                                   uint32_t *width_bits, uint32_t *n_segments)
{
    if (!matid || !left || !right || !stops || !width_bits) return NRAPS_ERR_NULL;
    if (N == 0 || N > 65535) return NRAPS_ERR_SHAPE;
    uint32_t count = 0;
    for (uint32_t i = 0; i < N;) {
        // Edges accumulate in f32 upstream (src/main.rs:119-140), so the width of a fuel or water cell changes by an ulp
        // where the position crosses a power of two: a material run is one segment, or two around such a point.
        const float w = right[i] - left[i];
        uint32_t wbits;
        std::memcpy(&wbits, &w, sizeof(wbits));
        uint32_t j = i + 1;
        while (j < N && matid[j] == matid[i]) {
            const float wj = right[j] - left[j];
            if (std::memcmp(&wj, &w, sizeof(float)) != 0) break;
            ++j;
        }
        const uint32_t stop_right = std::min(j, N - 1) + 1, stop_left = i > 0 ? i - 1 : 0;
        for (uint32_t q = i; q < j; ++q) {
            stops[q] = stop_left | (stop_right << 16);
            width_bits[q] = wbits;
        }
        ++count;
        i = j;
    }
    if (n_segments) *n_segments = count;
    return NRAPS_OK;
}
