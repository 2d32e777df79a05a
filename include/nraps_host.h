/*
 * nraps_host.h -- C ABI of the host side around the Monte Carlo path: the same
 * input decks in, the same CSV files out.  CPU only; no CUDA needed.
 *
 *   nraps_process_input ... src/process_input.rs:85-175 (+ scanner :44-83)
 *   nraps_mesh_gen ........ src/main.rs:85-143
 *   nraps_walk_segments ... (no counterpart) segments of equal-width cells the surface kernel walks and scores by range
 *   nraps_plot_solution ... src/plot_solution.rs:7-58 (without spawning plot.py)
 *   nraps_average_assembly  src/mc_code.rs:259-274
 *   nraps_k_fund .......... src/mc_code.rs:368-376
 *   nraps_diffusion_run ... src/discrete.rs:181-356 (nalgebra_method, the reference's other solver; cross-check only)
 */
#ifndef NRAPS_HOST_H
#define NRAPS_HOST_H

#include <stddef.h>
#include <stdint.h>

#include "nraps_mc.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Variables + XSData + raw MatID list + DeltaX (src/main.rs:22-56); arrays owned by the struct. */
typedef struct nraps_deck {
    uint32_t analk, mattypes, energygroups, numass, numrods;
    uint64_t generations, histories, skip, mpfr, mpwr;
    float roddia, rodpitch; /* rodpitch = RodPitch - RodDia, src/process_input.rs:102 */
    float boundl, boundr;
    float dx_fuel, dx_water; /* DeltaX, src/process_input.rs:109-112 */
    uint32_t n_xs, n_scat, n_matid;
    float *sigt, *sigs, *mu, *siga, *sigf, *nut, *chit, *inv_sigtr; /* [n_xs] */
    float *scat;                                                      /* [n_scat] */
    uint8_t *matid;                                                   /* [n_matid] per-pin list */
    int32_t solution; /* raw "Solution" value (1 = Monte Carlo) */
    int32_t solver;   /* raw "solver" value */
} nraps_deck;

/* Vec<Mesh> as structure-of-arrays + fuel_indices; arrays owned by the struct. */
typedef struct nraps_mesh {
    uint32_t N, NF;
    uint8_t *matid;
    float *dx, *left, *right;
    uint64_t *fuel_indices;
} nraps_mesh;

int nraps_process_input(const char *path, nraps_deck *out);
void nraps_deck_free(nraps_deck *d);
int nraps_mesh_gen(const uint8_t *matid, uint32_t n_matid, uint64_t mpfr, uint64_t mpwr, uint32_t numass,
                   float dx_fuel, float dx_water, nraps_mesh *out);
void nraps_mesh_free(nraps_mesh *m);
/* fill an nraps_problem with borrowed pointers into deck + mesh */
int nraps_problem_from(const nraps_deck *d, const nraps_mesh *m, float k0, nraps_problem *out);

/* Rust `f32::to_string()` / `f64::to_string()`: shortest round-trip digits, positional notation. */
size_t nraps_format_f32(float v, char *buf, size_t cap);
size_t nraps_format_f64(double v, char *buf, size_t cap);
/* writes <dir>/vars.csv, interface.csv, k_eff.csv; with r->fission_source or r->k_fund NULL (diffusion results) only
 * vars.csv and the 2G flux rows of interface.csv, which is what the reference's writer leaves behind in that case */
int nraps_plot_solution(const nraps_results *r, uint32_t G, uint64_t generations, uint32_t N,
                        double assembly_length, const char *dir);

/* Walk segments of the surface-tracking kernel (no reference counterpart: it is how the kernel books the scores of
 * src/mc_code.rs:163,173 as range updates).  A segment = a maximal range of cells of one material run whose widths
 * right[i] - left[i] are the same binary32 number.  Per cell i of segment [a, b):
 *   stops[i]      = (edge index at which a walk to the left stops) | (the same to the right) << 16: the neutron has then
 *                   entered cell a-1 (edge a-1 ahead) / cell b (edge b+1 ahead); the domain's boundary cells are not
 *                   entered by the walk loop, so a segment holding cell 0 stops at edge 0 and one holding cell N-1 at edge N
 *   width_bits[i] = the bits of that width
 * Returns the number of segments through *n_segments (optional). */
int nraps_walk_segments(const uint8_t *matid, const float *left, const float *right, uint32_t N, uint32_t *stops,
                        uint32_t *width_bits, uint32_t *n_segments);

void nraps_average_assembly(const float *flux, uint32_t G, uint32_t N, uint32_t numass, float *out);
void nraps_k_fund(const float *k, uint64_t generations, uint64_t skip, float *out);

/* Finite-difference diffusion eigenvalue solve of the same problem (src/discrete.rs:181-356): fills r->flux[G*N],
 * r->assembly_average[G*N] and r->k[0]; fission_source and k_fund are not produced (the reference returns empty
 * vectors, :351-353).  generations / histories / skip of the problem are ignored.  max_iterations = 0 means 100000
 * (the reference loops without bound); *iterations (optional) receives the number of power iterations.  The dense
 * f32 inverse of the reference is replaced by an f64 tridiagonal elimination: equal to the rounding noise of that
 * inverse, not bit for bit. */
int nraps_diffusion_run(const nraps_problem *p, nraps_results *r, uint64_t max_iterations, uint64_t *iterations);

#ifdef __cplusplus
}
#endif
#endif
