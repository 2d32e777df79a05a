// `nraps` driver: the reference binary's intended pipeline (its commented-out
// main, src/main.rs:332-364) with the Monte Carlo solver running on a B200:
//   process_input -> mesh_gen -> monte_carlo -> plot_solution (three CSV files)
// Usage: nraps [deck] [--out DIR] [--generations N] [--histories N] [--skip N]
//              [--seed S --stream Q --stride T] [--device D] [--scatter single_xi|rust_pre182|rust_182]
//              [--generation-log]  (one JSON line per generation on stderr: k, k_fund, bank size, source entropy)
//              [--fix-stale-xs] [--quiet] [--gpus N] [--tracking surface|woodcock] [--source uniform_fuel|fission_bank]
//              [--solution mc|diffusion|deck]  (diffusion = the reference's finite-difference solver, host code, cross-check
//              only; deck = what the deck's `Solution` key says, 1 = Monte Carlo, anything else diffusion, the dispatch the
//              reference's commented-out main intends (src/main.rs:336-346).  Default mc: the shipped decks say Solution = 0
//              and BASELINE's configurations are Monte Carlo runs of them)
// The deck defaults to ./TestCaseC.txt like the reference (src/process_input.rs:86).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <dlfcn.h>

#include "../../include/nraps_host.h"
#include "../../include/nraps_multi.h"

int main(int argc, char **argv)
{
    std::string deck_path = "./TestCaseC.txt", out_dir = ".";
    nraps_options opt{};
    opt.stale_xs = 1;
    long long gens = -1, hist = -1, skip = -1;
    int gpus = 1;
    bool gen_log = false, diffusion = false, solution_from_deck = false;
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        auto next = [&](const char *what) -> const char * {
            if (i + 1 >= argc) { std::fprintf(stderr, "missing value for %s\n", what); std::exit(2); }
            return argv[++i];
        };
        if (a == "--out") out_dir = next("--out");
        else if (a == "--generations") gens = std::atoll(next("--generations"));
        else if (a == "--histories") hist = std::atoll(next("--histories"));
        else if (a == "--skip") skip = std::atoll(next("--skip"));
        else if (a == "--seed") opt.seed = std::strtoull(next("--seed"), nullptr, 10);
        else if (a == "--stream") opt.stream = std::strtoull(next("--stream"), nullptr, 10);
        else if (a == "--stride") opt.stride = std::strtoull(next("--stride"), nullptr, 10);
        else if (a == "--device") opt.device = std::atoi(next("--device"));
        else if (a == "--fix-stale-xs") opt.stale_xs = 0;
        else if (a == "--quiet") opt.quiet = 1;
        else if (a == "--generation-log") gen_log = true;
        else if (a == "--gpus") gpus = std::atoi(next("--gpus"));
        else if (a == "--tracking") opt.tracking_mode = std::string(next("--tracking")) == "woodcock" ? NRAPS_TRACK_WOODCOCK : NRAPS_TRACK_SURFACE;
        else if (a == "--source") opt.source_mode = std::string(next("--source")) == "fission_bank" ? NRAPS_SOURCE_FISSION_BANK : NRAPS_SOURCE_UNIFORM_FUEL;
        else if (a == "--solution") {
            const std::string m = next("--solution");
            diffusion = m == "diffusion";
            solution_from_deck = m == "deck";
        }
        else if (a == "--scatter") {
            const std::string m = next("--scatter");
            opt.scatter_mode = m == "rust_pre182" ? NRAPS_SCATTER_RUST_PRE182 : m == "rust_182" ? NRAPS_SCATTER_RUST_182 : NRAPS_SCATTER_SINGLE_XI;
        } else if (a[0] != '-') deck_path = a;
        else { std::fprintf(stderr, "unknown option %s\n", a.c_str()); return 2; }
    }

    const auto t0 = std::chrono::steady_clock::now();
    nraps_deck deck;
    int rc = nraps_process_input(deck_path.c_str(), &deck);
    if (rc != NRAPS_OK) { std::fprintf(stderr, "process_input(%s): %s\n", deck_path.c_str(), nraps_strerror(rc)); return 1; }
    if (solution_from_deck) diffusion = deck.solution != 1;
    if (gens >= 0) deck.generations = (uint64_t)gens;
    if (hist >= 0) deck.histories = (uint64_t)hist;
    if (skip >= 0) deck.skip = (uint64_t)skip;

    nraps_mesh mesh;
    rc = nraps_mesh_gen(deck.matid, deck.n_matid, deck.mpfr, deck.mpwr, deck.numass, deck.dx_fuel, deck.dx_water, &mesh);
    if (rc != NRAPS_OK) { std::fprintf(stderr, "mesh_gen: %s\n", nraps_strerror(rc)); return 1; }

    nraps_problem prob;
    rc = nraps_problem_from(&deck, &mesh, 1.0f, &prob); // k_new = 1.0, src/main.rs:337
    if (rc != NRAPS_OK) { std::fprintf(stderr, "problem: %s\n", nraps_strerror(rc)); return 1; }

    const size_t GN = (size_t)prob.G * prob.N;
    std::vector<float> flux(GN), avg(GN), fis(prob.N), k(prob.generations), kf(prob.generations);
    nraps_results res{};
    res.flux = flux.data(); res.assembly_average = avg.data(); res.fission_source = fis.data();
    res.k = k.data(); res.k_fund = kf.data();
    std::vector<uint64_t> bank_sizes(prob.generations);
    std::vector<double> entropy(prob.generations);
    res.bank_sizes = bank_sizes.data(); res.entropy = entropy.data();
    if (diffusion) { // the reference's other solver (src/main.rs:338-346, src/discrete.rs), host code, cross-check only
        uint64_t iterations = 0;
        rc = nraps_diffusion_run(&prob, &res, 0, &iterations);
        if (rc != NRAPS_OK) { std::fprintf(stderr, "nalgebra_method: %s\n", nraps_strerror(rc)); return 1; }
        std::printf("%.10f\n", (double)k[0]); // src/discrete.rs:347
        res.fission_source = nullptr; res.k_fund = nullptr;
        rc = nraps_plot_solution(&res, prob.G, prob.generations, prob.N, (double)mesh.right[mesh.N - 1], out_dir.c_str());
        if (rc != NRAPS_OK) { std::fprintf(stderr, "plot_solution: %s\n", nraps_strerror(rc)); return 1; }
        const double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        std::printf("Run was completed in %lld milliseconds \n", (long long)(wall * 1e3));
        std::fprintf(stderr, "{\"k\": %.7g, \"iterations\": %llu}\n", (double)k[0], (unsigned long long)iterations);
        nraps_mesh_free(&mesh);
        nraps_deck_free(&deck);
        return 0;
    }
    if (gpus > 1) { // all GPUs of the box through libnraps_b200_nccl.so (loaded on demand: the core has no NCCL dependency)
        void *h = dlopen("libnraps_b200_nccl.so", RTLD_NOW);
        auto fn = h ? reinterpret_cast<decltype(&nraps_mc_run_multi)>(dlsym(h, "nraps_mc_run_multi")) : nullptr;
        if (!fn) { std::fprintf(stderr, "--gpus %d needs libnraps_b200_nccl.so: %s\n", gpus, dlerror()); return 1; }
        rc = fn(&prob, &opt, &res, gpus, nullptr);
    } else {
        rc = nraps_mc_run(&prob, &opt, &res);
    }
    if (rc != NRAPS_OK) {
        std::fprintf(stderr, "monte_carlo: %s %s\n", nraps_strerror(rc), rc == NRAPS_ERR_CUDA ? nraps_last_cuda_error() : "");
        return 1;
    }
    rc = nraps_plot_solution(&res, prob.G, prob.generations, prob.N, (double)mesh.right[mesh.N - 1], out_dir.c_str());
    if (rc != NRAPS_OK) { std::fprintf(stderr, "plot_solution: %s\n", nraps_strerror(rc)); return 1; }

    if (gen_log)
        for (uint64_t g = 0; g < prob.generations; ++g)
            std::fprintf(stderr, "{\"generation\": %llu, \"k\": %.7g, \"k_fund\": %.7g, \"bank\": %llu, \"entropy_bits\": %.6f}\n",
                         (unsigned long long)g, (double)k[g], (double)kf[g], (unsigned long long)bank_sizes[g], entropy[g]);
    const double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    const double total = (double)prob.histories * (double)prob.generations;
    std::printf("Run was completed in %lld milliseconds \n", (long long)(wall * 1e3)); // src/main.rs:366-369
    std::fprintf(stderr, "{\"k_fund\": %.7g, \"histories_per_s\": %.6g, \"device_s\": %.6g, \"collisions_per_history\": %.6g}\n",
                 (double)kf[prob.generations - 1], total / (res.seconds_device > 0 ? res.seconds_device : wall), res.seconds_device,
                 (double)res.counters[NRAPS_CT_COLLISIONS] / total);
    nraps_mesh_free(&mesh);
    nraps_deck_free(&deck);
    return 0;
}
