// CSV writer producing the reference's three output files byte-for-byte.
//
//   layout .......... src/plot_solution.rs:36-58: vars.csv = L, N, gens (one field per
//                     record); interface.csv = G flux rows, G assembly-average rows,
//                     one fission-source row; k_eff.csv = k row then k_fund row;
//                     csv-crate defaults: ',' delimiter, '\n' terminator, no quoting needed
//   number format ... Rust `to_string()` (:14-34): shortest round-trip digits,
//                     positional notation only, no ".0" on integral values
// plot.py is not spawned (:60 is presentation, out of scope).
#include "nraps_host.h"

#include <charconv>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>

namespace {

// digits + decimal exponent from the shortest round-trip scientific form
template <typename T> std::string rust_display(T v)
{
    if (std::isnan(v)) return "NaN";
    if (std::isinf(v)) return v > 0 ? "inf" : "-inf";
    if (v == 0) return std::signbit(v) ? "-0" : "0";
    char sci[64];
    auto res = std::to_chars(sci, sci + sizeof(sci), v, std::chars_format::scientific);
    std::string s(sci, res.ptr);
    const bool neg = s[0] == '-';
    if (neg) s.erase(0, 1);
    const size_t epos = s.find('e');
    const int exp10 = std::atoi(s.c_str() + epos + 1);
    std::string digits;
    for (size_t i = 0; i < epos; ++i)
        if (s[i] != '.') digits += s[i];
    // value = 0.d1d2d3... * 10^(exp10+1)
    const int point = exp10 + 1; // number of digits before the decimal point
    std::string out = neg ? "-" : "";
    if (point <= 0) {
        out += "0.";
        out.append((size_t)(-point), '0');
        out += digits;
    } else if ((size_t)point >= digits.size()) {
        out += digits;
        out.append((size_t)point - digits.size(), '0');
    } else {
        out += digits.substr(0, (size_t)point);
        out += '.';
        out += digits.substr((size_t)point);
    }
    return out;
}

size_t emit(const std::string &s, char *buf, size_t cap)
{
    if (buf && cap) {
        const size_t n = s.size() < cap - 1 ? s.size() : cap - 1;
        std::memcpy(buf, s.data(), n);
        buf[n] = '\0';
    }
    return s.size();
}

bool write_row(std::FILE *fh, const float *v, size_t n)
{
    std::string line;
    for (size_t i = 0; i < n; ++i) {
        if (i) line += ',';
        line += rust_display(v[i]);
    }
    line += '\n';
    return std::fwrite(line.data(), 1, line.size(), fh) == line.size();
}

} // namespace

extern "C" size_t nraps_format_f32(float v, char *buf, size_t cap) { return emit(rust_display(v), buf, cap); }
extern "C" size_t nraps_format_f64(double v, char *buf, size_t cap) { return emit(rust_display(v), buf, cap); }

extern "C" int nraps_plot_solution(const nraps_results *r, uint32_t G, uint64_t generations, uint32_t N,
                                   double assembly_length, const char *dir)
{
    if (!r || !r->flux || !r->assembly_average || !r->k) return NRAPS_ERR_NULL;
    // Diffusion results carry no fission source and no k_fund (src/discrete.rs:351-353).  The reference's writer then
    // fails on the empty fission record (csv UnequalLengths) after the 2G flux rows and never opens k_eff.csv; the
    // caller discards the error (`let _ =`, src/main.rs:358).  NULL fission_source / k_fund reproduce those files.
    const bool diffusion = !r->fission_source || !r->k_fund;
    const std::string base = (dir && *dir) ? std::string(dir) + "/" : std::string("./");

    std::FILE *fh = std::fopen((base + "vars.csv").c_str(), "wb");
    if (!fh) return NRAPS_ERR_IO;
    std::fprintf(fh, "%s\n%u\n%llu\n", rust_display(assembly_length).c_str(), N, (unsigned long long)generations);
    std::fclose(fh);

    fh = std::fopen((base + "interface.csv").c_str(), "wb");
    if (!fh) return NRAPS_ERR_IO;
    bool ok = true;
    for (uint32_t g = 0; g < G; ++g) ok = ok && write_row(fh, r->flux + (size_t)g * N, N);
    for (uint32_t g = 0; g < G; ++g) ok = ok && write_row(fh, r->assembly_average + (size_t)g * N, N);
    if (!diffusion) ok = ok && write_row(fh, r->fission_source, N);
    std::fclose(fh);
    if (!ok) return NRAPS_ERR_IO;
    if (diffusion) return NRAPS_OK;

    fh = std::fopen((base + "k_eff.csv").c_str(), "wb");
    if (!fh) return NRAPS_ERR_IO;
    ok = write_row(fh, r->k, generations) && write_row(fh, r->k_fund, generations);
    std::fclose(fh);
    return ok ? NRAPS_OK : NRAPS_ERR_IO;
}
