// Deck reader: same `key = value` input files as the reference, same quirks.
//
// Behaviour contract (reference src/process_input.rs):
//   * a '#' anywhere discards the rest of the line, including a key=value that
//     precedes it on that line (:56-59), and the first byte of the next line is
//     not examined by the scanner (:80);
//   * the key is the text from line start to ONE BYTE BEFORE '=' (:61), trimmed
//     and lower-cased (:66-68); keys are recognised by (last two chars, length)
//     only (:13-42);
//   * a repeated key appends " value" to its slot (:73) -- SigT on two lines,
//     one Scat line per material, one MatID line per assembly;
//   * rodpitch := RodPitch - RodDia (:102); dx = roddia/mpfr, gap/mpwr (:109-112);
//   * inv_sigtr = 1 / (sigt - mu*sigs) in f32 (:152-156).
// The deck path is an argument here (the reference hard-codes ./TestCaseC.txt, :86).
#include "nraps_host.h"

#include <algorithm>
#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace {

enum Slot {
    S_ANALK, S_MATTYPES, S_GROUPS, S_GENERATIONS, S_HISTORIES, S_SKIP, S_NUMASS, S_NUMRODS, S_RODDIA,
    S_RODPITCH, S_MPFR, S_MPWR, S_BOUNDL, S_BOUNDR, S_SIGT, S_SIGS, S_MU, S_SIGA, S_SIGF, S_NUT, S_CHIT,
    S_SCAT, S_MATID, S_SOLUTION, S_SOLVER, S_JUNK, S_COUNT
};

struct KeySig { char a, b; unsigned len; Slot slot; };

// (second-to-last char, last char, key length) -> slot
const KeySig kSigs[] = {
    {'l','k',5,S_ANALK}, {'e','s',8,S_MATTYPES}, {'p','s',12,S_GROUPS}, {'n','s',11,S_GENERATIONS},
    {'e','s',9,S_HISTORIES}, {'i','p',4,S_SKIP}, {'s','s',6,S_NUMASS}, {'d','s',7,S_NUMRODS},
    {'i','a',6,S_RODDIA}, {'c','h',8,S_RODPITCH}, {'f','r',4,S_MPFR}, {'w','r',4,S_MPWR},
    {'d','l',6,S_BOUNDL}, {'d','r',6,S_BOUNDR}, {'g','t',4,S_SIGT}, {'g','s',4,S_SIGS}, {'m','u',2,S_MU},
    {'g','a',4,S_SIGA}, {'g','f',4,S_SIGF}, {'u','t',3,S_NUT}, {'i','t',4,S_CHIT}, {'a','t',4,S_SCAT},
    {'i','d',5,S_MATID}, {'o','n',8,S_SOLUTION}, {'e','r',6,S_SOLVER},
};

std::string trimmed(const char *b, const char *e)
{
    while (b < e && std::isspace(static_cast<unsigned char>(*b))) ++b;
    while (e > b && std::isspace(static_cast<unsigned char>(e[-1]))) --e;
    return std::string(b, e);
}

bool classify(const std::string &key, Slot *out)
{
    if (key.size() < 2) return false; // the reference panics slicing key[len-2..]
    for (const KeySig &s : kSigs)
        if (s.len == key.size() && s.a == key[key.size() - 2] && s.b == key[key.size() - 1]) { *out = s.slot; return true; }
    *out = S_JUNK;
    return true;
}

// One pass over the bytes; `slots[i]` receives " v1 v2 ..." exactly like the reference's String slots.
bool scan(const std::vector<char> &buf, std::string slots[S_COUNT])
{
    const size_t n = buf.size();
    size_t at = 0, line0 = 0, key_end = 0, val0 = 0;
    while (at < n) {
        const char c = buf[at];
        if (c == '#') {
            while (at < n && buf[at] != '\n') ++at;
            line0 = ++at; // start of the next line ...
            ++at;         // ... whose first byte is never inspected
            continue;
        }
        if (c == '=') {
            key_end = at ? at - 1 : 0;
            val0 = at + 1;
        } else if (c == '\n') {
            if (key_end > line0) {
                std::string key = trimmed(&buf[line0], &buf[key_end]);
                std::transform(key.begin(), key.end(), key.begin(), [](unsigned char ch) { return std::tolower(ch); });
                Slot s;
                if (!classify(key, &s)) return false;
                slots[s] += ' ';
                slots[s] += trimmed(&buf[std::min(val0, at)], &buf[at]);
            }
            line0 = at + 1;
        }
        ++at;
    }
    return true;
}

bool to_u64(const std::string &s, uint64_t *out)
{
    std::string t = trimmed(s.data(), s.data() + s.size());
    if (t.empty()) return false;
    char *end = nullptr;
    unsigned long long v = std::strtoull(t.c_str(), &end, 10);
    if (*end != '\0' || t[0] == '-') return false;
    *out = v;
    return true;
}

bool to_f32(const std::string &s, float *out)
{
    std::string t = trimmed(s.data(), s.data() + s.size());
    if (t.empty()) return false;
    char *end = nullptr;
    float v = std::strtof(t.c_str(), &end); // correctly rounded, like Rust's str::parse::<f32>
    if (*end != '\0') return false;
    *out = v;
    return true;
}

bool to_f32_list(const std::string &s, std::vector<float> *out)
{
    size_t i = 0;
    while (i < s.size()) {
        while (i < s.size() && std::isspace(static_cast<unsigned char>(s[i]))) ++i;
        size_t j = i;
        while (j < s.size() && !std::isspace(static_cast<unsigned char>(s[j]))) ++j;
        if (j > i) {
            float v;
            if (!to_f32(s.substr(i, j - i), &v)) return false;
            out->push_back(v);
        }
        i = j;
    }
    return true;
}

float *dup_f32(const std::vector<float> &v)
{
    float *p = static_cast<float *>(std::malloc(std::max<size_t>(1, v.size()) * sizeof(float)));
    if (p && !v.empty()) std::memcpy(p, v.data(), v.size() * sizeof(float));
    return p;
}

} // namespace

extern "C" int nraps_process_input(const char *path, nraps_deck *d)
{
    if (!path || !d) return NRAPS_ERR_NULL;
    std::memset(d, 0, sizeof(*d));
    std::FILE *fh = std::fopen(path, "rb");
    if (!fh) return NRAPS_ERR_IO;
    std::vector<char> buf;
    char tmp[1 << 16];
    size_t got;
    while ((got = std::fread(tmp, 1, sizeof(tmp), fh)) > 0) buf.insert(buf.end(), tmp, tmp + got);
    std::fclose(fh);

    std::string slots[S_COUNT];
    if (!scan(buf, slots)) return NRAPS_ERR_IO;

    uint64_t u[12];
    const Slot ints[] = {S_ANALK, S_MATTYPES, S_GROUPS, S_GENERATIONS, S_HISTORIES, S_SKIP,
                         S_NUMASS, S_NUMRODS, S_MPFR, S_MPWR, S_SOLUTION};
    for (size_t i = 0; i < sizeof(ints) / sizeof(ints[0]); ++i)
        if (!to_u64(slots[ints[i]], &u[i])) return NRAPS_ERR_IO;
    // u8 fields in the reference (src/main.rs:23-30)
    if (u[0] > 255 || u[1] > 255 || u[2] > 255 || u[6] > 255 || u[7] > 255 || u[10] > 255) return NRAPS_ERR_SHAPE;
    d->analk = (uint32_t)u[0];
    d->mattypes = (uint32_t)u[1];
    d->energygroups = (uint32_t)u[2];
    d->generations = u[3];
    d->histories = u[4];
    d->skip = u[5];
    d->numass = (uint32_t)u[6];
    d->numrods = (uint32_t)u[7];
    d->mpfr = u[8];
    d->mpwr = u[9];
    d->solution = (int32_t)u[10];
    {
        std::string sv = trimmed(slots[S_SOLVER].data(), slots[S_SOLVER].data() + slots[S_SOLVER].size());
        d->solver = (sv == "1") ? 1 : (sv == "2") ? 2 : (sv == "3") ? 3 : 0; // :168-173
    }
    float pitch;
    if (!to_f32(slots[S_RODDIA], &d->roddia) || !to_f32(slots[S_RODPITCH], &pitch) ||
        !to_f32(slots[S_BOUNDL], &d->boundl) || !to_f32(slots[S_BOUNDR], &d->boundr))
        return NRAPS_ERR_IO;
    d->rodpitch = pitch - d->roddia;
    d->dx_fuel = d->roddia / (float)d->mpfr;
    d->dx_water = d->rodpitch / (float)d->mpwr;

    std::vector<float> sigt, sigs, mu, siga, sigf, nut, chit, scat;
    if (!to_f32_list(slots[S_SIGT], &sigt) || !to_f32_list(slots[S_SIGS], &sigs) || !to_f32_list(slots[S_MU], &mu) ||
        !to_f32_list(slots[S_SIGA], &siga) || !to_f32_list(slots[S_SIGF], &sigf) || !to_f32_list(slots[S_NUT], &nut) ||
        !to_f32_list(slots[S_CHIT], &chit) || !to_f32_list(slots[S_SCAT], &scat))
        return NRAPS_ERR_IO;
    const size_t nxs = sigt.size();
    // sigs / mu shorter than sigt: the reference indexes out of bounds at :152-156; a short siga / sigf / nut / chit
    // panics later, on the first lookup past its end (src/mc_code.rs:121,347).  Every table is handed out as [n_xs],
    // so all of that is one shape error here; entries past n_xs (harmless upstream) are dropped.
    for (std::vector<float> *t : {&sigs, &mu, &siga, &sigf, &nut, &chit}) {
        if (t->size() < nxs) return NRAPS_ERR_SHAPE;
        t->resize(nxs);
    }
    std::vector<float> inv(nxs);
    for (size_t i = 0; i < nxs; ++i) {
        const float prod = mu[i] * sigs[i];
        const float tr = sigt[i] - prod;
        inv[i] = 1.0f / tr; // powi(-1)
    }
    std::vector<uint8_t> pins;
    {
        const std::string &s = slots[S_MATID];
        size_t i = 0;
        while (i < s.size()) {
            while (i < s.size() && std::isspace(static_cast<unsigned char>(s[i]))) ++i;
            size_t j = i;
            while (j < s.size() && !std::isspace(static_cast<unsigned char>(s[j]))) ++j;
            if (j > i) {
                uint64_t v;
                if (!to_u64(s.substr(i, j - i), &v) || v > 255) return NRAPS_ERR_IO;
                pins.push_back((uint8_t)v);
            }
            i = j;
        }
    }
    d->n_xs = (uint32_t)nxs;
    d->n_scat = (uint32_t)scat.size();
    d->n_matid = (uint32_t)pins.size();
    d->sigt = dup_f32(sigt); d->sigs = dup_f32(sigs); d->mu = dup_f32(mu); d->siga = dup_f32(siga);
    d->sigf = dup_f32(sigf); d->nut = dup_f32(nut); d->chit = dup_f32(chit); d->inv_sigtr = dup_f32(inv);
    d->scat = dup_f32(scat);
    d->matid = static_cast<uint8_t *>(std::malloc(std::max<size_t>(1, pins.size())));
    if (d->matid && !pins.empty()) std::memcpy(d->matid, pins.data(), pins.size());
    return NRAPS_OK;
}

extern "C" void nraps_deck_free(nraps_deck *d)
{
    if (!d) return;
    float **fs[] = {&d->sigt, &d->sigs, &d->mu, &d->siga, &d->sigf, &d->nut, &d->chit, &d->inv_sigtr, &d->scat};
    for (float **p : fs) { std::free(*p); *p = nullptr; }
    std::free(d->matid);
    d->matid = nullptr;
}
