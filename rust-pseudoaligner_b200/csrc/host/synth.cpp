// synth.cpp -- synthetic workloads of BASELINE.md section 4 and minimal FASTA/FASTQ readers
// (libpsa_host.so).  Everything is a pure function of (seed, index): a counter-based
// splitmix64 stream per object, so any rank / thread / machine regenerates identical data
// without shipping it.
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <string>
#include <thread>
#include <vector>

#include "../../../include/psa_host.h"

extern thread_local std::string g_host_err;

namespace {

inline uint64_t splitmix(uint64_t& s) {
    uint64_t z = (s += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
struct Rng {  // stream = f(seed, tag, index)
    uint64_t s;
    Rng(uint64_t seed, uint64_t tag, uint64_t index) {
        s = seed * 0xD1342543DE82EF95ULL + tag;
        s = splitmix(s) ^ (index * 0xA24BAED4963EE407ULL);
        splitmix(s);
    }
    uint64_t next() { return splitmix(s); }
    double uniform() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
    uint64_t below(uint64_t n) { return (uint64_t)(((unsigned __int128)next() * n) >> 64); }
    double normal() {  // Box-Muller
        double u1 = uniform(), u2 = uniform();
        if (u1 < 1e-300) u1 = 1e-300;
        return sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
    }
};

template <class F>
void parallel_for(int T, F f) {
    if (T <= 1) { f(0); return; }
    std::vector<std::thread> th;
    for (int t = 0; t < T; t++) th.emplace_back(f, t);
    for (auto& x : th) x.join();
}
int nthreads(int threads) {
    int T = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
    return std::max(1, std::min(T, 256));
}

}  // namespace

struct psa_transcriptome {
    std::vector<uint8_t> codes;
    std::vector<uint64_t> tx_off;
    // read sampling: transcripts with len >= L, cumulative number of start positions (per L, cached)
    uint32_t cached_len = 0;
    std::vector<uint32_t> elig;
    std::vector<uint64_t> cum;
    uint32_t cached_half[2] = {0, 0};
    std::vector<uint32_t> elig_h[2];
    std::vector<uint64_t> cum_h[2];
};

extern "C" psa_transcriptome* psa_synth_transcriptome(uint64_t seed, uint32_t n_genes, int threads) {
    const int T = nthreads(threads);
    constexpr int N_REP = 50, REP_LEN = 300;
    std::vector<std::vector<uint8_t>> reps(N_REP, std::vector<uint8_t>(REP_LEN));
    for (int r = 0; r < N_REP; r++) {
        Rng g(seed, 1, r);
        for (auto& c : reps[r]) c = (uint8_t)(g.next() >> 62);
    }
    // per gene: exons, isoforms (built independently, concatenated in gene order)
    std::vector<std::vector<uint8_t>> gene_codes(n_genes);
    std::vector<std::vector<uint64_t>> gene_lens(n_genes);
    std::atomic<uint32_t> next(0);
    parallel_for(T, [&](int) {
        for (;;) {
            uint32_t gi = next.fetch_add(1);
            if (gi >= n_genes) break;
            Rng g(seed, 2, gi);
            int n_ex = 4 + (int)g.below(21);  // 4..24
            std::vector<std::vector<uint8_t>> exons(n_ex);
            for (auto& e : exons) {
                double len = 140.0 * exp(0.8 * g.normal());
                int L = (int)std::max(30.0, std::min(3000.0, len));
                e.resize(L);
                for (auto& c : e) c = (uint8_t)(g.next() >> 62);
            }
            // geometric with mean 10 (p = 0.1), capped at 200
            int n_iso = 1;
            while (n_iso < 200 && g.uniform() >= 0.1) n_iso++;
            for (int iso = 0; iso < n_iso; iso++) {
                std::vector<uint8_t> s;
                int first = g.uniform() < 0.8 ? 0 : (int)g.below(std::min(3, n_ex));       // alternative first exon
                int last = g.uniform() < 0.8 ? n_ex - 1 : n_ex - 1 - (int)g.below(std::min(3, n_ex));
                if (last < first) last = first;
                for (int e = first; e <= last; e++) {
                    bool keep = (e == first || e == last) ? true : g.uniform() < 0.75;
                    if (keep) s.insert(s.end(), exons[e].begin(), exons[e].end());
                }
                if (g.uniform() < 0.05) {  // a diverged copy of a repeat element somewhere inside
                    std::vector<uint8_t> r = reps[g.below(N_REP)];
                    for (auto& c : r)
                        if (g.uniform() < 0.10) c = (uint8_t)((c + 1 + g.below(3)) & 3);
                    size_t p = (size_t)g.below(s.size() + 1);
                    s.insert(s.begin() + p, r.begin(), r.end());
                }
                gene_lens[gi].push_back(s.size());
                gene_codes[gi].insert(gene_codes[gi].end(), s.begin(), s.end());
            }
        }
    });
    psa_transcriptome* t = new psa_transcriptome();
    uint64_t total = 0, ntx = 0;
    for (uint32_t gi = 0; gi < n_genes; gi++) { total += gene_codes[gi].size(); ntx += gene_lens[gi].size(); }
    t->codes.resize(total);
    t->tx_off.reserve(ntx + 1);
    t->tx_off.push_back(0);
    uint64_t o = 0;
    for (uint32_t gi = 0; gi < n_genes; gi++) {
        if (!gene_codes[gi].empty()) memcpy(t->codes.data() + o, gene_codes[gi].data(), gene_codes[gi].size());
        uint64_t p = o;
        for (uint64_t l : gene_lens[gi]) { p += l; t->tx_off.push_back(p); }
        o += gene_codes[gi].size();
        std::vector<uint8_t>().swap(gene_codes[gi]);
    }
    return t;
}

extern "C" psa_transcriptome* psa_transcriptome_from_codes(const uint8_t* codes, const uint64_t* tx_off, uint32_t n_tx) {
    psa_transcriptome* t = new psa_transcriptome();
    t->tx_off.assign(tx_off, tx_off + n_tx + 1);
    t->codes.assign(codes + tx_off[0], codes + tx_off[n_tx]);
    for (auto& x : t->tx_off) x -= tx_off[0];
    return t;
}
extern "C" void psa_transcriptome_free(psa_transcriptome* t) { delete t; }
extern "C" uint32_t psa_transcriptome_n_tx(const psa_transcriptome* t) { return (uint32_t)(t->tx_off.size() - 1); }
extern "C" uint64_t psa_transcriptome_n_bases(const psa_transcriptome* t) { return t->codes.size(); }
extern "C" const uint8_t* psa_transcriptome_codes(const psa_transcriptome* t) { return t->codes.data(); }
extern "C" const uint64_t* psa_transcriptome_tx_off(const psa_transcriptome* t) { return t->tx_off.data(); }

static void eligibility(const psa_transcriptome* t, uint32_t L, std::vector<uint32_t>& elig, std::vector<uint64_t>& cum) {
    elig.clear();
    cum.clear();
    cum.push_back(0);
    const uint32_t n = (uint32_t)(t->tx_off.size() - 1);
    for (uint32_t i = 0; i < n; i++) {
        uint64_t len = t->tx_off[i + 1] - t->tx_off[i];
        if (len >= L && L > 0) {
            elig.push_back(i);
            cum.push_back(cum.back() + (len - L + 1));
        }
    }
}
// uniformly random start position among all substrings of length L: (transcript, offset)
static inline const uint8_t* pick(const psa_transcriptome* t, const std::vector<uint32_t>& elig,
                                  const std::vector<uint64_t>& cum, Rng& g) {
    uint64_t x = g.below(cum.back());
    size_t i = std::upper_bound(cum.begin(), cum.end(), x) - cum.begin() - 1;
    return t->codes.data() + t->tx_off[elig[i]] + (x - cum[i]);
}

extern "C" int psa_synth_reads(const psa_transcriptome* tc, uint64_t seed, uint64_t first, uint64_t n, uint32_t L,
                               uint8_t* out, uint64_t stride, uint8_t* kind_out, int threads) {
    if (!tc || !out || stride < L) { g_host_err = "bad argument"; return -1; }
    psa_transcriptome* t = const_cast<psa_transcriptome*>(tc);
    const uint32_t h0 = L / 2, h1 = L - L / 2;
    if (t->cached_len != L) { eligibility(t, L, t->elig, t->cum); t->cached_len = L; }
    if (t->cached_half[0] != h0) { eligibility(t, h0, t->elig_h[0], t->cum_h[0]); t->cached_half[0] = h0; }
    if (t->cached_half[1] != h1) { eligibility(t, h1, t->elig_h[1], t->cum_h[1]); t->cached_half[1] = h1; }
    const bool have_full = !t->elig.empty(), have_half = !t->elig_h[0].empty() && !t->elig_h[1].empty();
    static const char ACGT[4] = {'A', 'C', 'G', 'T'};
    const int T = nthreads(threads);
    parallel_for(T, [&](int th) {
        uint64_t b = n * th / T, e = n * (th + 1) / T;
        for (uint64_t i = b; i < e; i++) {
            Rng g(seed, 3, first + i);
            uint8_t* r = out + i * stride;
            double u = g.uniform();
            int kind = u < 0.90 ? 0 : (u < 0.95 ? 1 : 2);
            if (kind == 0 && !have_full) kind = 2;
            if (kind == 1 && !have_half) kind = 2;
            if (kind == 0) {
                const uint8_t* s = pick(t, t->elig, t->cum, g);
                for (uint32_t j = 0; j < L; j++) r[j] = s[j];
            } else if (kind == 1) {
                const uint8_t* a = pick(t, t->elig_h[0], t->cum_h[0], g);
                const uint8_t* c = pick(t, t->elig_h[1], t->cum_h[1], g);
                for (uint32_t j = 0; j < h0; j++) r[j] = a[j];
                for (uint32_t j = 0; j < h1; j++) r[h0 + j] = c[j];
            } else {
                for (uint32_t j = 0; j < L; j++) r[j] = (uint8_t)(g.next() >> 62);
            }
            if (kind != 2) {  // substitutions, p = 0.005 per base: geometric gaps
                const double lq = log(1.0 - 0.005);
                uint64_t j = (uint64_t)(log(1.0 - g.uniform()) / lq);
                while (j < L) {
                    r[j] = (uint8_t)((r[j] + 1 + g.below(3)) & 3);
                    j += 1 + (uint64_t)(log(1.0 - g.uniform()) / lq);
                }
            }
            for (uint32_t j = 0; j < L; j++) r[j] = (uint8_t)ACGT[r[j]];
            if (kind_out) kind_out[i] = (uint8_t)kind;
        }
    });
    return 0;
}

// ---------------------------------------------------------------------------------------------
// FASTA / FASTQ
// ---------------------------------------------------------------------------------------------
struct psa_seqfile {
    std::vector<std::string> names;
    std::vector<uint8_t> data;
    std::vector<uint64_t> off;
};
static bool slurp(const char* path, std::string& buf) {
    FILE* f = fopen(path, "rb");
    if (!f) { g_host_err = std::string("cannot open ") + path; return false; }
    char tmp[1 << 16];
    size_t n;
    while ((n = fread(tmp, 1, sizeof tmp, f)) > 0) buf.append(tmp, n);
    fclose(f);
    return true;
}
// record order = transcript index (ref src/utils.rs:71-88)
extern "C" psa_seqfile* psa_fasta_read(const char* path) {
    std::string buf;
    if (!slurp(path, buf)) return nullptr;
    psa_seqfile* s = new psa_seqfile();
    s->off.push_back(0);
    size_t i = 0, n = buf.size();
    bool open = false;
    while (i < n) {
        size_t e = buf.find('\n', i);
        if (e == std::string::npos) e = n;
        size_t le = e;
        if (le > i && buf[le - 1] == '\r') le--;
        if (le > i && buf[i] == '>') {
            if (open) s->off.push_back(s->data.size());
            s->names.emplace_back(buf.substr(i + 1, le - i - 1));
            open = true;
        } else if (open) {
            s->data.insert(s->data.end(), buf.begin() + i, buf.begin() + le);
        }
        i = e + 1;
    }
    if (open) s->off.push_back(s->data.size());
    return s;
}
// four-line records; id = header up to the first whitespace (bio::io::fastq Record::id)
extern "C" psa_seqfile* psa_fastq_read(const char* path) {
    std::string buf;
    if (!slurp(path, buf)) return nullptr;
    psa_seqfile* s = new psa_seqfile();
    s->off.push_back(0);
    size_t i = 0, n = buf.size();
    int line = 0;
    while (i < n) {
        size_t e = buf.find('\n', i);
        if (e == std::string::npos) e = n;
        size_t le = e;
        if (le > i && buf[le - 1] == '\r') le--;
        if (line == 0) {
            if (le == i) { i = e + 1; continue; }  // blank line between records
            if (buf[i] != '@') { delete s; g_host_err = "FASTQ: record does not start with '@'"; return nullptr; }
            size_t sp = i + 1;
            while (sp < le && buf[sp] != ' ' && buf[sp] != '\t') sp++;
            s->names.emplace_back(buf.substr(i + 1, sp - i - 1));
        } else if (line == 1) {
            s->data.insert(s->data.end(), buf.begin() + i, buf.begin() + le);
            s->off.push_back(s->data.size());
        } else if (line == 2) {
            if (le == i || buf[i] != '+') { delete s; g_host_err = "FASTQ: missing '+' line"; return nullptr; }
        }
        line = (line + 1) & 3;
        i = e + 1;
    }
    if (line != 0) { delete s; g_host_err = "FASTQ: truncated record"; return nullptr; }
    return s;
}
extern "C" void psa_seqfile_free(psa_seqfile* s) { delete s; }
extern "C" uint64_t psa_seqfile_n(const psa_seqfile* s) { return s->names.size(); }
extern "C" const char* psa_seqfile_name(const psa_seqfile* s, uint64_t i) { return s->names[i].c_str(); }
extern "C" const uint8_t* psa_seqfile_data(const psa_seqfile* s) { return s->data.data(); }
extern "C" const uint64_t* psa_seqfile_off(const psa_seqfile* s) { return s->off.data(); }
