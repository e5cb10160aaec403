// build_graph.cpp -- host builder of the coloured compacted de Bruijn graph (libpsa_host.so).
//
// Produces the arrays of psa_index_desc.  Semantics follow the reference
// (10XGenomics/rust-pseudoaligner @ 9d9cab8), the algorithm is this project's own:
//   - every k-mer of every transcript with len >= k, stranded (src/build_index.rs:127-151,
//     src/config.rs:14);
//   - colour of a k-mer = ascending, de-duplicated list of the transcripts containing it,
//     interned to a dense id in order of first appearance over the sorted k-mers
//     (CountFilterEqClass::summarize, src/equiv_classes.rs:62-91);
//   - exts of a k-mer = union over its occurrences of the neighbouring bases
//     (src/equiv_classes.rs:73, src/build_index.rs:144);
//   - unitig = maximal path whose every internal link is the unique right ext of its source,
//     the unique left ext of its target and joins equal colours (ScmapCompress,
//     src/build_index.rs:171,178); a closed cycle is cut at its smallest k-mer.
// The reference shards by minimizer and compacts with the debruijn crate; here the k-mer
// occurrences are radix-partitioned by prefix, sorted per bucket on all host threads, grouped,
// colours interned through a sharded signature table, and unitigs walked from their heads.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <functional>
#include <string>
#include <thread>
#include <vector>

#include "../../../include/psa_host.h"

typedef unsigned __int128 u128;

thread_local std::string g_host_err;
extern "C" const char* psa_host_last_error(void) { return g_host_err.c_str(); }

struct psa_graph {
    uint32_t k = 0;
    uint64_t n_kmers = 0, n_cycles = 0;
    std::vector<uint64_t> seq_words, node_start, eq_offsets;
    std::vector<uint32_t> node_len, node_eq, eq_members;
    std::vector<uint8_t> node_exts;
};

namespace {

inline uint64_t mix64(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL;
    x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL;
    x ^= x >> 33;
    return x;
}

template <class F>
void parallel_for(int T, F f) {  // f(thread_index)
    if (T <= 1) { f(0); return; }
    std::vector<std::thread> th;
    for (int t = 0; t < T; t++) th.emplace_back(f, t);
    for (auto& x : th) x.join();
}
// dynamic scheduling over [0, n) in grains
template <class F>
void parallel_chunks(int T, uint64_t n, uint64_t grain, F f) {  // f(begin, end, thread)
    std::atomic<uint64_t> next(0);
    parallel_for(T, [&](int t) {
        for (;;) {
            uint64_t b = next.fetch_add(grain);
            if (b >= n) break;
            f(b, std::min(n, b + grain), t);
        }
    });
}

template <class K>
struct Occ {
    K kmer;
    uint32_t tx;
    uint8_t exts;
};

constexpr uint32_t NONE32 = 0xFFFFFFFFu;

// signature table shard: open addressing on the 128-bit colour signature
struct ClassEntry {
    uint64_t sig_lo, sig_hi;
    uint64_t rep;   // smallest distinct-k-mer index with this colour
    uint32_t len;
    uint32_t id;
};
struct Shard {
    std::vector<ClassEntry> tab;
    uint64_t mask = 0, used = 0;
    void init(uint64_t cap_pow2) { tab.assign(cap_pow2, ClassEntry{0, 0, ~0ULL, 0, 0}); mask = cap_pow2 - 1; used = 0; }
    void grow() {
        std::vector<ClassEntry> old;
        old.swap(tab);
        init((mask + 1) * 2);
        for (auto& e : old)
            if (e.rep != ~0ULL) {
                uint64_t h = e.sig_lo & mask;
                while (tab[h].rep != ~0ULL) h = (h + 1) & mask;
                tab[h] = e;
                used++;
            }
    }
};

template <class K>
psa_graph* build(const uint8_t* codes, const uint64_t* tx_off, uint32_t n_tx, uint32_t k, int T) {
    const int KB = 2 * (int)k;
    const K kmask = (KB == (int)sizeof(K) * 8) ? ~(K)0 : (((K)1 << KB) - 1);
    const int PB = std::min(KB, 12);  // bucket = top PB bits of the k-mer
    const uint64_t NBK = 1ULL << PB;
    const int bshift = KB - PB;

    // ---- transcripts -> thread ranges of roughly equal size
    std::vector<uint32_t> tstart(T + 1, n_tx);
    {
        uint64_t total = tx_off[n_tx] - tx_off[0];
        uint32_t t = 0;
        tstart[0] = 0;
        for (int i = 1; i < T; i++) {
            uint64_t want = tx_off[0] + total * i / T;
            while (t < n_tx && tx_off[t] < want) t++;
            tstart[i] = t;
        }
        tstart[T] = n_tx;
    }
    // ---- pass 1: count occurrences per (thread, bucket)
    std::vector<std::vector<uint64_t>> cnt(T, std::vector<uint64_t>(NBK, 0));
    std::atomic<int> bad(0);
    auto scan = [&](int t, bool emit, Occ<K>* occ, std::vector<uint64_t>* pos) {
        for (uint32_t tx = tstart[t]; tx < tstart[t + 1]; tx++) {
            const uint8_t* s = codes + tx_off[tx];
            const uint64_t len = tx_off[tx + 1] - tx_off[tx];
            if (len < k) continue;
            K km = 0;
            for (uint64_t i = 0; i < len; i++) {
                if (s[i] > 3) { bad = 1; return; }
                km = ((km << 2) | s[i]) & kmask;
                if (i + 1 >= k) {
                    uint64_t b = (uint64_t)(km >> bshift);
                    if (!emit) { cnt[t][b]++; continue; }
                    uint64_t p = i + 1 - k;
                    uint8_t e = 0;
                    if (p > 0) e |= (uint8_t)(1u << (4 + s[p - 1]));
                    if (i + 1 < len) e |= (uint8_t)(1u << s[i + 1]);
                    Occ<K>& o = occ[(*pos)[b]++];
                    o.kmer = km; o.tx = tx; o.exts = e;
                }
            }
        }
    };
    parallel_for(T, [&](int t) { scan(t, false, nullptr, nullptr); });
    if (bad) { g_host_err = "base code > 3 in a transcript"; return nullptr; }
    std::vector<uint64_t> bstart(NBK + 1, 0);
    for (uint64_t b = 0; b < NBK; b++) {
        uint64_t s = 0;
        for (int t = 0; t < T; t++) s += cnt[t][b];
        bstart[b + 1] = bstart[b] + s;
    }
    const uint64_t n_occ = bstart[NBK];
    Occ<K>* occ = (Occ<K>*)malloc((n_occ + 1) * sizeof(Occ<K>));
    if (!occ) { g_host_err = "out of memory (occurrences)"; return nullptr; }
    {
        std::vector<std::vector<uint64_t>> pos(T, std::vector<uint64_t>(NBK));
        for (uint64_t b = 0; b < NBK; b++) {
            uint64_t p = bstart[b];
            for (int t = 0; t < T; t++) { pos[t][b] = p; p += cnt[t][b]; }
        }
        parallel_for(T, [&](int t) { scan(t, true, occ, &pos[t]); });
    }
    cnt.clear();
    // ---- sort each bucket by (k-mer, transcript)
    parallel_chunks(T, NBK, 1, [&](uint64_t b0, uint64_t b1, int) {
        for (uint64_t b = b0; b < b1; b++)
            std::sort(occ + bstart[b], occ + bstart[b + 1], [](const Occ<K>& x, const Occ<K>& y) {
                return x.kmer != y.kmer ? x.kmer < y.kmer : x.tx < y.tx;
            });
    });
    // ---- distinct k-mers per bucket
    std::vector<uint64_t> dstart(NBK + 1, 0);
    parallel_chunks(T, NBK, 4, [&](uint64_t b0, uint64_t b1, int) {
        for (uint64_t b = b0; b < b1; b++) {
            uint64_t d = 0;
            for (uint64_t i = bstart[b]; i < bstart[b + 1]; i++) d += (i == bstart[b] || occ[i].kmer != occ[i - 1].kmer);
            dstart[b + 1] = d;
        }
    });
    for (uint64_t b = 0; b < NBK; b++) dstart[b + 1] += dstart[b];
    const uint64_t n_dist = dstart[NBK];
    if (n_dist >= NONE32) { free(occ); g_host_err = "more than 2^32-2 distinct k-mers"; return nullptr; }
    std::vector<K> kmer(n_dist + 1);
    std::vector<uint8_t> exts(n_dist + 1);
    std::vector<uint32_t> eq(n_dist + 1);
    std::vector<uint64_t> sig_lo(n_dist + 1), sig_hi(n_dist + 1), first_occ(n_dist + 2);
    parallel_chunks(T, NBK, 4, [&](uint64_t b0, uint64_t b1, int) {
        for (uint64_t b = b0; b < b1; b++) {
            uint64_t d = dstart[b];
            for (uint64_t i = bstart[b]; i < bstart[b + 1];) {
                uint64_t j = i;
                uint8_t e = 0;
                uint64_t h1 = 0x243F6A8885A308D3ULL, h2 = 0x13198A2E03707344ULL;
                uint32_t prev = NONE32;
                while (j < bstart[b + 1] && occ[j].kmer == occ[i].kmer) {
                    e |= occ[j].exts;
                    if (occ[j].tx != prev) {  // sort + dedup (src/equiv_classes.rs:78-79)
                        prev = occ[j].tx;
                        h1 = mix64(h1 ^ prev);
                        h2 = mix64(h2 + 0x9E3779B97F4A7C15ULL * (prev + 1));
                    }
                    j++;
                }
                kmer[d] = occ[i].kmer; exts[d] = e; sig_lo[d] = h1; sig_hi[d] = h2; first_occ[d] = i;
                d++;
                i = j;
            }
        }
    });
    first_occ[n_dist] = n_occ;
    // colour list of distinct k-mer d = de-duplicated tx of occ[first_occ[d] .. first_occ[d+1])
    auto colour_equal = [&](uint64_t a, uint64_t b) {
        uint64_t i = first_occ[a], ie = first_occ[a + 1], j = first_occ[b], je = first_occ[b + 1];
        while (i < ie && j < je) {
            if (occ[i].tx != occ[j].tx) return false;
            uint32_t t = occ[i].tx;
            while (i < ie && occ[i].tx == t) i++;
            while (j < je && occ[j].tx == t) j++;
        }
        return i == ie && j == je;
    };
    auto colour_len = [&](uint64_t a) {
        uint32_t n = 0, prev = NONE32;
        for (uint64_t i = first_occ[a]; i < first_occ[a + 1]; i++)
            if (occ[i].tx != prev) { prev = occ[i].tx; n++; }
        return n;
    };
    // ---- intern colours: shard s owns the signatures with sig_hi % S == s
    const int S = T;
    std::vector<Shard> shards(S);
    parallel_for(S, [&](int s) {
        Shard& sh = shards[s];
        sh.init(1 << 12);
        for (uint64_t d = 0; d < n_dist; d++) {
            if ((int)(sig_hi[d] % (uint64_t)S) != s) continue;
            uint64_t h = sig_lo[d] & sh.mask;
            for (;;) {
                ClassEntry& e = sh.tab[h];
                if (e.rep == ~0ULL) {
                    e.sig_lo = sig_lo[d]; e.sig_hi = sig_hi[d]; e.rep = d; e.len = colour_len(d);
                    if (++sh.used * 2 > sh.mask) sh.grow();
                    break;
                }
                if (e.sig_lo == sig_lo[d] && e.sig_hi == sig_hi[d] && colour_equal(e.rep, d)) break;
                h = (h + 1) & sh.mask;
            }
        }
    });
    // dense ids in order of first appearance over the sorted k-mers
    std::vector<ClassEntry*> classes;
    for (auto& sh : shards)
        for (auto& e : sh.tab)
            if (e.rep != ~0ULL) classes.push_back(&e);
    std::sort(classes.begin(), classes.end(), [](const ClassEntry* a, const ClassEntry* b) { return a->rep < b->rep; });
    const uint64_t n_eq = classes.size();
    psa_graph* g = new psa_graph();
    g->k = k;
    g->n_kmers = n_dist;
    g->eq_offsets.assign(n_eq + 1, 0);
    for (uint64_t c = 0; c < n_eq; c++) {
        classes[c]->id = (uint32_t)c;
        g->eq_offsets[c + 1] = g->eq_offsets[c] + classes[c]->len;
    }
    g->eq_members.resize(g->eq_offsets[n_eq] + 1);
    parallel_chunks(T, n_eq, 1024, [&](uint64_t c0, uint64_t c1, int) {
        for (uint64_t c = c0; c < c1; c++) {
            uint64_t o = g->eq_offsets[c];
            uint32_t prev = NONE32;
            for (uint64_t i = first_occ[classes[c]->rep]; i < first_occ[classes[c]->rep + 1]; i++)
                if (occ[i].tx != prev) { prev = occ[i].tx; g->eq_members[o++] = prev; }
        }
    });
    g->eq_members.resize(g->eq_offsets[n_eq]);
    parallel_chunks(T, n_dist, 1 << 16, [&](uint64_t d0, uint64_t d1, int) {
        for (uint64_t d = d0; d < d1; d++) {
            const Shard& sh = shards[sig_hi[d] % (uint64_t)S];
            uint64_t h = sig_lo[d] & sh.mask;
            for (;;) {
                const ClassEntry& e = sh.tab[h];
                if (e.sig_lo == sig_lo[d] && e.sig_hi == sig_hi[d] && (e.rep == d || colour_equal(e.rep, d))) {
                    eq[d] = e.id;
                    break;
                }
                h = (h + 1) & sh.mask;
            }
        }
    });
    free(occ);
    occ = nullptr;
    shards.clear();
    { std::vector<uint64_t>().swap(sig_lo); std::vector<uint64_t>().swap(sig_hi); std::vector<uint64_t>().swap(first_occ); }

    // ---- k-mer lookup: prefix table over the sorted distinct k-mers
    int PL = 8;
    while ((1ULL << PL) < n_dist / 4 && PL < 26) PL++;
    PL = std::min(PL, KB);
    const int lshift = KB - PL;
    const uint64_t NPL = 1ULL << PL;
    std::vector<uint32_t> ptab(NPL + 1, 0);
    {
        // ptab[p] = first index with prefix >= p
        parallel_chunks(T, n_dist, 1 << 16, [&](uint64_t d0, uint64_t d1, int) {
            for (uint64_t d = d0; d < d1; d++) {
                uint64_t p = (uint64_t)(kmer[d] >> lshift);
                uint64_t q = d ? (uint64_t)(kmer[d - 1] >> lshift) + 1 : 0;
                for (uint64_t x = q; x <= p; x++) ptab[x] = (uint32_t)d;
            }
        });
        uint64_t lastp = n_dist ? (uint64_t)(kmer[n_dist - 1] >> lshift) + 1 : 0;
        for (uint64_t x = lastp; x <= NPL; x++) ptab[x] = (uint32_t)n_dist;
    }
    auto find = [&](K x) -> uint32_t {
        uint64_t p = (uint64_t)(x >> lshift);
        uint32_t lo = ptab[p], hi = ptab[p + 1];
        while (lo < hi) {
            uint32_t mid = lo + (hi - lo) / 2;
            if (kmer[mid] < x) lo = mid + 1;
            else hi = mid;
        }
        return (lo < n_dist && kmer[lo] == x) ? lo : NONE32;
    };
    auto one_bit = [](unsigned x) { return x && !(x & (x - 1)); };
    // ---- links: fwd[i] = j iff i -> j is an internal unitig link
    std::vector<uint32_t> fwd(n_dist + 1, NONE32);
    std::vector<uint8_t> is_target(n_dist + 1, 0);
    std::atomic<int> missing(0);
    parallel_chunks(T, n_dist, 1 << 15, [&](uint64_t d0, uint64_t d1, int) {
        for (uint64_t i = d0; i < d1; i++) {
            unsigned r = exts[i] & 0xf;
            if (!one_bit(r)) continue;
            unsigned b = (unsigned)__builtin_ctz(r);
            uint32_t j = find((K)(((kmer[i] << 2) | (K)b) & kmask));
            if (j == NONE32) { missing = 1; continue; }  // an observed neighbour must exist
            if (j == i) continue;                        // self loop (e.g. poly-A): a path of its own
            if (!one_bit(exts[j] >> 4)) continue;
            if (eq[j] != eq[i]) continue;
            fwd[i] = j;
            is_target[j] = 1;  // unique: j has exactly one left ext
        }
    });
    if (missing) { delete g; g_host_err = "k-mer neighbour missing (internal)"; return nullptr; }
    // ---- unitig heads
    std::vector<uint32_t> heads;
    {
        const uint64_t G = 1 << 16, nch = (n_dist + G - 1) / G;
        std::vector<uint64_t> hc(nch + 1, 0);
        parallel_chunks(T, nch, 1, [&](uint64_t c0, uint64_t c1, int) {
            for (uint64_t c = c0; c < c1; c++) {
                uint64_t n = 0;
                for (uint64_t i = c * G; i < std::min(n_dist, (c + 1) * G); i++) n += !is_target[i];
                hc[c + 1] = n;
            }
        });
        for (uint64_t c = 0; c < nch; c++) hc[c + 1] += hc[c];
        heads.resize(hc[nch]);
        parallel_chunks(T, nch, 1, [&](uint64_t c0, uint64_t c1, int) {
            for (uint64_t c = c0; c < c1; c++) {
                uint64_t o = hc[c];
                for (uint64_t i = c * G; i < std::min(n_dist, (c + 1) * G); i++)
                    if (!is_target[i]) heads[o++] = (uint32_t)i;
            }
        });
    }
    // ---- path lengths; every k-mer reachable from a head is marked
    std::vector<uint8_t> visited(n_dist + 1, 0);
    std::vector<uint32_t> plen(heads.size());
    parallel_chunks(T, heads.size(), 1 << 10, [&](uint64_t h0, uint64_t h1, int) {
        for (uint64_t h = h0; h < h1; h++) {
            uint32_t cur = heads[h], n = 0;
            while (cur != NONE32) { visited[cur] = 1; n++; cur = fwd[cur]; }
            plen[h] = n;
        }
    });
    // closed cycles: every member is a link target, so none was reached; cut at the smallest k-mer
    for (uint64_t s = 0; s < n_dist; s++) {
        if (visited[s]) continue;
        uint32_t cur = (uint32_t)s, n = 0;
        while (cur != NONE32 && !visited[cur]) { visited[cur] = 1; n++; cur = fwd[cur]; }
        heads.push_back((uint32_t)s);
        plen.push_back(n);
        g->n_cycles++;
    }
    // ---- emit nodes
    const uint64_t n_nodes = heads.size();
    g->node_start.resize(n_nodes);
    g->node_len.resize(n_nodes);
    g->node_exts.resize(n_nodes);
    g->node_eq.resize(n_nodes);
    uint64_t n_bases = 0;
    for (uint64_t i = 0; i < n_nodes; i++) {
        g->node_start[i] = n_bases;
        g->node_len[i] = k + plen[i] - 1;
        n_bases += g->node_len[i];
    }
    g->seq_words.assign((n_bases + 31) / 32 + 1, 0);
    uint64_t* W = g->seq_words.data();
    parallel_chunks(T, n_nodes, 1 << 10, [&](uint64_t n0, uint64_t n1, int) {
        for (uint64_t i = n0; i < n1; i++) {
            uint32_t cur = heads[i], last = cur;
            uint64_t pos = g->node_start[i];
            uint64_t acc = 0;  // bits accumulated for word pos/32
            auto put = [&](unsigned b) {
                acc |= (uint64_t)b << (62 - 2 * (pos & 31));
                pos++;
                if ((pos & 31) == 0) { __atomic_fetch_or(&W[(pos - 1) >> 5], acc, __ATOMIC_RELAXED); acc = 0; }
            };
            K first = kmer[cur];
            for (uint32_t t = 0; t < k; t++) put((unsigned)(first >> (2 * (k - 1 - t))) & 3u);
            uint32_t steps = plen[i];
            for (uint32_t q = 1; q < steps; q++) {
                cur = fwd[cur];
                put((unsigned)kmer[cur] & 3u);
                last = cur;
            }
            if (pos & 31) __atomic_fetch_or(&W[pos >> 5], acc, __ATOMIC_RELAXED);
            g->node_exts[i] = (uint8_t)((exts[heads[i]] & 0xf0) | (exts[last] & 0x0f));
            g->node_eq[i] = eq[heads[i]];
        }
    });
    g->seq_words.resize((n_bases + 31) / 32);
    return g;
}

}  // namespace

extern "C" psa_graph* psa_build_graph(const uint8_t* codes, const uint64_t* tx_off, uint32_t n_tx, uint32_t k, int threads) {
    if (!tx_off || (n_tx && !codes)) { g_host_err = "null argument"; return nullptr; }
    if (k < 2 || k > 64) { g_host_err = "k must be in 2..64"; return nullptr; }
    int T = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
    if (T < 1) T = 1;
    if (T > 256) T = 256;
    try {
        if (k <= 32) return build<uint64_t>(codes, tx_off, n_tx, k, T);
        return build<u128>(codes, tx_off, n_tx, k, T);
    } catch (const std::bad_alloc&) {
        g_host_err = "out of memory";
        return nullptr;
    }
}
extern "C" void psa_graph_free(psa_graph* g) { delete g; }
extern "C" uint32_t psa_graph_k(const psa_graph* g) { return g->k; }
extern "C" uint64_t psa_graph_n_nodes(const psa_graph* g) { return g->node_len.size(); }
extern "C" uint64_t psa_graph_n_kmers(const psa_graph* g) { return g->n_kmers; }
extern "C" uint64_t psa_graph_n_eq(const psa_graph* g) { return g->eq_offsets.size() - 1; }
extern "C" uint64_t psa_graph_n_seq_words(const psa_graph* g) { return g->seq_words.size(); }
extern "C" uint64_t psa_graph_n_cycles(const psa_graph* g) { return g->n_cycles; }
extern "C" const uint64_t* psa_graph_seq_words(const psa_graph* g) { return g->seq_words.data(); }
extern "C" const uint64_t* psa_graph_node_start(const psa_graph* g) { return g->node_start.data(); }
extern "C" const uint32_t* psa_graph_node_len(const psa_graph* g) { return g->node_len.data(); }
extern "C" const uint8_t* psa_graph_node_exts(const psa_graph* g) { return g->node_exts.data(); }
extern "C" const uint32_t* psa_graph_node_eq(const psa_graph* g) { return g->node_eq.data(); }
extern "C" const uint64_t* psa_graph_eq_offsets(const psa_graph* g) { return g->eq_offsets.data(); }
extern "C" const uint32_t* psa_graph_eq_members(const psa_graph* g) { return g->eq_members.data(); }
