// index_file.cpp -- the flat index on disk.
//
// The reference persists `Pseudoaligner<K>` as one bincode blob (ref src/utils.rs:22-43, used by the CLI
// at src/bin/pseudoaligner.rs:114,135).  That layout belongs to debruijn/boomphf internals that are not
// available here, and this project's index is rebuilt on the GPU from the graph anyway (include/psa.h),
// so what is stored is exactly the input of psa_index_create: the arrays of psa_index_desc, raw and
// 64-byte aligned behind a small versioned header, with a checksum.  Loading is one read per array.
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../../include/psa_host.h"

struct psa_graph {  // same definition as in build_graph.cpp
    uint32_t k = 0;
    uint64_t n_kmers = 0, n_cycles = 0;
    std::vector<uint64_t> seq_words, node_start, eq_offsets;
    std::vector<uint32_t> node_len, node_eq, eq_members;
    std::vector<uint8_t> node_exts;
};

extern thread_local std::string g_host_err;

namespace {

constexpr char kMagic[8] = {'P', 'S', 'A', 'I', 'D', 'X', '1', 0};
constexpr uint32_t kVersion = 1;

struct Header {
    char magic[8];
    uint32_t version, k;
    uint64_t n_nodes, n_seq_words, n_eq, n_eq_members, n_kmers, n_cycles;
    uint64_t checksum;
    uint64_t reserved[7];
};
static_assert(sizeof(Header) == 128, "header is two 64-byte lines");

uint64_t fold(uint64_t h, const void* p, size_t bytes) {  // order-sensitive 64-bit checksum
    const uint8_t* b = (const uint8_t*)p;
    size_t i = 0;
    for (; i + 8 <= bytes; i += 8) {
        uint64_t x;
        memcpy(&x, b + i, 8);
        h = (h ^ x) * 0x9E3779B97F4A7C15ULL;
        h ^= h >> 29;
    }
    uint64_t x = 0;
    if (i < bytes) memcpy(&x, b + i, bytes - i);
    h = (h ^ x ^ bytes) * 0xD6E8FEB86659FD93ULL;
    return h ^ (h >> 32);
}
uint64_t checksum_of(const psa_graph& g) {
    uint64_t h = 0x243F6A8885A308D3ULL ^ g.k;
    h = fold(h, g.seq_words.data(), g.seq_words.size() * 8);
    h = fold(h, g.node_start.data(), g.node_start.size() * 8);
    h = fold(h, g.node_len.data(), g.node_len.size() * 4);
    h = fold(h, g.node_exts.data(), g.node_exts.size());
    h = fold(h, g.node_eq.data(), g.node_eq.size() * 4);
    h = fold(h, g.eq_offsets.data(), g.eq_offsets.size() * 8);
    h = fold(h, g.eq_members.data(), g.eq_members.size() * 4);
    return h;
}
bool put(FILE* f, const void* p, size_t bytes) {
    static const char zeros[64] = {0};
    if (bytes && fwrite(p, 1, bytes, f) != bytes) return false;
    const size_t pad = (64 - bytes % 64) % 64;
    return !pad || fwrite(zeros, 1, pad, f) == pad;
}
template <class T>
bool get(FILE* f, std::vector<T>& v, uint64_t n) {
    v.resize(n);
    const size_t bytes = n * sizeof(T);
    if (bytes && fread(v.data(), 1, bytes, f) != bytes) return false;
    const size_t pad = (64 - bytes % 64) % 64;
    return !pad || fseek(f, (long)pad, SEEK_CUR) == 0;
}

}  // namespace

extern "C" psa_graph* psa_graph_from_arrays(uint32_t k, uint64_t n_nodes, const uint64_t* seq_words, uint64_t n_seq_words,
                                            const uint64_t* node_start, const uint32_t* node_len, const uint8_t* node_exts,
                                            const uint32_t* node_eq, uint64_t n_eq, const uint64_t* eq_offsets,
                                            const uint32_t* eq_members) {
    if (!eq_offsets || (n_nodes && (!seq_words || !node_start || !node_len || !node_exts || !node_eq)) ||
        (eq_offsets[n_eq] && !eq_members) || (n_seq_words && !seq_words)) {
        g_host_err = "psa_graph_from_arrays: null array";
        return nullptr;
    }
    for (uint64_t c = 0; c < n_eq; c++)
        if (eq_offsets[c + 1] < eq_offsets[c]) {
            g_host_err = "psa_graph_from_arrays: eq_offsets not monotone";
            return nullptr;
        }
    psa_graph* g = nullptr;
    try {
        g = new psa_graph();
    } catch (...) {
        g_host_err = "psa_graph_from_arrays: out of memory";
        return nullptr;
    }
    try {
    g->k = k;
    g->seq_words.assign(seq_words, seq_words + n_seq_words);
    g->node_start.assign(node_start, node_start + n_nodes);
    g->node_len.assign(node_len, node_len + n_nodes);
    g->node_exts.assign(node_exts, node_exts + n_nodes);
    g->node_eq.assign(node_eq, node_eq + n_nodes);
    g->eq_offsets.assign(eq_offsets, eq_offsets + n_eq + 1);
    g->eq_members.assign(eq_members, eq_members + eq_offsets[n_eq]);
    for (uint64_t i = 0; i < n_nodes; i++) g->n_kmers += node_len[i] >= k ? node_len[i] - k + 1 : 0;
    } catch (...) {   // nothing may be thrown through the C boundary
        delete g;
        g_host_err = "psa_graph_from_arrays: out of memory";
        return nullptr;
    }
    return g;
}

extern "C" int psa_graph_save(const psa_graph* g, const char* path) {
    if (!g || !path) { g_host_err = "psa_graph_save: null argument"; return -1; }
    FILE* f = fopen(path, "wb");
    if (!f) { g_host_err = std::string("cannot create ") + path; return -7; }
    Header h{};
    memcpy(h.magic, kMagic, 8);
    h.version = kVersion;
    h.k = g->k;
    h.n_nodes = g->node_len.size();
    h.n_seq_words = g->seq_words.size();
    h.n_eq = g->eq_offsets.size() - 1;
    h.n_eq_members = g->eq_members.size();
    h.n_kmers = g->n_kmers;
    h.n_cycles = g->n_cycles;
    h.checksum = checksum_of(*g);
    bool ok = fwrite(&h, sizeof h, 1, f) == 1 && put(f, g->seq_words.data(), g->seq_words.size() * 8) &&
              put(f, g->node_start.data(), g->node_start.size() * 8) && put(f, g->node_len.data(), g->node_len.size() * 4) &&
              put(f, g->node_exts.data(), g->node_exts.size()) && put(f, g->node_eq.data(), g->node_eq.size() * 4) &&
              put(f, g->eq_offsets.data(), g->eq_offsets.size() * 8) && put(f, g->eq_members.data(), g->eq_members.size() * 4);
    ok = (fclose(f) == 0) && ok;
    if (!ok) { g_host_err = std::string("write error on ") + path; return -7; }
    return 0;
}

// what the arrays of a header occupy in the file (each padded to 64 bytes); false if a count is absurd
static bool payload_bytes(const Header& h, uint64_t& total) {
    const uint64_t lim = 1ull << 56;  // no count this large fits a file: keeps the sums below from wrapping
    if (h.n_nodes > lim || h.n_seq_words > lim || h.n_eq >= lim || h.n_eq_members > lim) return false;
    auto padded = [](uint64_t bytes) { return (bytes + 63) / 64 * 64; };
    total = padded(h.n_seq_words * 8) + padded(h.n_nodes * 8) + padded(h.n_nodes * 4) + padded(h.n_nodes) + padded(h.n_nodes * 4) +
            padded((h.n_eq + 1) * 8) + padded(h.n_eq_members * 4);
    return true;
}

extern "C" psa_graph* psa_graph_load(const char* path) {
    if (!path) { g_host_err = "psa_graph_load: null path"; return nullptr; }
    FILE* f = fopen(path, "rb");
    if (!f) { g_host_err = std::string("cannot open ") + path; return nullptr; }
    fseek(f, 0, SEEK_END);
    const uint64_t file_size = (uint64_t)ftell(f);
    rewind(f);
    Header h{};
    psa_graph* g = nullptr;
    const char* why = nullptr;
    uint64_t need = 0;
    try {
        if (fread(&h, sizeof h, 1, f) != 1 || memcmp(h.magic, kMagic, 8) != 0) why = "not a psa index file";
        else if (h.version != kVersion) why = "unsupported index file version";
        else if (h.k < 2 || h.k > 64) why = "corrupt header";
        // the header's counts are believed only as far as the file is long: nothing is allocated before this
        else if (!payload_bytes(h, need) || need > file_size - sizeof h) why = "header counts exceed the file size";
        else {
            g = new psa_graph();
            g->k = h.k;
            g->n_kmers = h.n_kmers;
            g->n_cycles = h.n_cycles;
            if (!get(f, g->seq_words, h.n_seq_words) || !get(f, g->node_start, h.n_nodes) || !get(f, g->node_len, h.n_nodes) ||
                !get(f, g->node_exts, h.n_nodes) || !get(f, g->node_eq, h.n_nodes) || !get(f, g->eq_offsets, h.n_eq + 1) ||
                !get(f, g->eq_members, h.n_eq_members))
                why = "truncated index file";
            else if (g->eq_offsets[h.n_eq] != h.n_eq_members || checksum_of(*g) != h.checksum)
                why = "index file checksum mismatch";
            else {
                // the arrays are what psa_index_create and the graph accessors index with: check them here
                uint64_t nk = 0;
                if (g->eq_offsets[0] != 0) why = "eq_offsets does not start at 0";
                for (uint64_t c = 0; c < h.n_eq && !why; c++)
                    if (g->eq_offsets[c + 1] < g->eq_offsets[c]) why = "eq_offsets not monotone";
                for (uint64_t i = 0; i < h.n_nodes && !why; i++) {
                    if (g->node_len[i] < h.k) why = "node shorter than k";
                    else if (g->node_eq[i] >= h.n_eq) why = "node class id out of range";
                    else if (g->node_start[i] > h.n_seq_words * 32 || g->node_len[i] > h.n_seq_words * 32 - g->node_start[i])
                        why = "node outside the sequence";
                    else nk += g->node_len[i] - h.k + 1;
                }
                if (!why && nk != h.n_kmers) why = "k-mer count does not match the nodes";
            }
        }
    } catch (...) {   // (std::bad_alloc / std::length_error: nothing may be thrown through the C boundary)
        why = "out of memory";
    }
    fclose(f);
    if (why) {
        delete g;
        g_host_err = std::string(why) + ": " + path;
        return nullptr;
    }
    return g;
}

// ---------------------------------------------------------------------------------------------
// Reader for the reference's own index file: `Pseudoaligner<K>` written by bincode 1.3 with default
// options (ref src/utils.rs:22-32: fixed-width little-endian integers, u64 length prefixes, struct
// fields in declaration order, usize as u64, bool as one byte).  Only the first two fields are read --
// `dbg` and `eq_classes` (ref src/pseudoaligner.rs:27-29); `dbg_index` (boomphf's bit-vectors),
// `tx_names` and `tx_gene_mapping` that follow are not needed, because the dictionary is rebuilt on the
// GPU.  The field order inside `dbg` is that of debruijn 0.3.4 @ 8d9a5c5 AS RECALLED (the crate is not
// available in this environment, so this has not been checked against a reference-built file):
//   DebruijnGraph { base: BaseGraph { sequences: PackedDnaStringSet { sequence: DnaString { storage:
//   Vec<u64>, len: usize }, start: Vec<usize>, length: Vec<u32> }, exts: Vec<Exts{val:u8}>, data: Vec<u32>,
//   stranded: bool }, left_order: Vec<u32>, right_order: Vec<u32> }
// Every length is cross-checked, so a different layout is refused rather than misread.  K is a type
// parameter of the reference (not stored): the caller passes k, as the CLI's -k does.
// ---------------------------------------------------------------------------------------------
namespace {
struct Cursor {
    FILE* f;
    uint64_t left;  // bytes left in the file
    bool u64(uint64_t& v) {
        if (left < 8 || fread(&v, 8, 1, f) != 1) return false;
        left -= 8;
        return true;
    }
    template <class T>
    bool vec(std::vector<T>& v, uint64_t max_elems) {
        uint64_t n;
        if (!u64(n) || n > max_elems || n * sizeof(T) > left) return false;
        v.resize(n);
        if (n && fread(v.data(), sizeof(T), n, f) != n) return false;
        left -= n * sizeof(T);
        return true;
    }
};
}  // namespace

extern "C" psa_graph* psa_graph_load_bincode(const char* path, uint32_t k) {
    if (!path || k < 2 || k > 64) { g_host_err = "psa_graph_load_bincode: bad argument"; return nullptr; }
    FILE* f = fopen(path, "rb");
    if (!f) { g_host_err = std::string("cannot open ") + path; return nullptr; }
    fseek(f, 0, SEEK_END);
    Cursor c{f, (uint64_t)ftell(f)};
    rewind(f);
    psa_graph* g = nullptr;
    const char* why = nullptr;
    try {
    g = new psa_graph();
    g->k = k;
    const uint64_t cap = c.left;  // no array can have more elements than the file has bytes
    uint64_t n_bases = 0;
    std::vector<uint32_t> left_order, right_order;
    uint8_t stranded = 0;
    uint64_t n_classes = 0;
    if (!c.vec(g->seq_words, cap) || !c.u64(n_bases)) why = "truncated (sequence)";
    else if (g->seq_words.size() != (n_bases + 31) / 32) why = "DnaString storage does not match its length";
    else if (!c.vec(g->node_start, cap) || !c.vec(g->node_len, cap) || !c.vec(g->node_exts, cap) || !c.vec(g->node_eq, cap))
        why = "truncated (node arrays)";
    else if (g->node_start.size() != g->node_len.size() || g->node_exts.size() != g->node_len.size() ||
             g->node_eq.size() != g->node_len.size())
        why = "node arrays of different lengths";
    else if (c.left < 1 || fread(&stranded, 1, 1, f) != 1) why = "truncated (stranded)";
    else {
        c.left -= 1;
        if (stranded != 1) why = "graph is not stranded (the reference builds with STRANDED = true, src/config.rs:14)";
        else if (!c.vec(left_order, cap) || !c.vec(right_order, cap)) why = "truncated (node orders)";
        else if (left_order.size() != g->node_len.size() || right_order.size() != g->node_len.size())
            why = "node order arrays do not match the node count";
        else if (!c.u64(n_classes) || n_classes > cap) why = "truncated (eq_classes)";
    }
    if (!why) {
        g->eq_offsets.assign(1, 0);
        std::vector<uint32_t> cls;
        for (uint64_t i = 0; i < n_classes && !why; i++) {
            if (!c.vec(cls, cap)) { why = "truncated (eq_classes)"; break; }
            for (size_t j = 1; j < cls.size(); j++)
                if (cls[j] <= cls[j - 1]) { why = "equivalence class not ascending"; break; }
            g->eq_members.insert(g->eq_members.end(), cls.begin(), cls.end());
            g->eq_offsets.push_back(g->eq_members.size());
        }
    }
    if (!why) {
        for (size_t i = 0; i < g->node_len.size() && !why; i++) {
            if (g->node_len[i] < k) why = "node shorter than k (wrong k?)";
            else if (g->node_start[i] + g->node_len[i] > n_bases) why = "node outside the sequence";
            else if (g->node_eq[i] >= n_classes) why = "node class id out of range";
            else g->n_kmers += g->node_len[i] - k + 1;
        }
    }
    } catch (...) {
        why = "out of memory";
    }
    fclose(f);
    if (why) {
        delete g;
        g_host_err = std::string("not a recognised Pseudoaligner bincode file (") + why + "): " + path;
        return nullptr;
    }
    return g;
}
