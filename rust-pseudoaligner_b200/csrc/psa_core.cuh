// psa_core.cuh -- arithmetic of the pseudoalignment hot path, shared by every kernel.
//
// Everything here is scalar __host__ __device__ code: 2-bit sequence access, the k-mer hash,
// the sector-block minimal perfect hash probe, the `values` packing, the k-mer verification,
// the per-word mismatch masks of the two extension loops and the map_read state machine
// (templated on a "warp" policy that supplies the cooperative steps).  The CUDA kernels in
// psa_kernels.cu instantiate it with warp-shuffle policies; tests/hostsim instantiates the
// same text with a serial policy so that the state machine and the hash layout are checked
// against the oracle on machines without a GPU.  That host instantiation is a unit-test
// harness only -- the product library (psa_api.cu) has no CPU path.
//
// Reference being restated: 10XGenomics/rust-pseudoaligner @ 9d9cab8
//   src/pseudoaligner.rs:64-319   map_read_to_nodes_with_mismatch   -> map_read_nodes()
//   src/pseudoaligner.rs:91-114   find_kmer_match                   -> W::find_seed + dict_get()
//   src/pseudoaligner.rs:99-107   MPHF answer verification          -> dict_get()
//   src/config.rs:16-18           constants
// and, from the un-vendored crates (published algorithms, see oracle/psa_oracle.h):
//   debruijn DnaString packing / get_kmer, boomphf Mphf::try_hash (structure only: cascaded
//   bit-vectors + rank; hash function, block layout and fingerprints are this project's own).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define PSA_HD __host__ __device__ __forceinline__
#else
#define PSA_HD inline
#endif

namespace psa {

constexpr uint32_t kNone = 0xFFFFFFFFu;
constexpr int kMaxLevels = 48;        // MPHF cascade depth cap (1e8 keys need ~25 at gamma 1.7)
constexpr uint32_t kBlockBits = 192;  // 3 data words per 32-byte block; word 0 is the rank header
constexpr uint32_t kSeedStride = 3;   // ref src/pseudoaligner.rs:110
constexpr uint32_t kCoverageThreshold = 32;  // ref src/config.rs:16
constexpr double kLeftExtendFraction = 0.2;  // ref src/config.rs:17

// ---------------------------------------------------------------------------------------------
// bit helpers
// ---------------------------------------------------------------------------------------------
PSA_HD int popc64(uint64_t x) {
#ifdef __CUDA_ARCH__
    return __popcll(x);
#else
    return __builtin_popcountll(x);
#endif
}
PSA_HD int ctz64(uint64_t x) {  // x != 0
#ifdef __CUDA_ARCH__
    return __ffsll((long long)x) - 1;
#else
    return __builtin_ctzll(x);
#endif
}
PSA_HD int clz64(uint64_t x) {  // x != 0
#ifdef __CUDA_ARCH__
    return __clzll((long long)x);
#else
    return __builtin_clzll(x);
#endif
}
PSA_HD uint64_t mulhi64(uint64_t a, uint64_t b) {
#ifdef __CUDA_ARCH__
    return __umul64hi(a, b);
#else
    return (uint64_t)(((unsigned __int128)a * b) >> 64);
#endif
}
PSA_HD uint64_t mix64(uint64_t x) {  // murmur3 finaliser (a bijection on 64 bits)
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL;
    x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL;
    x ^= x >> 33;
    return x;
}

// ---------------------------------------------------------------------------------------------
// word loaders: the same text reads the index through the read-only path on the device
// ---------------------------------------------------------------------------------------------
struct GLoad {  // immutable index memory (global, ld.global.nc)
    const uint64_t* p;
    PSA_HD uint64_t operator()(uint64_t i) const {
#ifdef __CUDA_ARCH__
        return __ldg(p + i);
#else
        return p[i];
#endif
    }
};
struct PLoad {  // plain pointer (shared memory tile or global read buffer)
    const uint64_t* p;
    PSA_HD uint64_t operator()(uint64_t i) const { return p[i]; }
};

// ---------------------------------------------------------------------------------------------
// 2-bit sequences: debruijn DnaString packing, base i in word i/32 at bits 62-2*(i%32)
// (ref call sites src/pseudoaligner.rs:93,103,156,182,241,265)
// ---------------------------------------------------------------------------------------------
template <class L>
PSA_HD uint32_t seq_get(L ld, uint64_t i) {  // DnaString::get
    return (uint32_t)(ld(i >> 5) >> (62 - 2 * (i & 31))) & 3u;
}
// n (1..32) bases from base pos, right-aligned: first base in the most significant used bits.
// Never touches a word that holds none of the requested bases.
template <class L>
PSA_HD uint64_t seq_bits(L ld, uint64_t pos, uint32_t n) {
    uint64_t wi = pos >> 5;
    uint32_t in_word = (uint32_t)(pos & 31);
    uint64_t v = ld(wi) << (2 * in_word);
    if (n > 32 - in_word) v |= ld(wi + 1) >> (64 - 2 * in_word);
    return v >> (64 - 2 * n);
}

// k-mer integer (DnaString::get_kmer): base 0 most significant, low 2k bits used.
template <int KW>
struct Kmer;
template <>
struct Kmer<1> {
    uint64_t lo;
    PSA_HD bool operator==(const Kmer& o) const { return lo == o.lo; }
};
template <>
struct Kmer<2> {
    uint64_t lo, hi;
    PSA_HD bool operator==(const Kmer& o) const { return lo == o.lo && hi == o.hi; }
};

template <class L>
PSA_HD Kmer<1> get_kmer1(L ld, uint64_t pos, uint32_t k) {
    Kmer<1> r;
    r.lo = seq_bits(ld, pos, k);
    return r;
}
template <class L>
PSA_HD Kmer<2> get_kmer2(L ld, uint64_t pos, uint32_t k) {
    Kmer<2> r;
    if (k <= 32) {
        r.hi = 0;
        r.lo = seq_bits(ld, pos, k);
    } else {
        r.hi = seq_bits(ld, pos, k - 32);
        r.lo = seq_bits(ld, pos + (k - 32), 32);
    }
    return r;
}
template <int KW>
struct KmerOps;
template <>
struct KmerOps<1> {
    template <class L>
    static PSA_HD Kmer<1> get(L ld, uint64_t pos, uint32_t k) { return get_kmer1(ld, pos, k); }
    static PSA_HD uint64_t fold(Kmer<1> x) { return x.lo; }
    // successor / predecessor k-mers for the edge tables (debruijn Kmer::extend_right/left)
    static PSA_HD Kmer<1> extend_right(Kmer<1> x, uint32_t b, uint32_t k) {
        Kmer<1> r;
        uint64_t mask = k == 32 ? ~0ULL : ((1ULL << (2 * k)) - 1);
        r.lo = ((x.lo << 2) | b) & mask;
        return r;
    }
    static PSA_HD Kmer<1> extend_left(Kmer<1> x, uint32_t b, uint32_t k) {
        Kmer<1> r;
        r.lo = (x.lo >> 2) | ((uint64_t)b << (2 * (k - 1)));
        return r;
    }
};
template <>
struct KmerOps<2> {
    template <class L>
    static PSA_HD Kmer<2> get(L ld, uint64_t pos, uint32_t k) { return get_kmer2(ld, pos, k); }
    static PSA_HD uint64_t fold(Kmer<2> x) { return x.lo ^ mix64(x.hi + 0x9e3779b97f4a7c15ULL); }
    static PSA_HD Kmer<2> extend_right(Kmer<2> x, uint32_t b, uint32_t k) {
        Kmer<2> r;
        r.hi = (x.hi << 2) | (x.lo >> 62);
        r.lo = (x.lo << 2) | b;
        if (k <= 32) {
            r.hi = 0;
            r.lo &= (k == 32 ? ~0ULL : ((1ULL << (2 * k)) - 1));
        } else {
            r.hi &= (k == 64 ? ~0ULL : ((1ULL << (2 * (k - 32))) - 1));
        }
        return r;
    }
    static PSA_HD Kmer<2> extend_left(Kmer<2> x, uint32_t b, uint32_t k) {
        Kmer<2> r;
        r.lo = (x.lo >> 2) | (x.hi << 62);
        r.hi = x.hi >> 2;
        if (k <= 32) r.lo |= (uint64_t)b << (2 * (k - 1));
        else r.hi |= (uint64_t)b << (2 * (k - 33));
        return r;
    }
};

// ---------------------------------------------------------------------------------------------
// The index as the kernels see it (all pointers device memory, immutable)
// ---------------------------------------------------------------------------------------------
struct NodeRec {        // 64 bytes, 64-byte aligned: one L2 line per node visit
    uint64_t start;     // first base in the concatenated sequence
    uint32_t len;       // bases
    uint32_t eq;        // equivalence-class id (*node.data())
    uint32_t class_len; // |eq_classes[eq]|
    uint32_t exts;      // debruijn Exts byte
    uint64_t class_off; // eq_classes[eq] starts at eq_mem[class_off]
    uint32_t succ[4];   // node reached by right extension b (kNone if the ext bit is clear)
    uint32_t pred[4];   // node reached by left extension b
};
static_assert(sizeof(NodeRec) == 64, "NodeRec must be one 64-byte line");

struct Mphf {
    const uint64_t* blocks;  // 4 words per block: [rank:48 | c1:7 | c2:8] w1 w2 w3
    uint32_t n_levels;
    uint32_t pad;
    uint64_t level_nblk[kMaxLevels];
    uint64_t level_base[kMaxLevels];  // first block of the level
};

struct DevIndex {
    uint32_t k;
    uint32_t node_bits, off_bits, fp_bits;  // `values` entry = node | off << node_bits | fp << (node_bits+off_bits)
    uint64_t n_nodes, n_kmers, n_eq;
    const uint64_t* values;
    const NodeRec* nodes;
    const uint64_t* seq;
    const uint64_t* eq_off;
    const uint32_t* eq_mem;
    Mphf mphf;
};

// Two 64-bit hashes per k-mer, computed once; level l probes h1 + l*h2 (double hashing, as
// in BBHash), the fingerprint is the top bits of h2.
struct KeyHash {
    uint64_t h1, h2;
};
PSA_HD KeyHash make_hash(uint64_t folded) {
    KeyHash kh;
    kh.h1 = mix64(folded + 0x9E3779B97F4A7C15ULL);
    kh.h2 = mix64(folded ^ 0xD6E8FEB86659FD93ULL) | 1ULL;
    return kh;
}
PSA_HD uint64_t level_hash(KeyHash kh, uint32_t lvl) { return kh.h1 + (uint64_t)lvl * kh.h2; }
PSA_HD uint64_t fp_of(KeyHash kh, uint32_t fp_bits) { return kh.h2 >> (64 - fp_bits); }

// position of a key at a level: (block within level, bit 0..191 within block); nblk < 2^32
PSA_HD void level_pos(uint64_t h, uint64_t nblk, uint64_t& blk, uint32_t& bit) {
    blk = ((h >> 32) * (uint64_t)(uint32_t)nblk) >> 32;
    bit = (uint32_t)(((h & 0xffffffffULL) * kBlockBits) >> 32);
}

struct Block {
    uint64_t w[4];
};
PSA_HD Block load_block(const uint64_t* blocks, uint64_t b) {
    Block r;
#ifdef __CUDA_ARCH__
    // one 32-byte sector per probe: header + 3 bit-vector words (LDG.E.256)
    asm("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];"
        : "=l"(r.w[0]), "=l"(r.w[1]), "=l"(r.w[2]), "=l"(r.w[3])
        : "l"(blocks + 4 * b));
#else
    for (int i = 0; i < 4; i++) r.w[i] = blocks[4 * b + i];
#endif
    return r;
}
constexpr uint64_t kRankMask = (1ULL << 48) - 1;
PSA_HD uint64_t make_header(uint64_t rank, uint32_t c1, uint32_t c2) {
    return rank | ((uint64_t)c1 << 48) | ((uint64_t)c2 << 55);
}
// rank of bit `bit` of a block whose bit is set = number of set bits before it in the cascade
PSA_HD uint64_t block_rank(const Block& b, uint32_t bit) {
    uint32_t wi = bit >> 6, bi = bit & 63;
    uint64_t hdr = b.w[0];
    uint64_t r = hdr & kRankMask;
    if (wi == 1) r += (hdr >> 48) & 0x7f;
    else if (wi == 2) r += (hdr >> 55) & 0xff;
    r += (uint64_t)popc64(b.w[1 + wi] & ((1ULL << bi) - 1));
    return r;
}

struct ProbeStats {  // sequential-equivalent event counts of one dictionary probe
    uint32_t levels, hit, verified;
};

// Mphf::try_hash: cascade of bit-vectors; the first level whose bit is set gives the slot.
PSA_HD bool mphf_lookup(const Mphf& m, KeyHash hk, uint64_t& slot, uint32_t& levels) {
    levels = 0;
    for (uint32_t lvl = 0; lvl < m.n_levels; lvl++) {
        uint64_t blk;
        uint32_t bit;
        level_pos(level_hash(hk, lvl), m.level_nblk[lvl], blk, bit);
        Block b = load_block(m.blocks, m.level_base[lvl] + blk);
        levels++;
        if ((b.w[1 + (bit >> 6)] >> (bit & 63)) & 1) {
            slot = block_rank(b, bit);
            return true;
        }
    }
    return false;
}

PSA_HD uint64_t pack_value(const DevIndex& ix, uint32_t node, uint32_t off, KeyHash hk) {
    uint64_t v = (uint64_t)node | ((uint64_t)off << ix.node_bits);
    if (ix.fp_bits) v |= fp_of(hk, ix.fp_bits) << (ix.node_bits + ix.off_bits);
    return v;
}

// dbg_index.get(kmer) followed by the reference's verification (src/pseudoaligner.rs:96-107).
// A fingerprint mismatch proves the slot's key differs from `key`, so skipping the unitig
// fetch cannot change the outcome of the reference's `read_kmer == ref_kmer` test.
template <int KW>
PSA_HD bool dict_get(const DevIndex& ix, Kmer<KW> key, uint32_t& node, uint32_t& off, ProbeStats* st) {
    KeyHash hk = make_hash(KmerOps<KW>::fold(key));
    uint64_t slot;
    uint32_t levels;
    bool in = mphf_lookup(ix.mphf, hk, slot, levels);
    if (st) { st->levels = levels; st->hit = in; st->verified = 0; }
    if (!in) return false;
#ifdef __CUDA_ARCH__
    uint64_t v = __ldg(ix.values + slot);
#else
    uint64_t v = ix.values[slot];
#endif
    if (ix.fp_bits) {
        if ((v >> (ix.node_bits + ix.off_bits)) != fp_of(hk, ix.fp_bits)) return false;
    }
    uint32_t n = (uint32_t)(v & ((1ULL << ix.node_bits) - 1));
    uint32_t o = (uint32_t)((v >> ix.node_bits) & ((1ULL << ix.off_bits) - 1));
    if (st) st->verified = 1;
#ifdef __CUDA_ARCH__
    uint64_t start = __ldg(&ix.nodes[n].start);
#else
    uint64_t start = ix.nodes[n].start;
#endif
    Kmer<KW> ref = KmerOps<KW>::get(GLoad{ix.seq}, start + o, ix.k);
    if (!(ref == key)) return false;
    node = n;
    off = o;
    return true;
}

// ---------------------------------------------------------------------------------------------
// extension compares (ref src/pseudoaligner.rs:151-170 and :236-255), 32 bases per word.
// Both return a mask with bit 2*t set iff the t-th compared base (in the order the
// reference's loop visits them) mismatches, for t < n <= 32.
// ---------------------------------------------------------------------------------------------
PSA_HD uint64_t rev_pairs(uint64_t x) {  // reverse the order of the 32 2-bit fields
#ifdef __CUDA_ARCH__
    x = __brevll(x);
    return ((x >> 1) & 0x5555555555555555ULL) | ((x & 0x5555555555555555ULL) << 1);
#else
    x = ((x >> 2) & 0x3333333333333333ULL) | ((x & 0x3333333333333333ULL) << 2);
    x = ((x >> 4) & 0x0F0F0F0F0F0F0F0FULL) | ((x & 0x0F0F0F0F0F0F0F0FULL) << 4);
    return __builtin_bswap64(x);
#endif
}
PSA_HD uint64_t fold_pairs(uint64_t x) { return (x | (x >> 1)) & 0x5555555555555555ULL; }

// forward: t-th base = read[rpos+t] vs ref[spos+t]
template <class LR, class LS>
PSA_HD uint64_t mismatch_fwd(LR rd, uint64_t rpos, LS sq, uint64_t spos, uint32_t n) {
    uint64_t x = seq_bits(rd, rpos, n) ^ seq_bits(sq, spos, n);  // base t at bits 2(n-1-t)
    x = rev_pairs(x) >> (64 - 2 * n);                            // base t at bits 2t
    return fold_pairs(x);
}
// backward: t-th base = read[rend-t] vs ref[send-t]
template <class LR, class LS>
PSA_HD uint64_t mismatch_bwd(LR rd, uint64_t rend, LS sq, uint64_t send, uint32_t n) {
    uint64_t x = seq_bits(rd, rend + 1 - n, n) ^ seq_bits(sq, send + 1 - n, n);  // base t at bits 2t
    return fold_pairs(x);
}
// index t of the j-th (1-based) set bit of a fold_pairs mask
PSA_HD uint32_t nth_mismatch(uint64_t m, uint32_t j) {
    for (uint32_t i = 1; i < j; i++) m &= m - 1;
    return (uint32_t)ctz64(m) >> 1;
}

// ---------------------------------------------------------------------------------------------
// map_read_to_nodes_with_mismatch, ref src/pseudoaligner.rs:64-319.
//
// The control flow is the reference's, statement by statement; the three inner loops are
// delegated to the warp policy W, whose results are warp-uniform:
//   W::find_seed(kmer_pos, last, node, off)      :91-114  first stride-3 position >= kmer_pos whose k-mer
//                                                         is in the graph; leaves kmer_pos as the loop would
//   W::node(id)                                  get_node: NodeRec fields
//   W::cmp_fwd(rpos, spos, m, A, premature)      :236-255 -> matched_bases
//   W::cmp_bwd(rend, send, m, A, premature)      :151-170 -> matched_bases
//   W::succ(id, b) / W::pred(id, b)              r_edges()[..].0 / l_edges()[..].0
//   W::push(node)                                nodes.push
// Returns false for None.  read_coverage is returned through `coverage`.
// ---------------------------------------------------------------------------------------------
struct NodeView {
    uint64_t start;
    uint32_t len, eq, class_len, exts;
    uint64_t class_off;
};

#ifdef __CUDACC__
#pragma nv_exec_check_disable
#endif
template <class W>
PSA_HD bool map_read_nodes(W& w, uint32_t k, uint64_t read_length, uint32_t allowed_mismatches,
                           uint32_t& coverage) {
    uint64_t read_coverage = 0;                                                 // :71
    const uint64_t kmer_length = k;                                             // :80
    uint64_t left_extend_threshold = (uint64_t)(kLeftExtendFraction * (double)read_length);  // :77
    uint64_t kmer_pos = 0;                                                      // :79
    if (read_length < kmer_length) return false;                                // :82-84
    const uint64_t last_kmer_pos = read_length - kmer_length;                   // :86
    uint32_t n_pushed = 0;

    uint32_t node_id = 0, kmer_offset = 0;
    bool have = w.find_seed(kmer_pos, last_kmer_pos, node_id, kmer_offset);     // :118-121

    // left extension, :124-205.  (kmer_pos >= 1 only fails for read_length < 5, where the
    // reference's `kmer_pos - 1` would underflow; unreachable for k >= 5.)
    if (have && kmer_pos >= left_extend_threshold && kmer_pos >= 1) {
        uint64_t last_pos = kmer_pos - 1;                                       // :127
        uint32_t prev_node_id = node_id;                                        // :128
        uint64_t prev_kmer_offset = kmer_offset > 0 ? kmer_offset - 1 : 0;      // :129 (sic)
        for (;;) {                                                              // :131
            NodeView nv = w.node(prev_node_id);                                 // :132
            uint64_t skipped_read = last_pos + 1;                               // :139
            uint64_t skipped_ref = prev_kmer_offset + 1;                        // :142
            uint64_t max_matchable_pos = skipped_read < skipped_ref ? skipped_read : skipped_ref;  // :145
            bool premature_break = false;                                       // :148
            uint64_t matched_bases =                                            // :149-170
                w.cmp_bwd(last_pos, nv.start + prev_kmer_offset, max_matchable_pos, allowed_mismatches,
                          premature_break);
            read_coverage += matched_bases;                                     // :169
            if (last_pos + 1 - matched_bases == 0 || premature_break) break;    // :173-175
            last_pos -= matched_bases;                                          // :178
            uint32_t next_base = w.read_base(last_pos);                         // :182
            if ((nv.exts >> (4 + next_base)) & 1) {                             // :183
                prev_node_id = w.pred(prev_node_id, next_base);                 // :185-194
                NodeView pv = w.node(prev_node_id);                             // :195
                prev_kmer_offset = pv.len - kmer_length;                        // :196
                w.push(prev_node_id, pv);                                       // :199
                n_pushed++;
            } else {
                break;                                                          // :201
            }
        }
    }

    // forward search, :208-302
    if (kmer_pos <= last_kmer_pos) {                                            // :208
        for (;;) {                                                              // :209
            NodeView nv = w.node(node_id);                                      // :210
            kmer_pos += kmer_length;                                            // :215
            read_coverage += kmer_length;                                       // :216
            w.push(node_id, nv);                                                // :219
            n_pushed++;
            uint64_t remaining_read = read_length - kmer_pos;                   // :222
            uint64_t ref_length = nv.len;                                       // :226
            uint64_t ref_offset = kmer_offset + kmer_length;                    // :227
            uint64_t informative_ref = ref_length - ref_offset;                 // :228
            uint64_t max_matchable_pos = remaining_read < informative_ref ? remaining_read : informative_ref;  // :231
            bool premature_break = false;                                       // :233
            uint64_t matched_bases =                                            // :234-255
                w.cmp_fwd(kmer_pos, nv.start + ref_offset, max_matchable_pos, allowed_mismatches,
                          premature_break);
            read_coverage += matched_bases;                                     // :254
            kmer_pos += matched_bases;                                          // :257
            if (kmer_pos >= read_length) break;                                 // :259-261
            uint32_t next_base = w.read_base(kmer_pos);                         // :265
            if (!premature_break && ((nv.exts >> next_base) & 1)) {             // :267
                node_id = w.succ(node_id, next_base);                           // :269-278
                kmer_offset = 0;                                                // :279
                kmer_pos -= kmer_length - 1;                                    // :282
                read_coverage -= kmer_length - 1;                               // :283
            } else {
                if (kmer_pos > last_kmer_pos) break;                            // :287-290
                if (!w.find_seed(kmer_pos, last_kmer_pos, node_id, kmer_offset)) break;  // :293-299
            }
        }
    }
    if (n_pushed == 0) return false;                                            // :305-314
    coverage = (uint32_t)read_coverage;
    return true;                                                                // :317
}

// Rust slice::binary_search on an ascending slice (used by intersect, ref :404)
template <class T>
PSA_HD bool contains_sorted(const T* v, uint64_t n, T x) {
    uint64_t lo = 0, hi = n;
    while (lo < hi) {
        uint64_t mid = lo + ((hi - lo) >> 1);
#ifdef __CUDA_ARCH__
        T y = __ldg(v + mid);
#else
        T y = v[mid];
#endif
        if (y < x) lo = mid + 1;
        else hi = mid;
    }
    if (lo >= n) return false;
#ifdef __CUDA_ARCH__
    return __ldg(v + lo) == x;
#else
    return v[lo] == x;
#endif
}

// ASCII -> 2-bit code, DnaString::from_dna_string (call site ref src/pseudoaligner.rs:450):
// A/a 0, C/c 1, G/g 2, T/t 3, every other byte 0.
PSA_HD uint32_t base_code(uint8_t c) {
    c &= 0xDF;  // fold case (only matters for letters)
    return c == 'C' ? 1u : c == 'G' ? 2u : c == 'T' ? 3u : 0u;
}

}  // namespace psa
