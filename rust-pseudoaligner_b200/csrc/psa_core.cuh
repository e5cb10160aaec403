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
#ifndef PSA_TWO_AHEAD
#define PSA_TWO_AHEAD 0
#endif
#if defined(__CUDA_ARCH__)
#define PSA_UNROLL _Pragma("unroll")
#else
#define PSA_UNROLL
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
// L2 residency hints (PSA_L2_HINTS): the big single-use table (`values`, 8 bytes used per 32-byte
// sector, no reuse) is loaded evict-first so that it does not push the small hot tables (node
// records, unitig sequence, class windows) out of the 126 MB L2; those are loaded evict-last.
#ifndef PSA_L2_HINTS
#define PSA_L2_HINTS 1
#endif
#if defined(__CUDA_ARCH__) && PSA_L2_HINTS
__device__ __forceinline__ uint64_t l2_policy_first() {
    uint64_t p;
    asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_last() {
    uint64_t p;
    asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t ld_u64_first(const uint64_t* a) {
    uint64_t v;
    asm("ld.global.nc.L1::no_allocate.L2::cache_hint.u64 %0, [%1], %2;" : "=l"(v) : "l"(a), "l"(l2_policy_first()));
    return v;
}
__device__ __forceinline__ uint64_t ld_u64_last(const uint64_t* a) {
    uint64_t v;
    asm("ld.global.nc.L2::cache_hint.u64 %0, [%1], %2;" : "=l"(v) : "l"(a), "l"(l2_policy_last()));
    return v;
}
__device__ __forceinline__ void ld_v4_last(const void* a, uint64_t& x, uint64_t& y, uint64_t& z, uint64_t& w) {
    asm("ld.global.nc.L2::cache_hint.v4.u64 {%0,%1,%2,%3}, [%4], %5;"
        : "=l"(x), "=l"(y), "=l"(z), "=l"(w) : "l"(a), "l"(l2_policy_last()));
}
#elif defined(__CUDA_ARCH__)
__device__ __forceinline__ uint64_t ld_u64_first(const uint64_t* a) { return __ldg(a); }
__device__ __forceinline__ uint64_t ld_u64_last(const uint64_t* a) { return __ldg(a); }
__device__ __forceinline__ void ld_v4_last(const void* a, uint64_t& x, uint64_t& y, uint64_t& z, uint64_t& w) {
    asm("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(x), "=l"(y), "=l"(z), "=l"(w) : "l"(a));
}
#endif

#ifndef PSA_PF_SUCC
#define PSA_PF_SUCC 0   // thread-per-read walk: prefetch the successor's node record before the compare
#endif
#ifndef PSA_PF_SPAN
#define PSA_PF_SPAN 0   // thread-per-read walk: prefetch the last sector of the unitig span before the compare
#endif
PSA_HD void prefetch_l2(const void* a) {
#ifdef __CUDA_ARCH__
    asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
#else
    (void)a;
#endif
}

struct GLoad {  // immutable index memory (global, ld.global.nc): the unitig sequence
    const uint64_t* p;
    PSA_HD uint64_t operator()(uint64_t i) const {
#ifdef __CUDA_ARCH__
        return ld_u64_last(p + i);
#else
        return p[i];
#endif
    }
};
struct PLoad {  // plain pointer (shared memory tile or global read buffer)
    const uint64_t* p;
    PSA_HD uint64_t operator()(uint64_t i) const { return p[i]; }
};
#ifndef PSA_READS_EVICT_LAST
#define PSA_READS_EVICT_LAST 0   // experiment: the reads' packed words keep L2 priority while their read is in flight
#endif
struct RLoad {  // the packed words of the reads (global): touched ~10 times over a read's life
    const uint64_t* p;
    PSA_HD uint64_t operator()(uint64_t i) const {
#if defined(__CUDA_ARCH__) && PSA_READS_EVICT_LAST && PSA_L2_HINTS
        return ld_u64_last(p + i);
#else
        return p[i];
#endif
    }
};
struct RegLoad6 {  // a read of at most 192 bases held in registers (selected, never indexed)
    uint64_t w0, w1, w2, w3, w4, w5;
    PSA_HD uint64_t operator()(uint64_t i) const { return i == 0 ? w0 : i == 1 ? w1 : i == 2 ? w2 : i == 3 ? w3 : i == 4 ? w4 : w5; }
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
// 64 bytes in two 32-byte sectors.  Sector 0 is all the forward walk needs (one LDG.E.256 per
// node visit: sequence, class, successors); sector 1 is read only by the left extension, the
// cooperative kernel's class list and the index self-checks.
struct NodeRec {
    uint64_t start_len;  // first base in the concatenated sequence (low 40 bits) | bases << 40
    uint32_t eq;         // equivalence-class id (*node.data())
    uint32_t class_len;  // |eq_classes[eq]|
    uint32_t succ[4];    // node reached by right extension b (kNone if the ext bit is clear)
    uint32_t pred[4];    // node reached by left extension b
    uint32_t exts;       // debruijn Exts byte
    uint32_t pad;
    uint64_t class_off;  // eq_classes[eq] starts at eq_mem[class_off]
};
static_assert(sizeof(NodeRec) == 64, "NodeRec must be one 64-byte line");
constexpr uint64_t kStartMask = (1ULL << 40) - 1;
constexpr uint32_t kMaxNodeLen = (1u << 24) - 1;
PSA_HD uint64_t pack_start_len(uint64_t start, uint32_t len) { return start | ((uint64_t)len << 40); }

struct Mphf {
    const uint64_t* blocks;  // 4 words per block: [rank:48 | c1:7 | c2:8] w1 w2 w3
    uint32_t n_levels;
    uint32_t pad;
    uint64_t level_nblk[kMaxLevels];
    uint64_t level_base[kMaxLevels];  // first block of the level
};

// Absent-key prefilter: a split-block Bloom filter over every k-mer of the graph.  One 32-byte
// block per key, one bit in each of its eight 32-bit words.  No false negatives, so consulting it
// before the MPHF never changes dict_get's answer; it lets a seed scan drop an absent k-mer after
// one sector instead of ~3 MPHF levels + the `values` sector.  (Not part of the reference: boomphf
// has nothing comparable; it pays off because unmappable reads probe 43 absent k-mers each.)
struct Bloom {
    const uint32_t* words;  // 8 per block; nullptr = no filter
    uint64_t n_blocks;
};
#ifndef PSA_BLOOM_BITS
#define PSA_BLOOM_BITS 12
#endif
constexpr uint32_t kBloomBitsPerKey = PSA_BLOOM_BITS;

struct DevIndex {
    uint32_t k;
    uint32_t node_bits, pos_bits, fp_bits;  // `values` entry = node | pos << node_bits | fp << (node_bits+pos_bits),
                                            // pos = absolute base position of the k-mer in `seq`
    uint64_t n_nodes, n_kmers, n_eq;
    const uint64_t* values;
    const NodeRec* nodes;
    const uint64_t* seq;
    const uint64_t* eq_off;
    const uint32_t* eq_mem;
    const struct ClassWin* class_win;  // one 32-byte window per class (see below)
    Bloom bloom;
    Mphf mphf;
};

// Two 64-bit hashes per k-mer, computed once; level l probes h1 + l*h2 (double hashing, as
// in BBHash), the fingerprint is the top bits of h2.
struct KeyHash {
    uint64_t h1, h2;
};
PSA_HD KeyHash make_hash(uint64_t folded) {
    KeyHash kh;
    kh.h1 = mix64(folded + 0x9E3779B97F4A7C15ULL);  // a bijection of the key
    uint64_t t = kh.h1 * 0xD6E8FEB86659FD93ULL;     // second hash derived from the first: distinct keys have
    kh.h2 = (t ^ (t >> 32)) | 1ULL;                 // distinct h1, hence (pseudo-)independent h2
    return kh;
}
PSA_HD uint64_t level_hash(KeyHash kh, uint32_t lvl) { return kh.h1 + (uint64_t)lvl * kh.h2; }
PSA_HD uint64_t fp_of(KeyHash kh, uint32_t fp_bits) { return kh.h2 >> (64 - fp_bits); }

// Bloom position of a key: block, and eight 5-bit in-word positions packed in `bits`
PSA_HD void bloom_pos(KeyHash kh, uint64_t n_blocks, uint64_t& blk, uint64_t& bits) {
    const uint64_t g = kh.h1 * 0x9E3779B97F4A7C15ULL + kh.h2;
    blk = mulhi64(g, n_blocks);
    bits = g * 0xD6E8FEB86659FD93ULL >> 24;  // 40 bits
}
PSA_HD bool bloom_maybe(const Bloom& b, KeyHash kh) {
    uint64_t blk, bits;
    bloom_pos(kh, b.n_blocks, blk, bits);
    uint32_t w[8];
#ifdef __CUDA_ARCH__
    const uint4* p = reinterpret_cast<const uint4*>(b.words + 8 * blk);
    const uint4 lo = __ldg(p), hi = __ldg(p + 1);
    w[0] = lo.x; w[1] = lo.y; w[2] = lo.z; w[3] = lo.w; w[4] = hi.x; w[5] = hi.y; w[6] = hi.z; w[7] = hi.w;
#else
    for (int i = 0; i < 8; i++) w[i] = b.words[8 * blk + i];
#endif
    uint32_t all = 1;
    PSA_UNROLL
    for (int i = 0; i < 8; i++) all &= w[i] >> ((bits >> (5 * i)) & 31);
    return all & 1;
}

// position of a key at a level: (block within level, bit 0..191 within block); nblk < 2^32
PSA_HD void level_pos(uint64_t h, uint64_t nblk, uint64_t& blk, uint32_t& bit) {
    blk = ((h >> 32) * (uint64_t)(uint32_t)nblk) >> 32;
    bit = (uint32_t)(((h & 0xffffffffULL) * kBlockBits) >> 32);
}

struct Block {  // named words, never indexed dynamically: the block must stay in registers
    uint64_t hdr, w1, w2, w3;
};
PSA_HD Block load_block(const uint64_t* blocks, uint64_t b) {
    Block r;
#ifdef __CUDA_ARCH__
    // one 32-byte sector per probe: header + 3 bit-vector words (LDG.E.256)
    asm("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];"
        : "=l"(r.hdr), "=l"(r.w1), "=l"(r.w2), "=l"(r.w3)
        : "l"(blocks + 4 * b));
#else
    r.hdr = blocks[4 * b]; r.w1 = blocks[4 * b + 1]; r.w2 = blocks[4 * b + 2]; r.w3 = blocks[4 * b + 3];
#endif
    return r;
}
PSA_HD uint64_t block_word(const Block& b, uint32_t wi) { return wi == 0 ? b.w1 : wi == 1 ? b.w2 : b.w3; }
constexpr uint64_t kRankMask = (1ULL << 48) - 1;
PSA_HD uint64_t make_header(uint64_t rank, uint32_t c1, uint32_t c2) {
    return rank | ((uint64_t)c1 << 48) | ((uint64_t)c2 << 55);
}
// rank of bit `bit` of a block whose bit is set = number of set bits before it in the cascade
PSA_HD uint64_t block_rank(const Block& b, uint32_t bit) {
    uint32_t wi = bit >> 6, bi = bit & 63;
    uint64_t hdr = b.hdr;
    uint64_t r = hdr & kRankMask;
    r += wi == 0 ? 0 : wi == 1 ? ((hdr >> 48) & 0x7f) : ((hdr >> 55) & 0xff);
    r += (uint64_t)popc64(block_word(b, wi) & ((1ULL << bi) - 1));
    return r;
}

struct ProbeStats {  // sequential-equivalent event counts of one dictionary probe
    uint32_t levels, hit, verified;
};

// Mphf::try_hash: cascade of bit-vectors; the first level whose bit is set gives the slot.
// two_ahead: fetch the blocks of levels 0 and 1 together (a likely-present key resolves at level
// 0 with p ~ 0.55 and at level 1 with p ~ 0.25: one round trip instead of two for the latter, at
// the price of one extra sector for the former).  `levels` stays the sequential count.
// Measured on B200 (config 3): k_map_thread 3.74 ms with it, 3.40 ms without -- the kernel is
// bound by the random-sector rate of HBM, not by the length of the dependent chain, so the extra
// sector costs more than the saved round trip.  Kept for reference, off (PSA_TWO_AHEAD=0).
PSA_HD bool mphf_lookup(const Mphf& m, KeyHash hk, uint64_t& slot, uint32_t& levels, bool two_ahead = false) {
    levels = 0;
    uint32_t lvl = 0;
    if (two_ahead && m.n_levels >= 2) {
        uint64_t blk0, blk1;
        uint32_t bit0, bit1;
        level_pos(level_hash(hk, 0), m.level_nblk[0], blk0, bit0);
        level_pos(level_hash(hk, 1), m.level_nblk[1], blk1, bit1);
        const Block b0 = load_block(m.blocks, m.level_base[0] + blk0);
        const Block b1 = load_block(m.blocks, m.level_base[1] + blk1);
        levels = 1;
        if ((block_word(b0, bit0 >> 6) >> (bit0 & 63)) & 1) {
            slot = block_rank(b0, bit0);
            return true;
        }
        levels = 2;
        if ((block_word(b1, bit1 >> 6) >> (bit1 & 63)) & 1) {
            slot = block_rank(b1, bit1);
            return true;
        }
        lvl = 2;
    }
    for (; lvl < m.n_levels; lvl++) {
        uint64_t blk;
        uint32_t bit;
        level_pos(level_hash(hk, lvl), m.level_nblk[lvl], blk, bit);
        Block b = load_block(m.blocks, m.level_base[lvl] + blk);
        levels++;
        if ((block_word(b, bit >> 6) >> (bit & 63)) & 1) {
            slot = block_rank(b, bit);
            return true;
        }
    }
    return false;
}

PSA_HD uint64_t pack_value(const DevIndex& ix, uint32_t node, uint64_t pos, KeyHash hk) {
    uint64_t v = (uint64_t)node | (pos << ix.node_bits);
    if (ix.fp_bits) v |= fp_of(hk, ix.fp_bits) << (ix.node_bits + ix.pos_bits);
    return v;
}

// dbg_index.get(kmer) followed by the reference's verification (src/pseudoaligner.rs:96-107).
// A fingerprint mismatch proves the slot's key differs from `key`, so skipping the unitig
// fetch cannot change the outcome of the reference's `read_kmer == ref_kmer` test.
// prefilter: ask the Bloom filter first (callers do when the key is likely absent).
template <int KW>
PSA_HD bool dict_get(const DevIndex& ix, Kmer<KW> key, uint32_t& node, uint32_t& off, ProbeStats* st,
                     bool prefilter = false, bool two_ahead = false) {
    KeyHash hk = make_hash(KmerOps<KW>::fold(key));
    if (prefilter && ix.bloom.words && !bloom_maybe(ix.bloom, hk)) return false;
    uint64_t slot;
    uint32_t levels;
    bool in = mphf_lookup(ix.mphf, hk, slot, levels, two_ahead);
    if (st) { st->levels = levels; st->hit = in; st->verified = 0; }
    if (!in) return false;
#ifdef __CUDA_ARCH__
    uint64_t v = ld_u64_first(ix.values + slot);
#else
    uint64_t v = ix.values[slot];
#endif
    if (ix.fp_bits) {
        if ((v >> (ix.node_bits + ix.pos_bits)) != fp_of(hk, ix.fp_bits)) return false;
    }
    uint32_t n = (uint32_t)(v & ((1ULL << ix.node_bits) - 1));
    uint64_t pos = (v >> ix.node_bits) & ((1ULL << ix.pos_bits) - 1);
    if (st) st->verified = 1;
    // the unitig k-mer and the node's start are independent loads: both addresses come from `v`
#ifdef __CUDA_ARCH__
    uint64_t start = ld_u64_last(&ix.nodes[n].start_len) & kStartMask;
#else
    uint64_t start = ix.nodes[n].start_len & kStartMask;
#endif
    Kmer<KW> ref = KmerOps<KW>::get(GLoad{ix.seq}, pos, ix.k);
    if (!(ref == key)) return false;
    node = n;
    off = (uint32_t)(pos - start);
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
//   view_succ(node view, b) / W::pred(id, b)     r_edges()[..].0 / l_edges()[..].0; kNone when the node has no
//                                                such extension (Exts::has_ext, :183 / :267); W::jumped() counts
//   W::push(node)                                nodes.push
//   W::abort()                                   true once the policy has given the read up (never for
//                                                the cooperative policies; see ThreadCtx)
// Returns false for None.  read_coverage is returned through `coverage`.
// ---------------------------------------------------------------------------------------------
struct NodeView {
    uint64_t start;
    uint32_t len, eq, class_len;
    uint32_t succ[4];  // never indexed dynamically (view_succ)
};
// sector 0 of a node record
PSA_HD NodeView load_node_view(const NodeRec* r) {
    NodeView v;
#ifdef __CUDA_ARCH__
    uint64_t a, b, c, d;
    ld_v4_last(r, a, b, c, d);
    v.start = a & kStartMask;
    v.len = (uint32_t)(a >> 40);
    v.eq = (uint32_t)b;
    v.class_len = (uint32_t)(b >> 32);
    v.succ[0] = (uint32_t)c; v.succ[1] = (uint32_t)(c >> 32); v.succ[2] = (uint32_t)d; v.succ[3] = (uint32_t)(d >> 32);
#else
    v.start = r->start_len & kStartMask;
    v.len = (uint32_t)(r->start_len >> 40);
    v.eq = r->eq;
    v.class_len = r->class_len;
    for (int i = 0; i < 4; i++) v.succ[i] = r->succ[i];
#endif
    return v;
}
PSA_HD uint32_t view_succ(const NodeView& v, uint32_t b) { return b == 0 ? v.succ[0] : b == 1 ? v.succ[1] : b == 2 ? v.succ[2] : v.succ[3]; }

// Optional policy hook: w.prefetch_succ(nv, pos) -- the forward walk is about to compare the rest of unitig nv
// and, if every base matches, will continue with the successor selected by read base `pos`.  Policies without
// the member get the no-op.
template <class W, class P>
PSA_HD auto hint_succ(W& w, const NodeView& nv, P pos, int) -> decltype(w.prefetch_succ(nv, pos), void()) {
    w.prefetch_succ(nv, pos);
}
template <class W, class P>
PSA_HD void hint_succ(W&, const NodeView&, P, long) {}

#ifdef __CUDACC__
#pragma nv_exec_check_disable
#endif
// P: the integer type of read positions (usize in the reference).  The kernels use uint32_t -- read lengths
// are 32-bit in the C ABI -- which halves the registers and instructions of the position arithmetic.
template <class P = uint64_t, class W>
PSA_HD bool map_read_nodes(W& w, uint32_t k, P read_length, uint32_t allowed_mismatches,
                           uint32_t& coverage) {
    P read_coverage = 0;                                                 // :71
    const P kmer_length = k;                                             // :80
    P left_extend_threshold = (P)(kLeftExtendFraction * (double)read_length);  // :77
    P kmer_pos = 0;                                                      // :79
    if (read_length < kmer_length) return false;                                // :82-84
    const P last_kmer_pos = read_length - kmer_length;                   // :86
    uint32_t n_pushed = 0;

    uint32_t node_id = 0, kmer_offset = 0;
    bool have = w.find_seed(kmer_pos, last_kmer_pos, node_id, kmer_offset);     // :118-121

    // left extension, :124-205.  (kmer_pos >= 1 only fails for read_length < 5, where the
    // reference's `kmer_pos - 1` would underflow; unreachable for k >= 5.)
    if (have && kmer_pos >= left_extend_threshold && kmer_pos >= 1) {
        P last_pos = kmer_pos - 1;                                       // :127
        uint32_t prev_node_id = node_id;                                        // :128
        P prev_kmer_offset = kmer_offset > 0 ? kmer_offset - 1 : 0;      // :129 (sic)
        for (;;) {                                                              // :131
            NodeView nv = w.node(prev_node_id);                                 // :132
            P skipped_read = last_pos + 1;                               // :139
            P skipped_ref = prev_kmer_offset + 1;                        // :142
            P max_matchable_pos = skipped_read < skipped_ref ? skipped_read : skipped_ref;  // :145
            bool premature_break = false;                                       // :148
            P matched_bases =                                            // :149-170
                w.cmp_bwd(last_pos, nv.start + prev_kmer_offset, max_matchable_pos, allowed_mismatches,
                          premature_break);
            read_coverage += matched_bases;                                     // :169
            if (last_pos + 1 - matched_bases == 0 || premature_break) break;    // :173-175
            last_pos -= matched_bases;                                          // :178
            uint32_t next_base = w.read_base(last_pos);                         // :182
            const uint32_t pred_id = w.pred(prev_node_id, next_base);          // kNone <=> no such left ext
            if (pred_id != kNone) {                                             // :183
                prev_node_id = pred_id;                                         // :185-194
                NodeView pv = w.node(prev_node_id);                             // :195
                prev_kmer_offset = pv.len - kmer_length;                        // :196
                w.push(prev_node_id, pv);                                       // :199
                n_pushed++;
                if (w.abort()) return false;  // policy gave the read up (it is redone by another kernel)
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
            if (w.abort()) return false;
            P remaining_read = read_length - kmer_pos;                   // :222
            P ref_length = nv.len;                                       // :226
            P ref_offset = kmer_offset + kmer_length;                    // :227
            P informative_ref = ref_length - ref_offset;                 // :228
            P max_matchable_pos = remaining_read < informative_ref ? remaining_read : informative_ref;  // :231
            if (remaining_read > informative_ref) hint_succ(w, nv, (P)(kmer_pos + informative_ref), 0);
            bool premature_break = false;                                       // :233
            P matched_bases =                                            // :234-255
                w.cmp_fwd(kmer_pos, nv.start + ref_offset, max_matchable_pos, allowed_mismatches,
                          premature_break);
            read_coverage += matched_bases;                                     // :254
            kmer_pos += matched_bases;                                          // :257
            if (kmer_pos >= read_length) break;                                 // :259-261
            uint32_t next_base = w.read_base(kmer_pos);                         // :265
            const uint32_t succ_id = view_succ(nv, next_base);                  // kNone <=> no such right ext
            if (!premature_break && succ_id != kNone) {                         // :267
                w.jumped();
                node_id = succ_id;                                              // :269-278
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

// ---------------------------------------------------------------------------------------------
// Class windows: nodes_to_eq_class (ref src/pseudoaligner.rs:323-356) as bit-parallel ANDs.
//
// An equivalence class is an ascending list of transcript ids (ref src/equiv_classes.rs:78-79).
// Transcripts of one gene are neighbours in the FASTA, so most classes span a narrow id range.
// Every class whose max - min < 192 also gets a 32-byte window {min, 192-bit membership map};
// the intersection of such classes is the AND of their windows aligned to a common base -- one
// sector per class and a few shifts instead of a binary search per member per class.  Classes
// that do not fit ("wide") keep the reference's list search (:399-404) against the candidates
// that survive the windows, so the result is exact in every case.
// ---------------------------------------------------------------------------------------------
constexpr uint32_t kWinBits = 192;
constexpr uint32_t kWinWide = 0xFFFFFFFFu;
struct ClassWin {      // 32 bytes, 32-byte aligned
    uint32_t lo;       // smallest member
    uint32_t len;      // |class|, or kWinWide (bits unused)
    uint64_t bits[3];  // bit t of the 192-bit map set iff lo + t is a member
};
static_assert(sizeof(ClassWin) == 32, "ClassWin must be one 32-byte sector");

struct Win {
    uint64_t w0, w1, w2;
};
PSA_HD Win win_and(Win a, Win b) { return Win{a.w0 & b.w0, a.w1 & b.w1, a.w2 & b.w2}; }
PSA_HD uint32_t win_popc(Win a) { return (uint32_t)(popc64(a.w0) + popc64(a.w1) + popc64(a.w2)); }
PSA_HD bool win_empty(Win a) { return (a.w0 | a.w1 | a.w2) == 0; }
// logical shift right of the 192-bit map by d (any d; >= 192 gives 0)
PSA_HD Win win_shr(Win a, uint32_t d) {
    if (d >= kWinBits) return Win{0, 0, 0};
    const uint32_t ws = d >> 6, bs = d & 63;
    uint64_t x0 = ws == 0 ? a.w0 : ws == 1 ? a.w1 : a.w2;
    uint64_t x1 = ws == 0 ? a.w1 : ws == 1 ? a.w2 : 0;
    uint64_t x2 = ws == 0 ? a.w2 : 0;
    if (bs) {
        x0 = (x0 >> bs) | (x1 << (64 - bs));
        x1 = (x1 >> bs) | (x2 << (64 - bs));
        x2 >>= bs;
    }
    return Win{x0, x1, x2};
}
// the window of one class (index construction; members ascending, unique)
PSA_HD ClassWin make_class_win(const uint32_t* members, uint64_t len) {
    ClassWin c;
    c.lo = len ? members[0] : 0;
    c.len = (uint32_t)len;
    c.bits[0] = c.bits[1] = c.bits[2] = 0;
    if (len && members[len - 1] - members[0] >= kWinBits) {
        c.len = kWinWide;
        return c;
    }
    for (uint64_t i = 0; i < len; i++) {
        uint32_t t = members[i] - c.lo;
        c.bits[t >> 6] |= 1ULL << (t & 63);
    }
    return c;
}
PSA_HD ClassWin load_class_win(const ClassWin* p) {
    ClassWin c;
#ifdef __CUDA_ARCH__
    uint64_t a, b0, b1, b2;
    ld_v4_last(p, a, b0, b1, b2);
    c.lo = (uint32_t)a;
    c.len = (uint32_t)(a >> 32);
    c.bits[0] = b0; c.bits[1] = b1; c.bits[2] = b2;
#else
    c = *p;
#endif
    return c;
}
// Running intersection of narrow classes: {base, map} means the set {base + t : bit t set}.
struct WinAcc {
    uint32_t base;
    Win map;
    bool have;
};
PSA_HD void winacc_merge(WinAcc& a, const WinAcc& b) {
    if (!b.have) return;
    if (!a.have) {
        a = b;
    } else if (b.base >= a.base) {
        a.map = win_and(win_shr(a.map, b.base - a.base), b.map);
        a.base = b.base;
    } else {
        a.map = win_and(a.map, win_shr(b.map, a.base - b.base));
    }
}
PSA_HD void winacc_and(WinAcc& a, const ClassWin& c) {  // c narrow
    WinAcc b;
    b.base = c.lo;
    b.map = Win{c.bits[0], c.bits[1], c.bits[2]};
    b.have = true;
    winacc_merge(a, b);
}
// members of the accumulated set, ascending
PSA_HD uint32_t win_write(const WinAcc& a, uint32_t* out) {
    uint32_t n = 0;
    uint64_t w = a.map.w0;
    while (w) { out[n++] = a.base + (uint32_t)ctz64(w); w &= w - 1; }
    w = a.map.w1;
    while (w) { out[n++] = a.base + 64 + (uint32_t)ctz64(w); w &= w - 1; }
    w = a.map.w2;
    while (w) { out[n++] = a.base + 128 + (uint32_t)ctz64(w); w &= w - 1; }
    return n;
}
PSA_HD uint32_t ld_mem(const uint32_t* p) {
#ifdef __CUDA_ARCH__
    return __ldg(p);
#else
    return *p;
#endif
}
PSA_HD uint64_t ld_off(const uint64_t* p) {
#ifdef __CUDA_ARCH__
    return __ldg(p);
#else
    return *p;
#endif
}
// The members of the ascending list v[0..n) that fall in [base, base + 192), as a window map over
// `base`: one binary search for the first such member (the reference's own search, ref :404), then
// the few that follow.  This is how a wide class is applied to the candidates the narrow classes'
// windows have left: exact, and one search per class instead of one per candidate.
PSA_HD Win win_of_list_range(const uint32_t* v, uint32_t n, uint32_t base) {
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
        uint32_t mid = lo + ((hi - lo) >> 1);
        if (ld_mem(v + mid) < base) lo = mid + 1;
        else hi = mid;
    }
    Win m{0, 0, 0};
    for (; lo < n; lo++) {
        const uint32_t x = ld_mem(v + lo);
        if (x - base >= kWinBits) break;  // (x >= base here)
        const uint32_t t = x - base;
        const uint64_t bit = 1ULL << (t & 63);
        if (t < 64) m.w0 |= bit;
        else if (t < 128) m.w1 |= bit;
        else m.w2 |= bit;
    }
    return m;
}
// drop from the accumulated set every member the (wide) class list v[0..n) lacks
PSA_HD void winacc_filter_list(WinAcc& a, const uint32_t* v, uint32_t n) {
    a.map = win_and(a.map, win_of_list_range(v, n, a.base));
}

// ---------------------------------------------------------------------------------------------
// One thread = one read: the policy of the fast kernel (k_map_thread).  It runs the same
// map_read_nodes text with every step done serially by the calling thread, and gives a read
// up ("defer") as soon as it needs something a single thread does badly: a seed scan longer
// than max_probes positions, more than kThreadWide distinct WIDE classes, or only wide classes
// with a smallest one of more than max_small members.  Deferred reads are redone from scratch
// by the cooperative kernel (k_map over the deferred list), so the split never changes a result.
//
// Classes are intersected ONLINE: the window of every newly visited class is ANDed into one
// accumulator as the walk goes (any number of narrow classes in constant registers, and the
// window load overlaps the next compare); only wide classes are listed, for the final filter.
// eq_id needs no list either: the result equals a visited class iff its size equals the smallest
// visited class length, and then it is the smallest-id class of that length.
// ---------------------------------------------------------------------------------------------
constexpr int kThreadWide = 3;
#ifndef PSA_WIDE_INLINE
#define PSA_WIDE_INLINE 2
#endif
constexpr uint32_t kThreadWideInline = PSA_WIDE_INLINE;  // further wide classes applied on arrival before giving up
constexpr int kThreadRecent = 2;
constexpr uint32_t kReseedProbes = 8;
constexpr uint32_t kFlagAligned = 1u, kFlagMapped = 2u;

struct HitRec {  // == psa_hit
    uint32_t coverage, n_tx;
    uint64_t tx_off;
    uint32_t eq_id, flags;
};

struct ThreadEvents {
    uint32_t lookups, levels, hits, verifs, visits, bases, jumps, members;
};

#ifndef PSA_CMP_CARRY
#define PSA_CMP_CARRY 0
#endif
// Sequential 32-base chunks of a 2-bit sequence, each word loaded once (see ThreadCtx::cmp).  next(n, more)
// returns the n (1..32) bases at the current position right-aligned, exactly as seq_bits does, and moves on
// by 32 bases; `more` says whether another chunk follows (only then may the following word be touched).
template <class L>
struct CarryStream {
    uint64_t wi, cur;
    uint32_t in_word;
    PSA_HD void start(L ld, uint64_t pos) {
        wi = pos >> 5;
        in_word = (uint32_t)(pos & 31);
        cur = ld(wi);
    }
    PSA_HD uint64_t next(L ld, uint32_t n, bool more) {
        uint64_t v = cur << (2 * in_word);
        if (n > 32 - in_word) {          // the chunk runs into the next word: that word starts the next chunk
            const uint64_t nx = ld(wi + 1);
            v |= nx >> (64 - 2 * in_word);
            cur = nx;
        } else if (more) {               // (in_word == 0 and a full chunk: the next chunk is the next word)
            cur = ld(wi + 1);
        }
        wi++;
        return v >> (64 - 2 * n);
    }
};

template <int KW, bool EV, class RD = PLoad>
struct ThreadCtx {
    const DevIndex& ix;
    RD rd;
    uint32_t max_probes;
    // online class state
    bool multi;                    // more than one distinct class visited (until then min_eq/min_len ARE the one
                                   // class seen, whose window is fetched only when a second class shows up)
    uint32_t min_len, min_eq;      // smallest class length seen, and the smallest id among the classes of that length
    WinAcc acc;                    // AND of the narrow classes' windows
    uint32_t wide_eq[kThreadWide], n_wide, n_inline;  // (their lengths are re-read from eq_off when needed)
    uint32_t recent[kThreadRecent];  // last few class ids (skips most repeated window loads; repeats are harmless)
    bool defer;
    uint32_t why;  // diagnostic: 0 first seed search, 1 re-seed search, 2 class list full, 3 smallest class too long
    bool seeded;
    // answer of the read's first seed search when k_seed_scan has already made it
    bool has_hint;
    uint32_t hint_pos, hint_node, hint_off;
    ThreadEvents ev;

    PSA_HD ThreadCtx(const DevIndex& ix_, RD rd_, uint32_t max_probes_)
        : ix(ix_), rd(rd_), max_probes(max_probes_), multi(false),
          min_len(kNone), min_eq(kNone), n_wide(0), n_inline(0), defer(false), why(0), seeded(false),
          has_hint(false), hint_pos(0), hint_node(0), hint_off(0), ev{} {
        acc.base = 0; acc.map = Win{0, 0, 0}; acc.have = false;
    PSA_UNROLL
        for (int j = 0; j < kThreadWide; j++) wide_eq[j] = kNone;
    PSA_UNROLL
        for (int j = 0; j < kThreadRecent; j++) recent[j] = kNone;
    }
    PSA_HD bool abort() const { return defer; }
    PSA_HD uint32_t read_base(uint64_t pos) const { return seq_get(rd, pos); }

    // find_kmer_match, ref src/pseudoaligner.rs:91-114, at most max_probes positions
    template <class P>
    PSA_HD bool find_seed(P& kmer_pos, P last, uint32_t& node, uint32_t& o) {
        if (kmer_pos > last) return false;
        if (has_hint) {  // the first search of the read (it starts at 0), done by k_seed_scan
            has_hint = false;
            kmer_pos = hint_pos;
            node = hint_node;
            o = hint_off;
            seeded = true;
            return true;
        }
        const P start = kmer_pos;
        P p = start;
        for (uint32_t probes = 0;; probes++, p += kSeedStride) {
            if (p > last) {
                kmer_pos = start + kSeedStride * ((last - start) / kSeedStride + 1);  // where the loop at :92-111 stops
                return false;
            }
            // re-seed searches (ref :293) are short as a rule and have no scan kernel of their own: allow them more
            if (probes >= (seeded ? (max_probes > kReseedProbes ? max_probes : kReseedProbes) : max_probes)) {
                defer = true;
                why = seeded ? 1 : 0;
                kmer_pos = last + 1;  // keeps map_read_nodes out of the forward loop
                return false;
            }
            ProbeStats st;
            // after a miss the next positions are likely absent too: Bloom first (never when counting
            // events, which are defined on the MPHF path)
            bool hit = dict_get<KW>(ix, KmerOps<KW>::get(rd, p, ix.k), node, o, EV ? &st : nullptr, !EV && probes > 0,
                                    PSA_TWO_AHEAD && probes == 0);
            if (EV) { ev.lookups++; ev.levels += st.levels; ev.hits += st.hit; ev.verifs += st.verified; }
            if (hit) {
                kmer_pos = p;
                seeded = true;
                return true;
            }
        }
    }
    PSA_HD NodeView node(uint32_t id) const { return load_node_view(ix.nodes + id); }
#if PSA_PF_SUCC
    // the node record the walk needs next if the rest of this unitig matches: fetched under the compare
    template <class P>
    PSA_HD void prefetch_succ(const NodeView& nv, P pos) {
        const uint32_t s = view_succ(nv, read_base(pos));
        if (s != kNone) prefetch_l2(ix.nodes + s);
    }
#endif
    PSA_HD void jumped() {
        if (EV) ev.jumps++;
    }
    PSA_HD uint32_t pred(uint32_t id, uint32_t b) {
#ifdef __CUDA_ARCH__
        const uint32_t p = __ldg(&ix.nodes[id].pred[b]);
#else
        const uint32_t p = ix.nodes[id].pred[b];
#endif
        if (EV && p != kNone) ev.jumps++;
        return p;
    }
    // ref src/pseudoaligner.rs:234-255 (FWD) and :149-170 (backward), 32 bases per step
    template <bool FWD, class P>
    PSA_HD P cmp(P rp, uint64_t sp, P m, uint32_t A, bool& premature) {
        uint32_t snp = 0;
#if PSA_PF_SPAN
        if (m > 32) {  // the far end of the unitig span, if it lies in another 32-byte sector (128 bases)
            const uint64_t far = FWD ? sp + m - 1 : sp - (m - 1);
            if ((far >> 7) != (sp >> 7)) prefetch_l2(ix.seq + (far >> 5));
        }
#endif
#if PSA_CMP_CARRY
        // forward compare with every word of the read and of the unitig loaded ONCE: consecutive 32-base chunks
        // start one word apart at a constant in-word offset, so a chunk's second word is the next chunk's first.
        // (Experiment switch, off: hostsim-verified, not yet measured on the GPU.)
        CarryStream<RD> rs;
        CarryStream<GLoad> ss;
        if (FWD && m > 0) {
            rs.start(rd, rp);
            ss.start(GLoad{ix.seq}, sp);
        }
#endif
        for (P my = 0; my < m; my += 32) {
            uint32_t n = m - my < 32 ? (uint32_t)(m - my) : 32u;
#if PSA_CMP_CARRY
            uint64_t mask;
            if (FWD) {
                const bool more = my + 32 < m;
                uint64_t x = rs.next(rd, n, more) ^ ss.next(GLoad{ix.seq}, n, more);  // base t at bits 2(n-1-t)
                x = rev_pairs(x) >> (64 - 2 * n);
                mask = fold_pairs(x);
            } else {
                mask = mismatch_bwd(rd, rp - my, GLoad{ix.seq}, sp - my, n);
            }
#else
            uint64_t mask = FWD ? mismatch_fwd(rd, rp + my, GLoad{ix.seq}, sp + my, n)
                                : mismatch_bwd(rd, rp - my, GLoad{ix.seq}, sp - my, n);
#endif
            uint32_t c = (uint32_t)popc64(mask);
            if (snp + c > A) {
                premature = true;
                P matched = my + nth_mismatch(mask, A + 1 - snp);
                if (EV) ev.bases += (uint32_t)matched + 1;
                return matched;
            }
            snp += c;
        }
        if (EV) ev.bases += (uint32_t)m;
        return m;
    }
    template <class P>
    PSA_HD P cmp_fwd(P rp, uint64_t sp, P m, uint32_t A, bool& pb) { return cmp<true>(rp, sp, m, A, pb); }
    template <class P>
    PSA_HD P cmp_bwd(P rp, uint64_t sp, P m, uint32_t A, bool& pb) { return cmp<false>(rp, sp, m, A, pb); }
    // AND one class into the running intersection (narrow), or list it (wide)
    PSA_HD void and_class(uint32_t e, uint32_t l) {
        const ClassWin c = load_class_win(ix.class_win + e);
        if (c.len != kWinWide) {
            winacc_and(acc, c);
            return;
        }
        bool dup = false;
    PSA_UNROLL
        for (int j = 0; j < kThreadWide; j++) dup |= (j < (int)n_wide && wide_eq[j] == e);
        if (dup) return;
        if (n_wide >= (uint32_t)kThreadWide) {
            // no room to remember it: apply it now to the candidates the windows have left (the filter
            // is idempotent and commutes with the ANDs still to come); without any window yet, give up
            if (acc.have && n_inline < kThreadWideInline) {
                n_inline++;
                winacc_filter_list(acc, ix.eq_mem + ld_off(ix.eq_off + e), l);
                return;
            }
            defer = true;
            why = 2;
            return;
        }
    PSA_UNROLL
        for (int j = 0; j < kThreadWide; j++)
            if (j == (int)n_wide) wide_eq[j] = e;
        n_wide++;
    }
    // nodes.push: only the classes matter, and the intersection is idempotent (ref :352-355)
    PSA_HD void push(uint32_t /*node_id*/, const NodeView& nv) {
        if (EV) ev.visits++;
        bool dup = false;
    PSA_UNROLL
        for (int j = 0; j < kThreadRecent; j++) dup |= (recent[j] == nv.eq);
        if (dup) return;
    PSA_UNROLL
        for (int j = kThreadRecent - 1; j > 0; j--) recent[j] = recent[j - 1];
        recent[0] = nv.eq;
        if (EV) ev.members += nv.class_len;  // (a class revisited after kThreadRecent others is counted again)
        const bool first = min_eq == kNone;
        if (!first && !multi) {
            if (nv.eq == min_eq) return;   // still the one class seen so far
            multi = true;
            and_class(min_eq, min_len);
        }
        if (nv.class_len < min_len || (nv.class_len == min_len && nv.eq < min_eq)) {
            min_len = nv.class_len;
            min_eq = nv.eq;
        }
        if (multi) and_class(nv.eq, nv.class_len);
    }
};

// All visited classes are wide: the reference's list scheme by one thread.  Members of the
// smallest listed class (index s) that every other listed class contains, ascending; each other
// class is searched only in the suffix after its previous match (ref :399-404).
// out == nullptr counts.
template <int KW, bool EV, class RD>
PSA_HD uint32_t thread_intersect_lists(const ThreadCtx<KW, EV, RD>& w, int s, uint32_t* out) {
    const uint32_t* mem = w.ix.eq_mem;
    uint32_t cur[kThreadWide];
    PSA_UNROLL
    for (int j = 0; j < kThreadWide; j++) cur[j] = 0;
    uint32_t s_eq = 0;
    PSA_UNROLL
    for (int j = 0; j < kThreadWide; j++)
        if (j == s) s_eq = w.wide_eq[j];
    const uint64_t s_off = ld_off(w.ix.eq_off + s_eq);
    const uint32_t s_len = (uint32_t)(ld_off(w.ix.eq_off + s_eq + 1) - s_off);
    uint32_t count = 0;
    for (uint32_t i = 0; i < s_len; i++) {
        const uint32_t x = ld_mem(mem + s_off + i);
        bool alive = true;
    PSA_UNROLL
        for (int j = 0; j < kThreadWide; j++) {
            if (j == s || j >= (int)w.n_wide || !alive) continue;
            const uint64_t o = ld_off(w.ix.eq_off + w.wide_eq[j]);
            const uint32_t n = (uint32_t)(ld_off(w.ix.eq_off + w.wide_eq[j] + 1) - o);
            const uint32_t* v = mem + o;
            uint32_t lo = cur[j], hi = n;
            while (lo < hi) {
                uint32_t mid = lo + ((hi - lo) >> 1);
                if (ld_mem(v + mid) < x) lo = mid + 1;
                else hi = mid;
            }
            cur[j] = lo;
            alive = lo < n && ld_mem(v + lo) == x;
        }
        if (alive) {
            if (out) out[count] = x;
            count++;
        }
    }
    return count;
}

// Result of map_read for one read as the kernels store it (flag = ref :453-462).
struct ThreadResult {
    HitRec hit;
    uint64_t count_slot;  // index into counts[]: eq id, n_eq (no visited class), n_eq + 1 (None)
    bool deferred;
    bool novel_overflow;
    uint32_t why;  // ThreadCtx::why when deferred
};

// map_read + the process_reads flag for one read, by one thread.  NovelAlloc::operator()(count,
// off_out) returns room for `count` members of a set that is no visited class (nullptr: no
// room / members not wanted).
template <int KW, bool EV, class NovelAlloc, class RD>
PSA_HD ThreadResult map_read_thread(const DevIndex& ix, RD words, uint32_t L, uint32_t allowed,
                                    uint32_t max_probes, uint32_t max_small, NovelAlloc& novel, bool want_members,
                                    ThreadEvents* ev_out, const uint32_t* hint = nullptr /* pos, node, off */) {
    ThreadResult res;
    res.hit.coverage = 0; res.hit.n_tx = 0; res.hit.tx_off = 0; res.hit.eq_id = kNone; res.hit.flags = 0;
    res.count_slot = ix.n_eq + 1;
    res.deferred = false;
    res.novel_overflow = false;
    ThreadCtx<KW, EV, RD> w(ix, words, max_probes);
    if (hint) {
        w.has_hint = true;
        w.hint_pos = hint[0]; w.hint_node = hint[1]; w.hint_off = hint[2];
    }
    uint32_t coverage = 0;
    bool some = map_read_nodes<uint32_t>(w, ix.k, L, allowed, coverage);
    res.why = w.why;
    if (w.defer) {
        res.deferred = true;
        return res;
    }
    if (some) {
        uint32_t count, eq_id;
        int s = 0;  // smallest wide class (only used when every class is wide)
        if (!w.multi) {
            count = w.min_len;
            eq_id = w.min_eq;
        } else {
            if (w.acc.have) {
                // wide classes filter what survived the windows (ref :399-404 on the candidates)
    PSA_UNROLL
                for (int j = 0; j < kThreadWide; j++) {
                    if (j >= (int)w.n_wide || win_empty(w.acc.map)) continue;
                    const uint64_t o = ld_off(ix.eq_off + w.wide_eq[j]);
                    winacc_filter_list(w.acc, ix.eq_mem + o, (uint32_t)(ld_off(ix.eq_off + w.wide_eq[j] + 1) - o));
                }
                count = win_popc(w.acc.map);
            } else {
                // smallest class first (ref :331-334)
                uint32_t s_eq = w.wide_eq[0];
                uint32_t s_len = (uint32_t)(ld_off(ix.eq_off + s_eq + 1) - ld_off(ix.eq_off + s_eq));
    PSA_UNROLL
                for (int j = 1; j < kThreadWide; j++) {
                    if (j >= (int)w.n_wide) continue;
                    const uint32_t e = w.wide_eq[j];
                    const uint32_t l = (uint32_t)(ld_off(ix.eq_off + e + 1) - ld_off(ix.eq_off + e));
                    if (l < s_len || (l == s_len && e < s_eq)) { s = j; s_len = l; s_eq = e; }
                }
                if (s_len > max_small) {  // long lists are the cooperative kernel's job
                    res.why = 3;
                    res.deferred = true;
                    return res;
                }
                count = thread_intersect_lists(w, s, (uint32_t*)nullptr);
            }
            // the result equals a visited class iff it has as many members as the smallest visited class
            eq_id = count == w.min_len ? w.min_eq : kNone;
        }
        res.hit.coverage = coverage;
        res.hit.n_tx = count;
        res.hit.eq_id = eq_id;
        res.hit.flags = kFlagAligned | ((coverage >= kCoverageThreshold && count == 0) ? kFlagMapped : 0u);  // ref :455 (sic)
        if (eq_id != kNone) {
            res.hit.tx_off = ld_off(ix.eq_off + eq_id);  // members are read from the index by k_expand
            res.count_slot = eq_id;
        } else {
            res.count_slot = ix.n_eq;
            if (count && want_members) {
                uint64_t o = 0;
                uint32_t* dst = novel(count, o);
                if (!dst) res.novel_overflow = true;
                else if (w.acc.have) win_write(w.acc, dst);
                else thread_intersect_lists(w, s, dst);
                res.hit.tx_off = o;
            }
        }
    }
    if (EV && ev_out) *ev_out = w.ev;
    return res;
}

// ASCII -> 2-bit code, DnaString::from_dna_string (call site ref src/pseudoaligner.rs:450):
// A/a 0, C/c 1, G/g 2, T/t 3, every other byte 0.
PSA_HD uint32_t base_code(uint8_t c) {
    c &= 0xDF;  // fold case (only matters for letters)
    return c == 'C' ? 1u : c == 'G' ? 2u : c == 'T' ? 3u : 0u;
}

}  // namespace psa
