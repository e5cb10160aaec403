// psa_core.cuh -- arithmetic of the pseudoalignment hot path, shared by every kernel.
//
// Everything here is scalar __host__ __device__ code: 2-bit sequence access, the k-mer hash,
// the bucket-cascade k-mer dictionary, the k-mer verification, the per-word mismatch masks of
// the two extension loops, the class-window intersection and the map_read state machine in its
// blocking form, templated on a policy that supplies the loads: WarpCtx (psa_kernels.cuh, a group of lanes per
// read: the cooperative kernels) and ThreadCtx (psa_thread.cuh, one thread per read).
// The CUDA kernels in psa_kernels.cuh instantiate this text; tests/hostsim instantiates the same
// text with serial policies so that it is checked against the oracle on machines without a GPU.
// That host instantiation is a unit-test harness only -- the product library (psa_api.cu) has
// no CPU path.
//
// Reference being restated: 10XGenomics/rust-pseudoaligner @ 9d9cab8
//   src/pseudoaligner.rs:64-319   map_read_to_nodes_with_mismatch   -> map_read_nodes()
//   src/pseudoaligner.rs:91-114   find_kmer_match                   -> W::find_seed + dict_get()
//   src/pseudoaligner.rs:99-107   dictionary answer verification    -> dict_get()
//   src/pseudoaligner.rs:323-356  nodes_to_eq_class                 -> ClassAcc
//   src/config.rs:16-18           constants
// and, from the un-vendored crates (published algorithms, see oracle/psa_oracle.h):
//   debruijn DnaString packing / get_kmer; boomphf's NoKeyBoomHashMap::get is replaced by an
//   exact k-mer -> (node, position) dictionary of this project's own design (any exact dictionary
//   gives the reference's output because every answer is verified against the unitig).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define PSA_HD __host__ __device__ __forceinline__
#else
#define PSA_HD inline
#endif
#if defined(__CUDA_ARCH__)
#define PSA_UNROLL _Pragma("unroll")
#else
#define PSA_UNROLL
#endif

namespace psa {

constexpr uint32_t kNone = 0xFFFFFFFFu;
constexpr int kMaxLevels = 32;        // dictionary cascade depth cap (1e8 keys need ~9 at 1.7 slots per key)
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
// loaders: the same text reads the index through the read-only path on the device
// ---------------------------------------------------------------------------------------------
// L2 residency hints: the dictionary (one use per sector, no reuse, 10x the size of L2) is loaded
// evict-first and kept out of L1 so that it does not push the small hot tables (node records,
// unitig sequence) out of the 126 MB L2; those are loaded evict-last.
#if defined(__CUDACC__)
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
__device__ __forceinline__ uint64_t ld_u64_last(const uint64_t* a) {
    uint64_t v;
    asm("ld.global.nc.L2::cache_hint.u64 %0, [%1], %2;" : "=l"(v) : "l"(a), "l"(l2_policy_last()));
    return v;
}
__device__ __forceinline__ void ld_v4_last(const void* a, uint64_t& x, uint64_t& y, uint64_t& z, uint64_t& w) {
    asm("ld.global.nc.L2::cache_hint.v4.u64 {%0,%1,%2,%3}, [%4], %5;"
        : "=l"(x), "=l"(y), "=l"(z), "=l"(w) : "l"(a), "l"(l2_policy_last()));
}
__device__ __forceinline__ void ld_v4_first(const void* a, uint64_t& x, uint64_t& y, uint64_t& z, uint64_t& w) {
    asm("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u64 {%0,%1,%2,%3}, [%4], %5;"
        : "=l"(x), "=l"(y), "=l"(z), "=l"(w) : "l"(a), "l"(l2_policy_first()));
}
#endif

// one 32-byte sector (LDG.E.256 on the device) as four named words: never indexed dynamically,
// so that it stays in registers
struct Sector {
    uint64_t w0, w1, w2, w3;
};
PSA_HD uint64_t sector_word(const Sector& s, uint32_t i) { return i == 0 ? s.w0 : i == 1 ? s.w1 : i == 2 ? s.w2 : s.w3; }
PSA_HD Sector load_sector_hot(const void* p) {  // node records, class windows: reused across reads
    Sector s;
#ifdef __CUDA_ARCH__
    ld_v4_last(p, s.w0, s.w1, s.w2, s.w3);
#else
    const uint64_t* q = static_cast<const uint64_t*>(p);
    s.w0 = q[0]; s.w1 = q[1]; s.w2 = q[2]; s.w3 = q[3];
#endif
    return s;
}
#ifndef PSA_DICT_LOAD
#define PSA_DICT_LOAD 0   // 0: evict-first + no L1 allocation, 1: plain read-only load, 2: evict-first only
#endif
PSA_HD Sector load_sector_stream(const void* p) {  // dictionary buckets: one use
    Sector s;
#ifdef __CUDA_ARCH__
#if PSA_DICT_LOAD == 1
    asm("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(s.w0), "=l"(s.w1), "=l"(s.w2), "=l"(s.w3) : "l"(p));
#elif PSA_DICT_LOAD == 2
    asm("ld.global.nc.L2::cache_hint.v4.u64 {%0,%1,%2,%3}, [%4], %5;"
        : "=l"(s.w0), "=l"(s.w1), "=l"(s.w2), "=l"(s.w3) : "l"(p), "l"(l2_policy_first()));
#else
    ld_v4_first(p, s.w0, s.w1, s.w2, s.w3);
#endif
#else
    const uint64_t* q = static_cast<const uint64_t*>(p);
    s.w0 = q[0]; s.w1 = q[1]; s.w2 = q[2]; s.w3 = q[3];
#endif
    return s;
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
struct PLoad {  // plain pointer (global read buffer, host memory)
    const uint64_t* p;
    PSA_HD uint64_t operator()(uint64_t i) const { return p[i]; }
};
struct WLoad {  // four consecutive words held in registers, the first one being word `base`
    Sector s;
    uint64_t base;
    PSA_HD uint64_t operator()(uint64_t i) const { return sector_word(s, (uint32_t)(i - base)); }
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
// 64 bytes in two 32-byte sectors, both fetched at every node visit (adjacent: one DRAM row).
// Sector 0 is the node itself -- sequence span, class, all four successors (the reference does up
// to 4 binary searches over all nodes per jump, ref src/pseudoaligner.rs:275).  Sector 1 is the
// window of the node's class (see "Class windows"): the intersection needs nothing else for a
// narrow class, so nodes_to_eq_class costs no gather of its own.
struct NodeRec {
    uint64_t start_len;  // first base in the concatenated sequence (low 40 bits) | bases << 40
    uint32_t eq;         // equivalence-class id (*node.data())
    uint32_t class_len;  // |eq_classes[eq]|
    uint32_t succ[4];    // node reached by right extension b (kNone if the ext bit is clear)
    uint32_t win_lo;     // ClassWin of eq
    uint32_t win_len;
    uint64_t win_bits[3];
};
static_assert(sizeof(NodeRec) == 64, "NodeRec must be one 64-byte line");
// Read only by the left extension (ref :183-196) and the index self-checks.
struct NodeCold {
    uint32_t pred[4];    // node reached by left extension b
    uint32_t exts;       // debruijn Exts byte
    uint32_t pad;
    uint64_t class_off;  // eq_classes[eq] starts at eq_mem[class_off]
};
static_assert(sizeof(NodeCold) == 32, "NodeCold must be one 32-byte sector");
constexpr uint64_t kStartMask = (1ULL << 40) - 1;
constexpr uint32_t kMaxNodeLen = (1u << 24) - 1;
PSA_HD uint64_t pack_start_len(uint64_t start, uint32_t len) { return start | ((uint64_t)len << 40); }

// k-mer dictionary: dbg_index of the reference (NoKeyBoomHashMap<K, (u32, u32)>, ref
// src/pseudoaligner.rs:31, built at src/build_index.rs:182-221) as a cascade of bucket tables.
// A bucket is one 32-byte sector of four 64-bit entries
//     entry = node | pos << node_bits | fingerprint << (node_bits + pos_bits)   (bit 63 clear)
// (pos = absolute base position of the k-mer in `seq`); unused entries are all ones.  A key lives in
// the first level whose bucket had room for it and no entry with the same fingerprint; bit 63 of
// entry 0 says that some key of this bucket was passed on to the next level.  So a probe is ONE
// sector for ~94 % of the present keys and ~91 % of the absent ones (1.7 slots per key), a
// fingerprint matches at most one entry of a bucket, and a matching entry is verified against the
// unitig exactly as the reference verifies its MPHF's answer (ref :99-107) -- dictionary membership
// is exact.  (boomphf needs ~1.8 bit-vector words + a rank + the `values` entry per present key and
// ~3 levels + a false-positive `values` entry per absent one; round 1 of this project did the same
// with 32-byte blocks and paid 2.8 / 4 sectors.)
constexpr uint64_t kEmptyEntry = ~0ULL;
constexpr uint64_t kMoreBit = 1ULL << 63;
constexpr uint32_t kBucketSlots = 4;
struct Dict {
    const uint64_t* buckets;  // 4 words per bucket
    uint32_t n_levels;
    uint32_t pad;
    uint64_t level_nbkt[kMaxLevels];  // buckets of the level (< 2^32)
    uint64_t level_base[kMaxLevels];  // first bucket of the level
};

struct DevIndex {
    uint32_t k;
    uint32_t node_bits, pos_bits, fp_bits;  // dictionary entry layout
    uint64_t n_nodes, n_kmers, n_eq;
    const NodeRec* nodes;
    const NodeCold* nodes_cold;
    const uint64_t* seq;
    const uint64_t* eq_off;
    const uint32_t* eq_mem;
    const struct ClassWin* class_win;  // one 32-byte window per class (cooperative kernel, wide classes)
    // absent-k-mer filter of k_seed_scan: a split-block Bloom filter over all k-mers, 32-byte blocks (nullptr: none)
    const uint32_t* bloom;
    uint64_t bloom_blocks;
    Dict dict;
};

// Two 64-bit hashes per k-mer, computed once; level l probes h1 + l*h2 (double hashing), the
// fingerprint is the top bits of h2.
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
// bucket of a key at a level; nbkt < 2^32
PSA_HD uint64_t level_bucket(KeyHash kh, uint32_t lvl, uint64_t nbkt) {
    return ((level_hash(kh, lvl) >> 32) * (uint64_t)(uint32_t)nbkt) >> 32;
}
PSA_HD const uint64_t* bucket_addr(const Dict& d, KeyHash kh, uint32_t lvl) {
    return d.buckets + 4 * (d.level_base[lvl] + level_bucket(kh, lvl, d.level_nbkt[lvl]));
}
PSA_HD uint64_t pack_entry(const DevIndex& ix, uint32_t node, uint64_t pos, KeyHash hk) {
    return (uint64_t)node | (pos << ix.node_bits) | (fp_of(hk, ix.fp_bits) << (ix.node_bits + ix.pos_bits));
}
PSA_HD uint32_t entry_node(const DevIndex& ix, uint64_t e) { return (uint32_t)(e & ((1ULL << ix.node_bits) - 1)); }
PSA_HD uint64_t entry_pos(const DevIndex& ix, uint64_t e) { return (e >> ix.node_bits) & ((1ULL << ix.pos_bits) - 1); }
// the entry of bucket b that carries hk's fingerprint (at most one by construction), or kEmptyEntry
PSA_HD uint64_t bucket_find(const DevIndex& ix, const Sector& b, KeyHash hk) {
    const uint32_t shift = ix.node_bits + ix.pos_bits;
    const uint64_t mask = ((1ULL << ix.fp_bits) - 1) << shift;
    const uint64_t want = fp_of(hk, ix.fp_bits) << shift;
    uint64_t r = kEmptyEntry;
    if (((b.w0 ^ want) & mask) == 0 && b.w0 != kEmptyEntry) r = b.w0 & ~kMoreBit;
    if (((b.w1 ^ want) & mask) == 0 && b.w1 != kEmptyEntry) r = b.w1;
    if (((b.w2 ^ want) & mask) == 0 && b.w2 != kEmptyEntry) r = b.w2;
    if (((b.w3 ^ want) & mask) == 0 && b.w3 != kEmptyEntry) r = b.w3;
    return r;
}
// true if a key that hashes to this bucket may live in a later level
PSA_HD bool bucket_more(const Sector& b) { return b.w0 != kEmptyEntry && (b.w0 & kMoreBit) != 0; }

// ---- the filter in front of the dictionary for searches that are mostly misses (k_seed_scan: an unmappable read probes
// 43 absent k-mers, and a dictionary probe -- any random access -- is a 128-byte DRAM line).  Split-block Bloom filter: a
// key sets one bit in each of the eight 32-bit words of ONE 32-byte block, so a query is one sector; at ~10 bits per key
// the filter of a human-scale index (74 MB) stays in L2 for the length of the scan kernel, where nothing else is hot.
// No false negatives, so dict_get's answer is unchanged; ~1 % of absent keys go on to the dictionary.
PSA_HD uint64_t bloom_block_of(KeyHash kh, uint64_t n_blocks) { return mulhi64(kh.h2, n_blocks); }
PSA_HD uint64_t bloom_bits_of(KeyHash kh) { return kh.h1; }   // eight 5-bit fields: bits 0..39 (the dictionary uses the top 32)
PSA_HD bool bloom_test(const DevIndex& ix, KeyHash kh) {
    const Sector b = load_sector_hot(reinterpret_cast<const char*>(ix.bloom) + 32 * bloom_block_of(kh, ix.bloom_blocks));
    const uint64_t x = bloom_bits_of(kh);
    const uint32_t xl = (uint32_t)x, xh = (uint32_t)(x >> 30);   // fields 0..5 in xl, 6..7 in xh
    // words 0..7 of the block are the two halves of w0..w3; 32-bit shifts only
    const uint32_t hit = ((uint32_t)b.w0 >> (xl & 31)) & ((uint32_t)(b.w0 >> 32) >> ((xl >> 5) & 31)) &
                         ((uint32_t)b.w1 >> ((xl >> 10) & 31)) & ((uint32_t)(b.w1 >> 32) >> ((xl >> 15) & 31)) &
                         ((uint32_t)b.w2 >> ((xl >> 20) & 31)) & ((uint32_t)(b.w2 >> 32) >> ((xl >> 25) & 31)) &
                         ((uint32_t)b.w3 >> (xh & 31)) & ((uint32_t)(b.w3 >> 32) >> ((xh >> 5) & 31));
    return (hit & 1u) != 0;
}

struct ProbeStats {  // sequential-equivalent event counts of one dictionary probe
    uint32_t levels, hit, verified;
};

// dbg_index.get(kmer) followed by the reference's verification (src/pseudoaligner.rs:96-107), in
// its blocking form (cooperative kernels, index construction).  A fingerprint mismatch proves that
// the entry's key differs from `key`, so skipping its unitig fetch cannot change the outcome of
// the reference's `read_kmer == ref_kmer` test.
template <int KW>
PSA_HD bool dict_get_hashed(const DevIndex& ix, Kmer<KW> key, KeyHash hk, uint32_t& node, uint32_t& off, ProbeStats* st);
template <int KW>
PSA_HD bool dict_get(const DevIndex& ix, Kmer<KW> key, uint32_t& node, uint32_t& off, ProbeStats* st) {
    return dict_get_hashed<KW>(ix, key, make_hash(KmerOps<KW>::fold(key)), node, off, st);
}
template <int KW>
PSA_HD bool dict_get_hashed(const DevIndex& ix, Kmer<KW> key, KeyHash hk, uint32_t& node, uint32_t& off, ProbeStats* st) {
    if (st) { st->levels = 0; st->hit = 0; st->verified = 0; }
    for (uint32_t lvl = 0; lvl < ix.dict.n_levels; lvl++) {
        const Sector b = load_sector_stream(bucket_addr(ix.dict, hk, lvl));
        if (st) st->levels++;
        const uint64_t e = bucket_find(ix, b, hk);
        if (e != kEmptyEntry) {
            const uint32_t n = entry_node(ix, e);
            const uint64_t pos = entry_pos(ix, e);
            if (st) { st->hit++; st->verified++; }
            // the unitig k-mer and the node's start are independent loads: both addresses come from the entry
#ifdef __CUDA_ARCH__
            const uint64_t start = ld_u64_last(&ix.nodes[n].start_len) & kStartMask;
#else
            const uint64_t start = ix.nodes[n].start_len & kStartMask;
#endif
            const Kmer<KW> ref = KmerOps<KW>::get(GLoad{ix.seq}, pos, ix.k);
            if (ref == key) {
                node = n;
                off = (uint32_t)(pos - start);
                return true;
            }
        }
        if (!bucket_more(b)) return false;
    }
    return false;
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
PSA_HD NodeView node_view_of(const Sector& s) {
    NodeView v;
    v.start = s.w0 & kStartMask;
    v.len = (uint32_t)(s.w0 >> 40);
    v.eq = (uint32_t)s.w1;
    v.class_len = (uint32_t)(s.w1 >> 32);
    v.succ[0] = (uint32_t)s.w2; v.succ[1] = (uint32_t)(s.w2 >> 32); v.succ[2] = (uint32_t)s.w3; v.succ[3] = (uint32_t)(s.w3 >> 32);
    return v;
}
PSA_HD NodeView load_node_view(const NodeRec* r) { return node_view_of(load_sector_hot(r)); }
PSA_HD uint32_t view_succ(const NodeView& v, uint32_t b) { return b == 0 ? v.succ[0] : b == 1 ? v.succ[1] : b == 2 ? v.succ[2] : v.succ[3]; }

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
PSA_HD ClassWin class_win_of(const Sector& s) {  // a ClassWin sector, or sector 1 of a node record
    ClassWin c;
    c.lo = (uint32_t)s.w0;
    c.len = (uint32_t)(s.w0 >> 32);
    c.bits[0] = s.w1; c.bits[1] = s.w2; c.bits[2] = s.w3;
    return c;
}
PSA_HD ClassWin load_class_win(const ClassWin* p) { return class_win_of(load_sector_hot(p)); }
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
// nodes_to_eq_class (ref src/pseudoaligner.rs:323-356) ONLINE, for one thread: the window of
// every newly visited class is ANDed into one accumulator as the walk goes (any number of narrow
// classes in constant space); only wide classes are listed, for the final filter.  eq_id needs
// no list either: the result is contained in every visited class, so it equals a visited class
// iff its size equals the smallest visited class length, and then it is the smallest-id class of
// that length.  A read whose classes do not fit (more than kThreadWide wide ones before any
// narrow one, ...) sets `defer`: the cooperative kernel redoes it from scratch.
// ---------------------------------------------------------------------------------------------
#ifndef PSA_THREAD_WIDE
#define PSA_THREAD_WIDE 3
#endif
constexpr int kThreadWide = PSA_THREAD_WIDE;  // wide classes one read may list
constexpr uint32_t kThreadWideInline = 2;  // further wide classes applied on arrival before giving up
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

struct ClassAcc {  // 60 bytes: it lives in shared memory, one per read in flight
    uint32_t min_len, min_eq;  // smallest class length seen, and the smallest id among the classes of that length
    uint32_t last_eq;          // class of the previous push (consecutive repeats cost nothing)
    uint32_t acc_base;         // AND of the narrow classes' windows: {acc_base + t : bit t of acc_map set}
    uint64_t acc_w0, acc_w1, acc_w2;
    uint32_t wide_eq[kThreadWide];
    uint8_t n_wide, n_inline;
    uint8_t bits;              // kAccHave | kAccMulti | kAccDefer

    static constexpr uint8_t kAccHave = 1, kAccMulti = 2, kAccDefer = 4;
    PSA_HD bool have() const { return bits & kAccHave; }
    PSA_HD bool multi() const { return bits & kAccMulti; }   // more than one distinct class visited
    PSA_HD bool defer() const { return bits & kAccDefer; }
    PSA_HD WinAcc acc() const { return WinAcc{acc_base, Win{acc_w0, acc_w1, acc_w2}, have()}; }
    PSA_HD void set_acc(const WinAcc& a) {
        acc_base = a.base; acc_w0 = a.map.w0; acc_w1 = a.map.w1; acc_w2 = a.map.w2;
        if (a.have) bits |= kAccHave;
    }

    PSA_HD void init() {
        min_len = kNone; min_eq = kNone; last_eq = kNone;
        acc_base = 0; acc_w0 = 0; acc_w1 = 0; acc_w2 = 0;
        PSA_UNROLL
        for (int j = 0; j < kThreadWide; j++) wide_eq[j] = kNone;
        n_wide = 0; n_inline = 0; bits = 0;
    }
    // AND one class into the running intersection (narrow), or list it (wide)
    PSA_HD void and_class(const DevIndex& ix, uint32_t e, uint32_t l, const ClassWin& c) {
        if (c.len != kWinWide) {
            WinAcc a = acc();
            winacc_and(a, c);
            set_acc(a);
            return;
        }
        bool dup = false;
        PSA_UNROLL
        for (int j = 0; j < kThreadWide; j++) dup |= (j < (int)n_wide && wide_eq[j] == e);
        if (dup) return;
        if (n_wide >= kThreadWide) {
            // no room to remember it: apply it now to the candidates the windows have left (the filter
            // is idempotent and commutes with the ANDs still to come); without any window yet, give up
            if (have() && n_inline < kThreadWideInline) {
                n_inline++;
                WinAcc a = acc();
                winacc_filter_list(a, ix.eq_mem + ld_off(ix.eq_off + e), l);
                set_acc(a);
                return;
            }
            bits |= kAccDefer;
            return;
        }
        PSA_UNROLL
        for (int j = 0; j < kThreadWide; j++)
            if (j == (int)n_wide) wide_eq[j] = e;
        n_wide++;
    }
    // nodes.push: only the classes matter, and the intersection is idempotent (ref :352-355).
    // Returns true when the class was not the previous push's (event counting).
    PSA_HD bool push(const DevIndex& ix, uint32_t eq, uint32_t class_len, const ClassWin& win) {
        if (eq == last_eq) return false;
        if (min_eq != kNone) bits |= kAccMulti;
        last_eq = eq;
        if (class_len < min_len || (class_len == min_len && eq < min_eq)) {
            min_len = class_len;
            min_eq = eq;
        }
        and_class(ix, eq, class_len, win);
        return true;
    }
};

// All visited classes are wide: the reference's list scheme by one thread.  Members of the
// smallest listed class (index s) that every other listed class contains, ascending; each other
// class is searched only in the suffix after its previous match (ref :399-404).
// out == nullptr counts.
PSA_HD uint32_t thread_intersect_lists(const DevIndex& ix, const ClassAcc& w, int s, uint32_t* out) {
    const uint32_t* mem = ix.eq_mem;
    uint32_t cur[kThreadWide];
    PSA_UNROLL
    for (int j = 0; j < kThreadWide; j++) cur[j] = 0;
    uint32_t s_eq = 0;
    PSA_UNROLL
    for (int j = 0; j < kThreadWide; j++)
        if (j == s) s_eq = w.wide_eq[j];
    const uint64_t s_off = ld_off(ix.eq_off + s_eq);
    const uint32_t s_len = (uint32_t)(ld_off(ix.eq_off + s_eq + 1) - s_off);
    uint32_t count = 0;
    for (uint32_t i = 0; i < s_len; i++) {
        const uint32_t x = ld_mem(mem + s_off + i);
        bool alive = true;
        PSA_UNROLL
        for (int j = 0; j < kThreadWide; j++) {
            if (j == s || j >= (int)w.n_wide || !alive) continue;
            const uint64_t o = ld_off(ix.eq_off + w.wide_eq[j]);
            const uint32_t n = (uint32_t)(ld_off(ix.eq_off + w.wide_eq[j] + 1) - o);
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

// The eq_class of a read from its accumulated classes: |set| and the id of the visited class it
// equals (kNone if none).  `s` receives the smallest wide class (used when every class is wide).
// Returns false when the lists are too long for one thread (smallest class > max_small).
PSA_HD bool class_result(const DevIndex& ix, ClassAcc& w, uint32_t max_small, uint32_t& count, uint32_t& eq_id, int& s) {
    s = 0;
    if (!w.multi()) {
        count = w.min_len;
        eq_id = w.min_eq;
        return true;
    }
    if (w.have()) {
        // wide classes filter what survived the windows (ref :399-404 on the candidates)
        WinAcc a = w.acc();
        PSA_UNROLL
        for (int j = 0; j < kThreadWide; j++) {
            if (j >= (int)w.n_wide || win_empty(a.map)) continue;
            const uint64_t o = ld_off(ix.eq_off + w.wide_eq[j]);
            winacc_filter_list(a, ix.eq_mem + o, (uint32_t)(ld_off(ix.eq_off + w.wide_eq[j] + 1) - o));
        }
        w.set_acc(a);
        count = win_popc(a.map);
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
        if (s_len > max_small) return false;  // long lists are the cooperative kernel's job
        count = thread_intersect_lists(ix, w, s, (uint32_t*)nullptr);
    }
    // the result equals a visited class iff it has as many members as the smallest visited class
    eq_id = count == w.min_len ? w.min_eq : kNone;
    return true;
}
// the members of a result that is no visited class, ascending
PSA_HD void class_members(const DevIndex& ix, const ClassAcc& w, int s, uint32_t* dst) {
    if (w.have()) win_write(w.acc(), dst);
    else thread_intersect_lists(ix, w, s, dst);
}

// ASCII -> 2-bit code, DnaString::from_dna_string (call site ref src/pseudoaligner.rs:450):
// A/a 0, C/c 1, G/g 2, T/t 3, every other byte 0.
PSA_HD uint32_t base_code(uint8_t c) {
    c &= 0xDF;  // fold case (only matters for letters)
    return c == 'C' ? 1u : c == 'G' ? 2u : c == 'T' ? 3u : 0u;
}

}  // namespace psa
