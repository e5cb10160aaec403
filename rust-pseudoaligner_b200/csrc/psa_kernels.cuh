// psa_kernels.cuh -- the sm_100a kernels: index construction on the device (k-mer enumeration,
// bucket-cascade dictionary, successor/predecessor tables, class windows), ASCII -> 2-bit
// packing, the map kernels, and the result expansion.
//
// Reference items replaced (10XGenomics/rust-pseudoaligner @ 9d9cab8):
//   k_map_thread / k_seed_scan / k_map   Pseudoaligner::map_read + the per-record body of
//                    process_reads (src/pseudoaligner.rs:64-384, :449-462)
//   k_pack_ascii*    DnaString::from_dna_string at src/pseudoaligner.rs:449-450
//   k_dict_*         make_dbg_index (src/build_index.rs:182-221)
//   k_build_edges    what Node::r_edges()/l_edges() compute per call (src/pseudoaligner.rs:191,275)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "psa_core.cuh"
#include "psa_thread.cuh"
#include "psa_thread.cuh"

namespace psa {

constexpr unsigned kFull = 0xffffffffu;

// ---------------------------------------------------------------------------------------------
// index construction
// ---------------------------------------------------------------------------------------------
// first pass over the nodes: record everything that does not need the dictionary
__global__ void k_node_basics(NodeRec* nodes, NodeCold* cold, uint64_t n_nodes, const uint64_t* node_start,
                              const uint32_t* node_len, const uint8_t* node_exts, const uint32_t* node_eq,
                              const uint64_t* eq_off, uint64_t n_eq, uint32_t k, uint32_t* err) {
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= n_nodes) return;
    NodeRec r;
    NodeCold c;
    const uint32_t len = node_len[i];
    r.start_len = pack_start_len(node_start[i], len);
    r.eq = node_eq[i];
    c.exts = node_exts[i];
    c.pad = 0;
    if (r.eq >= n_eq || len < k || len > kMaxNodeLen || node_start[i] > kStartMask) {
        atomicOr(err, 1u);
        r.class_len = 0;
        c.class_off = 0;
    } else {
        c.class_off = eq_off[r.eq];
        r.class_len = (uint32_t)(eq_off[r.eq + 1] - c.class_off);
    }
    for (int b = 0; b < 4; b++) r.succ[b] = c.pred[b] = kNone;
    r.win_lo = 0; r.win_len = 0; r.win_bits[0] = r.win_bits[1] = r.win_bits[2] = 0;
    nodes[i] = r;
    cold[i] = c;
}
// sector 1 of every node record: the window of the node's class
__global__ void k_node_windows(NodeRec* nodes, uint64_t n_nodes, uint64_t n_eq, const ClassWin* win) {
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= n_nodes) return;
    const uint32_t e = nodes[i].eq;
    if (e >= n_eq) return;
    const ClassWin c = win[e];
    nodes[i].win_lo = c.lo; nodes[i].win_len = c.len;
    nodes[i].win_bits[0] = c.bits[0]; nodes[i].win_bits[1] = c.bits[1]; nodes[i].win_bits[2] = c.bits[2];
}

// every k-mer of every node: key and its (node, offset); koff = exclusive scan of (len-k+1)
template <int KW>
__global__ void k_enumerate_keys(const uint64_t* seq, const uint64_t* node_start, const uint64_t* koff,
                                 uint64_t n_nodes, uint64_t n_kmers, uint32_t k, uint64_t* key_lo,
                                 uint64_t* key_hi, uint64_t* val) {
    uint64_t g = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (g >= n_kmers) return;
    uint64_t lo = 0, hi = n_nodes;  // last node with koff[node] <= g
    while (hi - lo > 1) {
        uint64_t mid = (lo + hi) >> 1;
        if (koff[mid] <= g) lo = mid;
        else hi = mid;
    }
    uint64_t off = g - koff[lo];
    Kmer<KW> key = KmerOps<KW>::get(GLoad{seq}, node_start[lo] + off, k);
    key_lo[g] = key.lo;
    if constexpr (KW == 2) key_hi[g] = key.hi;
    val[g] = (lo << 32) | off;
}

template <int KW>
__device__ __forceinline__ KeyHash key_hash_at(const uint64_t* key_lo, const uint64_t* key_hi, uint64_t i) {
    if (KW == 1) {
        Kmer<1> x; x.lo = key_lo[i];
        return make_hash(KmerOps<1>::fold(x));
    } else {
        Kmer<2> x; x.lo = key_lo[i]; x.hi = key_hi[i];
        return make_hash(KmerOps<2>::fold(x));
    }
}

// ---- the dictionary (psa_core.cuh Dict), one cascade level at a time --------------------------
// step 1: the bucket of every remaining key at this level (sort key) and its index (sort value)
template <int KW>
__global__ void k_dict_bucket_ids(const uint64_t* key_lo, const uint64_t* key_hi, uint32_t n, uint32_t lvl, uint64_t nbkt,
                                  uint32_t* bid, uint32_t* idx) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    bid[i] = (uint32_t)level_bucket(key_hash_at<KW>(key_lo, key_hi, i), lvl, nbkt);
    idx[i] = i;
}
// step 2 (after a stable sort by bucket): the thread at the head of a bucket's run fills the bucket with
// the first keys of the run, at most kBucketSlots of them and no two with the same fingerprint; the rest
// are flagged for the next level and the bucket's "more" bit is set.  Deterministic: the order of the
// keys is the order of their enumeration.
template <int KW>
__global__ void k_dict_fill(const uint64_t* key_lo, const uint64_t* key_hi, const uint64_t* val, uint32_t n,
                            const uint32_t* bid_sorted, const uint32_t* idx_sorted, DevIndex ix, uint64_t base,
                            uint64_t* buckets, uint8_t* passed_on) {
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const uint32_t b = bid_sorted[j];
    if (j > 0 && bid_sorted[j - 1] == b) return;  // not the head of its run
    uint64_t e[kBucketSlots];
    uint64_t fps[kBucketSlots];
    uint32_t cnt = 0;
    bool more = false;
    for (uint32_t t = j; t < n && bid_sorted[t] == b; t++) {
        const uint32_t i = idx_sorted[t];
        const KeyHash hk = key_hash_at<KW>(key_lo, key_hi, i);
        const uint64_t fp = fp_of(hk, ix.fp_bits);
        bool clash = cnt >= kBucketSlots;
        for (uint32_t q = 0; q < cnt && q < kBucketSlots; q++) clash |= fps[q] == fp;
        if (clash) {
            passed_on[i] = 1;
            more = true;
            continue;
        }
        passed_on[i] = 0;
        const uint64_t v = val[i];
        const uint32_t node = (uint32_t)(v >> 32);
        e[cnt] = pack_entry(ix, node, (ix.nodes[node].start_len & kStartMask) + (uint32_t)v, hk);
        fps[cnt] = fp;
        cnt++;
    }
    uint64_t* dst = buckets + 4 * (base + b);
    for (uint32_t q = 0; q < cnt; q++) dst[q] = q == 0 && more ? (e[q] | kMoreBit) : e[q];
}
// step 3: the keys passed on, compacted in order (sel = their indices, ascending)
__global__ void k_dict_gather(const uint64_t* key_lo, const uint64_t* key_hi, const uint64_t* val, const uint32_t* sel, uint32_t n,
                              uint64_t* out_lo, uint64_t* out_hi, uint64_t* out_val) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t s = sel[i];
    out_lo[i] = key_lo[s];
    if (key_hi) out_hi[i] = key_hi[s];
    out_val[i] = val[s];
}
// self-check after the build: every key resolves to its own (node, offset)
template <int KW>
__global__ void k_dict_check(const uint64_t* key_lo, const uint64_t* key_hi, const uint64_t* val, uint64_t n, DevIndex ix,
                             uint32_t* err) {
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    Kmer<KW> key;
    key.lo = key_lo[i];
    if constexpr (KW == 2) key.hi = key_hi[i];
    uint32_t node = 0, off = 0;
    const uint64_t v = val[i];
    if (!dict_get<KW>(ix, key, node, off, nullptr) || node != (uint32_t)(v >> 32) || off != (uint32_t)v) atomicOr(err, 2u);
}

// succ[b] / pred[b]: the node whose first k-mer is last(k-1)+b, resp. whose last k-mer is
// b+first(k-1).  debruijn's find_link expects every ext bit to resolve ("missing link").
template <int KW>
__global__ void k_build_edges(DevIndex ix, NodeRec* nodes, NodeCold* cold, uint32_t* err) {
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= ix.n_nodes) return;
    const NodeRec r = nodes[i];
    const uint32_t exts = cold[i].exts;
    const uint64_t r_start = r.start_len & kStartMask;
    const uint32_t r_len = (uint32_t)(r.start_len >> 40);
    if (r_len < ix.k) return;
    Kmer<KW> first = KmerOps<KW>::get(GLoad{ix.seq}, r_start, ix.k);
    Kmer<KW> last = KmerOps<KW>::get(GLoad{ix.seq}, r_start + r_len - ix.k, ix.k);
    for (uint32_t b = 0; b < 4; b++) {
        uint32_t n, o;
        if ((exts >> b) & 1) {
            if (dict_get<KW>(ix, KmerOps<KW>::extend_right(last, b, ix.k), n, o, nullptr) && o == 0)
                nodes[i].succ[b] = n;
            else
                atomicOr(err, 4u);
        }
        if ((exts >> (4 + b)) & 1) {
            if (dict_get<KW>(ix, KmerOps<KW>::extend_left(first, b, ix.k), n, o, nullptr) &&
                o == (uint32_t)(nodes[n].start_len >> 40) - ix.k)
                cold[i].pred[b] = n;
            else
                atomicOr(err, 4u);
        }
    }
}

// one 32-byte window per class (psa_core.cuh "Class windows")
__global__ void k_build_class_win(const uint64_t* eq_off, const uint32_t* eq_mem, uint64_t n_eq, ClassWin* out) {
    uint64_t c = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (c >= n_eq) return;
    const uint64_t o = eq_off[c];
    out[c] = make_class_win(eq_mem + o, eq_off[c + 1] - o);
}

template <int KW>
__global__ void k_lookup(DevIndex ix, const uint64_t* kmer_words, uint64_t n, uint8_t* found, uint32_t* node,
                         uint32_t* off) {
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    Kmer<KW> key = KmerOps<KW>::get(PLoad{kmer_words + (uint64_t)KW * i}, 0, ix.k);
    uint32_t nn = 0, oo = 0;
    bool f = dict_get<KW>(ix, key, nn, oo, nullptr);
    found[i] = f;
    node[i] = nn;
    off[i] = oo;
}

// ---------------------------------------------------------------------------------------------
// reads: ASCII -> DnaString words (DnaString::from_dna_string, ref src/pseudoaligner.rs:449-450)
// ---------------------------------------------------------------------------------------------
struct ReadsView {          // device-resident batch
    const uint64_t* words;  // packed
    const uint64_t* woff;   // word offset of read i; nullptr: i * wstride
    const uint32_t* len;    // nullptr: fixed_len
    uint64_t wstride;
    uint32_t fixed_len;
    uint64_t n;
};

__global__ void k_words_per_read(const uint32_t* len, uint64_t n, uint64_t* nw) {
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i < n) nw[i] = ((uint64_t)len[i] + 31) >> 5;
    else if (i == n) nw[i] = 0;  // the scan runs over n + 1 items so that its last output is the total
}

// 4 ASCII bytes (first base in the low byte) -> 8 bits of 2-bit codes, first base in bits 7:6.
// A/a 0, C/c 1, G/g 2, T/t 3, any other byte 0 -- bytewise, two 8-entry table lookups (PRMT):
// the low three bits of A C G T (1 3 7 4) are distinct, so they select both the code and the
// letter the byte must be (case folded) for the code to stand.
__device__ __forceinline__ uint32_t codes4(uint32_t w) {
    uint32_t t = w & 0x07070707u;
    t |= t >> 4;                                              // byte 0: idx0 | idx1 << 4, byte 2: idx2 | idx3 << 4
    const uint32_t sel = __byte_perm(t, 0u, 0x4420u);         // four selector nibbles (bit 3 of each is clear)
    uint32_t x = __byte_perm(0x01000000u, 0x02000003u, sel);  // idx 1 (A) 0, 3 (C) 1, 7 (G) 2, 4 (T) 3, else 0
    const uint32_t want = __byte_perm(0x43FF41FFu, 0x47FFFF54u, sel);  // the letter with those low bits (0xFF: none)
    const uint32_t d = (w & 0xDFDFDFDFu) ^ want;              // byte == 0 iff the (case-folded) byte is that letter
    const uint32_t nz = (((d & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | d) & 0x80808080u;  // bit 7 of every non-zero byte
    x &= ~((nz >> 7) | (nz >> 6));
    return (x * 0x40100401u) >> 24;       // gather the four 2-bit fields (no carries between them)
}
// nb (1..32) bases at s -> one DnaString word.  Reads only the aligned 32-bit words that hold
// at least one of the nb bytes.
__device__ __forceinline__ uint64_t pack32(const uint8_t* s, uint32_t nb) {
    const uint32_t mis = (uint32_t)((uintptr_t)s & 3);
    const uint32_t* a = reinterpret_cast<const uint32_t*>(s - mis);
    const uint32_t nwords = (mis + nb + 3) >> 2;  // <= 9
    uint32_t w[9];
#pragma unroll
    for (int i = 0; i < 9; i++) w[i] = (uint32_t)i < nwords ? __ldg(a + i) : 0u;
    uint64_t v = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint32_t c = __funnelshift_r(w[i], w[i + 1], 8 * mis);
        v |= (uint64_t)codes4(c) << (56 - 8 * i);
    }
    if (nb < 32) v &= ~0ULL << (64 - 2 * nb);
    return v;
}
// ---- bulk asynchronous copy (TMA, non-tensor form) + mbarrier, for the read tiles ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.b32 %0, 1, 0, P1;\n\t}"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(phase) : "memory");
    }
}
// fixed read length and stride, tile form: the ASCII of the CTA's 128 reads arrives in shared memory as ONE
// bulk asynchronous copy (cp.async.bulk + mbarrier: TMA), then every thread packs its own read from there
// (nine aligned 32-bit shared loads per word) and writes its words.  Replaces nine 4-byte global loads per
// word whose lanes touch 32 different sectors each.
constexpr int kPackTileReads = 128;
__device__ __forceinline__ uint64_t pack32_smem(const uint8_t* s, uint32_t nb) {
    const uint32_t mis = (uint32_t)(smem_u32(s) & 3);
    const uint32_t* a = reinterpret_cast<const uint32_t*>(s - mis);
    uint32_t w[9];
#pragma unroll
    for (int i = 0; i < 9; i++) w[i] = a[i];
    uint64_t v = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint32_t c = __funnelshift_r(w[i], w[i + 1], 8 * mis);
        v |= (uint64_t)codes4(c) << (56 - 8 * i);
    }
    if (nb < 32) v &= ~0ULL << (64 - 2 * nb);
    return v;
}
__global__ void __launch_bounds__(kPackTileReads) k_pack_ascii_tile(const uint8_t* ascii, uint32_t astride, uint32_t len,
                                                                   uint64_t n, uint64_t* words, uint32_t tile_cap) {
    // layout: [16 B slack | tile (tile_cap, multiple of 16) | 48 B slack | mbarrier]
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* tile = smem + 16;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 16 + tile_cap + 48);
    const uint32_t nw = (len + 31) >> 5;
    const uint64_t r0 = blockIdx.x * (uint64_t)kPackTileReads;
    const uint32_t nr = (uint32_t)min((uint64_t)kPackTileReads, n - r0);
    const uint8_t* src = ascii + r0 * astride;
    // a full tile is one bulk copy (the host guarantees 16-byte multiples and alignment); the batch's last
    // tile may not read the padding after its last read, which need not exist
    const bool full = nr == kPackTileReads && (r0 + kPackTileReads < n || astride == len);
    if (full) {
        if (threadIdx.x == 0) {
            mbar_init(bar, 1);
            mbar_expect_tx(bar, kPackTileReads * astride);
            bulk_g2s(tile, src, kPackTileReads * astride, bar);
        }
        __syncthreads();  // the barrier is initialised before anyone polls it
        mbar_wait(bar, 0);
    } else {
        const uint32_t bytes = (nr - 1) * astride + len;
        for (uint32_t i = threadIdx.x; i < bytes; i += kPackTileReads) tile[i] = src[i];
        __syncthreads();
    }
    if (threadIdx.x < nr) {
        const uint8_t* s = tile + threadIdx.x * astride;
        uint64_t* w = words + (r0 + threadIdx.x) * nw;
        for (uint32_t j = 0; j < nw; j++) w[j] = pack32_smem(s + 32 * j, min(32u, len - 32 * j));
    }
}
// fixed read length: one thread per output word (32-bit index arithmetic: the host splits larger batches)
__global__ void k_pack_ascii_fixed(const uint8_t* ascii, uint64_t astride, uint32_t len, uint64_t n, uint64_t* words) {
    const uint32_t nw = (len + 31) >> 5;
    const uint32_t total = (uint32_t)(n * nw);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const uint32_t r = i / nw;
        const uint32_t j = i - r * nw;
        words[i] = pack32(ascii + r * astride + 32 * j, min(32u, len - 32 * j));
    }
}
// ragged: warp per read, lane-strided over its words
__global__ void k_pack_ascii(const uint8_t* ascii, const uint64_t* aoff, uint64_t astride, const uint32_t* len,
                             uint32_t fixed_len, const uint64_t* woff, uint64_t wstride, uint64_t n,
                             uint64_t* words) {
    uint64_t warp = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
    uint32_t lane = threadIdx.x & 31;
    uint64_t nwarps = (gridDim.x * (uint64_t)blockDim.x) >> 5;
    for (uint64_t r = warp; r < n; r += nwarps) {
        const uint8_t* s = ascii + (aoff ? aoff[r] : r * astride);
        uint32_t L = len ? len[r] : fixed_len;
        uint64_t* w = words + (woff ? woff[r] : r * wstride);
        uint32_t nw = (L + 31) >> 5;
        for (uint32_t j = lane; j < nw; j += 32) w[j] = pack32(s + 32 * j, min(32u, L - 32 * j));
    }
}

// ---------------------------------------------------------------------------------------------
// the map kernel: G lanes (a power-of-two slice of a warp) cooperate on one read
// ---------------------------------------------------------------------------------------------

struct MapParams {
    ReadsView reads;
    HitRec* hits;
    unsigned long long* counts;   // n_eq + 2, or nullptr
    uint32_t* novel;              // members of sets that are no index class
    unsigned long long novel_cap;
    unsigned long long* novel_cursor;
    uint32_t* novel_list;         // reads whose eq_class is no visited class, in no particular order
    unsigned long long* novel_list_count;
    uint4* spill;                 // per-group overflow of the visited-class list
    uint32_t spill_cap;           // entries per group
    uint4* pool;                  // bump-allocated overflow of `spill` (very long reads)
    unsigned long long pool_cap;  // entries
    unsigned long long* pool_cursor;
    uint32_t allowed_mismatches;
    // deferred reads: written by k_map_thread, consumed by k_map (list != nullptr: map list[0..*list_count))
    uint32_t* list;
    unsigned long long* list_count;
    unsigned long long* work_cursor;  // k_map over the list: next entry to claim (zeroed per batch), or nullptr
    unsigned long long* hint_cursor;  // second pass of k_map_thread: next 32 entries of `seeded` to claim (zeroed per batch)
    // k_map in two launches (the first one overlaps k_seed_scan and the second pass of the thread kernel on another stream):
    const unsigned long long* list_first;  // first entry of this launch (nullptr: 0)
    const unsigned long long* list_end;    // one past its last entry (nullptr: *list_count)
    // reads whose FIRST seed search was too long for one thread: k_map_thread -> k_seed_scan
    uint32_t* scan_list;
    unsigned long long* scan_count;
    // reads k_seed_scan found a seed for: {read, pos, node, off} -> second pass of k_map_thread
    uint4* seeded;
    unsigned long long* seeded_count;
    uint4* seeded_ev;             // event counting only: {lookups, levels, hits, verifs} of that search
    uint32_t max_probes;          // seed positions one thread tries in a read's first search
    uint32_t reseed_probes;       // k_map_thread: ... and in a re-seed search (first pass: small, the read is redone by the second pass)
    uint32_t max_small;           // k_map_thread: largest smallest-class one thread intersects
    uint32_t* status;             // bit0: novel buffer overflow, bit1: spill pool overflow
    unsigned long long* events;   // 3 x psa_events layout ([0] k_map_thread, [1] k_map, [2] k_seed_scan), or nullptr
};

struct LaneEvents {
    uint32_t lookups, levels, hits, verifs, visits, bases, jumps, members;
};

// the lanes of one read: a G-wide slice of the warp (cooperative-groups tile, hand-rolled)
template <int G>
struct Grp {
    uint32_t lane, shift;
    unsigned mask;
    static constexpr unsigned kLow = G == 32 ? 0xffffffffu : ((1u << (G & 31)) - 1u);
    __device__ __forceinline__ Grp() {
        uint32_t wl = threadIdx.x & 31;
        lane = wl & (G - 1);
        shift = wl & ~(uint32_t)(G - 1);
        mask = kLow << shift;
    }
    __device__ __forceinline__ unsigned ballot(bool p) const { return (__ballot_sync(mask, p) >> shift) & kLow; }
    __device__ __forceinline__ bool any(bool p) const { return __any_sync(mask, p) != 0; }
    template <class T> __device__ __forceinline__ T shfl(T v, int src) const { return __shfl_sync(mask, v, src, G); }
    template <class T> __device__ __forceinline__ T shfl_up(T v, int d) const { return __shfl_up_sync(mask, v, d, G); }
    template <class T> __device__ __forceinline__ T shfl_xor(T v, int d) const { return __shfl_xor_sync(mask, v, d, G); }
    __device__ __forceinline__ void sync() const { __syncwarp(mask); }
};

template <int KW, bool EV, int G>
struct WarpCtx {
    const DevIndex& ix;
    PLoad rd;
    Grp<G> g;
    uint32_t lane;
    uint32_t k;
    // visited distinct classes: entry j < G lives in lane j, further ones in `spill`
    uint32_t my_eq, my_len, n_list;
    uint64_t my_off;
    uint4* spill;
    uint32_t spill_cap;
    bool spill_overflow;
    // a read can push at most read_len+1 nodes, so one pool allocation of that size always suffices
    uint4* pool;
    unsigned long long pool_cap;
    unsigned long long* pool_cursor;
    uint32_t read_len;
    bool scan_mode;
    uint32_t scan_skip = 0;  // scan mode: first position not yet known to be absent
    // scan mode, reads of at most 32 * G bases: lane j holds word j of the read, k-mers are cut from shuffled words
    // (the scan otherwise re-fetches two read words from L2 for every probe)
    bool staged = false;
    uint64_t my_word = 0;
    LaneEvents ev;

    __device__ __forceinline__ WarpCtx(const DevIndex& ix_, const uint64_t* read_words, uint32_t read_len_,
                                       const MapParams& p, uint64_t group_id)
        : ix(ix_), rd{read_words}, g(), lane(g.lane), k(ix_.k), my_eq(kNone), my_len(0), n_list(0), my_off(0),
          spill(p.spill + group_id * p.spill_cap), spill_cap(p.spill_cap), spill_overflow(false), pool(p.pool),
          pool_cap(p.pool_cap), pool_cursor(p.pool_cursor), read_len(read_len_), scan_mode(false), ev{} {}

    __device__ __forceinline__ uint32_t read_base(uint64_t pos) const { return seq_get(rd, pos); }
    __device__ __forceinline__ bool abort() const { return false; }
    __device__ __forceinline__ void stage_words() {
        const uint32_t nw = (read_len + 31) >> 5;
        staged = nw <= (uint32_t)G;
        if (staged) my_word = lane < nw ? rd(lane) : 0;
    }
    // the k-mer at position p from the staged words; every lane of the group calls it (p may differ per lane)
    __device__ __forceinline__ Kmer<KW> staged_kmer(uint32_t p) const {
        const uint32_t wi = p >> 5;
        Sector s;
        s.w0 = g.shfl(my_word, (int)wi);
        s.w1 = g.shfl(my_word, (int)min(wi + 1, (uint32_t)G - 1));
        s.w2 = KW == 2 ? g.shfl(my_word, (int)min(wi + 2, (uint32_t)G - 1)) : 0;
        s.w3 = 0;
        return KmerOps<KW>::get(WLoad{s, wi}, p, k);
    }

    // find_kmer_match, ref src/pseudoaligner.rs:91-114.  The first position is probed by the
    // whole group on one address (the common case: it hits); after a miss, G stride-3
    // positions are probed at once, one per lane, and the lowest hitting lane wins -- the
    // same answer as the sequential scan.
    template <class P>
    __device__ __forceinline__ bool find_seed(P& kmer_pos, P last, uint32_t& node, uint32_t& off) {
        if (kmer_pos > last) return false;
        ProbeStats st;
        const P start = kmer_pos;
        P first = start;
        if (!scan_mode) {  // (k_seed_scan's reads have already missed here: all lanes speculate from `start`)
            Kmer<KW> key = KmerOps<KW>::get(rd, kmer_pos, k);
            bool hit = dict_get<KW>(ix, key, node, off, EV ? &st : nullptr);
            if (EV && lane == 0) { ev.lookups++; ev.levels += st.levels; ev.hits += st.hit; ev.verifs += st.verified; }
            if (hit) return true;
            first = start + kSeedStride;
        } else {
            first = start + (P)scan_skip;  // positions the thread-per-read kernel has already found absent
        }
        for (P cur = first; cur <= last; cur += G * kSeedStride) {
            P p = cur + (P)kSeedStride * lane;
            bool h = false;
            uint32_t n = 0, o = 0;
            st.levels = st.hit = st.verified = 0;
            const bool in = p <= last;
            Kmer<KW> key{};
            if (staged) key = staged_kmer(in ? (uint32_t)p : 0u);
            else if (in) key = KmerOps<KW>::get(rd, p, k);
            unsigned b;
            if (!EV && scan_mode && ix.bloom) {
                // k_seed_scan: most of these k-mers are absent, and a dictionary probe is a 128-byte DRAM line plus the
                // verification loads.  Every lane asks the L2-resident filter; then only the LOWEST position that may be
                // present goes to the dictionary (for a mappable read that is its seed: one probe instead of up to G, the
                // positions behind a seed being present as well), the next one only after a false positive.
                KeyHash hk{0, 0};
                bool maybe = false;
                if (in) {
                    hk = make_hash(KmerOps<KW>::fold(key));
                    maybe = bloom_test(ix, hk);
                }
                unsigned mb = g.ballot(maybe);
                b = 0;
                while (mb) {
                    const int jj = __ffs(mb) - 1;
                    if ((int)lane == jj) h = dict_get_hashed<KW>(ix, key, hk, n, o, nullptr);
                    if (g.shfl((int)h, jj)) {
                        b = 1u << jj;
                        break;
                    }
                    mb &= mb - 1;
                }
            } else {
                if (in) h = dict_get<KW>(ix, key, n, o, EV ? &st : nullptr);
                b = g.ballot(h);
            }
            int j = b ? (__ffs(b) - 1) : G;
            if (EV) {  // sequential-equivalent events: the probes up to and including the first hit
                const bool counted = p <= last && (int)lane <= j;
                uint32_t c0 = counted, c1 = counted ? st.levels : 0, c2 = counted ? st.hit : 0, c3 = counted ? st.verified : 0;
#pragma unroll
                for (int d = G / 2; d; d >>= 1) {
                    c0 += g.shfl_xor(c0, d); c1 += g.shfl_xor(c1, d); c2 += g.shfl_xor(c2, d); c3 += g.shfl_xor(c3, d);
                }
                if (lane == 0) { ev.lookups += c0; ev.levels += c1; ev.hits += c2; ev.verifs += c3; }
            }
            if (b) {
                node = g.shfl(n, j);
                off = g.shfl(o, j);
                kmer_pos = cur + (P)kSeedStride * j;
                return true;
            }
        }
        kmer_pos = start + kSeedStride * ((last - start) / kSeedStride + 1);  // where the loop at :92-111 stops
        return false;
    }

    __device__ __forceinline__ NodeView node(uint32_t id) const { return load_node_view(ix.nodes + id); }
    __device__ __forceinline__ void jumped() {
        if (EV && lane == 0) ev.jumps++;
    }
    __device__ __forceinline__ uint32_t pred(uint32_t id, uint32_t b) {
        const uint32_t p = __ldg(&ix.nodes_cold[id].pred[b]);
        if (EV && lane == 0 && p != kNone) ev.jumps++;
        return p;
    }

    // shared tail of the two compare loops: lane `lane` holds the mismatch mask of bases
    // [base+32*lane, base+32*lane+n); find the (A+1)-th mismatch in scan order.
    __device__ __forceinline__ bool locate_break(uint64_t mask, uint32_t& snp, uint32_t A, uint32_t& t_out, int& j_out) {
        uint32_t c = (uint32_t)popc64(mask);
        uint32_t incl = c;
#pragma unroll
        for (int d = 1; d < G; d <<= 1) {
            uint32_t t = g.shfl_up(incl, d);
            if ((int)lane >= d) incl += t;
        }
        unsigned b = g.ballot(snp + incl > A);
        if (b) {
            int j = __ffs(b) - 1;
            uint32_t t = 0;
            if ((int)lane == j) t = nth_mismatch(mask, A + 1 - (snp + incl - c));
            t_out = g.shfl(t, j);
            j_out = j;
            return true;
        }
        snp += g.shfl(incl, G - 1);
        return false;
    }

    // ref src/pseudoaligner.rs:234-255 (FWD) and :149-170 (backward)
    template <bool FWD, class P>
    __device__ __forceinline__ P cmp(P rp, uint64_t sp, P m, uint32_t A, bool& premature) {
        uint32_t snp = 0;
        for (P base = 0; base < m; base += 32 * G) {
            P my = base + 32 * lane;
            uint32_t n = my < m ? (uint32_t)min((P)32, (P)(m - my)) : 0;
            uint64_t mask = 0;
            if (n) mask = FWD ? mismatch_fwd(rd, rp + my, GLoad{ix.seq}, sp + my, n)
                              : mismatch_bwd(rd, rp - my, GLoad{ix.seq}, sp - my, n);
            if (!g.any(mask != 0)) continue;
            uint32_t t; int j;
            if (locate_break(mask, snp, A, t, j)) {
                premature = true;
                P matched = base + 32 * (P)j + t;
                if (EV && lane == 0) ev.bases += (uint32_t)matched + 1;
                return matched;
            }
        }
        if (EV && lane == 0) ev.bases += (uint32_t)m;
        return m;
    }
    template <class P>
    __device__ __forceinline__ P cmp_fwd(P rp, uint64_t sp, P m, uint32_t A, bool& pb) { return cmp<true>(rp, sp, m, A, pb); }
    template <class P>
    __device__ __forceinline__ P cmp_bwd(P rp, uint64_t sp, P m, uint32_t A, bool& pb) { return cmp<false>(rp, sp, m, A, pb); }

    // nodes.push (ref :199, :219), keeping only what nodes_to_eq_class needs: the distinct
    // classes of the visited nodes (intersection is idempotent, ref :352-355).
    __device__ __forceinline__ void push(uint32_t /*node_id*/, const NodeView& nv) {
        if (EV && lane == 0) ev.visits++;
        if (g.ballot(lane < n_list && my_eq == nv.eq)) return;
        if (n_list < G) {
            if (lane == n_list) { my_eq = nv.eq; my_len = nv.class_len; my_off = __ldg(ix.eq_off + nv.eq); }
            n_list++;
            return;
        }
        // rare: more than G distinct classes
        uint32_t ns = n_list - G;
        bool dup = false;
        for (uint32_t j = lane; j < ns; j += G) dup |= (spill[j].x == nv.eq);
        if (g.any(dup)) return;
        if (ns >= spill_cap && !grow_spill(ns)) { spill_overflow = true; return; }
        if (lane == 0) {
            const uint64_t coff = __ldg(ix.eq_off + nv.eq);
            spill[ns] = make_uint4(nv.eq, nv.class_len, (uint32_t)coff, (uint32_t)(coff >> 32));
        }
        g.sync();
        n_list++;
    }
    // the per-group spill is full: move the list to a pool allocation sized for this read
    __device__ __forceinline__ bool grow_spill(uint32_t ns) {
        const unsigned long long need = (unsigned long long)read_len + 2;
        if (need <= spill_cap) return false;  // already grown: cannot happen (pushes <= read_len + 1)
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(pool_cursor, need);
        base = g.shfl(base, 0);
        if (base + need > pool_cap) return false;
        uint4* dst = pool + base;
        for (uint32_t j = lane; j < ns; j += G) dst[j] = spill[j];
        g.sync();
        spill = dst;
        spill_cap = (uint32_t)need;
        return true;
    }
    __device__ __forceinline__ void entry(uint32_t j, uint32_t& eq, uint32_t& len, uint64_t& off) const {  // uniform j
        if (j < G) {
            eq = g.shfl(my_eq, j);
            len = g.shfl(my_len, j);
            off = g.shfl(my_off, j);
        } else {
            uint4 e = spill[j - G];
            eq = e.x;
            len = e.y;
            off = (uint64_t)e.z | ((uint64_t)e.w << 32);
        }
    }
};

// nodes_to_eq_class, ref src/pseudoaligner.rs:323-356, on the distinct visited classes.
// The result is the ascending intersection (intersect keeps v1's order, :406); the sort by
// length (:331-334) only picks the smallest class as v1.  Lane l tests member c0+l of the
// smallest class against every other class by binary search (the reference's own search,
// :404).  out == nullptr counts, otherwise the survivors are written to out[].
template <int KW, bool EV, int G>
__device__ __forceinline__ uint32_t intersect_pass(WarpCtx<KW, EV, G>& w, uint32_t s_eq, uint32_t s_len, uint64_t s_off,
                                                   uint32_t* out) {
    const DevIndex& ix = w.ix;
    uint32_t count = 0;
    for (uint32_t c0 = 0; c0 < s_len; c0 += G) {
        bool alive = c0 + w.lane < s_len;
        uint32_t mem = alive ? __ldg(ix.eq_mem + s_off + c0 + w.lane) : 0;
        for (uint32_t j = 0; j < w.n_list; j++) {
            uint32_t e, l;
            uint64_t o;
            w.entry(j, e, l, o);
            if (e == s_eq) continue;
            if (alive) alive = contains_sorted(ix.eq_mem + o, (uint64_t)l, mem);
            if (!w.g.any(alive)) break;
        }
        unsigned b = w.g.ballot(alive);
        if (out && alive) out[count + __popc(b & ((1u << w.lane) - 1))] = mem;
        count += __popc(b);
    }
    return count;
}

// index construction: every k-mer sets its eight bits in the seed-scan filter (psa_core.cuh bloom_*)
template <int KW>
__global__ void k_bloom_set(const uint64_t* key_lo, const uint64_t* key_hi, uint64_t n, uint32_t* bloom, uint64_t n_blocks) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    Kmer<KW> key;
    key.lo = key_lo[i];
    if constexpr (KW == 2) key.hi = key_hi[i];
    const KeyHash kh = make_hash(KmerOps<KW>::fold(key));
    uint32_t* blk = bloom + 8 * bloom_block_of(kh, n_blocks);
    const uint64_t x = bloom_bits_of(kh);
#pragma unroll
    for (int j = 0; j < 8; j++) atomicOr(blk + j, 1u << ((x >> (5 * j)) & 31));
}

// the length the hand-over list has now (k_map's first launch maps exactly these entries while the list keeps growing)
__global__ void k_list_snapshot(const unsigned long long* list_count, unsigned long long* out) {
    if (threadIdx.x == 0) *out = *list_count;
}

template <int KW, bool EV, int G>
__global__ void __launch_bounds__(256) k_map(const __grid_constant__ DevIndex ix, const __grid_constant__ MapParams p) {
    const uint64_t gid = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) / G;
    const uint64_t ngroups = (gridDim.x * (uint64_t)blockDim.x) / G;
    LaneEvents tot{};
    uint64_t ev_reads = 0, ev_bases = 0, ev_out = 0, ev_aligned = 0;

    const uint64_t n_todo = p.list ? (uint64_t)(p.list_end ? *p.list_end : *p.list_count) : p.reads.n;
    const uint64_t first = p.list && p.list_first ? (uint64_t)*p.list_first : 0;
    // The handed-over reads differ widely in cost (tens to hundreds of dependent loads): groups claim them one
    // at a time from a counter instead of striding over the list, so that no group is left with several slow ones.
    const bool dynamic = p.list != nullptr && p.work_cursor != nullptr;
    const Grp<G> wg;
    auto claim = [&]() -> uint64_t {
        unsigned long long at = 0;
        if (wg.lane == 0) at = atomicAdd(p.work_cursor, 1ULL);
        return first + wg.shfl(at, 0);
    };
    for (uint64_t it = dynamic ? claim() : first + gid; it < n_todo; it = dynamic ? claim() : it + ngroups) {
        const uint64_t r = p.list ? (uint64_t)p.list[it] : it;
        const uint64_t wo = p.reads.woff ? p.reads.woff[r] : r * p.reads.wstride;
        const uint32_t L = p.reads.len ? p.reads.len[r] : p.reads.fixed_len;
        WarpCtx<KW, EV, G> w(ix, p.reads.words + wo, L, p, gid);
        const uint32_t lane = w.lane;
        uint32_t coverage = 0;
        bool some = map_read_nodes<uint32_t>(w, ix.k, L, p.allowed_mismatches, coverage);

        HitRec h;
        h.coverage = 0; h.n_tx = 0; h.tx_off = 0; h.eq_id = kNone; h.flags = 0;
        uint64_t count_slot = ix.n_eq + 1;  // None
        if (some) {
            // smallest class first (ref :331-334); ties broken by id so every lane agrees
            uint64_t key = lane < w.n_list ? (((uint64_t)w.my_len << 32) | w.my_eq) : ~0ULL;
            uint64_t koff = w.my_off;
            for (uint32_t j = G + lane; j < w.n_list; j += G) {
                uint4 e = w.spill[j - G];
                uint64_t kk = ((uint64_t)e.y << 32) | e.x;
                if (kk < key) { key = kk; koff = (uint64_t)e.z | ((uint64_t)e.w << 32); }
            }
#pragma unroll
            for (int d = G / 2; d; d >>= 1) {
                uint64_t o = w.g.shfl_xor(key, d);
                uint64_t oo = w.g.shfl_xor(koff, d);
                if (o < key) { key = o; koff = oo; }
            }
            const uint32_t s_len = (uint32_t)(key >> 32), s_eq = (uint32_t)key;
            const uint64_t s_off = koff;
            uint32_t count, eq_id;
            WinAcc acc;
            acc.base = 0; acc.map = Win{0, 0, 0}; acc.have = false;
            if (w.n_list == 1) {
                count = s_len;
                eq_id = s_eq;
                if (EV && lane == 0) w.ev.members += s_len;
            } else {
                // windows of this lane's entries, then a butterfly over the group (psa_core.cuh "Class windows")
                uint32_t n_wide = 0;
                if (lane < w.n_list) {
                    ClassWin c = load_class_win(ix.class_win + w.my_eq);
                    if (c.len == kWinWide) n_wide++;
                    else winacc_and(acc, c);
                }
                for (uint32_t j = G + lane; j < w.n_list; j += G) {
                    ClassWin c = load_class_win(ix.class_win + w.spill[j - G].x);
                    if (c.len == kWinWide) n_wide++;
                    else winacc_and(acc, c);
                }
#pragma unroll
                for (int d = G / 2; d; d >>= 1) {
                    WinAcc o;
                    o.base = w.g.shfl_xor(acc.base, d);
                    o.map.w0 = w.g.shfl_xor(acc.map.w0, d);
                    o.map.w1 = w.g.shfl_xor(acc.map.w1, d);
                    o.map.w2 = w.g.shfl_xor(acc.map.w2, d);
                    o.have = w.g.shfl_xor((int)acc.have, d) != 0;
                    winacc_merge(acc, o);
                    n_wide += w.g.shfl_xor(n_wide, d);
                }
                if (acc.have) {
                    if (n_wide) {
                        // wide classes filter the surviving candidates: every lane looks its own entries up
                        // (one range lookup per class, psa_core.cuh win_of_list_range), a butterfly ANDs the masks
                        Win keep{~0ULL, ~0ULL, ~0ULL};
                        if (lane < w.n_list && __ldg(&ix.class_win[w.my_eq].len) == kWinWide)
                            keep = win_of_list_range(ix.eq_mem + w.my_off, w.my_len, acc.base);
                        for (uint32_t j = G + lane; j < w.n_list; j += G) {
                            const uint4 e = w.spill[j - G];
                            if (__ldg(&ix.class_win[e.x].len) != kWinWide) continue;
                            keep = win_and(keep, win_of_list_range(ix.eq_mem + ((uint64_t)e.z | ((uint64_t)e.w << 32)), e.y, acc.base));
                        }
#pragma unroll
                        for (int d = G / 2; d; d >>= 1) {
                            keep.w0 &= w.g.shfl_xor(keep.w0, d);
                            keep.w1 &= w.g.shfl_xor(keep.w1, d);
                            keep.w2 &= w.g.shfl_xor(keep.w2, d);
                        }
                        acc.map = win_and(acc.map, keep);
                    }
                    count = win_popc(acc.map);
                } else {
                    count = intersect_pass(w, s_eq, s_len, s_off, nullptr);  // every class is wide: the list scheme
                }
                // the result equals a visited class iff that class has `count` members
                uint32_t cand = (lane < w.n_list && w.my_len == count) ? w.my_eq : kNone;
                for (uint32_t j = G + lane; j < w.n_list; j += G) {
                    uint4 e = w.spill[j - G];
                    if (e.y == count && e.x < cand) cand = e.x;
                }
#pragma unroll
                for (int d = G / 2; d; d >>= 1) cand = min(cand, w.g.shfl_xor(cand, d));
                eq_id = cand;
                if (EV) {
                    for (uint32_t j = 0; j < w.n_list; j++) {
                        uint32_t e, l;
                        uint64_t o;
                        w.entry(j, e, l, o);
                        if (lane == 0) w.ev.members += l;
                    }
                }
            }
            h.coverage = coverage;
            h.n_tx = count;
            h.eq_id = eq_id;
            h.flags = kFlagAligned | ((coverage >= kCoverageThreshold && count == 0) ? kFlagMapped : 0u);  // ref :455 (sic)
            if (eq_id != kNone) {
                count_slot = eq_id;  // (members: k_expand reads them from the index through eq_id)
            } else {
                count_slot = ix.n_eq;
                if (p.novel) {
                    unsigned long long base = 0;
                    if (lane == 0) {
                        base = atomicAdd(p.novel_cursor, (unsigned long long)count);
                        if (p.novel_list) p.novel_list[atomicAdd(p.novel_list_count, 1ULL)] = (uint32_t)r;
                    }
                    base = w.g.shfl(base, 0);
                    if (base + count > p.novel_cap) {
                        if (lane == 0) atomicOr(p.status, 1u);
                    } else if (acc.have) {
                        if (lane == 0) win_write(acc, p.novel + base);
                    } else {
                        intersect_pass(w, s_eq, s_len, s_off, p.novel + base);
                    }
                    h.tx_off = base;
                }
            }
            if (w.spill_overflow && lane == 0) atomicOr(p.status, 2u);
        }
        if (lane == 0) {
            p.hits[r] = h;
            if (p.counts) atomicAdd(p.counts + count_slot, 1ULL);
        }
        if (EV && lane == 0) {
            tot.lookups += w.ev.lookups; tot.levels += w.ev.levels; tot.hits += w.ev.hits; tot.verifs += w.ev.verifs;
            tot.visits += w.ev.visits; tot.bases += w.ev.bases; tot.jumps += w.ev.jumps; tot.members += w.ev.members;
            ev_reads++; ev_bases += L; ev_out += h.n_tx; ev_aligned += some;
        }
    }
    if (EV && p.events && (threadIdx.x & (G - 1)) == 0) {
        unsigned long long v[12] = {ev_reads, ev_bases, tot.lookups, tot.levels, tot.hits, tot.verifs,
                                    tot.visits, tot.bases, tot.jumps, tot.members, ev_out, ev_aligned};
#pragma unroll
        for (int i = 0; i < 12; i++)
            if (v[i]) atomicAdd(p.events + 12 + i, v[i]);
    }
}

// where the thread-per-read kernel puts a finished read's result
struct DevSink {
    const MapParams& p;
    __device__ __forceinline__ void result(uint32_t r, const HitRec& h, uint64_t count_slot) {
        // psa_hit is 24 bytes at an 8-byte aligned address: three 8-byte stores
        uint64_t* out = reinterpret_cast<uint64_t*>(p.hits + r);
        out[0] = (uint64_t)h.coverage | ((uint64_t)h.n_tx << 32);
        out[1] = h.tx_off;
        out[2] = (uint64_t)h.eq_id | ((uint64_t)h.flags << 32);
        if (p.counts) atomicAdd(p.counts + count_slot, 1ULL);
    }
    __device__ __forceinline__ uint32_t* novel(uint32_t r, uint32_t count, uint64_t& off) {
        unsigned long long base = atomicAdd(p.novel_cursor, (unsigned long long)count);
        if (p.novel_list) p.novel_list[atomicAdd(p.novel_list_count, 1ULL)] = r;
        off = base;
        return base + count <= p.novel_cap ? p.novel + base : nullptr;
    }
    __device__ __forceinline__ void novel_overflow() { atomicOr(p.status, 1u); }
};

// ---------------------------------------------------------------------------------------------
// the thread-per-read kernel in its blocking form (psa_thread.cuh): read r = global thread id, one
// call of map_read per thread, 64-thread CTAs at 20 per SM.  HINT = true: the reads of p.seeded,
// persistent warps striding over that list.  TILE (first pass, fixed word stride): the packed words of
// the CTA's 64 reads arrive in shared memory as one bulk asynchronous copy (cp.async.bulk + mbarrier:
// TMA) and every thread maps its read from there.
// ---------------------------------------------------------------------------------------------
#ifndef PSA_THREAD_BLOCK
#define PSA_THREAD_BLOCK 64
#endif
constexpr int kThreadBlock = PSA_THREAD_BLOCK;
#ifndef PSA_THREAD_MIN_BLOCKS
#define PSA_THREAD_MIN_BLOCKS 16
#endif
template <int KW, bool EV, bool HINT, bool TILE = false>
__global__ void __launch_bounds__(kThreadBlock, PSA_THREAD_MIN_BLOCKS) k_map_thread(const __grid_constant__ DevIndex ix,
                                                                                     const __grid_constant__ MapParams p) {
    const unsigned lane = threadIdx.x & 31;
    extern __shared__ __align__(128) uint8_t smem[];
    const uint64_t* my_words = nullptr;
    if (TILE) {
        // layout: [mbarrier (16 B slot) | packed words of the CTA's reads]
        uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
        uint64_t* pw = reinterpret_cast<uint64_t*>(smem + 16);
        const uint64_t r0 = blockIdx.x * (uint64_t)kThreadBlock;
        const uint32_t nr = (uint32_t)min((uint64_t)kThreadBlock, p.reads.n - r0);
        const uint32_t nw = (uint32_t)p.reads.wstride;
        const uint64_t* src = p.reads.words + r0 * nw;
        const uint32_t bytes = nr * nw * 8;
        if ((bytes & 15) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
            if (threadIdx.x == 0) {
                mbar_init(bar, 1);
                mbar_expect_tx(bar, bytes);
                bulk_g2s(pw, src, bytes, bar);
            }
            __syncthreads();  // the barrier is initialised before anyone polls it
            mbar_wait(bar, 0);
        } else {  // odd-sized last tile: plain loads
            for (uint32_t i = threadIdx.x; i < nr * nw; i += kThreadBlock) pw[i] = src[i];
            __syncthreads();
        }
        my_words = pw + threadIdx.x * nw;
    }
    DevSink sink{p};
    const uint64_t n_todo = HINT ? (uint64_t)*p.seeded_count : p.reads.n;
    // Second pass: persistent warps that CLAIM their next 32 entries from a counter (the entries differ widely in cost;
    // striding over the list left the slowest warp several entries behind).  Warp-uniform trip count either way: the
    // hand-over below uses full-warp votes.
    auto next_base = [&](uint64_t) -> uint64_t {
        unsigned long long at = 0;
        if (lane == 0) at = atomicAdd(p.hint_cursor, 32ULL);
        return __shfl_sync(kFull, at, 0);
    };
    for (uint64_t base = HINT ? next_base(0) : blockIdx.x * (uint64_t)blockDim.x + (threadIdx.x & ~31u); base < n_todo;
         base = HINT ? next_base(base) : ~0ULL >> 1) {
        const uint64_t it = base + lane;
        const bool live = it < n_todo;
        bool defer = false;
        uint32_t why = 0;
        uint64_t r = it;
        ThreadEvents ev{};
        uint4 sev = make_uint4(0, 0, 0, 0);
        uint32_t L = 0, n_tx = 0, aligned = 0;
        if (live) {
            uint32_t hint[3];
            bool hinted = false;
            if (HINT) {
                const uint4 e = p.seeded[it];
                r = e.x;
                hint[0] = e.y; hint[1] = e.z; hint[2] = e.w;
                hinted = e.y != kNone;   // (kNone: a read the first pass gave up at a re-seed search -- no answer to reuse)
                if (EV) sev = p.seeded_ev[it];
            }
            const uint64_t wo = p.reads.woff ? p.reads.woff[r] : r * p.reads.wstride;
            L = p.reads.len ? p.reads.len[r] : p.reads.fixed_len;
            ThreadResult res = TILE ? map_read_thread<KW, EV>(ix, PLoad{my_words}, (uint32_t)r, L, p.allowed_mismatches, p.max_probes,
                                                              p.reseed_probes, p.max_small, sink, p.novel != nullptr, EV ? &ev : nullptr,
                                                              hinted ? hint : nullptr)
                                    : map_read_thread<KW, EV>(ix, PLoad{p.reads.words + wo}, (uint32_t)r, L, p.allowed_mismatches, p.max_probes,
                                                              p.reseed_probes, p.max_small, sink, p.novel != nullptr, EV ? &ev : nullptr,
                                                              hinted ? hint : nullptr);
            defer = res.deferred;
            why = res.why;
            if (EV && defer && p.events) atomicAdd(p.events + 36 + res.why, 1ULL);
            if (!defer) {
                sink.result((uint32_t)r, res.hit, res.count_slot);
                if (res.novel_overflow) sink.novel_overflow();
                n_tx = res.hit.n_tx;
                aligned = res.hit.flags & kFlagAligned;
            } else if (EV && HINT && p.events) {
                // k_map redoes this read from scratch and counts its first search again
                atomicAdd(p.events + 24 + 2, 0ULL - sev.x); atomicAdd(p.events + 24 + 3, 0ULL - sev.y);
                atomicAdd(p.events + 24 + 4, 0ULL - sev.z); atomicAdd(p.events + 24 + 5, 0ULL - sev.w);
            }
        }
        // hand the given-up reads over (one atomic per warp and list): a too long FIRST seed search
        // goes to k_seed_scan, everything else to the cooperative kernel
        const bool to_scan = defer && !HINT && why == 0 && p.scan_list != nullptr;
        const unsigned bs = __ballot_sync(kFull, to_scan);
        if (bs) {
            unsigned long long at = 0;
            if (lane == (unsigned)(__ffs(bs) - 1)) at = atomicAdd(p.scan_count, (unsigned long long)__popc(bs));
            at = __shfl_sync(kFull, at, __ffs(bs) - 1);
            if (to_scan) p.scan_list[at + __popc(bs & ((1u << lane) - 1))] = (uint32_t)r;
        }
        // a re-seed search beyond the first pass's budget: the second pass redoes the read with the long one
        const bool to_retry = defer && !HINT && why == 1 && p.seeded != nullptr;
        const unsigned br = __ballot_sync(kFull, to_retry);
        if (br) {
            unsigned long long at = 0;
            if (lane == (unsigned)(__ffs(br) - 1)) at = atomicAdd(p.seeded_count, (unsigned long long)__popc(br));
            at = __shfl_sync(kFull, at, __ffs(br) - 1);
            if (to_retry) {
                const unsigned long long slot = at + __popc(br & ((1u << lane) - 1));
                p.seeded[slot] = make_uint4((uint32_t)r, kNone, 0u, 0u);
                if (EV) p.seeded_ev[slot] = make_uint4(0u, 0u, 0u, 0u);
            }
        }
        const bool to_coop = defer && !to_scan && !to_retry;
        const unsigned bc = __ballot_sync(kFull, to_coop);
        if (bc) {
            unsigned long long at = 0;
            if (lane == (unsigned)(__ffs(bc) - 1)) at = atomicAdd(p.list_count, (unsigned long long)__popc(bc));
            at = __shfl_sync(kFull, at, __ffs(bc) - 1);
            if (to_coop) p.list[at + __popc(bc & ((1u << lane) - 1))] = (uint32_t)r;
        }
        if (EV && p.events) {
            const bool cnt = live && !defer;
            unsigned long long v[12] = {cnt ? 1ull : 0ull, cnt ? L : 0ull, ev.lookups, ev.levels, ev.hits, ev.verifs,
                                        ev.visits, ev.bases, ev.jumps, ev.members, n_tx, aligned};
#pragma unroll
            for (int i = 0; i < 12; i++) {
                unsigned long long x = cnt ? v[i] : 0ull;
#pragma unroll
                for (int d = 16; d; d >>= 1) x += __shfl_xor_sync(kFull, x, d);
                if (lane == 0 && x) atomicAdd(p.events + i, x);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// k_seed_scan: the first seed search (find_kmer_match from position 0, ref :91-114) of the reads
// whose search was too long for one thread, G lanes probing G stride-3 positions at once.  A read
// with no seed at all is finished here (map_read = None); a seeded one goes back to the
// thread-per-read kernel with the answer.
// ---------------------------------------------------------------------------------------------
template <int KW, bool EV, int G>
__global__ void __launch_bounds__(256) k_seed_scan(const __grid_constant__ DevIndex ix, const __grid_constant__ MapParams p) {
    const uint64_t gid = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) / G;
    const uint64_t ngroups = (gridDim.x * (uint64_t)blockDim.x) / G;
    const uint64_t n_todo = (uint64_t)*p.scan_count;
    for (uint64_t it = gid; it < n_todo; it += ngroups) {
        const uint64_t r = p.scan_list[it];
        const uint64_t wo = p.reads.woff ? p.reads.woff[r] : r * p.reads.wstride;
        const uint32_t L = p.reads.len ? p.reads.len[r] : p.reads.fixed_len;  // >= k: shorter reads never search
        WarpCtx<KW, EV, G> w(ix, p.reads.words + wo, L, p, gid);
        w.scan_mode = true;
        w.stage_words();
        // a read is on this list because its own thread probed positions 0, 3, ..., 3 (max_probes - 1) and missed
        // them all: start behind them (when counting events the whole search is redone: the counters are
        // sequential-equivalent and the thread kernel did not count a search it gave up)
        w.scan_skip = EV ? 0u : kSeedStride * p.max_probes;
        uint32_t kmer_pos = 0;
        uint32_t node = 0, off = 0;
        const bool found = w.find_seed(kmer_pos, L - ix.k, node, off);
        if (w.lane == 0) {
            if (found) {
                const unsigned long long at = atomicAdd(p.seeded_count, 1ULL);
                p.seeded[at] = make_uint4((uint32_t)r, (uint32_t)kmer_pos, node, off);
                if (EV) p.seeded_ev[at] = make_uint4(w.ev.lookups, w.ev.levels, w.ev.hits, w.ev.verifs);
            } else {
                uint64_t* out = reinterpret_cast<uint64_t*>(p.hits + r);
                out[0] = 0;
                out[1] = 0;
                out[2] = (uint64_t)kNone;  // eq_id = none, flags = 0
                if (p.counts) atomicAdd(p.counts + ix.n_eq + 1, 1ULL);
            }
            if (EV && p.events) {
                unsigned long long* e = p.events + 24;
                if (!found) { atomicAdd(e + 0, 1ULL); atomicAdd(e + 1, (unsigned long long)L); }
                atomicAdd(e + 2, (unsigned long long)w.ev.lookups); atomicAdd(e + 3, (unsigned long long)w.ev.levels);
                atomicAdd(e + 4, (unsigned long long)w.ev.hits); atomicAdd(e + 5, (unsigned long long)w.ev.verifs);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Self-test entry: the three device routines that compute nodes_to_eq_class / intersect (ref
// src/pseudoaligner.rs:323-356, :389-418) applied to two ascending lists in isolation -- one thread's
// list scheme (thread_intersect_lists), the lane group's (intersect_pass) and the class windows.
// ix holds the two lists as classes 0 and 1.  out: 3 x cap entries, n_out[3] (kNone: not applicable).
// ---------------------------------------------------------------------------------------------
__global__ void k_selftest_intersect(DevIndex ix, MapParams p, uint32_t* out, uint32_t cap, uint32_t* n_out) {
    const uint32_t n0 = (uint32_t)(ix.eq_off[1] - ix.eq_off[0]), n1 = (uint32_t)(ix.eq_off[2] - ix.eq_off[1]);
    const int s = n1 < n0 ? 1 : 0;  // smallest class first, ties by id (ref :331-334)
    if (threadIdx.x == 0) {
        ClassAcc w;
        w.init();
        w.wide_eq[0] = 0; w.wide_eq[1] = 1; w.n_wide = 2;
        const uint32_t c = thread_intersect_lists(ix, w, s, (uint32_t*)nullptr);
        n_out[0] = c;
        if (c <= cap) thread_intersect_lists(ix, w, s, out);
        // windows
        const ClassWin c0 = make_class_win(ix.eq_mem + ix.eq_off[0], n0), c1 = make_class_win(ix.eq_mem + ix.eq_off[1], n1);
        if (c0.len == kWinWide || c1.len == kWinWide) {
            n_out[2] = kNone;
        } else {
            WinAcc a;
            a.base = 0; a.map = Win{0, 0, 0}; a.have = false;
            winacc_and(a, c0);
            winacc_and(a, c1);
            n_out[2] = win_popc(a.map);
            if (n_out[2] <= cap) win_write(a, out + 2 * (uint64_t)cap);
        }
    }
    if (threadIdx.x < 8) {  // one group of the cooperative kernel
        WarpCtx<1, false, 8> w(ix, nullptr, 0, p, 0);
        w.n_list = 2;
        if (w.lane < 2) {
            w.my_eq = w.lane;
            w.my_len = w.lane == 0 ? n0 : n1;
            w.my_off = ix.eq_off[w.lane];
        }
        const uint32_t s_len = s ? n1 : n0;
        const uint32_t c = intersect_pass(w, (uint32_t)s, s_len, ix.eq_off[s], (uint32_t*)nullptr);
        if (w.lane == 0) n_out[1] = c;
        if (c <= cap) intersect_pass(w, (uint32_t)s, s_len, ix.eq_off[s], out + cap);
    }
}

// ---------------------------------------------------------------------------------------------
// Measurement aid: independent random gathers of `bytes`-sized aligned chunks from a table, the
// access pattern of the index lookups without their dependencies.  Its rate is the practical
// ceiling ("random-sector roofline") the map kernels are compared with in DESIGN.md.
// ---------------------------------------------------------------------------------------------
// BYTES = 32/64/128: every thread gathers its own chunk (32-byte LDG.E.256 pieces, as the index
// lookups do).  BYTES = 0: every warp gathers one 128-byte line, 4 bytes per lane (the rate at
// which HBM serves random lines, without divergent addresses inside a load instruction).
template <int BYTES>
__global__ void k_gather_probe(const uint64_t* table, uint64_t n_chunks, uint32_t iters, uint64_t seed, uint64_t* sink) {
    const uint64_t tid = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    uint64_t acc = 0;
    uint64_t x = mix64(seed + (BYTES ? tid : tid >> 5) * 0x9E3779B97F4A7C15ULL);
    for (uint32_t i = 0; i < iters; i++) {
        x = x * 6364136223846793005ULL + 1442695040888963407ULL;
        const uint64_t c = mulhi64(x, n_chunks);
        if (BYTES == 0) {
            acc += __ldg(reinterpret_cast<const uint32_t*>(table + c * 16) + (threadIdx.x & 31));
        } else {
#pragma unroll
            for (int j = 0; j < BYTES / 32; j++) {
                uint64_t a, b2, c2, d;
                asm volatile("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b2), "=l"(c2), "=l"(d)
                             : "l"(table + c * (BYTES / 8) + 4 * j));
                acc += a ^ b2 ^ c2 ^ d;
            }
        }
    }
    if (acc == seed) sink[0] = acc;  // keeps the loads alive
}

// ---------------------------------------------------------------------------------------------
// result expansion: hits[i].tx_off (a source offset) -> members copied to tx_buf in read order
// ---------------------------------------------------------------------------------------------
struct TxLenN {  // iterator adaptor for the scan of n_tx; item n is a zero so that out[n] = total
    const HitRec* h;
    uint64_t n;
    int novel_only;  // compact results: only sets that are no index class travel as members
    __host__ __device__ uint64_t operator()(uint64_t i) const {
        if (i >= n) return 0;
        return novel_only && h[i].eq_id != kNone ? 0 : (uint64_t)h[i].n_tx;
    }
};
struct CastU64 {
    __host__ __device__ uint64_t operator()(uint32_t x) const { return x; }
};

// rel_off[i] = offset of read i's members inside this batch's tx_buf (an exclusive scan of n_tx); running[0] =
// members emitted by the batches before this one (so that tx_off is global across a chunked call).
// A warp takes 32 consecutive reads, whose members are one contiguous range of tx_buf; lane t of every round writes element t of that range -- it finds the read
// the element belongs to by a binary search over the lanes' start offsets (shuffles) and fetches the source
// pointer from that lane.  Stores are fully coalesced and no lane idles behind a read with a long class.
__global__ void __launch_bounds__(256) k_expand_balanced(HitRec* hits, uint64_t n, const uint64_t* rel_off, const uint64_t* running,
                                                         const uint64_t* eq_off, const uint32_t* eq_mem, const uint32_t* novel,
                                                         uint32_t* tx_buf, uint64_t tx_cap, int novel_only) {
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t base = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) & ~31ull;
    if (base >= n) return;  // warp-uniform
    const uint64_t i = base + lane;
    const bool live = i < n;
    uint32_t cnt = 0;
    uint64_t rel = 0;
    const uint32_t* src = nullptr;
    bool fits = false;
    if (live) {
        const HitRec h = hits[i];
        rel = rel_off[i];
        cnt = novel_only && h.eq_id != kNone ? 0 : h.n_tx;
        fits = tx_buf != nullptr && rel + cnt <= tx_cap;
        // members of an index class come from the index (through eq_id), any other set from the novel-set buffer
        src = h.eq_id != kNone ? eq_mem + __ldg(eq_off + h.eq_id) : novel + h.tx_off;
        hits[i].tx_off = running[0] + rel;
    }
    const uint64_t rel0 = __shfl_sync(kFull, rel, 0);
    const uint32_t start = live ? (uint32_t)(rel - rel0) : 0xFFFFFFFFu;  // ascending over the lanes; dead lanes last
    const uint32_t total = __reduce_max_sync(kFull, live ? start + cnt : 0u);
    const unsigned fitmask = __ballot_sync(kFull, fits);
    if (!fitmask) return;
    for (uint32_t t0 = 0; t0 < total; t0 += 32) {
        const uint32_t t = t0 + lane;
        uint32_t q = 0;  // last lane whose range starts at or before element t (empty ranges are passed over)
#pragma unroll
        for (int s = 16; s; s >>= 1) {
            const uint32_t sv = __shfl_sync(kFull, start, (q + s) & 31);
            if (sv <= t) q += s;
        }
        const uint64_t sp = __shfl_sync(kFull, (uint64_t)(uintptr_t)src, q);
        const uint32_t ss = __shfl_sync(kFull, start, q);
        if (t < total && ((fitmask >> q) & 1u)) tx_buf[rel0 + t] = __ldg(reinterpret_cast<const uint32_t*>((uintptr_t)sp) + (t - ss));
    }
}
// ---------------------------------------------------------------------------------------------
// Novel sets: an eq_class that is no index class (the reference returns the set itself, ref
// src/pseudoaligner.rs:323-356, :381-384) is counted per distinct set in a table that lives with the
// mapper.  After the map kernels of a batch, one thread per read whose set is no index class hashes
// the members, finds or claims the set's entry (open addressing, 64-bit compare-and-swap on the hash)
// and adds one to its count; the claimer copies the members into the table's pool.  k_novel_verify
// then compares every such read's members with its entry's: two different sets behind one 64-bit
// hash would be reported (status bit 8), never merged silently.
// ---------------------------------------------------------------------------------------------
struct NovelEntry {
    unsigned long long key;    // hash of the set (0 = empty slot)
    unsigned long long count;  // reads whose eq_class is this set
    unsigned long long off;    // members at pool[off ..) -- written by the claimer
    uint32_t len;              // 0xFFFFFFFF: claimed, members not stored (pool full)
    uint32_t pad;
};
static_assert(sizeof(NovelEntry) == 32, "NovelEntry");
struct NovelTable {
    NovelEntry* tab;
    uint64_t cap;              // entries, a power of two
    uint32_t* pool;
    uint64_t pool_cap;
    unsigned long long* cursors;  // [0] members used in pool, [1] entries claimed
};
constexpr uint32_t kStatusNovelClash = 8u, kStatusNovelFull = 16u;
__device__ __forceinline__ unsigned long long novel_hash(const uint32_t* m, uint32_t n) {
    unsigned long long h = mix64(0x9E3779B97F4A7C15ULL + n);
    for (uint32_t i = 0; i < n; i++) h = mix64(h ^ ((unsigned long long)m[i] + 0xD6E8FEB86659FD93ULL));
    return h ? h : 1ULL;
}
// step 1, over the batch's list of novel reads (written by the map kernels): find or claim the set's entry
__global__ void k_novel_claim(const uint32_t* list, const unsigned long long* list_count, const HitRec* hits, const uint32_t* novel,
                              NovelTable t, uint32_t* slot_out, uint32_t* status) {
    const uint64_t n = *list_count;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += gridDim.x * (uint64_t)blockDim.x) {
        const HitRec h = hits[list[i]];
        const uint32_t* m = novel + h.tx_off;
        const unsigned long long key = novel_hash(m, h.n_tx);
        uint32_t found = kNone;
        for (uint64_t probe = 0, slot = key & (t.cap - 1); probe < t.cap; probe++, slot = (slot + 1) & (t.cap - 1)) {
            unsigned long long cur = t.tab[slot].key;
            if (cur == 0) cur = atomicCAS(&t.tab[slot].key, 0ULL, key);
            if (cur == 0) {  // claimed: publish the members
                const unsigned long long off = atomicAdd(t.cursors, (unsigned long long)h.n_tx);
                const unsigned long long ne = atomicAdd(t.cursors + 1, 1ULL);
                if (off + h.n_tx > t.pool_cap || 2 * (ne + 1) > t.cap) {
                    atomicOr(status, kStatusNovelFull);   // the host grows the table and redoes the batch
                    t.tab[slot].len = 0xFFFFFFFFu;
                } else {
                    for (uint32_t j = 0; j < h.n_tx; j++) t.pool[off + j] = m[j];
                    t.tab[slot].off = off;
                    t.tab[slot].len = h.n_tx;
                }
                found = (uint32_t)slot;
                break;
            }
            if (cur == key) {
                found = (uint32_t)slot;
                break;
            }
        }
        if (found == kNone) atomicOr(status, kStatusNovelFull);
        slot_out[i] = found;
    }
}
// step 2: every listed read's members against its entry's -- two sets behind one hash are reported, never merged
__global__ void k_novel_verify(const uint32_t* list, const unsigned long long* list_count, const HitRec* hits, const uint32_t* novel,
                               NovelTable t, const uint32_t* slot_in, uint32_t* status) {
    const uint64_t n = *list_count;
    if (*status & kStatusNovelFull) return;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += gridDim.x * (uint64_t)blockDim.x) {
        const HitRec h = hits[list[i]];
        const uint32_t* m = novel + h.tx_off;
        const NovelEntry e = t.tab[slot_in[i]];
        bool same = e.len == h.n_tx;
        for (uint32_t j = 0; same && j < h.n_tx; j++) same = t.pool[e.off + j] == m[j];
        if (!same) atomicOr(status, kStatusNovelClash);
    }
}
// step 3, once the batch is known to stand (no overflow of any buffer, tx_buf large enough): count
__global__ void k_novel_add(const unsigned long long* list_count, NovelTable t, const uint32_t* slot_in, const uint32_t* status,
                            const uint64_t* batch_total, uint64_t tx_cap, int has_tx) {
    const uint64_t n = *list_count;
    if ((*status & 31u) || (has_tx && *batch_total > tx_cap)) return;
    // popular sets (the empty one above all) would serialise one atomic per read on one address: the lanes of a warp
    // that count the same entry add once
    const uint64_t n_round = (n + 31) & ~31ULL;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n_round; i += gridDim.x * (uint64_t)blockDim.x) {
        const uint32_t slot = i < n ? slot_in[i] : kNone;
        const unsigned same = __match_any_sync(kFull, slot);
        if (slot != kNone && (threadIdx.x & 31) == (unsigned)(__ffs(same) - 1))
            atomicAdd(&t.tab[slot].count, (unsigned long long)__popc(same));
    }
}
// growing the table: every used entry of the old one moves to the new one
__global__ void k_novel_rehash(const NovelEntry* old_tab, uint64_t old_cap, NovelEntry* tab, uint64_t cap) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= old_cap) return;
    const NovelEntry e = old_tab[i];
    if (e.key == 0 || e.len == 0xFFFFFFFFu) return;
    for (uint64_t slot = e.key & (cap - 1);; slot = (slot + 1) & (cap - 1)) {
        if (atomicCAS(&tab[slot].key, 0ULL, e.key) == 0ULL) {
            tab[slot].count = e.count; tab[slot].off = e.off; tab[slot].len = e.len;
            return;
        }
    }
}

// Verification aid: order-independent checksum of a result batch -- the sum over reads of a hash chain over
// (global read index, coverage, flags, eq_id, members in order).  oracle/psa_oracle.c restates it for host buffers.
__global__ void k_result_checksum(const HitRec* hits, const uint32_t* tx, uint64_t n, uint64_t first_index, uint64_t tx_base,
                                  unsigned long long* out) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    uint64_t h = 0;
    if (i < n) {
        const HitRec r = hits[i];
        h = mix64(first_index + i + 0x9E3779B97F4A7C15ULL);
        h = mix64(h ^ (((uint64_t)r.coverage << 32) | r.flags));
        h = mix64(h ^ (((uint64_t)r.n_tx << 32) | r.eq_id));
        for (uint32_t j = 0; j < r.n_tx; j++) h = mix64(h ^ ((uint64_t)tx[r.tx_off - tx_base + j] + 0xD6E8FEB86659FD93ULL));
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) h += __shfl_xor_sync(kFull, h, d);
    if ((threadIdx.x & 31) == 0 && h) atomicAdd(out, (unsigned long long)h);
}

// Compact results (psa.h PSA_RESULT_COMPACT): eight bytes per read.  eq_or_n = the index class of the read's
// eq_class | 0x80000000 + |eq_class| when it is no index class (its members travel in tx_buf) | 0xFFFFFFFF for None;
// cov_flags = coverage | PSA_FLAG_* << 28.
struct HitCompact {
    uint32_t eq_or_n, cov_flags;
};
constexpr uint32_t kCompactNovel = 0x80000000u, kCompactCovMask = (1u << 28) - 1;
__global__ void k_compact_hits(const HitRec* hits, uint64_t n, HitCompact* out, uint32_t* status) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const HitRec h = hits[i];
    HitCompact c;
    c.eq_or_n = !(h.flags & kFlagAligned) ? kNone : h.eq_id != kNone ? h.eq_id : (kCompactNovel | h.n_tx);
    c.cov_flags = (h.coverage & kCompactCovMask) | (h.flags << 28);
    if (h.coverage > kCompactCovMask || ((h.flags & kFlagAligned) && h.eq_id != kNone && h.eq_id >= kCompactNovel)) atomicOr(status, 32u);
    out[i] = c;
}

// after k_expand: advance the running total, publish {running, status} for the host.  sticky (may be
// nullptr): the OR of the status words of every batch since the host last cleared it -- several batches may
// be queued between two psa_mapper_sync calls and none of their overflows may go unnoticed.
__global__ void k_advance(uint64_t* running, const uint64_t* batch_total, uint64_t tx_cap, int has_tx,
                          const uint32_t* status, uint64_t* meta_out, uint32_t* sticky) {
    if (threadIdx.x || blockIdx.x) return;
    uint64_t t = *batch_total;
    uint32_t st = *status;
    if (has_tx && t > tx_cap) st |= 4u;
    if (sticky) {
        *sticky |= st;
        st = *sticky;
    }
    running[0] += t;
    meta_out[0] = running[0];
    meta_out[1] = st;
}

}  // namespace psa
