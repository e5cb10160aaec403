// psa_build.cu -- the coloured compacted de Bruijn graph built ON THE DEVICE:
// k-mer enumeration and sort, colour interning, unitig compaction -- what make_dbg / the debruijn crate's
// compression do on the host in the reference (ref src/build_index.rs:27-179, src/equiv_classes.rs:62-91).
// Semantics are the reference's, exactly as csrc/host/build_graph.cpp states them (the arrays produced here are
// bit-identical to that builder's, which the tests check):
//   - every k-mer of every transcript with len >= k, stranded (ref src/build_index.rs:127-151, src/config.rs:14);
//   - colour of a k-mer = ascending, de-duplicated list of the transcripts containing it, interned to a dense id in
//     order of first appearance over the sorted k-mers (CountFilterEqClass::summarize, src/equiv_classes.rs:62-91);
//   - exts of a k-mer = union over its occurrences of the neighbouring bases (src/equiv_classes.rs:73);
//   - unitig = maximal path whose every internal link is the unique right ext of its source, the unique left ext of
//     its target and joins equal colours (ScmapCompress, src/build_index.rs:171,178); a closed cycle is cut at its
//     smallest k-mer (cycles are resolved on the host: they are a handful at most).
// Algorithm: one (k-mer, transcript, exts) record per occurrence, radix-sorted by k-mer (stable, so the transcripts of a
// k-mer stay ascending); runs -> distinct k-mers with colour signatures; a second sort by signature groups equal colours,
// whose lists are then compared exactly; successor links by binary search in the sorted k-mers; unitigs walked from their
// heads.  All of it cub + a dozen small kernels; the host only sizes buffers and copies the result back.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>
#include <cub/iterator/counting_input_iterator.cuh>
#include <cub/iterator/transform_input_iterator.cuh>
#include <string>
#include <vector>

#include "../../include/psa.h"
#include "psa_core.cuh"

extern "C" int psa_internal_fail(int code, const char* msg);   // psa_api.cu: records the message for psa_last_error

namespace {
using psa::mix64;
constexpr uint32_t NONE32 = 0xFFFFFFFFu;

struct Buf {  // device allocation freed with its owner
    void* p = nullptr;
    size_t bytes = 0;
    cudaError_t alloc(size_t n) {
        bytes = n ? n : 1;
        return cudaMalloc(&p, bytes);
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
    }
    template <class T>
    T* as() const { return reinterpret_cast<T*>(p); }
    ~Buf() { release(); }
};
#define BCU(call)                                                                                   \
    do {                                                                                            \
        cudaError_t e_ = (call);                                                                    \
        if (e_ != cudaSuccess) {                                                                    \
            (void)cudaGetLastError();                                                               \
            return psa_internal_fail(PSA_ERR_CUDA, (std::string(#call) + ": " + cudaGetErrorString(e_)).c_str()); \
        }                                                                                           \
    } while (0)
inline unsigned blocks(uint64_t n, unsigned bs) { return (unsigned)((n + bs - 1) / bs); }

// ---- k-mers as one or two 64-bit words (KW = 2: k in 33..64, hi holds the first k - 32 bases)
template <int KW>
struct Km {
    uint64_t lo, hi;
    __host__ __device__ bool operator==(const Km& o) const { return lo == o.lo && (KW == 1 || hi == o.hi); }
    __host__ __device__ bool operator<(const Km& o) const { return KW == 2 && hi != o.hi ? hi < o.hi : lo < o.lo; }
};
template <int KW>
__device__ __forceinline__ Km<KW> km_push(Km<KW> x, uint32_t b, uint32_t k) {   // append base b, keep the last k bases
    Km<KW> r;
    if (KW == 1) {
        r.hi = 0;
        r.lo = ((x.lo << 2) | b) & (k == 32 ? ~0ULL : ((1ULL << (2 * k)) - 1));
    } else {
        r.hi = ((x.hi << 2) | (x.lo >> 62)) & (k == 64 ? ~0ULL : ((1ULL << (2 * (k - 32))) - 1));
        r.lo = (x.lo << 2) | b;
    }
    return r;
}

// one record per k-mer occurrence: occurrence i belongs to transcript tx = upper_bound(occ_off, i) - 1
template <int KW>
__global__ void k_occ_enumerate(const uint8_t* codes, const uint64_t* tx_off, const uint64_t* occ_off, uint32_t n_tx, uint32_t k, uint64_t n_occ,
                                uint64_t* key_lo, uint64_t* key_hi, uint64_t* payload, uint32_t* status) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= n_occ) return;
    uint32_t lo = 0, hi = n_tx;   // the last tx with occ_off[tx] <= i
    while (hi - lo > 1) {
        const uint32_t mid = lo + (hi - lo) / 2;
        if (occ_off[mid] <= i) lo = mid;
        else hi = mid;
    }
    const uint32_t tx = lo;
    const uint64_t p = i - occ_off[tx], len = tx_off[tx + 1] - tx_off[tx];
    const uint8_t* s = codes + tx_off[tx] + p;
    Km<KW> km{0, 0};
    uint32_t bad = 0;
    for (uint32_t t = 0; t < k; t++) {
        const uint32_t b = s[t];
        bad |= b > 3;
        km = km_push<KW>(km, b & 3u, k);
    }
    uint32_t e = 0;
    if (p > 0) { bad |= s[-1] > 3; e |= 1u << (4 + (s[-1] & 3)); }
    if (p + k < len) { bad |= s[k] > 3; e |= 1u << (s[k] & 3); }
    if (bad) atomicOr(status, 1u);
    key_lo[i] = km.lo;
    if (KW == 2) key_hi[i] = km.hi;
    payload[i] = ((uint64_t)tx << 8) | e;
}
__global__ void k_iota32(uint32_t* a, uint64_t n) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i < n) a[i] = (uint32_t)i;
}
template <class T>
__global__ void k_gather(const T* src, const uint32_t* idx, uint64_t n, T* dst) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[idx[i]];
}
// head[i] = 1 iff occurrence i starts a run of equal k-mers
template <int KW>
__global__ void k_run_heads(const uint64_t* key_lo, const uint64_t* key_hi, uint64_t n, uint32_t* head) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    head[i] = (i == 0 || key_lo[i] != key_lo[i - 1] || (KW == 2 && key_hi[i] != key_hi[i - 1])) ? 1u : 0u;
}
// first_occ[d] = i for the d-th run head (rank = exclusive scan of head)
__global__ void k_first_occ(const uint32_t* head, const uint32_t* rank, uint64_t n, uint64_t* first_occ) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i < n && head[i]) first_occ[rank[i]] = i;
}
// per distinct k-mer: the k-mer, the union of its occurrences' exts, the signature and length of its colour
template <int KW>
__global__ void k_distinct(const uint64_t* key_lo, const uint64_t* key_hi, const uint64_t* payload, const uint64_t* first_occ, uint64_t n_dist,
                           uint64_t* kmer_lo, uint64_t* kmer_hi, uint8_t* exts, uint64_t* sig_lo, uint64_t* sig_hi, uint32_t* clen) {
    const uint64_t d = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (d >= n_dist) return;
    const uint64_t i0 = first_occ[d], i1 = first_occ[d + 1];
    uint32_t e = 0, prev = NONE32, n = 0;
    uint64_t h1 = 0x243F6A8885A308D3ULL, h2 = 0x13198A2E03707344ULL;
    for (uint64_t i = i0; i < i1; i++) {
        const uint64_t pl = payload[i];
        e |= (uint32_t)(pl & 0xff);
        const uint32_t tx = (uint32_t)(pl >> 8);
        if (tx != prev) {   // sort + dedup (ref src/equiv_classes.rs:78-79)
            prev = tx;
            h1 = mix64(h1 ^ tx);
            h2 = mix64(h2 + 0x9E3779B97F4A7C15ULL * ((uint64_t)tx + 1));
            n++;
        }
    }
    kmer_lo[d] = key_lo[i0];
    if (KW == 2) kmer_hi[d] = key_hi[i0];
    exts[d] = (uint8_t)e;
    sig_lo[d] = h1; sig_hi[d] = h2; clen[d] = n;
}
// runs of equal signature in `order` (distinct k-mers sorted by signature, ascending index within a run): the first of a
// run is the colour's representative; every other member's list is compared with it exactly
__global__ void k_colour_reps(const uint32_t* order, const uint64_t* sig_lo, const uint64_t* sig_hi, uint64_t n_dist, uint32_t* run_head) {
    const uint64_t q = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (q >= n_dist) return;
    const uint32_t d = order[q];
    bool head = q == 0;
    if (!head) {
        const uint32_t pd = order[q - 1];
        head = sig_lo[d] != sig_lo[pd] || sig_hi[d] != sig_hi[pd];
    }
    run_head[q] = head ? 1u : 0u;
}
// rep_at[q] = position in `order` of the head of q's run (inclusive max-scan of head positions, done as: scan of heads -> run id,
// then heads scatter their position)
__global__ void k_scatter_run_pos(const uint32_t* run_head, const uint32_t* run_id, uint64_t n, uint32_t* run_pos) {
    const uint64_t q = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (q < n && run_head[q]) run_pos[run_id[q]] = (uint32_t)q;
}
__global__ void k_assign_reps(const uint32_t* order, const uint32_t* run_head, const uint32_t* run_id, const uint32_t* run_pos, const uint64_t* payload,
                              const uint64_t* first_occ, uint64_t n_dist, uint32_t* rep_of, uint32_t* is_rep, uint32_t* status) {
    const uint64_t q = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (q >= n_dist) return;
    const uint32_t d = order[q], rep = order[run_pos[run_id[q]]];
    rep_of[d] = rep;
    is_rep[d] = run_head[q];
    if (rep == d) return;
    // colour_equal(rep, d): both lists de-duplicated on the fly
    uint64_t i = first_occ[rep], ie = first_occ[rep + 1], j = first_occ[d], je = first_occ[d + 1];
    bool same = true;
    while (i < ie && j < je) {
        const uint32_t a = (uint32_t)(payload[i] >> 8), b = (uint32_t)(payload[j] >> 8);
        if (a != b) { same = false; break; }
        while (i < ie && (uint32_t)(payload[i] >> 8) == a) i++;
        while (j < je && (uint32_t)(payload[j] >> 8) == a) j++;
    }
    if (!same || i != ie || j != je) atomicOr(status, 2u);   // two colours with one 128-bit signature
}
__global__ void k_eq_of(const uint32_t* rep_of, const uint32_t* class_rank, uint64_t n_dist, uint32_t* eq) {
    const uint64_t d = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (d < n_dist) eq[d] = class_rank[rep_of[d]];
}
__global__ void k_class_lens(const uint32_t* reps, const uint32_t* clen, uint64_t n_eq, uint64_t* len64) {
    const uint64_t c = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (c < n_eq) len64[c] = clen[reps[c]];
    if (c == n_eq) len64[c] = 0;
}
__global__ void k_class_members(const uint32_t* reps, const uint64_t* first_occ, const uint64_t* payload, const uint64_t* eq_off, uint64_t n_eq,
                                uint32_t* members) {
    const uint64_t c = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (c >= n_eq) return;
    uint64_t o = eq_off[c];
    uint32_t prev = NONE32;
    for (uint64_t i = first_occ[reps[c]]; i < first_occ[reps[c] + 1]; i++) {
        const uint32_t tx = (uint32_t)(payload[i] >> 8);
        if (tx != prev) { prev = tx; members[o++] = tx; }
    }
}
template <int KW>
__device__ __forceinline__ uint32_t km_find(const uint64_t* kmer_lo, const uint64_t* kmer_hi, uint64_t n, Km<KW> x) {
    uint64_t lo = 0, hi = n;
    while (lo < hi) {
        const uint64_t mid = lo + ((hi - lo) >> 1);
        const Km<KW> y{kmer_lo[mid], KW == 2 ? kmer_hi[mid] : 0};
        if (y < x) lo = mid + 1;
        else hi = mid;
    }
    if (lo >= n) return NONE32;
    const Km<KW> y{kmer_lo[lo], KW == 2 ? kmer_hi[lo] : 0};
    return y == x ? (uint32_t)lo : NONE32;
}
__device__ __forceinline__ bool one_bit(uint32_t x) { return x && !(x & (x - 1)); }
// fwd[i] = j iff i -> j is an internal unitig link
template <int KW>
__global__ void k_links(const uint64_t* kmer_lo, const uint64_t* kmer_hi, const uint8_t* exts, const uint32_t* eq, uint64_t n_dist, uint32_t k,
                        uint32_t* fwd, uint32_t* not_target, uint32_t* status) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= n_dist) return;
    const uint32_t r = exts[i] & 0xf;
    if (!one_bit(r)) return;
    const uint32_t b = (uint32_t)__ffs((int)r) - 1;
    const Km<KW> me{kmer_lo[i], KW == 2 ? kmer_hi[i] : 0};
    const uint32_t j = km_find<KW>(kmer_lo, kmer_hi, n_dist, km_push<KW>(me, b, k));
    if (j == NONE32) { atomicOr(status, 4u); return; }   // an observed neighbour must exist
    if (j == i) return;                                   // self loop (e.g. poly-A): a path of its own
    if (!one_bit((uint32_t)exts[j] >> 4)) return;
    if (eq[j] != eq[i]) return;
    fwd[i] = j;
    not_target[j] = 0;   // unique writer: j has exactly one left ext
}
__global__ void k_fill32(uint32_t* a, uint64_t n, uint32_t v) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i < n) a[i] = v;
}
__global__ void k_walk_len(const uint32_t* heads, uint64_t n_heads, const uint32_t* fwd, uint32_t* unvisited, uint32_t* plen) {
    const uint64_t h = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (h >= n_heads) return;
    uint32_t cur = heads[h], n = 0;
    while (cur != NONE32) { unvisited[cur] = 0; n++; cur = fwd[cur]; }
    plen[h] = n;
}
__global__ void k_node_lens(const uint32_t* plen, uint64_t n_nodes, uint32_t k, uint32_t* node_len, uint64_t* len64) {
    const uint64_t v = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (v < n_nodes) { node_len[v] = k + plen[v] - 1; len64[v] = k + plen[v] - 1; }
    if (v == n_nodes) len64[v] = 0;
}
template <int KW>
__global__ void k_emit_nodes(const uint32_t* heads, const uint32_t* plen, const uint64_t* node_start, uint64_t n_nodes, uint32_t k, const uint64_t* kmer_lo,
                             const uint64_t* kmer_hi, const uint8_t* exts, const uint32_t* eq, const uint32_t* fwd, unsigned long long* W,
                             uint8_t* node_exts, uint32_t* node_eq) {
    const uint64_t v = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (v >= n_nodes) return;
    uint32_t cur = heads[v], last = cur;
    uint64_t pos = node_start[v];
    unsigned long long acc = 0;   // bits accumulated for word pos / 32
    auto put = [&](uint32_t b) {
        acc |= (unsigned long long)b << (62 - 2 * (pos & 31));
        pos++;
        if ((pos & 31) == 0) { atomicOr(W + ((pos - 1) >> 5), acc); acc = 0; }
    };
    for (uint32_t t = 0; t < k; t++) {   // base t of the first k-mer
        const uint32_t sh = 2 * (k - 1 - t);
        const uint64_t w = (KW == 2 && sh >= 64) ? kmer_hi[cur] >> (sh - 64) : kmer_lo[cur] >> sh;
        put((uint32_t)w & 3u);
    }
    const uint32_t steps = plen[v];
    for (uint32_t q = 1; q < steps; q++) {
        cur = fwd[cur];
        put((uint32_t)kmer_lo[cur] & 3u);
        last = cur;
    }
    if (pos & 31) atomicOr(W + (pos >> 5), acc);
    node_exts[v] = (uint8_t)((exts[heads[v]] & 0xf0) | (exts[last] & 0x0f));
    node_eq[v] = eq[heads[v]];
}
struct U32ToU64 {
    const uint32_t* a;
    __host__ __device__ uint64_t operator()(uint64_t i) const { return a[i]; }
};

template <class T>
T* host_copy(const void* dev, uint64_t n, cudaError_t& e) {
    T* h = (T*)malloc((n ? n : 1) * sizeof(T));
    if (h && n && e == cudaSuccess) e = cudaMemcpy(h, dev, n * sizeof(T), cudaMemcpyDeviceToHost);
    return h;
}

template <int KW>
int build(int device, const uint8_t* codes_h, const uint64_t* tx_off_h, uint32_t n_tx, uint32_t k, psa_built_graph* out) {
    BCU(cudaSetDevice(device));
    // ---- occurrences per transcript (host: n_tx additions)
    std::vector<uint64_t> occ_off(n_tx + 1, 0);
    for (uint32_t t = 0; t < n_tx; t++) {
        const uint64_t len = tx_off_h[t + 1] - tx_off_h[t];
        occ_off[t + 1] = occ_off[t] + (len >= k ? len - k + 1 : 0);
    }
    const uint64_t n_occ = occ_off[n_tx], n_bases_in = tx_off_h[n_tx] - tx_off_h[0];
    if (n_occ >= (1ull << 31)) return psa_internal_fail(PSA_ERR_ARG, "more than 2^31 k-mer occurrences: build the graph on the host");
    memset(out, 0, sizeof *out);
    out->k = k;
    Buf codes, tx_off, occ_offd, status;
    BCU(codes.alloc(n_bases_in + 1));
    BCU(tx_off.alloc((n_tx + 1) * 8));
    BCU(occ_offd.alloc((n_tx + 1) * 8));
    BCU(status.alloc(8));
    BCU(cudaMemcpy(codes.p, codes_h + tx_off_h[0], n_bases_in, cudaMemcpyHostToDevice));
    {
        std::vector<uint64_t> rel(tx_off_h, tx_off_h + n_tx + 1);
        for (auto& x : rel) x -= tx_off_h[0];
        BCU(cudaMemcpy(tx_off.p, rel.data(), (n_tx + 1) * 8, cudaMemcpyHostToDevice));
    }
    BCU(cudaMemcpy(occ_offd.p, occ_off.data(), (n_tx + 1) * 8, cudaMemcpyHostToDevice));
    BCU(cudaMemset(status.p, 0, 8));
    auto status_bits = [&](uint32_t& bits) -> cudaError_t { return cudaMemcpy(&bits, status.p, 4, cudaMemcpyDeviceToHost); };

    // ---- enumerate + sort by k-mer (stable: the transcripts of a k-mer stay ascending)
    Buf key_lo, key_hi, payload, tmp;
    BCU(key_lo.alloc(n_occ * 8));
    BCU(payload.alloc(n_occ * 8));
    if (KW == 2) BCU(key_hi.alloc(n_occ * 8));
    if (n_occ) k_occ_enumerate<KW><<<blocks(n_occ, 256), 256>>>(codes.as<uint8_t>(), tx_off.as<uint64_t>(), occ_offd.as<uint64_t>(), n_tx, k, n_occ,
                                                                key_lo.as<uint64_t>(), key_hi.as<uint64_t>(), payload.as<uint64_t>(), status.as<uint32_t>());
    BCU(cudaGetLastError());
    {
        uint32_t bits = 0;
        BCU(status_bits(bits));
        if (bits & 1u) return psa_internal_fail(PSA_ERR_ARG, "base code > 3 in a transcript");
    }
    codes.release();
    const int nbits = 2 * (int)std::min<uint32_t>(k, 32);
    if (KW == 1) {
        Buf k2, p2;
        BCU(k2.alloc(n_occ * 8));
        BCU(p2.alloc(n_occ * 8));
        size_t tb = 0;
        BCU(cub::DeviceRadixSort::SortPairs(nullptr, tb, key_lo.as<uint64_t>(), k2.as<uint64_t>(), payload.as<uint64_t>(), p2.as<uint64_t>(), (int)n_occ, 0, nbits));
        BCU(tmp.alloc(tb));
        BCU(cub::DeviceRadixSort::SortPairs(tmp.p, tb, key_lo.as<uint64_t>(), k2.as<uint64_t>(), payload.as<uint64_t>(), p2.as<uint64_t>(), (int)n_occ, 0, nbits));
        BCU(cudaDeviceSynchronize());
        std::swap(key_lo.p, k2.p);
        std::swap(payload.p, p2.p);
        tmp.release();
    } else {
        // 128-bit keys: least significant word first, then the (stable) sort by the most significant word
        Buf idx0, idx1, idx2, lo1, hi1, hi2;
        BCU(idx0.alloc(n_occ * 4)); BCU(idx1.alloc(n_occ * 4)); BCU(idx2.alloc(n_occ * 4));
        BCU(lo1.alloc(n_occ * 8)); BCU(hi1.alloc(n_occ * 8)); BCU(hi2.alloc(n_occ * 8));
        if (n_occ) k_iota32<<<blocks(n_occ, 256), 256>>>(idx0.as<uint32_t>(), n_occ);
        size_t tb = 0;
        BCU(cub::DeviceRadixSort::SortPairs(nullptr, tb, key_lo.as<uint64_t>(), lo1.as<uint64_t>(), idx0.as<uint32_t>(), idx1.as<uint32_t>(), (int)n_occ, 0, 64));
        BCU(tmp.alloc(tb));
        BCU(cub::DeviceRadixSort::SortPairs(tmp.p, tb, key_lo.as<uint64_t>(), lo1.as<uint64_t>(), idx0.as<uint32_t>(), idx1.as<uint32_t>(), (int)n_occ, 0, 64));
        if (n_occ) k_gather<uint64_t><<<blocks(n_occ, 256), 256>>>(key_hi.as<uint64_t>(), idx1.as<uint32_t>(), n_occ, hi1.as<uint64_t>());
        const int hbits = 2 * (int)(k - 32);
        BCU(cub::DeviceRadixSort::SortPairs(tmp.p, tb, hi1.as<uint64_t>(), hi2.as<uint64_t>(), idx1.as<uint32_t>(), idx2.as<uint32_t>(), (int)n_occ, 0, hbits));
        // permute lo and payload by the final order
        if (n_occ) {
            k_gather<uint64_t><<<blocks(n_occ, 256), 256>>>(key_lo.as<uint64_t>(), idx2.as<uint32_t>(), n_occ, lo1.as<uint64_t>());
            k_gather<uint64_t><<<blocks(n_occ, 256), 256>>>(payload.as<uint64_t>(), idx2.as<uint32_t>(), n_occ, hi1.as<uint64_t>());
        }
        BCU(cudaDeviceSynchronize());
        std::swap(key_lo.p, lo1.p);
        std::swap(payload.p, hi1.p);
        std::swap(key_hi.p, hi2.p);
        tmp.release();
    }

    // ---- distinct k-mers
    Buf head, rank, first_occ;
    BCU(head.alloc((n_occ + 1) * 4));
    BCU(rank.alloc((n_occ + 1) * 4));
    BCU(cudaMemset(head.p, 0, (n_occ + 1) * 4));
    if (n_occ) k_run_heads<KW><<<blocks(n_occ, 256), 256>>>(key_lo.as<uint64_t>(), key_hi.as<uint64_t>(), n_occ, head.as<uint32_t>());
    {
        size_t tb = 0;
        BCU(cub::DeviceScan::ExclusiveSum(nullptr, tb, head.as<uint32_t>(), rank.as<uint32_t>(), (int)(n_occ + 1)));
        BCU(tmp.alloc(tb));
        BCU(cub::DeviceScan::ExclusiveSum(tmp.p, tb, head.as<uint32_t>(), rank.as<uint32_t>(), (int)(n_occ + 1)));
        tmp.release();
    }
    uint32_t n_dist32 = 0;
    BCU(cudaMemcpy(&n_dist32, rank.as<uint32_t>() + n_occ, 4, cudaMemcpyDeviceToHost));
    const uint64_t n_dist = n_dist32;
    out->n_kmers = n_dist;
    BCU(first_occ.alloc((n_dist + 2) * 8));
    if (n_occ) k_first_occ<<<blocks(n_occ, 256), 256>>>(head.as<uint32_t>(), rank.as<uint32_t>(), n_occ, first_occ.as<uint64_t>());
    BCU(cudaMemcpy(first_occ.as<uint64_t>() + n_dist, &n_occ, 8, cudaMemcpyHostToDevice));
    head.release();
    rank.release();
    Buf kmer_lo, kmer_hi, exts, sig_lo, sig_hi, clen;
    BCU(kmer_lo.alloc(n_dist * 8));
    if (KW == 2) BCU(kmer_hi.alloc(n_dist * 8));
    BCU(exts.alloc(n_dist));
    BCU(sig_lo.alloc(n_dist * 8));
    BCU(sig_hi.alloc(n_dist * 8));
    BCU(clen.alloc((n_dist + 1) * 4));
    if (n_dist)
        k_distinct<KW><<<blocks(n_dist, 256), 256>>>(key_lo.as<uint64_t>(), key_hi.as<uint64_t>(), payload.as<uint64_t>(), first_occ.as<uint64_t>(), n_dist,
                                                     kmer_lo.as<uint64_t>(), kmer_hi.as<uint64_t>(), exts.as<uint8_t>(), sig_lo.as<uint64_t>(),
                                                     sig_hi.as<uint64_t>(), clen.as<uint32_t>());
    BCU(cudaGetLastError());
    key_lo.release();
    key_hi.release();

    // ---- intern colours: sort the distinct k-mers by signature (stable: ascending index within equal signatures)
    Buf order, rep_of, is_rep, eq;
    BCU(order.alloc((n_dist + 1) * 4));
    BCU(rep_of.alloc((n_dist + 1) * 4));
    BCU(is_rep.alloc((n_dist + 1) * 4));
    BCU(eq.alloc((n_dist + 1) * 4));
    uint64_t n_eq = 0;
    Buf reps, eq_off, members;
    {
        Buf i0, i1, s1, s2, g1, run_head, run_id, run_pos;
        BCU(i0.alloc((n_dist + 1) * 4)); BCU(i1.alloc((n_dist + 1) * 4));
        BCU(s1.alloc((n_dist + 1) * 8)); BCU(s2.alloc((n_dist + 1) * 8)); BCU(g1.alloc((n_dist + 1) * 8));
        BCU(run_head.alloc((n_dist + 1) * 4)); BCU(run_id.alloc((n_dist + 1) * 4)); BCU(run_pos.alloc((n_dist + 1) * 4));
        if (n_dist) k_iota32<<<blocks(n_dist, 256), 256>>>(i0.as<uint32_t>(), n_dist);
        size_t tb = 0;
        BCU(cub::DeviceRadixSort::SortPairs(nullptr, tb, sig_lo.as<uint64_t>(), s1.as<uint64_t>(), i0.as<uint32_t>(), i1.as<uint32_t>(), (int)n_dist, 0, 64));
        BCU(tmp.alloc(tb));
        BCU(cub::DeviceRadixSort::SortPairs(tmp.p, tb, sig_lo.as<uint64_t>(), s1.as<uint64_t>(), i0.as<uint32_t>(), i1.as<uint32_t>(), (int)n_dist, 0, 64));
        if (n_dist) k_gather<uint64_t><<<blocks(n_dist, 256), 256>>>(sig_hi.as<uint64_t>(), i1.as<uint32_t>(), n_dist, g1.as<uint64_t>());
        BCU(cub::DeviceRadixSort::SortPairs(tmp.p, tb, g1.as<uint64_t>(), s2.as<uint64_t>(), i1.as<uint32_t>(), order.as<uint32_t>(), (int)n_dist, 0, 64));
        tmp.release();
        BCU(cudaMemset(run_head.p, 0, (n_dist + 1) * 4));
        if (n_dist) k_colour_reps<<<blocks(n_dist, 256), 256>>>(order.as<uint32_t>(), sig_lo.as<uint64_t>(), sig_hi.as<uint64_t>(), n_dist, run_head.as<uint32_t>());
        {   // run id = inclusive scan of heads - 1
            size_t tb2 = 0;
            BCU(cub::DeviceScan::InclusiveSum(nullptr, tb2, run_head.as<uint32_t>(), run_id.as<uint32_t>(), (int)(n_dist + 1)));
            BCU(tmp.alloc(tb2));
            BCU(cub::DeviceScan::InclusiveSum(tmp.p, tb2, run_head.as<uint32_t>(), run_id.as<uint32_t>(), (int)(n_dist + 1)));
            tmp.release();
        }
        // run_id is 1-based now; run_pos is indexed with it
        BCU(cudaMemset(is_rep.p, 0, (n_dist + 1) * 4));
        if (n_dist) {
            k_scatter_run_pos<<<blocks(n_dist, 256), 256>>>(run_head.as<uint32_t>(), run_id.as<uint32_t>(), n_dist, run_pos.as<uint32_t>());
            k_assign_reps<<<blocks(n_dist, 256), 256>>>(order.as<uint32_t>(), run_head.as<uint32_t>(), run_id.as<uint32_t>(), run_pos.as<uint32_t>(),
                                                        payload.as<uint64_t>(), first_occ.as<uint64_t>(), n_dist, rep_of.as<uint32_t>(), is_rep.as<uint32_t>(),
                                                        status.as<uint32_t>());
        }
        BCU(cudaGetLastError());
        uint32_t bits = 0;
        BCU(status_bits(bits));
        if (bits & 2u) return psa_internal_fail(PSA_ERR_INTERNAL, "two colours share a 128-bit signature (never merged silently): build the graph on the host");
        // dense class ids in order of first appearance over the sorted k-mers = rank of the representative
        Buf class_rank;
        BCU(class_rank.alloc((n_dist + 1) * 4));
        size_t tb3 = 0;
        BCU(cub::DeviceScan::ExclusiveSum(nullptr, tb3, is_rep.as<uint32_t>(), class_rank.as<uint32_t>(), (int)(n_dist + 1)));
        BCU(tmp.alloc(tb3));
        BCU(cub::DeviceScan::ExclusiveSum(tmp.p, tb3, is_rep.as<uint32_t>(), class_rank.as<uint32_t>(), (int)(n_dist + 1)));
        tmp.release();
        uint32_t n_eq32 = 0;
        BCU(cudaMemcpy(&n_eq32, class_rank.as<uint32_t>() + n_dist, 4, cudaMemcpyDeviceToHost));
        n_eq = n_eq32;
        if (n_dist) k_eq_of<<<blocks(n_dist, 256), 256>>>(rep_of.as<uint32_t>(), class_rank.as<uint32_t>(), n_dist, eq.as<uint32_t>());
        // representatives in id order, class lengths, offsets, members
        Buf nsel, len64;
        BCU(reps.alloc((n_eq + 1) * 4));
        BCU(nsel.alloc(8));
        BCU(len64.alloc((n_eq + 2) * 8));
        BCU(eq_off.alloc((n_eq + 2) * 8));
        {
            cub::CountingInputIterator<uint32_t> it(0);
            size_t tb4 = 0;
            BCU(cub::DeviceSelect::Flagged(nullptr, tb4, it, is_rep.as<uint32_t>(), reps.as<uint32_t>(), nsel.as<uint32_t>(), (int)n_dist));
            BCU(tmp.alloc(tb4));
            BCU(cub::DeviceSelect::Flagged(tmp.p, tb4, it, is_rep.as<uint32_t>(), reps.as<uint32_t>(), nsel.as<uint32_t>(), (int)n_dist));
            tmp.release();
        }
        k_class_lens<<<blocks(n_eq + 1, 256), 256>>>(reps.as<uint32_t>(), clen.as<uint32_t>(), n_eq, len64.as<uint64_t>());
        size_t tb5 = 0;
        BCU(cub::DeviceScan::ExclusiveSum(nullptr, tb5, len64.as<uint64_t>(), eq_off.as<uint64_t>(), (int)(n_eq + 1)));
        BCU(tmp.alloc(tb5));
        BCU(cub::DeviceScan::ExclusiveSum(tmp.p, tb5, len64.as<uint64_t>(), eq_off.as<uint64_t>(), (int)(n_eq + 1)));
        tmp.release();
        uint64_t n_mem = 0;
        BCU(cudaMemcpy(&n_mem, eq_off.as<uint64_t>() + n_eq, 8, cudaMemcpyDeviceToHost));
        out->n_eq = n_eq;
        out->n_eq_members = n_mem;
        BCU(members.alloc((n_mem + 1) * 4));
        if (n_eq) k_class_members<<<blocks(n_eq, 128), 128>>>(reps.as<uint32_t>(), first_occ.as<uint64_t>(), payload.as<uint64_t>(), eq_off.as<uint64_t>(), n_eq,
                                                            members.as<uint32_t>());
        BCU(cudaGetLastError());
        BCU(cudaDeviceSynchronize());
    }
    payload.release();
    first_occ.release();
    sig_lo.release();
    sig_hi.release();
    order.release();
    rep_of.release();
    clen.release();

    // ---- links, heads, path lengths
    Buf fwd, not_target, heads, nheads, unvisited, plen;
    BCU(fwd.alloc((n_dist + 1) * 4));
    BCU(not_target.alloc((n_dist + 1) * 4));
    BCU(unvisited.alloc((n_dist + 1) * 4));
    BCU(nheads.alloc(8));
    if (n_dist) {
        k_fill32<<<blocks(n_dist, 256), 256>>>(fwd.as<uint32_t>(), n_dist, NONE32);
        k_fill32<<<blocks(n_dist, 256), 256>>>(not_target.as<uint32_t>(), n_dist, 1u);
        k_fill32<<<blocks(n_dist, 256), 256>>>(unvisited.as<uint32_t>(), n_dist, 1u);
        k_links<KW><<<blocks(n_dist, 256), 256>>>(kmer_lo.as<uint64_t>(), kmer_hi.as<uint64_t>(), exts.as<uint8_t>(), eq.as<uint32_t>(), n_dist, k,
                                                  fwd.as<uint32_t>(), not_target.as<uint32_t>(), status.as<uint32_t>());
    }
    BCU(cudaGetLastError());
    {
        uint32_t bits = 0;
        BCU(status_bits(bits));
        if (bits & 4u) return psa_internal_fail(PSA_ERR_INTERNAL, "k-mer neighbour missing (internal)");
    }
    BCU(heads.alloc((n_dist + 1) * 4));
    {
        cub::CountingInputIterator<uint32_t> it(0);
        size_t tb = 0;
        BCU(cub::DeviceSelect::Flagged(nullptr, tb, it, not_target.as<uint32_t>(), heads.as<uint32_t>(), nheads.as<uint32_t>(), (int)n_dist));
        BCU(tmp.alloc(tb));
        BCU(cub::DeviceSelect::Flagged(tmp.p, tb, it, not_target.as<uint32_t>(), heads.as<uint32_t>(), nheads.as<uint32_t>(), (int)n_dist));
        tmp.release();
    }
    uint32_t n_heads32 = 0;
    BCU(cudaMemcpy(&n_heads32, nheads.p, 4, cudaMemcpyDeviceToHost));
    uint64_t n_nodes = n_heads32;
    BCU(plen.alloc((n_dist + 1) * 4));
    if (n_nodes) k_walk_len<<<blocks(n_nodes, 128), 128>>>(heads.as<uint32_t>(), n_nodes, fwd.as<uint32_t>(), unvisited.as<uint32_t>(), plen.as<uint32_t>());
    BCU(cudaGetLastError());
    // closed cycles: every member is a link target, so none was reached; cut at the smallest k-mer (host: they are rare)
    {
        Buf left, nleft;
        BCU(left.alloc((n_dist + 1) * 4));
        BCU(nleft.alloc(8));
        cub::CountingInputIterator<uint32_t> it(0);
        size_t tb = 0;
        BCU(cub::DeviceSelect::Flagged(nullptr, tb, it, unvisited.as<uint32_t>(), left.as<uint32_t>(), nleft.as<uint32_t>(), (int)n_dist));
        BCU(tmp.alloc(tb));
        BCU(cub::DeviceSelect::Flagged(tmp.p, tb, it, unvisited.as<uint32_t>(), left.as<uint32_t>(), nleft.as<uint32_t>(), (int)n_dist));
        tmp.release();
        uint32_t n_left = 0;
        BCU(cudaMemcpy(&n_left, nleft.p, 4, cudaMemcpyDeviceToHost));
        if (n_left) {
            std::vector<uint32_t> lv(n_left), nxt(n_left);
            BCU(cudaMemcpy(lv.data(), left.p, (size_t)n_left * 4, cudaMemcpyDeviceToHost));
            Buf nd;
            BCU(nd.alloc((size_t)n_left * 4));
            k_gather<uint32_t><<<blocks(n_left, 256), 256>>>(fwd.as<uint32_t>(), left.as<uint32_t>(), n_left, nd.as<uint32_t>());
            BCU(cudaMemcpy(nxt.data(), nd.p, (size_t)n_left * 4, cudaMemcpyDeviceToHost));
            // lv is ascending; a cycle's members are all in lv
            std::vector<uint8_t> seen(n_left, 0);
            std::vector<uint32_t> ch, cl;
            for (uint32_t a = 0; a < n_left; a++) {
                if (seen[a]) continue;
                uint32_t cur = a, n = 0;
                while (!seen[cur]) {
                    seen[cur] = 1;
                    n++;
                    const uint32_t j = nxt[cur];
                    if (j == NONE32) break;
                    const auto itj = std::lower_bound(lv.begin(), lv.end(), j);
                    if (itj == lv.end() || *itj != j) break;
                    cur = (uint32_t)(itj - lv.begin());
                }
                ch.push_back(lv[a]);
                cl.push_back(n);
                out->n_cycles++;
            }
            BCU(cudaMemcpy(heads.as<uint32_t>() + n_nodes, ch.data(), ch.size() * 4, cudaMemcpyHostToDevice));
            BCU(cudaMemcpy(plen.as<uint32_t>() + n_nodes, cl.data(), cl.size() * 4, cudaMemcpyHostToDevice));
            n_nodes += ch.size();
        }
    }
    out->n_nodes = n_nodes;

    // ---- emit nodes
    Buf node_len, node_start, len64, seq, node_exts, node_eq;
    BCU(node_len.alloc((n_nodes + 1) * 4));
    BCU(len64.alloc((n_nodes + 2) * 8));
    BCU(node_start.alloc((n_nodes + 2) * 8));
    BCU(node_exts.alloc(n_nodes + 1));
    BCU(node_eq.alloc((n_nodes + 1) * 4));
    k_node_lens<<<blocks(n_nodes + 1, 256), 256>>>(plen.as<uint32_t>(), n_nodes, k, node_len.as<uint32_t>(), len64.as<uint64_t>());
    {
        size_t tb = 0;
        BCU(cub::DeviceScan::ExclusiveSum(nullptr, tb, len64.as<uint64_t>(), node_start.as<uint64_t>(), (int)(n_nodes + 1)));
        BCU(tmp.alloc(tb));
        BCU(cub::DeviceScan::ExclusiveSum(tmp.p, tb, len64.as<uint64_t>(), node_start.as<uint64_t>(), (int)(n_nodes + 1)));
        tmp.release();
    }
    uint64_t n_bases = 0;
    BCU(cudaMemcpy(&n_bases, node_start.as<uint64_t>() + n_nodes, 8, cudaMemcpyDeviceToHost));
    const uint64_t n_words = (n_bases + 31) / 32;
    out->n_seq_words = n_words;
    BCU(seq.alloc((n_words + 1) * 8));
    BCU(cudaMemset(seq.p, 0, (n_words + 1) * 8));
    if (n_nodes)
        k_emit_nodes<KW><<<blocks(n_nodes, 128), 128>>>(heads.as<uint32_t>(), plen.as<uint32_t>(), node_start.as<uint64_t>(), n_nodes, k, kmer_lo.as<uint64_t>(),
                                                        kmer_hi.as<uint64_t>(), exts.as<uint8_t>(), eq.as<uint32_t>(), fwd.as<uint32_t>(),
                                                        seq.as<unsigned long long>(), node_exts.as<uint8_t>(), node_eq.as<uint32_t>());
    BCU(cudaGetLastError());
    BCU(cudaDeviceSynchronize());

    // ---- results to the host
    cudaError_t e = cudaSuccess;
    out->seq_words = host_copy<uint64_t>(seq.p, n_words, e);
    out->node_start = host_copy<uint64_t>(node_start.p, n_nodes, e);
    out->node_len = host_copy<uint32_t>(node_len.p, n_nodes, e);
    out->node_exts = host_copy<uint8_t>(node_exts.p, n_nodes, e);
    out->node_eq = host_copy<uint32_t>(node_eq.p, n_nodes, e);
    out->eq_offsets = host_copy<uint64_t>(eq_off.p, n_eq + 1, e);
    out->eq_members = host_copy<uint32_t>(members.p, out->n_eq_members, e);
    if (!out->seq_words || !out->node_start || !out->node_len || !out->node_exts || !out->node_eq || !out->eq_offsets || !out->eq_members) {
        psa_built_graph_free(out);
        return psa_internal_fail(PSA_ERR_NOMEM, "out of host memory");
    }
    if (e != cudaSuccess) {
        psa_built_graph_free(out);
        return psa_internal_fail(PSA_ERR_CUDA, cudaGetErrorString(e));
    }
    return PSA_OK;
}
}  // namespace

extern "C" void psa_built_graph_free(psa_built_graph* g) {
    if (!g) return;
    free(g->seq_words); free(g->node_start); free(g->node_len); free(g->node_exts); free(g->node_eq); free(g->eq_offsets); free(g->eq_members);
    g->seq_words = nullptr; g->node_start = nullptr; g->node_len = nullptr; g->node_exts = nullptr; g->node_eq = nullptr;
    g->eq_offsets = nullptr; g->eq_members = nullptr;
}

extern "C" int psa_build_graph_device(int device, const uint8_t* codes, const uint64_t* tx_off, uint32_t n_tx, uint32_t k, psa_built_graph* out) {
    if (!tx_off || !out || (n_tx && !codes)) return psa_internal_fail(PSA_ERR_ARG, "null argument");
    if (k < 2 || k > 64) return psa_internal_fail(PSA_ERR_ARG, "k must be in 2..64");
    return k <= 32 ? build<1>(device, codes, tx_off, n_tx, k, out) : build<2>(device, codes, tx_off, n_tx, k, out);
}
