// psa_api.cu -- the C ABI of include/psa.h: index upload + on-device MPHF/edge construction,
// the mapper (batch pipeline around k_map), counts, NCCL all-reduce of the counts.
// There is no CPU implementation of any of this: without a CUDA device every compute entry
// returns PSA_ERR_CUDA.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>
#include <cub/iterator/counting_input_iterator.cuh>
#include <cub/iterator/transform_input_iterator.cuh>
#include <mutex>
#include <string>
#include <utility>
#include <vector>

#include "../../include/psa.h"
#include "psa_kernels.cuh"
#include "psa_fastq.cuh"
#include "psa_fastq.h"

using namespace psa;

static_assert(sizeof(HitRec) == sizeof(psa_hit), "HitRec must mirror psa_hit");

// ---------------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}
#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            (void)cudaGetLastError();                                                              \
            return fail(PSA_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));         \
        }                                                                                          \
    } while (0)

extern "C" int psa_internal_fail(int code, const char* msg) { return fail(code, msg ? msg : ""); }   // for the other translation units

extern "C" const char* psa_strerror(int code) {
    switch (code) {
        case PSA_OK: return "ok";
        case PSA_ERR_ARG: return "bad argument";
        case PSA_ERR_CUDA: return "CUDA error";
        case PSA_ERR_NOMEM: return "out of host memory";
        case PSA_ERR_CAPACITY: return "output buffer too small";
        case PSA_ERR_INDEX: return "invalid index graph";
        case PSA_ERR_NCCL: return "NCCL error";
        case PSA_ERR_IO: return "I/O error";
        case PSA_ERR_INTERNAL: return "internal self-check failed";
        default: return "unknown error";
    }
}
extern "C" const char* psa_last_error(void) { return g_err.c_str(); }
extern "C" int psa_abi_version(void) { return PSA_ABI_VERSION; }

static inline unsigned nblocks(uint64_t n, unsigned bs) { return (unsigned)((n + bs - 1) / bs); }

// ---------------------------------------------------------------------------------------------
// small device buffer
// ---------------------------------------------------------------------------------------------
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    bool view = false;  // points into another allocation: never freed here
    void set_view(void* q, size_t bytes) {
        p = q;
        cap = bytes;
        view = true;
    }
    int ensure(size_t bytes) {  // contents are NOT preserved on growth
        if (bytes <= cap) return PSA_OK;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        CU(cudaMalloc(&p, want));
        cap = want;
        return PSA_OK;
    }
    void release() {
        if (p && !view) cudaFree(p);
        p = nullptr;
        cap = 0;
        view = false;
    }
    template <class T>
    T* as() const { return reinterpret_cast<T*>(p); }
};

// ---------------------------------------------------------------------------------------------
// index
// ---------------------------------------------------------------------------------------------
struct psa_index {
    int device = 0;
    DevIndex d{};
    DevBuf buckets, nodes, nodes_cold, seq, eq_off, eq_mem, class_win, bloom;
    std::vector<uint64_t> h_eq_off;   // eq_classes on the host too: compact results are expanded from them
    std::vector<uint32_t> h_eq_mem;
    psa_index_info info{};
    int kw = 1;
    int sms = 148;   // multiprocessors of the device
    // idle FASTQ text lanes (psa_fastq.h) kept for the next psa_process_reads call: pinning their buffers and creating
    // their mappers costs 0.1-0.2 s per call otherwise.  Freed with the index.
    std::mutex lanes_mu;
    std::vector<struct psa_fq_lane*> idle_lanes;
};

static uint32_t bits_for(uint64_t max_value) {  // bits needed to store 0..max_value
    uint32_t b = 1;
    while (b < 64 && (max_value >> b)) b++;
    return b;
}

template <int KW>
static int build_on_device(psa_index* ix, const psa_index_desc* d, double gamma, DevBuf& node_start,
                           const std::vector<uint64_t>& koff_host, uint64_t n_kmers) {
    const uint32_t k = d->k;
    cudaStream_t st = 0;
    DevBuf koff, key_lo[3], key_hi[3], val[3], bid[2], idx[2], sel, passed, nsel, err, tmp;
    auto cleanup = [&]() {
        koff.release(); sel.release(); passed.release(); nsel.release(); err.release(); tmp.release();
        for (int i = 0; i < 3; i++) { key_lo[i].release(); key_hi[i].release(); val[i].release(); }
        for (int i = 0; i < 2; i++) { bid[i].release(); idx[i].release(); }
    };
#define CUB_(call)                                                                                 \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            (void)cudaGetLastError();                                                              \
            cleanup();                                                                             \
            return fail(PSA_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));         \
        }                                                                                          \
    } while (0)
#define RC_(call)                      \
    do {                               \
        int rc_ = (call);              \
        if (rc_ != PSA_OK) {           \
            cleanup();                 \
            return rc_;                \
        }                              \
    } while (0)

    RC_(koff.ensure((d->n_nodes + 1) * 8));
    CUB_(cudaMemcpy(koff.p, koff_host.data(), (d->n_nodes + 1) * 8, cudaMemcpyHostToDevice));
    RC_(err.ensure(4));
    CUB_(cudaMemset(err.p, 0, 4));
    RC_(nsel.ensure(8));

    // 1. every k-mer of every node
    RC_(key_lo[0].ensure(n_kmers * 8 + 8));
    if (KW == 2) RC_(key_hi[0].ensure(n_kmers * 8 + 8));
    RC_(val[0].ensure(n_kmers * 8 + 8));
    if (n_kmers)
        k_enumerate_keys<KW><<<nblocks(n_kmers, 256), 256, 0, st>>>(
            ix->seq.as<uint64_t>(), node_start.as<uint64_t>(), koff.as<uint64_t>(), d->n_nodes, n_kmers, k,
            key_lo[0].as<uint64_t>(), key_hi[0].as<uint64_t>(), val[0].as<uint64_t>());
    CUB_(cudaGetLastError());

    // 1b. the absent-k-mer filter of k_seed_scan (PSA_BLOOM_BITS per key, default 10; 0: none)
    {
        double bits = 10.0;
        if (const char* e = getenv("PSA_BLOOM_BITS")) bits = atof(e);
        ix->d.bloom = nullptr;
        ix->d.bloom_blocks = 0;
        if (bits > 0 && n_kmers) {
            const uint64_t nb = std::max<uint64_t>(1, (uint64_t)((double)n_kmers * bits / 256.0) + 1);
            RC_(ix->bloom.ensure(nb * 32));
            CUB_(cudaMemsetAsync(ix->bloom.p, 0, nb * 32, st));
            k_bloom_set<KW><<<nblocks(n_kmers, 256), 256, 0, st>>>(key_lo[0].as<uint64_t>(), key_hi[0].as<uint64_t>(), n_kmers, ix->bloom.as<uint32_t>(), nb);
            CUB_(cudaGetLastError());
            ix->d.bloom = ix->bloom.as<uint32_t>();
            ix->d.bloom_blocks = nb;
        }
    }

    // 2. the dictionary: a cascade of bucket tables (psa_core.cuh Dict).  Level sizes depend on the keys the
    // level before passed on (~5 % at 1.7 slots per key); allocate for the geometric bound and check.
    std::vector<uint64_t> lvl_nbkt, lvl_base;
    uint64_t total_bkt = 0;
    {
        auto level_size = [&](uint64_t n) { return std::max<uint64_t>(1, (uint64_t)(gamma * (double)n / kBucketSlots) + 1); };
        const uint64_t cap_bkt = 2 * level_size(n_kmers) + 4 * kMaxLevels;
        RC_(ix->buckets.ensure(cap_bkt * 32));
        CUB_(cudaMemsetAsync(ix->buckets.p, 0xFF, ix->buckets.cap, st));  // kEmptyEntry
        RC_(passed.ensure(n_kmers + 8));
        RC_(sel.ensure(n_kmers * 4 + 8));
        for (int i = 0; i < 2; i++) {
            RC_(bid[i].ensure(n_kmers * 4 + 8));
            RC_(idx[i].ensure(n_kmers * 4 + 8));
        }
        uint64_t rem = n_kmers;
        int src = 0;  // buffer holding the remaining keys (0 = the full list, kept for the self-check)
        while (rem > 0) {
            if ((int)lvl_nbkt.size() >= kMaxLevels) {
                cleanup();
                return fail(PSA_ERR_INDEX, "dictionary did not converge: duplicate k-mer in the graph");
            }
            const uint64_t nbkt = level_size(rem);
            if (nbkt >= 0xFFFFFFFFull || total_bkt + nbkt > cap_bkt) {
                cleanup();
                return fail(PSA_ERR_INDEX, "dictionary bucket budget exceeded");
            }
            const uint32_t lvl = (uint32_t)lvl_nbkt.size();
            const uint32_t n32 = (uint32_t)rem;
            k_dict_bucket_ids<KW><<<nblocks(rem, 256), 256, 0, st>>>(key_lo[src].as<uint64_t>(), key_hi[src].as<uint64_t>(), n32, lvl,
                                                                      nbkt, bid[0].as<uint32_t>(), idx[0].as<uint32_t>());
            int end_bit = 1;
            while (end_bit < 32 && (nbkt >> end_bit)) end_bit++;
            size_t tb = 0;
            CUB_(cub::DeviceRadixSort::SortPairs(nullptr, tb, bid[0].as<uint32_t>(), bid[1].as<uint32_t>(), idx[0].as<uint32_t>(),
                                                 idx[1].as<uint32_t>(), (int)n32, 0, end_bit, st));
            RC_(tmp.ensure(tb));
            CUB_(cub::DeviceRadixSort::SortPairs(tmp.p, tb, bid[0].as<uint32_t>(), bid[1].as<uint32_t>(), idx[0].as<uint32_t>(),
                                                 idx[1].as<uint32_t>(), (int)n32, 0, end_bit, st));
            k_dict_fill<KW><<<nblocks(rem, 256), 256, 0, st>>>(key_lo[src].as<uint64_t>(), key_hi[src].as<uint64_t>(),
                                                                val[src].as<uint64_t>(), n32, bid[1].as<uint32_t>(),
                                                                idx[1].as<uint32_t>(), ix->d, total_bkt, ix->buckets.as<uint64_t>(),
                                                                passed.as<uint8_t>());
            CUB_(cudaGetLastError());
            // the keys passed on, in order
            cub::CountingInputIterator<uint32_t> count_it(0);
            CUB_(cub::DeviceSelect::Flagged(nullptr, tb, count_it, passed.as<uint8_t>(), sel.as<uint32_t>(), nsel.as<uint32_t>(), (int)n32, st));
            RC_(tmp.ensure(tb));
            CUB_(cub::DeviceSelect::Flagged(tmp.p, tb, count_it, passed.as<uint8_t>(), sel.as<uint32_t>(), nsel.as<uint32_t>(), (int)n32, st));
            uint32_t next = 0;
            CUB_(cudaMemcpyAsync(&next, nsel.p, 4, cudaMemcpyDeviceToHost, st));
            CUB_(cudaStreamSynchronize(st));
            lvl_nbkt.push_back(nbkt);
            lvl_base.push_back(total_bkt);
            total_bkt += nbkt;
            if (next) {
                const int dst = src == 1 ? 2 : 1;
                RC_(key_lo[dst].ensure((uint64_t)next * 8 + 8));
                if (KW == 2) RC_(key_hi[dst].ensure((uint64_t)next * 8 + 8));
                RC_(val[dst].ensure((uint64_t)next * 8 + 8));
                k_dict_gather<<<nblocks(next, 256), 256, 0, st>>>(key_lo[src].as<uint64_t>(), KW == 2 ? key_hi[src].as<uint64_t>() : nullptr,
                                                                  val[src].as<uint64_t>(), sel.as<uint32_t>(), next,
                                                                  key_lo[dst].as<uint64_t>(), key_hi[dst].as<uint64_t>(),
                                                                  val[dst].as<uint64_t>());
                CUB_(cudaGetLastError());
                src = dst;
            }
            rem = next;
        }
    }
    ix->d.dict.buckets = ix->buckets.as<uint64_t>();
    ix->d.dict.n_levels = (uint32_t)lvl_nbkt.size();
    for (size_t i = 0; i < lvl_nbkt.size(); i++) {
        ix->d.dict.level_nbkt[i] = lvl_nbkt[i];
        ix->d.dict.level_base[i] = lvl_base[i];
    }
    ix->info.dict_levels = (uint32_t)lvl_nbkt.size();
    ix->info.dict_bytes = total_bkt * 32;

    // 3. self-check (every key resolves to itself), then the successor / predecessor tables
    if (n_kmers)
        k_dict_check<KW><<<nblocks(n_kmers, 256), 256, 0, st>>>(key_lo[0].as<uint64_t>(), key_hi[0].as<uint64_t>(),
                                                                 val[0].as<uint64_t>(), n_kmers, ix->d, err.as<uint32_t>());
    if (d->n_nodes)
        k_build_edges<KW><<<nblocks(d->n_nodes, 128), 128, 0, st>>>(ix->d, ix->nodes.as<NodeRec>(), ix->nodes_cold.as<NodeCold>(),
                                                                   err.as<uint32_t>());
    CUB_(cudaGetLastError());
    uint32_t e = 0;
    CUB_(cudaMemcpyAsync(&e, err.p, 4, cudaMemcpyDeviceToHost, st));
    CUB_(cudaStreamSynchronize(st));
    cleanup();
    if (e & 2u) return fail(PSA_ERR_INDEX, "dictionary self-check failed (internal)");
    if (e & 4u) return fail(PSA_ERR_INDEX, "missing link: an extension bit has no neighbouring node");
    return PSA_OK;
#undef CUB_
#undef RC_
}

extern "C" int psa_index_create(const psa_index_desc* d, int device, double gamma, psa_index** out) {
    if (!d || !out) return fail(PSA_ERR_ARG, "null argument");
    *out = nullptr;
    if (d->k < 2 || d->k > 64) return fail(PSA_ERR_ARG, "k must be in 2..64");
    if (d->n_nodes >= 0xFFFFFFFFull) return fail(PSA_ERR_ARG, "too many nodes (node ids are u32)");
    if (d->n_nodes && (!d->seq_words || !d->node_start || !d->node_len || !d->node_exts || !d->node_eq))
        return fail(PSA_ERR_ARG, "null node array");
    if (!d->eq_offsets || (d->n_eq && d->eq_offsets[d->n_eq] && !d->eq_members)) return fail(PSA_ERR_ARG, "null eq array");
    if (gamma <= 0) gamma = 1.7;  // ref src/build_index.rs:197
    if (gamma < 1.0) return fail(PSA_ERR_ARG, "gamma must be >= 1");
    CU(cudaSetDevice(device));

    // host-side pass over the node table: k-mer counts, value packing widths, validation
    std::vector<uint64_t> koff(d->n_nodes + 1);
    uint64_t n_kmers = 0, max_off = 0, max_pos = 0;
    for (uint64_t i = 0; i < d->n_nodes; i++) {
        koff[i] = n_kmers;
        if (d->node_len[i] < d->k) return fail(PSA_ERR_INDEX, "node shorter than k");
        if (d->node_len[i] > kMaxNodeLen) return fail(PSA_ERR_INDEX, "node longer than 2^24 - 1 bases");
        if (d->node_start[i] > kStartMask) return fail(PSA_ERR_INDEX, "sequence longer than 2^40 bases");
        if (d->node_eq[i] >= d->n_eq) return fail(PSA_ERR_INDEX, "node eq id out of range");
        if ((d->node_start[i] + d->node_len[i] + 31) / 32 > d->n_seq_words)
            return fail(PSA_ERR_INDEX, "node sequence outside seq_words");
        uint64_t nk = (uint64_t)d->node_len[i] - d->k + 1;
        max_off = std::max(max_off, nk - 1);
        max_pos = std::max(max_pos, d->node_start[i] + nk - 1);
        n_kmers += nk;
    }
    koff[d->n_nodes] = n_kmers;
    uint32_t max_class = 0;
    for (uint64_t c = 0; c < d->n_eq; c++) {
        if (d->eq_offsets[c + 1] < d->eq_offsets[c]) return fail(PSA_ERR_INDEX, "eq_offsets not monotone");
        uint64_t l = d->eq_offsets[c + 1] - d->eq_offsets[c];
        if (l > 0xFFFFFFFFull) return fail(PSA_ERR_INDEX, "class too large");
        max_class = std::max<uint32_t>(max_class, (uint32_t)l);
    }

    psa_index* ix = new (std::nothrow) psa_index();
    if (!ix) return fail(PSA_ERR_NOMEM, "out of memory");
    ix->device = device;
    ix->kw = d->k <= 32 ? 1 : 2;
    cudaDeviceGetAttribute(&ix->sms, cudaDevAttrMultiProcessorCount, device);
    if (ix->sms < 1) ix->sms = 148;
    auto bail = [&](int rc) {
        psa_index_destroy(ix);
        return rc;
    };
    int rc;
    const uint64_t n_mem = d->n_eq ? d->eq_offsets[d->n_eq] : 0;
    DevBuf node_start, node_len, node_exts, node_eq, err;
    auto rel = [&]() { node_start.release(); node_len.release(); node_exts.release(); node_eq.release(); err.release(); };
#define RCI(call)            \
    do {                     \
        rc = (call);         \
        if (rc != PSA_OK) {  \
            rel();           \
            return bail(rc); \
        }                    \
    } while (0)
#define CUI(call)                                                                                          \
    do {                                                                                                   \
        cudaError_t e_ = (call);                                                                           \
        if (e_ != cudaSuccess) {                                                                           \
            (void)cudaGetLastError();                                                                      \
            rel();                                                                                         \
            return bail(fail(PSA_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)));           \
        }                                                                                                  \
    } while (0)
    // (the unitig sequence is padded: the kernels fetch it four words at a time)
    RCI(ix->nodes.ensure((d->n_nodes + 1) * sizeof(NodeRec)));
    RCI(ix->nodes_cold.ensure((d->n_nodes + 1) * sizeof(NodeCold)));
    RCI(ix->seq.ensure((d->n_seq_words + 8) * 8));
    RCI(ix->class_win.ensure((d->n_eq + 1) * sizeof(ClassWin)));
    RCI(ix->eq_off.ensure((d->n_eq + 1) * 8));
    CUI(cudaMemset(ix->seq.p, 0, (d->n_seq_words + 8) * 8));
    if (d->n_seq_words) CUI(cudaMemcpy(ix->seq.p, d->seq_words, d->n_seq_words * 8, cudaMemcpyHostToDevice));
    CUI(cudaMemcpy(ix->eq_off.p, d->eq_offsets, (d->n_eq + 1) * 8, cudaMemcpyHostToDevice));
    RCI(ix->eq_mem.ensure(n_mem * 4 + 4));
    if (n_mem) CUI(cudaMemcpy(ix->eq_mem.p, d->eq_members, n_mem * 4, cudaMemcpyHostToDevice));
    RCI(node_start.ensure(d->n_nodes * 8 + 8));
    RCI(node_len.ensure(d->n_nodes * 4 + 4));
    RCI(node_exts.ensure(d->n_nodes + 4));
    RCI(node_eq.ensure(d->n_nodes * 4 + 4));
    RCI(err.ensure(4));
    CUI(cudaMemset(err.p, 0, 4));
    if (d->n_nodes) {
        CUI(cudaMemcpy(node_start.p, d->node_start, d->n_nodes * 8, cudaMemcpyHostToDevice));
        CUI(cudaMemcpy(node_len.p, d->node_len, d->n_nodes * 4, cudaMemcpyHostToDevice));
        CUI(cudaMemcpy(node_exts.p, d->node_exts, d->n_nodes, cudaMemcpyHostToDevice));
        CUI(cudaMemcpy(node_eq.p, d->node_eq, d->n_nodes * 4, cudaMemcpyHostToDevice));
    }

    DevIndex& D = ix->d;
    D.k = d->k;
    D.node_bits = bits_for(d->n_nodes ? d->n_nodes - 1 : 0);
    D.pos_bits = bits_for(max_pos + 1);  // (so that no entry is all ones: that is the empty slot)
    // a dictionary entry is node | pos | fingerprint in 63 bits; fewer than 8 fingerprint bits would let
    // absent k-mers through to the verification too often
    const int fp = 63 - (int)D.node_bits - (int)D.pos_bits;
    if (fp < 8) {
        rel();
        return bail(fail(PSA_ERR_INDEX, "graph too large for 64-bit dictionary entries (node bits + position bits > 55)"));
    }
    if (n_kmers >= 0xFFFFFFFFull) {
        rel();
        return bail(fail(PSA_ERR_INDEX, "too many k-mers (the dictionary build indexes them with 32 bits)"));
    }
    D.fp_bits = (uint32_t)std::min(fp, 32);
    D.n_nodes = d->n_nodes;
    D.n_kmers = n_kmers;
    D.n_eq = d->n_eq;
    D.nodes = ix->nodes.as<NodeRec>();
    D.nodes_cold = ix->nodes_cold.as<NodeCold>();
    D.seq = ix->seq.as<uint64_t>();
    D.eq_off = ix->eq_off.as<uint64_t>();
    D.eq_mem = ix->eq_mem.as<uint32_t>();
    D.class_win = ix->class_win.as<ClassWin>();

    cudaEvent_t e0, e1;
    CUI(cudaEventCreate(&e0));
    CUI(cudaEventCreate(&e1));
    CUI(cudaEventRecord(e0, 0));
    if (d->n_nodes)
        k_node_basics<<<nblocks(d->n_nodes, 256), 256>>>(ix->nodes.as<NodeRec>(), ix->nodes_cold.as<NodeCold>(), d->n_nodes, node_start.as<uint64_t>(),
                                                         node_len.as<uint32_t>(), node_exts.as<uint8_t>(),
                                                         node_eq.as<uint32_t>(), ix->eq_off.as<uint64_t>(), d->n_eq,
                                                         d->k, err.as<uint32_t>());
    if (d->n_eq)
        k_build_class_win<<<nblocks(d->n_eq, 256), 256>>>(ix->eq_off.as<uint64_t>(), ix->eq_mem.as<uint32_t>(), d->n_eq,
                                                          ix->class_win.as<ClassWin>());
    if (d->n_nodes)
        k_node_windows<<<nblocks(d->n_nodes, 256), 256>>>(ix->nodes.as<NodeRec>(), d->n_nodes, d->n_eq, ix->class_win.as<ClassWin>());
    CUI(cudaGetLastError());
    {
        uint32_t e = 0;
        CUI(cudaMemcpy(&e, err.p, 4, cudaMemcpyDeviceToHost));
        if (e) {
            rel();
            return bail(fail(PSA_ERR_INDEX, "invalid node table"));
        }
    }
    if (ix->kw == 1) RCI(build_on_device<1>(ix, d, gamma, node_start, koff, n_kmers));
    else RCI(build_on_device<2>(ix, d, gamma, node_start, koff, n_kmers));
    CUI(cudaEventRecord(e1, 0));
    CUI(cudaEventSynchronize(e1));
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    rel();
#undef RCI
#undef CUI

    ix->h_eq_off.assign(d->eq_offsets, d->eq_offsets + d->n_eq + 1);
    ix->h_eq_mem.assign(d->eq_members, d->eq_members + n_mem);
    psa_index_info& I = ix->info;
    I.k = d->k;
    I.n_nodes = d->n_nodes;
    I.n_kmers = n_kmers;
    I.n_eq = d->n_eq;
    I.n_eq_members = n_mem;
    I.n_seq_words = d->n_seq_words;
    I.node_bytes = d->n_nodes * (sizeof(NodeRec) + sizeof(NodeCold));
    I.seq_bytes = d->n_seq_words * 8;
    I.eq_bytes = (d->n_eq + 1) * 8 + n_mem * 4 + d->n_eq * sizeof(ClassWin);
    I.node_bits = D.node_bits;
    I.pos_bits = D.pos_bits;
    I.fp_bits = D.fp_bits;
    I.max_class_len = max_class;
    I.gamma = gamma;
    I.build_ms = ms;
    *out = ix;
    return PSA_OK;
}

struct psa_fq_lane;
static void fq_lane_free(psa_fq_lane* l);
extern "C" void psa_index_destroy(psa_index* ix) {
    if (!ix) return;
    cudaSetDevice(ix->device);
    for (auto l : ix->idle_lanes) fq_lane_free(l);
    ix->idle_lanes.clear();
    ix->buckets.release(); ix->nodes.release(); ix->nodes_cold.release();
    ix->seq.release(); ix->eq_off.release(); ix->eq_mem.release(); ix->class_win.release(); ix->bloom.release();
    delete ix;
}

// eq_classes as the index holds them on the host (borrowed until psa_index_destroy); used by psa_process_reads
extern "C" int psa_index_host_classes(const psa_index* ix, const uint64_t** eq_offsets, const uint32_t** eq_members, uint64_t* n_eq) {
    if (!ix || !eq_offsets || !eq_members || !n_eq) return fail(PSA_ERR_ARG, "null argument");
    *eq_offsets = ix->h_eq_off.data();
    *eq_members = ix->h_eq_mem.data();
    *n_eq = ix->d.n_eq;
    return PSA_OK;
}

extern "C" int psa_index_get_info(const psa_index* ix, psa_index_info* out) {
    if (!ix || !out) return fail(PSA_ERR_ARG, "null argument");
    *out = ix->info;
    return PSA_OK;
}

extern "C" int psa_index_lookup(psa_index* ix, const uint64_t* kmer_words, uint64_t n, uint8_t* found,
                                uint32_t* node, uint32_t* off) {
    if (!ix || (n && (!kmer_words || !found || !node || !off))) return fail(PSA_ERR_ARG, "null argument");
    if (!n) return PSA_OK;
    CU(cudaSetDevice(ix->device));
    DevBuf w, f, nn, oo;
    int rc = PSA_OK;
    auto rel = [&]() { w.release(); f.release(); nn.release(); oo.release(); };
    if ((rc = w.ensure(n * 8 * ix->kw + 8)) || (rc = f.ensure(n)) || (rc = nn.ensure(n * 4)) || (rc = oo.ensure(n * 4))) {
        rel();
        return rc;
    }
    cudaError_t e = cudaMemcpy(w.p, kmer_words, n * 8 * ix->kw, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        if (ix->kw == 1) k_lookup<1><<<nblocks(n, 128), 128>>>(ix->d, w.as<uint64_t>(), n, f.as<uint8_t>(), nn.as<uint32_t>(), oo.as<uint32_t>());
        else k_lookup<2><<<nblocks(n, 128), 128>>>(ix->d, w.as<uint64_t>(), n, f.as<uint8_t>(), nn.as<uint32_t>(), oo.as<uint32_t>());
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(found, f.p, n, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(node, nn.p, n * 4, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(off, oo.p, n * 4, cudaMemcpyDeviceToHost);
    rel();
    if (e != cudaSuccess) return fail(PSA_ERR_CUDA, cudaGetErrorString(e));
    return PSA_OK;
}

// ---------------------------------------------------------------------------------------------
// mapper
// ---------------------------------------------------------------------------------------------
struct Slot {  // staging of one pipeline chunk
    DevBuf in_data, in_off, in_len;     // host batch copied as is
    DevBuf hits, hits_c, tx, meta_dev;  // results of the chunk (hits_c: compact records); meta_dev = {running total, status}
    cudaEvent_t in_ready = nullptr, in_free = nullptr, comp_done = nullptr, meta_done = nullptr, out_free = nullptr;
    bool in_free_rec = false, out_free_rec = false;
    unsigned long long* meta_host = nullptr;  // pinned: [0] running tx total after the chunk, [1] status
};

// Host batches: chunk c uses slot c % kSlots; its host-side completion (reading its totals, enqueueing
// the D2H of its members) happens kSlots - 1 chunks later, so that the host thread keeps submitting
// ahead of the GPU instead of waiting for each chunk in turn.
constexpr int kSlots = 4;

struct psa_mapper {
    psa_index* ix = nullptr;
    uint64_t chunk_reads = 0;
    uint32_t allowed = PSA_DEFAULT_ALLOWED_MISMATCHES;
    cudaStream_t st = nullptr, st_h2d = nullptr, st_d2h = nullptr, st_aux = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;   // k_map's first launch runs on st_aux beside k_seed_scan + second pass
    bool overlap_coop = true;   // PSA_OVERLAP_COOP=0: one k_map launch after everything else, on the mapper's stream
    DevBuf counts, counts_backup, status, novel_cursor, events, novel, spill, pool, running;
    DevBuf hits_full;   // compact results on device batches: the kernels' 24-byte working records
    DevBuf ntab, ntab_backup, ncur_backup, npool, ncur, nlist, nslot;   // the novel-set table (NovelTable), this batch's list of novel reads and their entries
    uint64_t ntab_cap = 1ull << 21, npool_cap = 1ull << 23;   // 64 MB + 32 MB to start with; both grow on demand
    DevBuf words, woff, nwords, dst_off, scan_tmp, meta, deferred, scan_list, seeded, seeded_ev;
    uint64_t novel_cap = 0;
    uint32_t spill_cap = 56;          // visited-class list entries per group beyond its lanes
    uint64_t pool_cap = 1ull << 18;   // entries (uint4) of the shared overflow pool; grows on demand
    uint32_t group = 8;  // lanes cooperating on one read (8, 16 or 32)
    uint32_t fast_probes = 10;  // 0: every read goes to the cooperative kernel; default set from k at creation
    uint32_t fast_max_small = 32;
    uint32_t reseed_first = 2;  // re-seed positions a thread of the FIRST pass tries (PSA_RESEED_FIRST) before it leaves the read to the second pass,
                                // which allows max(fast_probes, 8): 2.36 vs 2.39 ms per batch on B200 (profiles/r2_exp_walk_kernel.md)
    bool tile_pack = true;      // PSA_TILE_PACK=0: pack fixed-stride ASCII without the shared-memory tiles
    uint32_t scan_width = 8;    // lanes per read of k_seed_scan (0: long first searches go to k_map)
    int grid = 0;
    Slot slot[kSlots];
    uint64_t launches = 0;
    // map-kernel timing (psa_mapper_profile_*)
    bool profiling = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_events[3];  // [0] k_map_thread, [1] k_map, [2] k_seed_scan
    // pending async call
    psa_result_batch* pending = nullptr;
    unsigned long long* pin = nullptr;  // pinned scratch: [0] tx total, [1] status
};

// the thread-per-read kernel
static uint32_t tile_smem_bytes(const ReadsView& rv) {  // shared memory of the TILE variant, 0 = not eligible
    if (rv.woff || rv.len || !rv.wstride) return 0;
    const uint64_t total = 16 + (uint64_t)kThreadBlock * rv.wstride * 8;
    return total > 12 * 1024 ? 0 : (uint32_t)total;  // longer reads keep the L1 for the index instead
}
template <bool EV>
static void launch_map_thread(psa_mapper* m, cudaStream_t st, const MapParams& p, bool hint) {
    if (hint) {  // second pass over the reads k_seed_scan seeded: persistent warps, the list length is on the device
        const unsigned grid = (unsigned)std::min<uint64_t>(nblocks(p.reads.n, kThreadBlock), 148 * PSA_THREAD_MIN_BLOCKS);
        if (m->ix->kw == 1) k_map_thread<1, EV, true><<<grid, kThreadBlock, 0, st>>>(m->ix->d, p);
        else k_map_thread<2, EV, true><<<grid, kThreadBlock, 0, st>>>(m->ix->d, p);
        return;
    }
    const unsigned grid = nblocks(p.reads.n, kThreadBlock);
    const uint32_t smem = tile_smem_bytes(p.reads);
    if (smem) {
        if (m->ix->kw == 1) k_map_thread<1, EV, false, true><<<grid, kThreadBlock, smem, st>>>(m->ix->d, p);
        else k_map_thread<2, EV, false, true><<<grid, kThreadBlock, smem, st>>>(m->ix->d, p);
    } else {
        if (m->ix->kw == 1) k_map_thread<1, EV, false><<<grid, kThreadBlock, 0, st>>>(m->ix->d, p);
        else k_map_thread<2, EV, false><<<grid, kThreadBlock, 0, st>>>(m->ix->d, p);
    }
}
template <bool EV>
static void launch_seed_scan(psa_mapper* m, cudaStream_t st, const MapParams& p) {
    const DevIndex& d = m->ix->d;
    const unsigned grid = (unsigned)std::min<uint64_t>(nblocks(p.reads.n * m->scan_width, 256), 148 * 4);
#define PSA_SCAN(KW, G) k_seed_scan<KW, EV, G><<<grid, 256, 0, st>>>(d, p)
    if (m->ix->kw == 1) {
        if (m->scan_width == 8) PSA_SCAN(1, 8);
        else if (m->scan_width == 16) PSA_SCAN(1, 16);
        else PSA_SCAN(1, 32);
    } else {
        if (m->scan_width == 8) PSA_SCAN(2, 8);
        else if (m->scan_width == 16) PSA_SCAN(2, 16);
        else PSA_SCAN(2, 32);
    }
#undef PSA_SCAN
}

template <bool EV>
static void launch_map(psa_mapper* m, int grid, cudaStream_t st, const MapParams& p) {
    const DevIndex& d = m->ix->d;
#define PSA_LAUNCH(KW, G) k_map<KW, EV, G><<<grid, 256, 0, st>>>(d, p)
    if (m->ix->kw == 1) {
        if (m->group == 8) PSA_LAUNCH(1, 8);
        else if (m->group == 16) PSA_LAUNCH(1, 16);
        else PSA_LAUNCH(1, 32);
    } else {
        if (m->group == 8) PSA_LAUNCH(2, 8);
        else if (m->group == 16) PSA_LAUNCH(2, 16);
        else PSA_LAUNCH(2, 32);
    }
#undef PSA_LAUNCH
}

static int mapper_grid(psa_mapper* m) {
    if (m->grid) return m->grid;
    int sms = 148, per_sm = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, m->ix->device);
#define PSA_OCC(KW, G) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_map<KW, false, G>, 256, 0)
    if (m->ix->kw == 1) {
        if (m->group == 8) PSA_OCC(1, 8);
        else if (m->group == 16) PSA_OCC(1, 16);
        else PSA_OCC(1, 32);
    } else {
        if (m->group == 8) PSA_OCC(2, 8);
        else if (m->group == 16) PSA_OCC(2, 16);
        else PSA_OCC(2, 32);
    }
#undef PSA_OCC
    if (per_sm < 1) per_sm = 1;
    m->grid = sms * per_sm;
    return m->grid;
}

static int mapper_alloc_spill(psa_mapper* m) {
    const int grid = mapper_grid(m);
    int rc = m->spill.ensure((size_t)grid * (256 / m->group) * m->spill_cap * sizeof(uint4));
    if (rc) return rc;
    return m->pool.ensure(m->pool_cap * sizeof(uint4));
}

extern "C" int psa_mapper_create(psa_index* ix, uint64_t chunk_reads, psa_mapper** out) {
    if (!ix || !out) return fail(PSA_ERR_ARG, "null argument");
    *out = nullptr;
    CU(cudaSetDevice(ix->device));
    psa_mapper* m = new (std::nothrow) psa_mapper();
    if (!m) return fail(PSA_ERR_NOMEM, "out of memory");
    m->ix = ix;
    m->chunk_reads = chunk_reads ? chunk_reads : (1ull << 19);  // 512 Ki: best H2D|kernel|D2H overlap measured on B200
    // initial size of the novel-set table (it grows on demand; the tests start it tiny)
    if (const char* e = getenv("PSA_NOVEL_TABLE_CAP")) {
        uint64_t c = 16;
        while (c < (uint64_t)std::max(16, atoi(e))) c <<= 1;
        m->ntab_cap = c;
        m->npool_cap = 4 * c;
    }
    int rc = PSA_OK;
    cudaError_t e = cudaStreamCreateWithFlags(&m->st, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&m->st_h2d, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&m->st_d2h, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&m->st_aux, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&m->ev_fork, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&m->ev_join, cudaEventDisableTiming);
    for (int s = 0; s < kSlots && e == cudaSuccess; s++) {
        Slot& S = m->slot[s];
        cudaEvent_t* evs[5] = {&S.in_ready, &S.in_free, &S.comp_done, &S.meta_done, &S.out_free};
        for (auto ev : evs)
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(ev, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaHostAlloc((void**)&S.meta_host, 64, cudaHostAllocDefault);
    }
    if (e == cudaSuccess) e = cudaHostAlloc((void**)&m->pin, 64, cudaHostAllocDefault);
    if (e != cudaSuccess) {
        psa_mapper_destroy(m);
        return fail(PSA_ERR_CUDA, cudaGetErrorString(e));
    }
    const uint64_t nc = ix->d.n_eq + 2;
    if ((rc = m->counts.ensure(nc * 8)) || (rc = m->counts_backup.ensure(nc * 8)) || (rc = m->status.ensure(8)) ||
        (rc = m->novel_cursor.ensure(128)) || (rc = m->ntab.ensure(m->ntab_cap * sizeof(NovelEntry))) ||
        (rc = m->npool.ensure(m->npool_cap * 4)) || (rc = m->ncur.ensure(16)) || (rc = m->events.ensure(40 * 8)) || (rc = m->running.ensure(16)) || (rc = m->meta.ensure(16)) ||
        (rc = m->slot[0].meta_dev.ensure(16)) || (rc = m->slot[1].meta_dev.ensure(16)) ||
        (rc = m->slot[2].meta_dev.ensure(16)) || (rc = m->slot[3].meta_dev.ensure(16))) {
        psa_mapper_destroy(m);
        return rc;
    }
    cudaMemset(m->counts.p, 0, nc * 8);
    cudaMemset(m->ntab.p, 0, m->ntab_cap * sizeof(NovelEntry));
    cudaMemset(m->ncur.p, 0, 16);
    cudaMemset(m->events.p, 0, 40 * 8);
    cudaMemset(m->status.p, 0, 8);
    if (const char* e = getenv("PSA_GROUP_WIDTH")) {
        int g = atoi(e);
        if (g == 8 || g == 16 || g == 32) m->group = (uint32_t)g;
    }
    // one substitution knocks out the seeds at positions (e-k, e]: ceil(k/3) stride-3 positions; two more and a
    // read with a single error in its head is seeded by its own thread (measured best on B200)
    m->fast_probes = (ix->d.k + 2) / 3 + 2;
    // ... unless k_seed_scan has its filter: then a search that misses three times is cheaper there (eight lanes ask the
    // L2-resident filter, one dictionary probe for the seed) than in a thread that probes alone: 3.62 vs 3.65 ms per batch
    if (ix->d.bloom) m->fast_probes = 3;
    if (const char* e = getenv("PSA_FAST_PROBES")) m->fast_probes = (uint32_t)std::max(0, atoi(e));
    if (const char* e = getenv("PSA_FAST_MAX_SMALL")) m->fast_max_small = (uint32_t)std::max(0, atoi(e));
    if (const char* e = getenv("PSA_OVERLAP_COOP")) m->overlap_coop = atoi(e) != 0;
    if (const char* e = getenv("PSA_RESEED_FIRST")) m->reseed_first = (uint32_t)std::max(0, atoi(e));
    if (const char* e = getenv("PSA_TILE_PACK")) m->tile_pack = atoi(e) != 0;
    if (const char* e = getenv("PSA_SCAN_WIDTH")) {
        int g = atoi(e);
        if (g == 0 || g == 8 || g == 16 || g == 32) m->scan_width = (uint32_t)g;
    }
    if ((rc = mapper_alloc_spill(m))) {
        psa_mapper_destroy(m);
        return rc;
    }
    *out = m;
    return PSA_OK;
}

extern "C" void psa_mapper_destroy(psa_mapper* m) {
    if (!m) return;
    cudaSetDevice(m->ix->device);
    if (m->st) cudaStreamSynchronize(m->st);
    if (m->st_h2d) cudaStreamSynchronize(m->st_h2d);
    if (m->st_d2h) cudaStreamSynchronize(m->st_d2h);
    if (m->st_aux) cudaStreamSynchronize(m->st_aux);
    DevBuf* bufs[] = {&m->counts, &m->counts_backup, &m->status, &m->novel_cursor, &m->events, &m->novel, &m->spill, &m->pool,
                      &m->running, &m->hits_full, &m->ntab, &m->ntab_backup, &m->ncur_backup, &m->npool, &m->ncur, &m->nlist, &m->nslot, &m->words, &m->woff, &m->nwords, &m->dst_off, &m->scan_tmp, &m->meta, &m->deferred, &m->scan_list, &m->seeded, &m->seeded_ev};
    for (auto b : bufs) b->release();
    for (int s = 0; s < kSlots; s++) {
        Slot& S = m->slot[s];
        S.in_data.release(); S.in_off.release(); S.in_len.release(); S.hits.release(); S.hits_c.release(); S.tx.release(); S.meta_dev.release();
        cudaEvent_t evs[5] = {S.in_ready, S.in_free, S.comp_done, S.meta_done, S.out_free};
        for (auto ev : evs)
            if (ev) cudaEventDestroy(ev);
        if (S.meta_host) cudaFreeHost(S.meta_host);
    }
    if (m->pin) cudaFreeHost(m->pin);
    if (m->st) cudaStreamDestroy(m->st);
    if (m->st_h2d) cudaStreamDestroy(m->st_h2d);
    if (m->st_d2h) cudaStreamDestroy(m->st_d2h);
    if (m->st_aux) cudaStreamDestroy(m->st_aux);
    if (m->ev_fork) cudaEventDestroy(m->ev_fork);
    if (m->ev_join) cudaEventDestroy(m->ev_join);
    delete m;
}

extern "C" int psa_mapper_set_allowed_mismatches(psa_mapper* m, uint32_t allowed) {
    if (!m) return fail(PSA_ERR_ARG, "null argument");
    m->allowed = allowed;
    return PSA_OK;
}
extern "C" int psa_mapper_set_group_width(psa_mapper* m, uint32_t lanes) {
    if (!m || (lanes != 8 && lanes != 16 && lanes != 32)) return fail(PSA_ERR_ARG, "group width must be 8, 16 or 32");
    CU(cudaSetDevice(m->ix->device));
    CU(cudaStreamSynchronize(m->st));
    m->group = lanes;
    m->grid = 0;
    return mapper_alloc_spill(m);
}
extern "C" int psa_mapper_set_fast_path(psa_mapper* m, uint32_t max_probes, uint32_t max_small) {
    if (!m) return fail(PSA_ERR_ARG, "null argument");
    m->fast_probes = max_probes;
    m->fast_max_small = max_small;
    return PSA_OK;
}
extern "C" int psa_mapper_set_scan_width(psa_mapper* m, uint32_t lanes) {
    if (!m || (lanes != 0 && lanes != 8 && lanes != 16 && lanes != 32))
        return fail(PSA_ERR_ARG, "scan width must be 0, 8, 16 or 32");
    m->scan_width = lanes;
    return PSA_OK;
}
extern "C" void* psa_mapper_stream(psa_mapper* m) { return m ? (void*)m->st : nullptr; }
extern "C" void* psa_mapper_counts_device(psa_mapper* m) { return m ? m->counts.p : nullptr; }
extern "C" uint64_t psa_mapper_launch_count(const psa_mapper* m) { return m ? m->launches : 0; }

// Enqueue the kernels for one device-resident batch on m->st.
//   reads: device pointers.  hits/tx_buf: device.  tx_base_dev: device u64 holding the offset at
//   which this batch's members start in tx_buf (nullptr = 0); it is advanced by the batch total.
//   total_dev: device u64[2] receiving {running total after the batch, status}.
struct DeviceBatch {
    const psa_read_batch* reads;
    HitRec* hits;
    uint32_t* tx_buf;
    uint64_t tx_cap;
    uint64_t* meta_out;    // device u64[2]: {running total after the batch, status}
    uint64_t total_words;  // ragged ASCII only: sum of ceil(len/32) if the caller knows it, else 0
    bool sticky = false;   // accumulate the status word over the batches queued since the last psa_mapper_sync
    HitCompact* hits_c = nullptr;  // compact results: where the 8-byte records go (hits stays the kernels' working array)
};

template <bool EV>
static int enqueue_device_batch(psa_mapper* m, const DeviceBatch& b, bool want_counts) {
    const psa_read_batch* r = b.reads;
    const uint64_t n = r->n_reads;
    psa_index* ix = m->ix;
    cudaStream_t st = m->st;
    ReadsView rv{};
    rv.n = n;
    rv.len = r->read_len;
    rv.fixed_len = r->fixed_len;
    int rc;
    if (r->format == PSA_READS_PACKED) {
        rv.words = (const uint64_t*)r->data;
        rv.woff = r->read_off;
        rv.wstride = r->stride;
    } else {
        // ASCII -> DnaString words (ref src/pseudoaligner.rs:449-450), in HBM
        if (r->read_len) {
            if ((rc = m->nwords.ensure((n + 2) * 8)) || (rc = m->woff.ensure((n + 2) * 8))) return rc;
            k_words_per_read<<<nblocks(n + 1, 256), 256, 0, st>>>(r->read_len, n, m->nwords.as<uint64_t>());
            m->launches++;
            size_t tb = 0;
            CU(cub::DeviceScan::ExclusiveSum(nullptr, tb, m->nwords.as<uint64_t>(), m->woff.as<uint64_t>(), n + 1, st));
            if ((rc = m->scan_tmp.ensure(tb))) return rc;
            CU(cub::DeviceScan::ExclusiveSum(m->scan_tmp.p, tb, m->nwords.as<uint64_t>(), m->woff.as<uint64_t>(), n + 1, st));
            uint64_t total_words = b.total_words;
            if (!total_words) {  // unknown: read it back (one small sync per ragged ASCII device batch)
                CU(cudaMemcpyAsync(m->pin + 2, m->woff.as<uint64_t>() + n, 8, cudaMemcpyDeviceToHost, st));
                CU(cudaStreamSynchronize(st));
                total_words = m->pin[2];
            }
            if ((rc = m->words.ensure(total_words * 8 + 16))) return rc;
            rv.woff = m->woff.as<uint64_t>();
        } else {
            rv.woff = nullptr;
            rv.wstride = ((uint64_t)r->fixed_len + 31) / 32;
            if ((rc = m->words.ensure(n * rv.wstride * 8 + 16))) return rc;
        }
        // tile form (TMA): stride and base 16-byte friendly, tile small enough for several CTAs per SM
        const uint64_t tile_bytes = (uint64_t)kPackTileReads * r->stride;
        if (!r->read_len && !r->read_off && m->tile_pack && r->fixed_len && r->stride >= r->fixed_len && tile_bytes % 16 == 0 &&
            ((uintptr_t)r->data & 15) == 0 && tile_bytes <= 40 * 1024) {
            const uint32_t smem = (uint32_t)(16 + tile_bytes + 48 + 16);
            k_pack_ascii_tile<<<nblocks(n, kPackTileReads), kPackTileReads, smem, st>>>(
                (const uint8_t*)r->data, (uint32_t)r->stride, r->fixed_len, n, m->words.as<uint64_t>(), (uint32_t)tile_bytes);
        } else if (!r->read_len && !r->read_off && rv.wstride == 0) {
            // fixed_len == 0: nothing to pack (every read is None, ref src/pseudoaligner.rs:82-84)
        } else if (!r->read_len && !r->read_off) {
            // 32-bit word indices inside the kernel: slices of at most 2^31 words
            const uint64_t per = std::max<uint64_t>(1, (1ull << 31) / rv.wstride);
            for (uint64_t r0 = 0; r0 < n; r0 += per) {
                const uint64_t nr = std::min(per, n - r0);
                k_pack_ascii_fixed<<<(unsigned)std::min<uint64_t>(nblocks(nr * rv.wstride, 256), 148 * 32), 256, 0, st>>>(
                    (const uint8_t*)r->data + r0 * r->stride, r->stride, r->fixed_len, nr, m->words.as<uint64_t>() + r0 * rv.wstride);
            }
        }
        else
            k_pack_ascii<<<(unsigned)std::min<uint64_t>(nblocks(n * 32, 256), 148 * 64), 256, 0, st>>>(
                (const uint8_t*)r->data, r->read_off, r->stride, r->read_len, r->fixed_len, rv.woff, rv.wstride, n,
                m->words.as<uint64_t>());
        m->launches++;
        CU(cudaGetLastError());
        rv.words = m->words.as<uint64_t>();
    }
    if (!m->novel_cap) {
        m->novel_cap = std::max<uint64_t>(1 << 20, 32 * std::min<uint64_t>(n, 1 << 22));
        if ((rc = m->novel.ensure(m->novel_cap * 4))) return rc;
    }
    if ((rc = m->nlist.ensure((n + 1) * 4)) || (rc = m->nslot.ensure((n + 1) * 4))) return rc;
    CU(cudaMemsetAsync(m->novel_cursor.p, 0, 128, st));  // [0] novel, [1] pool, [2] deferred, [3] to scan, [4] seeded, [5] k_map's claim counter, [6] second pass's claim counter, [7] unused, [8] novel reads listed, [9] list length when k_map's first launch began, [10] claim counter of its second launch
    CU(cudaMemsetAsync(m->status.p, 0, 4, st));

    MapParams p{};
    p.reads = rv;
    p.hits = b.hits;
    p.counts = want_counts ? m->counts.as<unsigned long long>() : nullptr;
    p.novel = m->novel.as<uint32_t>();   // sets that are no visited class are always materialised: they are counted per set
    p.novel_list = m->nlist.as<uint32_t>();
    p.novel_list_count = m->novel_cursor.as<unsigned long long>() + 8;
    p.novel_cap = m->novel_cap;
    p.novel_cursor = m->novel_cursor.as<unsigned long long>();
    p.spill = m->spill.as<uint4>();
    p.spill_cap = m->spill_cap;
    p.pool = m->pool.as<uint4>();
    p.pool_cap = m->pool_cap;
    p.pool_cursor = m->novel_cursor.as<unsigned long long>() + 1;
    p.allowed_mismatches = m->allowed;
    p.status = m->status.as<uint32_t>();
    p.events = EV ? m->events.as<unsigned long long>() : nullptr;
    const int grid = mapper_grid(m);
    auto timed = [&](int which, auto&& launch, cudaStream_t on = nullptr) -> int {
        if (!on) on = st;
        cudaEvent_t e0 = nullptr, e1 = nullptr;
        if (m->profiling) {
            CU(cudaEventCreate(&e0));
            CU(cudaEventCreate(&e1));
            CU(cudaEventRecord(e0, on));
        }
        launch();
        if (m->profiling) {
            CU(cudaEventRecord(e1, on));
            m->prof_events[which].emplace_back(e0, e1);
        }
        m->launches++;
        CU(cudaGetLastError());
        return PSA_OK;
    };
    if (n) {
        // fast kernel: one thread per read; the reads it gives up go to the cooperative kernel
        if (m->fast_probes && n < 0xFFFFFFFFull) {
            if ((rc = m->deferred.ensure(n * 4))) return rc;
            p.list = m->deferred.as<uint32_t>();
            p.list_count = m->novel_cursor.as<unsigned long long>() + 2;
            p.work_cursor = m->novel_cursor.as<unsigned long long>() + 5;
            p.hint_cursor = m->novel_cursor.as<unsigned long long>() + 6;
            p.max_probes = m->fast_probes;
            p.max_small = m->fast_max_small;
            if (m->scan_width) {
                if ((rc = m->scan_list.ensure(n * 4)) || (rc = m->seeded.ensure(n * sizeof(uint4)))) return rc;
                if (EV && (rc = m->seeded_ev.ensure(n * sizeof(uint4)))) return rc;
                p.scan_list = m->scan_list.as<uint32_t>();
                p.scan_count = m->novel_cursor.as<unsigned long long>() + 3;
                p.seeded = m->seeded.as<uint4>();
                p.seeded_count = m->novel_cursor.as<unsigned long long>() + 4;
                p.seeded_ev = EV ? m->seeded_ev.as<uint4>() : nullptr;
            }
            const uint32_t reseed_long = std::max(m->fast_probes, (uint32_t)kReseedProbes);
            auto fast = [&](bool hint) {
                // first pass: a read that has to search for a new seed mid-way is handed to the second pass (when there
                // is one), so that the other 31 reads of its warp do not wait for the search; second pass: long budget
                p.reseed_probes = (!hint && m->scan_width) ? std::min(m->reseed_first, reseed_long) : reseed_long;
                // (a read the first pass handed over at a re-seed search makes its first search again in the second pass:
                // whatever the first pass's budget was, it finds the seed there)
                p.max_probes = (hint && m->scan_width) ? std::max(m->fast_probes, (m->ix->d.k + 2) / 3 + 2) : m->fast_probes;
                launch_map_thread<EV>(m, st, p, hint);
            };
            if ((rc = timed(0, [&]() { fast(false); }))) return rc;
            if (m->scan_width) {
                // The reads the first pass handed to k_map are few and slow (a latency-bound tail): they are mapped on a
                // second stream while k_seed_scan and the second pass run; what the second pass hands over follows below.
                if (m->overlap_coop) {
                    unsigned long long* cur = m->novel_cursor.as<unsigned long long>();
                    CU(cudaEventRecord(m->ev_fork, st));
                    CU(cudaStreamWaitEvent(m->st_aux, m->ev_fork, 0));
                    k_list_snapshot<<<1, 32, 0, m->st_aux>>>(p.list_count, cur + 9);
                    MapParams p1 = p;
                    p1.list_end = cur + 9;
                    if ((rc = timed(1, [&]() { launch_map<EV>(m, grid, m->st_aux, p1); }, m->st_aux))) return rc;
                    CU(cudaEventRecord(m->ev_join, m->st_aux));
                    m->launches++;
                    p.list_first = cur + 9;
                    p.work_cursor = cur + 10;
                }
                if ((rc = timed(2, [&]() { launch_seed_scan<EV>(m, st, p); }))) return rc;
                if ((rc = timed(0, [&]() { fast(true); }))) return rc;
                if (m->overlap_coop) CU(cudaStreamWaitEvent(st, m->ev_join, 0));
            }
        }
        if ((rc = timed(1, [&]() { launch_map<EV>(m, grid, st, p); }))) return rc;
    }
    // sets that are no index class: find or claim their entries in the mapper's table (counted below, once the
    // batch is known to stand)
    NovelTable nt{m->ntab.as<NovelEntry>(), m->ntab_cap, m->npool.as<uint32_t>(), m->npool_cap, m->ncur.as<unsigned long long>()};
    // (random accesses to a table in HBM: a full wave of threads hides them; 296 CTAs took 0.25 ms per 8 Mi-read batch)
    const unsigned novel_grid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(nblocks(n, 256), (uint64_t)ix->sms * 8));
    if (n && want_counts) {
        k_novel_claim<<<novel_grid, 256, 0, st>>>(p.novel_list, p.novel_list_count, b.hits, p.novel, nt, m->nslot.as<uint32_t>(), p.status);
        k_novel_verify<<<novel_grid, 256, 0, st>>>(p.novel_list, p.novel_list_count, b.hits, p.novel, nt, m->nslot.as<uint32_t>(), p.status);
        m->launches += 2;
        CU(cudaGetLastError());
    }
    // exclusive scan of n_tx (n+1 items: the last one is the batch total), seeded by the running total
    if ((rc = m->dst_off.ensure((n + 2) * 8))) return rc;
    {
        cub::CountingInputIterator<uint64_t> cnt(0);
        TxLenN f{b.hits, n, b.hits_c != nullptr};
        cub::TransformInputIterator<uint64_t, TxLenN, cub::CountingInputIterator<uint64_t>> in(cnt, f);
        size_t tb = 0;
        CU(cub::DeviceScan::ExclusiveSum(nullptr, tb, in, m->dst_off.as<uint64_t>(), n + 1, st));
        if ((rc = m->scan_tmp.ensure(tb))) return rc;
        CU(cub::DeviceScan::ExclusiveSum(m->scan_tmp.p, tb, in, m->dst_off.as<uint64_t>(), n + 1, st));
    }
    if (n) {
        k_expand_balanced<<<nblocks(n, 256), 256, 0, st>>>(b.hits, n, m->dst_off.as<uint64_t>(), m->running.as<uint64_t>(),
                                                           ix->d.eq_off, ix->d.eq_mem, m->novel.as<uint32_t>(), b.tx_buf, b.tx_cap,
                                                           b.hits_c != nullptr);
        if (b.hits_c) {
            k_compact_hits<<<nblocks(n, 256), 256, 0, st>>>(b.hits, n, b.hits_c, m->status.as<uint32_t>());
            m->launches++;
        }
    }
    if (n && want_counts) {
        k_novel_add<<<novel_grid, 256, 0, st>>>(p.novel_list_count, nt, m->nslot.as<uint32_t>(), p.status, m->dst_off.as<uint64_t>() + n, b.tx_cap,
                                         b.tx_buf != nullptr);
        m->launches++;
    }
    k_advance<<<1, 32, 0, st>>>(m->running.as<uint64_t>(), m->dst_off.as<uint64_t>() + n, b.tx_cap, b.tx_buf != nullptr,
                                m->status.as<uint32_t>(), b.meta_out, b.sticky ? m->status.as<uint32_t>() + 1 : nullptr);
    m->launches++;
    m->launches++;
    CU(cudaGetLastError());
    return PSA_OK;
}

static int check_batch_args(const psa_read_batch* r, const psa_result_batch* o) {
    if (!r || !o) return fail(PSA_ERR_ARG, "null argument");
    if (r->format != PSA_READS_ASCII && r->format != PSA_READS_PACKED) return fail(PSA_ERR_ARG, "unknown read format");
    if (r->n_reads && (!r->data || !o->hits)) return fail(PSA_ERR_ARG, "null data / hits");
    if (!r->read_len && r->n_reads && r->read_off == nullptr && r->stride == 0 && r->fixed_len != 0)
        return fail(PSA_ERR_ARG, "stride required when read_off is NULL");
    if (o->tx_cap && !o->tx_buf) return fail(PSA_ERR_ARG, "tx_cap without tx_buf");
    return PSA_OK;
}

// grow the novel-set buffer / the class-list pool after an overflow
static int grow_novel(psa_mapper* m) {
    m->novel_cap *= 4;
    return m->novel.ensure(m->novel_cap * 4);
}
// the novel-set table filled up (status bit 16): four times the entries and members, the used entries rehashed
static int grow_novel_table(psa_mapper* m) {
    DevBuf tab, pool;
    const uint64_t cap = m->ntab_cap * 4, pcap = m->npool_cap * 4;
    int rc;
    if ((rc = tab.ensure(cap * sizeof(NovelEntry))) || (rc = pool.ensure(pcap * 4))) {
        tab.release(); pool.release();
        return rc;
    }
    CU(cudaMemsetAsync(tab.p, 0, cap * sizeof(NovelEntry), m->st));
    k_novel_rehash<<<nblocks(m->ntab_cap, 256), 256, 0, m->st>>>(m->ntab.as<NovelEntry>(), m->ntab_cap, tab.as<NovelEntry>(), cap);
    // entries claimed by the failed attempt whose members did not fit are dropped by the rehash; the member cursor may
    // have run past the old pool: clamp it
    unsigned long long cur[2];
    CU(cudaMemcpyAsync(cur, m->ncur.p, 16, cudaMemcpyDeviceToHost, m->st));
    CU(cudaStreamSynchronize(m->st));
    cur[0] = std::min<unsigned long long>(cur[0], m->npool_cap);
    CU(cudaMemcpyAsync(pool.p, m->npool.p, cur[0] * 4, cudaMemcpyDeviceToDevice, m->st));   // the members in use
    CU(cudaMemcpyAsync(m->ncur.p, cur, 16, cudaMemcpyHostToDevice, m->st));
    CU(cudaStreamSynchronize(m->st));
    m->ntab.release(); m->npool.release();
    m->ntab = tab; m->npool = pool;
    m->ntab_cap = cap; m->npool_cap = pcap;
    return PSA_OK;
}
// the table as it was when the call began (a call that overflows somewhere is redone as a whole)
static int novel_backup(psa_mapper* m) {
    int rc;
    if ((rc = m->ntab_backup.ensure(m->ntab_cap * sizeof(NovelEntry))) || (rc = m->ncur_backup.ensure(16))) return rc;
    CU(cudaMemcpyAsync(m->ntab_backup.p, m->ntab.p, m->ntab_cap * sizeof(NovelEntry), cudaMemcpyDeviceToDevice, m->st));
    CU(cudaMemcpyAsync(m->ncur_backup.p, m->ncur.p, 16, cudaMemcpyDeviceToDevice, m->st));
    return PSA_OK;
}
static int novel_restore(psa_mapper* m) {
    CU(cudaMemcpyAsync(m->ntab.p, m->ntab_backup.p, m->ntab_cap * sizeof(NovelEntry), cudaMemcpyDeviceToDevice, m->st));
    CU(cudaMemcpyAsync(m->ncur.p, m->ncur_backup.p, 16, cudaMemcpyDeviceToDevice, m->st));
    CU(cudaStreamSynchronize(m->st));
    return PSA_OK;
}
static int novel_clash() {
    return fail(PSA_ERR_INTERNAL, "two different transcript sets share a 64-bit hash in the novel-set table (never merged silently)");
}
static int grow_pool(psa_mapper* m) {
    m->pool_cap *= 8;
    return m->pool.ensure(m->pool_cap * sizeof(uint4));
}

template <bool EV>
static int map_device_sync(psa_mapper* m, const psa_read_batch* r, psa_result_batch* o) {
    const uint64_t nc = m->ix->d.n_eq + 2;
    for (int attempt = 0; attempt < 6; attempt++) {
        CU(cudaMemcpyAsync(m->counts_backup.p, m->counts.p, nc * 8, cudaMemcpyDeviceToDevice, m->st));
        {
            int rcb = novel_backup(m);
            if (rcb) return rcb;
        }
        CU(cudaMemsetAsync(m->running.p, 0, 16, m->st));
        DeviceBatch b{r, (HitRec*)o->hits, o->tx_buf, o->tx_cap, m->meta.as<uint64_t>(), 0};
        if (o->flags & PSA_RESULT_COMPACT) {
            int rch = m->hits_full.ensure((r->n_reads + 1) * sizeof(HitRec));
            if (rch) return rch;
            b.hits = m->hits_full.as<HitRec>();
            b.hits_c = (HitCompact*)o->hits;
        }
        int rc = enqueue_device_batch<EV>(m, b, true);
        if (rc) return rc;
        CU(cudaMemcpyAsync(m->pin, m->meta.p, 16, cudaMemcpyDeviceToHost, m->st));
        CU(cudaStreamSynchronize(m->st));
        uint32_t status = (uint32_t)m->pin[1];
        o->tx_used = m->pin[0];
        if (status & 8u) return novel_clash();
        if (status & 32u) return fail(PSA_ERR_ARG, "compact results need coverage < 2^28 and class ids < 2^31");
        if (status & 19u) {  // novel-set buffer / class-list pool / novel-set table overflow: undo the counts, retry with larger buffers
            CU(cudaMemcpyAsync(m->counts.p, m->counts_backup.p, nc * 8, cudaMemcpyDeviceToDevice, m->st));
            CU(cudaStreamSynchronize(m->st));
            if ((rc = novel_restore(m))) return rc;
            if ((status & 1u) && (rc = grow_novel(m))) return rc;
            if ((status & 2u) && (rc = grow_pool(m))) return rc;
            if ((status & 16u) && (rc = grow_novel_table(m))) return rc;
            continue;
        }
        if (o->tx_buf && o->tx_used > o->tx_cap) {  // the caller resubmits: leave the counts as they were
            CU(cudaMemcpyAsync(m->counts.p, m->counts_backup.p, nc * 8, cudaMemcpyDeviceToDevice, m->st));
            CU(cudaStreamSynchronize(m->st));
            if ((rc = novel_restore(m))) return rc;
            return fail(PSA_ERR_CAPACITY, "tx_buf too small");
        }
        return PSA_OK;
    }
    return fail(PSA_ERR_CAPACITY, "novel-set buffer / class-list pool kept overflowing");
}

extern "C" int psa_mapper_map_async(psa_mapper* m, const psa_read_batch* r, psa_result_batch* o) {
    if (!m) return fail(PSA_ERR_ARG, "null argument");
    int rc = check_batch_args(r, o);
    if (rc) return rc;
    if (r->location != PSA_MEM_DEVICE || o->location != PSA_MEM_DEVICE)
        return fail(PSA_ERR_ARG, "psa_mapper_map_async needs device-resident batches");
    CU(cudaSetDevice(m->ix->device));
    CU(cudaMemsetAsync(m->running.p, 0, 16, m->st));
    DeviceBatch b{r, (HitRec*)o->hits, o->tx_buf, o->tx_cap, m->meta.as<uint64_t>(), 0};
    if (o->flags & PSA_RESULT_COMPACT) {
        if ((rc = m->hits_full.ensure((r->n_reads + 1) * sizeof(HitRec)))) return rc;
        b.hits = m->hits_full.as<HitRec>();
        b.hits_c = (HitCompact*)o->hits;
    }
    b.sticky = true;  // earlier batches may still be queued: their overflow bits stay visible until the next sync
    if ((rc = enqueue_device_batch<false>(m, b, true))) return rc;
    CU(cudaMemcpyAsync(m->pin, m->meta.p, 16, cudaMemcpyDeviceToHost, m->st));
    m->pending = o;
    return PSA_OK;
}

extern "C" int psa_mapper_sync(psa_mapper* m) {
    if (!m) return fail(PSA_ERR_ARG, "null argument");
    CU(cudaSetDevice(m->ix->device));
    CU(cudaStreamSynchronize(m->st));
    if (m->pending) {
        psa_result_batch* o = m->pending;
        m->pending = nullptr;
        CU(cudaMemsetAsync(m->status.as<uint32_t>() + 1, 0, 4, m->st));  // the sticky status restarts here
        o->tx_used = m->pin[0];
        uint32_t status = (uint32_t)m->pin[1];
        if (status & 8u) return novel_clash();
        if (status & 19u) {
            int rc = PSA_OK;
            if ((status & 1u) && (rc = grow_novel(m))) return rc;
            if ((status & 2u) && (rc = grow_pool(m))) return rc;
            if ((status & 16u) && (rc = grow_novel_table(m))) return rc;
            return fail(PSA_ERR_CAPACITY, "novel-set buffer / class-list pool / novel-set table overflow (buffers grown: reset the counts and resubmit the batch)");
        }
        if ((status & 4u) || (o->tx_buf && o->tx_used > o->tx_cap))
            return fail(PSA_ERR_CAPACITY, "tx_buf too small (for one of the batches queued since the last sync)");
    }
    return PSA_OK;
}

// Host-resident batch: chunks pipelined over three streams (H2D | kernels | D2H).
static int map_host(psa_mapper* m, const psa_read_batch* r, psa_result_batch* o) {
    const uint64_t n = r->n_reads;
    const uint64_t nc = m->ix->d.n_eq + 2;
    const bool ascii = r->format == PSA_READS_ASCII;
    const bool compact = (o->flags & PSA_RESULT_COMPACT) != 0;
    const uint64_t unit = ascii ? 1 : 8;  // bytes per data element
    const uint64_t C = m->chunk_reads;
    const uint64_t nchunks = (n + C - 1) / C;
    o->tx_used = 0;
    if (!n) return PSA_OK;

    struct ChunkPlan { uint64_t r0, nr, d0, dn, words; };  // reads [r0, r0+nr), data elements [d0, d0+dn)
    std::vector<ChunkPlan> plan(nchunks);
    uint64_t max_dn = 0;
    for (uint64_t c = 0; c < nchunks; c++) {
        ChunkPlan& P = plan[c];
        P.r0 = c * C;
        P.nr = std::min(C, n - P.r0);
        P.words = 0;
        uint64_t lo = ~0ull, hi = 0, max_span = 0;
        if (r->read_off || r->read_len) {
            for (uint64_t i = P.r0; i < P.r0 + P.nr; i++) {
                uint64_t L = r->read_len ? r->read_len[i] : r->fixed_len;
                uint64_t span = ascii ? L : (L + 31) / 32;
                P.words += (L + 31) / 32;
                max_span = std::max(max_span, span);
                if (r->read_off) {
                    lo = std::min(lo, r->read_off[i]);
                    hi = std::max(hi, r->read_off[i] + span);
                }
            }
        } else {
            max_span = ascii ? r->fixed_len : ((uint64_t)r->fixed_len + 31) / 32;
        }
        if (r->read_off) {
            P.d0 = lo;
            P.dn = hi - lo;
        } else {
            P.d0 = P.r0 * r->stride;
            P.dn = (P.nr - 1) * r->stride + max_span;
        }
        if (P.d0 + P.dn > r->data_len) return fail(PSA_ERR_ARG, "read extends past data_len");
        max_dn = std::max(max_dn, P.dn);
    }

    for (int attempt = 0; attempt < 6; attempt++) {
        int rc;
        CU(cudaMemcpyAsync(m->counts_backup.p, m->counts.p, nc * 8, cudaMemcpyDeviceToDevice, m->st));
        if ((rc = novel_backup(m))) return rc;
        CU(cudaMemsetAsync(m->running.p, 0, 16, m->st));
        const bool verbose = getenv("PSA_VERBOSE") != nullptr;
        auto now_s = []() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
        const double ta0 = verbose ? now_s() : 0;
        // staging for as many slots as the call has chunks (a one-chunk call uses slot 0 only)
        for (int s = 0; s < kSlots && (uint64_t)s < nchunks; s++) {
            Slot& S = m->slot[s];
            if ((rc = S.in_data.ensure(max_dn * unit + 64)) || (rc = S.hits.ensure(std::min(C, n) * sizeof(HitRec)))) return rc;
            if (compact && (rc = S.hits_c.ensure(std::min(C, n) * sizeof(HitCompact)))) return rc;
            if (r->read_off && (rc = S.in_off.ensure(C * 8))) return rc;
            if (r->read_len && (rc = S.in_len.ensure(C * 4))) return rc;
            if (o->tx_buf && (rc = S.tx.ensure(std::max<uint64_t>(S.tx.cap, std::min(C, n) * 16 * 4)))) return rc;
        }
        for (int s = 0; s < kSlots; s++) m->slot[s].in_free_rec = m->slot[s].out_free_rec = false;
        const double t_alloc = verbose ? now_s() - ta0 : 0;
        bool novel_overflow = false, spill_overflow = false, stage_overflow = false, ntab_overflow = false, clash = false, bad_compact = false;
        uint64_t stage_need = 0;
        uint64_t tx_prev_total = 0;  // host copy of the running total before the chunk being finished

        double t_wait = 0, t_submit = 0;
        auto finish = [&](uint64_t c) -> int {  // host side of chunk c: totals, members D2H
            Slot& S = m->slot[c % kSlots];
            const double tw0 = verbose ? now_s() : 0;
            CU(cudaEventSynchronize(S.meta_done));
            if (verbose) t_wait += now_s() - tw0;
            uint64_t total = S.meta_host[0];
            uint32_t status = (uint32_t)S.meta_host[1];
            if (status & 1u) novel_overflow = true;
            if (status & 2u) spill_overflow = true;
            if (status & 8u) clash = true;
            if (status & 16u) ntab_overflow = true;
            if (status & 32u) bad_compact = true;
            uint64_t cnt = total - tx_prev_total;
            if (status & 4u) {  // the chunk produced more members than its staging buffer holds
                stage_overflow = true;
                stage_need = std::max(stage_need, cnt);
            } else if (o->tx_buf && cnt && total <= o->tx_cap) {
                CU(cudaMemcpyAsync(o->tx_buf + tx_prev_total, S.tx.p, cnt * 4, cudaMemcpyDeviceToHost, m->st_d2h));
            }
            CU(cudaEventRecord(S.out_free, m->st_d2h));
            S.out_free_rec = true;
            tx_prev_total = total;
            return PSA_OK;
        };

        for (uint64_t c = 0; c < nchunks; c++) {
            const ChunkPlan& P = plan[c];
            Slot& S = m->slot[c % kSlots];
            const double ts0 = verbose ? now_s() : 0;
            // H2D
            if (S.in_free_rec) CU(cudaStreamWaitEvent(m->st_h2d, S.in_free, 0));
            CU(cudaMemcpyAsync(S.in_data.p, (const uint8_t*)r->data + P.d0 * unit, P.dn * unit, cudaMemcpyHostToDevice, m->st_h2d));
            // (the pack kernels read whole aligned words: the bytes behind the last read are masked out, but defined)
            CU(cudaMemsetAsync((uint8_t*)S.in_data.p + P.dn * unit, 0, 16, m->st_h2d));
            if (r->read_off) CU(cudaMemcpyAsync(S.in_off.p, r->read_off + P.r0, P.nr * 8, cudaMemcpyHostToDevice, m->st_h2d));
            if (r->read_len) CU(cudaMemcpyAsync(S.in_len.p, r->read_len + P.r0, P.nr * 4, cudaMemcpyHostToDevice, m->st_h2d));
            CU(cudaEventRecord(S.in_ready, m->st_h2d));
            // kernels
            CU(cudaStreamWaitEvent(m->st, S.in_ready, 0));
            if (S.out_free_rec) CU(cudaStreamWaitEvent(m->st, S.out_free, 0));
            psa_read_batch rb = *r;
            rb.location = PSA_MEM_DEVICE;
            rb.n_reads = P.nr;
            rb.data_len = P.dn;
            if (r->read_off) {  // offsets stay absolute: bias the base pointer by the chunk's first element
                rb.data = (const uint8_t*)S.in_data.p - P.d0 * unit;
                rb.read_off = S.in_off.as<uint64_t>();
            } else {
                rb.data = S.in_data.p;
                rb.read_off = nullptr;
            }
            rb.read_len = r->read_len ? S.in_len.as<uint32_t>() : nullptr;
            DeviceBatch b{&rb, S.hits.as<HitRec>(), o->tx_buf ? S.tx.as<uint32_t>() : nullptr,
                          o->tx_buf ? (uint64_t)(S.tx.cap / 4) : 0, S.meta_dev.as<uint64_t>(), P.words};
            if (compact) b.hits_c = S.hits_c.as<HitCompact>();
            if ((rc = enqueue_device_batch<false>(m, b, true))) return rc;
            CU(cudaEventRecord(S.comp_done, m->st));
            CU(cudaEventRecord(S.in_free, m->st));
            S.in_free_rec = true;
            // D2H: the two meta words first (the host waits on them), then the hits
            CU(cudaStreamWaitEvent(m->st_d2h, S.comp_done, 0));
            CU(cudaMemcpyAsync(S.meta_host, S.meta_dev.p, 16, cudaMemcpyDeviceToHost, m->st_d2h));
            CU(cudaEventRecord(S.meta_done, m->st_d2h));
            if (compact)
                CU(cudaMemcpyAsync((HitCompact*)o->hits + P.r0, S.hits_c.p, P.nr * sizeof(HitCompact), cudaMemcpyDeviceToHost, m->st_d2h));
            else
                CU(cudaMemcpyAsync(o->hits + P.r0, S.hits.p, P.nr * sizeof(HitRec), cudaMemcpyDeviceToHost, m->st_d2h));
            if (verbose) t_submit += now_s() - ts0;
            if (c >= (uint64_t)(kSlots - 1) && (rc = finish(c - (kSlots - 1)))) return rc;
        }
        for (uint64_t c = nchunks > (uint64_t)(kSlots - 1) ? nchunks - (kSlots - 1) : 0; c < nchunks; c++)
            if ((rc = finish(c))) return rc;
        const double td0 = verbose ? now_s() : 0;
        CU(cudaStreamSynchronize(m->st_d2h));
        CU(cudaStreamSynchronize(m->st));
        if (verbose)
            fprintf(stderr, "psa: map_host %llu chunks: staging buffers %.2f ms, submit %.2f ms, waits on chunk totals %.2f ms, final drain %.2f ms\n",
                    (unsigned long long)nchunks, 1e3 * t_alloc, 1e3 * t_submit, 1e3 * t_wait, 1e3 * (now_s() - td0));
        o->tx_used = tx_prev_total;
        if (clash) return novel_clash();
        if (bad_compact) return fail(PSA_ERR_ARG, "compact results need coverage < 2^28 and class ids < 2^31");
        if (novel_overflow || stage_overflow || spill_overflow || ntab_overflow) {
            // (a chunk that overflowed counted nothing, but the chunks around it did: the novel-set counts of this call
            // are undone with the per-class counts -- the table's counts are part of the backup below)
            CU(cudaMemcpyAsync(m->counts.p, m->counts_backup.p, nc * 8, cudaMemcpyDeviceToDevice, m->st));
            CU(cudaStreamSynchronize(m->st));
            if ((rc = novel_restore(m))) return rc;
            if (novel_overflow && m->novel_cap && (rc = grow_novel(m))) return rc;
            if (spill_overflow && (rc = grow_pool(m))) return rc;
            if (ntab_overflow && (rc = grow_novel_table(m))) return rc;
            if (stage_overflow)
                for (int s = 0; s < kSlots; s++)
                    if ((rc = m->slot[s].tx.ensure(stage_need * 4 + 4096))) return rc;
            continue;
        }
        if (o->tx_buf && o->tx_used > o->tx_cap) {  // the caller resubmits: leave the counts as they were
            CU(cudaMemcpyAsync(m->counts.p, m->counts_backup.p, nc * 8, cudaMemcpyDeviceToDevice, m->st));
            CU(cudaStreamSynchronize(m->st));
            if ((rc = novel_restore(m))) return rc;
            return fail(PSA_ERR_CAPACITY, "tx_buf too small");
        }
        return PSA_OK;
    }
    return fail(PSA_ERR_CAPACITY, "staging buffers kept overflowing");
}

extern "C" int psa_mapper_map(psa_mapper* m, const psa_read_batch* r, psa_result_batch* o) {
    if (!m) return fail(PSA_ERR_ARG, "null argument");
    int rc = check_batch_args(r, o);
    if (rc) return rc;
    if (r->location != o->location) return fail(PSA_ERR_ARG, "reads and results must live on the same side");
    CU(cudaSetDevice(m->ix->device));
    if (r->location == PSA_MEM_DEVICE) return map_device_sync<false>(m, r, o);
    return map_host(m, r, o);
}

extern "C" int psa_mapper_map_events(psa_mapper* m, const psa_read_batch* r, psa_result_batch* o, psa_events* out) {
    if (!m || !out) return fail(PSA_ERR_ARG, "null argument");
    int rc = check_batch_args(r, o);
    if (rc) return rc;
    if (r->location != PSA_MEM_DEVICE || o->location != PSA_MEM_DEVICE)
        return fail(PSA_ERR_ARG, "psa_mapper_map_events needs device-resident batches");
    CU(cudaSetDevice(m->ix->device));
    CU(cudaMemsetAsync(m->events.p, 0, 40 * 8, m->st));
    if ((rc = map_device_sync<true>(m, r, o))) return rc;
    static_assert(sizeof(psa_events) == 12 * 8, "psa_events layout");
    CU(cudaMemcpy(out, m->events.p, 3 * sizeof(psa_events), cudaMemcpyDeviceToHost));
    return PSA_OK;
}

extern "C" int psa_mapper_defer_reasons(psa_mapper* m, uint64_t out[4]) {
    if (!m || !out) return fail(PSA_ERR_ARG, "null argument");
    CU(cudaSetDevice(m->ix->device));
    CU(cudaMemcpy(out, m->events.as<uint64_t>() + 36, 4 * 8, cudaMemcpyDeviceToHost));
    return PSA_OK;
}

extern "C" int psa_mapper_map_read(psa_mapper* m, const uint64_t* read_words, uint32_t read_len, uint32_t* tx_out,
                                   uint64_t tx_cap, uint32_t* n_tx, uint32_t* coverage) {
    if (!m || !n_tx || !coverage || (read_len && !read_words)) return fail(PSA_ERR_ARG, "null argument");
    uint64_t dummy = 0;
    uint32_t len = read_len;
    uint64_t off = 0;
    psa_read_batch r{};
    r.format = PSA_READS_PACKED;
    r.location = PSA_MEM_HOST;
    r.data = read_len ? (const void*)read_words : (const void*)&dummy;
    r.data_len = ((uint64_t)read_len + 31) / 32;
    r.read_off = &off;
    r.read_len = &len;
    r.n_reads = 1;
    psa_hit h{};
    psa_result_batch o{};
    o.location = PSA_MEM_HOST;
    o.hits = &h;
    o.tx_buf = tx_out;
    o.tx_cap = tx_out ? tx_cap : 0;
    int rc = psa_mapper_map(m, &r, &o);
    *n_tx = h.n_tx;
    *coverage = h.coverage;
    if (rc) return rc;
    return (h.flags & PSA_FLAG_ALIGNED) ? 1 : 0;
}

extern "C" int psa_mapper_profile_enable(psa_mapper* m, int on) {
    if (!m) return fail(PSA_ERR_ARG, "null argument");
    m->profiling = on != 0;
    return PSA_OK;
}
extern "C" int psa_mapper_profile_read(psa_mapper* m, double map_kernel_ms[3], uint64_t map_launches[3]) {
    if (!m || !map_kernel_ms || !map_launches) return fail(PSA_ERR_ARG, "null argument");
    CU(cudaSetDevice(m->ix->device));
    CU(cudaStreamSynchronize(m->st));
    for (int w = 0; w < 3; w++) {
        double ms = 0;
        for (auto& pr : m->prof_events[w]) {
            float t = 0;
            CU(cudaEventElapsedTime(&t, pr.first, pr.second));
            ms += t;
            cudaEventDestroy(pr.first);
            cudaEventDestroy(pr.second);
        }
        map_kernel_ms[w] = ms;
        map_launches[w] = m->prof_events[w].size();
        m->prof_events[w].clear();
    }
    return PSA_OK;
}

extern "C" int psa_mapper_counts_get(psa_mapper* m, uint64_t* counts_host) {
    if (!m || !counts_host) return fail(PSA_ERR_ARG, "null argument");
    CU(cudaSetDevice(m->ix->device));
    CU(cudaStreamSynchronize(m->st));
    CU(cudaMemcpy(counts_host, m->counts.p, (m->ix->d.n_eq + 2) * 8, cudaMemcpyDeviceToHost));
    return PSA_OK;
}
extern "C" int psa_mapper_counts_reset(psa_mapper* m) {
    if (!m) return fail(PSA_ERR_ARG, "null argument");
    CU(cudaSetDevice(m->ix->device));
    CU(cudaMemsetAsync(m->counts.p, 0, (m->ix->d.n_eq + 2) * 8, m->st));
    CU(cudaMemsetAsync(m->ntab.p, 0, m->ntab_cap * sizeof(NovelEntry), m->st));
    CU(cudaMemsetAsync(m->ncur.p, 0, 16, m->st));
    CU(cudaStreamSynchronize(m->st));
    return PSA_OK;
}

// ---------------------------------------------------------------------------------------------
// novel sets: the table as a sorted host view, the merge of several views
// ---------------------------------------------------------------------------------------------
namespace {
struct SetRef {
    const uint32_t* m;
    uint32_t len;
    uint64_t count;
};
bool set_less(const SetRef& a, const SetRef& b) {   // (length, contents)
    if (a.len != b.len) return a.len < b.len;
    return std::lexicographical_compare(a.m, a.m + a.len, b.m, b.m + b.len);
}
bool set_equal(const SetRef& a, const SetRef& b) { return a.len == b.len && std::equal(a.m, a.m + a.len, b.m); }
// sorted, duplicates merged (counts added) -> malloc'ed arrays of *out
int build_sets(std::vector<SetRef>& v, psa_novel_sets* out) {
    std::sort(v.begin(), v.end(), set_less);
    size_t n = 0;
    for (size_t i = 0; i < v.size(); i++) {
        if (n && set_equal(v[n - 1], v[i])) v[n - 1].count += v[i].count;
        else v[n++] = v[i];
    }
    v.resize(n);
    uint64_t nm = 0;
    for (auto& e : v) nm += e.len;
    out->n_sets = n;
    out->n_members = nm;
    out->offsets = (uint64_t*)malloc((n + 1) * 8);
    out->members = (uint32_t*)malloc((nm + 1) * 4);
    out->counts = (uint64_t*)malloc((n + 1) * 8);
    if (!out->offsets || !out->members || !out->counts) {
        psa_novel_sets_free(out);
        return fail(PSA_ERR_NOMEM, "out of memory");
    }
    uint64_t o = 0;
    for (size_t i = 0; i < n; i++) {
        out->offsets[i] = o;
        if (v[i].len) memcpy(out->members + o, v[i].m, (size_t)v[i].len * 4);
        o += v[i].len;
        out->counts[i] = v[i].count;
    }
    out->offsets[n] = o;
    return PSA_OK;
}
}  // namespace

extern "C" void psa_novel_sets_free(psa_novel_sets* s) {
    if (!s) return;
    free(s->offsets); free(s->members); free(s->counts);
    s->offsets = nullptr; s->members = nullptr; s->counts = nullptr;
    s->n_sets = s->n_members = 0;
}

extern "C" int psa_novel_sets_merge(const psa_novel_sets* parts, uint32_t n_parts, psa_novel_sets* out) {
    if (!out || (n_parts && !parts)) return fail(PSA_ERR_ARG, "null argument");
    memset(out, 0, sizeof *out);
    std::vector<SetRef> v;
    for (uint32_t p = 0; p < n_parts; p++)
        for (uint64_t i = 0; i < parts[p].n_sets; i++) {
            const uint64_t a = parts[p].offsets[i], b = parts[p].offsets[i + 1];
            if (b < a || b > parts[p].n_members) return fail(PSA_ERR_ARG, "inconsistent offsets");
            v.push_back(SetRef{parts[p].members + a, (uint32_t)(b - a), parts[p].counts[i]});
        }
    return build_sets(v, out);
}

extern "C" int psa_mapper_novel_sets(psa_mapper* m, psa_novel_sets* out) {
    if (!m || !out) return fail(PSA_ERR_ARG, "null argument");
    memset(out, 0, sizeof *out);
    CU(cudaSetDevice(m->ix->device));
    CU(cudaStreamSynchronize(m->st));
    unsigned long long cur[2] = {0, 0};
    CU(cudaMemcpy(cur, m->ncur.p, 16, cudaMemcpyDeviceToHost));
    std::vector<NovelEntry> tab(m->ntab_cap);
    const uint64_t used = std::min<uint64_t>(cur[0], m->npool_cap);
    std::vector<uint32_t> pool(used + 1);
    CU(cudaMemcpy(tab.data(), m->ntab.p, m->ntab_cap * sizeof(NovelEntry), cudaMemcpyDeviceToHost));
    if (used) CU(cudaMemcpy(pool.data(), m->npool.p, used * 4, cudaMemcpyDeviceToHost));
    std::vector<SetRef> v;
    for (const NovelEntry& e : tab)
        if (e.key && e.count && e.len != 0xFFFFFFFFu && e.off + e.len <= used) v.push_back(SetRef{pool.data() + e.off, e.len, e.count});
    return build_sets(v, out);
}

// ---------------------------------------------------------------------------------------------
// NCCL: resolved at run time so that the library loads on machines without libnccl and uses
// whichever libnccl.so.2 the process already holds (torch's, when called from bench.py).
// ---------------------------------------------------------------------------------------------
struct Id128 {
    char b[128];
};
struct NcclApi {
    void* h = nullptr;
    int (*GetUniqueId)(void*) = nullptr;
    int (*CommInitRank)(void**, int, /*ncclUniqueId by value*/ Id128, int) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;
static int nccl_load() {
    if (g_nccl.h) return PSA_OK;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    for (auto nm : names)
        if ((h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL))) break;
    if (!h) return fail(PSA_ERR_NCCL, std::string("cannot load libnccl: ") + dlerror());
    g_nccl.GetUniqueId = (int (*)(void*))dlsym(h, "ncclGetUniqueId");
    g_nccl.CommInitRank = (int (*)(void**, int, Id128, int))dlsym(h, "ncclCommInitRank");
    g_nccl.CommDestroy = (int (*)(void*))dlsym(h, "ncclCommDestroy");
    g_nccl.AllReduce = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))dlsym(h, "ncclAllReduce");
    g_nccl.AllGather = (int (*)(const void*, void*, size_t, int, void*, cudaStream_t))dlsym(h, "ncclAllGather");
    g_nccl.GetErrorString = (const char* (*)(int))dlsym(h, "ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.CommDestroy || !g_nccl.AllReduce || !g_nccl.AllGather)
        return fail(PSA_ERR_NCCL, "libnccl lacks a required symbol");
    g_nccl.h = h;
    return PSA_OK;
}
static int nccl_fail(const char* what, int code) {
    return fail(PSA_ERR_NCCL, std::string(what) + ": " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(code) : "?"));
}
struct psa_comm {
    void* comm = nullptr;
    int world = 1, rank = 0, device = 0;
};
extern "C" int psa_comm_unique_id(uint8_t id[PSA_NCCL_UNIQUE_ID_BYTES]) {
    if (!id) return fail(PSA_ERR_ARG, "null argument");
    int rc = nccl_load();
    if (rc) return rc;
    int e = g_nccl.GetUniqueId(id);
    if (e) return nccl_fail("ncclGetUniqueId", e);
    return PSA_OK;
}
extern "C" int psa_comm_create(const uint8_t id[PSA_NCCL_UNIQUE_ID_BYTES], int world, int rank, int device,
                               psa_comm** out) {
    if (!id || !out || world < 1 || rank < 0 || rank >= world) return fail(PSA_ERR_ARG, "bad argument");
    int rc = nccl_load();
    if (rc) return rc;
    CU(cudaSetDevice(device));
    psa_comm* c = new (std::nothrow) psa_comm();
    if (!c) return fail(PSA_ERR_NOMEM, "out of memory");
    Id128 uid;
    memcpy(uid.b, id, 128);
    int e = g_nccl.CommInitRank(&c->comm, world, uid, rank);
    if (e) {
        delete c;
        return nccl_fail("ncclCommInitRank", e);
    }
    c->world = world; c->rank = rank; c->device = device;
    *out = c;
    return PSA_OK;
}
extern "C" void psa_comm_destroy(psa_comm* c) {
    if (!c) return;
    if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
    delete c;
}
extern "C" int psa_mapper_counts_allreduce(psa_mapper* m, psa_comm* c) {
    if (!m || !c) return fail(PSA_ERR_ARG, "null argument");
    CU(cudaSetDevice(m->ix->device));
    // ncclUint64 = 5, ncclSum = 0
    int e = g_nccl.AllReduce(m->counts.p, m->counts.p, m->ix->d.n_eq + 2, 5, 0, c->comm, m->st);
    if (e) return nccl_fail("ncclAllReduce", e);
    CU(cudaStreamSynchronize(m->st));
    return PSA_OK;
}

// The novel-set tables of all ranks: every rank serialises its view as u64 words
// [n_sets, n_members, offsets (n_sets+1), counts (n_sets), members (packed two per word)], the sizes are
// all-gathered, then the tables padded to the largest one; the merge is psa_novel_sets_merge on every rank.
extern "C" int psa_mapper_novel_allgather(psa_mapper* m, psa_comm* c, psa_novel_sets* out) {
    if (!m || !c || !out) return fail(PSA_ERR_ARG, "null argument");
    psa_novel_sets mine{};
    int rc = psa_mapper_novel_sets(m, &mine);
    if (rc) return rc;
    std::vector<uint64_t> ser;
    ser.push_back(mine.n_sets);
    ser.push_back(mine.n_members);
    ser.insert(ser.end(), mine.offsets, mine.offsets + mine.n_sets + 1);
    ser.insert(ser.end(), mine.counts, mine.counts + mine.n_sets);
    const size_t mw = (mine.n_members + 1) / 2;
    const size_t at = ser.size();
    ser.resize(at + mw, 0);
    if (mine.n_members) memcpy(ser.data() + at, mine.members, mine.n_members * 4);
    psa_novel_sets_free(&mine);
    const int W = c->world;
    DevBuf dsz, dall_sz, dsend, drecv;
    auto rel = [&]() { dsz.release(); dall_sz.release(); dsend.release(); drecv.release(); };
    unsigned long long my_words = ser.size();
    std::vector<unsigned long long> sizes(W);
    if ((rc = dsz.ensure(8)) || (rc = dall_sz.ensure(8 * (size_t)W))) { rel(); return rc; }
    cudaMemcpyAsync(dsz.p, &my_words, 8, cudaMemcpyHostToDevice, m->st);
    int e = g_nccl.AllGather(dsz.p, dall_sz.p, 1, 5 /* ncclUint64 */, c->comm, m->st);
    if (e) { rel(); return nccl_fail("ncclAllGather", e); }
    cudaMemcpyAsync(sizes.data(), dall_sz.p, 8 * (size_t)W, cudaMemcpyDeviceToHost, m->st);
    if (cudaStreamSynchronize(m->st) != cudaSuccess) { rel(); return fail(PSA_ERR_CUDA, "novel all-gather (sizes)"); }
    unsigned long long max_words = 0;
    for (auto x : sizes) max_words = std::max(max_words, x);
    ser.resize(max_words, 0);
    if ((rc = dsend.ensure(max_words * 8)) || (rc = drecv.ensure(max_words * 8 * (size_t)W))) { rel(); return rc; }
    cudaMemcpyAsync(dsend.p, ser.data(), max_words * 8, cudaMemcpyHostToDevice, m->st);
    e = g_nccl.AllGather(dsend.p, drecv.p, max_words, 5, c->comm, m->st);
    if (e) { rel(); return nccl_fail("ncclAllGather", e); }
    std::vector<uint64_t> all(max_words * (size_t)W);
    cudaMemcpyAsync(all.data(), drecv.p, all.size() * 8, cudaMemcpyDeviceToHost, m->st);
    if (cudaStreamSynchronize(m->st) != cudaSuccess) { rel(); return fail(PSA_ERR_CUDA, "novel all-gather (tables)"); }
    rel();
    std::vector<psa_novel_sets> parts(W);
    for (int r = 0; r < W; r++) {
        uint64_t* b = all.data() + (size_t)r * max_words;
        parts[r].n_sets = b[0];
        parts[r].n_members = b[1];
        parts[r].offsets = b + 2;
        parts[r].counts = b + 2 + parts[r].n_sets + 1;
        parts[r].members = reinterpret_cast<uint32_t*>(b + 2 + 2 * parts[r].n_sets + 1);
    }
    return psa_novel_sets_merge(parts.data(), (uint32_t)W, out);
}

// ---------------------------------------------------------------------------------------------
// measurement aid
// ---------------------------------------------------------------------------------------------
extern "C" int psa_gather_probe(int device, uint64_t table_bytes, uint32_t chunk_bytes, uint32_t iters, double* gbytes_per_s) {
    if (!gbytes_per_s || (chunk_bytes != 0 && chunk_bytes != 32 && chunk_bytes != 64 && chunk_bytes != 128) ||
        table_bytes < (1u << 20) || !iters)
        return fail(PSA_ERR_ARG, "bad argument");
    CU(cudaSetDevice(device));
    DevBuf table, sink;
    int rc;
    if ((rc = table.ensure(table_bytes)) || (rc = sink.ensure(8))) {
        table.release();
        return rc;
    }
    cudaMemset(table.p, 1, table_bytes);
    const uint64_t n_chunks = table_bytes / (chunk_bytes ? chunk_bytes : 128);
    const unsigned grid = 148 * 16, block = 256;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0, 0);
        if (chunk_bytes == 0) k_gather_probe<0><<<grid, block>>>(table.as<uint64_t>(), n_chunks, iters, 17 + rep, sink.as<uint64_t>());
        else if (chunk_bytes == 32) k_gather_probe<32><<<grid, block>>>(table.as<uint64_t>(), n_chunks, iters, 17 + rep, sink.as<uint64_t>());
        else if (chunk_bytes == 64) k_gather_probe<64><<<grid, block>>>(table.as<uint64_t>(), n_chunks, iters, 17 + rep, sink.as<uint64_t>());
        else k_gather_probe<128><<<grid, block>>>(table.as<uint64_t>(), n_chunks, iters, 17 + rep, sink.as<uint64_t>());
        cudaEventRecord(e1, 0);
        cudaError_t e = cudaEventSynchronize(e1);
        if (e != cudaSuccess) {
            table.release(); sink.release();
            return fail(PSA_ERR_CUDA, cudaGetErrorString(e));
        }
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    table.release();
    sink.release();
    *gbytes_per_s = (double)grid * block * iters * (chunk_bytes ? chunk_bytes : 4) / (best / 1e3) / 1e9;
    return PSA_OK;
}

// ---------------------------------------------------------------------------------------------
// measurement aid: the synthetic read stream of include/psa_host.h (psa_synth_reads, csrc/host/synth.cpp)
// generated on the device -- the same counter-based generator, read i = f(seed, i), so that a billion-read
// configuration needs no host generation and any sample of it can be regenerated on the host for the oracle.
// ---------------------------------------------------------------------------------------------
namespace {
struct DevRng {  // == synth.cpp Rng
    uint64_t s;
    __device__ static uint64_t splitmix(uint64_t& s) {
        uint64_t z = (s += 0x9E3779B97F4A7C15ULL);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
        return z ^ (z >> 31);
    }
    __device__ DevRng(uint64_t seed, uint64_t tag, uint64_t index) {
        s = seed * 0xD1342543DE82EF95ULL + tag;
        s = splitmix(s) ^ (index * 0xA24BAED4963EE407ULL);
        splitmix(s);
    }
    __device__ uint64_t next() { return splitmix(s); }
    __device__ double uniform() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
    __device__ uint64_t below(uint64_t n) { return __umul64hi(next(), n); }
};
__device__ const uint8_t* synth_pick(const psa_synth_tables& t, int w, DevRng& g) {
    const uint64_t* cum = t.cum[w];
    const uint64_t x = g.below(cum[t.n_elig[w]]);
    uint64_t lo = 0, hi = t.n_elig[w] + 1;   // upper_bound(cum, x) - 1
    while (lo < hi) {
        const uint64_t mid = lo + ((hi - lo) >> 1);
        if (cum[mid] <= x) lo = mid + 1;
        else hi = mid;
    }
    const uint64_t i = lo - 1;
    return t.codes + t.tx_off[t.elig[w][i]] + (x - cum[i]);
}
__global__ void k_synth_reads(psa_synth_tables t, uint64_t seed, uint64_t first, uint64_t n, uint32_t L, uint8_t* out, uint64_t stride) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    DevRng g(seed, 3, first + i);
    uint8_t* r = out + i * stride;
    const uint32_t h0 = L / 2, h1 = L - L / 2;
    const double u = g.uniform();
    int kind = u < 0.90 ? 0 : (u < 0.95 ? 1 : 2);
    if (kind == 0 && !t.n_elig[0]) kind = 2;
    if (kind == 1 && !(t.n_elig[1] && t.n_elig[2])) kind = 2;
    if (kind == 0) {
        const uint8_t* s = synth_pick(t, 0, g);
        for (uint32_t j = 0; j < L; j++) r[j] = s[j];
    } else if (kind == 1) {
        const uint8_t* a = synth_pick(t, 1, g);
        const uint8_t* c = synth_pick(t, 2, g);
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
    for (uint32_t j = 0; j < L; j++) r[j] = (uint8_t)"ACGT"[r[j]];
}
}  // namespace
extern "C" int psa_synth_reads_device(int device, const psa_synth_tables* t, uint64_t seed, uint64_t first, uint64_t n, uint32_t L,
                                      uint8_t* out_dev, uint64_t stride) {
    if (!t || !out_dev || stride < L || !t->codes || !t->tx_off) return fail(PSA_ERR_ARG, "bad argument");
    CU(cudaSetDevice(device));
    if (n) k_synth_reads<<<nblocks(n, 128), 128>>>(*t, seed, first, n, L, out_dev, stride);
    CU(cudaGetLastError());
    CU(cudaDeviceSynchronize());
    return PSA_OK;
}

extern "C" int psa_selftest_intersect(int device, const uint32_t* v1, uint32_t n1, const uint32_t* v2, uint32_t n2,
                                      uint32_t* out, uint32_t cap, uint32_t n_out[3]) {
    if (!out || !n_out || (n1 && !v1) || (n2 && !v2)) return fail(PSA_ERR_ARG, "null argument");
    CU(cudaSetDevice(device));
    DevBuf mem, off, dout, dn;
    int rc;
    auto rel = [&]() { mem.release(); off.release(); dout.release(); dn.release(); };
    if ((rc = mem.ensure(((uint64_t)n1 + n2 + 1) * 4)) || (rc = off.ensure(3 * 8)) || (rc = dout.ensure(3 * (uint64_t)cap * 4 + 4)) ||
        (rc = dn.ensure(3 * 4))) {
        rel();
        return rc;
    }
    const uint64_t offs[3] = {0, n1, (uint64_t)n1 + n2};
    cudaMemcpy(off.p, offs, sizeof offs, cudaMemcpyHostToDevice);
    if (n1) cudaMemcpy(mem.p, v1, (uint64_t)n1 * 4, cudaMemcpyHostToDevice);
    if (n2) cudaMemcpy(mem.as<uint32_t>() + n1, v2, (uint64_t)n2 * 4, cudaMemcpyHostToDevice);
    cudaMemset(dn.p, 0, 12);
    DevIndex d{};
    d.n_eq = 2;
    d.eq_off = off.as<uint64_t>();
    d.eq_mem = mem.as<uint32_t>();
    MapParams p{};
    k_selftest_intersect<<<1, 32>>>(d, p, dout.as<uint32_t>(), cap, dn.as<uint32_t>());
    cudaError_t e = cudaMemcpy(n_out, dn.p, 12, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(out, dout.p, 3 * (uint64_t)cap * 4, cudaMemcpyDeviceToHost);
    rel();
    if (e != cudaSuccess) return fail(PSA_ERR_CUDA, cudaGetErrorString(e));
    return PSA_OK;
}

// Compact results -> the full form, on the host: index classes are expanded from the caller's copy of eq_classes.
extern "C" int psa_expand_compact(const psa_hit_compact* in, uint64_t n, const uint32_t* novel_tx, uint64_t novel_tx_len,
                                  const uint64_t* eq_offsets, const uint32_t* eq_members, uint64_t n_eq, psa_hit* hits,
                                  uint32_t* tx_buf, uint64_t tx_cap, uint64_t* tx_used) {
    if (!in || !hits || !tx_used || !eq_offsets || (n_eq && eq_offsets[n_eq] && !eq_members)) return fail(PSA_ERR_ARG, "null argument");
    uint64_t used = 0, nov = 0;
    for (uint64_t i = 0; i < n; i++) {
        const psa_hit_compact c = in[i];
        psa_hit h;
        h.coverage = c.cov_flags & ((1u << 28) - 1);
        h.flags = c.cov_flags >> 28;
        h.tx_off = used;
        const uint32_t* src = nullptr;
        if (!(h.flags & PSA_FLAG_ALIGNED)) {
            h.eq_id = PSA_EQ_NONE; h.n_tx = 0;
        } else if (c.eq_or_n & 0x80000000u) {
            h.eq_id = PSA_EQ_NONE;
            h.n_tx = c.eq_or_n & 0x7FFFFFFFu;
            if (nov + h.n_tx > novel_tx_len) return fail(PSA_ERR_ARG, "novel members missing");
            src = novel_tx + nov;
            nov += h.n_tx;
        } else {
            if (c.eq_or_n >= n_eq) return fail(PSA_ERR_ARG, "class id out of range");
            h.eq_id = c.eq_or_n;
            h.n_tx = (uint32_t)(eq_offsets[h.eq_id + 1] - eq_offsets[h.eq_id]);
            src = eq_members + eq_offsets[h.eq_id];
        }
        if (tx_buf && used + h.n_tx <= tx_cap && h.n_tx) memcpy(tx_buf + used, src, (size_t)h.n_tx * 4);
        used += h.n_tx;
        hits[i] = h;
    }
    *tx_used = used;
    if (tx_buf && used > tx_cap) return fail(PSA_ERR_CAPACITY, "tx_buf too small");
    return PSA_OK;
}

extern "C" int psa_result_checksum(int device, const psa_hit* hits_dev, const uint32_t* tx_dev, uint64_t n,
                                   uint64_t first_index, uint64_t* out) {
    if (!out || (n && (!hits_dev || !tx_dev))) return fail(PSA_ERR_ARG, "null argument");
    CU(cudaSetDevice(device));
    DevBuf acc;
    int rc = acc.ensure(8);
    if (rc) return rc;
    cudaMemset(acc.p, 0, 8);
    if (n) k_result_checksum<<<nblocks(n, 256), 256>>>((const HitRec*)hits_dev, tx_dev, n, first_index, 0, acc.as<unsigned long long>());
    cudaError_t e = cudaMemcpy(out, acc.p, 8, cudaMemcpyDeviceToHost);
    acc.release();
    if (e != cudaSuccess) return fail(PSA_ERR_CUDA, cudaGetErrorString(e));
    return PSA_OK;
}

// ---------------------------------------------------------------------------------------------
// mappability::analyze_graph (ref src/mappability.rs:120-156) on the device: for every node, its k-mer
// count goes into the histograms of every transcript of its class, binned by the class's number of
// transcripts resp. distinct genes (ref :59-73: bin = multiplicity - 1, everything above the last bin
// into the last one).
// ---------------------------------------------------------------------------------------------
namespace {
// distinct genes per class (ref :136-143: eq_class.iter().map(gene).unique().count())
__global__ void k_class_genes(const uint64_t* eq_off, const uint32_t* eq_mem, uint64_t n_eq, const uint32_t* tx_gene, uint32_t* n_genes) {
    const uint64_t c = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (c >= n_eq) return;
    const uint64_t o = eq_off[c], n = eq_off[c + 1] - o;
    uint32_t distinct = 0;
    for (uint64_t i = 0; i < n; i++) {
        const uint32_t g = tx_gene[eq_mem[o + i]];
        bool seen = false;
        for (uint64_t j = 0; j < i && !seen; j++) seen = tx_gene[eq_mem[o + j]] == g;
        distinct += seen ? 0u : 1u;
    }
    n_genes[c] = distinct;
}
__global__ void k_mappability(const NodeRec* nodes, uint64_t n_nodes, uint32_t k, const uint64_t* eq_off, const uint32_t* eq_mem,
                              const uint32_t* n_genes, uint32_t bins, unsigned long long* tx_mult, unsigned long long* gene_mult) {
    const uint64_t v = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (v >= n_nodes) return;
    const NodeRec nr = nodes[v];
    const unsigned long long num_kmer = (uint32_t)(nr.start_len >> 40) - k + 1;   // :128
    const uint64_t o = eq_off[nr.eq], num_tx = eq_off[nr.eq + 1] - o;             // :130-133
    if (!num_tx) return;
    const uint32_t num_genes = n_genes[nr.eq];
    const uint32_t tb = num_tx > bins ? bins - 1 : (uint32_t)num_tx - 1;          // :59-65
    const uint32_t gb = num_genes > bins ? bins - 1 : num_genes - 1;              // :67-73
    for (uint64_t i = 0; i < num_tx; i++) {                                       // :145-149
        const uint64_t tx = eq_mem[o + i];
        atomicAdd(tx_mult + tx * bins + tb, num_kmer);
        atomicAdd(gene_mult + tx * bins + gb, num_kmer);
    }
}
}  // namespace
extern "C" int psa_index_mappability(psa_index* ix, const uint32_t* tx_gene, uint32_t n_tx, uint32_t bins, uint64_t* tx_multiplicity,
                                     uint64_t* gene_multiplicity) {
    if (!ix || !tx_gene || !bins || !tx_multiplicity || !gene_multiplicity) return fail(PSA_ERR_ARG, "null argument");
    for (uint64_t i = 0; i < ix->h_eq_mem.size(); i++)
        if (ix->h_eq_mem[i] >= n_tx) return fail(PSA_ERR_ARG, "a class names a transcript >= n_tx");
    CU(cudaSetDevice(ix->device));
    DevBuf genes, ng, tm, gm;
    int rc;
    const size_t hist = (size_t)n_tx * bins * 8;
    if ((rc = genes.ensure((size_t)n_tx * 4 + 4)) || (rc = ng.ensure(ix->d.n_eq * 4 + 4)) || (rc = tm.ensure(hist + 8)) || (rc = gm.ensure(hist + 8))) {
        genes.release(); ng.release(); tm.release(); gm.release();
        return rc;
    }
    cudaError_t e = cudaMemcpy(genes.p, tx_gene, (size_t)n_tx * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemset(tm.p, 0, hist);
    if (e == cudaSuccess) e = cudaMemset(gm.p, 0, hist);
    if (e == cudaSuccess && ix->d.n_eq)
        k_class_genes<<<nblocks(ix->d.n_eq, 128), 128>>>(ix->d.eq_off, ix->d.eq_mem, ix->d.n_eq, genes.as<uint32_t>(), ng.as<uint32_t>());
    if (e == cudaSuccess && ix->d.n_nodes)
        k_mappability<<<nblocks(ix->d.n_nodes, 128), 128>>>(ix->d.nodes, ix->d.n_nodes, ix->d.k, ix->d.eq_off, ix->d.eq_mem, ng.as<uint32_t>(), bins,
                                                            tm.as<unsigned long long>(), gm.as<unsigned long long>());
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpy(tx_multiplicity, tm.p, hist, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(gene_multiplicity, gm.p, hist, cudaMemcpyDeviceToHost);
    genes.release(); ng.release(); tm.release(); gm.release();
    if (e != cudaSuccess) return fail(PSA_ERR_CUDA, cudaGetErrorString(e));
    return PSA_OK;
}

// ---------------------------------------------------------------------------------------------
// FASTQ text lanes (psa_fastq.h): raw FASTQ block in, `{:?}` lines out, everything in between on the
// device -- newline index, record table, ASCII -> 2-bit, map_read, decimal formatting (psa_fastq.cuh).
// One lane = one mapper (own stream) + the buffers of one block; psa_process_reads runs several lanes
// side by side so that one lane's copies overlap another's kernels.
// ---------------------------------------------------------------------------------------------
struct psa_fq_lane {
    psa_index* ix = nullptr;
    psa_mapper* m = nullptr;
    uint64_t block_bytes = 0, tail_bytes = 0;
    uint8_t* h_text = nullptr;            // pinned: block + tail (+ 64)
    char* h_out = nullptr;                // pinned: result lines
    uint64_t h_out_cap = 0;
    unsigned long long* h_small = nullptr;  // pinned scratch for the small read-backs
    DevBuf d_text, d_tile, d_tile_off, d_nl, d_seq_off, d_seq_len, d_id_off, d_id_len, d_hits, d_tx, d_line_len, d_line_off, d_out,
        d_small, d_scan;
    uint64_t len = 0, n_tiles = 0, nl_total = 0;
};

// a lane that is done goes back to its index (at most kIdleLanes are kept)
constexpr size_t kIdleLanes = 4;
extern "C" void psa_fq_lane_destroy(psa_fq_lane* l) {
    if (!l) return;
    if (l->m && !getenv("PSA_FQ_NO_LANE_CACHE")) {
        cudaSetDevice(l->ix->device);
        cudaStreamSynchronize(l->m->st);
        std::lock_guard<std::mutex> g(l->ix->lanes_mu);
        if (l->ix->idle_lanes.size() < kIdleLanes) {
            l->ix->idle_lanes.push_back(l);
            return;
        }
    }
    fq_lane_free(l);
}
static void fq_lane_free(psa_fq_lane* l) {
    if (!l) return;
    if (l->m) {
        cudaSetDevice(l->ix->device);
        cudaStreamSynchronize(l->m->st);
    }
    DevBuf* bufs[] = {&l->d_text, &l->d_tile, &l->d_tile_off, &l->d_nl, &l->d_seq_off, &l->d_seq_len, &l->d_id_off, &l->d_id_len,
                      &l->d_hits, &l->d_tx, &l->d_line_len, &l->d_line_off, &l->d_out, &l->d_small, &l->d_scan};
    for (auto b : bufs) b->release();
    if (l->h_text) cudaFreeHost(l->h_text);
    if (l->h_out) cudaFreeHost(l->h_out);
    if (l->h_small) cudaFreeHost(l->h_small);
    if (l->m) psa_mapper_destroy(l->m);
    delete l;
}

extern "C" int psa_fq_lane_create(psa_index* ix, uint64_t block_bytes, uint64_t tail_bytes, psa_fq_lane** out) {
    if (!ix || !out || !block_bytes || block_bytes % kFqTile) return fail(PSA_ERR_ARG, "block_bytes must be a positive multiple of 4096");
    if (block_bytes + tail_bytes >= (1ull << 32) - 2 * kFqTile) return fail(PSA_ERR_ARG, "a FASTQ block is at most 4 GB");
    *out = nullptr;
    CU(cudaSetDevice(ix->device));
    {   // an idle lane of the same geometry?
        std::lock_guard<std::mutex> g(ix->lanes_mu);
        for (size_t i = 0; i < ix->idle_lanes.size(); i++) {
            psa_fq_lane* c = ix->idle_lanes[i];
            if (c->block_bytes == block_bytes && c->tail_bytes == tail_bytes) {
                ix->idle_lanes.erase(ix->idle_lanes.begin() + (long)i);
                psa_mapper_counts_reset(c->m);
                *out = c;
                return PSA_OK;
            }
        }
    }
    psa_fq_lane* l = new (std::nothrow) psa_fq_lane();
    if (!l) return fail(PSA_ERR_NOMEM, "out of memory");
    l->ix = ix;
    l->block_bytes = block_bytes;
    l->tail_bytes = tail_bytes;
    int rc = psa_mapper_create(ix, 0, &l->m);
    const uint64_t text_cap = block_bytes + tail_bytes + 64;
    cudaError_t e = cudaSuccess;
    if (!rc) e = cudaHostAlloc((void**)&l->h_text, text_cap, cudaHostAllocDefault);
    if (!rc && e == cudaSuccess) e = cudaHostAlloc((void**)&l->h_small, 256, cudaHostAllocDefault);
    l->h_out_cap = std::max<uint64_t>(block_bytes / 2, 1u << 16);
    if (!rc && e == cudaSuccess) e = cudaHostAlloc((void**)&l->h_out, l->h_out_cap, cudaHostAllocDefault);
    if (!rc && e != cudaSuccess) rc = fail(PSA_ERR_CUDA, cudaGetErrorString(e));
    if (!rc) rc = l->d_text.ensure(text_cap + 2 * kFqTile);
    if (!rc) rc = l->d_small.ensure(256);
    if (!rc) rc = l->d_out.ensure(l->h_out_cap + 64);
    if (rc) {
        fq_lane_free(l);
        return rc;
    }
    *out = l;
    return PSA_OK;
}

extern "C" uint8_t* psa_fq_lane_text(psa_fq_lane* l) { return l ? l->h_text : nullptr; }

extern "C" int psa_fq_lane_index(psa_fq_lane* l, uint64_t len, uint64_t own_bytes, uint64_t* nl_own, uint64_t* nl_total) {
    if (!l || !nl_own || !nl_total || len > l->block_bytes + l->tail_bytes + 64 || own_bytes > len ||
        (own_bytes != len && own_bytes % kFqTile))
        return fail(PSA_ERR_ARG, "bad FASTQ block geometry");
    CU(cudaSetDevice(l->ix->device));
    cudaStream_t st = l->m->st;
    l->len = len;
    l->n_tiles = (len + kFqTile - 1) / kFqTile;
    l->nl_total = 0;
    *nl_own = *nl_total = 0;
    if (!len) return PSA_OK;
    int rc;
    if ((rc = l->d_tile.ensure((l->n_tiles + 1) * 4)) || (rc = l->d_tile_off.ensure((l->n_tiles + 1) * 4))) return rc;
    CU(cudaMemcpyAsync(l->d_text.p, l->h_text, len, cudaMemcpyHostToDevice, st));
    CU(cudaMemsetAsync(l->d_text.as<uint8_t>() + len, 0, l->n_tiles * kFqTile - len, st));   // the last tile is read whole
    CU(cudaMemsetAsync(l->d_tile.as<uint32_t>() + l->n_tiles, 0, 4, st));
    k_fq_count<<<(unsigned)l->n_tiles, 256, 0, st>>>(l->d_text.as<uint8_t>(), l->d_tile.as<uint32_t>());
    size_t tb = 0;
    CU(cub::DeviceScan::ExclusiveSum(nullptr, tb, l->d_tile.as<uint32_t>(), l->d_tile_off.as<uint32_t>(), l->n_tiles + 1, st));
    if ((rc = l->d_scan.ensure(tb))) return rc;
    CU(cub::DeviceScan::ExclusiveSum(l->d_scan.p, tb, l->d_tile.as<uint32_t>(), l->d_tile_off.as<uint32_t>(), l->n_tiles + 1, st));
    const uint64_t own_tiles = own_bytes == len ? l->n_tiles : own_bytes / kFqTile;
    uint32_t* hs = reinterpret_cast<uint32_t*>(l->h_small);
    CU(cudaMemcpyAsync(hs, l->d_tile_off.as<uint32_t>() + own_tiles, 4, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(hs + 1, l->d_tile_off.as<uint32_t>() + l->n_tiles, 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    *nl_own = hs[0];
    *nl_total = l->nl_total = hs[1];
    if ((rc = l->d_nl.ensure((l->nl_total + 1) * 4))) return rc;
    k_fq_positions<<<(unsigned)l->n_tiles, 256, 0, st>>>(l->d_text.as<uint8_t>(), l->d_tile_off.as<uint32_t>(), l->d_nl.as<uint32_t>());
    l->m->launches += 3;
    CU(cudaGetLastError());
    return PSA_OK;
}

extern "C" int psa_fq_lane_run(psa_fq_lane* l, int64_t j0, uint64_t n, const uint64_t* tick_at, uint32_t n_ticks, psa_fq_result* out) {
    if (!l || !out || n_ticks > PSA_FQ_MAX_TICKS || (n_ticks && !tick_at)) return fail(PSA_ERR_ARG, "bad argument");
    memset(out, 0, sizeof *out);
    out->plain = 1;
    out->out_text = l->h_out;
    if (!n) return PSA_OK;
    if (j0 < -1 || (uint64_t)(j0 + 4 * (int64_t)(n - 1) + 4) >= l->nl_total) return fail(PSA_ERR_ARG, "records beyond the indexed newlines");
    CU(cudaSetDevice(l->ix->device));
    cudaStream_t st = l->m->st;
    int rc;
    if ((rc = l->d_seq_off.ensure(n * 8)) || (rc = l->d_seq_len.ensure(n * 4)) || (rc = l->d_id_off.ensure(n * 4)) ||
        (rc = l->d_id_len.ensure(n * 4)) || (rc = l->d_hits.ensure((n + 1) * sizeof(HitRec))) || (rc = l->d_line_len.ensure((n + 1) * 4)) ||
        (rc = l->d_line_off.ensure((n + 2) * 8)))
        return rc;
    if (!l->d_tx.cap && (rc = l->d_tx.ensure(std::max<uint64_t>(n * 16 * 4, 1 << 20)))) return rc;
    unsigned long long* ds = l->d_small.as<unsigned long long>();   // [0] status, [1] mapped, [2] aligned, [3] end offset, [8..16) tick positions, [16..24) tick counts
    CU(cudaMemsetAsync(ds, 0, 256, st));
    k_fq_records<<<nblocks(n, 128), 128, 0, st>>>(l->d_text.as<uint8_t>(), l->d_nl.as<uint32_t>(), j0, n, l->d_seq_off.as<uint64_t>(),
                                                  l->d_seq_len.as<uint32_t>(), l->d_id_off.as<uint32_t>(), l->d_id_len.as<uint32_t>(),
                                                  reinterpret_cast<uint32_t*>(ds));
    l->m->launches++;
    CU(cudaGetLastError());
    // map_read for every record: the sequences are addressed by offset in the block's text
    psa_read_batch r{};
    r.format = PSA_READS_ASCII;
    r.location = PSA_MEM_DEVICE;
    r.data = l->d_text.p;
    r.data_len = l->len;
    r.read_off = l->d_seq_off.as<uint64_t>();
    r.read_len = l->d_seq_len.as<uint32_t>();
    r.n_reads = n;
    psa_result_batch o{};
    o.location = PSA_MEM_DEVICE;
    for (int attempt = 0;; attempt++) {
        o.hits = reinterpret_cast<psa_hit*>(l->d_hits.p);
        o.tx_buf = l->d_tx.as<uint32_t>();
        o.tx_cap = l->d_tx.cap / 4;
        rc = map_device_sync<false>(l->m, &r, &o);
        if (rc == PSA_ERR_CAPACITY && o.tx_used > o.tx_cap && attempt == 0) {
            if ((rc = l->d_tx.ensure(o.tx_used * 4 + 4096))) return rc;
            continue;
        }
        break;
    }
    if (rc) return rc;
    // (map_device_sync has synchronised the stream: the record kernel's verdict is there)
    CU(cudaMemcpyAsync(l->h_small, ds, 8, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (l->h_small[0] & 1ull) {
        out->plain = 0;
        return PSA_OK;
    }
    // line lengths -> offsets
    k_fq_line_len<<<nblocks(n, 128), 128, 0, st>>>(l->d_text.as<uint8_t>(), l->d_id_off.as<uint32_t>(), l->d_id_len.as<uint32_t>(),
                                                   l->d_hits.as<HitRec>(), l->d_tx.as<uint32_t>(), n, l->d_line_len.as<uint32_t>(), ds + 1);
    {
        cub::CountingInputIterator<uint64_t> cnt(0);
        FqLenToU64 f{l->d_line_len.as<uint32_t>(), n};
        cub::TransformInputIterator<uint64_t, FqLenToU64, cub::CountingInputIterator<uint64_t>> in(cnt, f);
        size_t tb = 0;
        CU(cub::DeviceScan::ExclusiveSum(nullptr, tb, in, l->d_line_off.as<uint64_t>(), n + 1, st));
        if ((rc = l->d_scan.ensure(tb))) return rc;
        CU(cub::DeviceScan::ExclusiveSum(l->d_scan.p, tb, in, l->d_line_off.as<uint64_t>(), n + 1, st));
    }
    CU(cudaMemcpyAsync(l->h_small, l->d_line_off.as<uint64_t>() + n, 8, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(l->h_small + 1, ds + 1, 16, cudaMemcpyDeviceToHost, st));
    {   // end of the last record = the newline that ends its quality line, + 1
        const uint64_t last_nl = (uint64_t)(j0 + 4 * (int64_t)(n - 1) + 4);
        CU(cudaMemcpyAsync(reinterpret_cast<uint32_t*>(l->h_small + 3), l->d_nl.as<uint32_t>() + last_nl, 4, cudaMemcpyDeviceToHost, st));
    }
    CU(cudaStreamSynchronize(st));
    const uint64_t total = l->h_small[0];
    out->mapped = l->h_small[1];
    out->aligned = l->h_small[2];
    out->end_off = (uint64_t)*reinterpret_cast<uint32_t*>(l->h_small + 3) + 1;
    if (total > l->h_out_cap) {
        cudaFreeHost(l->h_out);
        l->h_out = nullptr;
        l->h_out_cap = total + total / 4 + 4096;
        CU(cudaHostAlloc((void**)&l->h_out, l->h_out_cap, cudaHostAllocDefault));
        out->out_text = l->h_out;
    }
    if ((rc = l->d_out.ensure(total + 64))) return rc;
    k_fq_format<<<nblocks(n, kFqFmtBlock), kFqFmtBlock, 0, st>>>(l->d_text.as<uint8_t>(), l->d_id_off.as<uint32_t>(), l->d_id_len.as<uint32_t>(),
                                                                 l->d_hits.as<HitRec>(), l->d_tx.as<uint32_t>(), l->d_line_off.as<uint64_t>(), n,
                                                                 l->d_out.as<char>());
    l->m->launches += 3;
    if (n_ticks) {
        CU(cudaMemcpyAsync(ds + 8, tick_at, n_ticks * 8, cudaMemcpyHostToDevice, st));
        k_fq_mapped_prefix<<<148, 256, 0, st>>>(l->d_hits.as<HitRec>(), reinterpret_cast<const uint64_t*>(ds + 8), n_ticks, ds + 16);
        CU(cudaMemcpyAsync(l->h_small + 8, ds + 16, n_ticks * 8, cudaMemcpyDeviceToHost, st));
        l->m->launches++;
    }
    CU(cudaMemcpyAsync(l->h_out, l->d_out.p, total, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    CU(cudaGetLastError());
    for (uint32_t t = 0; t < n_ticks; t++) out->tick_mapped[t] = l->h_small[8 + t];
    out->out_bytes = total;
    return PSA_OK;
}

// ---------------------------------------------------------------------------------------------
// memory helpers
// ---------------------------------------------------------------------------------------------
extern "C" int psa_host_alloc(void** out, uint64_t bytes) {
    if (!out) return fail(PSA_ERR_ARG, "null argument");
    CU(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault));
    return PSA_OK;
}
extern "C" void psa_host_free(void* p) {
    if (p) cudaFreeHost(p);
}
extern "C" int psa_device_alloc(void** out, uint64_t bytes) {
    if (!out) return fail(PSA_ERR_ARG, "null argument");
    CU(cudaMalloc(out, bytes ? bytes : 1));
    return PSA_OK;
}
extern "C" void psa_device_free(void* p) {
    if (p) cudaFree(p);
}
extern "C" int psa_memcpy_h2d(void* dst, const void* src, uint64_t bytes) {
    CU(cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice));
    return PSA_OK;
}
extern "C" int psa_memcpy_d2h(void* dst, const void* src, uint64_t bytes) {
    CU(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost));
    return PSA_OK;
}
