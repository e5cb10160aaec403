// hostsim.cpp -- UNIT-TEST HARNESS, not product code.
//
// Compiles rust-pseudoaligner_b200/csrc/psa_core.cuh and psa_thread.cuh (the __host__ __device__
// arithmetic the CUDA kernels are made of: 2-bit access, k-mer hash, the bucket-cascade dictionary,
// fingerprint + verification, per-word mismatch masks, class windows, the map_read state machine under
// the cooperative policy and under the thread-per-read policy with its hand-overs) with g++ and
// serial drivers, so that the tests that run without a GPU can compare that text against the oracle.
// It is built into tests/hostsim/libhostsim.so, loaded only by tests/test_hostsim.py, and never
// linked into libpsa_b200.so -- the product has no CPU path.
// The dictionary here is built serially on the host with the same layout rules the device builder
// uses (level size = floor(gamma*n/4)+1 buckets, keys in enumeration order, at most four per bucket
// and no two with one fingerprint, the rest passed on), so the probe code sees a structure of
// identical shape.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "../../rust-pseudoaligner_b200/csrc/psa_core.cuh"
#include "../../rust-pseudoaligner_b200/csrc/psa_thread.cuh"
#include "../../rust-pseudoaligner_b200/csrc/psa_thread.cuh"

using namespace psa;

struct HsIndex {
    DevIndex d{};
    std::vector<uint64_t> buckets, seq, eq_off;
    std::vector<NodeRec> nodes;
    std::vector<NodeCold> cold;
    std::vector<uint32_t> eq_mem;
    std::vector<ClassWin> class_win;
    int kw = 1;
    int error = 0;
};

struct HsHit {  // == psa_hit
    uint32_t coverage, n_tx;
    uint64_t tx_off;
    uint32_t eq_id, flags;
};

template <int KW>
static void build(HsIndex* ix, uint32_t k, uint64_t n_nodes, const uint64_t* node_start, const uint32_t* node_len,
                  const uint8_t* node_exts, const uint32_t* node_eq, double gamma) {
    DevIndex& D = ix->d;
    std::vector<Kmer<KW>> keys;
    std::vector<uint64_t> vals;
    uint64_t max_pos = 0;
    for (uint64_t i = 0; i < n_nodes; i++) {
        uint64_t nk = node_len[i] - k + 1;
        max_pos = std::max(max_pos, node_start[i] + nk - 1);
        for (uint64_t o = 0; o < nk; o++) {
            keys.push_back(KmerOps<KW>::get(PLoad{ix->seq.data()}, node_start[i] + o, k));
            vals.push_back((i << 32) | o);
        }
    }
    const uint64_t n_kmers = keys.size();
    auto bits_for = [](uint64_t v) { uint32_t b = 1; while (b < 64 && (v >> b)) b++; return b; };
    D.k = k;
    D.node_bits = bits_for(n_nodes ? n_nodes - 1 : 0);
    D.pos_bits = bits_for(max_pos + 1);
    int fp = 63 - (int)D.node_bits - (int)D.pos_bits;
    if (fp < 8) { ix->error = 5; return; }
    D.fp_bits = (uint32_t)std::min(fp, 32);
    D.n_nodes = n_nodes;
    D.n_kmers = n_kmers;

    // nodes (the dictionary entries need their starts)
    ix->nodes.resize(n_nodes + 1);
    ix->cold.resize(n_nodes + 1);
    for (uint64_t i = 0; i < n_nodes; i++) {
        NodeRec& r = ix->nodes[i];
        NodeCold& c = ix->cold[i];
        r.start_len = pack_start_len(node_start[i], node_len[i]); r.eq = node_eq[i]; c.exts = node_exts[i]; c.pad = 0;
        r.class_len = (uint32_t)(ix->eq_off[r.eq + 1] - ix->eq_off[r.eq]);
        c.class_off = ix->eq_off[r.eq];
        for (int b = 0; b < 4; b++) r.succ[b] = c.pred[b] = kNone;
        const ClassWin& w = ix->class_win[r.eq];
        r.win_lo = w.lo; r.win_len = w.len; r.win_bits[0] = w.bits[0]; r.win_bits[1] = w.bits[1]; r.win_bits[2] = w.bits[2];
    }
    D.nodes = ix->nodes.data();
    D.nodes_cold = ix->cold.data();

    // cascade
    std::vector<uint64_t> cur(n_kmers);
    for (uint64_t i = 0; i < n_kmers; i++) cur[i] = i;
    uint64_t total_bkt = 0;
    uint32_t lvl = 0;
    while (!cur.empty()) {
        if (lvl >= (uint32_t)kMaxLevels) { ix->error = 1; return; }
        const uint64_t nbkt = std::max<uint64_t>(1, (uint64_t)(gamma * (double)cur.size() / kBucketSlots) + 1);
        ix->buckets.resize(4 * (total_bkt + nbkt), kEmptyEntry);
        std::vector<uint8_t> cnt(nbkt, 0);
        std::vector<uint64_t> next;
        for (uint64_t i : cur) {
            const KeyHash hk = make_hash(KmerOps<KW>::fold(keys[i]));
            const uint64_t b = level_bucket(hk, lvl, nbkt);
            uint64_t* e = &ix->buckets[4 * (total_bkt + b)];
            const uint64_t fpv = fp_of(hk, D.fp_bits);
            bool clash = cnt[b] >= kBucketSlots;
            for (uint32_t q = 0; q < cnt[b]; q++)
                clash |= ((e[q] & ~kMoreBit) >> (D.node_bits + D.pos_bits)) == fpv;
            if (clash) {
                next.push_back(i);
                e[0] |= kMoreBit;  // (cnt[b] >= 1 here)
                continue;
            }
            const uint32_t node = (uint32_t)(vals[i] >> 32);
            e[cnt[b]] = pack_entry(D, node, node_start[node] + (uint32_t)vals[i], hk);
            cnt[b]++;
        }
        D.dict.level_nbkt[lvl] = nbkt;
        D.dict.level_base[lvl] = total_bkt;
        total_bkt += nbkt;
        lvl++;
        cur.swap(next);
    }
    D.dict.n_levels = lvl;
    D.dict.buckets = ix->buckets.data();
    for (uint64_t i = 0; i < n_kmers; i++) {
        uint32_t n = 0, o = 0;
        if (!dict_get<KW>(D, keys[i], n, o, nullptr) || n != (uint32_t)(vals[i] >> 32) || o != (uint32_t)vals[i]) { ix->error = 3; return; }
    }
    // edges
    for (uint64_t i = 0; i < n_nodes; i++) {
        NodeRec& r = ix->nodes[i];
        NodeCold& c = ix->cold[i];
        Kmer<KW> first = KmerOps<KW>::get(PLoad{ix->seq.data()}, node_start[i], k);
        Kmer<KW> last = KmerOps<KW>::get(PLoad{ix->seq.data()}, node_start[i] + node_len[i] - k, k);
        for (uint32_t b = 0; b < 4; b++) {
            uint32_t n, o;
            if ((c.exts >> b) & 1) {
                if (dict_get<KW>(D, KmerOps<KW>::extend_right(last, b, k), n, o, nullptr) && o == 0) r.succ[b] = n;
                else ix->error = 4;
            }
            if ((c.exts >> (4 + b)) & 1) {
                if (dict_get<KW>(D, KmerOps<KW>::extend_left(first, b, k), n, o, nullptr) && o == node_len[n] - k) c.pred[b] = n;
                else ix->error = 4;
            }
        }
    }
}

// serial stand-in for the warp-cooperative steps: same per-lane functions, lanes looped
template <int KW>
struct SerialWarp {
    const DevIndex& ix;
    PLoad rd;
    uint32_t k;
    std::vector<uint32_t> eqs, lens;  // distinct visited classes
    uint64_t lookups = 0;

    template <class P>
    bool find_seed(P& kmer_pos, P last, uint32_t& node, uint32_t& off) {
        if (kmer_pos > last) return false;
        const P start = kmer_pos;
        // lane order == position order, so the first hitting lane is the sequential first hit
        for (P p = start; p <= last; p += kSeedStride) {
            lookups++;
            if (dict_get<KW>(ix, KmerOps<KW>::get(rd, p, k), node, off, nullptr)) { kmer_pos = p; return true; }
        }
        kmer_pos = start + kSeedStride * ((last - start) / kSeedStride + 1);
        return false;
    }
    uint32_t read_base(uint64_t pos) const { return seq_get(rd, pos); }
    bool abort() const { return false; }
    NodeView node(uint32_t id) const { return load_node_view(ix.nodes + id); }
    void jumped() {}
    uint32_t pred(uint32_t id, uint32_t b) { return ix.nodes_cold[id].pred[b]; }
    // the two compare loops, chunked by 32 bases exactly as the kernel lanes are
    template <bool FWD, class P>
    P cmp(P rp, uint64_t sp, P m, uint32_t A, bool& premature) {
        uint32_t snp = 0;
        for (P my = 0; my < m; my += 32) {
            uint32_t n = (uint32_t)std::min<P>(32, m - my);
            uint64_t mask = FWD ? mismatch_fwd(rd, rp + my, GLoad{ix.seq}, sp + my, n)
                                : mismatch_bwd(rd, rp - my, GLoad{ix.seq}, sp - my, n);
            uint32_t c = (uint32_t)popc64(mask);
            if (snp + c > A) {
                premature = true;
                return my + nth_mismatch(mask, A + 1 - snp);
            }
            snp += c;
        }
        return m;
    }
    template <class P>
    P cmp_fwd(P rp, uint64_t sp, P m, uint32_t A, bool& pb) { return cmp<true>(rp, sp, m, A, pb); }
    template <class P>
    P cmp_bwd(P rp, uint64_t sp, P m, uint32_t A, bool& pb) { return cmp<false>(rp, sp, m, A, pb); }
    void push(uint32_t, const NodeView& nv) {
        if (std::find(eqs.begin(), eqs.end(), nv.eq) != eqs.end()) return;
        eqs.push_back(nv.eq);
        lens.push_back(nv.class_len);
    }
};

template <int KW>
static void map_one(const HsIndex* ix, const uint64_t* words, uint32_t L, uint32_t allowed, HsHit& h,
                    std::vector<uint32_t>& tx) {
    SerialWarp<KW> w{ix->d, PLoad{words}, ix->d.k, {}, {}};
    uint32_t cov = 0;
    h.coverage = 0; h.n_tx = 0; h.tx_off = tx.size(); h.eq_id = kNone; h.flags = 0;
    if (!map_read_nodes(w, ix->d.k, (uint64_t)L, allowed, cov)) return;
    // smallest class, then keep its members found in every other class
    size_t s = 0;
    for (size_t j = 1; j < w.eqs.size(); j++)
        if (w.lens[j] < w.lens[s] || (w.lens[j] == w.lens[s] && w.eqs[j] < w.eqs[s])) s = j;
    const uint32_t* sm = ix->eq_mem.data() + ix->eq_off[w.eqs[s]];
    uint32_t count = 0;
    for (uint32_t i = 0; i < w.lens[s]; i++) {
        bool alive = true;
        for (size_t j = 0; j < w.eqs.size() && alive; j++)
            if (j != s) alive = contains_sorted(ix->eq_mem.data() + ix->eq_off[w.eqs[j]], (uint64_t)w.lens[j], sm[i]);
        if (alive) { tx.push_back(sm[i]); count++; }
    }
    uint32_t eq_id = kNone;
    for (size_t j = 0; j < w.eqs.size(); j++)
        if (w.lens[j] == count && w.eqs[j] < eq_id) eq_id = w.eqs[j];
    h.coverage = cov; h.n_tx = count; h.eq_id = eq_id;
    h.flags = 1u | ((cov >= kCoverageThreshold && count == 0) ? 2u : 0u);
}

extern "C" {

HsIndex* hs_index_create(uint32_t k, uint64_t n_nodes, const uint64_t* seq_words, uint64_t n_seq_words,
                         const uint64_t* node_start, const uint32_t* node_len, const uint8_t* node_exts,
                         const uint32_t* node_eq, uint64_t n_eq, const uint64_t* eq_offsets,
                         const uint32_t* eq_members, double gamma) {
    HsIndex* ix = new HsIndex();
    ix->seq.assign(seq_words, seq_words + n_seq_words);
    for (int i = 0; i < 8; i++) ix->seq.push_back(0);
    ix->eq_off.assign(eq_offsets, eq_offsets + n_eq + 1);
    ix->eq_mem.assign(eq_members, eq_members + eq_offsets[n_eq]);
    ix->d.seq = ix->seq.data();
    ix->d.eq_off = ix->eq_off.data();
    ix->d.eq_mem = ix->eq_mem.data();
    ix->d.n_eq = n_eq;
    ix->class_win.resize(n_eq + 1);
    for (uint64_t c = 0; c < n_eq; c++)
        ix->class_win[c] = make_class_win(ix->eq_mem.data() + eq_offsets[c], eq_offsets[c + 1] - eq_offsets[c]);
    ix->d.class_win = ix->class_win.data();
    ix->kw = k <= 32 ? 1 : 2;
    if (gamma <= 0) gamma = 1.7;
    if (ix->kw == 1) build<1>(ix, k, n_nodes, node_start, node_len, node_exts, node_eq, gamma);
    else build<2>(ix, k, n_nodes, node_start, node_len, node_exts, node_eq, gamma);
    return ix;
}
int hs_index_error(const HsIndex* ix) { return ix->error; }
uint32_t hs_index_levels(const HsIndex* ix) { return ix->d.dict.n_levels; }
uint32_t hs_index_fp_bits(const HsIndex* ix) { return ix->d.fp_bits; }
void hs_index_destroy(HsIndex* ix) { delete ix; }

int hs_lookup(const HsIndex* ix, const uint64_t* kmer_words, uint32_t* node, uint32_t* off) {
    if (ix->kw == 1) return dict_get<1>(ix->d, KmerOps<1>::get(PLoad{kmer_words}, 0, ix->d.k), *node, *off, nullptr);
    return dict_get<2>(ix->d, KmerOps<2>::get(PLoad{kmer_words}, 0, ix->d.k), *node, *off, nullptr);
}

// returns the number of tx entries needed; fills up to tx_cap
uint64_t hs_map_batch(const HsIndex* ix, const uint64_t* words, const uint64_t* read_off, const uint32_t* read_len,
                      uint64_t n, uint32_t allowed, HsHit* hits, uint32_t* tx_buf, uint64_t tx_cap) {
    std::vector<uint32_t> tx;
    for (uint64_t i = 0; i < n; i++) {
        if (ix->kw == 1) map_one<1>(ix, words + read_off[i], read_len[i], allowed, hits[i], tx);
        else map_one<2>(ix, words + read_off[i], read_len[i], allowed, hits[i], tx);
    }
    memcpy(tx_buf, tx.data(), std::min<uint64_t>(tx.size(), tx_cap) * 4);
    return tx.size();
}

}  // extern "C"

// what the thread-per-read policy delivers a finished read to
struct HostSink {
    HsHit* hit;
    std::vector<uint32_t> novel_buf;
    bool got = false;
    void result(uint32_t, const HitRec& h, uint64_t) {
        hit->coverage = h.coverage; hit->n_tx = h.n_tx; hit->tx_off = h.tx_off; hit->eq_id = h.eq_id; hit->flags = h.flags;
        got = true;
    }
    uint32_t* novel(uint32_t, uint32_t count, uint64_t& off) {
        off = 0;
        novel_buf.assign(count, 0);
        return novel_buf.data();
    }
    void novel_overflow() {}
};

extern "C" {
// The blocking thread-per-read policy (psa_thread.cuh ThreadCtx / map_read_thread), hand-overs redone by the
// serial stand-in of the cooperative kernel.
uint64_t hs_map_batch_thread(const HsIndex* ix, const uint64_t* words, const uint64_t* read_off,
                             const uint32_t* read_len, uint64_t n, uint32_t allowed, uint32_t max_probes,
                             uint32_t max_small, HsHit* hits, uint32_t* tx_buf, uint64_t tx_cap, uint64_t* n_deferred) {
    std::vector<uint32_t> tx;
    uint64_t nd = 0;
    for (uint64_t i = 0; i < n; i++) {
        HostSink sink{&hits[i], {}};
        ThreadResult r;
        // first pass: a small re-seed budget (0, 1 or 2 positions, varied with max_probes so that the tests see all three);
        // a read that runs out of it is redone by the second pass with the long budget, as the kernels do
        const uint32_t reseed1 = max_probes % 3, reseed2 = max_probes > kReseedProbes ? max_probes : kReseedProbes;
        for (int pass = 0; pass < 2; pass++) {
            sink.novel_buf.clear();
            const uint32_t rs = pass == 0 ? reseed1 : reseed2;
            if (ix->kw == 1) r = map_read_thread<1, false>(ix->d, PLoad{words + read_off[i]}, (uint32_t)i, read_len[i], allowed, max_probes, rs, max_small, sink, true, nullptr);
            else r = map_read_thread<2, false>(ix->d, PLoad{words + read_off[i]}, (uint32_t)i, read_len[i], allowed, max_probes, rs, max_small, sink, true, nullptr);
            if (!(r.deferred && r.why == 1)) break;
        }
        if (r.deferred) {
            nd++;
            if (ix->kw == 1) map_one<1>(ix, words + read_off[i], read_len[i], allowed, hits[i], tx);
            else map_one<2>(ix, words + read_off[i], read_len[i], allowed, hits[i], tx);
            continue;
        }
        HsHit& h = hits[i];
        h.coverage = r.hit.coverage; h.n_tx = r.hit.n_tx; h.eq_id = r.hit.eq_id; h.flags = r.hit.flags;
        h.tx_off = tx.size();
        const uint32_t* src = r.hit.eq_id != kNone ? ix->eq_mem.data() + ix->eq_off[r.hit.eq_id] : sink.novel_buf.data();
        tx.insert(tx.end(), src, src + r.hit.n_tx);
    }
    if (n_deferred) *n_deferred = nd;
    memcpy(tx_buf, tx.data(), std::min<uint64_t>(tx.size(), tx_cap) * 4);
    return tx.size();
}

// ASCII -> DnaString words with the product's base_code()
void hs_pack_ascii(const uint8_t* s, uint64_t len, uint64_t* words) {
    for (uint64_t j = 0; j < (len + 31) / 32; j++) {
        uint64_t v = 0;
        for (uint64_t t = 0; t < 32 && 32 * j + t < len; t++) v |= (uint64_t)base_code(s[32 * j + t]) << (62 - 2 * t);
        words[j] = v;
    }
}
}
