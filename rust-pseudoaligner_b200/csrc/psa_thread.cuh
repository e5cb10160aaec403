// psa_thread.cuh -- one thread = one read, blocking loads: map_read_nodes (psa_core.cuh) under a policy in
// which the calling thread does every step serially -- dictionary probe, node fetch, 32-base XOR compares,
// successor jump, the class intersection online (ClassAcc) -- and gives a read up ("defer") as soon as it
// needs something a single thread does badly: a seed search longer than max_probes positions, more wide
// classes than ClassAcc keeps, or only wide classes with a smallest one of more than max_small members.
// Deferred reads are redone from scratch by the cooperative kernels, so the split never changes a result.
// (A formulation cut at every load -- persistent warps over pools of reads in flight -- and one with a read per lane
// refilled as reads end were measured and rejected: profiles/r2_exp_walk_kernel.md.)
#pragma once
#include "psa_core.cuh"

namespace psa {

template <int KW, bool EV, class RD = PLoad>
struct ThreadCtx {
    const DevIndex& ix;
    RD rd;
    uint32_t max_probes;
    uint32_t reseed_probes;  // positions a re-seed search (ref :293) may try before the read is given up
    ClassAcc cls;
    uint32_t first_node;   // the node that brought the first class (its window is fetched only when a second class shows up)
    bool defer;
    uint32_t why;  // diagnostic: 0 first seed search, 1 re-seed search, 2 class list full, 3 smallest class too long
    bool seeded;
    // answer of the read's first seed search when k_seed_scan has already made it
    bool has_hint;
    uint32_t hint_pos, hint_node, hint_off;
    ThreadEvents ev;

    PSA_HD ThreadCtx(const DevIndex& ix_, RD rd_, uint32_t max_probes_, uint32_t reseed_probes_)
        : ix(ix_), rd(rd_), max_probes(max_probes_), reseed_probes(reseed_probes_), first_node(kNone), defer(false), why(0), seeded(false),
          has_hint(false), hint_pos(0), hint_node(0), hint_off(0), ev{} {
        cls.init();
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
            // re-seed searches (ref :293) have no scan kernel of their own: the first pass may hand the read to the
            // second one early (small reseed_probes: its warp is not held up by the search), which allows them more
            if (probes >= (seeded ? reseed_probes : max_probes)) {
                defer = true;
                why = seeded ? 1 : 0;
                kmer_pos = last + 1;  // keeps map_read_nodes out of the forward loop
                return false;
            }
            ProbeStats st;
            bool hit = dict_get<KW>(ix, KmerOps<KW>::get(rd, p, ix.k), node, o, EV ? &st : nullptr);
            if (EV) { ev.lookups++; ev.levels += st.levels; ev.hits += st.hit; ev.verifs += st.verified; }
            if (hit) {
                kmer_pos = p;
                seeded = true;
                return true;
            }
        }
    }
    PSA_HD NodeView node(uint32_t id) const { return load_node_view(ix.nodes + id); }
    PSA_HD void jumped() {
        if (EV) ev.jumps++;
    }
    PSA_HD uint32_t pred(uint32_t id, uint32_t b) {
#ifdef __CUDA_ARCH__
        const uint32_t p = __ldg(&ix.nodes_cold[id].pred[b]);
#else
        const uint32_t p = ix.nodes_cold[id].pred[b];
#endif
        if (EV && p != kNone) ev.jumps++;
        return p;
    }
    // ref src/pseudoaligner.rs:234-255 (FWD) and :149-170 (backward), 32 bases per step.  Forward, the mismatches of a
    // window are only counted; the bit reversal that puts them in base order is needed for the one window in which the
    // budget runs out.  (Carrying each word into the next step instead of loading it twice was measured slower: 2.49 vs
    // 2.36 ms per batch -- the second load is an L1 hit and the carried words cost registers.)
    template <bool FWD, class P>
    PSA_HD P cmp(P rp, uint64_t sp, P m, uint32_t A, bool& premature) {
        uint32_t snp = 0;
        for (P my = 0; my < m; my += 32) {
            const uint32_t n = m - my < 32 ? (uint32_t)(m - my) : 32u;
            uint64_t mask;
            uint32_t c;
            if (FWD) {
                mask = seq_bits(rd, rp + my, n) ^ seq_bits(GLoad{ix.seq}, sp + my, n);   // base t of the window at bits 2(n-1-t)
                c = (uint32_t)popc64(fold_pairs(mask));
            } else {
                mask = mismatch_bwd(rd, rp - my, GLoad{ix.seq}, sp - my, n);            // base t at bit 2t
                c = (uint32_t)popc64(mask);
            }
            if (snp + c > A) {
                premature = true;
                if (FWD) mask = fold_pairs(rev_pairs(mask) >> (64 - 2 * n));              // base t at bit 2t
                const P matched = my + nth_mismatch(mask, A + 1 - snp);
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
    // the window of a node's class: sector 1 of its record
    PSA_HD ClassWin node_window(uint32_t id) const {
        return class_win_of(load_sector_hot(reinterpret_cast<const char*>(ix.nodes + id) + 32));
    }
    // nodes.push: only the classes matter, and the intersection is idempotent (ref :352-355).  Until a second
    // distinct class shows up no window is fetched at all.
    PSA_HD void push(uint32_t node_id, const NodeView& nv) {
        if (EV) ev.visits++;
        if (nv.eq == cls.last_eq) return;
        if (EV) ev.members += nv.class_len;
        if (cls.min_eq == kNone) {              // the first class: remembered, not yet ANDed
            cls.last_eq = nv.eq;
            cls.min_len = nv.class_len;
            cls.min_eq = nv.eq;
            first_node = node_id;
            return;
        }
        if (first_node != kNone) {              // the second distinct class: the first one's window is due now
            const uint32_t fe = cls.min_eq, fl = cls.min_len;
            cls.and_class(ix, fe, fl, node_window(first_node));
            first_node = kNone;
        }
        cls.push(ix, nv.eq, nv.class_len, node_window(node_id));
        if (cls.defer()) {
            defer = true;
            why = 2;
        }
    }
};

// Result of map_read for one read as the kernels store it (flag = ref :453-462).
struct ThreadResult {
    HitRec hit;
    uint64_t count_slot;  // index into counts[]: eq id, n_eq (no visited class), n_eq + 1 (None)
    bool deferred;
    bool novel_overflow;
    uint32_t why;  // ThreadCtx::why when deferred
};

// map_read + the process_reads flag for one read, by one thread.  Sink::novel(r, count, off&) returns room
// for `count` members of a set that is no visited class (nullptr: no room).
template <int KW, bool EV, class Sink, class RD>
PSA_HD ThreadResult map_read_thread(const DevIndex& ix, RD words, uint32_t r, uint32_t L, uint32_t allowed,
                                    uint32_t max_probes, uint32_t reseed_probes, uint32_t max_small, Sink& sink, bool want_members,
                                    ThreadEvents* ev_out, const uint32_t* hint = nullptr /* pos, node, off */) {
    ThreadResult res;
    res.hit.coverage = 0; res.hit.n_tx = 0; res.hit.tx_off = 0; res.hit.eq_id = kNone; res.hit.flags = 0;
    res.count_slot = ix.n_eq + 1;
    res.deferred = false;
    res.novel_overflow = false;
    ThreadCtx<KW, EV, RD> w(ix, words, max_probes, reseed_probes);
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
        int s;
        if (!class_result(ix, w.cls, max_small, count, eq_id, s)) {
            res.why = 3;
            res.deferred = true;
            return res;
        }
        res.hit.coverage = coverage;
        res.hit.n_tx = count;
        res.hit.eq_id = eq_id;
        res.hit.flags = kFlagAligned | ((coverage >= kCoverageThreshold && count == 0) ? kFlagMapped : 0u);  // ref :455 (sic)
        if (eq_id != kNone) {
            res.count_slot = eq_id;   // (members: k_expand reads them from the index through eq_id)
        } else {
            res.count_slot = ix.n_eq;
            if (want_members) {       // (the empty set too: it is listed and counted like any other)
                uint64_t o = 0;
                uint32_t* dst = sink.novel(r, count, o);
                if (!dst) res.novel_overflow = true;
                else if (count) class_members(ix, w.cls, s, dst);
                res.hit.tx_off = o;
            }
        }
    }
    if (EV && ev_out) *ev_out = w.ev;
    return res;
}

}  // namespace psa
