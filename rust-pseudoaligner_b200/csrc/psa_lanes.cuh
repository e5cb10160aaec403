// psa_lanes.cuh -- map_read as a per-lane state machine: the thread-per-read kernel's text.
//
// map_read_to_nodes_with_mismatch (ref src/pseudoaligner.rs:64-319) is a chain of dependent
// gathers: dictionary bucket -> unitig k-mer + node record -> unitig words -> successor's node
// record -> ...  Written as one call per read (round 1), the 32 reads of a warp drift apart at the
// first data-dependent branch and from then on mostly wait for each other at reconvergence points:
// ncu saw 6.9 active threads per instruction and an effective memory-level parallelism of ~4 per
// warp.  Here the same control flow is cut at every load: a lane carries an explicit state, every
// state ends by REQUESTING the sectors the next one needs, and the kernel's loop issues the
// requests of all 32 lanes with the same (converged) load instructions before any lane consumes
// its data.  A lane that finishes its read takes the next one from a counter in the same loop --
// no lane waits for the slowest read of its warp.
//
// The blocks of step() below are the reference's statements in the reference's order (line numbers
// in the margin; compare psa_core.cuh map_read_nodes, the blocking form of the same text, which the
// cooperative kernels and the host simulation run).  A lane moves from block to block within one
// call as long as it needs no new data (`wait` is set by every request).
//
// Payloads: A = one sector (dictionary bucket | node record sector 0 | NodeCold), B = node record
// sector 1 (the class window), C = four consecutive words of unitig sequence.  Which addresses a
// state needs is a function of the lane's fields (requests()); LS_READ alone uses A and C as the
// eight words of a new read.  The lane record is 136 bytes: the kernel keeps one per read in
// flight in shared memory.
#pragma once
#include "psa_core.cuh"

namespace psa {

enum : uint32_t {
    LS_NEW = 0,   // needs a read (the kernel's loop assigns one)
    LS_READ,      // payload: the read's packed words
    LS_BUCKET,    // payload A: dictionary bucket of the probed k-mer
    LS_VERIFY,    // payload A, B: node record of the candidate, C: unitig words at the candidate position
    LS_NODE,      // payload A, B: node record of node_id (forward loop body, ref :209-231)
    LS_LPRED,     // payload A: NodeCold of prev_node (left extension, ref :183-194)
    LS_LNODE,     // payload A, B: node record of the predecessor (ref :195-199)
    LS_LCMP,      // payload C: unitig words of the backward compare (ref :151-170)
    LS_LAFTER,    // (no payload) ref :173-202
    LS_CMP,       // payload C: unitig words of the forward compare (ref :236-255)
    LS_AFTER,     // (no payload) ref :257-300
    LS_PROBE,     // (no payload) next position of find_kmer_match (ref :92-111)
    LS_FIN,       // (no payload) nodes_to_eq_class + the process_reads flag (ref :323-356, :453-462)
    LS_TOOLONG,   // read longer than the lane's shared-memory slot: handed to the cooperative kernel
    LS_IDLE       // no reads left
};
constexpr uint32_t LF_SEEDED = 1u;     // a seed has been found for this read (re-seed searches get a larger budget)
constexpr uint32_t LF_FIRST = 2u;      // the read's first seed search is running
constexpr uint32_t LF_FRESH = 4u;      // node_id is the first search's seed: left-extension test pending (ref :124-126)
constexpr uint32_t LF_PUSHED = 8u;     // nodes is not empty
constexpr uint32_t LF_PREMATURE = 16u; // premature_break of the running compare
constexpr uint32_t LF_MORE = 32u;      // the candidate's bucket has keys in later levels
constexpr uint32_t LF_HINT = 64u;      // the first search's answer is given (k_seed_scan made it)

// what a step produced for the kernel's loop
constexpr uint64_t kNoWords = 1ULL << 62;  // c_base when C holds no unitig words (no word index comes near it)
constexpr uint32_t LE_NONE = 0, LE_RESULT = 1, LE_TO_SCAN = 2, LE_TO_COOP = 3;

struct LaneParams {
    uint32_t allowed;       // allowed_mismatches (ref :361)
    uint32_t max_probes;    // seed positions one lane tries per search before handing the read over
    uint32_t max_small;     // largest smallest-class one lane intersects by list
    bool want_members;      // members of sets that are no visited class are wanted
    bool to_scan;           // a too long FIRST search goes to the seed-scan kernel (else: cooperative kernel)
};

struct LaneRequests {
    const void* a;        // one sector (nullptr: none)
    const void* b;        // one sector
    const uint64_t* c;    // four words
    bool a_stream;        // a is a dictionary bucket (one use: evict-first)
};

template <bool EV>
struct LaneEv {};
template <>
struct LaneEv<true> {
    ThreadEvents ev;  // sequential-equivalent events of the read in flight
};
struct StepOut {
    uint32_t emit;            // LE_*
    uint32_t n_tx, aligned;   // with LE_RESULT
};

template <int KW, bool EV>
struct Lane : LaneEv<EV> {
    uint32_t st;
    uint32_t r, L, flags;
    uint32_t kmer_pos, cov, node_id, kmer_offset;     // the reference's variables of the same names
    uint32_t cand;          // forward: successor taken if the whole span matches; left: the base selecting the predecessor
    uint32_t why;           // when a read is handed over: 0 first seed search, 1 re-seed search, 2 class list full,
                            // 3 smallest class too long, 4 read too long
    uint64_t cm_s;          // unitig position of the running compare / of the candidate k-mer being verified
    union {
        struct {            // find_kmer_match
            uint32_t f_start, f_p, f_probes, f_lvl;
            KeyHash hk;
        };
        struct {            // the running compare: t-th base = read[cm_r +- t] vs seq[cm_s +- t]; left extension
            uint32_t cm_r, cm_left, cm_done, snp;
            uint32_t last_pos, prev_node, prev_off, pad_;
        };
    };
    ClassAcc cls;

    PSA_HD void idle() {
        st = LS_NEW; why = 0;
        r = 0; L = 0; flags = 0;
    }
    // a new read for this lane; hint = {pos, node, off} of its first seed or nullptr
    PSA_HD void begin(uint32_t r_, uint32_t L_, uint32_t slot_words, const uint32_t* hint) {
        r = r_; L = L_;
        flags = hint ? LF_HINT : 0u;
        if (hint) { kmer_pos = hint[0]; node_id = hint[1]; kmer_offset = hint[2]; }
        st = ((L + 31) >> 5) > slot_words ? LS_TOOLONG : LS_READ;
    }

    // first word of the four the backward compare fetches
    static PSA_HD uint64_t left_base(uint64_t pos) {
        const uint64_t w = pos >> 5;
        return w >= 3 ? w - 3 : 0;
    }
    // what the lane's state needs before its next step (LS_READ: the read's words, fetched by the caller)
    PSA_HD LaneRequests requests(const DevIndex& ix) const {
        LaneRequests q;
        q.a = nullptr; q.b = nullptr; q.c = nullptr; q.a_stream = false;
        if (st == LS_BUCKET) {
            q.a = bucket_addr(ix.dict, hk, f_lvl);
            q.a_stream = true;
        } else if (st == LS_VERIFY || st == LS_NODE) {
            q.a = ix.nodes + node_id;
            q.b = reinterpret_cast<const char*>(ix.nodes + node_id) + 32;
            if (st == LS_VERIFY) q.c = ix.seq + (cm_s >> 5);
        } else if (st == LS_LNODE) {
            q.a = ix.nodes + prev_node;
            q.b = reinterpret_cast<const char*>(ix.nodes + prev_node) + 32;
        } else if (st == LS_LPRED) {
            q.a = ix.nodes_cold + prev_node;
        } else if (st == LS_CMP) {
            q.c = ix.seq + (cm_s >> 5);
        } else if (st == LS_LCMP) {
            q.c = ix.seq + left_base(cm_s);
        }
        return q;
    }

    template <class RW>
    PSA_HD uint32_t read_base(const RW& rw, uint32_t pos) const { return seq_get(rw, pos); }

    // ref :139-150: the span the backward compare may cover in the node that starts at `start`
    PSA_HD void left_span(uint64_t start) {
        const uint32_t skipped_read = last_pos + 1;                                     // :139
        const uint32_t skipped_ref = prev_off + 1;                                      // :142
        cm_left = skipped_read < skipped_ref ? skipped_read : skipped_ref;              // :145
        cm_r = last_pos; cm_s = start + prev_off; cm_done = 0; snp = 0;                 // :148-150
        flags &= ~LF_PREMATURE;
        st = LS_LCMP;
    }

    // One step.  A, B, C: the payloads requests() named before the call.  rw: the lane's read words
    // (rw(i) loads word i, rw.store(i, v) stores it).  sink: result(r, hit, count_slot),
    // novel(r, count, off&) -> room for the members of a set that is no visited class (or nullptr),
    // novel_overflow().
    template <class RW, class Sink>
    PSA_HD StepOut step(const DevIndex& ix, const LaneParams& lp, RW& rw, const Sector& A, const Sector& B, const Sector& C,
                        Sink& sink) {
        StepOut out;
        out.emit = LE_NONE; out.n_tx = 0; out.aligned = 0;
        bool wait = false;  // the lane's next block needs data it does not have: no further block this step
        const uint32_t k = ix.k;
        // the first of the four words C holds
        uint64_t c_base = (st == LS_VERIFY || st == LS_CMP) ? cm_s >> 5 : st == LS_LCMP ? left_base(cm_s) : kNoWords;

        if (st == LS_TOOLONG) {
            why = 4; out.emit = LE_TO_COOP; st = LS_NEW;
            return out;
        }
        if (st == LS_READ) {
            const uint32_t nw = (L + 31) >> 5;
            if (nw > 0) rw.store(0, A.w0);
            if (nw > 1) rw.store(1, A.w1);
            if (nw > 2) rw.store(2, A.w2);
            if (nw > 3) rw.store(3, A.w3);
            if (nw > 4) rw.store(4, C.w0);
            if (nw > 5) rw.store(5, C.w1);
            if (nw > 6) rw.store(6, C.w2);
            if (nw > 7) rw.store(7, C.w3);
            cov = 0;                                                                   // :71
            if (!(flags & LF_HINT)) kmer_pos = 0;                                      // :79
            cls.init();                                                                // :75
            if constexpr (EV) this->ev = ThreadEvents{};
            if (L < k) {                                                               // :82-84
                flags = 0;
                st = LS_FIN;
            } else if (flags & LF_HINT) {                                              // :118-121, answer given (begin())
                flags = LF_SEEDED | LF_FRESH;
                wait = true;
                st = LS_NODE;
            } else {
                flags = LF_FIRST;
                f_start = 0; f_p = 0; f_probes = 0;
                st = LS_PROBE;
            }
        }

        // ---- dbg_index.get (ref :96): one bucket of the cascade
        if (st == LS_BUCKET && !wait) {
            if constexpr (EV) this->ev.levels++;
            const uint64_t e = bucket_find(ix, A, hk);
            const bool more = bucket_more(A) && f_lvl + 1 < ix.dict.n_levels;
            if (e != kEmptyEntry) {
                if constexpr (EV) { this->ev.hits++; this->ev.verifs++; }
                node_id = entry_node(ix, e);                      // (the previous node is done with: find runs between nodes)
                cm_s = entry_pos(ix, e);
                flags = more ? (flags | LF_MORE) : (flags & ~LF_MORE);
                wait = true;
                st = LS_VERIFY;
            } else if (more) {
                f_lvl++;
                wait = true;
            } else {
                f_probes++; f_p += kSeedStride;                                        // :110
                st = LS_PROBE;
            }
        }
        // ---- ref :99-107: the answer is verified against the unitig
        if (st == LS_VERIFY && !wait) {
            const Kmer<KW> key = KmerOps<KW>::get(rw, f_p, k);                         // :93
            const Kmer<KW> ref = KmerOps<KW>::get(WLoad{C, c_base}, cm_s, k);          // :101-103
            if (ref == key) {                                                          // :105-107
                kmer_offset = (uint32_t)(cm_s - (A.w0 & kStartMask));
                kmer_pos = f_p;
                flags |= LF_SEEDED;
                if (flags & LF_FIRST) flags |= LF_FRESH;
                flags &= ~LF_FIRST;
                st = LS_NODE;                                                          // A, B are node_id's record
            } else if (flags & LF_MORE) {
                f_lvl++;
                wait = true;
                st = LS_BUCKET;
            } else {
                f_probes++; f_p += kSeedStride;                                        // :110
                st = LS_PROBE;
            }
        }
        // ---- forward loop body up to the compare (ref :209-235); the left-extension test first (:124-126)
        if (st == LS_NODE && !wait) {
            const NodeView nv = node_view_of(A);                                       // :210
            bool left = false;
            if (flags & LF_FRESH) {
                flags &= ~LF_FRESH;
                const uint32_t thr = (uint32_t)(kLeftExtendFraction * (double)L);      // :77
                if (kmer_pos >= thr && kmer_pos >= 1) {                                // :124-126
                    last_pos = kmer_pos - 1;                                           // :127
                    prev_node = node_id;                                               // :128
                    prev_off = kmer_offset > 0 ? kmer_offset - 1 : 0;                  // :129 (sic)
                    left_span(nv.start);                                               // :132-150
                    wait = true;
                    left = true;
                }
            }
            if (!left) {
                kmer_pos += k;                                                         // :215
                cov += k;                                                              // :216
                if constexpr (EV) this->ev.visits++;
                const bool fresh = cls.push(ix, nv.eq, nv.class_len, class_win_of(B)); // :219
                if constexpr (EV) if (fresh) this->ev.members += nv.class_len;
                flags |= LF_PUSHED;
                if (cls.defer()) {
                    why = 2; out.emit = LE_TO_COOP; st = LS_NEW;
                    return out;
                }
                const uint32_t remaining_read = L - kmer_pos;                          // :222
                const uint32_t ref_offset = kmer_offset + k;                           // :227
                const uint32_t informative_ref = nv.len - ref_offset;                  // :228
                cm_left = remaining_read < informative_ref ? remaining_read : informative_ref;  // :231
                // the successor the walk takes if the rest of this unitig matches (:265-275): its selecting
                // base is known now, the node view is not kept
                cand = remaining_read > informative_ref ? view_succ(nv, read_base(rw, kmer_pos + informative_ref)) : kNone;
                cm_r = kmer_pos; cm_s = nv.start + ref_offset; cm_done = 0; snp = 0;   // :233-235
                flags &= ~LF_PREMATURE;
                if (cm_left == 0) {
                    st = LS_AFTER;
                } else {
                    st = LS_CMP;
                    if ((cm_s >> 5) - c_base >= 4) wait = true;                        // not in C (it is, right after a verification)
                }
            }
        }
        // ---- left extension (ref :131-202)
        if (st == LS_LPRED && !wait) {
            const uint32_t pred_id = cand == 0 ? (uint32_t)A.w0 : cand == 1 ? (uint32_t)(A.w0 >> 32)
                                   : cand == 2 ? (uint32_t)A.w1 : (uint32_t)(A.w1 >> 32);
            if (pred_id != kNone) {                                                    // :183
                if constexpr (EV) this->ev.jumps++;
                prev_node = pred_id;                                                   // :185-194
                st = LS_LNODE;
            } else {                                                                   // :201
                st = LS_NODE;                                                          // on to the forward search
            }
            wait = true;
        }
        if (st == LS_LNODE && !wait) {
            const NodeView pv = node_view_of(A);                                       // :195
            prev_off = pv.len - k;                                                     // :196
            if constexpr (EV) this->ev.visits++;
            const bool fresh = cls.push(ix, pv.eq, pv.class_len, class_win_of(B));     // :199
            if constexpr (EV) if (fresh) this->ev.members += pv.class_len;
            flags |= LF_PUSHED;
            if (cls.defer()) {
                why = 2; out.emit = LE_TO_COOP; st = LS_NEW;
                return out;
            }
            left_span(pv.start);                                                       // :132-150 of the next round
            wait = true;
        }
        if (st == LS_LCMP && !wait) {                                                  // :151-170, up to 128 bases
            const uint32_t e0 = (uint32_t)(cm_s - (c_base << 5));  // the first compared base within C's 128 bases
            const uint32_t n = cm_left < e0 + 1 ? cm_left : e0 + 1;
            uint32_t matched_here = n;
            bool prem = false;
            PSA_UNROLL
            for (int j = 3; j >= 0; j--) {
                const uint32_t wlo = 32u * (uint32_t)j;
                const uint32_t lo = e0 + 1 - n > wlo ? e0 + 1 - n : wlo;
                const uint32_t hi = e0 + 1 < wlo + 32 ? e0 + 1 : wlo + 32;
                if (lo < hi && !prem) {
                    const uint32_t nn = hi - lo;
                    const uint64_t cw = j == 0 ? C.w0 : j == 1 ? C.w1 : j == 2 ? C.w2 : C.w3;
                    const uint64_t refb = (cw << (2 * (lo - wlo))) >> (64 - 2 * nn);
                    const uint64_t rdb = seq_bits(rw, cm_r - (e0 - lo), nn);
                    const uint64_t mask = fold_pairs(rdb ^ refb);      // t-th scanned base of the chunk (t = 0: hi - 1) at bit 2t
                    const uint32_t c = (uint32_t)popc64(mask);
                    if (snp + c > lp.allowed) {                                        // :161-165
                        prem = true;
                        matched_here = (e0 - (hi - 1)) + nth_mismatch(mask, lp.allowed + 1 - snp);
                    } else {
                        snp += c;
                    }
                }
            }
            if (prem) {
                cm_done += matched_here;
                if constexpr (EV) this->ev.bases += matched_here + 1;
                flags |= LF_PREMATURE;
                st = LS_LAFTER;
            } else {
                cm_done += n; cm_left -= n; cm_r -= n; cm_s -= n;
                if constexpr (EV) this->ev.bases += n;
                if (cm_left == 0) st = LS_LAFTER;
                else wait = true;
            }
        }
        if (st == LS_LAFTER && !wait) {
            cov += cm_done;                                                            // :169
            if (last_pos + 1 - cm_done == 0 || (flags & LF_PREMATURE)) {               // :173-175
                st = LS_NODE;                                                          // on to the forward search
            } else {
                last_pos -= cm_done;                                                   // :178
                cand = read_base(rw, last_pos);                                        // :182
                st = LS_LPRED;
            }
            wait = true;
        }
        // ---- forward compare (ref :236-255), up to 128 bases
        if (st == LS_CMP && !wait) {
            const uint32_t s0 = (uint32_t)(cm_s - (c_base << 5));  // the first compared base within C's 128 bases
            const uint32_t n = cm_left < 128 - s0 ? cm_left : 128 - s0;
            uint32_t matched_here = n;
            bool prem = false;
            PSA_UNROLL
            for (int j = 0; j < 4; j++) {
                const uint32_t wlo = 32u * (uint32_t)j;
                const uint32_t lo = s0 > wlo ? s0 : wlo;
                const uint32_t hi = s0 + n < wlo + 32 ? s0 + n : wlo + 32;
                if (lo < hi && !prem) {
                    const uint32_t nn = hi - lo;
                    const uint64_t cw = j == 0 ? C.w0 : j == 1 ? C.w1 : j == 2 ? C.w2 : C.w3;
                    const uint64_t refb = (cw << (2 * (lo - wlo))) >> (64 - 2 * nn);
                    const uint64_t rdb = seq_bits(rw, cm_r + (lo - s0), nn);
                    const uint64_t mask = fold_pairs(rev_pairs(rdb ^ refb) >> (64 - 2 * nn));  // t-th base of the chunk at bit 2t
                    const uint32_t c = (uint32_t)popc64(mask);
                    if (snp + c > lp.allowed) {                                        // :246-250
                        prem = true;
                        matched_here = (lo - s0) + nth_mismatch(mask, lp.allowed + 1 - snp);
                    } else {
                        snp += c;
                    }
                }
            }
            if (prem) {
                cm_done += matched_here;
                if constexpr (EV) this->ev.bases += matched_here + 1;
                flags |= LF_PREMATURE;
                st = LS_AFTER;
            } else {
                cm_done += n; cm_left -= n; cm_r += n; cm_s += n;
                if constexpr (EV) this->ev.bases += n;
                if (cm_left == 0) st = LS_AFTER;
                else wait = true;
            }
        }
        // ---- after the compare (ref :254-300)
        if (st == LS_AFTER && !wait) {
            cov += cm_done;                                                            // :254
            kmer_pos += cm_done;                                                       // :257
            if (kmer_pos >= L) {                                                       // :259-261
                st = LS_FIN;
            } else if (!(flags & LF_PREMATURE) && cand != kNone) {                     // :267
                if constexpr (EV) this->ev.jumps++;
                node_id = cand;                                                        // :269-278
                kmer_offset = 0;                                                       // :279
                kmer_pos -= k - 1;                                                     // :282
                cov -= k - 1;                                                          // :283
                wait = true;
                st = LS_NODE;
            } else if (kmer_pos > L - k) {                                             // :287-290
                st = LS_FIN;
            } else {                                                                   // :293
                f_start = kmer_pos; f_p = kmer_pos; f_probes = 0;
                st = LS_PROBE;
            }
        }
        // ---- find_kmer_match, one position (ref :92-111)
        if (st == LS_PROBE && !wait) {
            const uint32_t last = L - k;
            if (f_p > last) {                                                          // :92 exhausted, :113
                kmer_pos = f_start + kSeedStride * ((last - f_start) / kSeedStride + 1);
                st = LS_FIN;                                                           // first search: None (:305-314); re-seed: break (:296-298)
            } else {
                const uint32_t reseed = lp.max_probes > kReseedProbes ? lp.max_probes : kReseedProbes;
                if (f_probes >= ((flags & LF_SEEDED) ? reseed : lp.max_probes)) {      // hand the read over
                    why = (flags & LF_SEEDED) ? 1u : 0u;
                    out.emit = (why == 0 && lp.to_scan) ? LE_TO_SCAN : LE_TO_COOP;
                    st = LS_NEW;
                    return out;
                }
                if constexpr (EV) this->ev.lookups++;                                                  // :95
                hk = make_hash(KmerOps<KW>::fold(KmerOps<KW>::get(rw, f_p, k)));       // :93
                f_lvl = 0;
                wait = true;
                st = LS_BUCKET;
            }
        }
        // ---- map_read_with_mismatch + the process_reads flag (ref :361-376, :453-462)
        if (st == LS_FIN && !wait) {
            HitRec h;
            h.coverage = 0; h.n_tx = 0; h.tx_off = 0; h.eq_id = kNone; h.flags = 0;
            uint64_t count_slot = ix.n_eq + 1;  // None
            if (flags & LF_PUSHED) {                                                   // :305
                uint32_t count, eq_id;
                int s;
                if (!class_result(ix, cls, lp.max_small, count, eq_id, s)) {
                    why = 3; out.emit = LE_TO_COOP; st = LS_NEW;
                    return out;
                }
                h.coverage = cov;
                h.n_tx = count;
                h.eq_id = eq_id;
                h.flags = kFlagAligned | ((cov >= kCoverageThreshold && count == 0) ? kFlagMapped : 0u);  // ref :455 (sic)
                if (eq_id != kNone) {
                    count_slot = eq_id;  // (members: k_expand reads them from the index through eq_id)
                } else {
                    count_slot = ix.n_eq;
                    if (lp.want_members) {   // (the empty set too: it is listed and counted like any other)
                        uint64_t o = 0;
                        uint32_t* dst = sink.novel(r, count, o);
                        if (!dst) sink.novel_overflow();
                        else if (count) class_members(ix, cls, s, dst);
                        h.tx_off = o;
                    }
                }
            }
            sink.result(r, h, count_slot);
            out.n_tx = h.n_tx;
            out.aligned = h.flags & kFlagAligned;
            out.emit = LE_RESULT;
            st = LS_NEW;
        }
        return out;
    }
};

}  // namespace psa
