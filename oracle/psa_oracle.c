/*
 * psa_oracle.c -- CPU ORACLE (test infrastructure; see psa_oracle.h for the scope rules).
 *
 * Every function cites the reference lines it restates.  "ref" = /root/reference at
 * commit 9d9cab8 (10XGenomics/rust-pseudoaligner, crate debruijn_mapping 0.6.0).
 * Written for clarity first; the only concessions to speed are word-wise k-mer
 * extraction and an O(1) k-mer table, so that it is a fair single-thread CPU baseline.
 */
#include "psa_oracle.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;

/* ref src/config.rs:16-18 and the stride literal at src/pseudoaligner.rs:110 */
#define READ_COVERAGE_THRESHOLD 32u
#define LEFT_EXTEND_FRACTION 0.2
#define DEFAULT_ALLOWED_MISMATCHES 2u
#define SEED_STRIDE 3u

#define EXT_RIGHT(b) (1u << (b))       /* bit b    : right extension with base b */
#define EXT_LEFT(b) (1u << (4 + (b)))  /* bit 4+b  : left extension with base b  */

static __thread char g_err[256];
const char* orc_last_error(void) { return g_err; }
#define FAIL(...)                                   \
    do {                                            \
        snprintf(g_err, sizeof g_err, __VA_ARGS__); \
    } while (0)

struct orc_index {
    uint32_t k;
    uint64_t n_nodes, n_seq_words, n_eq, n_kmers, n_pure_cycles;
    uint64_t* seq_words;  /* all unitigs concatenated, DnaString packing */
    uint64_t* node_start; /* first base of node i in the concatenation   */
    uint32_t* node_len;   /* bases                                       */
    uint8_t* node_exts;   /* EXT_RIGHT / EXT_LEFT bits                   */
    uint32_t* node_eq;    /* colour (eq class id)                        */
    uint64_t* eq_offsets; /* CSR, n_eq+1                                 */
    uint32_t* eq_members; /* sorted unique transcript ids per class      */
    uint32_t *succ, *pred; /* 4 per node                                 */
    /* exact k-mer -> (node, offset) table, linear probing */
    uint64_t tab_mask;
    uint64_t* tab_lo;
    uint64_t* tab_hi;  /* NULL when k <= 32 */
    uint64_t* tab_val; /* node << 32 | offset ; ~0 = empty */
};

/* ------------------------------------------------------------------------------------
 * 2-bit packed sequences (debruijn::dna_string::DnaString; ref call sites
 * src/pseudoaligner.rs:93,103,156,182,241,265,450)
 * ---------------------------------------------------------------------------------- */
uint64_t orc_words_for(uint64_t n_bases) { return (n_bases + 31) / 32; }

static inline unsigned base_code(char c) {
    switch (c) {
        case 'A': case 'a': return 0;
        case 'C': case 'c': return 1;
        case 'G': case 'g': return 2;
        case 'T': case 't': return 3;
        default: return 0; /* QUIRK-6: non-ACGT read characters become A */
    }
}

void orc_pack_ascii(const char* s, uint64_t len, uint64_t* words) {
    /* one table lookup per base, 32 bases per word (a fair scalar DnaString::from_dna_string) */
    static uint8_t lut[256];
    static int lut_ready = 0;
    if (!lut_ready) {
        for (int c = 0; c < 256; c++) lut[c] = (uint8_t)base_code((char)c);
        lut_ready = 1;
    }
    const unsigned char* u = (const unsigned char*)s;
    uint64_t nw = orc_words_for(len);
    for (uint64_t w = 0; w < nw; w++) {
        uint64_t n = len - 32 * w < 32 ? len - 32 * w : 32;
        uint64_t v = 0;
        for (uint64_t i = 0; i < n; i++) v = (v << 2) | lut[u[32 * w + i]];
        words[w] = v << (64 - 2 * n);
    }
}

/* DnaString::get(i) */
static inline unsigned seq_get(const uint64_t* w, uint64_t i) {
    return (unsigned)(w[i >> 5] >> (62 - 2 * (i & 31))) & 3u;
}

static inline void seq_set(uint64_t* w, uint64_t i, unsigned b) {
    w[i >> 5] |= (uint64_t)b << (62 - 2 * (i & 31));
}

/* n (1..32) bases starting at base pos, right-aligned (first base most significant). */
static inline uint64_t seq_bits(const uint64_t* w, uint64_t pos, unsigned n) {
    uint64_t wi = pos >> 5;
    unsigned in_word = (unsigned)(pos & 31);
    uint64_t v = w[wi] << (2 * in_word);
    if (n > 32 - in_word) v |= w[wi + 1] >> (64 - 2 * in_word);
    return v >> (64 - 2 * n);
}

/* DnaString::get_kmer(pos): base 0 most significant, low 2k bits used. */
static inline u128 seq_kmer(const uint64_t* w, uint64_t pos, unsigned k) {
    if (k <= 32) return seq_bits(w, pos, k);
    return ((u128)seq_bits(w, pos, k - 32) << 64) | seq_bits(w, pos + (k - 32), 32);
}

static inline u128 kmer_mask(unsigned k) { return k == 64 ? ~(u128)0 : (((u128)1 << (2 * k)) - 1); }

/* ------------------------------------------------------------------------------------
 * k-mer dictionary.  Stands in for NoKeyBoomHashMap::get (ref src/pseudoaligner.rs:96);
 * membership is exact, so the reference's verification step (:99-107) always passes,
 * but the oracle still performs it.
 * ---------------------------------------------------------------------------------- */
static inline uint64_t mix64(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL;
    x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL;
    x ^= x >> 33;
    return x;
}
static inline uint64_t kmer_hash(u128 km) { return mix64((uint64_t)km ^ mix64((uint64_t)(km >> 64) + 0x9e3779b97f4a7c15ULL)); }

static int dict_alloc(orc_index* ix, uint64_t n_keys) {
    uint64_t cap = 16;
    while (cap < n_keys * 2) cap <<= 1;
    ix->tab_mask = cap - 1;
    ix->tab_lo = (uint64_t*)malloc(cap * 8);
    ix->tab_val = (uint64_t*)malloc(cap * 8);
    ix->tab_hi = ix->k > 32 ? (uint64_t*)malloc(cap * 8) : NULL;
    if (!ix->tab_lo || !ix->tab_val || (ix->k > 32 && !ix->tab_hi)) return -1;
    memset(ix->tab_val, 0xff, cap * 8);
    return 0;
}

static int dict_put(orc_index* ix, u128 km, uint32_t node, uint32_t off) {
    uint64_t h = kmer_hash(km) & ix->tab_mask;
    uint64_t lo = (uint64_t)km, hi = (uint64_t)(km >> 64);
    while (ix->tab_val[h] != ~0ULL) {
        if (ix->tab_lo[h] == lo && (!ix->tab_hi || ix->tab_hi[h] == hi)) return -1; /* duplicate k-mer */
        h = (h + 1) & ix->tab_mask;
    }
    ix->tab_lo[h] = lo;
    if (ix->tab_hi) ix->tab_hi[h] = hi;
    ix->tab_val[h] = ((uint64_t)node << 32) | off;
    return 0;
}

static inline int dict_get(const orc_index* ix, u128 km, uint32_t* node, uint32_t* off) {
    uint64_t h = kmer_hash(km) & ix->tab_mask;
    uint64_t lo = (uint64_t)km, hi = (uint64_t)(km >> 64);
    for (;;) {
        uint64_t v = ix->tab_val[h];
        if (v == ~0ULL) return 0;
        if (ix->tab_lo[h] == lo && (!ix->tab_hi || ix->tab_hi[h] == hi)) {
            *node = (uint32_t)(v >> 32);
            *off = (uint32_t)v;
            return 1;
        }
        h = (h + 1) & ix->tab_mask;
    }
}

/* ------------------------------------------------------------------------------------
 * Finishing an index given its flat arrays: k-mer table (what make_dbg_index produces,
 * ref src/build_index.rs:182-221: every k-mer of every node -> (node_id, offset)) and the
 * successor / predecessor tables (what Node::r_edges()/l_edges() compute on the fly in
 * debruijn: the node whose first k-mer is last(k-1)+b, resp. whose last k-mer is
 * b+first(k-1); stranded, so no reverse-complement links -- ref src/config.rs:14).
 * ---------------------------------------------------------------------------------- */
static int index_finish(orc_index* ix) {
    const unsigned k = ix->k;
    uint64_t n_kmers = 0;
    for (uint64_t i = 0; i < ix->n_nodes; i++) {
        if (ix->node_len[i] < k) { FAIL("node %llu shorter than k", (unsigned long long)i); return -1; }
        n_kmers += ix->node_len[i] - k + 1;
    }
    ix->n_kmers = n_kmers;
    if (dict_alloc(ix, n_kmers)) { FAIL("out of memory (dict)"); return -1; }
    for (uint64_t i = 0; i < ix->n_nodes; i++) {
        uint64_t nk = ix->node_len[i] - k + 1;
        for (uint64_t o = 0; o < nk; o++) {
            u128 km = seq_kmer(ix->seq_words, ix->node_start[i] + o, k);
            if (dict_put(ix, km, (uint32_t)i, (uint32_t)o)) { FAIL("duplicate k-mer in graph (node %llu off %llu)", (unsigned long long)i, (unsigned long long)o); return -1; }
        }
    }
    ix->succ = (uint32_t*)malloc(ix->n_nodes * 16 + 16);
    ix->pred = (uint32_t*)malloc(ix->n_nodes * 16 + 16);
    if (!ix->succ || !ix->pred) { FAIL("out of memory (edges)"); return -1; }
    memset(ix->succ, 0xff, ix->n_nodes * 16 + 16);
    memset(ix->pred, 0xff, ix->n_nodes * 16 + 16);
    const u128 mask = kmer_mask(k);
    for (uint64_t i = 0; i < ix->n_nodes; i++) {
        u128 first = seq_kmer(ix->seq_words, ix->node_start[i], k);
        u128 last = seq_kmer(ix->seq_words, ix->node_start[i] + ix->node_len[i] - k, k);
        for (unsigned b = 0; b < 4; b++) {
            uint32_t n, o;
            if (ix->node_exts[i] & EXT_RIGHT(b)) {
                u128 next = ((last << 2) | b) & mask;
                /* debruijn find_link expects every ext to resolve ("missing link") */
                if (!dict_get(ix, next, &n, &o) || o != 0) { FAIL("missing right link: node %llu base %u", (unsigned long long)i, b); return -1; }
                ix->succ[4 * i + b] = n;
            }
            if (ix->node_exts[i] & EXT_LEFT(b)) {
                u128 prev = (first >> 2) | ((u128)b << (2 * (k - 1)));
                if (!dict_get(ix, prev, &n, &o) || o != ix->node_len[n] - k) { FAIL("missing left link: node %llu base %u", (unsigned long long)i, b); return -1; }
                ix->pred[4 * i + b] = n;
            }
        }
    }
    return 0;
}

static void* dup_mem(const void* p, size_t n, size_t extra) {
    void* q = calloc(1, n + extra + 8);
    if (q && n) memcpy(q, p, n);
    return q;
}

orc_index* orc_index_from_flat(uint32_t k, uint64_t n_nodes, const uint64_t* seq_words,
                               uint64_t n_seq_words, const uint64_t* node_start,
                               const uint32_t* node_len, const uint8_t* node_exts,
                               const uint32_t* node_eq, uint64_t n_eq,
                               const uint64_t* eq_offsets, const uint32_t* eq_members) {
    if (k < 1 || k > 64) { FAIL("k out of range"); return NULL; }
    orc_index* ix = (orc_index*)calloc(1, sizeof *ix);
    if (!ix) return NULL;
    ix->k = k; ix->n_nodes = n_nodes; ix->n_seq_words = n_seq_words; ix->n_eq = n_eq;
    ix->seq_words = (uint64_t*)dup_mem(seq_words, n_seq_words * 8, 16); /* +pad word for seq_bits */
    ix->node_start = (uint64_t*)dup_mem(node_start, n_nodes * 8, 0);
    ix->node_len = (uint32_t*)dup_mem(node_len, n_nodes * 4, 0);
    ix->node_exts = (uint8_t*)dup_mem(node_exts, n_nodes, 0);
    ix->node_eq = (uint32_t*)dup_mem(node_eq, n_nodes * 4, 0);
    ix->eq_offsets = (uint64_t*)dup_mem(eq_offsets, (n_eq + 1) * 8, 0);
    ix->eq_members = (uint32_t*)dup_mem(eq_members, eq_offsets[n_eq] * 4, 0);
    if (!ix->seq_words || !ix->node_start || !ix->node_len || !ix->node_exts || !ix->node_eq ||
        !ix->eq_offsets || !ix->eq_members) { FAIL("out of memory"); orc_index_free(ix); return NULL; }
    for (uint64_t i = 0; i < n_nodes; i++)
        if (node_eq[i] >= n_eq) { FAIL("node %llu: eq id out of range", (unsigned long long)i); orc_index_free(ix); return NULL; }
    if (index_finish(ix)) { orc_index_free(ix); return NULL; }
    return ix;
}

void orc_index_free(orc_index* ix) {
    if (!ix) return;
    free(ix->seq_words); free(ix->node_start); free(ix->node_len); free(ix->node_exts);
    free(ix->node_eq); free(ix->eq_offsets); free(ix->eq_members); free(ix->succ); free(ix->pred);
    free(ix->tab_lo); free(ix->tab_hi); free(ix->tab_val);
    free(ix);
}

uint32_t orc_index_k(const orc_index* ix) { return ix->k; }
uint64_t orc_index_n_nodes(const orc_index* ix) { return ix->n_nodes; }
uint64_t orc_index_n_kmers(const orc_index* ix) { return ix->n_kmers; }
uint64_t orc_index_n_eq(const orc_index* ix) { return ix->n_eq; }
uint64_t orc_index_n_seq_words(const orc_index* ix) { return ix->n_seq_words; }
uint64_t orc_index_n_pure_cycles(const orc_index* ix) { return ix->n_pure_cycles; }
const uint64_t* orc_index_seq_words(const orc_index* ix) { return ix->seq_words; }
const uint64_t* orc_index_node_start(const orc_index* ix) { return ix->node_start; }
const uint32_t* orc_index_node_len(const orc_index* ix) { return ix->node_len; }
const uint8_t* orc_index_node_exts(const orc_index* ix) { return ix->node_exts; }
const uint32_t* orc_index_node_eq(const orc_index* ix) { return ix->node_eq; }
const uint64_t* orc_index_eq_offsets(const orc_index* ix) { return ix->eq_offsets; }
const uint32_t* orc_index_eq_members(const orc_index* ix) { return ix->eq_members; }
const uint32_t* orc_index_succ(const orc_index* ix) { return ix->succ; }
const uint32_t* orc_index_pred(const orc_index* ix) { return ix->pred; }

int orc_index_lookup(const orc_index* ix, const uint64_t* kmer_words, uint32_t* node, uint32_t* offset) {
    uint64_t tmp[4] = {kmer_words[0], ix->k > 32 ? kmer_words[1] : 0, 0, 0};
    return dict_get(ix, seq_kmer(tmp, 0, ix->k), node, offset);
}

/* ------------------------------------------------------------------------------------
 * Naive index builder.  Semantics, not algorithm, of ref src/build_index.rs:27-221:
 *  - every k-mer of every transcript with len >= k (partition_contigs, :134), stranded
 *    (src/config.rs:14);
 *  - colour of a k-mer = sorted, de-duplicated list of transcript indices containing it,
 *    interned to a dense id (CountFilterEqClass::summarize, src/equiv_classes.rs:62-91);
 *  - exts of a k-mer = union over its occurrences of the base before / after it in the
 *    transcript (Exts::from_dna_string at :144 + filter_kmers, all_exts.add at
 *    src/equiv_classes.rs:73);
 *  - unitigs = maximal paths whose every internal link is the unique right ext of its
 *    source AND the unique left ext of its target AND joins equal colours
 *    (compress_kmers_with_hash / compress_graph with ScmapCompress, :171,:178; debruijn's
 *    path extension also stops at a k-mer that is already used, which cuts cycles).
 * ---------------------------------------------------------------------------------- */
typedef struct { u128 kmer; uint32_t tx; uint8_t exts; } occ_t;

static int occ_cmp(const void* a, const void* b) {
    const occ_t* x = (const occ_t*)a; const occ_t* y = (const occ_t*)b;
    if (x->kmer != y->kmer) return x->kmer < y->kmer ? -1 : 1;
    if (x->tx != y->tx) return x->tx < y->tx ? -1 : 1;
    return 0;
}

/* distinct k-mer table used during the build */
typedef struct {
    uint64_t n;
    u128* kmer;     /* ascending */
    uint8_t* exts;
    uint32_t* eq;
} kmer_tab;

static int64_t ktab_find(const kmer_tab* t, u128 km) {
    uint64_t lo = 0, hi = t->n;
    while (lo < hi) {
        uint64_t mid = lo + (hi - lo) / 2;
        if (t->kmer[mid] < km) lo = mid + 1; else hi = mid;
    }
    return (lo < t->n && t->kmer[lo] == km) ? (int64_t)lo : -1;
}

static inline int one_bit(unsigned x) { return x && !(x & (x - 1)); }

/* If k-mer i links forward to a unique same-colour k-mer whose only left ext is i, return it. */
static int64_t link_fwd(const kmer_tab* t, uint64_t i, unsigned k, u128 mask) {
    unsigned r = t->exts[i] & 0xf;
    if (!one_bit(r)) return -1;
    unsigned b = (unsigned)__builtin_ctz(r);
    int64_t j = ktab_find(t, ((t->kmer[i] << 2) | b) & mask);
    if (j < 0) return -2; /* observed neighbour must exist */
    if (!one_bit(t->exts[j] >> 4)) return -1;
    if (t->eq[j] != t->eq[i]) return -1;
    (void)k;
    return j;
}
static int64_t link_bwd(const kmer_tab* t, uint64_t i, unsigned k, u128 mask) {
    unsigned l = t->exts[i] >> 4;
    if (!one_bit(l)) return -1;
    unsigned b = (unsigned)__builtin_ctz(l);
    (void)mask;
    int64_t j = ktab_find(t, (t->kmer[i] >> 2) | ((u128)b << (2 * (k - 1))));
    if (j < 0) return -2;
    if (!one_bit(t->exts[j] & 0xf)) return -1;
    if (t->eq[j] != t->eq[i]) return -1;
    return j;
}

/* colour interning: chained hash on the member list */
typedef struct { uint64_t off; uint32_t len; uint32_t id; int64_t next; } cls_ent;

orc_index* orc_index_build(const uint8_t* codes, const uint64_t* tx_off, uint32_t n_tx, uint32_t k) {
    if (k < 2 || k > 64) { FAIL("k out of range"); return NULL; }
    const u128 mask = kmer_mask(k);
    orc_index* ix = NULL;
    occ_t* occ = NULL; kmer_tab t = {0, NULL, NULL, NULL};
    uint32_t* members = NULL; cls_ent* cls = NULL; int64_t* buckets = NULL;
    uint8_t* used = NULL; uint64_t* path = NULL;
    uint64_t* eq_off = NULL;

    uint64_t n_occ = 0;
    for (uint32_t tx = 0; tx < n_tx; tx++) {
        uint64_t len = tx_off[tx + 1] - tx_off[tx];
        if (len >= k) n_occ += len - k + 1;
    }
    occ = (occ_t*)malloc((n_occ + 1) * sizeof *occ);
    if (!occ) { FAIL("out of memory"); goto fail; }
    uint64_t w = 0;
    for (uint32_t tx = 0; tx < n_tx; tx++) {
        const uint8_t* s = codes + tx_off[tx];
        uint64_t len = tx_off[tx + 1] - tx_off[tx];
        if (len < k) continue;
        u128 km = 0;
        for (uint64_t i = 0; i < len; i++) {
            if (s[i] > 3) { FAIL("base code > 3 in transcript %u", tx); goto fail; }
            km = ((km << 2) | s[i]) & mask;
            if (i + 1 >= k) {
                uint64_t p = i + 1 - k; /* k-mer start */
                uint8_t e = 0;
                if (p > 0) e |= EXT_LEFT(s[p - 1]);
                if (i + 1 < len) e |= EXT_RIGHT(s[i + 1]);
                occ[w].kmer = km; occ[w].tx = tx; occ[w].exts = e; w++;
            }
        }
    }
    qsort(occ, n_occ, sizeof *occ, occ_cmp);

    /* group occurrences -> distinct k-mers, colours, exts */
    uint64_t n_dist = 0;
    for (uint64_t i = 0; i < n_occ; i++) if (i == 0 || occ[i].kmer != occ[i - 1].kmer) n_dist++;
    t.n = n_dist;
    t.kmer = (u128*)malloc((n_dist + 1) * sizeof(u128));
    t.exts = (uint8_t*)malloc(n_dist + 1);
    t.eq = (uint32_t*)malloc((n_dist + 1) * 4);
    members = (uint32_t*)malloc((n_occ + 1) * 4); /* upper bound on total class size */
    cls = (cls_ent*)malloc((n_dist + 1) * sizeof *cls);
    uint64_t n_buckets = 1; while (n_buckets < n_dist + 1) n_buckets <<= 1;
    buckets = (int64_t*)malloc(n_buckets * 8);
    if (!t.kmer || !t.exts || !t.eq || !members || !cls || !buckets) { FAIL("out of memory"); goto fail; }
    memset(buckets, 0xff, n_buckets * 8);
    uint64_t n_cls = 0, n_mem = 0, d = 0;
    for (uint64_t i = 0; i < n_occ;) {
        uint64_t j = i; uint8_t e = 0;
        uint64_t m0 = n_mem; /* tentative member list */
        while (j < n_occ && occ[j].kmer == occ[i].kmer) {
            e |= occ[j].exts;
            if (n_mem == m0 || members[n_mem - 1] != occ[j].tx) members[n_mem++] = occ[j].tx; /* sort+dedup */
            j++;
        }
        uint32_t len = (uint32_t)(n_mem - m0);
        uint64_t h = 1469598103934665603ULL;
        for (uint32_t q = 0; q < len; q++) h = mix64(h ^ members[m0 + q]);
        int64_t c = buckets[h & (n_buckets - 1)];
        while (c >= 0 && !(cls[c].len == len && !memcmp(members + cls[c].off, members + m0, (size_t)len * 4))) c = cls[c].next;
        if (c < 0) { /* new class: ids dense in first-appearance order (equiv_classes.rs:84-89) */
            c = (int64_t)n_cls++;
            cls[c].off = m0; cls[c].len = len; cls[c].id = (uint32_t)c;
            cls[c].next = buckets[h & (n_buckets - 1)]; buckets[h & (n_buckets - 1)] = c;
        } else {
            n_mem = m0; /* already interned: drop the tentative copy */
        }
        t.kmer[d] = occ[i].kmer; t.exts[d] = e; t.eq[d] = (uint32_t)c; d++;
        i = j;
    }
    free(occ); occ = NULL;

    /* unitigs */
    used = (uint8_t*)calloc(n_dist + 1, 1);
    path = (uint64_t*)malloc((n_dist + 1) * 8);
    if (!used || !path) { FAIL("out of memory"); goto fail; }
    ix = (orc_index*)calloc(1, sizeof *ix);
    if (!ix) { FAIL("out of memory"); goto fail; }
    ix->k = k;
    /* first pass counts nodes/bases, second pass fills; do it in one pass with growth */
    uint64_t cap_nodes = 1024, cap_words = 1024, n_nodes = 0, n_bases = 0;
    ix->node_start = (uint64_t*)malloc(cap_nodes * 8);
    ix->node_len = (uint32_t*)malloc(cap_nodes * 4);
    ix->node_exts = (uint8_t*)malloc(cap_nodes);
    ix->node_eq = (uint32_t*)malloc(cap_nodes * 4);
    ix->seq_words = (uint64_t*)calloc(cap_words + 2, 8);
    if (!ix->node_start || !ix->node_len || !ix->node_exts || !ix->node_eq || !ix->seq_words) { FAIL("out of memory"); goto fail; }
    for (uint64_t s = 0; s < n_dist; s++) {
        if (used[s]) continue;
        /* walk left to the start of the path containing s */
        uint64_t head = s; int pure_cycle = 0;
        for (;;) {
            int64_t p = link_bwd(&t, head, k, mask);
            if (p == -2) { FAIL("k-mer neighbour missing (left)"); goto fail; }
            if (p < 0) break;
            if ((uint64_t)p == s) { pure_cycle = 1; head = s; break; } /* closed cycle: cut at its smallest k-mer */
            if (used[p]) break; /* cannot happen for a maximal path; kept for safety */
            head = (uint64_t)p;
        }
        if (pure_cycle) ix->n_pure_cycles++;
        /* walk right collecting the path */
        uint64_t plen = 0, cur = head;
        for (;;) {
            path[plen++] = cur; used[cur] = 1;
            int64_t nx = link_fwd(&t, cur, k, mask);
            if (nx == -2) { FAIL("k-mer neighbour missing (right)"); goto fail; }
            if (nx < 0 || used[nx]) break; /* used: self-loop or closed cycle */
            cur = (uint64_t)nx;
        }
        /* emit node */
        uint64_t nlen = k + plen - 1;
        if (n_nodes == cap_nodes) {
            cap_nodes *= 2;
            ix->node_start = (uint64_t*)realloc(ix->node_start, cap_nodes * 8);
            ix->node_len = (uint32_t*)realloc(ix->node_len, cap_nodes * 4);
            ix->node_exts = (uint8_t*)realloc(ix->node_exts, cap_nodes);
            ix->node_eq = (uint32_t*)realloc(ix->node_eq, cap_nodes * 4);
            if (!ix->node_start || !ix->node_len || !ix->node_exts || !ix->node_eq) { FAIL("out of memory"); goto fail; }
        }
        while (orc_words_for(n_bases + nlen) > cap_words) {
            uint64_t nc = cap_words * 2;
            uint64_t* nw = (uint64_t*)calloc(nc + 2, 8);
            if (!nw) { FAIL("out of memory"); goto fail; }
            memcpy(nw, ix->seq_words, cap_words * 8);
            free(ix->seq_words); ix->seq_words = nw; cap_words = nc;
        }
        ix->node_start[n_nodes] = n_bases;
        ix->node_len[n_nodes] = (uint32_t)nlen;
        ix->node_exts[n_nodes] = (uint8_t)((t.exts[path[0]] & 0xf0) | (t.exts[path[plen - 1]] & 0x0f));
        ix->node_eq[n_nodes] = t.eq[path[0]];
        u128 first = t.kmer[path[0]];
        for (unsigned i = 0; i < k; i++) seq_set(ix->seq_words, n_bases + i, (unsigned)(first >> (2 * (k - 1 - i))) & 3u);
        for (uint64_t q = 1; q < plen; q++) seq_set(ix->seq_words, n_bases + k - 1 + q, (unsigned)t.kmer[path[q]] & 3u);
        n_bases += nlen; n_nodes++;
    }
    ix->n_nodes = n_nodes;
    ix->n_seq_words = orc_words_for(n_bases);
    /* classes -> CSR in id order */
    ix->n_eq = n_cls;
    eq_off = (uint64_t*)malloc((n_cls + 1) * 8);
    ix->eq_members = (uint32_t*)malloc((n_mem + 1) * 4);
    if (!eq_off || !ix->eq_members) { FAIL("out of memory"); goto fail; }
    uint64_t o = 0;
    for (uint64_t c = 0; c < n_cls; c++) {
        eq_off[c] = o;
        memcpy(ix->eq_members + o, members + cls[c].off, (size_t)cls[c].len * 4);
        o += cls[c].len;
    }
    eq_off[n_cls] = o;
    ix->eq_offsets = eq_off; eq_off = NULL;
    free(t.kmer); free(t.exts); free(t.eq); free(members); free(cls); free(buckets); free(used); free(path);
    t.kmer = NULL; t.exts = NULL; t.eq = NULL; members = NULL; cls = NULL; buckets = NULL; used = NULL; path = NULL;
    if (index_finish(ix)) { orc_index_free(ix); return NULL; }
    if (ix->n_kmers != n_dist) { FAIL("k-mer count mismatch after compaction"); orc_index_free(ix); return NULL; }
    return ix;
fail:
    free(occ); free(t.kmer); free(t.exts); free(t.eq); free(members); free(cls); free(buckets);
    free(used); free(path); free(eq_off);
    orc_index_free(ix);
    return NULL;
}

/* ------------------------------------------------------------------------------------
 * intersect -- ref src/pseudoaligner.rs:389-418, statement by statement.  In place on v1.
 * Note the reference's `v2[idx2..].binary_search(..)` returns a position RELATIVE to the
 * suffix it searched and the reference then ASSIGNS it (`idx2 = pos + 1`, :409; `idx2 = pos`,
 * :413) instead of adding it to the suffix's start; the next search therefore starts from a
 * smaller (or equal) index than where the previous match was -- it merely searches a larger
 * suffix than necessary, which cannot change the outcome on ascending inputs.  Restated as is.
 * ---------------------------------------------------------------------------------- */
/* Rust slice::binary_search on a strictly ascending slice: found -> (1,pos); else (0, insertion pos) */
static inline int bsearch_u32(const uint32_t* v, uint32_t n, uint32_t x, uint32_t* pos) {
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
        uint32_t mid = lo + (hi - lo) / 2;
        if (v[mid] < x) lo = mid + 1; else hi = mid;
    }
    *pos = lo;
    return lo < n && v[lo] == x;
}

uint32_t orc_intersect(uint32_t* v1, uint32_t n1, const uint32_t* v2, uint32_t n2) {
    if (n1 == 0) return 0;                 /* :390-392 */
    if (n2 == 0) return 0;                 /* :394-396 v1.clear() */
    uint32_t fill_idx1 = 0, idx1 = 0, idx2 = 0; /* :398-400 */
    while (idx1 < n1 && idx2 < n2) {       /* :402 */
        uint32_t pos;
        if (bsearch_u32(v2 + idx2, n2 - idx2, v1[idx1], &pos)) { /* :403-404 v2[idx2..].binary_search(&v1[idx1]) */
            uint32_t tmp = v1[fill_idx1]; v1[fill_idx1] = v1[idx1]; v1[idx1] = tmp; /* :406 swap */
            fill_idx1++;                   /* :407 */
            idx1++;                        /* :408 */
            idx2 = pos + 1;                /* :409 (sic: suffix-relative position assigned) */
        } else {
            idx1++;                        /* :412 */
            idx2 = pos;                    /* :413 (sic) */
        }
    }
    return fill_idx1;                      /* :417 truncate */
}

/* ------------------------------------------------------------------------------------
 * map_read_to_nodes_with_mismatch -- ref src/pseudoaligner.rs:64-319, line by line.
 * nodes[] receives node ids in push order; returns 1 (Some) / 0 (None).
 * ---------------------------------------------------------------------------------- */
typedef struct { uint32_t* v; uint32_t n, cap; int overflow; } node_vec;
static inline void nv_push(node_vec* nv, uint32_t x) {
    if (nv->n < nv->cap) nv->v[nv->n] = x; else nv->overflow = 1;
    nv->n++;
}

/* find_kmer_match closure, :91-114 */
static int find_kmer_match(const orc_index* ix, const uint64_t* read, uint64_t* kmer_pos,
                           uint64_t last_kmer_pos, uint32_t* nid, uint32_t* offset, orc_events* ev) {
    const unsigned k = ix->k;
    while (*kmer_pos <= last_kmer_pos) {                       /* :92 */
        u128 read_kmer = seq_kmer(read, *kmer_pos, k);         /* :93 */
        ev->kmer_lookups++;                                    /* :95 */
        uint32_t n, o;
        if (dict_get(ix, read_kmer, &n, &o)) {                 /* :96-98 */
            ev->dict_hits++;
            u128 ref_kmer = seq_kmer(ix->seq_words, ix->node_start[n] + o, k); /* :101-103 */
            if (read_kmer == ref_kmer) { *nid = n; *offset = o; return 1; }    /* :105-107 */
        }
        *kmer_pos += SEED_STRIDE;                              /* :110 */
    }
    return 0;                                                  /* :113 */
}

static int map_read_to_nodes(const orc_index* ix, const uint64_t* read, uint64_t read_length,
                             node_vec* nodes, uint64_t allowed_mismatches, uint64_t* coverage_out,
                             orc_events* ev) {
    const uint64_t kmer_length = ix->k;
    uint64_t read_coverage = 0;                                /* :71 */
    nodes->n = 0;                                              /* :75 */
    uint64_t left_extend_threshold = (uint64_t)(LEFT_EXTEND_FRACTION * (double)read_length); /* :77 */
    uint64_t kmer_pos = 0;                                     /* :79 */
    if (read_length < kmer_length) return 0;                   /* :82-84 */
    uint64_t last_kmer_pos = read_length - kmer_length;        /* :86 */

    uint32_t node_id = 0, kmer_offset = 0;
    int have_node = find_kmer_match(ix, read, &kmer_pos, last_kmer_pos, &node_id, &kmer_offset, ev); /* :118-121 */

    /* left extension, :124-205 */
    if (have_node && kmer_pos >= left_extend_threshold && kmer_pos >= 1) {
        /* (kmer_pos >= 1 can only fail when read_length < 5, where the reference's
         *  `kmer_pos - 1` at :127 would underflow; unreachable for k >= 5.) */
        uint64_t last_pos = kmer_pos - 1;                      /* :127 */
        uint32_t prev_node_id = node_id;                       /* :128 */
        uint64_t prev_kmer_offset = kmer_offset > 0 ? kmer_offset - 1 : 0; /* :129 QUIRK-1 */
        for (;;) {                                             /* :131 */
            const uint64_t ref_start = ix->node_start[prev_node_id];
            uint64_t skipped_read = last_pos + 1;              /* :139 */
            uint64_t skipped_ref = prev_kmer_offset + 1;       /* :142 */
            uint64_t max_matchable_pos = skipped_read < skipped_ref ? skipped_read : skipped_ref; /* :145 */
            int premature_break = 0;                           /* :148 */
            uint64_t matched_bases = 0, seen_snp = 0;          /* :149-150 QUIRK-2: per node */
            for (uint64_t idx = 0; idx < max_matchable_pos; idx++) { /* :151 */
                uint64_t ref_pos = prev_kmer_offset - idx;     /* :152 */
                uint64_t read_offset = last_pos - idx;         /* :153 */
                ev->bases_compared++;
                if (seq_get(ix->seq_words, ref_start + ref_pos) != seq_get(read, read_offset)) { /* :156 */
                    seen_snp++;                                /* :161 */
                    if (seen_snp > allowed_mismatches) { premature_break = 1; break; } /* :162-165 */
                }
                matched_bases++;                               /* :168 */
                read_coverage++;                               /* :169 */
            }
            if (last_pos + 1 - matched_bases == 0 || premature_break) break; /* :173-175 */
            last_pos -= matched_bases;                         /* :178 */
            unsigned next_base = seq_get(read, last_pos);      /* :182 */
            if (ix->node_exts[prev_node_id] & EXT_LEFT(next_base)) { /* :183 */
                ev->edge_jumps++;
                prev_node_id = ix->pred[4 * (uint64_t)prev_node_id + next_base]; /* :185-194 */
                prev_kmer_offset = ix->node_len[prev_node_id] - kmer_length;     /* :196 */
                nv_push(nodes, prev_node_id);                  /* :199 QUIRK-3 */
                ev->node_visits++;
            } else {
                break;                                         /* :201 */
            }
        }
    }

    /* forward search, :208-302 */
    if (kmer_pos <= last_kmer_pos) {                           /* :208 */
        for (;;) {                                             /* :209 */
            kmer_pos += kmer_length;                           /* :215 */
            read_coverage += kmer_length;                      /* :216 */
            nv_push(nodes, node_id);                           /* :219 */
            ev->node_visits++;
            uint64_t remaining_read = read_length - kmer_pos;  /* :222 */
            const uint64_t ref_start = ix->node_start[node_id];
            uint64_t ref_length = ix->node_len[node_id];       /* :226 */
            uint64_t ref_offset = kmer_offset + kmer_length;   /* :227 */
            uint64_t informative_ref = ref_length - ref_offset; /* :228 */
            uint64_t max_matchable_pos = remaining_read < informative_ref ? remaining_read : informative_ref; /* :231 */
            int premature_break = 0;                           /* :233 */
            uint64_t matched_bases = 0, seen_snp = 0;          /* :234-235 */
            for (uint64_t idx = 0; idx < max_matchable_pos; idx++) { /* :236 */
                uint64_t ref_pos = ref_offset + idx;           /* :237 */
                uint64_t read_offset = kmer_pos + idx;         /* :238 */
                ev->bases_compared++;
                if (seq_get(ix->seq_words, ref_start + ref_pos) != seq_get(read, read_offset)) { /* :241 */
                    seen_snp++;                                /* :246 */
                    if (seen_snp > allowed_mismatches) { premature_break = 1; break; } /* :247-250 */
                }
                matched_bases++;                               /* :253 */
                read_coverage++;                               /* :254 */
            }
            kmer_pos += matched_bases;                         /* :257 */
            if (kmer_pos >= read_length) break;                /* :259-261 */
            unsigned next_base = seq_get(read, kmer_pos);      /* :265 */
            if (!premature_break && (ix->node_exts[node_id] & EXT_RIGHT(next_base))) { /* :267 */
                ev->edge_jumps++;
                node_id = ix->succ[4 * (uint64_t)node_id + next_base]; /* :269-278 */
                kmer_offset = 0;                               /* :279 */
                kmer_pos -= kmer_length - 1;                   /* :282 */
                read_coverage -= kmer_length - 1;              /* :283 */
            } else {
                if (kmer_pos > last_kmer_pos) break;           /* :287-290 */
                if (!find_kmer_match(ix, read, &kmer_pos, last_kmer_pos, &node_id, &kmer_offset, ev)) break; /* :293-299 QUIRK-5 */
            }
        }
    }

    if (nodes->n == 0) {                                       /* :305 */
        /* :306-312 would panic on read_coverage != 0; unreachable (coverage only grows after a push) */
        return 0;                                              /* :314 */
    }
    *coverage_out = read_coverage;
    return 1;                                                  /* :317 */
}

/* nodes_to_eq_class -- ref :323-356.  Returns |eq_class|, or UINT32_MAX if tx_out is too small. */
static uint32_t nodes_to_eq_class(const orc_index* ix, node_vec* nodes, uint32_t* tx_out, uint64_t tx_cap,
                                  orc_events* ev) {
    if (nodes->n == 0) return 0;                               /* :326-328 */
    /* :331-334 stable sort by class length (insertion sort is stable) */
    for (uint32_t i = 1; i < nodes->n; i++) {
        uint32_t x = nodes->v[i];
        uint64_t lx = ix->eq_offsets[ix->node_eq[x] + 1] - ix->eq_offsets[ix->node_eq[x]];
        uint32_t j = i;
        while (j > 0) {
            uint32_t y = nodes->v[j - 1];
            uint64_t ly = ix->eq_offsets[ix->node_eq[y] + 1] - ix->eq_offsets[ix->node_eq[y]];
            if (ly <= lx) break;
            nodes->v[j] = y; j--;
        }
        nodes->v[j] = x;
    }
    uint32_t first_color = ix->node_eq[nodes->v[0]];           /* :346-349 */
    uint64_t n = ix->eq_offsets[first_color + 1] - ix->eq_offsets[first_color];
    ev->class_members += n;
    if (n > tx_cap) return UINT32_MAX;
    memcpy(tx_out, ix->eq_members + ix->eq_offsets[first_color], (size_t)n * 4); /* :350 */
    uint32_t len = (uint32_t)n;
    for (uint32_t i = 1; i < nodes->n; i++) {                  /* :352 */
        uint32_t color = ix->node_eq[nodes->v[i]];             /* :353 */
        uint64_t m = ix->eq_offsets[color + 1] - ix->eq_offsets[color];
        ev->class_members += m;
        len = orc_intersect(tx_out, len, ix->eq_members + ix->eq_offsets[color], (uint32_t)m); /* :354 */
    }
    return len;
}

int orc_map_read(const orc_index* ix, const uint64_t* read_words, uint32_t read_len,
                 uint32_t* tx_out, uint64_t tx_cap, uint32_t* n_tx, uint32_t* coverage,
                 uint32_t* eq_id, uint32_t* nodes_out, uint32_t nodes_cap, uint32_t* n_nodes,
                 orc_events* ev) {
    /* map_read, :381-384: map_read_with_mismatch with DEFAULT_ALLOWED_MISMATCHES */
    return orc_map_read_with_mismatch(ix, read_words, read_len, DEFAULT_ALLOWED_MISMATCHES, tx_out, tx_cap, n_tx,
                                      coverage, eq_id, nodes_out, nodes_cap, n_nodes, ev);
}

int orc_map_read_with_mismatch(const orc_index* ix, const uint64_t* read_words, uint32_t read_len,
                               uint32_t allowed_mismatches, uint32_t* tx_out, uint64_t tx_cap, uint32_t* n_tx,
                               uint32_t* coverage, uint32_t* eq_id, uint32_t* nodes_out, uint32_t nodes_cap,
                               uint32_t* n_nodes, orc_events* ev) {
    orc_events local; if (!ev) { memset(&local, 0, sizeof local); ev = &local; }
    /* map_read_with_mismatch, :361-376 */
    uint32_t stack_nodes[512];
    node_vec nv; nv.n = 0; nv.overflow = 0;
    uint32_t need = 2 * read_len + 2;
    uint32_t* heap = NULL;
    if (need <= 512) { nv.v = stack_nodes; nv.cap = 512; }
    else { heap = (uint32_t*)malloc((size_t)need * 4); nv.v = heap; nv.cap = need; }
    uint64_t cov = 0;
    ev->reads++; ev->read_bases += read_len;
    int some = map_read_to_nodes(ix, read_words, read_len, &nv, allowed_mismatches, &cov, ev);
    *n_tx = 0; *coverage = 0; if (eq_id) *eq_id = ORC_EQ_NONE; if (n_nodes) *n_nodes = 0;
    if (!some) { free(heap); return 0; }
    if (n_nodes) {
        *n_nodes = nv.n;
        for (uint32_t i = 0; i < nv.n && i < nodes_cap; i++) nodes_out[i] = nv.v[i];
    }
    uint32_t len = nodes_to_eq_class(ix, &nv, tx_out, tx_cap, ev);
    if (len == UINT32_MAX) { free(heap); return -1; }
    /* eq id of the result when it coincides with a visited class (not a reference
     * output; our per-class counting key -- see include/psa.h) */
    uint32_t id = ORC_EQ_NONE;
    for (uint32_t i = 0; i < nv.n; i++) {
        uint32_t c = ix->node_eq[nv.v[i]];
        if (ix->eq_offsets[c + 1] - ix->eq_offsets[c] == len && c < id) id = c;
    }
    if (eq_id) *eq_id = id;
    *n_tx = len; *coverage = (uint32_t)cov;
    ev->aligned++; ev->out_members += len;
    free(heap);
    return 1;
}

/* process_reads inner loop -- ref :449-462 (the per-record work of one worker thread). */
int orc_map_batch(const orc_index* ix, const uint64_t* read_words, const uint64_t* read_off,
                  const uint32_t* read_len, uint64_t n_reads, orc_hit* hits, uint32_t* tx_buf,
                  uint64_t tx_cap, uint64_t* tx_used, uint64_t* counts, orc_events* ev) {
    return orc_map_batch_with_mismatch(ix, read_words, read_off, read_len, n_reads, DEFAULT_ALLOWED_MISMATCHES, hits,
                                       tx_buf, tx_cap, tx_used, counts, ev);
}

int orc_map_batch_with_mismatch(const orc_index* ix, const uint64_t* read_words, const uint64_t* read_off,
                                const uint32_t* read_len, uint64_t n_reads, uint32_t allowed_mismatches, orc_hit* hits,
                                uint32_t* tx_buf, uint64_t tx_cap, uint64_t* tx_used, uint64_t* counts, orc_events* ev) {
    orc_events local; if (!ev) { memset(&local, 0, sizeof local); ev = &local; }
    uint64_t used = 0; int overflow = 0;
    for (uint64_t i = 0; i < n_reads; i++) {
        uint32_t n_tx = 0, cov = 0, eq = ORC_EQ_NONE;
        uint64_t room = used <= tx_cap ? tx_cap - used : 0;
        int r = orc_map_read_with_mismatch(ix, read_words + read_off[i], read_len[i], allowed_mismatches,
                                           tx_buf ? tx_buf + used : NULL, overflow ? 0 : room, &n_tx, &cov, &eq, NULL, 0, NULL, ev);
        if (r < 0) { overflow = 1; /* keep counting the need with a scratch pass */
            uint32_t* tmp = (uint32_t*)malloc(((size_t)ix->eq_offsets[ix->n_eq] + 1) * 4);
            r = orc_map_read_with_mismatch(ix, read_words + read_off[i], read_len[i], allowed_mismatches, tmp, ix->eq_offsets[ix->n_eq], &n_tx, &cov, &eq, NULL, 0, NULL, NULL);
            free(tmp);
        }
        hits[i].coverage = cov; hits[i].n_tx = n_tx; hits[i].tx_off = used; hits[i].eq_id = eq;
        uint32_t flags = 0;
        if (r > 0) {
            flags |= ORC_FLAG_ALIGNED;
            if (cov >= READ_COVERAGE_THRESHOLD && n_tx == 0) flags |= ORC_FLAG_MAPPED; /* :455 QUIRK-4 */
        }
        hits[i].flags = flags;
        used += n_tx;
        if (counts) {
            if (r <= 0) counts[ix->n_eq + 1]++;          /* unaligned */
            else if (eq == ORC_EQ_NONE) counts[ix->n_eq]++; /* aligned, set is not a visited class */
            else counts[eq]++;
        }
    }
    *tx_used = used;
    return overflow ? -1 : 0;
}

/* process_reads worker body over ASCII records (ref :449-462): DnaString::from_dna_string (:449-450)
 * then map_read (:451), reads of `len` bytes at ascii + i*stride.  Same outputs as orc_map_batch. */
int orc_map_ascii_batch(const orc_index* ix, const uint8_t* ascii, uint64_t stride, uint32_t len, uint64_t n_reads,
                        uint32_t allowed_mismatches, orc_hit* hits, uint32_t* tx_buf, uint64_t tx_cap,
                        uint64_t* tx_used, uint64_t* counts, orc_events* ev) {
    uint64_t nw = orc_words_for(len) + 2;
    uint64_t* words = (uint64_t*)calloc(nw, 8);
    if (!words) return -2;
    uint64_t used = 0; int overflow = 0;
    const uint64_t zero = 0;
    for (uint64_t i = 0; i < n_reads; i++) {
        orc_pack_ascii((const char*)ascii + i * stride, len, words);             /* :449-450 */
        uint64_t u = 0;
        int r = orc_map_batch_with_mismatch(ix, words, &zero, &len, 1, allowed_mismatches, hits + i,
                                            tx_buf ? tx_buf + used : NULL, overflow || used > tx_cap ? 0 : tx_cap - used,
                                            &u, counts, ev);                        /* :451-462 */
        if (r < 0) overflow = 1;
        hits[i].tx_off = used;
        used += u;
    }
    free(words);
    *tx_used = used;
    return overflow ? -1 : 0;
}

/* Order-independent checksum of a result batch (verification aid; include/psa.h psa_result_checksum
 * restates it for device buffers): sum over reads of a hash chain over (global read index,
 * coverage, flags, eq_id, members in order). */
uint64_t orc_result_checksum(const orc_hit* hits, const uint32_t* tx, uint64_t n, uint64_t first_index) {
    uint64_t sum = 0;
    for (uint64_t i = 0; i < n; i++) {
        uint64_t h = mix64(first_index + i + 0x9E3779B97F4A7C15ULL);
        h = mix64(h ^ (((uint64_t)hits[i].coverage << 32) | hits[i].flags));
        h = mix64(h ^ (((uint64_t)hits[i].n_tx << 32) | hits[i].eq_id));
        for (uint32_t j = 0; j < hits[i].n_tx; j++) h = mix64(h ^ ((uint64_t)tx[hits[i].tx_off + j] + 0xD6E8FEB86659FD93ULL));
        sum += h;
    }
    return sum;
}
