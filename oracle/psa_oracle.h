/*
 * psa_oracle.h -- CPU ORACLE for the per-read pseudoalignment hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, the smoke check in
 * __graft_entry__.py and bench.py's cpu_baseline / --impl reference legs may load it.
 * The product (rust-pseudoaligner_b200/, include/psa.h) never links or calls it.
 *
 * It restates, in plain C, the algorithm of 10XGenomics/rust-pseudoaligner
 * (crate debruijn_mapping 0.6.0, commit 9d9cab8):
 *   src/pseudoaligner.rs:64-319   map_read_to_nodes_with_mismatch
 *   src/pseudoaligner.rs:323-356  nodes_to_eq_class
 *   src/pseudoaligner.rs:389-418  intersect
 *   src/pseudoaligner.rs:453-462  process_reads "mapped" flag
 *   src/config.rs:16-18           READ_COVERAGE_THRESHOLD / LEFT_EXTEND_FRACTION / DEFAULT_ALLOWED_MISMATCHES
 * plus a deliberately simple (sort + binary search) builder of the coloured compacted
 * de Bruijn graph whose *semantics* follow src/build_index.rs:127-221 and
 * src/equiv_classes.rs:62-91 (colour = sorted set of transcripts containing the k-mer,
 * exts = union of observed neighbours, stranded, unitig = maximal non-branching
 * same-colour path).
 *
 * Third-party arithmetic that is NOT under /root/reference (so restated from the
 * published algorithms, not transliterated): debruijn 0.3.4 @ 8d9a5c5 (2-bit packing,
 * k-mer integer form, Exts, ScmapCompress compaction rule), boomphf 0.6.0 / wyhash 0.5.0
 * (k-mer -> slot dictionary; replaced here by an exact open-addressing table -- legal
 * because every MPHF answer is verified against the unitig, src/pseudoaligner.rs:99-107,
 * so only dictionary *membership* reaches the output).
 *
 * Parity pins (see tests/test_oracle_*.py): validate_dbg (a)+(b) of
 * src/build_index.rs:262-368 at k=20 and k=64 on test/gencode_small.fa, the
 * test_alignment known answers (src/build_index.rs:429-441), the intersect vectors and
 * property test (src/pseudoaligner.rs:542-586).  Behaviour the reference never tests
 * (left extension, re-seed, >2 mismatches per unitig, small.fq) is pinned only by the
 * line-by-line restatement: "parity unpinned" for those.
 */
#ifndef PSA_ORACLE_H
#define PSA_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_EQ_NONE 0xFFFFFFFFu

#define ORC_FLAG_ALIGNED 1u /* map_read returned Some(..)                          */
#define ORC_FLAG_MAPPED 2u  /* process_reads flag: coverage >= 32 && eq_class.is_empty() (sic, :455) */

typedef struct orc_index orc_index;

/* Same field layout as psa_hit in include/psa.h so tests can compare raw buffers. */
typedef struct {
    uint32_t coverage; /* read_coverage (0 for None)                                        */
    uint32_t n_tx;     /* |eq_class|                                                        */
    uint64_t tx_off;   /* offset of the members in the caller's tx buffer                   */
    uint32_t eq_id;    /* min id of a visited class equal to the result set, else ORC_EQ_NONE */
    uint32_t flags;    /* ORC_FLAG_*                                                        */
} orc_hit;

/* Per-read event counters behind the algorithmic-bytes figure. */
typedef struct {
    uint64_t reads;
    uint64_t read_bases;     /* sum of read lengths                                    */
    uint64_t kmer_lookups;   /* P  (the reference's own counter, pseudoaligner.rs:95)  */
    uint64_t dict_hits;      /* lookups whose k-mer is in the graph                    */
    uint64_t node_visits;    /* V  nodes.push                                          */
    uint64_t bases_compared; /* Bc base-by-base compares in both extension loops       */
    uint64_t edge_jumps;     /* J  l_edges()/r_edges() uses                            */
    uint64_t class_members;  /* E  sum of |class(node)| over visited nodes             */
    uint64_t out_members;    /* R  sum of |result|                                     */
    uint64_t aligned;        /* reads with Some(..)                                    */
} orc_events;

/* ---- sequence packing (debruijn DnaString; call site pseudoaligner.rs:449-450) ---- */
/* ASCII -> 2-bit, A0 C1 G2 T3 (case-insensitive), anything else -> 0 (QUIRK-6).
 * Base i sits in word i/32 at bits 62-2*(i%32).  words must hold (len+31)/32 entries. */
void orc_pack_ascii(const char* s, uint64_t len, uint64_t* words);
uint64_t orc_words_for(uint64_t n_bases);

/* ---- index ---- */
/* Naive builder: transcripts given as one byte per base (values 0..3), transcript t
 * spanning codes[tx_off[t] .. tx_off[t+1]).  Transcripts shorter than k contribute
 * nothing (build_index.rs:134).  Returns NULL on failure (message via orc_last_error). */
orc_index* orc_index_build(const uint8_t* codes, const uint64_t* tx_off, uint32_t n_tx, uint32_t k);

/* Wrap an index given in the flat form of psa_index_desc (include/psa.h); arrays are copied. */
orc_index* orc_index_from_flat(uint32_t k, uint64_t n_nodes, const uint64_t* seq_words,
                               uint64_t n_seq_words, const uint64_t* node_start,
                               const uint32_t* node_len, const uint8_t* node_exts,
                               const uint32_t* node_eq, uint64_t n_eq,
                               const uint64_t* eq_offsets, const uint32_t* eq_members);
void orc_index_free(orc_index*);
const char* orc_last_error(void);

/* Flat-form accessors (borrowed pointers, valid until orc_index_free). */
uint32_t orc_index_k(const orc_index*);
uint64_t orc_index_n_nodes(const orc_index*);
uint64_t orc_index_n_kmers(const orc_index*);
uint64_t orc_index_n_eq(const orc_index*);
uint64_t orc_index_n_seq_words(const orc_index*);
uint64_t orc_index_n_pure_cycles(const orc_index*);
const uint64_t* orc_index_seq_words(const orc_index*);
const uint64_t* orc_index_node_start(const orc_index*);
const uint32_t* orc_index_node_len(const orc_index*);
const uint8_t* orc_index_node_exts(const orc_index*);
const uint32_t* orc_index_node_eq(const orc_index*);
const uint64_t* orc_index_eq_offsets(const orc_index*);
const uint32_t* orc_index_eq_members(const orc_index*);
const uint32_t* orc_index_succ(const orc_index*); /* 4 per node, ORC_EQ_NONE = no edge */
const uint32_t* orc_index_pred(const orc_index*);

/* k-mer dictionary probe (dbg_index.get + verification): k-mer given as 2-bit packed
 * words (k bases from base 0).  Returns 1 and (node, offset) if the k-mer is in the graph. */
int orc_index_lookup(const orc_index*, const uint64_t* kmer_words, uint32_t* node, uint32_t* offset);

/* ---- the hot path ---- */
/* map_read (pseudoaligner.rs:381-384).  Returns 1 for Some, 0 for None.  tx_out must
 * have room for the smallest visited class; *n_tx receives |eq_class|.  nodes_out (may
 * be NULL) receives the visited node ids in push order (map_read_to_nodes, :54-61). */
int orc_map_read(const orc_index*, const uint64_t* read_words, uint32_t read_len,
                 uint32_t* tx_out, uint64_t tx_cap, uint32_t* n_tx, uint32_t* coverage,
                 uint32_t* eq_id, uint32_t* nodes_out, uint32_t nodes_cap, uint32_t* n_nodes,
                 orc_events* ev);

/* map_read_with_mismatch (pseudoaligner.rs:361-376): the same with any allowed_mismatches. */
int orc_map_read_with_mismatch(const orc_index*, const uint64_t* read_words, uint32_t read_len,
                               uint32_t allowed_mismatches, uint32_t* tx_out, uint64_t tx_cap, uint32_t* n_tx,
                               uint32_t* coverage, uint32_t* eq_id, uint32_t* nodes_out, uint32_t nodes_cap,
                               uint32_t* n_nodes, orc_events* ev);

/* process_reads inner loop over a batch (pseudoaligner.rs:449-462), single thread, input
 * order.  Read i occupies read_words[read_off[i] ...] (offsets in 64-bit words).
 * Returns 0, or -1 if tx_buf is too small (tx_used then holds the required size). */
int orc_map_batch(const orc_index*, const uint64_t* read_words, const uint64_t* read_off,
                  const uint32_t* read_len, uint64_t n_reads, orc_hit* hits, uint32_t* tx_buf,
                  uint64_t tx_cap, uint64_t* tx_used, uint64_t* counts /* n_eq+2 or NULL */,
                  orc_events* ev /* may be NULL */);

int orc_map_batch_with_mismatch(const orc_index*, const uint64_t* read_words, const uint64_t* read_off,
                                const uint32_t* read_len, uint64_t n_reads, uint32_t allowed_mismatches, orc_hit* hits,
                                uint32_t* tx_buf, uint64_t tx_cap, uint64_t* tx_used, uint64_t* counts, orc_events* ev);

/* The same over ASCII records (DnaString::from_dna_string at :449-450 inside the loop, as the
 * reference's worker has it): read i = `len` bytes at ascii + i*stride. */
int orc_map_ascii_batch(const orc_index*, const uint8_t* ascii, uint64_t stride, uint32_t len, uint64_t n_reads,
                        uint32_t allowed_mismatches, orc_hit* hits, uint32_t* tx_buf, uint64_t tx_cap,
                        uint64_t* tx_used, uint64_t* counts, orc_events* ev);

/* Order-independent 64-bit checksum of a result batch (read index, coverage, flags, eq_id, members). */
uint64_t orc_result_checksum(const orc_hit* hits, const uint32_t* tx, uint64_t n, uint64_t first_index);

/* process_reads as the reference shapes it (c3_driver.c): num_threads workers, one FASTQ record per mutex
 * acquisition, a bounded channel of num_threads tuples, the main thread printing one `{:?}` line per read in
 * arrival order.  Returns 0, or -7 on an I/O error / malformed record. */
int orc_process_reads_c3(const orc_index*, const char* fastq_path, const char* out_path, uint32_t num_threads,
                         uint32_t allowed_mismatches, uint64_t* reads, uint64_t* mapped);

/* intersect (pseudoaligner.rs:389-418): in-place on v1, returns the new length. */
uint32_t orc_intersect(uint32_t* v1, uint32_t n1, const uint32_t* v2, uint32_t n2);

#ifdef __cplusplus
}
#endif
#endif
