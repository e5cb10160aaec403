/*
 * psa_host.h -- host-side companions of the GPU path (libpsa_host.so, plain C++/threads).
 *
 * In the reference these stay on the host in Rust (BASELINE.json north_star: "src/build_index.rs
 * and FASTQ I/O stay on the host"); there is no Rust toolchain in this image, so the pieces the
 * tests, the benchmark and a stand-alone user need are provided here in C++ behind a C ABI:
 *   - psa_build_*      a coloured compacted de Bruijn graph builder producing exactly the
 *                      arrays psa_index_desc wants (semantics of ref src/build_index.rs:27-221,
 *                      src/equiv_classes.rs:62-91; algorithm is this project's own)
 *   - psa_synth_*      the counter-based synthetic transcriptome / read generators of
 *                      BASELINE.md section 4
 *   - psa_fasta_* / psa_fastq_*   minimal readers (ref src/utils.rs:61-97, bio::io::fastq)
 * (The map driver, psa_process_reads, is part of libpsa_b200.so: include/psa.h.)
 * None of this is on the GPU hot path and none of it is a CPU implementation of map_read.
 */
#ifndef PSA_HOST_H
#define PSA_HOST_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

const char* psa_host_last_error(void);

/* ---- index construction (host) ---- */
typedef struct psa_graph psa_graph; /* owns the flat arrays */

/* codes: one byte per base (0..3), transcript t = codes[tx_off[t] .. tx_off[t+1]).
 * Transcripts shorter than k contribute nothing (ref src/build_index.rs:134).
 * threads <= 0: all hardware threads.  Returns NULL on failure. */
psa_graph* psa_build_graph(const uint8_t* codes, const uint64_t* tx_off, uint32_t n_tx, uint32_t k, int threads);
void psa_graph_free(psa_graph*);
uint32_t psa_graph_k(const psa_graph*);
uint64_t psa_graph_n_nodes(const psa_graph*);
uint64_t psa_graph_n_kmers(const psa_graph*);
uint64_t psa_graph_n_eq(const psa_graph*);
uint64_t psa_graph_n_seq_words(const psa_graph*);
uint64_t psa_graph_n_cycles(const psa_graph*); /* closed same-colour cycles cut at their smallest k-mer */
const uint64_t* psa_graph_seq_words(const psa_graph*);
const uint64_t* psa_graph_node_start(const psa_graph*);
const uint32_t* psa_graph_node_len(const psa_graph*);
const uint8_t* psa_graph_node_exts(const psa_graph*);
const uint32_t* psa_graph_node_eq(const psa_graph*);
const uint64_t* psa_graph_eq_offsets(const psa_graph*);
const uint32_t* psa_graph_eq_members(const psa_graph*);

/* ---- the flat index on disk (stands where the reference's bincode blob does, ref src/utils.rs:22-43):
 * exactly the arrays of psa_index_desc, raw and 64-byte aligned behind a versioned header with a
 * checksum.  psa_graph_load returns NULL on a missing, foreign, truncated or corrupt file. ---- */
psa_graph* psa_graph_from_arrays(uint32_t k, uint64_t n_nodes, const uint64_t* seq_words, uint64_t n_seq_words,
                                 const uint64_t* node_start, const uint32_t* node_len, const uint8_t* node_exts,
                                 const uint32_t* node_eq, uint64_t n_eq, const uint64_t* eq_offsets,
                                 const uint32_t* eq_members);
int psa_graph_save(const psa_graph*, const char* path); /* 0, or < 0 (psa_host_last_error) */
psa_graph* psa_graph_load(const char* path);
/* The reference's own index file (`pseudoaligner index` output: bincode 1.3 of Pseudoaligner<K>, ref
 * src/utils.rs:22-32): reads `dbg` and `eq_classes`, ignores the rest (the dictionary is rebuilt on the
 * GPU).  k = the K the index was built with (not stored in the file).  The field order of debruijn's
 * graph types is as recalled for 0.3.4 @ 8d9a5c5 -- not verifiable in this environment; every length is
 * cross-checked so that a different layout is refused (NULL) rather than misread. */
psa_graph* psa_graph_load_bincode(const char* path, uint32_t k);

/* ---- synthetic data; everything is a pure function of the seed ---- */
typedef struct psa_transcriptome psa_transcriptome;
/* GENCODE-shaped transcriptome: n_genes genes of 4..24 exons (log-normal lengths, median 140,
 * sigma 0.8, clamp 30..3000), geometric isoform counts (mean 10, max 200), each isoform an
 * ordered subset of its gene's exons (inclusion 0.75), 5 % of transcripts carry one of 50
 * repeat elements (300 b, 10 % divergence per copy).  n_genes = 20000 gives ~200 k transcripts. */
psa_transcriptome* psa_synth_transcriptome(uint64_t seed, uint32_t n_genes, int threads);
/* Wrap existing transcripts (one byte per base, 0..3) so reads can be sampled from them. */
psa_transcriptome* psa_transcriptome_from_codes(const uint8_t* codes, const uint64_t* tx_off, uint32_t n_tx);
void psa_transcriptome_free(psa_transcriptome*);
uint32_t psa_transcriptome_n_tx(const psa_transcriptome*);
uint64_t psa_transcriptome_n_bases(const psa_transcriptome*);
const uint8_t* psa_transcriptome_codes(const psa_transcriptome*);
const uint64_t* psa_transcriptome_tx_off(const psa_transcriptome*); /* n_tx + 1 */

/* Reads first .. first+n-1 of the stream `seed` as ASCII, read i at out[(i-first)*stride ..):
 * 90 % transcript reads (transcript chosen in proportion to its number of start positions,
 * each base substituted with p = 0.005), 5 % chimeric (two halves from independent
 * transcripts), 5 % uniform random.  kind_out (may be NULL) receives 0/1/2 per read. */
int psa_synth_reads(const psa_transcriptome*, uint64_t seed, uint64_t first, uint64_t n, uint32_t length,
                    uint8_t* out, uint64_t stride, uint8_t* kind_out, int threads);

/* ---- FASTA / FASTQ (plain or gzip is NOT handled here: plain text only) ---- */
typedef struct psa_seqfile psa_seqfile;
psa_seqfile* psa_fasta_read(const char* path); /* records in file order = transcript index */
psa_seqfile* psa_fastq_read(const char* path);
void psa_seqfile_free(psa_seqfile*);
uint64_t psa_seqfile_n(const psa_seqfile*);
const char* psa_seqfile_name(const psa_seqfile*, uint64_t i); /* FASTA: header after '>' ; FASTQ: id up to first space */
const uint8_t* psa_seqfile_data(const psa_seqfile*);           /* all sequences back to back (ASCII) */
const uint64_t* psa_seqfile_off(const psa_seqfile*);           /* n + 1 */

#ifdef __cplusplus
}
#endif
#endif
