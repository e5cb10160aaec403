/*
 * psa.h -- C ABI of the B200-native pseudoalignment hot path.
 *
 * This is the drop-in boundary for ONE path of 10XGenomics/rust-pseudoaligner (crate
 * debruijn_mapping 0.6.0, commit 9d9cab8): Pseudoaligner::map_read and the process_reads
 * inner loop.  The reference has no FFI of its own (everything is generic Rust), so every
 * entry point below names the reference item it replaces; "ref" = /root/reference.
 * INTEGRATION.md shows the Rust `extern "C"` block and the patch to process_reads that
 * binds them.
 *
 * Conventions: plain pointers and sizes, no C++/torch types; 0 = success, negative = error
 * (psa_strerror); no exceptions or aborts cross the ABI; the caller owns every host buffer,
 * the library owns every device buffer; nothing here ever falls back to the CPU -- without a
 * CUDA device the calls that need one return PSA_ERR_CUDA.
 */
#ifndef PSA_H
#define PSA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PSA_ABI_VERSION 2

/* ---- status codes ---- */
#define PSA_OK 0
#define PSA_ERR_ARG (-1)      /* bad argument                                              */
#define PSA_ERR_CUDA (-2)     /* CUDA runtime error (psa_last_error has the text)          */
#define PSA_ERR_NOMEM (-3)    /* host allocation failed                                    */
#define PSA_ERR_CAPACITY (-4) /* tx_buf too small; *tx_used holds the size needed          */
#define PSA_ERR_INDEX (-5)    /* graph is not a valid index: duplicate k-mer, ext bit whose
                                 neighbour is missing ("missing link" in debruijn), eq id
                                 out of range                                              */
#define PSA_ERR_NCCL (-6)     /* NCCL error or libnccl not loadable                        */
#define PSA_ERR_IO (-7)       /* file could not be read / parsed                           */
#define PSA_ERR_INTERNAL (-8) /* an internal self-check failed (never a wrong answer)      */

const char* psa_strerror(int code);
/* Text of the last error raised on the calling thread ("" if none). */
const char* psa_last_error(void);
int psa_abi_version(void);

/* ---- constants of the path (ref src/config.rs:16-18, stride literal src/pseudoaligner.rs:110) */
#define PSA_READ_COVERAGE_THRESHOLD 32u
#define PSA_DEFAULT_ALLOWED_MISMATCHES 2u
#define PSA_SEED_STRIDE 3u
#define PSA_EQ_NONE 0xFFFFFFFFu

/* ------------------------------------------------------------------------------------------
 * The index.  Replaces the read-only use of `Pseudoaligner<K>` (ref src/pseudoaligner.rs:26-33)
 * at map time.  The caller flattens `dbg` and `eq_classes` through debruijn's public accessors
 * (dbg.len(), get_node(i).sequence()/exts()/data()); `dbg_index` (boomphf's private
 * bit-vectors) is NOT passed: psa_index_create rebuilds a minimal perfect hash of every
 * k-mer -> (node, offset) on the GPU (what make_dbg_index does, ref src/build_index.rs:182-221)
 * together with the successor/predecessor tables that Node::r_edges()/l_edges() compute on
 * the fly.  That is legal because every MPHF answer is verified against the unitig
 * (ref src/pseudoaligner.rs:99-107), so only dictionary membership reaches the output.
 * ---------------------------------------------------------------------------------------- */
typedef struct psa_index psa_index; /* opaque, immutable after create, shareable across threads */

typedef struct psa_index_desc { /* all pointers host memory, borrowed for the call only */
    uint32_t k;                 /* K::k(), 2..64.  k <= 32 runs the 64-bit k-mer kernels,
                                   k <= 64 the 128-bit ones (Kmer64 of the CLI)             */
    uint32_t reserved;
    uint64_t n_nodes;           /* dbg.len()                                                */
    const uint64_t* seq_words;  /* all node sequences concatenated, debruijn DnaString
                                   packing: base i in word i/32 at bits 62-2*(i%32),
                                   A0 C1 G2 T3                                              */
    uint64_t n_seq_words;
    const uint64_t* node_start; /* first base of node i in the concatenation                */
    const uint32_t* node_len;   /* node.sequence().len(), >= k                              */
    const uint8_t* node_exts;   /* bit b = right extension with base b, bit 4+b = left
                                   extension with base b (debruijn Exts{val})               */
    const uint32_t* node_eq;    /* *node.data(): equivalence-class id                       */
    uint64_t n_eq;              /* eq_classes.len()                                         */
    const uint64_t* eq_offsets; /* n_eq+1, CSR over eq_members                              */
    const uint32_t* eq_members; /* eq_classes[c]: ascending, unique transcript ids
                                   (ref src/equiv_classes.rs:78-79)                         */
} psa_index_desc;

typedef struct psa_index_info {
    uint32_t k;
    uint32_t dict_levels; /* levels of the k-mer dictionary's bucket cascade                 */
    uint64_t n_nodes, n_kmers, n_eq, n_eq_members, n_seq_words;
    uint64_t dict_bytes, node_bytes, seq_bytes, eq_bytes; /* device residency               */
    uint32_t node_bits, pos_bits, fp_bits; /* layout of one 64-bit dictionary entry         */
    uint32_t max_class_len;
    double gamma;
    double build_ms; /* device time of the dictionary/edge build                            */
} psa_index_info;

/* gamma: dictionary slots per k-mer at every level of the cascade; <= 0 selects 1.7, the
 * reference's MPHF load parameter (ref src/build_index.rs:197). */
int psa_index_create(const psa_index_desc* desc, int device, double gamma, psa_index** out);
void psa_index_destroy(psa_index*);
int psa_index_get_info(const psa_index*, psa_index_info* out);
/* eq_classes as the index keeps them on the host (CSR, borrowed until psa_index_destroy): what compact results
 * (below) are expanded from. */
int psa_index_host_classes(const psa_index*, const uint64_t** eq_offsets, const uint32_t** eq_members, uint64_t* n_eq);
/* dbg_index.get(kmer) + verification (ref src/pseudoaligner.rs:96-107) for n k-mers given
 * as packed words (k bases from base 0; 1 word per k-mer if k<=32 else 2).  found[i] = 1 and
 * node/off filled when the k-mer is in the graph.  Diagnostic / test entry. */
int psa_index_lookup(psa_index*, const uint64_t* kmer_words, uint64_t n, uint8_t* found,
                     uint32_t* node, uint32_t* off);

/* ------------------------------------------------------------------------------------------
 * Reads and results
 * ---------------------------------------------------------------------------------------- */
#define PSA_READS_ASCII 0u  /* record.seq() bytes; packed on the GPU exactly as
                               DnaString::from_dna_string does at ref src/pseudoaligner.rs:449-450:
                               A/a 0, C/c 1, G/g 2, T/t 3, anything else 0                  */
#define PSA_READS_PACKED 1u /* DnaString storage words, read i starting at word read_off[i] */
#define PSA_MEM_HOST 0u
#define PSA_MEM_DEVICE 1u

typedef struct psa_read_batch {
    uint32_t format;          /* PSA_READS_*                                                */
    uint32_t location;        /* PSA_MEM_*: where data/read_off/read_len live               */
    const void* data;         /* bytes (ASCII) or uint64 words (PACKED)                     */
    uint64_t data_len;        /* in bytes / words                                           */
    const uint64_t* read_off; /* n_reads offsets into data (bytes / words); NULL: read i
                                 starts at i*stride                                          */
    const uint32_t* read_len; /* n_reads lengths in bases; NULL: every read has fixed_len    */
    uint64_t stride;          /* used when read_off == NULL                                 */
    uint32_t fixed_len;       /* used when read_len == NULL                                 */
    uint32_t reserved;
    uint64_t n_reads;
} psa_read_batch;

#define PSA_FLAG_ALIGNED 1u /* map_read returned Some(..)                                    */
#define PSA_FLAG_MAPPED 2u  /* the bool process_reads prints: coverage >= 32 &&
                               eq_class.is_empty() (sic, ref src/pseudoaligner.rs:455)       */

typedef struct psa_hit {    /* one per read, input order                                    */
    uint32_t coverage;      /* read_coverage; 0 for None                                    */
    uint32_t n_tx;          /* eq_class.len()                                               */
    uint64_t tx_off;        /* eq_class = tx_buf[tx_off .. tx_off+n_tx), ascending          */
    uint32_t eq_id;         /* id of the index class equal to eq_class when one of the
                               visited nodes carries it (smallest such id), else PSA_EQ_NONE */
    uint32_t flags;         /* PSA_FLAG_*                                                   */
} psa_hit;

/* Compact results: what crosses PCIe per read is 8 bytes instead of a psa_hit and every member.  eq_or_n is the
 * index class equal to the read's eq_class, or 0x80000000 + |eq_class| when the set is no index class (only THOSE
 * sets' members are then delivered in tx_buf, in read order), or PSA_EQ_NONE for None; cov_flags = coverage (28 bits)
 * | PSA_FLAG_* << 28.  psa_expand_compact turns them into psa_hit + members on the host from a copy of eq_classes. */
typedef struct psa_hit_compact {
    uint32_t eq_or_n;
    uint32_t cov_flags;
} psa_hit_compact;
#define PSA_RESULT_COMPACT 1u

typedef struct psa_result_batch {
    uint32_t location; /* PSA_MEM_*: where hits / tx_buf live                                */
    uint32_t flags;    /* PSA_RESULT_COMPACT: hits points to psa_hit_compact[n_reads]        */
    psa_hit* hits;     /* n_reads                                                            */
    uint32_t* tx_buf;  /* members of every eq_class back to back in read order; NULL = skip  */
    uint64_t tx_cap;   /* capacity of tx_buf in entries                                      */
    uint64_t tx_used;  /* out: entries written (or needed, with PSA_ERR_CAPACITY)            */
} psa_result_batch;

/* ------------------------------------------------------------------------------------------
 * The mapper: mutable per-caller state (device staging, per-class counts).  One per host
 * thread or pipeline; many mappers may share one index.  Replaces the worker-thread body of
 * process_reads (ref src/pseudoaligner.rs:440-472) for a whole batch of records at once.
 * ---------------------------------------------------------------------------------------- */
typedef struct psa_mapper psa_mapper;

/* chunk_reads: reads per internal pipeline chunk for host-resident batches (0 = default). */
int psa_mapper_create(psa_index*, uint64_t chunk_reads, psa_mapper** out);
void psa_mapper_destroy(psa_mapper*);
/* ref src/pseudoaligner.rs:382: DEFAULT_ALLOWED_MISMATCHES; map_read_with_mismatch (:361) takes any. */
int psa_mapper_set_allowed_mismatches(psa_mapper*, uint32_t allowed);
/* Tuning: lanes of a warp that cooperate on one read (8, 16 or 32; default 8, or the
 * PSA_GROUP_WIDTH environment variable).  Results do not depend on it. */
int psa_mapper_set_group_width(psa_mapper*, uint32_t lanes);
/* Tuning: the map step is two kernels.  k_map_thread gives every read one thread and hands a
 * read over to the cooperative kernel (k_map, group_width lanes per read) when its seed search
 * needs more than max_probes positions, it meets more wide classes than a thread keeps, or all
 * its classes are wide and the smallest has more than max_small members.
 * max_probes = 0 sends every read to the cooperative kernel.  Defaults 3 (ceil(k/3)+2 for an index without the seed-scan filter) / 32
 * (PSA_FAST_PROBES / PSA_FAST_MAX_SMALL).  Results do not depend on it. */
int psa_mapper_set_fast_path(psa_mapper*, uint32_t max_probes, uint32_t max_small);
/* Tuning: reads whose FIRST seed search is too long for one thread go to k_seed_scan, where
 * `lanes` lanes (8, 16 or 32; default 8, PSA_SCAN_WIDTH) probe as many stride-3 positions at
 * once; a read without any seed ends there, a seeded one returns to the thread-per-read kernel
 * with the answer.  0: such reads go to the cooperative kernel instead.  Results do not depend
 * on it. */
int psa_mapper_set_scan_width(psa_mapper*, uint32_t lanes);

/* Pseudoalign a batch.  For every read i: hits[i] and its members in tx_buf are exactly
 * what `index.map_read(&DnaString::from_dna_string(seq_i))` returns (ref :381-384, :449-462).
 * Host batches are cut into chunks and pipelined (H2D | kernels | D2H on separate streams);
 * device batches run in place on the mapper's stream.  Synchronous: results are complete
 * on return.  Per-class counts accumulate across calls until psa_mapper_counts_reset. */
int psa_mapper_map(psa_mapper*, const psa_read_batch* reads, psa_result_batch* results);

/* Host function (no device): compact records + the members of the non-class sets -> psa_hit[] and every member. */
int psa_expand_compact(const psa_hit_compact* in, uint64_t n, const uint32_t* novel_tx, uint64_t novel_tx_len,
                       const uint64_t* eq_offsets, const uint32_t* eq_members, uint64_t n_eq, psa_hit* hits,
                       uint32_t* tx_buf, uint64_t tx_cap, uint64_t* tx_used);

/* Device-resident batch, asynchronous on the mapper's stream, fixed shape helpers for
 * benchmarking: same as psa_mapper_map with location == PSA_MEM_DEVICE but does not
 * synchronise; results->tx_used is valid after psa_mapper_sync.  Several batches may be queued
 * before one psa_mapper_sync: it then reports an overflow of ANY of them (PSA_ERR_CAPACITY; the
 * buffers have been grown -- reset the counts and resubmit), and tx_used of the last one. */
int psa_mapper_map_async(psa_mapper*, const psa_read_batch* reads, psa_result_batch* results);
int psa_mapper_sync(psa_mapper*);
void* psa_mapper_stream(psa_mapper*); /* cudaStream_t */

/* Pseudoaligner::map_read for one read (ref src/pseudoaligner.rs:381): returns 1 = Some,
 * 0 = None, <0 error.  Convenience over psa_mapper_map; one launch per call. */
int psa_mapper_map_read(psa_mapper*, const uint64_t* read_words, uint32_t read_len,
                        uint32_t* tx_out, uint64_t tx_cap, uint32_t* n_tx, uint32_t* coverage);

/* counts[c] for c < n_eq: reads whose eq_class equals index class c; counts[n_eq+1]: None.
 * Reads whose eq_class is NO index class are counted per distinct set in the mapper's novel-set
 * table (below); counts[n_eq] is the sum of that table, kept so that the array always sums to the
 * reads mapped.  The array has n_eq+2 entries. */
int psa_mapper_counts_get(psa_mapper*, uint64_t* counts_host);
int psa_mapper_counts_reset(psa_mapper*);
void* psa_mapper_counts_device(psa_mapper*); /* uint64[n_eq+2] in HBM */

/* Novel sets.  map_read returns the transcript SET (ref src/pseudoaligner.rs:323-356, :381-384); most
 * sets equal the class of a visited node and carry that class's id (psa_hit.eq_id), the others --
 * intersections that no k-mer of the index has as its colour, incl. the empty set -- have
 * eq_id == PSA_EQ_NONE in psa_hit and are counted here, per distinct set, since the last
 * psa_mapper_counts_reset.  The view is sorted by (length, contents): the id of set i is n_eq + i.
 * Ids of this kind are final only once the run is complete (a later batch may bring a set that sorts
 * earlier); they are deterministic -- the same reads give the same table whatever the batching, the
 * kernel split or the number of GPUs (psa_novel_sets_merge / psa_mapper_novel_allgather). */
typedef struct psa_novel_sets {
    uint64_t n_sets, n_members;
    uint64_t* offsets; /* n_sets + 1                                                          */
    uint32_t* members; /* set i = members[offsets[i] .. offsets[i+1]), ascending              */
    uint64_t* counts;  /* reads whose eq_class is set i                                       */
} psa_novel_sets;
int psa_mapper_novel_sets(psa_mapper*, psa_novel_sets* out); /* arrays malloc'ed: psa_novel_sets_free */
/* The union of several views (e.g. one per GPU), equal sets merged, counts added, sorted again. */
int psa_novel_sets_merge(const psa_novel_sets* parts, uint32_t n_parts, psa_novel_sets* out);
void psa_novel_sets_free(psa_novel_sets*);

/* Work actually done by the mapper's kernels since the last reset, in the units of the
 * algorithmic-bytes model of DESIGN.md (sequential-equivalent events: speculative probes
 * of the warp-wide seed scan are not counted). Filled only by psa_mapper_map_events. */
typedef struct psa_events {
    uint64_t reads, read_bases;
    uint64_t kmer_lookups;   /* P: the reference's own counter (src/pseudoaligner.rs:95)     */
    uint64_t dict_levels;    /* dictionary buckets probed                                    */
    uint64_t dict_hits;      /* buckets with an entry carrying the k-mer's fingerprint       */
    uint64_t verifications;  /* unitig k-mer fetched and compared                            */
    uint64_t node_visits;    /* nodes.push                                                   */
    uint64_t bases_compared; /* base compares of both extension loops                        */
    uint64_t edge_jumps;
    uint64_t class_members;  /* members of visited distinct classes read by the intersection */
    uint64_t out_members;    /* sum |eq_class|                                               */
    uint64_t aligned;
} psa_events;
/* Same as psa_mapper_map on a device batch, with event counting compiled in (slower).
 * out[0]: the reads completed by k_map_lanes, out[1]: by k_map, out[2]: by k_seed_scan (reads
 * without a seed) plus the first seed searches it made for reads k_map_lanes completed. */
int psa_mapper_map_events(psa_mapper*, const psa_read_batch* reads, psa_result_batch* results,
                          psa_events out[3]);

/* After psa_mapper_map_events: why k_map_lanes handed reads over -- [0] first seed search
 * longer than max_probes, [1] re-seed search longer than max_probes, [2] wide-class list
 * full, [3] smallest class longer than max_small.  Diagnostic. */
int psa_mapper_defer_reasons(psa_mapper*, uint64_t out[4]);

/* Kernels launched by this mapper since creation (bench.py's gpu_launches). */
uint64_t psa_mapper_launch_count(const psa_mapper*);
/* Device timing of the map kernels alone: when enabled, every launch is bracketed by CUDA
 * events on the mapper's stream.  psa_mapper_profile_read synchronises, returns the summed
 * kernel time and launch count since the last read ([0] k_map_lanes, both passes; [1] k_map;
 * [2] k_seed_scan), and resets them. */
int psa_mapper_profile_enable(psa_mapper*, int on);
int psa_mapper_profile_read(psa_mapper*, double map_kernel_ms[3], uint64_t map_launches[3]);

/* ------------------------------------------------------------------------------------------
 * process_reads (ref src/pseudoaligner.rs:420-514): FASTQ file (plain or gzip) in, one line per
 * read out -- `(flag, "id", [tx, ...], coverage)`, the `{:?}` of the tuple the reference
 * println!s at :490 -- in input order.  out_path NULL or "-" = stdout.  num_threads = host
 * threads that format the lines (the reference's num_threads are its mapping workers; mapping
 * is on the GPU here) and read / write the files.  batch_reads 0 = 512 Ki reads per block.  progress != 0 prints the
 * reference's stderr tick at every 1 000 000th read (:497-504).  Records are cut exactly as the reference's reader
 * (bio::io::fastq 1.5) cuts them: the id is the header without its '@' up to the first SPACE (tabs stay), sequence
 * lines may be wrapped (every line up to the '+' line, each stripped of trailing white space), as many quality lines
 * follow.  A malformed record -- no '@', no quality, a blank line, a header that is not UTF-8 -- returns PSA_ERR_IO
 * after the records before it were processed (the reference panics, :446).  Sequence and quality bytes >= 0x80 are
 * not checked for UTF-8 validity (the reference's reader would fail on them).
 * ---------------------------------------------------------------------------------------- */
typedef struct psa_process_stats {
    uint64_t reads;   /* read_counter, :476                                                  */
    uint64_t mapped;  /* mapped_read_counter, :477 (the flag of :455, sic)                   */
    uint64_t aligned; /* reads for which map_read returned Some                              */
    double seconds;   /* wall time of the call                                               */
    double reader_seconds, mapper_seconds, writer_seconds; /* busy time of the three pipeline
                         stages (FASTQ parse | psa_mapper_map | format + write)              */
} psa_process_stats;
/* `{:?}` of a string exactly as the drivers print ids (Rust's escape_debug; ASCII exact, see process_reads.cpp for
 * the non-ASCII subset): returns the bytes written to out (cap >= 10 n + 2), negative if s is not valid UTF-8. */
int64_t psa_debug_str(const char* s, uint64_t n, char* out, uint64_t cap);
int psa_process_reads(psa_index*, const char* fastq_path, const char* out_path, uint32_t num_threads,
                      uint64_t batch_reads, int progress, psa_process_stats* stats);

/* ---- multi-GPU: reads shard across ranks, one all-reduce of the per-class counts ---- */
typedef struct psa_comm psa_comm;
#define PSA_NCCL_UNIQUE_ID_BYTES 128
int psa_comm_unique_id(uint8_t id[PSA_NCCL_UNIQUE_ID_BYTES]); /* rank 0; ship to the others */
int psa_comm_create(const uint8_t id[PSA_NCCL_UNIQUE_ID_BYTES], int world, int rank, int device,
                    psa_comm** out);
void psa_comm_destroy(psa_comm*);
/* ncclAllReduce(sum, uint64, n_eq+2) in place on the mapper's counts, on its stream. */
int psa_mapper_counts_allreduce(psa_mapper*, psa_comm*);
/* The novel-set tables of all ranks merged (two ncclAllGather: sizes, then the serialised tables):
 * every rank receives the same view, ids n_eq + i valid across the whole job. */
int psa_mapper_novel_allgather(psa_mapper*, psa_comm*, psa_novel_sets* out);

/* ---- measurement aid: GB/s of independent random gathers of chunk_bytes-sized (32, 64 or 128)
 * aligned chunks from a table_bytes table in HBM, one chunk per thread -- the access pattern of
 * the index lookups without their dependencies; the practical ceiling the map kernels are
 * compared with.  chunk_bytes = 0: one random 128-byte line per warp (4 bytes per lane). ---- */
int psa_gather_probe(int device, uint64_t table_bytes, uint32_t chunk_bytes, uint32_t iters, double* gbytes_per_s);

/* ---- index construction on the device (the reference: make_dbg / debruijn's compression on the host, ref
 * src/build_index.rs:27-179, src/equiv_classes.rs:62-91): every k-mer of every transcript of at least k bases (stranded),
 * radix-sorted; colour = ascending de-duplicated transcript list, interned in order of first appearance over the sorted k-mers;
 * exts = union over occurrences; unitigs = maximal paths of unique, same-colour links (ScmapCompress).  codes: one byte per
 * base (0..3), transcript t = codes[tx_off[t] .. tx_off[t+1]).  The arrays returned (host memory, psa_built_graph_free) are
 * exactly the fields of psa_index_desc and bit-identical to those of the host builder (psa_build_graph, psa_host.h). ---- */
typedef struct psa_built_graph {
    uint32_t k, reserved;
    uint64_t n_nodes, n_kmers, n_eq, n_seq_words, n_eq_members, n_cycles;
    uint64_t* seq_words;
    uint64_t* node_start;
    uint32_t* node_len;
    uint8_t* node_exts;
    uint32_t* node_eq;
    uint64_t* eq_offsets;  /* n_eq + 1 */
    uint32_t* eq_members;
} psa_built_graph;
int psa_build_graph_device(int device, const uint8_t* codes, const uint64_t* tx_off, uint32_t n_tx, uint32_t k, psa_built_graph* out);
void psa_built_graph_free(psa_built_graph*);

/* ---- mappability::analyze_graph (ref src/mappability.rs:120-156) on the device.  tx_gene[t] = an integer naming the
 * gene of transcript t (ref: tx_gene_mapping, src/utils.rs:83-86); bins = MAPPABILITY_COUNTS_LEN (ref src/config.rs: 11).
 * Out (host, n_tx x bins each, row-major): tx_multiplicity[t][j] = k-mers of unitigs whose class contains t and has j + 1
 * transcripts (the last bin: bins or more -- ref :59-65 puts multiplicity == bins there too), gene_multiplicity likewise
 * by the class's number of distinct genes.  total_kmer_count / fraction_unique_* (ref :54-82) are sums and ratios of a row. ---- */
#define PSA_MAPPABILITY_COUNTS_LEN 11u
int psa_index_mappability(psa_index*, const uint32_t* tx_gene, uint32_t n_tx, uint32_t bins, uint64_t* tx_multiplicity,
                          uint64_t* gene_multiplicity);

/* ---- measurement aid: the synthetic read stream of psa_host.h (psa_synth_reads: 90 % transcript reads with 0.5 %
 * substitutions, 5 % chimeric, 5 % random) generated ON THE DEVICE by the same counter-based generator -- read i is a
 * function of (seed, i) only, so BASELINE config 5 (10^9 reads) needs no host generation and any sample of the stream can
 * be regenerated on the host for the oracle.  All pointers are device memory: codes / tx_off = the transcriptome (2-bit
 * codes, one per byte); elig[w] / cum[w] (n_elig[w] and n_elig[w] + 1 entries) = the transcripts of at least len_w bases
 * and the running count of their start positions, for len_0 = L, len_1 = L / 2, len_2 = L - L / 2.  out: n reads of L
 * ASCII bytes, `stride` bytes apart.  Synchronous. ---- */
typedef struct psa_synth_tables {
    const uint8_t* codes;
    const uint64_t* tx_off;
    const uint32_t* elig[3];
    const uint64_t* cum[3];
    uint64_t n_elig[3];
} psa_synth_tables;
int psa_synth_reads_device(int device, const psa_synth_tables* tables, uint64_t seed, uint64_t first, uint64_t n, uint32_t L,
                           uint8_t* out_dev, uint64_t stride);

/* ---- self-test entry: the device routines behind nodes_to_eq_class / intersect (ref src/pseudoaligner.rs:323-356,
 * :389-418) on two ascending lists in isolation: out[0..cap) one thread's list scheme, out[cap..2cap) the cooperative
 * kernel's lane-group scheme, out[2cap..3cap) the class windows (n_out[2] = PSA_EQ_NONE when a list spans >= 192 ids
 * and has no window).  The reference's own intersect vectors (ref :544-559) are run through it by the tests. ---- */
int psa_selftest_intersect(int device, const uint32_t* v1, uint32_t n1, const uint32_t* v2, uint32_t n2,
                           uint32_t* out, uint32_t cap, uint32_t n_out[3]);

/* ---- verification aid: order-independent 64-bit checksum of a device-resident result batch (the sum over
 * reads of a hash chain over global read index first_index + i, coverage, flags, eq_id and the members in
 * tx_buf[tx_off ..)); bench.py compares it with the same function of the oracle's results. ---- */
int psa_result_checksum(int device, const psa_hit* hits_dev, const uint32_t* tx_dev, uint64_t n, uint64_t first_index,
                        uint64_t* out);

/* ---- pinned host memory for the batch buffers (pageable memory works, slower) ---- */
int psa_host_alloc(void** out, uint64_t bytes);
void psa_host_free(void*);
/* ---- raw device memory, for callers that keep batches resident in HBM ---- */
int psa_device_alloc(void** out, uint64_t bytes);
void psa_device_free(void*);
int psa_memcpy_h2d(void* dst_dev, const void* src_host, uint64_t bytes);
int psa_memcpy_d2h(void* dst_host, const void* src_dev, uint64_t bytes);

#ifdef __cplusplus
}
#endif
#endif /* PSA_H */
