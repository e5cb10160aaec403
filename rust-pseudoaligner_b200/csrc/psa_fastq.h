/* psa_fastq.h -- internal interface between psa_process_reads (process_reads.cpp) and the device-side FASTQ
 * text pipeline (psa_api.cu + psa_fastq.cuh).  Not part of the public ABI of include/psa.h: a "lane" is one
 * mapper plus the buffers of one raw FASTQ block in flight; process_reads.cpp runs a few of them side by side.
 * tests/hostsim/process_stub.cpp implements the same five functions serially on the CPU (unit-test harness). */
#ifndef PSA_FASTQ_H
#define PSA_FASTQ_H
#include <stdint.h>

#include "../../include/psa.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct psa_fq_lane psa_fq_lane;

#define PSA_FQ_MAX_TICKS 8u

typedef struct psa_fq_result {
    int plain;            /* 0: some record of the block is not a plain four-line ASCII record -- nothing was produced */
    uint64_t out_bytes;   /* bytes of result lines */
    const char* out_text; /* pinned host memory owned by the lane, valid until its next psa_fq_lane_index */
    uint64_t mapped;      /* reads with the flag of ref src/pseudoaligner.rs:455 */
    uint64_t aligned;     /* reads for which map_read returned Some */
    uint64_t end_off;     /* offset in the block of the first byte after its last record */
    uint64_t tick_mapped[PSA_FQ_MAX_TICKS]; /* "mapped" reads among the first tick_at[t] records */
} psa_fq_result;

/* block_bytes must be a multiple of 4096.  The lane's text buffer holds block_bytes + tail_bytes + 64 bytes. */
int psa_fq_lane_create(psa_index* index, uint64_t block_bytes, uint64_t tail_bytes, psa_fq_lane** out);
void psa_fq_lane_destroy(psa_fq_lane* lane);
/* pinned host buffer the caller fills with the raw bytes of the block (and its tail) */
uint8_t* psa_fq_lane_text(psa_fq_lane* lane);
/* Stage 1: copy text[0, len) to the device and index its newlines.  own_bytes (a multiple of 4096, or len) is the part of
 * the block that is not tail.  Returns the number of newlines in [0, own_bytes) and in [0, len). */
int psa_fq_lane_index(psa_fq_lane* lane, uint64_t len, uint64_t own_bytes, uint64_t* nl_own, uint64_t* nl_total);
/* Stage 2: cut the n_records records whose headers follow newlines j0, j0 + 4, ... (j0 = -1: byte 0), map them, format
 * their lines.  tick_at[t] (ascending, <= n_records): prefix lengths whose "mapped" counts are wanted. */
int psa_fq_lane_run(psa_fq_lane* lane, int64_t j0, uint64_t n_records, const uint64_t* tick_at, uint32_t n_ticks,
                    psa_fq_result* out);

#ifdef __cplusplus
}
#endif
#endif
