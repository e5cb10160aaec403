// process_stub.cpp -- UNIT-TEST HARNESS, not product code.
//
// Links rust-pseudoaligner_b200/csrc/process_reads.cpp (the C++ mirror of the reference's map driver)
// against a stand-in for the five C-ABI entry points it calls, so that its FASTQ reader, record table,
// batching, ordering and line writer can be tested without a GPU.  The stand-in "mapper" derives a fake
// result from each read's bytes (coverage = length, one transcript id = byte sum, flag from the first
// base), which lets the test check that every record reaches the mapper intact and in order.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/psa.h"

extern "C" {
int psa_host_alloc(void** out, uint64_t bytes) { *out = malloc(bytes ? bytes : 1); return *out ? PSA_OK : PSA_ERR_NOMEM; }
void psa_host_free(void* p) { free(p); }
int psa_mapper_create(psa_index*, uint64_t, psa_mapper** out) { *out = (psa_mapper*)1; return PSA_OK; }
void psa_mapper_destroy(psa_mapper*) {}
const char* psa_last_error(void) { return ""; }
// the stand-in index: class c = {c} for c < 2^20 (so that a compact record's class id expands to one member)
int psa_index_host_classes(const psa_index*, const uint64_t** eq_offsets, const uint32_t** eq_members, uint64_t* n_eq) {
    static uint64_t* off = nullptr;
    static uint32_t* mem = nullptr;
    const uint64_t n = 1u << 20;
    if (!off) {
        off = (uint64_t*)malloc((n + 1) * 8);
        mem = (uint32_t*)malloc(n * 4);
        for (uint64_t i = 0; i <= n; i++) off[i] = i;
        for (uint64_t i = 0; i < n; i++) mem[i] = (uint32_t)i;
    }
    *eq_offsets = off; *eq_members = mem; *n_eq = n;
    return PSA_OK;
}
int psa_mapper_map(psa_mapper*, const psa_read_batch* r, psa_result_batch* o) {
    if (r->format != PSA_READS_ASCII || r->location != PSA_MEM_HOST || !r->read_off || !r->read_len) return PSA_ERR_ARG;
    if (!(o->flags & PSA_RESULT_COMPACT)) return PSA_ERR_ARG;
    psa_hit_compact* hc = (psa_hit_compact*)o->hits;
    uint64_t used = 0;
    // PSA_STUB_NTX=n (throughput runs of scripts/host_process_bench.py only): n ids per read instead of one
    const char* e = getenv("PSA_STUB_NTX");
    const uint32_t ntx = e ? (uint32_t)atoi(e) : 1;
    const bool spread = getenv("PSA_STUB_SPREAD") != nullptr;
    for (uint64_t i = 0; i < r->n_reads; i++) {
        const uint8_t* s = (const uint8_t*)r->data + r->read_off[i];
        if (r->read_off[i] + r->read_len[i] > r->data_len) return PSA_ERR_ARG;
        uint32_t sum = 0;
        for (uint32_t j = 0; j < r->read_len[i]; j++) sum += s[j];
        // one id (= byte sum) travels as the class id of the stand-in index; several ids, or a sum beyond its classes,
        // travel as a non-class set; an empty read has the empty set
        const uint32_t n_tx = r->read_len[i] ? ntx : 0;
        const uint32_t flags = PSA_FLAG_ALIGNED | ((r->read_len[i] && s[0] == 'T') ? PSA_FLAG_MAPPED : 0);
        hc[i].cov_flags = r->read_len[i] | (flags << 28);
        if (n_tx == 1 && !spread && sum < (1u << 20) && (i & 1)) {
            hc[i].eq_or_n = sum;
            continue;
        }
        hc[i].eq_or_n = 0x80000000u | n_tx;
        for (uint32_t j = 0; j < n_tx; j++) {
            if (used >= o->tx_cap) { o->tx_used = used + (r->n_reads - i) * ntx; return PSA_ERR_CAPACITY; }
            // PSA_STUB_SPREAD=1: ids of every decimal length (the formatter's digit paths)
            o->tx_buf[used++] = spread ? (uint32_t)(((uint64_t)(sum + j) * 2654435761u) >> (j % 30)) : sum + 1000 * j;
        }
    }
    o->tx_used = used;
    return PSA_OK;
}
}
