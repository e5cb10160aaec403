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
int psa_mapper_map(psa_mapper*, const psa_read_batch* r, psa_result_batch* o) {
    if (r->format != PSA_READS_ASCII || r->location != PSA_MEM_HOST || !r->read_off || !r->read_len) return PSA_ERR_ARG;
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
        psa_hit& h = o->hits[i];
        h.coverage = r->read_len[i];
        h.n_tx = r->read_len[i] ? ntx : 0;
        h.tx_off = used;
        h.eq_id = 0;
        h.flags = PSA_FLAG_ALIGNED | ((r->read_len[i] && s[0] == 'T') ? PSA_FLAG_MAPPED : 0);
        for (uint32_t j = 0; j < h.n_tx; j++) {
            if (used >= o->tx_cap) return PSA_ERR_CAPACITY;
            // PSA_STUB_SPREAD=1: ids of every decimal length (the formatter's digit paths)
            o->tx_buf[used++] = spread ? (uint32_t)(((uint64_t)(sum + j) * 2654435761u) >> (j % 30)) : sum + 1000 * j;
        }
    }
    o->tx_used = used;
    return PSA_OK;
}
}
