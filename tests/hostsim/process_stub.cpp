// process_stub.cpp -- UNIT-TEST HARNESS, not product code.
//
// Links rust-pseudoaligner_b200/csrc/process_reads.cpp (the C++ mirror of the reference's map driver)
// against a stand-in for the C-ABI entry points it calls (the mapper, and the FASTQ text lanes of psa_fastq.h run
// serially on the CPU with the very routines the kernels are made of, psa_fastq.cuh), so that its FASTQ reader, record table,
// batching, ordering and line writer can be tested without a GPU.  The stand-in "mapper" derives a fake
// result from each read's bytes (coverage = length, one transcript id = byte sum, flag from the first
// base), which lets the test check that every record reaches the mapper intact and in order.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../include/psa.h"
#include "../../rust-pseudoaligner_b200/csrc/psa_fastq.cuh"
#include "../../rust-pseudoaligner_b200/csrc/psa_fastq.h"

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

// ---- the FASTQ text lanes (psa_fastq.h), serially: the same record cutting and line formatting text as the kernels
struct psa_fq_lane {
    uint64_t block_bytes, tail_bytes, len;
    std::vector<uint8_t> text;
    std::vector<uint32_t> nl;
    std::vector<char> out;
};
int psa_fq_lane_create(psa_index*, uint64_t block_bytes, uint64_t tail_bytes, psa_fq_lane** out) {
    if (!block_bytes || block_bytes % 4096) return PSA_ERR_ARG;
    psa_fq_lane* l = new psa_fq_lane();
    l->block_bytes = block_bytes; l->tail_bytes = tail_bytes; l->len = 0;
    l->text.resize(block_bytes + tail_bytes + 64);
    *out = l;
    return PSA_OK;
}
void psa_fq_lane_destroy(psa_fq_lane* l) { delete l; }
uint8_t* psa_fq_lane_text(psa_fq_lane* l) { return l->text.data(); }
int psa_fq_lane_index(psa_fq_lane* l, uint64_t len, uint64_t own_bytes, uint64_t* nl_own, uint64_t* nl_total) {
    if (len > l->text.size() || own_bytes > len || (own_bytes != len && own_bytes % 4096)) return PSA_ERR_ARG;
    l->len = len;
    l->nl.clear();
    *nl_own = 0;
    for (uint64_t i = 0; i < len; i++)
        if (l->text[i] == '\n') {
            l->nl.push_back((uint32_t)i);
            if (i < own_bytes) (*nl_own)++;
        }
    *nl_total = l->nl.size();
    return PSA_OK;
}
int psa_fq_lane_run(psa_fq_lane* l, int64_t j0, uint64_t n, const uint64_t* tick_at, uint32_t n_ticks, psa_fq_result* out) {
    memset(out, 0, sizeof *out);
    out->plain = 1;
    if (!n) return PSA_OK;
    if (j0 < -1 || (uint64_t)(j0 + 4 * (int64_t)(n - 1) + 4) >= l->nl.size() || n_ticks > PSA_FQ_MAX_TICKS) return PSA_ERR_ARG;
    const char* e = getenv("PSA_STUB_NTX");
    const uint32_t ntx = e ? (uint32_t)atoi(e) : 1;
    const bool spread = getenv("PSA_STUB_SPREAD") != nullptr;
    const uint8_t* x = l->text.data();
    std::vector<psa::FqRecord> rec(n);
    for (uint64_t r = 0; r < n; r++)
        if (!psa::fq_cut_record(x, l->nl.data(), j0 + 4 * (int64_t)r, rec[r])) {
            out->plain = 0;
            return PSA_OK;
        }
    l->out.clear();
    std::vector<uint32_t> tx;
    uint32_t t = 0;
    for (uint64_t r = 0; r < n; r++) {
        const uint8_t* s = x + rec[r].seq_off;
        uint32_t sum = 0;
        for (uint32_t j = 0; j < rec[r].seq_len; j++) sum += s[j];
        tx.clear();
        for (uint32_t j = 0; j < (rec[r].seq_len ? ntx : 0); j++)
            tx.push_back(spread ? (uint32_t)(((uint64_t)(sum + j) * 2654435761u) >> (j % 30)) : sum + 1000 * j);
        const bool flag = rec[r].seq_len && s[0] == 'T';
        const uint32_t need = psa::fq_line_len(flag, x + rec[r].id_off, rec[r].id_len, tx.data(), (uint32_t)tx.size(), rec[r].seq_len);
        const size_t at = l->out.size();
        l->out.resize(at + need);
        char* end = psa::fq_format_line(l->out.data() + at, flag, x + rec[r].id_off, rec[r].id_len, tx.data(), (uint32_t)tx.size(), rec[r].seq_len);
        if ((size_t)(end - l->out.data()) != at + need) return PSA_ERR_INTERNAL;   // length pass and write pass must agree
        while (t < n_ticks && tick_at[t] == r) out->tick_mapped[t++] = out->mapped;
        out->mapped += flag;
        out->aligned++;
    }
    while (t < n_ticks) out->tick_mapped[t++] = out->mapped;
    out->out_bytes = l->out.size();
    out->out_text = l->out.data();
    out->end_off = rec[n - 1].end;
    return PSA_OK;
}
}
