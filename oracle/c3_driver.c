/*
 * c3_driver.c -- CPU ORACLE, the reference's map DRIVER restated (test infrastructure; see psa_oracle.h).
 *
 * ref src/pseudoaligner.rs:420-514 process_reads, shaped as the reference shapes it:
 *   - num_threads worker threads (crossbeam scope, :434-474);
 *   - ONE record per mutex acquisition (utils::get_next_record, src/utils.rs:152-157; :431, :442): the FASTQ
 *     iterator is behind a mutex and every worker locks it to parse exactly one record;
 *   - per record: DnaString::from_dna_string (:449-450), map_read (:451), the "mapped" flag (:453-462);
 *   - the tuple goes through a BOUNDED channel of capacity num_threads (sync_channel, :430, :464);
 *   - the main thread receives and println!s one `{:?}` line per read (:480-507), serially.
 * This is BASELINE.md's "C3" arm: what the reference's own driver costs around map_read.  The record reader
 * follows bio::io::fastq (id up to the first space, trimmed lines, wrapped sequences).  Output order is arrival
 * order, as in the reference.
 */
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "psa_oracle.h"

typedef struct {
    int flag;
    char* id;        /* malloc'ed */
    uint32_t* tx;    /* malloc'ed */
    uint32_t n_tx, coverage;
    int done;        /* a worker's end-of-stream marker (the reference sends None, :468) */
} c3_msg;

typedef struct {
    /* the FASTQ "iterator": an in-memory copy of the file walked under the mutex */
    const char* text;
    size_t len, pos;
    pthread_mutex_t reader_mu;
    int reader_error;
    /* bounded channel */
    c3_msg* ring;
    uint32_t cap, head, count;
    pthread_mutex_t ch_mu;
    pthread_cond_t not_full, not_empty;
    const orc_index* ix;
    uint32_t allowed;
} c3_state;

static size_t trim_end(const char* s, size_t n) {
    while (n && (s[n - 1] == ' ' || (s[n - 1] >= 9 && s[n - 1] <= 13))) n--;
    return n;
}
/* one line [*pos, end of line); advances past the '\n'.  Returns 0 at the end of the text. */
static int next_line(c3_state* st, const char** s, size_t* n) {
    if (st->pos >= st->len) return 0;
    const char* p = st->text + st->pos;
    const char* e = memchr(p, '\n', st->len - st->pos);
    size_t l = e ? (size_t)(e - p) : st->len - st->pos;
    *s = p; *n = l;
    st->pos += l + (e ? 1 : 0);
    return 1;
}
/* get_next_record: lock, parse ONE record (bio::io::fastq::Reader::read), unlock.  1 record, 0 end, -1 error. */
static int get_next_record(c3_state* st, char** id, char** seq, size_t* seq_len) {
    int rc = 0;
    pthread_mutex_lock(&st->reader_mu);
    const char* s; size_t n;
    if (!st->reader_error && next_line(st, &s, &n)) {
        if (n == 0 || s[0] != '@') { st->reader_error = 1; rc = -1; }
        else {
            size_t h = trim_end(s + 1, n - 1), ie = 0;
            while (ie < h && s[1 + ie] != ' ') ie++;
            *id = (char*)malloc(ie + 1);
            memcpy(*id, s + 1, ie); (*id)[ie] = 0;
            size_t cap = 256, used = 0, lines = 0, qual = 0;
            char* sq = (char*)malloc(cap);
            int have;
            while ((have = next_line(st, &s, &n)) && !(n && s[0] == '+')) {
                size_t t = trim_end(s, n);
                if (used + t + 1 > cap) { while (used + t + 1 > cap) cap *= 2; sq = (char*)realloc(sq, cap); }
                memcpy(sq + used, s, t); used += t; lines++;
            }
            for (size_t q = 0; q < lines; q++)
                if (next_line(st, &s, &n)) qual += trim_end(s, n);
            if (qual == 0) { st->reader_error = 1; rc = -1; free(*id); free(sq); }
            else { *seq = sq; *seq_len = used; rc = 1; }
        }
    }
    pthread_mutex_unlock(&st->reader_mu);
    return rc;
}
static void ch_send(c3_state* st, c3_msg m) {
    pthread_mutex_lock(&st->ch_mu);
    while (st->count == st->cap) pthread_cond_wait(&st->not_full, &st->ch_mu);
    st->ring[(st->head + st->count) % st->cap] = m;
    st->count++;
    pthread_cond_signal(&st->not_empty);
    pthread_mutex_unlock(&st->ch_mu);
}
static c3_msg ch_recv(c3_state* st) {
    pthread_mutex_lock(&st->ch_mu);
    while (st->count == 0) pthread_cond_wait(&st->not_empty, &st->ch_mu);
    c3_msg m = st->ring[st->head];
    st->head = (st->head + 1) % st->cap;
    st->count--;
    pthread_cond_signal(&st->not_full);
    pthread_mutex_unlock(&st->ch_mu);
    return m;
}
static void* worker(void* arg) {
    c3_state* st = (c3_state*)arg;
    uint64_t* words = NULL; size_t words_cap = 0;
    uint32_t* tmp = NULL; size_t tmp_cap = 0;
    for (;;) {
        char *id = NULL, *seq = NULL; size_t L = 0;
        int r = get_next_record(st, &id, &seq, &L);                         /* :442 */
        if (r <= 0) break;
        size_t nw = orc_words_for(L) + 2;
        if (nw > words_cap) { words_cap = nw * 2; words = (uint64_t*)realloc(words, words_cap * 8); }
        orc_pack_ascii(seq, L, words);                                       /* :449-450 */
        words[orc_words_for(L)] = 0;
        uint32_t n_tx = 0, cov = 0, eq = 0;
        size_t cap = 4096;
        int some;
        for (;;) {
            if (cap > tmp_cap) { tmp_cap = cap; tmp = (uint32_t*)realloc(tmp, tmp_cap * 4); }
            some = orc_map_read_with_mismatch(st->ix, words, (uint32_t)L, st->allowed, tmp, tmp_cap, &n_tx, &cov, &eq, NULL, 0, NULL, NULL);  /* :451 */
            if (some >= 0) break;
            cap *= 16;
        }
        c3_msg m; memset(&m, 0, sizeof m);
        m.id = id;
        if (some > 0) {                                                       /* :453-460 */
            m.flag = cov >= 32 && n_tx == 0;                                  /* :455 (sic) */
            m.n_tx = n_tx; m.coverage = cov;
            m.tx = (uint32_t*)malloc((size_t)(n_tx ? n_tx : 1) * 4);
            memcpy(m.tx, tmp, (size_t)n_tx * 4);
        } else {                                                              /* :461 */
            m.flag = 0; m.n_tx = 0; m.coverage = 0; m.tx = NULL;
        }
        free(seq);
        ch_send(st, m);                                                       /* :464 */
    }
    c3_msg end; memset(&end, 0, sizeof end); end.done = 1;
    ch_send(st, end);                                                         /* :468 */
    free(words); free(tmp);
    return NULL;
}

/* Returns 0 (or -7 for an I/O / malformed-record error); *reads, *mapped receive the counters of :476-477. */
int orc_process_reads_c3(const orc_index* ix, const char* fastq_path, const char* out_path, uint32_t num_threads,
                         uint32_t allowed_mismatches, uint64_t* reads, uint64_t* mapped) {
    FILE* f = fopen(fastq_path, "rb");
    if (!f) return -7;
    fseek(f, 0, SEEK_END);
    long sz = ftell(f);
    rewind(f);
    char* text = (char*)malloc((size_t)sz + 1);
    if (!text || fread(text, 1, (size_t)sz, f) != (size_t)sz) { fclose(f); free(text); return -7; }
    fclose(f);
    FILE* out = fopen(out_path, "wb");
    if (!out) { free(text); return -7; }
    if (!num_threads) num_threads = 1;
    c3_state st; memset(&st, 0, sizeof st);
    st.text = text; st.len = (size_t)sz; st.ix = ix; st.allowed = allowed_mismatches;
    st.cap = num_threads; st.ring = (c3_msg*)calloc(st.cap, sizeof(c3_msg));
    pthread_mutex_init(&st.reader_mu, NULL); pthread_mutex_init(&st.ch_mu, NULL);
    pthread_cond_init(&st.not_full, NULL); pthread_cond_init(&st.not_empty, NULL);
    pthread_t* th = (pthread_t*)malloc(num_threads * sizeof(pthread_t));
    for (uint32_t t = 0; t < num_threads; t++) pthread_create(&th[t], NULL, worker, &st);
    uint64_t n_reads = 0, n_mapped = 0; uint32_t dead = 0;
    while (dead < num_threads) {                                              /* :480 */
        c3_msg m = ch_recv(&st);
        if (m.done) { dead++; continue; }                                     /* :483-487 */
        n_reads++; n_mapped += m.flag;
        fprintf(out, "(%s, \"%s\", [", m.flag ? "true" : "false", m.id);      /* :490 (ids here are plain ASCII) */
        for (uint32_t j = 0; j < m.n_tx; j++) fprintf(out, j ? ", %u" : "%u", m.tx[j]);
        fprintf(out, "], %u)\n", m.coverage);
        free(m.id); free(m.tx);
    }
    for (uint32_t t = 0; t < num_threads; t++) pthread_join(th[t], NULL);
    fclose(out);
    free(th); free(st.ring); free(text);
    if (reads) *reads = n_reads;
    if (mapped) *mapped = n_mapped;
    return st.reader_error ? -7 : 0;
}
