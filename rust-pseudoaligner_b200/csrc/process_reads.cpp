// process_reads.cpp -- psa_process_reads: the C++ mirror of the reference's map driver
// (ref src/pseudoaligner.rs:420-514) on top of the C ABI of include/psa.h.
//
// The reference pulls one FASTQ record per mutex acquisition (src/utils.rs:152-157), maps it on a
// worker thread, sends the tuple through a bounded channel and println!s it on the main thread
// (:480-507).  Here the same work is a three-stage pipeline over batches: a reader thread parses
// FASTQ text into pinned batch buffers, the calling thread runs psa_mapper_map (GPU), formatter
// threads turn psa_hit[] + tx_buf into the reference's `{:?}` lines and the writer emits them in
// INPUT order (a legal instance of the reference's "arrival order").  Nothing here maps reads on
// the CPU.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

#include <chrono>
#include <condition_variable>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/psa.h"

namespace {

struct Pinned {  // growable cudaHostAlloc buffer (contents preserved on growth)
    uint8_t* p = nullptr;
    uint64_t cap = 0;
    int reserve(uint64_t want, uint64_t keep) {
        if (want <= cap) return PSA_OK;
        uint64_t ncap = want + want / 2 + 4096;
        void* q = nullptr;
        int rc = psa_host_alloc(&q, ncap);
        if (rc) return rc;
        if (keep) memcpy(q, p, keep);
        if (p) psa_host_free(p);
        p = (uint8_t*)q;
        cap = ncap;
        return PSA_OK;
    }
    void release() {
        if (p) psa_host_free(p);
        p = nullptr;
        cap = 0;
    }
};

struct Batch {
    Pinned seq, hits, tx;
    uint64_t seq_len = 0;
    std::vector<uint64_t> off;
    std::vector<uint32_t> len;
    std::string ids;  // ids back to back
    std::vector<uint64_t> id_off;
    uint64_t n = 0, tx_used = 0;
    int state = 0;  // 0 free, 1 filled, 2 mapped
    bool last = false;
    void clear() {
        seq_len = 0; off.clear(); len.clear(); ids.clear(); id_off.clear(); n = 0; tx_used = 0; last = false;
    }
};

// buffered line reader over zlib (reads plain and gzip files alike)
struct LineReader {
    gzFile f = nullptr;
    std::vector<char> buf;
    size_t pos = 0, end = 0;
    bool eof = false;
    bool open(const char* path) {
        f = gzopen(path, "rb");
        if (!f) return false;
        gzbuffer(f, 1 << 20);
        buf.resize(1 << 22);
        return true;
    }
    void close() {
        if (f) gzclose(f);
        f = nullptr;
    }
    // returns false at end of file; line excludes the terminator (\n or \r\n)
    bool next(const char*& s, size_t& n) {
        for (;;) {
            char* nl = (char*)memchr(buf.data() + pos, '\n', end - pos);
            if (nl) {
                s = buf.data() + pos;
                n = (size_t)(nl - s);
                pos += n + 1;
                if (n && s[n - 1] == '\r') n--;
                return true;
            }
            if (eof) {
                if (pos == end) return false;
                s = buf.data() + pos;
                n = end - pos;
                pos = end;
                if (n && s[n - 1] == '\r') n--;
                return true;
            }
            if (pos > 0) {  // keep the partial line, refill
                memmove(buf.data(), buf.data() + pos, end - pos);
                end -= pos;
                pos = 0;
            }
            if (end == buf.size()) buf.resize(buf.size() * 2);
            int got = gzread(f, buf.data() + end, (unsigned)std::min<size_t>(buf.size() - end, 1u << 30));
            if (got <= 0) eof = true;
            else end += (size_t)got;
        }
    }
};

// Rust's `{:?}` of a String: quotes, with \" \\ \n \r \t \0 and \u{..} for other control chars
void debug_str(std::string& out, const char* s, size_t n) {
    out.push_back('"');
    for (size_t i = 0; i < n; i++) {
        unsigned char c = (unsigned char)s[i];
        switch (c) {
            case '"': out += "\\\""; break;
            case '\\': out += "\\\\"; break;
            case '\n': out += "\\n"; break;
            case '\r': out += "\\r"; break;
            case '\t': out += "\\t"; break;
            case 0: out += "\\0"; break;
            default:
                if (c < 0x20 || c == 0x7f) {
                    char t[16];
                    snprintf(t, sizeof t, "\\u{%x}", c);
                    out += t;
                } else {
                    out.push_back((char)c);
                }
        }
    }
    out.push_back('"');
}

void put_u64(std::string& out, uint64_t v) {
    char t[24];
    int i = 24;
    do { t[--i] = (char)('0' + v % 10); v /= 10; } while (v);
    out.append(t + i, 24 - i);
}

// `(flag, "id", [tx, ...], coverage)` -- the tuple printed at ref src/pseudoaligner.rs:490
void format_range(const Batch& b, uint64_t r0, uint64_t r1, std::string& out, uint64_t& mapped) {
    const psa_hit* hits = (const psa_hit*)b.hits.p;
    const uint32_t* tx = (const uint32_t*)b.tx.p;
    out.clear();
    out.reserve((r1 - r0) * 64);
    for (uint64_t i = r0; i < r1; i++) {
        const psa_hit& h = hits[i];
        const bool flag = (h.flags & PSA_FLAG_MAPPED) != 0;
        mapped += flag;
        out += flag ? "(true, " : "(false, ";
        debug_str(out, b.ids.data() + b.id_off[i], (size_t)(b.id_off[i + 1] - b.id_off[i]));
        out += ", [";
        for (uint32_t j = 0; j < h.n_tx; j++) {
            if (j) out += ", ";
            put_u64(out, tx[h.tx_off + j]);
        }
        out += "], ";
        put_u64(out, h.coverage);
        out += ")\n";
    }
}

}  // namespace

extern "C" int psa_process_reads(psa_index* index, const char* fastq_path, const char* out_path, uint32_t num_threads,
                                 uint64_t batch_reads, int progress, psa_process_stats* stats) {
    if (!index || !fastq_path) return PSA_ERR_ARG;
    if (!num_threads) num_threads = 1;
    if (!batch_reads) batch_reads = 1ull << 20;
    const auto t0 = std::chrono::steady_clock::now();

    LineReader in;
    if (!in.open(fastq_path)) return PSA_ERR_IO;
    FILE* out = stdout;
    const bool own_out = out_path && strcmp(out_path, "-") != 0;
    if (own_out) {
        out = fopen(out_path, "wb");
        if (!out) {
            in.close();
            return PSA_ERR_IO;
        }
    }
    psa_mapper* mapper = nullptr;
    int rc = psa_mapper_create(index, 0, &mapper);
    if (rc) {
        in.close();
        if (own_out) fclose(out);
        return rc;
    }

    constexpr int kSlots = 3;
    Batch slot[kSlots];
    std::mutex mu;
    std::condition_variable cv;
    int reader_rc = PSA_OK;
    bool abort_all = false;

    // stage 1: FASTQ text -> batch buffers (four-line records; id = header up to the first blank,
    // as bio::io::fastq::Record::id)
    std::thread reader([&]() {
        int s = 0;
        bool done = false;
        while (!done) {
            Batch& b = slot[s];
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&]() { return b.state == 0 || abort_all; });
                if (abort_all) return;
            }
            b.clear();
            b.id_off.push_back(0);
            int err = PSA_OK;
            while (b.n < batch_reads) {
                const char* l;
                size_t n;
                if (!in.next(l, n)) { done = true; break; }
                if (n == 0) continue;  // blank line between records
                if (l[0] != '@') { err = PSA_ERR_IO; break; }
                size_t e = 1;
                while (e < n && l[e] != ' ' && l[e] != '\t') e++;
                b.ids.append(l + 1, e - 1);
                b.id_off.push_back(b.ids.size());
                if (!in.next(l, n)) { err = PSA_ERR_IO; break; }
                if ((err = b.seq.reserve(b.seq_len + n + 64, b.seq_len))) break;
                memcpy(b.seq.p + b.seq_len, l, n);
                b.off.push_back(b.seq_len);
                b.len.push_back((uint32_t)n);
                b.seq_len += n;
                const char* q;
                size_t qn;
                if (!in.next(q, qn) || qn == 0 || q[0] != '+') { err = PSA_ERR_IO; break; }
                if (!in.next(q, qn)) { err = PSA_ERR_IO; break; }
                b.n++;
            }
            std::unique_lock<std::mutex> lk(mu);
            if (err) {
                reader_rc = err;  // the complete records before the bad one are still processed
                done = true;
            }
            b.last = done;
            b.state = 1;
            cv.notify_all();
            s = (s + 1) % kSlots;
        }
    });

    // stage 3: format + write, in input order
    uint64_t n_reads = 0, n_mapped = 0, n_aligned = 0, next_tick = 1000000;
    int writer_rc = PSA_OK;
    std::thread writer([&]() {
        int s = 0;
        std::vector<std::string> parts(num_threads);
        std::vector<uint64_t> mapped(num_threads);
        for (;;) {
            Batch& b = slot[s];
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&]() { return b.state == 2 || abort_all; });
                if (abort_all) return;
            }
            if (b.n) {
                std::vector<std::thread> th;
                for (uint32_t t = 0; t < num_threads; t++) {
                    mapped[t] = 0;
                    uint64_t r0 = b.n * t / num_threads, r1 = b.n * (t + 1) / num_threads;
                    th.emplace_back([&, t, r0, r1]() { format_range(b, r0, r1, parts[t], mapped[t]); });
                }
                for (auto& x : th) x.join();
                for (uint32_t t = 0; t < num_threads; t++) {
                    if (fwrite(parts[t].data(), 1, parts[t].size(), out) != parts[t].size()) writer_rc = PSA_ERR_IO;
                    n_mapped += mapped[t];
                }
                const psa_hit* hits = (const psa_hit*)b.hits.p;
                for (uint64_t i = 0; i < b.n; i++) n_aligned += hits[i].flags & PSA_FLAG_ALIGNED;
                n_reads += b.n;
                if (progress && n_reads >= next_tick) {  // ref :497-504
                    fprintf(stderr, "\rDone Mapping %llu reads w/ Rate: %g", (unsigned long long)n_reads,
                            (double)(float)((float)n_mapped * 100.0f / (float)n_reads));
                    fflush(stderr);
                    next_tick = (n_reads / 1000000 + 1) * 1000000;
                }
            }
            const bool last = b.last;
            {
                std::unique_lock<std::mutex> lk(mu);
                b.state = 0;
                cv.notify_all();
            }
            if (last) return;
            s = (s + 1) % kSlots;
        }
    });

    // stage 2 (this thread): the GPU
    int map_rc = PSA_OK;
    for (int s = 0;; s = (s + 1) % kSlots) {
        Batch& b = slot[s];
        {
            std::unique_lock<std::mutex> lk(mu);
            cv.wait(lk, [&]() { return b.state == 1; });
        }
        if (b.n) {
            map_rc = b.hits.reserve(b.n * sizeof(psa_hit), 0);
            if (!map_rc) map_rc = b.tx.reserve(std::max<uint64_t>(b.n * 16 * 4, 4096), 0);
            psa_read_batch r{};
            r.format = PSA_READS_ASCII;
            r.location = PSA_MEM_HOST;
            r.data = b.seq.p;
            r.data_len = b.seq_len;
            r.read_off = b.off.data();
            r.read_len = b.len.data();
            r.n_reads = b.n;
            psa_result_batch o{};
            o.location = PSA_MEM_HOST;
            for (int attempt = 0; attempt < 2 && !map_rc; attempt++) {
                o.hits = (psa_hit*)b.hits.p;
                o.tx_buf = (uint32_t*)b.tx.p;
                o.tx_cap = b.tx.cap / 4;
                map_rc = psa_mapper_map(mapper, &r, &o);
                if (map_rc == PSA_ERR_CAPACITY && o.tx_used > o.tx_cap && attempt == 0)
                    map_rc = b.tx.reserve(o.tx_used * 4 + 4096, 0);  // resubmit with the size asked for
                else
                    break;
            }
            b.tx_used = o.tx_used;
        }
        const bool last = b.last;
        {
            std::unique_lock<std::mutex> lk(mu);
            if (map_rc) {
                abort_all = true;
            } else {
                b.state = 2;
            }
            cv.notify_all();
        }
        if (map_rc || last) break;
    }
    reader.join();
    writer.join();
    if (progress) fprintf(stderr, "\n");
    fflush(out);
    if (own_out) fclose(out);
    in.close();
    psa_mapper_destroy(mapper);
    for (auto& b : slot) { b.seq.release(); b.hits.release(); b.tx.release(); }
    if (stats) {
        stats->reads = n_reads;
        stats->mapped = n_mapped;
        stats->aligned = n_aligned;
        stats->seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    }
    if (map_rc) return map_rc;
    if (reader_rc) return reader_rc;  // the reference panics on a malformed record (:446)
    return writer_rc;
}
