// process_reads.cpp -- psa_process_reads: the C++ mirror of the reference's map driver
// (ref src/pseudoaligner.rs:420-514) on top of the C ABI of include/psa.h.
//
// The reference pulls one FASTQ record per mutex acquisition (src/utils.rs:152-157), maps it on a
// worker thread, sends the tuple through a bounded channel and println!s it on the main thread
// (:480-507).  Here the same work is a three-stage pipeline over batches: a reader thread fills a
// pinned block with raw FASTQ text and indexes its lines (newline scan split over num_threads
// threads); the records are NOT copied -- the sequences are handed to psa_mapper_map as offsets
// into the text block (PSA_READS_ASCII with read_off), ids are formatted straight from it; the
// calling thread runs psa_mapper_map (GPU); formatter threads turn psa_hit[] + tx_buf into the
// reference's `{:?}` lines and the writer emits them in INPUT order (a legal instance of the
// reference's "arrival order").  Nothing here maps reads on the CPU.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <sys/types.h>
#include <unistd.h>
#include <zlib.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/psa.h"
#include "psa_fastq.cuh"
#include "psa_fastq.h"

namespace {

struct Pinned {  // growable cudaHostAlloc buffer (contents preserved on growth)
    uint8_t* p = nullptr;
    uint64_t cap = 0;
    int reserve(uint64_t want, uint64_t keep) {
        if (want <= cap) return PSA_OK;
        // the first allocation is what was asked for (pinning costs ~0.4 s per GB); regrowth leaves headroom
        uint64_t ncap = p ? want + want / 2 + 4096 : want + 4096;
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
    Pinned text, hits, tx;       // text: raw FASTQ bytes of this batch (records start at byte 0); hits: psa_hit_compact[]
    uint64_t text_len = 0;       // bytes of complete records
    std::vector<uint64_t> off;   // sequence start (byte offset into text) per read
    std::vector<uint32_t> len;   // sequence length per read
    std::vector<uint64_t> id_off;  // id start per read
    std::vector<uint32_t> id_len;
    uint64_t n = 0, tx_used = 0;
    int state = 0;  // 0 free, 1 filled, 2 mapped
    bool last = false;
};

// raw byte source: regular files through pread(2) (any number of threads at once), everything else
// (pipes, gzip) through a serial stream
struct ByteSource {
    gzFile gz = nullptr;    // gzip files, and every stream that is not a regular file (zlib passes plain text through)
    int fd = -1;            // regular plain file: positional reads
    uint64_t size = 0, pos = 0;
    bool open(const char* path) {
        const int f = ::open(path, O_RDONLY | O_CLOEXEC);
        if (f < 0) return false;
        struct stat st;
        bool regular = fstat(f, &st) == 0 && S_ISREG(st.st_mode);
        unsigned char magic[2] = {0, 0};
        const bool gzip = regular && pread(f, magic, 2, 0) == 2 && magic[0] == 0x1f && magic[1] == 0x8b;
        if (regular && !gzip) {
            fd = f;
            size = (uint64_t)st.st_size;
            return true;
        }
        gz = gzdopen(f, "rb");  // owns f from here on
        if (!gz) {
            ::close(f);
            return false;
        }
        gzbuffer(gz, 1 << 20);
        return true;
    }
    bool parallel() const { return fd >= 0; }
    // positional read of [at, at + n) of a regular file; returns bytes read
    size_t read_at(uint8_t* dst, size_t n, uint64_t at) const {
        size_t total = 0;
        while (total < n) {
            ssize_t g = pread(fd, dst + total, n - total, (off_t)(at + total));
            if (g <= 0) break;
            total += (size_t)g;
        }
        return total;
    }
    // serial stream: fills up to n bytes; returns bytes read (0 at end of file)
    size_t read(uint8_t* dst, size_t n) {
        size_t total = 0;
        while (total < n && gz) {
            int g = gzread(gz, dst + total, (unsigned)std::min<size_t>(n - total, 1u << 30));
            if (g <= 0) break;
            total += (size_t)g;
        }
        return total;
    }
    void close() {
        if (gz) gzclose(gz);
        if (fd >= 0) ::close(fd);
        gz = nullptr;
        fd = -1;
    }
};

// Uninitialised growable array (std::vector would zero-fill tens of megabytes per block on one thread)
struct U64Buf {
    uint64_t* p = nullptr;
    uint64_t cap = 0, n = 0;
    bool resize(uint64_t want) {
        if (want > cap) {
            uint64_t ncap = want + want / 4 + 1024;
            void* q = realloc(p, ncap * 8);
            if (!q) return false;
            p = (uint64_t*)q;
            cap = ncap;
        }
        n = want;
        return true;
    }
    uint64_t size() const { return n; }
    uint64_t operator[](uint64_t i) const { return p[i]; }
    ~U64Buf() { free(p); }
};

// Appends up to `room` bytes of the source to text[have, ...) and finds every '\n' of text[0, have + got):
// `threads` threads each read one slice of the new bytes (regular files: pread, all at once; streams: one
// serial read first) and scan it while it is still in their cache; thread 0 also scans the bytes that were
// there before.  Returns the bytes appended; nl = the newline positions, ascending.
uint64_t fill_and_index(ByteSource& in, uint8_t* text, uint64_t have, uint64_t room, uint32_t threads, bool& eof,
                        std::vector<std::vector<uint64_t>>& part, U64Buf& nl, bool& oom) {
    uint64_t got = 0;
    const bool par = in.parallel();
    if (!eof) {
        if (par) {
            got = std::min<uint64_t>(room, in.size > in.pos ? in.size - in.pos : 0);
        } else {
            got = in.read(text + have, room);
            if (got < room) eof = true;
        }
    }
    part.resize(threads + 1);
    std::vector<uint64_t> short_by(threads, 0);
    std::vector<std::thread> th;
    auto scan = [&](std::vector<uint64_t>& v, uint64_t lo, uint64_t hi) {
        v.reserve(v.size() + (hi - lo) / 64 + 16);
        const uint8_t* p = text + lo;
        const uint8_t* end = text + hi;
        while (p < end) {
            const uint8_t* q = (const uint8_t*)memchr(p, '\n', (size_t)(end - p));
            if (!q) break;
            v.push_back((uint64_t)(q - text));
            p = q + 1;
        }
    };
    for (uint32_t t = 0; t < threads; t++) {
        th.emplace_back([&, t]() {
            part[t + 1].clear();
            if (t == 0) {
                part[0].clear();
                scan(part[0], 0, have);
            }
            const uint64_t lo = have + got * t / threads, hi = have + got * (t + 1) / threads;
            if (par && hi > lo) {
                const size_t g = in.read_at(text + lo, hi - lo, in.pos + (lo - have));
                short_by[t] = (hi - lo) - g;
            }
            scan(part[t + 1], lo, hi);
        });
    }
    for (auto& x : th) x.join();
    if (par) {
        for (uint32_t t = 0; t < threads; t++)
            if (short_by[t]) {  // the file shrank under us: keep the bytes before the first hole
                uint64_t keep = got * t / threads + ((got * (t + 1) / threads - got * t / threads) - short_by[t]);
                for (uint32_t u = t; u < threads; u++) {
                    auto& v = part[u + 1];
                    while (!v.empty() && v.back() >= have + keep) v.pop_back();
                }
                got = keep;
                in.size = in.pos + got;
                break;
            }
        in.pos += got;
        if (in.pos >= in.size) eof = true;
    }
    // ascending positions in one array, copied by the threads that found them
    std::vector<uint64_t> at(threads + 2, 0);
    for (uint32_t t = 0; t <= threads; t++) at[t + 1] = at[t] + part[t].size();
    if (!nl.resize(at[threads + 1] + 1)) {  // (+1: the caller may append the newline of an unterminated last line)
        oom = true;
        return got;
    }
    nl.n = at[threads + 1];
    th.clear();
    for (uint32_t t = 0; t < threads; t++)
        th.emplace_back([&, t]() {
            if (t == 0 && !part[0].empty()) memcpy(nl.p, part[0].data(), part[0].size() * 8);
            if (!part[t + 1].empty()) memcpy(nl.p + at[t + 1], part[t + 1].data(), part[t + 1].size() * 8);
        });
    for (auto& x : th) x.join();
    return got;
}

// Output text of one formatter thread: a raw growable buffer (no per-append capacity checks on the hot path:
// the caller reserves the worst case of a record before formatting it)
struct OutBuf {
    char* p = nullptr;
    size_t n = 0, cap = 0;
    bool ok = true;
    void reserve_more(size_t extra) {
        if (n + extra <= cap) return;
        size_t ncap = std::max(cap + cap / 2, n + extra + (1u << 16));
        void* q = realloc(p, ncap);
        if (!q) {
            ok = false;  // the caller stops formatting and reports the failure
        } else {
            p = (char*)q;
            cap = ncap;
        }
    }
    ~OutBuf() { free(p); }
};

// ---- text helpers shared with the Python mirror (psa_debug_str / psa_fastq_trim_end below) -----------------------
// Rust's str::trim_end removes White_Space code points: the ASCII ones and the multi-byte ones below.
// Returns the length of s[0, n) without its trailing white space.
size_t trim_end(const uint8_t* s, size_t n) {
    for (;;) {
        if (n >= 1 && (s[n - 1] == ' ' || (s[n - 1] >= 0x09 && s[n - 1] <= 0x0D))) { n--; continue; }
        if (n >= 2 && s[n - 2] == 0xC2 && (s[n - 1] == 0x85 || s[n - 1] == 0xA0)) { n -= 2; continue; }
        if (n >= 3) {
            const uint8_t a = s[n - 3], b = s[n - 2], c = s[n - 1];
            const bool ws = (a == 0xE1 && b == 0x9A && c == 0x80) || (a == 0xE2 && b == 0x80 && (c <= 0x8A && c >= 0x80)) ||
                            (a == 0xE2 && b == 0x80 && (c == 0xA8 || c == 0xA9 || c == 0xAF)) || (a == 0xE2 && b == 0x81 && c == 0x9F) ||
                            (a == 0xE3 && b == 0x80 && c == 0x80);
            if (ws) { n -= 3; continue; }
        }
        return n;
    }
}
// decodes one UTF-8 scalar at s[i] (i < n); returns its length, 0 if the bytes are not valid UTF-8
int utf8_decode(const uint8_t* s, size_t i, size_t n, uint32_t& cp) {
    const uint8_t c = s[i];
    if (c < 0x80) { cp = c; return 1; }
    int len = c >= 0xF0 ? 4 : c >= 0xE0 ? 3 : c >= 0xC2 ? 2 : 0;
    if (!len || c > 0xF4 || i + len > n) return 0;
    cp = c & (0x3F >> (len - 1));
    for (int j = 1; j < len; j++) {
        if ((s[i + j] & 0xC0) != 0x80) return 0;
        cp = (cp << 6) | (s[i + j] & 0x3F);
    }
    if ((len == 3 && cp < 0x800) || (len == 4 && (cp < 0x10000 || cp > 0x10FFFF)) || (cp >= 0xD800 && cp <= 0xDFFF)) return 0;
    return len;
}
bool utf8_valid(const uint8_t* s, size_t n) {
    for (size_t i = 0; i < n;) {
        uint32_t cp;
        const int l = utf8_decode(s, i, n, cp);
        if (!l) return false;
        i += l;
    }
    return true;
}
// char::escape_debug escapes what is not printable or extends a grapheme.  ASCII is exact; beyond it the code
// points below are escaped as Rust does (C1 controls, soft hyphen, combining marks of the common blocks, format
// and separator characters, private use, non-characters); everything else is printed as is.  (Rust's full
// is_printable / Grapheme_Extend tables are not reproduced: an id with an unassigned or rare combining code point
// would print it raw here and escaped there.)
bool escape_unicode(uint32_t cp) {
    if (cp < 0xA0) return cp >= 0x7F;
    if (cp == 0xAD) return true;
    if ((cp >= 0x300 && cp <= 0x36F) || (cp >= 0x483 && cp <= 0x489) || (cp >= 0x591 && cp <= 0x5BD) || (cp >= 0x610 && cp <= 0x61A) ||
        (cp >= 0x64B && cp <= 0x65F) || cp == 0x61C || (cp >= 0x1AB0 && cp <= 0x1AFF) || (cp >= 0x1DC0 && cp <= 0x1DFF) ||
        (cp >= 0x20D0 && cp <= 0x20FF) || (cp >= 0xFE00 && cp <= 0xFE0F) || (cp >= 0xFE20 && cp <= 0xFE2F))
        return true;
    if ((cp >= 0x200B && cp <= 0x200F) || (cp >= 0x2028 && cp <= 0x202E) || (cp >= 0x2060 && cp <= 0x206F) || cp == 0x180E ||
        cp == 0xFEFF || (cp >= 0xFFF0 && cp <= 0xFFFB))
        return true;
    if ((cp >= 0xE000 && cp <= 0xF8FF) || (cp >= 0xF0000) || (cp & 0xFFFE) == 0xFFFE || (cp >= 0xFDD0 && cp <= 0xFDEF)) return true;
    if (cp >= 0xE0000 && cp <= 0xE0FFF) return true;
    return false;
}
// Rust's `{:?}` of a String: quotes, \" \\ \n \r \t \0, \u{..} for other non-printables
// (at most 10 output bytes per input byte, + 2 quotes).  s must be valid UTF-8.
char* debug_str(char* o, const char* s, size_t n) {
    const uint8_t* u = (const uint8_t*)s;
    *o++ = '"';
    for (size_t i = 0; i < n;) {
        const unsigned char c = u[i];
        if (c >= 0x20 && c < 0x7f && c != '"' && c != '\\') {
            *o++ = (char)c;
            i++;
            continue;
        }
        if (c < 0x80) {
            switch (c) {
                case '"': *o++ = '\\'; *o++ = '"'; break;
                case '\\': *o++ = '\\'; *o++ = '\\'; break;
                case '\n': *o++ = '\\'; *o++ = 'n'; break;
                case '\r': *o++ = '\\'; *o++ = 'r'; break;
                case '\t': *o++ = '\\'; *o++ = 't'; break;
                case 0: *o++ = '\\'; *o++ = '0'; break;
                default: o += sprintf(o, "\\u{%x}", c);
            }
            i++;
            continue;
        }
        uint32_t cp = 0xFFFD;
        int l = utf8_decode(u, i, n, cp);
        if (!l) l = 1;  // (callers validate first)
        if (escape_unicode(cp)) o += sprintf(o, "\\u{%x}", cp);
        else { memcpy(o, u + i, (size_t)l); o += l; }
        i += (size_t)l;
    }
    *o++ = '"';
    return o;
}
// Rust's `{}` of an f32: the shortest decimal that reads back as the same f32, never in exponent form
int display_f32(char* o, float v) {
    for (int prec = 0; prec <= 12; prec++) {
        char buf[64];
        snprintf(buf, sizeof buf, "%.*f", prec, (double)v);
        if (strtof(buf, nullptr) == v) return sprintf(o, "%s", buf);
    }
    return sprintf(o, "%.12f", (double)v);
}

const char kDigits2[201] =
    "00010203040506070809101112131415161718192021222324252627282930313233343536373839404142434445464748495051525354555657585960616263646566676869707172737475767778798081828384858687888990919293949596979899";
// decimal of a 32-bit value: digit count first, then two digits per division written in place
inline char* put_u32(char* o, uint32_t v) {
    const int len = v < 10 ? 1 : v < 100 ? 2 : v < 1000 ? 3 : v < 10000 ? 4 : v < 100000 ? 5 : v < 1000000 ? 6
                    : v < 10000000 ? 7 : v < 100000000 ? 8 : v < 1000000000 ? 9 : 10;
    char* e = o + len;
    while (v >= 100) {
        const uint32_t q = v / 100, r = v - q * 100;
        e -= 2;
        memcpy(e, kDigits2 + 2 * r, 2);
        v = q;
    }
    if (v >= 10) memcpy(o, kDigits2 + 2 * v, 2);
    else *o = (char)('0' + v);
    return o + len;
}

struct HostClasses {  // eq_classes on the host: compact results carry class ids, the members are expanded here
    const uint64_t* off = nullptr;
    const uint32_t* mem = nullptr;
    uint64_t n_eq = 0;
};

// `(flag, "id", [tx, ...], coverage)` -- the tuple printed at ref src/pseudoaligner.rs:490.
// novel_at: offset in b.tx of the members of the first non-class set of the range.
void format_range(const Batch& b, const HostClasses& hc, uint64_t r0, uint64_t r1, uint64_t novel_at, OutBuf& out,
                  uint64_t& mapped, uint64_t& aligned) {
    const psa_hit_compact* hits = (const psa_hit_compact*)b.hits.p;
    const uint32_t* tx = (const uint32_t*)b.tx.p;
    out.n = 0;
    out.ok = true;
    for (uint64_t i = r0; i < r1; i++) {
        const psa_hit_compact h = hits[i];
        const uint32_t flags = h.cov_flags >> 28, coverage = h.cov_flags & ((1u << 28) - 1);
        const uint32_t* m = nullptr;
        uint32_t n_tx = 0;
        if (!(flags & PSA_FLAG_ALIGNED)) {
        } else if (h.eq_or_n & 0x80000000u) {
            n_tx = h.eq_or_n & 0x7FFFFFFFu;
            m = tx + novel_at;
            novel_at += n_tx;
        } else {
            m = hc.mem + hc.off[h.eq_or_n];
            n_tx = (uint32_t)(hc.off[h.eq_or_n + 1] - hc.off[h.eq_or_n]);
        }
        // the id (300+ bytes apart in the FASTQ text), the class's offsets and its members are three dependent cache
        // misses per read: fetch them 16 / 16 / 8 reads ahead
        if (i + 16 < r1) {
            __builtin_prefetch(b.text.p + b.id_off[i + 16]);
            const psa_hit_compact& a = hits[i + 16];
            if (((a.cov_flags >> 28) & PSA_FLAG_ALIGNED) && !(a.eq_or_n & 0x80000000u)) __builtin_prefetch(hc.off + a.eq_or_n);
            const psa_hit_compact& c = hits[i + 8];
            if (((c.cov_flags >> 28) & PSA_FLAG_ALIGNED) && !(c.eq_or_n & 0x80000000u)) __builtin_prefetch(hc.mem + hc.off[c.eq_or_n]);
        }
        // worst case of this record: flag 8, id 10 per byte + 2, 12 per member ("4294967295, "), brackets/coverage/newline 32
        out.reserve_more(8 + 10 * (size_t)b.id_len[i] + 2 + 12 * (size_t)n_tx + 32);
        if (!out.ok) return;
        char* o = out.p + out.n;
        const bool flag = (flags & PSA_FLAG_MAPPED) != 0;
        mapped += flag;
        aligned += flags & PSA_FLAG_ALIGNED;
        if (flag) { memcpy(o, "(true, ", 7); o += 7; }
        else { memcpy(o, "(false, ", 8); o += 8; }
        o = debug_str(o, (const char*)b.text.p + b.id_off[i], b.id_len[i]);
        *o++ = ','; *o++ = ' '; *o++ = '[';
        for (uint32_t j = 0; j < n_tx; j++) {
            if (j) { *o++ = ','; *o++ = ' '; }
            o = put_u32(o, m[j]);
        }
        *o++ = ']'; *o++ = ','; *o++ = ' ';
        o = put_u32(o, coverage);
        *o++ = ')'; *o++ = '\n';
        out.n = (size_t)(o - out.p);
    }
}

// Ordered output: a regular file is written by all formatter threads at once (each at its own offset,
// pwrite); pipes and append-mode descriptors take the parts one after the other.
struct OutFile {
    FILE* f = nullptr;
    int fd = -1;          // >= 0: positional writes
    uint64_t pos = 0;
    // Regular files that can be mapped are written through ONE shared mapping of the region the output is expected to
    // fill: write(2) holds the inode's lock for the whole call, so the pwrites of many threads to one file run one after
    // the other (tmpfs: ~2 GB/s in total), while stores into a mapping fault their pages in independently.  The file is
    // extended to the end of the window first and cut to its real size at the end; output beyond the window is pwritten.
    bool can_map = false;
    uint8_t* map_base = nullptr;
    uint64_t map_off = 0, map_len = 0;
    uint64_t grown_to = 0;   // length the file was extended to (0: not extended)
    void attach(FILE* file) {
        f = file;
        fflush(f);
        const int d = fileno(f);
        struct stat st;
        const int fl = fcntl(d, F_GETFL);
        if (d >= 0 && fstat(d, &st) == 0 && S_ISREG(st.st_mode) && fl != -1 && !(fl & O_APPEND)) {
            const off_t at = lseek(d, 0, SEEK_CUR);
            if (at >= 0) {
                fd = d;
                pos = (uint64_t)at;
                const char* e = getenv("PSA_OUT_MMAP");
                can_map = (fl & O_ACCMODE) == O_RDWR && !(e && atoi(e) == 0);
            }
        }
    }
    // maps [pos, pos + expect) (page-rounded); harmless when it fails: put() falls back to pwrite
    void open_window(uint64_t expect) {
        if (!can_map || map_base) return;
        struct stat st;
        if (fstat(fd, &st) != 0) return;
        map_off = pos & ~4095ull;
        const uint64_t len = ((pos - map_off) + std::max<uint64_t>(expect, 1ull << 24) + 4095) & ~4095ull;
        if ((uint64_t)st.st_size < map_off + len) {
            if (ftruncate(fd, (off_t)(map_off + len)) != 0) return;
            grown_to = map_off + len;
        }
        void* m = mmap(nullptr, len, PROT_READ | PROT_WRITE, MAP_SHARED, fd, (off_t)map_off);
        if (m == MAP_FAILED) {
            can_map = false;
            return;
        }
        map_base = (uint8_t*)m;
        map_len = len;
    }
    static bool write_at(int fd, const char* p, size_t n, uint64_t at) {
        while (n) {
            ssize_t g = pwrite(fd, p, n, (off_t)at);
            if (g <= 0) return false;
            p += g; n -= (size_t)g; at += (uint64_t)g;
        }
        return true;
    }
    // [at, at + n) of the file = p[0, n)
    bool put(const char* p, size_t n, uint64_t at) const {
        if (!n) return true;
        if (map_base && at >= map_off && at + n <= map_off + map_len) {
            uint8_t* dst = map_base + (at - map_off);
#ifdef MADV_POPULATE_WRITE
            {   // fault the pages in with one call instead of one trap per page (Linux 5.14+; ignored when unsupported)
                const uintptr_t a0 = (uintptr_t)dst & ~(uintptr_t)4095, a1 = ((uintptr_t)dst + n + 4095) & ~(uintptr_t)4095;
                (void)madvise((void*)a0, a1 - a0, MADV_POPULATE_WRITE);
            }
#endif
            memcpy(dst, p, n);
            return true;
        }
        return write_at(fd, p, n, at);
    }
    bool finish() {
        bool ok = true;
        if (map_base) {
            munmap(map_base, map_len);
            map_base = nullptr;
        }
        if (fd >= 0) {
            if (grown_to > pos) ok = ftruncate(fd, (off_t)pos) == 0;
            grown_to = 0;
            lseek(fd, (off_t)pos, SEEK_SET);  // stdio continues after what was written
        }
        return ok;
    }
};

// ---- the record table of one text block, exactly as bio::io::fastq::Reader::read (bio 1.5) cuts records ----------
// (the reference reads through it: ref src/pseudoaligner.rs:421,431,442-447, src/utils.rs:152-157)
//   header line: must start with '@' (else Error::MissingAt); id = header[1..].trim_end() up to the first ' '
//   sequence: every following line up to one that starts with '+', each trim_end()-ed and concatenated
//   quality: as many lines as the sequence had; an empty quality is Error::IncompleteRecord
// Lines end at '\n' only.  A record whose lines are not all in the block yet is left for the next block (`need_more`).
struct LineIndex {
    const uint8_t* x;
    const U64Buf& nl;
    uint64_t n_lines;
    uint64_t start(uint64_t i) const { return i ? nl[i - 1] + 1 : 0; }
    uint64_t end(uint64_t i) const { return nl[i]; }  // position of the '\n'
};
enum { REC_OK = 0, REC_NEED_MORE = 1, REC_BAD = 2 };
// Parses the record that starts at line `li`; wrapped sequences are moved together in place.  next = first line after it.
int parse_record(uint8_t* x, const LineIndex& L, uint64_t li, bool eof, uint64_t& next, uint64_t& id_off, uint32_t& id_len,
                 uint64_t& seq_off, uint32_t& seq_len) {
    if (li >= L.n_lines) return REC_NEED_MORE;
    const uint64_t h0 = L.start(li), h1 = L.end(li);
    if (x[h0] != '@') return REC_BAD;                                   // Error::MissingAt (a blank line too)
    const uint64_t he = h0 + 1 + trim_end(x + h0 + 1, h1 - h0 - 1);
    uint64_t ie = h0 + 1;
    while (ie < he && x[ie] != ' ') ie++;
    if (!utf8_valid(x + h0 + 1, h1 - h0 - 1)) return REC_BAD;           // read_line fails on invalid UTF-8
    id_off = h0 + 1;
    id_len = (uint32_t)(ie - h0 - 1);
    uint64_t j = li + 1;
    while (j < L.n_lines && x[L.start(j)] != '+') j++;                  // (an empty line holds its '\n' at start(j): not '+')
    if (j >= L.n_lines) return eof ? REC_BAD : REC_NEED_MORE;           // at the end of the file: no quality -> IncompleteRecord
    const uint64_t k = j - (li + 1);
    if (j + k >= L.n_lines) {
        if (!eof) return REC_NEED_MORE;
        // the file ends inside the quality lines: bio reads empty strings for the missing ones
    }
    uint64_t qual = 0;
    for (uint64_t q = j + 1; q <= j + k && q < L.n_lines; q++) qual += trim_end(x + L.start(q), L.end(q) - L.start(q));
    if (qual == 0) return REC_BAD;                                      // Error::IncompleteRecord (k == 0 included)
    seq_off = L.start(li + 1);
    uint64_t w = seq_off;
    for (uint64_t q = li + 1; q < j; q++) {
        const uint64_t s0 = L.start(q), n = trim_end(x + s0, L.end(q) - s0);
        if (w != s0) memmove(x + w, x + s0, n);
        w += n;
    }
    seq_len = (uint32_t)(w - seq_off);
    next = j + k + 1 < L.n_lines ? j + k + 1 : L.n_lines;
    return REC_OK;
}


// ---- the block pipeline with the text work on the device (psa_fastq.h) ------------------------------------------
// The file is cut into fixed byte ranges ("blocks") with no regard to record boundaries.  A lane thread reads block i
// plus a tail that reaches into block i + 1 straight into its lane's pinned buffer (several threads pread slices of
// it), the device indexes the newlines, cuts the records whose HEADER starts inside the block (their remaining lines may
// lie in the tail), maps them and formats their lines, and the lane writes the text at the block's place in the output.
// Two numbers chain the blocks: the count of newlines before a block -- line number mod 4 tells headers from the other
// lines as long as every record so far was four lines -- and the output offset.  Both are known early (right after the
// newline index resp. the formatting), so the lanes overlap almost completely.  The first block with anything but plain
// four-line ASCII records ends this pipeline: the blocks before it are committed, and the caller's host parser resumes at
// the first record that was not.
struct FastState {
    std::mutex mu;
    std::condition_variable cv;
    uint64_t next_block = 0;      // next block to hand to a lane
    uint64_t idx_turn = 0;        // block whose newline count is folded in next
    uint64_t lines_before = 0, reads_before = 0;   // ... before block idx_turn
    uint64_t commit_turn = 0;     // block that is committed (output offset assigned, counters added) next
    uint64_t out_pos = 0, resume_off = 0;
    uint64_t n_reads = 0, n_mapped = 0, n_aligned = 0, next_tick = 1000000;
    int64_t stop_block = -1;      // first block that is not plain (or failed): nothing from it on is committed
    int rc = PSA_OK;
    double busy_reader = 0, busy_mapper = 0, busy_writer = 0;
};
struct FastConfig {
    uint64_t block_bytes, tail_bytes;
    uint32_t lanes, io_threads, write_threads;   // lane threads; pread threads and writer threads of each lane
    bool verbose;
};
// runs `fn(t)` on `threads` threads (the calling one included)
template <class F>
void parallel_for(uint32_t threads, F&& fn) {
    std::vector<std::thread> th;
    for (uint32_t t = 1; t < threads; t++) th.emplace_back([&fn, t]() { fn(t); });
    fn(0);
    for (auto& x : th) x.join();
}
// Returns true when the whole file went through; false: the caller continues at st.resume_off (st.rc tells errors).
bool fast_blocks(psa_index* index, ByteSource& in, OutFile& of, FILE* out, int progress, const FastConfig& cfg, FastState& st) {
    const uint64_t S = in.size;
    bool needs_nl = false;   // the file does not end in '\n': bio reads the last line all the same
    if (S) {
        uint8_t lastb = 0;
        if (in.read_at(&lastb, 1, S - 1) != 1) {
            st.rc = PSA_ERR_IO;
            return false;
        }
        needs_nl = lastb != '\n';
    }
    const uint64_t S1 = S + (needs_nl ? 1 : 0);   // size with the virtual last newline
    const uint64_t B = cfg.block_bytes, T = cfg.tail_bytes;
    const uint64_t n_blocks = (S1 + B - 1) / B;
    st.out_pos = of.pos;
    if (!n_blocks) return true;
    auto now = []() { return std::chrono::steady_clock::now(); };
    auto secs = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
        return std::chrono::duration<double>(b - a).count();
    };
    const auto tc0 = now();
    const uint32_t lanes = (uint32_t)std::min<uint64_t>(cfg.lanes, n_blocks);
    std::vector<psa_fq_lane*> lane(lanes, nullptr);
    std::vector<int> lane_rc(lanes, PSA_OK);
    parallel_for(lanes, [&](uint32_t a) { lane_rc[a] = psa_fq_lane_create(index, B, T, &lane[a]); });   // (pinning the buffers is the slow part)
    for (uint32_t a = 0; a < lanes; a++)
        if (lane_rc[a]) {
            for (auto l : lane)
                if (l) psa_fq_lane_destroy(l);
            st.rc = lane_rc[a];
            return false;
        }
    if (cfg.verbose)
        fprintf(stderr, "psa: block pipeline: %llu blocks of %.1f MB (+ %.2f MB tail), %u lanes x %u I/O threads, lanes ready in %.1f ms, output %s\n",
                (unsigned long long)n_blocks, B / 1e6, T / 1e6, lanes, cfg.io_threads, 1e3 * secs(tc0, now()),
                of.fd < 0 ? "serial" : of.map_base ? "mapped" : "pwrite");
    auto lane_main = [&](uint32_t a) {
        psa_fq_lane* L = lane[a];
        uint8_t* text = psa_fq_lane_text(L);
        for (;;) {
            uint64_t i;
            {
                std::unique_lock<std::mutex> lk(st.mu);
                if (st.stop_block >= 0 || st.next_block >= n_blocks) return;
                i = st.next_block++;
            }
            // ---- read block i and its tail
            const auto t0 = now();
            const uint64_t lo = i * B, hi = std::min(S, lo + B + T);
            uint64_t len = hi - lo;
            std::vector<uint64_t> short_by(cfg.io_threads, 0);
            parallel_for(cfg.io_threads, [&](uint32_t t) {
                const uint64_t a0 = len * t / cfg.io_threads, a1 = len * (t + 1) / cfg.io_threads;
                if (a1 > a0) short_by[t] = (a1 - a0) - in.read_at(text + a0, a1 - a0, lo + a0);
            });
            bool io_fail = false;
            for (auto sb : short_by) io_fail |= sb != 0;   // (the file shrank under us)
            if (hi == S && needs_nl) text[len++] = '\n';
            const uint64_t own = std::min(B, S1 - lo);     // (== len in the last block, a multiple of 4096 elsewhere)
            const bool last = i + 1 == n_blocks;
            const auto t1 = now();
            // ---- newline index on the device
            uint64_t nl_own = 0, nl_total = 0;
            int rc = io_fail ? PSA_ERR_IO : psa_fq_lane_index(L, len, last ? len : own, &nl_own, &nl_total);
            const auto t2 = now();
            // ---- my turn in the line chain
            psa::FqOwned ow{};
            uint64_t reads_before = 0;
            bool plain = true;
            {
                std::unique_lock<std::mutex> lk(st.mu);
                st.cv.wait(lk, [&]() { return st.idx_turn == i; });
                st.busy_reader += secs(t0, t1);
                ow = psa::fq_owned_records(i, st.lines_before, nl_own);
                if (!rc) {
                    // a header right at the end of the file is no record
                    if (last && ow.n_rec && ow.j0 + 4 * (int64_t)(ow.n_rec - 1) == (int64_t)nl_total - 1) ow.n_rec--;
                    // every owned record needs its four lines inside block + tail; the file must end with a whole record
                    if (ow.n_rec && ow.j0 + 4 * (int64_t)(ow.n_rec - 1) + 4 > (int64_t)nl_total - 1) plain = false;
                    if (last && (st.lines_before + nl_total) % 4 != 0) plain = false;
                }
                reads_before = st.reads_before;
                st.lines_before += nl_own;
                st.reads_before += plain ? ow.n_rec : 0;
                st.idx_turn++;
                st.cv.notify_all();
            }
            const auto t2w = now();
            // ---- cut, map, format
            psa_fq_result res{};
            res.plain = plain ? 1 : 0;
            uint64_t tick_at[PSA_FQ_MAX_TICKS];
            uint32_t n_ticks = 0;
            if (!rc && plain) {
                if (progress)
                    for (uint64_t t = reads_before / 1000000 + 1; t * 1000000 <= reads_before + ow.n_rec && n_ticks < PSA_FQ_MAX_TICKS; t++)
                        tick_at[n_ticks++] = t * 1000000 - reads_before;
                rc = psa_fq_lane_run(L, ow.j0, ow.n_rec, tick_at, n_ticks, &res);
            }
            const auto t3 = now();
            // ---- commit in block order
            uint64_t my_out = 0;
            bool write_it = false;
            {
                std::unique_lock<std::mutex> lk(st.mu);
                st.cv.wait(lk, [&]() { return st.commit_turn == i; });
                st.busy_mapper += secs(t1, t2) + secs(t2w, t3);
                if (st.stop_block < 0) {
                    if (rc) {
                        st.rc = rc;
                        st.stop_block = (int64_t)i;
                    } else if (!res.plain) {
                        st.stop_block = (int64_t)i;
                    } else {
                        my_out = st.out_pos;
                        st.out_pos += res.out_bytes;
                        for (uint32_t t = 0; t < n_ticks; t++) {   // ref :497-504
                            char rate[64];
                            display_f32(rate, (float)(st.n_mapped + res.tick_mapped[t]) * 100.0f / (float)st.next_tick);
                            fprintf(stderr, "\rDone Mapping %llu reads w/ Rate: %s", (unsigned long long)st.next_tick, rate);
                            fflush(stderr);
                            st.next_tick += 1000000;
                        }
                        st.n_reads += ow.n_rec;
                        st.n_mapped += res.mapped;
                        st.n_aligned += res.aligned;
                        if (ow.n_rec) st.resume_off = lo + res.end_off;
                        write_it = res.out_bytes != 0;
                    }
                }
                if (of.fd >= 0 || !write_it) {   // positional writes need no order: pass the turn on at once
                    st.commit_turn++;
                    st.cv.notify_all();
                }
            }
            const auto t4 = now();
            if (write_it) {
                bool ok = true;
                if (of.fd >= 0) {
                    std::vector<char> failed(cfg.write_threads, 0);
                    parallel_for(cfg.write_threads, [&](uint32_t t) {
                        const uint64_t a0 = res.out_bytes * t / cfg.write_threads, a1 = res.out_bytes * (t + 1) / cfg.write_threads;
                        if (a1 > a0) failed[t] = !of.put(res.out_text + a0, a1 - a0, my_out + a0);
                    });
                    for (auto f : failed) ok &= !f;
                } else {
                    ok = fwrite(res.out_text, 1, res.out_bytes, out) == res.out_bytes;
                }
                std::unique_lock<std::mutex> lk(st.mu);
                st.busy_writer += secs(t4, now());
                if (!ok && !st.rc) st.rc = PSA_ERR_IO;
                if (of.fd < 0) {
                    st.commit_turn++;
                    st.cv.notify_all();
                }
            }
            if (cfg.verbose)
                fprintf(stderr, "psa: block %llu lane %u: %llu records, read %.1f ms, index %.1f ms, map+format %.1f ms (waited %.1f ms for its turns), write %.1f ms\n",
                        (unsigned long long)i, a, (unsigned long long)ow.n_rec, 1e3 * secs(t0, t1), 1e3 * secs(t1, t2), 1e3 * secs(t2w, t3),
                        1e3 * (secs(t2, t2w) + secs(t3, t4)), 1e3 * secs(t4, now()));
            std::unique_lock<std::mutex> lk(st.mu);
            if (st.stop_block >= 0) return;
        }
    };
    const auto tp0 = now();
    // The pages of the output mapping are faulted in ahead of the writers by two helper threads (a fresh page of a tmpfs /
    // page-cache file costs microseconds: allocation, zeroing, accounting), a bounded distance ahead of the committed output.
    std::atomic<bool> pipeline_done{false};
    std::vector<std::thread> prefault;
#ifdef MADV_POPULATE_WRITE
    if (of.map_base && !getenv("PSA_NO_PREFAULT")) {
        for (uint32_t h = 0; h < 2; h++)
            prefault.emplace_back([&, h]() {
                const uint64_t chunk = 4ull << 20, ahead = 160ull << 20;
                uint64_t at = (of.pos & ~(chunk - 1)) + h * chunk;
                while (!pipeline_done.load(std::memory_order_relaxed)) {
                    uint64_t committed;
                    {
                        std::unique_lock<std::mutex> lk(st.mu);
                        committed = st.out_pos;
                    }
                    while (at + chunk <= committed) at += 2 * chunk;   // never behind the writers
                    if (at + chunk > of.map_off + of.map_len) return;
                    if (at > committed + ahead) {
                        std::this_thread::sleep_for(std::chrono::microseconds(200));
                        continue;
                    }
                    const uint64_t lo = std::max(at, of.map_off);
                    if (madvise(of.map_base + (lo - of.map_off), (size_t)(at + chunk - lo), MADV_POPULATE_WRITE) != 0) return;
                    at += 2 * chunk;
                }
            });
    }
#endif
    {
        std::vector<std::thread> th;
        for (uint32_t a = 1; a < lanes; a++) th.emplace_back(lane_main, a);
        lane_main(0);
        for (auto& x : th) x.join();
    }
    pipeline_done.store(true);
    for (auto& x : prefault) x.join();
    const auto tp1 = now();
    parallel_for(lanes, [&](uint32_t a) { psa_fq_lane_destroy(lane[a]); });
    if (cfg.verbose)
        fprintf(stderr, "psa: block pipeline summary: lanes created in %.1f ms, blocks %.1f ms, lanes destroyed in %.1f ms\n", 1e3 * secs(tc0, tp0),
                1e3 * secs(tp0, tp1), 1e3 * secs(tp1, now()));
    of.pos = st.out_pos;
    return st.stop_block < 0 && st.rc == PSA_OK;
}

}  // namespace

// `{:?}` of a string as the drivers print it (shared with the Python mirror); returns the bytes written, or a
// negative value when `s` is not valid UTF-8 / `cap` is too small (10 * n + 2 always suffices)
extern "C" int64_t psa_debug_str(const char* s, uint64_t n, char* out, uint64_t cap) {
    if (!out || (n && !s) || cap < 10 * n + 2) return -1;
    if (!utf8_valid((const uint8_t*)s, n)) return -2;
    return (int64_t)(debug_str(out, s, n) - out);
}

extern "C" int psa_index_host_classes(const psa_index*, const uint64_t** eq_offsets, const uint32_t** eq_members, uint64_t* n_eq);

extern "C" int psa_process_reads(psa_index* index, const char* fastq_path, const char* out_path, uint32_t num_threads,
                                 uint64_t batch_reads, int progress, psa_process_stats* stats) {
    if (!index || !fastq_path) return PSA_ERR_ARG;
    if (!num_threads) num_threads = 1;
    const uint64_t batch_reads_arg = batch_reads;
    if (!batch_reads) batch_reads = 1ull << 19;  // one pipeline chunk of psa_mapper_map: small blocks keep the start-up short
    const auto t0 = std::chrono::steady_clock::now();

    ByteSource in;
    if (!in.open(fastq_path)) return PSA_ERR_IO;
    FILE* out = stdout;
    const bool own_out = out_path && strcmp(out_path, "-") != 0;
    if (own_out) {
        out = fopen(out_path, "w+b");   // (read access too: the writers map it)
        if (!out) {
            in.close();
            return PSA_ERR_IO;
        }
    }
    OutFile of;
    of.attach(out);
    if (in.parallel()) of.open_window(in.size);
    const double t_open = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    (void)t_open;   // (result lines are about a third of the FASTQ text; more than the window is pwritten)
    const bool verbose = getenv("PSA_VERBOSE") != nullptr;
    uint64_t n_reads = 0, n_mapped = 0, n_aligned = 0, next_tick = 1000000;
    double busy_reader = 0, busy_mapper = 0, busy_writer = 0;  // seconds spent working (not waiting) per stage

    // Plain files first go through the block pipeline whose text work is on the device (fast_blocks above); it stops at the
    // first record that is not a plain four-line ASCII record, and the host parser below takes over from there (gzip, pipes
    // and PSA_PROCESS_FAST=0 start there).  batch_reads sizes its blocks too (~320 bytes of FASTQ per 150-base record).
    bool fast_done = false;
    {
        const char* e = getenv("PSA_PROCESS_FAST");
        if (in.parallel() && !(e && atoi(e) == 0)) {
            FastConfig cfg;
            uint64_t bb = batch_reads_arg ? batch_reads_arg * 320 : (32ull << 20);
            if (const char* b = getenv("PSA_FQ_BLOCK_BYTES")) bb = strtoull(b, nullptr, 10);
            bb = std::min<uint64_t>(std::max<uint64_t>(bb, 4096), 1ull << 30);
            cfg.block_bytes = (bb + 4095) / 4096 * 4096;
            cfg.tail_bytes = std::min<uint64_t>(std::max<uint64_t>(cfg.block_bytes / 8, 4096), 1ull << 20);
            if (const char* t = getenv("PSA_FQ_TAIL_BYTES")) cfg.tail_bytes = strtoull(t, nullptr, 10);
            cfg.lanes = 3;
            if (const char* l = getenv("PSA_FQ_LANES")) cfg.lanes = (uint32_t)std::max(1, atoi(l));
            cfg.io_threads = std::max<uint32_t>(1, num_threads / cfg.lanes);
            cfg.write_threads = cfg.io_threads;
            if (const char* w = getenv("PSA_FQ_WRITE_THREADS")) cfg.write_threads = (uint32_t)std::max(1, atoi(w));
            cfg.verbose = verbose;
            FastState fs;
            fast_done = fast_blocks(index, in, of, out, progress, cfg, fs);
            n_reads = fs.n_reads; n_mapped = fs.n_mapped; n_aligned = fs.n_aligned; next_tick = fs.next_tick;
            busy_reader = fs.busy_reader; busy_mapper = fs.busy_mapper; busy_writer = fs.busy_writer;
            if (fs.rc) {
                of.finish();
                fflush(out);
                if (own_out) fclose(out);
                in.close();
                return fs.rc;
            }
            if (!fast_done) {
                in.pos = fs.resume_off;
                if (verbose) fprintf(stderr, "psa: block pipeline stopped after %llu reads; host parser resumes at byte %llu\n",
                                     (unsigned long long)n_reads, (unsigned long long)fs.resume_off);
            }
        }
    }

    int map_rc = PSA_OK, reader_rc = PSA_OK, writer_rc = PSA_OK;
    if (!fast_done) {
    HostClasses hc;
    psa_mapper* mapper = nullptr;
    int rc = psa_index_host_classes(index, &hc.off, &hc.mem, &hc.n_eq);
    if (!rc) rc = psa_mapper_create(index, 0, &mapper);
    if (rc) {
        of.finish();
        in.close();
        if (own_out) fclose(out);
        return rc;
    }

    constexpr int kSlots = 3;
    Batch slot[kSlots];
    std::mutex mu;
    std::condition_variable cv;
    bool abort_all = false;
    auto now = []() { return std::chrono::steady_clock::now(); };
    auto secs = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
        return std::chrono::duration<double>(b - a).count();
    };

    // stage 1: raw FASTQ text -> pinned block + record table, records cut exactly as bio's reader cuts them
    // (parse_record above; four-line records take the parallel fast path).  Bytes after the last complete
    // record of a block are carried over to the next one.
    // text per block: starts at 400 bytes per record, then follows the record size seen (+3 %), so that little
    // text is left over after the batch_reads-th record and has to be carried to the next block
    uint64_t block_bytes = std::max<uint64_t>(batch_reads * 400, 1ull << 22);
    std::thread reader([&]() {
        int s = 0;
        bool done = false, eof = false;
        std::vector<uint8_t> carry;
        U64Buf nl;
        std::vector<std::vector<uint64_t>> parts;
        while (!done) {
            Batch& b = slot[s];
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&]() { return b.state == 0 || abort_all; });
                if (abort_all) return;
            }
            const auto tr0 = now();
            double t_alloc = 0, t_read = 0;
            int err = PSA_OK;
            b.n = 0; b.tx_used = 0; b.last = false; b.text_len = 0;
            uint64_t have = 0;
            for (;;) {
                const uint64_t want = std::max<uint64_t>(block_bytes, carry.size() + (1u << 20));
                const auto ta0 = now();
                if ((err = b.text.reserve(want + 64, have))) break;
                t_alloc += secs(ta0, now());
                if (!have && !carry.empty()) {
                    memcpy(b.text.p, carry.data(), carry.size());
                    have = carry.size();
                    carry.clear();
                }
                const auto tq0 = now();
                const uint64_t room = std::min<uint64_t>(b.text.cap - 64 - have, want > have ? want - have : (1u << 20));
                bool oom = false;
                have += fill_and_index(in, b.text.p, have, room, num_threads, eof, parts, nl, oom);
                if (oom) {
                    err = PSA_ERR_NOMEM;
                    break;
                }
                if (eof && have && b.text.p[have - 1] != '\n') {  // unterminated last line
                    b.text.p[have] = '\n';
                    nl.p[nl.n++] = have++;
                }
                t_read += secs(tq0, now());
                if (nl.size() >= 4 || eof) break;
                // one record larger than the block: grow and read more
                block_bytes *= 2;
            }
            if (!err) {
                uint8_t* x = b.text.p;
                const uint64_t n_lines = nl.size();
                const LineIndex LI{x, nl, n_lines};
                // fast path: four-line records, cut and checked by all threads at once.  Record r = lines 4r .. 4r+3 is
                // what bio's reader would cut as long as every record before it was one too: header '@', a sequence
                // line that does not start with '+', a '+' line, a non-empty quality line.
                const uint64_t n4 = std::min<uint64_t>(n_lines / 4, batch_reads);
                b.off.resize(n4); b.len.resize(n4); b.id_off.resize(n4); b.id_len.resize(n4);
                std::vector<uint64_t> bad(num_threads, ~0ULL);
                std::vector<std::thread> th;
                for (uint32_t t = 0; t < num_threads; t++) {
                    th.emplace_back([&, t]() {
                        for (uint64_t r = n4 * t / num_threads; r < n4 * (t + 1) / num_threads; r++) {
                            const uint64_t l0 = LI.start(4 * r), e0 = LI.end(4 * r);
                            const uint64_t l1 = e0 + 1, e1 = LI.end(4 * r + 1);
                            const uint64_t l2 = e1 + 1, l3 = LI.end(4 * r + 2) + 1, e3 = LI.end(4 * r + 3);
                            bool ok = x[l0] == '@' && x[l1] != '+' && x[l2] == '+' && trim_end(x + l3, e3 - l3) > 0;
                            if (ok) {
                                uint8_t hi = 0;
                                for (uint64_t q = l0; q < e0; q++) hi |= x[q];
                                if ((hi & 0x80) && !utf8_valid(x + l0 + 1, e0 - l0 - 1)) ok = false;
                            }
                            if (!ok) {
                                bad[t] = r;
                                break;
                            }
                            const uint64_t he = l0 + 1 + trim_end(x + l0 + 1, e0 - l0 - 1);
                            uint64_t ie = l0 + 1;
                            while (ie < he && x[ie] != ' ') ie++;      // the id ends at the first SPACE (tabs stay in it)
                            b.id_off[r] = l0 + 1;
                            b.id_len[r] = (uint32_t)(ie - l0 - 1);
                            b.off[r] = l1;
                            b.len[r] = (uint32_t)trim_end(x + l1, e1 - l1);
                        }
                    });
                }
                for (auto& t : th) t.join();
                uint64_t n_rec = n4;
                for (uint32_t t = 0; t < num_threads; t++) n_rec = std::min(n_rec, bad[t]);
                uint64_t line = 4 * n_rec;
                // everything else -- wrapped sequences, malformed records, the end of the file -- one record at a time
                if (n_rec < n4 || (eof && line < n_lines && n_rec < batch_reads)) {
                    b.off.resize(n_rec); b.len.resize(n_rec); b.id_off.resize(n_rec); b.id_len.resize(n_rec);
                    while (n_rec < batch_reads && line < n_lines) {
                        uint64_t next = line, id_off = 0, seq_off = 0;
                        uint32_t id_len = 0, seq_len = 0;
                        const int pr = parse_record(x, LI, line, eof, next, id_off, id_len, seq_off, seq_len);
                        if (pr == REC_NEED_MORE) break;
                        if (pr == REC_BAD) {   // the complete records before the bad one are still processed
                            err = PSA_ERR_IO;
                            break;
                        }
                        b.off.push_back(seq_off); b.len.push_back(seq_len);
                        b.id_off.push_back(id_off); b.id_len.push_back(id_len);
                        n_rec++;
                        line = next;
                    }
                }
                b.n = n_rec;
                b.text_len = line ? nl[line - 1] + 1 : 0;
                if (n_rec == batch_reads)
                    block_bytes = std::max<uint64_t>(b.text_len + b.text_len / 32 + 4096, 1ull << 22);
                if (!err) {
                    carry.assign(b.text.p + b.text_len, b.text.p + have);
                    if (eof && carry.empty()) done = true;
                    if (!eof && n_rec == 0 && line == 0) block_bytes *= 2;   // not even one record in the block: read more
                }
            }
            busy_reader += secs(tr0, now());
            if (verbose)
                fprintf(stderr, "psa: reader block %llu records, %.1f MB: alloc %.0f ms, read + newline index %.0f ms, total %.0f ms\n",
                        (unsigned long long)b.n, b.text_len / 1e6, 1e3 * t_alloc, 1e3 * t_read, 1e3 * secs(tr0, now()));
            std::unique_lock<std::mutex> lk(mu);
            if (err) {
                reader_rc = err;
                done = true;
            }
            b.last = done;
            b.state = 1;
            cv.notify_all();
            s = (s + 1) % kSlots;
        }
    });

    // stage 3: format + write, in input order
    std::thread writer([&]() {
        int s = 0;
        std::vector<OutBuf> parts(num_threads);
        std::vector<uint64_t> mapped(num_threads), aligned(num_threads);
        std::vector<char> failed(num_threads);
        for (;;) {
            Batch& b = slot[s];
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&]() { return b.state == 2 || abort_all; });
                if (abort_all) return;
            }
            const auto tw0 = now();
            if (b.n) {
                std::vector<std::thread> th;
                // where each thread's share of the non-class sets' members starts in b.tx
                std::vector<uint64_t> nov(num_threads + 1, 0);
                const psa_hit_compact* ch = (const psa_hit_compact*)b.hits.p;
                for (uint32_t t = 0; t < num_threads; t++) {
                    uint64_t r0 = b.n * t / num_threads, r1 = b.n * (t + 1) / num_threads;
                    th.emplace_back([&, t, r0, r1]() {
                        uint64_t sum = 0;
                        for (uint64_t i = r0; i < r1; i++)
                            if (((ch[i].cov_flags >> 28) & PSA_FLAG_ALIGNED) && (ch[i].eq_or_n & 0x80000000u)) sum += ch[i].eq_or_n & 0x7FFFFFFFu;
                        nov[t + 1] = sum;
                    });
                }
                for (auto& x : th) x.join();
                th.clear();
                for (uint32_t t = 0; t < num_threads; t++) nov[t + 1] += nov[t];
                for (uint32_t t = 0; t < num_threads; t++) {
                    mapped[t] = aligned[t] = 0;
                    uint64_t r0 = b.n * t / num_threads, r1 = b.n * (t + 1) / num_threads;
                    th.emplace_back([&, t, r0, r1]() { format_range(b, hc, r0, r1, nov[t], parts[t], mapped[t], aligned[t]); });
                }
                for (auto& x : th) x.join();
                th.clear();
                const auto tw1 = now();
                const uint64_t mapped_before = n_mapped;
                for (uint32_t t = 0; t < num_threads; t++) {
                    if (!parts[t].ok) writer_rc = PSA_ERR_NOMEM;
                    n_mapped += mapped[t];
                    n_aligned += aligned[t];
                }
                if (of.fd >= 0) {
                    std::vector<uint64_t> at(num_threads + 1, of.pos);
                    for (uint32_t t = 0; t < num_threads; t++) at[t + 1] = at[t] + parts[t].n;
                    for (uint32_t t = 0; t < num_threads; t++)
                        th.emplace_back([&, t]() { failed[t] = !of.put(parts[t].p, parts[t].n, at[t]); });
                    for (auto& x : th) x.join();
                    for (uint32_t t = 0; t < num_threads; t++)
                        if (failed[t]) writer_rc = PSA_ERR_IO;
                    of.pos = at[num_threads];
                } else {
                    for (uint32_t t = 0; t < num_threads; t++)
                        if (parts[t].n && fwrite(parts[t].p, 1, parts[t].n, out) != parts[t].n) writer_rc = PSA_ERR_IO;
                }
                if (verbose)
                    fprintf(stderr, "psa: writer block %llu records: format %.0f ms, write %.0f ms (%s)\n", (unsigned long long)b.n,
                            1e3 * secs(tw0, tw1), 1e3 * secs(tw1, now()), of.fd >= 0 ? (of.map_base ? "every thread through the mapping" : "pwrite by every thread") : "serial");
                // ref :497-504: at every 1 000 000th read, the share of "mapped" reads so far (f32 arithmetic, `{}`)
                while (progress && n_reads + b.n >= next_tick) {
                    const uint64_t k = next_tick - n_reads;   // reads of this block up to the tick
                    uint64_t m = mapped_before;
                    for (uint64_t i = 0; i < k; i++) m += ((ch[i].cov_flags >> 28) & PSA_FLAG_MAPPED) != 0;
                    char rate[64];
                    display_f32(rate, (float)m * 100.0f / (float)next_tick);
                    fprintf(stderr, "\rDone Mapping %llu reads w/ Rate: %s", (unsigned long long)next_tick, rate);
                    fflush(stderr);
                    next_tick += 1000000;
                }
                n_reads += b.n;
            }
            busy_writer += secs(tw0, now());
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
    for (int s = 0;; s = (s + 1) % kSlots) {
        Batch& b = slot[s];
        {
            std::unique_lock<std::mutex> lk(mu);
            cv.wait(lk, [&]() { return b.state == 1; });
        }
        const auto tm0 = now();
        if (b.n) {
            map_rc = b.hits.reserve(b.n * sizeof(psa_hit_compact), 0);
            if (!map_rc) map_rc = b.tx.reserve(std::max<uint64_t>(b.n * 2 * 4, 4096), 0);   // members of the non-class sets only
            psa_read_batch r{};
            r.format = PSA_READS_ASCII;
            r.location = PSA_MEM_HOST;
            r.data = b.text.p;
            r.data_len = b.text_len;
            r.read_off = b.off.data();
            r.read_len = b.len.data();
            r.n_reads = b.n;
            psa_result_batch o{};
            o.location = PSA_MEM_HOST;
            o.flags = PSA_RESULT_COMPACT;   // 8 bytes per read back; class ids are expanded from the host's eq_classes
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
        busy_mapper += secs(tm0, now());
        if (verbose) fprintf(stderr, "psa: mapper block %llu records: %.0f ms\n", (unsigned long long)b.n, 1e3 * secs(tm0, now()));
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
    psa_mapper_destroy(mapper);
    for (auto& b : slot) { b.text.release(); b.hits.release(); b.tx.release(); }
    }  // host parser
    if (progress) fprintf(stderr, "\n");
    const auto tf0 = std::chrono::steady_clock::now();
    of.finish();
    fflush(out);
    if (own_out) fclose(out);
    in.close();
    if (verbose)
        fprintf(stderr, "psa: summary: files opened after %.1f ms, output finished (unmap, cut to size, close) in %.1f ms, total %.1f ms\n", 1e3 * t_open,
                1e3 * std::chrono::duration<double>(std::chrono::steady_clock::now() - tf0).count(),
                1e3 * std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
    if (stats) {
        stats->reads = n_reads;
        stats->mapped = n_mapped;
        stats->aligned = n_aligned;
        stats->seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        stats->reader_seconds = busy_reader;
        stats->mapper_seconds = busy_mapper;
        stats->writer_seconds = busy_writer;
    }
    if (map_rc) return map_rc;
    if (reader_rc) return reader_rc;  // the reference panics on a malformed record (:446)
    return writer_rc;
}
