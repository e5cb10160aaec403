// psa_fastq.cuh -- the text side of process_reads (ref src/pseudoaligner.rs:420-514) on the device:
// raw FASTQ bytes in, the reference's `{:?}` result lines out.  A block of the file is copied to HBM as it
// is; kernels find its newlines, cut the four-line records exactly as bio::io::fastq::Reader (bio 1.5, the
// reader the reference pulls records from, ref :421,431,442-447 and src/utils.rs:152-157) would cut them,
// hand the sequences to the map kernels by offset, and format one `(flag, "id", [tx, ...], coverage)` line
// per read (ref :490) into a text buffer that goes back to the host in one copy.
//
// Only what is certain to be cut identically is handled here: records of exactly four lines with an ASCII
// header; anything else (wrapped sequences, a '+' where a sequence should be, bytes >= 0x80 where bio would
// decode or trim UTF-8, a truncated file) makes the block report "not plain" and the host parser of
// process_reads.cpp -- the full restatement of bio's state machine -- takes over from that record on.
//
// The per-record routines are __host__ __device__: tests/hostsim/process_stub.cpp runs the very same text
// serially so that the block pipeline of process_reads.cpp is tested on machines without a GPU.
#pragma once
#include <stdint.h>

#include "psa_core.cuh"

namespace psa {

// ---- record cutting ---------------------------------------------------------------------------
// Rust's str::trim_end over ASCII white space (the multi-byte White_Space code points make the caller fall back)
PSA_HD bool fq_is_space(uint8_t c) { return c == ' ' || (c >= 0x09 && c <= 0x0D); }
PSA_HD uint32_t fq_trim_end(const uint8_t* s, uint32_t n) {
    while (n && fq_is_space(s[n - 1])) n--;
    return n;
}
struct FqRecord {
    uint64_t seq_off;  // offsets into the block's text
    uint32_t seq_len, id_off, id_len;
    uint32_t end;      // first byte after the record's last line
};
// The record whose header line starts right after newline j of the block (j = -1: at byte 0); nl = positions of
// the block's newlines, ascending; the caller guarantees that newlines j+1 .. j+4 exist.  Returns false when
// the four lines are not a record bio would cut as header / sequence / '+' / quality with ASCII-only trimming.
PSA_HD bool fq_cut_record(const uint8_t* x, const uint32_t* nl, int64_t j, FqRecord& r) {
    const uint32_t l0 = j < 0 ? 0u : nl[j] + 1, e0 = nl[j + 1];
    const uint32_t l1 = e0 + 1, e1 = nl[j + 2];
    const uint32_t l2 = e1 + 1, e2 = nl[j + 3];
    const uint32_t l3 = e2 + 1, e3 = nl[j + 4];
    r.seq_off = l1; r.seq_len = 0; r.id_off = l0 + 1; r.id_len = 0; r.end = e3 + 1;
    // header: '@' (else Error::MissingAt), ASCII only (bio reads lines as UTF-8 and trims Unicode white space)
    if (x[l0] != '@') return false;
    uint8_t hi = 0;
    for (uint32_t q = l0; q < e0; q++) hi |= x[q];
    if (hi & 0x80) return false;
    // sequence: one line that does not start with '+'; then the '+' line; then a non-empty quality line
    if (x[l1] == '+' || x[l2] != '+') return false;
    const uint32_t sl = fq_trim_end(x + l1, e1 - l1), ql = fq_trim_end(x + l3, e3 - l3);
    if (ql == 0) return false;                                  // Error::IncompleteRecord
    if ((sl && (x[l1 + sl - 1] & 0x80)) || (x[l3 + ql - 1] & 0x80)) return false;   // may end in multi-byte white space
    // id = header[1..].trim_end() up to the first ' ' (tabs stay in it)
    const uint32_t he = l0 + 1 + fq_trim_end(x + l0 + 1, e0 - l0 - 1);
    uint32_t ie = l0 + 1;
    while (ie < he && x[ie] != ' ') ie++;
    r.id_len = ie - (l0 + 1);
    r.seq_len = sl;
    return true;
}

// ---- line formatting: `(flag, "id", [tx, ...], coverage)\n`, the tuple printed at ref :490 --------------
PSA_HD uint32_t fq_dec_len(uint32_t v) {
    return v < 10 ? 1 : v < 100 ? 2 : v < 1000 ? 3 : v < 10000 ? 4 : v < 100000 ? 5 : v < 1000000 ? 6
         : v < 10000000 ? 7 : v < 100000000 ? 8 : v < 1000000000 ? 9 : 10;
}
PSA_HD char* fq_put_u32(char* o, uint32_t v) {
    const uint32_t len = fq_dec_len(v);
    char* e = o + len;
    do {
        const uint32_t q = v / 10;
        *--e = (char)('0' + (v - q * 10));
        v = q;
    } while (v);
    return o + len;
}
// Rust's `{:?}` of an ASCII string (char::escape_debug): \" \\ \n \r \t \0, \u{hex} for the other control
// characters and DEL, everything else as it is
PSA_HD uint32_t fq_escaped_len(const uint8_t* s, uint32_t n) {
    uint32_t len = 2;
    for (uint32_t i = 0; i < n; i++) {
        const uint8_t c = s[i];
        if (c >= 0x20 && c < 0x7f) len += (c == '"' || c == '\\') ? 2 : 1;
        else if (c == '\n' || c == '\r' || c == '\t' || c == 0) len += 2;
        else len += c < 0x10 ? 5 : 6;
    }
    return len;
}
PSA_HD char* fq_put_escaped(char* o, const uint8_t* s, uint32_t n) {
    *o++ = '"';
    for (uint32_t i = 0; i < n; i++) {
        const uint8_t c = s[i];
        if (c >= 0x20 && c < 0x7f && c != '"' && c != '\\') {
            *o++ = (char)c;
        } else if (c == '"' || c == '\\') {
            *o++ = '\\'; *o++ = (char)c;
        } else if (c == '\n') {
            *o++ = '\\'; *o++ = 'n';
        } else if (c == '\r') {
            *o++ = '\\'; *o++ = 'r';
        } else if (c == '\t') {
            *o++ = '\\'; *o++ = 't';
        } else if (c == 0) {
            *o++ = '\\'; *o++ = '0';
        } else {
            *o++ = '\\'; *o++ = 'u'; *o++ = '{';
            const uint32_t h = c >> 4, l = c & 15;
            if (h) *o++ = (char)(h < 10 ? '0' + h : 'a' + h - 10);
            *o++ = (char)(l < 10 ? '0' + l : 'a' + l - 10);
            *o++ = '}';
        }
    }
    *o++ = '"';
    return o;
}
PSA_HD uint32_t fq_line_len(bool flag, const uint8_t* id, uint32_t id_len, const uint32_t* tx, uint32_t n_tx, uint32_t coverage) {
    uint32_t len = (flag ? 7u : 8u) + fq_escaped_len(id, id_len) + 3;   // `(true, ` | `(false, `, the id, `, [`
    for (uint32_t j = 0; j < n_tx; j++) len += fq_dec_len(tx[j]);
    if (n_tx) len += 2 * (n_tx - 1);
    return len + 3 + fq_dec_len(coverage) + 2;                          // `], `, coverage, `)\n`
}
PSA_HD char* fq_format_line(char* o, bool flag, const uint8_t* id, uint32_t id_len, const uint32_t* tx, uint32_t n_tx, uint32_t coverage) {
    *o++ = '(';
    if (flag) { *o++ = 't'; *o++ = 'r'; *o++ = 'u'; *o++ = 'e'; }
    else { *o++ = 'f'; *o++ = 'a'; *o++ = 'l'; *o++ = 's'; *o++ = 'e'; }
    *o++ = ','; *o++ = ' ';
    o = fq_put_escaped(o, id, id_len);
    *o++ = ','; *o++ = ' '; *o++ = '[';
    for (uint32_t j = 0; j < n_tx; j++) {
        if (j) { *o++ = ','; *o++ = ' '; }
        o = fq_put_u32(o, tx[j]);
    }
    *o++ = ']'; *o++ = ','; *o++ = ' ';
    o = fq_put_u32(o, coverage);
    *o++ = ')'; *o++ = '\n';
    return o;
}

// ---- which records a block owns ------------------------------------------------------------------
// A block is a fixed byte range of the file plus a tail that reaches into the next one.  With `lines_before`
// newlines in the file before the block, the line that starts after the block's newline j is line
// lines_before + j + 1 of the file, and -- every record before it being four lines -- a header iff that is a
// multiple of 4.  The block owns the records whose header follows one of its first `nl_own` newlines (block 0:
// also the record at byte 0), i.e. whose header starts inside [1, own_bytes] of the block.
struct FqOwned {
    int64_t j0;       // newline index before the first owned header (-1: byte 0)
    uint64_t n_rec;
};
PSA_HD FqOwned fq_owned_records(uint64_t block, uint64_t lines_before, uint64_t nl_own) {
    FqOwned o;
    o.j0 = block == 0 ? -1 : (int64_t)((4 - (lines_before + 1) % 4) % 4);
    o.n_rec = o.j0 < (int64_t)nl_own ? (uint64_t)(((int64_t)nl_own - 1 - o.j0) / 4 + 1) : 0;
    return o;
}

#if defined(__CUDACC__)
// ---- kernels -------------------------------------------------------------------------------------
constexpr uint32_t kFqTile = 4096;       // bytes per CTA of the newline kernels (256 threads x 16 bytes)
__device__ __forceinline__ uint32_t fq_nl_mask16(uint4 v) {  // bit b set iff byte b of the 16 is '\n'
    uint32_t m = 0;
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const uint32_t eq = __vcmpeq4(w[i], 0x0A0A0A0Au);   // 0xFF per equal byte
        m |= ((eq & 1u) | ((eq >> 7) & 2u) | ((eq >> 14) & 4u) | ((eq >> 21) & 8u)) << (4 * i);
    }
    return m;
}
// newlines per tile; text is padded with zero bytes up to a multiple of kFqTile
__global__ void __launch_bounds__(256) k_fq_count(const uint8_t* text, uint32_t* tile_count) {
    const uint4 v = reinterpret_cast<const uint4*>(text + (uint64_t)blockIdx.x * kFqTile)[threadIdx.x];
    uint32_t c = __popc(fq_nl_mask16(v));
    c = __reduce_add_sync(0xffffffffu, c);
    __shared__ uint32_t ws[8];
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) tile_count[blockIdx.x] = ws[0] + ws[1] + ws[2] + ws[3] + ws[4] + ws[5] + ws[6] + ws[7];
}
// positions of the newlines, ascending (tile_off = exclusive sum of tile_count)
__global__ void __launch_bounds__(256) k_fq_positions(const uint8_t* text, const uint32_t* tile_off, uint32_t* nl) {
    const uint64_t base = (uint64_t)blockIdx.x * kFqTile + 16u * threadIdx.x;
    const uint4 v = reinterpret_cast<const uint4*>(text + (uint64_t)blockIdx.x * kFqTile)[threadIdx.x];
    uint32_t m = fq_nl_mask16(v);
    const uint32_t c = __popc(m);
    // exclusive scan of c over the CTA
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= (unsigned)d) incl += t;
    }
    __shared__ uint32_t ws[8];
    if (lane == 31) ws[warp] = incl;
    __syncthreads();
    uint32_t before = tile_off[blockIdx.x] + incl - c;
    for (unsigned w = 0; w < warp; w++) before += ws[w];
    while (m) {
        const uint32_t b = __ffs(m) - 1;
        nl[before++] = (uint32_t)(base + b);
        m &= m - 1;
    }
}
// cut the block's records; status[0] |= 1 when one of them is not a plain four-line record
__global__ void __launch_bounds__(128) k_fq_records(const uint8_t* text, const uint32_t* nl, int64_t j0, uint64_t n_rec, uint64_t* seq_off,
                                                    uint32_t* seq_len, uint32_t* id_off, uint32_t* id_len, uint32_t* status) {
    const uint64_t r = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (r >= n_rec) return;
    FqRecord rec;
    const bool ok = fq_cut_record(text, nl, j0 + 4 * (int64_t)r, rec);
    seq_off[r] = rec.seq_off;
    seq_len[r] = ok ? rec.seq_len : 0u;
    id_off[r] = rec.id_off;
    id_len[r] = ok ? rec.id_len : 0u;
    if (!ok) atomicOr(status, 1u);
}
// bytes of every read's line; counters[0] += reads with the "mapped" flag, counters[1] += aligned reads
__global__ void __launch_bounds__(128) k_fq_line_len(const uint8_t* text, const uint32_t* id_off, const uint32_t* id_len, const HitRec* hits,
                                                     const uint32_t* tx, uint64_t n, uint32_t* line_len, unsigned long long* counters) {
    const uint64_t r = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    uint32_t mapped = 0, aligned = 0;
    if (r < n) {
        const HitRec h = hits[r];
        mapped = (h.flags & kFlagMapped) ? 1u : 0u;
        aligned = h.flags & kFlagAligned;
        line_len[r] = fq_line_len(mapped != 0, text + id_off[r], id_len[r], tx + h.tx_off, h.n_tx, h.coverage);
    }
    mapped = __reduce_add_sync(0xffffffffu, mapped);
    aligned = __reduce_add_sync(0xffffffffu, aligned);
    if ((threadIdx.x & 31) == 0) {
        if (mapped) atomicAdd(counters, (unsigned long long)mapped);
        if (aligned) atomicAdd(counters + 1, (unsigned long long)aligned);
    }
}
// The lines themselves.  A CTA's 128 lines are one contiguous byte range of the output: every thread formats its
// line into shared memory, then the CTA copies the range out with aligned 16-byte stores (the staging area starts at
// the range's own offset modulo 16, so that shared and global addresses are congruent).  A range that does not fit
// (very long transcript lists) is written line by line straight to global memory.
constexpr int kFqFmtBlock = 128;
constexpr uint32_t kFqFmtStage = 32 * 1024;
__global__ void __launch_bounds__(kFqFmtBlock) k_fq_format(const uint8_t* text, const uint32_t* id_off, const uint32_t* id_len, const HitRec* hits,
                                                           const uint32_t* tx, const uint64_t* line_off, uint64_t n, char* out) {
    __shared__ __align__(16) char stage[kFqFmtStage + 16];
    const uint64_t r0 = blockIdx.x * (uint64_t)kFqFmtBlock;
    const uint32_t nr = (uint32_t)min((uint64_t)kFqFmtBlock, n - r0);
    const uint64_t o0 = line_off[r0], o1 = line_off[r0 + nr];
    const uint32_t pad = (uint32_t)(o0 & 15);
    const bool staged = o1 - o0 <= kFqFmtStage;
    if (threadIdx.x < nr) {
        const uint64_t r = r0 + threadIdx.x;
        const HitRec h = hits[r];
        const uint64_t at = line_off[r];
        char* dst = staged ? stage + pad + (uint32_t)(at - o0) : out + at;
        fq_format_line(dst, (h.flags & kFlagMapped) != 0, text + id_off[r], id_len[r], tx + h.tx_off, h.n_tx, h.coverage);
    }
    if (!staged) return;
    __syncthreads();
    uint64_t a0 = (o0 + 15) & ~15ULL;
    if (a0 > o1) a0 = o1;
    uint64_t a1 = o1 & ~15ULL;
    if (a1 < a0) a1 = a0;
    for (uint64_t i = o0 + threadIdx.x; i < a0; i += kFqFmtBlock) out[i] = stage[pad + (uint32_t)(i - o0)];
    for (uint64_t q = a0 + 16ULL * threadIdx.x; q < a1; q += 16ULL * kFqFmtBlock)
        *reinterpret_cast<uint4*>(out + q) = *reinterpret_cast<const uint4*>(stage + pad + (uint32_t)(q - o0));
    for (uint64_t i = a1 + threadIdx.x; i < o1; i += kFqFmtBlock) out[i] = stage[pad + (uint32_t)(i - o0)];
}
// reads with the "mapped" flag among the first k[t] reads of the block (the progress line of ref :497-504)
__global__ void k_fq_mapped_prefix(const HitRec* hits, const uint64_t* k, uint32_t n_ticks, unsigned long long* out) {
    for (uint32_t t = 0; t < n_ticks; t++) {
        unsigned long long c = 0;
        for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < k[t]; i += gridDim.x * (uint64_t)blockDim.x)
            c += (hits[i].flags & kFlagMapped) ? 1 : 0;
        c = __reduce_add_sync(0xffffffffu, (unsigned)c);
        if ((threadIdx.x & 31) == 0 && c) atomicAdd(out + t, c);
    }
}
struct FqLenToU64 {
    const uint32_t* len;
    uint64_t n;
    __host__ __device__ uint64_t operator()(uint64_t i) const { return i < n ? (uint64_t)len[i] : 0ull; }
};
#endif  // __CUDACC__

}  // namespace psa
