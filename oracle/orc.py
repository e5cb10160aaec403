"""ctypes binding of the CPU oracle (oracle/psa_oracle.c).

TEST INFRASTRUCTURE ONLY.  Imported by tests/, by __graft_entry__.smoke() as the checker
and by bench.py's cpu_baseline / --impl reference legs.  Nothing under
rust-pseudoaligner_b200/ may import this module.
"""
import ctypes as C
import gzip
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libpsa_oracle.so")

EQ_NONE = 0xFFFFFFFF
FLAG_ALIGNED = 1
FLAG_MAPPED = 2

HIT_DTYPE = np.dtype(
    [("coverage", "<u4"), ("n_tx", "<u4"), ("tx_off", "<u8"), ("eq_id", "<u4"), ("flags", "<u4")]
)

EVENT_FIELDS = (
    "reads", "read_bases", "kmer_lookups", "dict_hits", "node_visits",
    "bases_compared", "edge_jumps", "class_members", "out_members", "aligned",
)


class Events(C.Structure):
    _fields_ = [(f, C.c_uint64) for f in EVENT_FIELDS]

    def as_dict(self):
        return {f: int(getattr(self, f)) for f in EVENT_FIELDS}


def build_lib(force=False):
    """Compile the oracle with its committed Makefile (gcc only)."""
    srcs = [os.path.join(_HERE, n) for n in ("psa_oracle.c", "c3_driver.c", "psa_oracle.h")]
    if (not force and os.path.exists(_LIB_PATH)
            and os.path.getmtime(_LIB_PATH) >= max(os.path.getmtime(s) for s in srcs)):
        return _LIB_PATH
    subprocess.check_call(["make", "-C", _HERE, "-B", "libpsa_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build_lib()
    L = C.CDLL(_LIB_PATH)
    vp, u32, u64 = C.c_void_p, C.c_uint32, C.c_uint64
    L.orc_last_error.restype = C.c_char_p
    L.orc_words_for.restype = u64
    L.orc_words_for.argtypes = [u64]
    L.orc_pack_ascii.argtypes = [C.c_char_p, u64, vp]
    L.orc_index_build.restype = vp
    L.orc_index_build.argtypes = [vp, vp, u32, u32]
    L.orc_index_from_flat.restype = vp
    L.orc_index_from_flat.argtypes = [u32, u64, vp, u64, vp, vp, vp, vp, u64, vp, vp]
    L.orc_index_free.argtypes = [vp]
    for name in ("n_nodes", "n_kmers", "n_eq", "n_seq_words", "n_pure_cycles"):
        f = getattr(L, "orc_index_" + name)
        f.restype, f.argtypes = u64, [vp]
    L.orc_index_k.restype, L.orc_index_k.argtypes = u32, [vp]
    for name in ("seq_words", "node_start", "node_len", "node_exts", "node_eq", "eq_offsets",
                 "eq_members", "succ", "pred"):
        f = getattr(L, "orc_index_" + name)
        f.restype, f.argtypes = vp, [vp]
    L.orc_index_lookup.restype = C.c_int
    L.orc_index_lookup.argtypes = [vp, vp, C.POINTER(u32), C.POINTER(u32)]
    L.orc_map_read.restype = C.c_int
    L.orc_map_read.argtypes = [vp, vp, u32, vp, u64, C.POINTER(u32), C.POINTER(u32), C.POINTER(u32),
                               vp, u32, C.POINTER(u32), vp]
    L.orc_map_batch.restype = C.c_int
    L.orc_map_batch.argtypes = [vp, vp, vp, vp, u64, vp, vp, u64, C.POINTER(u64), vp, vp]
    L.orc_map_batch_with_mismatch.restype = C.c_int
    L.orc_map_batch_with_mismatch.argtypes = [vp, vp, vp, vp, u64, u32, vp, vp, u64, C.POINTER(u64), vp, vp]
    L.orc_map_read_with_mismatch.restype = C.c_int
    L.orc_map_read_with_mismatch.argtypes = [vp, vp, u32, u32, vp, u64, C.POINTER(u32), C.POINTER(u32), C.POINTER(u32),
                                             vp, u32, C.POINTER(u32), vp]
    L.orc_map_ascii_batch.restype = C.c_int
    L.orc_map_ascii_batch.argtypes = [vp, vp, u64, u32, u64, u32, vp, vp, u64, C.POINTER(u64), vp, vp]
    L.orc_result_checksum.restype = u64
    L.orc_result_checksum.argtypes = [vp, vp, u64, u64]
    L.orc_process_reads_c3.restype = C.c_int
    L.orc_process_reads_c3.argtypes = [vp, C.c_char_p, C.c_char_p, u32, u32, C.POINTER(u64), C.POINTER(u64)]
    L.orc_intersect.restype = u32
    L.orc_intersect.argtypes = [vp, u32, vp, u32]
    _lib = L
    return L


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _view(ptr, n, dtype):
    if n == 0:
        return np.zeros(0, dtype=dtype)
    nbytes = n * np.dtype(dtype).itemsize
    buf = (C.c_char * nbytes).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype, count=n)


# ---------------------------------------------------------------- sequences
_CODE = np.zeros(256, dtype=np.uint8)  # QUIRK-6: everything that is not ACGTacgt -> 0
for _c, _v in (("A", 0), ("C", 1), ("G", 2), ("T", 3)):
    _CODE[ord(_c)] = _v
    _CODE[ord(_c.lower())] = _v


def encode(seq):
    """ASCII (str/bytes) -> one byte per base, 0..3."""
    if isinstance(seq, str):
        seq = seq.encode()
    return _CODE[np.frombuffer(seq, dtype=np.uint8)]


def pack_codes(codes):
    """uint8 codes -> uint64 words in debruijn DnaString packing (base i at bits 62-2*(i%32))."""
    n = len(codes)
    nw = (n + 31) // 32
    pad = np.zeros(nw * 32, dtype=np.uint64)
    pad[:n] = codes
    shifts = (62 - 2 * np.arange(32, dtype=np.uint64)).astype(np.uint64)
    return np.bitwise_or.reduce(pad.reshape(nw, 32) << shifts, axis=1).astype(np.uint64)


def pack_ascii(seq):
    if isinstance(seq, str):
        seq = seq.encode()
    nw = (len(seq) + 31) // 32
    out = np.zeros(max(nw, 1) + 1, dtype=np.uint64)
    lib().orc_pack_ascii(seq, len(seq), _ptr(out))
    return out[:nw]


def pack_reads(seqs):
    """list of ASCII reads -> (words, read_off[words], read_len) with one pad word at the end."""
    lens = np.array([len(s) for s in seqs], dtype=np.uint32)
    nws = (lens.astype(np.uint64) + 31) // 32
    off = np.zeros(len(seqs), dtype=np.uint64)
    if len(seqs):
        off[1:] = np.cumsum(nws)[:-1]
    words = np.zeros(int(nws.sum()) + 1, dtype=np.uint64)
    for i, s in enumerate(seqs):
        w = pack_ascii(s)
        words[int(off[i]):int(off[i]) + len(w)] = w
    return words, off, lens


def read_fasta(path):
    """-> (names, [ascii bytes]) ; record order = transcript index (ref src/utils.rs:71-88)."""
    op = gzip.open if str(path).endswith(".gz") else open
    names, seqs, cur = [], [], []
    with op(path, "rt") as f:
        for line in f:
            line = line.rstrip("\r\n")
            if line.startswith(">"):
                if names:
                    seqs.append("".join(cur).encode())
                names.append(line[1:])
                cur = []
            elif line:
                cur.append(line)
    if names:
        seqs.append("".join(cur).encode())
    return names, seqs


def read_fastq(path):
    """-> [(id, ascii seq)]; 4-line records."""
    op = gzip.open if str(path).endswith(".gz") else open
    out = []
    with op(path, "rt") as f:
        lines = [l.rstrip("\r\n") for l in f]
    for i in range(0, len(lines) - 3, 4):
        assert lines[i].startswith("@") and lines[i + 2].startswith("+")
        out.append((lines[i][1:].split()[0], lines[i + 1]))
    return out


def concat_transcripts(seqs):
    """list of ASCII transcripts -> (codes uint8, tx_off uint64[n+1])."""
    lens = np.array([len(s) for s in seqs], dtype=np.uint64)
    off = np.zeros(len(seqs) + 1, dtype=np.uint64)
    off[1:] = np.cumsum(lens)
    codes = encode(b"".join(bytes(s) if not isinstance(s, str) else s.encode() for s in seqs)) if len(seqs) else np.zeros(0, np.uint8)
    return np.ascontiguousarray(codes), off


# ---------------------------------------------------------------- index
class OrcIndex:
    def __init__(self, handle):
        if not handle:
            raise RuntimeError("oracle index: " + lib().orc_last_error().decode())
        self.h = C.c_void_p(handle)
        L = lib()
        self.k = int(L.orc_index_k(self.h))
        self.n_nodes = int(L.orc_index_n_nodes(self.h))
        self.n_kmers = int(L.orc_index_n_kmers(self.h))
        self.n_eq = int(L.orc_index_n_eq(self.h))
        self.n_seq_words = int(L.orc_index_n_seq_words(self.h))
        self.n_pure_cycles = int(L.orc_index_n_pure_cycles(self.h))

    @classmethod
    def build(cls, seqs, k):
        codes, off = concat_transcripts(seqs)
        return cls.build_codes(codes, off, k)

    @classmethod
    def build_codes(cls, codes, tx_off, k):
        codes = np.ascontiguousarray(codes, dtype=np.uint8)
        tx_off = np.ascontiguousarray(tx_off, dtype=np.uint64)
        return cls(lib().orc_index_build(_ptr(codes), _ptr(tx_off), len(tx_off) - 1, k))

    @classmethod
    def from_flat(cls, flat):
        f = {k_: (np.ascontiguousarray(v) if isinstance(v, np.ndarray) else v) for k_, v in flat.items()}
        assert f["seq_words"].dtype == np.uint64 and f["node_start"].dtype == np.uint64
        assert f["node_len"].dtype == np.uint32 and f["node_exts"].dtype == np.uint8
        assert f["node_eq"].dtype == np.uint32 and f["eq_offsets"].dtype == np.uint64
        assert f["eq_members"].dtype == np.uint32
        h = lib().orc_index_from_flat(
            int(f["k"]), len(f["node_len"]), _ptr(f["seq_words"]), len(f["seq_words"]),
            _ptr(f["node_start"]), _ptr(f["node_len"]), _ptr(f["node_exts"]), _ptr(f["node_eq"]),
            len(f["eq_offsets"]) - 1, _ptr(f["eq_offsets"]), _ptr(f["eq_members"]))
        return cls(h)

    def flat(self):
        """The index in the flat form of psa_index_desc (copies)."""
        L = lib()
        eq_offsets = _view(L.orc_index_eq_offsets(self.h), self.n_eq + 1, np.uint64).copy()
        return {
            "k": self.k,
            "seq_words": _view(L.orc_index_seq_words(self.h), self.n_seq_words, np.uint64).copy(),
            "node_start": _view(L.orc_index_node_start(self.h), self.n_nodes, np.uint64).copy(),
            "node_len": _view(L.orc_index_node_len(self.h), self.n_nodes, np.uint32).copy(),
            "node_exts": _view(L.orc_index_node_exts(self.h), self.n_nodes, np.uint8).copy(),
            "node_eq": _view(L.orc_index_node_eq(self.h), self.n_nodes, np.uint32).copy(),
            "eq_offsets": eq_offsets,
            "eq_members": _view(L.orc_index_eq_members(self.h), int(eq_offsets[-1]), np.uint32).copy(),
        }

    def edges(self):
        L = lib()
        return (_view(L.orc_index_succ(self.h), 4 * self.n_nodes, np.uint32).reshape(-1, 4).copy(),
                _view(L.orc_index_pred(self.h), 4 * self.n_nodes, np.uint32).reshape(-1, 4).copy())

    def lookup(self, kmer_ascii):
        w = np.zeros(3, dtype=np.uint64)
        p = pack_ascii(kmer_ascii)
        w[:len(p)] = p
        n, o = C.c_uint32(), C.c_uint32()
        r = lib().orc_index_lookup(self.h, _ptr(w), C.byref(n), C.byref(o))
        return (n.value, o.value) if r else None

    def map_read(self, seq, want_nodes=False, allowed=2):
        """Pseudoaligner::map_read: None or (sorted tx list, coverage[, nodes])."""
        if isinstance(seq, str):
            seq = seq.encode()
        words = np.zeros((len(seq) + 31) // 32 + 2, dtype=np.uint64)
        w = pack_ascii(seq)
        words[:len(w)] = w
        return self.map_packed(words, len(seq), want_nodes, allowed)

    def map_packed(self, words, length, want_nodes=False, allowed=2):
        cap = 1 << 16
        while True:
            tx = np.zeros(cap, dtype=np.uint32)
            nodes = np.zeros(2 * length + 2, dtype=np.uint32)
            n_tx, cov, eq, nn = C.c_uint32(), C.c_uint32(), C.c_uint32(), C.c_uint32()
            r = lib().orc_map_read_with_mismatch(self.h, _ptr(words), length, int(allowed), _ptr(tx), cap, C.byref(n_tx),
                                                 C.byref(cov), C.byref(eq), _ptr(nodes), len(nodes), C.byref(nn), None)
            if r >= 0:
                break
            cap *= 16
        if r == 0:
            return None
        res = (tx[:n_tx.value].tolist(), cov.value)
        if want_nodes:
            res = res + (nodes[:nn.value].tolist(),)
        return res

    def map_batch(self, words, read_off, read_len, counts=False, start=0, stop=None, allowed=2):
        """process_reads inner loop over reads [start, stop): (hits, tx_buf, counts|None, events)."""
        words = np.ascontiguousarray(words, dtype=np.uint64)
        read_off = np.ascontiguousarray(read_off, dtype=np.uint64)
        read_len = np.ascontiguousarray(read_len, dtype=np.uint32)
        stop = len(read_len) if stop is None else stop
        n = stop - start
        hits = np.zeros(n, dtype=HIT_DTYPE)
        cap = max(16 * n, 1024)
        cnt = np.zeros(self.n_eq + 2, dtype=np.uint64) if counts else None
        while True:
            tx = np.zeros(cap, dtype=np.uint32)
            used = C.c_uint64()
            ev = Events()
            if cnt is not None:
                cnt[:] = 0
            r = lib().orc_map_batch_with_mismatch(self.h, _ptr(words), _ptr(read_off[start:stop]),
                                                  _ptr(read_len[start:stop]), n, int(allowed), _ptr(hits), _ptr(tx), cap,
                                                  C.byref(used), _ptr(cnt) if cnt is not None else None, C.byref(ev))
            if r == 0:
                break
            cap = int(used.value) + 16
        return hits, tx[:used.value].copy(), cnt, ev.as_dict()

    def map_ascii_fixed(self, data, n, length, stride=None, start=0, stop=None, allowed=2, counts=False):
        """process_reads worker body over fixed-stride ASCII reads [start, stop): packs, then maps."""
        data = np.ascontiguousarray(data, dtype=np.uint8)
        stride = length if stride is None else stride
        stop = n if stop is None else stop
        m = stop - start
        hits = np.zeros(m, dtype=HIT_DTYPE)
        cap = max(16 * m, 1024)
        cnt = np.zeros(self.n_eq + 2, dtype=np.uint64) if counts else None
        base = data.ctypes.data + start * stride
        while True:
            tx = np.zeros(cap, dtype=np.uint32)
            used = C.c_uint64()
            if cnt is not None:
                cnt[:] = 0
            r = lib().orc_map_ascii_batch(self.h, C.c_void_p(base), stride, length, m, int(allowed), _ptr(hits), _ptr(tx), cap,
                                          C.byref(used), _ptr(cnt) if cnt is not None else None, None)
            if r == 0:
                break
            if r == -2:
                raise MemoryError
            cap = int(used.value) + 16
        return hits, tx[:used.value].copy(), cnt

    def process_reads_c3(self, fastq_path, out_path, num_threads=2, allowed=2):
        """The reference's map driver restated (c3_driver.c): -> (reads, mapped); lines in arrival order."""
        reads, mapped = C.c_uint64(), C.c_uint64()
        rc = lib().orc_process_reads_c3(self.h, os.fsencode(fastq_path), os.fsencode(out_path), int(num_threads), int(allowed),
                                        C.byref(reads), C.byref(mapped))
        if rc:
            raise RuntimeError("orc_process_reads_c3: error %d" % rc)
        return int(reads.value), int(mapped.value)

    def close(self):
        if self.h:
            lib().orc_index_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def intersect(v1, v2):
    a = np.array(v1, dtype=np.uint32)
    b = np.array(v2, dtype=np.uint32)
    n = lib().orc_intersect(_ptr(a), len(a), _ptr(b), len(b))
    return a[:n].tolist()


def result_checksum(hits, tx, first_index=0):
    hits = np.ascontiguousarray(hits)
    tx = np.ascontiguousarray(tx, dtype=np.uint32)
    return int(lib().orc_result_checksum(_ptr(hits), _ptr(tx), len(hits), int(first_index)))


def hits_to_tuples(hits, tx):
    """[(aligned, mapped_flag, tuple(tx ids), coverage)] for exact comparisons."""
    out = []
    for h in hits:
        o, n = int(h["tx_off"]), int(h["n_tx"])
        out.append((bool(h["flags"] & FLAG_ALIGNED), bool(h["flags"] & FLAG_MAPPED), tuple(tx[o:o + n].tolist()),
                    int(h["coverage"])))
    return out


def mappability(flat, tx_gene, bins=11):
    """CPU restatement of mappability::analyze_graph, ref src/mappability.rs:120-156 (records' add_tx_count / add_gene_count
    :59-73): -> (tx_multiplicity, gene_multiplicity), uint64 [n_tx, bins]."""
    k = int(flat["k"])
    n_tx = len(tx_gene)
    tm = np.zeros((n_tx, bins), np.uint64)
    gm = np.zeros((n_tx, bins), np.uint64)
    off, mem = flat["eq_offsets"], flat["eq_members"]
    for length, eq in zip(flat["node_len"], flat["node_eq"]):          # :127 for node in index.dbg.iter_nodes()
        num_kmer = int(length) - k + 1                                  # :128
        members = mem[int(off[eq]):int(off[eq + 1])]                    # :130-131
        num_tx = len(members)                                           # :133
        if num_tx == 0:
            continue
        num_genes = len(set(int(tx_gene[t]) for t in members))          # :135-143
        tb = bins - 1 if num_tx > bins else num_tx - 1                  # :59-65
        gb = bins - 1 if num_genes > bins else num_genes - 1            # :67-73
        for t in members:                                               # :145-149
            tm[int(t), tb] += np.uint64(num_kmer)
            gm[int(t), gb] += np.uint64(num_kmer)
    return tm, gm
