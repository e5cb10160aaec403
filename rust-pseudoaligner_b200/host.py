"""ctypes binding of libpsa_host.so (include/psa_host.h): host graph builder, synthetic
workloads, FASTA/FASTQ readers.  These are the parts the reference keeps on the host
(ref src/build_index.rs, src/utils.rs); none of them maps reads."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def lib_path():
    return os.path.join(_HERE, "libpsa_host.so")


def lib():
    global _lib
    if _lib is not None:
        return _lib
    p = lib_path()
    if not os.path.exists(p):
        raise ImportError("%s is missing: run __graft_entry__.build() (make -C rust-pseudoaligner_b200/csrc)" % p)
    L = C.CDLL(p)
    vp, u32, u64, i32 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int
    L.psa_host_last_error.restype = C.c_char_p
    L.psa_build_graph.restype, L.psa_build_graph.argtypes = vp, [vp, vp, u32, u32, i32]
    L.psa_graph_free.argtypes = [vp]
    L.psa_graph_k.restype, L.psa_graph_k.argtypes = u32, [vp]
    for n in ("n_nodes", "n_kmers", "n_eq", "n_seq_words", "n_cycles"):
        f = getattr(L, "psa_graph_" + n)
        f.restype, f.argtypes = u64, [vp]
    for n in ("seq_words", "node_start", "node_len", "node_exts", "node_eq", "eq_offsets", "eq_members"):
        f = getattr(L, "psa_graph_" + n)
        f.restype, f.argtypes = vp, [vp]
    L.psa_graph_from_arrays.restype = vp
    L.psa_graph_from_arrays.argtypes = [u32, u64, vp, u64, vp, vp, vp, vp, u64, vp, vp]
    L.psa_graph_save.restype, L.psa_graph_save.argtypes = i32, [vp, C.c_char_p]
    L.psa_graph_load.restype, L.psa_graph_load.argtypes = vp, [C.c_char_p]
    L.psa_graph_load_bincode.restype, L.psa_graph_load_bincode.argtypes = vp, [C.c_char_p, u32]
    L.psa_synth_transcriptome.restype, L.psa_synth_transcriptome.argtypes = vp, [u64, u32, i32]
    L.psa_transcriptome_from_codes.restype, L.psa_transcriptome_from_codes.argtypes = vp, [vp, vp, u32]
    L.psa_transcriptome_free.argtypes = [vp]
    L.psa_transcriptome_n_tx.restype, L.psa_transcriptome_n_tx.argtypes = u32, [vp]
    L.psa_transcriptome_n_bases.restype, L.psa_transcriptome_n_bases.argtypes = u64, [vp]
    L.psa_transcriptome_codes.restype, L.psa_transcriptome_codes.argtypes = vp, [vp]
    L.psa_transcriptome_tx_off.restype, L.psa_transcriptome_tx_off.argtypes = vp, [vp]
    L.psa_synth_reads.restype = i32
    L.psa_synth_reads.argtypes = [vp, u64, u64, u64, u32, vp, u64, vp, i32]
    for n in ("psa_fasta_read", "psa_fastq_read"):
        f = getattr(L, n)
        f.restype, f.argtypes = vp, [C.c_char_p]
    L.psa_seqfile_free.argtypes = [vp]
    L.psa_seqfile_n.restype, L.psa_seqfile_n.argtypes = u64, [vp]
    L.psa_seqfile_name.restype, L.psa_seqfile_name.argtypes = C.c_char_p, [vp, u64]
    L.psa_seqfile_data.restype, L.psa_seqfile_data.argtypes = vp, [vp]
    L.psa_seqfile_off.restype, L.psa_seqfile_off.argtypes = vp, [vp]
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


def _err():
    return lib().psa_host_last_error().decode()


_CODE = np.zeros(256, dtype=np.uint8)
for _c, _v in (("A", 0), ("C", 1), ("G", 2), ("T", 3)):
    _CODE[ord(_c)] = _v
    _CODE[ord(_c.lower())] = _v


def encode_transcripts(seqs):
    """list of ASCII transcripts -> (codes uint8, tx_off uint64[n+1]).  ACGT only: a reference
    FASTA with other letters needs debruijn's hashed-N substitution (ref src/utils.rs:76), which
    is outside this path."""
    lens = np.array([len(s) for s in seqs], dtype=np.uint64)
    off = np.zeros(len(seqs) + 1, dtype=np.uint64)
    off[1:] = np.cumsum(lens)
    joined = b"".join(s if isinstance(s, (bytes, bytearray)) else s.encode() for s in seqs)
    return np.ascontiguousarray(_CODE[np.frombuffer(joined, dtype=np.uint8)]), off


def _flat_of_graph(g):
    """psa_graph handle -> (flat index dict in the form of psa_index_desc, stats); frees the handle."""
    L = lib()
    k = int(L.psa_graph_k(g))
    try:
        n_nodes, n_eq, n_words = L.psa_graph_n_nodes(g), L.psa_graph_n_eq(g), L.psa_graph_n_seq_words(g)
        eq_offsets = _view(L.psa_graph_eq_offsets(g), n_eq + 1, np.uint64).copy()
        flat = {
            "k": k,
            "seq_words": _view(L.psa_graph_seq_words(g), n_words, np.uint64).copy(),
            "node_start": _view(L.psa_graph_node_start(g), n_nodes, np.uint64).copy(),
            "node_len": _view(L.psa_graph_node_len(g), n_nodes, np.uint32).copy(),
            "node_exts": _view(L.psa_graph_node_exts(g), n_nodes, np.uint8).copy(),
            "node_eq": _view(L.psa_graph_node_eq(g), n_nodes, np.uint32).copy(),
            "eq_offsets": eq_offsets,
            "eq_members": _view(L.psa_graph_eq_members(g), int(eq_offsets[-1]), np.uint32).copy(),
        }
        stats = {"n_kmers": int(L.psa_graph_n_kmers(g)), "n_cycles": int(L.psa_graph_n_cycles(g)),
                 "n_nodes": int(n_nodes), "n_eq": int(n_eq)}
    finally:
        L.psa_graph_free(g)
    return flat, stats


def build_graph(codes, tx_off, k, threads=0, copy=True):
    """Host graph builder -> flat index dict in the form of psa_index_desc."""
    codes = np.ascontiguousarray(codes, dtype=np.uint8)
    tx_off = np.ascontiguousarray(tx_off, dtype=np.uint64)
    g = lib().psa_build_graph(_ptr(codes), _ptr(tx_off), len(tx_off) - 1, int(k), int(threads))
    if not g:
        raise RuntimeError("psa_build_graph: " + _err())
    return _flat_of_graph(g)


def save_index(flat, path):
    """Write a flat index (the dict build_graph returns) as one file (psa_graph_save)."""
    f = {k_: (np.ascontiguousarray(v) if isinstance(v, np.ndarray) else v) for k_, v in flat.items()}
    L = lib()
    g = L.psa_graph_from_arrays(int(f["k"]), len(f["node_len"]), _ptr(f["seq_words"]), len(f["seq_words"]),
                                _ptr(f["node_start"]), _ptr(f["node_len"]), _ptr(f["node_exts"]), _ptr(f["node_eq"]),
                                len(f["eq_offsets"]) - 1, _ptr(f["eq_offsets"]), _ptr(f["eq_members"]))
    if not g:
        raise RuntimeError("psa_graph_from_arrays: " + _err())
    try:
        if L.psa_graph_save(g, os.fsencode(path)) != 0:
            raise RuntimeError("psa_graph_save: " + _err())
    finally:
        L.psa_graph_free(g)


def load_index(path):
    """Read a file written by save_index -> (flat, stats).  Raises on a foreign, truncated or corrupt file."""
    g = lib().psa_graph_load(os.fsencode(path))
    if not g:
        raise RuntimeError("psa_graph_load: " + _err())
    return _flat_of_graph(g)


def load_reference_index(path, k):
    """Read `dbg` + `eq_classes` of a reference-built index file (bincode of Pseudoaligner<K>) -> (flat, stats).
    Layout as recalled for debruijn 0.3.4 (see include/psa_host.h); refuses anything inconsistent."""
    g = lib().psa_graph_load_bincode(os.fsencode(path), int(k))
    if not g:
        raise RuntimeError("psa_graph_load_bincode: " + _err())
    return _flat_of_graph(g)


class Transcriptome:
    def __init__(self, handle):
        if not handle:
            raise RuntimeError("transcriptome: " + _err())
        self.h = C.c_void_p(handle)
        L = lib()
        self.n_tx = int(L.psa_transcriptome_n_tx(self.h))
        self.n_bases = int(L.psa_transcriptome_n_bases(self.h))

    @classmethod
    def synth(cls, seed, n_genes, threads=0):
        return cls(lib().psa_synth_transcriptome(int(seed), int(n_genes), int(threads)))

    @classmethod
    def from_codes(cls, codes, tx_off):
        codes = np.ascontiguousarray(codes, dtype=np.uint8)
        tx_off = np.ascontiguousarray(tx_off, dtype=np.uint64)
        return cls(lib().psa_transcriptome_from_codes(_ptr(codes), _ptr(tx_off), len(tx_off) - 1))

    def codes(self):
        return _view(lib().psa_transcriptome_codes(self.h), self.n_bases, np.uint8)

    def tx_off(self):
        return _view(lib().psa_transcriptome_tx_off(self.h), self.n_tx + 1, np.uint64)

    def reads(self, seed, first, n, length, out=None, stride=None, kinds=False, threads=0):
        """ASCII reads first..first+n-1 of stream `seed` -> uint8 [n, stride] (a view of `out`)."""
        stride = length if stride is None else stride
        if out is None:
            out = np.zeros(n * stride + 1, dtype=np.uint8)
        kind = np.zeros(n, np.uint8) if kinds else None
        rc = lib().psa_synth_reads(self.h, int(seed), int(first), int(n), int(length), _ptr(out), int(stride),
                                   _ptr(kind) if kinds else None, int(threads))
        if rc:
            raise RuntimeError("psa_synth_reads: " + _err())
        return (out, kind) if kinds else out

    def close(self):
        if self.h:
            lib().psa_transcriptome_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _seqfile(h):
    if not h:
        raise RuntimeError(_err())
    L = lib()
    try:
        n = int(L.psa_seqfile_n(h))
        off = _view(L.psa_seqfile_off(h), n + 1, np.uint64).copy()
        data = _view(L.psa_seqfile_data(h), int(off[-1]), np.uint8).copy()
        names = [L.psa_seqfile_name(h, i).decode() for i in range(n)]
    finally:
        L.psa_seqfile_free(h)
    return names, data, off


def read_fasta(path):
    """-> (names, ascii uint8 data, off[n+1]); record order = transcript index (ref src/utils.rs:71-88)."""
    return _seqfile(lib().psa_fasta_read(str(path).encode()))


def read_fastq(path):
    """-> (ids, ascii uint8 data, off[n+1])."""
    return _seqfile(lib().psa_fastq_read(str(path).encode()))
