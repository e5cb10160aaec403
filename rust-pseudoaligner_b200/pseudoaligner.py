"""ctypes binding of libpsa_b200.so and the host-side mirror of the reference's API for the path.

Names follow the crate: `Pseudoaligner` (ref src/pseudoaligner.rs:26-33) holds the flattened
index on the GPU, `map_read` (ref :381) and `process_reads` (ref :420-514) behave as the
reference's do, batch-at-a-time.  No torch types, no CPU fallback: if the shared library is
missing or there is no CUDA device the calls raise.
"""
import ctypes as C
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

EQ_NONE = 0xFFFFFFFF
FLAG_ALIGNED = 1
FLAG_MAPPED = 2
READS_ASCII, READS_PACKED = 0, 1
MEM_HOST, MEM_DEVICE = 0, 1
ERR_CAPACITY = -4

HIT_DTYPE = np.dtype(
    [("coverage", "<u4"), ("n_tx", "<u4"), ("tx_off", "<u8"), ("eq_id", "<u4"), ("flags", "<u4")]
)

HIT_COMPACT_DTYPE = np.dtype([("eq_or_n", "<u4"), ("cov_flags", "<u4")])
RESULT_COMPACT = 1

EVENT_FIELDS = ("reads", "read_bases", "kmer_lookups", "dict_levels", "dict_hits", "verifications",
                "node_visits", "bases_compared", "edge_jumps", "class_members", "out_members", "aligned")


class PsaError(RuntimeError):
    def __init__(self, code, text):
        super().__init__("psa error %d: %s" % (code, text))
        self.code = code


class _IndexDesc(C.Structure):
    _fields_ = [("k", C.c_uint32), ("reserved", C.c_uint32), ("n_nodes", C.c_uint64),
                ("seq_words", C.c_void_p), ("n_seq_words", C.c_uint64), ("node_start", C.c_void_p),
                ("node_len", C.c_void_p), ("node_exts", C.c_void_p), ("node_eq", C.c_void_p),
                ("n_eq", C.c_uint64), ("eq_offsets", C.c_void_p), ("eq_members", C.c_void_p)]


class _IndexInfo(C.Structure):
    _fields_ = [("k", C.c_uint32), ("dict_levels", C.c_uint32),
                ("n_nodes", C.c_uint64), ("n_kmers", C.c_uint64), ("n_eq", C.c_uint64),
                ("n_eq_members", C.c_uint64), ("n_seq_words", C.c_uint64),
                ("dict_bytes", C.c_uint64), ("node_bytes", C.c_uint64),
                ("seq_bytes", C.c_uint64), ("eq_bytes", C.c_uint64),
                ("node_bits", C.c_uint32), ("pos_bits", C.c_uint32), ("fp_bits", C.c_uint32),
                ("max_class_len", C.c_uint32), ("gamma", C.c_double), ("build_ms", C.c_double)]


class _ReadBatch(C.Structure):
    _fields_ = [("format", C.c_uint32), ("location", C.c_uint32), ("data", C.c_void_p),
                ("data_len", C.c_uint64), ("read_off", C.c_void_p), ("read_len", C.c_void_p),
                ("stride", C.c_uint64), ("fixed_len", C.c_uint32), ("reserved", C.c_uint32),
                ("n_reads", C.c_uint64)]


class _ResultBatch(C.Structure):
    _fields_ = [("location", C.c_uint32), ("flags", C.c_uint32), ("hits", C.c_void_p),
                ("tx_buf", C.c_void_p), ("tx_cap", C.c_uint64), ("tx_used", C.c_uint64)]


class _ProcessStats(C.Structure):
    _fields_ = [("reads", C.c_uint64), ("mapped", C.c_uint64), ("aligned", C.c_uint64), ("seconds", C.c_double),
                ("reader_seconds", C.c_double), ("mapper_seconds", C.c_double), ("writer_seconds", C.c_double)]


class _NovelSets(C.Structure):
    _fields_ = [("n_sets", C.c_uint64), ("n_members", C.c_uint64), ("offsets", C.POINTER(C.c_uint64)),
                ("members", C.POINTER(C.c_uint32)), ("counts", C.POINTER(C.c_uint64))]


class _Events(C.Structure):
    _fields_ = [(f, C.c_uint64) for f in EVENT_FIELDS]


EXPORTS = (
    "psa_strerror", "psa_last_error", "psa_abi_version",
    "psa_index_create", "psa_index_destroy", "psa_index_get_info", "psa_index_lookup",
    "psa_mapper_create", "psa_mapper_destroy", "psa_mapper_set_allowed_mismatches", "psa_mapper_set_group_width",
    "psa_mapper_set_fast_path", "psa_mapper_set_scan_width",
    "psa_mapper_map", "psa_mapper_map_async", "psa_mapper_sync", "psa_mapper_stream",
    "psa_mapper_map_read", "psa_mapper_counts_get", "psa_mapper_counts_reset",
    "psa_mapper_counts_device", "psa_mapper_map_events", "psa_mapper_defer_reasons", "psa_mapper_launch_count",
    "psa_mapper_profile_enable", "psa_mapper_profile_read",
    "psa_comm_unique_id", "psa_comm_create", "psa_comm_destroy", "psa_mapper_counts_allreduce",
    "psa_host_alloc", "psa_host_free", "psa_device_alloc", "psa_device_free",
    "psa_memcpy_h2d", "psa_memcpy_d2h", "psa_process_reads", "psa_gather_probe", "psa_result_checksum",
    "psa_selftest_intersect", "psa_mapper_novel_sets", "psa_novel_sets_merge", "psa_novel_sets_free",
    "psa_mapper_novel_allgather", "psa_expand_compact", "psa_debug_str", "psa_index_host_classes",
    "psa_synth_reads_device", "psa_index_mappability", "psa_build_graph_device", "psa_built_graph_free",
)


def lib_path():
    # PSA_LIB_PATH: an alternative build of the same library (kernel tuning experiments)
    return os.environ.get("PSA_LIB_PATH") or os.path.join(_HERE, "libpsa_b200.so")


_lib = None


def lib():
    """Load libpsa_b200.so.  Raises if it has not been built: there is nothing to fall back to."""
    global _lib
    if _lib is not None:
        return _lib
    p = lib_path()
    if not os.path.exists(p):
        raise ImportError("%s is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(make -C rust-pseudoaligner_b200/csrc). The CUDA library is the only implementation." % p)
    L = C.CDLL(p)
    vp, u32, u64, i32 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int
    L.psa_strerror.restype, L.psa_strerror.argtypes = C.c_char_p, [i32]
    L.psa_last_error.restype, L.psa_last_error.argtypes = C.c_char_p, []
    L.psa_abi_version.restype = i32
    L.psa_index_create.restype = i32
    L.psa_index_create.argtypes = [C.POINTER(_IndexDesc), i32, C.c_double, C.POINTER(vp)]
    L.psa_index_destroy.restype, L.psa_index_destroy.argtypes = None, [vp]
    L.psa_index_get_info.restype, L.psa_index_get_info.argtypes = i32, [vp, C.POINTER(_IndexInfo)]
    L.psa_index_lookup.restype, L.psa_index_lookup.argtypes = i32, [vp, vp, u64, vp, vp, vp]
    L.psa_mapper_create.restype, L.psa_mapper_create.argtypes = i32, [vp, u64, C.POINTER(vp)]
    L.psa_mapper_destroy.restype, L.psa_mapper_destroy.argtypes = None, [vp]
    L.psa_mapper_set_allowed_mismatches.restype = i32
    L.psa_mapper_set_allowed_mismatches.argtypes = [vp, u32]
    L.psa_mapper_set_group_width.restype, L.psa_mapper_set_group_width.argtypes = i32, [vp, u32]
    L.psa_mapper_set_fast_path.restype, L.psa_mapper_set_fast_path.argtypes = i32, [vp, u32, u32]
    L.psa_mapper_set_scan_width.restype, L.psa_mapper_set_scan_width.argtypes = i32, [vp, u32]
    for name in ("psa_mapper_map", "psa_mapper_map_async"):
        f = getattr(L, name)
        f.restype, f.argtypes = i32, [vp, C.POINTER(_ReadBatch), C.POINTER(_ResultBatch)]
    L.psa_mapper_map_events.restype = i32
    L.psa_mapper_map_events.argtypes = [vp, C.POINTER(_ReadBatch), C.POINTER(_ResultBatch), C.POINTER(_Events * 3)]
    L.psa_mapper_defer_reasons.restype, L.psa_mapper_defer_reasons.argtypes = i32, [vp, C.POINTER(u64 * 4)]
    L.psa_mapper_sync.restype, L.psa_mapper_sync.argtypes = i32, [vp]
    L.psa_mapper_stream.restype, L.psa_mapper_stream.argtypes = vp, [vp]
    L.psa_mapper_map_read.restype = i32
    L.psa_mapper_map_read.argtypes = [vp, vp, u32, vp, u64, C.POINTER(u32), C.POINTER(u32)]
    L.psa_mapper_counts_get.restype, L.psa_mapper_counts_get.argtypes = i32, [vp, vp]
    L.psa_mapper_counts_reset.restype, L.psa_mapper_counts_reset.argtypes = i32, [vp]
    L.psa_mapper_counts_device.restype, L.psa_mapper_counts_device.argtypes = vp, [vp]
    L.psa_mapper_launch_count.restype, L.psa_mapper_launch_count.argtypes = u64, [vp]
    L.psa_mapper_profile_enable.restype, L.psa_mapper_profile_enable.argtypes = i32, [vp, i32]
    L.psa_mapper_profile_read.restype = i32
    L.psa_mapper_profile_read.argtypes = [vp, C.POINTER(C.c_double * 3), C.POINTER(u64 * 3)]
    L.psa_comm_unique_id.restype, L.psa_comm_unique_id.argtypes = i32, [vp]
    L.psa_comm_create.restype, L.psa_comm_create.argtypes = i32, [vp, i32, i32, i32, C.POINTER(vp)]
    L.psa_comm_destroy.restype, L.psa_comm_destroy.argtypes = None, [vp]
    L.psa_mapper_counts_allreduce.restype, L.psa_mapper_counts_allreduce.argtypes = i32, [vp, vp]
    L.psa_host_alloc.restype, L.psa_host_alloc.argtypes = i32, [C.POINTER(vp), u64]
    L.psa_host_free.restype, L.psa_host_free.argtypes = None, [vp]
    L.psa_device_alloc.restype, L.psa_device_alloc.argtypes = i32, [C.POINTER(vp), u64]
    L.psa_device_free.restype, L.psa_device_free.argtypes = None, [vp]
    L.psa_memcpy_h2d.restype, L.psa_memcpy_h2d.argtypes = i32, [vp, vp, u64]
    L.psa_memcpy_d2h.restype, L.psa_memcpy_d2h.argtypes = i32, [vp, vp, u64]
    L.psa_gather_probe.restype = i32
    L.psa_gather_probe.argtypes = [i32, u64, u32, u32, C.POINTER(C.c_double)]
    L.psa_mapper_novel_sets.restype, L.psa_mapper_novel_sets.argtypes = i32, [vp, C.POINTER(_NovelSets)]
    L.psa_novel_sets_merge.restype, L.psa_novel_sets_merge.argtypes = i32, [C.POINTER(_NovelSets), u32, C.POINTER(_NovelSets)]
    L.psa_novel_sets_free.restype, L.psa_novel_sets_free.argtypes = None, [C.POINTER(_NovelSets)]
    L.psa_mapper_novel_allgather.restype, L.psa_mapper_novel_allgather.argtypes = i32, [vp, vp, C.POINTER(_NovelSets)]
    L.psa_debug_str.restype, L.psa_debug_str.argtypes = C.c_int64, [C.c_char_p, u64, C.c_char_p, u64]
    L.psa_expand_compact.restype = i32
    L.psa_expand_compact.argtypes = [vp, u64, vp, u64, vp, vp, u64, vp, vp, u64, C.POINTER(u64)]
    L.psa_selftest_intersect.restype = i32
    L.psa_selftest_intersect.argtypes = [i32, vp, u32, vp, u32, vp, u32, C.POINTER(u32 * 3)]
    L.psa_build_graph_device.restype = i32
    L.psa_build_graph_device.argtypes = [i32, vp, vp, u32, u32, vp]
    L.psa_built_graph_free.restype, L.psa_built_graph_free.argtypes = None, [vp]
    L.psa_index_mappability.restype = i32
    L.psa_index_mappability.argtypes = [vp, vp, u32, u32, vp, vp]
    L.psa_synth_reads_device.restype = i32
    L.psa_synth_reads_device.argtypes = [i32, vp, u64, u64, u64, u32, vp, u64]
    L.psa_result_checksum.restype = i32
    L.psa_result_checksum.argtypes = [i32, vp, vp, u64, u64, C.POINTER(u64)]
    L.psa_process_reads.restype = i32
    L.psa_process_reads.argtypes = [vp, C.c_char_p, C.c_char_p, u32, u64, i32, C.POINTER(_ProcessStats)]
    _lib = L
    return L


def _novel_to_py(ns):
    """psa_novel_sets -> [(tuple(members), count)] in id order (id = n_eq + position)."""
    n = int(ns.n_sets)
    off = np.ctypeslib.as_array(ns.offsets, shape=(n + 1,)).copy() if n else np.zeros(1, np.uint64)
    mem = np.ctypeslib.as_array(ns.members, shape=(max(int(ns.n_members), 1),)).copy()
    cnt = np.ctypeslib.as_array(ns.counts, shape=(n,)).copy() if n else np.zeros(0, np.uint64)
    return [(tuple(int(x) for x in mem[int(off[i]):int(off[i + 1])]), int(cnt[i])) for i in range(n)]


def novel_sets_merge(tables):
    """psa_novel_sets_merge over tables given as [(members tuple, count)] lists -> the merged, sorted table."""
    parts = (_NovelSets * max(len(tables), 1))()
    keep = []
    for p, t in zip(parts, tables):
        off = np.zeros(len(t) + 1, np.uint64)
        off[1:] = np.cumsum([len(m) for m, _ in t])
        mem = np.array([x for m, _ in t for x in m] + [0], dtype=np.uint32)
        cnt = np.array([c for _, c in t] + [0], dtype=np.uint64)
        keep.append((off, mem, cnt))
        p.n_sets, p.n_members = len(t), int(off[-1])
        p.offsets = off.ctypes.data_as(C.POINTER(C.c_uint64))
        p.members = mem.ctypes.data_as(C.POINTER(C.c_uint32))
        p.counts = cnt.ctypes.data_as(C.POINTER(C.c_uint64))
    out = _NovelSets()
    _check(lib().psa_novel_sets_merge(parts, len(tables), C.byref(out)))
    res = _novel_to_py(out)
    lib().psa_novel_sets_free(C.byref(out))
    return res


def selftest_intersect(v1, v2, device=0):
    """The three device intersection routines on two ascending lists -> (thread lists, lane group, windows | None)."""
    a = np.ascontiguousarray(v1, dtype=np.uint32)
    b = np.ascontiguousarray(v2, dtype=np.uint32)
    cap = max(len(a), len(b), 1)
    out = np.zeros(3 * cap, np.uint32)
    n = (C.c_uint32 * 3)()
    _check(lib().psa_selftest_intersect(int(device), _ptr(a), len(a), _ptr(b), len(b), _ptr(out), cap, C.byref(n)))
    res = [out[i * cap:i * cap + n[i]].tolist() for i in range(2)]
    res.append(None if n[2] == EQ_NONE else out[2 * cap:2 * cap + n[2]].tolist())
    return res


def gather_probe(device=0, table_bytes=1 << 30, chunk_bytes=32, iters=64):
    """GB/s of independent random chunk gathers from HBM (psa_gather_probe)."""
    out = C.c_double()
    _check(lib().psa_gather_probe(int(device), int(table_bytes), int(chunk_bytes), int(iters), C.byref(out)))
    return float(out.value)


def _check(rc):
    if rc < 0:
        raise PsaError(rc, lib().psa_last_error().decode() or lib().psa_strerror(rc).decode())
    return rc


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


# ------------------------------------------------------------------------------------ pinned / device memory
class PinnedArray:
    """numpy view over cudaHostAlloc memory (psa_host_alloc)."""

    def __init__(self, shape, dtype):
        self.dtype = np.dtype(dtype)
        self.shape = tuple(shape) if isinstance(shape, (tuple, list)) else (int(shape),)
        nbytes = int(np.prod(self.shape)) * self.dtype.itemsize
        p = C.c_void_p()
        _check(lib().psa_host_alloc(C.byref(p), max(nbytes, 1)))
        self.ptr = p.value
        buf = (C.c_char * max(nbytes, 1)).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=self.dtype, count=int(np.prod(self.shape))).reshape(self.shape)

    def free(self):
        if self.ptr:
            self.array = None
            lib().psa_host_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class DeviceBuffer:
    """Raw HBM allocation (psa_device_alloc) with numpy upload / download."""

    def __init__(self, nbytes):
        p = C.c_void_p()
        _check(lib().psa_device_alloc(C.byref(p), max(int(nbytes), 1)))
        self.ptr = p.value
        self.nbytes = int(nbytes)

    @classmethod
    def from_numpy(cls, a):
        a = np.ascontiguousarray(a)
        b = cls(a.nbytes)
        if a.nbytes:
            _check(lib().psa_memcpy_h2d(b.ptr, _ptr(a), a.nbytes))
        return b

    def to_numpy(self, dtype, count):
        out = np.empty(count, dtype=dtype)
        if out.nbytes:
            _check(lib().psa_memcpy_d2h(_ptr(out), self.ptr, out.nbytes))
        return out

    def free(self):
        if self.ptr:
            lib().psa_device_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


# ------------------------------------------------------------------------------------ index
class Index:
    """Device-resident flattened `Pseudoaligner<K>` (psa_index)."""

    def __init__(self, flat, device=0, gamma=0.0):
        f = {k_: (np.ascontiguousarray(v) if isinstance(v, np.ndarray) else v) for k_, v in flat.items()}
        assert f["seq_words"].dtype == np.uint64 and f["node_start"].dtype == np.uint64
        assert f["node_len"].dtype == np.uint32 and f["node_exts"].dtype == np.uint8
        assert f["node_eq"].dtype == np.uint32 and f["eq_offsets"].dtype == np.uint64
        assert f["eq_members"].dtype == np.uint32
        d = _IndexDesc()
        d.k = int(f["k"])
        d.n_nodes = len(f["node_len"])
        d.seq_words, d.n_seq_words = _ptr(f["seq_words"]), len(f["seq_words"])
        d.node_start, d.node_len = _ptr(f["node_start"]), _ptr(f["node_len"])
        d.node_exts, d.node_eq = _ptr(f["node_exts"]), _ptr(f["node_eq"])
        d.n_eq = len(f["eq_offsets"]) - 1
        d.eq_offsets, d.eq_members = _ptr(f["eq_offsets"]), _ptr(f["eq_members"])
        h = C.c_void_p()
        _check(lib().psa_index_create(C.byref(d), int(device), float(gamma), C.byref(h)))
        self.h = h
        self.k = d.k
        self.n_eq = int(d.n_eq)
        self.device = int(device)

    def info(self):
        i = _IndexInfo()
        _check(lib().psa_index_get_info(self.h, C.byref(i)))
        return {n: getattr(i, n) for n, _ in _IndexInfo._fields_}

    def lookup(self, kmer_words):
        """kmer_words: uint64 [n] (k<=32) or [n,2]; -> (found, node, off)."""
        w = np.ascontiguousarray(kmer_words, dtype=np.uint64)
        n = w.shape[0]
        found = np.zeros(n, np.uint8)
        node = np.zeros(n, np.uint32)
        off = np.zeros(n, np.uint32)
        _check(lib().psa_index_lookup(self.h, _ptr(w), n, _ptr(found), _ptr(node), _ptr(off)))
        return found.astype(bool), node, off

    def mappability(self, tx_gene, bins=11):
        """mappability::analyze_graph (ref src/mappability.rs:120-156) on the device -> (tx_multiplicity, gene_multiplicity),
        uint64 [n_tx, bins] each.  tx_gene[t]: an integer naming transcript t's gene."""
        tx_gene = np.ascontiguousarray(tx_gene, np.uint32)
        n_tx = len(tx_gene)
        tm = np.zeros((n_tx, bins), np.uint64)
        gm = np.zeros((n_tx, bins), np.uint64)
        _check(lib().psa_index_mappability(self.h, _ptr(tx_gene), n_tx, int(bins), _ptr(tm), _ptr(gm)))
        return tm, gm

    def close(self):
        if getattr(self, "h", None):
            lib().psa_index_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DeviceBatch:
    """A read batch resident in HBM plus its result buffers (for kernel-only timing)."""

    def __init__(self, fmt, data, n_reads, read_off=None, read_len=None, stride=0, fixed_len=0, tx_cap=0):
        self.fmt = fmt
        self.n = int(n_reads)
        data = np.ascontiguousarray(data)
        self.data = DeviceBuffer.from_numpy(data)
        self.data_len = data.size
        self.read_off = DeviceBuffer.from_numpy(np.ascontiguousarray(read_off, np.uint64)) if read_off is not None else None
        self.read_len = DeviceBuffer.from_numpy(np.ascontiguousarray(read_len, np.uint32)) if read_len is not None else None
        self.stride, self.fixed_len = int(stride), int(fixed_len)
        self.hits = DeviceBuffer(self.n * HIT_DTYPE.itemsize)
        self.tx_cap = int(tx_cap)
        self.tx = DeviceBuffer(self.tx_cap * 4) if tx_cap else None
        self.rb = _ReadBatch(fmt, MEM_DEVICE, self.data.ptr, self.data_len,
                             self.read_off.ptr if self.read_off else None,
                             self.read_len.ptr if self.read_len else None,
                             self.stride, self.fixed_len, 0, self.n)
        self.ob = _ResultBatch(MEM_DEVICE, 0, self.hits.ptr, self.tx.ptr if self.tx else None, self.tx_cap, 0)

    def download(self):
        hits = self.hits.to_numpy(HIT_DTYPE, self.n)
        used = int(self.ob.tx_used)
        tx = self.tx.to_numpy(np.uint32, min(used, self.tx_cap)) if self.tx else np.zeros(0, np.uint32)
        return hits, tx

    def checksum(self, first_index=0, device=0):
        """Order-independent checksum of the batch's results (psa_result_checksum)."""
        out = C.c_uint64()
        _check(lib().psa_result_checksum(int(device), self.hits.ptr, self.tx.ptr, self.n, int(first_index), C.byref(out)))
        return int(out.value)

    def free(self):
        for b in (self.data, self.read_off, self.read_len, self.hits, self.tx):
            if b is not None:
                b.free()


class _BuiltGraph(C.Structure):
    _fields_ = [("k", C.c_uint32), ("reserved", C.c_uint32), ("n_nodes", C.c_uint64), ("n_kmers", C.c_uint64), ("n_eq", C.c_uint64),
                ("n_seq_words", C.c_uint64), ("n_eq_members", C.c_uint64), ("n_cycles", C.c_uint64),
                ("seq_words", C.POINTER(C.c_uint64)), ("node_start", C.POINTER(C.c_uint64)), ("node_len", C.POINTER(C.c_uint32)),
                ("node_exts", C.POINTER(C.c_uint8)), ("node_eq", C.POINTER(C.c_uint32)), ("eq_offsets", C.POINTER(C.c_uint64)),
                ("eq_members", C.POINTER(C.c_uint32))]


def build_graph_device(codes, tx_off, k, device=0):
    """psa_build_graph_device: the coloured compacted de Bruijn graph built on the GPU -> (flat index dict in the form of
    psa_index_desc, stats), the same arrays host.build_graph returns."""
    codes = np.ascontiguousarray(codes, dtype=np.uint8)
    tx_off = np.ascontiguousarray(tx_off, dtype=np.uint64)
    g = _BuiltGraph()
    _check(lib().psa_build_graph_device(int(device), _ptr(codes), _ptr(tx_off), len(tx_off) - 1, int(k), C.byref(g)))
    try:
        def arr(p, n, dt):
            return np.ctypeslib.as_array(p, shape=(int(n),)).astype(dt, copy=True) if n else np.zeros(0, dt)
        flat = {"k": int(k), "seq_words": arr(g.seq_words, g.n_seq_words, np.uint64), "node_start": arr(g.node_start, g.n_nodes, np.uint64),
                "node_len": arr(g.node_len, g.n_nodes, np.uint32), "node_exts": arr(g.node_exts, g.n_nodes, np.uint8),
                "node_eq": arr(g.node_eq, g.n_nodes, np.uint32), "eq_offsets": arr(g.eq_offsets, g.n_eq + 1, np.uint64),
                "eq_members": arr(g.eq_members, g.n_eq_members, np.uint32)}
        stats = {"n_nodes": int(g.n_nodes), "n_kmers": int(g.n_kmers), "n_eq": int(g.n_eq), "n_cycles": int(g.n_cycles)}
    finally:
        lib().psa_built_graph_free(C.byref(g))
    return flat, stats


class _SynthTables(C.Structure):
    _fields_ = [("codes", C.c_void_p), ("tx_off", C.c_void_p), ("elig", C.c_void_p * 3), ("cum", C.c_void_p * 3),
                ("n_elig", C.c_uint64 * 3)]


class DeviceReadGenerator:
    """psa_synth_reads_device: the synthetic read stream of host.Transcriptome.reads generated in HBM (measurement aid
    for configurations too large to generate on the host).  `codes`, `tx_off`: the transcriptome (host.Transcriptome)."""

    def __init__(self, codes, tx_off, read_len, device=0):
        self.device, self.L = int(device), int(read_len)
        tx_off = np.ascontiguousarray(tx_off, np.uint64)
        lens = tx_off[1:] - tx_off[:-1]
        self.bufs = [DeviceBuffer.from_numpy(np.ascontiguousarray(codes, np.uint8)), DeviceBuffer.from_numpy(tx_off)]
        t = _SynthTables()
        t.codes, t.tx_off = self.bufs[0].ptr, self.bufs[1].ptr
        for w, lw in enumerate((self.L, self.L // 2, self.L - self.L // 2)):
            elig = np.nonzero((lens >= lw) & (lw > 0))[0].astype(np.uint32)
            cum = np.zeros(len(elig) + 1, np.uint64)
            cum[1:] = np.cumsum(lens[elig] - np.uint64(lw) + np.uint64(1))
            be, bc = DeviceBuffer.from_numpy(elig), DeviceBuffer.from_numpy(cum)
            self.bufs += [be, bc]
            t.elig[w], t.cum[w], t.n_elig[w] = be.ptr, bc.ptr, len(elig)
        self.tables = t

    def generate(self, seed, first, n, out_ptr, stride):
        _check(lib().psa_synth_reads_device(self.device, C.byref(self.tables), int(seed), int(first), int(n), self.L,
                                            C.c_void_p(out_ptr), int(stride)))

    def free(self):
        for b in self.bufs:
            b.free()
        self.bufs = []


class Comm:
    """NCCL communicator for the one collective of the path (psa_comm)."""

    @staticmethod
    def unique_id():
        b = np.zeros(128, np.uint8)
        _check(lib().psa_comm_unique_id(_ptr(b)))
        return b.tobytes()

    def __init__(self, uid, world, rank, device):
        b = np.frombuffer(uid, dtype=np.uint8).copy()
        h = C.c_void_p()
        _check(lib().psa_comm_create(_ptr(b), int(world), int(rank), int(device), C.byref(h)))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            lib().psa_comm_destroy(self.h)
            self.h = None


class Mapper:
    """psa_mapper: the per-caller state around the map kernel."""

    def __init__(self, index, chunk_reads=0):
        self.index = index
        h = C.c_void_p()
        _check(lib().psa_mapper_create(index.h, int(chunk_reads), C.byref(h)))
        self.h = h

    def set_allowed_mismatches(self, a):
        _check(lib().psa_mapper_set_allowed_mismatches(self.h, int(a)))

    def set_group_width(self, lanes):
        _check(lib().psa_mapper_set_group_width(self.h, int(lanes)))

    # ---- host batches -------------------------------------------------------------------
    def _map_host(self, fmt, data, n, read_off, read_len, stride, fixed_len, want_tx=True, tx_cap=None,
                  hits=None, tx=None, compact=False):
        rb = _ReadBatch(fmt, MEM_HOST, _ptr(data), data.size, _ptr(read_off), _ptr(read_len),
                        int(stride), int(fixed_len), 0, int(n))
        if hits is None:
            hits = np.zeros(n, dtype=HIT_COMPACT_DTYPE if compact else HIT_DTYPE)
        cap = int(tx_cap) if tx_cap is not None else max(16 * n, 1024)
        while True:
            if want_tx and (tx is None or tx.size < cap):
                tx = np.zeros(cap, dtype=np.uint32)
            ob = _ResultBatch(MEM_HOST, RESULT_COMPACT if compact else 0, _ptr(hits), _ptr(tx) if want_tx else None,
                              cap if want_tx else 0, 0)
            rc = lib().psa_mapper_map(self.h, C.byref(rb), C.byref(ob))
            if rc == ERR_CAPACITY and want_tx and ob.tx_used > cap:
                cap = int(ob.tx_used) + 16
                continue
            _check(rc)
            break
        return hits, (tx[:ob.tx_used] if want_tx else np.zeros(0, np.uint32))

    def map_ascii(self, seqs, want_tx=True, compact=False):
        """list of ASCII reads (bytes/str) -> (hits, tx_buf); the process_reads body for a batch.
        compact=True: 8-byte psa_hit_compact records, tx_buf = members of the non-class sets only."""
        bs = [s.encode() if isinstance(s, str) else bytes(s) for s in seqs]
        lens = np.array([len(b) for b in bs], dtype=np.uint32)
        off = np.zeros(len(bs), dtype=np.uint64)
        if len(bs):
            off[1:] = np.cumsum(lens.astype(np.uint64))[:-1]
        data = np.frombuffer(b"".join(bs) + b"\0", dtype=np.uint8)
        return self._map_host(READS_ASCII, data, len(bs), off, lens, 0, 0, want_tx, compact=compact)

    def map_ascii_fixed(self, data, n, length, stride=None, want_tx=True, **kw):
        """n reads of `length` bases at data[i*stride : i*stride+length] (uint8 array)."""
        stride = length if stride is None else stride
        return self._map_host(READS_ASCII, data, n, None, None, stride, length, want_tx, **kw)

    def map_packed(self, words, read_off, read_len, want_tx=True, **kw):
        words = np.ascontiguousarray(words, np.uint64)
        read_off = np.ascontiguousarray(read_off, np.uint64)
        read_len = np.ascontiguousarray(read_len, np.uint32)
        return self._map_host(READS_PACKED, words, len(read_len), read_off, read_len, 0, 0, want_tx, **kw)

    def map_packed_fixed(self, words, n, length, wstride=None, want_tx=True, **kw):
        """n reads of `length` bases as DnaString words, read i at words[i*wstride ...] (uint64 array)."""
        wstride = (length + 31) // 32 if wstride is None else wstride
        return self._map_host(READS_PACKED, words, n, None, None, wstride, length, want_tx, **kw)

    # ---- device batches -----------------------------------------------------------------
    def map_device(self, batch):
        _check(lib().psa_mapper_map(self.h, C.byref(batch.rb), C.byref(batch.ob)))

    def map_device_async(self, batch):
        _check(lib().psa_mapper_map_async(self.h, C.byref(batch.rb), C.byref(batch.ob)))

    def sync(self):
        _check(lib().psa_mapper_sync(self.h))

    def map_device_events(self, batch, split=False):
        """Event counts of one batch; split=True -> per kernel (k_map_thread, k_map, k_seed_scan)."""
        ev = (_Events * 3)()
        _check(lib().psa_mapper_map_events(self.h, C.byref(batch.rb), C.byref(batch.ob), C.byref(ev)))
        parts = [{f: int(getattr(e, f)) for f in EVENT_FIELDS} for e in ev]
        if split:
            return parts
        return {f: sum(p[f] for p in parts) for f in EVENT_FIELDS}

    def defer_reasons(self):
        out = (C.c_uint64 * 4)()
        _check(lib().psa_mapper_defer_reasons(self.h, C.byref(out)))
        return dict(zip(("first_seed_search", "reseed_search", "class_list_full", "smallest_class_long"), map(int, out)))

    def stream(self):
        return lib().psa_mapper_stream(self.h)

    # ---- single read: Pseudoaligner::map_read ------------------------------------------------
    def map_read_packed(self, words, length):
        words = np.ascontiguousarray(words, np.uint64)
        cap = 1 << 12
        while True:
            tx = np.zeros(cap, np.uint32)
            n_tx, cov = C.c_uint32(), C.c_uint32()
            rc = lib().psa_mapper_map_read(self.h, _ptr(words), int(length), _ptr(tx), cap, C.byref(n_tx), C.byref(cov))
            if rc == ERR_CAPACITY and n_tx.value > cap:
                cap = n_tx.value + 16
                continue
            _check(rc)
            break
        if rc == 0:
            return None
        return tx[:n_tx.value].tolist(), int(cov.value)

    # ---- counts ------------------------------------------------------------------------------
    def counts(self):
        out = np.zeros(self.index.n_eq + 2, dtype=np.uint64)
        _check(lib().psa_mapper_counts_get(self.h, _ptr(out)))
        return out

    def counts_reset(self):
        _check(lib().psa_mapper_counts_reset(self.h))

    def counts_allreduce(self, comm):
        _check(lib().psa_mapper_counts_allreduce(self.h, comm.h))

    def novel_sets(self, comm=None):
        """[(members tuple, count)] of the sets that are no index class, sorted by (length, contents): the id of
        entry i is n_eq + i.  With `comm`: the tables of all ranks merged (psa_mapper_novel_allgather)."""
        out = _NovelSets()
        if comm is None:
            _check(lib().psa_mapper_novel_sets(self.h, C.byref(out)))
        else:
            _check(lib().psa_mapper_novel_allgather(self.h, comm.h, C.byref(out)))
        res = _novel_to_py(out)
        lib().psa_novel_sets_free(C.byref(out))
        return res

    def launch_count(self):
        return int(lib().psa_mapper_launch_count(self.h))

    def profile_enable(self, on=True):
        _check(lib().psa_mapper_profile_enable(self.h, 1 if on else 0))

    def set_fast_path(self, max_probes, max_small=32):
        _check(lib().psa_mapper_set_fast_path(self.h, int(max_probes), int(max_small)))

    def set_scan_width(self, lanes):
        _check(lib().psa_mapper_set_scan_width(self.h, int(lanes)))

    def profile_read(self):
        """-> {kernel: (summed device ms, launches)} since the last read."""
        ms, n = (C.c_double * 3)(), (C.c_uint64 * 3)()
        _check(lib().psa_mapper_profile_read(self.h, C.byref(ms), C.byref(n)))
        return {"k_map_thread": (float(ms[0]), int(n[0])), "k_map": (float(ms[1]), int(n[1])),
                "k_seed_scan": (float(ms[2]), int(n[2]))}

    def close(self):
        if getattr(self, "h", None):
            lib().psa_mapper_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_CODE = np.zeros(256, dtype=np.uint8)
for _c, _v in (("A", 0), ("C", 1), ("G", 2), ("T", 3)):
    _CODE[ord(_c)] = _v
    _CODE[ord(_c.lower())] = _v


def pack_ascii_host(seq):
    """DnaString::from_dna_string for one read on the host (used only to feed map_read, which
    takes a DnaString like the reference's; batches are packed on the GPU)."""
    if isinstance(seq, str):
        seq = seq.encode()
    codes = _CODE[np.frombuffer(seq, dtype=np.uint8)].astype(np.uint64)
    n = len(codes)
    nw = (n + 31) // 32
    pad = np.zeros(max(nw, 1) * 32, dtype=np.uint64)
    pad[:n] = codes
    shifts = (62 - 2 * np.arange(32, dtype=np.uint64)).astype(np.uint64)
    return np.bitwise_or.reduce(pad.reshape(-1, 32) << shifts, axis=1).astype(np.uint64)


class Pseudoaligner:
    """Mirror of `Pseudoaligner<K>` (ref src/pseudoaligner.rs:26-33) with the index in HBM.

    flat: the index in the flat form of psa_index_desc; tx_names / tx_gene_mapping are carried
    along untouched like the reference's fields (unused at map time)."""

    def __init__(self, flat, tx_names=None, tx_gene_mapping=None, device=0, gamma=0.0, chunk_reads=0):
        self.index = Index(flat, device=device, gamma=gamma)
        self.mapper = Mapper(self.index, chunk_reads)
        self.tx_names = tx_names or []
        self.tx_gene_mapping = tx_gene_mapping or {}
        self.k = self.index.k

    def map_read(self, read_seq):
        """ref :381 -- `None` or `(eq_class, coverage)`; read_seq is ASCII (str/bytes)."""
        if isinstance(read_seq, str):
            read_seq = read_seq.encode()
        return self.mapper.map_read_packed(pack_ascii_host(read_seq), len(read_seq))

    def map_reads(self, seqs):
        """Batch form: [(None | (eq_class, coverage))] in input order."""
        hits, tx = self.mapper.map_ascii(seqs)
        out = []
        for h in hits:
            if not (h["flags"] & FLAG_ALIGNED):
                out.append(None)
            else:
                o, n = int(h["tx_off"]), int(h["n_tx"])
                out.append((tx[o:o + n].tolist(), int(h["coverage"])))
        return out

    def close(self):
        self.mapper.close()
        self.index.close()


def expand_compact(hits_c, novel_tx, eq_offsets, eq_members):
    """psa_expand_compact: compact records + members of the non-class sets -> (psa_hit array, every member)."""
    hits_c = np.ascontiguousarray(hits_c)
    novel_tx = np.ascontiguousarray(novel_tx, dtype=np.uint32)
    eq_offsets = np.ascontiguousarray(eq_offsets, dtype=np.uint64)
    eq_members = np.ascontiguousarray(eq_members, dtype=np.uint32)
    n = len(hits_c)
    hits = np.zeros(n, dtype=HIT_DTYPE)
    used = C.c_uint64()
    _check(lib().psa_expand_compact(_ptr(hits_c), n, _ptr(novel_tx), len(novel_tx), _ptr(eq_offsets), _ptr(eq_members),
                                    len(eq_offsets) - 1, _ptr(hits), None, 0, C.byref(used)))
    tx = np.zeros(int(used.value) + 1, np.uint32)
    _check(lib().psa_expand_compact(_ptr(hits_c), n, _ptr(novel_tx), len(novel_tx), _ptr(eq_offsets), _ptr(eq_members),
                                    len(eq_offsets) - 1, _ptr(hits), _ptr(tx), len(tx), C.byref(used)))
    return hits, tx[:int(used.value)]


def format_read_data(flag, read_id, eq_class, coverage):
    """The `{:?}` of `(bool, String, Vec<u32>, usize)` printed at ref src/pseudoaligner.rs:490 (the id through
    psa_debug_str, the escaping the native driver uses)."""
    raw = read_id.encode() if isinstance(read_id, str) else bytes(read_id)
    out = C.create_string_buffer(10 * len(raw) + 2)
    n = lib().psa_debug_str(raw, len(raw), out, len(out))
    if n < 0:
        raise ValueError("read id is not valid UTF-8")
    return "(%s, %s, [%s], %d)" % ("true" if flag else "false", out.raw[:n].decode(), ", ".join(str(int(t)) for t in eq_class),
                                   coverage)


def process_reads_file(fastq_path, index, out_path=None, num_threads=2, batch_reads=0, progress=False):
    """psa_process_reads: the native (C++) map driver, FASTQ file -> lines in `out_path` (stdout when
    None).  `index` is a Pseudoaligner or an Index.  Returns {reads, mapped, aligned, seconds}."""
    ix = index.index if isinstance(index, Pseudoaligner) else index
    st = _ProcessStats()
    _check(lib().psa_process_reads(ix.h, os.fsencode(fastq_path), os.fsencode(out_path) if out_path else None,
                                   int(num_threads), int(batch_reads), 1 if progress else 0, C.byref(st)))
    return {"reads": int(st.reads), "mapped": int(st.mapped), "aligned": int(st.aligned), "seconds": float(st.seconds),
            "reader_seconds": float(st.reader_seconds), "mapper_seconds": float(st.mapper_seconds),
            "writer_seconds": float(st.writer_seconds)}


def process_reads(records, index, outdir=None, num_threads=1, out=None, batch_reads=1 << 20):
    """ref src/pseudoaligner.rs:420-514 for an iterable of (id, seq) FASTQ records.

    Prints one line per read, `(flag, "id", [tx...], coverage)`, to `out` (stdout by default).
    The reference prints in arrival order of its worker threads; here lines come in input
    order.  num_threads / outdir are accepted for signature parity (outdir is only logged
    upstream, :428).  Returns (reads, mapped) like the counters at :476-477."""
    out = out or sys.stdout
    n_reads = n_mapped = 0
    batch = []

    def flush():
        nonlocal n_reads, n_mapped
        if not batch:
            return
        hits, tx = index.mapper.map_ascii([s for _, s in batch])
        lines = []
        for (rid, _), h in zip(batch, hits):
            o, n = int(h["tx_off"]), int(h["n_tx"])
            flag = bool(h["flags"] & FLAG_MAPPED)
            lines.append(format_read_data(flag, rid, tx[o:o + n], int(h["coverage"])))
            n_mapped += flag
        out.write("\n".join(lines) + "\n")
        n_reads += len(batch)
        batch.clear()

    for rec in records:
        batch.append(rec)
        if len(batch) >= batch_reads:
            flush()
    flush()
    return n_reads, n_mapped


def _rust_f64(x):
    """`{}` of an f64 as Rust prints it: the shortest decimal that round-trips, never in exponent form, NaN as `NaN`."""
    import decimal
    if x != x:
        return "NaN"
    if x in (float("inf"), float("-inf")):
        return "inf" if x > 0 else "-inf"
    t = format(decimal.Decimal(repr(float(x))), "f")
    if "." in t:
        t = t.rstrip("0").rstrip(".")
    return t


def mappability_tsv(tx_names, gene_names, tx_multiplicity, gene_multiplicity):
    """tx_mappability.tsv as the reference writes it (ref src/mappability.rs:30-31 header, :74-91 to_tsv / fractions)."""
    lines = ["tx_name\tgene_name\ttx_kmer_count\tfrac_kmer_unique_tx\tfrac_kmer_unique_gene"]
    for name, gene, tm, gm in zip(tx_names, gene_names, tx_multiplicity, gene_multiplicity):
        total = int(np.sum(tm, dtype=np.uint64))
        fu_tx = float(tm[0]) / total if total else float("nan")
        fu_gene = float(gm[0]) / total if total else float("nan")
        lines.append("%s\t%s\t%d\t%s\t%s" % (name, gene, total, _rust_f64(fu_tx), _rust_f64(fu_gene)))
    return "\n".join(lines) + "\n"
