"""CPU checks of the code the CUDA kernels are built from.

tests/hostsim/ compiles rust-pseudoaligner_b200/csrc/psa_core.cuh with g++ and a serial warp
policy; here its results are compared with the oracle read by read.  This pins, without a
GPU: the bucket-cascade dictionary layout + fingerprint + verification, the
successor/predecessor tables, the mismatch masks of both extension loops, the map_read state
machine under the cooperative policy AND under the thread-per-read policy (psa_thread.cuh, its
hand-overs redone by the serial stand-in of the cooperative kernel), class windows, and ASCII
packing.  (The warp-level glue itself is covered by the `-m gpu` parity tests.)"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import cases
import orc
import util

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostsim")


@pytest.fixture(scope="module")
def hs():
    subprocess.check_call(["make", "-C", _DIR, "libhostsim.so"], stdout=subprocess.DEVNULL)
    L = C.CDLL(os.path.join(_DIR, "libhostsim.so"))
    vp, u32, u64 = C.c_void_p, C.c_uint32, C.c_uint64
    L.hs_index_create.restype = vp
    L.hs_index_create.argtypes = [u32, u64, vp, u64, vp, vp, vp, vp, u64, vp, vp, C.c_double]
    L.hs_index_error.restype, L.hs_index_error.argtypes = C.c_int, [vp]
    L.hs_index_levels.restype, L.hs_index_levels.argtypes = u32, [vp]
    L.hs_index_fp_bits.restype, L.hs_index_fp_bits.argtypes = u32, [vp]
    L.hs_index_destroy.argtypes = [vp]
    L.hs_lookup.restype, L.hs_lookup.argtypes = C.c_int, [vp, vp, C.POINTER(u32), C.POINTER(u32)]
    L.hs_map_batch.restype = u64
    L.hs_map_batch.argtypes = [vp, vp, vp, vp, u64, u32, vp, vp, u64]
    L.hs_map_batch_thread.restype = u64
    L.hs_map_batch_thread.argtypes = [vp, vp, vp, vp, u64, u32, u32, u32, vp, vp, u64, C.POINTER(u64)]
    L.hs_pack_ascii.argtypes = [C.c_char_p, u64, vp]
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class HsIndex:
    def __init__(self, L, flat, gamma=0.0):
        self.L = L
        f = {k: (np.ascontiguousarray(v) if isinstance(v, np.ndarray) else v) for k, v in flat.items()}
        self.keep = f
        self.h = L.hs_index_create(int(f["k"]), len(f["node_len"]), _p(f["seq_words"]), len(f["seq_words"]),
                                   _p(f["node_start"]), _p(f["node_len"]), _p(f["node_exts"]), _p(f["node_eq"]),
                                   len(f["eq_offsets"]) - 1, _p(f["eq_offsets"]), _p(f["eq_members"]), gamma)
        assert L.hs_index_error(self.h) == 0, "hostsim index error %d" % L.hs_index_error(self.h)

    def map_batch(self, words, off, lens, allowed=2):
        n = len(lens)
        hits = np.zeros(n, dtype=orc.HIT_DTYPE)
        cap = max(64 * n, 4096)
        while True:
            tx = np.zeros(cap, np.uint32)
            need = self.L.hs_map_batch(self.h, _p(words), _p(off), _p(lens), n, allowed, _p(hits), _p(tx), cap)
            if need <= cap:
                return hits, tx[:need]
            cap = int(need)

    def map_batch_thread(self, words, off, lens, max_probes, max_small, allowed=2):
        """The blocking thread-per-read policy, hand-overs redone by the serial policy."""
        n = len(lens)
        hits = np.zeros(n, dtype=orc.HIT_DTYPE)
        cap = max(64 * n, 4096)
        nd = C.c_uint64()
        while True:
            tx = np.zeros(cap, np.uint32)
            need = self.L.hs_map_batch_thread(self.h, _p(words), _p(off), _p(lens), n, allowed, max_probes, max_small,
                                              _p(hits), _p(tx), cap, C.byref(nd))
            if need <= cap:
                return hits, tx[:need], int(nd.value)
            cap = int(need)

    def close(self):
        self.L.hs_index_destroy(self.h)


def _compare(ix_orc, ix_hs, reads, cfgs=((1, 4), (3, 64), (64, 1 << 30), (10, 32), (11, 32)), allowed=2):
    words, off, lens = orc.pack_reads(reads)
    h1, t1, _, _ = ix_orc.map_batch(words, off, lens, allowed=allowed)
    h2, t2 = ix_hs.map_batch(words, off, lens, allowed=allowed)
    a, b = orc.hits_to_tuples(h1, t1), orc.hits_to_tuples(h2, t2)
    for i, (x, y) in enumerate(zip(a, b)):
        assert x == y, (i, reads[i], x, y)
    assert np.array_equal(h1["eq_id"], h2["eq_id"])
    # the thread-per-read policy with its hand-overs: same results whatever the split
    for max_probes, max_small in cfgs:
        h4, t4, nd = ix_hs.map_batch_thread(words, off, lens, max_probes, max_small, allowed=allowed)
        d = orc.hits_to_tuples(h4, t4)
        for i, (x, y) in enumerate(zip(a, d)):
            assert x == y, ("thread", max_probes, max_small, i, reads[i], x, y)
        assert np.array_equal(h1["eq_id"], h4["eq_id"]) and np.array_equal(t1, t4)
        _DEFER[(max_probes, max_small)] = _DEFER.get((max_probes, max_small), 0) + nd
        _DEFER["reads"] = _DEFER.get("reads", 0) + len(reads)
    return a


_DEFER = {}


def test_thread_policy_hands_over_some_but_not_all(hs, orc_index_for, fixture_fasta):
    """With one probe per seed search a thread must hand the noisy reads over and keep the clean
    ones; with unbounded probes only class-list overflows are handed over."""
    ix = orc_index_for(20)
    hx = HsIndex(hs, ix.flat())
    rng = np.random.default_rng(5)
    reads = util.sample_reads(rng, fixture_fasta[1], 3000, 150, p_sub=0.005)
    words, off, lens = orc.pack_reads(reads)
    _, _, nd1 = hx.map_batch_thread(words, off, lens, 1, 64)
    _, _, nd64 = hx.map_batch_thread(words, off, lens, 64, 1 << 30)
    assert 0 <= nd64 <= nd1 < len(reads) // 2 and nd1 > 0, (nd1, nd64)
    hx.close()


@pytest.mark.parametrize("k,length", [(20, 150), (20, 60), (24, 91), (64, 150)])
def test_hostsim_vs_oracle_fixture(hs, orc_index_for, fixture_fasta, fixture_fastq, k, length):
    ix = orc_index_for(k)
    hx = HsIndex(hs, ix.flat())
    assert hs.hs_index_fp_bits(hx.h) >= 16
    rng = np.random.default_rng(77 * k + length)
    n_left = 0
    for name, reads in cases.read_sets(rng, fixture_fasta[1], length, k).items():
        res = _compare(ix, hx, reads)
        if name == "left_ext":
            n_left = sum(1 for r in res if r[0])
    assert n_left > 100
    if k == 20 and length == 60:
        _compare(ix, hx, [s for _, s in fixture_fastq])
    hx.close()


def test_hostsim_lookup_every_kmer(hs, orc_index_for):
    """validate_dbg (a)-style: every k-mer of every node resolves to its (node, offset); k-mers
    that are not in the graph resolve to nothing (ref src/build_index.rs:263-298, :99-107)."""
    ix = orc_index_for(20)
    flat = ix.flat()
    hx = HsIndex(hs, flat)
    nt = util.node_kmer_table(flat)
    rng = np.random.default_rng(3)
    pick = rng.choice(len(nt["lo"]), 20000, replace=False)
    n, o = C.c_uint32(), C.c_uint32()
    for i in pick.tolist():
        w = np.array([int(nt["lo"][i]) << (64 - 40), 0], dtype=np.uint64)
        assert hs.hs_lookup(hx.h, _p(w), C.byref(n), C.byref(o)) == 1
        assert (n.value, o.value) == (int(nt["node"][i]), int(nt["off"][i]))
    present = set(nt["lo"].tolist())
    miss = 0
    for v in rng.integers(0, 1 << 40, 20000).tolist():
        if v in present:
            continue
        w = np.array([v << 24, 0], dtype=np.uint64)
        assert hs.hs_lookup(hx.h, _p(w), C.byref(n), C.byref(o)) == 0
        miss += 1
    assert miss > 19000
    hx.close()


@pytest.mark.parametrize("k", [5, 19, 31, 32, 33, 47, 64])
def test_hostsim_vs_oracle_random_transcriptomes(hs, k):
    rng = np.random.default_rng(100 + k)
    seqs = util.random_transcriptome(rng, n_genes=8, k=k)
    ix = orc.OrcIndex.build(seqs, k)
    for gamma in (0.0, 1.0, 3.0):
        hx = HsIndex(hs, ix.flat(), gamma)
        for length in (k, k + 1, 2 * k + 3, 150, 1100):
            for name, reads in cases.read_sets(rng, seqs, length, k, scale=0.1).items():
                _compare(ix, hx, [r[:length] if name != "edge" else r for r in reads])
        hx.close()


def test_hostsim_pack_ascii_matches_oracle(hs):
    rng = np.random.default_rng(9)
    for _ in range(200):
        n = int(rng.integers(0, 200))
        s = bytes(rng.integers(0, 256, n, dtype=np.uint8))
        a = np.zeros((n + 31) // 32 + 1, np.uint64)
        b = np.zeros((n + 31) // 32 + 1, np.uint64)
        hs.hs_pack_ascii(s, n, _p(a))
        orc.lib().orc_pack_ascii(s, n, _p(b))
        assert np.array_equal(a, b), s


def test_hostsim_wide_classes(hs, fixture_fasta):
    """Transcripts in shuffled order: isoforms of a gene are no longer neighbours, so most
    multi-member classes span >= 192 ids and take the list path next to the windows."""
    rng = np.random.default_rng(12)
    seqs = list(fixture_fasta[1][:700])
    rng.shuffle(seqs)
    ix = orc.OrcIndex.build(seqs, 20)
    flat = ix.flat()
    eo, em = flat["eq_offsets"], flat["eq_members"]
    wide = sum(1 for c in range(len(eo) - 1) if eo[c + 1] > eo[c] and int(em[eo[c + 1] - 1]) - int(em[eo[c]]) >= 192)
    assert wide > 200
    hx = HsIndex(hs, flat)
    for name, reads in cases.read_sets(rng, seqs, 150, 20, scale=0.5).items():
        _compare(ix, hx, reads)
    hx.close()


def _assert_constructed_walks(ix, reads):
    """The constructed reads take the branches they were built for (asserted on the oracle's walk)."""
    d = {name: cases.describe_walk(ix, r) for name, r in reads.items()}
    q1 = d["quirk1_offset0_left_walk"]
    assert q1["first_seed"][0] == 30 and q1["first_seed"][2] == 0 and q1["left_steps"] >= 1      # QUIRK-1: seed at unitig offset 0, left walk
    assert d["left_walk_offset2"]["first_seed"][0] == 30 and d["left_walk_offset2"]["left_steps"] >= 2   # QUIRK-3: >= 2 predecessors
    assert d["left_to_read_start"]["first_seed"][0] == 30 and d["left_to_read_start"]["left_steps"] == 0
    assert d["left_to_read_start"]["result"][1] == 150                      # every base covered: the walk reached the read's start
    assert d["two_errors_per_node"]["result"][1] == 150 and len(d["two_errors_per_node"]["nodes"]) >= 3   # QUIRK-2
    assert d["reseed_into_visited_node"]["revisits"] >= 1                   # QUIRK-5 + re-seed into a visited node
    assert d["break_near_end"]["result"][1] < 150 and d["break_near_end"]["revisits"] == 0
    assert d["ends_at_unitig_end"]["result"][1] == 150
    assert d["no_right_ext"]["result"][1] == 100
    return d


def test_constructed_branches(hs):
    """QUIRK-1 (seed at unitig offset 0 + left extension, ref :129), QUIRK-3 (left walk over >= 2 predecessors,
    ref :199), QUIRK-2 (per-node budget), re-seed into a visited node (ref :293), break near the read's end
    (ref :287-290), read ending at a unitig end, missing right extension: constructed reads on a constructed
    graph, the branch asserted on the oracle's walk, then both policies against the oracle."""
    S, seqs = cases.constructed_transcriptome()
    ix = orc.OrcIndex.build(seqs, 20)
    reads = cases.constructed_reads(S)
    _assert_constructed_walks(ix, reads)
    hx = HsIndex(hs, ix.flat())
    _compare(ix, hx, [r.decode() for r in reads.values()])
    hx.close()


@pytest.mark.parametrize("allowed", [0, 1, 3, 7])
def test_allowed_mismatches(hs, orc_index_for, fixture_fasta, allowed):
    """map_read_with_mismatch (ref :361) with allowed_mismatches other than the default 2."""
    ix = orc_index_for(20)
    hx = HsIndex(hs, ix.flat())
    rng = np.random.default_rng(40 + allowed)
    reads = util.sample_reads(rng, fixture_fasta[1], 1500, 150, p_sub=0.03, mix=(0.9, 0.1, 0.0))
    reads += cases.left_extension_reads(rng, fixture_fasta[1], 300, 150, 20)
    res = _compare(ix, hx, reads, allowed=allowed)
    words, off, lens = orc.pack_reads(reads)
    base = ix.map_batch(words, off, lens)[0]                        # A = 2: the answers must actually depend on A
    h = ix.map_batch(words, off, lens, allowed=allowed)[0]
    assert not np.array_equal(h["coverage"], base["coverage"])
    assert len(res) == len(reads)
    hx.close()
