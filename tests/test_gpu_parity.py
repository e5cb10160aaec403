"""GPU parity: the CUDA path through the C ABI (libpsa_b200.so) against the oracle, read by
read, bit-exact on (aligned?, flag, sorted transcript set, coverage) and on eq_id.

Covers the reference's own pins for the path (validate_dbg (b), test_alignment known answers,
test/small.fq properties -- ref src/build_index.rs:300-367, :429-441) and every branch of
map_read the reference never tests (left extension, re-seed, >2 mismatches per unitig, N,
short reads), at k = 20 / 24 / 64 and read lengths 60 / 91 / 150 / 1100."""
import importlib
import os

import numpy as np
import pytest

import cases
import orc
import util
from conftest import GOLDEN

pytestmark = pytest.mark.gpu

pkg = importlib.import_module("rust-pseudoaligner_b200")
host = importlib.import_module("rust-pseudoaligner_b200.host")

_PA = {}


@pytest.fixture(scope="module")
def pa_for(orc_index_for):
    def get(k):
        if k not in _PA:
            _PA[k] = pkg.Pseudoaligner(orc_index_for(k).flat(), device=0)
        return _PA[k]
    yield get
    for p in _PA.values():
        p.close()
    _PA.clear()


def _oracle(ix, reads):
    words, off, lens = orc.pack_reads(reads)
    hits, tx, counts, ev = ix.map_batch(words, off, lens, counts=True)
    return hits, tx, counts, ev


def _assert_same(reads, got_hits, got_tx, want_hits, want_tx):
    a, b = orc.hits_to_tuples(got_hits, got_tx), orc.hits_to_tuples(want_hits, want_tx)
    bad = [i for i in range(len(reads)) if a[i] != b[i]]
    assert not bad, "%d/%d reads differ; first: read %d %r\n gpu    %r\n oracle %r" % (
        len(bad), len(reads), bad[0], reads[bad[0]], a[bad[0]], b[bad[0]])
    assert np.array_equal(got_hits["eq_id"], want_hits["eq_id"])
    # deterministic layout: members back to back in read order
    assert np.array_equal(got_hits["tx_off"], want_hits["tx_off"])
    assert np.array_equal(got_tx, want_tx)


def test_known_answers(pa_for):
    """test_alignment, ref src/build_index.rs:423-451."""
    pa = pa_for(20)
    ex1 = "GGCTGTCAACCAGTCCATAGGCAGGGCCATCAGGCACCAAAGGGATTCTGCCAGCATAGT"
    snp = "GGCTGTCAACCAGTCCATAGGCGGGGCCATCAGGCACCAAAGGGATTCTGCCAGCATAGT"
    assert pa.map_read(ex1) == ([1, 30], 60)
    assert pa.map_read(snp) == ([1, 30], 60)
    assert pa.map_read("ACGT") is None
    assert pa.map_reads([ex1, snp, "ACGT", ""]) == [([1, 30], 60), ([1, 30], 60), None, None]


@pytest.mark.parametrize("k", [20, 64])
def test_validate_dbg_part_b(pa_for, orc_index_for, fixture_fasta, k):
    """ref src/build_index.rs:300-367: every transcript maps to itself with coverage == len;
    singleton class == [i]; otherwise i is a member.  (Reads up to 16 355 bases.)"""
    pa = pa_for(k)
    seqs = [s for s in fixture_fasta[1]]
    res = pa.map_reads(seqs)
    for i, (s, r) in enumerate(zip(seqs, res)):
        if len(s) < k:
            assert r is None
            continue
        eq, cov = r
        assert cov == len(s), i
        assert i in eq
    hits, tx, _, _ = _oracle(orc_index_for(k), [s.decode() for s in seqs])
    want = [None if not a else (list(e), c) for a, _, e, c in orc.hits_to_tuples(hits, tx)]
    assert res == want


@pytest.mark.parametrize("k,length", [(20, 150), (20, 60), (24, 91), (64, 150)])
def test_parity_fixture_read_sets(pa_for, orc_index_for, fixture_fasta, k, length):
    ix, pa = orc_index_for(k), pa_for(k)
    rng = np.random.default_rng(31 * k + length)
    for name, reads in cases.read_sets(rng, fixture_fasta[1], length, k, scale=2.0).items():
        want_hits, want_tx, _, _ = _oracle(ix, reads)
        # tuning knobs that must not change results: lanes per read of the cooperative kernel,
        # and how much the thread-per-read kernel keeps for itself (0 probes = it is skipped)
        # and the lanes of the seed-scan kernel (0 = long first searches go to the cooperative kernel)
        for lanes, probes, max_small, scan in ((8, 0, 0, 16), (16, 1, 4, 8), (32, 3, 32, 0), (8, 64, 1 << 30, 32),
                                               (8, 1, 32, 32), (8, 2, 32, 16)):
            pa.mapper.set_group_width(lanes)
            pa.mapper.set_fast_path(probes, max_small)
            pa.mapper.set_scan_width(scan)
            got_hits, got_tx = pa.mapper.map_ascii(reads)
            _assert_same(reads, got_hits, got_tx, want_hits, want_tx)
    pa.mapper.set_group_width(8)
    pa.mapper.set_fast_path(3, 32)
    pa.mapper.set_scan_width(16)


def test_small_fq(pa_for, orc_index_for, fixture_fastq):
    """test/small.fq (BASELINE config 1): no golden output upstream; oracle parity + properties."""
    ix, pa = orc_index_for(20), pa_for(20)
    reads = [s for _, s in fixture_fastq]
    want_hits, want_tx, _, _ = _oracle(ix, reads)
    got_hits, got_tx = pa.mapper.map_ascii(reads)
    _assert_same(reads, got_hits, got_tx, want_hits, want_tx)
    exact = [i for i, (rid, _) in enumerate(fixture_fastq) if "_err" not in rid and not rid.endswith("_rev")]
    assert len(exact) == 3103 and (got_hits["coverage"][exact] == 60).all()


def test_long_reads_and_chunking(orc_index_for, fixture_fasta):
    """Reads of 1100 bases (35 words: the compare loop's second round), a pipeline cut into
    many small chunks, packed input, counts across calls."""
    ix = orc_index_for(20)
    pa = pkg.Pseudoaligner(ix.flat(), device=0, chunk_reads=257)
    rng = np.random.default_rng(5)
    reads = util.sample_reads(rng, fixture_fasta[1], 1500, 1100, p_sub=0.004, mix=(0.8, 0.15, 0.05))
    reads += util.sample_reads(rng, fixture_fasta[1], 700, 150, p_sub=0.01) + ["", "ACGT"]
    rng.shuffle(reads)
    want_hits, want_tx, want_counts, _ = _oracle(ix, reads)
    got_hits, got_tx = pa.mapper.map_ascii(reads)
    _assert_same(reads, got_hits, got_tx, want_hits, want_tx)
    assert np.array_equal(pa.mapper.counts(), want_counts)
    # packed input, same answer; counts accumulate
    words, off, lens = orc.pack_reads(reads)
    got_hits, got_tx = pa.mapper.map_packed(words, off, lens)
    _assert_same(reads, got_hits, got_tx, want_hits, want_tx)
    assert np.array_equal(pa.mapper.counts(), 2 * want_counts)
    pa.mapper.counts_reset()
    assert int(pa.mapper.counts().sum()) == 0
    # hits only (no member buffer)
    h2, t2 = pa.mapper.map_ascii(reads, want_tx=False)
    assert np.array_equal(h2["n_tx"], want_hits["n_tx"]) and np.array_equal(h2["coverage"], want_hits["coverage"])
    assert len(t2) == 0
    pa.close()


def _check_events(ev, want_ev):
    for key_gpu, key_orc in (("reads", "reads"), ("read_bases", "read_bases"), ("kmer_lookups", "kmer_lookups"),
                             ("node_visits", "node_visits"), ("bases_compared", "bases_compared"),
                             ("edge_jumps", "edge_jumps"), ("out_members", "out_members"), ("aligned", "aligned")):
        assert ev[key_gpu] == want_ev[key_orc], (key_gpu, ev, want_ev)
    assert ev["verifications"] >= want_ev["dict_hits"] and ev["dict_levels"] >= ev["kmer_lookups"]
    assert ev["dict_hits"] == ev["verifications"]


def test_device_batch_and_events(orc_index_for, fixture_fasta):
    """Device-resident fixed-stride ASCII batch (the bench's kernel-only arm), async entry,
    and the event counters against the oracle's (hash-independent ones must agree exactly)."""
    ix = orc_index_for(20)
    pa = pkg.Pseudoaligner(ix.flat(), device=0)
    rng = np.random.default_rng(8)
    L, n = 150, 6000
    reads = util.sample_reads(rng, fixture_fasta[1], n, L, p_sub=0.01)
    data = np.frombuffer("".join(reads).encode(), dtype=np.uint8)
    want_hits, want_tx, want_counts, want_ev = _oracle(ix, reads)
    b = pkg.DeviceBatch(pkg.pseudoaligner.READS_ASCII, data, n, stride=L, fixed_len=L, tx_cap=64 * n)
    pa.mapper.map_device(b)
    got_hits, got_tx = b.download()
    _assert_same(reads, got_hits, got_tx, want_hits, want_tx)
    pa.mapper.counts_reset()
    pa.mapper.map_device_async(b)
    pa.mapper.sync()
    got_hits, got_tx = b.download()
    _assert_same(reads, got_hits, got_tx, want_hits, want_tx)
    assert np.array_equal(pa.mapper.counts(), want_counts)
    for probes, scan in ((0, 16), (1, 0), (1, 8), (3, 16), (3, 32), (64, 16)):
        pa.mapper.set_fast_path(probes, 32)
        pa.mapper.set_scan_width(scan)
        parts = pa.mapper.map_device_events(b, split=True)
        ev = {key: sum(p[key] for p in parts) for key in parts[0]}
        assert (parts[0]["reads"] == 0) == (probes == 0) and parts[1]["reads"] > 0
        assert (parts[2]["kmer_lookups"] > 0) == (scan > 0 and 0 < probes < 64)
        _check_events(ev, want_ev)
    pa.mapper.set_fast_path(3, 32)
    pa.mapper.set_scan_width(16)
    ev = pa.mapper.map_device_events(b)
    _check_events(ev, want_ev)
    # several batches queued before one sync: an overflow of an EARLIER one is still reported
    b_small = pkg.DeviceBatch(pkg.pseudoaligner.READS_ASCII, data, n, stride=L, fixed_len=L, tx_cap=10)
    pa.mapper.map_device_async(b_small)
    pa.mapper.map_device_async(b)
    with pytest.raises(pkg.PsaError) as e:
        pa.mapper.sync()
    assert e.value.code == -4
    pa.mapper.counts_reset()
    pa.mapper.map_device_async(b)
    pa.mapper.sync()                      # and the flag does not stick beyond that sync
    b_small.free()
    # packed device batch
    words, off, lens = orc.pack_reads(reads)
    b2 = pkg.DeviceBatch(pkg.pseudoaligner.READS_PACKED, words, n, read_off=off, read_len=lens, tx_cap=64 * n)
    pa.mapper.map_device(b2)
    got_hits, got_tx = b2.download()
    _assert_same(reads, got_hits, got_tx, want_hits, want_tx)
    # tx_buf too small is reported, not truncated silently, and does not disturb the counts
    before = pa.mapper.counts()
    b3 = pkg.DeviceBatch(pkg.pseudoaligner.READS_ASCII, data, n, stride=L, fixed_len=L, tx_cap=10)
    with pytest.raises(pkg.PsaError) as e:
        pa.mapper.map_device(b3)
    assert e.value.code == -4 and b3.ob.tx_used == len(want_tx)
    assert np.array_equal(pa.mapper.counts(), before)
    for x in (b, b2, b3):
        x.free()
    pa.close()


def test_lookup_every_kmer(pa_for, orc_index_for):
    """dbg_index.get + verification over the whole dictionary and over absent k-mers."""
    ix, pa = orc_index_for(20), pa_for(20)
    nt = util.node_kmer_table(ix.flat())
    found, node, off = pa.index.lookup(nt["lo"] << np.uint64(24))
    assert found.all() and np.array_equal(node, nt["node"]) and np.array_equal(off, nt["off"])
    rng = np.random.default_rng(2)
    q = rng.integers(0, 1 << 40, 200000).astype(np.uint64)
    absent = ~np.isin(q, nt["lo"])
    found, _, _ = pa.index.lookup(q << np.uint64(24))
    assert not found[absent].any() and found[~absent].all()
    info = pa.index.info()
    assert info["n_kmers"] == len(nt["lo"]) and info["fp_bits"] >= 16 and 3 <= info["dict_levels"] <= 32


@pytest.mark.parametrize("k", [5, 19, 31, 32, 33, 47, 64])
def test_parity_random_transcriptomes(k):
    """Small adversarial transcriptomes (shared exons, repeats, poly-A self loop, tandem cycle)
    built by the PRODUCT host builder, every k-mer width edge, several MPHF gammas."""
    rng = np.random.default_rng(200 + k)
    seqs = util.random_transcriptome(rng, n_genes=8, k=k)
    codes, off = host.encode_transcripts(seqs)
    flat, _ = host.build_graph(codes, off, k)
    ix = orc.OrcIndex.from_flat(flat)
    for gamma, lanes, probes in ((0.0, 8, 3), (1.0, 16, 0), (4.0, 32, 1)):
        pa = pkg.Pseudoaligner(flat, device=0, gamma=gamma)
        pa.mapper.set_group_width(lanes)
        pa.mapper.set_fast_path(probes, 8)
        reads = []
        for length in (k, k + 1, 2 * k + 3, 150, 1100):
            for name, rs in cases.read_sets(rng, seqs, length, k, scale=0.1).items():
                reads += [r[:length] if name != "edge" else r for r in rs]
        want_hits, want_tx, want_counts, _ = _oracle(ix, reads)
        got_hits, got_tx = pa.mapper.map_ascii(reads)
        _assert_same(reads, got_hits, got_tx, want_hits, want_tx)
        assert np.array_equal(pa.mapper.counts(), want_counts)
        pa.close()


# the reference's own vectors for intersect, ref src/pseudoaligner.rs:544-559
INTERSECT_VECTORS = [
    [1, 2, 3, 4, 5, 6, 7, 8, 9], [1, 2, 3], [1, 4, 5], [7, 8, 9], [9], [], [1, 2, 3, 6, 7, 8, 9], [1, 7, 8, 9, 10],
    [10, 15, 20], [21, 22, 23], [0], [0, 1000, 5000], [0, 1000, 1000001], [5], [100000000], [1, 23, 45, 1000001, 100000000],
]


def test_intersect_vectors_on_device():
    """intersect_test + intersect_prop_test (ref src/pseudoaligner.rs:542-586) against the DEVICE routines in
    isolation: one thread's list scheme, the cooperative kernel's lane-group scheme and the class windows
    (psa_selftest_intersect), every ordered pair of the reference's 16 vectors, then random ascending lists."""
    psa = pkg.pseudoaligner

    def check(v1, v2):
        want = sorted(set(v1) & set(v2))
        assert orc.intersect(v1, v2) == want
        t, g, w = psa.selftest_intersect(v1, v2)
        assert t == want and g == want, (v1, v2, t, g, want)
        narrow = all((not v) or v[-1] - v[0] < 192 for v in (v1, v2))
        assert (w is not None) == narrow
        if w is not None:
            assert w == want, (v1, v2, w, want)

    for v1 in INTERSECT_VECTORS:
        for v2 in INTERSECT_VECTORS:
            check(v1, v2)
            check(v2, v1)
    rng = np.random.default_rng(6)
    for case in range(300):
        hi = (100, 150, 400, 100000)[case % 4]
        v1 = sorted(set(rng.integers(0, hi, int(rng.integers(0, 400))).tolist()))
        v2 = sorted(set(rng.integers(0, hi, int(rng.integers(0, 400))).tolist()))
        check(v1, v2)


def test_constructed_branches(fixture_fasta):
    """QUIRK-1 (seed at unitig offset 0 + left extension, ref :129), QUIRK-3 (left walk over >= 2 predecessors, ref
    :199), QUIRK-2, re-seed into a visited node (ref :293), break near the read's end (ref :287-290): constructed
    reads on a constructed graph.  The branch is asserted on the oracle's walk; the GPU must give the oracle's
    answer AND count the oracle's events (node visits, edge jumps, bases compared, k-mer lookups) -- the same walk."""
    from test_hostsim import _assert_constructed_walks
    S, seqs = cases.constructed_transcriptome()
    ix = orc.OrcIndex.build(seqs, 20)
    named = cases.constructed_reads(S)
    _assert_constructed_walks(ix, named)
    pa = pkg.Pseudoaligner(ix.flat(), device=0)
    for name, read in named.items():
        reads = [read.decode()]
        want_hits, want_tx, _, want_ev = _oracle(ix, reads)
        for probes, scan in ((64, 0), (3, 8), (0, 8)):        # the thread kernel alone / with the seed scan / cooperative kernel alone
            pa.mapper.set_fast_path(probes, 32)
            pa.mapper.set_scan_width(scan)
            got_hits, got_tx = pa.mapper.map_ascii(reads)
            _assert_same(reads, got_hits, got_tx, want_hits, want_tx)
            data = np.frombuffer(read, dtype=np.uint8)
            b = pkg.DeviceBatch(pkg.pseudoaligner.READS_ASCII, data, 1, stride=len(read), fixed_len=len(read), tx_cap=1024)
            ev = pa.mapper.map_device_events(b)
            b.free()
            _check_events(ev, want_ev)
    pa.close()


@pytest.mark.parametrize("allowed", [0, 1, 3, 7])
def test_allowed_mismatches(pa_for, orc_index_for, fixture_fasta, allowed):
    """map_read_with_mismatch (ref :361) with allowed_mismatches other than DEFAULT_ALLOWED_MISMATCHES:
    psa_mapper_set_allowed_mismatches against an oracle that takes A."""
    ix = orc_index_for(20)
    pa = pkg.Pseudoaligner(ix.flat(), device=0)
    rng = np.random.default_rng(40 + allowed)
    reads = util.sample_reads(rng, fixture_fasta[1], 3000, 150, p_sub=0.03, mix=(0.9, 0.1, 0.0))
    reads += cases.left_extension_reads(rng, fixture_fasta[1], 500, 150, 20)
    words, off, lens = orc.pack_reads(reads)
    want_hits, want_tx, want_counts, _ = ix.map_batch(words, off, lens, counts=True, allowed=allowed)
    base_hits = ix.map_batch(words, off, lens)[0]
    assert not np.array_equal(base_hits["coverage"], want_hits["coverage"])      # the answers do depend on A
    pa.mapper.set_allowed_mismatches(allowed)
    for probes, scan in ((None, None), (0, 8), (64, 0)):
        if probes is not None:
            pa.mapper.set_fast_path(probes, 32)
            pa.mapper.set_scan_width(scan)
        pa.mapper.counts_reset()
        got_hits, got_tx = pa.mapper.map_ascii(reads)
        _assert_same(reads, got_hits, got_tx, want_hits, want_tx)
        assert np.array_equal(pa.mapper.counts(), want_counts)
    pa.close()


def test_config2_one_million_reads(fixture_fasta):
    """BASELINE config 2 as stated: 1 M synthetic 150 bp reads (seed 1: 90 % transcript reads with 0.5 %
    substitutions, 5 % chimeric, 5 % random) vs the gencode_small.fa index, k = 20, one GPU, FULL per-read
    equality with the oracle on (aligned?, flag, transcript set, coverage, eq_id, member layout) and equal counts."""
    codes, off = host.encode_transcripts(fixture_fasta[1])
    tr = host.Transcriptome.from_codes(codes, off)
    flat, _ = host.build_graph(codes, off, 20)
    n, L = 1000000, 150
    data = tr.reads(1, 0, n, L)
    pa = pkg.Pseudoaligner(flat, device=0)
    got_hits, got_tx = pa.mapper.map_ascii_fixed(data, n, L)
    ox = orc.OrcIndex.from_flat(flat)
    import threading
    T = min(16, os.cpu_count() or 1)
    bounds = [n * t // T for t in range(T + 1)]
    res = [None] * T

    def work(t):
        res[t] = ox.map_ascii_fixed(data, n, L, start=bounds[t], stop=bounds[t + 1], counts=True)
    th = [threading.Thread(target=work, args=(t,)) for t in range(T)]
    for x in th:
        x.start()
    for x in th:
        x.join()
    want_hits = np.concatenate([r[0] for r in res])
    base, o = 0, 0
    for r in res:
        want_hits["tx_off"][o:o + len(r[0])] += np.uint64(base)
        base += len(r[1])
        o += len(r[0])
    want_tx = np.concatenate([r[1] for r in res])
    want_counts = sum(r[2] for r in res)
    for f in ("coverage", "n_tx", "eq_id", "flags", "tx_off"):
        assert np.array_equal(got_hits[f], want_hits[f]), f
    assert np.array_equal(got_tx, want_tx)
    assert np.array_equal(pa.mapper.counts(), want_counts)
    assert 0.93 * n < int((got_hits["flags"] & 1).sum()) < 0.97 * n       # ~5 % random reads do not align
    assert orc.result_checksum(want_hits, want_tx) == orc.result_checksum(got_hits, got_tx)
    pa.close()


def test_invalid_index(orc_index_for, fixture_fasta):
    ix = orc_index_for(20)
    flat = ix.flat()
    bad = dict(flat)
    bad["node_exts"] = flat["node_exts"].copy()
    # claim a right extension that no node provides -> "missing link"
    i = int(np.flatnonzero((flat["node_exts"] & 0x0F) == 0)[0])
    bad["node_exts"][i] |= 1
    with pytest.raises(pkg.PsaError) as e:
        pkg.Index(bad)
    assert e.value.code == -5
    dup = dict(flat)
    dup["node_eq"] = flat["node_eq"].copy()
    dup["node_eq"][0] = len(flat["eq_offsets"]) + 5
    with pytest.raises(pkg.PsaError):
        pkg.Index(dup)


def test_synthetic_scale_checksum():
    """A mid-size synthetic GENCODE-shaped index (BASELINE config 3 in miniature) built by the
    product builder: 200 k reads, full per-read equality and equal per-class counts."""
    t = host.Transcriptome.synth(2, 300)
    flat, stats = host.build_graph(t.codes(), t.tx_off(), 24)
    n, L = 200000, 150
    data = t.reads(3, 0, n, L)
    pa = pkg.Pseudoaligner(flat, device=0, chunk_reads=1 << 16)
    got_hits, got_tx = pa.mapper.map_ascii_fixed(data, n, L)
    ox = orc.OrcIndex.from_flat(flat)
    reads = [data[i * L:(i + 1) * L].tobytes().decode() for i in range(n)]
    want_hits, want_tx, want_counts, _ = _oracle(ox, reads)
    _assert_same(reads, got_hits, got_tx, want_hits, want_tx)
    assert np.array_equal(pa.mapper.counts(), want_counts)
    assert int((got_hits["flags"] & 1).sum()) > 0.9 * n
    pa.close()


def test_wide_classes(fixture_fasta):
    """Shuffled transcript order: most multi-member classes are too wide for a window, so the
    list path runs next to the windows in both kernels (and alone when every class is wide)."""
    rng = np.random.default_rng(12)
    seqs = list(fixture_fasta[1][:700])
    rng.shuffle(seqs)
    ix = orc.OrcIndex.build(seqs, 20)
    pa = pkg.Pseudoaligner(ix.flat(), device=0)
    for name, reads in cases.read_sets(rng, seqs, 150, 20, scale=1.0).items():
        want_hits, want_tx, want_counts, _ = _oracle(ix, reads)
        for lanes, probes, max_small, scan in ((8, 0, 0, 16), (16, 3, 2, 0), (32, 64, 1 << 30, 8), (8, 1, 32, 32)):
            pa.mapper.set_group_width(lanes)
            pa.mapper.set_fast_path(probes, max_small)
            pa.mapper.set_scan_width(scan)
            got_hits, got_tx = pa.mapper.map_ascii(reads)
            _assert_same(reads, got_hits, got_tx, want_hits, want_tx)
    pa.close()


def test_process_reads_native(pa_for, orc_index_for, fixture_fastq, tmp_path):
    """psa_process_reads (C++ mirror of ref src/pseudoaligner.rs:420-514) on test/small.fq: the
    printed tuples equal the oracle's, line by line in input order; plain and gzip input, small
    batches (pipeline wrap-around), a truncated file is an error after the good records."""
    import gzip
    ix, pa = orc_index_for(20), pa_for(20)
    want_hits, want_tx, _, _ = _oracle(ix, [s for _, s in fixture_fastq])
    want = []
    for (rid, _), (al, flag, eq, cov) in zip(fixture_fastq, orc.hits_to_tuples(want_hits, want_tx)):
        want.append(pkg.format_read_data(flag, rid, eq, cov))
    fq = tmp_path / "small.fq"
    with gzip.open(os.path.join(GOLDEN, "small.fq.gz"), "rb") as f:
        raw = f.read()
    fq.write_bytes(raw)
    out = tmp_path / "out.txt"
    st = pkg.process_reads_file(str(fq), pa, str(out), num_threads=3, batch_reads=1000)
    assert out.read_text().splitlines() == want
    assert st["reads"] == len(want) and st["mapped"] == sum(1 for l in want if l.startswith("(true"))
    st = pkg.process_reads_file(os.path.join(GOLDEN, "small.fq.gz"), pa, str(out), num_threads=1)
    assert out.read_text().splitlines() == want and st["reads"] == len(want)
    # CRLF line ends, a batch size that does not divide the file; sequences wrapped over two lines
    crlf = tmp_path / "crlf.fq"
    crlf.write_bytes(raw.replace(b"\n", b"\r\n"))
    st = pkg.process_reads_file(str(crlf), pa, str(out), num_threads=2, batch_reads=777)
    assert out.read_text().splitlines() == want and st["reads"] == len(want)
    lines_in = raw.split(b"\n")
    wrapped = tmp_path / "wrapped.fq"
    wrapped.write_bytes(b"".join(h + b"\n" + s[:25] + b"\n" + s[25:] + b"\n" + p_ + b"\n" + q[:25] + b"\n" + q[25:] + b"\n"
                                 for h, s, p_, q in zip(*[iter(lines_in[:4 * len(want)])] * 4)))
    st = pkg.process_reads_file(str(wrapped), pa, str(out), num_threads=3, batch_reads=500)
    assert out.read_text().splitlines() == want and st["reads"] == len(want)
    empty = tmp_path / "empty.fq"
    empty.write_bytes(b"")
    st = pkg.process_reads_file(str(empty), pa, str(out))
    assert st["reads"] == 0 and out.read_text() == ""
    # the Python driver prints the same lines
    import io
    buf = io.StringIO()
    pkg.process_reads(fixture_fastq, pa, out=buf)
    assert buf.getvalue().splitlines() == want
    # truncated record
    bad = tmp_path / "bad.fq"
    bad.write_bytes(raw[:raw.index(b"\n", len(raw) // 2) + 1] + b"@broken\nACGT\n")
    with pytest.raises(pkg.PsaError) as e:
        pkg.process_reads_file(str(bad), pa, str(out))
    assert e.value.code == -7
    n_good = raw[:raw.index(b"\n", len(raw) // 2) + 1].count(b"\n") // 4
    assert out.read_text().splitlines() == want[:n_good]      # the records before the bad one were processed


@pytest.mark.parametrize("tile", ["1", "0"])
def test_tma_read_tiles(orc_index_for, fixture_fasta, monkeypatch, tile):
    """PSA_TILE=1 (the default): the thread-per-read kernel stages the packed reads of a fixed-stride batch in
    shared memory with one bulk asynchronous copy (TMA) per CTA; PSA_TILE=0 reads them through L1.
    Same results, including batches whose last tile is partial and odd-sized."""
    monkeypatch.setenv("PSA_TILE", tile)
    ix = orc_index_for(20)
    pa = pkg.Pseudoaligner(ix.flat(), device=0)
    rng = np.random.default_rng(21)
    for L, n in ((150, 128 * 37 + 61), (91, 1000), (60, 127), (33, 3), (150, 64 * 5 + 1)):
        reads = util.sample_reads(rng, fixture_fasta[1], n, L, p_sub=0.01, mix=(0.85, 0.1, 0.05))
        data = np.frombuffer("".join(reads).encode(), dtype=np.uint8)
        want_hits, want_tx, _, _ = _oracle(ix, reads)
        got_hits, got_tx = pa.mapper.map_ascii_fixed(data, n, L)
        _assert_same(reads, got_hits, got_tx, want_hits, want_tx)
    pa.close()


def test_large_batch_invariants():
    """Size-independent properties at bench scale (2 Mi reads x 150 bp, a 1 500-gene synthetic index):
    the result of every read is the same whichever kernels do the work (default split, cooperative
    kernel alone, thread kernel without the seed scan), mapping a batch twice gives the same bits,
    the per-class counts sum to the number of reads and equal a recount from the hits, and a prefix
    equals the oracle."""
    t = host.Transcriptome.synth(2, 1500)
    flat, _ = host.build_graph(t.codes(), t.tx_off(), 24)
    n, L = 1 << 21, 150
    data = t.reads(3, 0, n, L)
    pa = pkg.Pseudoaligner(flat, device=0)
    n_eq = pa.index.n_eq
    results = []
    for probes, scan, lanes in ((None, None, 8), (0, 8, 8), (3, 0, 16), (1, 32, 32), (None, None, 8)):
        if probes is not None:
            pa.mapper.set_fast_path(probes, 32)
            pa.mapper.set_scan_width(scan)
        pa.mapper.set_group_width(lanes)
        pa.mapper.counts_reset()
        hits, tx = pa.mapper.map_ascii_fixed(data, n, L)
        counts = pa.mapper.counts()
        assert int(counts.sum()) == n
        slot = np.where(hits["flags"] & 1, np.where(hits["eq_id"] == 0xFFFFFFFF, n_eq, hits["eq_id"]), n_eq + 1)
        assert np.array_equal(np.bincount(slot, minlength=n_eq + 2).astype(np.uint64), counts)
        results.append((hits.copy(), tx.copy()))
    for hits, tx in results[1:]:
        assert np.array_equal(hits, results[0][0]) and np.array_equal(tx, results[0][1])
    m = 50000
    ox = orc.OrcIndex.from_flat(flat)
    reads = [data[i * L:(i + 1) * L].tobytes().decode() for i in range(m)]
    want_hits, want_tx, _, _ = _oracle(ox, reads)
    got_hits = results[0][0][:m]
    _assert_same(reads, got_hits, results[0][1][:len(want_tx)], want_hits, want_tx)
    pa.close()


def test_two_mappers_share_one_index_concurrently(orc_index_for, fixture_fasta):
    """`&Pseudoaligner` is shared immutably across the reference's worker threads (ref :35, :434-474);
    here: one psa_index, one psa_mapper per host thread, both mapping at the same time."""
    import threading
    ix = orc_index_for(20)
    index = pkg.Index(ix.flat(), device=0)
    rng = np.random.default_rng(77)
    jobs = []
    for t in range(2):
        reads = util.sample_reads(rng, fixture_fasta[1], 20000, 150 if t == 0 else 91, p_sub=0.01, mix=(0.85, 0.1, 0.05))
        jobs.append((reads, _oracle(ix, reads)))
    out = [None, None]

    def work(t):
        m = pkg.Mapper(index, chunk_reads=1500)
        res = []
        for _ in range(3):
            res.append(m.map_ascii(jobs[t][0]))
        out[t] = (res, m.counts())
        m.close()

    th = [threading.Thread(target=work, args=(t,)) for t in range(2)]
    for x in th:
        x.start()
    for x in th:
        x.join()
    for t in range(2):
        reads, (want_hits, want_tx, want_counts, _) = jobs[t]
        res, counts = out[t]
        for got_hits, got_tx in res:
            _assert_same(reads, got_hits, got_tx, want_hits, want_tx)
        assert np.array_equal(counts, 3 * want_counts)
    index.close()


def test_pack_every_byte_value(pa_for, orc_index_for, fixture_fasta):
    """DnaString::from_dna_string (ref src/pseudoaligner.rs:449-450): A/a C/c G/g T/t -> 0..3, EVERY other byte -> 0.
    Reads with bytes of the whole 0..255 range, through the three pack kernels (ragged, fixed stride, tile)."""
    ix, pa = orc_index_for(20), pa_for(20)
    rng = np.random.default_rng(77)
    length = 150
    base = util.sample_reads(rng, fixture_fasta[1], 1024, length, p_sub=0.0, mix=(1.0, 0.0, 0.0))
    reads = []
    for i, r in enumerate(base):
        b = bytearray(r.encode() if isinstance(r, str) else r)
        for j in range(6):                       # every byte value appears 24 times over the set
            b[int(rng.integers(0, length))] = (6 * i + j) % 256
        if i % 3 == 0:                           # and mixed case
            p = int(rng.integers(0, length - 20))
            b[p:p + 20] = bytes(b[p:p + 20]).lower()
        reads.append(bytes(b))
    want_hits, want_tx, _, _ = _oracle(ix, reads)
    got_hits, got_tx = pa.mapper.map_ascii(reads)                                   # ragged: k_pack_ascii
    _assert_same(reads, got_hits, got_tx, want_hits, want_tx)
    flat = np.frombuffer(b"".join(reads), dtype=np.uint8)
    tile = np.concatenate([flat, np.zeros(64, np.uint8)])                             # 128 x 150 bytes: tile kernel
    got_hits, got_tx = pa.mapper.map_ascii_fixed(tile, len(reads), length)
    _assert_same(reads, got_hits, got_tx, want_hits, want_tx)
    stride = length + 3                                                              # odd stride: k_pack_ascii_fixed
    padded = np.zeros(len(reads) * stride + 64, np.uint8)
    padded[:len(reads) * stride].reshape(len(reads), stride)[:, :length] = flat.reshape(len(reads), length)
    got_hits, got_tx = pa.mapper.map_ascii_fixed(padded, len(reads), length, stride=stride)
    _assert_same(reads, got_hits, got_tx, want_hits, want_tx)


def _oracle_novel_table(hits, tx):
    """[(members tuple, count)] sorted by (length, contents) of the oracle's results that are no visited class."""
    tab = {}
    for h in hits[(hits["flags"] & 1).astype(bool) & (hits["eq_id"] == 0xFFFFFFFF)]:
        key = tuple(int(x) for x in tx[int(h["tx_off"]):int(h["tx_off"]) + int(h["n_tx"])])
        tab[key] = tab.get(key, 0) + 1
    return sorted(tab.items(), key=lambda kv: (len(kv[0]), kv[0]))


@pytest.mark.parametrize("table_cap", ["", "16"])
def test_novel_sets(fixture_fasta, monkeypatch, table_cap):
    """Sets that are no index class (ref src/pseudoaligner.rs:323-356 returns the set itself): one entry per distinct
    set with its read count, ids by (length, contents) order, equal to the oracle's whatever the kernel split, the
    chunking of the call or the table's initial size (16 entries: it must grow several times); counts[n_eq] is the
    table's sum; a reset clears it."""
    if table_cap:
        monkeypatch.setenv("PSA_NOVEL_TABLE_CAP", table_cap)
    rng = np.random.default_rng(17)
    seqs = list(fixture_fasta[1][:500])
    ix = orc.OrcIndex.build(seqs, 20)
    reads = util.sample_reads(rng, seqs, 6000, 150, p_sub=0.02, mix=(0.6, 0.4, 0.0))     # chimeric + noisy: many novel sets
    want_hits, want_tx, want_counts, _ = _oracle(ix, reads)
    want_tab = _oracle_novel_table(want_hits, want_tx)
    assert len(want_tab) > 50 and any(len(m) == 0 for m, _ in want_tab) and any(c > 1 for _, c in want_tab)
    pa = pkg.Pseudoaligner(ix.flat(), device=0, chunk_reads=700)
    n_eq = pa.index.n_eq
    for probes, scan in ((None, None), (0, 8), (64, 0), (3, 8)):
        if probes is not None:
            pa.mapper.set_fast_path(probes, 32)
            pa.mapper.set_scan_width(scan)
        pa.mapper.counts_reset()
        got_hits, got_tx = pa.mapper.map_ascii(reads)
        _assert_same(reads, got_hits, got_tx, want_hits, want_tx)
        tab = pa.mapper.novel_sets()
        assert tab == want_tab
        counts = pa.mapper.counts()
        assert np.array_equal(counts, want_counts) and int(counts[n_eq]) == sum(c for _, c in tab)
        # a second call doubles every count; hits-only calls count too
        pa.mapper.map_ascii(reads, want_tx=False)
        assert pa.mapper.novel_sets() == [(m, 2 * c) for m, c in want_tab]
    pa.mapper.counts_reset()
    assert pa.mapper.novel_sets() == []
    pa.close()


def test_compact_results(orc_index_for, fixture_fasta):
    """PSA_RESULT_COMPACT: 8 bytes per read over PCIe (class id | size of a non-class set, coverage, flags) and only the
    members of the non-class sets; psa_expand_compact rebuilds psa_hit + every member on the host from eq_classes.
    Host batches (chunked pipeline), packed fixed-stride input, and a device batch."""
    ix = orc_index_for(20)
    flat = ix.flat()
    pa = pkg.Pseudoaligner(flat, device=0, chunk_reads=900)
    psa = pkg.pseudoaligner
    rng = np.random.default_rng(23)
    reads = util.sample_reads(rng, fixture_fasta[1], 5000, 150, p_sub=0.02, mix=(0.7, 0.25, 0.05)) + ["", "ACGT", "A" * 150]
    want_hits, want_tx, want_counts, _ = _oracle(ix, reads)
    hc, ntx = pa.mapper.map_ascii(reads, compact=True)
    assert hc.dtype == psa.HIT_COMPACT_DTYPE and hc.nbytes == 8 * len(reads)
    n_novel_members = int(want_hits["n_tx"][(want_hits["eq_id"] == 0xFFFFFFFF)].sum())
    assert len(ntx) == n_novel_members and 0 < len(ntx) < len(want_tx) // 4
    got_hits, got_tx = psa.expand_compact(hc, ntx, flat["eq_offsets"], flat["eq_members"])
    _assert_same(reads, got_hits, got_tx, want_hits, want_tx)
    assert np.array_equal(pa.mapper.counts(), want_counts)
    # packed fixed-stride input (DnaString words, map_read's own argument type), compact results
    fixed = [r for r in reads if len(r) == 150]
    words, off, lens = orc.pack_reads(fixed)
    w_hits, w_tx, _, _ = _oracle(ix, fixed)
    hc2, ntx2 = pa.mapper.map_packed_fixed(words, len(fixed), 150, compact=True)
    g2, t2 = psa.expand_compact(hc2, ntx2, flat["eq_offsets"], flat["eq_members"])
    _assert_same(fixed, g2, t2, w_hits, w_tx)
    pa.close()


@pytest.mark.gpu
def test_process_reads_block_pipeline(pa_for, orc_index_for, fixture_fastq, tmp_path, monkeypatch):
    """psa_process_reads' block pipeline (raw FASTQ blocks -> newline index, record table, map, `{:?}` formatting, all on
    the device: csrc/psa_fastq.cuh) against the oracle and against the host parser, line by line: default blocks, 64 KB
    blocks on 4 lanes (records across block boundaries, tails), ids that need escaping, empty sets, a file without a final
    newline, a wrapped record in mid-file (hand-over to the host parser), progress counters."""
    ix, pa = orc_index_for(20), pa_for(20)
    base = list(fixture_fastq)
    ids = [b'plain', b'with"quote', b'back\\slash', b'tab\there', b'ctl\x01\x1f\x7f', b"it's", b'', b'x' * 300]
    recs = []
    for rep in range(6):
        for i, (rid, seq) in enumerate(base):
            rid = rid.encode() if isinstance(rid, str) else rid
            seq = seq.encode() if isinstance(seq, str) else seq
            tag = ids[(i + rep) % len(ids)]
            recs.append((tag + b"#%d/%d" % (rep, i) if tag else (b"" if (i + rep) % 97 == 0 else b"r%d" % i), seq if (i + rep) % 53 else seq[:rep + 3]))
    want_hits, want_tx, _, _ = _oracle(ix, [s for _, s in recs])
    want = [pkg.format_read_data(flag, rid, eq, cov) for (rid, _), (al, flag, eq, cov) in zip(recs, orc.hits_to_tuples(want_hits, want_tx))]
    text = b"".join(b"@" + rid + b" desc\n" + seq + b"\n+\n" + b"I" * max(1, len(seq)) + b"\n" for rid, seq in recs)
    fq = tmp_path / "blocks.fq"
    fq.write_bytes(text)
    out = tmp_path / "out.txt"
    for env in ({}, {"PSA_FQ_BLOCK_BYTES": "65536", "PSA_FQ_TAIL_BYTES": "8192", "PSA_FQ_LANES": "4"}, {"PSA_PROCESS_FAST": "0"}):
        for k in ("PSA_FQ_BLOCK_BYTES", "PSA_FQ_TAIL_BYTES", "PSA_FQ_LANES", "PSA_PROCESS_FAST"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        st = pkg.process_reads_file(str(fq), pa, str(out), num_threads=4)
        got = out.read_bytes().decode().splitlines()
        assert len(got) == len(want), env
        bad = [i for i in range(len(want)) if got[i] != want[i]]
        assert not bad, (env, bad[:3], got[bad[0]], want[bad[0]])
        assert st["reads"] == len(want) and st["mapped"] == sum(1 for l in want if l.startswith("(true"))
        assert st["aligned"] == int((want_hits["flags"] & 1).sum())
    monkeypatch.setenv("PSA_FQ_BLOCK_BYTES", "65536")
    monkeypatch.setenv("PSA_FQ_TAIL_BYTES", "8192")
    # no final newline; a wrapped record in the middle of the file
    fq.write_bytes(text[:-1])
    st = pkg.process_reads_file(str(fq), pa, str(out), num_threads=3)
    assert out.read_bytes().decode().splitlines() == want
    half = len(recs) // 2
    cut = sum(len(b"@" + rid + b" desc\n" + seq + b"\n+\n" + b"I" * max(1, len(seq)) + b"\n") for rid, seq in recs[:half])
    rid, seq = recs[half]
    assert len(seq) > 10
    wrapped = b"@" + rid + b"\n" + seq[:7] + b"\n" + seq[7:] + b"\n+\n" + b"I" * 7 + b"\n" + b"I" * (len(seq) - 7) + b"\n"
    rest = cut + len(b"@" + rid + b" desc\n" + seq + b"\n+\n" + b"I" * max(1, len(seq)) + b"\n")
    fq.write_bytes(text[:cut] + wrapped + text[rest:])
    st = pkg.process_reads_file(str(fq), pa, str(out), num_threads=3)
    assert out.read_bytes().decode().splitlines() == want and st["reads"] == len(want)


@pytest.mark.gpu
def test_device_graph_builder_equals_host_builder(fixture_fasta):
    """psa_build_graph_device (csrc/psa_build.cu: k-mer sort, colour interning, unitig compaction on the GPU) returns
    the very arrays of the host builder -- the reference's fixture at the CLI's two k, adversarial transcriptomes
    (shared exons, repeats, poly-A self loop, tandem cycle) at every k-mer width edge, a synthetic GENCODE-shaped
    transcriptome -- and an index created from them maps like the oracle."""
    def same(codes, off, k):
        want, ws = host.build_graph(codes, off, k)
        got, gs = pkg.pseudoaligner.build_graph_device(codes, off, k)
        for key in ("seq_words", "node_start", "node_len", "node_exts", "node_eq", "eq_offsets", "eq_members"):
            assert np.array_equal(np.asarray(want[key]), got[key]), (k, key)
        assert gs["n_kmers"] == ws["n_kmers"] and gs["n_nodes"] == ws["n_nodes"] and gs["n_cycles"] == ws.get("n_cycles", gs["n_cycles"])
        return got, gs
    codes, off = host.encode_transcripts(fixture_fasta[1])
    for k in (20, 64):
        same(codes, off, k)
    cycles = 0
    for k in (5, 19, 31, 32, 33, 47, 64):
        rng = np.random.default_rng(300 + k)
        seqs = util.random_transcriptome(rng, n_genes=8, k=k)
        c2, o2 = host.encode_transcripts(seqs)
        _, gs = same(c2, o2, k)
        cycles += gs["n_cycles"]
    assert cycles > 0                                   # the tandem repeats close cycles: the host-side cut ran
    t = host.Transcriptome.synth(5, 300)
    flat, _ = same(t.codes(), t.tx_off(), 24)
    # degenerate inputs: nothing long enough, no transcripts at all
    got, gs = pkg.pseudoaligner.build_graph_device(np.array([0, 1, 2], np.uint8), np.array([0, 3], np.uint64), 5)
    assert gs["n_nodes"] == 0 and gs["n_kmers"] == 0 and len(got["eq_offsets"]) == 1
    with pytest.raises(pkg.PsaError):
        pkg.pseudoaligner.build_graph_device(np.array([0, 1, 7, 3, 2, 1], np.uint8), np.array([0, 6], np.uint64), 4)   # code > 3
    # the device-built index maps like the oracle
    ix = orc.OrcIndex.from_flat(flat)
    pa = pkg.Pseudoaligner(flat, device=0)
    reads = [r.tobytes() for r in t.reads(9, 0, 3000, 100)[:300000].reshape(3000, 100)]
    want_hits, want_tx, _, _ = _oracle(ix, reads)
    got_hits, got_tx = pa.mapper.map_ascii(reads)
    _assert_same(reads, got_hits, got_tx, want_hits, want_tx)
    pa.close()
    ix.close()


@pytest.mark.gpu
@pytest.mark.parametrize("k", [20, 64])
def test_device_mappability_equals_oracle(pa_for, orc_index_for, fixture_fasta, k):
    names = fixture_fasta[0]
    genes = [n.split("|")[1] if "|" in n else n for n in names]      # gencode headers: tx|gene|... (ref src/utils.rs:119-130)
    ids = {g: i for i, g in enumerate(dict.fromkeys(genes))}
    tx_gene = np.array([ids[g] for g in genes], np.uint32)
    assert 1 < len(ids) < len(names)
    ix, pa = orc_index_for(k), pa_for(k)
    want_tm, want_gm = orc.mappability(ix.flat(), tx_gene)
    for bins in (11, 3):
        w_tm, w_gm = (want_tm, want_gm) if bins == 11 else orc.mappability(ix.flat(), tx_gene, bins=bins)
        tm, gm = pa.index.mappability(tx_gene, bins=bins)
        assert np.array_equal(tm, w_tm) and np.array_equal(gm, w_gm)
    assert want_tm[:, 1:].sum() > 0 and (want_gm[:, 0] >= want_tm[:, 0]).all()   # isoforms share k-mers; genes are coarser
    text = pkg.pseudoaligner.mappability_tsv(names, genes, tm, gm)
    assert text.count("\n") == len(names) + 1
    with pytest.raises(pkg.PsaError):
        pa.index.mappability(tx_gene[:-1])          # a class names a transcript beyond the table
