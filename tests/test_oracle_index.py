"""Pins the oracle's index builder on the reference's own fixture and tests
(ref src/build_index.rs:262-368 validate_dbg, :394-409 test_gencode_small_build_20/_64)."""
import numpy as np
import pytest

import orc
import util

# SURVEY.md section 4 anchors: brute-force facts about test/gencode_small.fa
ANCHORS = {20: (1154378, 7404, 12916), 24: (1165762, 6577, 10968), 64: (1211466, 4773, 6577)}


@pytest.mark.parametrize("k", [20, 24, 64])
def test_fixture_anchor_counts(orc_index_for, k):
    ix = orc_index_for(k)
    assert (ix.n_kmers, ix.n_eq, ix.n_nodes) == ANCHORS[k]
    assert ix.n_pure_cycles == 0


@pytest.mark.parametrize("k", [20, 64])
def test_validate_dbg_part_a_exhaustive(orc_index_for, fixture_fasta, k):
    """validate_dbg (a): every k-mer's colour == eq_classes[data(node(dbg_index[kmer]))]."""
    codes = [orc.encode(s) for s in fixture_fasta[1]]
    cut = util.check_index_against_transcripts(orc_index_for(k).flat(), codes)
    assert cut == 0


@pytest.mark.parametrize("k", [20, 64])
def test_validate_dbg_part_b(orc_index_for, fixture_fasta, k):
    """validate_dbg (b), ref src/build_index.rs:300-367: every transcript maps to itself
    with full coverage; singleton classes are exactly [i]; otherwise the transcript is in
    the class and (unless shortest / identical pair) its node set is a subset of every
    other member's node set."""
    ix = orc_index_for(k)
    seqs = fixture_fasta[1]
    node_sets = {}

    def nodes_of(i):
        if i not in node_sets:
            node_sets[i] = set(ix.map_read(seqs[i], want_nodes=True)[2])
        return node_sets[i]

    n_checked = 0
    for i, s in enumerate(seqs):
        if len(s) < k:
            continue
        eq, cov = ix.map_read(s)
        assert cov == len(s)
        if len(eq) > 1:
            assert i in eq
            if len(eq) == 2 and seqs[eq[0]] == seqs[eq[1]]:
                continue
            shortest = min(len(seqs[x]) for x in eq)
            if len(s) != shortest:
                mine = nodes_of(i)
                for j in eq:
                    assert mine <= nodes_of(j)
        else:
            assert eq == [i]
        n_checked += 1
    assert n_checked > 1500
    if k == 64:
        assert sum(1 for s in seqs if len(s) < 64) == 5  # SURVEY 8(c) pin 1


def test_edges_are_consistent(orc_index_for):
    """succ/pred resolve every ext bit and are mutually inverse (debruijn find_link)."""
    ix = orc_index_for(20)
    flat = ix.flat()
    succ, pred = ix.edges()
    exts = flat["node_exts"]
    for b in range(4):
        has_r = (exts >> b) & 1
        has_l = (exts >> (4 + b)) & 1
        assert np.array_equal(succ[:, b] != orc.EQ_NONE, has_r.astype(bool))
        assert np.array_equal(pred[:, b] != orc.EQ_NONE, has_l.astype(bool))
    # u -> v via base b  implies  v has a left ext leading back to u
    us, bs = np.nonzero(succ != orc.EQ_NONE)
    vs = succ[us, bs]
    back = pred[vs]
    assert (back == us[:, None].astype(np.uint32)).any(axis=1).all()


@pytest.mark.parametrize("k", [5, 20, 31, 32, 33, 64])
def test_random_transcriptomes_structural(k):
    rng = np.random.default_rng(100 + k)
    for rep in range(3):
        seqs = util.random_transcriptome(rng, n_genes=6, k=k)
        ix = orc.OrcIndex.build(seqs, k)
        cut = util.check_index_against_transcripts(ix.flat(), [orc.encode(s) for s in seqs])
        assert cut == ix.n_pure_cycles + _self_loops(ix)
        for i, s in enumerate(seqs):          # validate_dbg (b) on synthetic input
            if len(s) < k:
                assert ix.map_read(s) is None
                continue
            eq, cov = ix.map_read(s)
            assert cov == len(s) and i in eq


def _self_loops(ix):
    """single-k-mer nodes with a unique link onto themselves count as cut cycles in the checker"""
    flat = ix.flat()
    succ, _ = ix.edges()
    n = 0
    for i in range(ix.n_nodes):
        e = int(flat["node_exts"][i])
        r, l = e & 0xF, e >> 4
        if r and not (r & (r - 1)) and l and not (l & (l - 1)):
            b = r.bit_length() - 1
            if succ[i, b] == i:
                n += 1
    return n - ix.n_pure_cycles if n >= ix.n_pure_cycles else 0


def test_flat_roundtrip(orc_index_for, fixture_fasta):
    ix = orc_index_for(20)
    ix2 = orc.OrcIndex.from_flat(ix.flat())
    assert (ix2.n_kmers, ix2.n_eq, ix2.n_nodes) == (ix.n_kmers, ix.n_eq, ix.n_nodes)
    for s in fixture_fasta[1][:50]:
        assert ix.map_read(s) == ix2.map_read(s)


def test_lookup_membership(orc_index_for, fixture_fasta):
    """dbg_index.get + verification behaves as an exact dictionary (ref src/pseudoaligner.rs:96-107)."""
    ix = orc_index_for(20)
    seqs = [s.decode() for s in fixture_fasta[1]]
    flat = ix.flat()
    s = seqs[0]
    assert len(s) > 200
    node, off = ix.lookup(s[100:120])
    codes = util.unpack_words(flat["seq_words"], 20, int(flat["node_start"][node]) + off)
    assert "".join("ACGT"[c] for c in codes) == s[100:120]
    all_kmers = set()
    for t in seqs[:200]:
        all_kmers.update(t[i:i + 20] for i in range(len(t) - 19))
    rng = np.random.default_rng(7)
    n_absent = 0
    for _ in range(2000):
        km = "".join("ACGT"[c] for c in rng.integers(0, 4, 20))
        if km not in all_kmers and not any(km in t for t in seqs):
            assert ix.lookup(km) is None
            n_absent += 1
    assert n_absent > 1900
    for km in list(all_kmers)[:2000]:
        assert ix.lookup(km) is not None
