"""mappability::analyze_graph (ref src/mappability.rs:120-156): the oracle's restatement on a graph small enough to
count by hand, and the TSV text.  (psa_index_mappability against the oracle: tests/test_gpu_parity.py.)"""
import importlib

import numpy as np
import pytest

import orc

pkg = importlib.import_module("rust-pseudoaligner_b200")


def _tiny():
    # k = 4.  Three transcripts: t0 and t1 share the prefix ACGTACC, t2 is on its own; genes: t0, t1 -> gene 7, t2 -> gene 9
    seqs = [b"ACGTACCGGA", b"ACGTACCTTG", b"GGGTTTCAAC"]
    ix = orc.OrcIndex.build(seqs, 4)
    return ix, seqs, np.array([7, 7, 9], np.uint32)


def test_oracle_counts_by_hand():
    ix, seqs, tx_gene = _tiny()
    flat = ix.flat()
    tm, gm = orc.mappability(flat, tx_gene, bins=11)
    # every k-mer of a transcript is in exactly one unitig of a class that contains the transcript: the row sums are the
    # transcripts' DISTINCT k-mer counts (ref total_kmer_count, :54-56)
    for t, s in enumerate(seqs):
        assert int(tm[t].sum()) == len({s[i:i + 4] for i in range(len(s) - 3)}) == int(gm[t].sum())
    # t0/t1 share ACGT CGTA GTAC TACC (4 k-mers, class {0, 1}: two transcripts, ONE gene); the rest is unique
    assert tm[0].tolist()[:3] == [3, 4, 0] and tm[1].tolist()[:3] == [3, 4, 0] and tm[2].tolist()[:2] == [7, 0]
    assert gm[0].tolist()[:2] == [7, 0] and gm[1].tolist()[:2] == [7, 0] and gm[2].tolist()[:2] == [7, 0]
    # bins: a class with as many or more transcripts than bins lands in the last one (ref :59-65)
    tm2, _ = orc.mappability(flat, tx_gene, bins=2)
    assert tm2[0].tolist() == [3, 4]
    tm1, _ = orc.mappability(flat, tx_gene, bins=1)
    assert tm1[:, 0].tolist() == [7, 7, 7]
    ix.close()


def test_tsv_text_follows_the_reference():
    ix, seqs, tx_gene = _tiny()
    tm, gm = orc.mappability(ix.flat(), tx_gene)
    text = pkg.pseudoaligner.mappability_tsv(["t0", "t1", "t2"], ["g7", "g7", "g9"], tm, gm)
    assert text.splitlines() == ["tx_name\tgene_name\ttx_kmer_count\tfrac_kmer_unique_tx\tfrac_kmer_unique_gene",
                                 "t0\tg7\t7\t0.42857142857142855\t1", "t1\tg7\t7\t0.42857142857142855\t1", "t2\tg9\t7\t1\t1"]
    f = pkg.pseudoaligner._rust_f64
    assert [f(x) for x in (0.5, 1.0, 0.0, 1e-7, float("nan"))] == ["0.5", "1", "0", "0.0000001", "NaN"]   # Rust's `{}` of an f64
    ix.close()
