import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def fixture_fasta():
    """test/gencode_small.fa of the reference (gzip of the unmodified file)."""
    import orc
    names, seqs = orc.read_fasta(os.path.join(GOLDEN, "gencode_small.fa.gz"))
    assert len(seqs) == 1832 and sum(map(len, seqs)) == 2836465
    return names, seqs


@pytest.fixture(scope="session")
def fixture_fastq():
    """test/small.fq of the reference (gzip of the unmodified file)."""
    import orc
    recs = orc.read_fastq(os.path.join(GOLDEN, "small.fq.gz"))
    assert len(recs) == 9309
    return recs


_ORC_CACHE = {}


@pytest.fixture(scope="session")
def orc_index_for(fixture_fasta):
    """k -> oracle index of the fixture transcriptome (naive oracle builder), cached."""
    import orc

    def get(k):
        if k not in _ORC_CACHE:
            _ORC_CACHE[k] = orc.OrcIndex.build(fixture_fasta[1], k)
        return _ORC_CACHE[k]

    return get
