"""Host logic of psa_process_reads (csrc/process_reads.cpp: FASTQ reader, record table, batching across
text blocks, ordered `{:?}` line writer -- the mirror of ref src/pseudoaligner.rs:420-514) without a GPU:
the file is linked against a stand-in mapper (tests/hostsim/process_stub.cpp) that derives a fake result
from each read's bytes, so every record must reach the mapper intact and come out in input order.  The
real mapper behind the same driver is covered by the `-m gpu` test test_process_reads_native."""
import ctypes as C
import gzip
import os
import subprocess

import numpy as np
import pytest

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostsim")


class _Stats(C.Structure):
    _fields_ = [("reads", C.c_uint64), ("mapped", C.c_uint64), ("aligned", C.c_uint64), ("seconds", C.c_double),
                ("reader_seconds", C.c_double), ("mapper_seconds", C.c_double), ("writer_seconds", C.c_double)]


@pytest.fixture(scope="module")
def stub():
    subprocess.check_call(["make", "-C", _DIR, "libprocess_stub.so"], stdout=subprocess.DEVNULL)
    L = C.CDLL(os.path.join(_DIR, "libprocess_stub.so"))
    L.psa_process_reads.restype = C.c_int
    L.psa_process_reads.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_uint32, C.c_uint64, C.c_int, C.POINTER(_Stats)]
    return L


@pytest.fixture(autouse=True, params=["blocks", "blocks_tiny", "host_parser"])
def pipeline(request, monkeypatch):
    """Every test runs three times: through the block pipeline whose text work belongs to the device (here: the stub's
    serial stand-in built from the same psa_fastq.cuh routines) with default and with 4 KB blocks on 4 lanes (block
    boundaries inside records, tails, hand-over to the host parser in mid-file), and through the host parser alone."""
    if request.param == "host_parser":
        monkeypatch.setenv("PSA_PROCESS_FAST", "0")
    elif request.param == "blocks_tiny":
        monkeypatch.setenv("PSA_FQ_BLOCK_BYTES", "4096")
        monkeypatch.setenv("PSA_FQ_TAIL_BYTES", "4096")
        monkeypatch.setenv("PSA_FQ_LANES", "4")
    return request.param


def _expected(records):
    out = []
    for rid, seq in records:
        flag = "true" if seq[:1] == b"T" else "false"
        esc = rid.decode().replace("\\", "\\\\").replace('"', '\\"').replace("\t", "\\t")
        members = "[%d]" % sum(seq) if seq else "[]"
        out.append('(%s, "%s", %s, %d)' % (flag, esc, members, len(seq)))
    return out


def _fastq(records, eol=b"\n", comment=b""):
    return b"".join(b"@" + rid + comment + eol + seq + eol + b"+" + eol + b"I" * len(seq) + eol for rid, seq in records)


def _run(stub, path, out, threads=2, batch=0):
    st = _Stats()
    rc = stub.psa_process_reads(C.c_void_p(1), str(path).encode(), str(out).encode(), threads, batch, 0, C.byref(st))
    return rc, st, open(out).read().splitlines()


def _records(rng, n, lmin=1, lmax=200):
    recs = []
    for i in range(n):
        L = int(rng.integers(lmin, lmax + 1))
        seq = bytes(rng.choice(np.frombuffer(b"ACGTN", np.uint8), L).tolist())
        recs.append((b"read%d/%d" % (i, L), seq))
    return recs


@pytest.mark.parametrize("batch,threads", [(0, 1), (7, 3), (1000, 2), (1, 2)])
def test_records_reach_the_mapper_in_order(stub, tmp_path, batch, threads):
    rng = np.random.default_rng(batch + threads)
    recs = _records(rng, 2500 if batch != 1 else 40)
    p = tmp_path / "a.fq"
    p.write_bytes(_fastq(recs, comment=b" some comment\tmore"))
    rc, st, lines = _run(stub, p, tmp_path / "out.txt", threads, batch)
    assert rc == 0 and lines == _expected(recs)
    assert st.reads == len(recs) and st.aligned == len(recs) and st.mapped == sum(1 for _, s in recs if s[:1] == b"T")


def test_gzip_crlf_blank_tail_and_no_final_newline(stub, tmp_path):
    rng = np.random.default_rng(3)
    recs = _records(rng, 300)
    want = _expected(recs)
    gz = tmp_path / "a.fq.gz"
    with gzip.open(gz, "wb") as f:
        f.write(_fastq(recs))
    assert _run(stub, gz, tmp_path / "o1", 2, 64)[2] == want
    crlf = tmp_path / "crlf.fq"
    crlf.write_bytes(_fastq(recs, eol=b"\r\n"))
    assert _run(stub, crlf, tmp_path / "o2", 2, 77)[2] == want
    # blank lines after the last record: bio's reader takes one for a header without '@' (Error::MissingAt, the
    # reference panics) -- an error here too, after every record was processed
    blank = tmp_path / "blank.fq"
    blank.write_bytes(_fastq(recs, eol=b"\r\n") + b"\r\n\r\n\n")
    rc, st, lines = _run(stub, blank, tmp_path / "o2b", 2, 77)
    assert rc == -7 and lines == want
    nonl = tmp_path / "nonl.fq"
    nonl.write_bytes(_fastq(recs)[:-1])                      # last quality line without its newline
    assert _run(stub, nonl, tmp_path / "o3", 1, 50)[2] == want
    empty = tmp_path / "empty.fq"
    empty.write_bytes(b"")
    rc, st, lines = _run(stub, empty, tmp_path / "o4")
    assert rc == 0 and st.reads == 0 and lines == []


def test_quotes_in_ids_and_records_larger_than_a_block(stub, tmp_path):
    recs = [(b'we"ird\\id', b"ACGT"), (b"big", b"ACGT" * 2_000_000), (b"after", b"TTTT")]   # 8 MB record > 4 MB block
    p = tmp_path / "big.fq"
    p.write_bytes(_fastq(recs))
    rc, st, lines = _run(stub, p, tmp_path / "out.txt", 2, 2)
    assert rc == 0 and lines == _expected(recs)


def test_malformed_files_are_errors_after_the_good_records(stub, tmp_path):
    rng = np.random.default_rng(4)
    recs = _records(rng, 100)
    good = _fastq(recs)
    for name, tail in (("truncated", b"@broken\nACGT\n"), ("no_at", b"x\nACGT\n+\nIIII\n")):
        p = tmp_path / (name + ".fq")
        p.write_bytes(good + tail + (_fastq(recs[:3]) if name != "truncated" else b""))
        rc, st, lines = _run(stub, p, tmp_path / "out.txt", 2, 30)
        assert rc == -7, name                                 # PSA_ERR_IO (the reference panics, :446)
        assert lines == _expected(recs), name                 # every complete record before the bad one
    # a record without its '+' line is NOT an error for bio's reader: it takes every line up to the next '+' as
    # sequence (here: its own quality, the next record's header and sequence) and as many lines as quality
    p = tmp_path / "no_plus.fq"
    p.write_bytes(good + b"@x\nACGT\nACGT\nIIII\n" + _fastq(recs[:3]))
    rc, st, lines = _run(stub, p, tmp_path / "out.txt", 2, 30)
    swallowed = b"ACGT" + b"ACGT" + b"IIII" + b"@" + recs[0][0] + recs[0][1]
    assert rc == 0 and lines == _expected(recs + [(b"x", swallowed), recs[2]])
    rc, _, _ = _run(stub, tmp_path / "missing.fq", tmp_path / "out.txt")
    assert rc == -7


def test_streams_pipes_and_append_mode(stub, tmp_path):
    """Regular files are read with pread and written with pwrite by every thread; everything else (FIFOs,
    append-mode descriptors) takes the serial ordered paths.  Same lines either way."""
    import threading
    rng = np.random.default_rng(5)
    recs = _records(rng, 1200)
    want = _expected(recs)
    text = _fastq(recs)
    # input through a FIFO: not seekable, read serially
    fifo_in = tmp_path / "in.fifo"
    os.mkfifo(fifo_in)
    feeder = threading.Thread(target=lambda: open(fifo_in, "wb").write(text))
    feeder.start()
    rc, st, lines = _run(stub, fifo_in, tmp_path / "o1", 3, 100)
    feeder.join()
    assert rc == 0 and lines == want
    # output through a FIFO: ordered serial writes
    p = tmp_path / "a.fq"
    p.write_bytes(text)
    fifo_out = tmp_path / "out.fifo"
    os.mkfifo(fifo_out)
    got = []
    drain = threading.Thread(target=lambda: got.append(open(fifo_out, "rb").read()))
    drain.start()
    st = _Stats()
    rc = stub.psa_process_reads(C.c_void_p(1), str(p).encode(), str(fifo_out).encode(), 4, 64, 0, C.byref(st))
    drain.join()
    assert rc == 0 and got[0].decode().splitlines() == want
    # more threads than records, one record per block
    rc, st, lines = _run(stub, p, tmp_path / "o3", 16, 1)
    assert rc == 0 and lines == want


def test_stdout_redirected_to_a_file_keeps_its_position(stub, tmp_path):
    """out_path NULL/"-" = stdout: when it is a regular file the lines go after what is already there, and what
    the process prints afterwards follows them (append mode falls back to ordered serial writes)."""
    import sys
    rng = np.random.default_rng(6)
    recs = _records(rng, 500)
    p = tmp_path / "a.fq"
    p.write_bytes(_fastq(recs))
    code = (
        "import ctypes as C, os, sys\n"
        "L = C.CDLL(%r)\n"
        "L.psa_process_reads.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_uint32, C.c_uint64, C.c_int, C.c_void_p]\n"
        "os.write(1, b'before\\n')\n"
        "rc = L.psa_process_reads(C.c_void_p(1), %r, b'-', 3, 70, 0, None)\n"
        "os.write(1, b'after %%d\\n' %% rc)\n" % (os.path.join(_DIR, "libprocess_stub.so"), str(p).encode()))
    for mode in ("wb", "ab"):
        out = tmp_path / ("stdout_" + mode)
        with open(out, mode) as f:
            subprocess.check_call([sys.executable, "-c", code], stdout=f)
        assert open(out).read().splitlines() == ["before"] + _expected(recs) + ["after 0"], mode


def test_control_characters_in_ids(stub, tmp_path):
    """Rust's {:?} of the id: \\u{..} for control characters, \\\\ and \\" escaped (ref src/pseudoaligner.rs:490)."""
    p = tmp_path / "c.fq"
    p.write_bytes(b"@a\x01b\x7f\\\"z\nTACG\n+\nIIII\n")
    rc, st, lines = _run(stub, p, tmp_path / "out.txt", 2, 0)
    assert rc == 0 and lines == ['(true, "a\\u{1}b\\u{7f}\\\\\\"z", [%d], 4)' % sum(b"TACG")]


def test_numbers_of_every_decimal_length(stub, tmp_path, monkeypatch):
    """Transcript ids from one to ten digits (the formatter writes two digits per division into a
    pre-counted field)."""
    monkeypatch.setenv("PSA_STUB_NTX", "31")
    monkeypatch.setenv("PSA_STUB_SPREAD", "1")
    rng = np.random.default_rng(8)
    recs = _records(rng, 400, lmin=1, lmax=120)
    p = tmp_path / "n.fq"
    p.write_bytes(_fastq(recs))
    rc, st, lines = _run(stub, p, tmp_path / "out.txt", 3, 50)
    assert rc == 0
    want, lens = [], set()
    for rid, seq in recs:
        ids = [((sum(seq) + j) * 2654435761 & 0xFFFFFFFFFFFFFFFF) >> (j % 30) & 0xFFFFFFFF for j in range(31)]
        lens.update(len(str(x)) for x in ids)
        want.append('(%s, "%s", [%s], %d)' % ("true" if seq[:1] == b"T" else "false", rid.decode(),
                                             ", ".join(str(x) for x in ids), len(seq)))
    assert lines == want
    assert lens >= set(range(3, 11))


def test_records_are_cut_as_bio_cuts_them(stub, tmp_path):
    """bio::io::fastq::Reader::read (bio 1.5, the reader behind ref src/pseudoaligner.rs:421-447): the id ends at the
    first SPACE only (a tab-tagged header as `samtools fastq -T` writes keeps its tags), the sequence line loses ALL
    trailing white space (a trailing blank or tab is not a base), sequences may be wrapped over several lines with
    as many quality lines, the separator line may repeat the id."""
    recs = [(b"r1\tBC:Z:ACGT\tXX:i:3", b"TACGTACG"), (b"r2", b"ACGT"), (b"r3", b"GGGGCCCC"), (b"r4", b"T" * 70), (b"r5", b"ACGTN")]
    text = (b"@r1\tBC:Z:ACGT\tXX:i:3 a comment\twith a tab\nTACGTACG \t\n+\nIIIIIIII\n"
            b"@r2  two spaces\nACGT\t\r\n+r2\nIIII\n"
            b"@r3\nGGGG\nCCCC\n+\nIIII\nIIII\n"                        # wrapped: two sequence lines, two quality lines
            + b"@r4\n" + b"\n".join([b"T" * 10] * 7) + b"\n+\n" + b"\n".join([b"I" * 10] * 7) + b"\n"
            + b"@r5\nACGTN\n+\n@@@@@\n")                                   # a quality line that starts with '@'
    p = tmp_path / "bio.fq"
    p.write_bytes(text)
    for threads, batch in ((1, 0), (3, 2), (2, 1)):
        rc, st, lines = _run(stub, p, tmp_path / "out.txt", threads, batch)
        assert rc == 0 and lines == _expected(recs), (threads, batch, lines)
    # wrapped records mixed into a long four-line file, across block boundaries
    rng = np.random.default_rng(9)
    many = _records(rng, 600, lmin=20, lmax=90)
    parts = []
    for i, (rid, seq) in enumerate(many):
        if i % 7 == 3:
            h = len(seq) // 2
            parts.append(b"@" + rid + b"\n" + seq[:h] + b"\n" + seq[h:] + b"\n+\n" + b"I" * h + b"\n" + b"I" * (len(seq) - h) + b"\n")
        else:
            parts.append(_fastq([(rid, seq)]))
    p2 = tmp_path / "mixed.fq"
    p2.write_bytes(b"".join(parts))
    for threads, batch in ((2, 50), (4, 0), (1, 3)):
        rc, st, lines = _run(stub, p2, tmp_path / "out.txt", threads, batch)
        assert rc == 0 and lines == _expected(many), (threads, batch)
    # errors of bio's reader: no quality at all, an empty sequence, a header that is not UTF-8
    for name, tail in (("no_quality", b"@x\nACGT\n+\n\n"), ("empty_record", b"@x\n+\n"), ("bad_utf8", b"@x\xff\nACGT\n+\nIIII\n")):
        q = tmp_path / (name + ".fq")
        q.write_bytes(_fastq(recs[1:2]) + tail)
        rc, st, lines = _run(stub, q, tmp_path / "out.txt", 2, 0)
        assert rc == -7 and lines == _expected(recs[1:2]), name


def test_debug_str_follows_rust(stub):
    """`{:?}` of the id: Rust's escape_debug -- ASCII exact, non-ASCII printable code points as they are, C1 controls,
    soft hyphen, zero-width and combining code points as \\u{..}."""
    stub.psa_debug_str.restype = C.c_int64
    stub.psa_debug_str.argtypes = [C.c_char_p, C.c_uint64, C.c_char_p, C.c_uint64]

    def dbg(b):
        out = C.create_string_buffer(10 * len(b) + 2)
        n = stub.psa_debug_str(b, len(b), out, len(out))
        return out.raw[:n].decode() if n >= 0 else None

    assert dbg(b"abc") == '"abc"' and dbg(b"") == '""'
    assert dbg(b"a'b\"c\\d") == '"a\'b\\"c\\\\d"'                    # ' is not escaped in a str
    assert dbg(b"\x00\t\n\r\x1b\x7f") == '"\\0\\t\\n\\r\\u{1b}\\u{7f}"'
    assert dbg("naïve-ß-日本".encode()) == '"naïve-ß-日本"'
    assert dbg("a\u00adb\u0085c\u200bd\u0301e\ufefff".encode()) == '"a\\u{ad}b\\u{85}c\\u{200b}d\\u{301}e\\u{feff}f"'
    assert dbg(b"\xff\xfe") is None and dbg(b"\xc3") is None                # not UTF-8: the reference panics


def test_block_pipeline_hands_over_in_mid_file(stub, tmp_path, pipeline):
    """Plain four-line records go through the block pipeline; the first record that is not plain (here: a header with a
    non-ASCII character, then a wrapped record, then a record longer than block + tail) hands the rest of the file to the
    host parser.  Lines, order and counters are the same whichever path a record took."""
    rng = np.random.default_rng(21)
    a, b, c = _records(rng, 700, 30, 120), _records(rng, 300, 30, 120), _records(rng, 200, 30, 120)
    uni = ("r\u00e9sum\u00e9".encode(), b"TTGCA")
    text = (_fastq(a) + b"@" + uni[0] + b"\n" + uni[1] + b"\n+\nIIIII\n" + _fastq(b)
            + b"@wrapped\nACGT\nACGT\n+\nIIII\nIIII\n" + _fastq(c) + _fastq([(b"long", b"ACGT" * 5000)]) + _fastq(a[:50]))
    p = tmp_path / "mid.fq"
    p.write_bytes(text)
    want = _expected(a) + ['(true, "r\u00e9sum\u00e9", [%d], 5)' % sum(uni[1])] + _expected(b) + _expected([(b"wrapped", b"ACGTACGT")]) + \
        _expected(c) + _expected([(b"long", b"ACGT" * 5000)]) + _expected(a[:50])
    for threads in (1, 5):
        rc, st, lines = _run(stub, p, tmp_path / "out.txt", threads, 0)
        assert rc == 0 and lines == want
        assert st.reads == len(want) and st.aligned == len(want)
    # a file that does not end in a newline, cut into blocks
    p2 = tmp_path / "nonl.fq"
    p2.write_bytes(_fastq(a)[:-1])
    rc, st, lines = _run(stub, p2, tmp_path / "out.txt", 3, 0)
    assert rc == 0 and lines == _expected(a)
    # block boundary exactly at a record boundary (4096-byte records)
    rec = (b"x" * 10, b"ACGT" * 509)     # '@' + 10 + '\n' + 2036 + '\n' + '+' + '\n' + 2036 + '\n' = 2 * 2036 + 16 = 4088
    pad = [(rec[0], rec[1] + b"AAAA")]    # 2 * 2040 + 16 = 4096 bytes
    assert len(_fastq(pad)) == 4096
    p3 = tmp_path / "aligned.fq"
    p3.write_bytes(_fastq(pad * 9 + a[:10]))
    rc, st, lines = _run(stub, p3, tmp_path / "out.txt", 2, 0)
    assert rc == 0 and lines == _expected(pad * 9 + a[:10])


def test_progress_ticks_are_the_same_on_both_paths(tmp_path):
    """ref src/pseudoaligner.rs:497-504: a line on stderr at every 1 000 000th read with the share of "mapped" reads so
    far.  The block pipeline gets the counts of the reads before each tick from the device; same text as the host parser."""
    import sys
    n = 2_100_000
    rec = b"@r\nTA\n+\nII\n@s\nAC\n+\nII\n@t\nCC\n+\nII\n"
    p = tmp_path / "ticks.fq"
    p.write_bytes(rec * (n // 3))
    code = ("import ctypes as C, sys\n"
            "L = C.CDLL(%r)\n"
            "L.psa_process_reads.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_uint32, C.c_uint64, C.c_int, C.c_void_p]\n"
            "sys.exit(L.psa_process_reads(C.c_void_p(1), %r, b'/dev/null', 4, 0, 1, None))\n"
            % (os.path.join(_DIR, "libprocess_stub.so"), str(p).encode()))
    subprocess.check_call(["make", "-C", _DIR, "libprocess_stub.so"], stdout=subprocess.DEVNULL)
    outs = []
    for env in ({"PSA_PROCESS_FAST": "0"}, {"PSA_FQ_BLOCK_BYTES": str(1 << 20)}, {}):
        r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, **env), capture_output=True)
        assert r.returncode == 0
        outs.append(r.stderr)
    assert outs[0] == outs[1] == outs[2]
    assert outs[0].count(b"Done Mapping") == 2 and b"Done Mapping 2000000 reads w/ Rate: 33.333" in outs[0]
