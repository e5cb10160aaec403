"""Read sets shared by the host-simulation tests and the GPU parity tests."""
import numpy as np

import util


def left_extension_reads(rng, seqs, n, length, k):
    """Reads whose first ~length/5+k bases carry substitutions every <k bases, so the first
    seed lands at or after the left-extension threshold (ref src/pseudoaligner.rs:126) and the
    backward walk (QUIRK-1/2/3 of SURVEY.md section 3.2) runs."""
    long_enough = [s for s in seqs if len(s) >= length]
    out = []
    for _ in range(n):
        s = long_enough[int(rng.integers(0, len(long_enough)))]
        p = int(rng.integers(0, len(s) - length + 1))
        r = bytearray(s[p:p + length])
        thr = length // 5
        n_err = int(rng.integers(1, 5))
        pos = sorted(set(int(x) for x in rng.integers(max(0, thr - k), thr + k, n_err) if x < length))
        # knock out every seed before the threshold
        step = max(1, k - int(rng.integers(1, 6)))
        pos += list(range(int(rng.integers(0, step)), thr, step))
        for i in sorted(set(pos)):
            if i < length:
                r[i] = b"ACGT"[(b"ACGT".index(bytes([r[i]]).upper()) + int(rng.integers(1, 4))) % 4]
        out.append(r.decode())
    return out


def read_sets(rng, seqs, length, k, scale=1.0):
    """name -> list of ASCII reads; `seqs` are ASCII transcripts (bytes)."""
    n = lambda x: max(8, int(x * scale))
    sets = {}
    if any(len(s) >= length for s in seqs):
        sets["clean_mix"] = util.sample_reads(rng, seqs, n(1500), length, p_sub=0.005)
        sets["noisy"] = util.sample_reads(rng, seqs, n(1500), length, p_sub=0.04, mix=(0.8, 0.2, 0.0))
        sets["with_N"] = util.sample_reads(rng, seqs, n(300), length, p_sub=0.01, n_rate=0.01)
        sets["left_ext"] = left_extension_reads(rng, seqs, n(600), length, k)
        sets["lower"] = [r.lower() for r in util.sample_reads(rng, seqs, n(50), length, p_sub=0.0)]
    sets["edge"] = ["", "A", "ACGT", "A" * (k - 1), "A" * k, "A" * length, "ACG" * (length // 3 + 1),
                    "T" * length, "N" * length, "C" * k + "N", "G" * (k + 1)]
    return sets
