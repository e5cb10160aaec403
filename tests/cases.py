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


# ------------------------------------------------------------------------------------------------
# Constructed (not sampled) cases for the branches of map_read no reference test reaches
# (SURVEY.md section 3.2 quirk list).  A tiny transcriptome whose graph shape is known by
# construction, reads placed on it by hand, and `describe_walk` to assert -- from the oracle's
# node list and the dictionary -- that the intended branch is the one the read takes.
# ------------------------------------------------------------------------------------------------
def constructed_transcriptome(k=20, seed=4242):
    """T0 = S (600 random bases).  T1..T3 = fresh 40-base heads + S[100:], S[104:], S[109:]: every head
    joins S at a different base, so S's path breaks into the unitigs ..S[:100+k-1], S[100:104+k-1],
    S[104:109+k-1], S[109:..] -- two SHORT nodes in a row with a longer one after them (each also
    changes colour).  T4 = S[300:420] + fresh tail: a branch that ends a unitig at S[420)."""
    rng = np.random.default_rng(seed)
    alphabet = np.frombuffer(b"ACGT", dtype=np.uint8)

    def rnd(n):
        return alphabet[rng.integers(0, 4, n)].tobytes()

    S = rnd(600)

    def head(j):  # 40 fresh bases whose last one differs from S[j-1] (else the junction would sit one base earlier)
        h = bytearray(rnd(40))
        h[-1] = b"ACGT"[(b"ACGT".index(S[j - 1:j]) + 1) % 4]
        return bytes(h)

    tail = bytearray(rnd(60))
    tail[0] = b"ACGT"[(b"ACGT".index(S[420:421]) + 2) % 4]
    seqs = [S, head(100) + S[100:], head(104) + S[104:], head(109) + S[109:], S[300:420] + bytes(tail)]
    return S, seqs


def _sub(read, pos):
    b = bytearray(read)
    for p in pos:
        b[p] = b"ACGT"[(b"ACGT".index(bytes([b[p]])) + 1) % 4]
    return bytes(b)


def constructed_reads(S, k=20, L=150):
    """name -> read (bytes).  Positions are chosen for k = 20, L = 150 (left-extension threshold 30)."""
    assert k == 20 and L == 150
    reads = {}
    # QUIRK-1 + QUIRK-3: the first stride-3 seed is read position 30 = S[109] = offset 0 of its unitig (errors at
    # 19 and 29 knock out the seeds at 0..27); the left walk then goes through the two short predecessors
    reads["quirk1_offset0_left_walk"] = _sub(S[79:79 + L], [19, 29])
    # the same seed position, at unitig offset 2 (read position 30 = S[111]): the ordinary left extension
    reads["left_walk_offset2"] = _sub(S[81:81 + L], [19, 29])
    # left extension that runs into the read's start inside one long node (no predecessor step)
    reads["left_to_read_start"] = _sub(S[200:200 + L], [19, 29])
    # QUIRK-2: the budget is per node -- two errors in each of two consecutive nodes still extend
    reads["two_errors_per_node"] = _sub(S[60:60 + L], [25, 33, 70, 90])
    # > A mismatches inside one unitig: premature break, re-seed (QUIRK-5) into the node already visited
    reads["reseed_into_visited_node"] = _sub(S[150:150 + L], [40, 44, 48])
    # premature break too close to the end for another seed (kmer_pos > last_kmer_pos, ref :287-290)
    reads["break_near_end"] = _sub(S[150:150 + L], [120, 124, 128])
    # read ending exactly at a unitig end (S[420) is where T4 branches off)
    reads["ends_at_unitig_end"] = S[420 - L:420]
    # missing right extension: T4's tail after S[420) continues in T4, the read follows S -- fine; the reverse:
    # a read that follows T0 up to S[420) and then diverges into random bases: no ext for that base -> re-seed fails
    reads["no_right_ext"] = S[420 - 100:420] + _sub(S[420:470], list(range(0, 50, 2)))
    return reads


def describe_walk(ix, read, k=20):
    """Facts about the path the oracle takes for `read`: first seed (pos, node, offset), pushed nodes in push
    order, how many were pushed before the seed's node (= left-extension steps), duplicates among them."""
    first = None
    for p in range(0, len(read) - k + 1, 3):
        hit = ix.lookup(read[p:p + k])
        if hit is not None:
            first = (p, hit[0], hit[1])
            break
    res = ix.map_read(read, want_nodes=True)
    nodes = res[2] if res is not None else []
    left_steps = nodes.index(first[1]) if first is not None and first[1] in nodes else 0
    return {"first_seed": first, "nodes": nodes, "left_steps": left_steps,
            "revisits": len(nodes) - len(set(nodes)), "result": None if res is None else (res[0], res[1])}
