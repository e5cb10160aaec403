"""numpy helpers that re-derive index facts independently of both the C oracle and the
product builder (used by validate_dbg-style tests, ref src/build_index.rs:262-368)."""
import numpy as np


def unpack_words(words, n_bases, start=0):
    """DnaString words -> uint8 codes of bases [start, start+n_bases)."""
    idx = np.arange(start, start + n_bases, dtype=np.uint64)
    w = words[(idx >> np.uint64(5)).astype(np.int64)]
    sh = (np.uint64(62) - np.uint64(2) * (idx & np.uint64(31))).astype(np.uint64)
    return ((w >> sh) & np.uint64(3)).astype(np.uint8)


def kmers_of(codes, k):
    """All k-mers of a code array as (hi, lo) uint64 pairs (base 0 most significant)."""
    n = len(codes) - k + 1
    if n <= 0:
        z = np.zeros(0, dtype=np.uint64)
        return z, z
    c = codes.astype(np.uint64)
    hi = np.zeros(n, dtype=np.uint64)
    lo = np.zeros(n, dtype=np.uint64)
    klo = min(k, 32)
    khi = k - klo
    for j in range(khi):
        hi = (hi << np.uint64(2)) | c[j:j + n]
    for j in range(klo):
        lo = (lo << np.uint64(2)) | c[khi + j:khi + j + n]
    return hi, lo


def transcript_kmer_table(seqs_codes, k):
    """Naive (kmer -> colour, exts) table straight from the transcripts.

    Returns dict with per-distinct-kmer arrays sorted by (hi, lo):
      hi, lo, left (4-bit mask), right (4-bit mask), grp_start/grp_end into (tx_sorted)
    colour of kmer i = unique(tx_sorted[grp_start[i]:grp_end[i]]).
    """
    his, los, txs, lefts, rights = [], [], [], [], []
    for t, c in enumerate(seqs_codes):
        n = len(c) - k + 1
        if n <= 0:
            continue
        hi, lo = kmers_of(c, k)
        left = np.zeros(n, dtype=np.uint8)
        right = np.zeros(n, dtype=np.uint8)
        left[1:] = np.uint8(1) << c[:n - 1]
        right[:n - 1] = np.uint8(1) << c[k:k + n - 1]
        his.append(hi); los.append(lo); lefts.append(left); rights.append(right)
        txs.append(np.full(n, t, dtype=np.uint32))
    hi = np.concatenate(his); lo = np.concatenate(los); tx = np.concatenate(txs)
    left = np.concatenate(lefts); right = np.concatenate(rights)
    order = np.lexsort((tx, lo, hi))
    hi, lo, tx, left, right = hi[order], lo[order], tx[order], left[order], right[order]
    new = np.ones(len(hi), dtype=bool)
    new[1:] = (hi[1:] != hi[:-1]) | (lo[1:] != lo[:-1])
    starts = np.flatnonzero(new)
    ends = np.append(starts[1:], len(hi))
    return {
        "hi": hi[starts], "lo": lo[starts],
        "left": np.bitwise_or.reduceat(left, starts), "right": np.bitwise_or.reduceat(right, starts),
        "grp_start": starts, "grp_end": ends, "tx_sorted": tx,
    }


def colour_of(tab, i):
    return np.unique(tab["tx_sorted"][tab["grp_start"][i]:tab["grp_end"][i]])


def node_kmer_table(flat):
    """Every k-mer of every node of a flat index, sorted by (hi, lo):
    dict(hi, lo, node, off)."""
    k = flat["k"]
    his, los, nodes, offs = [], [], [], []
    total = int(flat["node_start"][-1] + flat["node_len"][-1]) if len(flat["node_len"]) else 0
    codes = unpack_words(flat["seq_words"], total)
    for i, (s, l) in enumerate(zip(flat["node_start"].tolist(), flat["node_len"].tolist())):
        hi, lo = kmers_of(codes[s:s + l], k)
        his.append(hi); los.append(lo)
        nodes.append(np.full(len(hi), i, dtype=np.uint32))
        offs.append(np.arange(len(hi), dtype=np.uint32))
    hi = np.concatenate(his); lo = np.concatenate(los)
    node = np.concatenate(nodes); off = np.concatenate(offs)
    order = np.lexsort((lo, hi))
    return {"hi": hi[order], "lo": lo[order], "node": node[order], "off": off[order]}


def popcount4(x):
    x = x.astype(np.uint8)
    return ((x & 1) + ((x >> 1) & 1) + ((x >> 2) & 1) + ((x >> 3) & 1)).astype(np.uint8)


def check_index_against_transcripts(flat, seqs_codes):
    """validate_dbg part (a) (ref src/build_index.rs:263-298) made exhaustive, plus the
    structural facts SURVEY.md section 8(c) says the output depends on:
      1. node k-mers are distinct and are exactly the transcripts' k-mers;
      2. colour(node) equals the naive colour of every one of its k-mers, classes sorted+unique;
      3. node exts = naive left exts of its first k-mer | naive right exts of its last k-mer;
      4. interior links are unique in both directions;
      5. nodes are maximal (an end with a unique same-colour bidirectional link may only
         point back at the node's own first k-mer, i.e. a cut cycle).
    Returns the number of cut cycles found."""
    k = flat["k"]
    nt = node_kmer_table(flat)
    tt = transcript_kmer_table(seqs_codes, k)
    n = len(nt["hi"])
    assert n == len(tt["hi"]), "distinct k-mer count differs"
    assert np.array_equal(nt["hi"], tt["hi"]) and np.array_equal(nt["lo"], tt["lo"]), "k-mer sets differ"
    if n > 1:
        dup = (nt["hi"][1:] == nt["hi"][:-1]) & (nt["lo"][1:] == nt["lo"][:-1])
        assert not dup.any(), "a k-mer occurs in two node positions"

    eq_off, eq_mem = flat["eq_offsets"], flat["eq_members"]
    n_eq = len(eq_off) - 1
    for c in range(n_eq):
        m = eq_mem[int(eq_off[c]):int(eq_off[c + 1])]
        assert len(m) > 0 and (np.diff(m.astype(np.int64)) > 0).all(), "class not sorted/unique"
    # 2. colour agreement: signature per k-mer from the naive table vs class signature
    mixed = (tt["tx_sorted"].astype(np.uint64) + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15)
    mixed ^= mixed >> np.uint64(29)
    # de-duplicate (kmer, tx) repeats before summing
    first = np.ones(len(mixed), dtype=bool)
    grp_id = np.repeat(np.arange(n), tt["grp_end"] - tt["grp_start"])
    first[1:] = (tt["tx_sorted"][1:] != tt["tx_sorted"][:-1]) | (grp_id[1:] != grp_id[:-1])
    contrib = np.where(first, mixed, np.uint64(0))
    naive_sig = np.add.reduceat(contrib, tt["grp_start"])
    naive_cnt = np.add.reduceat(first.astype(np.int64), tt["grp_start"])
    cm = (eq_mem.astype(np.uint64) + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15)
    cm ^= cm >> np.uint64(29)
    cls_sig = np.add.reduceat(cm, eq_off[:-1].astype(np.int64)) if n_eq else np.zeros(0, np.uint64)
    cls_cnt = np.diff(eq_off.astype(np.int64))
    node_eq_of_kmer = flat["node_eq"][nt["node"]]
    assert np.array_equal(naive_sig, cls_sig[node_eq_of_kmer]), "k-mer colour differs from node class"
    assert np.array_equal(naive_cnt, cls_cnt[node_eq_of_kmer]), "k-mer colour size differs"
    # exact check on one representative k-mer per class
    rep = {}
    for i, c in enumerate(node_eq_of_kmer.tolist()):
        if c not in rep:
            rep[c] = i
    assert len(rep) == n_eq, "unused eq class id"
    for c, i in rep.items():
        assert np.array_equal(colour_of(tt, i), eq_mem[int(eq_off[c]):int(eq_off[c + 1])])
    # classes are interned: no two ids with the same member list
    seen = set()
    for c in range(n_eq):
        key = eq_mem[int(eq_off[c]):int(eq_off[c + 1])].tobytes()
        assert key not in seen, "duplicate eq class"
        seen.add(key)

    # 3-5 need per-(node, off) ordering
    order = np.lexsort((nt["off"], nt["node"]))
    node_o, off_o = nt["node"][order], nt["off"][order]
    left_o, right_o = tt["left"][order], tt["right"][order]
    hi_o, lo_o = nt["hi"][order], nt["lo"][order]
    is_first = off_o == 0
    is_last = np.ones(n, dtype=bool)
    is_last[:-1] = node_o[1:] != node_o[:-1]
    exts = flat["node_exts"]
    assert np.array_equal(exts[node_o[is_first]] >> 4, left_o[is_first]), "left exts differ"
    assert np.array_equal(exts[node_o[is_last]] & 0xF, right_o[is_last]), "right exts differ"
    interior_src = ~is_last
    interior_dst = ~is_first
    assert (popcount4(right_o[interior_src]) == 1).all(), "interior k-mer with non-unique right ext"
    assert (popcount4(left_o[interior_dst]) == 1).all(), "interior k-mer with non-unique left ext"

    # 5. maximality at right ends (left ends follow by symmetry of links)
    cut_cycles = 0
    key = {}
    if k <= 32:
        lookup = {int(l): i for i, l in enumerate(nt["lo"].tolist())}
    else:
        lookup = {(int(h), int(l)): i for i, (h, l) in enumerate(zip(nt["hi"].tolist(), nt["lo"].tolist()))}
    mask = (1 << (2 * k)) - 1
    first_kmer_of_node = {}
    for i in np.flatnonzero(is_first).tolist():
        first_kmer_of_node[int(node_o[i])] = (int(hi_o[i]) << 64) | int(lo_o[i])
    # sorted-position of each (node, off) row
    pos_sorted = order
    for i in np.flatnonzero(is_last).tolist():
        r = int(right_o[i])
        if r == 0 or (r & (r - 1)):
            continue
        b = r.bit_length() - 1
        km = ((((int(hi_o[i]) << 64) | int(lo_o[i])) << 2) | b) & mask
        j = lookup[km if k <= 32 else (km >> 64, km & ((1 << 64) - 1))]
        assert (int(tt["right"][j]) | 1) >= 0
        lj = int(tt["left"][j])
        same_colour = flat["node_eq"][nt["node"][j]] == flat["node_eq"][node_o[i]]
        if lj and not (lj & (lj - 1)) and same_colour:
            assert km == first_kmer_of_node[int(node_o[i])], "node is not maximal"
            cut_cycles += 1
    del key, pos_sorted
    return cut_cycles


def canonical_nodes(flat):
    """Order-free description of a flat index: sorted list of
    (unitig ascii-ish bytes, exts, class members bytes)."""
    total = int(flat["node_start"][-1] + flat["node_len"][-1]) if len(flat["node_len"]) else 0
    codes = unpack_words(flat["seq_words"], total)
    eq_off, eq_mem = flat["eq_offsets"], flat["eq_members"]
    out = []
    for s, l, e, c in zip(flat["node_start"].tolist(), flat["node_len"].tolist(),
                          flat["node_exts"].tolist(), flat["node_eq"].tolist()):
        out.append((codes[s:s + l].tobytes(), e, eq_mem[int(eq_off[c]):int(eq_off[c + 1])].tobytes()))
    out.sort()
    return out


def random_transcriptome(rng, n_genes=12, k=20, repeat=True):
    """Small synthetic transcriptome with shared exons (nested colour sets), a repeat
    element, a poly-A self loop and a tandem repeat -- exercises every builder branch."""
    alphabet = np.frombuffer(b"ACGT", dtype=np.uint8)

    def rnd(n):
        return alphabet[rng.integers(0, 4, n)].tobytes()

    rep = rnd(90)
    seqs = []
    for g in range(n_genes):
        exons = [rnd(int(rng.integers(25, 160))) for _ in range(int(rng.integers(2, 7)))]
        for iso in range(int(rng.integers(1, 5))):
            keep = [e for e in exons if rng.random() < 0.75] or [exons[0]]
            s = b"".join(keep)
            if repeat and rng.random() < 0.25:
                p = int(rng.integers(0, len(s)))
                s = s[:p] + rep + s[p:]
            seqs.append(s)
    seqs.append(rnd(40) + b"A" * (k + 7) + rnd(30))        # poly-A self loop inside a transcript
    seqs.append(b"ACG" * (k + 5))                           # pure tandem repeat (closed cycle)
    seqs.append(rnd(k - 1))                                 # shorter than k: contributes nothing
    seqs.append(seqs[0])                                    # exact duplicate transcript
    seqs.append(rnd(k))                                     # exactly one k-mer
    return seqs


def sample_reads(rng, seqs, n, length, p_sub=0.005, mix=(0.90, 0.05, 0.05), n_rate=0.0):
    """BASELINE config-2 style read mix over ASCII transcripts `seqs` (bytes):
    transcript reads with substitutions / chimeric halves / uniform random.
    Returns list of ASCII str.  (Test-side sampler; the bench uses the counter-based
    generator in the package so CPU and GPU agree without shipping data.)"""
    long_enough = [s for s in seqs if len(s) >= length]
    half = length // 2
    halves_ok = [s for s in seqs if len(s) >= length - half]
    w = np.array([len(s) - length + 1 for s in long_enough], dtype=np.float64)
    w /= w.sum()
    out = []
    kinds = rng.choice(3, size=n, p=np.array(mix) / sum(mix))
    for kind in kinds.tolist():
        if kind == 0 and long_enough:
            s = long_enough[int(rng.choice(len(long_enough), p=w))]
            p = int(rng.integers(0, len(s) - length + 1))
            r = bytearray(s[p:p + length])
        elif kind == 1 and halves_ok:
            parts = []
            for ln in (half, length - half):
                s = halves_ok[int(rng.integers(0, len(halves_ok)))]
                p = int(rng.integers(0, len(s) - ln + 1))
                parts.append(s[p:p + ln])
            r = bytearray(b"".join(parts))
        else:
            r = bytearray(b"ACGT"[c] for c in rng.integers(0, 4, length))
        if kind != 2 and p_sub > 0:
            for i in np.flatnonzero(rng.random(length) < p_sub).tolist():
                r[i] = b"ACGT"[(b"ACGT".index(bytes([r[i]]).upper()) + int(rng.integers(1, 4))) % 4]
        if n_rate > 0:
            for i in np.flatnonzero(rng.random(length) < n_rate).tolist():
                r[i] = ord("N")
        out.append(r.decode())
    return out
