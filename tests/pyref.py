"""Second, independent restatement of the reference hot path in pure Python (strings and
dicts, base by base) -- used only to cross-check the C oracle on small inputs, so that a
transcription slip in one restatement shows up as a disagreement.

Follows ref src/pseudoaligner.rs:64-418 (commit 9d9cab8) directly from the Rust source.
"""
import numpy as np

import util

READ_COVERAGE_THRESHOLD = 32       # ref src/config.rs:16
LEFT_EXTEND_FRACTION = 0.2         # ref src/config.rs:17
DEFAULT_ALLOWED_MISMATCHES = 2     # ref src/config.rs:18


class PyIndex:
    def __init__(self, flat):
        self.k = int(flat["k"])
        total = int(flat["node_start"][-1] + flat["node_len"][-1]) if len(flat["node_len"]) else 0
        codes = util.unpack_words(flat["seq_words"], total)
        letters = np.frombuffer(b"ACGT", dtype=np.uint8)[codes].tobytes().decode()
        self.seq = [letters[s:s + l] for s, l in zip(flat["node_start"].tolist(), flat["node_len"].tolist())]
        self.exts = flat["node_exts"].tolist()
        self.data = flat["node_eq"].tolist()
        eo, em = flat["eq_offsets"], flat["eq_members"]
        self.eq_classes = [em[int(eo[c]):int(eo[c + 1])].tolist() for c in range(len(eo) - 1)]
        self.kmers = {}
        self.first = {}
        self.last = {}
        for n, s in enumerate(self.seq):
            for o in range(len(s) - self.k + 1):
                self.kmers[s[o:o + self.k]] = (n, o)
            self.first[s[:self.k]] = n
            self.last[s[-self.k:]] = n

    # Exts::has_ext (bit b = right ext base b, bit 4+b = left ext base b; include/psa.h)
    def has_right(self, n, base):
        return (self.exts[n] >> "ACGT".index(base)) & 1

    def has_left(self, n, base):
        return (self.exts[n] >> (4 + "ACGT".index(base))) & 1

    # Node::r_edges / l_edges via debruijn find_link: stranded, by terminal k-mer
    def right_edge(self, n, base):
        return self.first[self.seq[n][-(self.k - 1):] + base] if self.k > 1 else None

    def left_edge(self, n, base):
        return self.last[base + self.seq[n][:self.k - 1]]


def clean_read(read):
    """DnaString::from_dna_string: non-ACGT -> A, case-insensitive (QUIRK-6)."""
    return "".join(c if c in "ACGT" else "A" for c in read.upper())


def map_read_to_nodes(ix, read, allowed_mismatches=DEFAULT_ALLOWED_MISMATCHES):
    read = clean_read(read)
    k = ix.k
    read_length = len(read)
    read_coverage = 0
    nodes = []
    left_extend_threshold = int(LEFT_EXTEND_FRACTION * read_length)
    kmer_pos = 0
    if read_length < k:
        return None
    last_kmer_pos = read_length - k

    def find_kmer_match(pos):
        while pos <= last_kmer_pos:
            hit = ix.kmers.get(read[pos:pos + k])
            if hit is not None:
                return pos, hit
            pos += 3
        return pos, None

    kmer_pos, hit = find_kmer_match(kmer_pos)
    node_id, kmer_offset = hit if hit else (None, None)

    if node_id is not None and kmer_pos >= left_extend_threshold:
        last_pos = kmer_pos - 1
        prev_node_id = node_id
        prev_kmer_offset = kmer_offset - 1 if kmer_offset > 0 else 0
        while True:
            ref = ix.seq[prev_node_id]
            max_matchable_pos = min(last_pos + 1, prev_kmer_offset + 1)
            premature_break = False
            matched_bases = 0
            seen_snp = 0
            for idx in range(max_matchable_pos):
                if ref[prev_kmer_offset - idx] != read[last_pos - idx]:
                    seen_snp += 1
                    if seen_snp > allowed_mismatches:
                        premature_break = True
                        break
                matched_bases += 1
                read_coverage += 1
            if last_pos + 1 - matched_bases == 0 or premature_break:
                break
            last_pos -= matched_bases
            next_base = read[last_pos]
            if ix.has_left(prev_node_id, next_base):
                prev_node_id = ix.left_edge(prev_node_id, next_base)
                prev_kmer_offset = len(ix.seq[prev_node_id]) - k
                nodes.append(prev_node_id)
            else:
                break

    if kmer_pos <= last_kmer_pos:
        while True:
            ref = ix.seq[node_id]
            kmer_pos += k
            read_coverage += k
            nodes.append(node_id)
            remaining_read = read_length - kmer_pos
            ref_offset = kmer_offset + k
            max_matchable_pos = min(remaining_read, len(ref) - ref_offset)
            premature_break = False
            matched_bases = 0
            seen_snp = 0
            for idx in range(max_matchable_pos):
                if ref[ref_offset + idx] != read[kmer_pos + idx]:
                    seen_snp += 1
                    if seen_snp > allowed_mismatches:
                        premature_break = True
                        break
                matched_bases += 1
                read_coverage += 1
            kmer_pos += matched_bases
            if kmer_pos >= read_length:
                break
            next_base = read[kmer_pos]
            if not premature_break and ix.has_right(node_id, next_base):
                node_id = ix.right_edge(node_id, next_base)
                kmer_offset = 0
                kmer_pos -= k - 1
                read_coverage -= k - 1
            else:
                if kmer_pos > last_kmer_pos:
                    break
                kmer_pos, hit = find_kmer_match(kmer_pos)
                if hit is None:
                    break
                node_id, kmer_offset = hit

    if not nodes:
        assert read_coverage == 0
        return None
    return read_coverage, nodes


def intersect(v1, v2):
    """ref :389-418, including the in-place swap/truncate bookkeeping."""
    import bisect
    if not v1:
        return v1
    if not v2:
        del v1[:]
    fill_idx1 = idx1 = idx2 = 0
    while idx1 < len(v1) and idx2 < len(v2):
        x = v1[idx1]
        pos = bisect.bisect_left(v2, x, idx2)
        if pos < len(v2) and v2[pos] == x:
            v1[fill_idx1], v1[idx1] = v1[idx1], v1[fill_idx1]
            fill_idx1 += 1
            idx1 += 1
            idx2 = pos + 1
        else:
            idx1 += 1
            idx2 = pos
    del v1[fill_idx1:]
    return v1


def map_read(ix, read):
    """Pseudoaligner::map_read: None | (eq_class, coverage)."""
    r = map_read_to_nodes(ix, read)
    if r is None:
        return None
    coverage, nodes = r
    nodes = sorted(nodes, key=lambda n: len(ix.eq_classes[ix.data[n]]))   # stable, :331-334
    eq_class = list(ix.eq_classes[ix.data[nodes[0]]])
    for n in nodes[1:]:
        intersect(eq_class, ix.eq_classes[ix.data[n]])
    return eq_class, coverage


def process_read(ix, read):
    """process_reads tuple minus the id: (flag, eq_class, coverage), ref :453-462."""
    r = map_read(ix, read)
    if r is None:
        return False, [], 0
    eq_class, coverage = r
    return (coverage >= READ_COVERAGE_THRESHOLD and not eq_class), eq_class, coverage
