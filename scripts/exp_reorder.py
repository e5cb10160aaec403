"""Experiment: renumber the graph's nodes in chain order (follow successor links) so that nodes a read
walks through are neighbours in memory; writes the permuted index into bench.py's cache directory."""
import importlib, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
host = importlib.import_module("rust-pseudoaligner_b200.host")
genes, k = int(sys.argv[1]) if len(sys.argv) > 1 else 20000, 24
tr = host.Transcriptome.synth(2, genes, threads=16)
flat, stats = host.build_graph(tr.codes(), tr.tx_off(), k, threads=16)
t0 = time.time()
n = len(flat["node_len"])
start, ln = flat["node_start"].astype(np.int64), flat["node_len"].astype(np.int64)
words = flat["seq_words"]
nb = int((start + ln).max())
codes = np.zeros(len(words) * 32, np.uint8)
for j in range(32):
    codes[j::32] = (words >> np.uint64(62 - 2 * j)) & np.uint64(3)
codes = codes[:nb]
def kmer_at(pos):  # pos: int64 array -> k-mer integer (k <= 32)
    v = np.zeros(len(pos), np.uint64)
    for t in range(k):
        v = (v << np.uint64(2)) | codes[pos + t].astype(np.uint64)
    return v
first = kmer_at(start)
last = kmer_at(start + ln - k)
order_by_first = np.argsort(first, kind="stable")
sorted_first = first[order_by_first]
mask = np.uint64((1 << (2 * k)) - 1)
succ = np.full((n, 4), -1, np.int64)
exts = flat["node_exts"]
for b in range(4):
    nxt = ((last << np.uint64(2)) | np.uint64(b)) & mask
    idx = np.searchsorted(sorted_first, nxt)
    idx[idx >= n] = n - 1
    ok = (sorted_first[idx] == nxt) & (((exts >> b) & 1) == 1)
    succ[ok, b] = order_by_first[idx[ok]]
# chain order
new_id = np.full(n, -1, np.int64)
order = []
succ_l = succ.tolist()
seen = bytearray(n)
for s in range(n):
    v = s
    while v >= 0 and not seen[v]:
        seen[v] = 1
        order.append(v)
        nv = -1
        for w in succ_l[v]:
            if w >= 0 and not seen[w]:
                nv = w
                break
        v = nv
order = np.array(order, np.int64)
assert len(order) == n
# permute
new_len = ln[order]
new_start = np.zeros(n, np.int64); new_start[1:] = np.cumsum(new_len)[:-1]
src = np.repeat(start[order] - new_start, new_len) + np.arange(int(new_len.sum()))
new_codes = codes[src]
pad = (-len(new_codes)) % 32
nc = np.concatenate([new_codes, np.zeros(pad, np.uint8)]).reshape(-1, 32).astype(np.uint64)
new_words = np.zeros(len(nc), np.uint64)
for j in range(32):
    new_words |= nc[:, j] << np.uint64(62 - 2 * j)
out = dict(flat)
out.update(seq_words=new_words, node_start=new_start.astype(np.uint64), node_len=new_len.astype(np.uint32),
           node_exts=flat["node_exts"][order], node_eq=flat["node_eq"][order])
print("reordered %d nodes in %.1f s" % (n, time.time() - t0))
d = "/dev/shm/psa_gencode_synth_g%d_k%d" % (genes, k)
os.makedirs(d, exist_ok=True)
for key in ("seq_words", "node_start", "node_len", "node_exts", "node_eq", "eq_offsets", "eq_members"):
    np.save(os.path.join(d, key + ".npy"), out[key])
open(os.path.join(d, "done"), "w").write(json.dumps(stats))
