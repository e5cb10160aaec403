#!/bin/bash
# round 2: device graph builder + mappability: parity tests, then the config 3 transcriptome timed (host builder beside it)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider -k "device_graph_builder or mappability" 2>&1 | tail -15
timeout 900 python - <<'PY' 2>&1 | tee gpurun_out/r2_build_device.txt
import importlib, time, numpy as np
pkg = importlib.import_module("rust-pseudoaligner_b200")
host = importlib.import_module("rust-pseudoaligner_b200.host")
psa = pkg.pseudoaligner
tr = host.Transcriptome.synth(2, 20000, threads=16)
codes, off = tr.codes(), tr.tx_off()
print("transcriptome: %d transcripts, %.1f M bases" % (tr.n_tx, tr.n_bases / 1e6))
for rep in range(2):
    t0 = time.time(); flat_d, sd = psa.build_graph_device(codes, off, 24); td = time.time() - t0
    print("device builder: %.2f s  %s" % (td, sd))
t0 = time.time(); flat_h, sh = host.build_graph(codes, off, 24, threads=16); th = time.time() - t0
print("host builder (16 threads): %.2f s  %s" % (th, sh))
same = all(np.array_equal(np.asarray(flat_h[k]), flat_d[k]) for k in ("seq_words", "node_start", "node_len", "node_exts", "node_eq", "eq_offsets", "eq_members"))
print("arrays identical:", same)
ix = pkg.Index(flat_d, device=0)
genes = np.arange(tr.n_tx, dtype=np.uint32) // 10
t0 = time.time(); tm, gm = ix.mappability(genes); print("mappability on the device: %.3f s, k-mers counted %d" % (time.time() - t0, int(tm.sum())))
PY
