"""Throughput of the drop-in entry psa_process_reads (FASTQ file -> `{:?}` lines file), config 3's index and read stream.
usage: python scripts/process_reads_bench.py [n_reads] [threads]   (env PR_CONFIGS: ';'-separated 'K=V,K=V' settings to compare)"""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("rust-pseudoaligner_b200")
host = importlib.import_module("rust-pseudoaligner_b200.host")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8_000_000
threads = int(sys.argv[2]) if len(sys.argv) > 2 else (os.cpu_count() or 16)
L = 150
tr = host.Transcriptome.synth(2, 20000, threads=threads)
flat, stats = host.build_graph(tr.codes(), tr.tx_off(), 24, threads=threads)
path, outp = "/dev/shm/psa_synth.fq", "/dev/shm/psa_out.txt"
t0 = time.time()
W = 12 + L + 3 + L + 1                                        # "@r%09d\n" seq "\n+\n" qual "\n"
with open(path, "wb") as f:
    for c0 in range(0, n, 4_000_000):                         # written in slices: the array of 32 M records would be 10 GB
        m = min(4_000_000, n - c0)
        data = tr.reads(3, c0, m, L, threads=threads)[:m * L].reshape(m, L)
        rec = np.empty((m, W), np.uint8)
        v = np.arange(c0, c0 + m, dtype=np.int64)
        for d in range(9):
            rec[:, 10 - d] = 48 + v % 10
            v //= 10
        rec[:, 0] = ord("@"); rec[:, 1] = ord("r"); rec[:, 11] = 10
        rec[:, 12:12 + L] = data; rec[:, 12 + L:15 + L] = np.frombuffer(b"\n+\n", np.uint8)
        rec[:, 15 + L:15 + 2 * L] = ord("I"); rec[:, 15 + 2 * L] = 10
        rec.tofile(f)
print("fastq: %d reads, %.2f GB, written in %.1f s; host cpus %d" % (n, n * W / 1e9, time.time() - t0, os.cpu_count()), flush=True)
pa = pkg.Pseudoaligner(flat, device=0)
KEYS = ("PSA_PROCESS_FAST", "PSA_FQ_BLOCK_BYTES", "PSA_FQ_TAIL_BYTES", "PSA_FQ_LANES", "PSA_VERBOSE", "PSA_FQ_WRITE_THREADS", "PSA_OUT_MMAP")
ref_sum = None
for cfg in os.environ.get("PR_CONFIGS", "").split(";"):
    for k in KEYS:
        os.environ.pop(k, None)
    reps = 2
    for kv in filter(None, cfg.split(",")):
        k, v = kv.split("=")
        if k == "REPS": reps = int(v)
        else: os.environ[k] = v
    for rep in range(reps):
        if os.path.exists(outp):
            os.unlink(outp)          # a fresh output file every time (truncating gigabytes of old output is not the driver's work)
        st = pkg.process_reads_file(path, pa, outp, num_threads=threads)
        print("[%-44s] threads %2d: %.3f s  %6.1f M reads/s  busy: reader %.2f mapper %.2f writer %.2f s  (reads %d, aligned %d, out %.2f GB)" % (
            cfg, threads, st["seconds"], st["reads"] / st["seconds"] / 1e6, st["reader_seconds"], st["mapper_seconds"],
            st["writer_seconds"], st["reads"], st["aligned"], os.path.getsize(outp) / 1e9), flush=True)
    # every configuration must write the same bytes
    import hashlib
    h = hashlib.blake2b(digest_size=16)
    with open(outp, "rb") as f:
        while True:
            b = f.read(1 << 24)
            if not b: break
            h.update(b)
    if ref_sum is None: ref_sum = h.hexdigest()
    print("    output digest %s %s" % (h.hexdigest(), "(same)" if h.hexdigest() == ref_sum else "DIFFERENT"), flush=True)
print(open(outp).readline().strip())
# BASELINE.md "C3": the reference's own driver shape (one record per mutex acquisition, bounded channel, serial print) around
# the CPU port of map_read, on a bounded sample of the same file, lines to /dev/null
if os.environ.get("PR_C3", "1") != "0":
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import orc
    m = min(n, int(os.environ.get("PR_C3_READS", 2_000_000)))
    sample = "/dev/shm/psa_synth_head.fq"
    with open(path, "rb") as f, open(sample, "wb") as g:
        g.write(f.read(m * W))
    ox = orc.OrcIndex.from_flat(flat)
    for t in (threads, 1):
        t0 = time.time()
        reads, mapped = ox.process_reads_c3(sample, "/dev/null", num_threads=t)
        dt = time.time() - t0
        print("[C3: reference-shaped CPU driver, %2d worker threads] %.2f s  %.2f M reads/s  (reads %d, mapped %d; first %d records of the file, lines to /dev/null)" % (
            t, dt, reads / dt / 1e6, reads, mapped, m), flush=True)
        m_small = min(m, 300_000)
        if t == threads and m_small < m:
            with open(path, "rb") as f, open(sample, "wb") as g:
                g.write(f.read(m_small * W))
            m = m_small
    os.unlink(sample)
os.unlink(path); os.unlink(outp)
