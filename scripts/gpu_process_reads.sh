# throughput of the native map driver (psa_process_reads): synthetic FASTQ in /dev/shm -> lines
mkdir -p gpurun_out
python - <<'PY' | tee gpurun_out/process_reads.txt
import importlib, os, time, numpy as np
pkg = importlib.import_module("rust-pseudoaligner_b200")
host = importlib.import_module("rust-pseudoaligner_b200.host")
tr = host.Transcriptome.synth(2, 20000, threads=16)
flat, stats = host.build_graph(tr.codes(), tr.tx_off(), 24, threads=16)
n, L = int(os.environ.get('N_READS', 8_000_000)), 150
data = tr.reads(3, 0, n, L, threads=16)[:n * L].reshape(n, L)
t0 = time.time()
rec = np.empty((n, 12 + L + 3 + L + 1), np.uint8)          # "@r%09d \n" seq "\n+\n" qual "\n"
ids = np.empty((n, 9), np.uint8)
v = np.arange(n, dtype=np.int64)
for d in range(9):
    ids[:, 8 - d] = 48 + v % 10
    v //= 10
rec[:, 0] = ord("@"); rec[:, 1] = ord("r"); rec[:, 2:11] = ids; rec[:, 11] = ord("\n")
rec[:, 12:12 + L] = data; rec[:, 12 + L:12 + L + 3] = np.frombuffer(b"\n+\n", np.uint8)
rec[:, 15 + L:15 + 2 * L] = ord("I"); rec[:, 15 + 2 * L] = ord("\n")
path = "/dev/shm/psa_synth.fq"
rec.tofile(path)
print("fastq: %d reads, %.2f GB, written in %.1f s" % (n, rec.nbytes / 1e9, time.time() - t0))
pa = pkg.Pseudoaligner(flat, device=0)
for threads in [int(x) for x in os.environ.get('PR_THREADS', '4,16').split(',')]:
    for rep in range(2):
        st = pkg.process_reads_file(path, pa, "/dev/shm/psa_out.txt", num_threads=threads)
        print("threads %2d: %.2f s  %.1f M reads/s  busy: reader %.2f mapper %.2f writer %.2f s  (reads %d, aligned %d, out %.2f GB)" % (
            threads, st["seconds"], st["reads"] / st["seconds"] / 1e6, st["reader_seconds"], st["mapper_seconds"],
            st["writer_seconds"], st["reads"], st["aligned"], os.path.getsize("/dev/shm/psa_out.txt") / 1e9))
print(open("/dev/shm/psa_out.txt").readline().strip())
os.unlink(path); os.unlink("/dev/shm/psa_out.txt")
PY
