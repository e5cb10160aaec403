"""Host side of psa_process_reads alone (reader / writer stages), timed against the stand-in mapper of
tests/hostsim (no GPU): python scripts/host_process_bench.py [n_reads] [threads]"""
import ctypes as C, os, subprocess, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
D = os.path.join(ROOT, "tests", "hostsim")
subprocess.check_call(["make", "-C", D, "libprocess_stub.so"], stdout=subprocess.DEVNULL)
L = C.CDLL(os.path.join(D, "libprocess_stub.so"))
class Stats(C.Structure):
    _fields_ = [("reads", C.c_uint64), ("mapped", C.c_uint64), ("aligned", C.c_uint64), ("seconds", C.c_double),
                ("reader_seconds", C.c_double), ("mapper_seconds", C.c_double), ("writer_seconds", C.c_double)]
L.psa_process_reads.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_uint32, C.c_uint64, C.c_int, C.POINTER(Stats)]
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
threads = int(sys.argv[2]) if len(sys.argv) > 2 else (os.cpu_count() or 4)
Lr = 150
rng = np.random.default_rng(1)
rec = np.empty((n, 12 + Lr + 3 + Lr + 1), np.uint8)
ids = np.char.zfill(np.arange(n).astype("U9"), 9).astype("S9").view(np.uint8).reshape(n, 9)
rec[:, 0] = ord("@"); rec[:, 1] = ord("r"); rec[:, 2:11] = ids; rec[:, 11] = ord("\n")
rec[:, 12:12 + Lr] = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, (n, Lr))]
rec[:, 12 + Lr:15 + Lr] = np.frombuffer(b"\n+\n", np.uint8); rec[:, 15 + Lr:15 + 2 * Lr] = ord("I"); rec[:, 15 + 2 * Lr] = 10
path, out = "/dev/shm/psa_hostbench.fq", "/dev/shm/psa_hostbench.out"
rec.tofile(path)
for rep in range(3):
    st = Stats()
    rc = L.psa_process_reads(C.c_void_p(1), path.encode(), out.encode(), threads, 0, 0, C.byref(st))
    assert rc == 0 and st.reads == n, (rc, st.reads)
    print("threads %d: %.2f s %.2f M reads/s | busy reader %.2f mapper(stub) %.2f writer %.2f | out %.2f GB" % (
        threads, st.seconds, n / st.seconds / 1e6, st.reader_seconds, st.mapper_seconds, st.writer_seconds, os.path.getsize(out) / 1e9))
os.unlink(path); os.unlink(out)
