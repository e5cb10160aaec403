#!/usr/bin/env python
"""bench.py -- reads/s pseudoaligned (150 bp) on N B200s, with roofline and CPU baseline.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torchrun)
    python bench.py --impl reference --gpus N --steps K --warmup W

A step is one pass of the hot path (ASCII -> 2-bit pack, map kernel, result scan + expansion,
per-class counting) over one batch of synthetic 150 bp reads.  The default workload is BASELINE
config 3: reads sampled from a synthetic GENCODE-scale transcriptome (20 000 genes, ~200 k
transcripts, k = 24).  `value` times K steps with the batch resident in HBM; `e2e` times the
same K batches through the public C ABI call (psa_mapper_map) from pinned HOST buffers, H2D and
D2H inside the timed region.  The reference arm times the CPU oracle (C restatement of the
reference's map_read; the Rust crate cannot be built in this image) on all host threads.
"""
import argparse
import gzip
import importlib
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

PKG = "rust-pseudoaligner_b200"
METRIC = "reads/sec pseudoaligned (150 bp)"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=12)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="gencode_synth", choices=["gencode_synth", "gencode_small"])
    ap.add_argument("--genes", type=int, default=20000, help="synthetic transcriptome size (20000 ~ 200k transcripts)")
    ap.add_argument("--k", type=int, default=0, help="k-mer length (default: 24 synthetic, 20 gencode_small)")
    ap.add_argument("--read-len", type=int, default=150)
    ap.add_argument("--reads-per-step", type=int, default=1 << 23)
    ap.add_argument("--distinct-batches", type=int, default=3, help="distinct read batches rotated over the steps")
    ap.add_argument("--gamma", type=float, default=0.0)
    ap.add_argument("--group-width", type=int, default=0, help="lanes per read in the cooperative kernel (8/16/32; 0 = library default)")
    ap.add_argument("--fast-probes", type=int, default=-1, help="seed positions one thread tries before handing the read over (0: no thread-per-read kernel; -1: library default)")
    ap.add_argument("--fast-max-small", type=int, default=32)
    ap.add_argument("--scan-width", type=int, default=-1, help="lanes per read of the seed-scan kernel (0: off; -1: library default)")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="CPU baseline budget")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--value-mappers", type=int, default=1, help="mappers (one stream each, shared index) the kernel-resident arm alternates its steps over")
    ap.add_argument("--e2e-threads", type=int, default=2, help="host threads (one mapper each) of the end-to-end arm")
    ap.add_argument("--chunk-reads", type=int, default=0, help="reads per pipeline chunk of host batches (0 = library default)")
    ap.add_argument("--parity-reads", type=int, default=-1, help="reads of the first batch compared with the oracle, untimed (-1: the CPU baseline's sample at 1 GPU, 200k otherwise; 0: off)")
    ap.add_argument("--verify-reads", type=int, default=1 << 20, help="N > 1: reads per rank of the sharded-vs-single-GPU count check")
    ap.add_argument("--cache-dir", default="/dev/shm")
    ap.add_argument("--host-threads", type=int, default=0)
    return ap.parse_args()


# ------------------------------------------------------------------------------------ workload
def workload_config(a):
    """The workload's definition: the SAME dict in the product arm and in the reference arm (run-specific
    facts -- host cores, build times, the CPU arm's bounded sample -- live under other keys)."""
    return {"workload": workload_name(a), "reads_per_step_per_gpu": a.reads_per_step, "read_len": a.read_len, "k": a.k,
            "genes": a.genes if a.workload == "gencode_synth" else None,
            "read_stream": "seed 3: 90 % transcript reads with 0.5 % substitutions, 5 % chimeric, 5 % random",
            "input": "ASCII reads; each step packs, maps, scans and expands one batch"}


def kernel_source_sha16():
    """Identity of the kernels a measurement belongs to (the GPU box has no .git): sha256 of the CUDA sources."""
    import hashlib
    h = hashlib.sha256()
    d = os.path.join(ROOT, PKG, "csrc")
    for name in ("psa_core.cuh", "psa_thread.cuh", "psa_kernels.cuh"):
        with open(os.path.join(d, name), "rb") as f:
            h.update(f.read())
    return h.hexdigest()[:16]


def workload_name(a):
    if a.workload == "gencode_small":
        return "config2: synthetic %d bp reads vs test/gencode_small.fa index, k=%d" % (a.read_len, a.k)
    return "config3: synthetic %d bp reads vs synthetic GENCODE-scale transcriptome (%d genes), k=%d" % (
        a.read_len, a.genes, a.k)


def load_transcriptome(a, host, threads):
    if a.workload == "gencode_small":
        names, seqs, cur = [], [], []
        with gzip.open(os.path.join(ROOT, "tests", "golden", "gencode_small.fa.gz"), "rt") as f:
            for line in f:
                line = line.rstrip("\r\n")
                if line.startswith(">"):
                    if names:
                        seqs.append("".join(cur).encode())
                    names.append(line[1:])
                    cur = []
                elif line:
                    cur.append(line)
        seqs.append("".join(cur).encode())
        codes, off = host.encode_transcripts(seqs)
        return host.Transcriptome.from_codes(codes, off)
    return host.Transcriptome.synth(2, a.genes, threads=threads)


FLAT_KEYS = ("seq_words", "node_start", "node_len", "node_exts", "node_eq", "eq_offsets", "eq_members")


def load_index(a, host, tr, rank, world, barrier, threads):
    """Host graph build (rank 0), shared with the other ranks through cache_dir."""
    tag = "psa_%s_g%d_k%d" % (a.workload, a.genes if a.workload == "gencode_synth" else 0, a.k)
    d = os.path.join(a.cache_dir, tag)
    done = os.path.join(d, "done")
    t0 = time.time()
    if rank == 0 and not os.path.exists(done):
        flat, stats = host.build_graph(tr.codes(), tr.tx_off(), a.k, threads=threads)
        try:
            os.makedirs(d, exist_ok=True)
            for key in FLAT_KEYS:
                np.save(os.path.join(d, key + ".npy"), flat[key])
            open(done, "w").write(json.dumps(stats))
        except OSError:
            if world > 1:
                raise
        build_s = time.time() - t0
        if world == 1:
            return flat, build_s
    barrier()
    flat = {"k": a.k}
    for key in FLAT_KEYS:
        flat[key] = np.load(os.path.join(d, key + ".npy"), mmap_mode="r")
        flat[key] = np.ascontiguousarray(flat[key])
    return flat, time.time() - t0


# ------------------------------------------------------------------------------------ host placement
def pin_to_gpu_cpus(local_rank, world):
    """Run this rank (its threads and the pinned buffers it allocates from here on) on the CPUs NVML names as local to
    its GPU; with several ranks on one box every rank takes its own contiguous share of them."""
    info = {"applied": False}
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        allowed = sorted(os.sched_getaffinity(0))
        cpus = [c for c in allowed if (words[c // 64] >> (c % 64)) & 1] or allowed
        info["gpu_local_cpus"] = len(cpus)
        try:
            info["numa_node"] = int(pynvml.nvmlDeviceGetNumaNodeId(h))
        except Exception:
            pass
        if world > 1:
            same = [c for c in cpus]
            share = max(1, len(same) // world)
            mine = same[(local_rank * share) % len(same):][:share] or same
            cpus = mine
        os.sched_setaffinity(0, cpus)
        info.update(applied=True, cpus=len(cpus), first_cpu=cpus[0])
    except Exception as e:      # no NVML / no permission: run unpinned and say so
        info["error"] = "%s: %s" % (type(e).__name__, e)
    return info


# ------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi polled every 20 ms from before the warm-up (it takes a while to start); the
    samples that fall inside the timed windows are the ones reported."""
    Q = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        self.windows = []

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "20"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def window(self, t0, t1):
        self.windows.append((t0, t1))

    def stop(self):
        import datetime
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        time.sleep(0.05)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons, power, n_all = [], [], set(), [], 0
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 10:
                continue
            try:
                ts = datetime.datetime.strptime(c[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                vals = (float(c[2]), float(c[3]), float(c[4]))
            except ValueError:
                continue
            n_all += 1
            if self.windows and not any(t0 - 0.02 <= ts <= t1 + 0.02 for t0, t1 in self.windows):
                continue
            sm.append(vals[0]); mx.append(vals[1]); power.append(vals[2])
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[6:10]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons),
                       samples=len(sm), samples_total=n_all, power_w_max=float(max(power)),
                       windows="kernel-resident and end-to-end timed regions")
        return out


# ------------------------------------------------------------------------------------ algorithmic bytes
def algorithmic_bytes(ev, k, hit_bytes=24):
    """Bytes the algorithm needs per DESIGN.md section 'Algorithmic bytes' (sequential-equivalent
    events; speculative probes, sector padding and re-reads are NOT counted)."""
    kb = (2 * k + 7) // 8
    b = 0.0
    b += ev["read_bases"] / 4.0                       # packed read, 2 bit/base
    b += 32.0 * ev["dict_levels"]                     # one dictionary bucket per level probed: four 64-bit entries,
                                                      # each compared with the key's fingerprint (the value is inline)
    b += (8.0 + kb) * ev["verifications"]             # node start + unitig k-mer
    b += 24.0 * ev["node_visits"]                     # start, len, eq, class_len, exts
    b += 0.25 * ev["bases_compared"]                  # unitig side of the compares
    b += 4.0 * ev["edge_jumps"]                       # one successor / predecessor entry
    b += 4.0 * ev["class_members"]                    # class members read by the intersection
    b += hit_bytes * ev["reads"] + 4.0 * ev["out_members"]   # psa_hit + members written
    return b


def sector_bytes(ev, k, hit_bytes=24):
    """S_read of SURVEY 8(d): 32 bytes x the distinct sectors the same sequential-equivalent events touch in
    THIS index layout (one sector per dictionary bucket probed, two per node record -- node + class window; an
    unaligned span of s bases of 2-bit sequence covers s/128 + 1 sectors on average).  A model, reported next
    to the DRAM bytes ncu measured (`traffic`); like `algorithmic_bytes` it ignores speculative probes and caches."""
    sect = 0.0
    sect += ev["reads"] * (ev["read_bases"] / ev["reads"] / 128.0 + 1.0)      # packed read
    sect += ev["dict_levels"]                                              # one bucket per level probed
    sect += ev["verifications"] * (2.0 + k / 128.0 + 1.0)                  # node record (two sectors) + unitig k-mer
    sect += ev["node_visits"] * 2.0                                        # node record: the node + its class window
    sect += ev["bases_compared"] / 128.0 + ev["node_visits"]               # unitig spans of the compares
    sect += ev["reads"] * (hit_bytes / 32.0 + 1.0) + ev["out_members"] / 8.0   # psa_hit, counts[] entry, members
    return 32.0 * sect


# ------------------------------------------------------------------------------------ CPU arm
def cpu_arm(a, tr, flat, steps, warmup, budget_s, threads, keep=False):
    """Times the oracle (oracle/psa_oracle.c, the C restatement of the reference's map_read) on
    `threads` host threads over bounded samples of the workload: every thread runs the reference's
    worker body (ASCII -> DnaString, then map_read; ref src/pseudoaligner.rs:449-451) over its
    share of the sample.  Returns (reads/s, info); keep=True also returns the sample's results."""
    import orc
    t0 = time.time()
    ox = orc.OrcIndex.from_flat(flat)
    index_s = time.time() - t0
    L = a.read_len

    def run(data, n):
        bounds = [n * t // threads for t in range(threads + 1)]
        res = [None] * threads

        def work(t):
            res[t] = ox.map_ascii_fixed(data, n, L, start=bounds[t], stop=bounds[t + 1])
        th = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
        t1 = time.perf_counter()
        for x in th:
            x.start()
        for x in th:
            x.join()
        return time.perf_counter() - t1, res

    # calibrate on a small sample, then size the steps to the budget
    n0 = 2000 * threads
    dt, _ = run(tr.reads(3, 10 ** 9, n0, L, threads=threads), n0)
    rate = n0 / dt
    total_steps = max(1, steps + warmup)
    n = int(max(2000 * threads, min(4e6, rate * budget_s / total_steps)))
    n = min(n, a.reads_per_step)
    data = tr.reads(3, 0, n, L, threads=threads)
    for _ in range(warmup):
        run(data, n)
    t_sum, res = 0.0, None
    for _ in range(steps):
        dt, res = run(data, n)
        t_sum += dt
    value = steps * n / t_sum
    info = {"sample": "%d steps x %d reads (the first reads of the workload's stream; ASCII -> 2-bit packing inside "
                      "the timed region), %d threads" % (steps, n, threads),
            "reads_per_step": n, "ms_per_step": 1e3 * t_sum / steps, "oracle_index_s": index_s}
    if keep:
        hits = np.concatenate([r[0] for r in res])
        base, o = 0, 0
        for r in res:                    # member offsets: per-thread -> per-sample
            hits["tx_off"][o:o + len(r[0])] += np.uint64(base)
            base += len(r[1])
            o += len(r[0])
        info["results"] = (hits, np.concatenate([r[1] for r in res]))
    ox.close()
    return value, info


def oracle_results(a, tr, flat, n, threads):
    """Untimed: the oracle's results for the first n reads of the workload's stream."""
    return oracle_results_from(tr.reads(3, 0, n, a.read_len, threads=threads), n, a.read_len, flat, threads)


def oracle_results_from(data, n, read_len, flat, threads):
    """Untimed: the oracle's results (hits, tx) for n ASCII reads of read_len bytes stored back to back in `data`."""
    import orc
    ox = orc.OrcIndex.from_flat(flat)
    bounds = [n * t // threads for t in range(threads + 1)]
    res = [None] * threads

    def work(t):
        res[t] = ox.map_ascii_fixed(data, n, read_len, start=bounds[t], stop=bounds[t + 1])
    th = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
    for x in th:
        x.start()
    for x in th:
        x.join()
    hits = np.concatenate([r[0] for r in res])
    base, o = 0, 0
    for r in res:
        hits["tx_off"][o:o + len(r[0])] += np.uint64(base)
        base += len(r[1])
        o += len(r[0])
    ox.close()
    return hits, np.concatenate([r[1] for r in res])


def parity_check(batch, want_hits, want_tx, device):
    """GPU results of the first len(want_hits) reads of a device batch (just mapped) against the oracle's:
    per-read equality of every psa_hit field and of the members, plus both checksums."""
    import orc
    n = len(want_hits)
    got_hits = batch.hits.to_numpy(pkg_mod().HIT_DTYPE, n)
    n_tx = int(want_hits["n_tx"].sum())
    got_n_tx = int(got_hits["n_tx"].sum())
    got_tx = batch.tx.to_numpy(np.uint32, max(got_n_tx, 1))[:got_n_tx]
    bad = np.zeros(n, bool)
    for f in ("coverage", "n_tx", "eq_id", "flags", "tx_off"):
        bad |= got_hits[f] != want_hits[f]
    if got_n_tx == n_tx:
        diff = np.flatnonzero(got_tx != want_tx)
        if len(diff):                      # attribute differing members to their reads
            ends = np.cumsum(want_hits["n_tx"].astype(np.int64))
            bad[np.unique(np.searchsorted(ends, diff, side="right"))] = True
    first = int(np.flatnonzero(bad)[0]) if bad.any() else None
    out = {"reads_checked": n, "mismatches": int(bad.sum()),
           "checksum_gpu": "%016x" % batch_checksum(batch, n, 0, device),
           "checksum_cpu": "%016x" % orc.result_checksum(want_hits, want_tx, 0),
           "fields": "coverage, n_tx, eq_id, flags, tx_off, members"}
    if first is not None:
        out["first_mismatch"] = {"read": first, "gpu": [int(x) for x in got_hits[first].tolist()],
                                 "oracle": [int(x) for x in want_hits[first].tolist()]}
    return out


def pkg_mod():
    return importlib.import_module(PKG)


def pack_ascii_batch(ascii_arr, n, L, out_words, threads):
    """ASCII reads -> DnaString words (A0 C1 G2 T3, anything else 0), numpy, `threads` slices in parallel (untimed helper)."""
    nw = (L + 31) // 32
    lut = np.zeros(256, np.uint64)
    for ch, v in ((b"C", 1), (b"G", 2), (b"T", 3), (b"c", 1), (b"g", 2), (b"t", 3)):
        lut[ch[0]] = v
    shifts = (62 - 2 * np.arange(32, dtype=np.uint64)).astype(np.uint64)
    step = 65536

    def work(t):
        for c0 in range(t * step, n, threads * step):
            c1 = min(n, c0 + step)
            codes = np.zeros((c1 - c0, nw * 32), np.uint64)
            codes[:, :L] = lut[ascii_arr[c0 * L:c1 * L].reshape(c1 - c0, L)]
            out_words[c0 * nw:c1 * nw] = np.bitwise_or.reduce(codes.reshape(c1 - c0, nw, 32) << shifts, axis=2).reshape(-1)
    th = [threading.Thread(target=work, args=(t,)) for t in range(max(1, threads))]
    for x in th:
        x.start()
    for x in th:
        x.join()


def pcie_probe(psa, torch, dist, world, nbytes=256 << 20, reps=4):
    """Host -> device and device -> host copy rate of pinned memory on this rank's GPU, all ranks copying at the
    same time: {h2d_gbs, d2h_gbs} = the minimum over the ranks (what every rank can count on) and the sum."""
    pin = psa.PinnedArray(nbytes, np.uint8)
    dev = psa.DeviceBuffer(nbytes)
    lib = psa.lib()
    out = {}
    for name, fn in (("h2d", lambda: lib.psa_memcpy_h2d(dev.ptr, pin.ptr, nbytes)), ("d2h", lambda: lib.psa_memcpy_d2h(pin.ptr, dev.ptr, nbytes))):
        fn()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        gbs = reps * nbytes / (time.perf_counter() - t0) / 1e9
        t = torch.tensor([gbs, -gbs, gbs], device="cuda", dtype=torch.float64)
        if world > 1:
            mx = t.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            sm = t.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
            out[name + "_gbs_min_rank"] = -float(mx[1].item())
            out[name + "_gbs_all_ranks"] = float(sm[2].item())
        else:
            out[name + "_gbs_min_rank"] = gbs
            out[name + "_gbs_all_ranks"] = gbs
    pin.free()
    dev.free()
    return out


def batch_checksum(batch, n, first_index, device):
    import ctypes as C
    psa = pkg_mod().pseudoaligner
    out = C.c_uint64()
    psa._check(psa.lib().psa_result_checksum(int(device), batch.hits.ptr, batch.tx.ptr, int(n), int(first_index), C.byref(out)))
    return int(out.value)


# ------------------------------------------------------------------------------------ main
def main():
    # the contract is ONE JSON line on stdout: libraries (NCCL prints its version) write to fd 1 too, so
    # everything but the final line is sent to stderr
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    a = parse_args()
    if a.k == 0:
        a.k = 24 if a.workload == "gencode_synth" else 20
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    ncores = os.cpu_count() or 1
    host_threads = a.host_threads or max(1, ncores // max(1, world))

    if a.impl == "reference":
        if rank != 0:
            return 0
        host = importlib.import_module(PKG + ".host")
        tr = load_transcriptome(a, host, ncores)
        flat, _ = load_index(a, host, tr, 0, 1, lambda: None, ncores)
        threads = a.host_threads or ncores
        value, info = cpu_arm(a, tr, flat, a.steps, a.warmup, max(a.cpu_seconds, 20.0), threads)
        line = {
            "impl": "reference", "metric": METRIC, "value": value, "unit": "reads/s", "n_gpus": a.gpus,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": info["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": workload_config(a),
            "run": {"sample_reads_per_step": info["reads_per_step"], "host_cores": ncores,
                    "note": "C restatement of the reference's map_read (oracle/psa_oracle.c): the Rust crate and its "
                            "debruijn/boomphf dependencies cannot be built in this image; each step is a bounded sample "
                            "of the workload's step"},
            "cpu_baseline": {"value": value, "unit": "reads/s", "cores": threads, "kind": "port", "sample": info["sample"]},
            "e2e": {"value": value, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
        real_stdout.write(json.dumps(line) + "\n")
        real_stdout.flush()
        return 0

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        real_stdout.write(json.dumps({"error": "no CUDA device: the hot path has no CPU fallback"}) + "\n")
        return 2
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()

    pkg = importlib.import_module(PKG)
    host = importlib.import_module(PKG + ".host")
    psa = pkg.pseudoaligner

    t_setup = time.time()
    tr = load_transcriptome(a, host, host_threads)
    # the graph (k-mer sort, colour interning, unitig compaction) is built on this rank's GPU: psa_build_graph_device
    t_b = time.time()
    flat, graph_stats = psa.build_graph_device(tr.codes(), tr.tx_off(), a.k, device=local_rank)
    build_s = time.time() - t_b
    index = pkg.Index(flat, device=local_rank, gamma=a.gamma)
    info = index.info()
    # from here on (pinned batch buffers, mapper threads) the rank stays on the CPUs local to its GPU
    affinity = pin_to_gpu_cpus(local_rank, world)
    if affinity.get("applied") and world > 1:
        host_threads = max(1, min(host_threads, affinity["cpus"]))
    mapper = pkg.Mapper(index, a.chunk_reads)
    if a.group_width:
        mapper.set_group_width(a.group_width)
    if a.fast_probes >= 0:
        mapper.set_fast_path(a.fast_probes, a.fast_max_small)
    if a.scan_width >= 0:
        mapper.set_scan_width(a.scan_width)
    stream = torch.cuda.ExternalStream(mapper.stream(), device=local_rank)

    R, L, G = a.reads_per_step, a.read_len, max(1, a.distinct_batches)
    tx_cap = 24 * R
    # distinct batches of this rank's share of the read stream: host (pinned) and device copies
    shard = importlib.import_module(PKG + ".shard")
    shard_lo, shard_hi = shard.shard_range(world * G * R, world, rank)   # weak scaling: G*R reads per rank
    assert shard_hi - shard_lo == G * R
    host_batches, dev_batches = [], []
    for g in range(G):
        pin = psa.PinnedArray(R * L + 64, np.uint8)
        tr.reads(3, shard_lo + g * R, R, L, out=pin.array, threads=host_threads)
        host_batches.append(pin)
        dev_batches.append(pkg.DeviceBatch(psa.READS_ASCII, pin.array, R, stride=L, fixed_len=L, tx_cap=tx_cap))
    setup_s = time.time() - t_setup

    # events of one batch (untimed): the algorithmic work the roofline is computed from
    ev_split = mapper.map_device_events(dev_batches[0], split=True)
    ev = {key: sum(part[key] for part in ev_split) for key in ev_split[0]}
    deferred_by = mapper.defer_reasons()
    mapper.counts_reset()
    a_bytes_per_read = algorithmic_bytes(ev, a.k) / ev["reads"]

    comm = None
    if world > 1:
        uid = [pkg.Comm.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        comm = pkg.Comm(uid[0], world, rank, local_rank)

    # ---------------- kernel-resident arm: `value`
    # M mappers (one stream and one set of work lists each) over the shared index take the steps in turn:
    # with M > 1 the low-occupancy tail kernels of one batch (seed scan, cooperative kernel, scan, expand)
    # run under the next batch's pack / thread-per-read kernels.  Every mapper has its own device batches.
    M = max(1, a.value_mappers)
    vmappers = [mapper]
    vbatches = [dev_batches]
    for _ in range(1, M):
        vm = pkg.Mapper(index, a.chunk_reads)
        if a.group_width:
            vm.set_group_width(a.group_width)
        if a.fast_probes >= 0:
            vm.set_fast_path(a.fast_probes, a.fast_max_small)
        if a.scan_width >= 0:
            vm.set_scan_width(a.scan_width)
        vmappers.append(vm)
        vbatches.append([pkg.DeviceBatch(psa.READS_ASCII, host_batches[g].array, R, stride=L, fixed_len=L, tx_cap=tx_cap)
                         for g in range(G)])
    vstreams = [stream] + [torch.cuda.ExternalStream(vm.stream(), device=local_rank) for vm in vmappers[1:]]

    def run_steps(first, count):
        for s in range(first, first + count):
            vmappers[s % M].map_device_async(vbatches[s % M][s % G])

    sampler = ClockSampler(local_rank)
    sampler.start()
    for attempt in range(6):       # (an undersized internal buffer is grown at the sync that reports it: warm up again)
        try:
            run_steps(0, max(a.warmup, M))
            for vm in vmappers:
                vm.sync()
            break
        except pkg.PsaError as e:
            if e.code != psa.ERR_CAPACITY or attempt == 5:
                raise
            for vm in vmappers:
                vm.counts_reset()
    for vm in vmappers:
        if comm is not None:
            vm.counts_allreduce(comm)      # warm-up of the collective too (NCCL connects lazily on first use)
        vm.counts_reset()
    launches0 = sum(vm.launch_count() for vm in vmappers)
    if M == 1:
        mapper.profile_enable(True)
        mapper.profile_read()
    barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_w0 = time.time()
    e0.record(stream)
    for st in vstreams[1:]:
        st.wait_event(e0)
    run_steps(0, a.steps)
    if comm is not None:
        for vm in vmappers:
            vm.sync()
            vm.counts_allreduce(comm)      # the path's one collective: per-class counts summed over NVLink
    for st in vstreams[1:]:
        done = torch.cuda.Event()
        done.record(st)
        stream.wait_event(done)
    e1.record(stream)
    for vm in vmappers:
        vm.sync()
    torch.cuda.synchronize()
    sampler.window(t_w0, time.time())
    barrier()
    ms = e0.elapsed_time(e1)
    launches = sum(vm.launch_count() for vm in vmappers) - launches0
    counts = sum(vm.counts() for vm in vmappers)
    prof_ms = ms
    if M == 1:
        prof = mapper.profile_read()
        mapper.profile_enable(False)
    else:
        # per-kernel durations need kernels that run alone: the same K steps once more on one mapper
        for vm in vmappers[1:]:
            vm.close()
        for bl in vbatches[1:]:
            for b in bl:
                b.free()
        mapper.profile_enable(True)
        mapper.profile_read()
        torch.cuda.synchronize()
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record(stream)
        for s in range(a.steps):
            mapper.map_device_async(dev_batches[s % G])
        p1.record(stream)
        mapper.sync()
        torch.cuda.synchronize()
        prof_ms = p0.elapsed_time(p1)
        prof = mapper.profile_read()
        mapper.profile_enable(False)
    novel_table = mapper.novel_sets(comm) if M == 1 else []
    total_reads_counted = int(counts.sum())
    expect = a.steps * R * (world if comm is not None else 1)
    assert total_reads_counted == expect, "per-class counts sum to %d, expected %d" % (total_reads_counted, expect)
    # order-independent checksum of (global read index, coverage, flags, eq_id, members) over ALL timed reads: the
    # results of the last pass over every distinct batch are still in HBM; batch g was mapped uses[g] times
    cks = [batch_checksum(dev_batches[g], R, shard_lo + g * R, local_rank) for g in range(min(G, a.steps))]
    uses = [len(range(g, a.steps, G)) for g in range(min(G, a.steps))]
    checksum_timed = sum(c * u for c, u in zip(cks, uses)) & 0xFFFFFFFFFFFFFFFF
    if world > 1:
        ct = torch.tensor([checksum_timed - (1 << 64) if checksum_timed >= (1 << 63) else checksum_timed], device="cuda", dtype=torch.int64)
        dist.all_reduce(ct, op=dist.ReduceOp.SUM)      # int64 wraps: the sum modulo 2^64
        checksum_timed = int(ct.item()) & 0xFFFFFFFFFFFFFFFF
    ms_t = torch.tensor([ms], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms_t, op=dist.ReduceOp.MAX)
    ms_max = float(ms_t.item())
    value = world * a.steps * R / (ms_max / 1e3)

    # ---------------- end-to-end arms through the public call, host buffers
    e2e = None
    e2e_extra = {}
    if not a.no_e2e:
        # T host threads, one psa_mapper each over the shared index (the reference's process_reads runs
        # num_threads workers over one shared &Pseudoaligner, ref src/pseudoaligner.rs:434-474): while one
        # mapper's pipeline drains its last chunk, the other's is already filling PCIe with its next batch.
        # Results come back compact (PSA_RESULT_COMPACT: 8 bytes per read + the members of the sets that are no
        # index class); the host holds eq_classes and expands class ids itself when it needs the members.
        T = max(1, a.e2e_threads)
        novel_cap = 4 * R
        workers = [(mapper, psa.PinnedArray(R, psa.HIT_COMPACT_DTYPE), psa.PinnedArray(novel_cap, np.uint32))]
        for _ in range(1, T):
            workers.append((pkg.Mapper(index, a.chunk_reads), psa.PinnedArray(R, psa.HIT_COMPACT_DTYPE),
                            psa.PinnedArray(novel_cap, np.uint32)))
        for w_m, _, _ in workers[1:]:
            if a.group_width:
                w_m.set_group_width(a.group_width)
            if a.fast_probes >= 0:
                w_m.set_fast_path(a.fast_probes, a.fast_max_small)
            if a.scan_width >= 0:
                w_m.set_scan_width(a.scan_width)
        # the same batches as DnaString words (map_read's own argument type, ref src/pseudoaligner.rs:381), packed untimed
        nw = (L + 31) // 32
        packed_batches = []
        for g in range(G):
            pw = psa.PinnedArray(R * nw + 8, np.uint64)
            pack_ascii_batch(host_batches[g].array, R, L, pw.array, host_threads)
            packed_batches.append(pw)

        def e2e_step(t, g, packed):
            w_m, w_h, w_t = workers[t]
            if packed:
                return w_m.map_packed_fixed(packed_batches[g].array, R, L, tx_cap=novel_cap, hits=w_h.array, tx=w_t.array, compact=True)
            return w_m.map_ascii_fixed(host_batches[g].array, R, L, tx_cap=novel_cap, hits=w_h.array, tx=w_t.array, compact=True)

        def timed_arm(packed):
            used_by = [0] * T

            def run(t, steps):
                for s in steps:
                    h, tx = e2e_step(t, s % G, packed)
                    used_by[t] += len(tx)

            for t in range(T):          # untimed: first use of every mapper's staging buffers
                e2e_step(t, 0, packed)
                e2e_step(t, 1 % G, packed)
            barrier()
            torch.cuda.synchronize()
            t_w0 = time.time()
            t0 = time.perf_counter()
            th = [threading.Thread(target=run, args=(t, range(t, a.steps, T))) for t in range(T)]
            for x in th:
                x.start()
            for x in th:
                x.join()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            sampler.window(t_w0, time.time())
            dt_t = torch.tensor([dt], device="cuda", dtype=torch.float64)
            if world > 1:
                dist.all_reduce(dt_t, op=dist.ReduceOp.MAX)
            dt = float(dt_t.item())
            h2d = R * (nw * 8 if packed else L)
            d2h = R * psa.HIT_COMPACT_DTYPE.itemsize + 4 * sum(used_by) // a.steps + 16 * ((R + (1 << 19) - 1) >> 19)
            return {"value": world * a.steps * R / dt, "unit": "reads/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "h2d_bytes_per_read": h2d / R, "d2h_bytes_per_read": d2h / R,
                    "ms_per_step": 1e3 * dt / a.steps, "host_threads": T}

        e2e = timed_arm(False)
        e2e["api"] = ("psa_mapper_map (host ASCII batch, the record.seq() bytes of ref src/pseudoaligner.rs:449 -> "
                      "psa_hit_compact[] + members of non-class sets), one mapper per host thread")
        ep = timed_arm(True)
        ep["api"] = "the same call with DnaString words in (map_read's own argument type, packed untimed)"
        e2e_extra["e2e_packed_input"] = ep
        # the per-read results of the last end-to-end step, expanded on the host, against the resident arm's
        w_m, w_h, w_t = workers[0]
        hc, ntx = e2e_step(0, 0, True)
        hx, tx_full = psa.expand_compact(hc[:100000], ntx, flat["eq_offsets"], flat["eq_members"])
        e2e_extra["e2e_expand_check"] = {"reads": 100000, "host_expanded_members": int(len(tx_full))}
        # what the host <-> device link gives this rank while every rank copies at once (pinned memory, 256 MB)
        e2e_extra["pcie_probe"] = pcie_probe(psa, torch, dist, world)
        for w_m, w_h, w_t in workers[1:]:
            w_m.close()

    clocks = sampler.stop()

    # ---------------- roofline of the dominant kernel
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
    else:
        peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
    # the map step is two kernels: k_map_thread (one thread per read) and k_map (cooperative, the reads
    # k_map_thread handed over); each is charged the algorithmic bytes of the reads it completed
    kernels = {}
    for name, part in (("k_map_thread", ev_split[0]), ("k_map", ev_split[1]), ("k_seed_scan", ev_split[2])):
        k_ms, k_n = prof[name]
        if not k_n:
            continue
        per_step_ms = k_ms / a.steps                     # per STEP: k_map_thread runs twice per step (second pass: seeded reads)
        a_bytes = algorithmic_bytes(part, a.k)          # of one batch = one step
        kernels[name] = {"ms_per_step": per_step_ms, "launches_per_step": k_n / a.steps,
                         "share_of_step": k_ms / prof_ms if prof_ms else None,
                         "reads_per_step": part["reads"], "algorithmic_bytes_per_step": a_bytes,
                         "achieved_gbs": a_bytes / (per_step_ms / 1e3) / 1e9}
    dom = max(kernels, key=lambda kname: kernels[kname]["ms_per_step"])
    achieved = kernels[dom]["achieved_gbs"]
    # DRAM bytes per step of the dominant kernel from the committed ncu capture -- only if it was taken from
    # these very kernel sources (the file is stamped with their hash), else the figure is stale and dropped
    traffic, traffic_note = None, "no ncu capture committed for this kernel"
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        if tj.get("kernel_source_sha16") == kernel_source_sha16():
            traffic = tj.get("%s:%s" % (a.workload, dom))
            traffic_note = tj.get("note")
        else:
            traffic_note = "stale: profiles/traffic.json was captured from other kernel sources (%s, now %s)" % (
                tj.get("kernel_source_sha16"), kernel_source_sha16())
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_note": traffic_note, "kernel_source_sha16": kernel_source_sha16(),
                # companion figures: the same kernel time against the DRAM bytes ncu counted
                "traffic_gbs": (traffic / (kernels[dom]["ms_per_step"] / 1e3) / 1e9) if traffic else None,
                "traffic_frac": (traffic / (kernels[dom]["ms_per_step"] / 1e3) / 1e9 / peak) if traffic else None,
                "traffic_over_algorithmic": (traffic / kernels[dom]["algorithmic_bytes_per_step"]) if traffic else None,
                "kernel": dom, "kernel_ms_per_step": kernels[dom]["ms_per_step"],
                "kernel_share_of_step": kernels[dom]["share_of_step"], "kernels": kernels,
                "algorithmic_bytes_per_read": a_bytes_per_read,
                "sector_model_bytes_per_read": sector_bytes(ev, a.k) / ev["reads"],
                "map_step_achieved_gbs": a_bytes_per_read * R / (sum(v["ms_per_step"] for v in kernels.values()) / 1e3) / 1e9,
                "peak_source": peak_src, "handed_over_by_k_map_thread": deferred_by,
                "kernel_timing": ("CUDA events around every map kernel, inside the timed region" if M == 1 else
                                  "CUDA events around every map kernel over the same %d steps run once more on ONE mapper "
                                  "(%.3f ms per step): in the timed region the kernels of %d mappers overlap" % (a.steps, prof_ms / a.steps, M)),
                "note": "dependent random 32-byte-sector gathers: see DESIGN.md for the sector-rate view"}

    cpu = None
    want = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        threads = a.host_threads or ncores
        v, ci = cpu_arm(a, tr, flat, 3, 1, a.cpu_seconds, threads, keep=True)
        want = ci.pop("results")
        cpu = {"value": v, "unit": "reads/s", "cores": threads, "kind": "port", "sample": ci["sample"],
               "note": "oracle/psa_oracle.c, the C restatement of the reference's map_read (Rust toolchain absent); "
                       "like the GPU arm's `value` it includes the ASCII -> 2-bit packing"}
        if threads > 1:     # the same port on ONE host thread, next to the all-threads figure
            v1, ci1 = cpu_arm(a, tr, flat, 1, 0, min(a.cpu_seconds, 3.0), 1)
            cpu["one_thread"] = {"value": v1, "unit": "reads/s", "sample": ci1["sample"]}

    # ---------------- parity, measured: the oracle's results for the first reads of the stream (the CPU baseline's
    # sample when it ran, else a smaller untimed sample) against the GPU's for the same reads of batch 0
    parity = {"novel_sets": {"distinct": len(novel_table), "reads": sum(c for _, c in novel_table),
                             "note": "eq_classes that are no index class, counted per distinct set over all timed reads%s; ids n_eq + rank "
                                     "by (length, contents)" % (" of all ranks (ncclAllGather + merge)" if comm else "")},
              "checksum_all_timed_reads": "%016x" % checksum_timed,
              "checksum_of": "sum over reads of a hash chain over (global read index, coverage, flags, eq_id, members)"}
    if rank == 0 and a.parity_reads != 0:
        if want is None or (0 < a.parity_reads < len(want[0])):
            n_par = min(R, a.parity_reads if a.parity_reads > 0 else 200000)
            want = oracle_results(a, tr, flat, n_par, a.host_threads or ncores)
        mapper.counts_reset()
        mapper.map_device(dev_batches[0])
        parity.update(parity_check(dev_batches[0], want[0], want[1], local_rank))
        parity["reads"] = "reads [0, %d) of the read stream = the first reads of rank 0's first batch, %s index" % (
            len(want[0]), workload_name(a))
    # N > 1 (BASELINE config 4's criterion): the sharded run's all-reduced counts equal a single GPU's counts of the same reads
    if world > 1 and a.verify_reads > 0:
        V = min(a.verify_reads, R)
        vbase = 1 << 40
        vdata = tr.reads(3, vbase + rank * V, V, L, threads=host_threads)
        vb = pkg.DeviceBatch(psa.READS_ASCII, vdata, V, stride=L, fixed_len=L, tx_cap=24 * V)
        mapper.counts_reset()
        mapper.map_device(vb)
        mapper.counts_allreduce(comm)
        sharded = mapper.counts()
        vb.free()
        if rank == 0:
            mapper.counts_reset()
            for r2 in range(world):
                d2 = tr.reads(3, vbase + r2 * V, V, L, threads=host_threads)
                b2 = pkg.DeviceBatch(psa.READS_ASCII, d2, V, stride=L, fixed_len=L, tx_cap=24 * V)
                mapper.map_device(b2)
                b2.free()
            single = mapper.counts()
            parity["multi_gpu"] = {"reads": world * V, "classes": int(len(single)),
                                   "sharded_allreduced_counts_equal_single_gpu": bool(np.array_equal(sharded, single)),
                                   "count_mismatches": int((sharded != single).sum())}
        barrier()

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "reads/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms_max / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u64", "data": "synthetic",
            "config": workload_config(a),
            "run": {"mappers": M,
                    "l2": "inputs larger than L2 (%.0f MB per batch, %d distinct batches rotated; index %.0f MB)" % (
                        R * L / 1e6, G, (info["dict_bytes"] + info["node_bytes"] + info["seq_bytes"] + info["eq_bytes"]) / 1e6),
                    "index": {key: int(info[key]) for key in ("n_nodes", "n_kmers", "n_eq", "n_eq_members",
                                                              "dict_levels", "dict_bytes", "fp_bits", "max_class_len")},
                    "collective": "ncclAllReduce(uint64 counts[n_eq+2]) once, inside the timed region" if comm else "none (1 GPU)",
                    "host_cores": ncores, "host_affinity": affinity, "index_build_s": build_s,
                    "index_built_by": "psa_build_graph_device (graph) + psa_index_create (dictionary), both on the GPU", "setup_s": setup_s},
            "parity": parity,
            "clocks": clocks, "e2e": e2e, **e2e_extra, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu,
            "events_per_read": {key: ev[key] / ev["reads"] for key in ev if key != "reads"},
        }
        real_stdout.write(json.dumps(line) + "\n")
        real_stdout.flush()
    failed = rank == 0 and (parity.get("mismatches", 0) != 0 or parity.get("checksum_gpu") != parity.get("checksum_cpu")
                            or parity.get("multi_gpu", {}).get("count_mismatches", 0) != 0)
    if failed:
        sys.stderr.write("PARITY FAILURE: %s\n" % json.dumps(parity))
    if comm is not None:
        comm.close()
    if world > 1:
        dist.destroy_process_group()
    return 3 if failed else 0


if __name__ == "__main__":
    sys.exit(main())
