"""The N>1 host logic on CPU: world_size-2 (and 3) gloo process groups.

The ranks shard a read set with rust-pseudoaligner_b200/shard.py exactly as bench.py and a
multi-GPU caller do, map their range (the oracle stands in for the GPU mapper here -- this test
is about the sharding and the count reduction, the GPU path has its own parity tests), sum the
per-class counts with the path's one collective, and rank 0 checks counts and gathered per-read
results against a single-process run."""
import importlib
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

import orc
import util

shard = importlib.import_module("rust-pseudoaligner_b200.shard")


def test_shard_ranges_tile_exactly():
    for n in (0, 1, 7, 8, 9, 1000, 10 ** 9 + 7):
        for world in (1, 2, 3, 4, 8):
            edges = [shard.shard_range(n, world, r) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[r][1] == edges[r + 1][0] for r in range(world - 1))
            sizes = [hi - lo for lo, hi in edges]
            assert max(sizes) - min(sizes) <= 1
    for n, world in ((10, 3), (1000, 8), (7, 2)):
        for i in range(n):
            lo, hi = shard.shard_range(n, world, shard.shard_of_read(i, n, world))
            assert lo <= i < hi
    with pytest.raises(ValueError):
        shard.shard_range(10, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, flat, reads, out_path):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ix = orc.OrcIndex.from_flat(flat)
        lo, hi = shard.shard_range(len(reads), world, rank)
        words, off, lens = orc.pack_reads(reads[lo:hi])
        hits, tx, counts, _ = ix.map_batch(words, off, lens, counts=True)
        assert int(counts.sum()) == hi - lo
        shard.allreduce_counts(counts)
        gathered = shard.gather_hits(hits, tx, len(reads), world, rank)
        if rank == 0:
            np.savez(out_path, counts=counts, hits=gathered[0], tx=gathered[1])
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_counts_equal_single_process(tmp_path, orc_index_for, fixture_fasta, world):
    ix = orc_index_for(20)
    rng = np.random.default_rng(40 + world)
    reads = util.sample_reads(rng, fixture_fasta[1], 3001, 150, p_sub=0.01, mix=(0.85, 0.1, 0.05))
    words, off, lens = orc.pack_reads(reads)
    want_hits, want_tx, want_counts, _ = ix.map_batch(words, off, lens, counts=True)
    out = str(tmp_path / "rank0.npz")
    mp.spawn(_worker, args=(world, _free_port(), ix.flat(), reads, out), nprocs=world, join=True)
    got = np.load(out)
    assert np.array_equal(got["counts"], want_counts)
    assert orc.hits_to_tuples(got["hits"], got["tx"]) == orc.hits_to_tuples(want_hits, want_tx)
    assert np.array_equal(got["hits"]["tx_off"], want_hits["tx_off"])


def test_novel_set_tables_merge_like_one_table():
    """psa_novel_sets_merge (host code of libpsa_b200.so): per-rank tables of novel sets merge into the table a
    single rank would have built -- equal sets once, counts added, order by (length, contents), whatever the split."""
    import importlib
    import numpy as np
    psa = importlib.import_module("rust-pseudoaligner_b200").pseudoaligner
    rng = np.random.default_rng(3)
    universe = [tuple(sorted(set(rng.integers(0, 50, int(rng.integers(0, 6))).tolist()))) for _ in range(200)]
    draws = [universe[int(i)] for i in rng.integers(0, len(universe), 5000)]

    def table(sample):
        t = {}
        for s in sample:
            t[s] = t.get(s, 0) + 1
        return sorted(t.items(), key=lambda kv: (len(kv[0]), kv[0]))

    whole = table(draws)
    for world in (1, 2, 3, 8):
        parts = [table(draws[r::world]) for r in range(world)]
        assert psa.novel_sets_merge(parts) == whole
    assert psa.novel_sets_merge([]) == [] and psa.novel_sets_merge([[], []]) == []
    ids = {m: i for i, (m, _) in enumerate(whole)}
    assert ids[()] == 0 and all(len(whole[i][0]) <= len(whole[i + 1][0]) for i in range(len(whole) - 1))
