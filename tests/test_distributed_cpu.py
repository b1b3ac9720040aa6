"""N>1 host logic on CPU: world_size-2 gloo run of the batch sharding, the MAX-over-ranks timing reduction and
the graph-boundary gather (the data path itself has no collective)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cyberether_b200.sharding import (all_shards, exchange_fir_halo, gather_rows, max_over_ranks, scatter_rows,
                                      shard_bounds)


def test_shards_partition_the_batch_exactly():
    for total in (1, 7, 8, 65536, 2 ** 20, 1000003):
        for world in (1, 2, 3, 4, 8):
            shards = all_shards(total, world)
            assert shards[0][0] == 0 and shards[-1][1] == total
            assert all(shards[i][1] == shards[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in shards]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total_rows, n, results):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        begin, end = shard_bounds(total_rows, world, rank)
        # each rank "processes" its slab: row r of the result holds r (so order is checkable)
        local = torch.arange(begin, end, dtype=torch.float32)[:, None].repeat(1, n)
        # graph-boundary scatter: rank 0 owns the whole input, every rank must receive exactly its slab
        whole_in = torch.arange(total_rows, dtype=torch.float32)[:, None].repeat(1, n) if rank == 0 else None
        mine = scatter_rows(whole_in, total_rows, (n,), torch.float32, "cpu", src=0)
        assert mine.shape == (end - begin, n) and torch.equal(mine, local)
        # FIR time sharding: each rank must receive the taps-1 samples that precede its slab of the stream
        stream = torch.arange(total_rows * n, dtype=torch.float32).reshape(total_rows, n)
        stream = torch.complex(stream, -stream)
        taps = 7
        halo = exchange_fir_halo(stream[begin:end], taps)
        flat = stream.reshape(-1)
        expect = flat[begin * n - (taps - 1):begin * n] if begin > 0 else torch.zeros(taps - 1, dtype=flat.dtype)
        assert torch.equal(halo, expect)
        assert torch.equal(halo.own_tail, flat[end * n - (taps - 1):end * n])
        # complex slabs travel as (re, im) pairs through the boundary collectives
        cwhole = stream if rank == 0 else None
        cmine = scatter_rows(cwhole, total_rows, (n,), torch.complex64, "cpu", src=0)
        assert cmine.dtype == torch.complex64 and torch.equal(cmine, stream[begin:end])
        cback = gather_rows(cmine, total_rows, dst=0)
        assert (cback is None) if rank else torch.equal(cback, stream)
        slowest = max_over_ranks(10.0 + rank)                 # rank 1 is slower: everyone must see 11.0
        whole = gather_rows(local, total_rows, dst=0)
        dist.barrier()
        results[rank] = (slowest, None if whole is None else whole[:, 0].numpy().copy())
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_sharding_and_gather():
    world, total_rows, n = 2, 11, 4
    manager = mp.Manager()
    results = manager.dict()
    mp.spawn(_worker, args=(world, _free_port(), total_rows, n, results), nprocs=world, join=True)
    assert results[0][0] == 11.0 and results[1][0] == 11.0
    assert results[1][1] is None
    assert np.array_equal(results[0][1], np.arange(total_rows, dtype=np.float32))
