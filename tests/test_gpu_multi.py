"""Multi-GPU paths that need two devices on one box (skipped on a single-GPU box; `gpurun --gpus 2 -- 'python -m pytest
tests/test_gpu_multi.py -m gpu'`): NCCL halo exchange + time-sharded FIR, and the graph-boundary scatter / gather
around the batch-sharded chain. Passed on 2 x B200 at the end of round 1."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _two_gpus():
    import torch
    return torch.cuda.is_available() and torch.cuda.device_count() >= 2


def _worker(rank, world, port, results):
    import torch
    import torch.distributed as dist
    import cyberether_b200 as cb
    from cyberether_b200.blocks import Filter, SpectrumEngine
    from cyberether_b200.sharding import exchange_fir_halo, gather_rows, scatter_rows, shard_bounds
    from cyberether_b200.synthetic import gaussian_cf32, spectral_rows
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        # ---- time-sharded FIR: each rank filters its slab of frames, primed with the halo from the previous rank
        frames, frame_len, taps = 32, 8192, 129
        stream = gaussian_cf32((frames, frame_len), 5)
        begin, end = shard_bounds(frames, world, rank)
        inp = cb.Tensor.from_numpy(stream[begin:end], device=dev, sampleAxis=1, batchAxis=0)
        block = Filter(sampleRate=8e6, bandwidth=1e6, taps=taps)      # modules allocate on the current device (set above)
        assert block.create("f", {"signal": inp}) == cb.Result.SUCCESS, cb.last_error()
        halo = exchange_fir_halo(inp.data, taps)
        assert block.modules["fir"].set_history(halo, frames_before=begin) == cb.Result.SUCCESS
        assert block.compute() == cb.Result.SUCCESS, cb.last_error()
        fir_whole = gather_rows(block.output("buffer").data.reshape(end - begin, -1), frames, dst=0)
        block.destroy()
        # ---- batch-sharded chain with the graph-boundary collectives: rank 0 owns input and result
        rows = 37
        x = torch.from_numpy(spectral_rows(0, rows)).to(dev) if rank == 0 else None
        mine = scatter_rows(x, rows, (4096,), torch.complex64, dev, src=0)
        t = cb.Tensor(mine.contiguous())
        t.set_attribute("sampleAxis", 1)
        t.set_attribute("batchAxis", 0)
        chain = SpectrumEngine(enableScale=True)
        assert chain.create("s", {"buffer": t}) == cb.Result.SUCCESS, cb.last_error()
        assert chain.compute() == cb.Result.SUCCESS, cb.last_error()
        spec_whole = gather_rows(chain.output("buffer").data, rows, dst=0)
        chain.destroy()
        if rank == 0:
            results["fir"] = fir_whole.cpu().numpy()
            results["spec"] = spec_whole.cpu().numpy()
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(not _two_gpus(), reason="needs two GPUs on one box")
def test_two_gpu_halo_fir_and_boundary_collectives():
    import torch
    import torch.multiprocessing as mp
    import cyberether_b200 as cb
    from cyberether_b200.blocks import Filter, SpectrumEngine
    from cyberether_b200.synthetic import gaussian_cf32, spectral_rows
    manager = mp.Manager()
    results = manager.dict()
    mp.spawn(_worker, args=(2, _free_port(), results), nprocs=2, join=True)

    def single(block, port, x):
        inp = cb.Tensor.from_numpy(x, sampleAxis=1, batchAxis=0)
        assert block.create("b", {port: inp}) == cb.Result.SUCCESS, cb.last_error()
        assert block.compute() == cb.Result.SUCCESS, cb.last_error()
        out = block.output("buffer").numpy().copy()
        block.destroy()
        return out

    fir_one = single(Filter(sampleRate=8e6, bandwidth=1e6, taps=129), "signal", gaussian_cf32((32, 8192), 5))
    spec_one = single(SpectrumEngine(enableScale=True), "buffer", spectral_rows(0, 37))
    assert np.array_equal(results["fir"], fir_one.reshape(32, -1))          # sharding in time changes nothing
    assert np.array_equal(results["spec"], spec_one)                         # nor does sharding the batch
