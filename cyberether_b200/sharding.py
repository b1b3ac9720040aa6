"""Batch sharding of the spectral chain across the GPUs of one box.

Every FFT row is independent (the fft transforms along the sample axis only, the window broadcasts over all
other axes, amplitude/range are elementwise), so the outermost non-sample axis is cut into contiguous slabs,
one per rank, with NO data-path collective. torch.distributed is used for the plumbing only: the barrier that
brackets a timed region, the MAX-over-ranks reduction of device times, and an optional gather of results at
the graph boundary for a consumer that wants the whole spectrum on one rank."""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.distributed as dist


def shard_bounds(total_rows: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous slab [begin, end) of `total_rows` owned by `rank`; slabs differ by at most one row."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    base, extra = divmod(total_rows, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def all_shards(total_rows: int, world: int) -> List[Tuple[int, int]]:
    return [shard_bounds(total_rows, world, r) for r in range(world)]


def max_over_ranks(value: float, device=None) -> float:
    """MAX reduction of a per-rank scalar (device time in ms); identity when not distributed."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def scatter_rows(whole, total_rows: int, row_shape, dtype, device, src: int = 0) -> torch.Tensor:
    """Graph-boundary scatter: `src` holds the whole [total_rows, ...] batch (None elsewhere); every rank receives
    its contiguous slab (`shard_bounds`). The counterpart of `gather_rows`; one collective per graph input, none
    inside the data path. Complex tensors travel as (re, im) pairs (NCCL and gloo have no complex types)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return whole
    world, rank = dist.get_world_size(), dist.get_rank()
    bounds = all_shards(total_rows, world)
    largest = max(e - b for b, e in bounds)
    is_complex = dtype.is_complex
    wire_dtype = torch.float32 if dtype == torch.complex64 else (torch.float64 if dtype == torch.complex128 else dtype)
    wire_shape = (largest,) + tuple(row_shape) + ((2,) if is_complex else ())
    recv = torch.empty(wire_shape, dtype=wire_dtype, device=device)
    parts = None
    if rank == src:
        source = torch.view_as_real(whole) if is_complex else whole
        parts = []
        for b, e in bounds:
            part = torch.zeros_like(recv)
            part[: e - b] = source[b:e]
            parts.append(part)
    dist.scatter(recv, parts, src=src)
    begin, end = bounds[rank]
    mine = recv[: end - begin]
    return torch.view_as_complex(mine.contiguous()) if is_complex else mine


def exchange_fir_halo(local_frames: torch.Tensor, taps: int, previous_cycle_tail=None) -> torch.Tensor:
    """Time sharding of the `filter` path: the frames of a stream are cut into contiguous slabs, one per
    rank (`shard_bounds` over the frame axis), and a FIR with `taps` coefficients needs the `taps - 1` input samples
    that precede its slab. One neighbour exchange per cycle: rank r sends the tail of its slab to rank r + 1; rank 0
    takes `previous_cycle_tail` (the tail the LAST rank kept from the previous cycle, None / zeros at stream start).
    Returns the halo [taps - 1] for this rank and leaves this rank's own tail in `.own_tail` of the result for the
    caller to keep (only the last rank's matters). A slab shorter than taps - 1 samples would need samples from TWO
    ranks back (a chain dependency): that is rejected, not zero-padded — cut the stream into longer slabs."""
    need = taps - 1
    flat = local_frames.reshape(-1)
    if flat.numel() < need:
        raise ValueError(f"exchange_fir_halo: a slab of {flat.numel()} samples is shorter than the {need}-sample halo "
                         f"a {taps}-tap filter needs; use fewer ranks or longer frames")
    own_tail = flat[-need:].clone() if need > 0 else flat[:0].clone()
    if previous_cycle_tail is None:
        previous_cycle_tail = torch.zeros(need, dtype=flat.dtype, device=flat.device)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1 or need == 0:
        halo = previous_cycle_tail.clone()
        halo.own_tail = own_tail
        return halo
    world, rank = dist.get_world_size(), dist.get_rank()
    halo = torch.empty(need, dtype=flat.dtype, device=flat.device)
    real_view = lambda t: torch.view_as_real(t) if t.is_complex() else t      # gloo has no complex send/recv
    ops = []
    if rank + 1 < world:
        ops.append(dist.P2POp(dist.isend, real_view(own_tail), rank + 1))
    if rank > 0:
        ops.append(dist.P2POp(dist.irecv, real_view(halo), rank - 1))
    for req in dist.batch_isend_irecv(ops) if ops else []:
        req.wait()
    if rank == 0:
        halo.copy_(previous_cycle_tail)
    halo.own_tail = own_tail
    return halo


def gather_rows(local: torch.Tensor, total_rows: int, dst: int = 0):
    """Graph-boundary gather: concatenates every rank's [rows_r, n] slab on `dst` in rank order (None elsewhere).
    Slabs may differ by one row, so they are padded to the largest slab for the collective; complex tensors travel
    as (re, im) pairs."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    bounds = all_shards(total_rows, world)
    largest = max(e - b for b, e in bounds)
    is_complex = local.is_complex()
    wire = torch.view_as_real(local.contiguous()) if is_complex else local
    padded = wire
    if wire.shape[0] < largest:
        padded = torch.zeros((largest,) + tuple(wire.shape[1:]), dtype=wire.dtype, device=wire.device)
        padded[: wire.shape[0]] = wire
    parts = [torch.empty_like(padded) for _ in range(world)] if rank == dst else None
    dist.gather(padded.contiguous(), parts, dst=dst)
    if rank != dst:
        return None
    whole = torch.cat([p[: e - b] for p, (b, e) in zip(parts, bounds)], dim=0)
    return torch.view_as_complex(whole.contiguous()) if is_complex else whole
