"""shard.py — multi-GPU decomposition of the DAS path.

One process per GPU (torch.distributed; NCCL over NVLink on the box, gloo in the CPU tests).
Two one-step decompositions, both legal because DAS is a plain sum over independent pixels/transmits
(kern/das_spec.m:476,480):

  * pixel sharding (default): rank g beamforms a contiguous slab of the slow image axis (I2 for a 2-D
    scan, I3 for a volume) so the fast axis stays contiguous; inputs replicated; NO collective on the
    data path — slabs are only concatenated if the caller wants the full image on one rank.
  * transmit partition: rank g holds x(:,:,m in M_g) (natural after a tx-sharded greens), computes a
    full-size partial image, then one all-reduce(sum) of the I complex pixels.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np
import torch


def slab(n: int, rank: int, world: int, align: int = 1) -> Tuple[int, int]:
    """(start, count) of rank's contiguous share of n items; boundaries aligned to `align` where possible."""
    if world <= 1:
        return 0, n
    units = -(-n // align)
    lo = (units * rank) // world * align
    hi = (units * (rank + 1)) // world * align
    lo, hi = min(lo, n), min(hi, n)
    if rank == world - 1:
        hi = n
    return lo, hi - lo


def pixel_shard(Pi: np.ndarray, rank: int, world: int, align: int = 32):
    """Slice the 3 x I1 x I2 x I3 pixel grid along its slowest non-singleton axis. Returns (Pi_slab, axis, start, count)."""
    Pi = np.asarray(Pi)
    Pi = Pi.reshape(Pi.shape + (1,) * (4 - Pi.ndim))
    axis = 3 if Pi.shape[3] > 1 else 2
    a = align if axis == 2 else 1
    s, c = slab(Pi.shape[axis], rank, world, a)
    sl = [slice(None)] * 4
    sl[axis] = slice(s, s + c)
    return Pi[tuple(sl)], axis, s, c


def pixel_shard_interleaved(Pi: np.ndarray, rank: int, world: int, group: int = 32):
    """Round-robin pixel sharding: the slow image axis (I2 columns of a 2-D scan, I3 planes of a volume) is cut into groups
    of `group` consecutive indices (one tile width of the staged kernel for columns, single planes for volumes) and rank g
    takes groups g, g + world, g + 2*world, ...  Contiguous slabs give every rank a different mix of cheap (out-of-range,
    shallow) and expensive pixels — 11 % spread in kernel time at 8 GPUs on the headline grid; interleaved groups give
    every rank the same mix.  The kernel takes arbitrary pixel positions, so the rank's grid is simply the concatenation of
    its groups.  Returns (Pi_sub, axis, index) with `index` the positions of the rank's columns/planes in the full grid."""
    Pi = np.asarray(Pi)
    Pi = Pi.reshape(Pi.shape + (1,) * (4 - Pi.ndim))
    axis = 3 if Pi.shape[3] > 1 else 2
    n = Pi.shape[axis]
    g = group if axis == 2 else 1
    if world <= 1:
        return Pi, axis, np.arange(n)
    ngroups = -(-n // g)
    idx = np.concatenate([np.arange(k * g, min((k + 1) * g, n)) for k in range(rank, ngroups, world)] or [np.arange(0)])
    return np.take(Pi, idx, axis=axis), axis, idx


def tx_shard(M: int, rank: int, world: int) -> Tuple[int, int]:
    return slab(M, rank, world, 1)


def allreduce_image(b: torch.Tensor, group=None) -> torch.Tensor:
    """Sum partial images over ranks in place (the only collective of the tx-partition mode)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return b
    buf = torch.view_as_real(b.contiguous()) if b.is_complex() else b.contiguous()
    dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    out = torch.view_as_complex(buf) if b.is_complex() else buf
    if out.data_ptr() != b.data_ptr():
        b.copy_(out)
    return b


def gather_slabs(b_local: torch.Tensor, axis: int, counts, group=None, dst: int = 0):
    """Concatenate pixel slabs on rank `dst` (not on the timed data path; for consumers that want one image)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return b_local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    real = torch.view_as_real(b_local.contiguous()) if b_local.is_complex() else b_local.contiguous()
    shapes = []
    for c in counts:
        s = list(real.shape)
        s[axis] = c
        shapes.append(s)
    outs = [torch.empty(s, dtype=real.dtype, device=real.device) for s in shapes]
    dist.all_gather(outs, real, group=group) if len(set(map(tuple, shapes))) == 1 else _uneven_gather(outs, real, group)
    full = torch.cat(outs, dim=axis)
    return torch.view_as_complex(full) if b_local.is_complex() else full


def _uneven_gather(outs, mine, group):
    import torch.distributed as dist
    rank = dist.get_rank(group)
    for r, o in enumerate(outs):
        if r == rank:
            o.copy_(mine)
        dist.broadcast(o, src=r, group=group)
