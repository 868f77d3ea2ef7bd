"""Multi-GPU sharding of the hot path: one process per GPU, `torch.distributed` for the plumbing.

The path shards in two natural ways:

* exposures  - independent raw images go to different ranks; no data-path collective at all;
* slices     - the refocus parameters `a` of ONE light field are split over the ranks; every rank holds the
               4-D light field, computes its share of the stack, and the finished slices are exchanged with a
               single all-gather (NCCL over NVLink on GPUs, gloo in the CPU tests).  The contrast bounds come
               from slice index 0 (lfp_refocuser/top_level.py:69), so the rank that owns it broadcasts two
               float64 values before anybody applies the contrast/sRGB epilogue.

Everything in this file is host logic on top of torch.distributed; it runs unchanged on the gloo backend with
CPU tensors, which is how tests/test_dist_gloo.py covers it.
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(n, rank, world):
    """Contiguous, balanced split of range(n): sizes differ by at most one, lower ranks get the extras."""
    base, extra = divmod(int(n), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_exposures(n, rank, world):
    lo, hi = shard_bounds(n, rank, world)
    return list(range(lo, hi))


def shard_slices(a_list, rank, world):
    """(indices, values) of the refocus parameters this rank computes (contiguous block of the stack, so the
    all-gather reassembles the stack in order)."""
    a_list = list(a_list)
    lo, hi = shard_bounds(len(a_list), rank, world)
    return list(range(lo, hi)), a_list[lo:hi]


def owner_of(index, n, world):
    """Rank that owns element `index` under shard_bounds."""
    for r in range(world):
        lo, hi = shard_bounds(n, r, world)
        if lo <= index < hi:
            return r
    raise IndexError(index)


def broadcast_bounds(lohi, src, group=None):
    """Share the two contrast bounds (float64 tensor of 2) computed on the owner of slice 0."""
    dist.broadcast(lohi, src=src, group=group)
    return lohi


def gather_slices(local, n_total, group=None):
    """All-gather per-rank blocks of slices (n_local, S, T, C) into the full (n_total, S, T, C) stack.
    Blocks may differ by one slice; they are padded to the largest block for the collective."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    lo, hi = shard_bounds(n_total, rank, world)
    assert local.shape[0] == hi - lo, "local block does not match shard_bounds"
    n_max = -(-n_total // world)
    tail = tuple(local.shape[1:])
    send = local
    if local.shape[0] != n_max:
        send = torch.zeros((n_max,) + tail, dtype=local.dtype, device=local.device)
        send[:local.shape[0]] = local
    recv = torch.empty((world * n_max,) + tail, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(recv, send.contiguous(), group=group)
    if n_total == world * n_max:
        return recv
    parts = []
    for r in range(world):
        rlo, rhi = shard_bounds(n_total, r, world)
        parts.append(recv[r * n_max:r * n_max + (rhi - rlo)])
    return torch.cat(parts, dim=0)


def sharded_refocus_stack(vp, a_list, group=None, postprocess=True, refi=False, view_zeroed=None):
    """Slice-sharded LfpRefocuser.shift_and_sum on device tensors: every rank passes the same light field
    `vp` (M,M,S,T,3) and gets the full finished stack back.  One all-gather; two tiny broadcasts.
    refi=True: `a_list` are the up-sampled-pixel parameters of the refinement path (range(M*r0, M*r1));
    each rank prefilters the light field itself (shift independent, ~3 ms) and evaluates its own slices."""
    from . import ops
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    a_list = list(a_list)
    _, mine = shard_slices(a_list, rank, world)
    mm = ops.new_minmax()
    if not mine:
        local = torch.empty((0,) + tuple(vp.shape[2:]), dtype=torch.float32, device=vp.device)
    elif refi:
        from .lfp_refocuser.refinement import refocus_refi
        mask = ops.aperture_mask(vp.shape[0]) if view_zeroed is None else view_zeroed
        local = refocus_refi(vp, mine, mask, mm=mm)
    else:
        local = ops.refocus_int(vp, mine, view_zeroed=view_zeroed, mm=mm)
    if postprocess:
        src = owner_of(0, len(a_list), world)
        lohi = torch.empty(2, dtype=torch.float64, device=vp.device)
        if rank == src:
            ops.contrast_bounds(local[0], 0.005, 99.9, lohi=lohi)
        broadcast_bounds(lohi, src, group)
        # Normalizer's falsy-bound fallback needs the min/max of the WHOLE stack (misc/normalizer.py:40-41)
        keys = mm.view(torch.int32).clone()
        kmin = (keys[0:1].to(torch.int64) & 0xFFFFFFFF)
        kmax = (keys[1:2].to(torch.int64) & 0xFFFFFFFF)
        if not mine:
            kmin.fill_(0xFFFFFFFF)
            kmax.fill_(0)
        dist.all_reduce(kmin, op=dist.ReduceOp.MIN, group=group)
        dist.all_reduce(kmax, op=dist.ReduceOp.MAX, group=group)
        both = torch.cat([kmin, kmax])
        mm_all = torch.where(both >= 2 ** 31, both - 2 ** 32, both).to(torch.int32)
        if mine:
            local = ops.contrast_srgb(local, lohi, mm=mm_all)
    return gather_slices(local, len(a_list), group)
