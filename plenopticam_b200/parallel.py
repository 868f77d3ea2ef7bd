"""Multi-GPU sharding of the hot path: one process per GPU, `torch.distributed` for the plumbing.

The path shards in two natural ways:

* exposures  - independent raw images go to different ranks; no data-path collective at all;
* slices     - the refocus parameters `a` of ONE light field are split over the ranks in contiguous blocks; every rank
               holds the 4-D light field and computes its block of the LINEAR stack.  The exchange of the finished
               slices is the only data-path communication; it comes in two forms:

               transport='p2p'   (the product, NVLink / NVSwitch): every rank owns a peer-mapped stack buffer
                                 (`PeerStack`); the kernel that finishes a slice stores it into the stack of EVERY rank
                                 (pcb_refocus_refi_fanout: compute and exchange in one kernel, the stores overlap the
                                 arithmetic of the tiles still running), followed by one 1-element all-reduce as the
                                 "all writes have landed" barrier;
               transport='nccl'  (baseline; also what the gloo CPU tests drive): one `all_gather_into_tensor` of the
                                 blocks.

               The colour management of LfpRefocuser (lfp_refocuser/top_level.py:68-70) needs the percentile bounds of
               slice 0 and, for Normalizer's falsy-bound fallback, the min / max of the whole stack.  Both are cheap
               (0.1 ms for a 60-slice stack) next to a collective's latency, so every rank derives them itself from
               the gathered linear stack (`finish_stack`) instead of exchanging them: no broadcast, no all-reduce.

`local_block` / `finish_stack` are the per-rank halves without any communication; tests/test_sharded_gpu.py runs them
for several world sizes on one GPU and compares with the unsharded stack.  The host logic (`shard_bounds`,
`gather_slices`) runs unchanged on the gloo backend with CPU tensors (tests/test_dist_gloo.py).
"""
import ctypes as C

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
    exchange reassembles the stack in order)."""
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
    """Share the two contrast bounds (float64 tensor of 2) computed on the owner of slice 0 (used by callers that
    post-process before the exchange; `sharded_refocus_stack` does not need it)."""
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


# ----------------------------------------------------------------------------------------------------------
# per-rank halves (no communication)
# ----------------------------------------------------------------------------------------------------------
ROW_GROUP = 8        # row bands start on multiples of the vertical pass' group size (RF_GROUP)


def row_band(S, rank, world, halo=0):
    """[y_lo, y_hi) of this rank in a ROW-sharded stack: whole groups of ROW_GROUP rows.  `halo` (rows a band needs beyond
    each of its inner edges: the reach of the vertical windows) balances band + halo instead of the band alone: the first
    and the last rank have one inner edge only and get correspondingly taller bands."""
    S, world = int(S), int(world)
    G = -(-S // ROW_GROUP)
    if world == 1:
        return 0, S
    hg = float(halo) / ROW_GROUP
    cost = (G + hg * (2 * world - 2)) / world
    sizes = [max(cost - hg * (1 if r in (0, world - 1) else 2), 0.0) for r in range(world)]
    scale = G / sum(sizes) if sum(sizes) > 0 else 0.0
    edges = [0.0]
    for v in sizes:
        edges.append(edges[-1] + v * scale)
    b = [int(round(e)) for e in edges]
    b[0], b[-1] = 0, G
    for r in range(1, world + 1):                      # monotone
        b[r] = max(b[r], b[r - 1])
    return b[rank] * ROW_GROUP, min(b[rank + 1] * ROW_GROUP, S)


def local_block(vp, a_list, rank, world, refi=False, view_zeroed=None, out=None, fanout=None, coef=None, shard='slices'):
    """This rank's share of the LINEAR stack of light field `vp` (M,M,S,T,3).

    shard='slices': a contiguous block of slices, float32 (n_local,S,T,3);
        out    : optional destination for the block (e.g. `stack[lo:hi]` of a full-size stack buffer)
        fanout : refinement path only - [(address of slice `lo` in a destination stack, None), ...]: the kernel stores
                 every finished slice into all of them (see PeerStack.fanout).
    shard='rows' (refinement path): the band `row_band(S, rank, world)` of EVERY slice - the prefilter and the horizontal
        pass then shrink with the band (plus the halo the vertical windows reach) instead of being repeated by every rank,
        and the horizontal pass keeps all its slice quads busy;
        out    : the FULL (N,S,T,3) stack the band is written into (returned); fanout addresses slice 0 of each stack."""
    from . import ops
    a_list = list(a_list)
    lo, hi = shard_bounds(len(a_list), rank, world)
    S, T, Cn = (int(v) for v in vp.shape[2:])
    if shard == 'rows':
        if not refi:
            raise NotImplementedError("row sharding exists for the refinement path (BASELINE configs[3])")
        from .lfp_refocuser.refinement import device_plan, refocus_refi
        mask = ops.aperture_mask(vp.shape[0]) if view_zeroed is None else view_zeroed
        plan = device_plan(S, T, int(vp.shape[0]), a_list, mask, vp.device)
        y_lo, y_hi = row_band(S, rank, world, halo=plan.halo_rows())
        if out is None and fanout is None:
            out = torch.empty((len(a_list), S, T, Cn), dtype=torch.float32, device=vp.device)
        if y_hi > y_lo:
            refocus_refi(vp, a_list, mask, plan=plan, out=out, fanout=fanout, coef=coef, row_range=(y_lo, y_hi))
        return out
    if shard != 'slices':
        raise ValueError("shard must be 'slices' or 'rows'")
    if hi == lo:
        return torch.empty((0, S, T, Cn), dtype=torch.float32, device=vp.device)
    if refi:
        from .lfp_refocuser.refinement import device_plan, refocus_refi
        mask = ops.aperture_mask(vp.shape[0]) if view_zeroed is None else view_zeroed
        # the operator tables of the WHOLE range: a block is then bit-identical to the same slices of the unsharded stack
        plan = device_plan(S, T, int(vp.shape[0]), a_list, mask, vp.device)
        return refocus_refi(vp, a_list, mask, plan=plan, out=out, fanout=fanout, coef=coef, n_range=(lo, hi))
    if fanout is not None:
        raise NotImplementedError("the fan-out epilogue exists for the refinement path (BASELINE configs[3]); the "
                                  "integer stack uses transport='nccl'")
    return ops.refocus_int(vp, a_list[lo:hi], view_zeroed=view_zeroed, out=out)


def finish_stack(stack, postprocess=True, out=None):
    """LfpRefocuser's colour management on the COMPLETE linear stack, done by every rank for itself: bounds from slice 0
    (lfp_refocuser/top_level.py:69), min / max of the whole stack for Normalizer's fallback, contrast + sRGB."""
    from . import ops
    if not postprocess:
        return stack
    return ops.refocuser_postprocess(stack, out=out)


# ----------------------------------------------------------------------------------------------------------
# peer-mapped stack buffers (transport='p2p')
# ----------------------------------------------------------------------------------------------------------
class _CudaArray(object):
    """Minimal __cuda_array_interface__ carrier so that torch can alias memory this library allocated."""

    def __init__(self, ptr, shape, typestr, owner):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2}
        self._owner = owner


class PeerStack(object):
    """One (N, S, T, C) float32 stack per rank, each mapped into every other rank of the group (CUDA IPC over
    NVLink / NVSwitch peer access).  `fanout(lo)` lists, for the kernel's epilogue, the address of slice `lo` in every
    rank's stack; `local` is this rank's stack as a torch tensor.  Collective: every rank of the group must construct
    (and `close`) it together."""

    def __init__(self, shape, group=None):
        from . import _native as nat
        self._nat = nat
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.shape = tuple(int(v) for v in shape)
        if self.world > nat.PCB_MAX_PEERS:
            raise ValueError("at most %d ranks per peer group" % nat.PCB_MAX_PEERS)
        self.nbytes = int(np.prod(self.shape)) * 4
        L = nat.lib()
        ptr = C.c_void_p()
        nat.check(L.pcb_peer_alloc(self.nbytes, C.byref(ptr)))
        self._own = ptr.value
        handle = C.create_string_buffer(64)
        nat.check(L.pcb_peer_export(C.c_void_p(self._own), handle))
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(handle.raw), group=group)
        self.ptrs = []
        for r, h in enumerate(handles):
            if r == self.rank:
                self.ptrs.append(self._own)
            else:
                p = C.c_void_p()
                nat.check(L.pcb_peer_open(C.create_string_buffer(h, 64), C.byref(p)))
                self.ptrs.append(p.value)
        dev = torch.device("cuda", torch.cuda.current_device())
        self.local = torch.as_tensor(_CudaArray(self._own, self.shape, "<f4", self), device=dev)
        self._flag = torch.zeros(1, dtype=torch.int32, device=dev)
        self._closed = False

    def fanout(self, first_slice):
        off = int(first_slice) * int(np.prod(self.shape[1:])) * 4
        return [(p + off, None) for p in self.ptrs]

    def barrier(self):
        """Stream-ordered: returns at once on the host; the work queued after it on the current stream starts when every
        rank's preceding kernels (and with them their peer stores) have completed."""
        dist.all_reduce(self._flag, group=self.group)

    def close(self):
        if self._closed:
            return
        self._closed = True
        torch.cuda.synchronize()
        dist.barrier(group=self.group)          # nobody unmaps while a peer may still be writing
        L = self._nat.lib()
        for r, p in enumerate(self.ptrs):
            if r != self.rank:
                L.pcb_peer_close(C.c_void_p(p))
        dist.barrier(group=self.group)
        self.local = None
        L.pcb_peer_free(C.c_void_p(self._own))


# ----------------------------------------------------------------------------------------------------------
# the sharded stack
# ----------------------------------------------------------------------------------------------------------
def gather_rows(full, group=None, halo=0):
    """All-gather of the row bands of a ROW-sharded stack (baseline transport): every rank's band of `full` (N,S,T,C) is
    sent to everybody and written into place."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    N, S = int(full.shape[0]), int(full.shape[1])
    bands = [row_band(S, r, world, halo) for r in range(world)]
    hmax = max(b - a for a, b in bands)
    y_lo, y_hi = bands[rank]
    send = torch.zeros((N, hmax) + tuple(full.shape[2:]), dtype=full.dtype, device=full.device)
    send[:, :y_hi - y_lo] = full[:, y_lo:y_hi]
    recv = torch.empty((world,) + tuple(send.shape), dtype=full.dtype, device=full.device)
    dist.all_gather_into_tensor(recv, send, group=group)
    for r, (a, b) in enumerate(bands):
        if r != rank and b > a:
            full[:, a:b] = recv[r, :, :b - a]
    return full


def sharded_refocus_stack(vp, a_list, group=None, postprocess=True, refi=False, view_zeroed=None, transport='nccl',
                          peer_stack=None, coef=None, shard='slices'):
    """Slice-sharded LfpRefocuser.shift_and_sum on device tensors: every rank passes the same light field `vp`
    (M,M,S,T,3) and gets the full finished stack back.
    refi=True: `a_list` are the up-sampled-pixel parameters of the refinement path (range(M*r0, M*r1)); each rank
    prefilters the light field itself (shift independent) and evaluates its own slices.
    transport='p2p' needs a `PeerStack` of the stack's shape (reusable across light fields) and refi=True.
    shard='rows' (refi only): every rank computes a band of image rows of ALL slices instead of a block of slices."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    a_list = list(a_list)
    n_total = len(a_list)
    lo, hi = shard_bounds(n_total, rank, world)
    if transport == 'p2p':
        if peer_stack is None:
            raise ValueError("transport='p2p' needs a PeerStack")
        peer_stack.barrier()                 # every rank has finished reading the previous contents of its stack
        local_block(vp, a_list, rank, world, refi=refi, view_zeroed=view_zeroed, coef=coef, shard=shard,
                    fanout=peer_stack.fanout(lo if shard == 'slices' else 0))
        peer_stack.barrier()                 # ... and all stores of this stack have landed
        full = peer_stack.local
    elif transport == 'nccl':
        local = local_block(vp, a_list, rank, world, refi=refi, view_zeroed=view_zeroed, coef=coef, shard=shard)
        if shard == 'slices':
            full = gather_slices(local, n_total, group)
        else:
            from . import ops
            from .lfp_refocuser.refinement import device_plan
            mask = ops.aperture_mask(vp.shape[0]) if view_zeroed is None else view_zeroed
            plan = device_plan(int(vp.shape[2]), int(vp.shape[3]), int(vp.shape[0]), a_list, mask, vp.device)
            full = gather_rows(local, group, halo=plan.halo_rows())
    else:
        raise ValueError("transport must be 'p2p' or 'nccl'")
    return finish_stack(full, postprocess)
