"""World-size-2 (and 3) gloo tests of the multi-GPU host logic in plenopticam_b200/parallel.py, on CPU tensors.
The per-rank compute is stood in for by the numpy oracle so that the sharding, the contrast-bound broadcast and
the ragged all-gather are checked end to end against the unsharded result."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import oracle_np as onp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_slices, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from plenopticam_b200 import parallel as par
        rng = np.random.default_rng(7)
        M, S, T = 5, 6, 7
        vp = rng.integers(0, 65536, size=(M, M, S, T, 3), dtype=np.uint16)
        a_list = list(range(-(n_slices // 2), n_slices - n_slices // 2))
        idx, mine = par.shard_slices(a_list, rank, world)
        assert idx == list(range(*par.shard_bounds(len(a_list), rank, world)))
        vp64 = onp._apply_aperture(vp, M) / M
        local = np.asarray([onp.refocus_int_slice(vp64, a, M) for a in mine]).reshape((len(mine), S, T, 3))
        # contrast bounds from slice 0, owned by exactly one rank
        src = par.owner_of(0, len(a_list), world)
        lohi = torch.zeros(2, dtype=torch.float64)
        if rank == src:
            lo, hi = onp.hist_percentiles(local[0])
            lohi[0], lohi[1] = float(lo), float(hi)
        par.broadcast_bounds(lohi, src)
        lo, hi = float(lohi[0]), float(lohi[1])
        post = onp.srgb_conv(np.asarray(onp.norm_fun(local.astype('float32'), np.float64(lo), np.float64(hi)),
                                        dtype='float32')) if len(mine) else local.astype('float32')
        full = par.gather_slices(torch.from_numpy(np.ascontiguousarray(post, dtype=np.float32)), len(a_list))
        ref = onp.refocuser_postprocess(np.asarray([onp.refocus_int_slice(vp64, a, M) for a in a_list]))
        ok = full.shape == ref.shape and np.array_equal(full.numpy(), ref)
        # round-2 flow of parallel.sharded_refocus_stack: exchange the LINEAR blocks, every rank post-processes the whole
        # stack itself (bounds from slice 0, no broadcast / all-reduce)
        lin = par.gather_slices(torch.from_numpy(np.ascontiguousarray(local, dtype=np.float64)), len(a_list))
        ok = ok and np.array_equal(onp.refocuser_postprocess(lin.numpy()), ref)
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_slices", [(2, 6), (2, 7), (3, 4)])
def test_slice_sharding_matches_unsharded(world, n_slices):
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, n_slices, ret), nprocs=world, join=True)
    assert all(ret.get(r) for r in range(world)), dict(ret)


def test_shard_bounds_cover_everything():
    from plenopticam_b200 import parallel as par
    for n in (0, 1, 5, 64, 65):
        for world in (1, 2, 3, 8):
            spans = [par.shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
            assert sum(len(par.shard_exposures(n, r, world)) for r in range(world)) == n
            for i in range(n):
                lo, hi = spans[par.owner_of(i, n, world)]
                assert lo <= i < hi
