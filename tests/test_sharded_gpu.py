"""Slice-sharded refocus stack (BASELINE configs[3]) on ONE GPU: the per-rank halves of parallel.sharded_refocus_stack
(`local_block`, `finish_stack`) are run for every rank of several world sizes in turn and the reassembled stack must
equal the unsharded one bit for bit - integer and refinement path, ragged splits, the bounds coming from whoever owns
slice 0.  The same halves joined by the real exchange (NCCL all-gather / peer stores over NVLink) are checked on 2-8
GPUs by bench.py's `sharded_stack` record."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _lf(cuda, M, S, T, seed):
    from plenopticam_b200 import ops
    rng = np.random.default_rng(seed)
    return ops.to_device((rng.random((M, M, S, T, 3)) ** 2 * 65535).astype(np.uint16))


@pytest.mark.parametrize("world", [2, 3, 8])
@pytest.mark.parametrize("refi", [False, True])
def test_blocks_of_every_rank_reassemble_the_unsharded_stack(cuda, world, refi):
    import torch
    from plenopticam_b200 import ops, parallel as par
    from plenopticam_b200.lfp_refocuser.refinement import refocus_refi
    M, S, T = 7, 33, 41
    vp = _lf(cuda, M, S, T, 5)
    a_list = list(range(-M, M + 3)) if refi else list(range(-6, 7))        # 17 / 13 slices: ragged for 2, 3 and 8 ranks
    mask = ops.aperture_mask(M)
    ref = refocus_refi(vp, a_list, mask) if refi else ops.refocus_int(vp, a_list)
    full = torch.full((len(a_list), S, T, 3), float('nan'), dtype=torch.float32, device=vp.device)
    sizes = []
    for rank in range(world):
        lo, hi = par.shard_bounds(len(a_list), rank, world)
        sizes.append(hi - lo)
        if hi > lo:
            blk = par.local_block(vp, a_list, rank, world, refi=refi, out=full[lo:hi])
            assert blk.data_ptr() == full[lo:hi].data_ptr() and tuple(blk.shape) == (hi - lo, S, T, 3)
        else:
            assert par.local_block(vp, a_list, rank, world, refi=refi).shape[0] == 0
    assert sum(sizes) == len(a_list) and max(sizes) - min(sizes) <= 1
    assert torch.equal(full, ref)
    # every rank finishes the gathered stack by itself: same result as the single-GPU post-process
    assert torch.equal(par.finish_stack(full), ops.refocuser_postprocess(ref))


@pytest.mark.parametrize("world,S", [(2, 101), (3, 101), (8, 101), (8, 434), (5, 300)])
def test_row_bands_of_every_rank_reassemble_the_unsharded_refined_stack(cuda, world, S):
    """shard='rows': every rank computes a band of image rows of ALL slices from a prefilter restricted to the rows the band
    needs; the bands written into one stack must equal the unsharded stack bit for bit, min / max included."""
    import torch
    from plenopticam_b200 import ops, parallel as par
    from plenopticam_b200.lfp_refocuser.refinement import device_plan, refocus_refi
    M, T = 7, 47                                            # S = 101: 13 row groups, ragged for 2, 3 and 8 ranks; the tall
                                                            # images have bands far from both borders (interior strips)
    vp = _lf(cuda, M, S, T, 6)
    a_list = list(range(-2 * M, 2 * M + 1))
    mask = ops.aperture_mask(M)
    halo = device_plan(S, T, M, a_list, mask, vp.device).halo_rows()
    assert 0 < halo < 40
    mm_ref = ops.new_minmax()
    ref = refocus_refi(vp, a_list, mask, mm=mm_ref)
    full = torch.full((len(a_list), S, T, 3), float('nan'), dtype=torch.float32, device=vp.device)
    mm = ops.new_minmax()
    covered = 0
    for rank in range(world):
        y_lo, y_hi = par.row_band(S, rank, world, halo)
        assert y_lo % par.ROW_GROUP == 0 and (y_hi % par.ROW_GROUP == 0 or y_hi == S)
        covered += y_hi - y_lo
        out = par.local_block(vp, a_list, rank, world, refi=True, shard='rows',
                              fanout=[(full.data_ptr(), mm.data_ptr())])
        assert out is None
    assert covered == S
    assert torch.equal(full, ref)
    assert ops.minmax_values(mm) == ops.minmax_values(mm_ref)
    # a band through `out=` touches its own rows only
    canvas = torch.full_like(full, -7.0)
    y_lo, y_hi = par.row_band(S, 1, world, halo)
    par.local_block(vp, a_list, 1, world, refi=True, shard='rows', out=canvas)
    assert torch.equal(canvas[:, y_lo:y_hi], ref[:, y_lo:y_hi])
    assert bool((canvas[:, :y_lo] == -7.0).all()) and bool((canvas[:, y_hi:] == -7.0).all())
    with pytest.raises(NotImplementedError):
        par.local_block(vp, a_list, 0, world, refi=False, shard='rows')


def test_row_band_at_illum_angular_size(cuda):
    """M = 15 with shifts up to +-30: the vertical windows reach 14 rows beyond a band; the band plan must cover them."""
    import torch
    from plenopticam_b200 import ops, parallel as par
    from plenopticam_b200.lfp_refocuser.refinement import refocus_refi
    M, S, T = 15, 90, 33
    vp = _lf(cuda, M, S, T, 7)
    a_list = list(range(-2 * M, 2 * M))
    mask = ops.aperture_mask(M)
    ref = refocus_refi(vp, a_list, mask)
    full = torch.zeros_like(ref)
    for rank in range(4):
        par.local_block(vp, a_list, rank, 4, refi=True, shard='rows', out=full)
    assert torch.equal(full, ref)


def test_refinement_fanout_writes_every_destination(cuda):
    """pcb_refocus_refi_fanout: the kernel stores each finished slice (and folds the min / max) into all destinations -
    here three buffers on the same GPU stand in for the peer-mapped stacks of other GPUs."""
    import torch
    from plenopticam_b200 import ops, parallel as par
    from plenopticam_b200.lfp_refocuser.refinement import device_plan, refocus_refi
    M, S, T = 5, 24, 37
    vp = _lf(cuda, M, S, T, 9)
    a_list = list(range(-M, M))
    mask = ops.aperture_mask(M)
    mm_ref = ops.new_minmax()
    ref = refocus_refi(vp, a_list, mask, mm=mm_ref)
    plan = device_plan(S, T, M, a_list, mask, vp.device)
    stacks = [torch.zeros((len(a_list), S, T, 3), dtype=torch.float32, device=vp.device) for _ in range(3)]
    mms = [ops.new_minmax() for _ in range(3)]
    for rank in range(2):                                      # two "ranks", each storing into all three stacks
        lo, hi = par.shard_bounds(len(a_list), rank, 2)
        fan = [(s[lo].data_ptr(), m.data_ptr()) for s, m in zip(stacks, mms)]
        refocus_refi(vp, a_list, mask, plan=plan, fanout=fan, n_range=(lo, hi))
    for s, m in zip(stacks, mms):
        assert torch.equal(s, ref)
        assert ops.minmax_values(m) == ops.minmax_values(mm_ref)
    # argument checks of the C-ABI entry point
    from plenopticam_b200 import _native as nat
    with pytest.raises(ValueError):
        refocus_refi(vp, a_list, mask, plan=plan, fanout=[(stacks[0].data_ptr(), None)] * (nat.PCB_MAX_PEERS + 1))
    with pytest.raises(ValueError):
        refocus_refi(vp, a_list, mask, plan=plan, fanout=[(stacks[0].data_ptr(), mms[0].data_ptr()),
                                                          (stacks[1].data_ptr(), None)])


def test_peer_buffer_round_trip_in_one_process(cuda):
    """pcb_peer_alloc / export / free and the torch alias of the buffer (opening a handle needs a second process:
    bench.py's multi-GPU record does that)."""
    import torch
    from plenopticam_b200 import _native as nat
    L = nat.lib()
    ptr = C.c_void_p()
    nat.check(L.pcb_peer_alloc(4096, C.byref(ptr)))
    handle = C.create_string_buffer(64)
    nat.check(L.pcb_peer_export(ptr, handle))
    assert any(handle.raw)
    from plenopticam_b200.parallel import _CudaArray
    t = torch.as_tensor(_CudaArray(ptr.value, (1024,), "<f4", None), device=cuda)
    assert t.data_ptr() == ptr.value
    t.fill_(3.0)
    assert float(t.sum()) == 3072.0
    del t
    nat.check(L.pcb_peer_free(ptr))
    with pytest.raises(ValueError):
        nat.check(L.pcb_peer_alloc(0, C.byref(ptr)))
