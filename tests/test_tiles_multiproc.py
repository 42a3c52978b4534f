"""N > 1 host-side logic on CPU: the interleaved-tile partition (vkrt_b200/csrc/tiles.h, exported as VKRT_tilePartition) and the
gather layout used by vkrt_cuda_gather (rank-major, each rank padded to the largest local pixel count), exercised with two real
processes over torch.distributed's gloo backend."""
import ctypes as C
import os
import socket
import sys

import numpy as np
import pytest

import conftest


def _partition(lib, w, h, tw, th, rank, world):
    n, tx, ty = C.c_uint32(), C.c_uint32(), C.c_uint32()
    assert lib.VKRT_tilePartition(w, h, tw, th, rank, world, C.byref(n), None, 0, C.byref(tx), C.byref(ty)) == 0
    l2g = np.zeros(max(n.value, 1), np.uint32)
    assert lib.VKRT_tilePartition(w, h, tw, th, rank, world, C.byref(n), l2g.ctypes.data_as(C.c_void_p), len(l2g), None, None) == 0
    return l2g[:n.value], tx.value, ty.value


def _local_pixel_ids(l2g, tx, w, h, tw, th):
    """Global pixel index of every local (tile-compact) pixel, -1 for the padding of edge tiles — what k_untile inverts."""
    out = np.full(len(l2g) * tw * th, -1, np.int64)
    for lt, gt in enumerate(l2g):
        x0, y0 = (gt % tx) * tw, (gt // tx) * th
        ys, xs = np.meshgrid(np.arange(th), np.arange(tw), indexing="ij")
        gx, gy = x0 + xs, y0 + ys
        ok = (gx < w) & (gy < h)
        ids = np.where(ok, gy * w + gx, -1).reshape(-1)
        out[lt * tw * th:(lt + 1) * tw * th] = ids
    return out


@pytest.mark.parametrize("w,h,world", [(200, 120, 2), (1920, 1080, 8), (33, 31, 3), (64, 64, 4), (5, 3, 2)])
def test_partition_covers_every_pixel_once(w, h, world):
    from vkrt_b200 import host
    lib = host.load_host_library()
    seen = np.zeros(w * h, np.int32)
    counts = []
    for r in range(world):
        l2g, tx, ty = _partition(lib, w, h, 32, 32, r, world)
        assert np.all(np.diff(l2g.astype(np.int64)) > 0)  # ascending
        ids = _local_pixel_ids(l2g, tx, w, h, 32, 32)
        np.add.at(seen, ids[ids >= 0], 1)
        counts.append(len(l2g))
    assert np.all(seen == 1)
    assert max(counts) - min(counts) <= max(1, (w + 31) // 32)  # interleaving balances the tile counts
    if world > 1 and (w + 31) // 32 >= world:
        # the row stagger keeps tile columns from aliasing onto one rank
        l2g, tx, ty = _partition(lib, w, h, 32, 32, 0, world)
        cols = set((l2g % tx).tolist())
        assert len(cols) > 1


def _worker(rank, world, port, w, h, out_path):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, conftest.ROOT)
    from vkrt_b200 import host
    lib = host.load_host_library()
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    tw = th = 32
    parts = [_partition(lib, w, h, tw, th, r, world) for r in range(world)]
    stride = max(len(p[0]) for p in parts) * tw * th  # vkrt_cuda_max_local_pixels
    l2g, tx, ty = parts[rank]
    ids = _local_pixel_ids(l2g, tx, w, h, tw, th)
    # this rank's tile-compact "film": value = f(global pixel), padding = -7
    film = np.where(ids >= 0, ids.astype(np.float64) * 3.0 + 1.0, -7.0)
    padded = np.full(stride, -7.0)
    padded[:len(film)] = film
    t = torch.from_numpy(padded)
    gathered = [torch.zeros(stride, dtype=torch.float64) for _ in range(world)] if rank == 0 else None
    dist.gather(t, gathered, dst=0)
    if rank == 0:
        full = np.zeros(w * h)
        for r in range(world):  # the un-tile step of vkrt_cuda_gather / vkrt_cuda_import_gathered
            rl2g, rtx, _ = parts[r]
            rids = _local_pixel_ids(rl2g, rtx, w, h, tw, th)
            buf = gathered[r].numpy()[:len(rids)]
            full[rids[rids >= 0]] = buf[rids >= 0]
        np.save(out_path, full)
    dist.barrier()
    dist.destroy_process_group()


def test_two_process_gather_reassembles_the_frame(tmp_path):
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    w, h, world = 200, 120, 2
    out = str(tmp_path / "full.npy")
    mp.spawn(_worker, args=(world, port, w, h, out), nprocs=world, join=True)
    full = np.load(out)
    assert np.array_equal(full, np.arange(w * h, dtype=np.float64) * 3.0 + 1.0)
