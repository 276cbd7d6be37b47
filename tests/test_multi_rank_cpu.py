"""N > 1 host logic on CPU: two gloo ranks each render their share of one image with the CPU oracle (the C-ABI partition
fields are backend independent) and the sum over ranks must equal the single-rank render."""
import os
import socket
import sys

import numpy as np
import pytest
from oracle import loader as oracle_loader

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, mode, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from vviewer_b200 import capi
    import partition_util as parallel
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    eng = capi.HostEngine()
    eng.build_scene("Cornell")
    eng.set_render_info(width=48, height=40, samples=8, batch_size=2, depth=5)
    ctx = capi.Context(oracle_loader.load_oracle())
    ctx.upload_scene(eng.scene_desc())
    ctx.build_accel()
    rp = parallel.partition(eng.render_params(), rank, world, mode, tile_size=16)
    rad, alb, nrm = ctx.render(rp)
    t = torch.from_numpy(np.stack([rad, alb, nrm]).copy())
    seg = torch.tensor([ctx.stats()["segments"]], dtype=torch.int64)
    dist.reduce(t, dst=0)  # the sum of the parts IS the image, alpha included (rank 0 alone writes the 1)
    dist.reduce(seg, dst=0)
    if rank == 0:
        np.save(os.path.join(out_dir, "sum_%s.npy" % mode), t.numpy())
        np.save(os.path.join(out_dir, "seg_%s.npy" % mode), seg.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["tile", "sample"])
def test_two_ranks_sum_to_full_image(capi, tmp_path, mode):
    import torch.multiprocessing as mp
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, mode, str(tmp_path)), nprocs=world, join=True)
    total = np.load(str(tmp_path / ("sum_%s.npy" % mode)))
    seg = int(np.load(str(tmp_path / ("seg_%s.npy" % mode)))[0])
    # single-rank reference
    eng = capi.HostEngine()
    eng.build_scene("Cornell")
    eng.set_render_info(width=48, height=40, samples=8, batch_size=2, depth=5)
    ctx = capi.Context(oracle_loader.load_oracle())
    ctx.upload_scene(eng.scene_desc())
    ctx.build_accel()
    rad, alb, nrm = ctx.render(eng.render_params())
    full = np.stack([rad, alb, nrm])
    assert np.all(total[..., 3] == 1.0) and np.all(full[..., 3] == 1.0)  # alpha survives the sum
    assert seg == ctx.stats()["segments"]
    # same samples, only the floating-point summation order differs
    assert np.allclose(total, full, rtol=1e-5, atol=1e-6)
    ctx.close()
    eng.close()


def test_batches_of_rank():
    import partition_util as parallel
    assert [parallel.batches_of_rank(1024, 16, r, 8, "sample") for r in range(8)] == [8] * 8
    assert [parallel.batches_of_rank(70, 16, r, 3, "sample") for r in range(3)] == [2, 1, 1]
    assert parallel.batches_of_rank(70, 16, 1, 3, "tile") == 4


@pytest.mark.parametrize("w,h,tile,world", [(50, 37, 16, 3), (37, 29, 32, 3), (33, 17, 0, 2)])
def test_ragged_tiles_and_ranks_without_pixels(capi, w, h, tile, world):
    """tile split on image sizes that are no multiple of the tile edge, with more ranks than tiles (a rank that owns no pixel renders
    nothing and returns zeros) and with tile_size 0 (= 32): the parts still sum to the full render, alpha included (oracle; the
    CUDA core runs the same cases in test_gpu_parity.test_partition_sums_to_full_render)"""
    import partition_util as parallel
    from oracle import loader as oracle_loader
    eng = capi.HostEngine()
    eng.build_scene("MeshLight")
    eng.set_render_info(width=w, height=h, samples=4, batch_size=2)
    ctx = capi.Context(oracle_loader.load_oracle())
    ctx.upload_scene(eng.scene_desc())
    ctx.build_accel()
    full = np.stack(ctx.render(eng.render_params()))
    total = np.zeros_like(full)
    owned = []
    for r in range(world):
        part = np.stack(ctx.render(parallel.partition(eng.render_params(), r, world, "tile", tile_size=tile)))
        owned.append(int(np.count_nonzero(part[1].reshape(-1, 4)[:, :3].any(axis=1) | part[0].reshape(-1, 4)[:, :3].any(axis=1))))
        total += part
    assert np.allclose(total[..., :3], full[..., :3], rtol=1e-5, atol=1e-6)
    assert np.all(total[..., 3] == 1.0)
    edge = tile or 32
    n_tiles = -(-w // edge) * -(-h // edge)
    if n_tiles < world:
        assert owned.count(0) >= world - n_tiles
    ctx.close()
    eng.close()
