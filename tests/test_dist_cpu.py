"""Multi-rank host logic on CPU (gloo, world size 2): image sharding, seed assignment, ordered tile merge, gather
and max-over-ranks timing. The GPU path uses the same functions with NCCL (bench.py --gpus N)."""
import os, sys, socket
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mlimgsynth_b200 import dist as D


def test_image_slices_cover_batch():
    for n in (1, 7, 8, 16):
        for world in (1, 2, 3, 8):
            got = [i for r in range(world) for i in D.image_slice(n, r, world)]
            assert got == list(range(n))
    # seeds do not depend on the world size: image i always gets seed + i (generate.sh:55-61 semantics)
    assert [s for r in range(4) for s in D.image_seeds(42, 8, r, 4)] == [42 + i for i in range(8)]


def test_tile_grid_matches_reference_geometry():
    # SDXL 2048^2, --vae-tile 512: latent 256, tile 512/8 + 16 = 80, step 64, offsets min(t*64, 176) (vae.c:335,366-368)
    assert D.tile_grid(256, 80, 8) == [0, 64, 128, 176]
    assert D.tile_grid(64, 64, 8) == [0]
    assert len(D.tile_list(256, 256, 80, 80, 8)) == 16


def _serial_reference(w, h, tw, th, k, up, tiles_px):
    """The reference's serial loop: tiles visited row-major, later ones overwrite earlier ones."""
    canvas = np.zeros((3, h * up, w * up), np.float32)
    return D.merge_tiles(canvas, tiles_px, w, h, tw, th, k, up)


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        w = h = 24; tw = th = 16; k = 2; up = 2
        tiles = D.tile_list(w, h, tw, th, k)
        mine = D.tiles_of_rank(tiles, rank, world)
        # "decode": a deterministic function of the tile index, so every rank can rebuild the expected result
        def fake(idx):
            return np.full((3, th * up, tw * up), float(idx + 1), np.float32) + np.arange(tw * up, dtype=np.float32)[None, None, :] * 1e-3
        n_max = (len(tiles) + world - 1) // world
        buf = np.zeros((n_max, 3, th * up, tw * up), np.float32)
        ids = np.full((n_max,), -1, np.int64)
        for j, (idx, _, _) in enumerate(mine):
            buf[j] = fake(idx); ids[j] = idx
        g_buf, g_ids = D.gather_arrays(buf), D.gather_arrays(ids)
        tmax = D.all_max(float(rank + 1))
        ok = True
        if rank == 0:
            decoded = {int(i): g_buf[r][j] for r in range(world) for j, i in enumerate(g_ids[r]) if i >= 0}
            got = D.merge_tiles(np.zeros((3, h * up, w * up), np.float32), decoded, w, h, tw, th, k, up)
            want = _serial_reference(w, h, tw, th, k, up, {idx: fake(idx) for idx, _, _ in tiles})
            ok = np.array_equal(got, want) and (got != 0).all()
        imgs = D.gather_arrays(np.full((2, 4, 4, 3), rank, np.uint8))
        if rank == 0:
            ok = ok and [int(x[0, 0, 0, 0]) for x in imgs] == list(range(world))
        q.put((rank, ok and tmax == float(world)))
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo():
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps: p.start()
    res = [q.get(timeout=120) for _ in ps]
    for p in ps: p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]
