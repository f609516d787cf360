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


def _serial_reference(lw, lh, tile_px, tiles, f=8, k=8):
    """Independent restatement of the reference's serial loop (vae.c:331-387), written from that code alone: tile extent
    min(tile/f + 2k, full), step n - 2k, offsets min(t*step, full - n), tiles visited t1 outer / t0 inner, of every tile
    the region of (n-k)*f pixels starting k*f in (0 at the origin) is copied to (i+d)*f -- later tiles overwrite earlier
    ones. `tiles[i]` is the decoded tile i [3, n1*f, n0*f]. Does not use mlimgsynth_b200.dist."""
    tile_px = (tile_px + 63) // 64 * 64
    n0, n1 = min(tile_px // f + 2 * k, lw), min(tile_px // f + 2 * k, lh)
    step0, step1 = n0 - 2 * k, n1 - 2 * k
    nt0, nt1 = (lw + step0 - 1) // step0, (lh + step1 - 1) // step1
    canvas = np.zeros((3, lh * f, lw * f), np.float32)
    i_tile = 0
    for t1 in range(nt1):
        i1 = min(t1 * step1, lh - n1)
        for t0 in range(nt0):
            i0 = min(t0 * step0, lw - n0)
            d0, d1 = (k if i0 else 0), (k if i1 else 0)
            cw, ch = (n0 - k) * f, (n1 - k) * f
            canvas[:, (i1 + d1) * f:(i1 + d1) * f + ch, (i0 + d0) * f:(i0 + d0) * f + cw] = tiles[i_tile][:, d1 * f:d1 * f + ch, d0 * f:d0 * f + cw]
            i_tile += 1
    return canvas, i_tile, (n0 * f, n1 * f)


LW, LH, TILE_PX = 40, 56, 64          # 24x24-latent tiles, step 8: 5 x 7 = 35 overlapping tiles


def _fake_tile(idx, tw, th):
    """Stands for a decoded tile: depends on the tile index AND the position inside the tile."""
    y, x = np.mgrid[0:th, 0:tw].astype(np.float32)
    return np.stack([idx + 1 + x * 1e-3, idx + 1 + y * 1e-3, idx + 1 + (x + y) * 1e-3]).astype(np.float32)


def _host_lib_dryrun():
    """The product host library in dry-run mode: no GPU, "device" pointers are host memory, only data movement runs."""
    import ctypes as C
    import mlimgsynth_b200
    os.environ["GGML_B200_DRYRUN"] = "1"; os.environ["GGML_B200_QUIET"] = "1"
    L = C.CDLL(mlimgsynth_b200.HOST_LIB)

    class Plan(C.Structure):
        _fields_ = [(n, C.c_int) for n in ("n0", "n1", "nt0", "nt1", "step0", "step1", "k", "f")] + [("tile_elems", C.c_size_t)]
    L.sdvae_tile_plan.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(Plan)]
    L.sdvae_merge_tiles.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    P = C.addressof(C.c_char.in_dll(L, "g_vae_sd1"))
    return L, P, Plan


def _merge_with_product(L, P, gathered, world, slots):
    out = np.zeros((3, LH * 8, LW * 8), np.float32)
    g = np.ascontiguousarray(gathered, dtype=np.float32)
    nt = L.sdvae_merge_tiles(P, LW, LH, TILE_PX, g.ctypes.data, world, slots, out.ctypes.data)
    return out, nt


def _dryrun_merge_worker(q):
    import ctypes as C
    L, P, Plan = _host_lib_dryrun()
    plan = Plan()
    nt = L.sdvae_tile_plan(P, LW, LH, TILE_PX, C.byref(plan))
    tw, th = plan.n0 * 8, plan.n1 * 8
    tiles = [_fake_tile(i, tw, th) for i in range(nt)]
    want, n_ref, tile_px = _serial_reference(LW, LH, TILE_PX, tiles)
    want = (want + np.float32(1)) * np.float32(0.5)             # the merge applies the decoder's (x+1)/2 on the way (vae.h:43-47)
    ok = (nt == n_ref == 35) and tile_px == (tw, th) and plan.tile_elems == 3 * tw * th
    for world in (1, 2, 3, 8):
        slots = (nt + world - 1) // world
        g = np.zeros((world, slots, 3, th, tw), np.float32)
        for t in range(nt):
            g[t % world, t // world] = tiles[t]          # rank r decodes tiles r, r + world, ... into consecutive slots
        got, n = _merge_with_product(L, P, g, world, slots)
        ok = ok and n == nt and np.array_equal(got, want) and (got != 0).all()
    q.put(bool(ok))


def test_product_tile_merge_equals_serial_reference_loop():
    """sdvae_merge_tiles (the product's C merge, run here in dry-run mode on host buffers) == the reference's serial loop,
    for every world size: the slot mapping and the last-writer-wins order do not depend on how tiles were distributed."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")       # fresh process: GGML_B200_DRYRUN is read once per process
    q = ctx.Queue()
    p = ctx.Process(target=_dryrun_merge_worker, args=(q,)); p.start()
    assert q.get(timeout=120) is True
    p.join(timeout=60)


def _worker(rank, world, port, q):
    import ctypes as C
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # the multi-GPU tiled decode of mlimgsynth_b200.dist.vae_decode_tiled with the GPU pieces replaced: "decode" is a
        # deterministic function of the tile index, the gather runs over gloo, the merge is the product's C code (dry run)
        L, P, Plan = _host_lib_dryrun()
        plan = Plan()
        nt = L.sdvae_tile_plan(P, LW, LH, TILE_PX, C.byref(plan))
        tw, th = plan.n0 * 8, plan.n1 * 8
        slots = (nt + world - 1) // world
        mine = np.zeros((slots, 3, th, tw), np.float32)
        for slot, t in enumerate(range(rank, nt, world)):
            mine[slot] = _fake_tile(t, tw, th)
        g = D.gather_arrays(mine)
        tmax = D.all_max(float(rank + 1))
        ok = True
        if rank == 0:
            got, n = _merge_with_product(L, P, np.stack(g), world, slots)
            want, _, _ = _serial_reference(LW, LH, TILE_PX, [_fake_tile(t, tw, th) for t in range(nt)])
            want = (want + np.float32(1)) * np.float32(0.5)
            ok = n == nt and np.array_equal(got, want) and (got != 0).all()
            # the Python mirrors of the geometry (used for planning / docs) agree with the C plan
            ok = ok and len(D.tile_list(LW, LH, plan.n0, plan.n1, 8)) == nt
        imgs = D.gather_arrays(np.full((2, 4, 4, 3), rank, np.uint8))
        if rank == 0:
            ok = ok and [int(x[0, 0, 0, 0]) for x in imgs] == list(range(world))
        q.put((rank, bool(ok and tmax == float(world))))
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
