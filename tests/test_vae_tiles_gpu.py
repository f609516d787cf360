"""VAE decode tiles spread across GPUs (SURVEY 8e / BASELINE configs[4]): the product's split + merge path against the
serial tiled decode of the same engine (bit-identical RGB8 image) -- every rank's share decoded in this process for any
world size, and the real thing (one process per GPU, NCCL gather of device buffers) when the box has >= 2 GPUs."""
import os, subprocess, sys
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def same_image(a, b, what):
    """Same tiles, same order, same kernels: the two images agree to the engine's run-to-run reproducibility. With
    GGML_B200_DETERMINISTIC GroupNorm statistics (fixed-order reductions, the default) that is bit for bit."""
    d = np.abs(a.astype(np.int16) - b.astype(np.int16))
    print("%s: %d of %d pixels differ, max |diff| %d" % (what, int((d > 0).any(-1).sum()), d.shape[0] * d.shape[1], int(d.max())))
    assert np.array_equal(a, b), what


@pytest.fixture(scope="module")
def ctx():
    import bench
    from mlimgsynth_b200 import api
    c = api.Ctx(model=bench.weights_path("sd1"))
    yield c
    c.close()


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_tile_split_merge_equals_serial_tiled_decode(ctx, world):
    import torch
    lat = (np.random.default_rng(5).standard_normal((1, 4, 72, 56)) * 0.18).astype(np.float32)     # lw 56, lh 72
    ctx.set("vae_tile", 192)          # tiles of 24 + 16 = 40 latent pixels, step 24: 3 x 3 = 9 tiles... (56: 0,16 clipped)
    img = ctx.decode(lat)             # serial tiled decode (run_tiled), float image [1,3,H,W]
    serial_u8 = ctx.image(0).copy()
    n_tiles, tw, th = ctx.vae_tile_plan(56, 72)
    assert n_tiles > 1 and (tw, th) == (320, 320)
    slots = (n_tiles + world - 1) // world
    g = torch.zeros((world, slots, 3 * th * tw), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    for r in range(world):            # what rank r of `world` would do on its own GPU
        ctx.vae_tiles_decode(lat, r, world, g[r].data_ptr())
    fimg = ctx.vae_tiles_merge(56, 72, g.data_ptr(), world, slots, want_float=True)
    split_u8 = ctx.image(0)
    ctx.set("vae_tile", 0)
    assert split_u8.shape == serial_u8.shape == (576, 448, 3)
    same_image(split_u8, serial_u8, "world %d" % world)
    assert np.abs(fimg - img).max() <= 2e-3


def test_tiles_across_gpus_nccl(tmp_path):
    """One process per GPU (torchrun), tiles round-robin, NCCL gather of the device buffers, merge on rank 0 == serial."""
    import torch
    n = min(torch.cuda.device_count(), 8)
    if n < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    out = str(tmp_path / "tiles.npz")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr", "127.0.0.1",
           "--master-port", "29631", os.path.join(ROOT, "tools", "tiles_multi_gpu.py"), "--out", out]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=1200)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    z = np.load(out)
    same_image(z["multi"], z["serial"], "%d GPUs" % n)
