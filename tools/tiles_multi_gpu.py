#!/usr/bin/env python3
"""torchrun worker of tests/test_vae_tiles_gpu.py: tiled VAE decode with the tiles spread across the ranks (one process
per GPU, NCCL gather) vs the serial tiled decode on rank 0. Writes both RGB8 images to --out (rank 0)."""
import argparse, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", required=True)
    ap.add_argument("--kind", default="sd1")
    ap.add_argument("--latent", type=int, default=96)
    ap.add_argument("--tile", type=int, default=256)
    a = ap.parse_args()
    import torch, torch.distributed as dist
    import bench
    from mlimgsynth_b200 import api, dist as D
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if rank == 0:
        bench.weights_path(a.kind)
    dist.barrier()
    os.environ.setdefault("GGML_B200_QUIET", "1")
    ctx = api.Ctx(backend="B200:%d" % local, model=bench.weights_path(a.kind), vae_tile=a.tile)
    lat = (np.random.default_rng(5).standard_normal((1, 4, a.latent, a.latent)) * 0.18).astype(np.float32)
    multi = D.vae_decode_tiled(ctx, lat)
    if rank == 0:
        ctx.decode(lat)
        serial = ctx.image(0)
        np.savez(a.out, multi=multi, serial=serial)
        print("tiles: %d on %d GPUs, identical: %s" % (ctx.vae_tile_plan(a.latent, a.latent)[0], world, np.array_equal(multi, serial)))
    dist.barrier()
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
