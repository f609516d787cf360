#!/usr/bin/env python3
"""Worker of tests/test_cfg_split_gpu.py: one of the TWO ranks of a cross-GPU CFG split (mlimgsynth_b200.dist.cfg_split_enable).
Each rank evaluates one CFG half per UNet evaluation and exchanges it with the peer. Backend nccl = one GPU per rank;
gloo = both ranks on GPU 0 (host-staged exchange), which is how the single-GPU test box runs it."""
import argparse, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools")); sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", required=True)
    ap.add_argument("--backend", default="gloo")
    a = ap.parse_args()
    import torch, torch.distributed as dist
    import bench
    import golden_cases as G
    from mlimgsynth_b200 import api, dist as D
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    dev = local if a.backend == "nccl" else 0
    torch.cuda.set_device(dev)
    dist.init_process_group(a.backend)
    os.environ.setdefault("GGML_B200_QUIET", "1")
    if rank == 0:
        bench.weights_path("sd1")
    dist.barrier()
    ctx = api.Ctx(backend="B200:%d" % dev, model=bench.weights_path("sd1"))
    for k, v in G.CASES["euler"]["opts"].items():
        ctx.set(k, v)
    ctx.set("image_dim", (128, 128)); ctx.set("batch_size", 1); ctx.set("seed", "42,0"); ctx.set("prompt", G.PROMPT)
    half = D.cfg_split_enable(ctx)
    ctx.generate()
    lat, img = ctx.tensor(api.TENSOR_LATENT), ctx.image(0)
    ctx.cfg_split(None)
    gathered = [None, None]
    dist.all_gather_object(gathered, (half, lat, img))
    if rank == 0:
        np.savez(a.out, lat0=gathered[0][1], lat1=gathered[1][1], img0=gathered[0][2], img1=gathered[1][2])
    dist.barrier()
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
