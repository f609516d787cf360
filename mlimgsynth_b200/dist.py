"""Multi-GPU plumbing of the denoising path: one process per GPU, no data-path collective.

The path shards only across independent units:
  * images of a batch -- image i is generated with seed + i and a fresh noise offset (the semantics of the
    reference's generate.sh:55-61 loop; the reference itself rejects batch > 1, mlimgsynth.c:1640-1641);
  * VAE decode tiles -- independent given the fixed tile graph (vae.c:343-346); the merged image must be what the
    reference's serial row-major loop produces: later tiles overwrite earlier ones (vae.c:365-387).
torch.distributed (NCCL on the GPUs, gloo in the CPU tests) is used for the barrier, max-over-ranks timing and one
gather of the final images / tiles. Nothing here touches the kernels: ranks never exchange activations.
"""
import numpy as np


def image_slice(n_images, rank, world):
    """Contiguous block of image indices owned by `rank` (blocks differ by at most one image)."""
    base, rem = divmod(n_images, world)
    start = rank * base + min(rank, rem)
    return range(start, start + base + (1 if rank < rem else 0))


def image_seeds(seed, n_images, rank, world):
    """Seeds of the images of this rank: image i of the global batch always gets seed + i, whatever the world size."""
    return [seed + i for i in image_slice(n_images, rank, world)]


def tile_grid(full, tile, k):
    """Tile origins along one axis, as the reference computes them (vae.c:332-391): tile extent `tile` (already
    including the 2k overlap and clipped to `full`), step tile - 2k, last tile shifted back inside the tensor."""
    if tile >= full:
        return [0]
    step = tile - 2 * k
    n = (full + step - 1) // step
    return [min(t * step, full - tile) for t in range(n)]


def tile_list(w, h, tw, th, k):
    """All tiles in the reference's visiting order (row-major: y outer, x inner) as (index, x0, y0)."""
    out = []
    for y0 in tile_grid(h, th, k):
        for x0 in tile_grid(w, tw, k):
            out.append((len(out), x0, y0))
    return out


def tiles_of_rank(tiles, rank, world):
    """Round-robin assignment of tiles to ranks."""
    return [t for t in tiles if t[0] % world == rank]


def merge_tiles(canvas, decoded, w, h, tw, th, k, up):
    """Paste decoded tiles into `canvas` [C, h*up, w*up] in the reference's order so that overlaps resolve identically
    (the kept region of a tile is [d, d + n - k) with d = k except at the left/top border; vae.c:365-387).
    `decoded` maps tile index -> array [C, th*up, tw*up]."""
    for idx, x0, y0 in tile_list(w, h, tw, th, k):
        t = decoded[idx]
        d0, d1 = (k if x0 else 0), (k if y0 else 0)
        c0 = tw if tw == w else tw - k
        c1 = th if th == h else th - k
        ys, xs = (y0 + d1) * up, (x0 + d0) * up
        canvas[:, ys:ys + c1 * up, xs:xs + c0 * up] = t[:, d1 * up:(d1 + c1) * up, d0 * up:(d0 + c0) * up]
    return canvas


def vae_decode_tiled(ctx, latent, rank=None, world=None, dst=0):
    """Tiled VAE decode with the tiles spread round-robin across the ranks of the default process group (BASELINE
    configs[4]; reference loop vae.c:331-391). Each rank decodes its tiles on its own GPU into a device buffer, ONE NCCL
    gather moves the tiles device-to-device to rank `dst`, which pastes them in the reference's row-major order
    (mlis_b200_vae_tiles_merge). Returns the RGB8 image [H,W,3] on rank `dst`, None elsewhere. `latent`: [1,4,lh,lw].
    Without an initialised process group (or world 1) this is the serial tiled decode through the same code."""
    import torch
    import torch.distributed as dist
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
        rank = dist.get_rank() if dist.is_initialized() else 0
    lh, lw = latent.shape[-2:]
    n_tiles, tw, th = ctx.vae_tile_plan(lw, lh)
    slots = (n_tiles + world - 1) // world
    mine = torch.empty((slots, 3 * th * tw), dtype=torch.float32, device="cuda")
    torch.cuda.current_stream().synchronize()
    ctx.vae_tiles_decode(latent, rank, world, mine.data_ptr())         # synchronises the engine stream before returning
    if world > 1:
        gathered = torch.empty((world, slots, 3 * th * tw), dtype=torch.float32, device="cuda") if rank == dst else None
        dist.gather(mine, list(gathered.unbind(0)) if rank == dst else None, dst=dst)
        torch.cuda.current_stream().synchronize()
    else:
        gathered = mine
    if rank != dst:
        return None
    ctx.vae_tiles_merge(lw, lh, gathered.data_ptr(), world, slots)
    return ctx.image(0)


def cfg_split_enable(ctx, group=None):
    """Opt-in cross-GPU CFG split for fewer images than GPUs: the two ranks of `group` (default: the world,
    which must then have exactly two ranks) run the SAME generation -- same options, seed, prompts -- and each evaluates one
    CFG half per UNet evaluation; the halves are exchanged with one 2-rank NCCL all-gather of the UNet output
    (64-256 KB per image). Both ranks end with the same latent / image. Call ctx.cfg_split(None) to switch it off."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    if world != 2:
        raise ValueError("the CFG split pairs exactly two ranks")
    half = dist.get_rank(group)
    bufs = {}

    def exchange(mine_ptr, other_ptr, n):
        if n not in bufs:
            bufs[n] = torch.empty((2, n), dtype=torch.float32, device="cuda")
        mine = torch.as_tensor(_DeviceWords(mine_ptr, n), device="cuda")
        other = torch.as_tensor(_DeviceWords(other_ptr, n), device="cuda")
        if dist.get_backend(group) == "nccl":
            dist.all_gather_into_tensor(bufs[n], mine, group=group)
            other.copy_(bufs[n][1 - half])
        else:                           # host-staged (gloo): used by the single-GPU test, two processes on one device
            parts = [torch.empty(n, dtype=torch.float32) for _ in range(2)]
            dist.all_gather(parts, mine.cpu(), group=group)
            other.copy_(parts[1 - half])
        torch.cuda.current_stream().synchronize()
    ctx.cfg_split(half, exchange)
    return half


class _DeviceWords:
    """Zero-copy f32 view of foreign device memory for torch (CUDA array interface)."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 2}


class _DeviceBytes:
    """Zero-copy view of foreign device memory for torch (CUDA array interface)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


_pinned = {}


def gather_images_device(ctx, dst=0):
    """Gather the RGB8 images of the last generation of every rank on rank `dst`: NCCL gather of the device buffers the
    pack kernel wrote (no re-upload of host copies), then ONE device-to-host copy into pinned memory on `dst`.
    Returns [world * n, h, w, 3] uint8 (numpy) there, None elsewhere."""
    import torch
    import torch.distributed as dist
    ptr, n, h, w = ctx.images_device()
    nbytes = n * h * w * 3
    mine = torch.as_tensor(_DeviceBytes(ptr, nbytes), device="cuda")
    world, rank = dist.get_world_size(), dist.get_rank()
    if rank == dst:
        g = torch.empty((world, nbytes), dtype=torch.uint8, device="cuda")
        dist.gather(mine, list(g.unbind(0)), dst=dst)
        key = (world, nbytes)
        if key not in _pinned:
            _pinned.clear(); _pinned[key] = torch.empty((world, nbytes), dtype=torch.uint8, pin_memory=True)
        host = _pinned[key]
        host.copy_(g, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return host.numpy().reshape(world * n, h, w, 3)
    dist.gather(mine, None, dst=dst)
    return None


def gather_arrays(arr, dst=0, device=None):
    """Gather equally-shaped numpy arrays from all ranks to rank `dst` (returns the list there, None elsewhere).
    Works on any initialised torch.distributed backend; with NCCL pass the rank's CUDA device."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return [arr]
    t = torch.from_numpy(np.ascontiguousarray(arr))
    if device is not None:
        t = t.to(device)
    out = [torch.empty_like(t) for _ in range(dist.get_world_size())] if dist.get_rank() == dst else None
    dist.gather(t, out, dst=dst)
    return [o.cpu().numpy() for o in out] if out is not None else None


def all_max(x, device=None):
    """Max of a Python float over all ranks (timing is reported as the slowest rank's)."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(x)
    t = torch.tensor([x], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
