#!/usr/bin/env python3
"""Generates tests/golden/r2/*.npz: results of the REFERENCE's own library (oracle/_ref/libmlimgsynth_cpu.so = the
unmodified reference objects compiled from /root/reference + the CPU restatement of the ggml ops) at the shapes
BASELINE.json names. Cases and driver functions: tests/golden_r2.py (shared with the GPU tests).
Weights: tools/gen_weights.py (seed 1234), identical on the GPU box. Run in the build container (CPU, ~40 min);
the fixtures are committed.   usage: gen_golden_r2.py [case ...] [--force]"""
import ctypes as C, os, subprocess, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import gen_weights
import golden_r2 as G
from mlimgsynth_b200 import api

REF = os.path.join(ROOT, "oracle", "_ref")
tmp = "/tmp/mlis_golden"; os.makedirs(tmp, exist_ok=True); os.makedirs(G.GOLD, exist_ok=True)


class Rng(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("offset", C.c_uint32)]


def weights(kind):
    p = os.path.join(tmp, kind + ".safetensors")
    if not os.path.exists(p):
        gen_weights.write_safetensors(p, gen_weights.build_spec(kind), 1234, "f16")
    return p


L = api.bind(C.CDLL(os.path.join(REF, "libmlimgsynth_cpu.so")), extensions=False)
L.unet_t_to_sigma.restype = C.c_float
L.unet_t_to_sigma.argtypes = [C.c_void_p, C.c_float]
g_rng = Rng.in_dll(L, "g_rng")


def seed_set(ctx, seed):
    ctx.set("seed", seed)
    g_rng.offset = 0


def randn(seed, n):
    r = Rng(seed, 0); buf = (C.c_float * n)()
    L.rng_philox_randn(C.byref(r), n, buf)
    return np.frombuffer(buf, dtype=np.float32).copy()


def gen_unet(name):
    c = G.UNET_CASES[name]
    ctx = api.Ctx(_lib=L, model=weights(c["model"]), log_level="info")
    ctx.setup()
    sigma0 = float(L.unet_t_to_sigma(C.addressof(C.c_char.in_dll(L, {"sd1": "g_unet_sd1", "sd2": "g_unet_sd2", "sdxl": "g_unet_sdxl"}[c["model"]])), 999.0 * c["f_t_ini"]))
    lat = G.unet_inputs(name)[0]
    outs, xs, outs_u = [], [], []
    for i in range(c["n"]):
        t0 = time.time()
        if c["cfg"] > 1:      # the two halves separately (cfg 1 each): the per-evaluation outputs; the test forms the combine
            outs.append(G.unet_step(ctx, api, name, [i], seed_set, half="cond")[0])
            outs_u.append(G.unet_step(ctx, api, name, [i], seed_set, half="ncond")[0])
        else:
            outs.append(G.unet_step(ctx, api, name, [i], seed_set)[0])
        xs.append(lat[i] + sigma0 * randn(42 + i, lat[i].size).reshape(lat[i].shape))
        print("  %s image %d: %.1f s" % (name, i, time.time() - t0), flush=True)
    ctx.close()
    out, x = np.stack(outs), np.stack(xs).astype(np.float32)
    extra = {"out_u": np.stack(outs_u)} if outs_u else {}
    np.savez_compressed(os.path.join(G.GOLD, name + ".npz"), out=out, x=x, sigma0=np.float32(sigma0), **extra)
    dx = (x - out) / sigma0
    print(name, "sigma0 %.4f  |dx| max %.3f rms %.3f" % (sigma0, np.abs(dx).max(), np.sqrt((dx ** 2).mean())), flush=True)


def gen_c1():
    """Config 1: SD1.5 txt2img 512x512, 20 Euler steps, cfg 7, seed 42 -- the whole path, prompt to RGB8."""
    ctx = api.Ctx(_lib=L, model=weights("sd1"), log_level="info", image_dim=(512, 512), steps=20, method="euler", cfg_scale=7)
    seed_set(ctx, 42); ctx.set("prompt", G.PROMPT)
    ctx.generate()
    np.savez_compressed(os.path.join(G.GOLD, "c1_sd15_512_euler20.npz"), latent=ctx.tensor(api.TENSOR_LATENT), image=ctx.image(0))
    ctx.close()


def lora_paths():
    ps = []
    for i, l in enumerate(G.C4_LORAS):
        p = os.path.join(tmp, "lora_c4_%d.safetensors" % i)
        if not os.path.exists(p):
            gen_weights.write_lora(p, "sd1", rank=l["rank"], alpha=l["alpha"], seed=l["seed"])
        ps.append(p)
    return ps


def gen_c4():
    """Config 4: SD1.5 img2img + inpainting, f_t_ini 0.7, 512x768, two LoRAs merged one after the other (lora.c:97-138)."""
    rgb, mask = G.c4_inputs()
    ctx = api.Ctx(_lib=L, model=weights("sd1"), log_level="info", steps=20, method="euler", cfg_scale=7)
    for p, l in zip(lora_paths(), G.C4_LORAS):
        ctx.set("lora", (p, l["mult"]))
    ctx.set("f_t_ini", 0.7)
    seed_set(ctx, 42)
    ctx.set_image(rgb); ctx.set_image(mask, mask=True); ctx.set("prompt", G.PROMPT)
    ctx.generate()
    np.savez_compressed(os.path.join(G.GOLD, "c4_img2img_inpaint_2lora_512x768.npz"), latent=ctx.tensor(api.TENSOR_LATENT), image=ctx.image(0))
    ctx.close()


def gen_c5():
    """Config 5: SDXL VAE, tiled decode (vae-tile 512: 16 tiles of 80x80, vae.c:331-391) of a 256x256 latent -> 2048x2048."""
    lat = G.c5_latent()
    ctx = api.Ctx(_lib=L, model=weights("sdxl"), log_level="info", vae_tile=512)
    img = ctx.decode(lat)        # [1,3,2048,2048] in [0,1]
    u8 = np.clip(np.transpose(img[0], (1, 2, 0)) * 255.0, 0, 255).astype(np.uint8)      # truncation as mlimgsynth.c:112-129
    np.savez_compressed(os.path.join(G.GOLD, "c5_sdxl_vae_tile512_2048.npz"), image_sub=G.sub(u8))
    ctx.close()


def gen_c5_tae():
    """Config 5, second half: TAE decode of the same latent, full frame (tae.c:117)."""
    lat = G.c5_latent()
    ctx = api.Ctx(_lib=L, model=weights("sd1"), tae=weights("tae"), log_level="info")
    img = ctx.decode(lat)
    u8 = np.clip(np.transpose(img[0], (1, 2, 0)) * 255.0, 0, 255).astype(np.uint8)
    np.savez_compressed(os.path.join(G.GOLD, "c5_tae_2048.npz"), image_sub=G.sub(u8))
    ctx.close()


JOBS = {n: (lambda n=n: gen_unet(n)) for n in G.UNET_CASES}
JOBS.update({"c1_sd15_512_euler20": gen_c1, "c4_img2img_inpaint_2lora_512x768": gen_c4, "c5_sdxl_vae_tile512_2048": gen_c5, "c5_tae_2048": gen_c5_tae})

if __name__ == "__main__":
    names = [a for a in sys.argv[1:] if not a.startswith("-")] or list(JOBS)
    if len(names) > 1:      # one process per case: the reference keeps one global noise stream and caches per process
        for n in names:
            if os.path.exists(os.path.join(G.GOLD, n + ".npz")) and "--force" not in sys.argv:
                continue
            t0 = time.time()
            r = subprocess.run([sys.executable, os.path.abspath(__file__), n, "--force"])
            print("== %s rc %d %.0f s" % (n, r.returncode, time.time() - t0), flush=True)
    else:
        JOBS[names[0]]()
