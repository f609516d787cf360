"""Parity cases at the shapes BASELINE.json names (round 2), shared by tools/gen_golden_r2.py (the reference's own
library on the CPU oracle -> committed fixtures tests/golden/r2/*.npz) and tests/test_parity_r2_gpu.py (the engine vs
those fixtures). The SAME driver functions run on both sides through the mlis_* C API (api.Ctx over either library), so a
fixture and its test cannot drift apart.

One UNet evaluation through the public API (reference: unet_denoise_run, unet.c:460-497, reached through
mlis_generate -> dnsamp_step -> solver_euler_step): a ONE-step Euler generation with caller latent + conditioning
(MLIS_TUF_LATENT | MLIS_TUF_CONDITIONING, mlimgsynth.c:1664-1700), no decode.  sampling.c:129-133 forms
x = latent + sigma0 * randn(seed, offset 0), solvers.c:82-98 returns x - sigma0 * dx, so
    dx = (x - out) / sigma0        with sigma0 = unet_t_to_sigma(999 * f_t_ini)
is the (CFG-combined, mlimgsynth.c:1572-1587) UNet output of that evaluation. x is identical on both sides because the
Philox stream is bit-exact (tests/test_host_cpu.py), and is stored in the fixture.
"""
import os
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden", "r2")
PROMPT = "a photograph of an astronaut riding a horse"

# name -> model kind, latent (w,h), images (seeds 42+i), cfg scale, f_t_ini (sigma0 = sigma(999 f)), latent0 scale
UNET_CASES = {
    # SD1.5 512x512 (config 1): latent 64x64, sigma0 = sigma_max = 14.61 (the first evaluation of every txt2img run)
    "unet_sd15_64_hi":   dict(model="sd1", lw=64, lh=64, n=1, cfg=1.0, f_t_ini=1.0, lat_scale=0.0),
    # ... and in the middle of the trajectory, around a non-trivial latent
    "unet_sd15_64_mid":  dict(model="sd1", lw=64, lh=64, n=1, cfg=1.0, f_t_ini=0.5, lat_scale=0.8),
    # the BENCHMARKED shape: 8 images x (cond | uncond) = 16 latents in one evaluation, CFG 7 combine
    "unet_sd15_64_b8cfg": dict(model="sd1", lw=64, lh=64, n=8, cfg=7.0, f_t_ini=1.0, lat_scale=0.0),
    # SD2.1 768x768 v-prediction (config 2): latent 96x96, 9216-token attention with 64-wide heads
    "unet_sd21_96":      dict(model="sd2", lw=96, lh=96, n=1, cfg=1.0, f_t_ini=0.6, lat_scale=0.8),
    # SDXL 1024x1024 (config 3): latent 128x128, label vector, transformer depth 2 / 10
    "unet_sdxl_128":     dict(model="sdxl", lw=128, lh=128, n=1, cfg=1.0, f_t_ini=0.7, lat_scale=0.8),
}

N_CTX = {"sd1": 768, "sd2": 1024, "sdxl": 2048}
N_ADM = {"sd1": 0, "sd2": 0, "sdxl": 2816}


def unet_inputs(name):
    """Deterministic synthetic inputs: initial latent [n,4,h,w], cond / ncond [77,n_ctx], label / nlabel [adm]."""
    c = UNET_CASES[name]
    r = np.random.default_rng(abs(hash_name(name)))
    lat = (r.standard_normal((c["n"], 4, c["lh"], c["lw"])) * c["lat_scale"]).astype(np.float32)
    cond = (r.standard_normal((77, N_CTX[c["model"]])) * 0.5).astype(np.float32)
    ncond = (r.standard_normal((77, N_CTX[c["model"]])) * 0.5).astype(np.float32)
    adm = N_ADM[c["model"]]
    label = (r.standard_normal((adm,)) * 0.5).astype(np.float32) if adm else None
    nlabel = (r.standard_normal((adm,)) * 0.5).astype(np.float32) if adm else None
    return lat, cond, ncond, label, nlabel


def hash_name(name):
    h = 0
    for ch in name.encode():
        h = (h * 131 + ch) % 1000003
    return h


def unet_step(ctx, api, name, images, seed_setter, half=None):
    """Run the one-step generation of `name` for image indices `images` (one generate() call; the reference takes one
    image per call). seed_setter(ctx, seed) must set seed AND reset the noise offset to 0. Returns out [len,4,h,w].
    half = "cond" / "ncond": ONE CFG half alone (cfg scale 1 with that conditioning) -- the per-evaluation outputs the
    CFG combine is made of."""
    c = UNET_CASES[name]
    lat, cond, ncond, label, nlabel = unet_inputs(name)
    cfg = c["cfg"]
    if half is not None:
        cfg = 1.0
        if half == "ncond":
            cond, label = ncond, nlabel
    nb = len(images)
    for k, v in dict(method="euler", scheduler="uniform", s_noise=0, s_ancestral=0, steps=1, cfg_scale=cfg,
                     image_dim=(c["lw"] * 8, c["lh"] * 8), no_decode=1).items():
        ctx.set(k, v)
    if nb > 1:
        ctx.set("batch_size", nb)
    ctx.set("f_t_ini", c["f_t_ini"])
    ctx.tensor_set(api.TENSOR_LATENT, lat[images])
    ctx.tensor_set(api.TENSOR_COND, cond)
    ctx.tensor_set(api.TENSOR_NCOND, ncond)
    if label is not None:
        ctx.tensor_set(api.TENSOR_LABEL, label)
        ctx.tensor_set(api.TENSOR_NLABEL, nlabel)
    ctx.set("tensor_use_flags", api.TUF_LATENT | api.TUF_CONDITIONING)
    seed_setter(ctx, 42 + images[0])
    ctx.generate()
    out = ctx.tensor(api.TENSOR_LATENT)
    return out.reshape(nb, 4, c["lh"], c["lw"])


# ---- full-size end-to-end cases (BASELINE configs 1, 4, 5)
def c4_inputs():
    """Config 4: 512x768 RGB8 input (uniform random bytes, seed 7) + mask whose centre rectangle is repainted."""
    rng = np.random.default_rng(7)
    w, h = 512, 768
    rgb = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
    mask = np.full((h, w), 255, dtype=np.uint8)
    mask[h // 4: 3 * h // 4, w // 4: 3 * w // 4] = 0
    return rgb, mask


C4_LORAS = [dict(rank=16, alpha=16.0, seed=77, mult=0.8), dict(rank=16, alpha=16.0, seed=78, mult=0.5)]


def c5_latent(kind="sdxl"):
    """Config 5: latent [256,256,4] = N(0,1) * 0.18, seed 42."""
    return (np.random.default_rng(42).standard_normal((1, 4, 256, 256)) * 0.18).astype(np.float32)


def sub(img, s=3):
    """Fixture-size reduction for the 2048x2048 images: every s-th pixel of every s-th row (PSNR is computed on the same
    sub-lattice of the engine's image; it still crosses every tile and every tile seam)."""
    return np.ascontiguousarray(img[::s, ::s])


def load(name):
    p = os.path.join(GOLD, name + ".npz")
    if not os.path.exists(p):
        import pytest
        pytest.skip("fixture %s not generated yet (tools/gen_golden_r2.py)" % name)
    return np.load(p)


def psnr_u8(a, b):
    mse = float(((a.astype(np.float32) - b.astype(np.float32)) ** 2).mean()) / 255.0 ** 2
    return 10 * np.log10(1.0 / mse) if mse > 0 else 99.0


def rel_errs(a, b):
    """Relative errors of a against the reference b, three readings of "max-relative error":
    global  = max|a-b| / max|b|                         (the bar used since round 1: <= 1e-2)
    mixed   = max over ALL elements of |a-b| / (|b| + 0.1 max|b|)   (element-wise with an absolute floor of 10 % of the
              largest value, i.e. numpy.allclose(rtol = bar, atol = 0.1 bar max|b|); <= 1e-2)
    element = max over elements with |b| >= 0.1 max|b| of |a-b| / |b|     (pure element-wise ratio where the reference value
              is not small; for smaller |b| the ratio of two f16-operand computations is unbounded: both sides round
              operands to f16, an absolute error of ~1e-3 max|b| is the floor of the arithmetic, not of this engine)"""
    a = a.astype(np.float64); b = b.astype(np.float64)
    d = np.abs(a - b); m = np.abs(b).max()
    sel = np.abs(b) >= 0.1 * m
    return float(d.max() / m), float((d / (np.abs(b) + 0.1 * m)).max()), float((d[sel] / np.abs(b[sel])).max())
