"""Parity at the shapes BASELINE.json names and bench.py measures (round 2).

Fixtures tests/golden/r2/*.npz come from the reference's own library on the CPU oracle (tools/gen_golden_r2.py); the
cases and the driver functions are shared (tests/golden_r2.py). Bars (BASELINE.json north_star):
  * per-step UNet output: max-relative error <= 1e-2 in the global norm max|a-b|/max|b| (THE bar, asserted). The two
    element-wise readings of golden_r2.rel_errs (absolute floor of 10 % of max|ref|; pure ratio over |ref| >= 10 % of
    max|ref|) are printed for every case and asserted <= 3e-2: with f16 operands on both sides an absolute error of
    ~1e-3 max|ref| is the floor of the arithmetic, so those ratios sit around 1e-2 by construction (measured 2e-3 ...
    1.1e-2) and are a sanity bound, not the bar;
  * final latent <= 1e-2, decoded image PSNR >= 35 dB.
"""
import os, sys
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import golden_r2 as G   # noqa: E402

_ctx_cache = {}


def engine_ctx(kind):
    """One context per checkpoint kind, kept for the module (weights stay resident)."""
    import bench
    from mlimgsynth_b200 import api
    if kind not in _ctx_cache:
        for c in _ctx_cache.values():       # one model resident at a time: SDXL + SD2 + SD1 weights and arenas add up
            c.close()
        _ctx_cache.clear()
        _ctx_cache[kind] = api.Ctx(model=bench.weights_path(kind))
    return _ctx_cache[kind]


@pytest.fixture(scope="module", autouse=True)
def _cleanup():
    yield
    for c in _ctx_cache.values():
        c.close()
    _ctx_cache.clear()


def seed_set(ctx, seed):
    ctx.set("seed", "%d,0" % seed)


def report(what, a, b, bar=1e-2):
    g, m, e = G.rel_errs(a, b)
    print("PARITY %s: max-rel err global %.3e | element-wise with floor %.3e | element-wise (|ref| >= 10%% max) %.3e" % (what, g, m, e))
    assert np.isfinite(a).all()
    assert g <= bar, (what, g)
    assert m <= 3 * bar, (what, m)
    assert e <= 3 * bar, (what, e)
    return g, m, e


@pytest.mark.parametrize("name", list(G.UNET_CASES))
def test_unet_evaluation_matches_reference(name):
    """ONE UNet evaluation (unet.c:460-497 + the CFG combine mlimgsynth.c:1572-1587) at full size, through mlis_generate
    with a one-step Euler run -- the same call sequence that produced the fixture on the reference."""
    from mlimgsynth_b200 import api
    c = G.UNET_CASES[name]
    z = G.load(name)
    ctx = engine_ctx(c["model"])
    out = G.unet_step(ctx, api, name, list(range(c["n"])), seed_set)
    ctx.set("no_decode", 0); ctx.set("batch_size", 1)
    s0 = float(z["sigma0"])
    dx_eng = (z["x"] - out) / s0
    if c["cfg"] > 1:
        # CFG combine of the two halves (mlimgsynth.c:1583: dx = f dx_c + (1 - f) dx_u). The fixture holds the reference's
        # per-evaluation outputs; an error e on each half appears as up to (f + |1 - f|) e = 13 e in the combination, so the
        # combined error is measured against f max|dx_c| + |1 - f| max|dx_u| (and the plain ratio is printed).
        f = c["cfg"]
        dc, du = (z["x"] - z["out"]) / s0, (z["x"] - z["out_u"]) / s0
        dx_ref = f * dc + (1 - f) * du
        d = np.abs(dx_eng - dx_ref)
        budget = f * np.abs(dc).max() + abs(1 - f) * np.abs(du).max()
        print("PARITY %s CFG-combined dx: max|err| / (f max|dx_c| + |1-f| max|dx_u|) %.3e | plain max|err|/max|dx| %.3e" % (name, d.max() / budget, d.max() / np.abs(dx_ref).max()))
        assert np.isfinite(dx_eng).all() and d.max() / budget <= 1e-2
        for i in range(c["n"]):     # every image of the batch on its own (a swapped / duplicated image cannot hide in the max)
            assert np.abs(dx_eng[i] - dx_ref[i]).max() / budget <= 1e-2, (name, i)
        # ... and each of the 16 evaluations of the batched launch on its own, through mlis_unet_eval
        _, cond, ncond, _, _ = G.unet_inputs(name)
        n = c["n"]
        xb = np.concatenate([z["x"], z["x"]])
        cb = np.concatenate([np.repeat(cond[None], n, 0), np.repeat(ncond[None], n, 0)])
        dxb = ctx.unet_eval(xb, cb, None, s0)
        report(name + " 16 evaluations in one launch", dxb, np.concatenate([dc, du]))
        for i in range(2 * n):
            assert G.rel_errs(dxb[i], np.concatenate([dc, du])[i])[0] <= 1e-2, (name, i)
        return
    dx_ref = (z["x"] - z["out"]) / s0
    report(name + " dx", dx_eng, dx_ref)


@pytest.mark.parametrize("name", ["unet_sd15_64_hi", "unet_sd15_64_mid", "unet_sd21_96", "unet_sdxl_128"])
def test_mlis_unet_eval_matches_reference(name):
    """The entry point bench.py times (mlis_unet_eval, include/mlimgsynth_b200.h) on the fixture's x, alone and as a batch
    of 16 copies (the benchmarked UNet batch): dx vs the reference's dx."""
    c = G.UNET_CASES[name]
    z = G.load(name)
    ctx = engine_ctx(c["model"])
    _, cond, _, label, _ = G.unet_inputs(name)
    s0 = float(z["sigma0"])
    dx_ref = (z["x"] - z["out"]) / s0
    dx = ctx.unet_eval(z["x"], cond[None], label[None] if label is not None else None, s0)
    report(name + " unet_eval", dx, dx_ref)
    if c["model"] == "sd1":
        nb = 16
        xb = np.repeat(z["x"], nb, axis=0)
        dxb = ctx.unet_eval(xb, np.repeat(cond[None], nb, axis=0), None, s0)
        for i in range(nb):
            g = G.rel_errs(dxb[i], dx_ref[0])[0]
            assert g <= 1e-2, (name, "batch16", i, g)
        print("%s batch-16 evaluation: every image within 1e-2" % name)


def check_image(what, lat_g, img_g, lat_c, img_c):
    g, m, e = G.rel_errs(lat_g, lat_c)
    p = G.psnr_u8(img_g, img_c)
    print("PARITY %s: final latent max-rel err global %.3e | with floor %.3e | element-wise (|ref| >= 10%% max) %.3e | image PSNR %.1f dB" % (what, g, m, e, p))
    assert np.isfinite(lat_g).all() and g <= 1e-2, g
    assert p >= 35.0, p


def test_config1_sd15_512_euler20_cfg7_seed42():
    """BASELINE configs[0] in full: prompt -> tokens -> CLIP -> 20 Euler steps at cfg 7 (40 UNet evaluations) -> VAE -> RGB8."""
    from mlimgsynth_b200 import api
    z = G.load("c1_sd15_512_euler20")
    ctx = engine_ctx("sd1")
    for k, v in dict(method="euler", scheduler="uniform", s_noise=0, s_ancestral=0, steps=20, cfg_scale=7, image_dim=(512, 512),
                     batch_size=1, no_decode=0, vae_tile=0).items():
        ctx.set(k, v)
    seed_set(ctx, 42); ctx.set("prompt", G.PROMPT)
    ctx.generate()
    check_image("config 1", ctx.tensor(api.TENSOR_LATENT), ctx.image(0), z["latent"], z["image"])


def test_config4_img2img_inpaint_two_loras_512x768(tmp_path):
    """BASELINE configs[3]: img2img + inpainting (f_t_ini 0.7 -> 14 steps) at 512x768 with TWO LoRAs merged on the device
    one after the other, one f16 rounding per merge (lora.c:46-78, :97-138)."""
    import bench, gen_weights
    from mlimgsynth_b200 import api
    z = G.load("c4_img2img_inpaint_2lora_512x768")
    rgb, mask = G.c4_inputs()
    c = api.Ctx(model=bench.weights_path("sd1"), steps=20, method="euler", cfg_scale=7)
    try:
        for i, l in enumerate(G.C4_LORAS):
            p = str(tmp_path / ("lora_c4_%d.safetensors" % i))
            gen_weights.write_lora(p, "sd1", rank=l["rank"], alpha=l["alpha"], seed=l["seed"])
            c.set("lora", (p, l["mult"]))
        c.set("f_t_ini", 0.7)
        seed_set(c, 42)
        c.set_image(rgb); c.set_image(mask, mask=True); c.set("prompt", G.PROMPT)
        c.generate()
        check_image("config 4", c.tensor(api.TENSOR_LATENT), c.image(0), z["latent"], z["image"])
    finally:
        c.close()


def _decode_u8(ctx, lat):
    img = ctx.decode(lat)
    return np.clip(np.transpose(img[0], (1, 2, 0)) * 255.0, 0, 255).astype(np.uint8)


def test_config5_sdxl_tiled_vae_decode_2048():
    """BASELINE configs[4]: SDXL VAE decode of a 256x256 latent with vae-tile 512 = 16 tiles of 80x80 (vae.c:331-391)."""
    z = G.load("c5_sdxl_vae_tile512_2048")
    ctx = engine_ctx("sdxl")
    ctx.set("vae_tile", 512)
    u8 = _decode_u8(ctx, G.c5_latent())
    ctx.set("vae_tile", 0)
    assert u8.shape == (2048, 2048, 3)
    p = G.psnr_u8(G.sub(u8), z["image_sub"])
    print("PARITY config 5 tiled decode 2048x2048: PSNR %.1f dB" % p)
    assert p >= 35.0


def test_config5_tae_decode_2048():
    """BASELINE configs[4], second half: TAE decode of the same latent, full frame (tae.c:117-136)."""
    import bench
    from mlimgsynth_b200 import api
    z = G.load("c5_tae_2048")
    c = api.Ctx(model=bench.weights_path("sd1"), tae=bench.weights_path("tae"))
    try:
        u8 = _decode_u8(c, G.c5_latent())
        p = G.psnr_u8(G.sub(u8), z["image_sub"])
        print("PARITY config 5 TAE decode 2048x2048: PSNR %.1f dB" % p)
        assert p >= 35.0
    finally:
        c.close()
