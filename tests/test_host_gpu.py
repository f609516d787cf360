"""The B200 host layer (libmlimgsynth_b200.so: device-resident sampler, batched CFG, cached graphs)
against the reference's own host code running on the CPU oracle (oracle/_ref/mlimgsynth_cpu), same
random-init checkpoint, prompt, seed and options. Bars: latent max-rel error <= 1e-2, image PSNR >= 35 dB."""
import os, subprocess, sys
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from test_e2e_gpu import load_tensor, load_pnm   # noqa: E402
PROMPT = "a photograph of an (astronaut:1.2) riding a [horse]"


@pytest.fixture(scope="module")
def weights(tmp_path_factory):
    import gen_weights
    d = tmp_path_factory.mktemp("w")
    p = str(d / "sd1.safetensors")
    gen_weights.write_safetensors(p, gen_weights.build_spec("sd1"), 1234, "f16")
    lp = str(d / "lora1.safetensors")
    gen_weights.write_lora(lp, "sd1", rank=8, alpha=8.0, seed=5)
    return p, lp


@pytest.fixture(scope="module")
def ctx(weights):
    from mlimgsynth_b200 import api
    c = api.Ctx(model=weights[0])
    yield c
    c.close()


def ref_cli(model, out, extra):
    cmd = [os.path.join(ROOT, "oracle", "_ref", "mlimgsynth_cpu"), "generate", "-m", model, "-p", PROMPT, "-S", "42",
           "-o", out + ".pnm", "--olatent", out + ".tensor"] + extra
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=1200)
    assert r.returncode == 0, r.stderr[-1500:]
    return load_tensor(out + ".tensor"), load_pnm(out + ".pnm")


def compare(lat_g, img_g, lat_c, img_c, what):
    from mlimgsynth_b200 import api  # noqa
    err = np.abs(lat_g - lat_c).max() / np.abs(lat_c).max()
    mse = float(((img_g.astype(np.float32) / 255.0 - img_c) ** 2).mean())
    psnr = 10 * np.log10(1.0 / mse) if mse > 0 else 99.0
    print("%s: latent max-rel err %.3e, PSNR %.1f dB" % (what, err, psnr))
    assert np.isfinite(lat_g).all() and err <= 1e-2, err
    assert psnr >= 35.0, psnr


CASES = [
    ("euler", dict(method="euler", steps=3, cfg_scale=7), ["-s", "3", "--method", "euler", "--cfg-scale", "7"]),
    ("heun", dict(method="heun", steps=4, cfg_scale=3), ["-s", "4", "--method", "heun", "--cfg-scale", "3"]),
    ("taylor3", dict(method="taylor3", steps=4, cfg_scale=1), ["-s", "4", "--method", "taylor3", "--cfg-scale", "1"]),
    ("dpmpp2m_karras", dict(method="dpmpp2m", scheduler="karras", steps=4, cfg_scale=5), ["-s", "4", "--method", "dpm++2m", "--scheduler", "karras", "--cfg-scale", "5"]),
    ("dpmpp2s_a", dict(method="dpmpp2s", s_ancestral=1, steps=4, cfg_scale=2), ["-s", "4", "--method", "dpm++2s_a", "--cfg-scale", "2"]),
    ("euler_snoise", dict(method="euler", s_noise=1, steps=3, cfg_scale=1), ["-s", "3", "--method", "euler", "--s-noise", "1", "--cfg-scale", "1"]),
]


@pytest.mark.parametrize("name,opts,cli", CASES, ids=[c[0] for c in CASES])
def test_txt2img_matches_reference_host(ctx, weights, tmp_path, name, opts, cli):
    from mlimgsynth_b200 import api
    lat_c, img_c = ref_cli(weights[0], str(tmp_path / name), ["-d", "128,128"] + cli)
    for k in ("s_noise", "s_ancestral"):
        ctx.set(k, opts.get(k, 0))
    ctx.set("scheduler", opts.get("scheduler", "uniform"))
    for k, v in opts.items():
        ctx.set(k, v)
    ctx.set("image_dim", (128, 128)); ctx.set("batch_size", 1); ctx.set("seed", "42,0"); ctx.set("prompt", PROMPT)
    ctx.generate()
    compare(ctx.tensor(api.TENSOR_LATENT), ctx.image(0), lat_c, img_c, name)


def test_batch_equals_seed_loop(ctx):
    """Image i of a batch == a single generation with seed + i (generate.sh:55-61 semantics)."""
    from mlimgsynth_b200 import api
    for k, v in dict(method="euler", scheduler="uniform", s_noise=0, s_ancestral=0, steps=3, cfg_scale=7, image_dim=(128, 192)).items():
        ctx.set(k, v)
    ctx.set("batch_size", 3); ctx.set("seed", "100,0"); ctx.set("prompt", PROMPT)
    ctx.generate()
    lat_b = ctx.tensor(api.TENSOR_LATENT); imgs = [ctx.image(i) for i in range(3)]
    assert lat_b.shape == (3, 4, 24, 16)
    for i in range(3):
        ctx.set("batch_size", 1); ctx.set("seed", "%d,0" % (100 + i)); ctx.set("prompt", PROMPT)
        ctx.generate()
        lat = ctx.tensor(api.TENSOR_LATENT)
        assert np.abs(lat[0] - lat_b[i]).max() / np.abs(lat).max() <= 1e-2
        d = np.abs(ctx.image(0).astype(int) - imgs[i].astype(int))
        assert d.mean() < 1.0


def test_img2img_inpaint_lora(ctx, weights, tmp_path):
    """Config 4 shape: img2img + alpha-mask inpainting (f_t_ini 0.7) with a LoRA merged on device."""
    from mlimgsynth_b200 import api
    rng = np.random.default_rng(7)
    w, h = 128, 192
    rgb = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
    mask = np.full((h, w), 255, dtype=np.uint8); mask[h // 4: 3 * h // 4, w // 4: 3 * w // 4] = 0   # centre rectangle is repainted
    ppm, pgm = str(tmp_path / "in.ppm"), str(tmp_path / "mask.pgm")
    with open(ppm, "wb") as f:
        f.write(b"P6\n%d %d\n255\n" % (w, h)); f.write(rgb.tobytes())
    with open(pgm, "wb") as f:
        f.write(b"P5\n%d %d\n255\n" % (w, h)); f.write(mask.tobytes())
    cli = ["-i", ppm, "--imask", pgm, "--f-t-ini", "0.7", "-s", "5", "--method", "euler", "--cfg-scale", "4", "--lora", "%s,0.8" % weights[1]]
    lat_c, img_c = ref_cli(weights[0], str(tmp_path / "i2i"), cli)
    c2 = api.Ctx(model=weights[0])
    c2.set("lora", (weights[1], 0.8))
    for k, v in dict(method="euler", steps=5, cfg_scale=4, f_t_ini=0.7, seed="42,0").items():
        c2.set(k, v)
    c2.set_image(rgb); c2.set_image(mask, mask=True); c2.set("prompt", PROMPT)
    c2.generate()
    compare(c2.tensor(api.TENSOR_LATENT), c2.image(0), lat_c, img_c, "img2img+inpaint+lora")
    c2.close()


def test_tiled_vae_decode_matches_untiled_reference(ctx, weights, tmp_path):
    """VAE tiling geometry (vae.c:331-391): tiled decode on the engine == tiled decode of the reference."""
    from mlimgsynth_b200 import api
    lat = (np.random.default_rng(3).standard_normal((1, 4, 40, 40)) * 0.18).astype(np.float32)
    p = str(tmp_path / "lat.tensor")
    with open(p, "wb") as f:
        f.write(b"TENSOR F32 40 40 4 1\n"); f.write(lat.tobytes())
    out = str(tmp_path / "dec.pnm")
    r = subprocess.run([os.path.join(ROOT, "oracle", "_ref", "mlimgsynth_cpu"), "vae-decode", "-m", weights[0], "--ilatent", p,
                        "--vae-tile", "128", "-o", out], capture_output=True, text=True, timeout=1200)
    assert r.returncode == 0, r.stderr[-1500:]
    img_c = load_pnm(out)
    ctx.set("vae_tile", 128)
    img = ctx.decode(lat)       # [1,3,H,W] in [0,1]
    ctx.set("vae_tile", 0)
    img_g = np.clip(np.transpose(img[0], (1, 2, 0)), 0, 1)
    mse = float(((img_g - img_c) ** 2).mean())
    psnr = 10 * np.log10(1.0 / mse)
    print("tiled decode PSNR %.1f dB" % psnr)
    assert psnr >= 35.0
