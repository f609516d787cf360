"""End-to-end parity cases shared by tools/gen_golden_e2e.py (reference on the CPU oracle -> fixtures)
and the GPU tests (engine vs fixtures). `cli` are reference CLI arguments, `opts` the same settings for
the mlis_* API."""
import os
import numpy as np

PROMPT = "a photograph of an (astronaut:1.2) riding a [horse]"
PROMPT_PLAIN = "a photograph of an astronaut riding a horse"

CASES = {
    # reference host code on both backends (tests/test_e2e_gpu.py)
    "ref_euler_cfg": dict(prompt=PROMPT_PLAIN, cli=["-d", "128,128", "-s", "3", "--method", "euler", "--cfg-scale", "7"]),
    "ref_dpmpp2m_karras": dict(prompt=PROMPT_PLAIN, cli=["-d", "192,128", "-s", "4", "--method", "dpm++2m", "--scheduler", "karras", "--cfg-scale", "5"]),
    "ref_euler_a": dict(prompt=PROMPT_PLAIN, cli=["-d", "128,128", "-s", "3", "--method", "euler_a", "--cfg-scale", "1"]),
    # B200 host layer (tests/test_host_gpu.py)
    "euler": dict(opts=dict(method="euler", steps=3, cfg_scale=7), cli=["-d", "128,128", "-s", "3", "--method", "euler", "--cfg-scale", "7"]),
    "heun": dict(opts=dict(method="heun", steps=4, cfg_scale=3), cli=["-d", "128,128", "-s", "4", "--method", "heun", "--cfg-scale", "3"]),
    "taylor3": dict(opts=dict(method="taylor3", steps=4, cfg_scale=1), cli=["-d", "128,128", "-s", "4", "--method", "taylor3", "--cfg-scale", "1"]),
    "dpmpp2m_karras": dict(opts=dict(method="dpmpp2m", scheduler="karras", steps=4, cfg_scale=5),
                           cli=["-d", "128,128", "-s", "4", "--method", "dpm++2m", "--scheduler", "karras", "--cfg-scale", "5"]),
    "dpmpp2s_a": dict(opts=dict(method="dpmpp2s", s_ancestral=1, steps=4, cfg_scale=2), cli=["-d", "128,128", "-s", "4", "--method", "dpm++2s_a", "--cfg-scale", "2"]),
    "euler_snoise": dict(opts=dict(method="euler", s_noise=1, steps=3, cfg_scale=1), cli=["-d", "128,128", "-s", "3", "--method", "euler", "--s-noise", "1", "--cfg-scale", "1"]),
    "img2img_inpaint_lora": dict(opts=dict(method="euler", steps=5, cfg_scale=4, f_t_ini=0.7),
                                 cli=["-i", "@TMP@/in.ppm", "--imask", "@TMP@/mask.pgm", "--f-t-ini", "0.7", "-s", "5", "--method", "euler", "--cfg-scale", "4", "--lora", "@LORA@,0.8"]),
    "vae_tiled_decode": dict(cmd="vae-decode", cli=["--ilatent", "@TMP@/lat.tensor", "--vae-tile", "128"]),
    # the other model families of BASELINE.json's configs (model = checkpoint kind of tools/gen_weights.py)
    # config 2 shape: SD2.x v-prediction (unet.c:54,490-494), OpenCLIP ViT-H text encoder, d_head 64, DPM++(2M)
    "sd2_vpred_dpmpp2m": dict(model="sd2", opts=dict(method="dpmpp2m", steps=4, cfg_scale=5, image_dim=(128, 192)),
                              cli=["-d", "128,192", "-s", "4", "--method", "dpm++2m", "--cfg-scale", "5"]),
    # config 3 shape: SDXL base (two text encoders, pooled-feature + size label, transformer depth 2 / 10)
    "sdxl_euler_cfg": dict(model="sdxl", opts=dict(method="euler", steps=3, cfg_scale=7, image_dim=(192, 128)),
                           cli=["-d", "192,128", "-s", "3", "--method", "euler", "--cfg-scale", "7"]),
    # config 5: TAE decode instead of the VAE (tae.c:117)
    "sd1_tae_decode": dict(model="sd1", tae=True, opts=dict(method="euler", steps=2, cfg_scale=3, image_dim=(128, 128)),
                           cli=["-d", "128,128", "-s", "2", "--method", "euler", "--cfg-scale", "3", "--tae", "@TAE@"]),
}


def inputs():
    """Deterministic synthetic inputs of the img2img / decode cases."""
    rng = np.random.default_rng(7)
    w, h = 128, 192
    rgb = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
    mask = np.full((h, w), 255, dtype=np.uint8)
    mask[h // 4: 3 * h // 4, w // 4: 3 * w // 4] = 0      # centre rectangle is repainted
    lat = (np.random.default_rng(3).standard_normal((1, 4, 40, 40)) * 0.18).astype(np.float32)
    return rgb, mask, lat


def write_inputs(d):
    rgb, mask, lat = inputs()
    h, w = mask.shape
    with open(os.path.join(d, "in.ppm"), "wb") as f:
        f.write(b"P6\n%d %d\n255\n" % (w, h)); f.write(rgb.tobytes())
    with open(os.path.join(d, "mask.pgm"), "wb") as f:
        f.write(b"P5\n%d %d\n255\n" % (w, h)); f.write(mask.tobytes())
    with open(os.path.join(d, "lat.tensor"), "wb") as f:
        f.write(b"TENSOR F32 40 40 4 1\n"); f.write(lat.tobytes())


def load(name):
    p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "e2e", name + ".npz")
    z = np.load(p)
    return z["latent"], z["image"]
