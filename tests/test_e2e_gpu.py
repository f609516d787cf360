"""Drop-in proof: the reference's own unmodified host code (oracle/_ref/mlimgsynth_*: CLI + libmlimgsynth
objects compiled from /root/reference against include/ggml*.h) runs txt2img once on the CPU oracle
and once on the CUDA engine, same random-init weights, prompt and seed.

Bars (BASELINE.json north_star): final latent within max-relative error 1e-2 per UNet step
(checked on the sampled latent after the last step), decoded image PSNR >= 35 dB.
"""
import os, subprocess, sys
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def load_tensor(path):
    with open(path, "rb") as f:
        head = f.readline().split()
        ne = [int(x) for x in head[2:6]]
        return np.frombuffer(f.read(), dtype=np.float32).reshape(ne[::-1])


def load_pnm(path):
    with open(path, "rb") as f:
        toks = []
        while len(toks) < 4:
            toks += f.readline().split()
        w, h = int(toks[1]), int(toks[2])
        return np.frombuffer(f.read(), dtype=np.uint8).reshape(h, w, -1).astype(np.float32) / 255.0


@pytest.fixture(scope="module")
def sd1_weights(tmp_path_factory):
    import gen_weights
    d = tmp_path_factory.mktemp("w")
    p = str(d / "sd1.safetensors")
    gen_weights.write_safetensors(p, gen_weights.build_spec("sd1"), 1234, "f16")
    return p


def run_cli(binary, model, out_prefix, extra):
    cmd = [os.path.join(ROOT, "oracle", "_ref", binary), "generate", "-m", model,
           "-p", "a photograph of an astronaut riding a horse", "-S", "42",
           "-o", out_prefix + ".pnm", "--olatent", out_prefix + ".tensor"] + extra
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=1200)
    assert r.returncode == 0, r.stderr[-2000:]
    return load_tensor(out_prefix + ".tensor"), load_pnm(out_prefix + ".pnm")


@pytest.mark.parametrize("name,extra", [
    ("euler_cfg", ["-d", "128,128", "-s", "3", "--method", "euler", "--cfg-scale", "7"]),
    ("dpmpp2m_karras", ["-d", "192,128", "-s", "4", "--method", "dpm++2m", "--scheduler", "karras", "--cfg-scale", "5"]),
    ("euler_a", ["-d", "128,128", "-s", "3", "--method", "euler_a", "--cfg-scale", "1"]),
])
def test_reference_host_on_engine(sd1_weights, tmp_path, name, extra):
    lat_c, img_c = run_cli("mlimgsynth_cpu", sd1_weights, str(tmp_path / ("cpu_" + name)), extra)
    lat_g, img_g = run_cli("mlimgsynth_b200", sd1_weights, str(tmp_path / ("gpu_" + name)), extra)
    assert lat_c.shape == lat_g.shape and np.isfinite(lat_g).all()
    err = np.abs(lat_g - lat_c).max() / np.abs(lat_c).max()
    mse = float(((img_g - img_c) ** 2).mean())
    psnr = 10 * np.log10(1.0 / mse) if mse > 0 else 99.0
    print("%s: latent max-rel err %.3e, image PSNR %.1f dB" % (name, err, psnr))
    assert err <= 1e-2
    assert psnr >= 35.0
