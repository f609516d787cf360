"""Drop-in proof: the reference's own unmodified host code (oracle/_ref/mlimgsynth_b200: CLI +
libmlimgsynth objects compiled from /root/reference against include/ggml*.h, linked to the product
libggml_b200.so) runs txt2img on the CUDA engine; results are compared with the same code running on
the CPU oracle (committed fixtures tests/golden/e2e/ref_*.npz from tools/gen_golden_e2e.py; one case is
also recomputed live on the CPU oracle). Same random-init weights (tools/gen_weights.py, seed 1234),
prompt and seed.

Bars (BASELINE.json north_star): final latent within max-relative error 1e-2, image PSNR >= 35 dB.
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


@pytest.fixture(scope="session")
def sd1_weights(tmp_path_factory):
    import bench
    return bench.weights_path("sd1")


def run_cli(binary, model, out_prefix, prompt, extra):
    cmd = [os.path.join(ROOT, "oracle", "_ref", binary), "generate", "-m", model, "-p", prompt, "-S", "42",
           "-o", out_prefix + ".pnm", "--olatent", out_prefix + ".tensor"] + extra
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=1200)
    assert r.returncode == 0, r.stderr[-2000:]
    return load_tensor(out_prefix + ".tensor"), load_pnm(out_prefix + ".pnm")


def check(lat_g, img_g, lat_c, img_c, name):
    assert lat_c.shape == lat_g.shape and np.isfinite(lat_g).all()
    err = np.abs(lat_g - lat_c).max() / np.abs(lat_c).max()
    mse = float(((img_g - img_c) ** 2).mean())
    psnr = 10 * np.log10(1.0 / mse) if mse > 0 else 99.0
    print("%s: latent max-rel err %.3e, image PSNR %.1f dB" % (name, err, psnr))
    assert err <= 1e-2
    assert psnr >= 35.0


@pytest.mark.parametrize("name", ["ref_euler_cfg", "ref_dpmpp2m_karras", "ref_euler_a"])
def test_reference_host_on_engine_vs_fixture(sd1_weights, tmp_path, name):
    import golden_cases as G
    case = G.CASES[name]
    lat_c, img_c = G.load(name)
    lat_g, img_g = run_cli("mlimgsynth_b200", sd1_weights, str(tmp_path / ("gpu_" + name)), case["prompt"], case["cli"])
    check(lat_g, img_g, lat_c, img_c.astype(np.float32) / 255.0, name)


def test_reference_host_on_engine_vs_live_oracle(sd1_weights, tmp_path):
    import golden_cases as G
    extra = ["-d", "128,128", "-s", "2", "--method", "heun", "--cfg-scale", "4"]
    lat_c, img_c = run_cli("mlimgsynth_cpu", sd1_weights, str(tmp_path / "cpu"), G.PROMPT_PLAIN, extra)
    lat_g, img_g = run_cli("mlimgsynth_b200", sd1_weights, str(tmp_path / "gpu"), G.PROMPT_PLAIN, extra)
    check(lat_g, img_g, lat_c, img_c, "live heun")
