"""Opt-in cross-GPU CFG split (reference: the serial cond / uncond pair of mlimgsynth.c:1578-1583): two ranks run the
same generation, each evaluates ONE CFG half per UNet evaluation and exchanges it with its peer. Both must end with the same
latent, and that latent must match the reference fixture of the ordinary (batched-halves) path."""
import os, subprocess, sys
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import golden_cases as G   # noqa: E402


def test_cfg_halves_on_two_ranks(tmp_path):
    import torch
    backend = "nccl" if torch.cuda.device_count() >= 2 else "gloo"      # one GPU: both ranks share it, host-staged exchange
    out = str(tmp_path / "cfg.npz")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29641", os.path.join(ROOT, "tools", "cfg_split_worker.py"), "--out", out, "--backend", backend]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-3000:]
    z = np.load(out)
    lat_c, img_c = G.load("euler")
    assert np.array_equal(z["lat0"], z["lat1"]) and np.array_equal(z["img0"], z["img1"])      # identical sampler states on both ranks
    err = np.abs(z["lat0"] - lat_c).max() / np.abs(lat_c).max()
    mse = float(((z["img0"].astype(np.float32) - img_c.astype(np.float32)) ** 2).mean()) / 255.0 ** 2
    psnr = 10 * np.log10(1.0 / mse) if mse > 0 else 99.0
    print("CFG split over 2 ranks (%s): latent max-rel err %.3e, PSNR %.1f dB" % (backend, err, psnr))
    assert err <= 1e-2 and psnr >= 35.0
