"""Whole-model anchor for the VAE: the reference's own vae.c graph running on the CPU oracle
(oracle/_ref/mlimgsynth_cpu, CLI `vae-decode` / `vae-encode`) against a PyTorch fp32 evaluation of the CompVis
latent-diffusion autoencoder (ldm.modules.diffusionmodules.model Encoder / Decoder: the architecture the
`first_stage_model.*` checkpoint keys belong to) on the same random-init weights. Written with torch.nn.functional
only; it shares no code with oracle/ggml_ref.c or with the reference. Operands of convolutions are rounded to f16 as
in ggml's CPU path; what is left is accumulation order (images agree to < 1/255 mean, latents to 1e-2 of the range).

CPU only; runs the 19 MB decoder + 34 MB encoder part of tools/gen_weights.py's VAE at 16x16 latents."""
import json, os, struct, subprocess, sys
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def read_safetensors(path, prefix):
    import torch
    out = {}
    with open(path, "rb") as f:
        n = struct.unpack("<Q", f.read(8))[0]
        hdr = json.loads(f.read(n)); base = 8 + n
        for k, v in hdr.items():
            if k.startswith(prefix):
                f.seek(base + v["data_offsets"][0])
                raw = f.read(v["data_offsets"][1] - v["data_offsets"][0])
                out[k[len(prefix):]] = torch.from_numpy(np.frombuffer(raw, dtype=np.float16).reshape(v["shape"]).astype(np.float32))
    return out


class LdmVae:
    """first_stage_model of SD1.x / SD2.x: ch 128, ch_mult (1, 2, 4, 4), 2 res blocks per level, attention in the middle."""
    def __init__(self, W):
        import torch, torch.nn.functional as F
        self.W, self.t, self.F = W, torch, F

    def conv(self, x, name, stride=1, pad=1):
        w = self.W[name + ".weight"]
        return self.F.conv2d(x.half().float(), w, self.W[name + ".bias"], stride=stride, padding=pad if w.shape[-1] == 3 else 0)

    def gn(self, x, name):
        return self.F.group_norm(x, 32, self.W[name + ".weight"], self.W[name + ".bias"], eps=1e-6)

    def resnet(self, x, p):
        h = self.conv(self.F.silu(self.gn(x, p + "norm1")), p + "conv1")
        h = self.conv(self.F.silu(self.gn(h, p + "norm2")), p + "conv2")
        if (p + "nin_shortcut.weight") in self.W:
            x = self.conv(x, p + "nin_shortcut")
        return x + h

    def attn(self, x, p):
        t = self.t
        n, c, hh, ww = x.shape
        h = self.gn(x, p + "norm")
        q, k, v = (self.conv(h, p + s).reshape(n, c, hh * ww) for s in ("q", "k", "v"))
        w = t.softmax(q.transpose(1, 2) @ k * c ** -0.5, dim=-1)            # [n, hw_q, hw_k]
        o = (v @ w.transpose(1, 2)).reshape(n, c, hh, ww)
        return x + self.conv(o, p + "proj_out")

    def decode(self, z):
        h = self.conv(z, "post_quant_conv")
        p = "decoder."
        h = self.conv(h, p + "conv_in")
        h = self.resnet(h, p + "mid.block_1."); h = self.attn(h, p + "mid.attn_1."); h = self.resnet(h, p + "mid.block_2.")
        for i in (3, 2, 1, 0):
            for j in range(3):
                h = self.resnet(h, p + "up.%d.block.%d." % (i, j))
            if i:
                h = self.conv(self.F.interpolate(h, scale_factor=2, mode="nearest"), p + "up.%d.upsample.conv" % i)
        return self.conv(self.F.silu(self.gn(h, p + "norm_out")), p + "conv_out")

    def encode(self, x):
        p = "encoder."
        h = self.conv(x, p + "conv_in")
        for i in range(4):
            for j in range(2):
                h = self.resnet(h, p + "down.%d.block.%d." % (i, j))
            if i != 3:
                h = self.conv(self.F.pad(h, (0, 1, 0, 1)), p + "down.%d.downsample.conv" % i, stride=2, pad=0)   # asymmetric padding
        h = self.resnet(h, p + "mid.block_1."); h = self.attn(h, p + "mid.attn_1."); h = self.resnet(h, p + "mid.block_2.")
        h = self.conv(self.F.silu(self.gn(h, p + "norm_out")), p + "conv_out")
        return self.conv(h, "quant_conv")                                    # moments: mean | logvar


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
        return np.frombuffer(f.read(), dtype=np.uint8).reshape(h, w, -1)


@pytest.fixture(scope="module")
def vae_setup(oracle_built, tmp_path_factory):
    pytest.importorskip("torch")
    import gen_weights
    exe = os.path.join(oracle_built, "mlimgsynth_cpu")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/mlimgsynth_cpu not built")
    d = tmp_path_factory.mktemp("vae")
    wpath = str(d / "vae.safetensors")
    gen_weights.write_safetensors(wpath, gen_weights.build_spec("sd1", parts=("vae",)), 1234, "f16")
    return exe, wpath, d, LdmVae(read_safetensors(wpath, "first_stage_model."))


def test_reference_vae_decode_on_oracle_matches_ldm_decoder(vae_setup):
    import torch
    exe, wpath, d, vae = vae_setup
    lat = (np.random.default_rng(5).standard_normal((1, 4, 16, 16)) * 0.18215 * 4).astype(np.float32)   # SD-scaled latent
    with open(d / "lat.tensor", "wb") as f:
        f.write(b"TENSOR F32 16 16 4 1\n"); f.write(lat.tobytes())
    r = subprocess.run([exe, "vae-decode", "-m", wpath, "--model-type", "sd1", "--ilatent", str(d / "lat.tensor"), "-o", str(d / "out.pnm")],
                       capture_output=True, text=True, timeout=600)       # (a VAE-only file has no UNet keys to detect the model type from)
    assert r.returncode == 0, r.stderr[-1500:]
    got = load_pnm(str(d / "out.pnm")).astype(np.float32)                   # [H, W, 3] u8
    with torch.no_grad():
        y = vae.decode(torch.from_numpy(lat) / 0.18215)[0]                  # vae.c:31 scale factor
    want = ((y + 1) / 2).clamp(0, 1).permute(1, 2, 0).numpy() * 255.0       # vae.h:43-47, mlimgsynth.c:112-129
    assert got.shape == want.shape == (128, 128, 3)
    diff = np.abs(got - np.floor(want))
    print("VAE decode on the oracle vs LDM decoder in torch: mean |d| %.3f / 255, max %.0f" % (diff.mean(), diff.max()))
    assert diff.mean() <= 0.6 and diff.max() <= 6


def test_reference_vae_encode_on_oracle_matches_ldm_encoder(vae_setup):
    import torch
    exe, wpath, d, vae = vae_setup
    rgb = np.random.default_rng(9).integers(0, 256, size=(128, 128, 3), dtype=np.uint8)
    with open(d / "in.ppm", "wb") as f:
        f.write(b"P6\n128 128\n255\n"); f.write(rgb.tobytes())
    r = subprocess.run([exe, "vae-encode", "-m", wpath, "--model-type", "sd1", "-i", str(d / "in.ppm"), "--olatent", str(d / "enc.tensor"), "-S", "1"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-1500:]
    got = load_tensor(str(d / "enc.tensor"))                                 # [1, 4, 16, 16] sampled, scaled latent
    x = torch.from_numpy(rgb.astype(np.float32) / 255.0).permute(2, 0, 1)[None] * 2 - 1     # vae.h:36-41
    with torch.no_grad():
        mom = vae.encode(x)[0]
    mean, logvar = mom[:4].numpy(), mom[4:].clamp(-30, 20).numpy()
    # latent = (mean + exp(logvar / 2) * noise) * 0.18215 (vae.c:197-220), noise = the first Philox draw of seed 1 (bit-exact
    # generator of the host layer, tests/test_host_cpu.py)
    import ctypes as C
    import mlimgsynth_b200

    class Rng(C.Structure):
        _fields_ = [("seed", C.c_uint64), ("offset", C.c_uint32)]
    L = C.CDLL(mlimgsynth_b200.HOST_LIB)
    buf = (C.c_float * 1024)()
    L.rng_philox_randn(C.byref(Rng(1, 0)), 1024, buf)
    noise = np.frombuffer(buf, dtype=np.float32).reshape(4, 16, 16)
    want = (mean + np.exp(0.5 * logvar) * noise) * 0.18215
    err = np.abs(got[0] - want).max() / np.abs(want).max()
    print("VAE encode on the oracle vs LDM encoder in torch + Philox sample: max-rel err %.2e" % err)
    assert err <= 1e-2
