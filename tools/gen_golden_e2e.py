#!/usr/bin/env python3
"""Generates tests/golden/e2e/*.npz: final latents and images of the REFERENCE's own host code running
on the CPU oracle (oracle/_ref/mlimgsynth_cpu, built from /root/reference) for the end-to-end parity
cases of tests/test_e2e_gpu.py and tests/test_host_gpu.py. Weights are the deterministic random-init
checkpoint of tools/gen_weights.py (seed 1234), so the GPU box regenerates identical inputs.
Run in the build container; the .npz files (a few KB each) are committed."""
import os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import gen_weights
from test_e2e_gpu import load_tensor, load_pnm
import golden_cases as G

out_dir = os.path.join(ROOT, "tests", "golden", "e2e"); os.makedirs(out_dir, exist_ok=True)
tmp = "/tmp/mlis_golden"; os.makedirs(tmp, exist_ok=True)
def weights(kind):
    p = os.path.join(tmp, kind + ".safetensors")
    if not os.path.exists(p):
        gen_weights.write_safetensors(p, gen_weights.build_spec(kind), 1234, "f16")
    return p
lora = os.path.join(tmp, "lora1.safetensors")
gen_weights.write_lora(lora, "sd1", rank=8, alpha=8.0, seed=5)
G.write_inputs(tmp)
exe = os.path.join(ROOT, "oracle", "_ref", "mlimgsynth_cpu")
for name, case in G.CASES.items():
    dst = os.path.join(out_dir, name + ".npz")
    if os.path.exists(dst) and "--force" not in sys.argv:
        continue
    o = os.path.join(tmp, name)
    model = weights(case.get("model", "sd1"))
    tae = weights("tae") if case.get("tae") else ""
    cli = [a.replace("@TMP@", tmp).replace("@LORA@", lora).replace("@TAE@", tae) for a in case["cli"]]
    if case.get("cmd") == "vae-decode":
        cmd = [exe, "vae-decode", "-m", model, "-o", o + ".pnm"] + cli
    else:
        cmd = [exe, "generate", "-m", model, "-p", case.get("prompt", G.PROMPT), "-S", "42", "-o", o + ".pnm", "--olatent", o + ".tensor"] + cli
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    img = (load_pnm(o + ".pnm") * 255 + 0.5).astype(np.uint8)
    lat = load_tensor(o + ".tensor") if case.get("cmd") != "vae-decode" else np.zeros(1, np.float32)
    np.savez_compressed(dst, latent=lat, image=img)
    print(name, lat.shape, img.shape)
