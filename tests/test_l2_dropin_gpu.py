"""Level-2 drop-in proof: the reference's own CONSUMERS of include/mlimgsynth.h run unmodified on the product host library
(mlimgsynth_b200/lib/libmlimgsynth_b200.so):
  * the reference CLI (main_mlimgsynth.c objects compiled from /root/reference, linked by oracle/Makefile target
    `mlimgsynth_l2`), and
  * the reference's Python binding python/mlimgsynth.py (staged by the same Makefile into oracle/_ref/python, loaded with
    MLIS_LIB_PATH as its line 166 documents).
Results are compared with the fixtures of the reference's own library on the CPU oracle (tests/golden/e2e)."""
import importlib.util, os, subprocess, sys
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import golden_cases as G   # noqa: E402
from test_e2e_gpu import load_tensor, load_pnm, check   # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref")


def _binding():
    p = os.path.join(REF, "python", "mlimgsynth.py")
    if not os.path.exists(p):
        pytest.skip("reference python binding not staged (make -C oracle b200)")
    import mlimgsynth_b200
    os.environ["MLIS_LIB_PATH"] = mlimgsynth_b200.HOST_LIB
    spec = importlib.util.spec_from_file_location("ref_mlimgsynth_binding", p)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_reference_python_binding_loads_and_sets_options():
    """No GPU needed: the binding resolves every symbol it declares and the option calls of its self-test work."""
    m = _binding()
    s = m.MLImgSynth()
    s.option_set(m.MLIS_OPT_IMAGE_DIM, 512, 512)        # by id, variadic ints
    s.option_set("cfg-scale", 7.0)                      # by name
    s.option_set("method", "dpm++2m")
    with pytest.raises(RuntimeError):
        s.option_set("no-such-option", 1)
    assert "no-such-option" in s.errstr_get()


@pytest.mark.gpu
def test_reference_python_binding_generates_on_engine():
    import bench
    m = _binding()
    lat_c, img_c = G.load("euler")
    s = m.MLImgSynth()
    s.option_set("model", bench.weights_path("sd1"))
    s.option_set(m.MLIS_OPT_IMAGE_DIM, 128, 128)
    s.option_set("steps", 3); s.option_set("method", "euler"); s.option_set("cfg-scale", 7.0)
    s.option_set("seed", 42, 0)
    s.option_set("prompt", G.PROMPT)
    s.setup()
    s.generate()
    img = s.image_get()
    assert (img.w, img.h, img.c) == (128, 128, 3)
    a = np.frombuffer(img.data, dtype=np.uint8).reshape(128, 128, 3)
    mse = float(((a.astype(np.float32) - img_c.astype(np.float32)) ** 2).mean()) / 255.0 ** 2
    psnr = 10 * np.log10(1.0 / mse) if mse > 0 else 99.0
    print("reference python binding on the engine: PSNR %.1f dB; infotext: %s" % (psnr, s.infotext_get().splitlines()[-1][:80]))
    assert psnr >= 35.0
    lat = s.clip_text_encode("a blue cat")            # TMP tensors + mlis_clip_text_encode through the binding
    assert lat.n[0] == 768 and lat.n[1] == 77


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["ref_euler_cfg", "ref_dpmpp2m_karras"])
def test_reference_cli_on_host_library(tmp_path, name):
    import bench
    exe = os.path.join(REF, "mlimgsynth_l2")
    if not os.path.exists(exe):
        pytest.skip("reference CLI not linked against the host library (make -C oracle b200)")
    case = G.CASES[name]
    lat_c, img_c = G.load(name)
    o = str(tmp_path / name)
    cmd = [exe, "generate", "-m", bench.weights_path("sd1"), "-p", case["prompt"], "-S", "42", "-o", o + ".pnm", "--olatent", o + ".tensor"] + case["cli"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    check(load_tensor(o + ".tensor"), load_pnm(o + ".pnm"), lat_c, img_c.astype(np.float32) / 255.0, "reference CLI on libmlimgsynth_b200: " + name)
