"""The reference arm of bench.py (`--impl reference`) runs entirely on the host: the reference's own CPU implementation of
the path (its host code on the CPU oracle). Check here, without a GPU, that it produces the contract's JSON line."""
import json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line(oracle_built):
    exe = os.path.join(oracle_built, "mlimgsynth_cpu")
    if not os.path.exists(exe):
        import pytest
        pytest.skip("oracle/_ref/mlimgsynth_cpu not built")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "images_per_sec" and line["unit"] == "images/s"
    assert line["higher_is_better"] is True and line["n_gpus"] == 1 and line["steps"] == 1
    assert 0 < line["value"] < 1.0                     # a CPU takes seconds per UNet evaluation
    assert line["config"]["workload"].startswith("SD1.5 txt2img 512x512")
    cb = line["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["cores"] == os.cpu_count() and cb["value"] == line["value"] and "UNet evaluation" in cb["sample"]
    e2e = line["e2e"]
    assert e2e["value"] == line["value"] and e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0
