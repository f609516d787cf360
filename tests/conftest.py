import os, subprocess, sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

ORACLE_DIR = os.path.join(ROOT, "oracle")
REF_DIR = os.path.join(ORACLE_DIR, "_ref")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def oracle_built():
    """The CPU oracle (oracle/_ref). Built here when the reference tree is present (this container);
    on the GPU box the prebuilt files travel with the snapshot."""
    if os.path.isdir("/root/reference/src"):
        subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "cpu"])
    need = os.path.join(REF_DIR, "libggml_ref.so")
    if not os.path.exists(need):
        pytest.skip("oracle/_ref not built and /root/reference absent")
    return REF_DIR


@pytest.fixture(scope="session")
def ref(oracle_built):
    from mlimgsynth_b200.ggml import GGML
    return GGML(os.path.join(oracle_built, "libggml_ref.so"))


@pytest.fixture(scope="session")
def eng():
    import mlimgsynth_b200
    g = mlimgsynth_b200.load_engine()   # raises loudly if the CUDA library is missing
    g.init_backend()
    return g
