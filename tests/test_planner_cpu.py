"""Planner logic without a GPU: graphs of tests/blocks.py are planned in dry-run mode (GGML_B200_DRYRUN=1: host pointers,
nothing executes) and the fused step list is inspected through ggml_b200_last_plan_count. Checks that the fusions the
design relies on actually happen (and that their switches undo them). Runs in a subprocess because the dry-run switch
is read once per process."""
import json, os, subprocess, sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r'''
import ctypes, json, os, sys
sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, "tests"))
import mlimgsynth_b200, blocks
eng = mlimgsynth_b200.load_engine(); eng.init_backend()
lib = ctypes.CDLL(mlimgsynth_b200.ENGINE_LIB)
lib.ggml_b200_last_plan_count.argtypes = [ctypes.c_char_p]
def plan(build):
    G = blocks.Graph(eng); o = build(blocks.B(G, 0)); G.run(o)
    keys = ["steps", "GEMM_TC", "CONV_TC", "ATTENTION", "GROUPNORM", "LAYERNORM", "GEGLU", "COPY", "UNARY", "BINARY", "GEMM_SIMT",
            "name:linear_fused", "name:linear_geglu", "name:concat_a", "gn_from_epilogue"]
    r = {k: lib.ggml_b200_last_plan_count(k.encode()) for k in keys}
    G.free()
    return r
def three_resnets(b):
    emb = b.inp(3, 1280)
    h = b.conv2d(b.inp(3, 4, 16, 16), 128)
    h = b.resnet(h, emb, 128); h = b.resnet(h, emb, 256)
    return b.resnet(h, emb, 192)
def transformer(b):
    x = b.conv2d(b.inp(1, 4, 16, 16), 320)          # the residual input is an engine tensor (channels-last f16), as in the UNet
    return b.spatial_transf(x, b.inp(77, 768), 320, 8)
def resnet_only(b):
    return b.resnet(b.conv2d(b.inp(2, 4, 16, 16), 128), b.inp(2, 1280), 128)
out = {"resnets": plan(three_resnets), "transformer": plan(transformer), "resnet": plan(resnet_only),
       "ff": plan(lambda b: b.feed_forward(b.inp(256, 320), 320))}
print("RESULT " + json.dumps(out))
'''


def run(env_extra=None):
    env = dict(os.environ, GGML_B200_DRYRUN="1", GGML_B200_QUIET="1")
    env.update(env_extra or {})
    r = subprocess.run([sys.executable, "-c", SCRIPT % {"root": ROOT}], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")][-1]
    return json.loads(line[7:])


@pytest.fixture(scope="module")
def fused():
    return run()


def test_resnet_is_five_launch_classes(fused):
    """GroupNorm+SiLU -> conv(+bias+emb) -> GroupNorm+SiLU -> conv(+bias+residual): no stand-alone add / activation steps."""
    r = fused["resnet"]
    assert r["BINARY"] == 0 and r["GROUPNORM"] == 2
    assert r["CONV_TC"] == 2 and r["UNARY"] == 1          # the one unary is silu(emb)


def test_embedding_projections_fuse_into_one_gemm(fused):
    r = fused["resnets"]
    assert r["name:linear_fused"] == 1                     # three biased emb projections -> one GEMM
    assert r["UNARY"] == 1                                 # silu(emb) computed once (CSE)
    unf = run({"GGML_B200_NO_EMB_FUSION": "1"})["resnets"]
    assert unf["name:linear_fused"] == 0 and unf["GEMM_TC"] == r["GEMM_TC"] + 2


def test_transformer_block_fusions(fused):
    r = fused["transformer"]
    assert r["ATTENTION"] == 2 and r["LAYERNORM"] == 3 and r["GROUPNORM"] == 1
    assert r["name:linear_fused"] == 2                     # q/k/v of the self-attention, k/v of the cross-attention
    assert r["name:linear_geglu"] == 1 and r["GEGLU"] == 0  # gate inside the projection's epilogue
    assert r["BINARY"] == 0 and r["GEMM_SIMT"] == 0
    unf = run({"GGML_B200_NO_PROJ_FUSION": "1", "GGML_B200_NO_GEGLU_FUSION": "1"})["transformer"]
    assert unf["name:linear_fused"] == 0 and unf["GEGLU"] == 1 and unf["GEMM_TC"] == r["GEMM_TC"] + 3   # q,k,v and k,v apart


def test_feed_forward_is_two_gemms(fused):
    r = fused["ff"]
    assert r["GEMM_TC"] == 2 and r["name:linear_geglu"] == 1 and r["BINARY"] == 0


def test_groupnorm_statistics_come_from_the_producer(fused):
    """Every group_norm whose input was written by a tensor-core GEMM / conv asks that launch for its statistics
    (mlblock_nn.c:78): conv_in -> GN1, conv1 -> GN2 and resnet -> next GN1; the switch undoes it."""
    assert fused["resnet"]["gn_from_epilogue"] == 2
    assert fused["resnets"]["gn_from_epilogue"] == 6
    assert fused["transformer"]["gn_from_epilogue"] == 1
    plain = run({"GGML_B200_NO_GN_EPILOGUE": "1"})
    assert plain["resnets"]["gn_from_epilogue"] == 0 and plain["resnets"]["GROUPNORM"] == fused["resnets"]["GROUPNORM"]
