"""Build-quality guards that need no GPU: the register-heavy tcgen05 kernels must not spill to local memory.
A run-time switch in the hot loop of attn_ap_kernel once made ptxas spill the score registers (STACK 288, 336 LDL/STL): every
parity test still passed and the kernel ran at half speed (profiles/r2_attention_microbench.md). cuobjdump -res-usage shows it."""
import os, re, shutil, subprocess
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "mlimgsynth_b200", "build")


def res_usage(obj):
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe) or not os.path.exists(obj):
        pytest.skip("cuobjdump or %s not available" % os.path.basename(obj))
    out = subprocess.run([exe, "-res-usage", obj], capture_output=True, text=True, timeout=300).stdout
    res = {}
    for m in re.finditer(r"Function (\S+):\s*\n\s*(.*)", out):
        fields = dict(kv.split(":") for kv in m.group(2).split() if ":" in kv)
        res[m.group(1)] = {k: int(v) for k, v in fields.items() if v.isdigit()}
    return res


def test_attention_kernels_keep_their_scores_in_registers():
    res = res_usage(os.path.join(OBJ, "attn_tc.cu.o"))
    ap = {k: v for k, v in res.items() if "attn_ap_kernel" in k}
    assert ap, "attn_ap_kernel not found in the object"
    for k, v in ap.items():
        assert v["STACK"] <= 32, "%s spills: STACK %d" % (k, v["STACK"])          # 32 = the argument block of the time-out printf
        assert v["REG"] <= 184, "%s: %d registers do not fit 352 threads per SM" % (k, v["REG"])
    for k, v in res.items():
        if "attn_split_kernel" in k or "attn_tc_kernel" in k or "attn_kv1_kernel" in k:
            assert v["STACK"] <= 104, "%s spills: STACK %d" % (k, v["STACK"])      # trace + time-out argument blocks only


def test_gemm_kernels_do_not_spill():
    res = res_usage(os.path.join(OBJ, "gemm_tc.cu.o"))
    pk = {k: v for k, v in res.items() if "gemm_tc_persistent_kernel" in k}
    assert len(pk) >= 26, "expected the 14 plain and 12 GroupNorm-statistics variants, found %d" % len(pk)
    for k, v in pk.items():
        assert v["STACK"] <= 16 and v["REG"] <= 200, "%s: STACK %d REG %d" % (k, v["STACK"], v["REG"])
