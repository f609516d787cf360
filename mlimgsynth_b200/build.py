"""Build the B200 engine: mlimgsynth_b200/lib/libggml_b200.so (sm_100a only, in-tree).

nvcc cross-compiles without a GPU. The library is linked with the static CUDA runtime so it can be
loaded next to any other CUDA user (e.g. torch) without sharing a libcudart.
"""
import os, subprocess, sys, hashlib, glob

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "lib")
OBJ = os.path.join(HERE, "build")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O3", "-lineinfo",
         "-Xcompiler", "-fPIC,-Wall,-Wno-unused-function", "-I", os.path.join(ROOT, "include"), "-I", CSRC]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(verbose=False, force=False):
    os.makedirs(LIB, exist_ok=True)
    os.makedirs(OBJ, exist_ok=True)
    headers = glob.glob(os.path.join(CSRC, "*.h")) + glob.glob(os.path.join(ROOT, "include", "*.h"))
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cpp")) + glob.glob(os.path.join(CSRC, "*.cu")))
    objs = []
    procs = []
    for s in srcs:
        o = os.path.join(OBJ, os.path.basename(s) + ".o")
        objs.append(o)
        if force or _stale(o, [s] + headers):
            cmd = [NVCC] + FLAGS + ["-x", "cu", "-c", s, "-o", o]
            if verbose:
                print(" ".join(cmd))
            procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    for s, p in procs:
        out = p.communicate()[0].decode()
        if p.returncode:
            sys.stderr.write(out)
            raise RuntimeError("nvcc failed on %s" % s)
        if verbose and out.strip():
            print(out)
    so = os.path.join(LIB, "libggml_b200.so")
    if force or procs or _stale(so, objs):
        cmd = [NVCC, "-shared", "-o", so] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-Xlinker", "--no-undefined", "-Xlinker", "-Bsymbolic"]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    build_host(so, verbose, force)
    return so


def build_host(engine_so, verbose=False, force=False):
    """The C host layer (mlis_* API, model builders, sampler): libmlimgsynth_b200.so, plain gcc."""
    hdir = os.path.join(CSRC, "host")
    srcs = sorted(glob.glob(os.path.join(hdir, "*.c")))
    deps = srcs + glob.glob(os.path.join(hdir, "*.h")) + glob.glob(os.path.join(ROOT, "include", "*.h")) + [engine_so]
    out = os.path.join(LIB, "libmlimgsynth_b200.so")
    if not (force or _stale(out, deps)):
        return out
    cmd = ["gcc", "-std=gnu11", "-O2", "-g", "-fPIC", "-shared", "-Wall", "-Wno-unused-function", "-fvisibility=default",
           "-I", os.path.join(ROOT, "include"), "-I", hdir, "-o", out] + srcs + \
          ["-L", LIB, "-lggml_b200", "-Wl,-rpath,$ORIGIN", "-Wl,-soname,libmlimgsynth_b200.so", "-Wl,--no-undefined", "-lm", "-ldl"]
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return out


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv))
