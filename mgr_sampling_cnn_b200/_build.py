"""Builds libnrrt.so (hand-written sm_100a CUDA + the C ABI of include/nrrt.h) in-tree with nvcc."""
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(_HERE, "libnrrt.so")

# per-file extra flags: the planner/sampler arithmetic is bit-exact fp64/fp32 (no FMA contraction)
SOURCES = {
    "abi.cu": [],
    "planner.cu": ["--fmad=false"],
    "sampler.cu": ["--fmad=false"],
    "conv_tc.cu": [],
    "conv_aux.cu": [],
    "forward.cu": [],
    "generate.cu": [],
    "gridpath.cu": [],
    "masks.cu": [],
    "wgrad_tc.cu": [],
    "train_ops.cu": [],
}
COMMON = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-Xcompiler", "-fPIC",
          "-I", os.path.join(_ROOT, "include"), "-I", CSRC]


def _newer(src, dst):
    return not os.path.exists(dst) or os.path.getmtime(src) > os.path.getmtime(dst)


def build(force: bool = False, verbose: bool = False) -> str:
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objdir = os.path.join(_HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(_ROOT, "include", "nrrt.h"))
    # experiment builds (e.g. NRRT_EXTRA_NVCC_FLAGS=-DNRRT_DEBUG_SWITCHES) never mix with product objects: a change of the
    # flag set rebuilds everything
    extra_all = os.environ.get("NRRT_EXTRA_NVCC_FLAGS", "").split()
    stamp = os.path.join(objdir, "flags.txt")
    old_flags = open(stamp).read() if os.path.exists(stamp) else ""
    if old_flags != " ".join(extra_all):
        force = True
        with open(stamp, "w") as f:
            f.write(" ".join(extra_all))
    objs, relink = [], force or not os.path.exists(LIB_PATH)
    for name, extra in SOURCES.items():
        src = os.path.join(CSRC, name)
        obj = os.path.join(objdir, name.replace(".cu", ".o"))
        if force or _newer(src, obj) or any(_newer(h, obj) for h in headers):
            cmd = [nvcc] + COMMON + extra + extra_all + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
            subprocess.check_call(cmd)
            relink = True
        objs.append(obj)
    if relink:
        subprocess.check_call([nvcc, "-shared", "-o", LIB_PATH] + objs + ["-lcudart", "-lcuda"])
    return LIB_PATH


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
