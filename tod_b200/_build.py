"""Build recipe for libtod_b200.so (nvcc, sm_100a only).  Used by __graft_entry__.build() and `python -m tod_b200._build`.

The library is built IN-TREE (tod_b200/libtod_b200.so) so that it travels to the GPU box with the repo snapshot.
nvcc cross-compiles without a GPU.
"""
import concurrent.futures
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
# TOD_B200_VARIANT=<name> + TOD_B200_DEFINES="-DX=1 ..." build an instrumented / experimental copy next to the
# production library (libtod_b200_<name>.so, loaded when TOD_B200_LIB points at it); the default build has neither.
VARIANT = os.environ.get("TOD_B200_VARIANT", "")
DEFINES = os.environ.get("TOD_B200_DEFINES", "").split() if VARIANT else []
OBJ = os.path.join(PKG, "build" + ("_" + VARIANT if VARIANT else ""))
LIB = os.path.join(PKG, "libtod_b200%s.so" % ("_" + VARIANT if VARIANT else ""))

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-I", os.path.join(ROOT, "include"), "-I", CSRC,
          "-Xcompiler", "-fPIC,-ffp-contract=off,-Wall"]
# per-file extra flags: the adjacency kernel must not contract mul+add into FMA (bit-exact vs the reference's x86 code)
EXTRA = {"k2_adjacency.cu": ["--fmad=false"], "k3_score.cu": ["--fmad=false"], "orb.cu": ["--fmad=false"]}


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith((".cu", ".cpp")))


def _headers_mtime():
    m = 0.0
    for d in (CSRC, os.path.join(ROOT, "include")):
        for f in os.listdir(d):
            if f.endswith((".h", ".cuh")):
                m = max(m, os.path.getmtime(os.path.join(d, f)))
    return m


def _compile(src, verbose):
    obj = os.path.join(OBJ, src + ".o")
    cmd = [NVCC] + ARCH + COMMON + DEFINES + EXTRA.get(src, []) + ["-c", os.path.join(CSRC, src), "-o", obj]
    if src.endswith(".cu"):
        cmd += ["-Xptxas", "-v"] if verbose else []
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
    return obj, r.stderr


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    hm = _headers_mtime()
    todo, objs = [], []
    for s in sources():
        obj = os.path.join(OBJ, s + ".o")
        objs.append(obj)
        sm = max(os.path.getmtime(os.path.join(CSRC, s)), hm, os.path.getmtime(__file__))
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < sm:
            todo.append(s)
    if todo:
        with concurrent.futures.ThreadPoolExecutor(max_workers=8) as ex:
            for obj, log in ex.map(lambda s: _compile(s, verbose), todo):
                if verbose and log:
                    sys.stderr.write(log)
    if todo or not os.path.exists(LIB):
        link_extra = os.environ.get("TOD_B200_LINK_FLAGS", "").split() if VARIANT else []
        cmd = [NVCC] + ARCH + ["-shared", "-o", LIB] + objs + ["-Xlinker", "--exclude-libs,ALL"] + link_extra
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
