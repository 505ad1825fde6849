"""TEST INFRASTRUCTURE ONLY.  Recipe that compiles the reference's OWN geometry sources, unmodified and in place from
/root/reference/src/common, against oracle/shim/ into oracle/_ref/libtod_ref.so.

    python -m oracle.build_ref

The reference's build system (catkin + ecto + OpenCV + Boost) cannot run here; these two .cpp files plus four headers
need only 12 cv:: symbols and 3 Boost utilities, which oracle/shim provides.  Reference sources are never copied into
the repo; oracle/_ref/ is git-ignored and travels to the GPU box as a built artefact.  Where /root/reference is
absent (GPU box) the prebuilt .so is used as is.
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/src/common"
OUT_DIR = os.path.join(HERE, "_ref")
OUT = os.path.join(OUT_DIR, "libtod_ref.so")


def build(force=False):
    if not os.path.isdir(REF):
        return OUT if os.path.exists(OUT) else None
    os.makedirs(OUT_DIR, exist_ok=True)
    srcs = [os.path.join(REF, "adjacency_ransac.cpp"), os.path.join(REF, "maximum_clique.cpp"),
            os.path.join(HERE, "ref_harness.cpp")]
    deps = srcs + [os.path.join(HERE, "shim", "opencv2", "core", "core.hpp")]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in deps):
        return OUT
    # same flags as the reference's CMakeLists.txt:11 (-Wno-pragmas -fno-strict-aliasing -Wall) + -O2;
    # x86-64 baseline => no FMA contraction, like the reference's stock build
    cmd = ["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-fno-strict-aliasing", "-Wno-pragmas", "-ffp-contract=off",
           "-I", os.path.join(HERE, "shim"), "-I", REF, "-o", OUT] + srcs
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force=True))
