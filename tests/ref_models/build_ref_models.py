#!/usr/bin/env python
"""Compile the REFERENCE's own model sources, unchanged and where they lie under /root/reference, against the B200 host API
(ntrtsim_b200/host/compat: the reference's include paths forwarded to tgb_host.h / tgb_creator.h) into
tests/ref_models/_build/libref_models.so.  TEST INFRASTRUCTURE: nothing of the reference is copied into the repo, the output is
git-ignored and travels to the GPU box like the other built libraries.  Linked -Bsymbolic: the host library defines classes of the
same names (its own builds of the same examples), and each side must keep calling its own."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/src"
OUT = os.path.join(HERE, "_build")
LIB = os.path.join(OUT, "libref_models.so")

SOURCES = [
    "examples/3_prism/PrismModel.cpp",
    "examples/SUPERball/T6Model.cpp",
    "examples/craterEscape/EscapeModel.cpp",
    "examples/motorModel/tsTestRig.cpp",
    "examples/learningSpines/BaseSpineModelLearning.cpp",
    "examples/learningSpines/TetraSpine/TetraSpineLearningModel.cpp",
    "examples/learningSpines/TetrahedralComplex/FlemonsSpineModelLearning.cpp",
    "models/obstacles/tgCraterDeep.cpp",
    "models/obstacles/tgCraterShallow.cpp",
    "models/obstacles/tgStairs.cpp",
]


def available():
    return all(os.path.exists(os.path.join(REF, s)) for s in SOURCES)


def build(force=False):
    """returns the library's path, or None when the reference tree is not here and no library was built earlier"""
    if not available():
        return LIB if os.path.exists(LIB) else None
    host = os.path.join(ROOT, "ntrtsim_b200", "host")
    libdir = os.path.join(ROOT, "ntrtsim_b200", "lib")
    deps = [os.path.join(REF, s) for s in SOURCES] + [os.path.join(HERE, "factory.cpp"), os.path.join(libdir, "libtgb_host.so")]
    if not force and os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(d) for d in deps):
        return LIB
    os.makedirs(OUT, exist_ok=True)
    inc = ["-I" + os.path.join(host, "compat"), "-I" + host, "-I" + os.path.join(ROOT, "include"), "-I" + REF]
    inc += ["-I" + d for d in sorted(set(os.path.dirname(os.path.join(REF, s)) for s in SOURCES))]
    cmd = ["g++", "-std=c++17", "-O1", "-fPIC", "-shared", "-w", "-o", LIB] + inc + [os.path.join(REF, s) for s in SOURCES] + \
          [os.path.join(HERE, "factory.cpp"), "-L" + libdir, "-ltgb_host", "-Wl,-Bsymbolic", "-Wl,-rpath," + libdir]
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
