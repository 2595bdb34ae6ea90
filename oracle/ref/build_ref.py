#!/usr/bin/env python
"""Compile the reference's own cable / anchor / actuator / controller / CPG sources, unchanged and where they lie under
/root/reference/src, against the shim headers of oracle/ref/shim (Bullet and Boost are not vendored with the
reference), together with oracle/ref/ref_harness.cpp, into oracle/_ref/libntrt_ref.so (git-ignored).
No reference source is copied into this repository."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("NTRT_REFERENCE", "/root/reference")
OUT = os.path.join(ROOT, "oracle", "_ref")

REFERENCE_SOURCES = [
    "src/core/tgSpringCable.cpp",
    "src/core/tgBulletSpringCable.cpp",
    "src/core/tgBulletSpringCableAnchor.cpp",
    "src/core/tgBulletContactSpringCable.cpp",
    "src/core/tgSpringCableActuator.cpp",
    "src/core/tgBasicActuator.cpp",
    "src/core/tgKinematicActuator.cpp",
    "src/core/tgModel.cpp",
    "src/core/tgSenseable.cpp",
    "src/core/tgTagSearch.cpp",
    "src/core/tgTags.cpp",
    "src/core/abstractMarker.cpp",
    "src/controllers/tgBasicController.cpp",
    "src/controllers/tgTensionController.cpp",
    "src/controllers/tgImpedanceController.cpp",
    "src/controllers/tgPIDController.cpp",
    "src/util/CPGNode.cpp",
    "src/util/CPGEquations.cpp",
]


def build():
    if not os.path.isdir(os.path.join(REF, "src")):
        print("reference tree not present at %s: nothing to build" % REF)
        return False
    os.makedirs(OUT, exist_ok=True)
    srcs = [os.path.join(REF, s) for s in REFERENCE_SOURCES if os.path.exists(os.path.join(REF, s))]
    cmd = ["g++", "-std=c++11", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-w", "-DBT_USE_DOUBLE_PRECISION",
           # bind the reference's symbols inside this library: the host mirror (libtgb_host.so, loaded RTLD_GLOBAL) defines classes of
           # the same names and must not interpose on them
           "-Wl,-Bsymbolic",
           "-I", os.path.join(HERE, "shim"), "-I", os.path.join(ROOT, "ntrtsim_b200", "host"), "-I", os.path.join(REF, "src"),
           "-o", os.path.join(OUT, "libntrt_ref.so"), os.path.join(HERE, "ref_harness.cpp")] + srcs
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, universal_newlines=True)
    if r.returncode != 0:
        print(r.stdout[-6000:])
        raise RuntimeError("reference build failed")
    return True


if __name__ == "__main__":
    ok = build()
    sys.exit(0 if ok or not os.path.isdir(REF) else 1)
