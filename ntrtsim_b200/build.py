"""In-tree build of the CUDA library, the host API library and the CPU oracle (test infrastructure)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build_all(verbose=False):
    r = subprocess.run(["make", "-C", ROOT, "-j8", "all"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, universal_newlines=True)
    if verbose or r.returncode != 0:
        print(r.stdout)
    if r.returncode != 0:
        raise RuntimeError("build failed")
