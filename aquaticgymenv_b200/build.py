"""In-tree build of ``libaquagym.so`` (nvcc, ``-gencode arch=compute_100a,code=sm_100a -lineinfo``).
nvcc cross-compiles without a GPU; the .so stays next to the package so that it travels to the
GPU box with the repo snapshot."""
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB = os.path.join(_HERE, "libaquagym.so")


def build(force=False, verbose=False):
    cmd = ["make", "-j8", "-C", CSRC] + (["-B"] if force else [])
    out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or out.returncode != 0:
        print(out.stdout)
    if out.returncode != 0:
        raise RuntimeError("building libaquagym.so failed (see output above)")
    return LIB
