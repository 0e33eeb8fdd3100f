"""In-tree build of the sm_100a shared library (nvcc cross-compiles without a GPU)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.path.join(HERE, "libsisua_b200.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def _sources():
  return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh"))) + \
      [os.path.join(os.path.dirname(HERE), "include", "sisua_b200.h")]


def needs_build() -> bool:
  if not os.path.exists(LIB_PATH):
    return True
  t = os.path.getmtime(LIB_PATH)
  return any(os.path.getmtime(s) > t for s in _sources())


def build(force: bool = False, verbose: bool = False) -> str:
  if not force and not needs_build():
    return LIB_PATH
  nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
  if not os.path.exists(nvcc):
    raise RuntimeError("nvcc not found: cannot build libsisua_b200.so")
  defs = []
  if os.path.exists(os.path.join(CSRC, "kernels_tc.cuh")):
    defs.append("-DSISUA_WITH_TC")
  defs += os.environ.get("SISUA_NVCC_DEFS", "").split()      # tuning experiments, e.g. -DSISUA_OUT_KU=4
  cmd = [nvcc] + NVCC_FLAGS + defs + [os.path.join(CSRC, "abi.cu"), "-o", LIB_PATH, "-lcuda"]
  if verbose:
    cmd.insert(-4, "-Xptxas")
    cmd.insert(-4, "-v")
    print(" ".join(cmd))
  r = subprocess.run(cmd, capture_output=True, text=True)
  if r.returncode != 0:
    raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
  if verbose:
    print(r.stderr)
  return LIB_PATH


if __name__ == "__main__":
  print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
