"""Build libdeepfm_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo).

    python -m deepfm_b200.build [--force]
"""

from __future__ import annotations

import glob
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "build")
LIB = os.path.join(HERE, "libdeepfm_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
         "-Xptxas", "-v"]


def _sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _deps():
    return _sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + \
        glob.glob(os.path.join(HERE, "..", "include", "*.h")) + [os.path.abspath(__file__)]


def up_to_date() -> bool:
    if not os.path.exists(LIB):
        return False
    t = os.path.getmtime(LIB)
    return all(os.path.getmtime(p) <= t for p in _deps())


def _compile(src: str) -> str:
    obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
    newest = max(os.path.getmtime(p) for p in [src] + glob.glob(os.path.join(CSRC, "*.cuh")) +
                 glob.glob(os.path.join(HERE, "..", "include", "*.h")) + [os.path.abspath(__file__)])
    if os.path.exists(obj) and os.path.getmtime(obj) >= newest:
        return obj
    cmd = [NVCC, *FLAGS, "-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    log = os.path.join(OBJ, os.path.basename(src)[:-3] + ".ptxas.log")
    with open(log, "w") as fh:
        fh.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    return obj


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and up_to_date():
        return LIB
    if not os.path.exists(NVCC):
        raise RuntimeError(f"nvcc not found at {NVCC}; libdeepfm_b200.so cannot be built")
    os.makedirs(OBJ, exist_ok=True)
    if force:
        for o in glob.glob(os.path.join(OBJ, "*.o")):
            os.remove(o)
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(_compile, _sources()))
    cmd = [NVCC, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    if verbose:
        print(f"built {LIB} from {len(objs)} objects")
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
