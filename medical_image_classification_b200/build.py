"""Build libb200ssm.so in-tree with nvcc for sm_100a (no torch headers involved: the library is a
plain C-ABI shared object, see include/b200_ssm.h).

    python -m medical_image_classification_b200.build [--force] [--verbose]
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libb200ssm.so")
STAMP = os.path.join(LIB_DIR, "libb200ssm.stamp")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest() -> str:
    h = hashlib.sha256()
    inc = os.path.join(os.path.dirname(HERE), "include", "b200_ssm.h")
    files = sources() + sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))) + [inc]
    for f in files:
        h.update(os.path.basename(f).encode())   # not the path: the snapshot on a GPU box lives elsewhere
        with open(f, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def can_build() -> bool:
    return bool(shutil.which("nvcc")) or os.path.exists("/usr/local/cuda/bin/nvcc")


def nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: libb200ssm.so cannot be built")
    return exe


def _fresh(dig: str) -> bool:
    return os.path.exists(LIB_PATH) and os.path.exists(STAMP) and open(STAMP).read().strip() == dig


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/*.cu into lib/libb200ssm.so.  Safe under `torchrun` on a fresh clone: ranks serialise on a file lock, the
    winner builds into temporary names and renames them into place, the others find the stamp fresh and return."""
    import fcntl
    os.makedirs(LIB_DIR, exist_ok=True)
    dig = _digest()
    if not force and _fresh(dig):
        return LIB_PATH
    with open(os.path.join(LIB_DIR, ".lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and _fresh(dig):      # another process built it while this one waited
                return LIB_PATH
            return _build_locked(dig, verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(dig: str, verbose: bool) -> str:
    objs = []
    procs = []
    env = dict(os.environ)
    for src in sources():
        obj = os.path.join(LIB_DIR, os.path.basename(src)[:-3] + f".{os.getpid()}.o")
        objs.append(obj)
        cmd = [nvcc(), *NVCC_FLAGS, "-c", src, "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env)))
    log = []
    for src, pr in procs:
        out, _ = pr.communicate()
        log.append(f"==== {os.path.basename(src)}\n{out}")
        if pr.returncode != 0:
            sys.stderr.write("\n".join(log))
            raise RuntimeError(f"nvcc failed on {src}")
    with open(os.path.join(LIB_DIR, "ptxas.log"), "w") as fh:
        fh.write("\n".join(log))
    if verbose:
        print("\n".join(log))
    tmp_so = LIB_PATH + f".{os.getpid()}.tmp"
    subprocess.run([nvcc(), "-shared", "-o", tmp_so, *objs, "-lcudart"], check=True)
    for obj in objs:                                        # keep lib/<name>.o for cuobjdump, under the stable name
        os.replace(obj, obj.replace(f".{os.getpid()}.o", ".o"))
    os.replace(tmp_so, LIB_PATH)                            # atomic: a concurrent dlopen sees the old or the new file, never half of one
    with open(STAMP + ".tmp", "w") as fh:
        fh.write(dig)
    os.replace(STAMP + ".tmp", STAMP)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
