"""Build libnka_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libnka_b200.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC"]


def sources() -> list[str]:
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps() -> list[str]:
    inc = os.path.join(os.path.dirname(HERE), "include")
    out = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    out += [os.path.join(inc, f) for f in os.listdir(inc)]
    out.append(__file__)
    return out


def is_stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(p) > t for p in _deps())


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compile every .cu under csrc/ into one shared library.  Objects are built
    in parallel (one nvcc per translation unit)."""
    if not force and not is_stale():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    nvcc = os.environ.get("NVCC", "nvcc")
    objs, procs = [], []
    for src in sources():
        obj = os.path.join(LIB_DIR, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        cmd = [nvcc, *NVCC_FLAGS, "-c", src, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError("nvcc failed: %s\n%s" % (" ".join(cmd), out))
        if verbose:
            print(out)
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB_PATH, *objs, "-ldl"]
    r = subprocess.run(link, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed: %s\n%s%s" % (" ".join(link), r.stdout, r.stderr))
    return LIB_PATH


def build_variant(tag: str, defines: dict, instantiate_max: int = 11) -> str:
    """Tuning builds (tools/tune.py): the same sources with -D overrides, written next to
    the product library as lib/variants/libnka_b200_<tag>.so.  Never loaded by default."""
    vdir = os.path.join(LIB_DIR, "variants")
    os.makedirs(vdir, exist_ok=True)
    out = os.path.join(vdir, "libnka_b200_%s.so" % tag)
    nvcc = os.environ.get("NVCC", "nvcc")
    dflags = ["-D%s=%s" % kv for kv in defines.items()] + ["-DNKA_INSTANTIATE_MAX=%d" % instantiate_max]
    cmd = [nvcc, *NVCC_FLAGS, *dflags, "-shared", "-o", out, *sources(), "-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed: %s\n%s%s" % (" ".join(cmd), r.stdout, r.stderr))
    return out


if __name__ == "__main__":
    import sys
    print(build_library(force=True, verbose="-v" in sys.argv))
