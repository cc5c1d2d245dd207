"""
In-tree build of the engine's only native artefact, ``csrc/liblrc.so`` (sm_100a, C ABI of include/lrc.h).

    python -m <package>.build         or        __graft_entry__.build()

nvcc cross-compiles without a GPU; the shared object is git-ignored but travels with the repository
snapshot to the GPU box.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "liblrc.so")
SOURCES = ["bvh_build.cu", "scan.cu", "post.cu", "plan.cu", "nn.cu"]
HEADERS = ["common.cuh", "traverse.cuh", "scan_util.cuh", "exchange.cuh", os.path.join("..", "..", "include", "lrc.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC,-fvisibility=default",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the engine cannot be built (there is no CPU fallback)")


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build_native(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/*.cu into csrc/liblrc.so for sm_100a.  Returns the library path."""
    if not force and not _stale():
        return LIB
    nvcc = [_nvcc()]
    # the image exports CC=/opt/gcc/bin/gcc, which lacks some specs; nvcc must use the system g++
    if os.path.exists("/usr/bin/g++"):
        nvcc += ["-ccbin", "/usr/bin/g++"]
    objdir = os.path.join(CSRC, "build")
    os.makedirs(objdir, exist_ok=True)
    hdr_time = max(os.path.getmtime(os.path.join(CSRC, h)) for h in HEADERS)
    hdr_time = max(hdr_time, os.path.getmtime(os.path.abspath(__file__)))

    def compile_one(src):
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        path = os.path.join(CSRC, src)
        if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(path), hdr_time):
            return src, 0, "(up to date)\n"
        cmd = nvcc + NVCC_FLAGS + ["-c", "-o", obj, path]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        return src, r.returncode, " ".join(cmd) + "\n" + r.stdout

    # one translation unit per source, compiled side by side (scan.cu with its kernel variants dominates), then linked
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=len(SOURCES)) as pool:
        results = list(pool.map(compile_one, SOURCES))
    text = "".join(out for _, _, out in results)
    rc = max(code for _, code, _ in results)
    if rc == 0:
        cmd = nvcc + ["-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB] + [os.path.join(objdir, s.replace(".cu", ".o")) for s in SOURCES]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        text += " ".join(cmd) + "\n" + r.stdout
        rc = r.returncode
    log = os.path.join(CSRC, "build.log")
    with open(log, "w") as f:
        f.write(text)
    if verbose or rc != 0:
        sys.stderr.write(text)
    if rc != 0:
        raise RuntimeError(f"nvcc failed ({rc}); see {log}")
    return LIB


if __name__ == "__main__":
    print(build_native(force="--force" in sys.argv, verbose=True))
