"""Builds libmsgpu.so (hand-written CUDA for sm_100a + the C ABI of include/msgpu.h) in-tree with nvcc.

    python -m muchsalsa_b200.build [--force]

nvcc cross-compiles without a GPU; the resulting .so is git-ignored but travels to the GPU box with the snapshot.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
ROOT = os.path.dirname(HERE)
LIB = os.path.join(HERE, "libmsgpu.so")
STAMP = os.path.join(HERE, ".libmsgpu.stamp")

SOURCES = ["msgpu_api.cu", "ingest.cu", "pair_sort.cu", "pair_fused.cu", "chain.cu", "filter.cu", "shard.cu", "seq.cu", "digest.cu", "coverage.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
    "--fmad=false",  # the score arithmetic must not be contracted into FMAs (bit-exact doubles, MatchMap.cpp:195-202)
    "-Xcompiler", "-fPIC,-O2,-Wall,-Wno-unused-function", "-Xptxas", "-v", "--expt-relaxed-constexpr",
]


NVCC_FLAGS += os.environ.get("MSGPU_NVCC_EXTRA", "").split()  # development experiments only (e.g. -DMSGPU_IDX_BITS=13)


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.sep not in c or os.path.exists(c)):
            return c
    raise RuntimeError("nvcc not found")


def _sources() -> list[str]:
    return [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def _digest() -> str:
    h = hashlib.sha256()
    h.update(" ".join(NVCC_FLAGS).encode())
    deps = sorted(os.listdir(CSRC)) + [os.path.join(ROOT, "include", "msgpu.h")]
    for name in deps:
        p = name if os.path.isabs(name) else os.path.join(CSRC, name)
        if os.path.isfile(p):
            h.update(name.encode())
            with open(p, "rb") as f:
                h.update(f.read())
    return h.hexdigest()


def build_lib(force: bool = False, verbose: bool = False) -> str:
    """Compile every .cu into libmsgpu.so if the sources changed.  Returns the library path."""
    dig = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(STAMP) and open(STAMP).read().strip() == dig:
        return LIB
    objs = []
    logs = []
    procs = []
    for src in _sources():
        obj = os.path.join(CSRC, os.path.basename(src).replace(".cu", ".o"))
        objs.append(obj)
        cmd = [_nvcc(), *NVCC_FLAGS, "-c", src, "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        logs.append(f"==== {os.path.basename(src)}\n{out}")
        if p.returncode != 0:
            failed = True
    log = "\n".join(logs)
    with open(os.path.join(HERE, "build.log"), "w") as f:
        f.write(log)
    if failed or verbose:
        print(log, file=sys.stderr)
    if failed:
        raise RuntimeError("nvcc failed (see muchsalsa_b200/build.log)")
    subprocess.run([_nvcc(), "-shared", "-o", LIB, *objs, "-cudart", "static"], check=True)
    with open(STAMP, "w") as f:
        f.write(dig)
    return LIB


def build_variant(tag: str, extra_flags: list[str]) -> str:
    """Development only: a second copy of the library compiled with extra nvcc flags (-D...), as libmsgpu_<tag>.so —
    selected on the GPU box with MSGPU_LIB=... for A/B measurements of compile-time constants."""
    out = os.path.join(HERE, f"libmsgpu_{tag}.so")
    objdir = os.path.join(HERE, f".obj_{tag}")
    os.makedirs(objdir, exist_ok=True)
    procs = []
    for src in _sources():
        obj = os.path.join(objdir, os.path.basename(src).replace(".cu", ".o"))
        procs.append((obj, subprocess.Popen([_nvcc(), *NVCC_FLAGS, *extra_flags, "-c", src, "-o", obj],
                                            stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for obj, p in procs:
        out_text, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(out_text)
    subprocess.run([_nvcc(), "-shared", "-o", out, *[o for o, _ in procs], "-cudart", "static"], check=True)
    return out


if __name__ == "__main__":
    if "--variant" in sys.argv:
        i = sys.argv.index("--variant")
        print(build_variant(sys.argv[i + 1], sys.argv[i + 2:]))
    else:
        print(build_lib(force="--force" in sys.argv, verbose="-v" in sys.argv))
