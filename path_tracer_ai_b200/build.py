"""In-tree build of libb2pt.so (the C-ABI engine) for sm_100a.

    python -m path_tracer_ai_b200.build [--force] [--verbose]

nvcc cross-compiles without a GPU.  The library is written next to this file so it travels to the
GPU box with the repo snapshot; nothing is JIT-compiled at run time.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
HOST = os.path.join(HERE, "host")
LIB = os.path.join(HERE, "libb2pt.so")
CLI = os.path.join(HERE, "b2pt_cli")

CU_SOURCES = ["api.cu", "build.cu", "trace.cu", "render.cu", "multi.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    # Parity: never contract a*b+c into an FMA on either side (exact.cuh also uses the _rn
    # intrinsics, which are immune to this flag; this is the belt to those braces).
    "-fmad=false",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-O2",
    "-Xptxas", "-warn-spills",
]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (needed to build libb2pt.so)")


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _deps():
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps += [os.path.join(HOST, f) for f in os.listdir(HOST)]
    deps.append(os.path.join(os.path.dirname(HERE), "include", "b2pt_host.h"))
    deps.append(os.path.join(os.path.dirname(HERE), "include", "b2pt.h"))
    deps.append(os.path.abspath(__file__))
    return deps


def build_lib(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale(LIB, _deps()):
        return LIB
    extra = os.environ.get("B2PT_EXTRA_NVCC", "").split()   # experiments: e.g. -DB2PT_SHD_BLOCK=64
    out = os.environ.get("B2PT_LIB_OUT", LIB)
    cmd = [_nvcc(), *NVCC_FLAGS, *extra, "-shared", "-o", out] + [os.path.join(CSRC, s) for s in CU_SOURCES]
    cmd += [os.path.join(HOST, "host_api.cpp"), "-lz"]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
        print(" ".join(cmd))
    env = dict(os.environ)
    # The image exports CXX=/opt/gcc/bin/g++; nvcc should use the distro host compiler.
    if os.path.exists("/usr/bin/g++"):
        cmd[1:1] = ["-ccbin", "/usr/bin/g++"]
    res = subprocess.run(cmd, env=env, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stdout + res.stderr)
    return out


def build_cli(force: bool = False, verbose: bool = False) -> str:
    """The reference-compatible command line (host/main.cpp), linked against libb2pt.so."""
    main_cpp = os.path.join(HOST, "main.cpp")
    if not os.path.exists(main_cpp):
        return ""
    deps = [os.path.join(HOST, f) for f in os.listdir(HOST)] + [LIB]
    if not force and not _stale(CLI, deps):
        return CLI
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    cmd = [cxx, "-std=c++17", "-O2", "-ffp-contract=off", "-fopenmp", "-I", os.path.join(os.path.dirname(HERE), "include"),
           "-o", CLI, main_cpp, "-L", HERE, "-lb2pt", "-Wl,-rpath,$ORIGIN", "-lz"]
    if verbose:
        print(" ".join(cmd))
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("g++ failed:\n" + res.stdout + res.stderr)
    return CLI


def main(argv=None) -> int:
    argv = sys.argv[1:] if argv is None else argv
    force = "--force" in argv
    verbose = "--verbose" in argv
    print(build_lib(force=force, verbose=verbose))
    cli = build_cli(force=force, verbose=verbose)
    if cli:
        print(cli)
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
