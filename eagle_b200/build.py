"""Build the CUDA extension in-tree: eagle_b200/libeagle_b200.so (sm_100a only).

``python -m eagle_b200.build`` or ``__graft_entry__.build()``.  nvcc cross-compiles without a GPU.
The .so is git-ignored but travels with the source tree to the GPU box.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libeagle_b200.so")
SOURCES = ["common.cu", "preprocess.cu", "decode.cu", "synthesize.cu", "fit.cu", "project.cu", "flow.cu", "upload.cu"]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def assemble_module_path() -> str:
    import sysconfig
    return os.path.join(PKG, "_assemble" + sysconfig.get_config_var("EXT_SUFFIX"))


def build_assemble(force: bool = False) -> str:
    """The CPython extension that builds the result dicts (csrc/assemble.c), compiled with gcc."""
    import sysconfig
    out = assemble_module_path()
    src = os.path.join(CSRC, "assemble.c")
    if not force and os.path.exists(out) and os.path.getmtime(out) >= os.path.getmtime(src):
        return out
    cc = shutil.which("gcc") or shutil.which("cc")
    if not cc:
        raise RuntimeError("gcc not found")
    subprocess.check_call([cc, "-O2", "-fPIC", "-shared", "-Wall", "-I", sysconfig.get_paths()["include"], src, "-o", out])
    return out


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f != "assemble.c"] + [os.path.join(os.path.dirname(PKG), "include", "eagle_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build_native(force: bool = False, verbose: bool = False) -> str:
    """Compile every .cu under csrc/ into one shared object; returns its path."""
    build_assemble(force)
    if not force and not _stale():
        return LIB
    nvcc = nvcc_path()
    objs = []
    os.makedirs(os.path.join(PKG, "build"), exist_ok=True)
    common = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"] + ARCH
    if os.environ.get("EGL_BENCH_VARIANTS") == "1":   # A/B kernels + EGL_*_VARIANT switches (tools/sweep.py, tools/sanitize.sh)
        common += ["-DEGL_BENCH_VARIANTS"]
    if verbose:
        common += ["-Xptxas", "-v"]
    procs = []
    for src in SOURCES:
        obj = os.path.join(PKG, "build", src.replace(".cu", ".o"))
        cmd = [nvcc, "-c", os.path.join(CSRC, src), "-o", obj] + common
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write(f"--- nvcc {src} failed ---\n{out}\n")
        elif verbose or out.strip():
            sys.stderr.write(f"--- nvcc {src} ---\n{out}\n")
    if failed:
        raise RuntimeError("nvcc failed; see output above")
    subprocess.check_call([nvcc, "-shared", "-o", LIB] + objs + ARCH + ["-Xcompiler", "-fPIC", "-lcudart", "-lpthread"])
    return LIB


if __name__ == "__main__":
    print(build_native(force="--force" in sys.argv, verbose="-v" in sys.argv))
