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


VARIANTS_LIB = os.path.join(os.path.dirname(PKG), "tools", "_variants", "libeagle_b200_variants.so")


def build_variants(verbose: bool = False) -> str:
    """The measurement build (-DEGL_BENCH_VARIANTS: A/B kernels and EGL_*_VARIANT / EGL_UPLOAD_* switches) as a SEPARATE
    library under tools/_variants/, next to the shipped one: probes load it by path, the package never does."""
    os.makedirs(os.path.dirname(VARIANTS_LIB), exist_ok=True)
    return build_native(True, verbose, lib=VARIANTS_LIB, objdir=os.path.join(PKG, "build", "variants"), variants=True)


def build_native(force: bool = False, verbose: bool = False, *, lib: str = LIB, objdir: str | None = None, variants: bool = False) -> str:
    """Compile every .cu under csrc/ into one shared object; returns its path."""
    build_assemble(force)
    if not force and not _stale():
        return LIB
    nvcc = nvcc_path()
    objs = []
    objdir = objdir or os.path.join(PKG, "build")
    os.makedirs(objdir, exist_ok=True)
    common = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"] + ARCH
    if variants or os.environ.get("EGL_BENCH_VARIANTS") == "1":   # A/B kernels + EGL_*_VARIANT switches (tools/sweep.py, tools/sanitize.sh)
        common += ["-DEGL_BENCH_VARIANTS"]
    if verbose:
        common += ["-Xptxas", "-v"]
    procs = []
    for src in SOURCES:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
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
    subprocess.check_call([nvcc, "-shared", "-o", lib] + objs + ARCH + ["-Xcompiler", "-fPIC", "-lcudart", "-lpthread"])
    return lib


if __name__ == "__main__":
    if "--variants" in sys.argv:
        print(build_variants(verbose="-v" in sys.argv))
    else:
        print(build_native(force="--force" in sys.argv, verbose="-v" in sys.argv))
