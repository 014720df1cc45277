"""Build the sm_100a shared library (and, optionally, the parity oracle) in-tree.

`python -m vk_gaussian_splatting_b200.build` or `build_all()` from __graft_entry__.build().
nvcc cross-compiles without a GPU. Objects are rebuilt only when a source or header is newer.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
OUT_DIR = PKG / "lib"
OBJ_DIR = PKG / "build"
LIB = OUT_DIR / "libvkgs_b200.so"

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-ffp-contract=off,-fvisibility=hidden",
          f"-I{ROOT / 'include'}", f"-I{CSRC}"]
# per-file extra flags: the per-splat front end must not contract mul+add (bit-exact vs the oracle)
EXTRA = {"k_preprocess.cu": ["-fmad=false"], "k_metrics.cu": ["-fmad=false"]}
SOURCES = ["context.cu", "sort_api.cu", "k_preprocess.cu", "k_radix_sort.cu", "k_binning.cu", "k_blend.cu", "k_metrics.cu",
           "host_camera.cpp", "host_pack.cpp", "host_synth.cpp", "host_loader.cpp"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found; the CUDA path cannot be built")


def _host_cxx() -> str:
    # the image exports CXX=/opt/gcc/bin/g++ (a wrapper without OpenMP specs); prefer the system one
    return "/usr/bin/g++" if Path("/usr/bin/g++").exists() else "g++"


def _newer(target: Path, deps) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(Path(d).stat().st_mtime > t for d in deps)


def build_variant(tag: str, defines, verbose: bool = False) -> Path:
    """Tuning builds: lib/libvkgs_b200_<tag>.so compiled with extra -D defines (tools/ab_variants.py benchmarks several of
    them in one GPU call; a process picks one with VKGS_LIB=<path>). Never used by the product path."""
    return build_lib(verbose, force=False, tag=tag, defines=list(defines))


def build_lib(verbose: bool = False, force: bool = False, tag: str = "", defines=()) -> Path:
    OUT_DIR.mkdir(exist_ok=True)
    global OBJ_DIR, LIB
    obj_dir_saved, lib_saved = OBJ_DIR, LIB
    if tag:
        OBJ_DIR = PKG / "build" / f"variant_{tag}"
        LIB = OUT_DIR / f"libvkgs_b200_{tag}.so"
    try:
        return _build_lib(verbose, force, [f"-D{d}" for d in defines])
    finally:
        OBJ_DIR, LIB = obj_dir_saved, lib_saved


def _build_lib(verbose: bool, force: bool, extra_defines) -> Path:
    OBJ_DIR.mkdir(parents=True, exist_ok=True)
    nvcc = _nvcc()
    headers = list(CSRC.glob("*.hpp")) + list(CSRC.glob("*.cuh")) + [ROOT / "include" / "vkgs_b200.h"]
    objs = []
    for src in SOURCES:
        s = CSRC / src
        o = OBJ_DIR / (src + ".o")
        objs.append(o)
        if force or _newer(o, [s, *headers, Path(__file__)]):
            cmd = [nvcc, "-ccbin", _host_cxx(), *ARCH, *COMMON, *EXTRA.get(src, []), *extra_defines, "-x", "cu", "-c", str(s), "-o", str(o)]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
                print(" ".join(cmd), flush=True)
            subprocess.run(cmd, check=True)
    if force or _newer(LIB, objs):
        # zlib: gzip container of .spz (the image's libz.a is not PIC, so the system libz.so.1 is used)
        cmd = [nvcc, "-ccbin", _host_cxx(), *ARCH, "-shared", "-o", str(LIB), *map(str, objs), "-lz"]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.run(cmd, check=True)
    return LIB


def build_oracle(verbose: bool = False) -> Path:
    """Compile oracle/ (test infrastructure): the C restatement + the CPU sorter baseline.
    When /root/reference exists (build container only) also refresh the glm golden vectors."""
    odir = ROOT / "oracle"
    subprocess.run(["make", "-C", str(odir), "libvkgs_oracle.so"], check=True,
                   stdout=None if verbose else subprocess.DEVNULL)
    return odir / "libvkgs_oracle.so"


def build_all(verbose: bool = False) -> None:
    build_lib(verbose)
    build_oracle(verbose)


if __name__ == "__main__":
    build_all(verbose="-v" in sys.argv)
    print(LIB)
