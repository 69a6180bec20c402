"""In-tree build of the CUDA library (nvcc, sm_100a only) and of the CPU emulation used by tests."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.path.join(CSRC, "libtomatis_b200.so")
# The library is compiled once per frame size the fused kernels serve (csrc/fft4096.cuh, TMT_NFFT): the reference's default
# 4096 / 2048 and its documented faster setting 2048 / 1024 (docs/Tomatis技术说明.md:253-258, "pair mode" of the same kernels).
FUSED_SIZES = {4096: 2048, 2048: 1024}                       # n_fft -> hop
LIB_PATHS = {4096: LIB_PATH, 2048: os.path.join(CSRC, "libtomatis_b200_n2048.so")}
EMUL_PATH = os.path.join(CSRC, "libtomatis_emul.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def _stale(target: str, sources) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build_library(force: bool = False, verbose: bool = False, n_fft: int = 4096) -> str:
    """Compile csrc/tomatis_b200.cu -> csrc/libtomatis_b200.so (n_fft 4096) or csrc/libtomatis_b200_n2048.so for sm_100a."""
    lib_path = LIB_PATHS[n_fft]
    srcs = [os.path.join(CSRC, f) for f in ("tomatis_b200.cu", "fft4096.cuh", "spectrum.cuh", "calib.cuh", "generic.cuh", "host_tables.hpp")]
    srcs.append(os.path.join(os.path.dirname(HERE), "include", "tomatis_b200.h"))
    if force or _stale(lib_path, srcs):
        cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + [f"-DTMT_NFFT={n_fft}", "-o", lib_path, srcs[0]]
        res = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
        if verbose:
            print(res.stderr)
    return lib_path


def build_libraries(force: bool = False, verbose: bool = False):
    """Every fused frame size.  The two compilations are independent: run them side by side."""
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(len(LIB_PATHS)) as ex:
        return list(ex.map(lambda n: build_library(force, verbose, n), sorted(LIB_PATHS, reverse=True)))


def build_emulation(force: bool = False) -> str:
    """Host-only build of the FFT stage functions (csrc/host_emul.cu) for the CPU tests."""
    srcs = [os.path.join(CSRC, f) for f in ("host_emul.cu", "fft4096.cuh", "spectrum.cuh", "calib.cuh", "generic.cuh", "host_tables.hpp")]
    if force or _stale(EMUL_PATH, srcs):
        cmd = [_nvcc(), "-O2", "-std=c++17", "-Xcompiler", "-fPIC", "-shared", "-Wno-deprecated-gpu-targets",
               "-o", EMUL_PATH, srcs[0]]
        res = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("nvcc (host emulation) failed:\n" + res.stdout + res.stderr)
    return EMUL_PATH


if __name__ == "__main__":
    print(build_libraries(force=True, verbose=True))
    print(build_emulation(force=True))
