"""Builds libswb200.so (the sm_100a kernels + C ABI) in-tree with nvcc.

``python -m sw_solver_b200.build`` or ``__graft_entry__.build()``.  nvcc cross-compiles
without a GPU; the resulting shared object travels to the GPU box with the repo snapshot.
"""

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libswb200.so")
SOURCES = [os.path.join(CSRC, "swb200.cu")]
HEADERS = [os.path.join(CSRC, h) for h in ("swb_types.cuh", "swb_scheme.cuh", "swb_step.cuh", "swb_kernels.cuh")] + [
    os.path.join(ROOT, "include", "swb200.h")]


def nvcc_path() -> str:
    cand = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(cand):
        raise NvccMissing("nvcc not found: libswb200.so cannot be built (there is no CPU fallback)")
    return cand


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(f) > t for f in SOURCES + HEADERS + [os.path.abspath(__file__)])


class NvccMissing(RuntimeError):
    """nvcc is not installed here (a GPU box without the toolkit): a prebuilt library may still be used."""


def build(force: bool = False, verbose: bool = False, out: str = LIB) -> str:
    """Compile ``out``.  Safe when several processes call it at once (torchrun: every rank imports
    the package): one builds under an exclusive file lock, into a temporary file that is moved into
    place atomically, so nobody ever maps a half-written object; the others find it up to date once
    they get the lock."""
    import fcntl

    if out == LIB and not force and not needs_build():
        return LIB
    with open(out + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if out == LIB and not force and not needs_build():  # another process built it meanwhile
                return LIB
            return _build_locked(verbose, out)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(verbose: bool, out: str) -> str:
    tmp = f"{out}.tmp.{os.getpid()}"
    cmd = [
        nvcc_path(),
        "-gencode", "arch=compute_100a,code=sm_100a",
        "-O3", "-std=c++17", "-lineinfo",
        # no implicit FMA contraction: STRICT needs numpy's operation-by-operation rounding, and
        # FAST spells every fused multiply-add out, so that all instantiations of the step kernel
        # (hot / boundary rows, TMA / direct loads, any band decomposition) round identically
        "-fmad=false",
        "-Xcompiler", "-fPIC", "-shared",
        "-I", os.path.join(ROOT, "include"), "-I", CSRC,
        "-o", tmp,
    ] + SOURCES
    cmd[1:1] = os.environ.get("SWB_NVCC_FLAGS", "").split()  # dev builds, e.g. -DSWB_RES_TIMING
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd))
    try:
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
        os.replace(tmp, out)
    finally:
        if os.path.exists(tmp):
            os.unlink(tmp)
    if verbose:
        print(res.stdout + res.stderr)
    return out


if __name__ == "__main__":
    outs = [a.split("=", 1)[1] for a in sys.argv if a.startswith("--out=")]
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, out=outs[0] if outs else LIB))
