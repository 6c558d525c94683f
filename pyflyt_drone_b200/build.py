"""In-tree build of libfwsim.so (hand-written sm_100a CUDA behind the C ABI of include/fwsim.h).

nvcc cross-compiles for sm_100a without a GPU; the .so is git-ignored but travels to the GPU box.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libfwsim.so")
SOURCES = ["fw_api.cu", "fw_kernels.cu", "fw_render.cu", "ppo_kernels.cu", "ppo_tc.cu", "ppo_tc_a6.cu", "ppo_tc_d64.cu", "ppo_update_tc.cu", "ppo_update_tc_a6.cu", "ppo_update_tc_d64.cu", "ppo_api.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    # flush fp32 denormals: MUFU.RCP/RSQ/SIN/COS then need no 2^24 rescue sequences (4-6 instructions each)
    "--ftz=true",
    # IEEE-rounded division / sqrt cost 10-20 instructions each; the 2-ulp SFU forms are far inside the parity
    # tolerance (1e-4 relative per step) and are what the hot loops are written for
    "--prec-div=false", "--prec-sqrt=false",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-O2", "--threads", "4",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libfwsim.so cannot be built (there is no CPU fallback)")


def have_nvcc() -> bool:
    try:
        _nvcc()
        return True
    except RuntimeError:
        return False


def sources() -> list[str]:
    return [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "fwsim.h")]
    inc = os.path.join(HERE, "..", "include")
    deps += [os.path.join(inc, f) for f in os.listdir(inc)]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False, out: str | None = None, defines: tuple[str, ...] = ()) -> str:
    """`out` / `defines`: experiment builds (scripts/ab_build.py) next to the product library."""
    if out is None and not force and not is_stale():
        return LIB
    cmd = [_nvcc(), *NVCC_FLAGS, *[f"-D{d}" for d in defines], "-shared", "-o", out or LIB, *sources()]
    if verbose:
        cmd[1:1] = ["-Xptxas", "-v"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed ({r.returncode}):\n{' '.join(cmd)}\n{r.stdout}\n{r.stderr}")
    if verbose:
        print(r.stderr)
    return out or LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
