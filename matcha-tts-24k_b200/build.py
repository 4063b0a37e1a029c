"""Builds libcfm_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m matcha_tts_24k_b200.build [--force] [-v]

The tensor-core kernels are explicit template instantiations compiled one nvcc process each (csrc/gemm_inst.cu with
-DCFM_BN/-DCFM_NEW/-DCFM_PAIR, csrc/attn_inst.cu) in parallel; csrc/cfm.cu (C ABI, schedule, bandwidth kernels) only sees
`extern template` declarations.  Objects are cached under csrc/build/ and rebuilt when a source or header is newer.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "build")
LIB = os.path.join(HERE, "libcfm_b200.so")
HEADERS = ["ptx.cuh", "gemm.cuh", "kernels.cuh", "attn.cuh", "attn_tc.cuh", "attn_persist.cuh", "ff_fused.cuh", "rowln.cuh", os.path.join("..", "..", "include", "cfm_b200.h")]
TC = [(64, 8), (128, 8), (160, 8), (192, 8), (256, 8), (256, 12)]  # gemm_tc_kernel<BN, epilogue warps>  (gemm.cuh CFM_FOR_EACH_TC)
TC2 = [128, 160, 192, 256]                                          # gemm_tc2_kernel<BN>                  (CFM_FOR_EACH_TC2)
COMMON = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
          "-static-global-template-stub=false", "-diag-suppress", "20279,20281"]


def units():
    """(object name, source, extra defines)"""
    out = [("cfm.o", "cfm.cu", []), ("attn_inst.o", "attn_inst.cu", []), ("ff_inst.o", "ff_inst.cu", []), ("rowln_inst.o", "rowln_inst.cu", [])]
    out += [(f"gemm_tc_{bn}_{new}.o", "gemm_inst.cu", [f"-DCFM_BN={bn}", f"-DCFM_NEW={new}", "-DCFM_PAIR=0"]) for bn, new in TC]
    out += [(f"attn_simt_{bf}_{d}.o", "attn_simt_inst.cu", [f"-DCFM_BF16={bf}", f"-DCFM_D={d}"]) for bf in (0, 1) for d in (32, 64)]
    out += [(f"gemm_tc2_{bn}.o", "gemm_inst.cu", [f"-DCFM_BN={bn}", "-DCFM_NEW=8", "-DCFM_PAIR=1"]) for bn in TC2]
    return out


def find_nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: libcfm_b200.so cannot be built")
    return nvcc


def _newest_dep(src: str) -> float:
    deps = [os.path.join(CSRC, src)] + [os.path.join(CSRC, f) for f in HEADERS] + [os.path.abspath(__file__)]
    return max(os.path.getmtime(d) for d in deps)


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(_newest_dep(src) > t for _, src, _ in units())


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return LIB
    nvcc = find_nvcc()
    os.makedirs(OBJ, exist_ok=True)

    def compile_one(unit):
        obj, src, defs = unit
        out = os.path.join(OBJ, obj)
        if not force and os.path.exists(out) and os.path.getmtime(out) >= _newest_dep(src):
            return ""
        cmd = [nvcc, *COMMON, *defs, "-c", os.path.join(CSRC, src), "-o", out]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        proc = subprocess.run(cmd, capture_output=True, text=True)
        if proc.returncode != 0:
            raise RuntimeError(f"nvcc failed for {obj}:\n" + proc.stdout + proc.stderr)
        return proc.stderr if verbose else ""

    with ThreadPoolExecutor(max_workers=max(1, min(len(units()), os.cpu_count() or 1))) as pool:
        logs = list(pool.map(compile_one, units()))
    if verbose:
        print("".join(logs))
    objs = [os.path.join(OBJ, o) for o, _, _ in units()]
    proc = subprocess.run([nvcc, "-shared", "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB, *objs],
                          capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("link failed:\n" + proc.stdout + proc.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
