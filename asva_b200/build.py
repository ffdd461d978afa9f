"""Builds libasva_b200.so (all CUDA kernels + the C ABI) in-tree with nvcc for sm_100a.

The shared object is git-ignored but travels to the GPU box with the repo snapshot."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libasva_b200.so")
SOURCES = ["host_common.cu", "gemm_tc.cu", "attn_tc.cu", "attn_mma.cu", "norm.cu", "misc.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC", "--use_fast_math" if False else "-DASVA_NO_FAST_MATH",
]


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "asva_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, debug: bool = False) -> str:
    """debug=True builds lib/libasva_b200_dbg.so with -DASVA_DEBUG_SWITCHES (environment-driven plan overrides,
    work-skipping timing modes, in-kernel traces) next to the shipped library; select it with ASVA_LIB=<path>."""
    if debug:
        return _build(LIB.replace(".so", "_dbg.so"), ["-DASVA_DEBUG_SWITCHES"], "dbg_", verbose)
    if not force and not _stale():
        return LIB
    return _build(LIB, [], "", verbose)


def _build(lib_path: str, extra: list, obj_prefix: str, verbose: bool) -> str:
    os.makedirs(LIBDIR, exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(LIBDIR, obj_prefix + src.replace(".cu", ".o"))
        # ASVA_NVCC_EXTRA: extra nvcc flags for experiments (use with force=True / --force)
        cmd = [nvcc, *NVCC_FLAGS, *extra, *os.environ.get("ASVA_NVCC_EXTRA", "").split(), "-c",
               os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out}")
        if verbose and out:
            print(out)
    cmd = [nvcc, "-shared", "-o", lib_path, *objs, "-lcudart"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}")
    return lib_path


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, debug="--debug" in sys.argv))
