"""Builds the in-tree native libraries of spla_b200 with nvcc for sm_100a (no JIT cache, no pip).

    python -m spla_b200.build            # libsplacu.so (CUDA kernels + C ABI)
    python -m spla_b200.build --spla     # + the spla host framework with the src/cuda plug-in (needs /root/reference)

nvcc cross-compiles without a GPU, so this also is the "does it build" check on the CPU box.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libsplacu.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden",
    "-diag-suppress", "186",
]
SOURCES = ["runtime.cu", "vector_ops.cu", "mxv_pull.cu", "mxv_seg.cu", "mxv_scat.cu", "vxm_push.cu", "jit.cu", "user_ops.cu", "dist.cu", "profile.cu", "ingest.cu"]


def nvcc_path():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build libsplacu.so (set NVCC=/path/to/nvcc)")


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_splacu(force=False, verbose=False):
    """Compile spla_b200/csrc/*.cu -> spla_b200/lib/libsplacu.so (one object per source, then link)."""
    os.makedirs(LIBDIR, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(HERE, "..", "include", "splacu.h"))
    nvcc = nvcc_path()
    objs = []
    procs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(LIBDIR, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + headers):
            cmd = [nvcc] + NVCC_FLAGS + ["-c", s, "-o", o]
            if verbose:
                print(" ".join(cmd))
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = []
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed.append((src, out))
        elif verbose and out.strip():
            print(out)
    if failed:
        raise RuntimeError("nvcc failed:\n" + "\n".join(f"== {s} ==\n{o}" for s, o in failed))
    if force or procs or _stale(LIB, objs):
        cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-ldl"]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    build_splacu(force="--force" in sys.argv, verbose=True)
    if "--spla" in sys.argv:
        from spla_b200.integration import build_spla_cuda

        build_spla_cuda(verbose=True)
    print(LIB)
