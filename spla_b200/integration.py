"""Builds spla WITH the CUDA backend of this repository registered beside its CPU backend.

    python -m spla_b200.integration        # -> spla_b200/lib/libspla_cuda_x64.so + test / example binaries

spla's host framework (public API, objects, schedule, registry, storage manager, CPU backend) is compiled BY PATH from the
reference checkout -- nothing of it is copied into this repository. The backend itself is new code: spla_b200/src/cuda/*
(C++ plug-in classes, only the C ABI of include/splacu.h is visible to them) on top of spla_b200/lib/libsplacu.so (CUDA
kernels). The one reference file that needs edits to know about a new accelerator is src/library.cpp; INTEGRATION.md lists
those edits as the diff a maintainer would commit. Here they are applied on the fly to a scratch copy under build/ (git-
ignored) by `patched_library_cpp`, which fails loudly if an anchor no longer matches the reference.

On the GPU box /root/reference does not exist: the prebuilt binaries under spla_b200/lib travel with the repository snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("SPLA_REFERENCE", "/root/reference")
BUILD = os.path.join(ROOT, "build", "spla_cuda")
LIB = os.path.join(HERE, "lib", "libspla_cuda_x64.so")

# (anchor in reference src/library.cpp, text inserted AFTER the anchor)
LIBRARY_CPP_EDITS = [
    ("#include <cpu/cpu_algo_registry.hpp>\n",
     "\n#if defined(SPLA_BUILD_CUDA)\n    #include <cuda/cuda_accelerator.hpp>\n    #include <cuda/cuda_algo_registry.hpp>\n#endif\n"),
    ("        register_algo_cpu(m_registry.get());\n",
     "\n#ifdef SPLA_BUILD_CUDA\n        // Register cuda algo version\n        register_algo_cuda(m_registry.get());\n#endif\n"),
    ("    Status Library::set_accelerator(AcceleratorType accelerator) {\n",
     "#if defined(SPLA_BUILD_CUDA)\n"
     "        if (accelerator == ACCELERATOR_TYPE_CUDA) {\n"
     "            m_accelerator = std::make_unique<CudaAccelerator>();\n\n"
     "            if (m_accelerator->init() != Status::Ok) {\n"
     "                m_accelerator.reset();\n"
     "                return Status::NoAcceleration;\n"
     "            }\n\n"
     "            return Status::Ok;\n"
     "        }\n"
     "#endif\n"),
    ("            g_library->set_accelerator(AcceleratorType::OpenCL);\n#endif\n",
     "\n#ifdef SPLA_BUILD_CUDA\n"
     "            // On init we by default attempt to set up the CUDA runtime\n"
     "            // If no device is present the library stays CPU-only, like the OpenCL path above\n"
     "            g_library->set_accelerator(ACCELERATOR_TYPE_CUDA);\n"
     "#endif\n"),
]


def reference_available():
    return os.path.exists(os.path.join(REF, "src", "library.cpp"))


def patched_library_cpp():
    """reference src/library.cpp + the CUDA registration edits -> build/spla_cuda/library.cpp (scratch, git-ignored)"""
    src = open(os.path.join(REF, "src", "library.cpp")).read()
    for anchor, insert in LIBRARY_CPP_EDITS:
        if src.count(anchor) != 1:
            raise RuntimeError(f"integration: anchor not found exactly once in reference src/library.cpp: {anchor!r}")
        src = src.replace(anchor, anchor + insert)
    os.makedirs(BUILD, exist_ok=True)
    out = os.path.join(BUILD, "library.cpp")
    if not os.path.exists(out) or open(out).read() != src:
        open(out, "w").write(src)
    return out


def build_spla_cuda(verbose=False):
    """No-op (prebuilt binaries are used) when the reference checkout is absent."""
    if not reference_available():
        if verbose:
            print(f"integration: {REF} not present -- using prebuilt {LIB} (if any)")
        return LIB if os.path.exists(LIB) else None
    from . import build as b

    b.build_splacu(verbose=verbose)
    patched_library_cpp()
    cmd = ["make", "-j8", "-C", os.path.join(HERE, "src"), f"REF={REF}", f"BUILD={BUILD}"]
    if not verbose:
        cmd.insert(1, "-s")
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build_spla_cuda(verbose="-q" not in sys.argv))
