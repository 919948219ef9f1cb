"""spla itself, with the CUDA backend of this repository registered beside its CPU backend (spla_b200/integration.py):
the reference's OWN gtest binaries for the path and a differential test that runs the same public-API calls on the CPU
backend and on the CUDA backend in one process (tests/cpp/test_cuda_backend.cpp).

The binaries are built where the reference checkout exists (this container) and travel to the GPU box with the snapshot.
"""
import ctypes
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "spla_b200", "lib")
LIB = os.path.join(LIBDIR, "libspla_cuda_x64.so")

needs_build = pytest.mark.skipif(not os.path.exists(LIB), reason="libspla_cuda_x64.so not built (python -m spla_b200.integration needs the reference checkout)")


def run(binary, *args, env=None, timeout=900):
    e = dict(os.environ)
    e.update(env or {})
    p = subprocess.run([os.path.join(LIBDIR, binary), *args], cwd=LIBDIR, env=e, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=timeout)
    return p.returncode, p.stdout


@needs_build
def test_library_exports_c_api_and_cuda_registration():
    """The C API of the reference (include/spla.h:372-373) is intact and the CUDA registration entry points are linked in."""
    lib = ctypes.CDLL(LIB)
    for sym in ("spla_Exec_mxv_masked", "spla_Exec_vxm_masked", "spla_Library_set_accelerator", "spla_Vector_make", "spla_Matrix_make"):
        assert hasattr(lib, sym), sym
    out = subprocess.run(["nm", "-DC", LIB], stdout=subprocess.PIPE, text=True).stdout
    for needle in ("spla::register_algo_cuda(spla::Registry*)", "spla::CudaAccelerator::init()", "spla::register_formats_cuda()"):
        assert needle in out, needle
    # the plug-in reaches the device only through the C ABI of include/splacu.h
    undefined = subprocess.run(["nm", "-D", "--undefined-only", LIB], stdout=subprocess.PIPE, text=True).stdout
    assert "splacu_mxv_masked" in undefined and "splacu_vxm_masked_begin" in undefined
    assert "cudaMalloc" not in undefined and "cudaLaunchKernel" not in undefined


@needs_build
def test_differential_harness_dry_run_on_cpu():
    """Without a device spla stays on its CPU backend (reference src/library.cpp:230-234 behaviour); the harness itself must be sound."""
    rc, out = run("test_cuda_backend", "9", env={"SPLA_TEST_DRY_RUN": "1", "CUDA_VISIBLE_DEVICES": ""})
    assert rc == 0, out
    assert "0 failed" in out


@pytest.mark.gpu
@needs_build
@pytest.mark.parametrize("scale", ["10", "13"])
def test_cuda_backend_matches_cpu_backend(scale):
    rc, out = run("test_cuda_backend", scale)
    assert rc == 0, out
    assert "accelerator: CUDA device" in out and "0 failed" in out, out


@pytest.mark.gpu
@needs_build
@pytest.mark.parametrize("binary", ["cuda_test_mxv", "cuda_test_vxm", "cuda_test_vector"])
def test_reference_gtests_on_cuda_backend(binary):
    """reference tests/test_mxv.cpp, tests/test_vxm.cpp, tests/test_vector.cpp, unmodified, linked to the CUDA-enabled library."""
    rc, out = run(binary)
    assert rc == 0, out[-3000:]
    assert "[  PASSED  ]" in out and "FAILED" not in out, out[-3000:]


def _gpu_count():
    try:
        import torch

        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.gpu
@needs_build
@pytest.mark.parametrize("shards", ["2", "3"])
def test_cuda_backend_sharded_matches_cpu_backend(shards):
    """The same differential test with the two products SHARDED behind spla's public API (CudaAccelerator device group: rows
    nnz-balanced for mxv, columns for vxm; SPLA_CUDA_DEVICES). On a single-GPU box the shards share the device
    (SPLA_CUDA_SHARE_DEVICES); with several GPUs visible they sit on distinct devices and v travels by ncclBroadcast."""
    env = {"SPLA_CUDA_DEVICES": shards}
    if _gpu_count() < int(shards):
        env["SPLA_CUDA_SHARE_DEVICES"] = "1"
    rc, out = run("test_cuda_backend", "12", env=env)
    assert rc == 0, out
    assert "more shard(s) for mxv / vxm" in out and "0 failed" in out, out


@pytest.mark.gpu
@needs_build
@pytest.mark.parametrize("binary", ["cuda_test_mxv", "cuda_test_vxm"])
def test_reference_gtests_on_sharded_cuda_backend(binary):
    env = {"SPLA_CUDA_DEVICES": "2"}
    if _gpu_count() < 2:
        env["SPLA_CUDA_SHARE_DEVICES"] = "1"
    rc, out = run(binary, env=env)
    assert rc == 0, out[-3000:]
    assert "[  PASSED  ]" in out and "FAILED" not in out, out[-3000:]
