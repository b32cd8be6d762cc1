"""pytest configuration: `gpu` marker, in-tree build of the oracle and of the host harness of the device math."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def build_host_harness():
    out = os.path.join(ROOT, "tests", "libpvb_host_harness.so")
    src = os.path.join(ROOT, "tests", "host_harness.cpp")
    deps = [src] + [os.path.join(ROOT, "panovlm_b200", "csrc", f) for f in ("pvb_math.cuh", "pvb_knn.cuh", "pvb_host.hpp")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-march=x86-64-v3", "-ffp-contract=off", "-fopenmp", "-fPIC", "-std=c++17", "-x", "c++", "-shared", "-o", out, src])
    return out


def build_adapter_harness():
    """tests/adapter_harness.cpp: the Ceres bridge of include/ driven through the ceres stand-in of oracle/shim, linked against the product library."""
    out = os.path.join(ROOT, "tests", "libpvb_adapter_harness.so")
    src = os.path.join(ROOT, "tests", "adapter_harness.cpp")
    deps = [src, os.path.join(ROOT, "include", "panovlm_b200_ceres_adapter.hpp"), os.path.join(ROOT, "include", "panovlm_b200.h"), os.path.join(ROOT, "include", "panovlm_b200_reduced.hpp"),
            os.path.join(ROOT, "oracle", "shim", "pvo_shim_ceres.hpp"), os.path.join(ROOT, "oracle", "shim", "pvo_shim_eigen.hpp")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O1", "-fPIC", "-shared", "-I", os.path.join(ROOT, "oracle", "shim"), "-I", os.path.join(ROOT, "include"),
                               "-o", out, src, "-L", os.path.join(ROOT, "panovlm_b200"), "-lpanovlm_b200", "-Wl,-rpath,$ORIGIN/../panovlm_b200"])
    return out


def build_nccl_harness():
    """tests/nccl_harness.cpp: a C++ multi-GPU caller (one thread + one context per GPU, ncclCommInitAll, the NCCL hook of libpanovlm_b200_nccl.so)."""
    out = os.path.join(ROOT, "tests", "libpvb_nccl_harness.so")
    src = os.path.join(ROOT, "tests", "nccl_harness.cpp")
    deps = [src, os.path.join(ROOT, "include", "panovlm_b200_nccl.h"), os.path.join(ROOT, "include", "panovlm_b200.h")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O1", "-fPIC", "-shared", "-pthread", "-I", os.path.join(ROOT, "include"), "-I", "/usr/local/cuda/include", "-o", out, src,
                               "-L", os.path.join(ROOT, "panovlm_b200"), "-lpanovlm_b200_nccl", "-lpanovlm_b200", "-lnccl", "-L", "/usr/local/cuda/lib64", "-lcudart",
                               "-Wl,-rpath,$ORIGIN/../panovlm_b200"])
    return out


@pytest.fixture(scope="session")
def oracle():
    from oracle import pvo
    pvo.build()
    pvo.lib()
    return pvo


@pytest.fixture(scope="session")
def harness():
    import ctypes
    return ctypes.CDLL(build_host_harness())


@pytest.fixture(scope="session")
def gpu_ctx():
    import panovlm_b200
    ctx = panovlm_b200.Context(0)   # raises loudly if the extension or the GPU is missing
    yield ctx
    ctx.close()
