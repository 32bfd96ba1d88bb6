"""The compiled-language host layer (include/vdf.hpp) mirrors the crate API in C++ because the reference is compiled
Rust and no Rust toolchain exists here.  tests/cpp/test_find_all.cpp re-expresses the reference's search tests
against it.  CPU: it must compile and link against libvdf_b200.so, and report a clean error without a GPU.
GPU: all of its tests pass."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "vid_dup_finder_lib_b200")
EXE = os.path.join(ROOT, "tests", "cpp", "_build", "test_find_all")


def _build():
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    src = os.path.join(ROOT, "tests", "cpp", "test_find_all.cpp")
    deps = [src, os.path.join(ROOT, "include", "vdf.hpp"), os.path.join(ROOT, "include", "vdf_b200.h"),
            os.path.join(PKG, "libvdf_b200.so")]
    if os.path.exists(EXE) and all(os.path.getmtime(d) <= os.path.getmtime(EXE) for d in deps):
        return
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-I", os.path.join(ROOT, "include"), src, "-L", PKG,
                           "-lvdf_b200", "-lpthread", f"-Wl,-rpath,{PKG}", "-o", EXE])


def test_cpp_host_layer_compiles_and_fails_loudly_without_gpu():
    _build()
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu test")
    r = subprocess.run([EXE], capture_output=True, text=True, timeout=120)
    assert r.returncode == 2 and "vdf_ctx_create failed" in r.stdout  # DeviceError, no CPU fallback


def test_cpp_host_layer_sort_order_without_gpu():
    """Search::sort through the C ABI (vdf_sort_order) against std::stable_sort with the header's Rust-Path comparator"""
    _build()
    r = subprocess.run([EXE, "--host-only"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
def test_cpp_host_layer_reference_tests():
    _build()
    r = subprocess.run([EXE], capture_output=True, text=True, timeout=600)
    print(r.stdout)
    assert r.returncode == 0 and "ALL TESTS PASSED" in r.stdout, r.stdout + r.stderr
