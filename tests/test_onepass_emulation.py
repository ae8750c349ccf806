"""The one-pass dense GKL kernels (krylovkit.jl_b200/csrc/onepass_kernels.cuh) were written without GPU time, so their
SOURCE is executed here on host threads: tests/emu/onepass_emu.cpp includes the same header nvcc compiles, over
tests/emu/cuda_emu.h (one std::thread per CUDA thread, barriers for __syncthreads, slot exchange for the warp
shuffles, heap buffers for global and dynamic shared memory), with the launch geometries and shared-memory sizes of
the host code in spmv.cu, and checks y = A x, z = A'(A x) against double loops.

  * under AddressSanitizer: every global and shared-memory index of every thread is in range (11 shapes: both
    variants, Float32 / Float64, NZ = 1, 2, 4, 7, ld = 32 mod 64, fewer tiles than CTAs ...);
  * under ThreadSanitizer: no two threads touch a shared-memory word without a barrier between them — removing the
    barrier at the end of the tile loop makes this run fail (tried when the test was written).

What the emulation cannot see: the inline-PTX streaming loads (replaced by plain loads), performance, and anything
that depends on real warp scheduling."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU = os.path.join(ROOT, "tests", "emu")
CSRC = os.path.join(ROOT, "krylovkit.jl_b200", "csrc")


def _build(tmp_path, sanitizer):
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("no g++")
    exe = str(tmp_path / f"onepass_emu_{sanitizer}")
    cmd = [gxx, "-std=c++17", "-O1", "-g", f"-fsanitize={sanitizer}", "-fno-omit-frame-pointer", "-Wno-unknown-pragmas",
           "-I", EMU, "-I", CSRC, os.path.join(EMU, "onepass_emu.cpp"), "-o", exe, "-lpthread"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0 and "sanitize" in res.stderr and "cannot find" in res.stderr:
        pytest.skip(f"lib{sanitizer[0]}san not installed")
    assert res.returncode == 0, res.stderr[-3000:]
    return exe


def _skip_if_the_sanitizer_cannot_start(res):
    """a sanitizer runtime that cannot set up its shadow memory on this kernel / address-space layout says so before main()
    runs: that is the host's business, not the kernels'"""
    for msg in ("FATAL: ThreadSanitizer", "unexpected memory mapping", "Shadow memory range interleaves",
                "ReserveShadowMemoryRange failed", "failed to allocate", "Resource temporarily unavailable"):   # (pids limit)
        if msg in res.stderr and "ok m=" not in res.stdout:
            pytest.skip("sanitizer runtime cannot start here: " + msg)


def test_kernel_sources_run_clean_under_address_sanitizer(tmp_path):
    exe = _build(tmp_path, "address")
    res = subprocess.run([exe], capture_output=True, text=True, timeout=900)
    _skip_if_the_sanitizer_cannot_start(res)
    assert res.returncode == 0 and "all ok" in res.stdout, res.stdout[-2000:] + res.stderr[-3000:]
    assert res.stdout.count("ok m=") == 11 and "ERROR: AddressSanitizer" not in res.stderr


def test_kernel_sources_have_no_shared_memory_race(tmp_path):
    exe = _build(tmp_path, "thread")
    res = subprocess.run([exe, "1"], capture_output=True, text=True, timeout=900)
    _skip_if_the_sanitizer_cannot_start(res)
    assert res.returncode == 0 and "all ok" in res.stdout, res.stdout[-2000:] + res.stderr[-3000:]
    assert "ThreadSanitizer" not in res.stderr
