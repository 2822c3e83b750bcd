"""Builds tests/cpp/host_mirror_test.cpp against include/ptb200.hpp + libptb200.so and runs it:
the compiled-language host mirror of the reference's prelude (the reference is Rust; no Rust
toolchain exists in the image, so the host side above the C ABI is C++17)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "host_mirror_test")


def build():
    src = os.path.join(ROOT, "tests", "cpp", "host_mirror_test.cpp")
    deps = [src, os.path.join(ROOT, "include", "ptb200.hpp"), os.path.join(ROOT, "include", "ptb200.h"),
            os.path.join(ROOT, "oracle", "pt_oracle.hpp"), os.path.join(ROOT, "rust_pathtracer_b200", "libptb200.so")]
    if os.path.exists(EXE) and all(os.path.getmtime(d) <= os.path.getmtime(EXE) for d in deps):
        return
    libdir = os.path.join(ROOT, "rust_pathtracer_b200")
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fopenmp", "-I", os.path.join(ROOT, "include"), src,
                           "-L", libdir, "-lptb200", f"-Wl,-rpath,{libdir}", "-o", EXE])


def test_cpp_host_mirror_cpu():
    build()
    out = subprocess.run([EXE, "--cpu"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "cpu checks passed" in out.stdout


@pytest.mark.gpu
def test_cpp_host_mirror_gpu():
    build()
    out = subprocess.run([EXE, "--gpu"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "gpu checks passed" in out.stdout
