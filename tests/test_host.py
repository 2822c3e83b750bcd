"""CPU tests of the host mirror of the reference's prelude (rust_pathtracer_b200/prelude.py)."""
import ctypes as C

import numpy as np
import pytest


def test_prelude_names(rp):
    # rust-pathtracer/src/lib.rs:24-48
    for name in ("I", "F", "F3", "Camera3D", "Pinhole", "AnalyticalLight", "Material", "Light", "ColorBuffer", "Scene", "Tracer"):
        assert hasattr(rp, name), name
    assert rp.F == "f32"                                    # lib.rs:6 at this commit (README says f64)


def test_material_defaults_and_masks(rp):
    m = rp.Material.new()                                   # material.rs:82-114
    assert tuple(m.rgb) == (1.5, 1.5, 1.5) and m.roughness == 0.5 and m.ior == 1.45 and m.metallic == 0.0
    assert m.set_mask == rp._abi.PTB_MAT_ALL
    a = rp.Material.assigning(rgb=(1, 1, 1), roughness=0.05, metallic=1.0)
    assert a.set_mask == (rp._abi.PTB_MAT_RGB | rp._abi.PTB_MAT_ROUGHNESS | rp._abi.PTB_MAT_METALLIC)
    c = rp.Material.assigning(albedo_kind=rp._abi.PTB_ALBEDO_CHECKER_DIR_RATIO, roughness=1.0)
    assert c.set_mask & rp._abi.PTB_MAT_RGB


def test_pinhole_and_light(rp):
    p = rp.Pinhole.new()                                    # pinhole.rs:14-25
    assert tuple(p.origin) == (0, 0, 3) and tuple(p.center) == (0, 0, 0) and p.fov == 80.0
    p.set((1, 2, 3), (0, 0, -1)); p.set_fov(60)
    assert tuple(p.origin) == (1, 2, 3) and p.fov == 60.0
    l = rp.AnalyticalLight.spherical(rp.F3(3, 2, 2), 1.0, rp.F3(3, 3, 3))       # light.rs:13-28
    assert abs(l.light.area - 4 * np.pi) < 1e-12 and l.light.light_type == rp._abi.PTB_LIGHT_SPHERICAL


def test_demo_scene_export(rp, demo_export):
    e = demo_export
    assert len(e.spheres) == 2 and len(e.planes) == 1 and len(e.materials) == 3 and len(e.lights) == 1
    assert tuple(e.spheres[0].center) == (-1.1, 0, 0) and tuple(e.spheres[1].center) == (1.1, 0, 0)
    assert e.depth == 4 and e.eps == 0.005 and e.flags & rp._abi.PTB_SCENE_ANYHIT_IGNORES_MAX_DIST
    sc, keep = e.to_c("f32")
    assert sc.n_spheres == 2 and abs(sc.spheres[1].center[0] - 1.1) < 1e-7 and sc.materials[2].albedo_kind == 1
    assert abs(sc.camera.fov - 80.0) < 1e-6 and abs(sc.background.gamma - 2.2) < 1e-6
    sc64, _ = e.to_c("f64")
    assert sc64.materials[1].rgb[1] == 0.186 and sc64.eps == 0.005


def test_color_buffer(rp):
    b = rp.ColorBuffer.new(5, 3)                            # buffer.rs:18-32
    assert b.width == 5 and b.height == 3 and b.frames == 0 and b.pixels.shape == (60,) and b.pixels.dtype == np.float32
    b.pixels[(2 * 5 + 4) * 4 + 1] = 7.0
    assert b.at(4, 2)[1] == 7.0
    assert rp.ColorBuffer.new(2, 2, "f64").pixels.dtype == np.float64
    with pytest.raises(RuntimeError):
        b.convert_to_u8(np.zeros(60, np.uint8))             # device-only: needs a tracer


def test_scene_without_export_is_rejected(rp):
    class HostOnly(rp.Scene):
        pass
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        rp.Tracer.new(HostOnly())


def test_synthetic_scenes(rp):
    s = rp.sphere_field_scene(n_spheres=300, n_lights_side=2).device_export()
    assert len(s.spheres) == 300 and len(s.lights) == 4 and len(s.materials) == 301
    assert all(m.set_mask == rp._abi.PTB_MAT_ALL for m in s.materials)
    # A.5: no material mixes 0<metallic<1 with transmission
    assert not any(0 < m.metallic < 1 and m.spec_trans > 0 for m in s.materials)
    s2 = rp.sphere_field_scene(n_spheres=300, n_lights_side=2).device_export()
    assert [tuple(a.center) for a in s.spheres] == [tuple(a.center) for a in s2.spheres]     # seeded
    d = rp.divergence_stress_scene(side=8, depth=16).device_export()
    assert len(d.spheres) == 64 and d.depth == 16
