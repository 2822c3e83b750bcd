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


def _resolved(rp, export, chain, odd=0):
    lib = rp._abi.load()
    sc, keep = export.to_c("f32")
    out = (C.c_float * rp._abi.PTB_RMAT_FLOATS)()
    ch = (C.c_uint32 * len(chain))(*chain)
    rp._abi.check(lib.ptb_test_resolved_material_f32(C.byref(sc), ch, len(chain), odd, out))
    v = np.array(out[:], np.float32)
    names = ["anisotropic", "metallic", "roughness", "subsurface", "specular_tint", "sheen", "sheen_tint", "clearcoat", "clearcoat_gloss",
             "spec_trans", "ior", "clearcoat_roughness", "ax", "ay", "eta_enter", "eta_exit"]
    d = dict(rgb=v[0:3], emission=v[3:6], spec_enter=v[22:25], spec_exit=v[25:28], sheen_col=v[28:31], lum=v[31], wd0=v[32], wc0=v[33],
             lobe_class=int(v[34]))
    d.update({n: v[6 + i] for i, n in enumerate(names)})
    return d


def test_resolved_material_table_is_bit_exact_with_the_oracle(rp, po):
    """The wavefront integrator shades small scenes from a table the HOST evaluates in ptb_set_scene_f32 (RMat).
    Its entries must be the oracle's own values bit for bit: Material::finalize (material.rs:117-131), eta (globals.rs:58-61),
    get_spec_color (tracer.rs:335-341) and the material-only lobe weights (tracer.rs:423, 426) — plain f32, same operation order.
    No device involved: this is the host half of the shade path."""
    r = np.random.default_rng(5)
    M = rp.Material
    mats = []
    for k in range(24):
        m = M(rgb=rp.F3(*r.uniform(0.0, 1.0, 3)), roughness=float(r.uniform(0.0, 1.0)), ior=float(r.uniform(1.05, 2.0)),
              metallic=float(r.choice([0.0, 1.0, r.uniform(0, 1)])), anisotropic=float(r.uniform(0, 1)), subsurface=float(r.uniform(0, 1)),
              specular_tint=float(r.uniform(0, 1)), sheen=float(r.uniform(0, 1)), sheen_tint=float(r.uniform(0, 1)),
              clearcoat=float(r.choice([0.0, r.uniform(0, 1)])), clearcoat_gloss=float(r.uniform(0, 1)),
              spec_trans=float(r.choice([0.0, r.uniform(0, 1)])), emission=rp.F3(*r.uniform(0, 2, 3)))
        mats.append(m)
    mats.append(M(rgb=rp.F3(0, 0, 0)))                                                      # zero luminance: ctint = 1 (tracer.rs:337)
    mats.append(M(roughness=1.0, albedo_kind=rp._abi.PTB_ALBEDO_CHECKER_DIR_RATIO))          # the demo floor
    e = rp.DeviceScene(spheres=[rp.Sphere(rp.F3(0, 0, 0), 1.0, i) for i in range(len(mats))], planes=[], materials=mats, lights=[],
                       camera=rp.Pinhole.new(), background=rp.Background(), depth=4, flags=rp._abi.PTB_SCENE_NO_BVH, eps=0.005)
    orc = po.OracleScene(e)
    o = np.zeros((3, 2), np.float32); o[2] = 3.0
    # two rays towards the sphere at the origin: from outside (entering) and from its centre (exiting)
    origins = np.array([[0, 0], [0, 0], [3, 0]], np.float32)
    d = np.array([[0, 0], [0, 0], [-1, -1]], np.float32)
    nrm = np.array([[0, 0], [0, 0], [1, 1]], np.float32)        # n.d < 0: entering; second ray: normal (0,0,-1)·d > 0 below
    nrm[2, 1] = -1.0
    for mi in range(len(mats)):
        for odd in ((0, 1) if mats[mi].albedo_kind else (0,)):
            got = _resolved(rp, e, [mi], odd)
            # a ray direction whose checker cell is even / odd for the direction-ratio checker (analytical.rs:107-113)
            direction = (0.2, -1.0, 0.3) if odd == 0 else (2.2, -1.0, 0.3)
            fin = orc.finalize(mi, origins, d, np.array([2.0, 1.0], np.float32), nrm)
            for k_o, k_g in (("roughness", "roughness"), ("clearcoat_roughness", "clearcoat_roughness"), ("ax", "ax"), ("ay", "ay")):
                assert fin[k_o][0] == got[k_g], (mi, k_o, fin[k_o][0], got[k_g])
            assert fin["eta"][0] == got["eta_enter"] and fin["eta"][1] == got["eta_exit"], (mi, fin["eta"], got["eta_enter"], got["eta_exit"])
            for side, eta in (("spec_enter", got["eta_enter"]), ("spec_exit", got["eta_exit"])):
                sc = orc.spec_color(mi, float(eta), direction)
                assert np.array_equal(sc["spec_col"], got[side]), (mi, side, sc["spec_col"], got[side])
                assert np.array_equal(sc["sheen_col"], got["sheen_col"])
                assert sc["lum"] == got["lum"] and sc["wd0"] == got["wd0"] and sc["wc0"] == got["wc0"], (mi, sc, got)
            m = mats[mi]
            want_class = (1 if (1 - m.metallic) * (1 - m.spec_trans) > 0 else 0) | (2 if m.clearcoat * (1 - m.metallic) > 0 else 0) | \
                         (4 if m.spec_trans * (1 - m.metallic) > 0 else 0)
            assert got["lobe_class"] == want_class
    # checker cells: even -> checker_a, odd -> checker_b
    fl = len(mats) - 1
    assert np.all(_resolved(rp, e, [fl], 0)["rgb"] == np.float32(0.25)) and np.all(_resolved(rp, e, [fl], 1)["rgb"] == np.float32(0.1))


def test_resolved_material_chain_replays_the_cumulative_assignment(rp, demo_export):
    """Partial-mask materials (the reference's closest_hit assigns a few fields per primitive that is closest SO FAR,
    analytical.rs:56-58, 82-85, 115-116): the table entry of an accepted chain is the first material's resolved values
    patched by the later ones' masked fields — e.g. the orange sphere seen in front of the metal one keeps metallic = 1."""
    e = demo_export
    metal, orange, floor = e.spheres[0].material, e.spheres[1].material, e.planes[0].material
    alone = _resolved(rp, e, [orange])
    assert alone["metallic"] == 0.0 and alone["clearcoat"] == 1.0 and alone["lobe_class"] == 3
    both = _resolved(rp, e, [metal, orange])
    assert both["metallic"] == 1.0 and both["clearcoat"] == 1.0 and both["lobe_class"] == 0        # (1 - metallic) kills diffuse and clearcoat
    assert np.allclose(both["rgb"], [1.0, 0.186, 0.0]) and both["roughness"] == np.float32(0.1)
    assert both["clearcoat_roughness"] == np.float32(np.float32(1 - 1.0) * np.float32(0.1) + np.float32(0.001) * np.float32(1.0))
    # the floor behind both spheres: checker albedo from the LAST accepted primitive, roughness 1, metallic / clearcoat left over
    chain = _resolved(rp, e, [metal, orange, floor], 1)
    assert np.all(chain["rgb"] == np.float32(0.1)) and chain["roughness"] == 1.0 and chain["metallic"] == 1.0 and chain["clearcoat"] == 1.0
    only_floor = _resolved(rp, e, [floor], 0)
    assert np.all(only_floor["rgb"] == np.float32(0.25)) and only_floor["metallic"] == 0.0 and only_floor["lobe_class"] == 1
    assert only_floor["eta_enter"] == np.float32(1.0) / np.float32(1.45) and only_floor["eta_exit"] == np.float32(1.45)


def test_film_quotients_by_fma_are_exact_or_declined(rp):
    """Host half of the division-free film coordinates (wavefront integrator): the verdict `all_exact` must be true exactly when
    no column / row quotient differs from the IEEE one, and the common frame sizes take the fast path."""
    lib = rp._abi.load()
    ok, bad = C.c_uint32(), C.c_uint32()
    sizes = [(800, 600), (960, 540), (1280, 720), (1920, 1080), (3840, 2160), (7680, 4320), (1, 1), (37, 19), (97, 61), (4099, 3001),
             (12345, 6789), (65535, 3)]
    n_fast = 0
    for (w, h) in sizes:
        rp._abi.check(lib.ptb_test_film_quotients_f32(w, h, C.byref(ok), C.byref(bad)))
        assert (ok.value == 1) == (bad.value == 0), (w, h, ok.value, bad.value)
        n_fast += ok.value
        # independent check of the claim in numpy: q = RN(x * RN(1/W)); r = x - q*W exactly; q' = RN(q + r * RN(1/W))
        x = np.arange(w, dtype=np.float32)
        rw = np.float32(1.0) / np.float32(w)
        q = x * rw
        r = (x.astype(np.float64) - q.astype(np.float64) * np.float64(w)).astype(np.float32)          # exact: |r| is tiny and representable
        q2 = (q.astype(np.float64) + r.astype(np.float64) * np.float64(rw)).astype(np.float32)        # one rounding (the f64 sum is exact enough)
        mism = int((q2 != x / np.float32(w)).sum())
        if ok.value:
            assert mism == 0, (w, h, mism)
    for (w, h) in [(800, 600), (1920, 1080), (3840, 2160)]:
        rp._abi.check(lib.ptb_test_film_quotients_f32(w, h, C.byref(ok), C.byref(bad)))
        assert ok.value == 1, (w, h, bad.value)
    assert n_fast >= len(sizes) - 2
    rp._abi.check(lib.ptb_test_film_quotients_f32(1 << 24, 10, C.byref(ok), C.byref(bad)))
    assert ok.value == 0                                      # beyond the exactly representable integers: declined


def test_color_buffer_escape_tracking(rp):
    """`pixels` is a public field of the reference (buffer.rs:9): once the writable array has been handed out the buffer may
    have been edited, so Tracer.render() must upload it; a buffer that never escaped may skip the upload."""
    b = rp.ColorBuffer.new(4, 4)
    assert not b._escaped
    v = b.read_pixels()
    assert not b._escaped and not v.flags.writeable and v.shape == (64,)
    assert b.at(1, 1) == [0.0, 0.0, 0.0, 0.0] and not b._escaped
    _ = b.pixels
    assert b._escaped
    b2 = rp.ColorBuffer.new(2, 2)
    b2.pixels = np.arange(16, dtype=np.float32)
    assert b2._escaped and b2.read_pixels()[5] == 5.0


def test_bvh_builder_bounds_depth_on_adversarial_inputs(rp):
    """binned SAH on strongly non-uniform sizes / positions builds long chains; the device traversal has a fixed
    40-entry stack, so the builder must switch to median splits before the depth can exceed it (never silently drop subtrees)."""
    lib = rp._abi.load()
    S = rp._abi.TYPES["f32"]["Sphere"]

    def build(centers, radii):
        n = len(radii)
        arr = (S * n)()
        for i in range(n):
            arr[i].center = (C.c_float * 3)(*[float(x) for x in centers[i]]); arr[i].radius = float(radii[i]); arr[i].material = 0
        d, nn, ml = C.c_uint32(), C.c_uint32(), C.c_uint32()
        rp._abi.check(lib.ptb_test_bvh_build_f32(arr, n, C.byref(d), C.byref(nn), C.byref(ml)))
        return d.value, nn.value, ml.value

    rng = np.random.default_rng(5)
    # uniform field: shallow tree
    d, nn, ml = build(rng.uniform(-50, 50, (4000, 3)), rng.uniform(0.05, 0.5, 4000))
    assert d <= 24 and ml <= 7 and nn < 2 * 4000
    # geometric progression of positions AND sizes: each SAH split peels off one primitive (a chain of depth ~n)
    n = 600
    k = np.arange(n)
    pos = np.stack([1.07 ** k, np.zeros(n), np.zeros(n)], 1)
    d, nn, ml = build(pos, 0.3 * 1.07 ** k)
    assert d < 40 and ml <= 7, (d, ml)
    # nested concentric shells (all centroids coincide) + one far outlier
    cen = np.zeros((300, 3)); cen[-1] = (1e6, 0, 0)
    d, nn, ml = build(cen, np.linspace(0.1, 30, 300))
    assert d < 40 and ml <= 7, (d, ml)
    # exponentially clustered points on a line, tiny radii
    pos = np.stack([np.exp(rng.uniform(-20, 20, 5000)), np.zeros(5000), np.zeros(5000)], 1)
    d, nn, ml = build(pos, np.full(5000, 1e-3))
    assert d < 40 and ml <= 7, (d, ml)
