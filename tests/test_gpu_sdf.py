"""Signed-distance scenes: the device evaluator / sphere tracer against the oracle's on identical inputs, and
whole images of the SDF example scene through every integrator that carries the program.  The reference has no SDF code — its
README lists an SDF example scene as open —, so what is checked here is that the two statements of the extension (CUDA, C++)
agree; the STRICT build must agree bit for bit (the program is add / mul / min / max / sqrt only)."""
import numpy as np
import pytest

from conftest import unit_vectors
from devfn import DeviceFns

pytestmark = pytest.mark.gpu


def pix_rel(a, b):
    a = a.reshape(-1, 4)[:, :3].astype(np.float64); b = b.reshape(-1, 4)[:, :3].astype(np.float64)
    return np.abs(a - b).max(1) / np.maximum(np.abs(b).max(1), 1e-3)


@pytest.fixture(scope="module")
def sdf_scene(rp):
    return rp.sdf_demo_scene()


@pytest.mark.parametrize("strict", [True, False])
def test_sdf_eval_and_trace_match_the_oracle(rp, po, sdf_scene, strict):
    export = sdf_scene.device_export()
    dev, osc = DeviceFns(rp, export, strict=strict), po.OracleScene(export)
    rng = np.random.default_rng(5)
    n = 40000
    q = rng.uniform(-2.5, 2.5, size=(3, n)).astype(np.float32)
    dd, dm = dev.sdf_eval(q)
    od, om = osc.sdf_eval(q)
    if strict:
        assert np.array_equal(dd, od) and np.array_equal(dm, om)
    else:
        assert np.abs(dd - od).max() <= 2e-6                        # FMA contraction, approximate sqrt: ~1 ulp of O(1) distances
        assert (dm != om).mean() < 1e-4                             # material ties at the seams of a union
    # rays from around the body towards it, and from inside it outwards (the transmissive ball)
    o = rng.uniform(-3.0, 3.0, size=(3, n)).astype(np.float32); o[1] = np.abs(o[1]) * 0.8
    aim = rng.uniform(-1.0, 1.0, size=(3, n)).astype(np.float32) * np.array([[1.4], [0.7], [0.9]], np.float32)
    d = aim - o
    d = (d / np.linalg.norm(d, axis=0, keepdims=True)).astype(np.float32)
    lim = np.where(rng.random(n) < 0.3, rng.uniform(0.5, 4.0, n), 1e30).astype(np.float32)
    a, b = dev.sdf_trace(o, d, lim), osc.sdf_trace(o, d, lim)
    if strict:
        for k in ("t", "normal", "material"):
            assert np.array_equal(a[k], b[k]), k
    else:
        both = (a["t"] >= 0) & (b["t"] >= 0)
        assert ((a["t"] >= 0) != (b["t"] >= 0)).mean() < 2e-3      # rays that graze the surface within hit_eps
        assert np.abs(a["t"][both] - b["t"][both]).max() < 5e-4     # (a hit may be declared one step earlier or later: |step| < hit_eps)
        assert np.quantile(np.abs(a["t"][both] - b["t"][both]), 0.99) < 2e-5
        assert np.quantile(np.linalg.norm(a["normal"][:, both] - b["normal"][:, both], axis=0), 0.99) < 2e-3
    assert (b["t"] >= 0).mean() > 0.3 and (b["t"] < 0).mean() > 0.1  # the sample exercises hits and misses
    assert len(np.unique(b["material"][b["t"] >= 0])) >= 3
    dev.close()


@pytest.mark.parametrize("integrator", ["auto", "fused", "wavefront"])
def test_sdf_scene_image_parity_strict(rp, po, sdf_scene, integrator):
    """closest_hit / any_hit with the signed-distance body inside the whole path loop: strict build vs oracle, shared counter RNG."""
    W, H, S = 160, 90, 2
    integ = {"auto": rp._abi.PTB_INTEGRATOR_AUTO, "fused": rp._abi.PTB_INTEGRATOR_FUSED, "wavefront": rp._abi.PTB_INTEGRATOR_WAVEFRONT}[integrator]
    pt = rp.Tracer.new(sdf_scene, strict=True, integrator=integ)
    buf = rp.ColorBuffer.new(W, H)
    pt.render_spp(buf, S)
    used = pt.integrator_used()
    pt.close()
    assert used == ("fused" if integrator == "fused" else "wavefront"), used        # generic wavefront kernel: no material table
    ref, _, _, _ = po.OracleScene(sdf_scene.device_export()).render(W, H, S)
    ok = np.isfinite(buf.pixels.reshape(-1, 4)).all(1) & np.isfinite(ref.reshape(-1, 4)).all(1)
    rel = pix_rel(buf.pixels, ref)[ok]
    print(f"[sdf image parity] {integrator}: within 1e-4: {(rel < 1e-4).mean():.5f}, bit-identical {(rel == 0).mean():.5f}")
    assert (~ok).sum() <= 2
    assert (rel < 1e-4).mean() >= 0.99
    assert np.all(buf.pixels.reshape(-1, 4)[:, 3] == 1.0)
    # the body is actually in the picture: the image differs from the same scene without the program
    export = sdf_scene.device_export()
    import copy
    bare = copy.copy(export); bare.sdf = None
    ref_bare, _, _, _ = po.OracleScene(bare).render(W, H, 1)
    assert (pix_rel(ref_bare, po.OracleScene(export).render(W, H, 1)[0]) > 1e-2).mean() > 0.1


def test_sdf_scene_shipped_build_and_f64(rp, po, sdf_scene):
    W, H, S = 160, 90, 2
    ref, _, _, _ = po.OracleScene(sdf_scene.device_export()).render(W, H, S)
    pt = rp.Tracer.new(sdf_scene)
    buf = rp.ColorBuffer.new(W, H)
    pt.render_spp(buf, S)
    pt.close()
    rel = pix_rel(buf.pixels, ref)
    print(f"[sdf image parity] shipped build: within 1e-4: {(rel < 1e-4).mean():.5f}")
    assert (rel < 1e-4).mean() >= 0.95                             # branchy sphere tracing + glass: more flipped paths than analytic scenes
    lum = lambda im: (im.reshape(-1, 4)[:, :3].astype(np.float64) @ np.array([0.212671, 0.715160, 0.072169])).mean()
    assert abs(lum(buf.pixels) / lum(ref) - 1) < 5e-3
    # the f64 instantiation of `F` (lib.rs:5-6) against the f64 oracle
    pt = rp.Tracer.new(sdf_scene, precision="f64")
    b64 = rp.ColorBuffer.new(W, H, precision="f64")
    pt.render_spp(b64, S)
    pt.close()
    ref64, _, _, _ = po.OracleScene(sdf_scene.device_export(), "f64").render(W, H, S)
    assert (pix_rel(b64.pixels, ref64) < 1e-7).mean() >= 0.995


def test_sdf_program_validation(rp):
    import ctypes as C
    sc = rp.sdf_demo_scene()
    pt = rp.Tracer.new(sc)
    T = rp._abi.TYPES["f32"]

    def try_prog(nodes, **kw):
        prog = rp.SdfProgram(nodes=nodes, **kw)
        sd, keep = prog.to_c("f32")
        return pt._lib.ptb_set_sdf_f32(pt._handle(), C.byref(sd))
    N = rp.SdfNode
    assert try_prog([N.sphere((0, 0, 0), 1.0, 1)]) == 0
    assert try_prog([N.sphere((0, 0, 0), 1.0, 1), N.union()]) == rp._abi.PTB_E_INVALID             # combinator without two operands
    assert try_prog([N.sphere((0, 0, 0), 1.0, 1), N.sphere((1, 0, 0), 1.0, 1)]) == rp._abi.PTB_E_INVALID   # two values left
    assert try_prog([N.sphere((0, 0, 0), 1.0, 99)]) == rp._abi.PTB_E_INVALID                        # material out of range
    assert try_prog([N.sphere((0, 0, 0), 1.0, 1)] * 17) == rp._abi.PTB_E_INVALID                    # too many nodes
    assert try_prog([N.sphere((0, 0, 0), 1.0, 1), N.sphere((1, 0, 0), 1.0, 1), N.smooth_union(0.0)]) == rp._abi.PTB_E_INVALID
    assert try_prog([N.sphere((0, 0, 0), 1.0, 1)], hit_eps=0.0) == rp._abi.PTB_E_INVALID
    assert try_prog([]) == 0                                                                        # removes the program
    pt.close()
