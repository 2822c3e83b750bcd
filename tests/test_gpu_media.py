"""Media (material.rs:5-34, Readme.md:13 "Support of mediums / volumetric objects"): the reference declares
Medium / MediumType and carries them in Material and State but its tracer never reads them.  The library implements them as
an extension (PTB_MEDIUM_* in include/ptb200.h); the oracle states the same semantics on the CPU (Tracer::medium_step,
pinned by closed forms in tests/test_oracle.py), and these tests hold the device to the oracle."""
import copy

import numpy as np
import pytest

from test_gpu_image import lum, pix_rel

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("precision", ["f32", "f64"])
def test_media_demo_scene_matches_the_oracle(rp, po, precision):
    """absorbing, scattering and emissive balls: image and event counters, both integrators that carry the medium state"""
    sc = rp.media_demo_scene(depth=8)
    e = sc.device_export()
    W, H, S = 160, 96, 4
    ref, _, _, oc = po.OracleScene(e, precision=precision).render(W, H, S, counters=True)
    plain = copy.deepcopy(sc.device_export())
    for m in plain.materials:
        m.medium = rp.Medium()
    ref_plain, _, _, _ = po.OracleScene(plain, precision=precision).render(W, H, S)
    assert (pix_rel(ref, ref_plain) > 1e-2).mean() > 0.05            # the media are visible
    for integ in (rp._abi.PTB_INTEGRATOR_WAVEFRONT, rp._abi.PTB_INTEGRATOR_FUSED, rp._abi.PTB_INTEGRATOR_AUTO):
        pt = rp.Tracer.new(sc, integrator=integ, collect_counters=True, precision=precision)
        buf = rp.ColorBuffer.new(W, H, precision=precision)
        pt.render_spp(buf, S)
        c = pt.counters()
        used = pt.integrator_used()
        pt.close()
        assert used.replace("_f64", "") in ("wavefront", "fused"), used   # never the resolved-material or the streaming kernels
        rel = pix_rel(buf.pixels, ref)
        bar = 1e-4 if precision == "f32" else 1e-9
        assert (rel < bar).mean() >= 0.99, (integ, (rel < bar).mean())
        assert abs(lum(buf.pixels).mean() / lum(ref).mean() - 1) < (1e-4 if precision == "f32" else 1e-9)
        for k in ("closest_hit", "any_hit", "shade", "nee_contrib", "end_sky", "end_emitter", "end_depth", "end_pdf"):
            assert abs(c[k] - oc[k]) <= max(5, 1e-3 * W * H * S), (integ, k, c[k], oc[k])


@pytest.mark.parametrize("kind", ["ABSORB", "SCATTER", "EMISSIVE"])
def test_single_medium_ball_closed_forms_on_the_device(rp, po, kind):
    """the closed forms of tests/test_oracle.py, on the device: Beer-Lambert, density x length, white furnace"""
    from test_oracle import _ball_in_white_sky
    den = {"ABSORB": 0.8, "SCATTER": 1.5, "EMISSIVE": 0.3}[kind]
    col = {"ABSORB": (0.9, 0.5, 0.2), "SCATTER": (1.0, 1.0, 1.0), "EMISSIVE": (1.0, 0.5, 0.25)}[kind]
    e = _ball_in_white_sky(rp, rp.Medium(getattr(rp.MediumType, kind), den, rp.F3(*col), 0.5), depth=200 if kind == "SCATTER" else 6)
    for integ in (rp._abi.PTB_INTEGRATOR_WAVEFRONT, rp._abi.PTB_INTEGRATOR_FUSED):
        pt = rp.Tracer.new(rp.ExportedScene(e), integrator=integ)
        buf = rp.ColorBuffer.new(65, 65)
        pt.render_spp(buf, 64)
        pt.close()
        centre = buf.pixels.reshape(65, 65, 4)[31:34, 31:34, :3].mean((0, 1))
        want = {"ABSORB": np.exp(-(1 - np.array(col)) * 2.0 * den), "SCATTER": np.ones(3), "EMISSIVE": 1.0 + np.array(col) * 2.0 * den}[kind]
        assert np.allclose(centre, want, rtol=0.02), (kind, integ, centre, want)


def test_media_are_refused_by_the_streaming_integrator_and_rerouted_by_auto(rp):
    sc = rp.media_demo_scene(depth=4)
    pt = rp.Tracer.new(sc, integrator=rp._abi.PTB_INTEGRATOR_STREAM)
    buf = rp.ColorBuffer.new(64, 48)
    with pytest.raises(Exception, match="media"):
        pt.render_spp(buf, 1)
    pt.close()
    # a BVH scene large enough for AUTO to pick the streaming integrator: with a medium it runs on the shared-memory wavefront
    big = rp.sphere_field_scene(n_spheres=5000, n_lights_side=2).device_export()
    big.materials[2].spec_trans = 1.0
    big.materials[2].medium = rp.Medium(rp.MediumType.ABSORB, 2.0, rp.F3(0.3, 0.6, 0.9), 0.0)
    pt = rp.Tracer.new(rp.ExportedScene(big))
    pt.render_spp(buf, 1)
    assert pt.integrator_used().startswith("wavefront")
    assert np.isfinite(buf.pixels).all()
    pt.close()
