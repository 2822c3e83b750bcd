"""GPU parity at image level: the CUDA integrator (through the host mirror and the C ABI) against
the CPU oracle, plus the drop-in semantics of Tracer::render / ColorBuffer.

Tolerances (BASELINE.json north_star / SURVEY.md §7):
  * deterministic, shared counter RNG: >= 99 % of pixels within 1e-4 relative of the oracle
    (outliers are paths that took a different branch at a silhouette / CDF edge);
  * statistical, independent RNG streams: mean relative luminance error < 0.5 %, per-pixel RMSE
    <= 1.2 x the Monte Carlo bound sqrt(var_a/N_a + var_b/N_b) estimated from batch means.
"""
import numpy as np
import pytest

from devfn import DeviceFns

pytestmark = pytest.mark.gpu


def pix_rel(a, b):
    a = a.reshape(-1, 4)[:, :3].astype(np.float64); b = b.reshape(-1, 4)[:, :3].astype(np.float64)
    return np.abs(a - b).max(1) / np.maximum(np.abs(b).max(1), 1e-3)


def lum(img):
    p = img.reshape(-1, 4).astype(np.float64)
    return 0.212671 * p[:, 0] + 0.715160 * p[:, 1] + 0.072169 * p[:, 2]


@pytest.fixture(scope="module")
def scene(rp):
    return rp.AnalyticalScene.new()


@pytest.mark.parametrize("spp", [1, 4, 16])
def test_image_deterministic_parity(rp, scene, oracle_demo, spp):
    W, H = 200, 150
    buf = rp.ColorBuffer.new(W, H)
    pt = rp.Tracer.new(scene)
    for _ in range(spp):
        pt.render(buf)                                   # the drop-in call, 1 spp each, host pixels round trip
    ref, frames, _, _ = oracle_demo.render(W, H, spp)
    assert buf.frames == frames == spp
    rel = pix_rel(buf.pixels, ref)
    assert (rel < 1e-4).mean() >= 0.99, (rel < 1e-4).mean()
    assert np.median(rel) < 1e-6
    assert np.all(buf.pixels.reshape(-1, 4)[:, 3] == 1.0)
    assert abs(lum(buf.pixels).mean() / lum(ref).mean() - 1) < 1e-4
    pt.close()


def test_batch_render_equals_repeated_render(rp, scene):
    W, H, N = 128, 96, 8
    pt = rp.Tracer.new(scene)
    a = rp.ColorBuffer.new(W, H)
    for _ in range(N):
        pt.render(a)
    b = rp.ColorBuffer.new(W, H)
    pt.render_spp(b, N)
    assert a.frames == b.frames == N
    assert np.allclose(a.pixels, b.pixels, rtol=2e-6, atol=1e-7)     # same samples, different f32 summation order
    # bit-reproducible run to run
    c = rp.ColorBuffer.new(W, H)
    pt.render_spp(c, N)
    assert np.array_equal(b.pixels, c.pixels)
    pt.close()


def test_reset_and_resume_semantics(rp, scene, oracle_demo):
    W, H = 96, 64
    pt = rp.Tracer.new(scene)
    buf = rp.ColorBuffer.new(W, H)
    pt.render(buf); pt.render(buf); pt.render(buf)
    three = buf.pixels.copy()
    buf.frames = 0                                       # the reference's reset idiom
    pt.render(buf)
    one, _, _, _ = oracle_demo.render(W, H, 1)
    assert buf.frames == 1 and (pix_rel(buf.pixels, one) < 1e-4).mean() > 0.99
    # the app edits the public fields: a foreign image with frames = 1 is uploaded and averaged with sample #1
    buf.pixels[:] = 0.25; buf.frames = 1
    pt.render(buf)
    second, _, _, _ = oracle_demo.render(W, H, 1, sample_base=1)
    want = 0.5 * 0.25 + 0.5 * second.reshape(-1, 4)
    want[:, 3] = 1.0
    got = buf.pixels.reshape(-1, 4)
    assert buf.frames == 2 and (pix_rel(got, want.astype(np.float32)) < 1e-4).mean() > 0.99 and np.all(got[:, 3] == 1.0)
    # resume a checkpointed (pixels, frames) pair on a NEW tracer: identical to never having stopped
    pt2 = rp.Tracer.new(scene)
    chk = rp.ColorBuffer.new(W, H)
    pt2.render(chk); pt2.render(chk)
    saved = (chk.pixels.copy(), chk.frames)
    pt3 = rp.Tracer.new(scene)
    res = rp.ColorBuffer.new(W, H); res.pixels[:] = saved[0]; res.frames = saved[1]
    pt3.render(res)
    assert res.frames == 3 and np.allclose(res.pixels, three, rtol=1e-5, atol=1e-7)
    for t in (pt, pt2, pt3):
        t.close()


def test_sample_split_invariance(rp, scene):
    """the multi-GPU decomposition: disjoint sample ranges summed == one range"""
    W, H = 120, 80
    pt = rp.Tracer.new(scene)
    a = rp.ColorBuffer.new(W, H)
    pt.render_spp(a, 8)
    b = rp.ColorBuffer.new(W, H)
    pt._ensure_size(b); pt.clear()
    pt.render_samples(4, 4); pt.render_samples(3, 0); pt.render_samples(1, 3)      # out of order on purpose
    pt.download(b)
    assert np.allclose(a.pixels, b.pixels, rtol=2e-6, atol=1e-7)
    pt.close()


def test_statistical_parity_4096spp(rp, scene, po, demo_export):
    W, H, K, B = 160, 90, 8, 512                        # 8 batches x 512 spp = 4096 spp per side
    pt = rp.Tracer.new(scene, seed=0xC0FFEE)            # independent stream from the oracle's seed 0
    gb, ob = [], []
    osc = po.OracleScene(demo_export)
    for k in range(K):
        buf = rp.ColorBuffer.new(W, H)
        pt._ensure_size(buf); pt.clear(); pt.render_samples(B, k * B); pt.download(buf)
        gb.append(buf.pixels.reshape(-1, 4)[:, :3].astype(np.float64))
        ref, _, _, _ = osc.render(W, H, B, sample_base=k * B)
        ob.append(ref.reshape(-1, 4)[:, :3].astype(np.float64))
    gb, ob = np.stack(gb), np.stack(ob)
    # the reference never filters NaN radiance: an exactly grazing clearcoat hit gives 0/0 (tracer.rs:414-418) about once
    # per 6e8 samples and poisons the pixel; which sample does it hangs on the last bit, so compare the finite pixels
    finite = np.isfinite(gb).all((0, 2)) & np.isfinite(ob).all((0, 2))
    assert finite.mean() > 0.999
    gb, ob = gb[:, finite], ob[:, finite]
    gm, om = gb.mean(0), ob.mean(0)
    w = np.array([0.212671, 0.715160, 0.072169])
    lg, lo = gm @ w, om @ w
    rel_lum = abs(lg.mean() / lo.mean() - 1)
    assert rel_lum < 0.005, rel_lum                     # < 0.5 % mean relative luminance error
    var = gb.var(0, ddof=1) / K + ob.var(0, ddof=1) / K
    rmse = np.sqrt(((gm - om) ** 2).mean())
    bound = np.sqrt(var.mean())
    assert rmse <= 1.2 * bound, (rmse, bound)
    assert rmse >= 0.5 * bound                           # and the streams really are independent
    pt.close()


def test_counters_match_oracle(rp, scene, oracle_demo):
    W, H, S = 200, 150, 4
    pt = rp.Tracer.new(scene, collect_counters=True)
    buf = rp.ColorBuffer.new(W, H)
    pt.render_spp(buf, S)
    c = pt.counters()
    _, _, _, o = oracle_demo.render(W, H, S, counters=True)
    n = W * H * S
    assert c["samples"] == o["samples"] == n
    for k in ("closest_hit", "any_hit", "shade", "nee_contrib", "eval_calls", "lobe_diffuse", "lobe_clearcoat", "lobe_reflect",
              "lobe_refract", "end_sky", "end_emitter", "end_pdf", "end_depth"):
        assert abs(c[k] - o[k]) <= max(3, 2e-4 * n), (k, c[k], o[k])
    assert c["end_sky"] + c["end_emitter"] + c["end_pdf"] + c["end_depth"] == n
    # counting must not change the image
    pt2 = rp.Tracer.new(scene)
    b2 = rp.ColorBuffer.new(W, H)
    pt2.render_spp(b2, S)
    # (the counting kernel is a different template instantiation: same arithmetic, different FMA contraction)
    assert (pix_rel(buf.pixels, b2.pixels) < 1e-5).mean() > 0.995
    pt.close(); pt2.close()


@pytest.mark.parametrize("wh", [(1, 1), (37, 19), (16, 16), (17, 33)])
def test_odd_frame_sizes(rp, scene, oracle_demo, wh):
    W, H = wh
    pt = rp.Tracer.new(scene)
    buf = rp.ColorBuffer.new(W, H)
    pt.render_spp(buf, 3)
    ref, _, _, _ = oracle_demo.render(W, H, 3)
    assert np.all(buf.pixels.reshape(-1, 4)[:, 3] == 1.0)             # every pixel got its samples
    assert (pix_rel(buf.pixels, ref) < 1e-4).mean() >= 0.97
    pt.render_samples(0, 0)                                            # spp = 0 is a no-op
    pt.close()


def test_convert_to_u8_paths(rp, scene, po):
    W, H = 96, 64
    pt = rp.Tracer.new(scene)
    buf = rp.ColorBuffer.new(W, H)
    pt.render_spp(buf, 8)
    frame = np.zeros(W * H * 4, np.uint8)
    pt.convert_to_u8(frame)                                            # from the device-resident image
    ref = po.convert_to_u8(buf.pixels)
    d = np.abs(frame.astype(np.int32) - ref.astype(np.int32))
    assert d.max() <= 1 and (d != 0).mean() < 1e-3
    d = np.abs(buf.to_u8_vec().astype(np.int32) - ref.astype(np.int32))   # ColorBuffer method: uploads the host pixels
    assert d.max() <= 1 and (d != 0).mean() < 2e-3
    buf.pixels[:] = np.linspace(0, 1.2, buf.pixels.size, dtype=np.float32)   # host edit -> upload path
    frame2 = np.zeros(W * H * 4, np.uint8)
    buf.convert_to_u8(frame2)
    ref2 = po.convert_to_u8(buf.pixels)
    # upload multiplies by `frames` and the kernel divides again: +-1 LSB where that rounding crosses an integer
    d2 = np.abs(frame2.astype(np.int32) - ref2.astype(np.int32))
    assert d2.max() <= 1 and (d2 != 0).mean() < 2e-3
    # convert_to_u8_at: no gamma, strict bounds, rows counted from the bottom (buffer.rs:67-102)
    fw, fh = 128, 100
    big = np.full(fw * fh * 4, 9, np.uint8)
    buf.convert_to_u8_at(big, (10, 7, fw, fh))
    want = po.convert_to_u8_at(buf.pixels, W, H, np.full(fw * fh * 4, 9, np.uint8), 10, 7, fw, fh)
    d3 = np.abs(big.astype(np.int32) - want.astype(np.int32))
    assert d3.max() <= 1 and (d3 != 0).mean() < 2e-3 and (big != 9).sum() > 0
    pt.close()


def test_f64_switch(rp, scene, po, demo_export):
    """the F = f64 instantiation (lib.rs:6) against the f64 oracle"""
    W, H, S = 96, 64, 4
    pt = rp.Tracer.new(scene, precision="f64")
    buf = rp.ColorBuffer.new(W, H, "f64")
    for _ in range(S):
        pt.render(buf)
    ref, _, _, _ = po.OracleScene(demo_export, "f64").render(W, H, S)
    rel = pix_rel(buf.pixels, ref)
    assert (rel < 1e-9).mean() >= 0.99 and np.median(rel) < 1e-13
    # and f32 vs f64 agree statistically (different RNG grids, same distribution)
    pt32 = rp.Tracer.new(scene)
    b32 = rp.ColorBuffer.new(W, H)
    pt32.render_spp(b32, 256)
    b64 = rp.ColorBuffer.new(W, H, "f64")
    pt.render_spp(b64, 256)
    assert abs(lum(b32.pixels).mean() / lum(b64.pixels).mean() - 1) < 0.01
    with pytest.raises(rp._abi.PtbError):
        rp._abi.check(rp._abi.load().ptb_download_f32(pt._handle(), b32.pixels.ctypes.data))   # precision mismatch is an error
    pt.close(); pt32.close()


def _small_field(rp, n=1500, side=2):
    return rp.sphere_field_scene(n_spheres=n, n_lights_side=side)


def test_bvh_equals_bruteforce_and_oracle(rp, po):
    sc = _small_field(rp)
    export = sc.device_export()
    W, H, S = 96, 54, 2
    img = {}
    for name, flags in (("bvh", rp._abi.PTB_SCENE_FORCE_BVH), ("brute", rp._abi.PTB_SCENE_NO_BVH)):
        export.flags = flags
        pt = rp.Tracer.new(sc)
        buf = rp.ColorBuffer.new(W, H)
        pt.render_spp(buf, S)
        img[name] = buf.pixels.copy()
        pt.close()
    # same arithmetic per sphere and the same lowest-index tie rule; the two kernels are different template
    # instantiations, so FMA contraction may differ in the last bit
    assert (pix_rel(img["bvh"], img["brute"]) < 1e-5).mean() > 0.995
    assert abs(lum(img["bvh"]).mean() / lum(img["brute"]).mean() - 1) < 1e-5
    export.flags = 0
    ref, _, _, _ = po.OracleScene(export).render(W, H, S)              # linear scan on the CPU
    rel = pix_rel(img["bvh"], ref)
    # 1500 small spheres = many silhouettes, and the field's glass / high-gloss clearcoat lobes are the
    # ill-conditioned ones (see test_gpu_functions.py): more branch-flip outliers than the demo scene
    assert (rel < 1e-4).mean() >= 0.95, (rel < 1e-4).mean()
    assert abs(lum(img["bvh"]).mean() / lum(ref).mean() - 1) < 2e-3


def test_bvh_closest_hit_function_parity(rp, po):
    sc = _small_field(rp, 3000, 3)
    export = sc.device_export()
    dev = DeviceFns(rp, export)
    osc = po.OracleScene(export)
    rng = np.random.default_rng(77)
    n = 20000
    o = np.stack([rng.uniform(-60, 60, n), rng.uniform(0, 8, n), rng.uniform(-120, 10, n)]).astype(np.float32)
    d = rng.normal(size=(3, n)); d[1] = -np.abs(d[1]) * 0.3; d /= np.linalg.norm(d, axis=0); d = d.astype(np.float32)
    hd = np.full(n, -1.0, np.float32)
    ref, got = osc.closest_hit(o, d, hd), dev.closest_hit(o, d, hd)
    same = (ref["hit"] == got["hit"]) & (ref["material"] == got["material"]) & (ref["is_emitter"] == got["is_emitter"])
    assert (~same).sum() <= 5
    m = same & (ref["hit"] == 1)
    assert (ref["material"][m] > 0).sum() > 1000                       # plenty of sphere hits, not only the plane
    assert (np.abs(got["hit_dist"][m] - ref["hit_dist"][m]) / np.maximum(ref["hit_dist"][m], 1.0)).max() < 2e-5
    md = rng.uniform(0, 30, n).astype(np.float32)
    assert (osc.any_hit(o, d, md) != dev.any_hit(o, d, md)).sum() <= 5
    dev.close()


def test_russian_roulette_is_unbiased(rp):
    sc = rp.divergence_stress_scene(side=6, depth=16)
    W, H, S = 128, 72, 512
    means = []
    for rr in (0, 3):
        pt = rp.Tracer.new(sc, rr_start=rr, collect_counters=True)
        buf = rp.ColorBuffer.new(W, H)
        pt.render_spp(buf, S)
        means.append(lum(buf.pixels).mean())
        c = pt.counters()
        assert (c["end_rr"] > 0) == (rr > 0)
        pt.close()
    assert abs(means[0] / means[1] - 1) < 0.02, means


def test_degenerate_scenes(rp):
    # no primitives, no lights: every path is sky after one closest_hit
    e = rp.DeviceScene()
    pt = rp.Tracer.new(rp.ExportedScene(e), collect_counters=True)
    buf = rp.ColorBuffer.new(32, 16)
    pt.render_spp(buf, 2)
    c = pt.counters()
    assert c["closest_hit"] == c["samples"] == c["end_sky"] == 32 * 16 * 2 and c["any_hit"] == 0
    px = buf.pixels.reshape(-1, 4)
    assert np.all(px[:, 2] == 0.5) and np.all(px[:, 3] == 1.0)         # blue channel of the gradient sky is 1^2.2 * 0.5
    pt.close()
    # depth 1: exactly one closest_hit per sample
    e = rp.AnalyticalScene.new().device_export(); e.depth = 1
    pt = rp.Tracer.new(rp.ExportedScene(e), collect_counters=True)
    pt.render_spp(buf, 1)
    c = pt.counters()
    assert c["closest_hit"] == c["samples"]
    pt.close()
    # depth 0: the bounce loop never runs (tracer.rs:61) -> black image with alpha 1, no closest_hit at all
    e = rp.AnalyticalScene.new().device_export(); e.depth = 0
    for integ in (rp._abi.PTB_INTEGRATOR_FUSED, rp._abi.PTB_INTEGRATOR_WAVEFRONT, rp._abi.PTB_INTEGRATOR_STREAM):
        pt = rp.Tracer.new(rp.ExportedScene(e), collect_counters=True, integrator=integ)
        b0 = rp.ColorBuffer.new(32, 16)
        pt.render_spp(b0, 2)
        c = pt.counters()
        assert c["closest_hit"] == 0 and c["samples"] == 32 * 16 * 2
        px = b0.pixels.reshape(-1, 4)
        assert np.all(px[:, :3] == 0.0) and np.all(px[:, 3] == 1.0)
        pt.close()
    # lights but no geometry, and a light in front of the camera: stale hit_dist = -1 hides it (A.1) -> pure sky
    e = rp.DeviceScene(lights=[rp.AnalyticalLight.spherical((0, 0, 0), 1.0, (5, 5, 5))])
    pt = rp.Tracer.new(rp.ExportedScene(e), collect_counters=True)
    pt.render_spp(buf, 1)
    assert pt.counters()["end_emitter"] == 0
    pt.close()


@pytest.mark.parametrize("wh_spp", [(200, 150, 4), (97, 61, 3), (640, 360, 2)])
def test_wavefront_integrator_parity(rp, scene, oracle_demo, wh_spp):
    """the shared-memory wavefront integrator traces exactly the same paths as the fused one"""
    W, H, S = wh_spp
    img = {}
    for name, integ in (("fused", rp._abi.PTB_INTEGRATOR_FUSED), ("wave", rp._abi.PTB_INTEGRATOR_WAVEFRONT)):
        pt = rp.Tracer.new(scene, integrator=integ, collect_counters=True)
        buf = rp.ColorBuffer.new(W, H)
        pt.render_spp(buf, S)
        img[name] = (buf.pixels.copy(), pt.counters())
        pt.close()
    assert (pix_rel(img["wave"][0], img["fused"][0]) < 1e-5).mean() > 0.995
    assert np.all(img["wave"][0].reshape(-1, 4)[:, 3] == 1.0)
    cw, cf = img["wave"][1], img["fused"][1]
    assert cw["samples"] == cf["samples"] == W * H * S
    for k in cw:
        assert abs(cw[k] - cf[k]) <= max(3, 2e-4 * W * H * S), (k, cw[k], cf[k])
    ref, _, _, _ = oracle_demo.render(W, H, S)
    assert (pix_rel(img["wave"][0], ref) < 1e-4).mean() >= 0.99
    # bit-reproducible run to run
    pt = rp.Tracer.new(scene, integrator=rp._abi.PTB_INTEGRATOR_WAVEFRONT)
    b1 = rp.ColorBuffer.new(W, H); b2 = rp.ColorBuffer.new(W, H)
    pt.render_spp(b1, S); pt.render_spp(b2, S)
    assert np.array_equal(b1.pixels, b2.pixels)
    pt.close()


def test_wavefront_bvh_and_rr(rp):
    sc = _small_field(rp, 800, 2)
    W, H, S = 96, 54, 2
    img = {}
    for name, integ in (("fused", rp._abi.PTB_INTEGRATOR_FUSED), ("wave", rp._abi.PTB_INTEGRATOR_WAVEFRONT)):
        pt = rp.Tracer.new(sc, integrator=integ, bvh_threshold=1)
        buf = rp.ColorBuffer.new(W, H)
        pt.render_spp(buf, S)
        img[name] = buf.pixels.copy()
        pt.close()
    assert (pix_rel(img["wave"], img["fused"]) < 1e-5).mean() > 0.99
    sc = rp.divergence_stress_scene(side=6, depth=16)
    for name, integ in (("fused", rp._abi.PTB_INTEGRATOR_FUSED), ("wave", rp._abi.PTB_INTEGRATOR_WAVEFRONT)):
        pt = rp.Tracer.new(sc, integrator=integ, rr_start=3)
        buf = rp.ColorBuffer.new(W, H)
        pt.render_spp(buf, 8)
        img[name] = buf.pixels.copy()
        pt.close()
    assert (pix_rel(img["wave"], img["fused"]) < 1e-5).mean() > 0.99


@pytest.mark.parametrize("wh_spp_wave", [(200, 150, 4, 0), (97, 61, 3, 4096), (640, 360, 2, 100000), (64, 48, 5, 256)])
def test_stream_integrator_parity(rp, scene, oracle_demo, wh_spp_wave):
    """the global-memory wavefront (one kernel per stage over HBM queues) traces exactly the same paths as the fused
    integrator, whatever the wave shape: whole frame x several samples, pixel ranges x one sample, partial last waves"""
    W, H, S, wave = wh_spp_wave
    img = {}
    for name, integ in (("fused", rp._abi.PTB_INTEGRATOR_FUSED), ("stream", rp._abi.PTB_INTEGRATOR_STREAM)):
        pt = rp.Tracer.new(scene, integrator=integ, collect_counters=True, wave_paths=wave)
        buf = rp.ColorBuffer.new(W, H)
        pt.render_spp(buf, S)
        img[name] = (buf.pixels.copy(), pt.counters())
        pt.close()
    assert (pix_rel(img["stream"][0], img["fused"][0]) < 1e-5).mean() > 0.995
    assert np.all(img["stream"][0].reshape(-1, 4)[:, 3] == 1.0)
    cs, cf = img["stream"][1], img["fused"][1]
    assert cs["samples"] == cf["samples"] == W * H * S
    for k in cs:
        assert abs(cs[k] - cf[k]) <= max(3, 2e-4 * W * H * S), (k, cs[k], cf[k])
    ref, _, _, _ = oracle_demo.render(W, H, S)
    assert (pix_rel(img["stream"][0], ref) < 1e-4).mean() >= 0.99
    # bit-reproducible run to run although the queue order is not; the non-counting build prunes zero-pdf shadow rays
    pt = rp.Tracer.new(scene, integrator=rp._abi.PTB_INTEGRATOR_STREAM, wave_paths=wave)
    b1 = rp.ColorBuffer.new(W, H); b2 = rp.ColorBuffer.new(W, H)
    pt.render_spp(b1, S); pt.render_spp(b2, S)
    assert np.array_equal(b1.pixels, b2.pixels)
    assert (pix_rel(b1.pixels, img["stream"][0]) < 1e-5).mean() > 0.995
    pt.close()


def test_stream_bvh_and_rr(rp):
    sc = _small_field(rp, 800, 2)
    W, H, S = 96, 54, 2
    img = {}
    for name, integ in (("fused", rp._abi.PTB_INTEGRATOR_FUSED), ("stream", rp._abi.PTB_INTEGRATOR_STREAM)):
        pt = rp.Tracer.new(sc, integrator=integ, bvh_threshold=1, collect_counters=True)
        buf = rp.ColorBuffer.new(W, H)
        pt.render_spp(buf, S)
        img[name] = (buf.pixels.copy(), pt.counters())
        pt.close()
    assert (pix_rel(img["stream"][0], img["fused"][0]) < 1e-5).mean() > 0.99
    for k in img["fused"][1]:
        if k.startswith("bvh_"):        # traversal work: absolute tolerances below are per path event, not per node visit
            assert abs(img["stream"][1][k] - img["fused"][1][k]) <= 5e-3 * img["fused"][1][k], k
            continue
        assert abs(img["stream"][1][k] - img["fused"][1][k]) <= max(3, 1e-3 * W * H * S), k
    sc = rp.divergence_stress_scene(side=6, depth=16)
    for name, integ in (("fused", rp._abi.PTB_INTEGRATOR_FUSED), ("stream", rp._abi.PTB_INTEGRATOR_STREAM)):
        pt = rp.Tracer.new(sc, integrator=integ, rr_start=3)
        buf = rp.ColorBuffer.new(W, H)
        pt.render_spp(buf, 8)
        img[name] = buf.pixels.copy()
        pt.close()
    assert (pix_rel(img["stream"], img["fused"]) < 1e-5).mean() > 0.99
    # > 16384 spheres: from bounce 1 on the persistent-lane traversal kernels (k_stream_trace / k_stream_finish) take over
    sc = rp.sphere_field_scene(n_spheres=20000, n_lights_side=5)
    for name, integ in (("fused", rp._abi.PTB_INTEGRATOR_FUSED), ("stream", rp._abi.PTB_INTEGRATOR_STREAM)):
        pt = rp.Tracer.new(sc, integrator=integ, collect_counters=True)
        buf = rp.ColorBuffer.new(160, 90)
        pt.render_spp(buf, 3)
        img[name] = (buf.pixels.copy(), pt.counters())
        pt.close()
    # (the fused kernel of a scene without media / extended lights / SDF is a leaner instantiation than the streaming stages:
    #  FMA contraction differs in a last bit here and there, and this scene's glass / clearcoat paths amplify it — 98.97 % measured)
    assert (pix_rel(img["stream"][0], img["fused"][0]) < 1e-5).mean() > 0.98
    for k in img["fused"][1]:
        if k.startswith("bvh_"):        # the dedicated traversal kernels of large scenes visit the tree in another order
            assert 0.5 < img["stream"][1][k] / max(1, img["fused"][1][k]) < 2.0, k
            continue
        assert abs(img["stream"][1][k] - img["fused"][1][k]) <= max(3, 1e-3 * 160 * 90 * 3), k
    pt = rp.Tracer.new(sc, integrator=rp._abi.PTB_INTEGRATOR_STREAM)      # non-counting build: zero-pdf shadow rays pruned
    buf = rp.ColorBuffer.new(160, 90)
    pt.render_spp(buf, 3)
    pt.close()
    assert (pix_rel(buf.pixels, img["fused"][0]) < 1e-5).mean() > 0.99
    # 25 lights: light BVH inside k_stream_closest
    sc = rp.sphere_field_scene(n_spheres=300, n_lights_side=5)
    for name, integ in (("fused", rp._abi.PTB_INTEGRATOR_FUSED), ("stream", rp._abi.PTB_INTEGRATOR_STREAM)):
        pt = rp.Tracer.new(sc, integrator=integ)
        buf = rp.ColorBuffer.new(W, H)
        pt.render_spp(buf, 3)
        img[name] = buf.pixels.copy()
        pt.close()
    assert (pix_rel(img["stream"], img["fused"]) < 1e-5).mean() > 0.99


def test_light_bvh_matches_linear_scan(rp, po):
    """>= 16 spherical lights switch Scene::sample_lights (scene.rs:36-86) from the linear scan to a light BVH"""
    sc = rp.sphere_field_scene(n_spheres=200, n_lights_side=5)       # 25 lights
    export = sc.device_export()
    dev = DeviceFns(rp, export)
    osc = po.OracleScene(export)                                       # linear scan over the lights
    rng = np.random.default_rng(5)
    n = 40000
    o = np.stack([rng.uniform(-60, 60, n), rng.uniform(0, 8, n), rng.uniform(-120, 0, n)]).astype(np.float32)
    # aim most rays at a random light so that plenty of them hit one
    L = np.array([[*l.light.position] for l in export.lights], np.float64)
    tgt = L[rng.integers(0, len(L), n)].T + rng.normal(scale=0.35, size=(3, n))
    d = tgt - o
    d /= np.linalg.norm(d, axis=0)
    d = d.astype(np.float32)
    hd = rng.choice([-1.0, 5.0, 1e30], size=n, p=[0.1, 0.2, 0.7]).astype(np.float32)
    ref, got = osc.closest_hit(o, d, hd), dev.closest_hit(o, d, hd)
    same = (ref["hit"] == got["hit"]) & (ref["is_emitter"] == got["is_emitter"])
    assert (~same).sum() <= 5
    em = same & (ref["is_emitter"] == 1)
    assert em.sum() > 5000
    assert (np.abs(got["hit_dist"][em] - ref["hit_dist"][em]) / np.maximum(ref["hit_dist"][em], 1.0)).max() < 2e-5
    assert np.array_equal(got["light_emission"][:, em], ref["light_emission"][:, em])
    well = em & (ref["light_pdf"] < 1e6)
    assert np.percentile(np.abs(got["light_pdf"][well] / ref["light_pdf"][well] - 1), 99) < 1e-4
    assert np.array_equal(got["hit_dist"][same & (ref["hit"] == 0)], hd[same & (ref["hit"] == 0)])
    dev.close()
    # and at image level against the oracle
    W, H, S = 80, 45, 2
    pt = rp.Tracer.new(sc)
    buf = rp.ColorBuffer.new(W, H)
    pt.render_spp(buf, S)
    ref_img, _, _, _ = osc.render(W, H, S)
    assert (pix_rel(buf.pixels, ref_img) < 1e-4).mean() >= 0.95
    pt.close()


def _random_scene(rp, seed):
    """a random small scene that exercises the generic export: several spheres / planes / lights, full materials with
    emission, transmission, anisotropy, both background kinds, any_hit honouring max_dist, depth 1..6"""
    r = np.random.default_rng(seed)
    M = rp.Material
    mats = []
    for _ in range(int(r.integers(2, 7))):
        kind = r.integers(0, 5)
        m = M(rgb=rp.F3(*r.uniform(0.1, 1.0, 3)), roughness=float(r.uniform(0.15, 0.9)), ior=float(r.uniform(1.2, 1.7)))
        if kind == 0: m.metallic = 1.0; m.anisotropic = float(r.uniform(0, 0.8))
        elif kind == 1: m.spec_trans = float(r.uniform(0.5, 1.0))
        elif kind == 2: m.clearcoat = float(r.uniform(0.3, 1.0)); m.clearcoat_gloss = float(r.uniform(0.0, 0.6))
        elif kind == 3: m.sheen = float(r.uniform(0, 1)); m.sheen_tint = float(r.uniform(0, 1)); m.subsurface = float(r.uniform(0, 1)); m.specular_tint = float(r.uniform(0, 1))
        else: m.emission = rp.F3(*r.uniform(0.0, 2.0, 3))
        mats.append(m)
    floor = M(roughness=float(r.uniform(0.3, 1.0)))
    if r.random() < 0.5:
        floor.albedo_kind = rp._abi.PTB_ALBEDO_CHECKER_DIR_RATIO
    mats.append(floor)
    spheres = [rp.Sphere(rp.F3(*(r.uniform(-2.5, 2.5), r.uniform(-0.5, 1.5), r.uniform(-3.0, 0.5))), float(r.uniform(0.3, 1.0)),
                         int(r.integers(0, len(mats) - 1))) for _ in range(int(r.integers(1, 9)))]
    planes = [rp.Plane(rp.F3(0, -1, 0), rp.F3(0, 1, 0), len(mats) - 1)]
    if r.random() < 0.5:
        planes.append(rp.Plane(rp.F3(0, 0, -5), rp.F3(0, 0, 1), int(r.integers(0, len(mats)))))
    lights = [rp.AnalyticalLight.spherical(rp.F3(*(r.uniform(-4, 4), r.uniform(1.5, 5), r.uniform(-3, 3))), float(r.uniform(0.3, 1.2)),
                                           rp.F3(*r.uniform(1, 8, 3))) for _ in range(int(r.integers(0, 5)))]
    cam = rp.Pinhole.new()
    cam.set(rp.F3(float(r.uniform(-1, 1)), float(r.uniform(0, 2)), float(r.uniform(3, 5))), rp.F3(0, 0, 0))
    cam.set_fov(float(r.uniform(50, 90)))
    bg = rp.Background() if r.random() < 0.6 else rp.Background(kind=rp._abi.PTB_BG_CONSTANT, colour_a=rp.F3(*r.uniform(0, 0.6, 3)))
    return rp.DeviceScene(spheres=spheres, planes=planes, materials=mats, lights=lights, camera=cam, background=bg,
                          depth=int(r.integers(1, 7)), flags=0, eps=0.005)


@pytest.mark.parametrize("seed", [101, 102, 103, 104, 105, 106, 107, 108])
def test_random_scenes_parity(rp, po, seed):
    e = _random_scene(rp, seed)
    W, H, S = 120, 80, 3
    ref, _, _, oc = po.OracleScene(e).render(W, H, S, counters=True)
    for integ in (rp._abi.PTB_INTEGRATOR_FUSED, rp._abi.PTB_INTEGRATOR_WAVEFRONT, rp._abi.PTB_INTEGRATOR_STREAM):
        pt = rp.Tracer.new(rp.ExportedScene(e), integrator=integ, collect_counters=True)
        buf = rp.ColorBuffer.new(W, H)
        pt.render_spp(buf, S)
        c = pt.counters()
        pt.close()
        finite = np.isfinite(ref.reshape(-1, 4)).all(1) & np.isfinite(buf.pixels.reshape(-1, 4)).all(1)
        # NaN pixels (the reference never filters NaN radiance) must be NaN on both sides, and rare
        assert (np.isfinite(ref.reshape(-1, 4)).all(1) == np.isfinite(buf.pixels.reshape(-1, 4)).all(1)).mean() > 0.999
        rel = pix_rel(buf.pixels.reshape(-1, 4)[finite], ref.reshape(-1, 4)[finite])
        # transmission / high-gloss clearcoat scenes carry the ill-conditioned lobes (test_gpu_functions.py): 95 % bar
        assert (rel < 1e-4).mean() >= 0.95, (seed, integ, (rel < 1e-4).mean())
        assert np.median(rel) < 2e-6
        n = W * H * S
        for k in ("closest_hit", "any_hit", "shade", "end_sky", "end_emitter", "end_depth", "lobe_refract", "lobe_clearcoat"):
            assert abs(c[k] - oc[k]) <= max(5, 1e-3 * n), (seed, integ, k, c[k], oc[k])


def _partial_mask_scene(rp, seed):
    """a small scene whose materials assign only SOME fields (the reference's cumulative assignment): overlapping
    spheres in front of each other + checker floor, so that many accepted-primitive sets occur; <= 6 primitives, which is what
    the resolved-material table enumerates"""
    r = np.random.default_rng(seed)
    M = rp.Material
    mats = [
        M.assigning(rgb=rp.F3(*r.uniform(0.2, 1.0, 3)), roughness=float(r.uniform(0.05, 0.5)), metallic=1.0),
        M.assigning(rgb=rp.F3(*r.uniform(0.2, 1.0, 3)), clearcoat=1.0, clearcoat_gloss=float(r.uniform(0.2, 1.0)), roughness=float(r.uniform(0.1, 0.6))),
        M.assigning(roughness=float(r.uniform(0.2, 0.9)), sheen=float(r.uniform(0, 1)), specular_tint=float(r.uniform(0, 1))),
        M.assigning(rgb=rp.F3(*r.uniform(0.2, 1.0, 3)), anisotropic=float(r.uniform(0, 0.8)), ior=float(r.uniform(1.2, 1.7))),
        M.assigning(roughness=1.0, albedo_kind=rp._abi.PTB_ALBEDO_CHECKER_DIR_RATIO),
    ]
    n_sph = int(r.integers(3, 5))          # + the floor = 4..5 primitives: 15..31 accepted sets, the table fits its 12 KB
    spheres = [rp.Sphere(rp.F3(float(r.uniform(-1.5, 1.5)), float(r.uniform(-0.3, 0.8)), float(-1.2 * k + r.uniform(-0.3, 0.3))), float(r.uniform(0.6, 1.1)),
                         int(r.integers(0, 4))) for k in range(n_sph)]
    planes = [rp.Plane(rp.F3(0, -1, 0), rp.F3(0, 1, 0), 4)]
    lights = [rp.AnalyticalLight.spherical(rp.F3(3, 2.5, 2), 1.0, rp.F3(3, 3, 3))]
    cam = rp.Pinhole.new()
    cam.set(rp.F3(0.0, 0.3, 3.5), rp.F3(0, 0, -1))
    cam.set_fov(75.0)
    return rp.DeviceScene(spheres=spheres, planes=planes, materials=mats, lights=lights, camera=cam, background=rp.Background(),
                          depth=4, flags=rp._abi.PTB_SCENE_ANYHIT_IGNORES_MAX_DIST, eps=0.005)


@pytest.mark.parametrize("seed", [7, 8, 9])
def test_resolved_material_table_equals_generic_path_and_oracle(rp, po, seed, monkeypatch):
    """The wavefront integrator shades small scenes from a host-evaluated table (RMat, ptb_device.cuh) keyed by the accepted
    primitive set / material, checker parity and side.  It must agree with the generic shade path (same kernel without the
    table: PTB200_NO_RESOLVED_MATERIALS=1, read by ptb_set_scene_*) and with the oracle."""
    e = _partial_mask_scene(rp, seed)
    W, H, S = 160, 100, 4
    ref, _, _, oc = po.OracleScene(e).render(W, H, S, counters=True)
    imgs = {}
    for name, env in (("table", None), ("generic", "1")):
        if env is None:
            monkeypatch.delenv("PTB200_NO_RESOLVED_MATERIALS", raising=False)
        else:
            monkeypatch.setenv("PTB200_NO_RESOLVED_MATERIALS", env)
        pt = rp.Tracer.new(rp.ExportedScene(e), integrator=rp._abi.PTB_INTEGRATOR_WAVEFRONT, collect_counters=True)
        buf = rp.ColorBuffer.new(W, H)
        pt.render_spp(buf, S)
        c = pt.counters()
        pt.close()
        imgs[name] = buf.pixels.copy()
        rel = pix_rel(buf.pixels, ref)
        assert (rel < 1e-4).mean() >= 0.99, (name, seed, (rel < 1e-4).mean())
        assert np.median(rel) < 2e-6
        for k in ("closest_hit", "any_hit", "shade", "end_sky", "end_emitter", "end_depth", "lobe_clearcoat", "lobe_diffuse"):
            assert abs(c[k] - oc[k]) <= max(5, 1e-3 * W * H * S), (name, seed, k, c[k], oc[k])
    monkeypatch.delenv("PTB200_NO_RESOLVED_MATERIALS", raising=False)
    rel = pix_rel(imgs["table"], imgs["generic"])
    assert (rel < 1e-4).mean() >= 0.995
    # the two paths round differently (host-evaluated f32 without contraction vs device FMA): bit-identical images would mean
    # that the table was silently not built for this scene
    assert (rel > 0).any()


@pytest.mark.parametrize("wh_spp", [(97, 61, 11), (320, 200, 16), (64, 48, 37), (1280, 720, 9)])
def test_wavefront_tail_blocks_match_whole_pixels(rp, scene, oracle_demo, wh_spp):
    """The wavefront integrator hands the frame's last pixels (one per path slot) out as up to 8 sample blocks each and
    k_tail_combine adds the blocks in block order (ptb_wavefront.cuh).  The samples are the same as without the split, so
    the image must equal the fused integrator's (whole pixels only) up to f32 summation order — also when spp is not a
    multiple of the block count, when every pixel of a small frame is a tail pixel, and when only part of a large one is."""
    W, H, S = wh_spp
    imgs = []
    for integ in (rp._abi.PTB_INTEGRATOR_WAVEFRONT, rp._abi.PTB_INTEGRATOR_FUSED):
        pt = rp.Tracer.new(scene, integrator=integ)
        buf = rp.ColorBuffer.new(W, H)
        pt.render_spp(buf, S)
        assert buf.frames == S
        assert np.all(buf.pixels.reshape(-1, 4)[:, 3] == 1.0)          # every pixel received exactly S samples
        imgs.append(buf.pixels.copy())
        # a second batch continues the sample sequence: 2 S samples in two launches == the running mean of both
        pt.render_spp(buf, S)
        assert buf.frames == 2 * S and np.all(buf.pixels.reshape(-1, 4)[:, 3] == 1.0)
        pt.close()
    rel = pix_rel(imgs[0], imgs[1])
    assert (rel < 1e-4).mean() >= 0.999, (rel < 1e-4).mean()          # fused and wavefront round a few branch decisions differently
    assert np.median(rel) < 1e-6
    if W * H * S <= 200 * 150 * 16:
        ref, _, _, _ = oracle_demo.render(W, H, S)
        assert (pix_rel(imgs[0], ref) < 1e-4).mean() >= 0.99


def test_non_spherical_light_types_are_inert_like_the_reference(rp, po):
    """LightType::Rectangular / Distant exist in the reference (globals.rs:69-73) but Tracer::sample_light only implements the
    spherical arm (tracer.rs:217 `_ => {}`: direction and normal stay zero, so the cull at tracer.rs:148 drops the sample) and
    Scene::sample_lights only intersects spherical lights (scene.rs:69).  They still count in number_of_lights(): the uniform
    light pick lands on them and the spherical light's emission is scaled by the full count (tracer.rs:137-139, 214)."""
    e = rp.AnalyticalScene.new().device_export()
    rect = rp.AnalyticalLight.spherical(rp.F3(-2.0, 2.5, 1.0), 0.8, rp.F3(6, 5, 4)); rect.light.light_type = rp._abi.PTB_LIGHT_RECTANGULAR
    dist = rp.AnalyticalLight.spherical(rp.F3(0.0, 4.0, -1.0), 0.5, rp.F3(2, 2, 9)); dist.light.light_type = rp._abi.PTB_LIGHT_DISTANT
    e.lights = [rect, e.lights[0], dist]
    W, H, S = 160, 100, 4
    ref, _, _, oc = po.OracleScene(e).render(W, H, S, counters=True)
    assert oc["end_emitter"] >= 0
    for integ in (rp._abi.PTB_INTEGRATOR_WAVEFRONT, rp._abi.PTB_INTEGRATOR_FUSED, rp._abi.PTB_INTEGRATOR_STREAM):
        pt = rp.Tracer.new(rp.ExportedScene(e), integrator=integ, collect_counters=True)
        buf = rp.ColorBuffer.new(W, H)
        pt.render_spp(buf, S)
        c = pt.counters()
        pt.close()
        rel = pix_rel(buf.pixels, ref)
        assert (rel < 1e-4).mean() >= 0.99, (integ, (rel < 1e-4).mean())
        for k in ("closest_hit", "any_hit", "shade", "nee_contrib", "end_sky", "end_emitter"):
            assert abs(c[k] - oc[k]) <= max(5, 1e-3 * W * H * S), (integ, k, c[k], oc[k])
    # a third of the light picks land on the spherical light: fewer shadow rays than with that light alone
    only = rp.AnalyticalScene.new().device_export()
    _, _, _, oc1 = po.OracleScene(only).render(W, H, S, counters=True)
    assert oc["any_hit"] < 0.5 * oc1["any_hit"]


def extended_light_export(rp, quad=True, distant=True, spherical=True):
    """the demo scene lit by a downward-facing quad above the spheres and / or a distant light (PTB_SCENE_EXTENDED_LIGHTS)"""
    e = rp.AnalyticalScene.new().device_export()
    lights = []
    if quad:       # cross(u, v) = (0, -2.4, 0): emits downwards
        lights.append(rp.AnalyticalLight.rectangular(rp.F3(-1.0, 3.0, -0.6), rp.F3(2.0, 0.0, 0.0), rp.F3(0.0, 0.0, 1.2), rp.F3(9.0, 8.0, 7.0)))
    if spherical:
        lights.append(e.lights[0])
    if distant:
        lights.append(rp.AnalyticalLight.distant(rp.F3(0.4, 1.0, 0.3), rp.F3(0.5, 0.6, 0.9)))
    e.lights = lights
    e.flags |= rp._abi.PTB_SCENE_EXTENDED_LIGHTS
    return e


@pytest.mark.parametrize("precision", ["f32", "f64"])
def test_extended_light_types_match_the_oracle(rp, po, precision):
    """with PTB_SCENE_EXTENDED_LIGHTS rectangular and distant lights are sampled (and quads are hit by rays) with
    the semantics the reference's hooks were written for (tracer.rs:148 single-sided cull, tracer.rs:158 no MIS when area == 0,
    globals.rs:76-84 u / v / area).  The oracle states the same extension on the CPU; all integrators must agree with it."""
    e = extended_light_export(rp)
    W, H, S = 160, 100, 4
    ref, _, _, oc = po.OracleScene(e, precision=precision).render(W, H, S, counters=True)
    inert = extended_light_export(rp); inert.flags &= ~rp._abi.PTB_SCENE_EXTENDED_LIGHTS
    ref_inert, _, _, oc_inert = po.OracleScene(inert, precision=precision).render(W, H, S, counters=True)
    assert oc["any_hit"] > 2 * oc_inert["any_hit"]                 # all three lights cast shadow rays now
    assert lum(ref).mean() > 1.2 * lum(ref_inert).mean()
    integs = (rp._abi.PTB_INTEGRATOR_WAVEFRONT, rp._abi.PTB_INTEGRATOR_FUSED) + ((rp._abi.PTB_INTEGRATOR_STREAM,) if precision == "f32" else ())
    for integ in integs:
        pt = rp.Tracer.new(rp.ExportedScene(e), integrator=integ, collect_counters=True, precision=precision)
        buf = rp.ColorBuffer.new(W, H, precision=precision)
        pt.render_spp(buf, S)
        c = pt.counters()
        pt.close()
        rel = pix_rel(buf.pixels, ref)
        bar = 1e-4 if precision == "f32" else 1e-9
        assert (rel < bar).mean() >= 0.99, (integ, (rel < bar).mean())
        for k in ("closest_hit", "any_hit", "shade", "nee_contrib", "end_sky", "end_emitter"):
            assert abs(c[k] - oc[k]) <= max(5, 1e-3 * W * H * S), (integ, k, c[k], oc[k])


def test_quad_light_is_single_sided_and_visible(rp, po):
    """A quad in front of the left sphere, facing the camera.  Primary rays that hit it end there (tracer.rs:80-88) — with the
    reference's MIS weight power_heuristic(0, pdf) = 0 for a path that has not scattered yet, i.e. BLACK, exactly like a directly
    seen spherical light of the reference.  The same quad flipped (u and v swapped) is invisible from the camera and lights the
    sphere behind it instead.  (Lights are only ever hit in front of geometry: against the sky the stale hit_dist hides them,
    scene.rs:66-71.)"""
    def render(flip):
        e = rp.AnalyticalScene.new().device_export()
        u, v = rp.F3(0.6, 0.0, 0.0), rp.F3(0.0, 0.6, 0.0)
        if flip:
            u, v = v, u
        e.lights = [rp.AnalyticalLight.rectangular(rp.F3(-1.4, -0.3, 1.5), u, v, rp.F3(9.0, 8.0, 7.0))]
        e.flags |= rp._abi.PTB_SCENE_EXTENDED_LIGHTS
        pt = rp.Tracer.new(rp.ExportedScene(e))
        buf = rp.ColorBuffer.new(96, 96)
        pt.render_spp(buf, 8)
        pt.close()
        ref, _, _, _ = po.OracleScene(e).render(96, 96, 8)
        assert (pix_rel(buf.pixels, ref) < 1e-4).mean() >= 0.99
        return buf.pixels.reshape(-1, 4)[:, :3]
    front, back = render(False), render(True)
    black = (front == 0).all(1)
    assert 300 <= black.sum() <= 700, black.sum()          # 0.6 / 1.5 * 48 / tan(40 deg) = 23 pixels on a side
    assert not (back == 0).all(1).any()
    assert back.mean() > 1.2 * front.mean()                # flipped, it faces the sphere and lights it


@pytest.mark.parametrize("wh_spp", [(1920, 1080, 2), (3840, 2160, 2), (3840, 2160, 1)])
@pytest.mark.parametrize("strict", [False, True])
def test_benchmarked_kernel_full_frame_parity(rp, scene, oracle_demo, wh_spp, strict):
    """The instantiation bench.py measures (AUTO -> resolved-material wavefront kernel, tail blocks, FMA-corrected film
    coordinates) at the frame sizes of BASELINE.json configs[1] and configs[2], against the oracle with the shared counter RNG
    (VERDICT r1 weak #7).  spp = 2 takes the tail-block path (two one-sample blocks for the frame's last pixels), spp = 1 the
    whole-pixel path; both builds.  Bars: >= 99.5 % of pixels within 1e-4 relative for the shipped build (measured 99.80 % at
    1920x1080 x 2 spp: the outliers are paths through the clearcoat lobe at alpha = 0.001, whose sampled direction is conditioned
    to ~1e-4 in f32 — profiles/r02_function_parity.md), >= 99.95 % for the strict build (its residual is CUDA's double-rounded
    pow / sincos against glibc's), mean luminance within 1e-5."""
    W, H, S = wh_spp
    pt = rp.Tracer.new(scene, strict=strict)
    buf = rp.ColorBuffer.new(W, H)
    pt.render_spp(buf, S)
    assert buf.frames == S
    assert pt.integrator_used() == "wavefront_rm"
    ref, frames, _, _ = oracle_demo.render(W, H, S)
    assert frames == S
    px = buf.pixels.reshape(-1, 4)
    assert np.all(px[:, 3] == 1.0)
    ok = np.isfinite(px).all(1) & np.isfinite(ref.reshape(-1, 4)).all(1)     # a 0/0 clearcoat sample poisons a pixel on either side
    assert (~ok).sum() <= 2
    rel = pix_rel(buf.pixels, ref)[ok]
    frac = (rel < 1e-4).mean()
    print(f"[full-frame parity] {W}x{H}x{S} strict={strict}: within 1e-4: {frac:.6f}, within 1e-5: {(rel < 1e-5).mean():.6f}, "
          f"bit-identical: {(rel == 0).mean():.6f}, median {np.median(rel):.2e}, non-finite {(~ok).sum()}")
    assert frac >= (0.9995 if strict else 0.995), frac
    assert np.median(rel) < (1e-7 if strict else 1e-6)
    assert abs(lum(buf.pixels)[ok].mean() / lum(ref)[ok].mean() - 1) < 1e-5
    pt.close()


@pytest.mark.parametrize("strict", [True, False])
def test_config4_full_scene_parity(rp, po, strict):
    """BASELINE.json configs[3] at FULL scale — 100 000 spheres, 64 lights, sphere BVH + light BVH, the integrator AUTO picks for it
    (streaming wavefront with the dedicated traversal kernels) — against the oracle's linear scan over all spheres, shared counter
    RNG, 96x54 x 1 spp (the oracle's 2.5e9 sphere tests take ~15 s on 16 cores).  The strict build must put >= 99 % of the pixels
    within 1e-4; the shipped build's outliers are paths through the field's glass / high-gloss clearcoat lobes
    (conditioned to ~1e-4 in f32: profiles/r02_function_parity.md; a third of this scene's spheres are glass, a third gloss-0.5..1
    clearcoat), measured 94.3 %, bar 93 %, value printed."""
    sc = rp.sphere_field_scene()
    W, H, S = 96, 54, 1
    pt = rp.Tracer.new(sc, strict=strict)
    buf = rp.ColorBuffer.new(W, H)
    pt.render_spp(buf, S)
    used = pt.integrator_used()
    pt.close()
    assert used == "stream_split_bvh", used
    ref, frames, secs, _ = po.OracleScene(sc.device_export()).render(W, H, S)
    ok = np.isfinite(buf.pixels.reshape(-1, 4)).all(1) & np.isfinite(ref.reshape(-1, 4)).all(1)
    rel = pix_rel(buf.pixels, ref)[ok]
    frac = (rel < 1e-4).mean()
    print(f"[config 4 parity] strict={strict}: within 1e-4: {frac:.5f}, bit-identical {(rel == 0).mean():.5f}, median {np.median(rel):.2e}, "
          f"non-finite {(~ok).sum()}, oracle {secs:.1f} s")
    assert (~ok).sum() <= 2
    assert frac >= (0.99 if strict else 0.93), frac
    assert abs(lum(buf.pixels)[ok].mean() / lum(ref)[ok].mean() - 1) < (1e-4 if strict else 2e-3)


@pytest.mark.parametrize("wh_spp", [(96, 64, 4), (200, 150, 5), (37, 19, 9)])
def test_f64_on_the_wavefront_integrator(rp, scene, po, demo_export, wh_spp):
    """`F = f64` (lib.rs:5-6) on the shared-memory wavefront integrator: same paths as the f64 oracle to
    1e-9, same image as the fused f64 kernel, tail blocks and odd frame sizes included."""
    W, H, S = wh_spp
    imgs = {}
    for name, integ in (("wave", rp._abi.PTB_INTEGRATOR_WAVEFRONT), ("fused", rp._abi.PTB_INTEGRATOR_FUSED)):
        pt = rp.Tracer.new(scene, precision="f64", integrator=integ)
        buf = rp.ColorBuffer.new(W, H, "f64")
        pt.render_spp(buf, S)
        assert pt.integrator_used() == ("wavefront_f64" if name == "wave" else "fused_f64")
        assert buf.frames == S and np.all(buf.pixels.reshape(-1, 4)[:, 3] == 1.0)
        imgs[name] = buf.pixels.copy()
        pt.close()
    ref, _, _, _ = po.OracleScene(demo_export, "f64").render(W, H, S)
    rel = pix_rel(imgs["wave"], ref)
    assert (rel < 1e-9).mean() >= 0.99 and np.median(rel) < 1e-13
    assert (pix_rel(imgs["wave"], imgs["fused"]) < 1e-11).mean() >= 0.995


def test_f64_wavefront_on_a_bvh_scene_and_with_a_signed_distance_body(rp, po):
    sc = _small_field(rp, 800, 2)
    W, H, S = 96, 54, 2
    pt = rp.Tracer.new(sc, precision="f64", integrator=rp._abi.PTB_INTEGRATOR_WAVEFRONT)
    buf = rp.ColorBuffer.new(W, H, "f64")
    pt.render_spp(buf, S)
    assert pt.integrator_used() == "wavefront_bvh_f64"
    pt.close()
    ref, _, _, _ = po.OracleScene(sc.device_export(), "f64").render(W, H, S)
    assert (pix_rel(buf.pixels, ref) < 1e-8).mean() >= 0.99
    sd = rp.sdf_demo_scene()
    pt = rp.Tracer.new(sd, precision="f64", integrator=rp._abi.PTB_INTEGRATOR_WAVEFRONT)
    b2 = rp.ColorBuffer.new(W, H, "f64")
    pt.render_spp(b2, S)
    pt.close()
    ref2, _, _, _ = po.OracleScene(sd.device_export(), "f64").render(W, H, S)
    assert (pix_rel(b2.pixels, ref2) < 1e-7).mean() >= 0.995


def test_bvh_work_counters(rp):
    """ptb_counters.bvh_nodes / bvh_leaf_tests (bench.py's work model for BASELINE configs 4 and 5): the three integrators run the
    same traversal on scenes below the split threshold, so they must count the same work for the same rays; the dedicated
    traversal kernels of large scenes (persistent lanes, parked leaves) visit the same tree in another order."""
    sc = _small_field(rp, 3000, 2)
    W, H, S = 96, 54, 2
    got = {}
    for name, integ in (("fused", rp._abi.PTB_INTEGRATOR_FUSED), ("wave", rp._abi.PTB_INTEGRATOR_WAVEFRONT), ("stream", rp._abi.PTB_INTEGRATOR_STREAM)):
        pt = rp.Tracer.new(sc, integrator=integ, collect_counters=True)
        buf = rp.ColorBuffer.new(W, H)
        pt.render_spp(buf, S)
        got[name] = pt.counters()
        pt.close()
    rays = got["fused"]["closest_hit"] + got["fused"]["any_hit"]
    assert got["fused"]["bvh_nodes"] > 2 * rays and got["fused"]["bvh_leaf_tests"] > 0.2 * rays      # a 3000-sphere tree is ~11 levels deep
    for name in ("wave", "stream"):
        for k in ("bvh_nodes", "bvh_leaf_tests"):
            assert abs(got[name][k] - got["fused"][k]) <= 2e-3 * got["fused"][k], (name, k, got[name][k], got["fused"][k])
    big = rp.sphere_field_scene(n_spheres=20000, n_lights_side=2)                                  # above ST_SPLIT_MIN_SPHERES
    pt = rp.Tracer.new(big, collect_counters=True)
    buf = rp.ColorBuffer.new(W, H)
    pt.render_spp(buf, S)
    c = pt.counters()
    assert pt.integrator_used() == "stream_split_bvh"
    pt.close()
    pt = rp.Tracer.new(big, collect_counters=True, integrator=rp._abi.PTB_INTEGRATOR_FUSED)
    pt.render_spp(buf, S)
    f = pt.counters()
    pt.close()
    assert abs(c["closest_hit"] - f["closest_hit"]) <= 1e-2 * f["closest_hit"]      # (a few paths flip a branch between the two instantiations)
    assert 0.5 < c["bvh_nodes"] / f["bvh_nodes"] < 2.0 and 0.5 < c["bvh_leaf_tests"] / f["bvh_leaf_tests"] < 2.0


def test_resolved_material_kernel_with_and_without_embedded_primitives(rp, po):
    """The resolved-material wavefront kernel has two instantiations: scenes of up to 8 spheres / 4 planes / 4 lights read their
    primitives from the copies in the kernel parameter (constant bank), larger small scenes through the shared-memory scene view.
    Both against the oracle and against the fused integrator: the demo scene (embedded), a 16-sphere field without BVH (view),
    and a 9-sphere scene right above the limit."""
    def field(n):
        e = rp.sphere_field_scene(n_spheres=n, n_lights_side=2).device_export()
        e.flags |= rp._abi.PTB_SCENE_NO_BVH
        e.camera.set(rp.F3(0.0, 2.0, 6.0), rp.F3(0.0, 0.0, -20.0))
        return e
    W, H, S = 128, 72, 3
    for name, e in (("demo", rp.AnalyticalScene.new().device_export()), ("field16", field(16)), ("field9", field(9)), ("field8", field(8))):
        ref, _, _, _ = po.OracleScene(e).render(W, H, S)
        imgs = {}
        for integ in (rp._abi.PTB_INTEGRATOR_WAVEFRONT, rp._abi.PTB_INTEGRATOR_FUSED):
            pt = rp.Tracer.new(rp.ExportedScene(e), integrator=integ)
            buf = rp.ColorBuffer.new(W, H)
            pt.render_spp(buf, S)
            used = pt.integrator_used()
            pt.close()
            imgs[integ] = buf.pixels.copy()
            if integ == rp._abi.PTB_INTEGRATOR_WAVEFRONT:
                assert used == "wavefront_rm", (name, used)
            rel = pix_rel(buf.pixels, ref)
            assert (rel < 1e-4).mean() >= 0.97, (name, used, (rel < 1e-4).mean())
            assert np.median(rel) < 1e-6
        assert (pix_rel(imgs[rp._abi.PTB_INTEGRATOR_WAVEFRONT], imgs[rp._abi.PTB_INTEGRATOR_FUSED]) < 1e-4).mean() >= 0.99
