"""GPU tests of the C-ABI contract around the render path: precision switches on a live tracer, frame-size limits, peer-slot
invalidation, caller-owned page-locking, the asynchronous download and the upload-skipping drop-in call — plus image parity of
the BENCHMARKED instantiation (resolved-material wavefront kernel, tail blocks, division-free film coordinates) at the frame
sizes BASELINE.json quotes: 1920x1080 and 3840x2160."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def pix_rel(a, b):
    a = a.reshape(-1, 4)[:, :3].astype(np.float64); b = b.reshape(-1, 4)[:, :3].astype(np.float64)
    return np.abs(a - b).max(1) / np.maximum(np.abs(b).max(1), 1e-3)


@pytest.fixture(scope="module")
def scene(rp):
    return rp.AnalyticalScene.new()


# ---- parity of the benchmarked kernel at the benchmarked frame sizes ------------------------------------------------------
@pytest.mark.parametrize("whs", [(1920, 1080, 2), (3840, 2160, 1), (3840, 2160, 2)])
def test_benchmarked_kernel_parity_at_full_frame_sizes(rp, scene, oracle_demo, whs):
    """the kernel bench.py times (k_render_wavefront<RM>, tail items from 2 spp, FMA film quotients) against
    the oracle on the SAME samples at 1080p and 4K.  A 4K 1-spp oracle frame costs well under a second per host core-second."""
    W, H, spp = whs
    lib = rp._abi.load()
    exact, bad = C.c_uint32(), C.c_uint32()
    rp._abi.check(lib.ptb_test_film_quotients_f32(W, H, C.byref(exact), C.byref(bad)))
    pt = rp.Tracer.new(scene, integrator=rp._abi.PTB_INTEGRATOR_AUTO)           # AUTO = what bench.py runs
    buf = rp.ColorBuffer.new(W, H)
    pt.render_spp(buf, spp)
    ref, frames, _, _ = oracle_demo.render(W, H, spp)
    assert frames == spp == buf.frames
    got = buf.read_pixels()
    rel = pix_rel(got, ref)
    frac = float((rel < 1e-4).mean())
    assert frac >= 0.995, (W, H, spp, frac, int(exact.value))
    assert np.median(rel) < 1e-6
    assert np.all(got.reshape(-1, 4)[:, 3] == 1.0) and np.isfinite(got).all()
    la = (got.reshape(-1, 4)[:, :3].astype(np.float64) @ np.array([0.212671, 0.715160, 0.072169])).mean()
    lb = (ref.reshape(-1, 4)[:, :3].astype(np.float64) @ np.array([0.212671, 0.715160, 0.072169])).mean()
    assert abs(la / lb - 1) < 2e-5, (la, lb)
    # the same frame through the IEEE-division film path and without tail blocks: 1 spp never has tail blocks, 2 spp does;
    # both must agree with a launch that uses neither (fused integrator: IEEE film coordinates, whole pixels only)
    pf = rp.Tracer.new(scene, integrator=rp._abi.PTB_INTEGRATOR_FUSED)
    fb = rp.ColorBuffer.new(W, H)
    pf.render_spp(fb, spp)
    rel2 = pix_rel(got, fb.read_pixels())
    assert float((rel2 < 1e-4).mean()) >= 0.995
    pt.close(); pf.close()


# ---- precision switch on a live tracer -------------------------------------------------------------------------
def test_precision_switch_reallocates_the_frame(rp, scene):
    A = rp._abi
    lib = A.load()
    t32 = rp.Tracer.new(scene)
    buf = rp.ColorBuffer.new(64, 48)
    t32.render_spp(buf, 2)
    ref32 = buf.read_pixels().copy()
    # same handle, now an f64 scene: the float4 accumulators must not be reinterpreted as double4
    sc64, keep = scene.device_export().to_c("f64")
    A.check(lib.ptb_set_scene_f64(t32._handle(), C.byref(sc64)))
    px = np.zeros(64 * 48 * 4, np.float64)
    A.check(lib.ptb_render_frame_f64(t32._handle(), 64, 48, 0, px.ctypes.data))
    t64 = rp.Tracer.new(scene, precision="f64")
    b64 = rp.ColorBuffer.new(64, 48, "f64")
    t64.render(b64)
    assert np.array_equal(px, b64.read_pixels())
    f = C.c_uint64()
    A.check(lib.ptb_frames(t32._handle(), C.byref(f)))
    assert f.value == 1
    # calls of the old precision are refused, not misread
    assert lib.ptb_download_f32(t32._handle(), ref32.ctypes.data) == A.PTB_E_PRECISION
    # and back to f32: identical to a fresh f32 tracer
    sc32, keep32 = scene.device_export().to_c("f32")
    A.check(lib.ptb_set_scene_f32(t32._handle(), C.byref(sc32)))
    p32 = np.zeros(64 * 48 * 4, np.float32)
    A.check(lib.ptb_render_frame_f32(t32._handle(), 64, 48, 0, p32.ctypes.data))
    A.check(lib.ptb_render_frame_f32(t32._handle(), 64, 48, 1, p32.ctypes.data))
    fresh = rp.ColorBuffer.new(64, 48)
    tf = rp.Tracer.new(scene)
    tf.render(fresh); tf.render(fresh)
    assert np.array_equal(p32, fresh.read_pixels())
    for t in (t32, t64, tf):
        t.close()


def test_precision_switch_with_external_accumulator_is_refused(rp, scene):
    import torch
    A = rp._abi
    lib = A.load()
    t = rp.Tracer.new(scene)
    acc = torch.zeros(32 * 32 * 4, dtype=torch.float32, device="cuda")
    t.bind_accumulator(acc.data_ptr(), 32, 32)
    sc64, keep = scene.device_export().to_c("f64")
    assert lib.ptb_set_scene_f64(t._handle(), C.byref(sc64)) == A.PTB_E_PRECISION
    assert b"external accumulator" in lib.ptb_last_error()
    t.render_samples(1, 0); t.synchronize()              # still a working f32 tracer
    assert float(acc.view(-1, 4)[:, 3].min().item()) == 1.0
    A.check(lib.ptb_bind_accumulator(t._handle(), None, 32, 32))     # back to a library-owned frame: now the switch works
    A.check(lib.ptb_set_scene_f64(t._handle(), C.byref(sc64)))
    t.close()


# ---- frame-size limits ------------------------------------------------------------------------------------------
def test_frames_wider_than_16_bit_coordinates(rp, scene, oracle_demo):
    A = rp._abi
    lib = A.load()
    W, H = 70000, 3
    explicit = rp.Tracer.new(scene, integrator=A.PTB_INTEGRATOR_WAVEFRONT)
    A.check(lib.ptb_resize(explicit._handle(), W, H))
    assert lib.ptb_render(explicit._handle(), 1, 0) == A.PTB_E_UNSUPPORTED       # would alias pixels: refused
    assert b"65535" in lib.ptb_last_error()
    explicit.close()
    auto = rp.Tracer.new(scene)                                                  # AUTO falls back to the fused integrator
    buf = rp.ColorBuffer.new(W, H)
    auto.render_spp(buf, 1)
    ref, _, _, _ = oracle_demo.render(W, H, 1)
    assert (pix_rel(buf.read_pixels(), ref) < 1e-4).mean() > 0.99
    # padded work-item count beyond the 32-bit hand-out counter
    assert lib.ptb_resize(auto._handle(), 1, 1 << 29) == A.PTB_E_INVALID
    assert b"work items" in lib.ptb_last_error()
    auto.close()


def test_resize_invalidates_the_peer_target(rp, scene):
    """peer slots are sized for the frame they were created for; a resize must not leave the render kernel storing
    into them (single-GPU: the tracer is its own root)."""
    A = rp._abi
    lib = A.load()
    t = rp.Tracer.new(scene, integrator=A.PTB_INTEGRATOR_WAVEFRONT)
    A.check(lib.ptb_resize(t._handle(), 64, 32))
    t.peer_slots_create(1)
    t.peer_set_target(0, 0)
    t.render_samples(2, 0); t.peer_sum(0); t.synchronize()
    A.check(lib.ptb_resize(t._handle(), 128, 64))              # 4x the pixels: the old slot is too small
    t.render_samples(1, 0)                                     # flush target was cleared: accumulates locally, in bounds
    buf = rp.ColorBuffer.new(128, 64); buf._tracer = t
    t._size = (128, 64)
    t.download(buf)
    assert np.all(buf.read_pixels().reshape(-1, 4)[:, 3] == 1.0)
    assert lib.ptb_peer_set_target(t._handle(), 0, 0) == A.PTB_E_INVALID         # stale slots are refused
    t.peer_slots_close()
    t.close()


# ---- caller-owned page-locking, upload skipping, asynchronous download ---------------------------------------------------------
def test_pin_unpin_host(rp):
    lib = rp._abi.load()
    a = np.zeros(1 << 20, np.float32)
    assert lib.ptb_pin_host(C.c_void_p(a.ctypes.data), a.nbytes) == 0
    assert lib.ptb_pin_host(C.c_void_p(a.ctypes.data), a.nbytes) == 0          # already pinned: not an error
    assert lib.ptb_unpin_host(C.c_void_p(a.ctypes.data)) == 0
    assert lib.ptb_unpin_host(C.c_void_p(a.ctypes.data)) == 0                  # not pinned: not an error either
    assert lib.ptb_pin_host(None, 16) == rp._abi.PTB_E_INVALID


def test_drop_in_loop_skips_redundant_uploads_but_never_misses_an_edit(rp, scene, oracle_demo):
    W, H = 96, 64
    pt = rp.Tracer.new(scene)
    a = rp.ColorBuffer.new(W, H)            # never escapes: uploads skipped after the first call
    b = rp.ColorBuffer.new(W, H)            # escaped from the start: every call uploads (reference semantics)
    _ = b.pixels
    pb = rp.Tracer.new(scene)
    for _i in range(5):
        pt.render(a); pb.render(b)
    # (not bit-equal: an upload turns the running mean back into a sum, mean * frames, which rounds differently from the resident sum)
    assert np.allclose(a.read_pixels(), b.read_pixels(), rtol=2e-6, atol=1e-7)
    ref, _, _, _ = oracle_demo.render(W, H, 5)
    assert (pix_rel(a.read_pixels(), ref) < 1e-4).mean() > 0.99
    # an edit through the public field is seen even on the buffer that used the fast path so far
    a.pixels[:] = 0.5
    pt.render(a)
    b.pixels[:] = 0.5
    pb.render(b)
    assert np.allclose(a.read_pixels(), b.read_pixels(), rtol=2e-6, atol=1e-7) and a.frames == 6
    assert abs(float(a.read_pixels().reshape(-1, 4)[:, :3].mean()) - 0.5) < 0.2      # 5/6 of the edited grey is in the mean
    # editing `frames` alone (the reference's reset idiom) also defeats the skip
    c = rp.ColorBuffer.new(W, H)
    pt.render(c); pt.render(c)
    c.frames = 0
    pt.render(c)
    one, _, _, _ = oracle_demo.render(W, H, 1)
    assert (pix_rel(c.read_pixels(), one) < 1e-4).mean() > 0.99
    # another tracer rendered into the device image in between: the host copy is not what THIS tracer's device holds
    d = rp.ColorBuffer.new(W, H)
    pt.render(d)
    e = rp.ColorBuffer.new(W, H)
    pt.render(e); pt.render(e)             # the tracer's device image now belongs to `e`
    pt.render(d)                           # must upload d (frames 1), not continue from e
    two, _, _, _ = oracle_demo.render(W, H, 2)
    assert d.frames == 2 and (pix_rel(d.read_pixels(), two) < 1e-4).mean() > 0.99
    pt.close(); pb.close()


def test_async_download_matches_blocking_download(rp, scene):
    W, H = 320, 200
    pt = rp.Tracer.new(scene)
    bufs = [rp.ColorBuffer.new(W, H) for _ in range(3)]
    sync = []
    ref = rp.ColorBuffer.new(W, H)
    pt.render_spp(ref, 2); sync.append(ref.read_pixels().copy())
    pt.render_spp(ref, 3); sync.append(ref.read_pixels().copy())
    pt.render_spp(ref, 1); sync.append(ref.read_pixels().copy())
    # the same three steps, each download in flight while the next step is traced
    pt.clear()
    bufs[0].frames = 0
    pt.render_spp(bufs[0], 2, download="async")
    for k, (n, prev) in enumerate(((3, 0), (1, 1)), start=1):
        bufs[k].frames = bufs[prev].frames
        bufs[k]._tracer = pt
        pt.render_spp(bufs[k], n, download="async")
    pt.wait_download()
    for k in range(3):
        assert np.array_equal(bufs[k].read_pixels(), sync[k]), k
    pt.close()


def test_scene_reexport_reuses_the_bvh_only_for_unchanged_spheres(rp):
    """ptb_set_scene_* keys the device BVH by the sphere data: re-exporting the same scene (bench.py's e2e step does it every step,
    like the reference re-reads its scene per ray) must not rebuild it, and must rebuild it as soon as one sphere moves."""
    import time
    sc = rp.sphere_field_scene(n_spheres=30000, n_lights_side=2)
    pt = rp.Tracer.new(sc)
    W, H = 96, 54
    a = rp.ColorBuffer.new(W, H); pt.render_spp(a, 2)
    pod = pt.prepare_scene()
    t0 = time.perf_counter(); pt.sync_scene(pod); t_same = time.perf_counter() - t0
    b = rp.ColorBuffer.new(W, H); pt.render_spp(b, 2)
    assert np.array_equal(a.pixels, b.pixels)
    export = sc.device_export()
    export.spheres[0].center = rp.F3(0.0, 500.0, 0.0)
    big = export.spheres[1]; big.center = rp.F3(0.0, 1.5, 6.0); big.radius = 2.5      # a large sphere in front of the camera
    pod2 = pt.prepare_scene()
    t0 = time.perf_counter(); pt.sync_scene(pod2); t_changed = time.perf_counter() - t0
    c = rp.ColorBuffer.new(W, H); pt.render_spp(c, 2)
    assert not np.array_equal(a.pixels, c.pixels)
    fresh = rp.Tracer.new(sc)
    d = rp.ColorBuffer.new(W, H); fresh.render_spp(d, 2)
    assert np.array_equal(c.pixels, d.pixels)                      # the rebuilt tree is the tree a new tracer builds
    assert t_same < 0.5 * t_changed, (t_same, t_changed)
    pt.close(); fresh.close()
