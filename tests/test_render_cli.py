"""The headless renderer (python -m rust_pathtracer_b200.render; SURVEY.md §8 f4: replaces the reference's windowed demo app,
renderer/src/main.rs:74-193) and its image writers."""
import os

import numpy as np
import pytest


def test_exr_writer_round_trip(tmp_path):
    from rust_pathtracer_b200 import render as R
    rng = np.random.default_rng(0)
    w, h = 37, 19
    img = rng.uniform(0, 4, size=(h, w, 4)).astype(np.float32)
    img[0, 0] = [np.inf, 1e-30, 65504.0, 1.0]
    path = R.write_exr(str(tmp_path / "x.exr"), img.reshape(-1), w, h)
    back, w2, h2 = R.read_exr(path)
    assert (w2, h2) == (w, h) and np.array_equal(back, img)
    raw = open(path, "rb").read()
    assert raw[:4] == bytes([0x76, 0x2F, 0x31, 0x01])                   # OpenEXR magic
    assert os.path.getsize(path) == raw.index(b"screenWindowWidth") + len("screenWindowWidth") + 1 + 6 + 4 + 4 + 1 + 8 * h + h * (8 + 16 * w)


@pytest.mark.gpu
@pytest.mark.parametrize("scene_size", [("demo", 96, 54), ("sdf", 64, 36), ("media", 64, 36), ("lights", 64, 36)])
def test_render_cli_writes_the_converted_frame(tmp_path, rp, po, scene_size):
    """CLI -> Tracer.render_spp -> convert_to_u8 -> PNG / EXR: the bytes on disk are the reference's encoding (buffer.rs:55-64) of
    the image the oracle renders with the same counter RNG."""
    from rust_pathtracer_b200 import render as R
    scene, w, h = scene_size
    exr = str(tmp_path / "a.exr")
    R.main(["--scene", scene, "--size", f"{w}x{h}", "--spp", "4", "--batch", "3", "--out", exr])
    lin, w2, h2 = R.read_exr(exr)
    assert (w2, h2) == (w, h) and np.all(lin[..., 3] == 1.0)
    ref, _, _, _ = po.OracleScene(R.SCENES[scene]().device_export()).render(w, h, 4)
    ref = ref.reshape(h, w, 4)
    rel = np.abs(lin[..., :3] - ref[..., :3]).max(-1) / np.maximum(np.abs(ref[..., :3]).max(-1), 1e-3)
    assert (rel < 1e-4).mean() >= 0.97
    png = str(tmp_path / "a.png")
    R.main(["--scene", scene, "--size", f"{w}x{h}", "--spp", "4", "--out", png])
    out = png if os.path.exists(png) else png[:-4] + ".ppm"
    assert os.path.getsize(out) > 100
    try:
        from PIL import Image
        u8 = np.asarray(Image.open(out).convert("RGB"))
        want = po.convert_to_u8(lin.reshape(-1)).reshape(h, w, 4)[..., :3]
        assert (np.abs(u8.astype(int) - want.astype(int)) <= 1).mean() > 0.999
    except ImportError:
        pass


def _atrous_numpy(img, iterations, sigma):
    """numpy statement of k_atrous (ptb_kernels.cuh): the checker of the device denoiser"""
    h, w, _ = img.shape
    k = {0: 0.375, 1: 0.25, 2: 0.0625}
    cur = img.astype(np.float64)
    for it in range(iterations):
        step, inv = 1 << it, 1.0 / (sigma * sigma)
        c = cur[..., :3]
        c_ok = np.isfinite(c).all(-1)
        acc, wsum = np.zeros((h, w, 3)), np.zeros((h, w))
        for dy in range(-2, 3):
            for dx in range(-2, 3):
                sy, sx = dy * step, dx * step
                ys, xs = np.arange(h) + sy, np.arange(w) + sx
                vy, vx = (ys >= 0) & (ys < h), (xs >= 0) & (xs < w)
                q = np.full((h, w, 3), np.nan)
                q[np.ix_(vy, vx)] = c[np.ix_(ys[vy], xs[vx])]
                q_ok = np.isfinite(q).all(-1)
                wgt = np.full((h, w), k[abs(dx)] * k[abs(dy)])
                with np.errstate(invalid="ignore", over="ignore"):
                    d2 = ((np.where(q_ok[..., None], q, 0.0) - np.where(c_ok[..., None], c, 0.0)) ** 2).sum(-1)
                wgt = np.where(c_ok, wgt * np.exp(-d2 * inv), wgt) * q_ok
                acc += wgt[..., None] * np.where(q_ok[..., None], q, 0.0)
                wsum += wgt
        out = cur.copy()
        good = wsum > 0
        out[..., :3] = np.where(good[..., None], acc / np.where(good, wsum, 1.0)[..., None], c)
        cur = out
        sigma *= 0.5
    return cur


@pytest.mark.gpu
def test_denoiser_matches_its_numpy_statement_and_reduces_noise(rp, po):
    W, H = 160, 90
    scene = rp.AnalyticalScene.new()
    pt = rp.Tracer.new(scene)
    buf = rp.ColorBuffer.new(W, H)
    pt.render_spp(buf, 4)
    noisy = buf.read_pixels().reshape(H, W, 4).copy()
    den = pt.denoise(3, 0.4).reshape(H, W, 4)
    ref = _atrous_numpy(noisy, 3, 0.4)
    assert np.abs(den[..., :3] - ref[..., :3]).max() < 2e-5
    assert np.array_equal(den[..., 3], noisy[..., 3])
    assert np.array_equal(buf.read_pixels(), noisy.reshape(-1))          # the running mean itself is untouched
    # against a converged image the filtered 4-spp frame is closer than the raw one
    conv = rp.ColorBuffer.new(W, H)
    pt2 = rp.Tracer.new(scene, seed=99)
    pt2.render_spp(conv, 2048)
    truth = conv.read_pixels().reshape(H, W, 4)[..., :3]
    ok = np.isfinite(truth).all(-1)
    e_raw = np.sqrt(((noisy[..., :3] - truth)[ok] ** 2).mean()); e_den = np.sqrt(((den[..., :3] - truth)[ok] ** 2).mean())
    print(f"[denoise] RMSE vs 2048 spp: raw 4 spp {e_raw:.4f}, filtered {e_den:.4f}")
    assert e_den < 0.75 * e_raw
    # a poisoned pixel is repaired from its neighbours
    bad = noisy.copy(); bad[40, 80, :3] = np.nan
    b2 = rp.ColorBuffer.new(W, H); b2.pixels[:] = bad.reshape(-1); b2.frames = 4
    pt._upload(b2)
    rep = pt.denoise(2, 0.4).reshape(H, W, 4)
    assert np.isfinite(rep).all()
    pt.close(); pt2.close()
