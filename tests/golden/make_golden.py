"""Regenerates the committed fixtures under tests/golden/.

1. screenshot_means.json — client-area mean sRGB8 of the reference's ONLY verification artefact,
   /root/reference/images/spheres.png (a macOS window screenshot).  Needs the
   reference tree, so it runs in the build container only; the JSON travels to the GPU box.
2. oracle_kat_f32.npz — per-function known answers produced by the CPU oracle (f32) on seeded
   inputs.  The reference ships no golden vectors, so these pin the ORACLE against
   regressions; the oracle itself is pinned by the hand-derived values in tests/test_oracle.py.

Run:  python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def screenshot_means():
    from PIL import Image
    path = "/root/reference/images/spheres.png"
    im = np.asarray(Image.open(path).convert("RGB"), dtype=np.float64)
    h, w, _ = im.shape
    # the 800x600 client area is shown at 2x (1600x1200) below the title bar, inside the window shadow
    cw, ch = 1600, 1200
    rows = np.nonzero(im[:, w // 2].sum(1) > 0)[0]
    cols = np.nonzero(im[h // 2].sum(1) > 0)[0]
    x0 = int(cols[0])
    y0 = int(rows[-1]) + 1 - ch          # the title bar sits above the client area
    assert int(cols[-1]) + 1 - x0 == cw
    client = im[y0:y0 + ch, x0:x0 + cw]
    # 40x30 block means (each block = 20x20 pixels of the 800x600 render) as a structural pin
    thumb = client.reshape(30, 40, 40, 40, 3).mean(axis=(1, 3))
    return {"source": "images/spheres.png", "image_size": [w, h], "client_rect": [x0, y0, cw, ch],
            "mean_srgb8": client.reshape(-1, 3).mean(0).tolist(),
            "thumb_40x30_srgb8": np.round(thumb, 2).tolist()}


def oracle_kats():
    from oracle import pyoracle as po
    import rust_pathtracer_b200 as rp
    rng = np.random.default_rng(20261017)
    sc = po.OracleScene(rp.AnalyticalScene.new().device_export())
    n = 64

    def unit(n):
        v = rng.normal(size=(3, n)); v /= np.linalg.norm(v, axis=0); return v.astype(np.float32)
    out = {}
    o = (rng.uniform(-3, 3, size=(3, n))).astype(np.float32); o[1] = np.abs(o[1]) + 0.1
    d = unit(n)
    hd = rng.choice([-1.0, 0.5, 2.0, 100.0], size=n).astype(np.float32)
    ch = sc.closest_hit(o, d, hd, want_material=True)
    out.update(ch_o=o, ch_d=d, ch_hd=hd, **{f"ch_{k}": v for k, v in ch.items()})
    out["ah_hit"] = sc.any_hit(o, d, np.full(n, 3.0, np.float32))
    out["bg"] = sc.background(d)
    p2 = rng.uniform(0, 1, size=(2, n)).astype(np.float32); off = rng.uniform(0, 1, size=(2, n)).astype(np.float32)
    go, gd = sc.gen_ray(p2, off, 800, 600)
    out.update(gr_p2=p2, gr_off=off, gr_o=go, gr_d=gd)
    r1, r2, coin = (rng.uniform(0, 1, n).astype(np.float32) for _ in range(3))
    sl = sc.sample_light(0, o, r1, r2)
    out.update(sl_r1=r1, sl_r2=r2, **{f"sl_{k}": v for k, v in sl.items()})
    nrm = unit(n)
    v = unit(n); v = np.where((v * nrm).sum(0) < 0, -v, v).astype(np.float32)
    l = unit(n)
    eta = np.full(n, 1 / 1.45, np.float32)
    out.update(bs_n=nrm, bs_v=v, bs_l=l, bs_eta=eta, bs_coin=coin)
    for mi in range(3):
        f, pdf = sc.disney_eval(mi, eta, v, nrm, l)
        s = sc.disney_sample(mi, eta, v, nrm, l, r1, r2, coin)
        out.update({f"ev{mi}_f": f, f"ev{mi}_pdf": pdf, f"sm{mi}_lobe": s["lobe"], f"sm{mi}_l": s["l"], f"sm{mi}_f": s["f"], f"sm{mi}_pdf": s["pdf"]})
    px, frames, _, _ = sc.render(32, 24, 2)
    out["img_32x24_2spp"] = px
    out["rng_b0"] = po.rng(np.arange(8, dtype=np.uint32), np.arange(8, dtype=np.uint64) * 1000003, 0)
    out["rng_b3"] = po.rng(np.arange(8, dtype=np.uint32), np.arange(8, dtype=np.uint64) * 1000003, 3)
    return out


if __name__ == "__main__":
    if os.path.exists("/root/reference/images/spheres.png"):
        with open(os.path.join(HERE, "screenshot_means.json"), "w") as f:
            json.dump(screenshot_means(), f, indent=1)
    np.savez_compressed(os.path.join(HERE, "oracle_kat_f32.npz"), **oracle_kats())
    print("golden fixtures written")
