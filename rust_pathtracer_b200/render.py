"""Headless renderer — the device-path replacement for the reference's windowed demo app
(renderer/src/main.rs:74-193: `pt.render(&mut buffer); buffer.convert_to_u8(frame); present`).

    python -m rust_pathtracer_b200.render --scene demo --size 800x600 --spp 256 --out spheres.png

Writes a PNG (via PIL when available, else a binary PPM).  Needs a CUDA device: there is no CPU path.
"""
from __future__ import annotations

import argparse
import time

import numpy as np

from . import AnalyticalScene, ColorBuffer, Tracer, divergence_stress_scene, sdf_demo_scene, sphere_field_scene

SCENES = {"demo": AnalyticalScene.new, "field": sphere_field_scene, "stress": divergence_stress_scene, "sdf": sdf_demo_scene}


def write_image(path: str, rgba8: np.ndarray, w: int, h: int) -> str:
    img = rgba8.reshape(h, w, 4)[..., :3]
    try:
        from PIL import Image
        Image.fromarray(img, "RGB").save(path)
        return path
    except Exception:
        ppm = path.rsplit(".", 1)[0] + ".ppm"
        with open(ppm, "wb") as f:
            f.write(f"P6 {w} {h} 255\n".encode())
            f.write(img.tobytes())
        return ppm


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("--scene", default="demo", choices=sorted(SCENES))
    ap.add_argument("--size", default="800x600")
    ap.add_argument("--spp", type=int, default=256)
    ap.add_argument("--batch", type=int, default=64, help="samples per device pass (progressive)")
    ap.add_argument("--out", default="render.png")
    ap.add_argument("--precision", default="f32", choices=["f32", "f64"])
    args = ap.parse_args(argv)
    w, h = (int(x) for x in args.size.lower().split("x"))
    scene = SCENES[args.scene]()
    pt = Tracer.new(scene, precision=args.precision)
    buf = ColorBuffer.new(w, h, args.precision)
    t0 = time.perf_counter()
    while buf.frames < args.spp:
        pt.render_spp(buf, min(args.batch, args.spp - buf.frames), download=False)
    pt.download(buf)
    dt = time.perf_counter() - t0
    frame = np.zeros(w * h * 4, np.uint8)
    pt.convert_to_u8(frame)                       # buffer.rs:55-64 on the device-resident image
    out = write_image(args.out, frame, w, h)
    print(f"{args.scene} {w}x{h} {buf.frames} spp in {dt * 1e3:.1f} ms ({w * h * buf.frames / dt / 1e6:.1f} Msamples/s incl. transfers) -> {out}")
    pt.close()


if __name__ == "__main__":
    main()
