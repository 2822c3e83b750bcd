"""Headless renderer — the device-path replacement for the reference's windowed demo app
(renderer/src/main.rs:74-193: `pt.render(&mut buffer); buffer.convert_to_u8(frame); present`).

    python -m rust_pathtracer_b200.render --scene demo --size 800x600 --spp 256 --out spheres.png

Writes a PNG (via PIL when available, else a binary PPM) from the gamma-encoded bytes of `convert_to_u8`, or — `--out x.exr` — the
linear f32 running mean as an uncompressed scan-line OpenEXR file (writer below, no dependency).  Needs a CUDA device: there is no
CPU path.
"""
from __future__ import annotations

import argparse
import time

import numpy as np

from . import AnalyticalScene, ColorBuffer, Tracer, divergence_stress_scene, lights_demo_scene, media_demo_scene, sdf_demo_scene, sphere_field_scene

SCENES = {"demo": AnalyticalScene.new, "field": sphere_field_scene, "stress": divergence_stress_scene, "sdf": sdf_demo_scene,
          "media": media_demo_scene, "lights": lights_demo_scene}


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


def write_exr(path: str, rgba: np.ndarray, w: int, h: int) -> str:
    """Linear float image -> OpenEXR 2.0, single part, scan lines, no compression, FLOAT channels A B G R (the format stores
    channels alphabetically and, per scan line, one channel after the other)."""
    import struct
    px = np.ascontiguousarray(rgba, dtype=np.float32).reshape(h, w, 4)

    def attr(name: str, typ: str, data: bytes) -> bytes:
        return name.encode() + b"\0" + typ.encode() + b"\0" + struct.pack("<i", len(data)) + data

    chlist = b"".join(c.encode() + b"\0" + struct.pack("<iB3xii", 2, 0, 1, 1) for c in "ABGR") + b"\0"      # 2 = FLOAT
    box = struct.pack("<4i", 0, 0, w - 1, h - 1)
    head = struct.pack("<II", 20000630, 2)
    head += attr("channels", "chlist", chlist) + attr("compression", "compression", b"\0") + attr("dataWindow", "box2i", box)
    head += attr("displayWindow", "box2i", box) + attr("lineOrder", "lineOrder", b"\0") + attr("pixelAspectRatio", "float", struct.pack("<f", 1.0))
    head += attr("screenWindowCenter", "v2f", struct.pack("<2f", 0.0, 0.0)) + attr("screenWindowWidth", "float", struct.pack("<f", 1.0)) + b"\0"
    line_bytes = 4 * w * 4
    first = len(head) + 8 * h
    with open(path, "wb") as f:
        f.write(head)
        f.write(struct.pack(f"<{h}Q", *[first + y * (8 + line_bytes) for y in range(h)]))
        for y in range(h):
            f.write(struct.pack("<ii", y, line_bytes))
            f.write(np.ascontiguousarray(px[y][:, [3, 2, 1, 0]].T).tobytes())      # A, B, G, R planes of this line
    return path


def read_exr(path: str):
    """Reader for the files write_exr produces (tests): returns (rgba float32 (h, w, 4), w, h)."""
    import struct
    b = open(path, "rb").read()
    assert struct.unpack_from("<II", b, 0) == (20000630, 2)
    pos, attrs = 8, {}
    while b[pos] != 0:
        e = b.index(b"\0", pos); name = b[pos:e].decode(); pos = e + 1
        e = b.index(b"\0", pos); typ = b[pos:e].decode(); pos = e + 1
        (n,) = struct.unpack_from("<i", b, pos); pos += 4
        attrs[name] = (typ, b[pos:pos + n]); pos += n
    pos += 1
    x0, y0, x1, y1 = struct.unpack("<4i", attrs["dataWindow"][1])
    w, h = x1 - x0 + 1, y1 - y0 + 1
    assert attrs["compression"][1] == b"\0"
    offsets = struct.unpack_from(f"<{h}Q", b, pos)
    out = np.empty((h, w, 4), np.float32)
    for off in offsets:
        y, n = struct.unpack_from("<ii", b, off)
        planes = np.frombuffer(b, np.float32, 4 * w, off + 8).reshape(4, w)
        out[y] = planes[[3, 2, 1, 0]].T
    return out, w, h


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("--scene", default="demo", choices=sorted(SCENES))
    ap.add_argument("--size", default="800x600")
    ap.add_argument("--spp", type=int, default=256)
    ap.add_argument("--batch", type=int, default=64, help="samples per device pass (progressive)")
    ap.add_argument("--out", default="render.png")
    ap.add_argument("--precision", default="f32", choices=["f32", "f64"])
    ap.add_argument("--denoise", type=int, default=0, metavar="LEVELS", help="a-trous wavelet levels (0 = off); EXR output only carries the filtered image")
    ap.add_argument("--sigma", type=float, default=0.35, help="range sigma of the denoiser's first level")
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
    if args.out.lower().endswith(".exr"):
        px = pt.denoise(args.denoise, args.sigma) if args.denoise else buf.read_pixels()
        out = write_exr(args.out, px, w, h)                     # the linear running mean, f32
    else:
        frame = np.zeros(w * h * 4, np.uint8)
        pt.convert_to_u8(frame)                   # buffer.rs:55-64 on the device-resident image
        out = write_image(args.out, frame, w, h)
    print(f"{args.scene} {w}x{h} {buf.frames} spp in {dt * 1e3:.1f} ms ({w * h * buf.frames / dt / 1e6:.1f} Msamples/s incl. transfers) -> {out}")
    pt.close()


if __name__ == "__main__":
    main()
