"""ctypes wrapper of the CPU ORACLE (oracle/libptoracle.so).

TEST INFRASTRUCTURE ONLY — see the header of oracle/pt_oracle.hpp (parity unpinned by reference
tests).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module; nothing under rust_pathtracer_b200/ does.

vec3 arrays are numpy arrays of shape (3, n) (SoA), scalars of shape (n,), both C-contiguous.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libptoracle.so")
_lib = None

COUNTER_FIELDS = ["samples", "closest_hit", "any_hit", "shade", "nee_contrib", "eval_calls", "lobe_diffuse", "lobe_clearcoat",
                  "lobe_reflect", "lobe_refract", "end_sky", "end_emitter", "end_pdf", "end_depth", "ev_diffuse", "ev_reflect",
                  "ev_refract", "ev_clearcoat", "nee_culled", "nee_shadowed", "background", "finalize"]


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in ("oracle_capi.cpp", "pt_oracle.hpp")] + [os.path.join(_HERE, "..", "include", "ptb200.h")]
    if force or not os.path.exists(_LIB) or any(os.path.getmtime(s) > os.path.getmtime(_LIB) for s in srcs if os.path.exists(s)):
        subprocess.check_call(["make", "-C", _HERE, "-s", "libptoracle.so"] + (["-B"] if force else []))
    return _LIB


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            build()
        _lib = C.CDLL(_LIB)
        _lib.pto_max_threads.restype = C.c_int
        _lib.pto_counters_size.restype = C.c_size_t
        for sfx in ("f32", "f64"):
            getattr(_lib, f"pto_scene_literal_{sfx}").restype = C.c_void_p
            getattr(_lib, f"pto_scene_flat_{sfx}").restype = C.c_void_p
            getattr(_lib, f"pto_scene_flat_{sfx}").argtypes = [C.c_void_p]
            getattr(_lib, f"pto_scene_destroy_{sfx}").argtypes = [C.c_void_p]
            getattr(_lib, f"pto_render_{sfx}").restype = C.c_double
            getattr(_lib, f"pto_render_chacha_{sfx}").restype = C.c_double
        assert _lib.pto_counters_size() == 8 * len(COUNTER_FIELDS)
    return _lib


_NP = {"f32": np.float32, "f64": np.float64}
_CT = {"f32": C.c_float, "f64": C.c_double}


def _p(a):
    return C.c_void_p(a.ctypes.data) if a is not None else C.c_void_p(0)


def philox4x32_10(ctr, key):
    lib = load()
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    lib.pto_philox4x32_10(c, k, o)
    return [int(x) for x in o]


class OracleScene:
    """A scene on the oracle side: the literal AnalyticalScene restatement or a FlatScene built
    from the same POD export the device receives."""

    def __init__(self, export=None, precision: str = "f32"):
        self.lib = load()
        self.p = precision
        self.np = _NP[precision]
        self._keep = None
        if export is None:
            self.h = C.c_void_p(getattr(self.lib, f"pto_scene_literal_{precision}")())
        else:
            sc, keep = export.to_c(precision)
            self._keep = (sc, keep)
            self.h = C.c_void_p(getattr(self.lib, f"pto_scene_flat_{precision}")(C.byref(sc)))
            if getattr(export, "sdf", None) is not None and export.sdf.nodes:
                sd, sd_keep = export.sdf.to_c(precision)
                self._keep = (sc, keep, sd, sd_keep)
                assert getattr(self.lib, f"pto_scene_set_sdf_{precision}")(self.h, C.byref(sd)) == 0

    def close(self):
        if self.h:
            getattr(self.lib, f"pto_scene_destroy_{self.p}")(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _fn(self, name):
        return getattr(self.lib, f"pto_{name}_{self.p}")

    def _a(self, x, shape=None):
        a = np.ascontiguousarray(x, dtype=self.np)
        if shape is not None:
            assert a.shape == shape, (a.shape, shape)
        return a

    # -- per-function ---------------------------------------------------------------------------
    def gen_ray(self, p2, off2, w, h):
        p2, off2 = self._a(p2), self._a(off2)
        n = p2.shape[1]
        o, d = np.empty((3, n), self.np), np.empty((3, n), self.np)
        self._fn("gen_ray")(self.h, C.c_size_t(n), _p(p2), _p(off2), _CT[self.p](w), _CT[self.p](h), _p(o), _p(d))
        return o, d

    def closest_hit(self, o, d, hit_dist_in, want_material=False):
        o, d, hd = self._a(o), self._a(d), self._a(hit_dist_in)
        n = o.shape[1]
        hit, em, mat = np.empty(n, np.uint32), np.empty(n, np.uint32), np.empty(n, np.uint32)
        hdo, nrm, lpdf, lem = np.empty(n, self.np), np.empty((3, n), self.np), np.empty(n, self.np), np.empty((3, n), self.np)
        mf = np.empty((n, 17), self.np) if want_material else None
        self._fn("closest_hit")(self.h, C.c_size_t(n), _p(o), _p(d), _p(hd), _p(hit), _p(em), _p(hdo), _p(nrm), _p(mat), _p(lpdf), _p(lem), _p(mf))
        out = dict(hit=hit, is_emitter=em, hit_dist=hdo, normal=nrm, material=mat, light_pdf=lpdf, light_emission=lem)
        if want_material:
            out["material_fields"] = mf
        return out

    def any_hit(self, o, d, max_dist):
        o, d, md = self._a(o), self._a(d), self._a(max_dist)
        n = o.shape[1]
        hit = np.empty(n, np.uint32)
        self._fn("any_hit")(self.h, C.c_size_t(n), _p(o), _p(d), _p(md), _p(hit))
        return hit

    def background(self, d):
        d = self._a(d)
        n = d.shape[1]
        rgb = np.empty((3, n), self.np)
        self._fn("background")(self.h, C.c_size_t(n), _p(d), _p(rgb))
        return rgb

    def sample_light(self, li, pos, r1, r2):
        pos, r1, r2 = self._a(pos), self._a(r1), self._a(r2)
        n = pos.shape[1]
        nrm, em, dr = (np.empty((3, n), self.np) for _ in range(3))
        dist, pdf = np.empty(n, self.np), np.empty(n, self.np)
        self._fn("sample_light")(self.h, C.c_size_t(n), C.c_uint32(li), _p(pos), _p(r1), _p(r2), _p(nrm), _p(em), _p(dr), _p(dist), _p(pdf))
        return dict(normal=nrm, emission=em, direction=dr, dist=dist, pdf=pdf)

    def finalize(self, mi, o, d, hit_dist, normal):
        o, d, hd, nr = self._a(o), self._a(d), self._a(hit_dist), self._a(normal)
        n = o.shape[1]
        s = [np.empty(n, self.np) for _ in range(5)]
        ffn, fhp = np.empty((3, n), self.np), np.empty((3, n), self.np)
        self._fn("finalize")(self.h, C.c_size_t(n), C.c_uint32(mi), _p(o), _p(d), _p(hd), _p(nr), *[_p(x) for x in s], _p(ffn), _p(fhp))
        return dict(roughness=s[0], clearcoat_roughness=s[1], ax=s[2], ay=s[3], eta=s[4], ffnormal=ffn, fhp=fhp)

    def spec_color(self, mi, eta, direction=(0.0, -1.0, 0.0)):
        """get_spec_color + material-only lobe weights of material `mi` (finalized) for `eta`; `direction` = ray direction
        (decides the checker cell of a direction-ratio checker albedo)"""
        d, out = self._a(direction), np.empty(9, self.np)
        fn = self._fn("spec_color")
        fn.restype = None
        fn(self.h, C.c_uint32(mi), (C.c_float if self.p == "f32" else C.c_double)(eta), _p(d), _p(out))
        return dict(spec_col=out[0:3].copy(), sheen_col=out[3:6].copy(), lum=out[6], wd0=out[7], wc0=out[8])

    def disney_eval(self, mi, eta, v, n_, l):
        eta, v, n_, l = self._a(eta), self._a(v), self._a(n_), self._a(l)
        n = v.shape[1]
        f, pdf = np.empty((3, n), self.np), np.empty(n, self.np)
        self._fn("disney_eval")(self.h, C.c_size_t(n), C.c_uint32(mi), _p(eta), _p(v), _p(n_), _p(l), _p(f), _p(pdf))
        return f, pdf

    def disney_sample(self, mi, eta, v, n_, lprev, r1, r2, coin):
        eta, v, n_, lprev, r1, r2, coin = (self._a(x) for x in (eta, v, n_, lprev, r1, r2, coin))
        n = v.shape[1]
        lobe = np.empty(n, np.uint32)
        l, f, pdf = np.empty((3, n), self.np), np.empty((3, n), self.np), np.empty(n, self.np)
        self._fn("disney_sample")(self.h, C.c_size_t(n), C.c_uint32(mi), _p(eta), _p(v), _p(n_), _p(lprev), _p(r1), _p(r2), _p(coin),
                                  _p(lobe), _p(l), _p(f), _p(pdf))
        return dict(lobe=lobe, l=l, f=f, pdf=pdf)

    # -- images ---------------------------------------------------------------------------------
    def render(self, w, h, n_frames, pixels=None, frames=0, sample_base=0, seed=0, threads=0, counters=False):
        """n_frames calls of Tracer::render (1 spp each) into a running-mean buffer.
        Returns (pixels[h*w*4], frames, seconds, counters dict or None)."""
        px = np.zeros(w * h * 4, self.np) if pixels is None else np.ascontiguousarray(pixels, self.np).copy()
        fr = C.c_uint64(frames)
        ctr = (C.c_uint64 * len(COUNTER_FIELDS))() if counters else None
        secs = self._fn("render")(self.h, C.c_uint32(w), C.c_uint32(h), _p(px), C.byref(fr), C.c_uint32(n_frames), C.c_uint64(sample_base),
                                  C.c_uint64(seed), C.c_int(threads), ctr if counters else C.c_void_p(0))
        cd = {k: int(ctr[i]) for i, k in enumerate(COUNTER_FIELDS)} if counters else None
        return px, fr.value, float(secs), cd

    def render_chacha(self, w, h, n_frames, pixels=None, frames=0, seed=0, threads=0):
        """the same frame loop drawing from per-thread ChaCha12 generators in call order (the reference's RNG: rand 0.8.5
        thread_rng) — the CPU baseline closest to the rayon path's cost.  Returns (pixels, frames, seconds)."""
        px = np.zeros(w * h * 4, self.np) if pixels is None else np.ascontiguousarray(pixels, self.np).copy()
        fr = C.c_uint64(frames)
        secs = self._fn("render_chacha")(self.h, C.c_uint32(w), C.c_uint32(h), _p(px), C.byref(fr), C.c_uint32(n_frames), C.c_uint64(seed), C.c_int(threads))
        return px, fr.value, float(secs)

    def trace_samples(self, w, h, px, row, sample, seed=0):
        px = np.ascontiguousarray(px, np.uint32); row = np.ascontiguousarray(row, np.uint32); sample = np.ascontiguousarray(sample, np.uint64)
        n = px.size
        rgb = np.empty((3, n), self.np)
        self._fn("trace_samples")(self.h, C.c_uint32(w), C.c_uint32(h), C.c_size_t(n), _p(px), _p(row), _p(sample), C.c_uint64(seed), _p(rgb))
        return rgb


    def sdf_eval(self, q):
        q = self._a(q)
        n = q.shape[1]
        dist, mat = np.empty(n, self.np), np.empty(n, np.uint32)
        self._fn("sdf_eval")(self.h, C.c_size_t(n), _p(q), _p(dist), _p(mat))
        return dist, mat

    def sdf_trace(self, o, d, limit):
        o, d, limit = self._a(o), self._a(d), self._a(limit)
        n = o.shape[1]
        t, nrm, mat = np.empty(n, self.np), np.empty((3, n), self.np), np.empty(n, np.uint32)
        self._fn("sdf_trace")(self.h, C.c_size_t(n), _p(o), _p(d), _p(limit), _p(t), _p(nrm), _p(mat))
        return dict(t=t, normal=nrm, material=mat)

    def trace_scripted(self, w, h, px, row, draws):
        """radiance of ONE sample of pixel (px, row) traced on a recorded draw sequence (call order, tools/ref_kat);
        returns (rgb, draws consumed) — consumed = -1 if the sequence ran out"""
        d = self._a(draws)
        rgb = np.empty(3, self.np)
        fn = self._fn("trace_scripted")
        fn.restype = C.c_long
        k = fn(self.h, C.c_uint32(w), C.c_uint32(h), C.c_uint32(px), C.c_uint32(row), _p(d), C.c_size_t(d.size), _p(rgb))
        return rgb, int(k)


def chacha_block(key8, counter, stream, rounds):
    lib = load()
    k = (C.c_uint32 * 8)(*key8)
    o = (C.c_uint32 * 16)()
    lib.pto_chacha_block(k, C.c_uint64(counter), C.c_uint64(stream), C.c_int(rounds), o)
    return [int(x) for x in o]


def sample_hg(v, g, r1, r2, precision="f64"):
    """media extension: Henyey-Greenstein direction around v (3, n) and its pdf"""
    lib = load()
    t = _NP[precision]
    v, r1, r2 = (np.ascontiguousarray(x, t) for x in (v, r1, r2))
    n = v.shape[1]
    d, pdf = np.empty((3, n), t), np.empty(n, t)
    getattr(lib, f"pto_sample_hg_{precision}")(C.c_size_t(n), _p(v), _CT[precision](g), _p(r1), _p(r2), _p(d), _p(pdf))
    return d, pdf


def sphere_hit(o, d, c, r, precision="f32"):
    lib = load(); dt = _NP[precision]
    o, d, c, r = (np.ascontiguousarray(x, dt) for x in (o, d, c, r))
    n = o.shape[1]
    t = np.empty(n, dt)
    getattr(lib, f"pto_sphere_hit_{precision}")(C.c_size_t(n), _p(o), _p(d), _p(c), _p(r), _p(t))
    return t


def plane_hit(o, d, p, nn, precision="f32"):
    lib = load(); dt = _NP[precision]
    o, d, p, nn = (np.ascontiguousarray(x, dt) for x in (o, d, p, nn))
    n = o.shape[1]
    t = np.empty(n, dt)
    getattr(lib, f"pto_plane_hit_{precision}")(C.c_size_t(n), _p(o), _p(d), _p(p), _p(nn), _p(t))
    return t


def rng(pixel, sample, bounce, seed=0, precision="f32"):
    lib = load(); dt = _NP[precision]
    pixel = np.ascontiguousarray(pixel, np.uint32); sample = np.ascontiguousarray(sample, np.uint64)
    n = pixel.size
    out = np.empty((8, n), dt)
    getattr(lib, f"pto_rng_{precision}")(C.c_size_t(n), _p(pixel), _p(sample), C.c_uint32(bounce), C.c_uint64(seed), _p(out))
    return out


SCALAR_OPS = {"power_heuristic": 0, "schlick_fresnel": 1, "dielectric_fresnel": 2, "gtr1": 3, "smithg": 4, "gtr2aniso": 5,
              "smithganiso": 6, "cosine_sample_hemisphere": 7, "sample_gtr1": 8, "sample_ggxvndf": 9, "checker": 10, "luminance": 11}


def scalar(op: str, *args, precision="f32"):
    lib = load(); dt = _NP[precision]
    a = np.zeros(8, dt); a[:len(args)] = args
    out = np.zeros(3, dt)
    k = getattr(lib, f"pto_scalar_{precision}")(C.c_int(SCALAR_OPS[op]), _p(a), _p(out))
    assert k > 0
    return float(out[0]) if k == 1 else [float(x) for x in out[:k]]


def convert_to_u8(rgba, precision="f32"):
    lib = load(); dt = _NP[precision]
    rgba = np.ascontiguousarray(rgba, dt).reshape(-1)
    n = rgba.size // 4
    out = np.empty(n * 4, np.uint8)
    getattr(lib, f"pto_convert_to_u8_{precision}")(C.c_size_t(n), _p(rgba), _p(out))
    return out


def convert_to_u8_at(rgba, bw, bh, frame, x, y, fw, fh, precision="f32"):
    lib = load(); dt = _NP[precision]
    rgba = np.ascontiguousarray(rgba, dt).reshape(-1)
    frame = np.ascontiguousarray(frame, np.uint8).copy()
    getattr(lib, f"pto_convert_to_u8_at_{precision}")(_p(rgba), C.c_size_t(bw), C.c_size_t(bh), _p(frame), C.c_size_t(x), C.c_size_t(y),
                                                      C.c_size_t(fw), C.c_size_t(fh))
    return frame


def max_threads() -> int:
    return load().pto_max_threads()
