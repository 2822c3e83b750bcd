"""numpy wrappers around the per-function parity entry points of the C ABI (ptb_test_*_f32).
Same array conventions as oracle/pyoracle.py: vec3 = (3, n) float32, scalars = (n,)."""
import ctypes as C

import numpy as np


def _p(a):
    return C.c_void_p(a.ctypes.data)


def _a(x):
    return np.ascontiguousarray(x, np.float32)


class DeviceFns:
    def __init__(self, rp, export, **tracer_kw):
        class _S(rp.Scene):
            def device_export(self_inner):
                return export
        self.rp = rp
        self.tracer = rp.Tracer.new(_S(), **tracer_kw)      # strict=True: the IEEE build of the same device functions
        self.lib = self.tracer._lib
        self.h = self.tracer._handle()
        self.chk = self.tracer._check

    def close(self):
        self.tracer.close()

    def sphere_hit(self, o, d, c, r):
        o, d, c, r = _a(o), _a(d), _a(c), _a(r)
        n = o.shape[1]; t = np.empty(n, np.float32)
        self.chk(self.lib.ptb_test_sphere_hit_f32(self.h, n, _p(o), _p(d), _p(c), _p(r), _p(t)))
        return t

    def plane_hit(self, o, d, p, nn):
        o, d, p, nn = _a(o), _a(d), _a(p), _a(nn)
        n = o.shape[1]; t = np.empty(n, np.float32)
        self.chk(self.lib.ptb_test_plane_hit_f32(self.h, n, _p(o), _p(d), _p(p), _p(nn), _p(t)))
        return t

    def gen_ray(self, p2, off2, w, h):
        p2, off2 = _a(p2), _a(off2)
        n = p2.shape[1]; o = np.empty((3, n), np.float32); d = np.empty((3, n), np.float32)
        self.chk(self.lib.ptb_test_gen_ray_f32(self.h, n, _p(p2), _p(off2), C.c_float(w), C.c_float(h), _p(o), _p(d)))
        return o, d

    def closest_hit(self, o, d, hd):
        o, d, hd = _a(o), _a(d), _a(hd)
        n = o.shape[1]
        hit, em, mat = (np.empty(n, np.uint32) for _ in range(3))
        hdo, lpdf = np.empty(n, np.float32), np.empty(n, np.float32)
        nrm, lem = np.empty((3, n), np.float32), np.empty((3, n), np.float32)
        self.chk(self.lib.ptb_test_closest_hit_f32(self.h, n, _p(o), _p(d), _p(hd), _p(hit), _p(em), _p(hdo), _p(nrm), _p(mat), _p(lpdf), _p(lem)))
        return dict(hit=hit, is_emitter=em, hit_dist=hdo, normal=nrm, material=mat, light_pdf=lpdf, light_emission=lem)

    def sdf_eval(self, q):
        q = _a(q)
        n = q.shape[1]; dist = np.empty(n, np.float32); mat = np.empty(n, np.uint32)
        self.chk(self.lib.ptb_test_sdf_eval_f32(self.h, n, _p(q), _p(dist), _p(mat)))
        return dist, mat

    def sdf_trace(self, o, d, limit):
        o, d, limit = _a(o), _a(d), _a(limit)
        n = o.shape[1]
        t, nrm, mat = np.empty(n, np.float32), np.empty((3, n), np.float32), np.empty(n, np.uint32)
        self.chk(self.lib.ptb_test_sdf_trace_f32(self.h, n, _p(o), _p(d), _p(limit), _p(t), _p(nrm), _p(mat)))
        return dict(t=t, normal=nrm, material=mat)

    def any_hit(self, o, d, md):
        o, d, md = _a(o), _a(d), _a(md)
        n = o.shape[1]; hit = np.empty(n, np.uint32)
        self.chk(self.lib.ptb_test_any_hit_f32(self.h, n, _p(o), _p(d), _p(md), _p(hit)))
        return hit

    def background(self, d):
        d = _a(d); n = d.shape[1]; rgb = np.empty((3, n), np.float32)
        self.chk(self.lib.ptb_test_background_f32(self.h, n, _p(d), _p(rgb)))
        return rgb

    def sample_light(self, li, pos, r1, r2):
        pos, r1, r2 = _a(pos), _a(r1), _a(r2)
        n = pos.shape[1]
        nrm, em, dr = (np.empty((3, n), np.float32) for _ in range(3))
        dist, pdf = np.empty(n, np.float32), np.empty(n, np.float32)
        self.chk(self.lib.ptb_test_sample_light_f32(self.h, n, li, _p(pos), _p(r1), _p(r2), _p(nrm), _p(em), _p(dr), _p(dist), _p(pdf)))
        return dict(normal=nrm, emission=em, direction=dr, dist=dist, pdf=pdf)

    def finalize(self, mi, o, d, hd, nrm):
        o, d, hd, nrm = _a(o), _a(d), _a(hd), _a(nrm)
        n = o.shape[1]
        s = [np.empty(n, np.float32) for _ in range(5)]
        ffn, fhp = np.empty((3, n), np.float32), np.empty((3, n), np.float32)
        self.chk(self.lib.ptb_test_finalize_f32(self.h, n, mi, _p(o), _p(d), _p(hd), _p(nrm), *[_p(x) for x in s], _p(ffn), _p(fhp)))
        return dict(roughness=s[0], clearcoat_roughness=s[1], ax=s[2], ay=s[3], eta=s[4], ffnormal=ffn, fhp=fhp)

    def disney_eval(self, mi, eta, v, nrm, l):
        eta, v, nrm, l = _a(eta), _a(v), _a(nrm), _a(l)
        n = v.shape[1]; f = np.empty((3, n), np.float32); pdf = np.empty(n, np.float32)
        self.chk(self.lib.ptb_test_disney_eval_f32(self.h, n, mi, _p(eta), _p(v), _p(nrm), _p(l), _p(f), _p(pdf)))
        return f, pdf

    def disney_sample(self, mi, eta, v, nrm, lprev, r1, r2, coin):
        eta, v, nrm, lprev, r1, r2, coin = (_a(x) for x in (eta, v, nrm, lprev, r1, r2, coin))
        n = v.shape[1]
        lobe = np.empty(n, np.uint32); l = np.empty((3, n), np.float32); f = np.empty((3, n), np.float32); pdf = np.empty(n, np.float32)
        self.chk(self.lib.ptb_test_disney_sample_f32(self.h, n, mi, _p(eta), _p(v), _p(nrm), _p(lprev), _p(r1), _p(r2), _p(coin),
                                                      _p(lobe), _p(l), _p(f), _p(pdf)))
        return dict(lobe=lobe, l=l, f=f, pdf=pdf)

    def rng(self, pixel, sample, bounce):
        pixel = np.ascontiguousarray(pixel, np.uint32); sample = np.ascontiguousarray(sample, np.uint64)
        n = pixel.size; out = np.empty((8, n), np.float32)
        self.chk(self.lib.ptb_test_rng_f32(self.h, n, _p(pixel), _p(sample), bounce, _p(out)))
        return out

    def convert_to_u8(self, rgba):
        rgba = _a(rgba).reshape(-1); n = rgba.size // 4; out = np.empty(n * 4, np.uint8)
        self.chk(self.lib.ptb_convert_pixels_to_u8_f32(self.h, n, _p(rgba), _p(out)))
        return out
