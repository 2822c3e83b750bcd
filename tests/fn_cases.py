"""Per-function parity cases shared by tests/test_gpu_functions.py (asserts bounds) and tools/function_parity.py (measures
them for the shipped build and for the IEEE measurement build, profiles/r02_function_parity.md).

Every case evaluates one reference function through the C ABI's ptb_test_* entry points and through the CPU oracle on the SAME
seeded inputs and yields per-element relative errors (vectors: against the vector norm).  `floor` is what the oracle's own f32
evaluation differs from its f64 evaluation by on those inputs — the conditioning of the FORMULA, which no f32 implementation
can beat — so the tables separate "the reference's arithmetic is ill-conditioned here" from "the device approximates".
"""
import numpy as np

from conftest import rel_err, unit_vectors, vec_rel_err

N = 100_000


def zoo_export(rp):
    """the demo scene plus a material zoo that exercises every lobe at WELL-conditioned settings:
    3 clearcoat with a wide lobe, 4 rough glass (refraction), 5 anisotropic metal, 6 everything mixed"""
    e = rp.AnalyticalScene.new().device_export()
    M = rp.Material
    e.materials += [
        M(rgb=rp.F3(0.8, 0.3, 0.2), clearcoat=1.0, clearcoat_gloss=0.0, roughness=0.4),
        M(rgb=rp.F3(1.0, 1.0, 1.0), spec_trans=1.0, ior=1.45, roughness=0.3),
        M(rgb=rp.F3(0.9, 0.8, 0.5), metallic=1.0, roughness=0.35, anisotropic=0.7),
        M(rgb=rp.F3(0.6, 0.7, 0.3), metallic=0.4, spec_trans=0.5, roughness=0.5, sheen=0.6, sheen_tint=0.5, subsurface=0.3,
          specular_tint=0.4, clearcoat=0.5, clearcoat_gloss=0.2, anisotropic=0.3, ior=1.3),
    ]
    return e


def rays(rng, n):
    o = rng.uniform(-4, 4, (3, n)).astype(np.float32)
    o[1] = np.abs(o[1]) * 0.75 - 0.9
    d = unit_vectors(rng, n)
    o[:, : n // 2] = np.array([[0], [0], [3]], np.float32)          # half the rays start at the camera like primary rays
    t = rng.uniform(-1, 1, (3, n // 2)); t[2] = -1.2
    d[:, : n // 2] = (t / np.linalg.norm(t, axis=0)).astype(np.float32)
    return o, d


def bsdf_inputs(rng, n):
    nrm = unit_vectors(rng, n)
    v = unit_vectors(rng, n)
    v = np.where((v * nrm).sum(0) < 0, -v, v).astype(np.float32)      # v on the normal's side, as ffnormal guarantees
    l = unit_vectors(rng, n)
    eta = rng.choice([1 / 1.45, 1.45], size=n).astype(np.float32)
    return nrm, v, l, eta


def _f64(*xs):
    return [np.asarray(x, np.float64) for x in xs]


def _vmax_rel(a, b, floor=1e-7):
    """(3, n) colour triples: max abs component error over the largest reference component"""
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return np.abs(a - b).max(0) / np.maximum(np.abs(b).max(0), floor)


def geometry_cases(dev, osc, po):
    """yields (name, err, floor_err or None, note)"""
    rng = np.random.default_rng(2)
    o = rng.uniform(-4, 4, (3, N)).astype(np.float32)
    c = rng.uniform(-2, 2, (3, N)).astype(np.float32)
    r = rng.uniform(0.05, 1.5, N).astype(np.float32)
    d = unit_vectors(rng, N)
    d[:, : N // 2] = ((c - o) / np.linalg.norm(c - o, axis=0) + 0.3 * unit_vectors(rng, N))[:, : N // 2]
    d = (d / np.linalg.norm(d, axis=0)).astype(np.float32)
    ref, got = po.sphere_hit(o, d, c, r), dev.sphere_hit(o, d, c, r)
    l = (c - o).astype(np.float64); tca = (l * d).sum(0); d2 = (l * l).sum(0) - tca * tca
    margin = np.abs(d2 - r.astype(np.float64) ** 2) / r.astype(np.float64) ** 2
    both = (margin > 1e-3) & (ref >= 0) & (got >= 0)
    scale = np.maximum(np.abs(ref), np.abs(tca) + np.sqrt(np.maximum(r.astype(np.float64) ** 2 - d2, 0)))
    yield ("sphere.t", (np.abs(got.astype(np.float64) - ref) / scale)[both], None,
           "analytical.rs:166-190; error over the operands' scale |tca|+thc (t = tca - thc cancels near the surface)")

    rng = np.random.default_rng(3)
    o = rng.uniform(-4, 4, (3, N)).astype(np.float32); d = unit_vectors(rng, N)
    nn2 = unit_vectors(rng, N); p2 = rng.uniform(-1, 1, (3, N)).astype(np.float32)
    ref, got = po.plane_hit(o, d, p2, nn2), dev.plane_hit(o, d, p2, nn2)
    m = (np.abs((nn2 * d).sum(0)) > 2e-4) & (ref > 1e-3) & (got >= 0)
    yield ("plane.t", rel_err(got[m], ref[m]), None, "analytical.rs:193-204, general planes, |n.d| > 2e-4")

    for (w, h) in ((800, 600), (3840, 2160)):
        rng = np.random.default_rng(4)
        n = 50_000
        xs = rng.integers(0, w, n); ys = rng.integers(0, h, n)
        p = np.stack([xs / w, 1.0 - (h - ys) / h]).astype(np.float32)
        off = rng.uniform(0, 1, (2, n)).astype(np.float32)
        ro, rd = osc.gen_ray(p, off, w, h)
        go, gd = dev.gen_ray(p, off, w, h)
        yield (f"gen_ray.dir[{w}x{h}]", vec_rel_err(gd, rd), None, "pinhole.rs:38-61")

    rng = np.random.default_rng(5)
    o, d = rays(rng, N)
    hd = rng.choice([-1.0, 0.3, 2.0, 7.0, 1e30], size=N).astype(np.float32)
    ref, got = osc.closest_hit(o, d, hd), dev.closest_hit(o, d, hd)
    same = (ref["hit"] == got["hit"]) & (ref["is_emitter"] == got["is_emitter"]) & (ref["material"] == got["material"])
    m = same & (ref["hit"] == 1)
    yield ("closest_hit.hit_dist", np.abs(got["hit_dist"][m].astype(np.float64) - ref["hit_dist"][m]) / np.maximum(ref["hit_dist"][m], 1.0), None,
           "analytical.rs:36-127 + scene.rs:36-86; over max(t, 1)")
    geom = m & (ref["material"] != 0xFFFFFFFF)
    yield ("closest_hit.normal", vec_rel_err(got["normal"][:, geom], ref["normal"][:, geom]) / np.maximum(1.0, ref["hit_dist"][geom]), None,
           "(hp - c)/r inherits t's rounding: over max(t, 1)")
    e = m & (ref["is_emitter"] == 1)
    yield ("closest_hit.light_pdf", rel_err(got["light_pdf"][e], ref["light_pdf"][e]), None, "scene.rs:75, ~1/cos(theta)")

    rb, gb = osc.background(d), dev.background(d)
    yield ("background.rgb", rel_err(gb, rb).max(0), None, "analytical.rs:28-32 (powf 2.2)")

    rng = np.random.default_rng(7)
    pos = rng.uniform(-4, 4, (3, N)).astype(np.float32); pos[1] = rng.uniform(-1, 0.5, N)
    r1, r2 = rng.uniform(0, 1, N).astype(np.float32), rng.uniform(0, 1, N).astype(np.float32)
    ref, got = osc.sample_light(0, pos, r1, r2), dev.sample_light(0, pos, r1, r2)
    yield ("sample_light.direction", vec_rel_err(got["direction"], ref["direction"]), None, "tracer.rs:173-220")
    yield ("sample_light.normal", vec_rel_err(got["normal"], ref["normal"]), None, "")
    yield ("sample_light.dist", rel_err(got["dist"], ref["dist"]), None, "")
    well = np.abs((ref["normal"].astype(np.float64) * ref["direction"]).sum(0)) > 0.02
    yield ("sample_light.pdf", rel_err(got["pdf"][well], ref["pdf"][well]), None, "pdf ~ 1/|n.d|, |n.d| > 0.02")

    rng = np.random.default_rng(8)
    n = 20000
    o = rng.uniform(-3, 3, (3, n)).astype(np.float32); d = unit_vectors(rng, n); nrm = unit_vectors(rng, n)
    hd = rng.uniform(0.1, 9, n).astype(np.float32)
    errs = []
    for mi in range(3):
        ref, got = osc.finalize(mi, o, d, hd, nrm), dev.finalize(mi, o, d, hd, nrm)
        for k in ("roughness", "clearcoat_roughness", "ax", "ay", "eta"):
            errs.append(rel_err(got[k], ref[k]))
        errs.append(vec_rel_err(got["fhp"], ref["fhp"]))
    yield ("finalize.*", np.concatenate(errs), None, "globals.rs:50-62 + material.rs:117-131, all outputs, 3 materials")


def eval_case(dev, osc, osc64, mi, n=N, seed=None):
    """disney_eval of material mi: returns dict with f / pdf errors (device vs f32 oracle) and the f32-vs-f64 floor, over the
    well-posed inputs (away from grazing v / l and from the hemisphere boundary, finite reference)."""
    rng = np.random.default_rng((9 + mi) if seed is None else seed)
    nrm, v, l, eta = bsdf_inputs(rng, n)
    rf, rpdf = osc.disney_eval(mi, eta, v, nrm, l)
    gf, gpdf = dev.disney_eval(mi, eta, v, nrm, l)
    df, dpdf = osc64.disney_eval(mi, *_f64(eta, v, nrm, l))
    vz = (v.astype(np.float64) * nrm).sum(0); lz = (l.astype(np.float64) * nrm).sum(0)
    well = (vz > 0.05) & (np.abs(lz) > 0.05) & np.isfinite(rf).all(0) & np.isfinite(rpdf) & (np.abs(df).max(0) > 1e-6)
    out = dict(well=well, n_well=int(well.sum()), lz=lz,
               f=_vmax_rel(gf[:, well], rf[:, well]), pdf=rel_err(gpdf[well], rpdf[well]),
               f_floor=_vmax_rel(rf[:, well], df[:, well], 1e-30), pdf_floor=rel_err(rpdf[well], dpdf[well], 1e-30),
               zero_flips=int(((rpdf == 0) != (gpdf == 0)).sum()), rf=rf)
    return out


def sample_case(dev, osc, osc64, mi, n=N, seed=None, lprev_is_l=False):
    rng = np.random.default_rng((20 + mi) if seed is None else seed)
    nrm, v, lprev, eta = bsdf_inputs(rng, n)
    r1, r2, coin = (rng.uniform(0, 1, n).astype(np.float32) for _ in range(3))
    ref = osc.disney_sample(mi, eta, v, nrm, lprev, r1, r2, coin)
    got = dev.disney_sample(mi, eta, v, nrm, lprev, r1, r2, coin)
    r64 = osc64.disney_sample(mi, *_f64(eta, v, nrm, lprev, r1, r2, coin))
    vz = (v.astype(np.float64) * nrm).sum(0)
    ok = (ref["lobe"] == got["lobe"]) & (ref["lobe"] == r64["lobe"]) & (vz > 0.05) & np.isfinite(ref["pdf"]) & (ref["pdf"] > 0) \
        & np.isfinite(ref["f"]).all(0) & np.isfinite(ref["l"]).all(0)
    lz = (ref["l"].astype(np.float64) * nrm).sum(0)
    well = ok & (np.abs(lz) > 0.05)

    def w_of(x):
        return x["f"][:, well].astype(np.float64) / x["pdf"][well]
    out = dict(ok=ok, well=well, lobe=ref["lobe"], lobe_flips=int((ref["lobe"] != got["lobe"]).sum()),
               l=vec_rel_err(got["l"][:, ok], ref["l"][:, ok]), l_floor=vec_rel_err(ref["l"][:, ok], r64["l"][:, ok]),
               lobe_ok=ref["lobe"][ok], lobe_well=ref["lobe"][well],
               pdf=rel_err(got["pdf"][well], ref["pdf"][well]), pdf_floor=rel_err(ref["pdf"][well], r64["pdf"][well], 1e-30),
               f=_vmax_rel(got["f"][:, well], ref["f"][:, well]), f_floor=_vmax_rel(ref["f"][:, well], r64["f"][:, well], 1e-30),
               w=_vmax_rel(w_of(got), w_of(ref)), w_floor=_vmax_rel(w_of(ref), w_of(r64), 1e-30))
    return out


def stats(e):
    e = np.asarray(e, np.float64)
    if e.size == 0:
        return dict(n=0, max=0.0, p999=0.0, p99=0.0, p95=0.0)
    return dict(n=int(e.size), max=float(e.max()), p999=float(np.percentile(e, 99.9)), p99=float(np.percentile(e, 99)),
                p95=float(np.percentile(e, 95)))
