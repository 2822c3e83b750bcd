"""Replay of tools/ref_kat's known-answer cases through the C++ oracle (test infrastructure, like the oracle itself).

`replay(cases, osc, po)` evaluates every case `{"fn", "in", "out"}` with the oracle function that restates the reference
function of that name and returns, per function, the oracle's outputs in the layout tools/ref_kat/src/main.rs documents.
Used by tests/test_ref_kat.py (comparison against a dump of the reference) and by tools/ref_kat/selfcheck_from_oracle.py
(which WRITES a file in the same format from the oracle, so that the replay code itself is exercised in every CPU run)."""
import numpy as np

F = np.float32
PLANE_P, PLANE_N = (0.0, -1.0, 0.0), (0.0, 1.0, 0.0)       # analytical.rs:193-204


def dec(x):
    return F({"nan": np.nan, "inf": np.inf, "-inf": -np.inf}[x]) if isinstance(x, str) else F(x)


def decode(vals):
    return np.array([dec(v) for v in vals], F)


def enc(x):
    x = float(x)
    if np.isnan(x):
        return "nan"
    if np.isinf(x):
        return "inf" if x > 0 else "-inf"
    return float(np.format_float_scientific(F(x), unique=True))


def col(a):
    return np.asarray(a, F).reshape(-1, 1)


def one(name, a, osc, po):
    """oracle outputs (1-d float32 array) of case `name` on inputs a (1-d float32 array)"""
    if name in po.SCALAR_OPS and name != "checker":
        r = po.scalar(name, *[float(x) for x in a])
        return np.atleast_1d(np.asarray(r, F))
    if name == "gen_ray":
        o, d = osc.gen_ray(col(a[0:2]), col(a[2:4]), float(a[4]), float(a[5]))
        return np.concatenate([o[:, 0], d[:, 0]])
    if name == "sphere":
        return po.sphere_hit(col(a[0:3]), col(a[3:6]), col(a[6:9]), np.array([a[9]], F))
    if name == "plane":
        return po.plane_hit(col(a[0:3]), col(a[3:6]), col(PLANE_P), col(PLANE_N))
    if name == "closest_hit":
        h = osc.closest_hit(col(a[0:3]), col(a[3:6]), np.array([a[6]], F), want_material=True)
        return np.concatenate([[F(h["hit"][0]), F(h["is_emitter"][0]), h["hit_dist"][0]], h["normal"][:, 0], [h["light_pdf"][0]],
                               h["light_emission"][:, 0], h["material_fields"][0]]).astype(F)
    if name == "any_hit":
        return np.array([osc.any_hit(col(a[0:3]), col(a[3:6]), np.array([a[6]], F))[0]], F)
    if name == "background":
        return osc.background(col(a[0:3]))[:, 0]
    if name == "finalize":
        f = osc.finalize(int(a[0]), col(a[1:4]), col(a[4:7]), np.array([a[7]], F), col(a[8:11]))
        return np.concatenate([[f["roughness"][0], f["clearcoat_roughness"][0], f["ax"][0], f["ay"][0], f["eta"][0]], f["ffnormal"][:, 0],
                               f["fhp"][:, 0]]).astype(F)
    if name == "disney_eval":
        f, pdf = osc.disney_eval(int(a[0]), np.array([a[1]], F), col(a[2:5]), col(a[5:8]), col(a[8:11]))
        return np.concatenate([f[:, 0], pdf]).astype(F)
    if name == "disney_sample":
        r = osc.disney_sample(int(a[0]), np.array([a[1]], F), col(a[2:5]), col(a[5:8]), col(a[8:11]), np.array([a[11]], F), np.array([a[12]], F),
                              np.array([a[13]], F))
        consumed = 3 if r["lobe"][0] >= 2 else 2                       # the coin is drawn in the spec lobe only (tracer.rs:534)
        return np.concatenate([r["f"][:, 0], r["l"][:, 0], r["pdf"], [F(consumed)]]).astype(F)
    if name == "sample_light":
        r = osc.sample_light(0, col(a[0:3]), np.array([a[3]], F), np.array([a[4]], F))
        return np.concatenate([r["normal"][:, 0], r["emission"][:, 0], r["direction"][:, 0], r["dist"], r["pdf"]]).astype(F)
    if name == "convert_to_u8":
        return po.convert_to_u8(a[0:4]).astype(F)
    if name == "path_1x1":
        rgb, consumed = osc.trace_scripted(1, 1, 0, 0, a)
        return np.concatenate([rgb, [F(1.0), F(consumed)]]).astype(F)
    raise KeyError(name)


# outputs that are flags / counts / bytes: compared exactly
EXACT = {"closest_hit": (0, 1), "any_hit": (0,), "disney_sample": (7,), "convert_to_u8": (0, 1, 2, 3), "path_1x1": (3, 4)}


def compare(cases, osc, po):
    """per function: (n cases, n bit-identical, max relative error over finite outputs, n mismatching flags)"""
    res = {}
    for c in cases:
        name = c["fn"]
        want, got = decode(c["out"]), one(name, decode(c["in"]), osc, po)
        assert want.shape == got.shape, (name, want.shape, got.shape)
        n, same, worst, flags = res.get(name, (0, 0, 0.0, 0))
        ex = EXACT.get(name, ())
        bad_flag = any(not (want[i] == got[i]) for i in ex)
        rest = [i for i in range(len(want)) if i not in ex]
        w, g = want[rest].astype(np.float64), got[rest].astype(np.float64)
        both_nan = np.isnan(w) & np.isnan(g)
        ident = bool(np.all((want[rest].view(np.uint32) == got[rest].view(np.uint32)) | both_nan)) and not bad_flag
        with np.errstate(invalid="ignore", divide="ignore"):
            # vectors are measured against their largest component (a near-zero component of a unit vector has no relative meaning)
            e = np.abs(w - g) / np.maximum(np.abs(w).max(initial=0.0) if len(w) <= 11 else np.abs(w), 1e-6)
        e = np.where(both_nan | (w == g), 0.0, e)
        e = np.where(np.isnan(e), np.inf, e)                            # NaN on one side only
        res[name] = (n + 1, same + int(ident), max(worst, float(e.max(initial=0.0))), flags + int(bad_flag))
    return res
