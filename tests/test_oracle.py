"""CPU tests that PIN THE ORACLE (oracle/pt_oracle.hpp), the checker every GPU parity test uses.

The reference ships no tests, golden vectors or fixtures, so the oracle is pinned
against: (1) the published Philox4x32-10 known-answer vectors (Random123 kat_vectors) for the RNG;
(2) hand-derived closed-form values of the analytic intersections / BSDF terms (SURVEY.md
Appendix D); (3) the reference's only verification artefact, the screenshot images/spheres.png
(mean colour and a 40x30 block thumbnail, extracted by tests/golden/make_golden.py); (4) committed
oracle-generated fixtures that guard against regressions; (5) bit-equality of the literal
AnalyticalScene restatement with the data-driven FlatScene the device export feeds.
"""
import json
import math
import os

import numpy as np
import pytest

from conftest import unit_vectors

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


# ---- (1) RNG: published Philox4x32-10 known answers -------------------------------------------
def test_philox_known_answers(po):
    assert po.philox4x32_10([0, 0, 0, 0], [0, 0]) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert po.philox4x32_10([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert po.philox4x32_10([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0]) == \
        [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def test_rng_grid_and_layout(po):
    """draws lie on the 2^-24 grid in [0,1) (rand 0.8.5 Standard for f32) and follow the slot layout"""
    pix = np.arange(1000, dtype=np.uint32)
    smp = (np.arange(1000, dtype=np.uint64) * 7919) + (1 << 33)
    u = po.rng(pix, smp, 2, seed=5)
    assert u.shape == (8, 1000) and u.min() >= 0.0 and u.max() < 1.0
    assert np.all(u * 16777216.0 == np.floor(u * 16777216.0))
    # slot s of bounce b = word s%4 of Philox block (2b + s//4), key (pixel, sample lo), ctr (block, sample hi, seed lo, seed hi)
    i = 17
    for s in range(8):
        w = po.philox4x32_10([2 * 2 + s // 4, int(smp[i] >> 32), 5, 0], [int(pix[i]), int(smp[i] & 0xFFFFFFFF)])[s % 4]
        assert u[s, i] == np.float32((w >> 8) * 2.0 ** -24)
    assert 0.48 < u.mean() < 0.52
    u64 = po.rng(pix, smp, 1, precision="f64")
    assert np.all(u64 * 2.0 ** 53 == np.floor(u64 * 2.0 ** 53)) and u64.max() < 1.0


# ---- (2) hand-derived known answers -----------------------------------------
def test_scalar_known_answers(po):
    A = lambda x, y, tol=2e-6: abs(x - y) <= tol * max(1.0, abs(y))
    assert A(po.scalar("power_heuristic", 2, 1), 0.8) and po.scalar("power_heuristic", 0, 5) == 0.0
    assert A(po.scalar("schlick_fresnel", 0.3), 0.7 ** 5)
    # Fresnel at normal incidence: ((1-eta)/(1+eta))^2 with eta = 1/1.45
    e = 1 / 1.45
    assert A(po.scalar("dielectric_fresnel", 1.0, e), ((1 - e) / (1 + e)) ** 2)
    assert A(po.scalar("dielectric_fresnel", 0.5, 1 / 1.5), 0.0891867)
    assert po.scalar("dielectric_fresnel", 0.2, 1.45) == 1.0                      # total internal reflection
    # gtr1 with log2 (quirk A.3): (a2-1)/(pi*log2(a2)*t) at ndoth=1, a=0.1 -> t = a2
    a2 = 0.01
    assert A(po.scalar("gtr1", 1.0, 0.1), (a2 - 1) / (math.pi * math.log2(a2) * a2), 1e-5)
    assert not A(po.scalar("gtr1", 1.0, 0.1), (a2 - 1) / (math.pi * math.log(a2) * a2), 1e-2)   # NOT the natural log
    assert A(po.scalar("gtr1", 0.3, 1.0), 1 / math.pi)
    assert A(po.scalar("smithg", 0.5, 0.25), 1.0 / (0.5 + math.sqrt(0.0625 + 0.25 - 0.0625 * 0.25)))
    c = (0.3 / 0.1) ** 2 + (0.2 / 0.2) ** 2 + 0.81
    assert A(po.scalar("gtr2aniso", .9, .3, .2, .1, .2), 1 / (math.pi * 0.1 * 0.2 * c * c))
    assert A(po.scalar("smithganiso", .7, .5, .5, .1, .2), 1.4 / (0.7 + math.sqrt(0.0025 + 0.01 + 0.49)))
    x, y, z = po.scalar("cosine_sample_hemisphere", 0.25, 0.5)
    assert A(x, -0.5) and abs(y) < 1e-6 and A(z, math.sqrt(0.75))
    g = po.scalar("sample_gtr1", 0.1, 0.3, 0.777)
    assert g == po.scalar("sample_gtr1", 0.1, 0.3, 0.123)                         # r2 is ignored (quirk A.4)
    assert np.allclose(g, [-0.0536230, 0.1650346, 0.9848290], atol=2e-6)
    assert np.allclose(po.scalar("sample_ggxvndf", .6, 0, .8, .2, .2, .3, .6), [0.1100154, -0.1117582, 0.9876268], atol=2e-6)
    assert po.scalar("checker", 100.2, 100.7, .25, .1) == np.float32(0.25) and po.scalar("checker", 101.2, 100.7, .25, .1) == np.float32(0.1)
    assert A(po.scalar("luminance", 1, 1, 1), 1.0)


def test_intersections_known_answers(po):
    o = np.array([[0, 0, 0], [0, 0, 0], [3, 0, 3]], np.float32)          # columns: (0,0,3), (0,0,0), (0,0,3)
    d = np.array([[0, 0, 0], [0, 0, 0], [-1, -1, 1]], np.float32)
    c = np.zeros((3, 3), np.float32)
    t = po.sphere_hit(o, d, c, np.ones(3, np.float32))
    assert t[0] == 2.0 and t[1] == 1.0 and t[2] == -1.0                    # front hit / from centre / pointing away
    # plane y=-1 from (0,0,3) along norm(0,-1,-1): t = sqrt(2); |d.y| <= 1e-4: None
    s = np.float32(1 / math.sqrt(2))
    o = np.array([[0, 0], [0, 0], [3, 3]], np.float32)
    d = np.array([[0, math.sqrt(1 - 2.5e-9)], [-s, 5e-5], [-s, 0]], np.float32)
    p = np.array([[0, 0], [-1, -1], [0, 0]], np.float32)
    n = np.array([[0, 0], [1, 1], [0, 0]], np.float32)
    t = po.plane_hit(o, d, p, n)
    assert abs(t[0] - math.sqrt(2)) < 1e-6 and t[1] == -1.0
    # a plane behind the ray is a miss (t < 0)
    assert po.plane_hit(o[:, :1], -d[:, :1], p[:, :1], n[:, :1])[0] == -1.0


def test_scene_known_answers(oracle_literal):
    sc = oracle_literal
    # pinhole: fov 80 horizontal; centre ray of an 800x600 frame and the lower-left corner
    o, d = sc.gen_ray(np.array([[.5, 0], [.5, 0]]), np.array([[.5, 0], [.5, 0]]), 800, 600)
    assert np.allclose(o.T, [[0, 0, 3]] * 2)
    assert np.allclose(d[:, 0], [0.0010489, 0.0010489, -0.9999989], atol=2e-6)
    hw = math.tan(math.radians(40)); hh = hw * 0.75
    ll = np.array([-hw, -hh, -1]); ll /= np.linalg.norm(ll)
    assert np.allclose(d[:, 1], ll, atol=2e-6)
    # background: to_linear(lerp(white, (0.5,0.7,1), t)) * 0.5
    bg = sc.background(np.array([[0, 1], [1, 0], [0, 0]], np.float32))
    assert np.allclose(bg[:, 0], [0.5 ** 2.2 * 0.5, 0.7 ** 2.2 * 0.5, 0.5], atol=1e-6)
    assert np.allclose(bg[:, 1], [0.75 ** 2.2 * 0.5, 0.85 ** 2.2 * 0.5, 0.5], atol=1e-6)
    # light sampling (Appendix D): demo light from (0,-1,0), draws (0.25, 0.5)
    sl = sc.sample_light(0, np.array([[0], [-1], [0]]), [0.25], [0.5])
    assert np.allclose(sl["direction"][:, 0], [0.4742712, 0.7755651, 0.4166121], atol=2e-6)
    assert abs(sl["dist"][0] - 4.5447544) < 1e-5 and abs(sl["pdf"][0] - 86.556696) < 2e-3
    assert np.allclose(sl["normal"][:, 0], [-0.8445537, 0.5247527, -0.1066004], atol=2e-6)
    assert np.allclose(sl["emission"][:, 0], [3, 3, 3])
    # pdf identity: dist^2 / (area * 0.5 * |n.d|), area = 4*pi
    nd = abs(float((sl["normal"][:, 0] * sl["direction"][:, 0]).sum()))
    assert abs(sl["pdf"][0] - sl["dist"][0] ** 2 / (4 * math.pi * 0.5 * nd)) < 1e-3


def test_closest_hit_quirks(oracle_literal):
    sc = oracle_literal
    # primary ray through the metal sphere's centre: t = |o-c| - 1
    o = np.array([[0.0], [0.0], [3.0]]); tgt = np.array([[-1.1], [0.0], [0.0]])
    d = (tgt - o) / np.linalg.norm(tgt - o)
    h = sc.closest_hit(o, d, [-1.0], want_material=True)
    assert h["hit"][0] == 1 and h["material"][0] == 0 and abs(h["hit_dist"][0] - (math.sqrt(1.21 + 9) - 1)) < 1e-6
    assert np.allclose(h["normal"][:, 0], -d[:, 0], atol=1e-6)
    mf = h["material_fields"][0]
    assert mf[7] == 1.0 and abs(mf[8] - 0.05) < 1e-7 and np.allclose(mf[:3], 1.0)      # metallic, roughness, rgb
    # A.1: a ray that hits only the light with the initial hit_dist = -1 reports NO hit
    o = np.array([[0.0], [2.0], [2.0]]); d = np.array([[1.0], [0.0], [0.0]])
    assert sc.closest_hit(o, d, [-1.0])["hit"][0] == 0
    h = sc.closest_hit(o, d, [5.0])                      # stale hit_dist of 5 lets the light (t=2) win
    assert h["hit"][0] == 1 and h["is_emitter"][0] == 1 and abs(h["hit_dist"][0] - 2.0) < 1e-6
    assert abs(h["light_pdf"][0] - 4.0 / (4 * math.pi * 1.0 * 0.5)) < 1e-5            # d^2/(area*cos*0.5), cos=1
    assert sc.closest_hit(o, d, [1.5])["hit"][0] == 0    # stale 1.5 < 2: the light is "behind" the stale distance
    # stale material: a ray that passes the orange sphere first and the metal sphere behind it keeps metallic=1
    o = np.array([[4.0], [0.0], [0.0]]); d = np.array([[-1.0], [0.0], [0.0]])
    h = sc.closest_hit(o, d, [-1.0], want_material=True)
    mf = h["material_fields"][0]
    assert h["material"][0] == 1 and abs(h["hit_dist"][0] - 1.9) < 1e-6
    assert mf[7] == 1.0 and mf[13] == 1.0 and abs(mf[8] - 0.1) < 1e-7 and abs(mf[1] - 0.186) < 1e-7
    # ... whereas the opposite direction (metal first) is a plain metal hit, clearcoat 0
    h = sc.closest_hit(-o, -d, [-1.0], want_material=True)
    assert h["material"][0] == 0 and h["material_fields"][0][13] == 0.0
    # checker is a function of the ray DIRECTION ratios (A.7): same hit point, different direction -> may differ
    assert sc.any_hit(np.array([[0.0], [5.0], [0.0]]), np.array([[0.0], [1.0], [0.0]]), [1e-3])[0] == 0   # up: plane t<0
    assert sc.any_hit(np.array([[0.0], [5.0], [0.0]]), np.array([[0.0], [-1.0], [0.0]]), [1e-3])[0] == 1  # ignores max_dist (A.8)


def test_bsdf_known_answers(oracle_demo):
    sc = oracle_demo
    nrm = np.array([[0.0], [0.0], [1.0]])
    v = np.array([[.3], [.1], [.9]]); v /= np.linalg.norm(v)
    l = np.array([[-.2], [.3], [.8]]); l /= np.linalg.norm(l)
    eta = [1 / 1.45]
    f, pdf = sc.disney_eval(1, eta, v, nrm, l)                      # orange clearcoat sphere
    assert np.allclose(f[:, 0], [0.3044437, 0.0682228, 0.0142460], rtol=2e-5) and abs(pdf[0] - 0.1978416) < 4e-6
    mirror = np.array([[-v[0, 0]], [-v[1, 0]], [v[2, 0]]])
    f, pdf = sc.disney_eval(0, eta, v, nrm, mirror)                 # metal sphere, mirror direction
    assert np.allclose(f[:, 0], 33.73349, rtol=3e-5) and abs(pdf[0] - 33.73609) < 2e-3
    s = sc.disney_sample(1, eta, v, nrm, np.zeros((3, 1)), [0.1], [0.7], [0.4])
    assert s["lobe"][0] == 0 and np.allclose(s["l"][:, 0], [-0.4162624, 0.1352518, 0.8991288], atol=3e-6)
    assert np.allclose(s["f"][:, 0], [0.2862006, 0.0532333, 0.0], atol=3e-6) and abs(s["pdf"][0] - 0.1493999) < 3e-6
    s = sc.disney_sample(1, eta, v, nrm, np.zeros((3, 1)), [0.9], [0.7], [0.4])
    assert s["lobe"][0] == 2 and np.allclose(s["l"][:, 0], [-0.2950880, -0.1026423, 0.9499409], atol=3e-6)
    assert np.allclose(s["f"][:, 0], 0.5490972, rtol=2e-5) and abs(s["pdf"][0] - 0.8296929) < 2e-5
    # finalize: orange -> clearcoat_roughness mix(0.1, 0.001, gloss=1) = 0.001, ax = ay = roughness
    fin = sc.finalize(1, np.zeros((3, 1)), -nrm, [1.0], nrm)
    assert abs(fin["clearcoat_roughness"][0] - 0.001) < 1e-8 and abs(fin["ax"][0] - 0.1) < 1e-8 and abs(fin["eta"][0] - 1 / 1.45) < 1e-7
    assert np.allclose(fin["ffnormal"][:, 0], [0, 0, 1]) and np.allclose(fin["fhp"][:, 0], [0, 0, -1])


def test_bsdf_energy_sanity(oracle_demo):
    """the eval pdf integrates to about one over the sphere (not exactly: disney_eval re-weights the
    lobes with a direction-dependent Fresnel, tracer.rs:596-597) and covers the sampled lobe's pdf"""
    sc = oracle_demo
    rng = np.random.default_rng(3)
    n = 200000
    l = unit_vectors(rng, n)
    nrm = np.tile(np.array([[0.0], [0.0], [1.0]], np.float32), (1, n))
    v = np.tile(np.array([[0.5], [0.0], [math.sqrt(0.75)]], np.float32), (1, n))
    eta = np.full(n, 1 / 1.45, np.float32)
    _, pdf = sc.disney_eval(2, eta, v, nrm, l)                        # rough plane: smooth pdf
    integral = float(pdf.astype(np.float64).mean() * 4 * math.pi)
    assert 0.6 < integral < 1.03, integral          # measured 0.815
    # eval at the sampled direction reproduces the sampled lobe's pdf share (diffuse-only when l from diffuse)
    r1, r2, coin = (rng.uniform(0, 1, 2000).astype(np.float32) for _ in range(3))
    s = sc.disney_sample(2, eta[:2000], v[:, :2000], nrm[:, :2000], np.zeros((3, 2000), np.float32), r1, r2, coin)
    f, pdf = sc.disney_eval(2, eta[:2000], v[:, :2000], nrm[:, :2000], s["l"])
    ok = s["pdf"] > 0
    assert np.all(pdf[ok] >= s["pdf"][ok] * (1 - 1e-3))               # total pdf >= the sampled lobe's share


# ---- (3) the reference's screenshot ----------------------------------------------------------------
def test_render_matches_reference_screenshot(po, oracle_literal):
    g = json.load(open(os.path.join(GOLD, "screenshot_means.json")))
    px, frames, _, _ = oracle_literal.render(800, 600, 16)
    assert frames == 16
    u8 = po.convert_to_u8(px).reshape(600, 800, 4)[..., :3].astype(np.float64)
    mean = u8.reshape(-1, 3).mean(0)
    assert np.all(np.abs(mean - np.array(g["mean_srgb8"])) < 0.03 * np.array(g["mean_srgb8"])), mean   # within 3 %
    thumb = u8.reshape(30, 20, 40, 20, 3).mean(axis=(1, 3))
    d = np.abs(thumb - np.array(g["thumb_40x30_srgb8"]))
    assert d.mean() < 5.0 and d.max() < 25.0, (d.mean(), d.max())     # measured: 3.2 mean, 14 max (display colour management)
    img = px.reshape(600, 800, 4)
    assert np.all(img[..., 3] == 1.0)                                   # alpha is exactly 1 (tracer.rs:59,105)
    lum = 0.212671 * img[..., 0] + 0.715160 * img[..., 1] + 0.072169 * img[..., 2]
    assert abs(lum.mean() - 0.1639) < 0.002                            # SURVEY.md Appendix D (independent f64 model)


# ---- (4) regression fixtures ------------------------------------------------------------------------
def test_oracle_matches_committed_fixtures(po, oracle_demo):
    k = np.load(os.path.join(GOLD, "oracle_kat_f32.npz"))
    sc = oracle_demo
    ch = sc.closest_hit(k["ch_o"], k["ch_d"], k["ch_hd"], want_material=True)
    for name, v in ch.items():
        assert np.array_equal(v, k[f"ch_{name}"], equal_nan=True), name
    assert np.array_equal(sc.any_hit(k["ch_o"], k["ch_d"], np.full(64, 3.0, np.float32)), k["ah_hit"])
    assert np.array_equal(sc.background(k["ch_d"]), k["bg"])
    o, d = sc.gen_ray(k["gr_p2"], k["gr_off"], 800, 600)
    assert np.array_equal(o, k["gr_o"]) and np.array_equal(d, k["gr_d"])
    sl = sc.sample_light(0, k["ch_o"], k["sl_r1"], k["sl_r2"])
    for name, v in sl.items():
        assert np.array_equal(v, k[f"sl_{name}"]), name
    for mi in range(3):
        f, pdf = sc.disney_eval(mi, k["bs_eta"], k["bs_v"], k["bs_n"], k["bs_l"])
        assert np.array_equal(f, k[f"ev{mi}_f"], equal_nan=True) and np.array_equal(pdf, k[f"ev{mi}_pdf"], equal_nan=True)
        s = sc.disney_sample(mi, k["bs_eta"], k["bs_v"], k["bs_n"], k["bs_l"], k["sl_r1"], k["sl_r2"], k["bs_coin"])
        assert np.array_equal(s["lobe"], k[f"sm{mi}_lobe"]) and np.array_equal(s["l"], k[f"sm{mi}_l"], equal_nan=True)
        assert np.array_equal(s["f"], k[f"sm{mi}_f"], equal_nan=True) and np.array_equal(s["pdf"], k[f"sm{mi}_pdf"], equal_nan=True)
    px, _, _, _ = sc.render(32, 24, 2)
    assert np.array_equal(px, k["img_32x24_2spp"])
    assert np.array_equal(po.rng(np.arange(8, dtype=np.uint32), np.arange(8, dtype=np.uint64) * 1000003, 0), k["rng_b0"])
    assert np.array_equal(po.rng(np.arange(8, dtype=np.uint32), np.arange(8, dtype=np.uint64) * 1000003, 3), k["rng_b3"])


# ---- (5) literal restatement == data-driven export ---------------------------------------------------
def test_flat_export_equals_literal_scene(oracle_literal, oracle_demo):
    rng = np.random.default_rng(11)
    n = 100000
    o = rng.uniform(-4, 4, size=(3, n)).astype(np.float32)
    o[1] = np.abs(o[1]) * 0.75 - 0.9                                   # mostly above the plane, some skimming it
    d = unit_vectors(rng, n)
    hd = rng.choice([-1.0, 0.3, 2.0, 7.0, 1e30], size=n).astype(np.float32)
    a = oracle_literal.closest_hit(o, d, hd, want_material=True)
    b = oracle_demo.closest_hit(o, d, hd, want_material=True)
    for k in a:
        assert np.array_equal(a[k], b[k], equal_nan=True), k
    # the stale-material case must be exercised by the sample (orange hit with metallic == 1)
    stale = (a["material"] == 1) & (a["material_fields"][:, 7] == 1.0)
    assert stale.sum() > 10
    md = rng.uniform(0, 5, n).astype(np.float32)
    assert np.array_equal(oracle_literal.any_hit(o, d, md), oracle_demo.any_hit(o, d, md))
    assert np.array_equal(oracle_literal.background(d), oracle_demo.background(d))
    pa, fa, _, ca = oracle_literal.render(96, 64, 3, counters=True)
    pb, fb, _, cb = oracle_demo.render(96, 64, 3, counters=True)
    assert np.array_equal(pa, pb) and fa == fb == 3 and ca == cb


def test_running_mean_and_sample_base(oracle_demo):
    """render() accumulates a running mean (tracer.rs:105-117); sample_base shifts the RNG stream"""
    p2, f2, _, _ = oracle_demo.render(48, 32, 2)
    p1, f1, _, _ = oracle_demo.render(48, 32, 1)
    pb, fb, _, _ = oracle_demo.render(48, 32, 1, sample_base=1)       # the second sample alone
    assert f2 == 2 and f1 == 1
    assert np.allclose(p2, 0.5 * p1 + 0.5 * pb, rtol=1e-6, atol=1e-7)
    pc, fc, _, _ = oracle_demo.render(48, 32, 1, pixels=p1, frames=1)  # resume from (pixels, frames)
    assert fc == 2 and np.array_equal(pc, p2)


def test_counters_call_rates(oracle_demo):
    """demo-scene call rates of SURVEY.md Appendix C (independent f64 model): 2.04 closest_hit, 0.91 any_hit per sample"""
    _, _, _, c = oracle_demo.render(200, 150, 8, counters=True)
    s = c["samples"]
    assert s == 200 * 150 * 8
    assert abs(c["closest_hit"] / s - 2.038) < 0.02 and abs(c["any_hit"] / s - 0.908) < 0.02
    assert abs(c["shade"] / s - 1.187) < 0.02 and abs(c["end_sky"] / s - 0.844) < 0.01
    assert c["end_sky"] + c["end_emitter"] + c["end_pdf"] + c["end_depth"] == s
    assert c["lobe_diffuse"] + c["lobe_clearcoat"] + c["lobe_reflect"] + c["lobe_refract"] == c["shade"]
    assert c["lobe_refract"] == 0


def test_convert_to_u8_edges(po):
    x = np.array([[0.0, 1.0, 0.5, 1.0], [np.nan, -1.0, 2.0, 0.999], [1e-8, 0.2176, 0.5, 0.0]], np.float32)
    out = po.convert_to_u8(x).reshape(3, 4)
    assert out[0].tolist() == [0, 255, int(np.float32(0.5) ** np.float32(0.4545) * 255), 255]
    assert out[1].tolist() == [0, 0, 255, 254]          # NaN -> 0, negative -> NaN powf -> 0, saturate, alpha truncates
    assert out[2, 3] == 0


def test_convert_to_u8_at_bounds(po):
    """buffer.rs:67-102: no gamma, strict `>` bounds: column at.0 and row at.1 are never written"""
    bw, bh, fw, fh = 4, 3, 8, 6
    px = np.full(bw * bh * 4, 0.5, np.float32)
    frame = po.convert_to_u8_at(px, bw, bh, np.zeros(fw * fh * 4, np.uint8), 1, 1, fw, fh).reshape(fh, fw, 4)
    written = frame[..., 0] == 127
    # x in (1, 5) -> columns 2,3,4; y in (1, 4) -> y = fh - j = 2,3 -> frame rows fh-1-j: j = 4,3 -> rows 1,2
    assert written.sum() == 3 * 2 and written[1:3, 2:5].all()


def test_chacha_block_rfc7539_vector(po):
    """the block function behind the baseline's thread_rng stand-in, pinned by RFC 7539 section 2.3.2 (ChaCha20: same quarter
    round and layout, 20 rounds; words 12..15 = counter 1, nonce 00000009 0000004a 00000000)"""
    import struct
    key = list(struct.unpack("<8I", bytes(range(32))))
    out = po.chacha_block(key, 1 | (0x09000000 << 32), 0x4A000000, 20)
    assert out == [0xE4E7F110, 0x15593BD1, 0x1FDD0F50, 0xC47120A3, 0xC7F4D1C7, 0x0368C033, 0x9AAA2204, 0x4E6CD4C3,
                   0x466482D2, 0x09AA9F07, 0x05D7C214, 0xA2028BD9, 0xD19C12B5, 0xB94E16DE, 0xE883D0CB, 0x4E3C50A2]
    assert po.chacha_block(key, 1, 0, 12) != po.chacha_block(key, 2, 0, 12)


def test_chacha_render_converges_to_the_counter_rng_image(po, oracle_demo):
    """the ChaCha12 call-order frame loop (bench.py's CPU baseline) renders the same image as the counter-RNG loop"""
    W, H, S = 96, 54, 48
    a, fa, _ = oracle_demo.render_chacha(W, H, S, seed=7)
    b, fb, _, _ = oracle_demo.render(W, H, S)
    assert fa == fb == S
    la, lb = a.reshape(-1, 4)[:, :3].mean(), b.reshape(-1, 4)[:, :3].mean()
    assert abs(la / lb - 1) < 0.01
    assert np.all(a.reshape(-1, 4)[:, 3] == 1.0)
    # per-pixel: two independent 48-spp estimates of the same image
    d = (a - b).reshape(-1, 4)[:, :3]
    assert np.sqrt((d ** 2).mean()) < 0.08


# ---- extended light kinds (PTB_SCENE_EXTENDED_LIGHTS): the library's extension, pinned by closed forms ------------------------
def _quad_scene(rp, a=2.0, b=1.2, h=1.5):
    """a quad of a x b, h above the origin, centred, facing down; plus a distant light"""
    e = rp.AnalyticalScene.new().device_export()
    e.lights = [rp.AnalyticalLight.rectangular(rp.F3(-a / 2, h, -b / 2), rp.F3(a, 0.0, 0.0), rp.F3(0.0, 0.0, b), rp.F3(2.0, 3.0, 4.0)),
                rp.AnalyticalLight.distant(rp.F3(0.0, 2.0, 0.0), rp.F3(1.0, 1.0, 1.0))]
    e.flags |= rp._abi.PTB_SCENE_EXTENDED_LIGHTS
    return e


def test_quad_light_sampling_integrates_to_the_form_factor(rp, po):
    """Irradiance of a uniform quad on a parallel element under its centre has a closed form (four corner form factors):
    F_corner = 1/(2 pi) [X/sqrt(1+X^2) atan(Y/sqrt(1+X^2)) + Y/sqrt(1+Y^2) atan(X/sqrt(1+Y^2))], X = a/h, Y = b/h.
    The estimator the tracer uses, mean(L cos / pdf) over sample_light draws, must converge to pi L F."""
    a, b, h = 2.0, 1.2, 1.5
    orc = po.OracleScene(_quad_scene(rp, a, b, h), "f64")
    n = 400_000
    rng = np.random.default_rng(3)
    r1, r2 = rng.uniform(0, 1, n), rng.uniform(0, 1, n)
    pos = np.zeros((3, n))
    ls = orc.sample_light(0, pos, r1, r2)
    assert np.allclose(ls["normal"], np.array([[0.0], [-1.0], [0.0]]))
    assert np.allclose(ls["emission"], np.array([[4.0], [6.0], [8.0]]))      # number_of_lights() x emission (tracer.rs:214)
    cos_surface = ls["direction"][1]                                        # element normal (0, 1, 0)
    est = (cos_surface / ls["pdf"]).mean()

    def corner(x, y):
        X, Y = x / h, y / h
        return (X / math.sqrt(1 + X * X) * math.atan(Y / math.sqrt(1 + X * X)) + Y / math.sqrt(1 + Y * Y) * math.atan(X / math.sqrt(1 + Y * Y))) / (2 * math.pi)
    want = math.pi * 4 * corner(a / 2, b / 2)
    assert abs(est / want - 1) < 5e-3, (est, want)
    # the sampled points lie on the quad and dist is their distance
    p = ls["direction"] * ls["dist"]
    assert np.allclose(p[1], h) and p[0].min() >= -a / 2 - 1e-9 and p[0].max() <= a / 2 + 1e-9 and np.abs(p[2]).max() <= b / 2 + 1e-9
    # the distant light: fixed direction, pdf 1, infinitely far
    ds = orc.sample_light(1, pos[:, :8], r1[:8], r2[:8])
    assert np.allclose(ds["direction"], np.array([[0.0], [1.0], [0.0]])) and np.all(ds["pdf"] == 1.0) and np.all(ds["dist"] > 1e300)


def test_quad_light_hit_pdf_is_the_sampling_pdf(rp, po):
    """A ray that hits the quad reports pdf = t^2 / (area cos), the density sample_light would have assigned to that direction —
    what the MIS weight at tracer.rs:84 needs; from behind the quad does not exist."""
    a, b, h = 2.0, 1.2, 1.5
    e = _quad_scene(rp, a, b, h)
    e.planes[0].point = rp.F3(0.0, 5.0, 0.0)               # the ground plane becomes a ceiling behind the quad: rays find a hit_dist
    e.planes[0].normal = rp.F3(0.0, -1.0, 0.0)
    e.spheres = []
    orc = po.OracleScene(e, "f64")
    o = np.array([[0.3, 0.3], [-2.0, 4.0], [0.1, 0.1]]); d = np.array([[0.0, 0.0], [1.0, -1.0], [0.0, 0.0]])
    o[1, 0] = -2.0
    tgt = np.array([0.5, h, -0.2])
    d[:, 0] = (tgt - o[:, 0]) / np.linalg.norm(tgt - o[:, 0])
    r = orc.closest_hit(o, d, np.array([-1.0, -1.0]))
    t = np.linalg.norm(tgt - o[:, 0])
    assert r["is_emitter"][0] == 1 and abs(r["hit_dist"][0] - t) < 1e-12
    assert abs(r["light_pdf"][0] / (t * t / (a * b * d[1, 0])) - 1) < 1e-12
    assert r["is_emitter"][1] == 0                          # from above: the back side
    ls = orc.sample_light(0, o[:, :1], np.array([(0.5 + a / 2) / a]), np.array([(-0.2 + b / 2) / b]))
    assert np.allclose(ls["direction"][:, 0], d[:, 0]) and abs(ls["pdf"][0] / r["light_pdf"][0] - 1) < 1e-12


# ---- media (PTB_MEDIUM_*): the library's extension, pinned by closed forms ----------------------------------------------------
def test_henyey_greenstein_sampling_matches_its_pdf(po):
    """sample_hg draws from phase_hg: directions are unit, E[1 / pdf] = 4 pi (the pdf integrates to one over the sphere),
    E[cos] = -g in the pbrt convention (v points back along the ray: forward scattering is cos = -1), and g = 0 is uniform."""
    rng = np.random.default_rng(11)
    n = 400_000
    v = np.ascontiguousarray(np.tile(np.array([[0.3], [-0.5], [0.81]]) / np.linalg.norm([0.3, -0.5, 0.81]), (1, n)))
    r1, r2 = rng.uniform(0, 1, n), rng.uniform(0, 1, n)
    for g in (-0.6, 0.0, 0.0005, 0.4, 0.9):
        d, pdf = po.sample_hg(v, g, r1, r2)
        assert np.allclose(np.linalg.norm(d, axis=0), 1.0, atol=1e-12)
        cos = (d * v).sum(0)
        assert abs(cos.mean() + g) < 4e-3, (g, cos.mean())
        assert abs((1.0 / pdf).mean() / (4 * math.pi) - 1) < (2e-2 if abs(g) > 0.8 else 5e-3), (g, (1.0 / pdf).mean())
        if g == 0.0:
            assert np.allclose(pdf, 1 / (4 * math.pi))


def _ball_in_white_sky(rp, medium, depth, ior=1.0002):
    """one glass ball of radius 1 in a constant white sky, no lights, camera looking at its centre"""
    e = rp.AnalyticalScene.new().device_export()
    g = rp.Material(); g.rgb = rp.F3(1.0, 1.0, 1.0); g.spec_trans = 1.0; g.roughness = 0.01; g.ior = ior; g.medium = medium
    e.materials = [g]
    e.spheres = [rp.Sphere(rp.F3(0.0, 0.0, 0.0), 1.0, 0)]
    e.planes = []; e.lights = []
    e.background = rp.Background(rp._abi.PTB_BG_CONSTANT, rp.F3(1.0, 1.0, 1.0), rp.F3(1.0, 1.0, 1.0), 1.0, 1.0)
    e.depth = depth; e.flags = 0
    e.camera.set_fov(60.0)                 # half width 1.73 at the ball: the corners see the sky directly
    return e


def test_absorbing_medium_follows_beer_lambert(rp, po):
    """straight through the centre of a ball with ior ~ 1: radiance = exp(-(1 - color) * 2 r * density) per channel"""
    den, col = 0.8, (0.9, 0.5, 0.2)
    e = _ball_in_white_sky(rp, rp.Medium(rp.MediumType.ABSORB, den, rp.F3(*col), 0.0), depth=6)
    img, _, _, _ = po.OracleScene(e, "f64").render(65, 65, 64)
    centre = img.reshape(65, 65, 4)[31:34, 31:34, :3].mean((0, 1))
    want = np.exp(-(1 - np.array(col)) * 2.0 * den)
    assert np.allclose(centre, want, rtol=0.02), (centre, want)
    # outside the ball's silhouette nothing is attenuated
    assert np.allclose(img.reshape(65, 65, 4)[0, 0, :3], 1.0)


def test_emissive_medium_adds_density_times_length(rp, po):
    den, col = 0.3, (1.0, 0.5, 0.25)
    e = _ball_in_white_sky(rp, rp.Medium(rp.MediumType.EMISSIVE, den, rp.F3(*col), 0.0), depth=6)
    img, _, _, _ = po.OracleScene(e, "f64").render(65, 65, 64)
    centre = img.reshape(65, 65, 4)[31:34, 31:34, :3].mean((0, 1))
    assert np.allclose(centre, 1.0 + np.array(col) * 2.0 * den, rtol=0.02), centre


def test_scattering_medium_conserves_energy_in_a_white_furnace(rp, po):
    """albedo 1, any anisotropy: every path ends in the white sky with throughput 1 unless the depth limit cuts it; with albedo a the
    radiance is below 1 and above a^(expected number of collisions bound)"""
    for g in (0.0, 0.6):
        e = _ball_in_white_sky(rp, rp.Medium(rp.MediumType.SCATTER, 1.5, rp.F3(1.0, 1.0, 1.0), g), depth=200)
        img, _, _, oc = po.OracleScene(e, "f64").render(65, 65, 16, counters=True)
        centre = img.reshape(65, 65, 4)[28:37, 28:37, :3]
        assert abs(centre.mean() - 1.0) < 0.01, (g, centre.mean())
        assert oc["end_depth"] < 1e-3 * oc["samples"]
    e = _ball_in_white_sky(rp, rp.Medium(rp.MediumType.SCATTER, 1.5, rp.F3(0.8, 0.8, 0.8), 0.0), depth=200)
    img, _, _, _ = po.OracleScene(e, "f64").render(65, 65, 16)
    c = img.reshape(65, 65, 4)[28:37, 28:37, :3].mean()
    # unscattered share exp(-3) keeps 1; single scattering or more loses at least one factor 0.8
    assert math.exp(-3.0) + 0.3 * (1 - math.exp(-3.0)) < c < math.exp(-3.0) + 0.8 * (1 - math.exp(-3.0)) + 0.01, c
