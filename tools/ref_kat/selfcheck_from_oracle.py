"""Writes tests/golden/ref_kat_selfcheck_f32.json: cases in tools/ref_kat's format whose OUTPUTS COME FROM THE ORACLE, not from
the reference (meta.source says so).  Its only purpose is to keep the replay code of tests/test_ref_kat.py exercised in every CPU
run; it pins nothing.  The real file, tests/golden/ref_kat_f32.json, is produced by tools/ref_kat/run.sh with a Rust toolchain."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ref_kat_replay as rk                     # noqa: E402
import rust_pathtracer_b200 as rp               # noqa: E402
from oracle import pyoracle as po               # noqa: E402

F = np.float32
g = np.random.default_rng(0xB200)


def u(n=1):
    return (g.integers(0, 1 << 24, n) / 16777216.0).astype(F)


def rng(a, b, n=1):
    return (F(a) + (F(b) - F(a)) * u(n)).astype(F)


def unit():
    while True:
        v = rng(-1, 1, 3)
        l = np.linalg.norm(v)
        if 0.1 < l <= 1:
            return (v / l).astype(F)


def main():
    po.load()
    osc = po.OracleScene(rp.AnalyticalScene.new().device_export())
    cases = []

    def add(name, a):
        a = np.asarray(a, F)
        out = rk.one(name, a, osc, po)
        cases.append({"fn": name, "in": [rk.enc(x) for x in a], "out": [rk.enc(x) for x in out]})

    for _ in range(8):
        add("power_heuristic", rng(0, 50, 2)); add("schlick_fresnel", rng(-0.2, 1.2)); add("dielectric_fresnel", [u()[0], 1 / 1.45])
        add("gtr1", [u()[0], rng(0.001, 1.2)[0]]); add("smithg", [rng(0.01, 1)[0], 0.25])
        h = unit(); h[2] = abs(h[2])
        add("gtr2aniso", [h[2], h[0], h[1], 0.1, 0.3]); add("smithganiso", [h[2], h[0], h[1], 0.1, 0.3]); add("luminance", rng(0, 2, 3))
        add("cosine_sample_hemisphere", u(2)); add("sample_gtr1", [0.001, u()[0], u()[0]]); add("sample_ggxvndf", [h[0], h[1], h[2], 0.1, 0.2, u()[0], u()[0]])
        add("gen_ray", np.concatenate([u(4), [800, 600]]))
        o, c = rng(-4, 4, 3), rng(-2, 2, 3)
        d = c + F(0.5) * unit() - o
        add("sphere", np.concatenate([o, d / np.linalg.norm(d), c, [1.0]])); add("plane", np.concatenate([o, unit()]))
        o2, d2 = osc.gen_ray(u(2).reshape(2, 1), np.zeros((2, 1), F), 800.0, 600.0)
        for hd in (-1.0, 2.0, 1e30):
            add("closest_hit", np.concatenate([o2[:, 0], d2[:, 0], [hd]]))
        add("any_hit", np.concatenate([o2[:, 0], d2[:, 0], [3.0]])); add("background", unit())
        add("sample_light", np.concatenate([rng(-3, 2, 3), u(2)])); add("convert_to_u8", rng(0, 1.3, 4)); add("path_1x1", u(26))
        for mi, (org, dr) in enumerate((([-1.1, 0.1, 3], [0, 0, -1]), ([1.1, -0.2, 3], [0, 0, -1]), ([0, 0, 3], [0, -0.8, -0.6]))):
            org, dr = np.asarray(org, F), np.asarray(dr, F)
            hit = osc.closest_hit(org.reshape(3, 1), dr.reshape(3, 1), np.array([-1.0], F))
            t, nrm = hit["hit_dist"][0], hit["normal"][:, 0]
            add("finalize", np.concatenate([[mi], org, dr, [t], nrm]))
            fin = osc.finalize(mi, org.reshape(3, 1), dr.reshape(3, 1), np.array([t], F), nrm.reshape(3, 1))
            l = unit()
            if np.dot(l, fin["ffnormal"][:, 0]) < 0:
                l = -l
            add("disney_eval", np.concatenate([[mi, fin["eta"][0]], -dr, fin["ffnormal"][:, 0], l]))
            add("disney_sample", np.concatenate([[mi, fin["eta"][0]], -dr, fin["ffnormal"][:, 0], dr, u(3)]))
    out = os.path.join(ROOT, "tests", "golden", "ref_kat_selfcheck_f32.json")
    json.dump({"meta": {"source": "oracle-selfcheck (NOT the reference: exercises the replay harness only)", "f": "f32", "cases": len(cases)},
               "cases": cases}, open(out, "w"))
    print(out, len(cases))


if __name__ == "__main__":
    main()
