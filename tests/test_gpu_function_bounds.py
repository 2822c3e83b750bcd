"""Per-function parity as BOUNDS, for both builds of the CUDA library (VERDICT r1 "next" 1b).

tests/golden/fn_bounds.json holds, for every case of tests/fn_cases.py, twice the maximum and twice the p99.9 relative error that
tools/function_parity.py measured on a B200 against the CPU oracle on these same seeded inputs (profiles/r02_function_parity.md
is the table).  This module asserts them as hard bounds — a maximum, not a percentile — and pins WHICH cases meet the north-star
1e-5 bar as a maximum:

  * strict build (libptb200_strict.so: IEEE operations in the reference's order, no contraction): every geometry function, every
    finalize output and all of disney_eval are BIT-IDENTICAL to the oracle (bound 0); disney_sample inherits the rare last-bit
    differences between glibc's and a correctly rounded sin / cos / pow;
  * shipped build (FMA contraction, 2-ulp reciprocal / sqrt / rsqrt, MUFU pow): the geometry side meets 1e-5 as a maximum; the
    BSDF side tracks the formula's own conditioning (the oracle's f32-vs-f64 distance on the same inputs, the table's last
    column) — 1e-5 at p99 or better for the well-conditioned lobes, up to percents in D(h) of the alpha = 0.001 clearcoat lobe,
    where the reference's f32 value is itself 8 % away from its f64 value.
"""
import json
import os

import numpy as np
import pytest

import fn_cases as fc
from devfn import DeviceFns

pytestmark = pytest.mark.gpu
BOUNDS_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fn_bounds.json")
BOUNDS = json.load(open(BOUNDS_PATH)) if os.path.exists(BOUNDS_PATH) else None
NORTH_STAR = 1e-5

# cases that must meet the north-star 1e-5 relative bound AS A MAXIMUM (measured; a regression here is a failure even if the
# doubled bound in the JSON would still hold)
GEOMETRY = ["sphere.t", "plane.t", "gen_ray.dir[800x600]", "gen_ray.dir[3840x2160]", "closest_hit.hit_dist", "closest_hit.normal",
            "closest_hit.light_pdf", "background.rgb", "sample_light.direction", "sample_light.normal", "sample_light.dist", "finalize.*"]
MEETS_1E5 = {
    "shipped": set(GEOMETRY) | {"disney_eval[2].pdf", "disney_eval[5].f", "disney_eval[5].pdf", "disney_sample[4].l", "disney_sample[5].pdf"},
    # everything except six disney_sample outputs whose maximum is 1.0e-5 .. 1.4e-5 (a last-bit difference between glibc's sinf /
    # cosf / powf and the correctly rounded value, amplified by the sampled lobe): those are held to 2e-5
    "strict": (set(GEOMETRY) | {"sample_light.pdf"} | {f"disney_eval[{m}].{k}" for m in range(7) for k in ("f", "pdf")}
               | {f"disney_sample[{m}].{k}" for m in range(7) for k in ("l", "pdf", "f", "w (= f/pdf)")})
    - {"disney_sample[0].pdf", "disney_sample[0].f", "disney_sample[1].pdf", "disney_sample[1].f", "disney_sample[1].w (= f/pdf)", "disney_sample[2].pdf"},
}
STRICT_CEILING = 2e-5
BIT_EXACT_STRICT = {"sphere.t", "plane.t", "gen_ray.dir[800x600]", "gen_ray.dir[3840x2160]", "closest_hit.hit_dist", "closest_hit.normal",
                    "closest_hit.light_pdf", "finalize.*"} | {f"disney_eval[{m}].{k}" for m in range(7) for k in ("f", "pdf")}


def check(build, name, err):
    assert BOUNDS is not None, "tests/golden/fn_bounds.json missing: run tools/function_parity.py on a GPU box"
    b = BOUNDS[build][name]
    err = np.asarray(err, np.float64)
    assert err.size > 0.5 * b["n"], (name, err.size, b["n"])
    mx, p999 = float(err.max()), float(np.percentile(err, 99.9))
    assert mx <= b["max"], (build, name, "max", mx, b["max"])
    assert p999 <= b["p999"], (build, name, "p99.9", p999, b["p999"])
    if name in MEETS_1E5[build]:
        assert mx <= NORTH_STAR, (build, name, mx)
    if build == "strict":
        assert mx <= STRICT_CEILING, (name, mx)
        if name in BIT_EXACT_STRICT:
            assert mx == 0.0, (name, mx)


@pytest.fixture(scope="module", params=["shipped", "strict"])
def build(request):
    return request.param


@pytest.fixture(scope="module")
def zoo(rp, po, build):
    e = fc.zoo_export(rp)
    d = DeviceFns(rp, e, strict=(build == "strict"))
    yield d, po.OracleScene(e), po.OracleScene(e, "f64")
    d.close()


def test_geometry_functions(rp, po, build, demo_export, oracle_demo):
    dev = DeviceFns(rp, demo_export, strict=(build == "strict"))
    seen = []
    for name, err, _, _ in fc.geometry_cases(dev, oracle_demo, po):
        check(build, name, err)
        seen.append(name)
    dev.close()
    assert set(GEOMETRY) <= set(seen) and "sample_light.pdf" in seen


@pytest.mark.parametrize("mi", range(7))
def test_disney_eval_bounds(build, zoo, mi):
    dev, osc, osc64 = zoo
    e = fc.eval_case(dev, osc, osc64, mi, seed=(9 + mi) if mi < 3 else (40 + mi))
    assert e["n_well"] > 40_000 and e["zero_flips"] <= 5
    for k in ("f", "pdf"):
        check(build, f"disney_eval[{mi}].{k}", e[k])
    if mi in (4, 6):                                                    # refraction really evaluated
        assert ((e["lz"] < -0.05) & (np.abs(e["rf"]).max(0) > 0))[e["well"]].sum() > 1000


@pytest.mark.parametrize("mi", range(7))
def test_disney_sample_bounds(build, zoo, mi):
    dev, osc, osc64 = zoo
    s = fc.sample_case(dev, osc, osc64, mi, seed=(20 + mi) if mi < 3 else (60 + mi))
    assert s["lobe_flips"] <= 3, s["lobe_flips"]                        # draws within an ulp of a CDF edge
    present = set(np.unique(s["lobe"]).tolist())
    assert {1: {0, 1, 2}, 4: {2, 3}, 5: {2}}.get(mi, present) <= present
    for k in ("l", "pdf", "f", "w"):
        check(build, f"disney_sample[{mi}].{k}" + (" (= f/pdf)" if k == "w" else ""), s[k])
