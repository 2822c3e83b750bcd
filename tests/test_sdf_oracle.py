"""CPU checks of the oracle's signed-distance evaluator (oracle/pt_oracle.hpp, SdfProgram) against closed forms: the statement of
the extension that the CUDA evaluator is compared with in tests/test_gpu_sdf.py."""
import copy

import numpy as np


def _scene(rp, nodes, **kw):
    sc = rp.sdf_demo_scene().device_export()
    sc = copy.copy(sc)
    sc.sdf = rp.SdfProgram(nodes=nodes, **kw)
    return sc


def test_primitive_distances_closed_form(rp, po):
    N = rp.SdfNode
    q = np.array([[2.0, 0.0, 0.0, 0.3], [0.0, 0.0, 3.0, 0.2], [0.0, 0.5, 0.0, -0.1]], np.float32)        # 4 points as columns
    d, m = po.OracleScene(_scene(rp, [N.sphere((0, 0, 0), 1.0, 2)])).sdf_eval(q)
    np.testing.assert_allclose(d, np.linalg.norm(q.astype(np.float64), axis=0) - 1.0, rtol=0, atol=1e-6)
    assert np.all(m == 2)
    d, _ = po.OracleScene(_scene(rp, [N.box((0, 0, 0), (1.0, 0.5, 0.25), 1)])).sdf_eval(q)
    np.testing.assert_allclose(d, [1.0, 0.25, 2.5, -0.15], atol=1e-6)           # beyond the x face; beyond the z face; beyond the y face; inside (nearest face: z)
    d, _ = po.OracleScene(_scene(rp, [N.torus((0, 0, 0), 1.0, 0.25, 1)])).sdf_eval(q)
    qd = q.astype(np.float64)
    np.testing.assert_allclose(d, np.hypot(np.hypot(qd[0], qd[2]) - 1.0, qd[1]) - 0.25, atol=1e-6)
    assert abs(d[0] - 0.75) < 1e-6                                               # (2, 0, 0): one ring radius outside the tube centre
    d, _ = po.OracleScene(_scene(rp, [N.plane((0, 1, 0), 1.0, 0)])).sdf_eval(q)
    np.testing.assert_allclose(d, q[1] + 1.0, atol=1e-7)


def test_combinators(rp, po):
    N = rp.SdfNode
    a, b = N.sphere((-0.5, 0, 0), 1.0, 1), N.sphere((0.5, 0, 0), 1.0, 2)
    q = np.array([[-1.0, 1.0, 0.0, 3.0], [0.0, 0.0, 0.2, 0.0], [0.0, 0.0, 0.0, 0.0]], np.float32)
    da = np.linalg.norm(q - np.array([[-0.5], [0], [0]], np.float32), axis=0) - 1.0
    db = np.linalg.norm(q - np.array([[0.5], [0], [0]], np.float32), axis=0) - 1.0
    d, m = po.OracleScene(_scene(rp, [a, b, N.union()])).sdf_eval(q)
    np.testing.assert_allclose(d, np.minimum(da, db), atol=1e-6)
    assert list(m) == [1, 2, 1, 2]                                   # nearer operand; a tie goes to the first
    d, m = po.OracleScene(_scene(rp, [a, b, N.intersect()])).sdf_eval(q)
    np.testing.assert_allclose(d, np.maximum(da, db), atol=1e-6)
    d, m = po.OracleScene(_scene(rp, [a, b, N.subtract()])).sdf_eval(q)
    np.testing.assert_allclose(d, np.maximum(da, -db), atol=1e-6)
    assert np.all(m == 1)
    k = 0.3
    d, m = po.OracleScene(_scene(rp, [a, b, N.smooth_union(k)])).sdf_eval(q)
    h = np.clip(0.5 + 0.5 * (db - da) / k, 0, 1)
    np.testing.assert_allclose(d, (1 - h) * db + da * h - k * h * (1 - h), atol=1e-6)
    assert np.all(d <= np.minimum(da, db) + 1e-6) and np.all(d >= np.minimum(da, db) - k / 4 - 1e-6)


def test_sphere_trace_agrees_with_the_analytic_sphere(rp, po):
    N = rp.SdfNode
    osc = po.OracleScene(_scene(rp, [N.sphere((0.2, 0.1, -0.3), 0.9, 1)], hit_eps=1e-5, max_dist=50.0, max_steps=256))
    rng = np.random.default_rng(3)
    n = 2000
    o = rng.uniform(-4, 4, size=(3, n)).astype(np.float32)
    aim = np.array([[0.2], [0.1], [-0.3]], np.float32) + rng.uniform(-1.1, 1.1, size=(3, n)).astype(np.float32)
    d = aim - o
    d = (d / np.linalg.norm(d, axis=0, keepdims=True)).astype(np.float32)
    outside = np.linalg.norm(o - np.array([[0.2], [0.1], [-0.3]], np.float32), axis=0) > 0.95
    r = osc.sdf_trace(o, d, np.full(n, 1e30, np.float32))
    t_ref = po.sphere_hit(o, d, np.repeat(np.array([[0.2], [0.1], [-0.3]], np.float32), n, 1), np.full(n, 0.9, np.float32))
    hit = (t_ref >= 0) & outside
    graze = np.abs(np.linalg.norm(np.cross((np.array([[0.2], [0.1], [-0.3]]) - o).T, d.T), axis=1) - 0.9) < 2e-2     # within ~hit_eps of tangent
    sel = hit & ~graze
    assert sel.sum() > 500
    assert np.all(r["t"][sel] >= 0)
    assert np.abs(r["t"][sel] - t_ref[sel]).max() < 2e-3            # the tracer stops hit_eps short of the surface along the ray: <= eps / cos
    hp = o[:, sel] + r["t"][sel] * d[:, sel]
    nref = (hp - np.array([[0.2], [0.1], [-0.3]], np.float32)) / 0.9
    assert np.abs(r["normal"][:, sel] - nref).max() < 5e-3          # tetrahedron gradient with h = 1e-3
    miss = (t_ref < 0) & outside & ~graze
    assert np.all(r["t"][miss] < 0)
    # a limit shorter than the hit distance turns the hit into a miss (shadow rays, tracer.rs:152)
    lim = (t_ref * 0.5).astype(np.float32)
    r2 = osc.sdf_trace(o[:, sel], d[:, sel], lim[sel])
    assert np.all(r2["t"] < 0)


def test_closest_hit_orders_the_body_after_the_planes_and_before_the_lights(rp, po):
    export = rp.sdf_demo_scene().device_export()
    osc = po.OracleScene(export)
    o = np.array([[0.0, 1.2, 0.0], [0.0, -0.45, 0.0], [3.0, 3.0, 3.0]], np.float32).T.copy()     # columns: towards the boxy body, the glass ball, the floor
    d = np.array([[-0.6, -0.1, -3.0], [0.0, 0.0, -3.0], [0.0, -1.0, -0.0001]], np.float32).T.copy()
    d /= np.linalg.norm(d, axis=0, keepdims=True)
    o = np.ascontiguousarray(np.array([[0.0, 0.0, 3.0], [1.2, -0.45, 3.0], [0.0, 3.0, 3.0]], np.float32).T)
    h = osc.closest_hit(o, d.astype(np.float32), np.full(3, -1.0, np.float32))
    assert list(h["hit"]) == [1, 1, 1]
    assert list(h["material"]) == [1, 3, 0]                           # paint (body), glass ball, checker plane
    assert abs(h["hit_dist"][1] - (3.0 - 0.3 - 0.55)) < 1e-3
