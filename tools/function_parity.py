"""Per-function error table: the shipped CUDA build and an IEEE measurement build against the CPU oracle on identical inputs
(VERDICT r1 "next" 1b).  A measurement tool like tests/: it runs the oracle as the checker.

  python tools/function_parity.py build     # here (CPU): nvcc cross-compiles rust_pathtracer_b200/variants/libptb200_ieee.so
  python tools/function_parity.py run       # GPU box: one subprocess per build -> gpurun_out/function_parity.{json,md}

IEEE build = -DPTB_IEEE (powf / sincosf / IEEE reciprocal, division, sqrt, normalize) -prec-div=true -prec-sqrt=true -ftz=false
-fmad=false: what is left between it and the oracle is CUDA libm vs glibc rounding.  `floor` columns: the oracle's own f32
evaluation against its f64 evaluation on the same inputs — the conditioning of the reference's formulas.
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
VDIR = os.path.join(ROOT, "rust_pathtracer_b200", "variants")
IEEE_LIB = os.path.join(ROOT, "rust_pathtracer_b200", "libptb200_strict.so")       # built by __graft_entry__.build()
# two more builds that separate the causes: the shipped approximations WITHOUT FMA contraction, and IEEE operations WITH it
NOFMA_LIB = os.path.join(VDIR, "libptb200_fast_nofma.so")
IEEEFMA_LIB = os.path.join(VDIR, "libptb200_ieee_fma.so")


def build():
    import __graft_entry__ as g
    os.makedirs(VDIR, exist_ok=True)
    procs = []
    for lib, flags in ((NOFMA_LIB, g.NVCC_FLAGS + ["-fmad=false"]), (IEEEFMA_LIB, [f for f in g.STRICT_FLAGS if f != "-fmad=false"])):
        cmd = ["nvcc"] + flags + ["-o", lib, os.path.join(g.CSRC, "ptb_api.cu")]
        print(" ".join(cmd), flush=True)
        procs.append(subprocess.Popen(cmd))
    g.build()
    for p in procs:
        assert p.wait() == 0


def one():
    import numpy as np
    import rust_pathtracer_b200 as rp
    from oracle import pyoracle as po
    import fn_cases as fc
    from devfn import DeviceFns
    po.load()
    rows = []
    demo = rp.AnalyticalScene.new().device_export()
    dev, osc = DeviceFns(rp, demo), po.OracleScene(demo)
    for name, err, floor, note in fc.geometry_cases(dev, osc, po):
        rows.append(dict(name=name, note=note, **fc.stats(err), frac_gt_1e5=float((np.asarray(err) > 1e-5).mean()) if len(err) else 0.0))
    dev.close()
    zoo = fc.zoo_export(rp)
    dev, osc, osc64 = DeviceFns(rp, zoo), po.OracleScene(zoo), po.OracleScene(zoo, "f64")
    names = {0: "metal a=0.05", 1: "orange clearcoat gloss 1 (a=0.001)", 2: "checker plane (diffuse)", 3: "zoo clearcoat wide", 4: "zoo rough glass",
             5: "zoo aniso metal", 6: "zoo mixed"}
    for mi in range(7):
        e = fc.eval_case(dev, osc, osc64, mi, seed=(9 + mi) if mi < 3 else (40 + mi))
        for k in ("f", "pdf"):
            err, fl = e[k], e[k + "_floor"]
            over = err > 1e-5
            rows.append(dict(name=f"disney_eval[{mi}].{k}", note=names[mi], **fc.stats(err), frac_gt_1e5=float(over.mean()),
                             floor=fc.stats(fl), max_err_over_floor_where_gt_1e5=float((err[over] / np.maximum(fl[over], 6e-8)).max()) if over.any() else 0.0))
        s = fc.sample_case(dev, osc, osc64, mi, seed=(20 + mi) if mi < 3 else (60 + mi))
        for k in ("l", "pdf", "f", "w"):
            err, fl = s[k], s[k + "_floor"]
            over = err > 1e-5
            rows.append(dict(name=f"disney_sample[{mi}].{k}" + (" (= f/pdf)" if k == "w" else ""), note=names[mi], **fc.stats(err),
                             frac_gt_1e5=float(over.mean()), floor=fc.stats(fl), lobe_flips=s["lobe_flips"],
                             max_err_over_floor_where_gt_1e5=float((err[over] / np.maximum(fl[over], 6e-8)).max()) if over.any() else 0.0))
    dev.close()
    print("JSON:" + json.dumps(rows), flush=True)


def fmt(x):
    return "0" if x == 0 else f"{x:.1e}"


def run():
    out_dir = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    res = {}
    for name, lib in (("shipped", os.path.join(ROOT, "rust_pathtracer_b200", "libptb200.so")), ("ieee", IEEE_LIB), ("fast_nofma", NOFMA_LIB),
                      ("ieee_fma", IEEEFMA_LIB)):
        if not os.path.exists(lib):
            print(name, "missing:", lib, flush=True)
            continue
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "one"], env=dict(os.environ, PTB200_LIB=lib), capture_output=True, text=True, timeout=1500)
        line = [l for l in r.stdout.splitlines() if l.startswith("JSON:")]
        if not line:
            print(name, "failed:", r.stderr[-2000:], flush=True)
            continue
        res[name] = json.loads(line[0][5:])
    json.dump(res, open(os.path.join(out_dir, "function_parity.json"), "w"))
    # bounds for tests/test_gpu_function_bounds.py: twice the measured maximum / p99.9 (exactly 0 stays 0: bit-exact is asserted as such)
    def bound(x):
        return 0.0 if x == 0 else max(2.0 * x, 2.5e-7)
    bounds = {tag: {r["name"]: {"max": bound(r["max"]), "p999": bound(r["p999"]), "measured_max": r["max"], "measured_p999": r["p999"], "n": r["n"]}
                    for r in res.get(src, [])} for tag, src in (("shipped", "shipped"), ("strict", "ieee"))}
    json.dump(bounds, open(os.path.join(out_dir, "fn_bounds.json"), "w"), indent=1)
    def extra(name):
        rows = {r["name"]: r for r in res.get(name, [])}
        return lambda n: rows.get(n)
    nofma, ieeefma = extra("fast_nofma"), extra("ieee_fma")
    lines = ["| function | n | shipped max | p99.9 | p99 | p95 | >1e-5 | IEEE max | p99.9 | p99 | >1e-5 | shipped ops, no FMA: max / p99.9 | IEEE ops + FMA: max / p99.9 | formula floor (oracle f32 vs f64) max / p99.9 / p99 | note |",
             "|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|"]
    ship = res.get("shipped", [])
    ieee = {r["name"]: r for r in res.get("ieee", [])}
    for r in ship:
        i = ieee.get(r["name"])
        fl = r.get("floor")
        nf, jf = nofma(r["name"]), ieeefma(r["name"])
        lines.append("| `{}` | {} | {} | {} | {} | {} | {:.2%} | {} | {} | {} | {} | {} | {} | {} | {} |".format(
            r["name"], r["n"], fmt(r["max"]), fmt(r["p999"]), fmt(r["p99"]), fmt(r["p95"]), r["frac_gt_1e5"],
            fmt(i["max"]) if i else "-", fmt(i["p999"]) if i else "-", fmt(i["p99"]) if i else "-", f"{i['frac_gt_1e5']:.2%}" if i else "-",
            f"{fmt(nf['max'])} / {fmt(nf['p999'])}" if nf else "-", f"{fmt(jf['max'])} / {fmt(jf['p999'])}" if jf else "-",
            f"{fmt(fl['max'])} / {fmt(fl['p999'])} / {fmt(fl['p99'])}" if fl else "-", r["note"]))
    open(os.path.join(out_dir, "function_parity.md"), "w").write("\n".join(lines) + "\n")
    print("\n".join(lines), flush=True)


if __name__ == "__main__":
    mode = sys.argv[1] if len(sys.argv) > 1 else "run"
    {"build": build, "one": one, "run": run}[mode]()
