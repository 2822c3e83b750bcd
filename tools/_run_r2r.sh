mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -12) > gpurun_out/r2r_pytest.log
for c in 4 5; do (timeout 600 python bench.py --config $c --steps 3 --no-cpu-baseline 2> gpurun_out/r2r_bench_c$c.err | tail -1) > gpurun_out/r2r_bench_c$c.json; done
tail -4 gpurun_out/r2r_pytest.log
python - <<'PY'
import json
for c in (4,5):
    d=json.load(open(f"gpurun_out/r2r_bench_c{c}.json")); r=d["roofline"]
    print(c, round(d["value"]), round(d["e2e"]["value"]), d["rays_per_s"], r and {k:r[k] for k in ("achieved","peak","frac","flops_per_sample")}, r and r.get("bvh"))
PY
