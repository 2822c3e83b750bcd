for i in 1 2 3; do echo "cfg5 integrator $i"; python tools/prof_cfg.py 5 $i 3840x2160x8; done
for n in 4096 8192 16384 32768; do echo "field $n"; python - <<PY
import sys; sys.path.insert(0,'.')
import rust_pathtracer_b200 as rp
sc = rp.sphere_field_scene(n_spheres=$n, n_lights_side=4)
for integ in (2, 3):
    pt = rp.Tracer.new(sc, integrator=integ); buf = rp.ColorBuffer.new(1920, 1080)
    pt.render_spp(buf, 1, download=False); pt.render_spp(buf, 8, download=False); ms = pt.last_render_ms(); u = pt.integrator_used(); pt.close()
    print("   ", u, round(1920*1080*8/ms/1e3, 1), "Msamples/s")
PY
done
