//! ref_kat — known-answer vectors from the reference's OWN arithmetic (markusmoenig/rust-pathtracer).
//!
//! Links a copy of the reference crate whose only edits are visibility (`fn` -> `pub fn` on `impl Tracer`) and the TYPE of
//! the RNG handed to the drawing functions (tools/ref_kat/run.sh lists and diffs the three edits), plus
//! `renderer/src/analytical.rs` included verbatim, and dumps per-function input/output pairs and whole-path radiances as
//! JSON.  `tests/test_ref_kat.py` replays every case through the C++ oracle (oracle/pt_oracle.hpp); a green run pins the
//! oracle — and through it the CUDA path — to the Rust code instead of to a reading of it.
//!
//! Every case is `{"fn": name, "in": [f32...], "out": [f32...]}`; f32 values are printed with `{:e}` (shortest
//! round-trip form), non-finite ones as the strings "nan" / "inf" / "-inf".  Layouts are documented at each call below.

use rust_pathtracer::kat_rng;
use rust_pathtracer::prelude::*;
use std::fmt::Write as _;

#[path = "../_work/analytical.rs"]
mod analytical;
use analytical::{AnalyticalIntersections, AnalyticalScene};

// ---- deterministic inputs on the 2^-24 grid ------------------------------------------------------
struct Lcg(u64);
impl Lcg {
    fn u(&mut self) -> f32 {
        self.0 = self.0.wrapping_mul(6364136223846793005).wrapping_add(1442695040888963407);
        ((self.0 >> 40) as u32) as f32 / 16777216.0
    }
    fn range(&mut self, a: f32, b: f32) -> f32 { a + (b - a) * self.u() }
    fn unit(&mut self) -> F3 {
        loop {
            let v = F3::new(self.range(-1.0, 1.0), self.range(-1.0, 1.0), self.range(-1.0, 1.0));
            let l = v.length();
            if l > 0.1 && l <= 1.0 { return v.normalize(); }
        }
    }
    fn unit_up(&mut self) -> F3 { let v = self.unit(); F3::new(v.x, v.y, v.z.abs().max(0.02)).normalize() }
}

// ---- JSON ------------------------------------------------------------------------------------------
fn fj(x: f32) -> String {
    if x.is_finite() { format!("{:e}", x) } else if x.is_nan() { "\"nan\"".to_string() } else if x > 0.0 { "\"inf\"".to_string() } else { "\"-inf\"".to_string() }
}
fn arr(v: &[f32]) -> String { format!("[{}]", v.iter().map(|x| fj(*x)).collect::<Vec<_>>().join(",")) }
struct Out { s: String, n: usize }
impl Out {
    fn case(&mut self, name: &str, inp: &[f32], out: &[f32]) {
        if self.n > 0 { self.s.push_str(",\n"); }
        self.n += 1;
        write!(self.s, "{{\"fn\":\"{}\",\"in\":{},\"out\":{}}}", name, arr(inp), arr(out)).unwrap();
    }
}
fn v3(v: &F3) -> [f32; 3] { [v.x, v.y, v.z] }
fn opt(t: Option<F>) -> f32 { match t { Some(d) => d, None => -1.0 } }

/// A ray that hits demo primitive `mi` (0 metal sphere, 1 clearcoat sphere, 2 checker plane) and nothing else first, so
/// that `closest_hit` leaves exactly that primitive's material assignments in `state.material` (analytical.rs:41-116).
fn ray_onto(mi: usize, g: &mut Lcg) -> Ray {
    match mi {
        0 => Ray::new(F3::new(-1.1 + g.range(-0.6, 0.6), g.range(-0.6, 0.6), 3.0), F3::new(g.range(-0.04, 0.04), g.range(-0.04, 0.04), -1.0).normalize()),
        1 => Ray::new(F3::new(1.1 + g.range(-0.6, 0.6), g.range(-0.6, 0.6), 3.0), F3::new(g.range(-0.04, 0.04), g.range(-0.04, 0.04), -1.0).normalize()),
        _ => Ray::new(F3::new(g.range(-0.05, 0.05), 0.0, 3.0), F3::new(g.range(-0.08, 0.08), -0.8, -0.6 + g.range(-0.08, 0.08)).normalize()),
    }
}
/// (ray, state after closest_hit + finalize) on primitive `mi`
fn shaded_state(scene: &AnalyticalScene, mi: usize, g: &mut Lcg) -> (Ray, State) {
    loop {
        let ray = ray_onto(mi, g);
        let mut state = State::new();
        let mut ls = LightSampleRec::new();
        state.material = Material::new();
        if scene.closest_hit(&ray, &mut state, &mut ls) && !state.is_emitter {
            state.finalize(&ray);
            return (ray, state);
        }
    }
}

fn main() {
    let path = std::env::args().nth(1).unwrap_or_else(|| "ref_kat_f32.json".to_string());
    assert_eq!(std::mem::size_of::<F>(), 4, "this dumper is written for the f32 build of `F` (lib.rs:6)");
    let scene = AnalyticalScene::new();
    let tracer = Tracer::new(Box::new(AnalyticalScene::new()));
    let mut g = Lcg(0xB200);
    let mut o = Out { s: String::new(), n: 0 };
    const N: usize = 64;

    // ---- scalar terms, tracer.rs:222-333.  in = the arguments in declaration order, out = the value(s) ----
    for _ in 0..N { let (a, b) = (g.range(0.0, 50.0), g.range(0.0, 50.0)); o.case("power_heuristic", &[a, b], &[tracer.power_heuristic(&a, &b)]); }
    for _ in 0..N { let u = g.range(-0.2, 1.2); o.case("schlick_fresnel", &[u], &[tracer.schlick_fresnel(u)]); }
    for k in 0..N {
        let c = g.u();
        let eta = match k % 4 { 0 => 1.0 / 1.45, 1 => 1.45, 2 => 1.0 / 1.5, _ => g.range(0.5, 2.0) };
        o.case("dielectric_fresnel", &[c, eta], &[tracer.dielectric_fresnel(c, eta)]);
    }
    for k in 0..N { let (h, a) = (g.u(), if k % 8 == 0 { 0.001 } else { g.range(0.001, 1.2) }); o.case("gtr1", &[h, a], &[tracer.gtr1(&h, a)]); }
    for _ in 0..N { let (v, a) = (g.range(0.01, 1.0), g.range(0.05, 1.0)); o.case("smithg", &[v, a], &[tracer.smithg(&v, a)]); }
    for _ in 0..N {
        let h = g.unit_up(); let (ax, ay) = (g.range(0.001, 1.0), g.range(0.001, 1.0));
        o.case("gtr2aniso", &[h.z, h.x, h.y, ax, ay], &[tracer.gtr2aniso(&h.z, &h.x, &h.y, &ax, &ay)]);
        o.case("smithganiso", &[h.z, h.x, h.y, ax, ay], &[tracer.smithganiso(&h.z, &h.x, &h.y, &ax, &ay)]);
    }
    for _ in 0..N { let c = F3::new(g.range(0.0, 2.0), g.range(0.0, 2.0), g.range(0.0, 2.0)); o.case("luminance", &v3(&c), &[tracer.luminance(&c)]); }
    for _ in 0..N { let (r1, r2) = (g.u(), g.u()); o.case("cosine_sample_hemisphere", &[r1, r2], &v3(&tracer.cosine_sample_hemisphere(r1, r2))); }
    for k in 0..N { let (a, r1, r2) = (if k % 4 == 0 { 0.001 } else { g.range(0.001, 0.5) }, g.u(), g.u()); o.case("sample_gtr1", &[a, r1, r2], &v3(&tracer.sample_gtr1(a, r1, r2))); }
    for _ in 0..N {
        let v = g.unit_up(); let (ax, ay, r1, r2) = (g.range(0.001, 1.0), g.range(0.001, 1.0), g.u(), g.u());
        o.case("sample_ggxvndf", &[v.x, v.y, v.z, ax, ay, r1, r2], &v3(&tracer.sample_ggxvndf(&v, ax, ay, r1, r2)));
    }

    // ---- Pinhole::gen_ray, pinhole.rs:38-61.  in = p.xy, offset.xy, width, height; out = origin, direction ----
    let cam = Pinhole::new();
    for k in 0..N {
        let (w, h) = if k % 2 == 0 { (800.0, 600.0) } else { (3840.0, 2160.0) };
        let (p, off) = (F2::new(g.u(), g.u()), F2::new(g.u(), g.u()));
        let r = cam.gen_ray(p, off, w, h);
        o.case("gen_ray", &[p.x, p.y, off.x, off.y, w, h], &[r.origin.x, r.origin.y, r.origin.z, r.direction.x, r.direction.y, r.direction.z]);
    }

    // ---- AnalyticalScene::sphere / plane, analytical.rs:166-204.  out = t, or -1 for None ----
    for _ in 0..4 * N {
        let c = F3::new(g.range(-2.0, 2.0), g.range(-2.0, 2.0), g.range(-2.0, 2.0));
        let rad = g.range(0.05, 1.5);
        let org = F3::new(g.range(-4.0, 4.0), g.range(-4.0, 4.0), g.range(-4.0, 4.0));
        // aim near the sphere so that hits, grazing hits and misses all occur
        let aim = c + F3::new_x(rad * 1.2) * g.unit();
        let ray = Ray::new(org, (aim - org).normalize());
        o.case("sphere", &[org.x, org.y, org.z, ray.direction.x, ray.direction.y, ray.direction.z, c.x, c.y, c.z, rad], &[opt(scene.sphere(&ray, c, rad))]);
    }
    for k in 0..2 * N {
        let org = F3::new(g.range(-4.0, 4.0), g.range(-3.0, 3.0), g.range(-4.0, 4.0));
        let mut d = g.unit();
        if k % 8 == 0 { d = F3::new(d.x, 5e-5, d.z).normalize(); }      // |denom| below the 1e-4 threshold
        let ray = Ray::new(org, d);
        o.case("plane", &[org.x, org.y, org.z, d.x, d.y, d.z], &[opt(scene.plane(&ray))]);
    }

    // ---- Scene::closest_hit incl. sample_lights with the stale hit_dist, analytical.rs:36-127 + scene.rs:36-86 ----
    // in = origin, direction, state.hit_dist before the call
    // out = hit, is_emitter, hit_dist, normal, light_sample.pdf, light_sample.emission, then the 17 material fields
    //       rgb, emission, anisotropic, metallic, roughness, subsurface, specular_tint, sheen, sheen_tint, clearcoat,
    //       clearcoat_gloss, spec_trans, ior (state.material was Material::new() before the call, tracer.rs:63)
    let hds = [-1.0f32, 0.3, 2.0, 7.0, 1.0e30];
    for k in 0..8 * N {
        let ray = if k % 2 == 0 {
            cam.gen_ray(F2::new(g.u(), g.u()), F2::new(0.0, 0.0), 800.0, 600.0)
        } else {
            // secondary-like rays: from points around the spheres / above the plane, any direction; a third aimed at the light
            let org = F3::new(g.range(-3.0, 3.5), g.range(-0.99, 2.5), g.range(-3.0, 3.0));
            let d = if k % 3 == 0 { (F3::new(3.0, 2.0, 2.0) + F3::new_x(0.9) * g.unit() - org).normalize() } else { g.unit() };
            Ray::new(org, d)
        };
        let hd = hds[(k / 2) % hds.len()];
        let mut state = State::new();
        let mut ls = LightSampleRec::new();
        state.material = Material::new();
        state.hit_dist = hd;
        let hit = scene.closest_hit(&ray, &mut state, &mut ls);
        let m = &state.material;
        o.case("closest_hit", &[ray.origin.x, ray.origin.y, ray.origin.z, ray.direction.x, ray.direction.y, ray.direction.z, hd],
               &[hit as u32 as f32, state.is_emitter as u32 as f32, state.hit_dist, state.normal.x, state.normal.y, state.normal.z, ls.pdf,
                 ls.emission.x, ls.emission.y, ls.emission.z,
                 m.rgb.x, m.rgb.y, m.rgb.z, m.emission.x, m.emission.y, m.emission.z, m.anisotropic, m.metallic, m.roughness, m.subsurface,
                 m.specular_tint, m.sheen, m.sheen_tint, m.clearcoat, m.clearcoat_gloss, m.spec_trans, m.ior]);
        let md = g.range(0.1, 8.0);
        o.case("any_hit", &[ray.origin.x, ray.origin.y, ray.origin.z, ray.direction.x, ray.direction.y, ray.direction.z, md], &[scene.any_hit(&ray, md) as u32 as f32]);
    }
    for _ in 0..N { let d = g.unit(); o.case("background", &v3(&d), &v3(&scene.background(&Ray::new(F3::zeros(), d)))); }

    // ---- State::finalize + Material::finalize, globals.rs:50-62, material.rs:117-131 ----
    // in = primitive (0, 1, 2), origin, direction, hit_dist, normal;  out = roughness, clearcoat_roughness, ax, ay, eta, ffnormal, fhp
    for k in 0..3 * N {
        let mi = k % 3;
        let (ray, st) = shaded_state(&scene, mi, &mut g);
        o.case("finalize", &[mi as f32, ray.origin.x, ray.origin.y, ray.origin.z, ray.direction.x, ray.direction.y, ray.direction.z, st.hit_dist,
                             st.normal.x, st.normal.y, st.normal.z],
               &[st.material.roughness, st.material.clearcoat_roughness, st.material.ax, st.material.ay, st.eta, st.ffnormal.x, st.ffnormal.y,
                 st.ffnormal.z, st.fhp.x, st.fhp.y, st.fhp.z]);
    }

    // ---- Tracer::disney_eval, tracer.rs:555-626.  in = primitive, eta, v, n, l;  out = f, pdf ----
    for k in 0..6 * N {
        let mi = k % 3;
        let (ray, st) = shaded_state(&scene, mi, &mut g);
        let v = -ray.direction;
        // light directions over the whole sphere, biased towards the upper hemisphere of n
        let mut l = g.unit();
        if k % 4 != 0 && dot(&l, &st.ffnormal) < 0.0 { l = -l; }
        let mut pdf: F = 0.0;
        let f = tracer.disney_eval(&st, v, &st.ffnormal, &l, &mut pdf);
        o.case("disney_eval", &[mi as f32, st.eta, v.x, v.y, v.z, st.ffnormal.x, st.ffnormal.y, st.ffnormal.z, l.x, l.y, l.z], &[f.x, f.y, f.z, pdf]);
    }

    // ---- Tracer::disney_sample, tracer.rs:441-553, on scripted draws (r1, r2 at 446-447; coin at 534, spec lobe only) ----
    // in = primitive, eta, v, n, l before the call (the stale `l` of quirk A.5), r1, r2, coin;  out = f, l, pdf, draws consumed
    for k in 0..6 * N {
        let mi = k % 3;
        let (ray, st) = shaded_state(&scene, mi, &mut g);
        let v = -ray.direction;
        let lprev = if k % 2 == 0 { F3::zeros() } else { ray.direction };
        let draws = [g.u(), g.u(), g.u()];
        kat_rng::script(&draws);
        let mut rng = kat_rng::thread_rng();
        let (mut l, mut pdf) = (lprev, 0.0 as F);
        let f = tracer.disney_sample(&st, v, &st.ffnormal, &mut l, &mut pdf, &mut rng);
        let consumed = 3 - kat_rng::remaining();
        o.case("disney_sample", &[mi as f32, st.eta, v.x, v.y, v.z, st.ffnormal.x, st.ffnormal.y, st.ffnormal.z, lprev.x, lprev.y, lprev.z,
                                  draws[0], draws[1], draws[2]],
               &[f.x, f.y, f.z, l.x, l.y, l.z, pdf, consumed as f32]);
    }

    // ---- Tracer::sample_light, tracer.rs:173-220.  in = scatter_pos, r1, r2;  out = normal, emission, direction, dist, pdf ----
    for _ in 0..2 * N {
        let pos = F3::new(g.range(-3.0, 2.0), g.range(-1.0, 1.0), g.range(-3.0, 3.0));
        let draws = [g.u(), g.u()];
        kat_rng::script(&draws);
        let mut rng = kat_rng::thread_rng();
        let mut ls = LightSampleRec::new();
        tracer.sample_light(&scene.light_at(0).light, &pos, &mut ls, &mut rng);
        assert_eq!(kat_rng::remaining(), 0);
        o.case("sample_light", &[pos.x, pos.y, pos.z, draws[0], draws[1]],
               &[ls.normal.x, ls.normal.y, ls.normal.z, ls.emission.x, ls.emission.y, ls.emission.z, ls.direction.x, ls.direction.y, ls.direction.z, ls.dist, ls.pdf]);
    }

    // ---- ColorBuffer::convert_to_u8, buffer.rs:55-64.  in = rgba;  out = the four bytes ----
    for _ in 0..N {
        let mut b = ColorBuffer::new(1, 1);
        let px = [g.range(0.0, 1.3), g.range(0.0, 1.3), g.range(0.0, 1.3), g.range(0.0, 1.0)];
        b.pixels.copy_from_slice(&px);
        let mut frame = [0u8; 4];
        b.convert_to_u8(&mut frame);
        o.case("convert_to_u8", &px, &[frame[0] as f32, frame[1] as f32, frame[2] as f32, frame[3] as f32]);
    }

    // ---- the whole per-pixel loop, tracer.rs:22-123: Tracer::render into a 1x1 ColorBuffer on scripted draws ----
    // A 1x1 frame has pixel_size 1, so the jitter (draws 0, 1) spans the camera's whole field of view (aspect 1).  The script is
    // consumed in call order: jitter x, y; per shaded bounce light pick, light r1, r2, bsdf r1, r2 and — spec lobe only — the
    // coin.  frames = 0, so the running mean returns the sample itself: mix(0, c, 1/1) = c (tracer.rs:108-115).
    // in = the 2 + 6 * depth scripted draws;  out = r, g, b, alpha, draws consumed
    for _ in 0..8 * N {
        let mut tr = Tracer::new(Box::new(AnalyticalScene::new()));
        let draws: Vec<f32> = (0..26).map(|_| g.u()).collect();
        kat_rng::script(&draws);
        let mut b = ColorBuffer::new(1, 1);
        tr.render(&mut b);
        let consumed = draws.len() - kat_rng::remaining();
        o.case("path_1x1", &draws, &[b.pixels[0], b.pixels[1], b.pixels[2], b.pixels[3], consumed as f32]);
    }

    let json = format!("{{\"meta\":{{\"source\":\"reference\",\"crate\":\"rust-pathtracer 0.2.4\",\"f\":\"f32\",\"cases\":{}}},\n\"cases\":[\n{}\n]}}\n", o.n, o.s);
    std::fs::write(&path, json).expect("write");
    eprintln!("[ref_kat] {} cases -> {}", o.n, path);
}
