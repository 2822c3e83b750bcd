// oracle_capi.cpp — C entry points of the CPU ORACLE for ctypes (test infrastructure only; see the
// header of pt_oracle.hpp: parity unpinned by reference tests, never linked into the product).
//
// Array convention is the same SoA one the device test entry points use (include/ptb200.h):
// a "vec3 array of n" is 3 consecutive blocks of n values (x block, y block, z block).
#include "pt_oracle.hpp"

#include <algorithm>
#include <chrono>
#ifdef _OPENMP
#include <omp.h>
#endif

using namespace pto;

namespace {

template <class R> struct SceneBox {
    std::unique_ptr<Scene<R>> scene;
    FlatScene<R>* flat = nullptr;   // non-null when data-driven
};

template <class R> inline V3<R> ld3(const R* a, size_t n, size_t i) { return V3<R>(a[i], a[n + i], a[2 * n + i]); }
template <class R> inline void st3(R* a, size_t n, size_t i, const V3<R>& v) { a[i] = v.x; a[n + i] = v.y; a[2 * n + i] = v.z; }

template <class R> void t_sphere_hit(size_t n, const R* o, const R* d, const R* c, const R* r, R* t_out) {
    for (size_t i = 0; i < n; ++i) {
        R t;
        bool h = isect_sphere(Ray<R>(ld3(o, n, i), ld3(d, n, i)), ld3(c, n, i), r[i], t);
        t_out[i] = h ? t : R(-1);
    }
}
template <class R> void t_plane_hit(size_t n, const R* o, const R* d, const R* p, const R* nn, R* t_out) {
    for (size_t i = 0; i < n; ++i) {
        R t;
        bool h = isect_plane(Ray<R>(ld3(o, n, i), ld3(d, n, i)), ld3(p, n, i), ld3(nn, n, i), t);
        t_out[i] = h ? t : R(-1);
    }
}
template <class R> void t_gen_ray(const SceneBox<R>* sb, size_t n, const R* p2, const R* off2, R w, R h, R* o_out, R* d_out) {
    for (size_t i = 0; i < n; ++i) {
        Ray<R> ray = sb->scene->camera().gen_ray(p2[i], p2[n + i], off2[i], off2[n + i], w, h);
        st3(o_out, n, i, ray.origin);
        st3(d_out, n, i, ray.direction);
    }
}
template <class R>
void t_closest_hit(const SceneBox<R>* sb, size_t n, const R* o, const R* d, const R* hit_dist_in, uint32_t* hit_out,
                   uint32_t* emitter_out, R* hit_dist_out, R* normal_out, uint32_t* material_out, R* light_pdf_out,
                   R* light_emission_out, R* mat_fields_out /* optional n*17: resolved material, AoS per element */) {
    for (size_t i = 0; i < n; ++i) {
        State<R> st;
        LightSampleRec<R> ls;
        st.hit_dist = hit_dist_in[i];
        st.material = Material<R>();
        bool hit = sb->scene->closest_hit(Ray<R>(ld3(o, n, i), ld3(d, n, i)), st, ls);
        hit_out[i] = hit;
        emitter_out[i] = st.is_emitter;
        hit_dist_out[i] = st.hit_dist;
        st3(normal_out, n, i, st.normal);
        material_out[i] = st.material_index < 0 ? 0xffffffffu : (uint32_t)st.material_index;
        light_pdf_out[i] = ls.pdf;
        st3(light_emission_out, n, i, ls.emission);
        if (mat_fields_out) {
            const Material<R>& m = st.material;
            R* f = mat_fields_out + i * 17;
            f[0] = m.rgb.x; f[1] = m.rgb.y; f[2] = m.rgb.z; f[3] = m.emission.x; f[4] = m.emission.y; f[5] = m.emission.z;
            f[6] = m.anisotropic; f[7] = m.metallic; f[8] = m.roughness; f[9] = m.subsurface; f[10] = m.specular_tint;
            f[11] = m.sheen; f[12] = m.sheen_tint; f[13] = m.clearcoat; f[14] = m.clearcoat_gloss; f[15] = m.spec_trans;
            f[16] = m.ior;
        }
    }
}
template <class R> void t_any_hit(const SceneBox<R>* sb, size_t n, const R* o, const R* d, const R* max_dist, uint32_t* hit_out) {
    for (size_t i = 0; i < n; ++i) hit_out[i] = sb->scene->any_hit(Ray<R>(ld3(o, n, i), ld3(d, n, i)), max_dist[i]);
}
template <class R> void t_background(const SceneBox<R>* sb, size_t n, const R* d, R* rgb_out) {
    for (size_t i = 0; i < n; ++i) st3(rgb_out, n, i, sb->scene->background(Ray<R>(V3<R>(), ld3(d, n, i))));
}
template <class R>
void t_sample_light(const SceneBox<R>* sb, size_t n, uint32_t li, const R* pos, const R* r1, const R* r2, R* normal_out,
                    R* emission_out, R* direction_out, R* dist_out, R* pdf_out) {
    Tracer<R> tr(sb->scene.get());
    for (size_t i = 0; i < n; ++i) {
        LightSampleRec<R> ls;
        tr.sample_light(sb->scene->light_at(li), ld3(pos, n, i), ls, r1[i], r2[i]);
        st3(normal_out, n, i, ls.normal);
        st3(emission_out, n, i, ls.emission);
        st3(direction_out, n, i, ls.direction);
        dist_out[i] = ls.dist;
        pdf_out[i] = ls.pdf;
    }
}
// Resolve material `mi` of a flat scene as closest_hit would for a ray of direction d (single hit).
template <class R> Material<R> resolve_material(const SceneBox<R>* sb, uint32_t mi, const Ray<R>& ray) {
    State<R> st;
    st.material = Material<R>();
    if (sb->flat) sb->flat->apply_material(st, mi, ray);
    return st.material;
}
template <class R>
void t_finalize(const SceneBox<R>* sb, size_t n, uint32_t mi, const R* o, const R* d, const R* hit_dist, const R* normal,
                R* rough_out, R* ccrough_out, R* ax_out, R* ay_out, R* eta_out, R* ffn_out, R* fhp_out) {
    for (size_t i = 0; i < n; ++i) {
        Ray<R> ray(ld3(o, n, i), ld3(d, n, i));
        State<R> st;
        st.material = resolve_material(sb, mi, ray);
        st.hit_dist = hit_dist[i];
        st.normal = ld3(normal, n, i);
        st.finalize(ray);
        rough_out[i] = st.material.roughness;
        ccrough_out[i] = st.material.clearcoat_roughness;
        ax_out[i] = st.material.ax;
        ay_out[i] = st.material.ay;
        eta_out[i] = st.eta;
        st3(ffn_out, n, i, st.ffnormal);
        st3(fhp_out, n, i, st.fhp);
    }
}
// get_spec_color (tracer.rs:335-341) + the material-only lobe weights (tracer.rs:423, 426) of material `mi` after
// Material::finalize, for a given eta: spec_col3, sheen_col3, luminance(rgb), diffuse weight, clearcoat weight
template <class R>
void t_spec_color(const SceneBox<R>* sb, uint32_t mi, R eta, const R* dir3, R* out9) {
    State<R> st;
    st.material = resolve_material(sb, mi, Ray<R>(V3<R>(), V3<R>(dir3[0], dir3[1], dir3[2])));
    st.material.finalize();
    V3<R> spec, sheen;
    Tracer<R>::get_spec_color(st.material, eta, spec, sheen);
    out9[0] = spec.x; out9[1] = spec.y; out9[2] = spec.z;
    out9[3] = sheen.x; out9[4] = sheen.y; out9[5] = sheen.z;
    R wd = 0, wr = 0, wt = 0, wc = 0;
    const R lum = Tracer<R>::luminance(st.material.rgb);
    out9[6] = lum;
    out9[7] = lum * (R(1) - st.material.metallic) * (R(1) - st.material.spec_trans);        // tracer.rs:423 before normalisation
    out9[8] = R(0.25) * st.material.clearcoat * (R(1) - st.material.metallic);              // tracer.rs:426
    (void)wd; (void)wr; (void)wt; (void)wc;
}
template <class R>
void t_disney_eval(const SceneBox<R>* sb, size_t n, uint32_t mi, const R* eta, const R* v, const R* nrm, const R* l, R* f_out,
                   R* pdf_out) {
    Tracer<R> tr(sb->scene.get());
    for (size_t i = 0; i < n; ++i) {
        State<R> st;
        st.material = resolve_material(sb, mi, Ray<R>(V3<R>(), -ld3(v, n, i)));
        st.material.finalize();
        st.eta = eta[i];
        R pdf;
        V3<R> f = tr.disney_eval(st, ld3(v, n, i), ld3(nrm, n, i), ld3(l, n, i), pdf, nullptr);
        st3(f_out, n, i, f);
        pdf_out[i] = pdf;
    }
}
template <class R>
void t_disney_sample(const SceneBox<R>* sb, size_t n, uint32_t mi, const R* eta, const R* v, const R* nrm, const R* lprev,
                     const R* r1, const R* r2, const R* coin, uint32_t* lobe_out, R* l_out, R* f_out, R* pdf_out) {
    Tracer<R> tr(sb->scene.get());
    for (size_t i = 0; i < n; ++i) {
        State<R> st;
        st.material = resolve_material(sb, mi, Ray<R>(V3<R>(), -ld3(v, n, i)));
        st.material.finalize();
        st.eta = eta[i];
        V3<R> l = ld3(lprev, n, i);
        R pdf;
        int lobe = -1;
        V3<R> f = tr.disney_sample(st, ld3(v, n, i), ld3(nrm, n, i), l, pdf, r1[i], r2[i], coin[i], &lobe, nullptr);
        lobe_out[i] = (uint32_t)lobe;
        st3(l_out, n, i, l);
        st3(f_out, n, i, f);
        pdf_out[i] = pdf;
    }
}
template <class R> void t_rng(size_t n, const uint32_t* pixel, const uint64_t* sample, uint32_t bounce, uint64_t seed, R* out8) {
    for (size_t i = 0; i < n; ++i) {
        CounterRng<R> rng(pixel[i], sample[i], seed);
        for (uint32_t s = 0; s < 8; ++s) out8[s * n + i] = rng.draw(bounce, s);
    }
}

// Scalar helpers exposed for the known-answer tests.  op codes below.
template <class R> int t_scalar(int op, const R* a, R* out) {
    using T = Tracer<R>;
    switch (op) {
        case 0: out[0] = T::power_heuristic(a[0], a[1]); return 1;
        case 1: out[0] = T::schlick_fresnel(a[0]); return 1;
        case 2: out[0] = T::dielectric_fresnel(a[0], a[1]); return 1;
        case 3: out[0] = T::gtr1(a[0], a[1]); return 1;
        case 4: out[0] = T::smithg(a[0], a[1]); return 1;
        case 5: out[0] = T::gtr2aniso(a[0], a[1], a[2], a[3], a[4]); return 1;
        case 6: out[0] = T::smithganiso(a[0], a[1], a[2], a[3], a[4]); return 1;
        case 7: { V3<R> v = T::cosine_sample_hemisphere(a[0], a[1]); out[0] = v.x; out[1] = v.y; out[2] = v.z; return 3; }
        case 8: { V3<R> v = T::sample_gtr1(a[0], a[1], a[2]); out[0] = v.x; out[1] = v.y; out[2] = v.z; return 3; }
        case 9: { V3<R> v = T::sample_ggxvndf(V3<R>(a[0], a[1], a[2]), a[3], a[4], a[5], a[6]); out[0] = v.x; out[1] = v.y; out[2] = v.z; return 3; }
        case 10: out[0] = checker<R>(a[0], a[1], a[2], a[3]); return 1;
        case 11: out[0] = T::luminance(V3<R>(a[0], a[1], a[2])); return 1;
        default: return -1;
    }
}

template <class R>
double t_render(const SceneBox<R>* sb, uint32_t w, uint32_t h, R* pixels, uint64_t* frames_inout, uint32_t n_frames,
                uint64_t sample_base, uint64_t seed, int threads, Counters* counters) {
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#else
    (void)threads;
#endif
    ColorBuffer<R> buf(w, h);
    std::copy(pixels, pixels + (size_t)w * h * 4, buf.pixels.begin());
    buf.frames = (size_t)*frames_inout;
    Tracer<R> tr(sb->scene.get());
    tr.seed = seed;
    if (sb->flat) tr.eps = sb->flat->eps;
    auto t0 = std::chrono::steady_clock::now();
    for (uint32_t f = 0; f < n_frames; ++f) tr.render(buf, counters, sample_base);
    auto t1 = std::chrono::steady_clock::now();
    std::copy(buf.pixels.begin(), buf.pixels.end(), pixels);
    *frames_inout = buf.frames;
    return std::chrono::duration<double>(t1 - t0).count();
}

// radiance of individual (pixel, sample) pairs — lets tests compare single paths
template <class R>
void t_trace_samples(const SceneBox<R>* sb, uint32_t w, uint32_t h, size_t n, const uint32_t* px, const uint32_t* py_row,
                     const uint64_t* sample, uint64_t seed, R* rgb_out) {
    Tracer<R> tr(sb->scene.get());
    tr.seed = seed;
    if (sb->flat) tr.eps = sb->flat->eps;
    for (size_t i = 0; i < n; ++i) {
        size_t row = py_row[i];
        size_t j = (size_t)h - 1 - row;
        V3<R> c = tr.trace_sample(px[i], j, w, (R)h, (uint32_t)(row * w + px[i]), sample[i], nullptr);
        st3(rgb_out, n, i, c);
    }
}

template <class R>
double t_render_chacha(const SceneBox<R>* sb, uint32_t w, uint32_t h, R* pixels, uint64_t* frames_inout, uint32_t n_frames, uint64_t seed, int threads) {
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#else
    (void)threads;
#endif
    ColorBuffer<R> buf(w, h);
    std::copy(pixels, pixels + (size_t)w * h * 4, buf.pixels.begin());
    buf.frames = (size_t)*frames_inout;
    Tracer<R> tr(sb->scene.get());
    if (sb->flat) tr.eps = sb->flat->eps;
    auto t0 = std::chrono::steady_clock::now();
    for (uint32_t f = 0; f < n_frames; ++f) tr.render_chacha(buf, seed);
    auto t1 = std::chrono::steady_clock::now();
    std::copy(buf.pixels.begin(), buf.pixels.end(), pixels);
    *frames_inout = buf.frames;
    return std::chrono::duration<double>(t1 - t0).count();
}

// signed-distance program: evaluation and tracing on their own (per-function parity of the extension)
template <class R> void t_sdf_eval(const SceneBox<R>* sb, size_t n, const R* q, R* dist_out, uint32_t* mat_out) {
    for (size_t i = 0; i < n; ++i) sb->flat->sdf.eval(ld3(q, n, i), dist_out[i], mat_out[i]);
}
template <class R> void t_sdf_trace(const SceneBox<R>* sb, size_t n, const R* o, const R* d, const R* limit, R* t_out, R* normal_out, uint32_t* mat_out) {
    for (size_t i = 0; i < n; ++i) {
        Ray<R> ray(ld3(o, n, i), ld3(d, n, i));
        uint32_t m = 0xffffffffu;
        const R t = sb->flat->sdf.trace(ray, limit[i], m);
        t_out[i] = t; mat_out[i] = m;
        st3(normal_out, n, i, t >= R(0) ? sb->flat->sdf.normal(ray.at(t)) : V3<R>(0, 0, 0));
    }
}

// One sample of pixel (px, row) of a w x h frame traced on a recorded draw sequence (SeqRng); returns the draws consumed,
// or -1 if the sequence was too short.
template <class R>
long t_trace_scripted(const SceneBox<R>* sb, uint32_t w, uint32_t h, uint32_t px, uint32_t row, const R* draws, size_t n_draws, R* rgb_out) {
    Tracer<R> tr(sb->scene.get());
    if (sb->flat) tr.eps = sb->flat->eps;
    SeqRng<R> rng(draws, n_draws);
    V3<R> c = tr.trace_sample_with(px, (size_t)h - 1 - row, w, (R)h, rng, nullptr);
    rgb_out[0] = c.x; rgb_out[1] = c.y; rgb_out[2] = c.z;
    return rng.overrun ? -1 : (long)rng.pos;
}

}  // namespace

extern "C" {

int pto_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void pto_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]};
    philox4x32_10(c, key[0], key[1]);
    for (int i = 0; i < 4; ++i) out[i] = c[i];
}

size_t pto_counters_size(void) { return sizeof(Counters); }

void pto_chacha_block(const uint32_t key[8], uint64_t counter, uint64_t stream, int rounds, uint32_t out[16]) {
    chacha_block(key, counter, stream, rounds, out);
}

#define PTO_INSTANTIATE(SFX, R)                                                                                         \
    void* pto_scene_literal_##SFX(void) {                                                                               \
        auto* sb = new SceneBox<R>();                                                                                   \
        sb->scene.reset(new AnalyticalSceneLiteral<R>());                                                               \
        return sb;                                                                                                      \
    }                                                                                                                   \
    void* pto_scene_flat_##SFX(const ptb_scene_##SFX* s) {                                                              \
        auto* sb = new SceneBox<R>();                                                                                   \
        auto* f = new FlatScene<R>(*s);                                                                                 \
        sb->scene.reset(f);                                                                                             \
        sb->flat = f;                                                                                                   \
        return sb;                                                                                                      \
    }                                                                                                                   \
    void pto_scene_destroy_##SFX(void* sb) { delete static_cast<SceneBox<R>*>(sb); }                                    \
    int pto_scene_set_sdf_##SFX(void* sb, const ptb_sdf_##SFX* sdf) {                                                   \
        auto* b = static_cast<SceneBox<R>*>(sb);                                                                        \
        if (!b->flat) return -1;                                                                                        \
        b->flat->set_sdf(sdf);                                                                                          \
        return 0;                                                                                                       \
    }                                                                                                                   \
    void pto_sdf_eval_##SFX(void* sb, size_t n, const R* q, R* dist, uint32_t* mat) { t_sdf_eval<R>(static_cast<SceneBox<R>*>(sb), n, q, dist, mat); } \
    void pto_sdf_trace_##SFX(void* sb, size_t n, const R* o, const R* d, const R* limit, R* t, R* nrm, uint32_t* mat) { \
        t_sdf_trace<R>(static_cast<SceneBox<R>*>(sb), n, o, d, limit, t, nrm, mat);                                     \
    }                                                                                                                   \
    void pto_sphere_hit_##SFX(size_t n, const R* o, const R* d, const R* c, const R* r, R* t) { t_sphere_hit<R>(n, o, d, c, r, t); } \
    void pto_plane_hit_##SFX(size_t n, const R* o, const R* d, const R* p, const R* nn, R* t) { t_plane_hit<R>(n, o, d, p, nn, t); } \
    void pto_gen_ray_##SFX(void* sb, size_t n, const R* p2, const R* off2, R w, R h, R* o, R* d) {                      \
        t_gen_ray<R>(static_cast<SceneBox<R>*>(sb), n, p2, off2, w, h, o, d);                                           \
    }                                                                                                                   \
    void pto_closest_hit_##SFX(void* sb, size_t n, const R* o, const R* d, const R* hd_in, uint32_t* hit, uint32_t* em, \
                               R* hd_out, R* nrm, uint32_t* mat, R* lpdf, R* lem, R* mat_fields) {                      \
        t_closest_hit<R>(static_cast<SceneBox<R>*>(sb), n, o, d, hd_in, hit, em, hd_out, nrm, mat, lpdf, lem, mat_fields); \
    }                                                                                                                   \
    void pto_any_hit_##SFX(void* sb, size_t n, const R* o, const R* d, const R* md, uint32_t* hit) {                    \
        t_any_hit<R>(static_cast<SceneBox<R>*>(sb), n, o, d, md, hit);                                                  \
    }                                                                                                                   \
    /* media extension: Henyey-Greenstein sample around v (3, n) and its phase value = pdf */                           \
    void pto_sample_hg_##SFX(size_t n, const R* v, R g, const R* r1, const R* r2, R* dir, R* pdf) {                      \
        for (size_t i = 0; i < n; ++i) {                                                                                \
            const V3<R> vv = ld3(v, n, i);                                                                              \
            const V3<R> d = Tracer<R>::sample_hg(vv, g, r1[i], r2[i]);                                                  \
            st3(dir, n, i, d);                                                                                          \
            pdf[i] = Tracer<R>::phase_hg(dot(vv, d), g);                                                                \
        }                                                                                                               \
    }                                                                                                                   \
    void pto_background_##SFX(void* sb, size_t n, const R* d, R* rgb) { t_background<R>(static_cast<SceneBox<R>*>(sb), n, d, rgb); } \
    void pto_sample_light_##SFX(void* sb, size_t n, uint32_t li, const R* pos, const R* r1, const R* r2, R* nrm, R* em, \
                                R* dir, R* dist, R* pdf) {                                                              \
        t_sample_light<R>(static_cast<SceneBox<R>*>(sb), n, li, pos, r1, r2, nrm, em, dir, dist, pdf);                  \
    }                                                                                                                   \
    void pto_finalize_##SFX(void* sb, size_t n, uint32_t mi, const R* o, const R* d, const R* hd, const R* nrm, R* rough, \
                            R* ccr, R* ax, R* ay, R* eta, R* ffn, R* fhp) {                                             \
        t_finalize<R>(static_cast<SceneBox<R>*>(sb), n, mi, o, d, hd, nrm, rough, ccr, ax, ay, eta, ffn, fhp);          \
    }                                                                                                                   \
    void pto_spec_color_##SFX(void* sb, uint32_t mi, R eta, const R* dir3, R* out9) {                                    \
        t_spec_color<R>(static_cast<SceneBox<R>*>(sb), mi, eta, dir3, out9);                                            \
    }                                                                                                                   \
    void pto_disney_eval_##SFX(void* sb, size_t n, uint32_t mi, const R* eta, const R* v, const R* nrm, const R* l, R* f, \
                               R* pdf) {                                                                                \
        t_disney_eval<R>(static_cast<SceneBox<R>*>(sb), n, mi, eta, v, nrm, l, f, pdf);                                 \
    }                                                                                                                   \
    void pto_disney_sample_##SFX(void* sb, size_t n, uint32_t mi, const R* eta, const R* v, const R* nrm, const R* lprev, \
                                 const R* r1, const R* r2, const R* coin, uint32_t* lobe, R* l, R* f, R* pdf) {         \
        t_disney_sample<R>(static_cast<SceneBox<R>*>(sb), n, mi, eta, v, nrm, lprev, r1, r2, coin, lobe, l, f, pdf);    \
    }                                                                                                                   \
    void pto_rng_##SFX(size_t n, const uint32_t* pixel, const uint64_t* sample, uint32_t bounce, uint64_t seed, R* out8) { \
        t_rng<R>(n, pixel, sample, bounce, seed, out8);                                                                 \
    }                                                                                                                   \
    int pto_scalar_##SFX(int op, const R* a, R* out) { return t_scalar<R>(op, a, out); }                                \
    void pto_convert_to_u8_##SFX(size_t n_pixels, const R* rgba, uint8_t* out) { convert_to_u8<R>(rgba, n_pixels, out); } \
    void pto_convert_to_u8_at_##SFX(const R* rgba, size_t bw, size_t bh, uint8_t* frame, size_t x, size_t y, size_t fw, \
                                    size_t fh) {                                                                        \
        convert_to_u8_at<R>(rgba, bw, bh, frame, x, y, fw, fh);                                                         \
    }                                                                                                                   \
    double pto_render_##SFX(void* sb, uint32_t w, uint32_t h, R* pixels, uint64_t* frames_inout, uint32_t n_frames,     \
                            uint64_t sample_base, uint64_t seed, int threads, void* counters) {                         \
        return t_render<R>(static_cast<SceneBox<R>*>(sb), w, h, pixels, frames_inout, n_frames, sample_base, seed, threads, \
                           static_cast<Counters*>(counters));                                                           \
    }                                                                                                                   \
    long pto_trace_scripted_##SFX(void* sb, uint32_t w, uint32_t h, uint32_t px, uint32_t row, const R* draws, size_t n_draws, R* rgb) { \
        return t_trace_scripted<R>(static_cast<SceneBox<R>*>(sb), w, h, px, row, draws, n_draws, rgb);                  \
    }                                                                                                                   \
    double pto_render_chacha_##SFX(void* sb, uint32_t w, uint32_t h, R* pixels, uint64_t* frames_inout, uint32_t n_frames,\
                                   uint64_t seed, int threads) {                                                        \
        return t_render_chacha<R>(static_cast<SceneBox<R>*>(sb), w, h, pixels, frames_inout, n_frames, seed, threads);  \
    }                                                                                                                   \
    void pto_trace_samples_##SFX(void* sb, uint32_t w, uint32_t h, size_t n, const uint32_t* px, const uint32_t* row,   \
                                 const uint64_t* sample, uint64_t seed, R* rgb) {                                       \
        t_trace_samples<R>(static_cast<SceneBox<R>*>(sb), w, h, n, px, row, sample, seed, rgb);                         \
    }

PTO_INSTANTIATE(f32, float)
PTO_INSTANTIATE(f64, double)

}  // extern "C"
