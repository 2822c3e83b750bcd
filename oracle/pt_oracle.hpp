// pt_oracle.hpp — CPU ORACLE (test infrastructure, NOT product code).
//
// A C++17 restatement of the per-pixel path-tracing loop of markusmoenig/rust-pathtracer
// (reference paths below are relative to /root/reference/).  It exists only so that tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs can check and time
// the CUDA path against the reference's algorithm.  Nothing under rust_pathtracer_b200/ may
// include, link or call it.
//
// PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures
// and cannot be compiled in this image (no Rust toolchain), so this restatement
// cannot be pinned against reference outputs.  It is pinned instead against (1) the published
// Philox4x32-10 known-answer vectors for the RNG, (2) hand-derived known answers for the
// analytic intersections / BSDF terms and (3) the mean colour of the
// reference's screenshot images/spheres.png.  See tests/test_oracle.py.
//
// Rules followed so the f32 instantiation produces what the Rust f32 build produces:
//   * operation order and associativity exactly as written in the Rust source;
//   * compiled with -ffp-contract=off -fno-fast-math (rustc never contracts a*b+c to FMA);
//   * glibc libm for sqrt/sin/cos/tan/pow/log2/floor/fmod (Rust std calls the platform libm);
//   * f32::max/min = fmaxf/fminf semantics (NaN-ignoring), f32::clamp = two compares.
//
// The third-party arithmetic on the path that is NOT in /root/reference is the RNG:
// rand 0.8.5 `thread_rng()` (ChaCha12, OS-seeded; Cargo.lock:1428-1446).  It is unseedable, so no
// image can be reproduced bit for bit even by the reference itself; only the distribution
// (uniform on the 2^-24 grid in [0,1) for f32, 2^-53 for f64) matters.  The oracle and the device
// share a counter-based Philox4x32-10 stream instead (SURVEY.md §8d "Seeds").
#pragma once
#ifdef _OPENMP
#include <omp.h>
#endif
#include <cmath>
#include <cstdint>
#include <cstddef>
#include <cstring>
#include <limits>
#include <vector>
#include <memory>

#include "../include/ptb200.h"


namespace pto {

// ------------------------------------------------------------------------------------------------
// lib.rs:5-10 — scalar switch and constants (PI is the F-typed constant; INV_PI = 1/PI and
// TWO_PI = PI*2 are evaluated in F).
template <class R> struct K;
template <> struct K<float> {
    static constexpr float PI = 3.14159265358979323846f;
    static constexpr float INV_PI = 1.0f / 3.14159265358979323846f;
    static constexpr float TWO_PI = 3.14159265358979323846f * 2.0f;
};
template <> struct K<double> {
    static constexpr double PI = 3.14159265358979323846;
    static constexpr double INV_PI = 1.0 / 3.14159265358979323846;
    static constexpr double TWO_PI = 3.14159265358979323846 * 2.0;
};

// Rust float method semantics
template <class R> inline R fmax_(R a, R b) { return std::fmax(a, b); }   // f32::max
template <class R> inline R clamp_(R x, R lo, R hi) {                    // f32::clamp
    if (x < lo) return lo;
    if (x > hi) return hi;
    return x;
}

// ------------------------------------------------------------------------------------------------
// fx.rs:209-515 — F3 and its operators (component-wise * and /, scalar*F3, normalize = 3 divides)
template <class R> struct V3 {
    R x, y, z;
    V3() : x(0), y(0), z(0) {}
    V3(R x_, R y_, R z_) : x(x_), y(y_), z(z_) {}
    static V3 zeros() { return V3(0, 0, 0); }            // fx.rs:225
    static V3 new_x(R v) { return V3(v, v, v); }          // fx.rs:233
    R length() const { return std::sqrt(x * x + y * y + z * z); }          // fx.rs:330
    V3 normalize() const { R l = length(); return V3(x / l, y / l, z / l); } // fx.rs:306
    R dot(const V3& o) const { return x * o.x + y * o.y + z * o.z; }       // fx.rs:334
    V3 cross(const V3& o) const {                                          // fx.rs:338
        return V3(y * o.z - z * o.y, z * o.x - x * o.z, x * o.y - y * o.x);
    }
    V3 mult_f(R f) const { return V3(x * f, y * f, z * f); }               // fx.rs:345
};
template <class R> inline V3<R> operator+(V3<R> a, V3<R> b) { return V3<R>(a.x + b.x, a.y + b.y, a.z + b.z); }
template <class R> inline V3<R> operator-(V3<R> a, V3<R> b) { return V3<R>(a.x - b.x, a.y - b.y, a.z - b.z); }
template <class R> inline V3<R> operator*(V3<R> a, V3<R> b) { return V3<R>(a.x * b.x, a.y * b.y, a.z * b.z); }
template <class R> inline V3<R> operator/(V3<R> a, V3<R> b) { return V3<R>(a.x / b.x, a.y / b.y, a.z / b.z); }
template <class R> inline V3<R> operator*(R s, V3<R> a) { return V3<R>(s * a.x, s * a.y, s * a.z); }  // fx.rs:476
template <class R> inline V3<R> operator-(V3<R> a) { return V3<R>(-a.x, -a.y, -a.z); }
template <class R> inline V3<R>& operator+=(V3<R>& a, V3<R> b) { a.x += b.x; a.y += b.y; a.z += b.z; return a; }
template <class R> inline V3<R>& operator/=(V3<R>& a, V3<R> b) { a.x /= b.x; a.y /= b.y; a.z /= b.z; return a; }

// math.rs:3-60
template <class R> inline R dot(const V3<R>& a, const V3<R>& b) { return a.dot(b); }
template <class R> inline V3<R> cross(const V3<R>& a, const V3<R>& b) { return a.cross(b); }
template <class R> inline V3<R> normalize(const V3<R>& a) { return a.normalize(); }
template <class R> inline R length(const V3<R>& a) { return a.length(); }
template <class R> inline V3<R> mix(const V3<R>& a, const V3<R>& b, R v) {        // math.rs:33-39
    return V3<R>((R(1) - v) * a.x + b.x * v, (R(1) - v) * a.y + b.y * v, (R(1) - v) * a.z + b.z * v);
}
template <class R> inline R mix_f(R a, R b, R v) { return (R(1) - v) * a + b * v; }  // math.rs:42
template <class R> inline V3<R> powv(const V3<R>& a, const V3<R>& e) {              // math.rs:53-59
    return V3<R>(std::pow(a.x, e.x), std::pow(a.y, e.y), std::pow(a.z, e.z));
}

// ------------------------------------------------------------------------------------------------
// ray.rs:6-33 (inv_direction / sign_* are never read on the path and go stale on bounce, A.9)
template <class R> struct Ray {
    V3<R> origin, direction;
    Ray() {}
    Ray(V3<R> o, V3<R> d) : origin(o), direction(d) {}
    V3<R> at(R dist) const { return origin + dist * direction; }   // ray.rs:31-33
};

// material.rs:48-131 — only the fields the tracer reads
// material.rs:7-34.  The reference's tracer never reads it; the semantics used here are the library's extension (PTB_MEDIUM_*
// in include/ptb200.h), stated in Tracer::medium_step below.
template <class R> struct Medium {
    uint32_t type = PTB_MEDIUM_NONE;
    R density = 0;
    V3<R> color{0, 0, 0};
    R anisotropy = 0;
};
template <class R> struct Material {
    V3<R> rgb{R(1.5), R(1.5), R(1.5)};     // material.rs:85
    V3<R> emission{0, 0, 0};
    R anisotropic = 0, metallic = 0, roughness = R(0.5), subsurface = 0, specular_tint = 0;
    R sheen = 0, sheen_tint = 0, clearcoat = 0, clearcoat_gloss = 0, clearcoat_roughness = 0;
    R spec_trans = 0, ior = R(1.45);
    R ax = 0, ay = 0;
    Medium<R> medium;                       // material.rs:75
    // material.rs:117-131
    void finalize() {
        medium.anisotropy = medium.anisotropy < R(-0.9) ? R(-0.9) : (medium.anisotropy > R(0.9) ? R(0.9) : medium.anisotropy);   // material.rs:126
        roughness = fmax_(roughness, R(0.01));
        clearcoat_roughness = mix_f(R(0.1), R(0.001), clearcoat_gloss);
        R aspect = std::sqrt(R(1) - anisotropic * R(0.9));
        ax = fmax_(roughness / aspect, R(0.001));
        ay = fmax_(roughness * aspect, R(0.001));
    }
};

// globals.rs:76-84, light.rs:13-28
template <class R> struct Light {
    uint32_t type = PTB_LIGHT_SPHERICAL;
    V3<R> position, emission;
    R radius = 0, area = 0;
    V3<R> u, v;                         // globals.rs:80-81 (rectangular lights)
    static Light spherical(V3<R> pos, R radius, V3<R> emission) {
        Light l;
        l.type = PTB_LIGHT_SPHERICAL;
        l.position = pos;
        l.emission = emission;
        l.radius = radius;
        l.area = R(4) * K<R>::PI * radius * radius;   // light.rs:22
        return l;
    }
};

// globals.rs:89-130
template <class R> struct ScatterSampleRec { V3<R> l, f; R pdf = 0; };
template <class R> struct LightSampleRec { V3<R> normal, emission, direction; R dist = 0, pdf = 0; };

// globals.rs:6-62
template <class R> struct State {
    uint16_t depth = 4;
    R eta = 0;
    R hit_dist = R(-1);                 // globals.rs:28 (quirk A.1)
    V3<R> fhp, normal, ffnormal;
    bool is_emitter = false;
    Material<R> material;
    int material_index = -1;            // oracle-only bookkeeping for the per-function tests
    // globals.rs:50-62
    void finalize(const Ray<R>& ray) {
        fhp = ray.at(hit_dist);
        if (dot(normal, ray.direction) <= R(0)) ffnormal = normal; else ffnormal = -normal;
        material.finalize();
        eta = dot(ray.direction, normal) < R(0) ? R(1) / material.ior : material.ior;
    }
};

// camera/pinhole.rs:5-61
template <class R> struct Pinhole {
    V3<R> origin{0, 0, 3}, center{0, 0, 0};
    R fov = 80;
    Ray<R> gen_ray(R px, R py, R offx, R offy, R width, R height) const {
        R ratio = width / height;
        R pixel_size_x = R(1) / width, pixel_size_y = R(1) / height;
        // f32::to_radians = self * (PI / 180)
        R half_width = std::tan((fov * (K<R>::PI / R(180))) * R(0.5));
        R half_height = half_width / ratio;
        V3<R> up(0, 1, 0);
        V3<R> w = (origin - center).normalize();
        V3<R> u = up.cross(w);
        V3<R> v = w.cross(u);
        V3<R> lower_left = origin - u.mult_f(half_width) - v.mult_f(half_height) - w;
        V3<R> horizontal = u.mult_f(half_width * R(2));
        V3<R> vertical = v.mult_f(half_height * R(2));
        V3<R> rd = lower_left - origin;
        rd += horizontal.mult_f(pixel_size_x * offx + px);
        rd += vertical.mult_f(pixel_size_y * offy + py);
        return Ray<R>(origin, rd.normalize());
    }
};

// ------------------------------------------------------------------------------------------------
// Event counters: the algorithmic-FLOP model multiplies these.
struct Counters {
    uint64_t samples = 0, closest_hit = 0, any_hit = 0, shade = 0, nee_contrib = 0, eval_calls = 0;
    uint64_t lobe_diffuse = 0, lobe_clearcoat = 0, lobe_reflect = 0, lobe_refract = 0;
    uint64_t end_sky = 0, end_emitter = 0, end_pdf = 0, end_depth = 0;
    // per-lobe evaluations (inside disney_eval and disney_sample)
    uint64_t ev_diffuse = 0, ev_reflect = 0, ev_refract = 0, ev_clearcoat = 0;
    uint64_t nee_culled = 0, nee_shadowed = 0, background = 0, finalize = 0;
    void add(const Counters& o) {
        const uint64_t* s = reinterpret_cast<const uint64_t*>(&o);
        uint64_t* d = reinterpret_cast<uint64_t*>(this);
        for (size_t i = 0; i < sizeof(Counters) / sizeof(uint64_t); ++i) d[i] += s[i];
    }
};

// ------------------------------------------------------------------------------------------------
// Counter RNG: Philox4x32-10 (Salmon et al., SC'11; Random123 reference constants).
// key = (pixel index, low 32 bits of the global sample index)
// ctr = (block, high 32 bits of the sample index, seed lo, seed hi)
// f32: block = bounce*2 + slot/4, word = slot%4, u = (word >> 8) * 2^-24
// f64: block = bounce*4 + slot/2, words (2*(slot%2), 2*(slot%2)+1), u = ((w0<<32|w1) >> 11) * 2^-53
// slots: 0,1 jitter (tracer.rs:45) | 2 light pick (137) | 3 reflect/refract coin (534)
//        | 4,5 light r1,r2 (191-192) | 6,7 bsdf r1,r2 (446-447)
// (the four draws every shaded bounce needs share the second block; the first block is only needed by bounce 0's jitter,
//  by scenes with several lights and by materials that can refract)
enum : uint32_t { SLOT_JITTER_X = 0, SLOT_JITTER_Y = 1, SLOT_LIGHT_PICK = 2, SLOT_COIN = 3, SLOT_LIGHT_R1 = 4, SLOT_LIGHT_R2 = 5,
                  SLOT_BSDF_R1 = 6, SLOT_BSDF_R2 = 7 };
inline void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)M0 * c[0];
        uint64_t p1 = (uint64_t)M1 * c[2];
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += W0; k1 += W1;
    }
}

template <class R> struct CounterRng {
    uint32_t pixel; uint64_t sample; uint64_t seed;
    CounterRng(uint32_t p, uint64_t s, uint64_t sd) : pixel(p), sample(s), seed(sd) {}
    R draw(uint32_t bounce, uint32_t slot) const;
    void coin_consumed() const {}
};
// Replay of a recorded draw SEQUENCE (tools/ref_kat: the reference run on a scripted RNG): draws are handed out in call
// order, whatever (bounce, slot) is asked for — valid because this file asks for them at the same program points as the
// reference calls rng.gen() (tracer.rs:45, 137, 191-192, 446-447, 534).  The coin of tracer.rs:534 is drawn by the
// reference only inside the spec lobe: the oracle asks for it before it knows the lobe, so the coin is PEEKED and the
// caller reports afterwards whether the reference would have consumed it.
template <class R> struct SeqRng {
    const R* seq; size_t n; mutable size_t pos = 0; mutable bool overrun = false;
    SeqRng(const R* s, size_t count) : seq(s), n(count) {}
    R draw(uint32_t, uint32_t slot) const {
        if (pos >= n) { overrun = true; return R(0); }
        return slot == 3u /* SLOT_COIN */ ? seq[pos] : seq[pos++];
    }
    void coin_consumed() const { if (pos < n) ++pos; else overrun = true; }
};
// ChaCha block function (D. J. Bernstein's original layout: 64-bit block counter in words 12-13, 64-bit stream id in 14-15) and a
// sequential generator over it.  The reference draws from rand 0.8.5's `thread_rng()` = ChaCha12 (rand_chacha 0.3.1,
// Cargo.lock:1428-1450), buffered four blocks at a time, and turns a word into an f32 as (w >> 8) * 2^-24 (f64: 53 bits of two
// words) — this generator reproduces that COST and distribution for the CPU baseline (bench.py cpu_baseline.chacha12); the
// parity runs use the counter RNG above.  `rounds` = 20 reproduces the RFC 7539 block test vector (tests/test_oracle.py).
inline void chacha_block(const uint32_t key[8], uint64_t counter, uint64_t stream, int rounds, uint32_t out[16]) {
    uint32_t x[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u, key[0], key[1], key[2], key[3], key[4], key[5], key[6], key[7],
                      (uint32_t)counter, (uint32_t)(counter >> 32), (uint32_t)stream, (uint32_t)(stream >> 32)};
    uint32_t in[16];
    for (int i = 0; i < 16; ++i) in[i] = x[i];
    auto rotl = [](uint32_t v, int c) { return (v << c) | (v >> (32 - c)); };
    auto qr = [&](int a, int b, int c, int d) {
        x[a] += x[b]; x[d] = rotl(x[d] ^ x[a], 16);
        x[c] += x[d]; x[b] = rotl(x[b] ^ x[c], 12);
        x[a] += x[b]; x[d] = rotl(x[d] ^ x[a], 8);
        x[c] += x[d]; x[b] = rotl(x[b] ^ x[c], 7);
    };
    for (int r = 0; r < rounds; r += 2) {
        qr(0, 4, 8, 12); qr(1, 5, 9, 13); qr(2, 6, 10, 14); qr(3, 7, 11, 15);
        qr(0, 5, 10, 15); qr(1, 6, 11, 12); qr(2, 7, 8, 13); qr(3, 4, 9, 14);
    }
    for (int i = 0; i < 16; ++i) out[i] = x[i] + in[i];
}
template <class R> struct ChaCha12Rng {
    uint32_t key[8];
    mutable uint64_t counter = 0;
    mutable uint32_t buf[64];
    mutable int pos = 64;
    explicit ChaCha12Rng(uint64_t seed) {
        // key expansion from a 64-bit seed (splitmix64, like rand's seed_from_u64; the reference seeds from the OS)
        uint64_t z = seed;
        for (int i = 0; i < 4; ++i) {
            z += 0x9E3779B97F4A7C15ull;
            uint64_t v = z;
            v = (v ^ (v >> 30)) * 0xBF58476D1CE4E5B9ull; v = (v ^ (v >> 27)) * 0x94D049BB133111EBull; v ^= v >> 31;
            key[2 * i] = (uint32_t)v; key[2 * i + 1] = (uint32_t)(v >> 32);
        }
    }
    uint32_t next_u32() const {
        if (pos >= 64) {                                            // rand_chacha refills four blocks (256 bytes) at a time
            for (int b = 0; b < 4; ++b) chacha_block(key, counter++, 0, 12, buf + 16 * b);
            pos = 0;
        }
        return buf[pos++];
    }
    R next() const;
    mutable bool have_coin = false;
    mutable R coin = 0;
    // call-order interface of the tracer (see SeqRng): the coin of tracer.rs:534 is drawn only in the spec lobe
    R draw(uint32_t, uint32_t slot) const {
        if (slot == 3u) { if (!have_coin) { coin = next(); have_coin = true; } return coin; }
        return next();
    }
    void coin_consumed() const { have_coin = false; }
};
template <> inline float ChaCha12Rng<float>::next() const { return (float)(next_u32() >> 8) * (1.0f / 16777216.0f); }
template <> inline double ChaCha12Rng<double>::next() const {
    const uint64_t lo = next_u32(), hi = next_u32();
    return (double)(((hi << 32) | lo) >> 11) * (1.0 / 9007199254740992.0);
}
template <> inline float CounterRng<float>::draw(uint32_t bounce, uint32_t slot) const {
    uint32_t c[4] = {bounce * 2u + (slot >> 2), (uint32_t)(sample >> 32), (uint32_t)seed, (uint32_t)(seed >> 32)};
    philox4x32_10(c, pixel, (uint32_t)sample);
    return (float)(c[slot & 3u] >> 8) * (1.0f / 16777216.0f);
}
template <> inline double CounterRng<double>::draw(uint32_t bounce, uint32_t slot) const {
    uint32_t c[4] = {bounce * 4u + (slot >> 1), (uint32_t)(sample >> 32), (uint32_t)seed, (uint32_t)(seed >> 32)};
    philox4x32_10(c, pixel, (uint32_t)sample);
    uint32_t w0 = c[(slot & 1u) * 2u], w1 = c[(slot & 1u) * 2u + 1u];
    uint64_t bits = (((uint64_t)w0 << 32) | w1) >> 11;
    return (double)bits * (1.0 / 9007199254740992.0);
}

// ------------------------------------------------------------------------------------------------
// scene.rs:5-90 — the plug-in trait
template <class R> struct Scene {
    virtual ~Scene() {}
    virtual V3<R> background(const Ray<R>& ray) const = 0;
    virtual bool closest_hit(const Ray<R>& ray, State<R>& state, LightSampleRec<R>& light) const = 0;
    virtual bool any_hit(const Ray<R>& ray, R max_dist) const = 0;
    virtual const Pinhole<R>& camera() const = 0;
    virtual size_t number_of_lights() const = 0;
    virtual const Light<R>& light_at(size_t i) const = 0;
    virtual uint16_t recursion_depth() const { return 4; }                  // scene.rs:28-30
    V3<R> to_linear(V3<R> c) const {                                         // scene.rs:32-34
        return V3<R>(std::pow(c.x, R(2.2)), std::pow(c.y, R(2.2)), std::pow(c.z, R(2.2)));
    }

    // scene.rs:39-63 — local ray/sphere test of sample_lights
    static bool light_sphere(const Ray<R>& ray, V3<R> center, R radius, R& t_out) {
        V3<R> l = center - ray.origin;
        R tca = l.dot(ray.direction);
        R d2 = l.dot(l) - tca * tca;
        R radius2 = radius * radius;
        if (d2 > radius2) return false;
        R thc = std::sqrt(radius2 - d2);
        R t0 = tca - thc, t1 = tca + thc;
        if (t0 > t1) { R tmp = t0; t0 = t1; t1 = tmp; }
        if (t0 < R(0)) { t0 = t1; if (t0 < R(0)) return false; }
        t_out = t0;
        return true;
    }
    // PTB_SCENE_EXTENDED_LIGHTS (include/ptb200.h): rectangular / distant lights take the upstream GLSL project's semantics
    // instead of being inert like in the reference.  Not reference code: the CPU statement of the library's extension.
    virtual bool extended_lights() const { return false; }
    // ray against a rectangular light: hidden from behind; t and the cosine at the quad
    static bool light_rect(const Ray<R>& ray, const Light<R>& L, R& t_out, R& cos_out) {
        const V3<R> n = normalize(cross(L.u, L.v));
        const R dn = dot(n, ray.direction);
        if (!(dn < R(0))) return false;
        const R t = dot(n, L.position - ray.origin) / dn;
        if (!(t > R(0))) return false;
        const V3<R> w = (ray.origin + t * ray.direction) - L.position;
        const R a1 = dot(L.u, w) / dot(L.u, L.u), a2 = dot(L.v, w) / dot(L.v, L.v);
        if (a1 < R(0) || a1 > R(1) || a2 < R(0) || a2 > R(1)) return false;
        t_out = t; cos_out = -dn;
        return true;
    }
    // scene.rs:36-86 — default method; `dist` starts from the (possibly stale) state.hit_dist
    bool sample_lights(const Ray<R>& ray, State<R>& state, LightSampleRec<R>& light_sample,
                       const std::vector<Light<R>>& lights) const {
        bool hit = sample_lights_spherical(ray, state, light_sample, lights);
        if (extended_lights() && state.hit_dist > R(0)) {         // quads after the spheres, nearest wins (extension)
            R dist = state.hit_dist;
            for (const Light<R>& light : lights) {
                if (light.type != PTB_LIGHT_RECTANGULAR) continue;
                R d, c;
                if (light_rect(ray, light, d, c) && d < dist) {
                    dist = d;
                    light_sample.pdf = (dist * dist) / (light.area * c);
                    light_sample.emission = light.emission;
                    state.is_emitter = true;
                    state.hit_dist = d;
                    hit = true;
                }
            }
        }
        return hit;
    }
    bool sample_lights_spherical(const Ray<R>& ray, State<R>& state, LightSampleRec<R>& light_sample,
                                 const std::vector<Light<R>>& lights) const {
        bool hit = false;
        R dist = state.hit_dist;
        for (const Light<R>& light : lights) {
            if (light.type == PTB_LIGHT_SPHERICAL) {
                R d;
                if (light_sphere(ray, light.position, light.radius, d)) {
                    if (d < dist) {
                        dist = d;
                        V3<R> hit_point = ray.at(d);
                        R cos_theta = dot(-ray.direction, normalize(hit_point - light.position));
                        light_sample.pdf = (dist * dist) / (light.area * cos_theta * R(0.5));
                        light_sample.emission = light.emission;
                        state.is_emitter = true;
                        state.hit_dist = d;
                        hit = true;
                    }
                }
            }
        }
        return hit;
    }
};

// analytical.rs:166-190
template <class R> inline bool isect_sphere(const Ray<R>& ray, V3<R> center, R radius, R& t_out) {
    V3<R> l = center - ray.origin;
    R tca = l.dot(ray.direction);
    R d2 = l.dot(l) - tca * tca;
    R radius2 = radius * radius;
    if (d2 > radius2) return false;
    R thc = std::sqrt(radius2 - d2);
    R t0 = tca - thc, t1 = tca + thc;
    if (t0 > t1) { R tmp = t0; t0 = t1; t1 = tmp; }
    if (t0 < R(0)) { t0 = t1; if (t0 < R(0)) return false; }
    t_out = t0;
    return true;
}
// analytical.rs:193-204 generalised to (point, normal); for point (0,-1,0), normal (0,1,0) the
// arithmetic reduces to the reference's bit for bit (the extra terms are exact zeros).
template <class R> inline bool isect_plane(const Ray<R>& ray, V3<R> point, V3<R> normal, R& t_out) {
    R denom = dot(normal, ray.direction);
    if (std::fabs(denom) > R(0.0001)) {
        R t = dot(point - ray.origin, normal) / denom;
        if (t >= R(0)) { t_out = t; return true; }
    }
    return false;
}
// analytical.rs:107-111
template <class R> inline R checker(R x, R y, R a, R b) {
    R x1 = std::fmod(std::floor(x), R(2));
    R y1 = std::fmod(std::floor(y), R(2));
    return (std::fmod(x1 + y1, R(2)) < R(1)) ? a : b;
}

// ------------------------------------------------------------------------------------------------
// renderer/src/analytical.rs — the reference's only Scene impl, restated LITERALLY (constants
// in place) so the data-driven FlatScene below can be checked against it bit for bit.
template <class R> struct AnalyticalSceneLiteral : Scene<R> {
    std::vector<Light<R>> lights;
    Pinhole<R> pinhole;
    AnalyticalSceneLiteral() {                                              // analytical.rs:13-22
        R em = 3;
        lights.push_back(Light<R>::spherical(V3<R>(3, 2, 2), R(1), V3<R>(em, em, em)));
    }
    const Pinhole<R>& camera() const override { return pinhole; }
    V3<R> background(const Ray<R>& ray) const override {                    // analytical.rs:28-32
        R t = R(0.5) * (ray.direction.y + R(1));
        return this->to_linear((R(1) - t) * V3<R>(1, 1, 1) + t * V3<R>(R(0.5), R(0.7), R(1.0))) * V3<R>::new_x(R(0.5));
    }
    bool closest_hit(const Ray<R>& ray, State<R>& state, LightSampleRec<R>& light_sample) const override {
        R dist = std::numeric_limits<R>::max();                             // analytical.rs:38
        bool hit = false;
        V3<R> center(R(-1.1), 0, 0);
        R d;
        if (isect_sphere(ray, center, R(1), d)) {                           // analytical.rs:43-68
            V3<R> hp = ray.at(d);
            V3<R> normal = normalize(hp - center);
            state.hit_dist = d;
            state.normal = normal;
            state.material.rgb = V3<R>::new_x(R(1));
            state.material.roughness = R(0.05);
            state.material.metallic = R(1);
            state.material_index = 0;
            hit = true;
            dist = d;
        }
        center = V3<R>(R(1.1), 0, 0);
        if (isect_sphere(ray, center, R(1), d)) {                           // analytical.rs:72-99
            if (d < dist) {
                V3<R> hp = ray.at(d);
                V3<R> normal = normalize(hp - center);
                state.hit_dist = d;
                state.normal = normal;
                state.material.rgb = V3<R>(R(1.0), R(0.186), R(0.0));
                state.material.clearcoat = R(1);
                state.material.clearcoat_gloss = R(1);
                state.material.roughness = R(0.1);
                state.material_index = 1;
                hit = true;
                dist = d;
            }
        }
        if (isect_plane(ray, V3<R>(0, -1, 0), V3<R>(0, 1, 0), d)) {         // analytical.rs:101-120
            if (d < dist) {
                state.hit_dist = d;
                state.normal = V3<R>(0, 1, 0);
                R c = checker(ray.direction.x / ray.direction.y * R(0.5) + R(100),
                              ray.direction.z / ray.direction.y * R(0.5) + R(100), R(0.25), R(0.1));
                state.material.rgb = V3<R>(c, c, c);
                state.material.roughness = R(1);
                state.material_index = 2;
                hit = true;
            }
        }
        if (this->sample_lights(ray, state, light_sample, lights)) hit = true;   // analytical.rs:122
        return hit;
    }
    bool any_hit(const Ray<R>& ray, R) const override {                      // analytical.rs:130-145
        R d;
        if (isect_sphere(ray, V3<R>(R(-1.1), 0, 0), R(1), d)) return true;
        if (isect_sphere(ray, V3<R>(R(1.1), 0, 0), R(1), d)) return true;
        if (isect_plane(ray, V3<R>(0, -1, 0), V3<R>(0, 1, 0), d)) return true;
        return false;
    }
    size_t number_of_lights() const override { return lights.size(); }
    const Light<R>& light_at(size_t i) const override { return lights[i]; }
};

// ------------------------------------------------------------------------------------------------
// FlatScene — the same trait implemented from the POD scene export (include/ptb200.h), i.e. what
// the device path is given.  Semantics generalise analytical.rs:36-145: spheres in index order
// (first unconditional, then `d < dist`), then planes, then sample_lights; each accepted hit
// assigns the material fields in its set_mask on top of what earlier (farther) hits assigned.
template <class R> struct FlatTypes;
template <> struct FlatTypes<float> {
    using scene = ptb_scene_f32; using material = ptb_material_f32; using sphere = ptb_sphere_f32;
    using plane = ptb_plane_f32; using light = ptb_light_f32; using sdf = ptb_sdf_f32; using sdf_node = ptb_sdf_node_f32;
};
template <> struct FlatTypes<double> {
    using scene = ptb_scene_f64; using material = ptb_material_f64; using sphere = ptb_sphere_f64;
    using plane = ptb_plane_f64; using light = ptb_light_f64; using sdf = ptb_sdf_f64; using sdf_node = ptb_sdf_node_f64;
};

// Signed-distance program (include/ptb200.h, PTB_SDF_*).  The reference has no SDF scene (Readme.md:18 lists one as open), so
// this is not a restatement of reference code: it is the CPU statement of the extension the library defines, written once
// here in plain C++ and checked against the device evaluator operation for operation (tests/test_gpu_sdf.py).
template <class R> struct SdfProgram {
    std::vector<typename FlatTypes<R>::sdf_node> nodes;
    R hit_eps = 0, max_dist = 0, normal_h = 0;
    uint32_t max_steps = 0;
    bool empty() const { return nodes.empty(); }
    static R len2(R x, R y) { return std::sqrt(x * x + y * y); }
    static R len3(R x, R y, R z) { return std::sqrt(x * x + y * y + z * z); }
    static R maxr(R a, R b) { return std::fmax(a, b); }
    static R minr(R a, R b) { return a < b ? a : b; }
    // (distance, material) at q
    void eval(const V3<R>& q, R& dist, uint32_t& material) const {
        R sd[PTB_SDF_MAX_STACK]; uint32_t sm[PTB_SDF_MAX_STACK];
        int sp = 0;
        for (const auto& n : nodes) {
            const R x = q.x - n.p[0], y = q.y - n.p[1], z = q.z - n.p[2];
            R d; uint32_t m = n.material;
            switch (n.op) {
                case PTB_SDF_SPHERE: d = len3(x, y, z) - n.a[0]; break;
                case PTB_SDF_BOX: {
                    const R dx = std::fabs(x) - n.a[0], dy = std::fabs(y) - n.a[1], dz = std::fabs(z) - n.a[2];
                    const R outside = len3(maxr(dx, R(0)), maxr(dy, R(0)), maxr(dz, R(0)));
                    d = outside + minr(maxr(dx, maxr(dy, dz)), R(0)) - n.a[3];
                    break;
                }
                case PTB_SDF_TORUS: d = len2(len2(x, z) - n.a[0], y) - n.a[1]; break;
                case PTB_SDF_PLANE: d = q.x * n.a[0] + q.y * n.a[1] + q.z * n.a[2] + n.a[3]; break;
                default: {
                    --sp; const R bd = sd[sp]; const uint32_t bm = sm[sp];
                    --sp; const R ad = sd[sp]; const uint32_t am = sm[sp];
                    if (n.op == PTB_SDF_UNION) { if (bd < ad) { d = bd; m = bm; } else { d = ad; m = am; } }
                    else if (n.op == PTB_SDF_INTERSECT) { if (bd > ad) { d = bd; m = bm; } else { d = ad; m = am; } }
                    else if (n.op == PTB_SDF_SUBTRACT) { d = maxr(ad, -bd); m = am; }
                    else {
                        R h = R(0.5) + R(0.5) * ((bd - ad) / n.a[0]);
                        h = h < R(0) ? R(0) : (h > R(1) ? R(1) : h);
                        d = ((R(1) - h) * bd + ad * h) - n.a[0] * h * (R(1) - h);
                        m = h >= R(0.5) ? am : bm;
                    }
                }
            }
            sd[sp] = d; sm[sp] = m; ++sp;
        }
        dist = sd[0]; material = sm[0];
    }
    // sphere tracing from t = 0; returns t >= 0 or -1
    R trace(const Ray<R>& ray, R limit, uint32_t& material) const {
        const R t_end = minr(limit, max_dist);
        R t = 0;
        for (uint32_t i = 0; i < max_steps; ++i) {
            R d; uint32_t m;
            eval(V3<R>(ray.origin.x + t * ray.direction.x, ray.origin.y + t * ray.direction.y, ray.origin.z + t * ray.direction.z), d, m);
            const R a = std::fabs(d);
            if (a < hit_eps) { material = m; return t; }
            t = t + a;
            if (!(t < t_end)) break;
        }
        return R(-1);
    }
    V3<R> normal(const V3<R>& q) const {
        const R h = normal_h;
        R d0, d1, d2, d3; uint32_t m;
        eval(V3<R>(q.x + h, q.y - h, q.z - h), d0, m); eval(V3<R>(q.x - h, q.y - h, q.z + h), d1, m);
        eval(V3<R>(q.x - h, q.y + h, q.z - h), d2, m); eval(V3<R>(q.x + h, q.y + h, q.z + h), d3, m);
        return normalize(V3<R>(((d0 - d1) - d2) + d3, ((d2 + d3) - d0) - d1, ((d1 + d3) - d0) - d2));
    }
};

template <class R> struct FlatScene : Scene<R> {
    using T = FlatTypes<R>;
    std::vector<typename T::sphere> spheres;
    std::vector<typename T::plane> planes;
    std::vector<typename T::material> materials;
    std::vector<Light<R>> lights;
    Pinhole<R> pinhole;
    uint32_t bg_kind = 0;
    V3<R> bg_a, bg_b; R bg_scale = 1, bg_gamma = 1;
    uint16_t depth = 4;
    uint32_t flags = 0;
    R eps = R(0.005);
    SdfProgram<R> sdf;               // optional signed-distance body, tested after the planes (ptb_set_sdf_*)

    void set_sdf(const typename T::sdf* p) {
        sdf = SdfProgram<R>();
        if (!p || p->n_nodes == 0) return;
        sdf.nodes.assign(p->nodes, p->nodes + p->n_nodes);
        sdf.hit_eps = p->hit_eps; sdf.max_dist = p->max_dist; sdf.normal_h = p->normal_h; sdf.max_steps = p->max_steps;
    }

    explicit FlatScene(const typename T::scene& s) {
        spheres.assign(s.spheres, s.spheres + s.n_spheres);
        planes.assign(s.planes, s.planes + s.n_planes);
        materials.assign(s.materials, s.materials + s.n_materials);
        for (uint32_t i = 0; i < s.n_lights; ++i) {
            const auto& l = s.lights[i];
            Light<R> L = Light<R>::spherical(V3<R>(l.position[0], l.position[1], l.position[2]), l.radius,
                                             V3<R>(l.emission[0], l.emission[1], l.emission[2]));
            L.type = l.type;
            L.u = V3<R>(l.u[0], l.u[1], l.u[2]); L.v = V3<R>(l.v[0], l.v[1], l.v[2]);
            if (l.type == PTB_LIGHT_RECTANGULAR) {
                const R cx = l.u[1] * l.v[2] - l.u[2] * l.v[1], cy = l.u[2] * l.v[0] - l.u[0] * l.v[2], cz = l.u[0] * l.v[1] - l.u[1] * l.v[0];
                L.area = std::sqrt(cx * cx + cy * cy + cz * cz);
            } else if (l.type == PTB_LIGHT_DISTANT) {
                L.area = R(0);
            }
            lights.push_back(L);
        }
        pinhole.origin = V3<R>(s.camera.origin[0], s.camera.origin[1], s.camera.origin[2]);
        pinhole.center = V3<R>(s.camera.center[0], s.camera.center[1], s.camera.center[2]);
        pinhole.fov = s.camera.fov;
        bg_kind = s.background.kind;
        bg_a = V3<R>(s.background.colour_a[0], s.background.colour_a[1], s.background.colour_a[2]);
        bg_b = V3<R>(s.background.colour_b[0], s.background.colour_b[1], s.background.colour_b[2]);
        bg_scale = s.background.scale;
        bg_gamma = s.background.gamma;
        depth = (uint16_t)s.depth;
        flags = s.flags;
        eps = s.eps;
    }
    const Pinhole<R>& camera() const override { return pinhole; }
    uint16_t recursion_depth() const override { return depth; }
    bool extended_lights() const override { return (flags & PTB_SCENE_EXTENDED_LIGHTS) != 0; }
    V3<R> background(const Ray<R>& ray) const override {
        if (bg_kind == PTB_BG_GRADIENT_Y) {
            R t = R(0.5) * (ray.direction.y + R(1));
            V3<R> c = (R(1) - t) * bg_a + t * bg_b;
            return V3<R>(std::pow(c.x, bg_gamma), std::pow(c.y, bg_gamma), std::pow(c.z, bg_gamma)) * V3<R>::new_x(bg_scale);
        }
        return bg_a;
    }
    void apply_material(State<R>& state, uint32_t mi, const Ray<R>& ray) const {
        const auto& m = materials[mi];
        Material<R>& o = state.material;
        uint32_t k = m.set_mask;
        if (k & PTB_MAT_RGB) {
            if (m.albedo_kind == PTB_ALBEDO_CHECKER_DIR_RATIO) {
                R c = checker(ray.direction.x / ray.direction.y * m.checker_scale + m.checker_offset,
                              ray.direction.z / ray.direction.y * m.checker_scale + m.checker_offset,
                              m.checker_a, m.checker_b);
                o.rgb = V3<R>(c, c, c);
            } else {
                o.rgb = V3<R>(m.rgb[0], m.rgb[1], m.rgb[2]);
            }
        }
        if (k & PTB_MAT_EMISSION) o.emission = V3<R>(m.emission[0], m.emission[1], m.emission[2]);
        if (k & PTB_MAT_ANISOTROPIC) o.anisotropic = m.anisotropic;
        if (k & PTB_MAT_METALLIC) o.metallic = m.metallic;
        if (k & PTB_MAT_ROUGHNESS) o.roughness = m.roughness;
        if (k & PTB_MAT_SUBSURFACE) o.subsurface = m.subsurface;
        if (k & PTB_MAT_SPECULAR_TINT) o.specular_tint = m.specular_tint;
        if (k & PTB_MAT_SHEEN) o.sheen = m.sheen;
        if (k & PTB_MAT_SHEEN_TINT) o.sheen_tint = m.sheen_tint;
        if (k & PTB_MAT_CLEARCOAT) o.clearcoat = m.clearcoat;
        if (k & PTB_MAT_CLEARCOAT_GLOSS) o.clearcoat_gloss = m.clearcoat_gloss;
        if (k & PTB_MAT_SPEC_TRANS) o.spec_trans = m.spec_trans;
        if (k & PTB_MAT_IOR) o.ior = m.ior;
        // the medium belongs to the body: always the hit primitive's own (no set_mask bit)
        o.medium.type = m.medium_type; o.medium.density = m.medium_density;
        o.medium.color = V3<R>(m.medium_color[0], m.medium_color[1], m.medium_color[2]); o.medium.anisotropy = m.medium_anisotropy;
        state.material_index = (int)mi;
    }
    bool closest_hit(const Ray<R>& ray, State<R>& state, LightSampleRec<R>& light_sample) const override {
        R dist = std::numeric_limits<R>::max();
        bool hit = false;
        for (size_t i = 0; i < spheres.size(); ++i) {
            const auto& s = spheres[i];
            V3<R> center(s.center[0], s.center[1], s.center[2]);
            R d;
            if (isect_sphere(ray, center, (R)s.radius, d)) {
                if (i == 0 || d < dist) {
                    V3<R> hp = ray.at(d);
                    state.hit_dist = d;
                    state.normal = normalize(hp - center);
                    apply_material(state, s.material, ray);
                    hit = true;
                    dist = d;
                }
            }
        }
        for (size_t i = 0; i < planes.size(); ++i) {
            const auto& p = planes[i];
            R d;
            if (isect_plane(ray, V3<R>(p.point[0], p.point[1], p.point[2]), V3<R>(p.normal[0], p.normal[1], p.normal[2]), d)) {
                if (d < dist) {
                    state.hit_dist = d;
                    state.normal = V3<R>(p.normal[0], p.normal[1], p.normal[2]);
                    apply_material(state, p.material, ray);
                    hit = true;
                    // analytical.rs:103-119 does not update `dist` after the plane; with more than
                    // one plane the generalisation must, or a farther plane could overwrite a
                    // nearer one.  Identical for the single-plane demo scene.
                    dist = d;
                }
            }
        }
        if (!sdf.empty()) {
            uint32_t mi = 0;
            const R d = sdf.trace(ray, dist, mi);
            if (d >= R(0) && d < dist) {
                state.hit_dist = d;
                state.normal = sdf.normal(ray.at(d));
                apply_material(state, mi, ray);
                hit = true;
                dist = d;
            }
        }
        if (this->sample_lights(ray, state, light_sample, lights)) hit = true;
        return hit;
    }
    bool any_hit(const Ray<R>& ray, R max_dist) const override {
        if (!sdf.empty()) {
            uint32_t mi;
            if (sdf.trace(ray, max_dist, mi) >= R(0)) return true;
        }
        const bool ignore = (flags & PTB_SCENE_ANYHIT_IGNORES_MAX_DIST) != 0;
        R d;
        for (const auto& s : spheres)
            if (isect_sphere(ray, V3<R>(s.center[0], s.center[1], s.center[2]), (R)s.radius, d))
                if (ignore || d < max_dist) return true;
        for (const auto& p : planes)
            if (isect_plane(ray, V3<R>(p.point[0], p.point[1], p.point[2]), V3<R>(p.normal[0], p.normal[1], p.normal[2]), d))
                if (ignore || d < max_dist) return true;
        return false;
    }
    size_t number_of_lights() const override { return lights.size(); }
    const Light<R>& light_at(size_t i) const override { return lights[i]; }
};

// ------------------------------------------------------------------------------------------------
// buffer.rs:6-64
template <class R> struct ColorBuffer {
    size_t width, height;
    std::vector<R> pixels;
    size_t frames = 0;
    ColorBuffer(size_t w, size_t h) : width(w), height(h), pixels(w * h * 4, R(0)) {}
};
// Rust `as u8`: saturating, NaN -> 0, truncating
template <class R> inline uint8_t as_u8(R v) {
    if (!(v == v)) return 0;
    if (v <= R(0)) return 0;
    if (v >= R(255)) return 255;
    return (uint8_t)v;
}
// buffer.rs:55-64
template <class R> inline void convert_to_u8(const R* pixels, size_t n_pixels, uint8_t* frame) {
    for (size_t p = 0; p < n_pixels; ++p) {
        size_t o = p * 4;
        frame[o + 0] = as_u8<R>(std::pow(pixels[o + 0], R(0.4545)) * R(255));
        frame[o + 1] = as_u8<R>(std::pow(pixels[o + 1], R(0.4545)) * R(255));
        frame[o + 2] = as_u8<R>(std::pow(pixels[o + 2], R(0.4545)) * R(255));
        frame[o + 3] = as_u8<R>(pixels[o + 3] * R(255));
    }
}
// buffer.rs:67-102 (frame is at.2 x at.3; no gamma; strict > bounds; row j counted from the END)
template <class R> inline void convert_to_u8_at(const R* pixels, size_t bw, size_t bh, uint8_t* frame,
                                                size_t at0, size_t at1, size_t width, size_t height) {
    for (size_t j = 0; j < height; ++j) {
        uint8_t* line = frame + (height - 1 - j) * width * 4;   // par_rchunks_exact_mut: j = 0 is the last row
        for (size_t ii = 0; ii < width; ++ii) {
            size_t i = j * width + ii;
            size_t x = i % width;
            size_t y = height - (i / width);
            if (x > at0 && x < at0 + bw) {
                if (y > at1 && y < at1 + bh) {
                    size_t o = (x - at0) * 4 + (y - at1) * bw * 4;
                    uint8_t* px = line + ii * 4;
                    px[0] = as_u8<R>(pixels[o] * R(255));
                    px[1] = as_u8<R>(pixels[o + 1] * R(255));
                    px[2] = as_u8<R>(pixels[o + 2] * R(255));
                    px[3] = as_u8<R>(pixels[o + 3] * R(255));
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// tracer.rs — the integrator
template <class R> struct Tracer {
    R eps = R(0.005);                                       // tracer.rs:16
    const Scene<R>* scene;
    uint64_t seed = 0;
    explicit Tracer(const Scene<R>* s) : scene(s) {}

    // tracer.rs:222-231
    static R power_heuristic(R a, R b) { R t = a * a; return t / (b * b + t); }
    static R mix_ptf(R a, R b, R v) { return (R(1) - v) * a + b * v; }
    // tracer.rs:233-240 — log2, not ln (quirk A.3)
    static R gtr1(R ndoth, R a) {
        if (a >= R(1)) return K<R>::INV_PI;
        R a2 = a * a;
        R t = R(1) + (a2 - R(1)) * ndoth * ndoth;
        return (a2 - R(1)) / (K<R>::PI * std::log2(a2) * t);
    }
    // tracer.rs:242-254 — r2 unused (quirk A.4)
    static V3<R> sample_gtr1(R rgh, R r1, R) {
        R a = fmax_(R(0.001), rgh);
        R a2 = a * a;
        R phi = r1 * K<R>::TWO_PI;
        R cos_theta = std::sqrt((R(1) - std::pow(a2, R(1) - r1)) / (R(1) - a2));
        R sin_theta = clamp_(std::sqrt(R(1) - (cos_theta * cos_theta)), R(0), R(1));
        R sin_phi = std::sin(phi);
        R cos_phi = std::cos(phi);
        return V3<R>(sin_theta * cos_phi, sin_theta * sin_phi, cos_theta);
    }
    // tracer.rs:256-274
    static V3<R> sample_ggxvndf(const V3<R>& v, R ax, R ay, R r1, R r2) {
        V3<R> vh = normalize(V3<R>(ax * v.x, ay * v.y, v.z));
        R lensq = vh.x * vh.x + vh.y * vh.y;
        V3<R> t_1 = lensq > R(0) ? V3<R>(-vh.y, vh.x, 0).mult_f(R(1) / std::sqrt(lensq)) : V3<R>(1, 0, 0);
        V3<R> t_2 = cross(vh, t_1);
        R r = std::sqrt(r1);
        R phi = R(2) * K<R>::PI * r2;
        R t1 = r * std::cos(phi);
        R t2 = r * std::sin(phi);
        R s = R(0.5) * (R(1) + vh.z);
        t2 = (R(1) - s) * std::sqrt(R(1) - t1 * t1) + s * t2;
        V3<R> nh = t1 * t_1 + t2 * t_2 + std::sqrt(fmax_(R(0), R(1) - t1 * t1 - t2 * t2)) * vh;
        return normalize(V3<R>(ax * nh.x, ay * nh.y, fmax_(R(0), nh.z)));
    }
    // tracer.rs:276-280
    static R smithg(R ndotv, R alphag) {
        R a = alphag * alphag;
        R b = ndotv * ndotv;
        return (R(2) * ndotv) / (ndotv + std::sqrt(a + b - a * b));
    }
    // tracer.rs:284-286
    static R luminance(const V3<R>& c) { return R(0.212671) * c.x + R(0.715160) * c.y + R(0.072169) * c.z; }
    // tracer.rs:288-292
    static R schlick_fresnel(R u) {
        R m = clamp_(R(1) - u, R(0), R(1));
        R m2 = m * m;
        return m2 * m2 * m;
    }
    // tracer.rs:294-299
    static R gtr2aniso(R ndoth, R hdotx, R hdoty, R ax, R ay) {
        R a = hdotx / ax;
        R b = hdoty / ay;
        R c = a * a + b * b + ndoth * ndoth;
        return R(1) / (K<R>::PI * ax * ay * c * c);
    }
    // tracer.rs:301-306
    static R smithganiso(R ndotv, R vdotx, R vdoty, R ax, R ay) {
        R a = vdotx * ax;
        R b = vdoty * ay;
        R c = ndotv;
        return (R(2) * ndotv) / (ndotv + std::sqrt(a * a + b * b + c * c));
    }
    // tracer.rs:308-322
    static R dielectric_fresnel(R cos_theta_i, R eta) {
        R sin_theta_tsq = eta * eta * (R(1) - cos_theta_i * cos_theta_i);
        if (sin_theta_tsq > R(1)) return R(1);
        R cos_theta_t = std::sqrt(fmax_(R(1) - sin_theta_tsq, R(0)));
        R rs = (eta * cos_theta_t - cos_theta_i) / (eta * cos_theta_t + cos_theta_i);
        R rp = (eta * cos_theta_i - cos_theta_t) / (eta * cos_theta_i + cos_theta_t);
        return R(0.5) * (rs * rs + rp * rp);
    }
    // tracer.rs:324-333
    static V3<R> cosine_sample_hemisphere(R r1, R r2) {
        V3<R> dir;
        R r = std::sqrt(r1);
        R phi = K<R>::TWO_PI * r2;
        dir.x = r * std::cos(phi);
        dir.y = r * std::sin(phi);
        dir.z = std::sqrt(fmax_(R(0), R(1) - dir.x * dir.x - dir.y * dir.y));
        return dir;
    }
    // tracer.rs:335-341
    static void get_spec_color(const Material<R>& material, R eta, V3<R>& spec_col, V3<R>& sheen_col) {
        R lum = luminance(material.rgb);
        V3<R> ctint = lum > R(0) ? material.rgb / V3<R>::new_x(lum) : V3<R>(1, 1, 1);
        R f0 = (R(1) - eta) / (R(1) + eta);
        spec_col = mix(f0 * f0 * mix(V3<R>(1, 1, 1), ctint, material.specular_tint), material.rgb, material.metallic);
        sheen_col = mix(V3<R>(1, 1, 1), ctint, material.sheen_tint);
    }
    // tracer.rs:435-439
    static R disney_fresnel(const Material<R>& material, R eta, R ldoth, R vdoth) {
        R metallic_fresnel = schlick_fresnel(ldoth);
        R dielectric = dielectric_fresnel(std::fabs(vdoth), eta);
        return mix_ptf(dielectric, metallic_fresnel, material.metallic);
    }
    // tracer.rs:343-366
    static V3<R> eval_diffuse(const Material<R>& material, const V3<R>& c_sheen, const V3<R>& v, const V3<R>& l,
                              const V3<R>& h, R& pdf) {
        pdf = 0;
        if (l.z <= R(0)) return V3<R>::zeros();
        R fl = schlick_fresnel(l.z);
        R fv = schlick_fresnel(v.z);
        R fh = schlick_fresnel(dot(l, h));
        R fd90 = R(0.5) + R(2) * dot(l, h) * dot(l, h) * material.roughness;
        R fd = mix_ptf(R(1), fd90, fl) * mix_ptf(R(1), fd90, fv);
        R fss90 = dot(l, h) * dot(l, h) * material.roughness;
        R fss = mix_ptf(R(1), fss90, fl) * mix_ptf(R(1), fss90, fv);
        R ss = R(1.25) * (fss * (R(1) / (l.z + v.z) - R(0.5)) + R(0.5));
        V3<R> fsheen = fh * material.sheen * c_sheen;
        pdf = l.z * K<R>::INV_PI;
        return (R(1) - material.metallic) * (R(1) - material.spec_trans) *
               (K<R>::INV_PI * mix_ptf(fd, ss, material.subsurface) * material.rgb + fsheen);
    }
    // tracer.rs:368-382
    static V3<R> eval_spec_reflection(const Material<R>& material, R eta, const V3<R>& spec_col, const V3<R>& v,
                                      const V3<R>& l, const V3<R>& h, R& pdf) {
        pdf = 0;
        if (l.z <= R(0)) return V3<R>::zeros();
        R fm = disney_fresnel(material, eta, dot(l, h), dot(v, h));
        V3<R> f = mix(spec_col, V3<R>(1, 1, 1), fm);
        R d = gtr2aniso(h.z, h.x, h.y, material.ax, material.ay);
        R g1 = smithganiso(std::fabs(v.z), v.x, v.y, material.ax, material.ay);
        R g2 = g1 * smithganiso(std::fabs(l.z), l.x, l.y, material.ax, material.ay);
        pdf = g1 * d / (R(4) * v.z);
        return d * g2 * f / V3<R>::new_x(R(4) * l.z * v.z);
    }
    // tracer.rs:384-402
    static V3<R> eval_spec_refraction(const Material<R>& material, R eta, const V3<R>& v, const V3<R>& l,
                                      const V3<R>& h, R& pdf) {
        pdf = 0;
        if (l.z >= R(0)) return V3<R>::zeros();
        R f = dielectric_fresnel(std::fabs(dot(v, h)), eta);
        R d = gtr2aniso(h.z, h.x, h.y, material.ax, material.ay);
        R g1 = smithganiso(std::fabs(v.z), v.x, v.y, material.ax, material.ay);
        R g2 = g1 * smithganiso(std::fabs(l.z), l.x, l.y, material.ax, material.ay);
        R denom = dot(l, h) + dot(v, h) * eta;
        denom *= denom;
        R eta2 = eta * eta;
        R jacobian = std::fabs(dot(l, h)) / denom;
        pdf = g1 * fmax_(R(0), dot(v, h)) * d * jacobian / v.z;
        // Rust precedence: ((((((((1-m)*st)*(1-f))*d)*g2)*|v.h|)*jac)*eta2) / |l.z*v.z|  then * pow(rgb, .5)
        R s = (R(1) - material.metallic) * material.spec_trans * (R(1) - f) * d * g2 * std::fabs(dot(v, h)) * jacobian *
              eta2 / std::fabs(l.z * v.z);
        return s * powv(material.rgb, V3<R>(R(0.5), R(0.5), R(0.5)));
    }
    // tracer.rs:404-419
    static V3<R> eval_clearcoat(const Material<R>& material, const V3<R>& v, const V3<R>& l, const V3<R>& h, R& pdf) {
        pdf = 0;
        if (l.z <= R(0)) return V3<R>::zeros();
        R fh = dielectric_fresnel(dot(v, h), R(1) / R(1.5));
        R f = mix_ptf(R(0.04), R(1), fh);
        R d = gtr1(h.z, material.clearcoat_roughness);
        R g = smithg(l.z, R(0.25)) * smithg(v.z, R(0.25));
        R jacobian = R(1) / (R(4) * dot(v, h));
        pdf = d * h.z * jacobian;
        return material.clearcoat * f * d * g / (R(4) * l.z * v.z) * V3<R>(R(0.25), R(0.25), R(0.25));
    }
    // tracer.rs:421-433
    static void get_lobe_probabilities(const Material<R>& material, const V3<R>& spec_col, R approx_fresnel,
                                       R& diffuse_wt, R& spec_reflect_wt, R& spec_refract_wt, R& clearcoat_wt) {
        diffuse_wt = luminance(material.rgb) * (R(1) - material.metallic) * (R(1) - material.spec_trans);
        spec_reflect_wt = luminance(mix(spec_col, V3<R>(1, 1, 1), approx_fresnel));
        spec_refract_wt = (R(1) - approx_fresnel) * (R(1) - material.metallic) * material.spec_trans * luminance(material.rgb);
        clearcoat_wt = R(0.25) * material.clearcoat * (R(1) - material.metallic);
        R total_wt = diffuse_wt + spec_reflect_wt + spec_refract_wt + clearcoat_wt;
        diffuse_wt /= total_wt;
        spec_reflect_wt /= total_wt;
        spec_refract_wt /= total_wt;
        clearcoat_wt /= total_wt;
    }
    // tracer.rs:449-454 (also 184-189, 559-564)
    static void onb(const V3<R>& n, V3<R>& t, V3<R>& b) {
        V3<R> up = std::fabs(n.z) < R(0.999) ? V3<R>(0, 0, 1) : V3<R>(1, 0, 0);
        t = normalize(cross(up, n));
        b = cross(n, t);
    }
    static V3<R> to_local(const V3<R>& x, const V3<R>& y, const V3<R>& z, const V3<R>& v) {
        return V3<R>(dot(v, x), dot(v, y), dot(v, z));
    }
    static V3<R> to_world(const V3<R>& x, const V3<R>& y, const V3<R>& z, const V3<R>& v) {
        return v.x * x + v.y * y + v.z * z;
    }
    static V3<R> reflect(V3<R> i, V3<R> n) { return i - V3<R>(2, 2, 2) * n * V3<R>::new_x(dot(n, i)); }  // tracer.rs:464
    static V3<R> refract(V3<R> i, V3<R> n, R eta) {                                                   // tracer.rs:468
        R k = R(1) - eta * eta * (R(1) - dot(n, i) * dot(n, i));
        if (k < R(0)) return V3<R>::zeros();
        return eta * i - (eta * dot(n, i) + std::sqrt(k)) * n;
    }

    // tracer.rs:441-553.  (r1, r2, coin) are the draws at 446, 447 and 534; `l` enters holding the
    // previous bounce's sampled direction (quirk A.5).  lobe_out: 0 diffuse, 1 clearcoat, 2 reflect,
    // 3 refract.
    V3<R> disney_sample(const State<R>& state, V3<R> v, const V3<R>& n, V3<R>& l, R& pdf, R r1, R r2, R coin,
                        int* lobe_out, Counters* ctr) const {
        pdf = 0;
        V3<R> f;
        V3<R> t, b;
        onb(n, t, b);
        v = to_local(t, b, n, v);
        V3<R> spec_col, sheen_col;
        get_spec_color(state.material, state.eta, spec_col, sheen_col);
        R diffuse_wt = 0, spec_reflect_wt = 0, spec_refract_wt = 0, clearcoat_wt = 0;
        R approx_fresnel = disney_fresnel(state.material, state.eta, v.z, v.z);
        get_lobe_probabilities(state.material, spec_col, approx_fresnel, diffuse_wt, spec_reflect_wt, spec_refract_wt,
                               clearcoat_wt);
        R cdf[4];
        cdf[0] = diffuse_wt;
        cdf[1] = cdf[0] + clearcoat_wt;
        cdf[2] = cdf[1] + spec_reflect_wt;
        cdf[3] = cdf[2] + spec_refract_wt;
        (void)cdf[3];
        int lobe;
        if (r1 < cdf[0]) {
            r1 /= cdf[0];
            l = cosine_sample_hemisphere(r1, r2);
            V3<R> h = normalize(l + v);
            f = eval_diffuse(state.material, sheen_col, v, l, h, pdf);
            pdf *= diffuse_wt;
            lobe = 0;
            if (ctr) { ctr->lobe_diffuse++; ctr->ev_diffuse++; }
        } else if (r1 < cdf[1]) {
            r1 = (r1 - cdf[0]) / (cdf[1] - cdf[0]);
            V3<R> h = sample_gtr1(state.material.clearcoat_roughness, r1, r2);
            if (h.z < R(0)) h = -h;
            l = normalize(reflect(-v, h));
            f = eval_clearcoat(state.material, v, l, h, pdf);
            pdf *= clearcoat_wt;
            lobe = 1;
            if (ctr) { ctr->lobe_clearcoat++; ctr->ev_clearcoat++; }
        } else {
            r1 = (r1 - cdf[1]) / (R(1) - cdf[1]);
            V3<R> h = sample_ggxvndf(v, state.material.ax, state.material.ay, r1, r2);
            if (h.z < R(0)) h = -h;
            R fresnel = disney_fresnel(state.material, state.eta, dot(l, h), dot(v, h));   // stale l (A.5)
            R ff = R(1) - ((R(1) - fresnel) * state.material.spec_trans * (R(1) - state.material.metallic));
            if (coin < ff) {
                l = normalize(reflect(-v, h));
                f = eval_spec_reflection(state.material, state.eta, spec_col, v, l, h, pdf);
                pdf *= ff;
                lobe = 2;
                if (ctr) { ctr->lobe_reflect++; ctr->ev_reflect++; }
            } else {
                l = normalize(refract(-v, h, state.eta));
                f = eval_spec_refraction(state.material, state.eta, v, l, h, pdf);
                pdf *= R(1) - ff;
                lobe = 3;
                if (ctr) { ctr->lobe_refract++; ctr->ev_refract++; }
            }
            pdf *= spec_reflect_wt + spec_refract_wt;
        }
        if (lobe_out) *lobe_out = lobe;
        l = to_world(t, b, n, l);
        return std::fabs(dot(n, l)) * f;
    }

    // tracer.rs:555-626
    V3<R> disney_eval(const State<R>& state, V3<R> v_world, const V3<R>& n, const V3<R>& l_world, R& bsdf_pdf,
                      Counters* ctr) const {
        bsdf_pdf = 0;
        V3<R> f = V3<R>::zeros();
        V3<R> t, b;
        onb(n, t, b);
        V3<R> v = to_local(t, b, n, v_world);
        V3<R> l = to_local(t, b, n, l_world);
        V3<R> h;
        if (l.z > R(0)) h = normalize(l + v); else h = normalize(l + state.eta * v);
        if (h.z < R(0)) h = -h;
        V3<R> spec_col, sheen_col;
        get_spec_color(state.material, state.eta, spec_col, sheen_col);
        R diffuse_wt = 0, spec_reflect_wt = 0, spec_refract_wt = 0, clearcoat_wt = 0;
        R fresnel = disney_fresnel(state.material, state.eta, dot(l, h), dot(v, h));
        get_lobe_probabilities(state.material, spec_col, fresnel, diffuse_wt, spec_reflect_wt, spec_refract_wt, clearcoat_wt);
        R pdf = 0;
        if (diffuse_wt > R(0) && l.z > R(0)) {
            f += eval_diffuse(state.material, sheen_col, v, l, h, pdf);
            bsdf_pdf += pdf * diffuse_wt;
            if (ctr) ctr->ev_diffuse++;
        }
        if (spec_reflect_wt > R(0) && l.z > R(0) && v.z > R(0)) {
            f += eval_spec_reflection(state.material, state.eta, spec_col, v, l, h, pdf);
            bsdf_pdf += pdf * spec_reflect_wt;
            if (ctr) ctr->ev_reflect++;
        }
        if (spec_refract_wt > R(0) && l.z < R(0)) {
            f += eval_spec_refraction(state.material, state.eta, v, l, h, pdf);
            bsdf_pdf += pdf * spec_refract_wt;
            if (ctr) ctr->ev_refract++;
        }
        if (clearcoat_wt > R(0) && l.z > R(0) && v.z > R(0)) {
            f += eval_clearcoat(state.material, v, l, h, pdf);
            bsdf_pdf += pdf * clearcoat_wt;
            if (ctr) ctr->ev_clearcoat++;
        }
        return std::fabs(l.z) * f;
    }

    // tracer.rs:173-220 — (r1, r2) are the draws at 191-192
    void sample_light(const Light<R>& light, const V3<R>& scatter_pos, LightSampleRec<R>& light_sample, R r1, R r2) const {
        if (light.type != PTB_LIGHT_SPHERICAL) {
            if (!scene->extended_lights()) return;                            // tracer.rs:217 `_ => {}`
            light_sample.emission = (R)scene->number_of_lights() * light.emission;
            if (light.type == PTB_LIGHT_RECTANGULAR) {                        // extension: uniform by area
                V3<R> surf = (light.position + r1 * light.u) + r2 * light.v;
                light_sample.direction = surf - scatter_pos;
                light_sample.dist = length(light_sample.direction);
                R dist_sq = light_sample.dist * light_sample.dist;
                light_sample.direction /= V3<R>::new_x(light_sample.dist);
                light_sample.normal = normalize(cross(light.u, light.v));
                light_sample.pdf = dist_sq / (light.area * std::fabs(dot(light_sample.normal, light_sample.direction)));
            } else {                                                          // distant
                light_sample.direction = normalize(light.position);
                light_sample.normal = normalize(scatter_pos - light.position);
                light_sample.dist = std::numeric_limits<R>::max();
                light_sample.pdf = R(1);
            }
            return;
        }
        V3<R> sphere_center_to_surface = scatter_pos - light.position;
        R dist_to_sphere_center = length(sphere_center_to_surface);
        // uniform_sample_hemisphere, tracer.rs:178-182
        R rr = std::sqrt(fmax_(R(0), R(1) - r1 * r1));
        R phi = K<R>::TWO_PI * r2;
        V3<R> sampled_dir(rr * std::cos(phi), rr * std::sin(phi), r1);
        sphere_center_to_surface /= V3<R>::new_x(dist_to_sphere_center);
        V3<R> t, b;
        onb(sphere_center_to_surface, t, b);
        sampled_dir = sampled_dir.x * t + sampled_dir.y * b + sampled_dir.z * sphere_center_to_surface;
        V3<R> light_surface_pos = light.position + light.radius * sampled_dir;
        light_sample.direction = light_surface_pos - scatter_pos;
        light_sample.dist = length(light_sample.direction);
        R dist_sq = light_sample.dist * light_sample.dist;
        light_sample.direction /= V3<R>::new_x(light_sample.dist);
        light_sample.normal = normalize(light_surface_pos - light.position);
        light_sample.emission = (R)scene->number_of_lights() * light.emission;
        light_sample.pdf = dist_sq / (light.area * R(0.5) * std::fabs(dot(light_sample.normal, light_sample.direction)));
    }

    // tracer.rs:126-170
    template <class RNG>
    V3<R> direct_light(const Ray<R>& ray, const State<R>& state, const RNG& rng, uint32_t bounce,
                       Counters* ctr) const {
        V3<R> ld = V3<R>::zeros();
        V3<R> scatter_pos = state.fhp + eps * state.ffnormal;
        ScatterSampleRec<R> scatter_sample;
        size_t number_lights = scene->number_of_lights();
        if (number_lights > 0) {
            R random = rng.draw(bounce, SLOT_LIGHT_PICK);
            random *= (R)scene->number_of_lights();
            size_t index = (size_t)random;
            const Light<R>& light = scene->light_at(index);
            LightSampleRec<R> light_sample;
            R r1 = rng.draw(bounce, SLOT_LIGHT_R1), r2 = rng.draw(bounce, SLOT_LIGHT_R2);
            sample_light(light, scatter_pos, light_sample, r1, r2);
            V3<R> li = light_sample.emission;
            if (dot(light_sample.direction, light_sample.normal) < R(0)) {
                Ray<R> shadow_ray(scatter_pos, light_sample.direction);
                if (ctr) ctr->any_hit++;
                bool in_shadow = scene->any_hit(shadow_ray, light_sample.dist - eps);
                if (!in_shadow) {
                    if (ctr) ctr->eval_calls++;
                    scatter_sample.f = disney_eval(state, -ray.direction, state.ffnormal, light_sample.direction,
                                                   scatter_sample.pdf, ctr);
                    R mis_weight = 1;
                    if (light.area > R(0)) mis_weight = power_heuristic(light_sample.pdf, scatter_sample.pdf);
                    if (scatter_sample.pdf > R(0)) {
                        ld += mis_weight * li * (scatter_sample.f / V3<R>::new_x(light_sample.pdf));
                        if (ctr) ctr->nee_contrib++;
                    }
                } else if (ctr) ctr->nee_shadowed++;
            } else if (ctr) ctr->nee_culled++;
        }
        return ld;
    }

    // ---- media (extension, PTB_MEDIUM_* in include/ptb200.h): Henyey-Greenstein, pbrt convention ----
    static R phase_hg(R cos_theta, R g) {
        const R denom = (R(1) + g * g) + (R(2) * g) * cos_theta;
        return (R(1) / (R(4) * K<R>::PI)) * ((R(1) - g * g) / (denom * std::sqrt(denom)));
    }
    static V3<R> sample_hg(const V3<R>& v, R g, R r1, R r2) {
        R cos_theta;
        if (std::fabs(g) < R(0.001)) cos_theta = R(1) - R(2) * r2;
        else {
            const R sqr = (R(1) - g * g) / ((R(1) + g) - (R(2) * g) * r2);
            cos_theta = -(((R(1) + g * g) - sqr * sqr) / (R(2) * g));
        }
        const R sin_theta = std::sqrt(fmax_(R(0), R(1) - cos_theta * cos_theta));
        const R phi = K<R>::TWO_PI * r1;
        V3<R> t, b;
        onb(v, t, b);
        return ((sin_theta * std::cos(phi)) * t + (sin_theta * std::sin(phi)) * b) + cos_theta * v;
    }
    // The medium's part of a bounce for a path inside `medium` that hit geometry at state.hit_dist.  Returns true when the bounce
    // happened in the medium (ray, throughput, radiance and scatter_sample.pdf updated), false when the surface is shaded next.
    template <class RNG>
    bool medium_step(const Medium<R>& medium, Ray<R>& ray, const State<R>& state, V3<R>& radiance, V3<R>& throughput,
                     ScatterSampleRec<R>& scatter_sample, const RNG& rng, uint32_t bounce, Counters* ctr) const {
        const R t = state.hit_dist, density = medium.density;
        const V3<R>& c = medium.color;
        if (medium.type == PTB_MEDIUM_ABSORB) {
            throughput = throughput * V3<R>(std::exp((-(R(1) - c.x) * t) * density), std::exp((-(R(1) - c.y) * t) * density),
                                            std::exp((-(R(1) - c.z) * t) * density));
            return false;
        }
        if (medium.type == PTB_MEDIUM_EMISSIVE) {
            radiance += ((t * density) * c) * throughput;
            return false;
        }
        const R u = rng.draw(bounce, SLOT_JITTER_Y);
        const R ff = -std::log(u) / density;
        const R sd = ff < t ? ff : t;
        if (!(sd < t)) return false;
        throughput = throughput * c;
        ray.origin = ray.origin + sd * ray.direction;
        // direct_light(ray, state, is_surface = false, rng): tracer.rs:125-170 with the phase function as f and pdf
        const size_t number_lights = scene->number_of_lights();
        const V3<R> back = -ray.direction;
        if (number_lights > 0) {
            R random = rng.draw(bounce, SLOT_LIGHT_PICK);
            random *= (R)number_lights;
            const Light<R>& light = scene->light_at((size_t)random);
            LightSampleRec<R> ls;
            const R r1 = rng.draw(bounce, SLOT_LIGHT_R1), r2 = rng.draw(bounce, SLOT_LIGHT_R2);
            sample_light(light, ray.origin, ls, r1, r2);
            if (dot(ls.direction, ls.normal) < R(0)) {
                if (ctr) ctr->any_hit++;
                if (!scene->any_hit(Ray<R>(ray.origin, ls.direction), ls.dist - eps)) {
                    const R ph = phase_hg(dot(back, ls.direction), medium.anisotropy);
                    R w = 1;
                    if (light.area > R(0)) w = power_heuristic(ls.pdf, ph);
                    if (ph > R(0)) {
                        radiance += ((w * ls.emission) * V3<R>::new_x(ph / ls.pdf)) * throughput;
                        if (ctr) ctr->nee_contrib++;
                    }
                } else if (ctr) ctr->nee_shadowed++;
            } else if (ctr) ctr->nee_culled++;
        }
        const R h1 = rng.draw(bounce, SLOT_BSDF_R1), h2 = rng.draw(bounce, SLOT_BSDF_R2);
        const V3<R> dir = sample_hg(back, medium.anisotropy, h1, h2);
        scatter_sample.pdf = phase_hg(dot(back, dir), medium.anisotropy);
        ray.direction = dir;
        return true;
    }

    // tracer.rs:44-103 — radiance of ONE sample of pixel (x, memory row r); W, H as F.
    // `pixel_id` keys the RNG (memory pixel index r*W + x), `sample` is the global sample index.
    V3<R> trace_sample(size_t x, size_t j, size_t width, R height, uint32_t pixel_id, uint64_t sample,
                       Counters* ctr) const {
        return trace_sample_with(x, j, width, height, CounterRng<R>(pixel_id, sample, seed), ctr);
    }
    template <class RNG>
    V3<R> trace_sample_with(size_t x, size_t j, size_t width, R height, const RNG& rng, Counters* ctr) const {
        // tracer.rs:34-46:  i = j*width + x with j counted from the LAST memory row
        size_t i = j * width + x;
        R xf = (R)(i % width);
        R yf = height - (R)(i / width);
        R xx = xf / (R)width;
        R yy = yf / height;
        R offx = rng.draw(0, SLOT_JITTER_X), offy = rng.draw(0, SLOT_JITTER_Y);
        Ray<R> ray = scene->camera().gen_ray(xx, R(1) - yy, offx, offy, (R)width, height);

        V3<R> radiance(0, 0, 0), throughput(1, 1, 1);
        State<R> state;
        LightSampleRec<R> light_sample;
        ScatterSampleRec<R> scatter_sample;
        state.depth = scene->recursion_depth();
        if (ctr) ctr->samples++;
        bool ended = false;
        bool in_medium = false;                                               // (extension) globals.rs:19 `State::medium`
        Medium<R> medium;
        for (uint32_t bounce = 0; bounce < state.depth; ++bounce) {
            state.material = Material<R>();                                   // tracer.rs:63
            if (ctr) ctr->closest_hit++;
            bool hit = scene->closest_hit(ray, state, light_sample);
            if (!hit) {
                radiance += scene->background(ray) * throughput;
                if (ctr) { ctr->end_sky++; ctr->background++; }
                ended = true;
                break;
            }
            state.finalize(ray);
            if (ctr) ctr->finalize++;
            if (in_medium && !state.is_emitter &&
                medium_step(medium, ray, state, radiance, throughput, scatter_sample, rng, bounce, ctr)) continue;   // (extension)
            radiance += state.material.emission * throughput;
            if (state.is_emitter) {
                R mis_weight = 1;
                if (state.depth > 0) mis_weight = power_heuristic(scatter_sample.pdf, light_sample.pdf);   // quirk A.2
                radiance += mis_weight * light_sample.emission * throughput;
                if (ctr) ctr->end_emitter++;
                ended = true;
                break;
            }
            if (ctr) ctr->shade++;
            radiance += direct_light(ray, state, rng, bounce, ctr) * throughput;
            R r1 = rng.draw(bounce, SLOT_BSDF_R1), r2 = rng.draw(bounce, SLOT_BSDF_R2), coin = rng.draw(bounce, SLOT_COIN);
            int lobe = -1;
            scatter_sample.f = disney_sample(state, -ray.direction, state.ffnormal, scatter_sample.l, scatter_sample.pdf,
                                             r1, r2, coin, &lobe, ctr);
            if (lobe >= 2) rng.coin_consumed();                                // tracer.rs:534: drawn in the spec lobe only
            if (scatter_sample.pdf > R(0)) {
                throughput = throughput * (scatter_sample.f / V3<R>::new_x(scatter_sample.pdf));
            } else {
                if (ctr) ctr->end_pdf++;
                ended = true;
                break;
            }
            ray.direction = scatter_sample.l;
            ray.origin = state.fhp + eps * ray.direction;
            if (state.material.medium.type != PTB_MEDIUM_NONE) {              // (extension) entering / leaving the body; no nesting
                in_medium = dot(ray.direction, state.normal) < R(0);
                medium = state.material.medium;
            }
        }
        if (!ended && ctr) ctr->end_depth++;
        return radiance;
    }

    // The same frame loop on per-thread ChaCha12 generators in call order — the reference's RNG cost and call pattern
    // (`thread_rng()` per pixel is a handle to the thread's generator, tracer.rs:44).  Not reproducible against the counter-RNG
    // image sample by sample (like the reference against itself); converges to the same image.
    void render_chacha(ColorBuffer<R>& buffer, uint64_t base_seed) const {
        const size_t width = buffer.width;
        const R height = (R)buffer.height;
        const size_t H = buffer.height;
        const R mixv = R(1) / (R)(buffer.frames + 1);
#pragma omp parallel
        {
#ifdef _OPENMP
            const uint64_t tid = (uint64_t)omp_get_thread_num();
#else
            const uint64_t tid = 0;
#endif
            ChaCha12Rng<R> rng(base_seed * 0x10001ull + tid * 0x9E3779B97F4A7C15ull + buffer.frames);
#pragma omp for schedule(dynamic, 1)
            for (long long jj = 0; jj < (long long)H; ++jj) {
                size_t j = (size_t)jj;
                size_t row = H - 1 - j;
                R* line = buffer.pixels.data() + row * width * 4;
                for (size_t x = 0; x < width; ++x) {
                    rng.have_coin = false;
                    V3<R> radiance = trace_sample_with(x, j, width, height, rng, nullptr);
                    R* pixel = line + x * 4;
                    R color[4] = {radiance.x, radiance.y, radiance.z, R(1)};
                    for (int c = 0; c < 4; ++c) pixel[c] = (R(1) - mixv) * pixel[c] + color[c] * mixv;
                }
            }
        }
        buffer.frames += 1;
    }

    // tracer.rs:22-123 — one frame (1 spp) accumulated as a running mean; rows are independent
    // tasks exactly like par_rchunks_exact_mut(width*4) (OpenMP schedule(dynamic,1)).
    void render(ColorBuffer<R>& buffer, Counters* total = nullptr, uint64_t sample_base = 0) const {
        const size_t width = buffer.width;
        const R height = (R)buffer.height;
        const size_t H = buffer.height;
        const uint64_t sample = sample_base + buffer.frames;
        const R mixv = R(1) / (R)(buffer.frames + 1);
#pragma omp parallel
        {
            Counters local;
#pragma omp for schedule(dynamic, 1)
            for (long long jj = 0; jj < (long long)H; ++jj) {
                size_t j = (size_t)jj;
                size_t row = H - 1 - j;                  // rchunks: chunk j is memory row H-1-j
                R* line = buffer.pixels.data() + row * width * 4;
                for (size_t x = 0; x < width; ++x) {
                    V3<R> radiance = trace_sample(x, j, width, height, (uint32_t)(row * width + x), sample,
                                                  total ? &local : nullptr);
                    R* pixel = line + x * 4;
                    R color[4] = {radiance.x, radiance.y, radiance.z, R(1)};
                    for (int c = 0; c < 4; ++c) pixel[c] = (R(1) - mixv) * pixel[c] + color[c] * mixv;   // tracer.rs:108-115
                }
            }
            if (total) {
#pragma omp critical
                total->add(local);
            }
        }
        buffer.frames += 1;
    }
};

}  // namespace pto
