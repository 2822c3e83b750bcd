// ptb_device.cuh — device-side building blocks of the B200 path-tracing loop (sm_100a).
//
// Everything here is __device__ code templated on the scalar type R (float | double), the device
// counterpart of the reference's `type F` switch (rust-pathtracer/src/lib.rs:5-6).  The functions
// are shared by the fused persistent integrator, the wavefront stage kernels and the per-function
// parity entry points, so what the tests measure is what the integrators run.
//
// Reference citations are relative to /root/reference/.  The code is NOT a transliteration:
// loop-invariant camera terms are hoisted to the host, the shading frame / specular colours are
// computed once per bounce and shared between next-event estimation and BSDF sampling, hit
// attributes are computed once for the winning primitive, and materials are resolved after the
// intersection loop from an "accepted primitive" bitmask.  All of these are value-preserving with
// respect to the reference's arithmetic (same operations on the same operands), up to FMA
// contraction and CUDA libm rounding, which the 1e-5 parity tests bound.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/ptb200.h"

namespace ptb {

// ------------------------------------------------------------------------------------------------
// scalar math wrappers
#define PTB_DEV __device__ __forceinline__
#define PTB_HD __host__ __device__ __forceinline__

PTB_DEV float m_sqrt(float x) { return sqrtf(x); }
PTB_DEV double m_sqrt(double x) { return sqrt(x); }
PTB_DEV float m_abs(float x) { return fabsf(x); }
PTB_DEV double m_abs(double x) { return fabs(x); }
PTB_DEV float m_max(float a, float b) { return fmaxf(a, b); }   // f32::max: NaN-ignoring
PTB_DEV double m_max(double a, double b) { return fmax(a, b); }
// powf as exp2(b*log2(a)) on the two MUFU instructions directly (lg2.approx / ex2.approx.ftz, no denormal pre-scaling:
// every base on this path is a colour or a squared roughness >= 1e-6): ~3 ulp for the |b*log2(a)| <= 20 this path produces,
// a fraction of the size of powf's special-case tree.  a == 0 -> 0, a < 0 -> NaN, a == 1 -> 1, as powf.
// PTB_IEEE: measurement build (tools/function_parity.py, profiles/r02_function_parity.md) — every deliberate approximation of
// the shipped build is replaced by the correctly rounded / libm operation (and the build adds -prec-div=true -prec-sqrt=true
// -ftz=false -fmad=false), so that what remains between it and the oracle is CUDA-libm-vs-glibc rounding alone.  The difference
// between the two builds' error tables is the cost of the approximations; what both share is the conditioning of the formulas.
// glibc's f32 powf / sinf / cosf / log2f (what Rust's f32 methods call on Linux) are correctly rounded on all but a ~1e-6 share of
// inputs; CUDA's are 1-2 ulp functions.  Evaluating in double and rounding once reproduces the correctly rounded value.
#ifdef PTB_IEEE
PTB_DEV float m_pow(float a, float b) { return (float)pow((double)a, (double)b); }
#else
PTB_DEV float m_pow(float a, float b) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(b * __log2f(a))); return r; }
#endif
PTB_DEV double m_pow(double a, double b) { return pow(a, b); }
#ifdef PTB_IEEE
PTB_DEV float m_log2(float a) { return (float)log2((double)a); }
#else
PTB_DEV float m_log2(float a) { return log2f(a); }
#endif
PTB_DEV double m_log2(double a) { return log2(a); }
// exp / ln of the medium code (PTB_MEDIUM_*): libm accuracy in both builds (media are not on the benchmarked path)
#ifdef PTB_IEEE
PTB_DEV float m_exp(float a) { return (float)exp((double)a); }
PTB_DEV float m_ln(float a) { return (float)log((double)a); }
#else
PTB_DEV float m_exp(float a) { return expf(a); }
PTB_DEV float m_ln(float a) { return logf(a); }
#endif
PTB_DEV double m_exp(double a) { return exp(a); }
PTB_DEV double m_ln(double a) { return log(a); }
PTB_DEV float m_floor(float a) { return floorf(a); }
PTB_DEV double m_floor(double a) { return floor(a); }
// IEEE division where a branch decision hangs on the last bit (checker cell, film coordinates)
PTB_DEV float div_rn(float a, float b) { return __fdiv_rn(a, b); }
PTB_DEV double div_rn(double a, double b) { return a / b; }
// sin/cos of an angle known to lie in [0, 2*pi] (every call site passes TWO_PI * u, u in [0,1)):
// quadrant reduction with a three-term Cody-Waite pi/2 and the Cephes sinf/cosf minimax kernels on
// [-pi/4, pi/4]; ~1 ulp, no large-argument slow path (sincosf's Payne-Hanek tail is dead code here).
#if defined(PTB_IEEE)
PTB_DEV void m_sincos(float x, float* s, float* c) { double sd, cd; sincos((double)x, &sd, &cd); *s = (float)sd; *c = (float)cd; }
#elif defined(PTB_MUFU_SINCOS)
// MUFU.SIN / MUFU.COS on the argument shifted into [-pi, pi) (sin x = -sin(x - pi), cos x = -cos(x - pi)): max abs error
// 2^-21.4 = 3.6e-7 there — inside the 1e-5 parity budget, outside the ~1 ulp of the minimax kernel below
PTB_DEV void m_sincos(float x, float* s, float* c) {
    float sn, cs;
    __sincosf(x - 3.14159265358979323846f, &sn, &cs);
    *s = -sn; *c = -cs;
}
#else
PTB_DEV void m_sincos(float x, float* s, float* c) {
    float kf = rintf(x * 0.636619772f);
    float r = fmaf(kf, -1.5703125f, x);
    r = fmaf(kf, -4.837512969970703125e-4f, r);
    r = fmaf(kf, -7.54978995489188216e-8f, r);
    float z = r * r;
    float sn = fmaf(fmaf(fmaf(-1.9515295891e-4f, z, 8.3321608736e-3f), z, -1.6666654611e-1f) * z, r, r);
    float cs = fmaf(fmaf(fmaf(2.443315711809948e-5f, z, -1.388731625493765e-3f), z, 4.166664568298827e-2f) * z, z, fmaf(-0.5f, z, 1.0f));
    int k = (int)kf;
    float a = (k & 1) ? cs : sn;
    float b = (k & 1) ? sn : cs;
    *s = (k & 2) ? -a : a;
    *c = ((k + 1) & 2) ? -b : b;
}
#endif
PTB_DEV void m_sincos(double a, double* s, double* c) { sincos(a, s, c); }
template <class R> PTB_DEV R m_clamp(R x, R lo, R hi) { return x < lo ? lo : (x > hi ? hi : x); }  // f32::clamp

template <class R> struct Const;
template <> struct Const<float> {
    static constexpr float PI = 3.14159265358979323846f;
    static constexpr float INV_PI = 1.0f / 3.14159265358979323846f;      // lib.rs:9, evaluated in F
    static constexpr float TWO_PI = 3.14159265358979323846f * 2.0f;      // lib.rs:10
    static constexpr float INV_4PI = 1.0f / (4.0f * 3.14159265358979323846f);
    static constexpr float MAXV = 3.402823466e+38f;
};
template <> struct Const<double> {
    static constexpr double PI = 3.14159265358979323846;
    static constexpr double INV_PI = 1.0 / 3.14159265358979323846;
    static constexpr double TWO_PI = 3.14159265358979323846 * 2.0;
    static constexpr double INV_4PI = 1.0 / (4.0 * 3.14159265358979323846);
    static constexpr double MAXV = 1.7976931348623157e+308;
};

// ------------------------------------------------------------------------------------------------
// F3 (fx.rs:209-515)
template <class R> struct V3 {
    R x, y, z;
    PTB_HD V3() {}
    PTB_HD V3(R a, R b, R c) : x(a), y(b), z(c) {}
};
template <class R> PTB_HD V3<R> operator+(V3<R> a, V3<R> b) { return V3<R>(a.x + b.x, a.y + b.y, a.z + b.z); }
template <class R> PTB_HD V3<R> operator-(V3<R> a, V3<R> b) { return V3<R>(a.x - b.x, a.y - b.y, a.z - b.z); }
template <class R> PTB_HD V3<R> operator*(V3<R> a, V3<R> b) { return V3<R>(a.x * b.x, a.y * b.y, a.z * b.z); }
template <class R> PTB_HD V3<R> operator*(R s, V3<R> a) { return V3<R>(s * a.x, s * a.y, s * a.z); }
template <class R> PTB_HD V3<R> operator-(V3<R> a) { return V3<R>(-a.x, -a.y, -a.z); }
template <class R> PTB_HD R dot(V3<R> a, V3<R> b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
template <class R> PTB_HD V3<R> cross(V3<R> a, V3<R> b) {
    return V3<R>(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
template <class R> PTB_DEV R length(V3<R> a) { return m_sqrt(dot(a, a)); }
// F3 / scalar and normalize (fx.rs:306-313: three divides by sqrt).  f32: one reciprocal (MUFU.RCP /
// MUFU.RSQ, <= 2 ulp) and three multiplies — a handful of instructions instead of three division
// sequences; f64 keeps IEEE division.
#ifdef PTB_IEEE
PTB_DEV float m_rcp(float x) { return __fdiv_rn(1.0f, x); }
#else
PTB_DEV float m_rcp(float x) { return __fdividef(1.0f, x); }
#endif
PTB_DEV double m_rcp(double x) { return 1.0 / x; }
// f32 quotient as MUFU.RCP + FMUL (div.approx: <= 2 ulp for 2^-126 <= |b| <= 2^126, the same error class as the
// div.full.f32 that `a / b` compiles to under -prec-div=false).  div.full spends two range compares and a predicated
// rescaling multiply per operand on denominators outside that range — 8 issue slots per quotient, a quarter of the shading
// stage's instructions — and no denominator of this path gets there without the reference's own result being inf / NaN
// already (lobe weight sums, roughness products, cosines, squared distances, pdfs).  f64 keeps IEEE division.
#if defined(PTB_IEEE)
PTB_DEV float m_div(float a, float b) { return __fdiv_rn(a, b); }
#elif defined(PTB_FULL_DIV)
PTB_DEV float m_div(float a, float b) { return a / b; }
#else
PTB_DEV float m_div(float a, float b) { return __fdividef(a, b); }
#endif
PTB_DEV double m_div(double a, double b) { return a / b; }
#ifdef PTB_IEEE
PTB_DEV V3<float> div_s(V3<float> a, float s) { return V3<float>(__fdiv_rn(a.x, s), __fdiv_rn(a.y, s), __fdiv_rn(a.z, s)); }
PTB_DEV V3<float> normalize(V3<float> a) { return div_s(a, __fsqrt_rn(dot(a, a))); }     // fx.rs:306-313
#else
PTB_DEV V3<float> div_s(V3<float> a, float s) { float r = m_rcp(s); return V3<float>(a.x * r, a.y * r, a.z * r); }
PTB_DEV V3<float> normalize(V3<float> a) { float r = rsqrtf(dot(a, a)); return V3<float>(a.x * r, a.y * r, a.z * r); }
#endif
PTB_DEV V3<double> div_s(V3<double> a, double s) { return V3<double>(a.x / s, a.y / s, a.z / s); }
PTB_DEV V3<double> normalize(V3<double> a) { return div_s(a, length(a)); }
template <class R> PTB_DEV V3<R> mix3(V3<R> a, V3<R> b, R v) {                          // math.rs:33-39
    R w = R(1) - v;
    return V3<R>(w * a.x + b.x * v, w * a.y + b.y * v, w * a.z + b.z * v);
}
template <class R> PTB_DEV R mix1(R a, R b, R v) { return (R(1) - v) * a + b * v; }     // tracer.rs:228-231

// ------------------------------------------------------------------------------------------------
// Counter RNG: Philox4x32-10, key = (pixel, sample lo), ctr = (block, sample hi, seed lo, seed hi).
// Eight draw slots per bounce (f32: two blocks of four):
//   0,1 jitter (tracer.rs:45; slot 0 of later bounces: Russian-roulette extension) | 2 light pick (137) | 3 reflect/refract coin (534)
//   4,5 light r1,r2 (191-192) | 6,7 bsdf r1,r2 (446-447)
// The four draws EVERY shaded bounce needs share the second block, so a shading stage computes one Philox block unless the
// scene has several lights or the material can refract (SURVEY.md §8d lists the slots in call order; same draws, regrouped).
// Bit-exact with the CPU checker by construction (integer arithmetic only; the uint->float conversions are exact).
enum : uint32_t { SLOT_JITTER_X = 0, SLOT_JITTER_Y = 1, SLOT_LIGHT_PICK = 2, SLOT_COIN = 3, SLOT_LIGHT_R1 = 4, SLOT_LIGHT_R2 = 5,
                  SLOT_BSDF_R1 = 6, SLOT_BSDF_R2 = 7 };
struct Philox4 { uint32_t v[4]; };
PTB_DEV Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#ifndef PTB_PHILOX_ROUNDS
#define PTB_PHILOX_ROUNDS 10      /* A/B knob only (tools/ab_variants.py): the oracle and every parity test use the 10-round generator */
#endif
#pragma unroll
    for (int r = 0; r < PTB_PHILOX_ROUNDS; ++r) {
        uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
        uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
        c0 = hi1 ^ c1 ^ k0;
        c1 = lo1;
        c2 = hi0 ^ c3 ^ k1;
        c3 = lo0;
        k0 += W0;
        k1 += W1;
    }
    Philox4 o;
    o.v[0] = c0; o.v[1] = c1; o.v[2] = c2; o.v[3] = c3;
    return o;
}

template <class R> struct Rng;
// f32: two blocks per bounce, four 24-bit draws each
template <> struct Rng<float> {
    uint32_t pixel, s_lo, s_hi, seed_lo, seed_hi;
    PTB_DEV Rng(uint32_t p, uint64_t sample, uint64_t seed)
        : pixel(p), s_lo((uint32_t)sample), s_hi((uint32_t)(sample >> 32)), seed_lo((uint32_t)seed), seed_hi((uint32_t)(seed >> 32)) {}
    PTB_DEV static float u(uint32_t w) { return (float)(w >> 8) * (1.0f / 16777216.0f); }
    // slots 4*half .. 4*half+3 of `bounce`
    PTB_DEV void block(uint32_t bounce, uint32_t half, float out[4]) const {
        Philox4 p = philox4x32_10(bounce * 2u + half, s_hi, seed_lo, seed_hi, pixel, s_lo);
#pragma unroll
        for (int i = 0; i < 4; ++i) out[i] = u(p.v[i]);
    }
    // all 8 slots of a bounce
    PTB_DEV void draws(uint32_t bounce, float out[8]) const {
        block(bounce, 0, out);
        block(bounce, 1, out + 4);
    }
};
// f64: four blocks per bounce, two 53-bit draws each
template <> struct Rng<double> {
    uint32_t pixel, s_lo, s_hi, seed_lo, seed_hi;
    PTB_DEV Rng(uint32_t p, uint64_t sample, uint64_t seed)
        : pixel(p), s_lo((uint32_t)sample), s_hi((uint32_t)(sample >> 32)), seed_lo((uint32_t)seed), seed_hi((uint32_t)(seed >> 32)) {}
    PTB_DEV static double u(uint32_t w0, uint32_t w1) {
        uint64_t bits = (((uint64_t)w0 << 32) | w1) >> 11;
        return (double)bits * (1.0 / 9007199254740992.0);
    }
    PTB_DEV void block(uint32_t bounce, uint32_t half, double out[4]) const {
#pragma unroll
        for (uint32_t q = 0; q < 2; ++q) {
            Philox4 p = philox4x32_10(bounce * 4u + half * 2u + q, s_hi, seed_lo, seed_hi, pixel, s_lo);
            out[q * 2 + 0] = u(p.v[0], p.v[1]);
            out[q * 2 + 1] = u(p.v[2], p.v[3]);
        }
    }
    PTB_DEV void draws(uint32_t bounce, double out[8]) const {
        block(bounce, 0, out);
        block(bounce, 1, out + 4);
    }
};

// The draws of a shaded bounce for the f32 staged integrators: slots 4..7 always; slots 2,3 (light pick, coin) only when
// they can matter — with one light the pick is index 0 whatever the draw ((u * 1) as usize, u < 1), and without a
// transmission lobe the coin decides nothing (ff = 1 - (1-F) * spec_trans * (1-metallic) = 1 > coin, tracer.rs:532-534).
PTB_DEV void shade_draws(const Rng<double>& rng, uint32_t bounce, bool, double u[8]) { rng.draws(bounce, u); }
PTB_DEV void shade_draws(const Rng<float>& rng, uint32_t bounce, bool need_first_block, float u[8]) {
    u[0] = u[1] = u[2] = u[3] = 0.0f;
    if (need_first_block) rng.block(bounce, 0, u);
    rng.block(bounce, 1, u + 4);
}

// ------------------------------------------------------------------------------------------------
#define PTB_EMB_SPHERES 8
#define PTB_EMB_PLANES 4
#define PTB_EMB_LIGHTS 4
// Device scene (what ptb_set_scene_* builds from the POD export)
template <class R> struct DMaterial {
    // resolved values: Material::new() defaults (material.rs:82-114) patched by this material
    R rgb[3], emission[3];
    R anisotropic, metallic, roughness, subsurface, specular_tint, sheen, sheen_tint, clearcoat, clearcoat_gloss, spec_trans, ior;
    uint32_t set_mask, albedo_kind;
    R checker_a, checker_b, checker_scale, checker_offset;
    // material.rs:15-22 (PTB_MEDIUM_*); med_g = anisotropy clamped to [-0.9, 0.9] (material.rs:126)
    uint32_t med_type, med_pad;
    R med_density, med_color[3], med_g;
};
constexpr uint32_t PTB_MEDIUM_MAX_INDEX = 127u;      // PathState::medium is kept in 7 bits of the wavefront slot's flag word
template <class R> struct alignas(4 * sizeof(R)) DSphere { R cx, cy, cz, r; };   // one 16-byte (f32) / 32-byte (f64) vector load
template <class R> struct DPlane { R px, py, pz, nx, ny, nz; };
template <class R> struct DLight { R px, py, pz, radius, ex, ey, ez, area; uint32_t type, pad; R ux, uy, uz, vx, vy, vz; };

// One instruction of the signed-distance program, in postfix order (ptb_sdf_node_*, include/ptb200.h): primitives push
// (distance, material), combinators pop two entries and push one.
template <class R> struct DSdfNode { uint32_t op, material; R p[3]; R a[4]; };

// 32-byte BVH node over spheres (f32 bounds even for the f64 build; bounds are conservative).
struct alignas(32) BvhNode {
    float lo[3]; uint32_t left_or_first;   // inner: left child index (right = left + 1); leaf: first prim
    float hi[3]; uint32_t count;           // 0 = inner, >0 = leaf prim count
};

template <class R> struct DScene {
    uint32_t n_spheres, n_planes, n_materials, n_lights;
    const void* blob;                   // packed scene arrays in HBM (see stage_scene)
    uint32_t blob_bytes, small_bytes, off_spheres, off_planes, off_materials, off_lights, off_sphere_material, off_plane_material;
    uint32_t off_rm_keys, off_rm_table, rm_entries;   // resolved-material table (RMat below); rm_entries == 0: not built
    const DSphere<R>* spheres;          // the same arrays, addressed directly (BVH leaves, parity kernels)
    const uint32_t* sphere_material;
    const DPlane<R>* planes;
    const uint32_t* plane_material;
    const DMaterial<R>* materials;
    const DLight<R>* lights;
    const BvhNode* bvh;                 // NULL when the scene is small
    const uint32_t* bvh_prim;           // sphere indices in leaf order
    const DSphere<R>* bvh_spheres;      // the spheres themselves in leaf order (leaf = one contiguous read)
    float light_lo[3], light_hi[3];     // bounding box of all spherical lights (cull for sample_lights)
    const BvhNode* light_bvh;           // BVH over the spherical lights (NULL below 16 lights: linear scan as in scene.rs:68)
    const DSphere<R>* light_bvh_spheres;
    const uint32_t* light_bvh_prim;
    uint32_t use_bvh;
    uint32_t has_emissive;              // 1 if any material has non-zero emission
    uint32_t has_media;                 // 1 if any material carries a medium (PTB_MEDIUM_*): the path loop tracks inside / outside
    uint32_t has_fx;                    // has_media, or rectangular / distant lights that PTB_SCENE_EXTENDED_LIGHTS switches on
    // small scenes: the primitives themselves in the kernel parameter (constant bank, uniform loads) — see PTB_EMB_SCENE
    uint32_t emb;                       // 1 if the three arrays below hold the whole scene
    DSphere<R> emb_spheres[PTB_EMB_SPHERES];
    DPlane<R> emb_planes[PTB_EMB_PLANES];
    DLight<R> emb_lights[PTB_EMB_LIGHTS];
    uint32_t patch_materials;           // 1 if any set_mask != PTB_MAT_ALL (order-dependent patching)
    uint32_t depth, flags;
    R eps;
    // camera/pinhole.rs:38-61 with the loop-invariant part hoisted (host, same op order, in R)
    R cam_origin[3], cam_base[3] /* lower_left - origin */, cam_horizontal[3], cam_vertical[3];
    // background
    uint32_t bg_kind;
    R bg_a[3], bg_b[3], bg_scale, bg_gamma;
    R n_lights_f;                       // number_of_lights() as F (tracer.rs:138,214)
    // signed-distance program (ptb_set_sdf_*): one more "primitive", tested after the planes.  The nodes live in the kernel
    // parameter (constant bank): every lane reads the same node at the same time, which is what that path is fast at.
    uint32_t n_rect_lights;             // rectangular lights in the scene (tested by rays only with PTB_SCENE_EXTENDED_LIGHTS)
    uint32_t n_sdf, sdf_max_steps;
    R sdf_hit_eps, sdf_max_dist, sdf_normal_h;
    DSdfNode<R> sdf[PTB_SDF_MAX_NODES];
};

// Scene arrays as the kernels read them.  The host packs [planes | lights | plane_material | spheres |
// sphere_material | materials] into ONE blob (DScene::blob); a CTA stages the whole blob in shared
// memory with a single loop when it fits, otherwise just the small head section (planes, lights),
// and reads the rest from HBM through L1/L2.
template <class R> struct SceneView {
    const DSphere<R>* spheres;
    const uint32_t* sphere_material;
    const DPlane<R>* planes;
    const uint32_t* plane_material;
    const DMaterial<R>* materials;
    const DLight<R>* lights;
};

constexpr uint32_t PTB_SMEM_SCENE_BYTES = 12 * 1024;    // capacity of the shared-memory scene copy

template <class R> struct alignas(16) SceneSmem { uint32_t words[PTB_SMEM_SCENE_BYTES / 4]; };

// Cooperative copy of the (small) scene blob into shared memory; call from all threads of the CTA.
template <class R> __device__ inline SceneView<R> stage_scene(const DScene<R>& s, void* smem_words, uint32_t capacity_bytes) {
    // whole blob if it fits, else at least its small head section (planes, lights, plane materials)
    const uint32_t n = s.blob_bytes <= capacity_bytes ? s.blob_bytes : (s.small_bytes <= capacity_bytes ? s.small_bytes : 0u);
    const uint32_t* src = (const uint32_t*)s.blob;
    uint32_t* dst = (uint32_t*)smem_words;
    for (uint32_t i = threadIdx.x; i < n / 4u; i += blockDim.x) dst[i] = src[i];
    __syncthreads();
    const char* g = (const char*)s.blob;
    const char* m = (const char*)smem_words;
    auto at = [&](uint32_t off) { return off < n ? m + off : g + off; };
    SceneView<R> v;
    v.planes = (const DPlane<R>*)at(s.off_planes);
    v.lights = (const DLight<R>*)at(s.off_lights);
    v.plane_material = (const uint32_t*)at(s.off_plane_material);
    v.spheres = (const DSphere<R>*)at(s.off_spheres);
    v.sphere_material = (const uint32_t*)at(s.off_sphere_material);
    v.materials = (const DMaterial<R>*)at(s.off_materials);
    return v;
}

// ------------------------------------------------------------------------------------------------
// Material in registers (material.rs:48-78 + derived fields of finalize, material.rs:117-131)
template <class R> struct Mat {
    V3<R> rgb, emission;
    R anisotropic, metallic, roughness, subsurface, specular_tint, sheen, sheen_tint, clearcoat, clearcoat_gloss, spec_trans, ior;
    R clearcoat_roughness, ax, ay;
};

template <class R> PTB_DEV void mat_defaults(Mat<R>& m) {    // material.rs:82-114
    m.rgb = V3<R>(R(1.5), R(1.5), R(1.5));
    m.emission = V3<R>(0, 0, 0);
    m.anisotropic = 0; m.metallic = 0; m.roughness = R(0.5); m.subsurface = 0; m.specular_tint = 0;
    m.sheen = 0; m.sheen_tint = 0; m.clearcoat = 0; m.clearcoat_gloss = 0; m.spec_trans = 0; m.ior = R(1.45);
    m.clearcoat_roughness = 0; m.ax = 0; m.ay = 0;
}

// analytical.rs:107-111
// `v % 2.0` for an integer-valued v: v - 2*trunc(v/2) is exact in floating point and keeps the
// dividend's sign like fmod (NaN/inf -> NaN), at a fraction of fmodf's code size.
PTB_DEV float fmod2_int(float v) { return v - 2.0f * truncf(v * 0.5f); }
PTB_DEV double fmod2_int(double v) { return v - 2.0 * trunc(v * 0.5); }
template <class R> PTB_DEV R checker(R x, R y, R a, R b) {
    R x1 = fmod2_int(m_floor(x));
    R y1 = fmod2_int(m_floor(y));
    return (fmod2_int(x1 + y1) < R(1)) ? a : b;
}

template <class R> PTB_DEV V3<R> material_rgb(const DMaterial<R>& dm, V3<R> rd) {
    if (dm.albedo_kind == PTB_ALBEDO_CHECKER_DIR_RATIO) {   // analytical.rs:113-115 (quirk A.7)
        R c = checker(div_rn(rd.x, rd.y) * dm.checker_scale + dm.checker_offset, div_rn(rd.z, rd.y) * dm.checker_scale + dm.checker_offset,
                      dm.checker_a, dm.checker_b);
        return V3<R>(c, c, c);
    }
    return V3<R>(dm.rgb[0], dm.rgb[1], dm.rgb[2]);
}

// first accepted primitive: defaults patched by dm == dm's resolved values
template <class R> PTB_DEV void mat_load(Mat<R>& m, const DMaterial<R>& dm, V3<R> rd) {
    m.rgb = (dm.set_mask & PTB_MAT_RGB) ? material_rgb(dm, rd) : V3<R>(dm.rgb[0], dm.rgb[1], dm.rgb[2]);
    m.emission = V3<R>(dm.emission[0], dm.emission[1], dm.emission[2]);
    m.anisotropic = dm.anisotropic; m.metallic = dm.metallic; m.roughness = dm.roughness; m.subsurface = dm.subsurface;
    m.specular_tint = dm.specular_tint; m.sheen = dm.sheen; m.sheen_tint = dm.sheen_tint; m.clearcoat = dm.clearcoat;
    m.clearcoat_gloss = dm.clearcoat_gloss; m.spec_trans = dm.spec_trans; m.ior = dm.ior;
}
// later accepted primitives assign only their masked fields (see PTB_MAT_* in ptb200.h)
template <class R> PTB_DEV void mat_patch(Mat<R>& m, const DMaterial<R>& dm, V3<R> rd) {
    const uint32_t k = dm.set_mask;
    if (k & PTB_MAT_RGB) m.rgb = material_rgb(dm, rd);
    if (k & PTB_MAT_EMISSION) m.emission = V3<R>(dm.emission[0], dm.emission[1], dm.emission[2]);
    if (k & PTB_MAT_ANISOTROPIC) m.anisotropic = dm.anisotropic;
    if (k & PTB_MAT_METALLIC) m.metallic = dm.metallic;
    if (k & PTB_MAT_ROUGHNESS) m.roughness = dm.roughness;
    if (k & PTB_MAT_SUBSURFACE) m.subsurface = dm.subsurface;
    if (k & PTB_MAT_SPECULAR_TINT) m.specular_tint = dm.specular_tint;
    if (k & PTB_MAT_SHEEN) m.sheen = dm.sheen;
    if (k & PTB_MAT_SHEEN_TINT) m.sheen_tint = dm.sheen_tint;
    if (k & PTB_MAT_CLEARCOAT) m.clearcoat = dm.clearcoat;
    if (k & PTB_MAT_CLEARCOAT_GLOSS) m.clearcoat_gloss = dm.clearcoat_gloss;
    if (k & PTB_MAT_SPEC_TRANS) m.spec_trans = dm.spec_trans;
    if (k & PTB_MAT_IOR) m.ior = dm.ior;
}
// material.rs:117-131
template <class R> PTB_DEV void mat_finalize(Mat<R>& m) {
    m.roughness = m_max(m.roughness, R(0.01));
    m.clearcoat_roughness = mix1(R(0.1), R(0.001), m.clearcoat_gloss);
    R aspect = m_sqrt(R(1) - m.anisotropic * R(0.9));
    m.ax = m_max(m_div(m.roughness, aspect), R(0.001));
    m.ay = m_max(m.roughness * aspect, R(0.001));
}

// ------------------------------------------------------------------------------------------------
// intersections
// analytical.rs:166-190 / scene.rs:39-63.  Returns t >= 0 or -1 for None.  The cancelling
// expression d2 = l.l - tca^2 is evaluated WITHOUT FMA contraction (measured: contracting it costs
// parity — hit distances drift to 1e-4 on grazing rays and six parity tests fail — for a 1 % speed-up).
PTB_DEV float mul_rn(float a, float b) { return __fmul_rn(a, b); }
PTB_DEV double mul_rn(double a, double b) { return __dmul_rn(a, b); }
PTB_DEV float add_rn(float a, float b) { return __fadd_rn(a, b); }
PTB_DEV double add_rn(double a, double b) { return __dadd_rn(a, b); }
PTB_DEV float sub_rn(float a, float b) { return __fsub_rn(a, b); }
PTB_DEV double sub_rn(double a, double b) { return __dsub_rn(a, b); }
template <class R> PTB_DEV R dot_rn(V3<R> a, V3<R> b) {
    return add_rn(add_rn(mul_rn(a.x, b.x), mul_rn(a.y, b.y)), mul_rn(a.z, b.z));
}
template <class R> PTB_DEV R isect_sphere(V3<R> o, V3<R> d, V3<R> c, R radius) {
    V3<R> l = c - o;
    R tca = dot_rn(l, d);
    R d2 = sub_rn(dot_rn(l, l), mul_rn(tca, tca));
    R radius2 = radius * radius;
    if (d2 > radius2) return R(-1);
    R thc = m_sqrt(radius2 - d2);
    R t0 = tca - thc, t1 = tca + thc;
    if (t0 > t1) { R tmp = t0; t0 = t1; t1 = tmp; }
    if (t0 < R(0)) {
        t0 = t1;
        if (t0 < R(0)) return R(-1);
    }
    return t0;
}
// analytical.rs:193-204 generalised to (point, normal)
template <class R> PTB_DEV R isect_plane(V3<R> o, V3<R> d, V3<R> p, V3<R> n) {
    R denom = dot_rn(n, d);
    if (m_abs(denom) > R(0.0001)) {
        R t = m_div(dot_rn(p - o, n), denom);
        if (t >= R(0)) return t;
    }
    return R(-1);
}

// Ray against a rectangular light (extension, see PTB_LIGHT_RECTANGULAR): t >= 0 and the cosine at the quad, or -1.  The quad
// emits from the side n = normalize(cross(u, v)) points to and is invisible from behind.
template <class R> struct DLight;
template <class R> PTB_DEV R isect_rect_light(const DLight<R>& L, V3<R> o, V3<R> d, R& cos_out) {
    const V3<R> u(L.ux, L.uy, L.uz), v(L.vx, L.vy, L.vz), p(L.px, L.py, L.pz);
    const V3<R> n = normalize(cross(u, v));
    const R dn = dot(n, d);
    if (!(dn < R(0))) return R(-1);
    const R t = m_div(dot(n, p - o), dn);
    if (!(t > R(0))) return R(-1);
    const V3<R> w = (o + t * d) - p;
    const R a1 = m_div(dot(u, w), dot(u, u)), a2 = m_div(dot(v, w), dot(v, v));
    if (a1 < R(0) || a1 > R(1) || a2 < R(0) || a2 > R(1)) return R(-1);
    cos_out = -dn;
    return t;
}

// ------------------------------------------------------------------------------------------------
// Signed-distance scenes (the reference's stated motivation, Readme.md:18,76-84, and its open "SDF based
// example scene" todo).  The reference has no SDF code, so there is nothing to restate: the semantics below are this library's
// extension of the Scene contract (scene.rs:12-16: closest_hit fills hit_dist / normal / material, any_hit answers shadow
// rays) and oracle/pt_oracle.hpp evaluates the identical program with the identical operation order.  All of it is
// add / mul / min / max / sqrt — no transcendental — so the strict build reproduces the oracle bit for bit.
template <class R> struct SdfSample { R d; uint32_t material; };
template <class R> PTB_DEV R sdf_len2(R x, R y) { return m_sqrt(x * x + y * y); }
template <class R> PTB_DEV R sdf_len3(R x, R y, R z) { return m_sqrt(x * x + y * y + z * z); }
template <class R> PTB_DEV R m_min(R a, R b) { return a < b ? a : b; }
template <class R> PTB_DEV SdfSample<R> sdf_eval(const DScene<R>& s, V3<R> q) {
    SdfSample<R> st[PTB_SDF_MAX_STACK];
    int sp = 0;
#pragma unroll 1
    for (uint32_t i = 0; i < s.n_sdf; ++i) {
        const DSdfNode<R>& n = s.sdf[i];
        const R x = q.x - n.p[0], y = q.y - n.p[1], z = q.z - n.p[2];
        SdfSample<R> v;
        v.material = n.material;
        switch (n.op) {
            case PTB_SDF_SPHERE: v.d = sdf_len3(x, y, z) - n.a[0]; break;
            case PTB_SDF_BOX: {                             // half extents a[0..2], corner radius a[3]
                const R dx = m_abs(x) - n.a[0], dy = m_abs(y) - n.a[1], dz = m_abs(z) - n.a[2];
                const R outside = sdf_len3(m_max(dx, R(0)), m_max(dy, R(0)), m_max(dz, R(0)));
                v.d = outside + m_min(m_max(dx, m_max(dy, dz)), R(0)) - n.a[3];
                break;
            }
            case PTB_SDF_TORUS: v.d = sdf_len2(sdf_len2(x, z) - n.a[0], y) - n.a[1]; break;       // ring of radius a[0] in the xz plane, tube a[1]
            case PTB_SDF_PLANE: v.d = q.x * n.a[0] + q.y * n.a[1] + q.z * n.a[2] + n.a[3]; break;   // unit normal a[0..2], offset a[3]
            default: {                                      // combinators: b = top of stack, a = the entry below it
                const SdfSample<R> b = st[--sp], a = st[--sp];
                if (n.op == PTB_SDF_UNION) v = b.d < a.d ? b : a;
                else if (n.op == PTB_SDF_INTERSECT) v = b.d > a.d ? b : a;
                else if (n.op == PTB_SDF_SUBTRACT) { v.d = m_max(a.d, -b.d); v.material = a.material; }
                else {                                      // PTB_SDF_SMOOTH_UNION, blend radius a[0] (polynomial smooth minimum)
                    const R h = m_clamp(R(0.5) + R(0.5) * div_rn(b.d - a.d, n.a[0]), R(0), R(1));
                    v.d = ((R(1) - h) * b.d + a.d * h) - n.a[0] * h * (R(1) - h);
                    v.material = h >= R(0.5) ? a.material : b.material;
                }
            }
        }
        st[sp++] = v;
    }
    return st[0];
}
// Sphere tracing from t = 0: steps of |distance| (a ray that starts inside a transmissive body walks out to the surface), hit
// when |distance| < sdf_hit_eps, given up beyond min(limit, sdf_max_dist) or after sdf_max_steps.  Returns t >= 0 or -1.
template <class R> PTB_DEV R sdf_trace(const DScene<R>& s, V3<R> o, V3<R> d, R limit, uint32_t& material) {
    const R t_end = m_min(limit, s.sdf_max_dist);
    R t = 0;
#pragma unroll 1
    for (uint32_t i = 0; i < s.sdf_max_steps; ++i) {
        const SdfSample<R> v = sdf_eval(s, V3<R>(o.x + t * d.x, o.y + t * d.y, o.z + t * d.z));
        const R a = m_abs(v.d);
        if (a < s.sdf_hit_eps) { material = v.material; return t; }
        t = t + a;
        if (!(t < t_end)) break;
    }
    return R(-1);
}
// gradient by four samples on a tetrahedron around the hit point
template <class R> PTB_DEV V3<R> sdf_normal(const DScene<R>& s, V3<R> q) {
    const R h = s.sdf_normal_h;
    const R d0 = sdf_eval(s, V3<R>(q.x + h, q.y - h, q.z - h)).d, d1 = sdf_eval(s, V3<R>(q.x - h, q.y - h, q.z + h)).d;
    const R d2 = sdf_eval(s, V3<R>(q.x - h, q.y + h, q.z - h)).d, d3 = sdf_eval(s, V3<R>(q.x + h, q.y + h, q.z + h)).d;
    return normalize(V3<R>(((d0 - d1) - d2) + d3, ((d2 + d3) - d0) - d1, ((d1 + d3) - d0) - d2));
}
template <class R> PTB_DEV uint32_t sdf_prim(const DScene<R>& s) { return s.n_spheres + s.n_planes; }   // its index in test order

// camera/pinhole.rs:38-61 (invariants hoisted into DScene by the host); p = film position,
// off = jitter, inv_w/inv_h = pixel_size
template <class R> PTB_DEV void gen_ray(const DScene<R>& s, R px, R py, R offx, R offy, R inv_w, R inv_h, V3<R>& o, V3<R>& d) {
    R a = inv_w * offx + px;
    R b = inv_h * offy + py;
    V3<R> rd(s.cam_base[0], s.cam_base[1], s.cam_base[2]);
    rd = rd + V3<R>(s.cam_horizontal[0] * a, s.cam_horizontal[1] * a, s.cam_horizontal[2] * a);
    rd = rd + V3<R>(s.cam_vertical[0] * b, s.cam_vertical[1] * b, s.cam_vertical[2] * b);
    o = V3<R>(s.cam_origin[0], s.cam_origin[1], s.cam_origin[2]);
    d = normalize(rd);
}

// film coordinates of tracer.rs:34-46 for pixel x of memory row `row` (0 = top)
template <class R> PTB_DEV void film_coords(uint32_t x, uint32_t row, uint32_t W, uint32_t H, R& px, R& py) {
    // j counts rows from the LAST memory row (par_rchunks_exact_mut): j = H-1-row; y = H - j
    R hf = (R)H;
    R y = hf - (R)(H - 1u - row);
    R yy = div_rn(y, hf);
    px = div_rn((R)x, (R)W);
    py = R(1) - yy;
}

// the same quotients from a precomputed reciprocal and one FMA correction step (q = a*y; r = a - q*b exactly; q + r*y):
// correctly rounded for the operands the host checked (RenderArgs::film_fast), three FMA-pipe instructions per quotient
PTB_HD float div_by_fma(float a, float b, float rcp_b) {
    float q = a * rcp_b;
    float r = fmaf(-q, b, a);
    return fmaf(r, rcp_b, q);
}
PTB_DEV void film_coords_fma(uint32_t x, uint32_t row, uint32_t W, uint32_t H, float rcp_w, float rcp_h, float& px, float& py) {
    px = div_by_fma((float)x, (float)W, rcp_w);
    py = 1.0f - div_by_fma((float)(row + 1u), (float)H, rcp_h);      // y = H - (H - 1 - row) = row + 1 exactly (H < 2^24)
}

// Scene::background, analytical.rs:28-32
template <class R> PTB_DEV V3<R> background(const DScene<R>& s, V3<R> d) {
    V3<R> a(s.bg_a[0], s.bg_a[1], s.bg_a[2]);
    if (s.bg_kind == PTB_BG_GRADIENT_Y) {
        R t = R(0.5) * (d.y + R(1));
        V3<R> b(s.bg_b[0], s.bg_b[1], s.bg_b[2]);
        V3<R> c = (R(1) - t) * a + t * b;
        return V3<R>(m_pow(c.x, s.bg_gamma) * s.bg_scale, m_pow(c.y, s.bg_gamma) * s.bg_scale, m_pow(c.z, s.bg_gamma) * s.bg_scale);
    }
    return a;
}

// Ray / box slab test in f32, inflated by a relative margin so that rounding can never cull a true
// hit.  Returns the entry distance, or +inf when the box is missed or starts beyond `limit`.
struct RayF { float ox, oy, oz, idx, idy, idz; };
PTB_DEV float box_entry(const float* lo, const float* hi, const RayF& r, float limit) {
    float tx0 = (lo[0] - r.ox) * r.idx, tx1 = (hi[0] - r.ox) * r.idx;
    float ty0 = (lo[1] - r.oy) * r.idy, ty1 = (hi[1] - r.oy) * r.idy;
    float tz0 = (lo[2] - r.oz) * r.idz, tz1 = (hi[2] - r.oz) * r.idz;
    float tn = fmaxf(fmaxf(fminf(tx0, tx1), fminf(ty0, ty1)), fmaxf(fminf(tz0, tz1), 0.0f));
    float tf = fminf(fminf(fmaxf(tx0, tx1), fmaxf(ty0, ty1)), fmaxf(tz0, tz1));
    tn *= 0.9999f;
    return (tn <= tf * 1.0001f && tn <= limit) ? tn : 3.0e38f;
}
PTB_DEV BvhNode load_node(const BvhNode* p) {
    const float4* q = reinterpret_cast<const float4*>(p);
    float4 a = __ldg(q), b = __ldg(q + 1);
    BvhNode n;
    n.lo[0] = a.x; n.lo[1] = a.y; n.lo[2] = a.z; n.left_or_first = __float_as_uint(a.w);
    n.hi[0] = b.x; n.hi[1] = b.y; n.hi[2] = b.z; n.count = __float_as_uint(b.w);
    return n;
}

// Sphere BVH traversal.  Children of an inner node are adjacent, so both boxes are fetched with one
// 64-byte read and the nearer child is descended first (the farther one is stacked with its entry
// distance and skipped on pop if the closest hit found meanwhile is nearer).  ANY = shadow-ray mode:
// first hit within max_dist returns.  Closest mode keeps the reference's tie rule (ascending index,
// strict `d < dist` => the lowest sphere index wins equal distances).  Returns the sphere index or -1.
// traversal stack entries; the host builder bounds the tree depth by this (BvhBuilder::build_node, ptb_api.cu)
constexpr int BVH_STACK = 40;
template <class R, bool ANY>
PTB_DEV int bvh_traverse(const BvhNode* __restrict__ nodes, const DSphere<R>* __restrict__ leaf_spheres, const uint32_t* __restrict__ leaf_prim,
                         V3<R> o, V3<R> d, R& best_t, uint32_t* stats = nullptr) {
    int best = -1;
    RayF r;
    r.ox = (float)o.x; r.oy = (float)o.y; r.oz = (float)o.z;
    r.idx = m_rcp((float)d.x); r.idy = m_rcp((float)d.y); r.idz = m_rcp((float)d.z);
    constexpr int STACK = BVH_STACK;
    // an entry = (node reference, entry distance): the reference carries the child's link and leaf count (count <= 7, builder: <= 4),
    // so a pop is ONE 8-byte local load and goes straight on to the children — the first version popped with two dependent local
    // loads and then re-read the popped node from global memory just to learn its link and count
    uint2 stack[STACK];
    int sp = 0;
    BvhNode root = load_node(nodes);
    if (box_entry(root.lo, root.hi, r, (float)best_t) >= 3.0e38f) return -1;
    uint32_t cur_lf = root.left_or_first, cur_cnt = root.count;
    while (true) {
        if (cur_cnt) {
            if (stats) stats[1] += cur_cnt;
            for (uint32_t i = 0; i < cur_cnt; ++i) {
                const DSphere<R> sph = leaf_spheres[cur_lf + i];
                R t = isect_sphere(o, d, V3<R>(sph.cx, sph.cy, sph.cz), sph.r);
                if (t >= R(0)) {
                    if (ANY) { if (t < best_t) return (int)leaf_prim[cur_lf + i]; }
                    else {
                        const int si = (int)leaf_prim[cur_lf + i];
                        if (t < best_t || (t == best_t && si < best)) { best_t = t; best = si; }
                    }
                }
            }
        } else {
            if (stats) stats[0]++;
            const BvhNode a = load_node(nodes + cur_lf), b = load_node(nodes + cur_lf + 1);
            const float ta = box_entry(a.lo, a.hi, r, (float)best_t), tb = box_entry(b.lo, b.hi, r, (float)best_t);
            const bool ha = ta < 3.0e38f, hb = tb < 3.0e38f;
            if (ha || hb) {
                const bool a_near = ha && (!hb || ta <= tb);
                const uint32_t ra = (a.left_or_first << 3) | a.count, rb = (b.left_or_first << 3) | b.count;
                const uint32_t near_ref = a_near ? ra : rb;
                if (ha && hb && sp < STACK) {
                    // (far child as ra ^ rb ^ near: see the note on nvcc 12.9 and mirrored selects in ptb_stream.cuh)
                    stack[sp] = make_uint2(ra ^ rb ^ near_ref, __float_as_uint(fmaxf(ta, tb)));
                    ++sp;
                }
                cur_lf = near_ref >> 3; cur_cnt = near_ref & 7u;
                continue;
            }
        }
        bool found = false;
        while (sp > 0) {
            --sp;
            const uint2 e = stack[sp];
            if (__uint_as_float(e.y) <= (float)best_t) {
                cur_lf = e.x >> 3; cur_cnt = e.x & 7u;
                found = true;
                break;
            }
        }
        if (!found) break;
    }
    return best;
}
template <class R> PTB_DEV int bvh_closest(const DScene<R>& s, V3<R> o, V3<R> d, R& best_t, uint32_t* stats = nullptr) {
    return bvh_traverse<R, false>(s.bvh, s.bvh_spheres, s.bvh_prim, o, d, best_t, stats);
}
template <class R> PTB_DEV bool bvh_any(const DScene<R>& s, V3<R> o, V3<R> d, R max_dist, bool ignore_max, uint32_t* stats = nullptr) {
    R limit = ignore_max ? Const<R>::MAXV : max_dist;
    return bvh_traverse<R, true>(s.bvh, s.bvh_spheres, s.bvh_prim, o, d, limit, stats) >= 0;
}

// Scene::closest_hit for the exported scene, geometry part (analytical.rs:36-127 + scene.rs:36-86):
// which primitive / light the ray hits.  `hit_dist_in` is State::hit_dist carried across bounces
// (quirk A.1).  Material and normal are resolved separately (hit_material / hit_normal) so the
// wavefront integrator can do that in its shading stage.
template <class R> struct HitCore {
    bool hit, is_emitter, geom;
    R hit_dist;            // state.hit_dist after the call (stale value kept when nothing was hit)
    int prim;              // closest geometric primitive: < n_spheres sphere, else plane; -1 none
    uint64_t accepted;     // bit i: primitive i was "closest so far" when tested (patching scenes only)
    R light_pdf;           // light_sample.pdf (valid if is_emitter)
    V3<R> light_emission;  // light_sample.emission
};

// sphere part of closest_hit: the closest sphere (index, distance) — the part a dedicated traversal kernel can run
// EMB (the resolved-material kernel on scenes of a few primitives): spheres / planes / lights are read from the copies in the
// kernel parameter (DScene::emb_*) instead of through the scene view.  The view's pointers are generic (a scene that does not fit
// is read from global memory through the same code), and under that kernel's 72-register cap ptxas rematerialised them at every
// use — S2UR SR_CgaCtaId, UMOV, ULEA, LDC, IADD3: 3 % of its instructions (capture r02-d); the parameter copies are read with
// uniform constant-bank loads and need no pointer at all.
#define SV_SPHERE(i) (EMB ? s.emb_spheres[i] : sv.spheres[i])
#define SV_PLANE(i) (EMB ? s.emb_planes[i] : sv.planes[i])
#define SV_LIGHT(i) (EMB ? s.emb_lights[i] : sv.lights[i])
template <class R, bool BVH, bool EMB = false>
PTB_DEV void closest_spheres(const DScene<R>& s, const SceneView<R>& sv, V3<R> o, V3<R> d, int& best, R& dist, uint64_t& accepted, uint32_t* bvh_stats = nullptr) {
    dist = Const<R>::MAXV;
    best = -1;
    accepted = 0;
    if (BVH) {
        best = bvh_closest(s, o, d, dist, bvh_stats);
    } else {
#pragma unroll 1
        for (uint32_t i = 0; i < s.n_spheres; ++i) {
            DSphere<R> sp = SV_SPHERE(i);
            R t = isect_sphere(o, d, V3<R>(sp.cx, sp.cy, sp.cz), sp.r);
            if (t >= R(0) && (i == 0 || t < dist)) {   // analytical.rs:43 (unconditional), :74 (d < dist)
                dist = t; best = (int)i; accepted |= (1ull << (i & 63u));
            }
        }
    }
}

// the rest of closest_hit given the sphere result: planes, then Scene::sample_lights
// SDF = false compiles the signed-distance test out (instantiations whose scenes cannot carry a program); XL = false the
// rectangular-light test (the resolved-material kernel: ptb_set_scene_* builds no table for scenes with extended lights)
template <class R, bool BVH, bool SDF = true, bool XL = true, bool EMB = false>
PTB_DEV HitCore<R> closest_hit_finish(const DScene<R>& s, const SceneView<R>& sv, V3<R> o, V3<R> d, R hit_dist_in, int best, R dist,
                                      uint64_t accepted) {
    HitCore<R> h;
    h.hit = false; h.is_emitter = false; h.geom = false; h.hit_dist = hit_dist_in;
    h.light_pdf = 0; h.light_emission = V3<R>(0, 0, 0);
#pragma unroll 1
    for (uint32_t i = 0; i < s.n_planes; ++i) {
        DPlane<R> pl = SV_PLANE(i);
        R t = isect_plane(o, d, V3<R>(pl.px, pl.py, pl.pz), V3<R>(pl.nx, pl.ny, pl.nz));
        if (t >= R(0) && t < dist) {                   // analytical.rs:101-103
            dist = t; best = (int)(s.n_spheres + i); accepted |= (1ull << ((s.n_spheres + i) & 63u));
        }
    }
    if (SDF && s.n_sdf) {                               // the signed-distance body: one more primitive after the planes
        uint32_t mi = 0;
        const R t = sdf_trace(s, o, d, dist, mi);
        // (scenes with a signed-distance program assign whole materials — ptb_set_sdf_* checks it —, so the accepted set is
        //  not needed and the word carries the material the program returned at the hit)
        if (t >= R(0) && t < dist) { dist = t; best = (int)sdf_prim(s); accepted = (uint64_t)mi; }
    }
    h.prim = best; h.accepted = accepted;
    if (best >= 0) { h.hit = true; h.geom = true; h.hit_dist = dist; }
    // Scene::sample_lights, scene.rs:36-86 — starts from the possibly stale state.hit_dist
    R ldist = h.hit_dist;
    int lbest = -1;
    // A light can only win with 0 <= t < ldist: nothing to test when ldist <= 0 (a fresh path that missed all
    // geometry: hit_dist is still -1) or — for scenes with many lights — when the ray misses the lights' bounding box.
    uint32_t n_test = ldist > R(0) ? s.n_lights : 0u;
    if (n_test >= 4u) {
        RayF r;
        r.ox = (float)o.x; r.oy = (float)o.y; r.oz = (float)o.z;
        r.idx = m_rcp((float)d.x); r.idy = m_rcp((float)d.y); r.idz = m_rcp((float)d.z);
        if (box_entry(s.light_lo, s.light_hi, r, (float)ldist) >= 3.0e38f) n_test = 0u;
    }
    if (BVH && n_test && s.light_bvh) {      // (only the BVH kernels carry traversal code and its stack)
        // many lights: same traversal, same tie rule (ascending index, strict `d < dist`) as the sphere BVH
        lbest = bvh_traverse<R, false>(s.light_bvh, s.light_bvh_spheres, s.light_bvh_prim, o, d, ldist);
        n_test = 0u;
    }
#pragma unroll 1
    for (uint32_t i = 0; i < n_test; ++i) {
        DLight<R> L = SV_LIGHT(i);
        if (L.type != PTB_LIGHT_SPHERICAL) continue;
        R t = isect_sphere(o, d, V3<R>(L.px, L.py, L.pz), L.radius);
        if (t >= R(0) && t < ldist) { ldist = t; lbest = (int)i; }
    }
    R rect_cos = 0;
    if (XL && (s.flags & PTB_SCENE_EXTENDED_LIGHTS) && s.n_rect_lights && ldist > R(0)) {
        // quads (extension, PTB_LIGHT_RECTANGULAR in ptb200.h): hidden from behind, nearest wins like the spheres above; tested in
        // light order AFTER the spherical lights, so a quad and a sphere at the same distance resolve to the sphere
#pragma unroll 1
        for (uint32_t i = 0; i < s.n_lights; ++i) {
            DLight<R> L = sv.lights[i];
            if (L.type != PTB_LIGHT_RECTANGULAR) continue;
            R c;
            R t = isect_rect_light(L, o, d, c);
            if (t >= R(0) && t < ldist) { ldist = t; lbest = (int)i; rect_cos = c; }
        }
    }
    if (lbest >= 0) {
        DLight<R> L = SV_LIGHT(lbest);
        if (XL && L.type == PTB_LIGHT_RECTANGULAR) {
            h.light_pdf = m_div(ldist * ldist, L.area * rect_cos);
        } else {
            V3<R> hp = o + ldist * d;
            R cos_theta = dot(-d, normalize(hp - V3<R>(L.px, L.py, L.pz)));
            h.light_pdf = m_div(ldist * ldist, L.area * cos_theta * R(0.5));       // scene.rs:75
        }
        h.light_emission = V3<R>(L.ex, L.ey, L.ez);
        h.is_emitter = true;
        h.hit_dist = ldist;
        h.hit = true;
    }
    return h;
}

// normal of primitive `prim` at distance t along (o, d): analytical.rs:45-46 (sphere), :105 (plane)
template <class R, bool BVH, bool SDF = true, bool XL = true, bool EMB = false>
PTB_DEV HitCore<R> closest_hit_core(const DScene<R>& s, const SceneView<R>& sv, V3<R> o, V3<R> d, R hit_dist_in, uint32_t* bvh_stats = nullptr) {
    int best;
    R dist;
    uint64_t accepted;
    closest_spheres<R, BVH, EMB>(s, sv, o, d, best, dist, accepted, bvh_stats);
    return closest_hit_finish<R, BVH, SDF, XL, EMB>(s, sv, o, d, hit_dist_in, best, dist, accepted);
}

template <class R, bool BVH, bool SDF = true, bool EMB = false>
PTB_DEV V3<R> hit_normal(const DScene<R>& s, const SceneView<R>& sv, int prim, V3<R> o, V3<R> d, R t) {
    if ((uint32_t)prim < s.n_spheres) {
        DSphere<R> sp = BVH ? s.spheres[prim] : SV_SPHERE(prim);
        V3<R> hp = o + t * d;
        return normalize(hp - V3<R>(sp.cx, sp.cy, sp.cz));
    }
    if (SDF && (uint32_t)prim == sdf_prim(s)) return sdf_normal(s, o + t * d);
    DPlane<R> pl = SV_PLANE(prim - s.n_spheres);
    return V3<R>(pl.nx, pl.ny, pl.nz);
}

template <class R, bool BVH> PTB_DEV uint32_t prim_material(const DScene<R>& s, const SceneView<R>& sv, int prim) {
    if ((uint32_t)prim < s.n_spheres) return BVH ? s.sphere_material[prim] : sv.sphere_material[prim];
    return sv.plane_material[prim - s.n_spheres];
}

// material at the hit (un-finalized): the closest primitive's, or — scenes with partial set_masks —
// the reference's assignment order replayed over the accepted primitives (see PTB_MAT_* in ptb200.h)
template <class R, bool BVH>
PTB_DEV uint32_t hit_material(const DScene<R>& s, const SceneView<R>& sv, int prim, uint64_t accepted, V3<R> d, Mat<R>& mat) {
    const uint32_t mi = (s.n_sdf && (uint32_t)prim == sdf_prim(s)) ? (uint32_t)accepted : prim_material<R, BVH>(s, sv, prim);
    if (BVH || !s.patch_materials) {
        mat_load(mat, BVH ? s.materials[mi] : sv.materials[mi], d);
    } else {
        bool first = true;
#pragma unroll 1
        while (accepted) {
            int i = __ffsll((long long)accepted) - 1;
            accepted &= accepted - 1;
            uint32_t m = prim_material<R, BVH>(s, sv, i);
            if (first) { mat_load(mat, sv.materials[m], d); first = false; }
            else mat_patch(mat, sv.materials[m], d);
        }
    }
    return mi;
}

// Lobe class of the material at the hit — which Disney lobes can carry weight (tracer.rs:423-426):
// bit 0 diffuse ((1-metallic)(1-spec_trans) > 0), bit 1 clearcoat (clearcoat(1-metallic) > 0),
// bit 2 transmission (spec_trans(1-metallic) > 0).  The wavefront integrator sorts its shading
// queue by this key so that a warp evaluates the same lobes.
template <class R> PTB_DEV uint32_t lobe_class_of(R metallic, R spec_trans, R clearcoat) {
    const R nm = R(1) - metallic;
    return (nm * (R(1) - spec_trans) > R(0) ? 1u : 0u) | (clearcoat * nm > R(0) ? 2u : 0u) | (spec_trans * nm > R(0) ? 4u : 0u);
}
template <class R, bool BVH>
PTB_DEV uint32_t hit_lobe_class(const DScene<R>& s, const SceneView<R>& sv, int prim, uint64_t accepted) {
    if (BVH || !s.patch_materials) {
        const uint32_t mi = (s.n_sdf && (uint32_t)prim == sdf_prim(s)) ? (uint32_t)accepted : prim_material<R, BVH>(s, sv, prim);
        const DMaterial<R>& dm = BVH ? s.materials[mi] : sv.materials[mi];
        return lobe_class_of(dm.metallic, dm.spec_trans, dm.clearcoat);
    }
    R metallic = 0, spec_trans = 0, clearcoat = 0;
    bool first = true;
#pragma unroll 1
    while (accepted) {
        int i = __ffsll((long long)accepted) - 1;
        accepted &= accepted - 1;
        const DMaterial<R>& dm = sv.materials[prim_material<R, BVH>(s, sv, i)];
        const uint32_t k = first ? (uint32_t)PTB_MAT_ALL : dm.set_mask;
        first = false;
        if (k & PTB_MAT_METALLIC) metallic = dm.metallic;
        if (k & PTB_MAT_SPEC_TRANS) spec_trans = dm.spec_trans;
        if (k & PTB_MAT_CLEARCOAT) clearcoat = dm.clearcoat;
    }
    return lobe_class_of(metallic, spec_trans, clearcoat);
}

// Result of the complete Scene::closest_hit (used by the parity kernels)
template <class R> struct HitRec {
    bool hit, is_emitter;
    R hit_dist;
    V3<R> normal;          // geometry normal of the closest geometric hit (valid if geom)
    bool geom;
    uint32_t material;     // material index of the final geometric hit, 0xffffffff if none
    R light_pdf;
    V3<R> light_emission;
};
template <class R, bool BVH>
PTB_DEV HitRec<R> closest_hit(const DScene<R>& s, const SceneView<R>& sv, V3<R> o, V3<R> d, R hit_dist_in, Mat<R>& mat) {
    HitCore<R> c = closest_hit_core<R, BVH>(s, sv, o, d, hit_dist_in);
    HitRec<R> h;
    h.hit = c.hit; h.is_emitter = c.is_emitter; h.hit_dist = c.hit_dist; h.geom = c.geom;
    h.light_pdf = c.light_pdf; h.light_emission = c.light_emission;
    h.normal = V3<R>(0, 0, 0); h.material = 0xffffffffu;
    if (c.geom) {
        // geometry distance: equals hit_dist unless a nearer light replaced it
        R tg = c.hit_dist;
        if (c.is_emitter) {
            // recompute the primitive's own t (only the parity kernels get here)
            if ((uint32_t)c.prim < s.n_spheres) {
                DSphere<R> sp = BVH ? s.spheres[c.prim] : sv.spheres[c.prim];
                tg = isect_sphere(o, d, V3<R>(sp.cx, sp.cy, sp.cz), sp.r);
            } else if ((uint32_t)c.prim == sdf_prim(s)) {
                uint32_t mi;
                tg = sdf_trace(s, o, d, Const<R>::MAXV, mi);
            } else {
                DPlane<R> pl = sv.planes[c.prim - s.n_spheres];
                tg = isect_plane(o, d, V3<R>(pl.px, pl.py, pl.pz), V3<R>(pl.nx, pl.ny, pl.nz));
            }
        }
        h.normal = hit_normal<R, BVH>(s, sv, c.prim, o, d, tg);
        h.material = hit_material<R, BVH>(s, sv, c.prim, c.accepted, d, mat);
    }
    return h;
}

// Scene::any_hit, analytical.rs:130-145 (+ max_dist unless the scene flag says the impl ignores it), in two parts
// so that a dedicated traversal kernel can run the sphere part
template <class R, bool BVH, bool EMB = false> PTB_DEV bool any_hit_spheres(const DScene<R>& s, const SceneView<R>& sv, V3<R> o, V3<R> d, R max_dist, uint32_t* bvh_stats = nullptr) {
    const bool ignore = (s.flags & PTB_SCENE_ANYHIT_IGNORES_MAX_DIST) != 0;
    if constexpr (BVH) {
        return bvh_any(s, o, d, max_dist, ignore, bvh_stats);
    } else {
#pragma unroll 1
        for (uint32_t i = 0; i < s.n_spheres; ++i) {
            DSphere<R> sp = SV_SPHERE(i);
            R t = isect_sphere(o, d, V3<R>(sp.cx, sp.cy, sp.cz), sp.r);
            if (t >= R(0) && (ignore || t < max_dist)) return true;
        }
        return false;
    }
}
template <class R, bool SDF = true, bool EMB = false> PTB_DEV bool any_hit_planes(const DScene<R>& s, const SceneView<R>& sv, V3<R> o, V3<R> d, R max_dist) {
    const bool ignore = (s.flags & PTB_SCENE_ANYHIT_IGNORES_MAX_DIST) != 0;
#pragma unroll 1
    for (uint32_t i = 0; i < s.n_planes; ++i) {
        DPlane<R> pl = SV_PLANE(i);
        R t = isect_plane(o, d, V3<R>(pl.px, pl.py, pl.pz), V3<R>(pl.nx, pl.ny, pl.nz));
        if (t >= R(0) && (ignore || t < max_dist)) return true;
    }
    if (SDF && s.n_sdf) {                               // the signed-distance body (always honours max_dist)
        uint32_t mi;
        if (sdf_trace(s, o, d, max_dist, mi) >= R(0)) return true;
    }
    return false;
}
template <class R, bool BVH, bool SDF = true, bool EMB = false> PTB_DEV bool any_hit(const DScene<R>& s, const SceneView<R>& sv, V3<R> o, V3<R> d, R max_dist, uint32_t* bvh_stats = nullptr) {
    return any_hit_spheres<R, BVH, EMB>(s, sv, o, d, max_dist, bvh_stats) || any_hit_planes<R, SDF, EMB>(s, sv, o, d, max_dist);
}

// ------------------------------------------------------------------------------------------------
// Disney BSDF terms (tracer.rs:222-439)
template <class R> PTB_DEV R power_heuristic(R a, R b) { R t = a * a; return m_div(t, b * b + t); }   // tracer.rs:223-226
template <class R> PTB_DEV R luminance(V3<R> c) { return R(0.212671) * c.x + R(0.715160) * c.y + R(0.072169) * c.z; }
// f32::clamp(0, 1) keeps a NaN input (tracer.rs:289), the hardware's .SAT modifier turns it into 0: the saturating add stays
// (it is free), and a NaN operand is put back with one compare + select — the reference never filters NaN radiance
// (SURVEY.md §5), so a poisoned path must stay poisoned here too.
PTB_DEV float sat01_of_one_minus(float u) {
#ifdef PTB_SAT_DROPS_NAN
    return __saturatef(1.0f - u);
#else
    const float m = __saturatef(1.0f - u);
    return u != u ? u : m;
#endif
}
PTB_DEV double sat01_of_one_minus(double u) { return m_clamp(1.0 - u, 0.0, 1.0); }
template <class R> PTB_DEV R schlick_fresnel(R u) {                                                // tracer.rs:288-292
    R m = sat01_of_one_minus(u);
    R m2 = m * m;
    return m2 * m2 * m;
}
template <class R> PTB_DEV R dielectric_fresnel(R cos_theta_i, R eta) {                            // tracer.rs:308-322
    R sin_theta_tsq = eta * eta * (R(1) - cos_theta_i * cos_theta_i);
    if (sin_theta_tsq > R(1)) return R(1);
    R cos_theta_t = m_sqrt(m_max(R(1) - sin_theta_tsq, R(0)));
    R rs = m_div(eta * cos_theta_t - cos_theta_i, eta * cos_theta_t + cos_theta_i);
    R rp = m_div(eta * cos_theta_i - cos_theta_t, eta * cos_theta_i + cos_theta_t);
    return R(0.5) * (rs * rs + rp * rp);
}
template <class R> PTB_DEV R gtr1(R ndoth, R a) {                                                  // tracer.rs:233-240 (log2: A.3)
    if (a >= R(1)) return Const<R>::INV_PI;
    R a2 = a * a;
    R t = R(1) + (a2 - R(1)) * ndoth * ndoth;
    return m_div(a2 - R(1), Const<R>::PI * m_log2(a2) * t);
}
template <class R> PTB_DEV R smithg(R ndotv, R alphag) {                                           // tracer.rs:276-280
    R a = alphag * alphag;
    R b = ndotv * ndotv;
    return m_div(R(2) * ndotv, ndotv + m_sqrt(a + b - a * b));
}
template <class R> PTB_DEV R gtr2aniso(R ndoth, R hdotx, R hdoty, R ax, R ay) {                    // tracer.rs:294-299
    R a = m_div(hdotx, ax);
    R b = m_div(hdoty, ay);
    R c = a * a + b * b + ndoth * ndoth;
    return m_rcp(Const<R>::PI * ax * ay * c * c);
}
template <class R> PTB_DEV R smithganiso(R ndotv, R vdotx, R vdoty, R ax, R ay) {                  // tracer.rs:301-306
    R a = vdotx * ax;
    R b = vdoty * ay;
    R c = ndotv;
    return m_div(R(2) * ndotv, ndotv + m_sqrt(a * a + b * b + c * c));
}
template <class R> PTB_DEV R disney_fresnel(const Mat<R>& m, R eta, R ldoth, R vdoth) {            // tracer.rs:435-439
    R metallic_fresnel = schlick_fresnel(ldoth);
    R dielectric = dielectric_fresnel(m_abs(vdoth), eta);
    return mix1(dielectric, metallic_fresnel, m.metallic);
}
// tracer.rs:449-454 (identical copies at 184-189 and 559-564)
template <class R> PTB_DEV void onb(V3<R> n, V3<R>& t, V3<R>& b) {
    V3<R> up = m_abs(n.z) < R(0.999) ? V3<R>(0, 0, 1) : V3<R>(1, 0, 0);
    t = normalize(cross(up, n));
    b = cross(n, t);
}

// samplers, tracer.rs:242-274, 324-333.  Each takes (sin, cos) of its azimuth so that ONE sincos
// serves whichever lobe was picked (code size: the fused kernel is instruction-cache bound).
template <class R> PTB_DEV V3<R> cosine_sample_hemisphere(R r1, R sn, R cs) {
    R r = m_sqrt(r1);
    R x = r * cs, y = r * sn;
    return V3<R>(x, y, m_sqrt(m_max(R(0), R(1) - x * x - y * y)));
}
template <class R> PTB_DEV V3<R> sample_gtr1(R rgh, R r1, R sn, R cs) {          // r2 unused: quirk A.4 (phi = r1 * TWO_PI)
    R a = m_max(R(0.001), rgh);
    R a2 = a * a;
    R cos_theta = m_sqrt(m_div(R(1) - m_pow(a2, R(1) - r1), R(1) - a2));
    R sin_theta = m_clamp(m_sqrt(R(1) - (cos_theta * cos_theta)), R(0), R(1));
    return V3<R>(sin_theta * cs, sin_theta * sn, cos_theta);
}
template <class R> PTB_DEV V3<R> sample_ggxvndf(V3<R> v, R ax, R ay, R r1, R sn, R cs) {
    V3<R> vh = normalize(V3<R>(ax * v.x, ay * v.y, v.z));
    R lensq = vh.x * vh.x + vh.y * vh.y;
    V3<R> t_1;
    if (lensq > R(0)) { R il = m_rcp(m_sqrt(lensq)); t_1 = V3<R>(-vh.y * il, vh.x * il, R(0)); }
    else t_1 = V3<R>(1, 0, 0);
    V3<R> t_2 = cross(vh, t_1);
    R r = m_sqrt(r1);
    R t1 = r * cs;
    R t2 = r * sn;
    R s = R(0.5) * (R(1) + vh.z);
    t2 = (R(1) - s) * m_sqrt(R(1) - t1 * t1) + s * t2;
    V3<R> nh = t1 * t_1 + t2 * t_2 + m_sqrt(m_max(R(0), R(1) - t1 * t1 - t2 * t2)) * vh;
    return normalize(V3<R>(ax * nh.x, ay * nh.y, m_max(R(0), nh.z)));
}

// Per-bounce shading context: everything disney_eval (tracer.rs:555-600) and disney_sample
// (tracer.rs:441-493) both derive from (material, eta, n, v).  The reference recomputes it in each
// call; computing it once is value-identical.
template <class R> struct ShadeCtx {
    V3<R> t, b, n;       // onb(n)
    V3<R> v;             // to_local(v_world)
    V3<R> spec_col, sheen_col;
    R eta;
    R lum, wd0, wc0;     // luminance(rgb), diffuse and clearcoat weights before normalisation (tracer.rs:423, 426)
};
template <class R> PTB_DEV V3<R> to_local(const ShadeCtx<R>& c, V3<R> w) { return V3<R>(dot(w, c.t), dot(w, c.b), dot(w, c.n)); }
// The view vector's cosine v.z = dot(v, n) is evaluated like the reference does (three rounded products, two rounded sums —
// no FMA contraction): sampling the clearcoat lobe at an EXACTLY grazing view (v.z == 0) divides 0 by 0 (smithg(0) = 0 over
// 4 l.z v.z, tracer.rs:414-418, no v.z guard on the sampling side, tracer.rs:510-520) and poisons the pixel for good — once per
// ~6e8 samples of the demo scene in the reference's arithmetic.  A contracted dot product practically never lands on an exact
// zero (the 48-bit products do not cancel), so the fused form would silently lose that behaviour.
template <class R> PTB_DEV V3<R> to_local_view(const ShadeCtx<R>& c, V3<R> w) {
#ifdef PTB_CONTRACT_VIEW_COSINE
    return to_local(c, w);
#else
    return V3<R>(dot(w, c.t), dot(w, c.b), dot_rn(w, c.n));
#endif
}
template <class R> PTB_DEV V3<R> to_world(const ShadeCtx<R>& c, V3<R> l) { return l.x * c.t + l.y * c.b + l.z * c.n; }

template <class R> PTB_DEV void shade_ctx_init(ShadeCtx<R>& c, const Mat<R>& m, R eta, V3<R> n, V3<R> v_world) {
    c.n = n;
    onb(n, c.t, c.b);
    c.v = to_local_view(c, v_world);
    c.eta = eta;
    // get_spec_color, tracer.rs:335-341
    R lum = luminance(m.rgb);
    V3<R> ctint = lum > R(0) ? div_s(m.rgb, lum) : V3<R>(1, 1, 1);
    R f0 = m_div(R(1) - eta, R(1) + eta);
    c.spec_col = mix3((f0 * f0) * mix3(V3<R>(1, 1, 1), ctint, m.specular_tint), m.rgb, m.metallic);
    c.sheen_col = mix3(V3<R>(1, 1, 1), ctint, m.sheen_tint);
    c.lum = lum;
    c.wd0 = lum * (R(1) - m.metallic) * (R(1) - m.spec_trans);
    c.wc0 = R(0.25) * m.clearcoat * (R(1) - m.metallic);
}

// tracer.rs:421-433
template <class R>
PTB_DEV void lobe_probabilities(const Mat<R>& m, const ShadeCtx<R>& c, R approx_fresnel, R& wd, R& wr, R& wt, R& wc) {
    wd = c.wd0;
    wr = luminance(mix3(c.spec_col, V3<R>(1, 1, 1), approx_fresnel));
    wt = (R(1) - approx_fresnel) * (R(1) - m.metallic) * m.spec_trans * c.lum;
    wc = c.wc0;
    R total = wd + wr + wt + wc;
    wd = m_div(wd, total); wr = m_div(wr, total); wt = m_div(wt, total); wc = m_div(wc, total);
}

// lobe evaluations in the local frame, tracer.rs:343-419
template <class R> PTB_DEV V3<R> eval_diffuse(const Mat<R>& m, V3<R> c_sheen, V3<R> v, V3<R> l, V3<R> h, R& pdf) {
    pdf = 0;
    if (l.z <= R(0)) return V3<R>(0, 0, 0);
    R ldh = dot(l, h);
    R fl = schlick_fresnel(l.z);
    R fv = schlick_fresnel(v.z);
    R fh = schlick_fresnel(ldh);
    R fd90 = R(0.5) + R(2) * ldh * ldh * m.roughness;
    R fd = mix1(R(1), fd90, fl) * mix1(R(1), fd90, fv);
    R fss90 = ldh * ldh * m.roughness;
    R fss = mix1(R(1), fss90, fl) * mix1(R(1), fss90, fv);
    R ss = R(1.25) * (fss * (m_rcp(l.z + v.z) - R(0.5)) + R(0.5));
    V3<R> fsheen = (fh * m.sheen) * c_sheen;
    pdf = l.z * Const<R>::INV_PI;
    return ((R(1) - m.metallic) * (R(1) - m.spec_trans)) * ((Const<R>::INV_PI * mix1(fd, ss, m.subsurface)) * m.rgb + fsheen);
}
template <class R> PTB_DEV V3<R> eval_spec_reflection(const Mat<R>& m, R fm, V3<R> spec_col, V3<R> v, V3<R> l, V3<R> h, R& pdf) {
    pdf = 0;
    if (l.z <= R(0)) return V3<R>(0, 0, 0);
    V3<R> f = mix3(spec_col, V3<R>(1, 1, 1), fm);
    R d = gtr2aniso(h.z, h.x, h.y, m.ax, m.ay);
    R g1 = smithganiso(m_abs(v.z), v.x, v.y, m.ax, m.ay);
    R g2 = g1 * smithganiso(m_abs(l.z), l.x, l.y, m.ax, m.ay);
    pdf = m_div(g1 * d, R(4) * v.z);
    return div_s((d * g2) * f, R(4) * l.z * v.z);
}
template <class R> PTB_DEV V3<R> eval_spec_refraction(const Mat<R>& m, R eta, R f, V3<R> v, V3<R> l, V3<R> h, R& pdf) {
    pdf = 0;
    if (l.z >= R(0)) return V3<R>(0, 0, 0);
    R vdh = dot(v, h), ldh = dot(l, h);
    R d = gtr2aniso(h.z, h.x, h.y, m.ax, m.ay);
    R g1 = smithganiso(m_abs(v.z), v.x, v.y, m.ax, m.ay);
    R g2 = g1 * smithganiso(m_abs(l.z), l.x, l.y, m.ax, m.ay);
    R denom = ldh + vdh * eta;
    denom *= denom;
    R eta2 = eta * eta;
    R jacobian = m_div(m_abs(ldh), denom);
    pdf = m_div(g1 * m_max(R(0), vdh) * d * jacobian, v.z);
    R s = m_div((R(1) - m.metallic) * m.spec_trans * (R(1) - f) * d * g2 * m_abs(vdh) * jacobian * eta2, m_abs(l.z * v.z));
    return s * V3<R>(m_pow(m.rgb.x, R(0.5)), m_pow(m.rgb.y, R(0.5)), m_pow(m.rgb.z, R(0.5)));
}
template <class R> PTB_DEV V3<R> eval_clearcoat(const Mat<R>& m, V3<R> v, V3<R> l, V3<R> h, R& pdf) {
    pdf = 0;
    if (l.z <= R(0)) return V3<R>(0, 0, 0);
    R vdh = dot(v, h);
    R fh = dielectric_fresnel(vdh, R(1) / R(1.5));
    R f = mix1(R(0.04), R(1), fh);
    R d = gtr1(h.z, m.clearcoat_roughness);
    R g = smithg(l.z, R(0.25)) * smithg(v.z, R(0.25));
    R jacobian = m_rcp(R(4) * vdh);
    pdf = d * h.z * jacobian;
    R s = m_div(m.clearcoat * f * d * g, R(4) * l.z * v.z);
    return s * V3<R>(R(0.25), R(0.25), R(0.25));
}

// lobe ids
enum { LOBE_DIFFUSE = 0, LOBE_CLEARCOAT = 1, LOBE_REFLECT = 2, LOBE_REFRACT = 3 };

// What to evaluate at a local direction pair (l, h): which lobes, and the factor each lobe's pdf
// is multiplied with.  disney_eval (NEE) and disney_sample (BSDF sampling) both reduce to this,
// so the four eval_* bodies exist ONCE in the integrator.
template <class R> struct LobeQuery {
    V3<R> l, h;          // local frame
    uint32_t mask;       // bit k: evaluate lobe k
    R pw[4];             // pdf weight per lobe id
    R diel;              // dielectric_fresnel(|v.h|, eta): tracer.rs:390 and the dielectric half of 374/437
    R fm;                // disney_fresnel(l.h, v.h), tracer.rs:374 (same operands as 531/596 => same value)
};

// f = sum of the enabled lobes (reference order: diffuse, spec reflection, spec refraction,
// clearcoat; tracer.rs:601-623), pdf = sum lobe_pdf * weight.
template <class R>
PTB_DEV void eval_lobes(const Mat<R>& m, const ShadeCtx<R>& c, const LobeQuery<R>& q, V3<R>& f, R& pdf_out, uint32_t* ev) {
    f = V3<R>(0, 0, 0);
    pdf_out = 0;
    R pdf;
    if (q.mask & (1u << LOBE_DIFFUSE)) { f = f + eval_diffuse(m, c.sheen_col, c.v, q.l, q.h, pdf); pdf_out += pdf * q.pw[LOBE_DIFFUSE]; if (ev) ev[LOBE_DIFFUSE]++; }
    if (q.mask & (1u << LOBE_REFLECT)) { f = f + eval_spec_reflection(m, q.fm, c.spec_col, c.v, q.l, q.h, pdf); pdf_out += pdf * q.pw[LOBE_REFLECT]; if (ev) ev[LOBE_REFLECT]++; }
    if (q.mask & (1u << LOBE_REFRACT)) { f = f + eval_spec_refraction(m, c.eta, q.diel, c.v, q.l, q.h, pdf); pdf_out += pdf * q.pw[LOBE_REFRACT]; if (ev) ev[LOBE_REFRACT]++; }
    if (q.mask & (1u << LOBE_CLEARCOAT)) { f = f + eval_clearcoat(m, c.v, q.l, q.h, pdf); pdf_out += pdf * q.pw[LOBE_CLEARCOAT]; if (ev) ev[LOBE_CLEARCOAT]++; }
}

// Query of Tracer::disney_eval, tracer.rs:555-600, for the world-space light direction.
template <class R> PTB_DEV void query_for_eval(const Mat<R>& m, const ShadeCtx<R>& c, V3<R> l_world, LobeQuery<R>& q) {
    const V3<R> v = c.v;
    const V3<R> l = to_local(c, l_world);
    V3<R> h = l.z > R(0) ? normalize(l + v) : normalize(l + c.eta * v);
    if (h.z < R(0)) h = -h;
    R wd, wr, wt, wc;
    q.diel = dielectric_fresnel(m_abs(dot(v, h)), c.eta);
    q.fm = mix1(q.diel, schlick_fresnel(dot(l, h)), m.metallic);                    // = disney_fresnel, tracer.rs:596
    lobe_probabilities(m, c, q.fm, wd, wr, wt, wc);
    q.l = l; q.h = h;
    q.pw[LOBE_DIFFUSE] = wd; q.pw[LOBE_CLEARCOAT] = wc; q.pw[LOBE_REFLECT] = wr; q.pw[LOBE_REFRACT] = wt;
    q.mask = 0;
    if (wd > R(0) && l.z > R(0)) q.mask |= 1u << LOBE_DIFFUSE;                       // tracer.rs:602
    if (wr > R(0) && l.z > R(0) && v.z > R(0)) q.mask |= 1u << LOBE_REFLECT;        // tracer.rs:608
    if (wt > R(0) && l.z < R(0)) q.mask |= 1u << LOBE_REFRACT;                      // tracer.rs:614
    if (wc > R(0) && l.z > R(0) && v.z > R(0)) q.mask |= 1u << LOBE_CLEARCOAT;      // tracer.rs:620
}

template <class R> PTB_DEV V3<R> reflect(V3<R> i, V3<R> n) {                        // tracer.rs:464-466
    R d = dot(n, i);
    return V3<R>(i.x - R(2) * n.x * d, i.y - R(2) * n.y * d, i.z - R(2) * n.z * d);
}
template <class R> PTB_DEV V3<R> refract(V3<R> i, V3<R> n, R eta) {                 // tracer.rs:468-475
    R ndi = dot(n, i);
    R k = R(1) - eta * eta * (R(1) - ndi * ndi);
    if (k < R(0)) return V3<R>(0, 0, 0);
    return eta * i - (eta * ndi + m_sqrt(k)) * n;
}

// Query of Tracer::disney_sample, tracer.rs:441-549: lobe pick by CDF order diffuse, clearcoat,
// specular (495-523), direction sampling, and the factor the chosen lobe's pdf is scaled with
// (507, 520, 540-548).  (r1, r2, coin) are the draws at 446, 447, 534; l_prev_world is the stale
// `l` read at 531 (quirk A.5).  Returns the lobe id.
template <class R>
PTB_DEV int query_for_sample(const Mat<R>& m, const ShadeCtx<R>& c, R r1, R r2, R coin, V3<R> l_prev_world, LobeQuery<R>& q) {
    R wd, wr, wt, wc;
    R approx_fresnel = disney_fresnel(m, c.eta, c.v.z, c.v.z);
    lobe_probabilities(m, c, approx_fresnel, wd, wr, wt, wc);
    const R cdf0 = wd;
    const R cdf1 = cdf0 + wc;
    int lobe;
    if (r1 < cdf0) { lobe = LOBE_DIFFUSE; r1 = m_div(r1, cdf0); }
    else if (r1 < cdf1) { lobe = LOBE_CLEARCOAT; r1 = m_div(r1 - cdf0, cdf1 - cdf0); }
    else { lobe = LOBE_REFLECT; r1 = m_div(r1 - cdf1, R(1) - cdf1); }
    // azimuth: TWO_PI * r2 (diffuse 328, spec 265), r1 * TWO_PI for the clearcoat lobe (246, quirk A.4)
    R sn, cs;
    m_sincos(Const<R>::TWO_PI * (lobe == LOBE_CLEARCOAT ? r1 : r2), &sn, &cs);
    V3<R> l, h;
    R w;
    q.diel = 0; q.fm = 0;
    if (lobe == LOBE_DIFFUSE) {
        l = cosine_sample_hemisphere(r1, sn, cs);
        h = normalize(l + c.v);
        w = wd;
    } else {
        h = lobe == LOBE_CLEARCOAT ? sample_gtr1(m.clearcoat_roughness, r1, sn, cs) : sample_ggxvndf(c.v, m.ax, m.ay, r1, sn, cs);
        if (h.z < R(0)) h = -h;
        w = wc;
        bool refr = false;
        if (lobe != LOBE_CLEARCOAT) {
            q.diel = dielectric_fresnel(m_abs(dot(c.v, h)), c.eta);
            R fresnel = mix1(q.diel, schlick_fresnel(dot(l_prev_world, h)), m.metallic);   // tracer.rs:531, stale l
            R ff = R(1) - ((R(1) - fresnel) * m.spec_trans * (R(1) - m.metallic));
            refr = !(coin < ff);
            w = (refr ? R(1) - ff : ff) * (wr + wt);
            if (refr) lobe = LOBE_REFRACT;
        }
        l = normalize(refr ? refract(-c.v, h, c.eta) : reflect(-c.v, h));
        q.fm = mix1(q.diel, schlick_fresnel(dot(l, h)), m.metallic);                       // tracer.rs:374 with the new l
    }
    q.l = l; q.h = h;
    q.mask = 1u << lobe;
    q.pw[0] = q.pw[1] = q.pw[2] = q.pw[3] = w;
    return lobe;
}

// Tracer::disney_eval, tracer.rs:555-626 (returns |l.z| * f) — standalone form for the parity kernels
template <class R> PTB_DEV V3<R> disney_eval(const Mat<R>& m, const ShadeCtx<R>& c, V3<R> l_world, R& bsdf_pdf, uint32_t* ev = nullptr) {
    LobeQuery<R> q;
    query_for_eval(m, c, l_world, q);
    V3<R> f;
    eval_lobes(m, c, q, f, bsdf_pdf, ev);
    return m_abs(q.l.z) * f;
}
// Tracer::disney_sample, tracer.rs:441-553: returns |n.l| * f, writes l (world), pdf, lobe — standalone form
template <class R>
PTB_DEV V3<R> disney_sample(const Mat<R>& m, const ShadeCtx<R>& c, R r1, R r2, R coin, V3<R> l_prev_world, V3<R>& l_world, R& pdf,
                            int& lobe) {
    LobeQuery<R> q;
    lobe = query_for_sample(m, c, r1, r2, coin, l_prev_world, q);
    V3<R> f;
    eval_lobes(m, c, q, f, pdf, (uint32_t*)nullptr);
    l_world = to_world(c, q.l);
    return m_abs(dot(c.n, l_world)) * f;
}

// Tracer::sample_light, tracer.rs:173-220
template <class R> struct LightSample { V3<R> normal, emission, direction; R dist, pdf; };
// the two kinds the reference leaves unimplemented (tracer.rs:217 `_ => {}`), with the upstream GLSL project's semantics
template <class R> PTB_DEV LightSample<R> sample_light_extended(const DLight<R>& L, R n_lights_f, V3<R> scatter_pos, R r1, R r2) {
    LightSample<R> ls;
    const V3<R> lp(L.px, L.py, L.pz);
    ls.emission = n_lights_f * V3<R>(L.ex, L.ey, L.ez);
    if (L.type == PTB_LIGHT_RECTANGULAR) {
        const V3<R> u(L.ux, L.uy, L.uz), v(L.vx, L.vy, L.vz);
        const V3<R> surf = (lp + r1 * u) + r2 * v;
        ls.direction = surf - scatter_pos;
        ls.dist = length(ls.direction);
        const R dist_sq = ls.dist * ls.dist;
        ls.direction = div_s(ls.direction, ls.dist);
        ls.normal = normalize(cross(u, v));
        ls.pdf = m_div(dist_sq, L.area * m_abs(dot(ls.normal, ls.direction)));
    } else {                                            // PTB_LIGHT_DISTANT
        ls.direction = normalize(lp);
        ls.normal = normalize(scatter_pos - lp);
        ls.dist = Const<R>::MAXV;
        ls.pdf = R(1);
    }
    return ls;
}
template <class R> PTB_DEV LightSample<R> sample_light(const DLight<R>& L, R n_lights_f, V3<R> scatter_pos, R r1, R r2) {
    LightSample<R> ls;
    V3<R> lp(L.px, L.py, L.pz);
    V3<R> c2s = scatter_pos - lp;
    R dist_c = length(c2s);
    R rr = m_sqrt(m_max(R(0), R(1) - r1 * r1));           // uniform_sample_hemisphere, tracer.rs:178-182
    R phi = Const<R>::TWO_PI * r2;
    R s, c;
    m_sincos(phi, &s, &c);
    V3<R> sd(rr * c, rr * s, r1);
    c2s = div_s(c2s, dist_c);
    V3<R> t, b;
    onb(c2s, t, b);
    V3<R> dir = sd.x * t + sd.y * b + sd.z * c2s;
    V3<R> surf = lp + L.radius * dir;
    ls.direction = surf - scatter_pos;
    ls.dist = length(ls.direction);
    R dist_sq = ls.dist * ls.dist;
    ls.direction = div_s(ls.direction, ls.dist);
    ls.normal = normalize(surf - lp);
    ls.emission = n_lights_f * V3<R>(L.ex, L.ey, L.ez);
    ls.pdf = m_div(dist_sq, L.area * R(0.5) * m_abs(dot(ls.normal, ls.direction)));
    return ls;
}

// State::finalize, globals.rs:50-62
template <class R> PTB_DEV void state_finalize(V3<R> o, V3<R> d, R hit_dist, V3<R> normal, Mat<R>& m, V3<R>& fhp, V3<R>& ffn, R& eta) {
    fhp = o + hit_dist * d;
    R nd = dot(normal, d);
    ffn = nd <= R(0) ? normal : -normal;
    mat_finalize(m);
    eta = nd < R(0) ? m_rcp(m.ior) : m.ior;
}

// ------------------------------------------------------------------------------------------------
// One path: the body of the per-pixel loop, tracer.rs:51-103.
template <class R> struct PathState {
    V3<R> o, d;             // ray (ray.rs)
    V3<R> thr, rad;         // throughput, radiance (tracer.rs:51-52)
    R hit_dist;             // State::hit_dist, carried across bounces (A.1)
    R prev_pdf;             // scatter_sample.pdf of the previous bounce (0 on the first)
    uint32_t bounce;
    uint32_t medium;        // 0 = outside; else 1 + index of the material whose medium the path is in (only maintained when s.has_media)
};

struct PathCounters {   // per-thread event counts (only when collect_counters)
    uint32_t closest_hit, any_hit, shade, nee_contrib, eval_calls, lobe[4], end_sky, end_emitter, end_pdf, end_depth, end_rr, ev[4];
    uint32_t bvh[2];    // sphere-BVH work: [0] inner nodes visited (one visit tests both children), [1] leaf spheres tested
};

template <class R> PTB_DEV void path_begin(const DScene<R>& s, PathState<R>& p, uint32_t x, uint32_t row, uint32_t W, uint32_t H,
                                           R inv_w, R inv_h, R j0, R j1, bool film_fast = false, float rcp_w = 0.0f, float rcp_h = 0.0f) {
    R px, py;
    if constexpr (sizeof(R) == 4) {
        if (film_fast) film_coords_fma(x, row, W, H, rcp_w, rcp_h, px, py);
        else film_coords<R>(x, row, W, H, px, py);
    } else {
        film_coords<R>(x, row, W, H, px, py);
    }
    gen_ray(s, px, py, j0, j1, inv_w, inv_h, p.o, p.d);
    p.thr = V3<R>(1, 1, 1);
    p.rad = V3<R>(0, 0, 0);
    p.hit_dist = R(-1);        // globals.rs:28
    p.prev_pdf = 0;
    p.bounce = 0;
    p.medium = 0;
}

// Second half of a bounce, tracer.rs:72-101, for a path that hit geometry (not a light), in three pieces so that
// the global-memory wavefront can run the shadow ray in its own kernel between them:
//   shade_setup       finalize (globals.rs:50-62), emission (tracer.rs:74), shading frame / specular colours
//   shade_nee_sample  direct_light up to the shadow ray (tracer.rs:131-150): light pick, light sample, cull
//   shade_finish      disney_eval + MIS towards the light (tracer.rs:155-164), disney_sample (92-97), next ray (100-101)
template <class R> struct ShadeSetup {
    V3<R> fhp, ffn;
    R eta;
    ShadeCtx<R> c;
};
template <class R> struct NeeSample {
    bool wants_shadow_ray;     // the sample passed the back-face cull: visibility decides whether it contributes
    LightSample<R> ls;
    R light_area;
    V3<R> scatter_pos;
};

// ADD_EMISSION = false: the caller has added `emission * throughput` itself (the shared-memory wavefront keeps radiance and
// throughput out of registers while it shades)
template <class R, bool COUNT, bool ADD_EMISSION = true>
PTB_DEV void shade_setup(const DScene<R>& s, PathState<R>& p, V3<R> normal, Mat<R>& mat, ShadeSetup<R>& su, PathCounters* pc) {
    state_finalize(p.o, p.d, p.hit_dist, normal, mat, su.fhp, su.ffn, su.eta);
    if (ADD_EMISSION) p.rad = p.rad + mat.emission * p.thr;     // tracer.rs:74
    if (COUNT) pc->shade++;
    shade_ctx_init(su.c, mat, su.eta, su.ffn, -p.d);
}

template <class R, bool XL = true, bool EMB = false>
PTB_DEV void shade_nee_sample(const DScene<R>& s, const SceneView<R>& sv, const ShadeSetup<R>& su, const R* u, NeeSample<R>& ns) {
    ns.wants_shadow_ray = false;
    ns.light_area = 0;
    if (s.n_lights > 0) {
        uint32_t li = (uint32_t)(u[SLOT_LIGHT_PICK] * s.n_lights_f);          // tracer.rs:137-139
        ns.scatter_pos = su.fhp + s.eps * su.ffn;
        const DLight<R> L = SV_LIGHT(li);
        const bool extended = XL && (s.flags & PTB_SCENE_EXTENDED_LIGHTS) != 0 && L.type != PTB_LIGHT_SPHERICAL;
        if (extended) ns.ls = sample_light_extended(L, s.n_lights_f, ns.scatter_pos, u[SLOT_LIGHT_R1], u[SLOT_LIGHT_R2]);
        else ns.ls = sample_light(L, s.n_lights_f, ns.scatter_pos, u[SLOT_LIGHT_R1], u[SLOT_LIGHT_R2]);
        ns.light_area = L.area;
        // tracer.rs:148; a non-spherical light of the reference leaves direction and normal at zero: never a shadow ray
        ns.wants_shadow_ray = (L.type == PTB_LIGHT_SPHERICAL || extended) && dot(ns.ls.direction, ns.ls.normal) < R(0);
    }
}

// `nee`: the light sample is visible (passed the cull and the shadow ray).  Returns true while the path continues.
// DEFER (global-memory wavefront): `nee` only says that the sample passed the cull; the contribution `ld * throughput`
// (tracer.rs:164 and :89) is handed to `sink(contribution, flags)` instead of being added, and the shadow-ray kernel adds
// it if the ray turns out unoccluded — the same additions in the same order, without keeping the shading context alive
// across kernels.  flags: bits 0..3 = lobes evaluated towards the light, bit 4 = pdf > 0 (the sample can contribute).
struct NoSink { template <class V> PTB_DEV void operator()(V, uint32_t) const {} };
// UNROLL: two copies of the lobe code, one per pass: the compiler overlaps the independent passes (+5 % in the resolved-material
// wavefront kernel, +4.5 % in the fused kernel on the demo scene, -1 % on the depth-16 stress scene; no effect in the streaming
// shade kernel, which keeps the single copy).  The single copy dates from the 142 KB fused kernel that was instruction-cache bound.
template <class R, bool COUNT, bool DEFER = false, class Sink = NoSink, bool UNROLL = false>
PTB_DEV bool shade_finish(const DScene<R>& s, PathState<R>& p, const Mat<R>& mat, const ShadeSetup<R>& su, bool nee, const LightSample<R>& ls,
                          R light_area, const R* u, PathCounters* pc, const Sink& sink = Sink()) {
    const ShadeCtx<R>& c = su.c;
    // BSDF evaluation, ONE copy of the lobe code for both uses:
    //   pass 0 = disney_eval towards the light sample (tracer.rs:155), only for un-shadowed lanes;
    //   pass 1 = disney_sample (tracer.rs:92); stale `l` = previous sampled direction = ray dir (A.5)
    V3<R> f, l_world;
    R pdf = 0;
    int lobe = 0;
#pragma unroll (UNROLL ? 2 : 1)
    for (int pass = 0; pass < 2; ++pass) {
        if (pass == 0 && !nee) continue;
        LobeQuery<R> q;
        if (pass == 0) {
            if (COUNT && !DEFER) pc->eval_calls++;
            query_for_eval(mat, c, ls.direction, q);
        } else {
            V3<R> l_prev = p.bounce == 0 ? V3<R>(0, 0, 0) : p.d;
            lobe = query_for_sample(mat, c, u[SLOT_BSDF_R1], u[SLOT_BSDF_R2], u[SLOT_COIN], l_prev, q);
            if (COUNT) pc->lobe[lobe]++;
        }
        eval_lobes(mat, c, q, f, pdf, COUNT && !(DEFER && pass == 0) ? pc->ev : (uint32_t*)nullptr);
        if (pass == 0) {
            f = m_abs(q.l.z) * f;                                   // tracer.rs:625
            R w = R(1);
            if (light_area > R(0)) w = power_heuristic(ls.pdf, pdf); // tracer.rs:157-160
            if (pdf > R(0)) {                                       // tracer.rs:162-164
                V3<R> ld = (w * ls.emission) * div_s(f, ls.pdf);
                if (DEFER) sink(ld * p.thr, q.mask | 16u);
                else {
                    p.rad = p.rad + ld * p.thr;
                    if (COUNT) pc->nee_contrib++;
                }
            } else if (DEFER) {
                sink(V3<R>(0, 0, 0), q.mask);
            }
        } else {
            l_world = to_world(c, q.l);                             // tracer.rs:551-552
            f = m_abs(dot(c.n, l_world)) * f;
        }
    }
    const V3<R> l = l_world;
    if (!(pdf > R(0))) {
        if (COUNT) pc->end_pdf++;
        return false;
    }
    p.thr = p.thr * div_s(f, pdf);
    p.prev_pdf = pdf;
    p.d = l;                                                    // tracer.rs:100-101
    p.o = su.fhp + s.eps * l;
    p.bounce++;
    if (p.bounce >= s.depth) {
        if (COUNT) pc->end_depth++;
        return false;
    }
    return true;
}

// the three pieces in one go (fused integrator, shared-memory wavefront)
template <class R, bool COUNT, bool BVH, bool ADD_EMISSION = true, bool SDF = true, bool EMB = false, bool XL = true>
PTB_DEV bool path_shade(const DScene<R>& s, const SceneView<R>& sv, PathState<R>& p, V3<R> normal, Mat<R>& mat, const R* u, PathCounters* pc) {
    ShadeSetup<R> su;
    shade_setup<R, COUNT, ADD_EMISSION>(s, p, normal, mat, su, pc);
    NeeSample<R> ns;
    shade_nee_sample<R, XL, EMB>(s, sv, su, u, ns);
    bool nee = false;
    if (ns.wants_shadow_ray) {
        if (COUNT) pc->any_hit++;
        nee = !any_hit<R, BVH, SDF, EMB>(s, sv, ns.scatter_pos, ns.ls.direction, ns.ls.dist - s.eps, COUNT ? pc->bvh : nullptr);   // tracer.rs:150-154
    }
    return shade_finish<R, COUNT, false, NoSink, true>(s, p, mat, su, nee, ns.ls, ns.light_area, u, pc);
}

// ------------------------------------------------------------------------------------------------
// Resolved-material table (f32 staged integrators, small scenes).  Everything a shaded bounce derives from the hit
// material ALONE is a function of (which primitives were accepted, checker cell parity, entering / exiting): the
// assignment replay of hit_material, Material::finalize (material.rs:117-131), get_spec_color (tracer.rs:335-341), the
// eta pick of State::finalize (globals.rs:58-61) and the material-only lobe weights (tracer.rs:423, 426).  The host
// evaluates these once per scene in plain f32 (x86-64, no contraction: the reference's own arithmetic) and the shade stage
// reads the entry from shared memory instead of recomputing it per bounce; fields are fetched where they are used, so the
// material does not occupy registers across the whole stage.
struct alignas(16) RMat {
    Mat<float> m;               // patched + finalized
    float eta[2];               // [0] entering (1 / ior), [1] exiting (ior)
    float spec_col[2][3];       // get_spec_color per side
    float sheen_col[3];
    float lum, wd0, wc0;
    uint32_t lobe_class;
};
// key word of the table: bits 0..15 first entry, bits 16..31 = 1 + index of the material whose checker decides between
// entry and entry + 1 (0: the albedo is constant)
PTB_DEV bool checker_odd(float x, float y) {              // analytical.rs:107-111: the cell that takes checker_b
    float x1 = fmod2_int(m_floor(x));
    float y1 = fmod2_int(m_floor(y));
    return !(fmod2_int(x1 + y1) < 1.0f);
}
PTB_DEV uint32_t rm_key_of(const DScene<float>& s, const SceneView<float>& sv, int prim, uint32_t accepted_lo) {
    return s.patch_materials ? accepted_lo : prim_material<float, false>(s, sv, prim);
}
PTB_DEV const RMat& rm_lookup(const DScene<float>& s, const SceneView<float>& sv, const uint32_t* rm_keys, const RMat* rm_table, uint32_t key, V3<float> rd) {
    const uint32_t k = rm_keys[key];
    uint32_t e = k & 0xffffu;
    const uint32_t cm = k >> 16;
    if (cm) {
        const DMaterial<float>& dm = sv.materials[cm - 1u];
        if (checker_odd(div_rn(rd.x, rd.y) * dm.checker_scale + dm.checker_offset, div_rn(rd.z, rd.y) * dm.checker_scale + dm.checker_offset)) e += 1u;
    }
    return rm_table[e];
}
// shade_setup for a table entry
template <bool COUNT, bool ADD_EMISSION = true>
PTB_DEV void shade_setup_rm(PathState<float>& p, V3<float> normal, const RMat& rm, ShadeSetup<float>& su, PathCounters* pc) {
    su.fhp = p.o + p.hit_dist * p.d;
    const float nd = dot(normal, p.d);
    su.ffn = nd <= 0.0f ? normal : -normal;
    const int side = nd < 0.0f ? 0 : 1;
    su.eta = rm.eta[side];
    if (ADD_EMISSION) p.rad = p.rad + rm.m.emission * p.thr;    // tracer.rs:74
    if (COUNT) pc->shade++;
    ShadeCtx<float>& c = su.c;
    c.n = su.ffn;
    onb(c.n, c.t, c.b);
    c.v = to_local_view(c, -p.d);
    c.eta = su.eta;
    c.spec_col = V3<float>(rm.spec_col[side][0], rm.spec_col[side][1], rm.spec_col[side][2]);
    c.sheen_col = V3<float>(rm.sheen_col[0], rm.sheen_col[1], rm.sheen_col[2]);
    c.lum = rm.lum; c.wd0 = rm.wd0; c.wc0 = rm.wc0;
}
template <bool COUNT, bool ADD_EMISSION = true, bool EMB = false>
PTB_DEV bool path_shade_rm(const DScene<float>& s, const SceneView<float>& sv, PathState<float>& p, V3<float> normal, const RMat& rm, const float* u,
                           PathCounters* pc) {
    ShadeSetup<float> su;
    shade_setup_rm<COUNT, ADD_EMISSION>(p, normal, rm, su, pc);
    NeeSample<float> ns;
    shade_nee_sample<float, false, EMB>(s, sv, su, u, ns);
    bool nee = false;
    if (ns.wants_shadow_ray) {
        if (COUNT) pc->any_hit++;
        nee = !any_hit<float, false, false, EMB>(s, sv, ns.scatter_pos, ns.ls.direction, ns.ls.dist - s.eps);   // tracer.rs:150-154
    }
#ifdef PTB_RM_SINGLE_COPY
    return shade_finish<float, COUNT, false, NoSink, false>(s, p, rm.m, su, nee, ns.ls, ns.light_area, u, pc);
#else
    return shade_finish<float, COUNT, false, NoSink, true>(s, p, rm.m, su, nee, ns.ls, ns.light_area, u, pc);
#endif
}

// Russian roulette EXTENSION at the start of bounce > 0 (the reference has none, quirk A.12; off in
// every parity run): survival probability from the throughput (GLSL-PathTracer's rule), decided by
// slot 0 of this bounce, which is free after bounce 0 (slots 0,1 are the camera jitter).
template <class R> PTB_DEV bool russian_roulette_survives(PathState<R>& p, R u0) {
    R q = m_max(p.thr.x, m_max(p.thr.y, p.thr.z)) + R(0.001);
    q = q > R(0.95) ? R(0.95) : q;
    if (u0 >= q) return false;
    p.thr = m_rcp(q) * p.thr;
    return true;
}

// tracer.rs:66-69: the path left the scene
template <class R> PTB_DEV void path_add_sky(const DScene<R>& s, PathState<R>& p) {
    V3<R> bg = background(s, p.d);
    p.rad = p.rad + bg * p.thr;
}
// tracer.rs:72-87: the path hit a light.  finalize() would run first but none of its outputs reach the
// radiance; the emission of the geometry's material (if geometry was also hit) is still added (tracer.rs:74)
template <class R, bool BVH>
PTB_DEV void path_add_emitter(const DScene<R>& s, const SceneView<R>& sv, PathState<R>& p, const HitCore<R>& h) {
    if (h.geom && s.has_emissive) {
        Mat<R> mat;
        hit_material<R, BVH>(s, sv, h.prim, h.accepted, p.d, mat);
        p.rad = p.rad + mat.emission * p.thr;
    }
    R w = power_heuristic(p.prev_pdf, h.light_pdf);             // `state.depth > 0` is always true (A.2)
    p.rad = p.rad + (w * h.light_emission) * p.thr;
}

// First half of a bounce, tracer.rs:63-87: closest_hit, background on a miss, MIS-weighted emission
// on a light hit.  Returns 0 = path ended, 1 = geometry hit (continue with path_shade).
// FX = false (fused integrator on scenes without a signed-distance program, media or live extended lights): none of that code
template <class R, bool COUNT, bool BVH, bool FX = true>
PTB_DEV int path_intersect(const DScene<R>& s, const SceneView<R>& sv, PathState<R>& p, HitCore<R>& h, PathCounters* pc) {
    if (p.bounce >= s.depth) {                                  // recursion_depth() == 0: `for _ in 0..0` never runs (tracer.rs:61)
        if (COUNT) pc->end_depth++;
        return 0;
    }
    if (COUNT) pc->closest_hit++;
    h = closest_hit_core<R, BVH, FX, FX>(s, sv, p.o, p.d, p.hit_dist, COUNT ? pc->bvh : nullptr);
    p.hit_dist = h.hit_dist;
    if (!h.hit) {
        path_add_sky(s, p);
        if (COUNT) pc->end_sky++;
        return 0;
    }
    if (h.is_emitter) {
        path_add_emitter<R, BVH>(s, sv, p, h);
        if (COUNT) pc->end_emitter++;
        return 0;
    }
    return 1;
}

// ------------------------------------------------------------------------------------------------
// Media (PTB_MEDIUM_* in ptb200.h; material.rs:5-34, State::medium globals.rs:19, the `_is_surface` hook of direct_light,
// tracer.rs:125).  The reference never reads its Medium: the semantics are the GLSL project's it was ported from, stated
// identically by the oracle (Tracer::medium_step).  Henyey-Greenstein in the pbrt convention: cos is taken between the
// direction back along the ray and the new direction, forward scattering (g > 0) peaks at cos = -1.
template <class R> PTB_DEV R phase_hg(R cos_theta, R g) {
    const R denom = (R(1) + g * g) + (R(2) * g) * cos_theta;
    return Const<R>::INV_4PI * m_div(R(1) - g * g, denom * m_sqrt(denom));
}
template <class R> PTB_DEV V3<R> sample_hg(V3<R> v, R g, R r1, R r2) {
    R cos_theta;
    if (m_abs(g) < R(0.001)) cos_theta = R(1) - R(2) * r2;
    else {
        const R sqr = m_div(R(1) - g * g, (R(1) + g) - (R(2) * g) * r2);
        cos_theta = -m_div((R(1) + g * g) - sqr * sqr, R(2) * g);
    }
    const R sin_theta = m_sqrt(m_max(R(0), R(1) - cos_theta * cos_theta));
    R sn, cs;
    m_sincos(Const<R>::TWO_PI * r1, &sn, &cs);
    V3<R> t, b;
    onb(v, t, b);
    return ((sin_theta * cs) * t + (sin_theta * sn) * b) + cos_theta * v;
}
// The medium's part of a bounce, for a path that is inside one (p.medium != 0) and has just hit geometry at p.hit_dist.
// Returns 0: go on with the surface (absorption / emission applied), 1: the bounce happened in the medium and the path
// continues from there, 2: it happened in the medium and the path ended (depth).  Draws: slot 1 (free after bounce 0; a path
// cannot be inside on bounce 0) for the free-flight distance, the light slots as at a surface, the two BSDF slots for the phase
// function.
enum : int { MED_SURFACE = 0, MED_SCATTERED = 1, MED_ENDED = 2 };
// `defer` (global-memory wavefront): the shadow ray of the in-medium light sample is not traced here; the sample is handed back
// (origin, direction, length, contribution already multiplied by the throughput) for the shadow-ray kernel to add if unoccluded
template <class R> struct MediumNee { bool wants; V3<R> pos, dir, contrib; R max_dist; };
template <class R, bool COUNT, bool BVH, bool SDF>
PTB_DEV int path_medium(const DScene<R>& s, const SceneView<R>& sv, PathState<R>& p, const R* u, PathCounters* pc, MediumNee<R>* defer = nullptr) {
    if (defer) defer->wants = false;
    const DMaterial<R>& mm = BVH ? s.materials[p.medium - 1u] : sv.materials[p.medium - 1u];
    const R density = mm.med_density, t = p.hit_dist;
    const V3<R> color(mm.med_color[0], mm.med_color[1], mm.med_color[2]);
    if (mm.med_type == PTB_MEDIUM_ABSORB) {
        p.thr = p.thr * V3<R>(m_exp((-(R(1) - color.x) * t) * density), m_exp((-(R(1) - color.y) * t) * density), m_exp((-(R(1) - color.z) * t) * density));
        return MED_SURFACE;
    }
    if (mm.med_type == PTB_MEDIUM_EMISSIVE) {
        p.rad = p.rad + ((t * density) * color) * p.thr;
        return MED_SURFACE;
    }
    const R sd = m_min(m_div(-m_ln(u[SLOT_JITTER_Y]), density), t);
    if (!(sd < t)) return MED_SURFACE;
    p.thr = p.thr * color;
    p.o = p.o + sd * p.d;
    // direct_light(.., is_surface = false): the sample leaves from the point itself, the phase function is value and pdf
    ShadeSetup<R> su;
    su.fhp = p.o; su.ffn = V3<R>(R(0), R(0), R(0));
    NeeSample<R> ns;
    shade_nee_sample(s, sv, su, u, ns);
    const V3<R> back = -p.d;
    if (ns.wants_shadow_ray) {
        if (COUNT) pc->any_hit++;
        if (defer || !any_hit<R, BVH, SDF>(s, sv, ns.scatter_pos, ns.ls.direction, ns.ls.dist - s.eps, COUNT ? pc->bvh : nullptr)) {
            const R ph = phase_hg(dot(back, ns.ls.direction), mm.med_g);
            R w = R(1);
            if (ns.light_area > R(0)) w = power_heuristic(ns.ls.pdf, ph);
            if (ph > R(0)) {
                const V3<R> c = ((w * ns.ls.emission) * V3<R>(m_div(ph, ns.ls.pdf), m_div(ph, ns.ls.pdf), m_div(ph, ns.ls.pdf))) * p.thr;
                if (defer) {
                    defer->wants = true; defer->pos = ns.scatter_pos; defer->dir = ns.ls.direction; defer->max_dist = ns.ls.dist - s.eps; defer->contrib = c;
                } else {
                    p.rad = p.rad + c;
                    if (COUNT) pc->nee_contrib++;
                }
            }
        }
    }
    const V3<R> dir = sample_hg(back, mm.med_g, u[SLOT_BSDF_R1], u[SLOT_BSDF_R2]);
    p.prev_pdf = phase_hg(dot(back, dir), mm.med_g);
    p.d = dir;
    p.bounce++;
    if (p.bounce >= s.depth) {
        if (COUNT) pc->end_depth++;
        return MED_ENDED;
    }
    return MED_SCATTERED;
}
// after a surface bounce: a path that leaves a surface with a medium towards its back side is inside it, towards the front
// outside; surfaces without a medium change nothing (no nesting).  p.d is the new direction, `normal` the geometry normal.
template <class R, bool BVH>
PTB_DEV void path_medium_update(const DScene<R>& s, const SceneView<R>& sv, PathState<R>& p, V3<R> normal, uint32_t mi) {
    const DMaterial<R>& mm = BVH ? s.materials[mi] : sv.materials[mi];
    if (mm.med_type != PTB_MEDIUM_NONE) p.medium = dot(p.d, normal) < R(0) ? mi + 1u : 0u;
}

// Runs ONE bounce (both halves); returns true while the path continues.  COUNT enables event counters.
template <class R, bool COUNT, bool BVH, bool FX = true>
PTB_DEV bool path_bounce(const DScene<R>& s, const SceneView<R>& sv, PathState<R>& p, const R* u, uint32_t rr_start, PathCounters* pc) {
    if (rr_start != 0 && p.bounce >= rr_start && p.bounce > 0) {
        if (!russian_roulette_survives(p, u[0])) {
            if (COUNT) pc->end_rr++;
            return false;
        }
    }
    HitCore<R> h;
    if (!path_intersect<R, COUNT, BVH, FX>(s, sv, p, h, pc)) return false;
    if (FX && s.has_media && p.medium) {
        const int med = path_medium<R, COUNT, BVH, true>(s, sv, p, u, pc);
        if (med != MED_SURFACE) return med == MED_SCATTERED;
    }
    Mat<R> mat;
    const uint32_t mi = hit_material<R, BVH>(s, sv, h.prim, h.accepted, p.d, mat);
    V3<R> normal = hit_normal<R, BVH, FX>(s, sv, h.prim, p.o, p.d, h.hit_dist);
    const bool alive = path_shade<R, COUNT, BVH, true, FX, false, FX>(s, sv, p, normal, mat, u, pc);
    if (FX && s.has_media) path_medium_update<R, BVH>(s, sv, p, normal, mi);
    return alive;
}

}  // namespace ptb
