// ptb_kernels.cuh — __global__ kernels: the fused persistent integrator, ColorBuffer kernels and
// the per-function parity kernels.  (The wavefront stage kernels live in ptb_wavefront.cuh.)
#pragma once
#include "ptb_device.cuh"

namespace ptb {

template <class R> struct Vec4T;
template <> struct Vec4T<float> { using type = float4; };
template <> struct Vec4T<double> { using type = double4; };
PTB_DEV float4 mk4(float a, float b, float c, float d) { return make_float4(a, b, c, d); }
PTB_DEV double4 mk4(double a, double b, double c, double d) { return make_double4(a, b, c, d); }

struct DeviceCounters {   // mirrors ptb_counters
    unsigned long long samples, closest_hit, any_hit, shade, nee_contrib, eval_calls, lobe[4], end_sky, end_emitter, end_pdf,
        end_depth, end_rr, ev[4], bvh_nodes, bvh_leaf_tests;
};

struct RenderArgs {
    void* accum;                 // W*H x (sum r, sum g, sum b, sample count) in R
    void* flush_dst;             // non-NULL: a pixel's partial sum (sum r, g, b, spp) of this launch is STORED here instead of
                                 // being added to `accum` — the root GPU's slot buffer, written over NVLink (ptb_peer_*)
    uint32_t W, H;
    uint32_t spp;
    uint64_t sample_base;
    uint64_t seed;
    uint32_t rr_start;
    uint32_t tiles_x, n_items;   // 16x16 pixel tiles; n_items = tiles * 256
    unsigned int* work_counter;  // global pixel-chunk dispenser
    DeviceCounters* counters;
    // film coordinates without a division (wavefront integrator): rcp_w / rcp_h are the correctly rounded 1/W, 1/H and
    // film_fast says that the host verified, for EVERY column and row of this frame size, that the FMA-corrected quotient
    // (film_coords_fma) equals the IEEE x / W and y / H of tracer.rs:34-46 bit for bit
    float rcp_w, rcp_h;
    double rcp_w64, rcp_h64;     // the same pixel size for the f64 instantiation
    uint32_t film_fast;
    // tail items (wavefront integrator): work items [n_whole, n_items) are the frame's last tail_zt pixels (in tile order) cut
    // into 2^tail_log2b sample blocks; item n_whole + b * tail_zt + z is block b of tail pixel z and its sum goes to tail_side
    uint32_t n_whole, tail_zt, tail_log2b;
    void* tail_side;
};

constexpr int FUSED_THREADS = 256;
// pixels a warp reserves per global atomic: two rows of a 16x16 tile (512 B of accumulator, flushed as two 256 B segments).
// A whole tile per atomic starved warps on small frames: 960x540 is 2040 tiles for 3552 (wavefront) / 2368 (fused) warps.
#ifndef PTB_CHUNK
#define PTB_CHUNK 32
#endif
constexpr uint32_t FUSED_CHUNK = PTB_CHUNK;

// ------------------------------------------------------------------------------------------------
// Fused persistent integrator.
//
// One lane = one path at a time.  A lane owns one pixel for `spp` consecutive samples (its sum
// lives in registers and is flushed with a single 16-byte read-modify-write), and every loop
// iteration runs exactly ONE bounce for all live lanes.  Lanes whose path ended regenerate in
// place (next sample of their pixel, or a new pixel from the warp's tile), so the warp stays full
// until the frame runs out of pixels: termination divergence — the reference's early `break`s,
// 84 % of paths end on the sky within two bounces — costs nothing, and there is no path state in
// HBM at all.  Pixels are handed out in 16x16-tile order, FUSED_CHUNK (two tile rows) per global atomic;
// within a warp they are distributed with ballot/popc prefix ranks.  Every pixel's samples are
// summed in sample order by one lane, so the image is bit-reproducible run to run.
// FX: scenes with a signed-distance program, media or live extended lights (their code inlined cost every other scene a CTA per SM)
template <class R, bool COUNT, bool BVH, bool FX>
__global__ void __launch_bounds__(FUSED_THREADS) k_render_fused(const __grid_constant__ DScene<R> s, const RenderArgs a) {
    __shared__ SceneSmem<R> sm;
    const SceneView<R> sv = stage_scene(s, sm.words, PTB_SMEM_SCENE_BYTES);
    using V4 = typename Vec4T<R>::type;
    V4* accum = reinterpret_cast<V4*>(a.accum);

    const unsigned FULL = 0xffffffffu;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const R inv_w = R(1) / (R)a.W, inv_h = R(1) / (R)a.H;    // pixel_size, pinhole.rs:41

    // warp-uniform tile cursor
    uint32_t w_next = 0, w_end = 0;
    // lane state
    bool have_pixel = false, alive = false, done = false;
    uint32_t pix = 0, px = 0, prow = 0, s_idx = 0;
    V3<R> acc(0, 0, 0);
    PathState<R> p;
    PathCounters pc;
    if (COUNT) {
        pc.closest_hit = pc.any_hit = pc.shade = pc.nee_contrib = pc.eval_calls = 0;
        pc.lobe[0] = pc.lobe[1] = pc.lobe[2] = pc.lobe[3] = 0;
        pc.end_sky = pc.end_emitter = pc.end_pdf = pc.end_depth = pc.end_rr = 0;
        pc.ev[0] = pc.ev[1] = pc.ev[2] = pc.ev[3] = 0;
        pc.bvh[0] = pc.bvh[1] = 0;
    }
    uint32_t n_samples = 0;

    while (true) {
        // ---- pixel hand-out -------------------------------------------------------------------
        bool want = !alive && !done && (!have_pixel || s_idx == a.spp);
        if (want && have_pixel) {
            if (a.flush_dst) {
                reinterpret_cast<V4*>(a.flush_dst)[pix] = mk4(acc.x, acc.y, acc.z, (R)a.spp);
            } else {
                V4 v = accum[pix];
                v.x += acc.x; v.y += acc.y; v.z += acc.z; v.w += (R)a.spp;
                accum[pix] = v;
            }
            have_pixel = false;
        }
        unsigned need = __ballot_sync(FULL, want);
        while (need) {
            if (w_next == w_end) {
                uint32_t base = 0;
                if (lane == 0) base = atomicAdd(a.work_counter, FUSED_CHUNK);
                base = __shfl_sync(FULL, base, 0);
                if (base >= a.n_items) {          // frame exhausted
                    if (want) { done = true; want = false; }
                    break;
                }
                w_next = base;
                w_end = base + FUSED_CHUNK;
            }
            const uint32_t avail = w_end - w_next;
            const uint32_t rank = __popc(need & lt_mask);
            if (want && rank < avail) {
                const uint32_t idx = w_next + rank;
                const uint32_t tile = idx >> 8, within = idx & 255u;
                px = (tile % a.tiles_x) * 16u + (within & 15u);
                prow = (tile / a.tiles_x) * 16u + (within >> 4);
                want = false;
                if (px < a.W && prow < a.H) {
                    pix = prow * a.W + px;
                    have_pixel = true;
                    s_idx = 0;
                    acc = V3<R>(0, 0, 0);
                } else {
                    want = true;                  // outside the frame (partial tile): take another
                }
            }
            const uint32_t n_need = __popc(need);
            w_next += n_need < avail ? n_need : avail;
            need = __ballot_sync(FULL, want);
        }
        if (__all_sync(FULL, done)) break;

        // ---- one bounce for every live lane -------------------------------------------------------
        const bool start = !alive && !done;
        if (start) p.bounce = 0;
        if (alive || start) {
            Rng<R> rng(pix, a.sample_base + s_idx, a.seed);
            R u[8];
            rng.draws(p.bounce, u);
            if (start) {
                path_begin(s, p, px, prow, a.W, a.H, inv_w, inv_h, u[0], u[1]);
                alive = true;
                if (COUNT) n_samples++;
            }
            alive = path_bounce<R, COUNT, BVH, FX>(s, sv, p, u, a.rr_start, &pc);
            if (!alive) {
                acc = acc + p.rad;
                s_idx++;
            }
        }
    }

    if (COUNT) {
        DeviceCounters* c = a.counters;
        atomicAdd(&c->samples, (unsigned long long)n_samples);
        atomicAdd(&c->closest_hit, (unsigned long long)pc.closest_hit);
        atomicAdd(&c->any_hit, (unsigned long long)pc.any_hit);
        atomicAdd(&c->shade, (unsigned long long)pc.shade);
        atomicAdd(&c->nee_contrib, (unsigned long long)pc.nee_contrib);
        atomicAdd(&c->eval_calls, (unsigned long long)pc.eval_calls);
        for (int i = 0; i < 4; ++i) atomicAdd(&c->lobe[i], (unsigned long long)pc.lobe[i]);
        for (int i = 0; i < 4; ++i) atomicAdd(&c->ev[i], (unsigned long long)pc.ev[i]);
        atomicAdd(&c->end_sky, (unsigned long long)pc.end_sky);
        atomicAdd(&c->end_emitter, (unsigned long long)pc.end_emitter);
        atomicAdd(&c->end_pdf, (unsigned long long)pc.end_pdf);
        atomicAdd(&c->end_depth, (unsigned long long)pc.end_depth);
        atomicAdd(&c->end_rr, (unsigned long long)pc.end_rr);
        atomicAdd(&c->bvh_nodes, (unsigned long long)pc.bvh[0]);
        atomicAdd(&c->bvh_leaf_tests, (unsigned long long)pc.bvh[1]);
    }
}

// ------------------------------------------------------------------------------------------------
// ColorBuffer kernels (buffer.rs)

// accumulators (sum, count) -> mean image, alpha 1 where samples landed (tracer.rs:59,105-117)
template <class R> __global__ void k_resolve(const typename Vec4T<R>::type* accum, typename Vec4T<R>::type* out, uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    auto v = accum[i];
    if (v.w > R(0)) out[i] = mk4(div_rn(v.x, v.w), div_rn(v.y, v.w), div_rn(v.z, v.w), R(1));
    else out[i] = mk4(R(0), R(0), R(0), R(0));
}
// mean image + frame count -> accumulators
template <class R> __global__ void k_unresolve(const typename Vec4T<R>::type* mean, typename Vec4T<R>::type* accum, R frames, uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    auto v = mean[i];
    accum[i] = mk4(v.x * frames, v.y * frames, v.z * frames, frames);
}

// Edge-avoiding a-trous wavelet filter, one level (Dammertz et al. 2010, guided by the colour alone: this path produces no
// normal / albedo buffers).  5x5 taps of the B3 spline (1/16, 1/4, 3/8, 1/4, 1/16) spread `step` pixels apart, each weighted by
// exp(-|c_p - c_q|^2 * inv_sigma2); non-finite neighbours are skipped, and a non-finite centre (a pixel the path loop poisoned
// with 0/0, which the reference never filters) takes the plain spline average of its finite neighbours.  The reference lists a
// denoiser as open (Readme.md:14): this is the device form of that item, checked against a numpy statement of the same filter.
template <class R>
__global__ void k_atrous(const typename Vec4T<R>::type* in, typename Vec4T<R>::type* out, uint32_t W, uint32_t H, int step, R inv_sigma2) {
    const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= W || y >= H) return;
    const R k[3] = {R(0.375), R(0.25), R(0.0625)};
    const auto c = in[(size_t)y * W + x];
    const bool c_ok = isfinite(c.x) && isfinite(c.y) && isfinite(c.z);
    R sx = 0, sy = 0, sz = 0, sw = 0;
    for (int dy = -2; dy <= 2; ++dy) {
        const int yy = (int)y + dy * step;
        if (yy < 0 || yy >= (int)H) continue;
        for (int dx = -2; dx <= 2; ++dx) {
            const int xx = (int)x + dx * step;
            if (xx < 0 || xx >= (int)W) continue;
            const auto q = in[(size_t)yy * W + xx];
            if (!(isfinite(q.x) && isfinite(q.y) && isfinite(q.z))) continue;
            R w = k[dx < 0 ? -dx : dx] * k[dy < 0 ? -dy : dy];
            if (c_ok) {
                const R ex = q.x - c.x, ey = q.y - c.y, ez = q.z - c.z;
                w *= (R)exp(-(double)((ex * ex + ey * ey + ez * ez) * inv_sigma2));
            }
            sx += w * q.x; sy += w * q.y; sz += w * q.z; sw += w;
        }
    }
    out[(size_t)y * W + x] = sw > R(0) ? mk4(div_rn(sx, sw), div_rn(sy, sw), div_rn(sz, sw), c.w) : c;
}

// multi-GPU gather: accumulators += sum over the n_slots partial-sum buffers of one step, in slot order (deterministic)
__global__ void k_peer_sum(float4* accum, const float4* slots, uint32_t n_slots, uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 v = accum[i];
    for (uint32_t r = 0; r < n_slots; ++r) {
        const float4 p = slots[(size_t)r * n + i];
        v.x += p.x; v.y += p.y; v.z += p.z; v.w += p.w;
    }
    accum[i] = v;
}

// Rust `as u8` (saturating, NaN -> 0, truncating)
PTB_DEV uint8_t as_u8(double v) {
    if (!(v == v)) return 0;
    if (v <= 0.0) return 0;
    if (v >= 255.0) return 255;
    return (uint8_t)v;
}
// x.powf(0.4545) * 255 in F.  powf is evaluated in double and rounded once, which reproduces
// glibc's (nearly always correctly rounded) f32 powf; the product is then formed in F.
PTB_DEV uint8_t encode_gamma(float x) { float g = (float)pow((double)x, (double)0.4545f); return as_u8((double)(g * 255.0f)); }
PTB_DEV uint8_t encode_gamma(double x) { return as_u8(pow(x, 0.4545) * 255.0); }
PTB_DEV uint8_t encode_linear(float x) { return as_u8((double)(x * 255.0f)); }
PTB_DEV uint8_t encode_linear(double x) { return as_u8(x * 255.0); }

// ColorBuffer::convert_to_u8, buffer.rs:55-64.  `from_accum`: input is (sum,count) instead of a mean image.
template <class R> __global__ void k_convert_u8(const typename Vec4T<R>::type* in, uchar4* out, uint32_t n, int from_accum) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    auto v = in[i];
    if (from_accum) {
        if (v.w > R(0)) v = mk4(div_rn(v.x, v.w), div_rn(v.y, v.w), div_rn(v.z, v.w), R(1));
        else v = mk4(R(0), R(0), R(0), R(0));
    }
    out[i] = make_uchar4(encode_gamma(v.x), encode_gamma(v.y), encode_gamma(v.z), encode_linear(v.w));
}
// ColorBuffer::convert_to_u8_at, buffer.rs:67-102: frame fw x fh, buffer bw x bh placed at (at0, at1);
// one thread per FRAME pixel; j counts frame rows from the END; strict `>` bounds; no gamma.
template <class R>
__global__ void k_convert_u8_at(const typename Vec4T<R>::type* accum, uint32_t bw, uint32_t bh, uchar4* frame, uint32_t at0, uint32_t at1,
                                uint32_t fw, uint32_t fh, int from_accum) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= fw * fh) return;
    uint32_t frow = t / fw, ii = t % fw;
    uint32_t j = fh - 1u - frow;
    uint32_t i = j * fw + ii;
    uint32_t x = i % fw;
    uint32_t y = fh - (i / fw);
    if (x > at0 && x < at0 + bw && y > at1 && y < at1 + bh) {
        auto v = accum[(x - at0) + (y - at1) * bw];
        if (from_accum) {
            if (v.w > R(0)) v = mk4(div_rn(v.x, v.w), div_rn(v.y, v.w), div_rn(v.z, v.w), R(1));
            else v = mk4(R(0), R(0), R(0), R(0));
        }
        frame[t] = make_uchar4(encode_linear(v.x), encode_linear(v.y), encode_linear(v.z), encode_linear(v.w));
    }
}

// ------------------------------------------------------------------------------------------------
// Per-function parity kernels: SoA in, SoA out, one thread per element; they call the SAME device
// functions as the integrators.
template <class R> PTB_DEV V3<R> ld3(const R* a, size_t n, size_t i) { return V3<R>(a[i], a[n + i], a[2 * n + i]); }
template <class R> PTB_DEV void st3(R* a, size_t n, size_t i, V3<R> v) { a[i] = v.x; a[n + i] = v.y; a[2 * n + i] = v.z; }

#define PTB_TEST_PROLOGUE size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;

template <class R> __global__ void k_test_sphere_hit(size_t n, const R* o, const R* d, const R* c, const R* r, R* t_out) {
    PTB_TEST_PROLOGUE
    if (i >= n) return;
    t_out[i] = isect_sphere(ld3(o, n, i), ld3(d, n, i), ld3(c, n, i), r[i]);
}
template <class R> __global__ void k_test_plane_hit(size_t n, const R* o, const R* d, const R* p, const R* nn, R* t_out) {
    PTB_TEST_PROLOGUE
    if (i >= n) return;
    t_out[i] = isect_plane(ld3(o, n, i), ld3(d, n, i), ld3(p, n, i), ld3(nn, n, i));
}
template <class R>
__global__ void k_test_gen_ray(const __grid_constant__ DScene<R> s, size_t n, const R* p2, const R* off2, R w, R h, R* o_out, R* d_out) {
    PTB_TEST_PROLOGUE
    if (i >= n) return;
    V3<R> o, d;
    gen_ray(s, p2[i], p2[n + i], off2[i], off2[n + i], R(1) / w, R(1) / h, o, d);
    st3(o_out, n, i, o);
    st3(d_out, n, i, d);
}
template <class R>
__global__ void k_test_closest_hit(const __grid_constant__ DScene<R> s, size_t n, const R* o, const R* d, const R* hd_in, uint32_t* hit,
                                   uint32_t* em, R* hd_out, R* nrm, uint32_t* mat_out, R* lpdf, R* lem) {
    __shared__ SceneSmem<R> sm;
    const SceneView<R> sv = stage_scene(s, sm.words, PTB_SMEM_SCENE_BYTES);
    PTB_TEST_PROLOGUE
    if (i >= n) return;
    Mat<R> m;
    HitRec<R> h = s.use_bvh ? closest_hit<R, true>(s, sv, ld3(o, n, i), ld3(d, n, i), hd_in[i], m)
                            : closest_hit<R, false>(s, sv, ld3(o, n, i), ld3(d, n, i), hd_in[i], m);
    hit[i] = h.hit; em[i] = h.is_emitter; hd_out[i] = h.hit_dist;
    st3(nrm, n, i, h.normal);
    mat_out[i] = h.material;
    lpdf[i] = h.light_pdf;
    st3(lem, n, i, h.light_emission);
}
template <class R>
__global__ void k_test_any_hit(const __grid_constant__ DScene<R> s, size_t n, const R* o, const R* d, const R* md, uint32_t* hit) {
    __shared__ SceneSmem<R> sm;
    const SceneView<R> sv = stage_scene(s, sm.words, PTB_SMEM_SCENE_BYTES);
    PTB_TEST_PROLOGUE
    if (i >= n) return;
    hit[i] = s.use_bvh ? any_hit<R, true>(s, sv, ld3(o, n, i), ld3(d, n, i), md[i]) : any_hit<R, false>(s, sv, ld3(o, n, i), ld3(d, n, i), md[i]);
}
// the signed-distance program on its own: value + material at a point; sphere trace + normal along a ray
template <class R> __global__ void k_test_sdf_eval(const __grid_constant__ DScene<R> s, size_t n, const R* q, R* dist, uint32_t* mat) {
    PTB_TEST_PROLOGUE
    if (i >= n) return;
    const SdfSample<R> v = sdf_eval(s, ld3(q, n, i));
    dist[i] = v.d; mat[i] = v.material;
}
template <class R>
__global__ void k_test_sdf_trace(const __grid_constant__ DScene<R> s, size_t n, const R* o, const R* d, const R* limit, R* t_out, R* nrm, uint32_t* mat) {
    PTB_TEST_PROLOGUE
    if (i >= n) return;
    const V3<R> oo = ld3(o, n, i), dd = ld3(d, n, i);
    uint32_t m = 0xffffffffu;
    const R t = sdf_trace(s, oo, dd, limit[i], m);
    t_out[i] = t; mat[i] = m;
    st3(nrm, n, i, t >= R(0) ? sdf_normal(s, oo + t * dd) : V3<R>(0, 0, 0));
}
template <class R> __global__ void k_test_background(const __grid_constant__ DScene<R> s, size_t n, const R* d, R* rgb) {
    PTB_TEST_PROLOGUE
    if (i >= n) return;
    st3(rgb, n, i, background(s, ld3(d, n, i)));
}
template <class R>
__global__ void k_test_sample_light(const __grid_constant__ DScene<R> s, size_t n, uint32_t li, const R* pos, const R* r1, const R* r2,
                                    R* nrm, R* em, R* dir, R* dist, R* pdf) {
    PTB_TEST_PROLOGUE
    if (i >= n) return;
    const DLight<R> L = s.lights[li];
    LightSample<R> ls;
    if ((s.flags & PTB_SCENE_EXTENDED_LIGHTS) && L.type != PTB_LIGHT_SPHERICAL) ls = sample_light_extended(L, s.n_lights_f, ld3(pos, n, i), r1[i], r2[i]);
    else if (L.type == PTB_LIGHT_SPHERICAL) ls = sample_light(L, s.n_lights_f, ld3(pos, n, i), r1[i], r2[i]);
    else { ls.normal = ls.emission = ls.direction = V3<R>(R(0), R(0), R(0)); ls.dist = ls.pdf = R(0); }     // tracer.rs:217 `_ => {}`
    st3(nrm, n, i, ls.normal); st3(em, n, i, ls.emission); st3(dir, n, i, ls.direction);
    dist[i] = ls.dist; pdf[i] = ls.pdf;
}
template <class R>
__global__ void k_test_finalize(const __grid_constant__ DScene<R> s, size_t n, uint32_t mi, const R* o, const R* d, const R* hd,
                                const R* nrm, R* rough, R* ccr, R* ax, R* ay, R* eta_out, R* ffn_out, R* fhp_out) {
    PTB_TEST_PROLOGUE
    if (i >= n) return;
    Mat<R> m;
    mat_load(m, s.materials[mi], ld3(d, n, i));
    V3<R> fhp, ffn;
    R eta;
    state_finalize(ld3(o, n, i), ld3(d, n, i), hd[i], ld3(nrm, n, i), m, fhp, ffn, eta);
    rough[i] = m.roughness; ccr[i] = m.clearcoat_roughness; ax[i] = m.ax; ay[i] = m.ay; eta_out[i] = eta;
    st3(ffn_out, n, i, ffn); st3(fhp_out, n, i, fhp);
}
template <class R>
__global__ void k_test_disney_eval(const __grid_constant__ DScene<R> s, size_t n, uint32_t mi, const R* eta, const R* v, const R* nrm,
                                   const R* l, R* f_out, R* pdf_out) {
    PTB_TEST_PROLOGUE
    if (i >= n) return;
    Mat<R> m;
    V3<R> vw = ld3(v, n, i);
    mat_load(m, s.materials[mi], -vw);
    mat_finalize(m);
    ShadeCtx<R> c;
    shade_ctx_init(c, m, eta[i], ld3(nrm, n, i), vw);
    R pdf;
    V3<R> f = disney_eval(m, c, ld3(l, n, i), pdf);
    st3(f_out, n, i, f);
    pdf_out[i] = pdf;
}
template <class R>
__global__ void k_test_disney_sample(const __grid_constant__ DScene<R> s, size_t n, uint32_t mi, const R* eta, const R* v, const R* nrm,
                                     const R* lprev, const R* r1, const R* r2, const R* coin, uint32_t* lobe_out, R* l_out, R* f_out,
                                     R* pdf_out) {
    PTB_TEST_PROLOGUE
    if (i >= n) return;
    Mat<R> m;
    V3<R> vw = ld3(v, n, i);
    mat_load(m, s.materials[mi], -vw);
    mat_finalize(m);
    ShadeCtx<R> c;
    shade_ctx_init(c, m, eta[i], ld3(nrm, n, i), vw);
    V3<R> l;
    R pdf;
    int lobe;
    V3<R> f = disney_sample(m, c, r1[i], r2[i], coin[i], ld3(lprev, n, i), l, pdf, lobe);
    lobe_out[i] = (uint32_t)lobe;
    st3(l_out, n, i, l); st3(f_out, n, i, f);
    pdf_out[i] = pdf;
}
template <class R>
__global__ void k_test_rng(size_t n, const uint32_t* pixel, const unsigned long long* sample, uint32_t bounce, uint64_t seed, R* out8) {
    PTB_TEST_PROLOGUE
    if (i >= n) return;
    Rng<R> rng(pixel[i], sample[i], seed);
    R u[8];
    rng.draws(bounce, u);
    for (int k = 0; k < 8; ++k) out8[(size_t)k * n + i] = u[k];
}

}  // namespace ptb
