// ptb_api.cu — the C ABI of include/ptb200.h: tracer handle, scene export -> device buffers
// (+ sphere BVH), ColorBuffer transfers, render dispatch, per-function parity entry points.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC
// (see __graft_entry__.build()).  There is deliberately no host fallback anywhere in this file.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <type_traits>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/ptb200.h"
#include "ptb_kernels.cuh"
#ifdef PTB_WF_ASYNC   // A/B only (tools/ab_variants.py): barrier-free per-key rings — measured slower: desynchronised warps thrash the
#include "ptb_wavefront_async.cuh"      // instruction cache (no_instruction 0.21 -> 2.6 stall cycles per issue; profiles/r02_ab_variants.txt)
#else
#include "ptb_wavefront.cuh"
#endif
#include "ptb_stream.cuh"

using namespace ptb;

// ------------------------------------------------------------------------------------------------
// error plumbing
static thread_local std::string g_err;
static int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}
#define CU(call)                                                                                           \
    do {                                                                                                   \
        cudaError_t e_ = (call);                                                                           \
        if (e_ != cudaSuccess) return fail(PTB_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

// ------------------------------------------------------------------------------------------------
// device buffers of one scene
template <class R> struct SceneBuffers {
    DScene<R> d{};
    void* blob = nullptr; void* bvh = nullptr; void* bvh_prim = nullptr; void* bvh_spheres = nullptr;
    void* lbvh = nullptr; void* lbvh_prim = nullptr; void* lbvh_spheres = nullptr;
    size_t cap_lbvh = 0, cap_lprim = 0, cap_lspheres = 0;
    size_t cap_blob = 0, cap_bvh = 0, cap_prim = 0, cap_spheres = 0;   // allocations are reused across set_scene calls:
    size_t bytes = 0;                                                  // cudaFree would synchronise the whole device
    uint32_t rm_entries_built = 0;
    // the sphere BVH on the device is keyed by the sphere data it was built from: re-exporting a scene whose spheres are unchanged
    // (ptb_set_scene_* per frame, the way the reference re-reads its scene per ray) re-uploads the arrays but not the tree
    uint64_t bvh_key = 0; size_t bvh_key_n = 0, bvh_bytes = 0; bool bvh_cached = false;                                     // resolved-material entries in the blob (d.rm_entries is 0 while a signed-distance program is attached)
    void release() {
        for (void** p : {&blob, &bvh, &bvh_prim, &bvh_spheres, &lbvh, &lbvh_prim, &lbvh_spheres}) {
            if (*p) cudaFree(*p);
            *p = nullptr;
        }
        bvh_cached = false;
        cap_blob = cap_bvh = cap_prim = cap_spheres = cap_lbvh = cap_lprim = cap_lspheres = 0;
        bytes = 0;
    }
};

// Host copy of the camera parameters (resize re-derives the basis because `ratio` depends on W/H).
template <class R> struct CamParams { R origin[3], center[3], fov; };

struct ptb_tracer {
    ptb_config cfg{};
    CamParams<float> c32{};
    CamParams<double> c64{};
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int precision = 0;   // 0 none, 4 f32, 8 f64
    SceneBuffers<float> s32;
    SceneBuffers<double> s64;
    uint32_t W = 0, H = 0;
    void* accum = nullptr;          // device, W*H*4 reals
    bool own_accum = false;
    size_t accum_bytes = 0;
    void* staging = nullptr;        // device, W*H*4 reals (mean image) or u8 frame
    size_t staging_bytes = 0;
    uint64_t frames = 0;
    unsigned int* work_counter = nullptr;
    DeviceCounters* counters = nullptr;
    uint64_t launches = 0;
    uint32_t last_integrator = 0, last_kernel_bits = 0;     // what the last ptb_render ran (ptb_last_integrator)
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool timed = false;
    int sm_count = 0;
    int fused_blocks_f32 = 0, fused_blocks_f64 = 0;
    WavefrontState wf;
    StreamState st;
    // multi-GPU gather over peer memory (ptb_peer_*): 2 x n_slots frame-sized partial-sum buffers living on the ROOT GPU
    void* peer_base = nullptr;      // root: own allocation; other ranks: the root's allocation mapped through CUDA IPC
    bool peer_owner = false, peer_ipc = false;
    uint32_t peer_slots = 0;
    size_t peer_frame_bytes = 0;
    void* flush_dst = nullptr;      // current target of ptb_render's per-pixel partial sums (NULL: add to the accumulators)
    // drop-in path: host buffer that holds exactly the device image as of `host_clean_frames` frames, because this tracer's
    // last render_frame / download wrote it (NULL: unknown).  With PTB_FRAME_HOST_UNCHANGED the caller vouches that it has
    // not touched the buffer since, and the upload of the next render_frame is skipped.
    const void* host_clean = nullptr;
    uint64_t host_clean_frames = 0;
    // asynchronous download (ptb_download_async_*): side stream + events, so the copy of step k overlaps the trace of step k+1
    cudaStream_t dl_stream = nullptr;
    cudaEvent_t dl_ready = nullptr;
    cudaEvent_t dl_free[2] = {nullptr, nullptr};   // the D2H copy out of dl_staging[k] has completed
    bool dl_used[2] = {false, false};
    void* dl_staging[2] = {nullptr, nullptr};
    size_t dl_staging_bytes = 0;
    uint32_t dl_parity = 0;
};

static size_t real_size(const ptb_tracer* t) { return (size_t)t->precision; }

// ------------------------------------------------------------------------------------------------
// BVH build over spheres (host, binned SAH, children stored adjacently)
namespace {
struct Box { float lo[3], hi[3]; };
inline Box box_empty() { return Box{{3e38f, 3e38f, 3e38f}, {-3e38f, -3e38f, -3e38f}}; }
inline void box_grow(Box& b, const Box& o) {
    for (int k = 0; k < 3; ++k) { b.lo[k] = std::min(b.lo[k], o.lo[k]); b.hi[k] = std::max(b.hi[k], o.hi[k]); }
}
inline float box_area(const Box& b) {
    float dx = b.hi[0] - b.lo[0], dy = b.hi[1] - b.lo[1], dz = b.hi[2] - b.lo[2];
    if (dx < 0 || dy < 0 || dz < 0) return 0.f;
    return 2.f * (dx * dy + dy * dz + dz * dx);
}
struct BvhBuilder {
    std::vector<Box> pbox;
    std::vector<float> pcen;   // 3 per prim
    std::vector<uint32_t> prim;
    std::vector<BvhNode> nodes;
    // The device traversal keeps a fixed stack of BVH_STACK entries (ptb_device.cuh) and pushes at most one entry per level,
    // so the tree may not be deeper than that.  Binned SAH on strongly non-uniform inputs can produce long chains; once the
    // levels a balanced subtree of `count` primitives still needs would no longer fit, the split falls back to the median
    // along the widest centroid axis, which halves `count` at every level and so keeps the bound.
    uint32_t max_depth = 0;
    static uint32_t ceil_log2(uint32_t n) { uint32_t l = 0; while ((1u << l) < n) ++l; return l; }
    void build_node(uint32_t ni, uint32_t first, uint32_t count, uint32_t depth = 0) {
        max_depth = std::max(max_depth, depth);
        Box nb = box_empty(), cb = box_empty();
        for (uint32_t i = first; i < first + count; ++i) {
            box_grow(nb, pbox[prim[i]]);
            const float* c = &pcen[3 * prim[i]];
            for (int k = 0; k < 3; ++k) { cb.lo[k] = std::min(cb.lo[k], c[k]); cb.hi[k] = std::max(cb.hi[k], c[k]); }
        }
        BvhNode& n = nodes[ni];
        for (int k = 0; k < 3; ++k) { n.lo[k] = nb.lo[k]; n.hi[k] = nb.hi[k]; }
        auto make_leaf = [&]() { nodes[ni].left_or_first = first; nodes[ni].count = count; };
        if (count <= 2) { make_leaf(); return; }
        constexpr int NB = 16;
        float best_cost = 3e38f; int best_axis = -1, best_split = 0;
        for (int ax = 0; ax < 3; ++ax) {
            float ext = cb.hi[ax] - cb.lo[ax];
            if (!(ext > 0.f)) continue;
            Box bb[NB]; uint32_t bc[NB];
            for (int b = 0; b < NB; ++b) { bb[b] = box_empty(); bc[b] = 0; }
            float scale = NB / ext;
            for (uint32_t i = first; i < first + count; ++i) {
                int b = std::min(NB - 1, (int)((pcen[3 * prim[i] + ax] - cb.lo[ax]) * scale));
                bc[b]++; box_grow(bb[b], pbox[prim[i]]);
            }
            float la[NB - 1], ra[NB - 1]; uint32_t lc[NB - 1], rc[NB - 1];
            Box acc = box_empty(); uint32_t cnt = 0;
            for (int b = 0; b < NB - 1; ++b) { box_grow(acc, bb[b]); cnt += bc[b]; la[b] = box_area(acc); lc[b] = cnt; }
            acc = box_empty(); cnt = 0;
            for (int b = NB - 1; b > 0; --b) { box_grow(acc, bb[b]); cnt += bc[b]; ra[b - 1] = box_area(acc); rc[b - 1] = cnt; }
            for (int b = 0; b < NB - 1; ++b) {
                if (lc[b] == 0 || rc[b] == 0) continue;
                float cost = la[b] * lc[b] + ra[b] * rc[b];
                if (cost < best_cost) { best_cost = cost; best_axis = ax; best_split = b; }
            }
        }
        // no useful split for a small node: keep it as a leaf
        if (count <= 4 && (best_axis < 0 || best_cost >= box_area(nb) * count)) { make_leaf(); return; }
        uint32_t mid;
        const bool force_median = depth + ceil_log2(count) + 2u >= (uint32_t)BVH_STACK;
        if (force_median) {
            int ax = 0;
            for (int k = 1; k < 3; ++k) if (cb.hi[k] - cb.lo[k] > cb.hi[ax] - cb.lo[ax]) ax = k;
            mid = first + count / 2;
            std::nth_element(prim.begin() + first, prim.begin() + mid, prim.begin() + first + count,
                             [&](uint32_t a, uint32_t b) { return pcen[3 * a + ax] < pcen[3 * b + ax]; });
        } else if (best_axis >= 0) {
            float ext = cb.hi[best_axis] - cb.lo[best_axis];
            float scale = NB / ext;
            auto it = std::partition(prim.begin() + first, prim.begin() + first + count, [&](uint32_t p) {
                int b = std::min(NB - 1, (int)((pcen[3 * p + best_axis] - cb.lo[best_axis]) * scale));
                return b <= best_split;
            });
            mid = (uint32_t)(it - prim.begin());
        } else {
            mid = first + count / 2;   // all centroids coincide
        }
        if (mid == first || mid == first + count) mid = first + count / 2;
        uint32_t left = (uint32_t)nodes.size();
        nodes.push_back(BvhNode{});
        nodes.push_back(BvhNode{});
        nodes[ni].left_or_first = left;
        nodes[ni].count = 0;
        build_node(left, first, mid - first, depth + 1);
        build_node(left + 1, mid, first + count - mid, depth + 1);
    }
};
}  // namespace

// ------------------------------------------------------------------------------------------------
// Resolved-material table (RMat, ptb_device.cuh): host-side evaluation, in f32 with the reference's operation order
// (x86-64 host code, no contraction).  `chain` = the materials of the accepted primitives in test order (one entry for
// scenes whose materials assign every field); `odd` = the checker cell of the material that supplies the albedo.
namespace {
inline float h_mix(float a, float b, float v) { return (1.0f - v) * a + b * v; }                  // math.rs:33-39 / tracer.rs:228-231
inline void h_rgb(const DMaterial<float>& dm, bool odd, float out[3]) {
    if (dm.albedo_kind == PTB_ALBEDO_CHECKER_DIR_RATIO) { float c = odd ? dm.checker_b : dm.checker_a; out[0] = out[1] = out[2] = c; }
    else { out[0] = dm.rgb[0]; out[1] = dm.rgb[1]; out[2] = dm.rgb[2]; }
}
// index (into chain) of the material whose albedo ends up in state.material.rgb; -1 if it is a constant colour
inline int rm_checker_source(const std::vector<const DMaterial<float>*>& chain) {
    int src = 0;
    bool src_is_rgb_assign = (chain[0]->set_mask & PTB_MAT_RGB) != 0;
    for (size_t i = 1; i < chain.size(); ++i)
        if (chain[i]->set_mask & PTB_MAT_RGB) { src = (int)i; src_is_rgb_assign = true; }
    return (src_is_rgb_assign && chain[src]->albedo_kind == PTB_ALBEDO_CHECKER_DIR_RATIO) ? src : -1;
}
inline RMat rm_resolve(const std::vector<const DMaterial<float>*>& chain, bool odd) {
    RMat r;
    memset(&r, 0, sizeof(r));
    Mat<float>& m = r.m;
    float rgb[3];
    {   // first accepted primitive: Material::new() patched by it == its resolved values (mat_load)
        const DMaterial<float>& dm = *chain[0];
        if (dm.set_mask & PTB_MAT_RGB) h_rgb(dm, odd, rgb); else { rgb[0] = dm.rgb[0]; rgb[1] = dm.rgb[1]; rgb[2] = dm.rgb[2]; }
        m.emission = V3<float>(dm.emission[0], dm.emission[1], dm.emission[2]);
        m.anisotropic = dm.anisotropic; m.metallic = dm.metallic; m.roughness = dm.roughness; m.subsurface = dm.subsurface;
        m.specular_tint = dm.specular_tint; m.sheen = dm.sheen; m.sheen_tint = dm.sheen_tint; m.clearcoat = dm.clearcoat;
        m.clearcoat_gloss = dm.clearcoat_gloss; m.spec_trans = dm.spec_trans; m.ior = dm.ior;
    }
    for (size_t i = 1; i < chain.size(); ++i) {     // later accepted primitives assign their masked fields (mat_patch)
        const DMaterial<float>& dm = *chain[i];
        const uint32_t k = dm.set_mask;
        if (k & PTB_MAT_RGB) h_rgb(dm, odd, rgb);
        if (k & PTB_MAT_EMISSION) m.emission = V3<float>(dm.emission[0], dm.emission[1], dm.emission[2]);
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
    m.rgb = V3<float>(rgb[0], rgb[1], rgb[2]);
    // lobe class from the un-finalized values (lobe_class_of)
    {
        const float nm = 1.0f - m.metallic;
        r.lobe_class = (nm * (1.0f - m.spec_trans) > 0.0f ? 1u : 0u) | (m.clearcoat * nm > 0.0f ? 2u : 0u) | (m.spec_trans * nm > 0.0f ? 4u : 0u);
    }
    // Material::finalize, material.rs:117-131
    m.roughness = std::fmax(m.roughness, 0.01f);
    m.clearcoat_roughness = h_mix(0.1f, 0.001f, m.clearcoat_gloss);
    const float aspect = std::sqrt(1.0f - m.anisotropic * 0.9f);
    m.ax = std::fmax(m.roughness / aspect, 0.001f);
    m.ay = std::fmax(m.roughness * aspect, 0.001f);
    // get_spec_color, tracer.rs:335-341, for both values State::finalize can give eta (globals.rs:58-61)
    const float lum = 0.212671f * rgb[0] + 0.715160f * rgb[1] + 0.072169f * rgb[2];
    float ctint[3];
    for (int k = 0; k < 3; ++k) ctint[k] = lum > 0.0f ? rgb[k] / lum : 1.0f;
    r.eta[0] = 1.0f / m.ior;
    r.eta[1] = m.ior;
    for (int side = 0; side < 2; ++side) {
        const float eta = r.eta[side];
        const float f0 = (1.0f - eta) / (1.0f + eta);
        for (int k = 0; k < 3; ++k) r.spec_col[side][k] = h_mix((f0 * f0) * h_mix(1.0f, ctint[k], m.specular_tint), rgb[k], m.metallic);
    }
    for (int k = 0; k < 3; ++k) r.sheen_col[k] = h_mix(1.0f, ctint[k], m.sheen_tint);
    r.lum = lum;
    r.wd0 = lum * (1.0f - m.metallic) * (1.0f - m.spec_trans);       // tracer.rs:423
    r.wc0 = 0.25f * m.clearcoat * (1.0f - m.metallic);               // tracer.rs:426
    return r;
}
constexpr uint32_t RM_MAX_PATCH_PRIMS = 6;      // partial-mask scenes: one key per subset of the primitives
}  // namespace

// ------------------------------------------------------------------------------------------------
// scene upload
template <class R> struct PodTypes;
template <> struct PodTypes<float> { using scene = ptb_scene_f32; using material = ptb_material_f32; using sdf = ptb_sdf_f32; };
template <> struct PodTypes<double> { using scene = ptb_scene_f64; using material = ptb_material_f64; using sdf = ptb_sdf_f64; };

template <class T> static cudaError_t upload_vec(void** dst, size_t& cap, const std::vector<T>& v, cudaStream_t st, size_t& bytes) {
    size_t n = std::max<size_t>(v.size(), 1) * sizeof(T);
    cudaError_t e = cudaSuccess;
    if (n > cap || !*dst) {
        if (*dst) cudaFree(*dst);
        *dst = nullptr; cap = 0;
        e = cudaMalloc(dst, n);
        if (e != cudaSuccess) return e;
        cap = n;
    }
    bytes += n;
    if (!v.empty()) e = cudaMemcpyAsync(*dst, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, st);
    return e;
}

// Material::new() defaults (material.rs:82-114) for everything a POD material does not assign
template <class R, class PodMaterial> static DMaterial<R> resolve_pod_material(const PodMaterial& m) {
    DMaterial<R> o;
    const uint32_t k = m.set_mask;
    for (int c = 0; c < 3; ++c) { o.rgb[c] = (k & PTB_MAT_RGB) ? m.rgb[c] : R(1.5); o.emission[c] = (k & PTB_MAT_EMISSION) ? m.emission[c] : R(0); }
    o.anisotropic = (k & PTB_MAT_ANISOTROPIC) ? m.anisotropic : R(0);
    o.metallic = (k & PTB_MAT_METALLIC) ? m.metallic : R(0);
    o.roughness = (k & PTB_MAT_ROUGHNESS) ? m.roughness : R(0.5);
    o.subsurface = (k & PTB_MAT_SUBSURFACE) ? m.subsurface : R(0);
    o.specular_tint = (k & PTB_MAT_SPECULAR_TINT) ? m.specular_tint : R(0);
    o.sheen = (k & PTB_MAT_SHEEN) ? m.sheen : R(0);
    o.sheen_tint = (k & PTB_MAT_SHEEN_TINT) ? m.sheen_tint : R(0);
    o.clearcoat = (k & PTB_MAT_CLEARCOAT) ? m.clearcoat : R(0);
    o.clearcoat_gloss = (k & PTB_MAT_CLEARCOAT_GLOSS) ? m.clearcoat_gloss : R(0);
    o.spec_trans = (k & PTB_MAT_SPEC_TRANS) ? m.spec_trans : R(0);
    o.ior = (k & PTB_MAT_IOR) ? m.ior : R(1.45);
    o.set_mask = k & PTB_MAT_ALL;
    o.albedo_kind = m.albedo_kind;
    o.checker_a = m.checker_a; o.checker_b = m.checker_b; o.checker_scale = m.checker_scale; o.checker_offset = m.checker_offset;
    // material.rs:15-22; finalize clamps the anisotropy (material.rs:126)
    o.med_type = m.medium_type;
    o.med_density = m.medium_density;
    for (int c = 0; c < 3; ++c) o.med_color[c] = m.medium_color[c];
    o.med_g = m.medium_anisotropy < R(-0.9) ? R(-0.9) : (m.medium_anisotropy > R(0.9) ? R(0.9) : m.medium_anisotropy);
    return o;
}

// BVH over n spheres (binned SAH, f32 bounds rounded outward); returns the depth of the tree
template <class CenterOf, class RadiusOf>
static uint32_t build_bvh(size_t n, CenterOf center_of, RadiusOf radius_of, std::vector<BvhNode>& nodes_out, std::vector<uint32_t>& prim_out) {
    BvhBuilder b;
    b.pbox.resize(n); b.pcen.resize(3 * n); b.prim.resize(n);
    for (size_t i = 0; i < n; ++i) {
        for (int k = 0; k < 3; ++k) {
            double c = center_of(i, k), r = std::fabs(radius_of(i));
            float lo = (float)(c - r), hi = (float)(c + r);
            lo = std::nextafterf(lo - std::fabs(lo) * 1e-6f, -3e38f);
            hi = std::nextafterf(hi + std::fabs(hi) * 1e-6f, 3e38f);
            b.pbox[i].lo[k] = lo; b.pbox[i].hi[k] = hi; b.pcen[3 * i + k] = (float)c;
        }
        b.prim[i] = (uint32_t)i;
    }
    b.nodes.reserve(2 * n + 1);
    b.nodes.push_back(BvhNode{});
    b.build_node(0, 0, (uint32_t)n);
    nodes_out.swap(b.nodes);
    prim_out.swap(b.prim);
    return b.max_depth;
}

template <class R> static int set_scene_impl(ptb_tracer* t, SceneBuffers<R>& sb, const typename PodTypes<R>::scene* sc) {
    if (!t || !sc) return fail(PTB_E_INVALID, "null tracer or scene");
    if ((sc->n_spheres && !sc->spheres) || (sc->n_planes && !sc->planes) || (sc->n_materials && !sc->materials) ||
        (sc->n_lights && !sc->lights))
        return fail(PTB_E_INVALID, "scene array pointer is NULL with a non-zero count");
    if (sc->depth > 65535u) return fail(PTB_E_INVALID, "depth must fit u16 (scene.rs:28)");
    for (uint32_t i = 0; i < sc->n_spheres; ++i)
        if (sc->spheres[i].material >= sc->n_materials) return fail(PTB_E_INVALID, "sphere %u: material index out of range", i);
    for (uint32_t i = 0; i < sc->n_planes; ++i)
        if (sc->planes[i].material >= sc->n_materials) return fail(PTB_E_INVALID, "plane %u: material index out of range", i);
    CU(cudaSetDevice(t->device));

    bool patch = false;
    for (uint32_t i = 0; i < sc->n_materials; ++i)
        if ((sc->materials[i].set_mask & PTB_MAT_ALL) != PTB_MAT_ALL) patch = true;
    const uint32_t thr = t->cfg.bvh_threshold ? t->cfg.bvh_threshold : 64u;
    bool use_bvh = !(sc->flags & PTB_SCENE_NO_BVH) && sc->n_spheres > 0 && (sc->n_spheres >= thr || (sc->flags & PTB_SCENE_FORCE_BVH));
    if (patch && use_bvh) return fail(PTB_E_INVALID, "BVH scenes need PTB_MAT_ALL on every material (order-independent)");
    if (patch && sc->n_spheres + sc->n_planes > 64u)
        return fail(PTB_E_INVALID, "partial material masks support at most 64 primitives");
    if (use_bvh && sc->n_spheres >= (1u << 27))         // node references pack (link << 3 | leaf count) into 32 bits; 2n - 1 nodes
        return fail(PTB_E_UNSUPPORTED, "sphere BVH: at most 2^27 - 1 spheres (%u given)", sc->n_spheres);

    std::vector<DSphere<R>> spheres(sc->n_spheres);
    std::vector<uint32_t> smat(sc->n_spheres);
    for (uint32_t i = 0; i < sc->n_spheres; ++i) {
        spheres[i] = DSphere<R>{sc->spheres[i].center[0], sc->spheres[i].center[1], sc->spheres[i].center[2], sc->spheres[i].radius};
        smat[i] = sc->spheres[i].material;
    }
    std::vector<DPlane<R>> planes(sc->n_planes);
    std::vector<uint32_t> pmat(sc->n_planes);
    for (uint32_t i = 0; i < sc->n_planes; ++i) {
        const auto& p = sc->planes[i];
        planes[i] = DPlane<R>{p.point[0], p.point[1], p.point[2], p.normal[0], p.normal[1], p.normal[2]};
        pmat[i] = p.material;
    }
    std::vector<DMaterial<R>> mats(sc->n_materials);
    for (uint32_t i = 0; i < sc->n_materials; ++i) mats[i] = resolve_pod_material<R>(sc->materials[i]);
    bool has_media = false;
    for (uint32_t i = 0; i < sc->n_materials; ++i) {
        if (mats[i].med_type == PTB_MEDIUM_NONE) continue;
        if (mats[i].med_type > PTB_MEDIUM_EMISSIVE) return fail(PTB_E_INVALID, "material %u: unknown medium type %u", i, mats[i].med_type);
        if (i + 1u > PTB_MEDIUM_MAX_INDEX) return fail(PTB_E_UNSUPPORTED, "materials with a medium need an index below %u (the path state keeps 7 bits for it)", PTB_MEDIUM_MAX_INDEX);
        has_media = true;
    }
    bool extended_lights = false;                   // scenes whose rectangular / distant lights are live: no resolved-material table
    if (sc->flags & PTB_SCENE_EXTENDED_LIGHTS)
        for (uint32_t i = 0; i < sc->n_lights; ++i) if (sc->lights[i].type != PTB_LIGHT_SPHERICAL) extended_lights = true;
    std::vector<DLight<R>> lights(sc->n_lights);
    const R PI_R = Const<R>::PI;
    for (uint32_t i = 0; i < sc->n_lights; ++i) {
        const auto& l = sc->lights[i];
        DLight<R>& o = lights[i];
        o.px = l.position[0]; o.py = l.position[1]; o.pz = l.position[2]; o.radius = l.radius;
        o.ex = l.emission[0]; o.ey = l.emission[1]; o.ez = l.emission[2];
        o.area = R(4) * PI_R * l.radius * l.radius;   // light.rs:22
        o.type = l.type; o.pad = 0;
        o.ux = l.u[0]; o.uy = l.u[1]; o.uz = l.u[2]; o.vx = l.v[0]; o.vy = l.v[1]; o.vz = l.v[2];
        if (l.type == PTB_LIGHT_RECTANGULAR) {          // |cross(u, v)|, in R with this operation order (the oracle's)
            const R cx = l.u[1] * l.v[2] - l.u[2] * l.v[1], cy = l.u[2] * l.v[0] - l.u[0] * l.v[2], cz = l.u[0] * l.v[1] - l.u[1] * l.v[0];
            o.area = std::sqrt(cx * cx + cy * cy + cz * cz);
        } else if (l.type == PTB_LIGHT_DISTANT) {
            o.area = R(0);                              // no MIS for distant lights (tracer.rs:158)
        }
    }

    // BVH over the spheres, and over the spherical lights when there are many (binned SAH, f32 bounds rounded outward)
    std::vector<BvhNode> nodes;
    std::vector<uint32_t> prim;
    uint32_t bvh_depth = 0, light_bvh_depth = 0;
    auto build_sphere_bvh = [&]() {
        bvh_depth = build_bvh(sc->n_spheres, [&](size_t i, int k) { return (double)sc->spheres[i].center[k]; },
                  [&](size_t i) { return (double)sc->spheres[i].radius; }, nodes, prim);
    };
    std::vector<BvhNode> lnodes;
    std::vector<uint32_t> lprim;
    std::vector<DSphere<R>> lleaf;
    {
        std::vector<uint32_t> sph_lights;          // indices of the spherical lights (others are never hit, scene.rs:69)
        for (uint32_t i = 0; i < sc->n_lights; ++i)
            if (sc->lights[i].type == PTB_LIGHT_SPHERICAL) sph_lights.push_back(i);
        // the light BVH lives in the BVH kernels only: scenes with partial material masks (no sphere BVH possible) keep the linear scan
        if (sph_lights.size() >= 16 && !patch && sc->n_spheres > 0 && !(sc->flags & PTB_SCENE_NO_BVH)) {
            use_bvh = true;
            light_bvh_depth = build_bvh(sph_lights.size(), [&](size_t i, int k) { return (double)sc->lights[sph_lights[i]].position[k]; },
                      [&](size_t i) { return (double)sc->lights[sph_lights[i]].radius; }, lnodes, lprim);
            lleaf.resize(lprim.size());
            for (size_t i = 0; i < lprim.size(); ++i) {
                const auto& l = sc->lights[sph_lights[lprim[i]]];
                lleaf[i] = DSphere<R>{l.position[0], l.position[1], l.position[2], l.radius};
                lprim[i] = sph_lights[lprim[i]];     // leaf order -> light index
            }
        }
    }
    uint64_t sphere_key = 1469598103934665603ull;          // FNV-1a over the sphere array
    {
        const unsigned char* b = reinterpret_cast<const unsigned char*>(spheres.data());
        for (size_t i = 0, n = spheres.size() * sizeof(DSphere<R>); i < n; ++i) { sphere_key ^= b[i]; sphere_key *= 1099511628211ull; }
    }
    const bool reuse_bvh = use_bvh && sb.bvh_cached && sb.bvh_key == sphere_key && sb.bvh_key_n == spheres.size() && sb.bvh && sb.bvh_prim && sb.bvh_spheres;
    if (use_bvh && !reuse_bvh) build_sphere_bvh();
    if (std::max(bvh_depth, light_bvh_depth) >= (uint32_t)BVH_STACK)
        return fail(PTB_E_INVALID, "internal: BVH depth %u exceeds the traversal stack (%d)", std::max(bvh_depth, light_bvh_depth), BVH_STACK);
    for (const auto* nv : {&nodes, &lnodes})
        for (const BvhNode& nd : *nv)
            if (nd.count > 7u) return fail(PTB_E_INVALID, "internal: BVH leaf with %u primitives (node references carry 3 count bits)", nd.count);

    // pack the arrays into one blob (16-byte aligned sections) so a CTA stages it with one loop
    std::vector<unsigned char> blob;
    auto append = [&](const void* src, size_t bytes) {
        size_t off = (blob.size() + 31) & ~size_t(31);
        blob.resize(off + bytes);
        if (bytes) memcpy(blob.data() + off, src, bytes);
        return (uint32_t)off;
    };
    DScene<R>& d = sb.d;
    d.off_planes = append(planes.data(), planes.size() * sizeof(DPlane<R>));
    d.off_lights = append(lights.data(), lights.size() * sizeof(DLight<R>));
    d.off_plane_material = append(pmat.data(), pmat.size() * sizeof(uint32_t));
    blob.resize((blob.size() + 31) & ~size_t(31));
    d.small_bytes = (uint32_t)blob.size();          // head section: always worth staging in shared memory
    d.off_spheres = append(spheres.data(), spheres.size() * sizeof(DSphere<R>));
    d.off_sphere_material = append(smat.data(), smat.size() * sizeof(uint32_t));
    d.off_materials = append(mats.data(), mats.size() * sizeof(DMaterial<R>));
    blob.resize((blob.size() + 31) & ~size_t(31));
    // resolved-material table: small f32 scenes whose whole blob (table included) fits the shared-memory scene copy
    d.off_rm_keys = d.off_rm_table = d.rm_entries = 0;
    d.n_sdf = 0;                                    // a new scene drops the signed-distance program (ptb_set_sdf_*)
    sb.rm_entries_built = 0;
    if constexpr (std::is_same<R, float>::value) {
        const uint32_t n_prims = sc->n_spheres + sc->n_planes;
        const char* off_env = getenv("PTB200_NO_RESOLVED_MATERIALS");     // A/B switch for profiling the generic shade path
        if (!use_bvh && !has_media && !extended_lights && sc->n_materials > 0 && n_prims > 0 && (!patch || n_prims <= RM_MAX_PATCH_PRIMS) && !(off_env && off_env[0] == '1')) {
            std::vector<uint32_t> keys;
            std::vector<RMat> table;
            auto add_key = [&](const std::vector<const DMaterial<float>*>& chain, const std::vector<uint32_t>& chain_index) {
                const int src = rm_checker_source(chain);
                keys.push_back((uint32_t)table.size() | (src >= 0 ? (chain_index[src] + 1u) << 16 : 0u));
                table.push_back(rm_resolve(chain, false));
                if (src >= 0) table.push_back(rm_resolve(chain, true));
            };
            if (patch) {
                keys.push_back(0u);                                           // the empty set is never looked up
                for (uint32_t mask = 1; mask < (1u << n_prims); ++mask) {
                    std::vector<const DMaterial<float>*> chain;
                    std::vector<uint32_t> idx;
                    for (uint32_t i = 0; i < n_prims; ++i)
                        if (mask & (1u << i)) { idx.push_back(i < sc->n_spheres ? smat[i] : pmat[i - sc->n_spheres]); chain.push_back(&mats[idx.back()]); }
                    add_key(chain, idx);
                }
            } else {
                for (uint32_t i = 0; i < sc->n_materials; ++i) add_key({&mats[i]}, {i});
            }
            const size_t total = ((blob.size() + 31) & ~size_t(31)) + ((keys.size() * 4 + 31) & ~size_t(31)) + table.size() * sizeof(RMat) + 32;
            if (table.size() < 0xffffu && total <= PTB_SMEM_SCENE_BYTES && total <= WF_SCENE_BYTES_RM) {
                d.off_rm_keys = append(keys.data(), keys.size() * sizeof(uint32_t));
                d.off_rm_table = append(table.data(), table.size() * sizeof(RMat));
                d.rm_entries = (uint32_t)table.size();
                sb.rm_entries_built = d.rm_entries;
                blob.resize((blob.size() + 31) & ~size_t(31));
            }
        }
    }

    sb.bytes = 0;
    CU(upload_vec(&sb.blob, sb.cap_blob, blob, t->stream, sb.bytes));
    if (reuse_bvh) {
        sb.bytes += sb.bvh_bytes;                       // still resident
    } else {
        const size_t before = sb.bytes;
        CU(upload_vec(&sb.bvh, sb.cap_bvh, nodes, t->stream, sb.bytes));
        CU(upload_vec(&sb.bvh_prim, sb.cap_prim, prim, t->stream, sb.bytes));
        std::vector<DSphere<R>> leaf_spheres(prim.size());
        for (size_t i = 0; i < prim.size(); ++i) leaf_spheres[i] = spheres[prim[i]];
        CU(upload_vec(&sb.bvh_spheres, sb.cap_spheres, leaf_spheres, t->stream, sb.bytes));
        CU(cudaStreamSynchronize(t->stream));           // leaf_spheres goes out of scope
        sb.bvh_bytes = sb.bytes - before;
        sb.bvh_cached = use_bvh; sb.bvh_key = sphere_key; sb.bvh_key_n = spheres.size();
    }
    CU(upload_vec(&sb.lbvh, sb.cap_lbvh, lnodes, t->stream, sb.bytes));
    CU(upload_vec(&sb.lbvh_prim, sb.cap_lprim, lprim, t->stream, sb.bytes));
    CU(upload_vec(&sb.lbvh_spheres, sb.cap_lspheres, lleaf, t->stream, sb.bytes));
    CU(cudaStreamSynchronize(t->stream));   // host vectors go out of scope

    const char* base = (const char*)sb.blob;
    d.blob = sb.blob; d.blob_bytes = (uint32_t)blob.size();
    d.n_spheres = sc->n_spheres; d.n_planes = sc->n_planes; d.n_materials = sc->n_materials; d.n_lights = sc->n_lights;
    d.spheres = (const DSphere<R>*)(base + d.off_spheres); d.sphere_material = (const uint32_t*)(base + d.off_sphere_material);
    d.planes = (const DPlane<R>*)(base + d.off_planes); d.plane_material = (const uint32_t*)(base + d.off_plane_material);
    d.materials = (const DMaterial<R>*)(base + d.off_materials); d.lights = (const DLight<R>*)(base + d.off_lights);
    d.bvh = use_bvh ? (const BvhNode*)sb.bvh : nullptr; d.bvh_prim = (const uint32_t*)sb.bvh_prim;
    d.bvh_spheres = (const DSphere<R>*)sb.bvh_spheres;
    d.light_bvh = lnodes.empty() ? nullptr : (const BvhNode*)sb.lbvh;
    d.light_bvh_spheres = (const DSphere<R>*)sb.lbvh_spheres;
    d.light_bvh_prim = (const uint32_t*)sb.lbvh_prim;
    for (int k = 0; k < 3; ++k) { d.light_lo[k] = 3e38f; d.light_hi[k] = -3e38f; }
    for (const auto& l : lights) {
        if (l.type != PTB_LIGHT_SPHERICAL) continue;
        const double c[3] = {(double)l.px, (double)l.py, (double)l.pz}, r = std::fabs((double)l.radius) * 1.0001 + 1e-6;
        for (int k = 0; k < 3; ++k) {
            d.light_lo[k] = std::min(d.light_lo[k], std::nextafterf((float)(c[k] - r), -3e38f));
            d.light_hi[k] = std::max(d.light_hi[k], std::nextafterf((float)(c[k] + r), 3e38f));
        }
    }
    d.n_rect_lights = 0;
    for (const auto& l : lights) if (l.type == PTB_LIGHT_RECTANGULAR) d.n_rect_lights++;
    d.use_bvh = use_bvh; d.patch_materials = patch;
    d.depth = sc->depth; d.flags = sc->flags; d.eps = sc->eps;
    d.n_lights_f = (R)sc->n_lights;
    d.has_emissive = 0;
    // small scenes: the primitives also go into the kernel parameter (the resolved-material kernel's EMB instantiation reads them there)
    d.emb = (sc->n_spheres <= PTB_EMB_SPHERES && sc->n_planes <= PTB_EMB_PLANES && sc->n_lights <= PTB_EMB_LIGHTS && !use_bvh) ? 1u : 0u;
#ifdef PTB_NO_EMB
    d.emb = 0u;
#endif
    std::memset(d.emb_spheres, 0, sizeof(d.emb_spheres)); std::memset(d.emb_planes, 0, sizeof(d.emb_planes)); std::memset(d.emb_lights, 0, sizeof(d.emb_lights));
    if (d.emb) {
        for (uint32_t i = 0; i < sc->n_spheres; ++i) d.emb_spheres[i] = spheres[i];
        for (uint32_t i = 0; i < sc->n_planes; ++i) d.emb_planes[i] = planes[i];
        for (uint32_t i = 0; i < sc->n_lights; ++i) d.emb_lights[i] = lights[i];
    }
    d.has_media = has_media ? 1u : 0u;
    d.has_fx = (has_media || extended_lights) ? 1u : 0u;          // (the resolved-material kernel is not built for them: no table above)
    for (const auto& m : mats)
        if (m.emission[0] != R(0) || m.emission[1] != R(0) || m.emission[2] != R(0)) d.has_emissive = 1;

    // the camera basis is derived by derive_camera() once the frame size is known
    for (int k = 0; k < 3; ++k) d.cam_origin[k] = sc->camera.origin[k];
    d.bg_kind = sc->background.kind;
    for (int k = 0; k < 3; ++k) { d.bg_a[k] = sc->background.colour_a[k]; d.bg_b[k] = sc->background.colour_b[k]; }
    d.bg_scale = sc->background.scale; d.bg_gamma = sc->background.gamma;
    return PTB_OK;
}

// camera/pinhole.rs:38-61, loop-invariant part, evaluated in R with the reference's operation
// order (x86-64 host code built without -mfma, so nothing is contracted).
template <class R> static void derive_camera(DScene<R>& d, const CamParams<R>& c, uint32_t W, uint32_t H) {
    R width = (R)W, height = (R)H;
    R ratio = width / height;
    R half_width = std::tan((c.fov * (Const<R>::PI / R(180))) * R(0.5));
    R half_height = half_width / ratio;
    R o[3] = {c.origin[0], c.origin[1], c.origin[2]};
    R wv[3] = {o[0] - c.center[0], o[1] - c.center[1], o[2] - c.center[2]};
    R len = std::sqrt(wv[0] * wv[0] + wv[1] * wv[1] + wv[2] * wv[2]);
    R w[3] = {wv[0] / len, wv[1] / len, wv[2] / len};
    const R up[3] = {0, 1, 0};
    R u[3] = {up[1] * w[2] - up[2] * w[1], up[2] * w[0] - up[0] * w[2], up[0] * w[1] - up[1] * w[0]};
    R v[3] = {w[1] * u[2] - w[2] * u[1], w[2] * u[0] - w[0] * u[2], w[0] * u[1] - w[1] * u[0]};
    for (int k = 0; k < 3; ++k) {
        R lower_left = ((o[k] - u[k] * half_width) - v[k] * half_height) - w[k];
        d.cam_base[k] = lower_left - o[k];
        d.cam_horizontal[k] = u[k] * (half_width * R(2));
        d.cam_vertical[k] = v[k] * (half_height * R(2));
        d.cam_origin[k] = o[k];
    }
}

static void refresh_camera(ptb_tracer* t) {
    if (!t->W || !t->H) return;
    if (t->precision == 4) derive_camera(t->s32.d, t->c32, t->W, t->H);
    if (t->precision == 8) derive_camera(t->s64.d, t->c64, t->W, t->H);
}

// ------------------------------------------------------------------------------------------------
extern "C" {

int ptb_abi_version(void) { return PTB_ABI_VERSION; }
const char* ptb_last_error(void) { return g_err.c_str(); }
int ptb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int ptb_create(const ptb_config* cfg, ptb_tracer** out) {
    if (!out) return fail(PTB_E_INVALID, "out is NULL");
    *out = nullptr;
    int n = ptb_device_count();
    if (n <= 0) return fail(PTB_E_NO_DEVICE, "no CUDA device visible; this library has no CPU fallback");
    ptb_config c{};
    if (cfg) c = *cfg;
    if (c.device < 0 || c.device >= n) return fail(PTB_E_NO_DEVICE, "device %d out of range (have %d)", c.device, n);
    if (c.integrator > PTB_INTEGRATOR_STREAM) return fail(PTB_E_INVALID, "unknown integrator %u", c.integrator);
    CU(cudaSetDevice(c.device));
    ptb_tracer* t = new ptb_tracer();
    t->cfg = c;
    t->device = c.device;
    cudaError_t e = cudaStreamCreateWithFlags(&t->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { delete t; return fail(PTB_E_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e)); }
    t->own_stream = true;
    cudaDeviceGetAttribute(&t->sm_count, cudaDevAttrMultiProcessorCount, c.device);
    cudaMalloc((void**)&t->work_counter, sizeof(unsigned int));
    cudaMalloc((void**)&t->counters, sizeof(DeviceCounters));
    cudaMemset(t->counters, 0, sizeof(DeviceCounters));
    cudaEventCreate(&t->ev0);
    cudaEventCreate(&t->ev1);
    e = cudaGetLastError();
    if (e != cudaSuccess) { ptb_destroy(t); return fail(PTB_E_CUDA, "tracer setup: %s", cudaGetErrorString(e)); }
    *out = t;
    return PTB_OK;
}

void ptb_destroy(ptb_tracer* t) {
    if (!t) return;
    cudaSetDevice(t->device);
    if (t->stream) cudaStreamSynchronize(t->stream);
    t->s32.release();
    t->s64.release();
    t->wf.release();
    t->st.release();
    if (t->peer_base) { if (t->peer_ipc) cudaIpcCloseMemHandle(t->peer_base); else if (t->peer_owner) cudaFree(t->peer_base); }
    if (t->dl_stream) { cudaStreamSynchronize(t->dl_stream); cudaStreamDestroy(t->dl_stream); }
    if (t->dl_ready) cudaEventDestroy(t->dl_ready);
    for (cudaEvent_t e : t->dl_free) if (e) cudaEventDestroy(e);
    for (void* p : t->dl_staging) if (p) cudaFree(p);
    if (t->own_accum && t->accum) cudaFree(t->accum);
    if (t->staging) cudaFree(t->staging);
    if (t->work_counter) cudaFree(t->work_counter);
    if (t->counters) cudaFree(t->counters);
    if (t->ev0) cudaEventDestroy(t->ev0);
    if (t->ev1) cudaEventDestroy(t->ev1);
    if (t->own_stream && t->stream) cudaStreamDestroy(t->stream);
    delete t;
}

int ptb_set_stream(ptb_tracer* t, void* cuda_stream) {
    if (!t) return fail(PTB_E_INVALID, "null tracer");
    CU(cudaSetDevice(t->device));
    CU(cudaStreamSynchronize(t->stream));
    if (t->own_stream && t->stream) cudaStreamDestroy(t->stream);
    t->stream = (cudaStream_t)cuda_stream;
    t->own_stream = false;
    return PTB_OK;
}

// A tracer may move between the two instantiations of `F` (lib.rs:5-6).  Everything sized in reals belongs to the old
// precision: the accumulators, the staging image, the frame count, the peer slots.  They are dropped (the next
// ptb_resize / ptb_render_frame allocates them anew); an accumulator the CALLER owns, or mapped peer slots, cannot be
// re-typed behind the caller's back, so the switch is refused there.
static int precision_change(ptb_tracer* t, int new_precision) {
    if (t->precision == 0 || t->precision == new_precision) return PTB_OK;
    if (t->accum && !t->own_accum)
        return fail(PTB_E_PRECISION, "an external accumulator of f%d values is bound; unbind it (ptb_bind_accumulator NULL) before switching precision",
                    t->precision * 8);
    if (t->peer_base) return fail(PTB_E_PRECISION, "peer slots of the f%d frame exist; ptb_peer_slots_close before switching precision", t->precision * 8);
    CU(cudaSetDevice(t->device));
    CU(cudaStreamSynchronize(t->stream));
    if (t->accum) cudaFree(t->accum);
    if (t->staging) cudaFree(t->staging);
    t->accum = nullptr; t->own_accum = false; t->accum_bytes = 0;
    t->staging = nullptr; t->staging_bytes = 0;
    if (t->dl_stream) cudaStreamSynchronize(t->dl_stream);
    for (int k = 0; k < 2; ++k) { if (t->dl_staging[k]) cudaFree(t->dl_staging[k]); t->dl_staging[k] = nullptr; t->dl_used[k] = false; }
    t->dl_staging_bytes = 0;
    t->W = t->H = 0; t->frames = 0; t->flush_dst = nullptr; t->timed = false;
    t->host_clean = nullptr;
    return PTB_OK;
}

int ptb_set_scene_f32(ptb_tracer* t, const ptb_scene_f32* sc) {
    if (!t || !sc) return fail(PTB_E_INVALID, "null tracer or scene");
    int r = precision_change(t, 4);
    if (r != PTB_OK) return r;
    r = set_scene_impl<float>(t, t->s32, sc);
    if (r != PTB_OK) return r;
    t->s64.release();
    t->precision = 4;
    t->fused_blocks_f32 = t->fused_blocks_f64 = 0;
    CamParams<float>& c = t->c32;
    for (int k = 0; k < 3; ++k) { c.origin[k] = sc->camera.origin[k]; c.center[k] = sc->camera.center[k]; }
    c.fov = sc->camera.fov;
    refresh_camera(t);
    return PTB_OK;
}
int ptb_set_scene_f64(ptb_tracer* t, const ptb_scene_f64* sc) {
    if (!t || !sc) return fail(PTB_E_INVALID, "null tracer or scene");
    int r = precision_change(t, 8);
    if (r != PTB_OK) return r;
    r = set_scene_impl<double>(t, t->s64, sc);
    if (r != PTB_OK) return r;
    t->s32.release();
    t->precision = 8;
    t->fused_blocks_f32 = t->fused_blocks_f64 = 0;
    CamParams<double>& c = t->c64;
    for (int k = 0; k < 3; ++k) { c.origin[k] = sc->camera.origin[k]; c.center[k] = sc->camera.center[k]; }
    c.fov = sc->camera.fov;
    refresh_camera(t);
    return PTB_OK;
}

}  // extern "C"
// Signed-distance program: validated (postfix stack discipline, material indices, parameters) and copied into the scene
// descriptor the kernels receive as their parameter.
template <class R> static int set_sdf_impl(ptb_tracer* t, SceneBuffers<R>& sb, const typename PodTypes<R>::sdf* sdf) {
    DScene<R>& d = sb.d;
    if (!sdf || sdf->n_nodes == 0) { d.n_sdf = 0; d.rm_entries = sb.rm_entries_built; return PTB_OK; }
    if (!sdf->nodes) return fail(PTB_E_INVALID, "sdf: nodes is NULL");
    if (sdf->n_nodes > PTB_SDF_MAX_NODES) return fail(PTB_E_INVALID, "sdf: %u nodes, at most %d", sdf->n_nodes, PTB_SDF_MAX_NODES);
    if (d.patch_materials) return fail(PTB_E_UNSUPPORTED, "sdf: scenes with a signed-distance program need PTB_MAT_ALL on every material");
    if (!(sdf->hit_eps > R(0)) || !(sdf->max_dist > R(0)) || !(sdf->normal_h > R(0)) || sdf->max_steps == 0)
        return fail(PTB_E_INVALID, "sdf: hit_eps, max_dist, normal_h and max_steps must be positive");
    int depth = 0;
    for (uint32_t i = 0; i < sdf->n_nodes; ++i) {
        const auto& n = sdf->nodes[i];
        if (n.op <= PTB_SDF_PLANE) {
            if (n.material >= d.n_materials) return fail(PTB_E_INVALID, "sdf node %u: material index out of range", i);
            if (++depth > PTB_SDF_MAX_STACK) return fail(PTB_E_INVALID, "sdf node %u: the program needs more than %d stack entries", i, PTB_SDF_MAX_STACK);
        } else if (n.op >= PTB_SDF_UNION && n.op <= PTB_SDF_INTERSECT) {
            if (depth < 2) return fail(PTB_E_INVALID, "sdf node %u: combinator with fewer than two operands on the stack", i);
            if (n.op == PTB_SDF_SMOOTH_UNION && !(n.a[0] > R(0))) return fail(PTB_E_INVALID, "sdf node %u: smooth union needs a positive blend radius", i);
            --depth;
        } else {
            return fail(PTB_E_INVALID, "sdf node %u: unknown op %u", i, n.op);
        }
    }
    if (depth != 1) return fail(PTB_E_INVALID, "sdf: the program leaves %d values on the stack, expected 1", depth);
    for (uint32_t i = 0; i < sdf->n_nodes; ++i) {
        const auto& n = sdf->nodes[i];
        DSdfNode<R>& o = d.sdf[i];
        o.op = n.op; o.material = n.material;
        for (int k = 0; k < 3; ++k) o.p[k] = n.p[k];
        for (int k = 0; k < 4; ++k) o.a[k] = n.a[k];
    }
    d.n_sdf = sdf->n_nodes; d.sdf_max_steps = sdf->max_steps;
    d.sdf_hit_eps = sdf->hit_eps; d.sdf_max_dist = sdf->max_dist; d.sdf_normal_h = sdf->normal_h;
    d.rm_entries = 0;            // the material at a hit on the body comes from the program: generic shade path
    (void)t;
    return PTB_OK;
}
extern "C" {
int ptb_set_sdf_f32(ptb_tracer* t, const ptb_sdf_f32* sdf) {
    if (!t) return fail(PTB_E_INVALID, "null tracer");
    if (t->precision != 4) return fail(t->precision ? PTB_E_PRECISION : PTB_E_NO_SCENE, "ptb_set_sdf_f32 needs an f32 scene (ptb_set_scene_f32 first)");
    return set_sdf_impl<float>(t, t->s32, sdf);
}
int ptb_set_sdf_f64(ptb_tracer* t, const ptb_sdf_f64* sdf) {
    if (!t) return fail(PTB_E_INVALID, "null tracer");
    if (t->precision != 8) return fail(t->precision ? PTB_E_PRECISION : PTB_E_NO_SCENE, "ptb_set_sdf_f64 needs an f64 scene (ptb_set_scene_f64 first)");
    return set_sdf_impl<double>(t, t->s64, sdf);
}

static int need_scene(ptb_tracer* t, int precision) {
    if (!t) return fail(PTB_E_INVALID, "null tracer");
    if (t->precision == 0) return fail(PTB_E_NO_SCENE, "no scene set (ptb_set_scene_f32/_f64)");
    if (precision && t->precision != precision) return fail(PTB_E_PRECISION, "tracer holds an f%d scene", t->precision * 8);
    return PTB_OK;
}

// Frames are handed out as 16x16-pixel tiles of 256 work items with 32-bit item indices (RenderArgs::n_items), the wavefront
// integrator appends up to grid x pool x 8 tail items, and its slots pack a pixel's column and row into 16 bits each.
static int check_frame_size(uint32_t width, uint32_t height) {
    if (!width || !height || (uint64_t)width * height > (1ull << 31)) return fail(PTB_E_INVALID, "bad frame size %ux%u", width, height);
    const uint64_t padded = (uint64_t)((width + 15u) / 16u) * ((height + 15u) / 16u) * 256u;
    if (padded > 0xffffffffull - (1ull << 26))
        return fail(PTB_E_INVALID, "frame %ux%u pads to %llu work items; the hand-out counter is 32 bits wide", width, height, (unsigned long long)padded);
    return PTB_OK;
}

static int ensure_staging(ptb_tracer* t, size_t bytes) {
    if (t->staging_bytes >= bytes) return PTB_OK;
    if (t->staging) cudaFree(t->staging);
    t->staging = nullptr; t->staging_bytes = 0;
    CU(cudaMalloc(&t->staging, bytes));
    t->staging_bytes = bytes;
    return PTB_OK;
}

int ptb_clear(ptb_tracer* t) {
    if (!t) return fail(PTB_E_INVALID, "null tracer");
    if (!t->accum) return fail(PTB_E_INVALID, "no frame allocated (ptb_resize)");
    CU(cudaSetDevice(t->device));
    CU(cudaMemsetAsync(t->accum, 0, t->accum_bytes, t->stream));
    t->frames = 0;
    t->host_clean = nullptr;
    return PTB_OK;
}

int ptb_resize(ptb_tracer* t, uint32_t width, uint32_t height) {
    int r = need_scene(t, 0);
    if (r) return r;
    if ((r = check_frame_size(width, height))) return r;
    CU(cudaSetDevice(t->device));
    CU(cudaStreamSynchronize(t->stream));
    t->flush_dst = nullptr;            // peer slots are sized for the old frame: ptb_peer_set_target must be called again
    t->host_clean = nullptr;
    if (t->own_accum && t->accum) cudaFree(t->accum);
    t->accum = nullptr; t->own_accum = false;
    t->accum_bytes = (size_t)width * height * 4 * real_size(t);
    CU(cudaMalloc(&t->accum, t->accum_bytes));
    t->own_accum = true;
    t->W = width; t->H = height;
    refresh_camera(t);
    return ptb_clear(t);
}

int ptb_bind_accumulator(ptb_tracer* t, void* device_ptr, uint32_t width, uint32_t height) {
    int r = need_scene(t, 0);
    if (r) return r;
    if (!device_ptr) return ptb_resize(t, width, height);
    if ((r = check_frame_size(width, height))) return r;
    CU(cudaSetDevice(t->device));
    CU(cudaStreamSynchronize(t->stream));
    t->flush_dst = nullptr;
    t->host_clean = nullptr;
    if (t->own_accum && t->accum) cudaFree(t->accum);
    t->accum = device_ptr; t->own_accum = false;
    t->accum_bytes = (size_t)width * height * 4 * real_size(t);
    t->W = width; t->H = height;
    t->frames = 0;
    refresh_camera(t);
    return PTB_OK;
}

}  // extern "C"
template <class R> static int upload_impl(ptb_tracer* t, const R* pixels, uint64_t frames) {
    if (!pixels) return fail(PTB_E_INVALID, "pixels is NULL");
    if (!t->accum) return fail(PTB_E_INVALID, "no frame allocated (ptb_resize)");
    CU(cudaSetDevice(t->device));
    int r = ensure_staging(t, t->accum_bytes);
    if (r) return r;
    CU(cudaMemcpyAsync(t->staging, pixels, t->accum_bytes, cudaMemcpyHostToDevice, t->stream));
    using V4 = typename Vec4T<R>::type;
    uint32_t n = t->W * t->H;
    k_unresolve<R><<<(n + 255) / 256, 256, 0, t->stream>>>((const V4*)t->staging, (V4*)t->accum, (R)frames, n);
    t->launches++;
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(t->stream));
    t->frames = frames;
    t->host_clean = nullptr;
    return PTB_OK;
}
extern "C" {
int ptb_upload_f32(ptb_tracer* t, const float* px, uint64_t frames) { int r = need_scene(t, 4); return r ? r : upload_impl<float>(t, px, frames); }
int ptb_upload_f64(ptb_tracer* t, const double* px, uint64_t frames) { int r = need_scene(t, 8); return r ? r : upload_impl<double>(t, px, frames); }

}  // extern "C"
template <class R> static int download_impl(ptb_tracer* t, R* pixels) {
    if (!pixels) return fail(PTB_E_INVALID, "pixels is NULL");
    if (!t->accum) return fail(PTB_E_INVALID, "no frame allocated (ptb_resize)");
    CU(cudaSetDevice(t->device));
    int r = ensure_staging(t, t->accum_bytes);
    if (r) return r;
    using V4 = typename Vec4T<R>::type;
    uint32_t n = t->W * t->H;
    k_resolve<R><<<(n + 255) / 256, 256, 0, t->stream>>>((const V4*)t->accum, (V4*)t->staging, n);
    t->launches++;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(pixels, t->staging, t->accum_bytes, cudaMemcpyDeviceToHost, t->stream));
    CU(cudaStreamSynchronize(t->stream));
    return PTB_OK;
}
// Asynchronous variant for callers that keep rendering: the mean image is resolved on the tracer's stream into one of two
// staging images and copied to the host on a SIDE stream behind an event, so the D2H of step k overlaps the trace of step k+1
// (the copy engines and the SMs are independent).  A staging image is reused only after its previous copy has completed.
template <class R> static int download_async_impl(ptb_tracer* t, R* pixels) {
    if (!pixels) return fail(PTB_E_INVALID, "pixels is NULL");
    if (!t->accum) return fail(PTB_E_INVALID, "no frame allocated (ptb_resize)");
    CU(cudaSetDevice(t->device));
    if (!t->dl_stream) {
        CU(cudaStreamCreateWithFlags(&t->dl_stream, cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&t->dl_ready, cudaEventDisableTiming));
        for (cudaEvent_t& e : t->dl_free) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    if (t->dl_staging_bytes < t->accum_bytes) {
        CU(cudaStreamSynchronize(t->dl_stream));
        for (int k = 0; k < 2; ++k) {
            if (t->dl_staging[k]) cudaFree(t->dl_staging[k]);
            t->dl_staging[k] = nullptr; t->dl_used[k] = false;
        }
        t->dl_staging_bytes = 0;
        for (int k = 0; k < 2; ++k) CU(cudaMalloc(&t->dl_staging[k], t->accum_bytes));
        t->dl_staging_bytes = t->accum_bytes;
    }
    const uint32_t k = t->dl_parity;
    t->dl_parity ^= 1u;
    if (t->dl_used[k]) CU(cudaStreamWaitEvent(t->stream, t->dl_free[k], 0));
    using V4 = typename Vec4T<R>::type;
    const uint32_t n = t->W * t->H;
    k_resolve<R><<<(n + 255) / 256, 256, 0, t->stream>>>((const V4*)t->accum, (V4*)t->dl_staging[k], n);
    t->launches++;
    CU(cudaGetLastError());
    CU(cudaEventRecord(t->dl_ready, t->stream));
    CU(cudaStreamWaitEvent(t->dl_stream, t->dl_ready, 0));
    CU(cudaMemcpyAsync(pixels, t->dl_staging[k], t->accum_bytes, cudaMemcpyDeviceToHost, t->dl_stream));
    CU(cudaEventRecord(t->dl_free[k], t->dl_stream));
    t->dl_used[k] = true;
    return PTB_OK;
}
extern "C" {
int ptb_download_async_f32(ptb_tracer* t, float* px) { int r = need_scene(t, 4); return r ? r : download_async_impl<float>(t, px); }
int ptb_download_async_f64(ptb_tracer* t, double* px) { int r = need_scene(t, 8); return r ? r : download_async_impl<double>(t, px); }
int ptb_wait_download(ptb_tracer* t) {
    if (!t) return fail(PTB_E_INVALID, "null tracer");
    if (!t->dl_stream) return PTB_OK;
    CU(cudaSetDevice(t->device));
    CU(cudaStreamSynchronize(t->dl_stream));
    return PTB_OK;
}
int ptb_download_f32(ptb_tracer* t, float* px) { int r = need_scene(t, 4); return r ? r : download_impl<float>(t, px); }
int ptb_download_f64(ptb_tracer* t, double* px) { int r = need_scene(t, 8); return r ? r : download_impl<double>(t, px); }

}  // extern "C"
template <class R> static int denoise_impl(ptb_tracer* t, uint32_t iterations, R sigma_color, R* pixels) {
    if (!pixels) return fail(PTB_E_INVALID, "pixels is NULL");
    if (!t->accum) return fail(PTB_E_INVALID, "no frame allocated (ptb_resize)");
    if (iterations == 0 || iterations > 8 || !(sigma_color > R(0))) return fail(PTB_E_INVALID, "denoise: 1..8 iterations, sigma_color > 0");
    CU(cudaSetDevice(t->device));
    int r = ensure_staging(t, 2 * t->accum_bytes);            // two mean images: ping and pong
    if (r) return r;
    using V4 = typename Vec4T<R>::type;
    const uint32_t n = t->W * t->H;
    V4* a = (V4*)t->staging;
    V4* b = a + n;
    k_resolve<R><<<(n + 255) / 256, 256, 0, t->stream>>>((const V4*)t->accum, a, n);
    t->launches++;
    const dim3 blk(32, 8), grd((t->W + 31) / 32, (t->H + 7) / 8);
    R sigma = sigma_color;
    for (uint32_t i = 0; i < iterations; ++i) {
        k_atrous<R><<<grd, blk, 0, t->stream>>>(a, b, t->W, t->H, 1 << i, R(1) / (sigma * sigma));
        t->launches++;
        std::swap(a, b);
        sigma *= R(0.5);                                        // the range weight tightens as the footprint grows
    }
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(pixels, a, t->accum_bytes, cudaMemcpyDeviceToHost, t->stream));
    CU(cudaStreamSynchronize(t->stream));
    return PTB_OK;
}
extern "C" {
int ptb_denoise_f32(ptb_tracer* t, uint32_t iterations, float sigma_color, float* px) { int r = need_scene(t, 4); return r ? r : denoise_impl<float>(t, iterations, sigma_color, px); }
int ptb_denoise_f64(ptb_tracer* t, uint32_t iterations, double sigma_color, double* px) { int r = need_scene(t, 8); return r ? r : denoise_impl<double>(t, iterations, sigma_color, px); }

int ptb_frames(ptb_tracer* t, uint64_t* frames) {
    if (!t || !frames) return fail(PTB_E_INVALID, "null argument");
    *frames = t->frames;
    return PTB_OK;
}

}  // extern "C"
template <class R> static int render_fused(ptb_tracer* t, DScene<R>& d, uint32_t spp, uint64_t sample_base) {
    RenderArgs a{};
    a.accum = t->accum; a.flush_dst = t->flush_dst; a.W = t->W; a.H = t->H; a.spp = spp; a.sample_base = sample_base; a.seed = t->cfg.seed;
    a.rr_start = t->cfg.rr_start;
    a.tiles_x = (t->W + 15u) / 16u;
    a.n_items = a.tiles_x * ((t->H + 15u) / 16u) * 256u;
    a.work_counter = t->work_counter;
    a.counters = t->counters;
    int& blocks = sizeof(R) == 4 ? t->fused_blocks_f32 : t->fused_blocks_f64;
    const bool count = t->cfg.collect_counters != 0;
    const bool fx = d.has_fx != 0 || d.n_sdf != 0;
    void (*kern)(const DScene<R>, const RenderArgs) =
        fx ? (d.use_bvh ? (count ? k_render_fused<R, true, true, true> : k_render_fused<R, false, true, true>)
                        : (count ? k_render_fused<R, true, false, true> : k_render_fused<R, false, false, true>))
           : (d.use_bvh ? (count ? k_render_fused<R, true, true, false> : k_render_fused<R, false, true, false>)
                        : (count ? k_render_fused<R, true, false, false> : k_render_fused<R, false, false, false>));
    static thread_local const void* last_kern = nullptr;            // (the cached grid belongs to one instantiation)
    if (last_kern != (const void*)kern) { blocks = 0; last_kern = (const void*)kern; }
    if (blocks == 0) {
        int per_sm = 0;
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, FUSED_THREADS, 0));
        if (per_sm < 1) per_sm = 1;
        blocks = per_sm * t->sm_count;
    }
    uint32_t max_useful = (a.n_items + FUSED_THREADS - 1) / FUSED_THREADS;
    int grid = std::max(1, std::min<int>(blocks, (int)max_useful));
    CU(cudaMemsetAsync(t->work_counter, 0, sizeof(unsigned int), t->stream));
    CU(cudaEventRecord(t->ev0, t->stream));
    kern<<<grid, FUSED_THREADS, 0, t->stream>>>(d, a);
    CU(cudaGetLastError());
    CU(cudaEventRecord(t->ev1, t->stream));
    t->timed = true;
    t->launches++;
    return PTB_OK;
}

extern "C" {
int ptb_render(ptb_tracer* t, uint32_t spp, uint64_t sample_base) {
    int r = need_scene(t, 0);
    if (r) return r;
    if (!t->accum) return fail(PTB_E_INVALID, "no frame allocated (ptb_resize)");
    CU(cudaSetDevice(t->device));
    if (t->flush_dst && t->peer_frame_bytes != (size_t)t->W * t->H * 16)
        return fail(PTB_E_INVALID, "frame size changed since the peer slots were created");
    if (spp == 0) {
        // nothing to trace; with a peer target the slot must still read as "no samples" for this step
        if (t->flush_dst) CU(cudaMemsetAsync(t->flush_dst, 0, t->peer_frame_bytes, t->stream));
        return PTB_OK;
    }
    // AUTO (measured, profiles/): f64 -> fused; f32 small scenes -> the shared-memory wavefront (queues on the SM, no HBM
    // traffic); f32 BVH scenes from 4096 spheres -> the global-memory wavefront, whose dedicated traversal kernels keep three
    // times as many warps resident on the dependent node fetches (100k spheres + 64 lights: 482 fused / 265 shared-memory
    // wavefront / 690 streaming Msamples/s)
    uint32_t integ = t->cfg.integrator;
    if (integ == PTB_INTEGRATOR_AUTO) {
        if (t->precision != 4) integ = PTB_INTEGRATOR_WAVEFRONT;           // f64: 1215 vs 696 Msamples/s (fused) at 4K on the demo scene
        else integ = (t->s32.d.use_bvh && t->s32.d.n_spheres >= 4096u) ? PTB_INTEGRATOR_STREAM : PTB_INTEGRATOR_WAVEFRONT;
        // the wavefront slots pack (column, row) into 16 bits each: wider or taller frames go to the fused integrator
        if (integ == PTB_INTEGRATOR_WAVEFRONT && (t->W > 65535u || t->H > 65535u)) integ = PTB_INTEGRATOR_FUSED;
        // signed-distance programs are compiled into the fused kernels and the generic (non-BVH) shared-memory wavefront kernel only
        if ((t->precision == 4 ? t->s32.d.n_sdf != 0 : t->s64.d.n_sdf != 0) &&
            (integ == PTB_INTEGRATOR_STREAM || (t->precision == 4 ? t->s32.d.use_bvh != 0 : t->s64.d.use_bvh != 0))) integ = PTB_INTEGRATOR_FUSED;
    }
    const bool has_sdf = t->precision == 4 ? t->s32.d.n_sdf != 0 : t->s64.d.n_sdf != 0;
    const bool bvh_scene = t->precision == 4 ? t->s32.d.use_bvh != 0 : t->s64.d.use_bvh != 0;
    if (has_sdf && (integ == PTB_INTEGRATOR_STREAM || (integ == PTB_INTEGRATOR_WAVEFRONT && bvh_scene)))
        return fail(PTB_E_UNSUPPORTED, "scenes with a signed-distance program run on the fused integrator or, without a sphere BVH, on the shared-memory wavefront");
    if (integ == PTB_INTEGRATOR_WAVEFRONT) {
        if (t->W > 65535u || t->H > 65535u)
            return fail(PTB_E_UNSUPPORTED, "the wavefront integrator packs pixel coordinates into 16 bits: frame %ux%u has a side over 65535", t->W, t->H);
        if (t->precision == 4)
            r = wavefront_render<float>(t->wf, t->s32.d, t->accum, t->flush_dst, t->W, t->H, spp, sample_base, t->cfg, t->stream, t->sm_count, t->counters,
                                        t->work_counter, t->ev0, t->ev1, &t->launches, g_err);
        else
            r = wavefront_render<double>(t->wf, t->s64.d, t->accum, t->flush_dst, t->W, t->H, spp, sample_base, t->cfg, t->stream, t->sm_count, t->counters,
                                         t->work_counter, t->ev0, t->ev1, &t->launches, g_err);
        if (r) return r;
        t->timed = true;
    } else if (integ == PTB_INTEGRATOR_STREAM) {
        if (t->precision != 4) return fail(PTB_E_UNSUPPORTED, "the streaming wavefront integrator is built for f32 only");
        r = stream_render(t->st, t->s32.d, t->accum, t->flush_dst, t->W, t->H, spp, sample_base, t->cfg, t->stream, t->sm_count, t->counters, t->ev0, t->ev1,
                          &t->launches, g_err);
        if (r) return r;
        t->timed = true;
    } else {
        r = t->precision == 4 ? render_fused<float>(t, t->s32.d, spp, sample_base) : render_fused<double>(t, t->s64.d, spp, sample_base);
        if (r) return r;
    }
    t->last_integrator = integ;
    t->last_kernel_bits = t->precision == 8 ? (PTB_KERNEL_F64 | (t->s64.d.use_bvh ? PTB_KERNEL_BVH : 0u))
                          : ((t->s32.d.use_bvh ? PTB_KERNEL_BVH : 0u) |
                             (integ == PTB_INTEGRATOR_WAVEFRONT && t->s32.d.rm_entries != 0 && !t->s32.d.use_bvh ? PTB_KERNEL_RM_TABLE : 0u) |
                             (integ == PTB_INTEGRATOR_STREAM && stream_uses_split(t->s32.d) ? PTB_KERNEL_SPLIT : 0u));
    t->frames += spp;
    t->host_clean = nullptr;
    return PTB_OK;
}

int ptb_last_integrator(ptb_tracer* t, uint32_t* integrator, uint32_t* kernel_bits) {
    if (!t || !integrator || !kernel_bits) return fail(PTB_E_INVALID, "null argument");
    if (!t->last_integrator) return fail(PTB_E_INVALID, "no render has been issued");
    *integrator = t->last_integrator; *kernel_bits = t->last_kernel_bits;
    return PTB_OK;
}

int ptb_synchronize(ptb_tracer* t) {
    if (!t) return fail(PTB_E_INVALID, "null tracer");
    CU(cudaSetDevice(t->device));
    CU(cudaStreamSynchronize(t->stream));
    return PTB_OK;
}

}  // extern "C"
template <class R> static int render_frame_impl(ptb_tracer* t, uint32_t w, uint32_t h, uint64_t frames_before, R* pixels, uint32_t flags) {
    if (!pixels) return fail(PTB_E_INVALID, "pixels is NULL");
    if (flags & ~(uint32_t)PTB_FRAME_HOST_UNCHANGED) return fail(PTB_E_INVALID, "unknown render_frame flags 0x%x", flags);
    int r;
    if (w != t->W || h != t->H || !t->accum) { if ((r = ptb_resize(t, w, h))) return r; }
    // the host buffer is the source of truth, exactly as in the reference where `pixels` and `frames`
    // are public fields the app may edit between calls — unless the caller vouches that it still holds what
    // this tracer wrote last (same buffer, same frame count, device image untouched since): then the upload is redundant
    const bool resident = (flags & PTB_FRAME_HOST_UNCHANGED) && t->host_clean == (const void*)pixels && t->host_clean_frames == frames_before &&
                          t->frames == frames_before;
    if (frames_before == 0) { if ((r = ptb_clear(t))) return r; }
    else if (!resident) { if ((r = upload_impl<R>(t, pixels, frames_before))) return r; }
    t->host_clean = nullptr;
    if ((r = ptb_render(t, 1, frames_before))) return r;
    if ((r = download_impl<R>(t, pixels))) return r;
    t->host_clean = pixels; t->host_clean_frames = t->frames;
    return PTB_OK;
}
extern "C" {
int ptb_render_frame_f32(ptb_tracer* t, uint32_t w, uint32_t h, uint64_t fb, float* px) { int r = need_scene(t, 4); return r ? r : render_frame_impl<float>(t, w, h, fb, px, 0); }
int ptb_render_frame_f64(ptb_tracer* t, uint32_t w, uint32_t h, uint64_t fb, double* px) { int r = need_scene(t, 8); return r ? r : render_frame_impl<double>(t, w, h, fb, px, 0); }
int ptb_render_frame_ex_f32(ptb_tracer* t, uint32_t w, uint32_t h, uint64_t fb, float* px, uint32_t flags) { int r = need_scene(t, 4); return r ? r : render_frame_impl<float>(t, w, h, fb, px, flags); }
int ptb_render_frame_ex_f64(ptb_tracer* t, uint32_t w, uint32_t h, uint64_t fb, double* px, uint32_t flags) { int r = need_scene(t, 8); return r ? r : render_frame_impl<double>(t, w, h, fb, px, flags); }

// Page-locking of caller-owned host memory is the CALLER's decision and lifetime (a registration cached inside the library
// would outlive a freed buffer and alias the next allocation at the same address).  Portable: valid for every device.
int ptb_pin_host(void* p, size_t bytes) {
    if (!p || !bytes) return fail(PTB_E_INVALID, "ptb_pin_host: null buffer");
    if (ptb_device_count() <= 0) return fail(PTB_E_NO_DEVICE, "no CUDA device visible");
    cudaError_t e = cudaHostRegister(p, bytes, cudaHostRegisterPortable);
    if (e == cudaErrorHostMemoryAlreadyRegistered) { cudaGetLastError(); return PTB_OK; }
    if (e != cudaSuccess) { cudaGetLastError(); return fail(PTB_E_CUDA, "cudaHostRegister: %s", cudaGetErrorString(e)); }
    return PTB_OK;
}
int ptb_unpin_host(void* p) {
    if (!p) return fail(PTB_E_INVALID, "ptb_unpin_host: null buffer");
    cudaError_t e = cudaHostUnregister(p);
    cudaGetLastError();            // never leave a sticky last-error for the next launch check
    if (e != cudaSuccess && e != cudaErrorHostMemoryNotRegistered) return fail(PTB_E_CUDA, "cudaHostUnregister: %s", cudaGetErrorString(e));
    return PTB_OK;
}

int ptb_convert_to_u8(ptb_tracer* t, uint8_t* rgba8) {
    int r = need_scene(t, 0);
    if (r) return r;
    if (!rgba8) return fail(PTB_E_INVALID, "rgba8 is NULL");
    if (!t->accum) return fail(PTB_E_INVALID, "no frame allocated (ptb_resize)");
    CU(cudaSetDevice(t->device));
    uint32_t n = t->W * t->H;
    if ((r = ensure_staging(t, (size_t)n * 4))) return r;
    if (t->precision == 4) k_convert_u8<float><<<(n + 255) / 256, 256, 0, t->stream>>>((const float4*)t->accum, (uchar4*)t->staging, n, 1);
    else k_convert_u8<double><<<(n + 255) / 256, 256, 0, t->stream>>>((const double4*)t->accum, (uchar4*)t->staging, n, 1);
    t->launches++;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(rgba8, t->staging, (size_t)n * 4, cudaMemcpyDeviceToHost, t->stream));
    CU(cudaStreamSynchronize(t->stream));
    return PTB_OK;
}

int ptb_convert_to_u8_at(ptb_tracer* t, uint8_t* frame, uint32_t x, uint32_t y, uint32_t fw, uint32_t fh) {
    int r = need_scene(t, 0);
    if (r) return r;
    if (!frame || !fw || !fh) return fail(PTB_E_INVALID, "bad frame");
    if (!t->accum) return fail(PTB_E_INVALID, "no frame allocated (ptb_resize)");
    CU(cudaSetDevice(t->device));
    size_t bytes = (size_t)fw * fh * 4;
    if ((r = ensure_staging(t, bytes))) return r;
    CU(cudaMemcpyAsync(t->staging, frame, bytes, cudaMemcpyHostToDevice, t->stream));
    uint32_t n = fw * fh;
    if (t->precision == 4) k_convert_u8_at<float><<<(n + 255) / 256, 256, 0, t->stream>>>((const float4*)t->accum, t->W, t->H, (uchar4*)t->staging, x, y, fw, fh, 1);
    else k_convert_u8_at<double><<<(n + 255) / 256, 256, 0, t->stream>>>((const double4*)t->accum, t->W, t->H, (uchar4*)t->staging, x, y, fw, fh, 1);
    t->launches++;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(frame, t->staging, bytes, cudaMemcpyDeviceToHost, t->stream));
    CU(cudaStreamSynchronize(t->stream));
    return PTB_OK;
}

}  // extern "C"
template <class R> static int convert_pixels_impl(ptb_tracer* t, size_t n, const R* rgba, uint8_t* out) {
    if (!t) return fail(PTB_E_INVALID, "null tracer");
    if (!rgba || !out) return fail(PTB_E_INVALID, "null buffer");
    if (n == 0) return PTB_OK;
    if (n > (1ull << 31)) return fail(PTB_E_INVALID, "too many pixels");
    CU(cudaSetDevice(t->device));
    using V4 = typename Vec4T<R>::type;
    void *din = nullptr, *dout = nullptr;
    CU(cudaMalloc(&din, n * sizeof(V4)));
    cudaError_t e = cudaMalloc(&dout, n * 4);
    if (e != cudaSuccess) { cudaFree(din); return fail(PTB_E_CUDA, "cudaMalloc: %s", cudaGetErrorString(e)); }
    cudaMemcpyAsync(din, rgba, n * sizeof(V4), cudaMemcpyHostToDevice, t->stream);
    k_convert_u8<R><<<(unsigned)((n + 255) / 256), 256, 0, t->stream>>>((const V4*)din, (uchar4*)dout, (uint32_t)n, 0);
    t->launches++;
    cudaMemcpyAsync(out, dout, n * 4, cudaMemcpyDeviceToHost, t->stream);
    e = cudaStreamSynchronize(t->stream);
    cudaFree(din); cudaFree(dout);
    if (e != cudaSuccess) return fail(PTB_E_CUDA, "convert_pixels: %s", cudaGetErrorString(e));
    return PTB_OK;
}
template <class R>
static int convert_pixels_at_impl(ptb_tracer* t, const R* rgba, uint32_t w, uint32_t h, uint8_t* frame, uint32_t x, uint32_t y, uint32_t fw,
                                  uint32_t fh) {
    if (!t) return fail(PTB_E_INVALID, "null tracer");
    if (!rgba || !frame || !w || !h || !fw || !fh) return fail(PTB_E_INVALID, "bad argument");
    CU(cudaSetDevice(t->device));
    using V4 = typename Vec4T<R>::type;
    size_t nin = (size_t)w * h, nf = (size_t)fw * fh;
    void *din = nullptr, *dfr = nullptr;
    CU(cudaMalloc(&din, nin * sizeof(V4)));
    cudaError_t e = cudaMalloc(&dfr, nf * 4);
    if (e != cudaSuccess) { cudaFree(din); return fail(PTB_E_CUDA, "cudaMalloc: %s", cudaGetErrorString(e)); }
    cudaMemcpyAsync(din, rgba, nin * sizeof(V4), cudaMemcpyHostToDevice, t->stream);
    cudaMemcpyAsync(dfr, frame, nf * 4, cudaMemcpyHostToDevice, t->stream);
    k_convert_u8_at<R><<<(unsigned)((nf + 255) / 256), 256, 0, t->stream>>>((const V4*)din, w, h, (uchar4*)dfr, x, y, fw, fh, 0);
    t->launches++;
    cudaMemcpyAsync(frame, dfr, nf * 4, cudaMemcpyDeviceToHost, t->stream);
    e = cudaStreamSynchronize(t->stream);
    cudaFree(din); cudaFree(dfr);
    if (e != cudaSuccess) return fail(PTB_E_CUDA, "convert_pixels_at: %s", cudaGetErrorString(e));
    return PTB_OK;
}
extern "C" {
int ptb_convert_pixels_to_u8_f32(ptb_tracer* t, size_t n, const float* rgba, uint8_t* out) { return convert_pixels_impl<float>(t, n, rgba, out); }
int ptb_convert_pixels_to_u8_f64(ptb_tracer* t, size_t n, const double* rgba, uint8_t* out) { return convert_pixels_impl<double>(t, n, rgba, out); }
int ptb_convert_pixels_to_u8_at_f32(ptb_tracer* t, const float* rgba, uint32_t w, uint32_t h, uint8_t* frame, uint32_t x, uint32_t y, uint32_t fw,
                                    uint32_t fh) { return convert_pixels_at_impl<float>(t, rgba, w, h, frame, x, y, fw, fh); }
int ptb_convert_pixels_to_u8_at_f64(ptb_tracer* t, const double* rgba, uint32_t w, uint32_t h, uint8_t* frame, uint32_t x, uint32_t y, uint32_t fw,
                                    uint32_t fh) { return convert_pixels_at_impl<double>(t, rgba, w, h, frame, x, y, fw, fh); }

int ptb_test_film_quotients_f32(uint32_t W, uint32_t H, uint32_t* all_exact, uint32_t* mismatches) {
    if (!all_exact || !mismatches) return fail(PTB_E_INVALID, "null argument");
    uint32_t bad = 0;
    if (W && H && W < (1u << 24) && H < (1u << 24)) {
        const float wf = (float)W, hf = (float)H, rw = 1.0f / wf, rh = 1.0f / hf;
        for (uint32_t x = 0; x < W; ++x) bad += div_by_fma((float)x, wf, rw) != (float)x / wf;
        for (uint32_t y = 1; y <= H; ++y) bad += div_by_fma((float)y, hf, rh) != (float)y / hf;
    }
    *mismatches = bad;
    *all_exact = film_coords_fma_exact(W, H) ? 1u : 0u;      // what wavefront_render decides
    return PTB_OK;
}
int ptb_test_bvh_build_f32(const ptb_sphere_f32* spheres, uint32_t n, uint32_t* depth_out, uint32_t* nodes_out, uint32_t* max_leaf_out) {
    if (!spheres || !n || !depth_out || !nodes_out || !max_leaf_out) return fail(PTB_E_INVALID, "null argument");
    std::vector<BvhNode> nodes;
    std::vector<uint32_t> prim;
    *depth_out = build_bvh(n, [&](size_t i, int k) { return (double)spheres[i].center[k]; }, [&](size_t i) { return (double)spheres[i].radius; }, nodes, prim);
    *nodes_out = (uint32_t)nodes.size();
    uint32_t ml = 0, covered = 0;
    for (const BvhNode& nd : nodes) { ml = std::max(ml, nd.count); covered += nd.count; }
    *max_leaf_out = ml;
    if (covered != n) return fail(PTB_E_INVALID, "internal: BVH leaves cover %u of %u primitives", covered, n);
    return PTB_OK;
}
int ptb_test_resolved_material_f32(const ptb_scene_f32* sc, const uint32_t* chain, uint32_t chain_len, uint32_t checker_odd, float* out) {
    if (!sc || !chain || !out || chain_len == 0) return fail(PTB_E_INVALID, "null argument or empty chain");
    std::vector<DMaterial<float>> mats(chain_len);
    std::vector<const DMaterial<float>*> ptrs(chain_len);
    for (uint32_t i = 0; i < chain_len; ++i) {
        if (chain[i] >= sc->n_materials) return fail(PTB_E_INVALID, "chain[%u]: material index out of range", i);
        mats[i] = resolve_pod_material<float>(sc->materials[chain[i]]);
        ptrs[i] = &mats[i];
    }
    const RMat r = rm_resolve(ptrs, checker_odd != 0);
    const Mat<float>& m = r.m;
    const float v[PTB_RMAT_FLOATS] = {m.rgb.x, m.rgb.y, m.rgb.z, m.emission.x, m.emission.y, m.emission.z, m.anisotropic, m.metallic, m.roughness,
                                      m.subsurface, m.specular_tint, m.sheen, m.sheen_tint, m.clearcoat, m.clearcoat_gloss, m.spec_trans, m.ior,
                                      m.clearcoat_roughness, m.ax, m.ay, r.eta[0], r.eta[1], r.spec_col[0][0], r.spec_col[0][1], r.spec_col[0][2],
                                      r.spec_col[1][0], r.spec_col[1][1], r.spec_col[1][2], r.sheen_col[0], r.sheen_col[1], r.sheen_col[2],
                                      r.lum, r.wd0, r.wc0, (float)r.lobe_class};
    memcpy(out, v, sizeof(v));
    return PTB_OK;
}
int ptb_get_counters(ptb_tracer* t, ptb_counters* out) {
    if (!t || !out) return fail(PTB_E_INVALID, "null argument");
    CU(cudaSetDevice(t->device));
    CU(cudaStreamSynchronize(t->stream));
    static_assert(sizeof(DeviceCounters) == sizeof(ptb_counters), "counter layouts must match");
    CU(cudaMemcpy(out, t->counters, sizeof(ptb_counters), cudaMemcpyDeviceToHost));
    return PTB_OK;
}
int ptb_reset_counters(ptb_tracer* t) {
    if (!t) return fail(PTB_E_INVALID, "null tracer");
    CU(cudaSetDevice(t->device));
    CU(cudaMemsetAsync(t->counters, 0, sizeof(DeviceCounters), t->stream));
    return PTB_OK;
}
int ptb_launch_count(ptb_tracer* t, uint64_t* launches) {
    if (!t || !launches) return fail(PTB_E_INVALID, "null argument");
    *launches = t->launches;
    return PTB_OK;
}
int ptb_last_render_ms(ptb_tracer* t, float* ms) {
    if (!t || !ms) return fail(PTB_E_INVALID, "null argument");
    if (!t->timed) return fail(PTB_E_INVALID, "no render has been issued");
    CU(cudaSetDevice(t->device));
    CU(cudaEventSynchronize(t->ev1));
    CU(cudaEventElapsedTime(ms, t->ev0, t->ev1));
    return PTB_OK;
}

}  // extern "C"
// ---- multi-GPU gather over peer memory ------------------------------------------------------------
// A progressive render shards by sample: every rank traces its samples of the WHOLE frame.  Instead of
// accumulating locally and NCCL-reducing 132.7 MB per step afterwards (measured 1.1 ms per step, 4 % of an 8-GPU step), a
// rank's render kernel STORES each pixel's partial sum — once per pixel per launch, 16 bytes, fire-and-forget — straight
// into a slot buffer in the ROOT GPU's memory over NVLink, so the transfer is spread over the whole trace and costs nothing.
// What is left on the critical path per step is a stream-ordered barrier and k_peer_sum on the root (N x 132.7 MB of local
// HBM reads, fixed slot order => deterministic image).  Slots are double-buffered by step parity so that the root may sum
// step k while the other ranks already write step k+1.
extern "C" {
int ptb_peer_slots_create(ptb_tracer* t, uint32_t n_slots, uint8_t* handle_out) {
    int r = need_scene(t, 4);
    if (r) return r;
    if (!t->accum || !n_slots || !handle_out) return fail(PTB_E_INVALID, "ptb_peer_slots_create: no frame, no slots or NULL handle");
    if (t->peer_base) return fail(PTB_E_INVALID, "peer slots already exist");
    CU(cudaSetDevice(t->device));
    const size_t frame = (size_t)t->W * t->H * 16;
    CU(cudaMalloc(&t->peer_base, 2 * (size_t)n_slots * frame));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, t->peer_base);
    if (e != cudaSuccess) { cudaFree(t->peer_base); t->peer_base = nullptr; return fail(PTB_E_CUDA, "cudaIpcGetMemHandle: %s", cudaGetErrorString(e)); }
    static_assert(sizeof(cudaIpcMemHandle_t) == PTB_PEER_HANDLE_BYTES, "IPC handle size");
    memcpy(handle_out, &h, sizeof h);
    t->peer_owner = true; t->peer_ipc = false; t->peer_slots = n_slots; t->peer_frame_bytes = frame;
    return PTB_OK;
}
int ptb_peer_slots_open(ptb_tracer* t, const uint8_t* handle, uint32_t n_slots) {
    int r = need_scene(t, 4);
    if (r) return r;
    if (!t->accum || !n_slots || !handle) return fail(PTB_E_INVALID, "ptb_peer_slots_open: no frame, no slots or NULL handle");
    if (t->peer_base) return fail(PTB_E_INVALID, "peer slots already exist");
    CU(cudaSetDevice(t->device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof h);
    cudaError_t e = cudaIpcOpenMemHandle(&t->peer_base, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) { t->peer_base = nullptr; cudaGetLastError(); return fail(PTB_E_UNSUPPORTED, "cudaIpcOpenMemHandle: %s", cudaGetErrorString(e)); }
    t->peer_owner = false; t->peer_ipc = true; t->peer_slots = n_slots; t->peer_frame_bytes = (size_t)t->W * t->H * 16;
    return PTB_OK;
}
int ptb_peer_set_target(ptb_tracer* t, uint32_t slot, uint32_t parity) {
    if (!t) return fail(PTB_E_INVALID, "null tracer");
    if (slot == 0xffffffffu) { t->flush_dst = nullptr; return PTB_OK; }
    if (!t->peer_base) return fail(PTB_E_INVALID, "no peer slots (ptb_peer_slots_create / _open)");
    if (slot >= t->peer_slots || parity > 1u) return fail(PTB_E_INVALID, "peer slot %u / parity %u out of range", slot, parity);
    if (t->peer_frame_bytes != (size_t)t->W * t->H * 16) return fail(PTB_E_INVALID, "frame size changed since the peer slots were created");
    t->flush_dst = (char*)t->peer_base + ((size_t)parity * t->peer_slots + slot) * t->peer_frame_bytes;
    return PTB_OK;
}
int ptb_peer_sum(ptb_tracer* t, uint32_t parity) {
    int r = need_scene(t, 4);
    if (r) return r;
    if (!t->peer_base || !t->peer_owner) return fail(PTB_E_INVALID, "ptb_peer_sum runs on the rank that created the slots");
    if (parity > 1u || !t->accum) return fail(PTB_E_INVALID, "bad parity or no frame");
    CU(cudaSetDevice(t->device));
    const uint32_t n = t->W * t->H;
    const float4* slots = (const float4*)((char*)t->peer_base + (size_t)parity * t->peer_slots * t->peer_frame_bytes);
    k_peer_sum<<<(n + 255) / 256, 256, 0, t->stream>>>((float4*)t->accum, slots, t->peer_slots, n);
    t->launches++;
    t->host_clean = nullptr;
    CU(cudaGetLastError());
    return PTB_OK;
}
int ptb_peer_slots_close(ptb_tracer* t) {
    if (!t) return fail(PTB_E_INVALID, "null tracer");
    if (!t->peer_base) return PTB_OK;
    CU(cudaSetDevice(t->device));
    CU(cudaStreamSynchronize(t->stream));
    if (t->peer_ipc) cudaIpcCloseMemHandle(t->peer_base); else if (t->peer_owner) cudaFree(t->peer_base);
    t->peer_base = nullptr; t->flush_dst = nullptr; t->peer_slots = 0; t->peer_owner = t->peer_ipc = false;
    return PTB_OK;
}
}  // extern "C"

// ---- per-function parity entry points ----------------------------------------------------------
namespace {
struct Dev {   // device scratch holder for the test entry points
    std::vector<void*> ptrs;
    cudaStream_t st;
    explicit Dev(cudaStream_t s) : st(s) {}
    ~Dev() { for (void* p : ptrs) cudaFree(p); }
    template <class T> T* in(const T* host, size_t n) {
        void* d = nullptr;
        if (cudaMalloc(&d, std::max<size_t>(n, 1) * sizeof(T)) != cudaSuccess) return nullptr;
        ptrs.push_back(d);
        if (n) cudaMemcpyAsync(d, host, n * sizeof(T), cudaMemcpyHostToDevice, st);
        return (T*)d;
    }
    template <class T> T* out(size_t n) {
        void* d = nullptr;
        if (cudaMalloc(&d, std::max<size_t>(n, 1) * sizeof(T)) != cudaSuccess) return nullptr;
        ptrs.push_back(d);
        cudaMemsetAsync(d, 0, std::max<size_t>(n, 1) * sizeof(T), st);
        return (T*)d;
    }
    template <class T> void back(T* host, const T* dev, size_t n) {
        if (n) cudaMemcpyAsync(host, dev, n * sizeof(T), cudaMemcpyDeviceToHost, st);
    }
};
inline unsigned grid_for(size_t n) { return (unsigned)std::max<size_t>(1, (n + 127) / 128); }
}  // namespace

extern "C" {
#define TEST_BEGIN(need_sc)                                  \
    int r_ = (need_sc) ? need_scene(t, 4) : (t ? PTB_OK : fail(PTB_E_INVALID, "null tracer")); \
    if (r_) return r_;                                       \
    CU(cudaSetDevice(t->device));                            \
    Dev dv(t->stream);
#define TEST_END()                           \
    t->launches++;                           \
    CU(cudaGetLastError());                  \
    CU(cudaStreamSynchronize(t->stream));    \
    return PTB_OK;

int ptb_test_sphere_hit_f32(ptb_tracer* t, size_t n, const float* o, const float* d, const float* c, const float* r, float* t_out) {
    TEST_BEGIN(false)
    auto *po = dv.in(o, 3 * n), *pd = dv.in(d, 3 * n), *pc = dv.in(c, 3 * n), *pr = dv.in(r, n);
    auto* pt = dv.out<float>(n);
    k_test_sphere_hit<float><<<grid_for(n), 128, 0, t->stream>>>(n, po, pd, pc, pr, pt);
    dv.back(t_out, pt, n);
    TEST_END()
}
int ptb_test_plane_hit_f32(ptb_tracer* t, size_t n, const float* o, const float* d, const float* p, const float* nn, float* t_out) {
    TEST_BEGIN(false)
    auto *po = dv.in(o, 3 * n), *pd = dv.in(d, 3 * n), *pp = dv.in(p, 3 * n), *pn = dv.in(nn, 3 * n);
    auto* pt = dv.out<float>(n);
    k_test_plane_hit<float><<<grid_for(n), 128, 0, t->stream>>>(n, po, pd, pp, pn, pt);
    dv.back(t_out, pt, n);
    TEST_END()
}
int ptb_test_gen_ray_f32(ptb_tracer* t, size_t n, const float* p2, const float* off2, float w, float h, float* o_out, float* d_out) {
    TEST_BEGIN(true)
    DScene<float> sc = t->s32.d;
    derive_camera(sc, t->c32, (uint32_t)w, (uint32_t)h);
    auto *pp = dv.in(p2, 2 * n), *pf = dv.in(off2, 2 * n);
    auto *po = dv.out<float>(3 * n), *pd = dv.out<float>(3 * n);
    k_test_gen_ray<float><<<grid_for(n), 128, 0, t->stream>>>(sc, n, pp, pf, w, h, po, pd);
    dv.back(o_out, po, 3 * n); dv.back(d_out, pd, 3 * n);
    TEST_END()
}
int ptb_test_closest_hit_f32(ptb_tracer* t, size_t n, const float* o, const float* d, const float* hd_in, uint32_t* hit, uint32_t* em,
                             float* hd_out, float* nrm, uint32_t* mat, float* lpdf, float* lem) {
    TEST_BEGIN(true)
    auto *po = dv.in(o, 3 * n), *pd = dv.in(d, 3 * n), *ph = dv.in(hd_in, n);
    auto *qh = dv.out<uint32_t>(n), *qe = dv.out<uint32_t>(n), *qm = dv.out<uint32_t>(n);
    auto *qd = dv.out<float>(n), *qn = dv.out<float>(3 * n), *qp = dv.out<float>(n), *ql = dv.out<float>(3 * n);
    k_test_closest_hit<float><<<grid_for(n), 128, 0, t->stream>>>(t->s32.d, n, po, pd, ph, qh, qe, qd, qn, qm, qp, ql);
    dv.back(hit, qh, n); dv.back(em, qe, n); dv.back(mat, qm, n); dv.back(hd_out, qd, n); dv.back(nrm, qn, 3 * n);
    dv.back(lpdf, qp, n); dv.back(lem, ql, 3 * n);
    TEST_END()
}
int ptb_test_any_hit_f32(ptb_tracer* t, size_t n, const float* o, const float* d, const float* md, uint32_t* hit) {
    TEST_BEGIN(true)
    auto *po = dv.in(o, 3 * n), *pd = dv.in(d, 3 * n), *pm = dv.in(md, n);
    auto* qh = dv.out<uint32_t>(n);
    k_test_any_hit<float><<<grid_for(n), 128, 0, t->stream>>>(t->s32.d, n, po, pd, pm, qh);
    dv.back(hit, qh, n);
    TEST_END()
}
int ptb_test_sdf_eval_f32(ptb_tracer* t, size_t n, const float* q, float* dist, uint32_t* mat) {
    TEST_BEGIN(true)
    if (!t->s32.d.n_sdf) return fail(PTB_E_INVALID, "no signed-distance program set (ptb_set_sdf_f32)");
    auto* pq = dv.in(q, 3 * n);
    auto* qd = dv.out<float>(n);
    auto* qm = dv.out<uint32_t>(n);
    k_test_sdf_eval<float><<<grid_for(n), 128, 0, t->stream>>>(t->s32.d, n, pq, qd, qm);
    dv.back(dist, qd, n); dv.back(mat, qm, n);
    TEST_END()
}
int ptb_test_sdf_trace_f32(ptb_tracer* t, size_t n, const float* o, const float* d, const float* limit, float* t_out, float* nrm, uint32_t* mat) {
    TEST_BEGIN(true)
    if (!t->s32.d.n_sdf) return fail(PTB_E_INVALID, "no signed-distance program set (ptb_set_sdf_f32)");
    auto *po = dv.in(o, 3 * n), *pd = dv.in(d, 3 * n), *pl = dv.in(limit, n);
    auto *qt = dv.out<float>(n), *qn = dv.out<float>(3 * n);
    auto* qm = dv.out<uint32_t>(n);
    k_test_sdf_trace<float><<<grid_for(n), 128, 0, t->stream>>>(t->s32.d, n, po, pd, pl, qt, qn, qm);
    dv.back(t_out, qt, n); dv.back(nrm, qn, 3 * n); dv.back(mat, qm, n);
    TEST_END()
}
int ptb_test_background_f32(ptb_tracer* t, size_t n, const float* d, float* rgb) {
    TEST_BEGIN(true)
    auto* pd = dv.in(d, 3 * n);
    auto* q = dv.out<float>(3 * n);
    k_test_background<float><<<grid_for(n), 128, 0, t->stream>>>(t->s32.d, n, pd, q);
    dv.back(rgb, q, 3 * n);
    TEST_END()
}
int ptb_test_sample_light_f32(ptb_tracer* t, size_t n, uint32_t li, const float* pos, const float* r1, const float* r2, float* nrm, float* em,
                              float* dir, float* dist, float* pdf) {
    TEST_BEGIN(true)
    if (li >= t->s32.d.n_lights) return fail(PTB_E_INVALID, "light index out of range");
    auto *pp = dv.in(pos, 3 * n), *p1 = dv.in(r1, n), *p2 = dv.in(r2, n);
    auto *qn = dv.out<float>(3 * n), *qe = dv.out<float>(3 * n), *qd = dv.out<float>(3 * n), *qs = dv.out<float>(n), *qp = dv.out<float>(n);
    k_test_sample_light<float><<<grid_for(n), 128, 0, t->stream>>>(t->s32.d, n, li, pp, p1, p2, qn, qe, qd, qs, qp);
    dv.back(nrm, qn, 3 * n); dv.back(em, qe, 3 * n); dv.back(dir, qd, 3 * n); dv.back(dist, qs, n); dv.back(pdf, qp, n);
    TEST_END()
}
int ptb_test_finalize_f32(ptb_tracer* t, size_t n, uint32_t mi, const float* o, const float* d, const float* hd, const float* nrm, float* rough,
                          float* ccr, float* ax, float* ay, float* eta, float* ffn, float* fhp) {
    TEST_BEGIN(true)
    if (mi >= t->s32.d.n_materials) return fail(PTB_E_INVALID, "material index out of range");
    auto *po = dv.in(o, 3 * n), *pd = dv.in(d, 3 * n), *ph = dv.in(hd, n), *pn = dv.in(nrm, 3 * n);
    auto *q1 = dv.out<float>(n), *q2 = dv.out<float>(n), *q3 = dv.out<float>(n), *q4 = dv.out<float>(n), *q5 = dv.out<float>(n);
    auto *q6 = dv.out<float>(3 * n), *q7 = dv.out<float>(3 * n);
    k_test_finalize<float><<<grid_for(n), 128, 0, t->stream>>>(t->s32.d, n, mi, po, pd, ph, pn, q1, q2, q3, q4, q5, q6, q7);
    dv.back(rough, q1, n); dv.back(ccr, q2, n); dv.back(ax, q3, n); dv.back(ay, q4, n); dv.back(eta, q5, n);
    dv.back(ffn, q6, 3 * n); dv.back(fhp, q7, 3 * n);
    TEST_END()
}
int ptb_test_disney_eval_f32(ptb_tracer* t, size_t n, uint32_t mi, const float* eta, const float* v, const float* nrm, const float* l,
                             float* f_out, float* pdf_out) {
    TEST_BEGIN(true)
    if (mi >= t->s32.d.n_materials) return fail(PTB_E_INVALID, "material index out of range");
    auto *pe = dv.in(eta, n), *pv = dv.in(v, 3 * n), *pn = dv.in(nrm, 3 * n), *pl = dv.in(l, 3 * n);
    auto *qf = dv.out<float>(3 * n), *qp = dv.out<float>(n);
    k_test_disney_eval<float><<<grid_for(n), 128, 0, t->stream>>>(t->s32.d, n, mi, pe, pv, pn, pl, qf, qp);
    dv.back(f_out, qf, 3 * n); dv.back(pdf_out, qp, n);
    TEST_END()
}
int ptb_test_disney_sample_f32(ptb_tracer* t, size_t n, uint32_t mi, const float* eta, const float* v, const float* nrm, const float* lprev,
                               const float* r1, const float* r2, const float* coin, uint32_t* lobe, float* l_out, float* f_out, float* pdf_out) {
    TEST_BEGIN(true)
    if (mi >= t->s32.d.n_materials) return fail(PTB_E_INVALID, "material index out of range");
    auto *pe = dv.in(eta, n), *pv = dv.in(v, 3 * n), *pn = dv.in(nrm, 3 * n), *pl = dv.in(lprev, 3 * n);
    auto *p1 = dv.in(r1, n), *p2 = dv.in(r2, n), *p3 = dv.in(coin, n);
    auto* ql = dv.out<uint32_t>(n);
    auto *qd = dv.out<float>(3 * n), *qf = dv.out<float>(3 * n), *qp = dv.out<float>(n);
    k_test_disney_sample<float><<<grid_for(n), 128, 0, t->stream>>>(t->s32.d, n, mi, pe, pv, pn, pl, p1, p2, p3, ql, qd, qf, qp);
    dv.back(lobe, ql, n); dv.back(l_out, qd, 3 * n); dv.back(f_out, qf, 3 * n); dv.back(pdf_out, qp, n);
    TEST_END()
}
int ptb_test_rng_f32(ptb_tracer* t, size_t n, const uint32_t* pixel, const uint64_t* sample, uint32_t bounce, float* out8) {
    TEST_BEGIN(false)
    auto* pp = dv.in(pixel, n);
    auto* ps = dv.in((const unsigned long long*)sample, n);
    auto* q = dv.out<float>(8 * n);
    k_test_rng<float><<<grid_for(n), 128, 0, t->stream>>>(n, pp, ps, bounce, t->cfg.seed, q);
    dv.back(out8, q, 8 * n);
    TEST_END()
}
}  // extern "C"
