/*
 * ptb200.h — C ABI of the B200-native path-tracing loop.
 *
 * This is the drop-in boundary for ONE hot path of markusmoenig/rust-pathtracer:
 * `Tracer::render` (rust-pathtracer/src/tracer.rs:22-123) and everything it calls.  The
 * reference has no FFI at all (it is 100 % safe Rust); these entry points are
 * what a `ptb200-sys` crate binds with `extern "C"` so that the crate's public API
 * (`Tracer::new`, `Tracer::render`, `ColorBuffer`, `Scene`) keeps its shape while the per-pixel
 * loop runs as hand-written sm_100a CUDA.  INTEGRATION.md shows the Rust-side binding.
 *
 * Conventions
 *   - every function returns PTB_OK (0) or a negative PTB_E_* code; nothing throws or unwinds
 *     across the boundary; `ptb_last_error()` returns a thread-local message for the last failure;
 *   - the library owns all device memory, the caller owns all host memory; scene descriptions are
 *     copied during `ptb_set_scene_*`;
 *   - a `ptb_tracer` handle is bound to one CUDA device and is NOT thread-safe (the reference's
 *     `render(&mut self, &mut ColorBuffer)` has the same one-at-a-time contract, tracer.rs:22);
 *   - there is NO CPU fallback: without a CUDA device `ptb_create` fails with PTB_E_NO_DEVICE;
 *   - `_f32` / `_f64` suffixes are the two instantiations of the reference's `F` type alias
 *     (rust-pathtracer/src/lib.rs:5-6).
 */
#ifndef PTB200_H
#define PTB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PTB_ABI_VERSION 4   /* 4: ptb_light_* gained u, v; ptb_material_* gained medium_* (struct sizes changed) */

/* ---- status codes ------------------------------------------------------------------------ */
enum {
    PTB_OK            =  0,
    PTB_E_INVALID     = -1,  /* bad argument (NULL, zero size, index out of range) */
    PTB_E_NO_DEVICE   = -2,  /* no CUDA device / device index out of range        */
    PTB_E_CUDA        = -3,  /* a CUDA runtime call or kernel failed               */
    PTB_E_NO_SCENE    = -4,  /* render called before ptb_set_scene_*               */
    PTB_E_PRECISION   = -5,  /* _f32 call on a tracer holding an f64 scene or v.v. */
    PTB_E_UNSUPPORTED = -6   /* feature not built                                  */
};

/* ---- enumerations ------------------------------------------------------------------------ */

/* How a material's `rgb` is obtained at a hit. */
enum {
    PTB_ALBEDO_CONSTANT = 0,
    /* renderer/src/analytical.rs:107-115 — checker evaluated on ray *direction ratios*
     * (d.x/d.y*scale+offset, d.z/d.y*scale+offset), rgb = (c,c,c) with c = checker_a when
     * ((floor(x)%2 + floor(y)%2) % 2 < 1) else checker_b.  Quirk A.7 of SURVEY.md. */
    PTB_ALBEDO_CHECKER_DIR_RATIO = 1
};

/* Which Material fields a primitive's closest_hit branch ASSIGNS (ptb_material.set_mask).
 * The reference resets `state.material = Material::new()` once per bounce (tracer.rs:63) and
 * each primitive branch of AnalyticalScene::closest_hit then assigns only a few fields
 * (analytical.rs:56-58, 82-85, 115-116) whenever that primitive is the closest SO FAR.  A ray
 * that passes the orange sphere and would also hit the metal sphere further away therefore
 * shades the orange sphere with `metallic = 1` left over from the metal branch.  To reproduce
 * this the export carries per-material assignment masks and the device applies them in
 * primitive order; fields never assigned keep the Material::new() defaults
 * (material.rs:82-114: rgb 1.5, roughness 0.5, ior 1.45, everything else 0).
 * Scenes that use the sphere BVH must use PTB_MAT_ALL (order-independent). */
enum {
    PTB_MAT_RGB = 1u << 0, PTB_MAT_EMISSION = 1u << 1, PTB_MAT_ANISOTROPIC = 1u << 2,
    PTB_MAT_METALLIC = 1u << 3, PTB_MAT_ROUGHNESS = 1u << 4, PTB_MAT_SUBSURFACE = 1u << 5,
    PTB_MAT_SPECULAR_TINT = 1u << 6, PTB_MAT_SHEEN = 1u << 7, PTB_MAT_SHEEN_TINT = 1u << 8,
    PTB_MAT_CLEARCOAT = 1u << 9, PTB_MAT_CLEARCOAT_GLOSS = 1u << 10, PTB_MAT_SPEC_TRANS = 1u << 11,
    PTB_MAT_IOR = 1u << 12, PTB_MAT_ALL = 0x1fffu
};

/* Light kinds, rust-pathtracer/src/globals.rs:69-73.  Only Spherical is implemented by the reference (tracer.rs:175-217,
 * scene.rs:69): by default the other two are inert exactly like there (they still count in number_of_lights()).  With
 * PTB_SCENE_EXTENDED_LIGHTS they get the semantics of the GLSL project the reference was ported from (the hooks are in place:
 * `Light.u / v / area`, globals.rs:76-84; the single-sided cull at tracer.rs:148; "no MIS for distant light", tracer.rs:158):
 *   RECTANGULAR  quad position + s*u + t*v, s, t in [0,1]; emits from the side cross(u, v) points to; sampled uniformly by area,
 *                pdf = dist^2 / (area * |n.d|); hit by rays (hidden from behind), pdf = t^2 / (area * cos) for the MIS weight
 *   DISTANT      direction normalize(position), emission constant, pdf 1, area 0 (no MIS), never hit by rays             */
enum { PTB_LIGHT_RECTANGULAR = 0, PTB_LIGHT_SPHERICAL = 1, PTB_LIGHT_DISTANT = 2 };

/* Medium kinds, rust-pathtracer/src/material.rs:7-13.  The reference declares the type, carries it in Material / State
 * (material.rs:75, globals.rs:19) and clamps its anisotropy (material.rs:126) but its tracer never reads it ("Support of
 * mediums / volumetric objects" is an open item, Readme.md:13; the hook is direct_light's unused `_is_surface`, tracer.rs:125).
 * A material with medium_type == NONE (the default) behaves exactly like the reference.  Otherwise the body behind the surface
 * is filled with the medium, with the semantics of the GLSL project the reference was ported from: a path is inside after it
 * leaves a surface of such a material towards its back side (dot(new direction, geometry normal) < 0) and outside again after
 * leaving one towards the front; no nesting.  While inside, at the next surface hit at distance t (before shading it):
 *   ABSORB    throughput *= exp(-(1 - color) * t * density)
 *   EMISSIVE  radiance   += color * t * density * throughput
 *   SCATTER   free-flight distance s = -ln(u) / density; if s < t the bounce happens in the medium instead of at the surface:
 *             throughput *= color, next-event estimation from the point with the Henyey-Greenstein phase function (value = pdf),
 *             new direction sampled from it (its pdf feeds the MIS weight of a light hit like a BSDF pdf would)                */
enum { PTB_MEDIUM_NONE = 0, PTB_MEDIUM_ABSORB = 1, PTB_MEDIUM_SCATTER = 2, PTB_MEDIUM_EMISSIVE = 3 };

/* Background kinds. */
enum {
    PTB_BG_CONSTANT = 0,     /* colour_a                                                           */
    /* renderer/src/analytical.rs:28-32: t = 0.5*(d.y+1); pow((1-t)*a + t*b, gamma) * scale         */
    PTB_BG_GRADIENT_Y = 1
};

/* Scene flags. */
enum {
    /* renderer/src/analytical.rs:130 — AnalyticalScene::any_hit ignores max_dist (quirk A.8).
     * Set for the exact demo-scene export; clear to honour the trait contract (scene.rs:15-16). */
    PTB_SCENE_ANYHIT_IGNORES_MAX_DIST = 1u << 0,
    /* Build / use the sphere BVH even for small sphere counts (otherwise: count >= bvh_threshold) */
    PTB_SCENE_FORCE_BVH               = 1u << 1,
    PTB_SCENE_NO_BVH                  = 1u << 2,
    /* Rectangular and distant lights are sampled / hit (see PTB_LIGHT_*) instead of being inert like in the reference.  A scene
     * that has such lights (or a material with a medium) runs on the generic kernels: no resolved-material table is built. */
    PTB_SCENE_EXTENDED_LIGHTS         = 1u << 3
};

/* Signed-distance program (ptb_set_sdf_*): instructions in postfix order.  Primitives push (distance, material), combinators
 * pop two entries (a below b) and push one.  No reference counterpart: the reference lists "Implement a SDF based example
 * scene" as open (Readme.md:18) — this is the device form of such a Scene impl, evaluated identically by the oracle. */
enum {
    PTB_SDF_SPHERE = 0,       /* |q - p| - a[0]                                                      */
    PTB_SDF_BOX = 1,          /* box at p, half extents a[0..2], corner radius a[3]                  */
    PTB_SDF_TORUS = 2,        /* ring of radius a[0] in the xz plane through p, tube radius a[1]     */
    PTB_SDF_PLANE = 3,        /* dot(q, a[0..2]) + a[3] (unit normal, offset); p unused              */
    PTB_SDF_UNION = 16,       /* min(a, b), material of the nearer                                   */
    PTB_SDF_SMOOTH_UNION = 17,/* polynomial smooth minimum with blend radius a[0]                    */
    PTB_SDF_SUBTRACT = 18,    /* max(a, -b), material of a                                           */
    PTB_SDF_INTERSECT = 19    /* max(a, b), material of the farther                                  */
};
#define PTB_SDF_MAX_NODES 16
#define PTB_SDF_MAX_STACK 8

/* Integrator selection (ptb_config.integrator). */
enum {
    PTB_INTEGRATOR_AUTO      = 0, /* shared-memory wavefront; streaming wavefront for large BVH scenes */
    PTB_INTEGRATOR_FUSED     = 1, /* persistent per-lane path loop with in-register regeneration   */
    PTB_INTEGRATOR_WAVEFRONT = 2, /* SoA path-state queues in shared memory, one stage per kind of
                                     work, queue sorted by lobe class between stages (f32 and f64)  */
    PTB_INTEGRATOR_STREAM    = 3  /* SoA ray / path-state queues in HBM, one KERNEL per kind of work
                                     (generate, closest_hit, shade per lobe class, any_hit,
                                     accumulate), ballot/popc compaction between them (f32 only)    */
};

/* ---- POD scene description, declared once per precision ----------------------------------- */

#define PTB_DECLARE_TYPES(SFX, REAL)                                                              \
    /* rust-pathtracer/src/material.rs:48-78, fields the tracer reads (tracer.rs:335-626).        \
     * Only fields named in set_mask are applied at a hit (see PTB_MAT_*);                         \
     * finalize() (material.rs:117-131) is applied on the device. */                               \
    typedef struct ptb_material_##SFX {                                                           \
        REAL rgb[3];                                                                              \
        REAL emission[3];                                                                         \
        REAL anisotropic, metallic, roughness, subsurface, specular_tint;                         \
        REAL sheen, sheen_tint, clearcoat, clearcoat_gloss, spec_trans, ior;                      \
        uint32_t set_mask;               /* PTB_MAT_* fields this material assigns            */  \
        uint32_t albedo_kind;            /* PTB_ALBEDO_*                                      */  \
        REAL checker_a, checker_b;       /* the two grey levels (0.25 / 0.1 in the demo)      */  \
        REAL checker_scale, checker_offset; /* 0.5 / 100 in the demo                          */  \
        /* material.rs:5-34 `Medium` (PTB_MEDIUM_*): what fills the body behind this surface  */  \
        uint32_t medium_type;                                                                     \
        REAL medium_density;                                                                      \
        REAL medium_color[3];                                                                     \
        REAL medium_anisotropy;          /* clamped to [-0.9, 0.9] (material.rs:126)          */  \
    } ptb_material_##SFX;                                                                         \
    /* analytical.rs:41-99,166-190 */                                                             \
    typedef struct ptb_sphere_##SFX {                                                             \
        REAL center[3];                                                                           \
        REAL radius;                                                                              \
        uint32_t material;                                                                        \
    } ptb_sphere_##SFX;                                                                           \
    /* analytical.rs:193-204: |dot(n,d)| > 1e-4, t = dot(point - o, n) / dot(n,d), t >= 0 */      \
    typedef struct ptb_plane_##SFX {                                                              \
        REAL point[3];                                                                            \
        REAL normal[3];                                                                           \
        uint32_t material;                                                                        \
    } ptb_plane_##SFX;                                                                            \
    /* light.rs:13-28 (area = 4*pi*r^2 is derived in F by the library) */                         \
    typedef struct ptb_light_##SFX {                                                              \
        REAL position[3];                                                                         \
        REAL radius;                                                                              \
        REAL emission[3];                                                                         \
        uint32_t type;                   /* PTB_LIGHT_*                                       */  \
        REAL u[3];                       /* RECTANGULAR: the two edges (globals.rs:80-81)     */  \
        REAL v[3];                                                                                \
    } ptb_light_##SFX;                                                                            \
    /* camera/pinhole.rs:5-25 (private fields origin/center/fov; fov in degrees, horizontal) */   \
    typedef struct ptb_camera_##SFX {                                                             \
        REAL origin[3];                                                                           \
        REAL center[3];                                                                           \
        REAL fov;                                                                                 \
    } ptb_camera_##SFX;                                                                           \
    typedef struct ptb_background_##SFX {                                                         \
        uint32_t kind;                   /* PTB_BG_*                                          */  \
        REAL colour_a[3];                                                                         \
        REAL colour_b[3];                                                                         \
        REAL scale;                                                                               \
        REAL gamma;                                                                               \
    } ptb_background_##SFX;                                                                       \
    typedef struct ptb_sdf_node_##SFX {                                                           \
        uint32_t op;                     /* PTB_SDF_*                                         */  \
        uint32_t material;               /* primitives: index into the scene's materials      */  \
        REAL p[3];                       /* primitives: position                              */  \
        REAL a[4];                       /* parameters, see PTB_SDF_*                         */  \
    } ptb_sdf_node_##SFX;                                                                         \
    typedef struct ptb_sdf_##SFX {                                                                \
        uint32_t n_nodes;                /* <= PTB_SDF_MAX_NODES; 0 removes the program       */  \
        const ptb_sdf_node_##SFX* nodes;                                                          \
        REAL hit_eps;                    /* |distance| below which the surface counts as hit  */  \
        REAL max_dist;                   /* tracing gives up beyond this distance             */  \
        REAL normal_h;                   /* offset of the four gradient samples               */  \
        uint32_t max_steps;                                                                       \
    } ptb_sdf_##SFX;                                                                              \
    /* What the new `Scene::device_export()` trait method returns. */             \
    typedef struct ptb_scene_##SFX {                                                              \
        uint32_t n_spheres, n_planes, n_materials, n_lights;                                      \
        const ptb_sphere_##SFX*   spheres;                                                        \
        const ptb_plane_##SFX*    planes;                                                         \
        const ptb_material_##SFX* materials;                                                      \
        const ptb_light_##SFX*    lights;                                                         \
        ptb_camera_##SFX          camera;                                                         \
        ptb_background_##SFX      background;                                                     \
        uint32_t depth;                  /* Scene::recursion_depth(), scene.rs:28-30 (4)      */  \
        uint32_t flags;                  /* PTB_SCENE_*                                       */  \
        REAL     eps;                    /* Tracer::eps, tracer.rs:16 (0.005)                 */  \
    } ptb_scene_##SFX;

PTB_DECLARE_TYPES(f32, float)
PTB_DECLARE_TYPES(f64, double)

/* ---- tracer configuration ---------------------------------------------------------------- */
typedef struct ptb_config {
    int32_t  device;          /* CUDA device ordinal                                              */
    uint32_t integrator;      /* PTB_INTEGRATOR_*                                                 */
    uint64_t seed;            /* base seed of the counter RNG (0 in all parity tests)             */
    uint32_t rr_start;        /* Russian-roulette start bounce; 0 = off (reference has none, A.12)*/
    uint32_t wave_paths;      /* STREAM integrator: paths carried per wave (0 = default 8 Mi)      */
    uint32_t bvh_threshold;   /* sphere count from which the BVH is used (0 = default 64)         */
    uint32_t collect_counters;/* 1 = count closest_hit/any_hit/lobe/ending events (slower)        */
} ptb_config;

/* Event counters (per tracer, cumulative since ptb_reset_counters). SURVEY.md Appendix C. */
typedef struct ptb_counters {
    uint64_t samples;
    uint64_t closest_hit;
    uint64_t any_hit;
    uint64_t shade;           /* bounces that reached direct_light + disney_sample                */
    uint64_t nee_contrib;     /* direct_light calls that added radiance                           */
    uint64_t eval_calls;      /* disney_eval calls                                                */
    uint64_t lobe_diffuse, lobe_clearcoat, lobe_reflect, lobe_refract;
    uint64_t end_sky, end_emitter, end_pdf, end_depth, end_rr;
    /* lobe evaluations, inside disney_eval and disney_sample together (tracer.rs:343-419) */
    uint64_t ev_diffuse, ev_clearcoat, ev_reflect, ev_refract;
    /* sphere-BVH work (0 for scenes without one): inner nodes visited — one visit tests both children — and leaf spheres tested */
    uint64_t bvh_nodes, bvh_leaf_tests;
} ptb_counters;

typedef struct ptb_tracer ptb_tracer;

/* ---- lifetime ------------------------------------------------------------------------------ */
/* replaces Tracer::new, tracer.rs:13-19 (the scene is supplied separately as data) */
int  ptb_create(const ptb_config* cfg, ptb_tracer** out);
void ptb_destroy(ptb_tracer* t);
int  ptb_abi_version(void);
const char* ptb_last_error(void);
/* number of visible CUDA devices (0 without a GPU; never fails) */
int  ptb_device_count(void);

/* Use an externally created CUDA stream (cudaStream_t as void*) for all work of this tracer. */
int  ptb_set_stream(ptb_tracer* t, void* cuda_stream);

/* ---- scene: replaces the dyn Scene callbacks (scene.rs:5-90) with exported data ------------ */
int  ptb_set_scene_f32(ptb_tracer* t, const ptb_scene_f32* scene);
int  ptb_set_scene_f64(ptb_tracer* t, const ptb_scene_f64* scene);

/* Attach a signed-distance program to the scene set last (copied; NULL or n_nodes == 0 removes it; ptb_set_scene_* removes it
 * too).  The body is tested after the planes in Scene::closest_hit order and by any_hit.  Requires materials that assign every
 * field (set_mask == PTB_MAT_ALL).  Scenes with a program shade through the generic path (no resolved-material table). */
int  ptb_set_sdf_f32(ptb_tracer* t, const ptb_sdf_f32* sdf);
int  ptb_set_sdf_f64(ptb_tracer* t, const ptb_sdf_f64* sdf);

/* ---- ColorBuffer on the device (buffer.rs:6-26) -------------------------------------------- */
/* (Re)allocate the accumulators for a w x h frame and clear them (ColorBuffer::new). */
int  ptb_resize(ptb_tracer* t, uint32_t width, uint32_t height);
/* Use caller-provided DEVICE memory (width*height*4 REALs: sum r,g,b and sample count) as the
 * accumulator, e.g. a torch tensor that is later NCCL-reduced.  NULL = library-owned again. */
int  ptb_bind_accumulator(ptb_tracer* t, void* device_ptr, uint32_t width, uint32_t height);
int  ptb_clear(ptb_tracer* t);
/* host ColorBuffer.pixels (running mean, RGBA, row 0 = top) + frames -> device accumulators */
int  ptb_upload_f32(ptb_tracer* t, const float*  pixels_rgba, uint64_t frames);
int  ptb_upload_f64(ptb_tracer* t, const double* pixels_rgba, uint64_t frames);
/* device accumulators -> mean image, alpha = 1 where any sample landed (tracer.rs:59,105) */
int  ptb_download_f32(ptb_tracer* t, float*  pixels_rgba);
int  ptb_download_f64(ptb_tracer* t, double* pixels_rgba);
/* The same download without blocking the caller: the mean image is resolved on the tracer's stream and copied to the host on a
 * side stream, so the copy of step k overlaps the tracing of step k+1 (bench.py's e2e arm; a progressive viewer does the same).
 * `pixels_rgba` must stay valid — and should be page-locked, ptb_pin_host — until ptb_wait_download returns.  Two downloads may be
 * in flight; a third waits for the first.  No reference counterpart: the reference's render() is synchronous (tracer.rs:22). */
int  ptb_download_async_f32(ptb_tracer* t, float*  pixels_rgba);
int  ptb_download_async_f64(ptb_tracer* t, double* pixels_rgba);
int  ptb_wait_download(ptb_tracer* t);
/* Page-lock / release caller-owned host memory (cudaHostRegister, portable) so that uploads and downloads run as DMA.  Opt-in and
 * owned by the caller: unpin BEFORE freeing the memory.  The host wrappers do this for ColorBuffer.pixels (pin on first render,
 * unpin in the buffer's destructor).  Already-pinned / not-pinned buffers are not an error. */
int  ptb_pin_host(void* host_ptr, size_t bytes);
int  ptb_unpin_host(void* host_ptr);
/* Denoised copy of the mean image (the accumulators are untouched): `iterations` levels (1..8) of an edge-avoiding a-trous
 * wavelet filter guided by the colour, range weight exp(-|dc|^2 / sigma^2) with sigma halving per level; NaN / inf pixels are
 * repaired from their finite neighbours.  No reference counterpart: the reference lists a denoiser as open (Readme.md:14). */
int  ptb_denoise_f32(ptb_tracer* t, uint32_t iterations, float sigma_color, float* pixels_rgba_out);
int  ptb_denoise_f64(ptb_tracer* t, uint32_t iterations, double sigma_color, double* pixels_rgba_out);
/* frames accumulated so far (ColorBuffer.frames) */
int  ptb_frames(ptb_tracer* t, uint64_t* frames);

/* ---- the hot path -------------------------------------------------------------------------- */
/* Add `spp` samples per pixel, with global sample indices [sample_base, sample_base + spp), to
 * the device accumulators.  Asynchronous on the tracer's stream.  Sample-split multi-GPU runs
 * give each rank a disjoint index range. */
int  ptb_render(ptb_tracer* t, uint32_t spp, uint64_t sample_base);
/* Drop-in for one `Tracer::render(&mut ColorBuffer)` call, tracer.rs:22-123: if frames_before
 * is 0 the accumulators are cleared, otherwise the host pixels are uploaded first (they are the
 * source of truth: `pixels` and `frames` are public fields the app may edit between calls);
 * then ONE sample per pixel is traced with sample index frames_before and the running mean is
 * written back to `pixels_rgba_inout`.  Synchronous.  Page-lock the buffer with ptb_pin_host for DMA-speed
 * copies (the host wrappers do).  Throughput-minded callers keep the image device-resident instead:
 * ptb_render + ptb_download. */
int  ptb_render_frame_f32(ptb_tracer* t, uint32_t width, uint32_t height, uint64_t frames_before,
                          float* pixels_rgba_inout);
int  ptb_render_frame_f64(ptb_tracer* t, uint32_t width, uint32_t height, uint64_t frames_before,
                          double* pixels_rgba_inout);
/* The same call with flags.  PTB_FRAME_HOST_UNCHANGED: the caller vouches that `pixels_rgba_inout` still holds exactly what this
 * tracer's previous render_frame wrote there (a wrapper knows this from a dirty flag on its ColorBuffer).  If in addition the
 * buffer address, the frame count and the device image are what that call left behind, the H2D upload — half of the PCIe traffic
 * of the drop-in loop (renderer/src/main.rs:113-122) — is skipped; otherwise the flag is ignored and the pixels are uploaded. */
enum { PTB_FRAME_HOST_UNCHANGED = 1u << 0 };
int  ptb_render_frame_ex_f32(ptb_tracer* t, uint32_t width, uint32_t height, uint64_t frames_before,
                             float* pixels_rgba_inout, uint32_t flags);
int  ptb_render_frame_ex_f64(ptb_tracer* t, uint32_t width, uint32_t height, uint64_t frames_before,
                             double* pixels_rgba_inout, uint32_t flags);
int  ptb_synchronize(ptb_tracer* t);

/* ColorBuffer::convert_to_u8, buffer.rs:55-64, of the current mean image (device kernel + D2H) */
int  ptb_convert_to_u8(ptb_tracer* t, uint8_t* rgba8);
/* ColorBuffer::convert_to_u8_at, buffer.rs:67-102: blit WITHOUT gamma into a frame of
 * frame_w x frame_h at (x, y) with the reference's strict `>` bounds. `frame_rgba8` is host
 * memory holding the existing frame contents (pixels outside the rectangle are kept). */
int  ptb_convert_to_u8_at(ptb_tracer* t, uint8_t* frame_rgba8, uint32_t x, uint32_t y,
                          uint32_t frame_w, uint32_t frame_h);

/* The same two conversions applied to HOST pixels (w*h RGBA reals in, bytes out): what
 * ColorBuffer::convert_to_u8 / convert_to_u8_at do to the public `pixels` field, whatever the
 * app stored there (any alpha).  H2D + kernel + D2H. */
int  ptb_convert_pixels_to_u8_f32(ptb_tracer* t, size_t n_pixels, const float* rgba, uint8_t* rgba8);
int  ptb_convert_pixels_to_u8_f64(ptb_tracer* t, size_t n_pixels, const double* rgba, uint8_t* rgba8);
int  ptb_convert_pixels_to_u8_at_f32(ptb_tracer* t, const float* rgba, uint32_t width, uint32_t height,
                                     uint8_t* frame_rgba8, uint32_t x, uint32_t y, uint32_t frame_w, uint32_t frame_h);
int  ptb_convert_pixels_to_u8_at_f64(ptb_tracer* t, const double* rgba, uint32_t width, uint32_t height,
                                     uint8_t* frame_rgba8, uint32_t x, uint32_t y, uint32_t frame_w, uint32_t frame_h);

/* ---- multi-GPU: sample-split partial sums gathered over peer memory ------------ */
/* The reference is single-process (rayon); the new build shards a progressive render by SAMPLE across the GPUs of one
 * box.  The baseline is one NCCL sum-reduce of the float4 accumulators per step (rust_pathtracer_b200/distributed.py).
 * These entry points fuse the transfer into the render kernel instead: a rank's kernel stores every pixel's partial sum
 * of a ptb_render call directly into a slot buffer in the ROOT GPU's memory (NVLink peer stores), and the root adds the
 * slots up in fixed order.  f32 scenes only; needs CUDA IPC + peer access between the GPUs (PTB_E_UNSUPPORTED otherwise —
 * callers then fall back to the NCCL reduce). */
#define PTB_PEER_HANDLE_BYTES 64
/* root: allocate 2 x n_slots frame-sized slot buffers (double-buffered by step parity) and export them (cudaIpcMemHandle_t) */
int  ptb_peer_slots_create(ptb_tracer* t, uint32_t n_slots, uint8_t* handle_out /* [PTB_PEER_HANDLE_BYTES] */);
/* every other rank: map the root's slots */
int  ptb_peer_slots_open(ptb_tracer* t, const uint8_t* handle /* [PTB_PEER_HANDLE_BYTES] */, uint32_t n_slots);
/* ptb_render now STORES its per-pixel partial sums into slot `slot` of buffer `parity` (0/1) instead of adding them to
 * the local accumulators; slot 0xffffffff restores local accumulation */
int  ptb_peer_set_target(ptb_tracer* t, uint32_t slot, uint32_t parity);
/* root, after all ranks' renders of the step have completed: accumulators += slot 0 + slot 1 + ... of buffer `parity` */
int  ptb_peer_sum(ptb_tracer* t, uint32_t parity);
int  ptb_peer_slots_close(ptb_tracer* t);

int  ptb_get_counters(ptb_tracer* t, ptb_counters* out);
int  ptb_reset_counters(ptb_tracer* t);
/* kernels launched by this tracer since creation (for bench.py's gpu_launches) */
int  ptb_launch_count(ptb_tracer* t, uint64_t* launches);
/* average device time (ms) of the render kernels launched by the last ptb_render call, measured
 * with CUDA events on the tracer's stream; synchronises. */
int  ptb_last_render_ms(ptb_tracer* t, float* ms);
/* which integrator the last ptb_render call actually ran (AUTO resolved: PTB_INTEGRATOR_FUSED / _WAVEFRONT / _STREAM) and
 * how it was instantiated: PTB_KERNEL_* bits.  Tests and bench.py name the measured kernel from this instead of guessing. */
#define PTB_KERNEL_BVH      1u   /* sphere BVH traversal instead of the linear scan                               */
#define PTB_KERNEL_RM_TABLE 2u   /* shades from the resolved-material table (small scenes, shared-memory wavefront) */
#define PTB_KERNEL_SPLIT    4u   /* streaming integrator: dedicated traversal kernels for bounce >= 1              */
#define PTB_KERNEL_F64      8u   /* the f64 instantiation of `F` (lib.rs:5-6)                                       */
int  ptb_last_integrator(ptb_tracer* t, uint32_t* integrator, uint32_t* kernel_bits);

/* ---- per-function parity entry points ------------------------------------------------------- */
/* Each runs the SAME __device__ functions the integrators call over n independent inputs that
 * live in HOST memory (SoA: component arrays are consecutive blocks of n), so that the 1e-5
 * per-function tests of SURVEY.md §7 go through this boundary.  The scene set on the tracer
 * supplies camera, lights and materials where needed.  Arrays named *_out are written. */

/* analytical.rs:166-190 — out_t[i] < 0 encodes None */
int  ptb_test_sphere_hit_f32(ptb_tracer* t, size_t n, const float* origin3, const float* dir3,
                             const float* center3, const float* radius, float* t_out);
/* analytical.rs:193-204 with the plane (point, normal) — out_t[i] < 0 encodes None */
int  ptb_test_plane_hit_f32(ptb_tracer* t, size_t n, const float* origin3, const float* dir3,
                            const float* point3, const float* normal3, float* t_out);
/* camera/pinhole.rs:38-61 — p2 = film position, offset2 = jitter; writes origin3/dir3 */
int  ptb_test_gen_ray_f32(ptb_tracer* t, size_t n, const float* p2, const float* offset2,
                          float width, float height, float* origin3_out, float* dir3_out);
/* Scene::closest_hit of the exported scene incl. sample_lights (analytical.rs:36-127,
 * scene.rs:36-86).  hit_dist_in is the stale State::hit_dist carried in (quirk A.1).
 * Outputs: hit flag (0/1), is_emitter (0/1), hit_dist, normal3, material index (or 0xffffffff),
 * light_pdf, light_emission3. */
int  ptb_test_closest_hit_f32(ptb_tracer* t, size_t n, const float* origin3, const float* dir3,
                              const float* hit_dist_in, uint32_t* hit_out, uint32_t* emitter_out,
                              float* hit_dist_out, float* normal3_out, uint32_t* material_out,
                              float* light_pdf_out, float* light_emission3_out);
/* Scene::any_hit of the exported scene (analytical.rs:130-145) */
int  ptb_test_any_hit_f32(ptb_tracer* t, size_t n, const float* origin3, const float* dir3,
                          const float* max_dist, uint32_t* hit_out);
/* the signed-distance program of ptb_set_sdf_f32: value and material at n points; sphere trace (t < 0 encodes a miss), gradient
 * normal at the hit and material along n rays */
int  ptb_test_sdf_eval_f32(ptb_tracer* t, size_t n, const float* point3, float* dist_out, uint32_t* material_out);
int  ptb_test_sdf_trace_f32(ptb_tracer* t, size_t n, const float* origin3, const float* dir3, const float* limit,
                            float* t_out, float* normal3_out, uint32_t* material_out);
/* Scene::background (analytical.rs:28-32) */
int  ptb_test_background_f32(ptb_tracer* t, size_t n, const float* dir3, float* rgb3_out);
/* Tracer::sample_light, tracer.rs:173-220, light `light_index` of the scene, draws (r1, r2) */
int  ptb_test_sample_light_f32(ptb_tracer* t, size_t n, uint32_t light_index, const float* pos3,
                               const float* r1, const float* r2, float* normal3_out,
                               float* emission3_out, float* direction3_out, float* dist_out,
                               float* pdf_out);
/* State::finalize + Material::finalize (globals.rs:50-62, material.rs:117-131) for material
 * `material_index`: outputs roughness, clearcoat_roughness, ax, ay, eta, ffnormal3, fhp3 */
int  ptb_test_finalize_f32(ptb_tracer* t, size_t n, uint32_t material_index, const float* origin3,
                           const float* dir3, const float* hit_dist, const float* normal3,
                           float* rough_out, float* ccrough_out, float* ax_out, float* ay_out,
                           float* eta_out, float* ffnormal3_out, float* fhp3_out);
/* HOST-ONLY: the wavefront integrator computes the film coordinates x / W and y / H of tracer.rs:34-46 (IEEE divisions in
 * the reference) as a correctly rounded reciprocal plus one FMA correction step, but only after checking on the host that this
 * is bit-identical to the IEEE quotient for EVERY column and row of the frame size; otherwise it divides.  This entry point
 * returns that verdict (1 = every quotient exact) and the number of columns / rows whose corrected quotient differs. */
int  ptb_test_film_quotients_f32(uint32_t width, uint32_t height, uint32_t* all_exact_out, uint32_t* mismatches_out);
/* HOST-ONLY: builds the sphere BVH the way ptb_set_scene_* does and reports its depth, node count and largest leaf.  The device
 * traversal keeps a fixed 40-entry stack; the builder falls back to median splits before a chain of SAH splits could exceed it
 * (tests/test_host.py feeds it adversarial size distributions). */
int  ptb_test_bvh_build_f32(const ptb_sphere_f32* spheres, uint32_t n_spheres, uint32_t* depth_out, uint32_t* nodes_out,
                            uint32_t* max_leaf_out);
/* HOST-ONLY (no device, no tracer): the entry of the resolved-material table that ptb_set_scene_f32 builds for small scenes
 *  for one accepted-primitive chain.  `chain` = material indices of the accepted primitives in test order
 * (one index for scenes whose materials assign every field); `checker_odd` = which checker cell supplies the albedo.
 * Replaces, per shaded bounce: the cumulative material assignment of closest_hit (analytical.rs:56-58, 82-85, 115-116),
 * Material::finalize (material.rs:117-131), the eta pick of State::finalize (globals.rs:58-61), get_spec_color
 * (tracer.rs:335-341) and the material-only lobe weights (tracer.rs:423, 426).
 * out[PTB_RMAT_FLOATS]: rgb3, emission3, anisotropic, metallic, roughness, subsurface, specular_tint, sheen, sheen_tint,
 * clearcoat, clearcoat_gloss, spec_trans, ior, clearcoat_roughness, ax, ay, eta_enter, eta_exit, spec_col_enter3,
 * spec_col_exit3, sheen_col3, luminance, diffuse_weight, clearcoat_weight, lobe_class (as float). */
#define PTB_RMAT_FLOATS 35
int  ptb_test_resolved_material_f32(const ptb_scene_f32* scene, const uint32_t* chain, uint32_t chain_len,
                                    uint32_t checker_odd, float* out);
/* Tracer::disney_eval, tracer.rs:555-626.  n3 = shading normal (ffnormal), v3 = -ray.direction,
 * l3 = light direction (world), eta per element.  Outputs f3 (already times |l.z|) and pdf. */
int  ptb_test_disney_eval_f32(ptb_tracer* t, size_t n, uint32_t material_index, const float* eta,
                              const float* v3, const float* n3, const float* l3, float* f3_out,
                              float* pdf_out);
/* Tracer::disney_sample, tracer.rs:441-553.  Draws (r1, r2, coin); lprev3 is the stale `l`
 * (quirk A.5).  Outputs lobe id (0 diffuse, 1 clearcoat, 2 reflect, 3 refract), l3 (world),
 * f3 (already times |n.l|) and pdf. */
int  ptb_test_disney_sample_f32(ptb_tracer* t, size_t n, uint32_t material_index, const float* eta,
                                const float* v3, const float* n3, const float* lprev3,
                                const float* r1, const float* r2, const float* coin,
                                uint32_t* lobe_out, float* l3_out, float* f3_out, float* pdf_out);
/* The counter RNG: Philox4x32-10 keyed (pixel, sample), SURVEY.md §8d "Seeds".  Writes the 8
 * slot values of `bounce` for n (pixel, sample) pairs: out[slot*n + i].  Bit-exact vs oracle. */
int  ptb_test_rng_f32(ptb_tracer* t, size_t n, const uint32_t* pixel, const uint64_t* sample,
                      uint32_t bounce, float* out8);

#ifdef __cplusplus
}
#endif
#endif /* PTB200_H */
