// ptb_wavefront.cuh — the wavefront integrator: SoA path-state queues, one stage per kind of work,
// queues compacted / sorted between stages, persistent CTAs — with the queues held in the SM's
// 227 KB of SHARED MEMORY instead of HBM.
//
// Why shared memory: the demo scene costs ~1.4 kFLOP and ~2 bounces per sample.
// A classic HBM wavefront streams ~1 KB of ray / path state per sample through the queues, which
// caps it at ~6 Gsamples/s on a 6.5 TB/s part before any arithmetic is done (SURVEY.md §7 "the
// roofline that actually binds").  One SM can hold 2048 paths x 96 B (or 2496 x 80 B) of state — 221 / 227 KB
// with the queue arrays and the scene copy —, enough for every stage to run with full warps, so the state
// never leaves the SM: HBM traffic stays at the 32 B per pixel of the accumulator read-modify-write.
//
// One CTA per SM (512 threads and 2048 slots; 832 threads and 2496 slots in the instantiation that shades from
// the resolved-material table, RMat in ptb_device.cuh) owns the pool and iterates over it; every iteration visits each
// slot once (the kernel's comment below describes the visit: pending event, regeneration, closest_hit) and ends with
// a counting sort of the slots by the key of what they hit — the LOBE CLASS of the material, WF_MISS, WF_REGEN — so
// that the next iteration's warps take 32-slot chunks of ONE key.  All warps of the SM run the same code at the same
// time, which is what keeps the instruction cache effective: a barrier-free variant with per-key rings
// (ptb_wavefront_async.cuh, A/B only) removes the 8 % of warp cycles spent at the iteration barrier and loses 17 %
// to instruction fetch (no_instruction 0.21 -> 2.6 stall cycles per issue; profiles/r02_ab_variants.txt).
//
// A slot owns one pixel for `spp` consecutive samples and sums them in sample order; the frame's last pixels (one
// per slot) are cut into sample blocks that k_tail_combine adds in block order (see wavefront_render).  Either way
// the image does not depend on scheduling: it is bit-reproducible run to run, like the fused integrator's.
#pragma once
#include <string>

#include "ptb_kernels.cuh"

namespace ptb {

// threads per CTA (one CTA per SM).  The generic and BVH instantiations need ~125 registers in the shade stage: 512 threads.
// With the resolved-material table the material is read from shared memory where it is used, and radiance / throughput /
// sample bookkeeping stay in shared memory while a path is shaded, so the kernel fits 72 registers with 24 bytes of spill:
// 832 threads = 26 warps.  Warps and pool go together: an iteration hands out pool / 32 chunks, and the last round of
// chunks leaves warps idle unless that is a multiple of the warp count — 26 warps x 3 rounds = 78 chunks = 2496 slots is
// what the 227 KB hold (measured at 4K, threads_pool: 512_2048 7622, 576_2304 8095, 640_2560 8454, 704_2112 8506,
// 768_2304 8489, 832_2496 8693, 896_2304 8492, 896_1792 8243, 1024_2304 8314 Msamples/s; profiles/r02_ab_variants.txt).
#ifndef PTB_WF_THREADS
#define PTB_WF_THREADS 512
#endif
#ifndef PTB_WF_THREADS_RM
#define PTB_WF_THREADS_RM 832
#endif
// f64 (`F = f64`, lib.rs:5-6): a slot is five 32-byte vectors, so the 227 KB hold 1024 of them = 32 chunks.  The shading code
// would like 255 registers, but warps pay more than spills cost: threads_pool 192_1152 / 256_1024 / 320_960 / 384_1152 /
// 512_1024 = 879 / 1029 / 1084 / 1166 / 1215 Msamples/s at 4K (128 registers, 2 rounds of 16 warps; the fused f64 kernel: 696)
#ifndef PTB_WF_THREADS_F64
#define PTB_WF_THREADS_F64 512
#endif
#ifndef PTB_WF_POOL_F64
#define PTB_WF_POOL_F64 1024
#endif
constexpr int WF_THREADS_F64 = PTB_WF_THREADS_F64;
constexpr uint32_t WF_POOL_F64 = PTB_WF_POOL_F64;
constexpr int WF_THREADS_GENERIC = PTB_WF_THREADS;
constexpr int WF_THREADS_RM = PTB_WF_THREADS_RM;
// shared-memory scene copy of the resolved-material instantiation (the host builds the table only if the blob fits)
#ifndef PTB_WF_SCENE_BYTES_RM
#define PTB_WF_SCENE_BYTES_RM PTB_SMEM_SCENE_BYTES
#endif
constexpr uint32_t WF_SCENE_BYTES_RM = PTB_WF_SCENE_BYTES_RM;
#ifndef PTB_WF_POOL
#define PTB_WF_POOL 2048
#endif
constexpr uint32_t WF_POOL_GENERIC = PTB_WF_POOL;       // path slots per CTA
// resolved-material instantiation: 2496 slots = 78 chunks = 3 rounds of 26 warps (see above)
#ifndef PTB_WF_POOL_RM
#define PTB_WF_POOL_RM 2496
#endif
constexpr uint32_t WF_POOL_RM = PTB_WF_POOL_RM;
#ifndef PTB_WF_REGEN_DEN
#define PTB_WF_REGEN_DEN 4u
#endif
constexpr uint32_t WF_MISS = 8;             // the path left the scene: background lookup, done with full warps in stage 2

// per-slot state: five 16-byte vectors (one LDS.128 / STS.128 each) + one packed queue word
//   ro = (o.xyz, State::hit_dist)        rd = (d.xyz, pixel column | row << 16)
//   tr = (throughput.xyz, flags)         ra = (radiance.xyz, sample index)
//   ac = (pixel sum.xyz, hit primitive / accepted set of the pending shading event)
// flags word: bit0 alive, bit1 have_pixel, bits 3..7 sample block (tail items), bits 8..23 bounce, bit 24 tail item,
//             bits 25..31 PathState::medium (0 = outside, else 1 + material index; scenes with media only)
constexpr uint32_t FL_ALIVE = 1u, FL_PIXEL = 2u, FL_BLOCK = 1u << 24, FL_BLOCK_BITS = FL_BLOCK | (31u << 3);
#ifndef PTB_WF_TAIL_LOG2
#define PTB_WF_TAIL_LOG2 3
#endif
constexpr uint32_t WF_TAIL_LOG2_BLOCKS = PTB_WF_TAIL_LOG2;  // the last pixels of a frame are traced as up to 2^this sample blocks each (see wavefront_render)
constexpr uint32_t WF_REGEN = 9;            // queue key of a slot without a live path: regenerate (next sample / next pixel)
constexpr uint32_t WF_NKEYS = 10;           // 8 lobe classes, WF_MISS, WF_REGEN
constexpr uint32_t WF_NOKEY = 0xffffu;      // slot left the queue for good (frame exhausted)
constexpr uint32_t PRIM_SKY = 0xffffffffu;

// GENERIC: scenes with partial material masks may have up to 64 primitives, so the accepted set needs 64 bits of its own
template <class R, uint32_t WF_POOL, uint32_t SCENE_BYTES, bool GENERIC> struct WfSmemT {
    using V4 = typename Vec4T<R>::type;
    uint32_t scene[SCENE_BYTES / 4];
    V4 ro[WF_POOL], rd[WF_POOL], tr[WF_POOL], ra[WF_POOL], ac[WF_POOL];
    uint32_t acc_lo[GENERIC ? WF_POOL : 1], acc_hi[GENERIC ? WF_POOL : 1];
    uint32_t kt[WF_POOL];           // queue key << 16 | ticket inside the key (WF_NOKEY: not queued)
    uint16_t order[WF_POOL];        // the queue: slot indices, key-ordered
    uint32_t cnt[2][WF_NKEYS + 2];  // tickets handed out per key for the NEXT queue (double-buffered by iteration parity)
    uint32_t cursor[2];             // next 32-entry chunk of the current queue (double-buffered like cnt)
    // PTB_WF_HALVES: the pool as two halves with a queue each (cnt[h], cursor[h] belong to half h); see the kernel
    uint32_t arrive[2];             // warps that have left the current queue of half h
    uint32_t epoch[2];              // sorts completed for half h (visit k may start once epoch >= k)
    uint32_t nq[2];                 // length of the current queue of half h
};

// a 32-bit word kept in the .w lane of a vector of reals (bit pattern only: the lane is never used arithmetically)
PTB_DEV float wf_word(float, uint32_t u) { return __uint_as_float(u); }
PTB_DEV double wf_word(double, uint32_t u) { return __hiloint2double(0, (int)u); }
PTB_DEV uint32_t wf_bits(float w) { return __float_as_uint(w); }
PTB_DEV uint32_t wf_bits(double w) { return (uint32_t)__double2loint(w); }
// shared-memory loads that stay where they are written (volatile: neither nvcc nor ptxas hoists them above the shading code)
PTB_DEV double4 lds128_late(const double4* p) {      // (256 bits here)
    double4 v;
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
    asm volatile("ld.volatile.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a) : "memory");
    asm volatile("ld.volatile.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.z), "=d"(v.w) : "r"(a + 16u) : "memory");
    return v;
}
PTB_DEV float4 lds128_late(const float4* p) {
    float4 v;
    asm volatile("ld.volatile.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"((uint32_t)__cvta_generic_to_shared(p)) : "memory");
    return v;
}
PTB_DEV uint32_t lds32_late(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"((uint32_t)__cvta_generic_to_shared(p)) : "memory");
    return v;
}

// One stage per iteration.  Every slot that still has work is in the queue exactly once, ordered by key; warps take 32-entry
// chunks and run, for their 32 slots,
//   A  the pending event of the slot's path: background lookup (WF_MISS) or shading (finalize, light sample + shadow ray,
//      Disney eval with MIS, Disney sample, next ray) — full warps of one lobe class;
//   B  for lanes whose path has ended (in A, or at an emitter / by roulette in C of the previous iteration): add the radiance to
//      the pixel sum and regenerate IN PLACE — next sample of the slot's pixel, or a new pixel handed out per warp with
//      ballot/popc from the global work counter (16x16-tile order), camera ray;
//   C  closest_hit (incl. lights with the stale hit_dist, MIS-weighted emission on a light hit) for EVERY lane — continuing
//      and regenerated paths alike, so the intersection code always runs on full warps —, then a ticket for the next queue
//      under the key of what was hit.
// Between iterations: one CTA barrier, the counting-sort scatter of the tickets into the next queue (offsets are recomputed
// by every warp from the 10 counters, no serial section), a second barrier.  Compared with the two-stage form (intersect all
// slots / sort / shade) a bounce costs one slot visit instead of two — half the state loads / stores and queue bookkeeping —,
// the previous bounce's pdf never leaves registers, and there are two CTA barriers per iteration instead of four.
//
// RM: the scene has a resolved-material table (RMat, ptb_device.cuh) and the host guarantees that the WHOLE blob sits in the
// shared-memory copy, so every scene read of this instantiation is a shared-memory load (LDS, not a generic LD).
template <class R> __host__ __device__ constexpr int wf_threads(bool rm) { return sizeof(R) == 8 ? WF_THREADS_F64 : (rm ? WF_THREADS_RM : WF_THREADS_GENERIC); }
template <class R> __host__ __device__ constexpr uint32_t wf_pool(bool rm) { return sizeof(R) == 8 ? WF_POOL_F64 : (rm ? WF_POOL_RM : WF_POOL_GENERIC); }
// EMB: instantiation for scenes whose primitives fit DScene::emb_* (see SV_SPHERE in ptb_device.cuh)
// FX:  the scene has media or live rectangular / distant lights (PTB_MEDIUM_*, PTB_SCENE_EXTENDED_LIGHTS).  A separate
//      instantiation: the in-medium bounce and the quad tests inlined into the shading stage cost the f64 kernel 7 % on scenes
//      that have neither.
template <class R, bool COUNT, bool BVH, bool RM, bool EMB = false, bool FX = false>
__global__ void __launch_bounds__(wf_threads<R>(RM), 1) k_render_wavefront(const __grid_constant__ DScene<R> s, const RenderArgs a) {
    constexpr int WF_THREADS = wf_threads<R>(RM);
    static_assert(!(RM && BVH), "the resolved-material table is for scenes that live in shared memory");
    static_assert(!RM || sizeof(R) == 4, "the resolved-material table is built in f32");
    constexpr uint32_t WF_POOL = wf_pool<R>(RM);
    using V4 = typename Vec4T<R>::type;
    using WfSmem = WfSmemT<R, WF_POOL, RM ? WF_SCENE_BYTES_RM : PTB_SMEM_SCENE_BYTES, !RM>;
    extern __shared__ __align__(16) unsigned char wf_raw[];
    WfSmem& sm = *reinterpret_cast<WfSmem*>(wf_raw);
    SceneView<R> sv;
    const uint32_t* rm_keys = nullptr;
    const RMat* rm_table = nullptr;
    if constexpr (RM) {
        const uint32_t* src = (const uint32_t*)s.blob;
        for (uint32_t i = threadIdx.x; i < s.blob_bytes / 4u; i += WF_THREADS) sm.scene[i] = src[i];
        const unsigned char* m = reinterpret_cast<const unsigned char*>(sm.scene);
        sv.planes = (const DPlane<R>*)(m + s.off_planes);
        sv.lights = (const DLight<R>*)(m + s.off_lights);
        sv.plane_material = (const uint32_t*)(m + s.off_plane_material);
        sv.spheres = (const DSphere<R>*)(m + s.off_spheres);
        sv.sphere_material = (const uint32_t*)(m + s.off_sphere_material);
        sv.materials = (const DMaterial<R>*)(m + s.off_materials);
        rm_keys = (const uint32_t*)(m + s.off_rm_keys);
        rm_table = (const RMat*)(m + s.off_rm_table);
    } else {
        sv = stage_scene(s, sm.scene, PTB_SMEM_SCENE_BYTES);
    }
    V4* accum = reinterpret_cast<V4*>(a.accum);

    const unsigned FULL = 0xffffffffu;
    const uint32_t tid = threadIdx.x, lane = tid & 31u;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const R inv_w = sizeof(R) == 8 ? (R)a.rcp_w64 : (R)a.rcp_w, inv_h = sizeof(R) == 8 ? (R)a.rcp_h64 : (R)a.rcp_h;     // pixel_size (pinhole.rs:41), correctly rounded on the host

    // the first queue: every slot, key WF_REGEN
    for (uint32_t i = tid; i < WF_POOL; i += WF_THREADS) {
        sm.ro[i] = mk4(R(0), R(0), R(0), R(-1));
        sm.rd[i] = mk4(R(0), R(0), R(1), wf_word(R(0), 0u));
        sm.tr[i] = mk4(R(0), R(0), R(0), wf_word(R(0), 0u));
        sm.ra[i] = mk4(R(0), R(0), R(0), wf_word(R(0), 0u));
        sm.ac[i] = mk4(R(0), R(0), R(0), wf_word(R(0), PRIM_SKY));
        sm.order[i] = (uint16_t)i;
        sm.kt[i] = WF_NOKEY << 16;
    }
    if (tid < 2u * (WF_NKEYS + 2u)) (&sm.cnt[0][0])[tid] = 0;
    if (tid < 2u) { sm.cursor[tid] = 0; sm.arrive[tid] = 0; sm.epoch[tid] = 0; sm.nq[tid] = WF_POOL / 2u; }
    __syncthreads();

#ifndef PTB_WF_HALVES
    uint32_t n_queue = WF_POOL, par = 0;      // par: which cnt / cursor buffer the CURRENT iteration's consumers use
#endif
    PathCounters pc;
    uint32_t n_samples = 0;
    if (COUNT) {
        pc.closest_hit = pc.any_hit = pc.shade = pc.nee_contrib = pc.eval_calls = 0;
        pc.lobe[0] = pc.lobe[1] = pc.lobe[2] = pc.lobe[3] = 0;
        pc.end_sky = pc.end_emitter = pc.end_pdf = pc.end_depth = pc.end_rr = 0;
        pc.ev[0] = pc.ev[1] = pc.ev[2] = pc.ev[3] = 0;
        pc.bvh[0] = pc.bvh[1] = 0;
    }

#ifdef PTB_WF_HALVES
    // Two half-pools with a queue each, and no CTA barrier: a warp that runs out of chunks in half A goes on to half B's queue
    // instead of waiting for the slowest chunk of A; the LAST warp to leave a queue sorts that half's tickets into its next queue
    // (alone: 39 steps of a warp) and publishes it.  All warps visit (half, round) in the same order, so "the queue of my k-th
    // visit of half h is ready" is `epoch[h] >= k`.
    static_assert((WF_POOL / 2u) % 32u == 0, "half pools of whole chunks");
    constexpr uint32_t HALF = WF_POOL / 2u, NWARPS = WF_THREADS / 32;
    uint32_t visit0 = 0, visit1 = 0;
    bool done0 = false, done1 = false;
    for (uint32_t turn = 0; !(done0 && done1); ++turn) {
        const uint32_t h = turn & 1u;
        if (h ? done1 : done0) continue;
        const uint32_t my_visit = h ? visit1 : visit0;
        if (lane == 0) { while (*(volatile uint32_t*)&sm.epoch[h] < my_visit) __nanosleep(40); }
        __syncwarp();
        __threadfence_block();
        const uint32_t n_queue = *(volatile uint32_t*)&sm.nq[h];
        if (n_queue == 0) { if (h) done1 = true; else done0 = true; continue; }
        const uint32_t WF_QBASE = h * HALF;
#define WF_CNT_IDX h
#define WF_CUR_IDX h
#else
    while (n_queue) {
        constexpr uint32_t WF_QBASE = 0;
        // (indices, not pointers: a pointer to a shared-memory member is a generic pointer, and forming one costs an S2R)
#define WF_CNT_IDX (par ^ 1u)
#define WF_CUR_IDX par
#endif
#pragma unroll 1
        while (true) {
            uint32_t chunk = 0;
            if (lane == 0) chunk = atomicAdd(&sm.cursor[WF_CUR_IDX], 1u);
            chunk = __shfl_sync(FULL, chunk, 0);
            if (chunk * 32u >= n_queue) break;
            const uint32_t j = chunk * 32u + lane;
            const bool valid = j < n_queue;
            const uint32_t i = valid ? sm.order[WF_QBASE + j] : 0u;
            // ---- load what the pending event needs: the ray, the hit, the flags ----
            // Radiance, throughput, pixel and sample index stay in shared memory while the event is evaluated: A works on a unit
            // throughput and a zero radiance, and what it returns is applied to the slot's values afterwards — `rad + x * thr` and
            // `thr * (f / pdf)` are the reference's operations on the reference's operands (tracer.rs:67, 89, 94), only the nine
            // registers are not live across the shading code, which is where the kernel's register pressure peaks.
            PathState<R> p;
            uint32_t fl0 = 0, prim_bits = PRIM_SKY, pxy0 = 0;
            uint64_t accepted = 0;
            p.o = V3<R>(0, 0, 0); p.d = V3<R>(0, 0, 1); p.thr = V3<R>(1, 1, 1); p.rad = V3<R>(0, 0, 0); p.hit_dist = R(-1); p.prev_pdf = 0; p.bounce = 0;
            if (valid) {
                const V4 q0 = sm.ro[i], q1 = sm.rd[i];
                p.o = V3<R>(q0.x, q0.y, q0.z); p.hit_dist = q0.w;
                p.d = V3<R>(q1.x, q1.y, q1.z); pxy0 = wf_bits(q1.w);
                fl0 = wf_bits(sm.tr[i].w);
                prim_bits = wf_bits(sm.ac[i].w);
                if constexpr (!RM) accepted = (uint64_t)sm.acc_lo[i] | ((uint64_t)sm.acc_hi[i] << 32);
            }
            p.bounce = (fl0 >> 8) & 0xffffu;
            p.medium = (RM || !FX) ? 0u : fl0 >> 25;                    // (media: the FX instantiations only)
            bool alive = valid && (fl0 & FL_ALIVE);
            const bool had_event = alive;

            // ================================ A: the path's pending event ================================
            if (alive) {
                if (prim_bits == PRIM_SKY) {                            // the path left the scene (tracer.rs:66-69)
                    path_add_sky(s, p);
                    if (COUNT) pc.end_sky++;
                    alive = false;
                } else {
                    Rng<R> rng((pxy0 >> 16) * a.W + (pxy0 & 0xffffu), a.sample_base + wf_bits(sm.ra[i].w), a.seed);
                    R u[8];
                    if constexpr (RM) {
                        const int prim = (int)(prim_bits & 0xffffu);
                        const RMat& rm = rm_lookup(s, sv, rm_keys, rm_table, rm_key_of(s, sv, prim, prim_bits >> 16), p.d);
                        if (s.has_emissive) {                           // tracer.rs:74, on the slot's own radiance and throughput
                            const V4 t4 = sm.tr[i];
                            V4 r4 = sm.ra[i];
                            r4.x = r4.x + rm.m.emission.x * t4.x; r4.y = r4.y + rm.m.emission.y * t4.y; r4.z = r4.z + rm.m.emission.z * t4.z;
                            sm.ra[i] = r4;
                        }
                        shade_draws(rng, p.bounce, s.n_lights > 1u || (rm.lobe_class & 4u) != 0u, u);
                        const V3<R> normal = hit_normal<R, BVH, false, EMB>(s, sv, prim, p.o, p.d, p.hit_dist);
                        alive = path_shade_rm<COUNT, false, EMB>(s, sv, p, normal, rm, u, &pc);
                    } else {
                        const int prim = (int)prim_bits;
                        Mat<R> mat;
                        const uint32_t mi = hit_material<R, BVH>(s, sv, prim, accepted, p.d, mat);
                        shade_draws(rng, p.bounce, s.n_lights > 1u || p.medium != 0u || (lobe_class_of(mat.metallic, mat.spec_trans, mat.clearcoat) & 4u) != 0u, u);
                        int med = MED_SURFACE;
                        if (FX && s.has_media && p.medium) med = path_medium<R, COUNT, BVH, !BVH>(s, sv, p, u, &pc);       // works on the unit throughput like the shading below
                        if (med != MED_SURFACE) {
                            alive = med == MED_SCATTERED;
                        } else {
                            if (s.has_emissive) {                       // tracer.rs:74 (behind a medium: on the attenuated throughput)
                                const V4 t4 = sm.tr[i];
                                V4 r4 = sm.ra[i];
                                r4.x = r4.x + mat.emission.x * (t4.x * p.thr.x); r4.y = r4.y + mat.emission.y * (t4.y * p.thr.y); r4.z = r4.z + mat.emission.z * (t4.z * p.thr.z);
                                sm.ra[i] = r4;
                            }
                            const V3<R> normal = hit_normal<R, BVH, !BVH, EMB>(s, sv, prim, p.o, p.d, p.hit_dist);
                            alive = path_shade<R, COUNT, BVH, false, !BVH, EMB, FX>(s, sv, p, normal, mat, u, &pc);
                            if (FX && s.has_media) path_medium_update<R, BVH>(s, sv, p, normal, mi);
                        }
                    }
                }
            }
            // ---- the rest of the slot; apply what A returned ----
            uint32_t fl = 0, sidx = 0, pxy = 0;
            if (valid) {
                const V4 q2 = lds128_late(&sm.tr[i]), q3 = lds128_late(&sm.ra[i]);
                V3<R> thr(q2.x, q2.y, q2.z), rad(q3.x, q3.y, q3.z);
                fl = wf_bits(q2.w); sidx = wf_bits(q3.w);
                pxy = lds32_late(reinterpret_cast<const uint32_t*>(&sm.rd[i].w));
                if (had_event) {
                    rad = rad + p.rad * thr;                            // background (tracer.rs:67) or next-event estimation (:89); p.rad is 0 without one
                    thr = thr * p.thr;                                  // throughput * (f / pdf) (:94); unused when the path ended
                }
                p.thr = thr; p.rad = rad;
            }
            bool have_pixel = fl & FL_PIXEL;
            uint32_t pix = (pxy >> 16) * a.W + (pxy & 0xffffu);
            if (alive && a.rr_start != 0 && p.bounce >= a.rr_start) {         // Russian-roulette extension (A.12), slot 0 of the new bounce
                Rng<R> rng(pix, a.sample_base + sidx, a.seed);
                R u4[4];
                rng.block(p.bounce, 0, u4);
                if (!russian_roulette_survives(p, u4[0])) { alive = false; if (COUNT) pc.end_rr++; }
            }

            // ================================ B: finish dead paths, regenerate in place ================================
            // Work items below a.n_whole are whole pixels (all spp samples); the items above are the frame's last a.tail_zt
            // pixels cut into sample blocks, so that the ramp-down at the end of the launch lasts one block, not one pixel.
            // A block's sum goes to a side buffer and k_tail_combine adds the blocks of a pixel in block order: the image
            // stays independent of which slot traced what.
            bool done = false;
            uint32_t blkbits = fl & FL_BLOCK_BITS;
            // In place only where B runs on (nearly) full warps: always in WF_MISS / WF_REGEN chunks, and in a shading chunk when at
            // least 1 / PTB_WF_REGEN_DEN of its paths ended there (pdf <= 0, depth: ~12 % on the demo scene).  Otherwise the few dead
            // lanes take a WF_REGEN ticket and are regenerated by a full warp next iteration instead of dragging this warp through
            // the pixel hand-out, Philox and camera-ray code at 4 of 32 lanes.
            const unsigned dead_mask = __ballot_sync(FULL, valid && !alive);
            const bool regen_here = (uint32_t)__popc(dead_mask) * PTB_WF_REGEN_DEN >= (uint32_t)__popc(__ballot_sync(FULL, valid));
            if (regen_here) {
                V3<R> acc(0, 0, 0);
                const bool dead = valid && !alive;
                if (dead) {
                    const V4 q4 = sm.ac[i];
                    acc = V3<R>(q4.x + p.rad.x, q4.y + p.rad.y, q4.z + p.rad.z);      // a fresh slot adds 0 to 0
                    sidx += have_pixel ? 1u : 0u;
                }
                const uint32_t blk = (fl >> 3) & 31u;
                const uint32_t s_end = (fl & FL_BLOCK) ? ((blk + 1u) * a.spp) >> a.tail_log2b : a.spp;
                bool want = dead && (!have_pixel || sidx == s_end);
                if (want && have_pixel) {
                    if (fl & FL_BLOCK) {
                        const uint32_t px0 = pxy & 0xffffu, pr0 = pxy >> 16;
                        const uint32_t pidx = (((pr0 >> 4) * a.tiles_x + (px0 >> 4)) << 8) | ((pr0 & 15u) << 4) | (px0 & 15u);
                        const uint32_t s_begin = (blk * a.spp) >> a.tail_log2b;
                        reinterpret_cast<V4*>(a.tail_side)[(pidx - a.n_whole) + blk * a.tail_zt] = mk4(acc.x, acc.y, acc.z, (R)(s_end - s_begin));
                    } else if (a.flush_dst) {       // multi-GPU: this launch's partial sum goes straight into the root GPU's slot (peer store)
                        reinterpret_cast<V4*>(a.flush_dst)[pix] = mk4(acc.x, acc.y, acc.z, (R)a.spp);
                    } else {
                        V4 v = accum[pix];
                        v.x += acc.x; v.y += acc.y; v.z += acc.z; v.w += (R)a.spp;
                        accum[pix] = v;
                    }
                    have_pixel = false;
                }
                // Hand-out: the lanes that want a pixel reserve exactly as many work items as they are (one warp-aggregated atomic on
                // the global counter).  Nothing is reserved ahead: with dynamic chunks a warp has no slots of its own that would be
                // sure to consume a leftover reservation before the queue drains, and an item held back would be a lost pixel.
                unsigned need = __ballot_sync(FULL, want);
                while (need) {
                    const uint32_t n_need = __popc(need);
                    uint32_t b = 0;
                    if (lane == 0) b = atomicAdd(a.work_counter, n_need);
                    b = __shfl_sync(FULL, b, 0);
                    if (want) {
                        uint32_t idx = b + __popc(need & lt_mask), nb = 0, s0 = 0;
                        if (idx >= a.n_items) {                                  // frame exhausted: this slot is finished for good
                            done = true; want = false;
                        } else {
                            if (idx >= a.n_whole) {                              // tail item: (pixel, sample block)
                                const uint32_t k = idx - a.n_whole, b2 = k / a.tail_zt;
                                idx = a.n_whole + (k - b2 * a.tail_zt);
                                nb = FL_BLOCK | (b2 << 3);
                                s0 = (b2 * a.spp) >> a.tail_log2b;
                            }
                            const uint32_t tile = idx >> 8, within = idx & 255u;
                            const uint32_t px = (tile % a.tiles_x) * 16u + (within & 15u);
                            const uint32_t prow = (tile / a.tiles_x) * 16u + (within >> 4);
                            if (px < a.W && prow < a.H) {                        // (items of a partial tile outside the frame: take another)
                                pix = prow * a.W + px; pxy = px | (prow << 16);
                                have_pixel = true; want = false; sidx = s0; blkbits = nb;
                                acc = V3<R>(0, 0, 0);
                            }
                        }
                    }
                    need = __ballot_sync(FULL, want);
                }
                if (dead) {
                    sm.ac[i] = mk4(acc.x, acc.y, acc.z, wf_word(R(0), PRIM_SKY));
                    if (!done) {                                        // next sample of the slot's pixel
                        Rng<R> rng(pix, a.sample_base + sidx, a.seed);
                        R u4[4];
                        rng.block(0, 0, u4);
                        path_begin(s, p, pxy & 0xffffu, pxy >> 16, a.W, a.H, inv_w, inv_h, u4[0], u4[1], sizeof(R) == 4 && a.film_fast != 0u, a.rcp_w, a.rcp_h);
                        alive = true;
                        if (COUNT) n_samples++;
                    }
                }
            }

            // ================================ C: closest_hit for every live path ================================
            uint32_t key = valid && !done ? WF_REGEN : WF_NOKEY;      // (a lane that is neither alive nor done waits for its regeneration)
            if (alive) {
                uint32_t new_prim = PRIM_SKY;
                if (p.bounce >= s.depth) {                              // recursion depth 0 (tracer.rs:61)
                    alive = false;
                    if (COUNT) pc.end_depth++;
                } else {
                    if (COUNT) pc.closest_hit++;
                    const HitCore<R> h = closest_hit_core<R, BVH, !RM && !BVH, !RM && FX, EMB>(s, sv, p.o, p.d, p.hit_dist, COUNT ? pc.bvh : nullptr);   // signed-distance programs: generic instantiation only
                    p.hit_dist = h.hit_dist;
                    if (!h.hit) {
                        key = WF_MISS;                                 // background lookup next iteration, with full warps
                    } else if (h.is_emitter) {
                        path_add_emitter<R, BVH>(s, sv, p, h);
                        alive = false;
                        if (COUNT) pc.end_emitter++;
                    } else {
                        if constexpr (RM) {
                            key = rm_table[rm_keys[rm_key_of(s, sv, h.prim, (uint32_t)h.accepted)] & 0xffffu].lobe_class;
                            new_prim = ((uint32_t)h.prim & 0xffffu) | ((uint32_t)h.accepted << 16);
                        } else {
                            key = hit_lobe_class<R, BVH>(s, sv, h.prim, h.accepted);
                            new_prim = (uint32_t)h.prim;
                            sm.acc_lo[i] = (uint32_t)h.accepted; sm.acc_hi[i] = (uint32_t)(h.accepted >> 32);
                        }
                    }
                }
                sm.ro[i] = mk4(p.o.x, p.o.y, p.o.z, p.hit_dist);
                sm.rd[i] = mk4(p.d.x, p.d.y, p.d.z, wf_word(R(0), pxy));
                sm.ac[i].w = wf_word(R(0), new_prim);
            }
            if (valid) {
                fl = (p.bounce << 8) | blkbits | (alive ? FL_ALIVE : 0u) | (have_pixel ? FL_PIXEL : 0u) | (p.medium << 25);
                sm.tr[i] = mk4(p.thr.x, p.thr.y, p.thr.z, wf_word(R(0), fl));
                sm.ra[i] = mk4(p.rad.x, p.rad.y, p.rad.z, wf_word(R(0), sidx));
                uint32_t ticket = 0;
                if (key != WF_NOKEY) ticket = atomicAdd(&sm.cnt[WF_CNT_IDX][key], 1u);
                sm.kt[i] = (key << 16) | ticket;
            }
        }
#ifdef PTB_WF_HALVES
        // ---- leave the queue; the last warp out sorts the half ----
        __syncwarp();
        uint32_t last = 0;
        if (lane == 0) { __threadfence_block(); last = atomicAdd(&sm.arrive[h], 1u) == NWARPS - 1u ? 1u : 0u; }
        last = __shfl_sync(FULL, last, 0);
        if (last) {
            __threadfence_block();
            uint32_t off_lane = 0, run = 0;
            {
                const int order_by_cost[WF_NKEYS] = {7, 3, 5, 6, 1, 2, 4, 0, (int)WF_MISS, (int)WF_REGEN};
#pragma unroll
                for (int k = 0; k < (int)WF_NKEYS; ++k) { const int c = order_by_cost[k]; off_lane = (uint32_t)c == lane ? run : off_lane; run += *(volatile uint32_t*)&sm.cnt[h][c]; }
            }
#pragma unroll 1
            for (uint32_t i = WF_QBASE + lane; i < WF_QBASE + HALF; i += 32u) {
                const uint32_t kt = *(volatile uint32_t*)&sm.kt[i], k = kt >> 16;
                const uint32_t o = __shfl_sync(FULL, off_lane, (int)(k & 31u));
                if (k != WF_NOKEY) {
                    sm.order[WF_QBASE + o + (kt & 0xffffu)] = (uint16_t)i;
                    sm.kt[i] = WF_NOKEY << 16;
                }
            }
            __syncwarp();
            if (lane < WF_NKEYS) sm.cnt[h][lane] = 0;
            if (lane == 0) { sm.cursor[h] = 0; sm.arrive[h] = 0; sm.nq[h] = run; }
            __syncwarp();
            __threadfence_block();
            if (lane == 0) *(volatile uint32_t*)&sm.epoch[h] = my_visit + 1u;
        }
        if (h) ++visit1; else ++visit0;
    }
#else
        __syncthreads();

        // ================================ sort: tickets -> key-ordered queue ================================
        // keys with more lobes cost more to process: queue them first so that the dynamic chunking ends on cheap chunks
        // (longest-processing-time-first); every thread derives the offsets itself from the ten counters
#ifdef PTB_WF_SORT_SELECT
        uint32_t off[WF_NKEYS];
        {
            const int order_by_cost[WF_NKEYS] = {7, 3, 5, 6, 1, 2, 4, 0, (int)WF_MISS, (int)WF_REGEN};
            uint32_t run = 0;
#pragma unroll
            for (int k = 0; k < (int)WF_NKEYS; ++k) { const int c = order_by_cost[k]; off[c] = run; run += sm.cnt[par ^ 1u][c]; }
            n_queue = run;
        }
#pragma unroll 1
        for (uint32_t i = tid; i < WF_POOL; i += WF_THREADS) {
            const uint32_t kt = sm.kt[i], k = kt >> 16;
            if (k != WF_NOKEY) {
                uint32_t o = 0;
#pragma unroll
                for (int c = 0; c < (int)WF_NKEYS; ++c) o = (k == (uint32_t)c) ? off[c] : o;
                sm.order[o + (kt & 0xffffu)] = (uint16_t)i;
                sm.kt[i] = WF_NOKEY << 16;
            }
        }
#else
        // lane c of every warp holds the queue offset of key c; a slot fetches its key's offset with one shuffle (the first
        // version selected it out of ten registers: 20 instructions per slot, 2 % of the kernel's)
        uint32_t off_lane = 0;
        {
            const int order_by_cost[WF_NKEYS] = {7, 3, 5, 6, 1, 2, 4, 0, (int)WF_MISS, (int)WF_REGEN};
            uint32_t run = 0;
#pragma unroll
            for (int k = 0; k < (int)WF_NKEYS; ++k) { const int c = order_by_cost[k]; off_lane = (uint32_t)c == lane ? run : off_lane; run += sm.cnt[par ^ 1u][c]; }
            n_queue = run;
        }
        static_assert(WF_POOL % 32u == 0, "whole warps in the sort pass (the shuffle below needs all 32 lanes)");
#pragma unroll 1
        for (uint32_t i = tid; i < WF_POOL; i += WF_THREADS) {
            const uint32_t kt = sm.kt[i], k = kt >> 16;
            const uint32_t o = __shfl_sync(FULL, off_lane, (int)(k & 31u));
            if (k != WF_NOKEY) {
                sm.order[o + (kt & 0xffffu)] = (uint16_t)i;
                sm.kt[i] = WF_NOKEY << 16;
            }
        }
#endif
        if (tid < WF_NKEYS) sm.cnt[par][tid] = 0;       // this iteration's consumers are done with it; it collects the tickets of the next one
        if (tid == 0) sm.cursor[par] = 0;
        par ^= 1u;
        __syncthreads();
    }
#endif

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

struct WavefrontState {
    bool configured = false;
    uint32_t film_w = 0, film_h = 0;     // frame size the film_fast verdict below was established for
    bool film_fast = false;
    void* tail_side = nullptr;           // float4[tail_zt << WF_TAIL_LOG2_BLOCKS]: block sums of the tail pixels
    size_t tail_side_bytes = 0;
    void release() { if (tail_side) cudaFree(tail_side); tail_side = nullptr; tail_side_bytes = 0; }
};

// Adds the sample blocks of every tail pixel in block order (fixed association: the result does not depend on which slot traced
// which block) to the accumulator, or stores the sum into the peer slot like the render kernel does for whole pixels.
template <class R> __global__ void k_tail_combine(const RenderArgs a) {
    using V4 = typename Vec4T<R>::type;
    const uint32_t z = blockIdx.x * blockDim.x + threadIdx.x;
    if (z >= a.tail_zt) return;
    const uint32_t idx = a.n_whole + z, tile = idx >> 8, within = idx & 255u;
    const uint32_t px = (tile % a.tiles_x) * 16u + (within & 15u), prow = (tile / a.tiles_x) * 16u + (within >> 4);
    if (px >= a.W || prow >= a.H) return;
    const V4* side = reinterpret_cast<const V4*>(a.tail_side);
    V4 sum = side[z];
    for (uint32_t b = 1; b < (1u << a.tail_log2b); ++b) {
        const V4 v = side[z + b * a.tail_zt];
        sum.x += v.x; sum.y += v.y; sum.z += v.z; sum.w += v.w;
    }
    const uint32_t pix = prow * a.W + px;
    if (a.flush_dst) {
        reinterpret_cast<V4*>(a.flush_dst)[pix] = sum;
    } else {
        V4* accum = reinterpret_cast<V4*>(a.accum);
        V4 v = accum[pix];
        v.x += sum.x; v.y += sum.y; v.z += sum.z; v.w += sum.w;
        accum[pix] = v;
    }
}

// RenderArgs::film_fast: every column / row quotient of this frame size, FMA-corrected vs IEEE (W + H checks per frame size)
inline bool film_coords_fma_exact(uint32_t W, uint32_t H) {
    if (W == 0 || H == 0 || W >= (1u << 24) || H >= (1u << 24)) return false;
    const float wf = (float)W, hf = (float)H, rw = 1.0f / wf, rh = 1.0f / hf;
    for (uint32_t x = 0; x < W; ++x)
        if (div_by_fma((float)x, wf, rw) != (float)x / wf) return false;
    for (uint32_t y = 1; y <= H; ++y)
        if (div_by_fma((float)y, hf, rh) != (float)y / hf) return false;
    return true;
}

// host launcher: one persistent CTA per SM
template <class R>
inline int wavefront_render(WavefrontState& wf, const DScene<R>& d, void* accum, void* flush_dst, uint32_t W, uint32_t H, uint32_t spp, uint64_t sample_base,
                            const ptb_config& cfg, cudaStream_t stream, int sm_count, DeviceCounters* counters, unsigned int* work_counter,
                            cudaEvent_t ev0, cudaEvent_t ev1, uint64_t* launches, std::string& err) {
    constexpr bool F32 = sizeof(R) == 4;
    using V4 = typename Vec4T<R>::type;
    RenderArgs a{};
    a.accum = accum; a.flush_dst = flush_dst; a.W = W; a.H = H; a.spp = spp; a.sample_base = sample_base; a.seed = cfg.seed; a.rr_start = cfg.rr_start;
    a.tiles_x = (W + 15u) / 16u;
    a.n_items = a.tiles_x * ((H + 15u) / 16u) * 256u;
    a.work_counter = work_counter;
    a.counters = counters;
    if (F32 && (wf.film_w != W || wf.film_h != H)) { wf.film_fast = film_coords_fma_exact(W, H); wf.film_w = W; wf.film_h = H; }
    a.film_fast = F32 && wf.film_fast ? 1u : 0u;
#ifdef PTB_NO_FILM_FMA
    a.film_fast = 0u;
#endif
    a.rcp_w = 1.0f / (float)W; a.rcp_h = 1.0f / (float)H;
    a.rcp_w64 = 1.0 / (double)W; a.rcp_h64 = 1.0 / (double)H;
    const bool count = cfg.collect_counters != 0;
    bool rm = false;
    void (*kern)(const DScene<R>, const RenderArgs);
    size_t smem_bytes;
#define WF_K(RR, B, RMv, E, F) (count ? k_render_wavefront<RR, true, B, RMv, E, F> : k_render_wavefront<RR, false, B, RMv, E, F>)
    const bool fx = d.has_fx != 0;          // media or live rectangular / distant lights: the FX instantiations (never with RM)
    if constexpr (F32) {
        rm = d.rm_entries != 0 && !d.use_bvh;
        kern = d.use_bvh ? (fx ? WF_K(float, true, false, false, true) : WF_K(float, true, false, false, false))
               : rm      ? (d.emb ? WF_K(float, false, true, true, false) : WF_K(float, false, true, false, false))
               : fx      ? WF_K(float, false, false, false, true)
               : d.emb   ? WF_K(float, false, false, true, false) : WF_K(float, false, false, false, false);
        smem_bytes = rm ? sizeof(WfSmemT<float, WF_POOL_RM, WF_SCENE_BYTES_RM, false>) : sizeof(WfSmemT<float, WF_POOL_GENERIC, PTB_SMEM_SCENE_BYTES, true>);
    } else {
        kern = d.use_bvh ? (fx ? WF_K(double, true, false, false, true) : WF_K(double, true, false, false, false))
               : fx      ? WF_K(double, false, false, false, true)
               : d.emb   ? WF_K(double, false, false, true, false) : WF_K(double, false, false, false, false);
        smem_bytes = sizeof(WfSmemT<double, WF_POOL_F64, PTB_SMEM_SCENE_BYTES, true>);
    }
#undef WF_K
    const uint32_t WF_POOL = wf_pool<R>(rm);
    const int threads = wf_threads<R>(rm);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
    if (e != cudaSuccess) { err = std::string("cudaFuncSetAttribute(wavefront smem): ") + cudaGetErrorString(e); return PTB_E_CUDA; }
    wf.configured = true;
    uint32_t max_useful = (a.n_items + WF_POOL - 1) / WF_POOL;
    int grid = (int)std::max<uint32_t>(1u, std::min<uint32_t>((uint32_t)sm_count, max_useful));
    // Tail items: when the work counter runs dry every slot is somewhere inside its last pixel, and the launch ramps down
    // over one pixel's worth of iterations (measured: -15 % at 1920x1080, -4 % at 3840x2160).  The last `grid * WF_POOL`
    // pixels (one per slot) are therefore handed out as 8 sample blocks each, which shortens the ramp eightfold.
    a.n_whole = a.n_items; a.tail_zt = 0; a.tail_log2b = 0; a.tail_side = nullptr;
#ifndef PTB_WF_NO_TAIL
    if (spp >= 2u && spp <= (1u << 24)) {
        uint32_t log2b = 1;
        while (log2b < WF_TAIL_LOG2_BLOCKS && (2u << log2b) <= spp) ++log2b;      // at least one sample per block
        const uint32_t zt = std::min<uint32_t>(a.n_items, (((uint32_t)grid * WF_POOL + 255u) / 256u) * 256u);
        const size_t need = ((size_t)zt << log2b) * sizeof(V4);
        if (need > wf.tail_side_bytes) {
            if (wf.tail_side) cudaFree(wf.tail_side);
            wf.tail_side = nullptr; wf.tail_side_bytes = 0;
            if ((e = cudaMalloc(&wf.tail_side, need)) != cudaSuccess) { err = std::string("cudaMalloc(tail blocks): ") + cudaGetErrorString(e); return PTB_E_CUDA; }
            wf.tail_side_bytes = need;
        }
        a.n_whole = a.n_items - zt; a.tail_zt = zt; a.tail_log2b = log2b; a.tail_side = wf.tail_side;
        a.n_items = a.n_whole + (zt << log2b);
    }
#endif
    if ((e = cudaMemsetAsync(work_counter, 0, sizeof(unsigned int), stream)) != cudaSuccess ||
        (e = cudaEventRecord(ev0, stream)) != cudaSuccess) { err = cudaGetErrorString(e); return PTB_E_CUDA; }
    kern<<<grid, threads, smem_bytes, stream>>>(d, a);
    if ((e = cudaGetLastError()) == cudaSuccess && a.tail_zt) {
        k_tail_combine<R><<<(a.tail_zt + 255u) / 256u, 256, 0, stream>>>(a);
        e = cudaGetLastError();
        (*launches)++;
    }
    if (e != cudaSuccess || (e = cudaEventRecord(ev1, stream)) != cudaSuccess) {
        err = std::string("k_render_wavefront launch: ") + cudaGetErrorString(e);
        return PTB_E_CUDA;
    }
    (*launches)++;
    return PTB_OK;
}

}  // namespace ptb
