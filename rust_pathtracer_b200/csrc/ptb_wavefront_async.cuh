// ptb_wavefront.cuh — the wavefront integrator: SoA path-state queues, one stage per kind of work,
// queues compacted / sorted between stages, persistent CTAs — with the queues held in the SM's
// 227 KB of SHARED MEMORY instead of HBM.
//
// Why shared memory: the demo scene costs ~1.4 kFLOP and ~2 bounces per sample.
// A classic HBM wavefront streams ~1 KB of ray / path state per sample through the queues, which
// caps it at ~6 Gsamples/s on a 6.5 TB/s part before any arithmetic is done (SURVEY.md §7 "the
// roofline that actually binds").  One SM can hold 2048 paths x 96 B (or 2304 x 88 B) of state — 221 / 229 KB
// with the queue arrays and the scene copy —, enough for every stage to run with full warps, so the state
// never leaves the SM: HBM traffic stays at the 32 B per pixel of the accumulator read-modify-write.
//
// One CTA per SM (512 threads and 2048 slots; 768 threads and 2304 slots in the instantiation that shades from
// the resolved-material table, RMat in ptb_device.cuh) owns the pool.  Each iteration runs two stages over it,
// separated by CTA barriers, so that ALL warps of the SM execute the same stage code at the same time (small
// instruction-cache footprint, the fused kernel's main stall):
//
//   stage 1  "generate + intersect"  — every slot: a finished slot regenerates in place (next
//            sample of its pixel, or a new pixel handed out per warp with ballot/popc from a global
//            tile counter), then camera-ray generation / closest_hit (+ spherical lights with the
//            stale hit_dist quirk), MIS-weighted emission on a light hit.  Surviving paths enter the
//            stage-2 queue: they take a ticket in the counter of their key — the LOBE CLASS of the
//            hit material (which Disney lobes it can express) or WF_MISS (the path left the scene).
//   sort     a counting sort by key turns the tickets into a compacted, key-ordered index list
//            (the queue proper), most expensive classes first.
//   stage 2  "shade" — warps take 32-entry chunks of the queue, so a warp shades paths of one lobe
//            class: finalize, light sampling + any_hit shadow ray, Disney eval with MIS, Disney
//            sample, throughput update, next ray (or termination); WF_MISS chunks do the background
//            lookup with full warps.
//
// A slot owns one pixel for `spp` consecutive samples and sums them in sample order; the frame's last pixels (one
// per slot) are cut into sample blocks that k_tail_combine adds in block order (see wavefront_render).  Either way
// the image does not depend on scheduling: it is bit-reproducible run to run, like the fused integrator's.
#pragma once
#include <string>

#include "ptb_kernels.cuh"

namespace ptb {

// threads per CTA (one CTA per SM).  The generic and BVH instantiations need ~125 registers in the shade stage: 512 threads.
// With the resolved-material table the material is read from shared memory where it is used and the kernel fits 80
// registers (8 bytes of spill), so 768 threads = 24 warps hide the stage's dependent-issue and barrier stalls better
// (measured, 4K demo scene: 512 / 640 / 768 threads = 6067 / 6151 / 6451 Msamples/s; profiles/r01_ab_variants.txt).
#ifndef PTB_WF_THREADS
#define PTB_WF_THREADS 512
#endif
#ifndef PTB_WF_THREADS_RM
#define PTB_WF_THREADS_RM 768
#endif
constexpr int WF_THREADS_GENERIC = PTB_WF_THREADS;
constexpr int WF_THREADS_RM = PTB_WF_THREADS_RM;
// shared-memory scene copy of the resolved-material instantiation (the host builds the table only if the blob fits)
#ifndef PTB_WF_SCENE_BYTES_RM
#define PTB_WF_SCENE_BYTES_RM PTB_SMEM_SCENE_BYTES
#endif
constexpr uint32_t WF_SCENE_BYTES_RM = PTB_WF_SCENE_BYTES_RM;
// generic / BVH instantiations: 10 KB (the two accepted-set words per slot take the rest of the 227 KB); stage_scene copies the
// whole blob when it fits, else the head section
constexpr uint32_t WF_SCENE_BYTES_GENERIC = 10 * 1024;
#ifndef PTB_WF_POOL
#define PTB_WF_POOL 2048
#endif
constexpr uint32_t WF_POOL_GENERIC = PTB_WF_POOL;       // path slots per CTA
// resolved-material instantiation: 2304 slots = 3 x 768, every warp owns exactly three 32-slot groups in stage 1 (2048 slots
// left a third of the warps idle for one group in three); the slot state is two words smaller there (see U_* below)
#ifndef PTB_WF_POOL_RM
#define PTB_WF_POOL_RM 2048
#endif
constexpr uint32_t WF_POOL_RM = PTB_WF_POOL_RM;
constexpr uint32_t WF_MISS = 8;             // the path left the scene: background lookup, done with full warps in stage 2

// per-slot state: five 16-byte vectors (one LDS.128 / STS.128 each)
//   ro = (o.xyz, State::hit_dist)        rd = (d.xyz, pixel column | row << 16)
//   tr = (throughput.xyz, flags)         ra = (radiance.xyz, sample index)
//   ac = (pixel sum.xyz, hit primitive / accepted set of the pending shading event)
// flags word: bit0 alive, bit1 have_pixel, bits 3..7 sample block (tail items), bits 8..23 bounce, bit 24 tail item
constexpr uint32_t FL_ALIVE = 1u, FL_PIXEL = 2u, FL_BLOCK = 1u << 24, FL_BLOCK_BITS = FL_BLOCK | (31u << 3);
#ifndef PTB_WF_TAIL_LOG2
#define PTB_WF_TAIL_LOG2 3
#endif
constexpr uint32_t WF_TAIL_LOG2_BLOCKS = PTB_WF_TAIL_LOG2;  // the last pixels of a frame are traced as up to 2^this sample blocks each (see wavefront_render)
constexpr uint32_t WF_REGEN = 9;            // queue key of a slot without a live path: regenerate (next sample / next pixel)
constexpr uint32_t WF_NKEYS = 10;           // 8 lobe classes, WF_MISS, WF_REGEN
constexpr uint32_t WF_NOKEY = 0xffffu;      // slot left the queues for good (frame exhausted)
constexpr uint32_t WF_EMPTY = 0xffffu;      // ring cell not (yet) written
constexpr uint32_t PRIM_SKY = 0xffffffffu;

// One ring of slot indices per queue key.  A slot is in at most one ring, so a ring of WF_POOL cells can never overflow; head and
// tail are monotone reservation counters (cell = counter mod WF_POOL, WF_POOL a power of two).
// GENERIC: scenes with partial material masks may have up to 64 primitives, so the accepted set needs 64 bits of its own
template <uint32_t WF_POOL, uint32_t SCENE_BYTES, bool GENERIC> struct WfSmemT {
    uint32_t scene[SCENE_BYTES / 4];
    float4 ro[WF_POOL], rd[WF_POOL], tr[WF_POOL], ra[WF_POOL], ac[WF_POOL];
    uint32_t acc_lo[GENERIC ? WF_POOL : 1], acc_hi[GENERIC ? WF_POOL : 1];
    uint16_t ring[WF_NKEYS][WF_POOL];
    uint32_t head[WF_NKEYS];        // cells handed to consumers so far
    uint32_t tail[WF_NKEYS];        // cells reserved by producers so far
    uint32_t n_live;                // slots that have not left the queues for good
};

// The queues never drain between "iterations": there are none.  Every slot that still has work sits in the ring of its key;
// a warp takes up to 32 slots of ONE key (a full chunk whenever some ring holds 32 — with 2048 slots, at most 32 per warp in
// flight and ten keys, one always does until the frame runs out), and runs, for its lanes,
//   A  the pending event of the slot's path: background lookup (WF_MISS) or shading (finalize, light sample + shadow ray,
//      Disney eval with MIS, Disney sample, next ray) — full warps of one lobe class;
//   B  for lanes whose path has ended: add the radiance to the pixel sum and regenerate IN PLACE — next sample of the slot's
//      pixel, or a new pixel handed out per warp with ballot/popc from the global work counter (16x16-tile order), camera
//      ray.  Runs where it runs on (nearly) full warps — WF_MISS and WF_REGEN chunks; the few lanes of a shading chunk whose
//      path ended there (pdf <= 0, depth) go to the WF_REGEN ring instead of dragging the whole warp through B at 4 of 32 lanes;
//   C  closest_hit (incl. lights with the stale hit_dist, MIS-weighted emission on a light hit) for every live lane, then the
//      slot enters the ring of what was hit (warp-aggregated: one shared-memory atomic per distinct key).
// There is NO CTA barrier after the prologue and no sort pass: a warp that finishes a chunk takes the next one at once, so no
// warp waits for the slowest chunk of an iteration (the iteration-synchronous form lost 8 % of its warp cycles at that barrier:
// profiles/r02_ncu_wavefront.md, r02-a), and chunks never straddle a key boundary.  Cell hand-over is per cell: a producer
// waits for WF_EMPTY before it writes, a consumer for non-EMPTY before it reads and then resets the cell; the slot state is
// published with a block-level fence before the cell is.
//
// RM: the scene has a resolved-material table (RMat, ptb_device.cuh) and the host guarantees that the WHOLE blob sits in the
// shared-memory copy, so every scene read of this instantiation is a shared-memory load (LDS, not a generic LD).
template <bool COUNT, bool BVH, bool RM>
__global__ void __launch_bounds__(RM ? WF_THREADS_RM : WF_THREADS_GENERIC, 1) k_render_wavefront(const __grid_constant__ DScene<float> s, const RenderArgs a) {
    using R = float;
    constexpr int WF_THREADS = RM ? WF_THREADS_RM : WF_THREADS_GENERIC;
    static_assert(!(RM && BVH), "the resolved-material table is for scenes that live in shared memory");
    constexpr uint32_t WF_POOL = RM ? WF_POOL_RM : WF_POOL_GENERIC;
    static_assert((WF_POOL & (WF_POOL - 1u)) == 0u && WF_POOL <= 32768u, "ring cells are addressed modulo a power of two; slot indices are 16 bits");
    using WfSmem = WfSmemT<WF_POOL, RM ? WF_SCENE_BYTES_RM : WF_SCENE_BYTES_GENERIC, !RM>;
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
        sv = stage_scene(s, sm.scene, WF_SCENE_BYTES_GENERIC);
    }
    float4* accum = reinterpret_cast<float4*>(a.accum);

    const unsigned FULL = 0xffffffffu;
    const uint32_t tid = threadIdx.x, lane = tid & 31u;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const R inv_w = a.rcp_w, inv_h = a.rcp_h;     // pixel_size (pinhole.rs:41), correctly rounded on the host

    // every slot starts in the WF_REGEN ring
    for (uint32_t i = tid; i < WF_POOL; i += WF_THREADS) {
        sm.ro[i] = make_float4(0.f, 0.f, 0.f, -1.f);
        sm.rd[i] = make_float4(0.f, 0.f, 1.f, __uint_as_float(0u));
        sm.tr[i] = make_float4(0.f, 0.f, 0.f, __uint_as_float(0u));
        sm.ra[i] = make_float4(0.f, 0.f, 0.f, __uint_as_float(0u));
        sm.ac[i] = make_float4(0.f, 0.f, 0.f, __uint_as_float(PRIM_SKY));
#pragma unroll
        for (uint32_t k = 0; k < WF_NKEYS; ++k) sm.ring[k][i] = (uint16_t)(k == WF_REGEN ? i : WF_EMPTY);
    }
    if (tid < WF_NKEYS) { sm.head[tid] = 0; sm.tail[tid] = tid == WF_REGEN ? WF_POOL : 0u; }
    if (tid == 0) sm.n_live = WF_POOL;
    __syncthreads();

    volatile uint32_t* v_head = sm.head;
    volatile uint32_t* v_tail = sm.tail;
    volatile uint32_t* v_live = &sm.n_live;

    PathCounters pc;
    uint32_t n_samples = 0;
    if (COUNT) {
        pc.closest_hit = pc.any_hit = pc.shade = pc.nee_contrib = pc.eval_calls = 0;
        pc.lobe[0] = pc.lobe[1] = pc.lobe[2] = pc.lobe[3] = 0;
        pc.end_sky = pc.end_emitter = pc.end_pdf = pc.end_depth = pc.end_rr = 0;
        pc.ev[0] = pc.ev[1] = pc.ev[2] = pc.ev[3] = 0;
        pc.bvh[0] = pc.bvh[1] = 0;
    }

#pragma unroll 1
    while (true) {
        // ================================ take a chunk: up to 32 slots of one key ================================
        // lane 0 scans the rings, most expensive keys first, for one that holds a full chunk; failing that it takes the fullest
        // (only at the end of the frame).  The reservation is a compare-and-swap on the ring's head.
        uint32_t ck = WF_NOKEY, cbase = 0, cn = 0;
        {
            // lane q looks at ring q; a ring with a full chunk is picked in an order rotated per warp (so that concurrent scans do
            // not all go for the same head), else the fullest one (only when the frame runs out)
            const uint32_t my_key = lane < WF_NKEYS ? lane : 0u;
            uint32_t tries = tid >> 5;
            while (true) {
                uint32_t h = 0, n = 0;
                if (lane < WF_NKEYS) { h = v_head[my_key]; n = v_tail[my_key] - h; }
                const unsigned full = __ballot_sync(FULL, n >= 32u);
                uint32_t pick;
                if (full) {
                    const uint32_t rot = tries % WF_NKEYS, m = full >> rot;
                    pick = m ? rot + (uint32_t)__ffs(m) - 1u : (uint32_t)__ffs(full) - 1u;
                } else {
                    const uint32_t mx = __reduce_max_sync(FULL, n);
                    if (mx == 0u) {
                        if (*v_live == 0u) break;               // every slot has left the queues: the frame is done
                        __nanosleep(200);
                        continue;
                    }
                    pick = (uint32_t)__ffs(__ballot_sync(FULL, n == mx)) - 1u;
                }
                const uint32_t bh = __shfl_sync(FULL, h, pick), bn = min(__shfl_sync(FULL, n, pick), 32u);
                uint32_t got = 0;
                if (lane == 0) got = atomicCAS(&sm.head[pick], bh, bh + bn) == bh ? 1u : 0u;
                if (__shfl_sync(FULL, got, 0)) { ck = pick; cbase = bh; cn = bn; break; }
                ++tries;
            }
        }
        if (ck == WF_NOKEY) break;
        const bool valid = lane < cn;
        uint32_t i = 0;
        if (valid) {
            volatile uint16_t* cell = &sm.ring[ck][(cbase + lane) & (WF_POOL - 1u)];
            uint32_t v;
            while ((v = *cell) == WF_EMPTY) {}                  // reserved by a producer that has not written it yet
            *cell = (uint16_t)WF_EMPTY;
            i = v;
        }
        __threadfence_block();                                  // the slot's state was published before its cell
        // ---- load the slot ----
        PathState<R> p;
        uint32_t fl = 0, sidx = 0, pxy = 0, prim_bits = PRIM_SKY;
        uint64_t accepted = 0;
        p.o = V3<R>(0, 0, 0); p.d = V3<R>(0, 0, 1); p.thr = V3<R>(0, 0, 0); p.rad = V3<R>(0, 0, 0); p.hit_dist = R(-1); p.prev_pdf = 0; p.bounce = 0;
        if (valid) {
            const float4 q0 = sm.ro[i], q1 = sm.rd[i], q2 = sm.tr[i], q3 = sm.ra[i];
            p.o = V3<R>(q0.x, q0.y, q0.z); p.hit_dist = q0.w;
            p.d = V3<R>(q1.x, q1.y, q1.z); pxy = __float_as_uint(q1.w);
            p.thr = V3<R>(q2.x, q2.y, q2.z); fl = __float_as_uint(q2.w);
            p.rad = V3<R>(q3.x, q3.y, q3.z); sidx = __float_as_uint(q3.w);
            prim_bits = __float_as_uint(sm.ac[i].w);
            if constexpr (!RM) accepted = (uint64_t)sm.acc_lo[i] | ((uint64_t)sm.acc_hi[i] << 32);
        }
        p.bounce = (fl >> 8) & 0xffffu;
        bool alive = valid && (fl & FL_ALIVE);
        bool have_pixel = fl & FL_PIXEL;
        uint32_t pix = (pxy >> 16) * a.W + (pxy & 0xffffu);

        // ================================ A: the path's pending event ================================
        if (alive) {
            if (prim_bits == PRIM_SKY) {                            // the path left the scene (tracer.rs:66-69)
                path_add_sky(s, p);
                if (COUNT) pc.end_sky++;
                alive = false;
            } else {
                Rng<R> rng(pix, a.sample_base + sidx, a.seed);
                R u[8];
                bool cont;
                if constexpr (RM) {
                    const int prim = (int)(prim_bits & 0xffffu);
                    const RMat& rm = rm_lookup(s, sv, rm_keys, rm_table, rm_key_of(s, sv, prim, prim_bits >> 16), p.d);
                    shade_draws(rng, p.bounce, s.n_lights > 1u || (rm.lobe_class & 4u) != 0u, u);
                    const V3<R> normal = hit_normal<R, BVH>(s, sv, prim, p.o, p.d, p.hit_dist);
                    cont = path_shade_rm<COUNT>(s, sv, p, normal, rm, u, &pc);
                } else {
                    const int prim = (int)prim_bits;
                    Mat<R> mat;
                    hit_material<R, BVH>(s, sv, prim, accepted, p.d, mat);
                    shade_draws(rng, p.bounce, s.n_lights > 1u || (lobe_class_of(mat.metallic, mat.spec_trans, mat.clearcoat) & 4u) != 0u, u);
                    const V3<R> normal = hit_normal<R, BVH>(s, sv, prim, p.o, p.d, p.hit_dist);
                    cont = path_shade<R, COUNT, BVH>(s, sv, p, normal, mat, u, &pc);
                }
                alive = cont;
                if (cont && a.rr_start != 0 && p.bounce >= a.rr_start) {      // Russian-roulette extension (A.12), slot 0 of the new bounce
                    R u4[4];
                    rng.block(p.bounce, 0, u4);
                    if (!russian_roulette_survives(p, u4[0])) { alive = false; if (COUNT) pc.end_rr++; }
                }
            }
        }

        // ================================ B: finish dead paths, regenerate in place ================================
        // Work items below a.n_whole are whole pixels (all spp samples); the items above are the frame's last a.tail_zt
        // pixels cut into sample blocks, so that the ramp-down at the end of the launch lasts one block, not one pixel.
        // A block's sum goes to a side buffer and k_tail_combine adds the blocks of a pixel in block order: the image
        // stays independent of which slot traced what.
        bool done = false;
        uint32_t blkbits = fl & FL_BLOCK_BITS;
        const unsigned dead_mask = __ballot_sync(FULL, valid && !alive);
        // in place when at least a quarter of the chunk needs it (always in WF_MISS / WF_REGEN chunks); else the dead lanes queue up
        // under WF_REGEN and are regenerated by a full warp later
        const bool regen_here = __popc(dead_mask) * 4u >= cn;
        if (regen_here) {
            V3<R> acc(0, 0, 0);
            const bool dead = valid && !alive;
            if (dead) {
                const float4 q4 = sm.ac[i];
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
                    reinterpret_cast<float4*>(a.tail_side)[(pidx - a.n_whole) + blk * a.tail_zt] = make_float4(acc.x, acc.y, acc.z, (R)(s_end - s_begin));
                } else if (a.flush_dst) {       // multi-GPU: this launch's partial sum goes straight into the root GPU's slot (peer store)
                    reinterpret_cast<float4*>(a.flush_dst)[pix] = make_float4(acc.x, acc.y, acc.z, (R)a.spp);
                } else {
                    float4 v = accum[pix];
                    v.x += acc.x; v.y += acc.y; v.z += acc.z; v.w += (R)a.spp;
                    accum[pix] = v;
                }
                have_pixel = false;
            }
            // Hand-out: the lanes that want a pixel reserve exactly as many work items as they are (one warp-aggregated atomic on
            // the global counter).  Nothing is reserved ahead: an item held back by a warp that then finds no more work would be a
            // lost pixel.
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
                sm.ac[i] = make_float4(acc.x, acc.y, acc.z, __uint_as_float(PRIM_SKY));
                p.rad = V3<R>(0, 0, 0);
                if (!done) {                                        // next sample of the slot's pixel
                    Rng<R> rng(pix, a.sample_base + sidx, a.seed);
                    R u4[4];
                    rng.block(0, 0, u4);
                    path_begin(s, p, pxy & 0xffffu, pxy >> 16, a.W, a.H, inv_w, inv_h, u4[0], u4[1], a.film_fast != 0u, a.rcp_w, a.rcp_h);
                    alive = true;
                    if (COUNT) n_samples++;
                }
            }
        }

        // ================================ C: closest_hit for every live path ================================
        uint32_t key = done || !valid ? WF_NOKEY : WF_REGEN;
        if (alive) {
            uint32_t new_prim = PRIM_SKY;
            if (p.bounce >= s.depth) {                              // recursion depth 0 (tracer.rs:61)
                alive = false;
                if (COUNT) pc.end_depth++;
            } else {
                if (COUNT) pc.closest_hit++;
                const HitCore<R> h = closest_hit_core<R, BVH>(s, sv, p.o, p.d, p.hit_dist);
                p.hit_dist = h.hit_dist;
                if (!h.hit) {
                    key = WF_MISS;                                 // background lookup with full warps
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
            sm.ro[i] = make_float4(p.o.x, p.o.y, p.o.z, p.hit_dist);
            sm.rd[i] = make_float4(p.d.x, p.d.y, p.d.z, __uint_as_float(pxy));
            reinterpret_cast<uint32_t*>(&sm.ac[i])[3] = new_prim;
        }
        if (valid) {
            fl = (p.bounce << 8) | blkbits | (alive ? FL_ALIVE : 0u) | (have_pixel ? FL_PIXEL : 0u);
            sm.tr[i] = make_float4(p.thr.x, p.thr.y, p.thr.z, __uint_as_float(fl));
            sm.ra[i] = make_float4(p.rad.x, p.rad.y, p.rad.z, __uint_as_float(sidx));
        }
        __threadfence_block();                                      // state before cells

        // ================================ hand the slots on: one reservation per distinct key ================================
        unsigned todo = __ballot_sync(FULL, key != WF_NOKEY);
        while (todo) {
            const uint32_t kk = __shfl_sync(FULL, key, __ffs(todo) - 1);
            const unsigned grp = __ballot_sync(FULL, key == kk);
            uint32_t t = 0;
            if (lane == (uint32_t)(__ffs(grp) - 1)) t = atomicAdd(&sm.tail[kk], (uint32_t)__popc(grp));
            t = __shfl_sync(FULL, t, __ffs(grp) - 1);
            if (key == kk) {
                volatile uint16_t* cell = &sm.ring[kk][(t + __popc(grp & lt_mask)) & (WF_POOL - 1u)];
                while (*cell != WF_EMPTY) {}                        // (its previous occupant has been reserved but not read yet)
                *cell = (uint16_t)i;
            }
            todo &= ~grp;
        }
        const unsigned gone = __ballot_sync(FULL, done);
        if (gone && lane == 0) atomicSub(&sm.n_live, (uint32_t)__popc(gone));
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
__global__ void k_tail_combine(const RenderArgs a) {
    const uint32_t z = blockIdx.x * blockDim.x + threadIdx.x;
    if (z >= a.tail_zt) return;
    const uint32_t idx = a.n_whole + z, tile = idx >> 8, within = idx & 255u;
    const uint32_t px = (tile % a.tiles_x) * 16u + (within & 15u), prow = (tile / a.tiles_x) * 16u + (within >> 4);
    if (px >= a.W || prow >= a.H) return;
    const float4* side = reinterpret_cast<const float4*>(a.tail_side);
    float4 sum = side[z];
    for (uint32_t b = 1; b < (1u << a.tail_log2b); ++b) {
        const float4 v = side[z + b * a.tail_zt];
        sum.x += v.x; sum.y += v.y; sum.z += v.z; sum.w += v.w;
    }
    const uint32_t pix = prow * a.W + px;
    if (a.flush_dst) {
        reinterpret_cast<float4*>(a.flush_dst)[pix] = sum;
    } else {
        float4* accum = reinterpret_cast<float4*>(a.accum);
        float4 v = accum[pix];
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
inline int wavefront_render(WavefrontState& wf, const DScene<float>& d, void* accum, void* flush_dst, uint32_t W, uint32_t H, uint32_t spp, uint64_t sample_base,
                            const ptb_config& cfg, cudaStream_t stream, int sm_count, DeviceCounters* counters, unsigned int* work_counter,
                            cudaEvent_t ev0, cudaEvent_t ev1, uint64_t* launches, std::string& err) {
    RenderArgs a{};
    a.accum = accum; a.flush_dst = flush_dst; a.W = W; a.H = H; a.spp = spp; a.sample_base = sample_base; a.seed = cfg.seed; a.rr_start = cfg.rr_start;
    a.tiles_x = (W + 15u) / 16u;
    a.n_items = a.tiles_x * ((H + 15u) / 16u) * 256u;
    a.work_counter = work_counter;
    a.counters = counters;
    if (wf.film_w != W || wf.film_h != H) { wf.film_fast = film_coords_fma_exact(W, H); wf.film_w = W; wf.film_h = H; }
    a.film_fast = wf.film_fast ? 1u : 0u;
#ifdef PTB_NO_FILM_FMA
    a.film_fast = 0u;
#endif
    a.rcp_w = 1.0f / (float)W; a.rcp_h = 1.0f / (float)H;
    const bool count = cfg.collect_counters != 0;
    const bool rm = d.rm_entries != 0 && !d.use_bvh;
    void (*kern)(const DScene<float>, const RenderArgs) =
        d.use_bvh ? (count ? k_render_wavefront<true, true, false> : k_render_wavefront<false, true, false>)
        : rm      ? (count ? k_render_wavefront<true, false, true> : k_render_wavefront<false, false, true>)
                  : (count ? k_render_wavefront<true, false, false> : k_render_wavefront<false, false, false>);
    const size_t smem_bytes = rm ? sizeof(WfSmemT<WF_POOL_RM, WF_SCENE_BYTES_RM, false>) : sizeof(WfSmemT<WF_POOL_GENERIC, WF_SCENE_BYTES_GENERIC, true>);
    const uint32_t WF_POOL = rm ? WF_POOL_RM : WF_POOL_GENERIC;
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
        const size_t need = ((size_t)zt << log2b) * sizeof(float4);
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
    kern<<<grid, rm ? WF_THREADS_RM : WF_THREADS_GENERIC, smem_bytes, stream>>>(d, a);
    if ((e = cudaGetLastError()) == cudaSuccess && a.tail_zt) {
        k_tail_combine<<<(a.tail_zt + 255u) / 256u, 256, 0, stream>>>(a);
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
