// ptb_stream.cuh — the global-memory ("streaming") wavefront integrator: one kernel per kind of work
// over SoA ray / path-state queues in HBM, compacted with warp ballot/popc, persistent warps that pull
// 32-entry chunks from the queues.
//
// The shared-memory wavefront (ptb_wavefront.cuh) keeps a pool of 2048 paths per SM; that is the right
// shape while the scene fits next to the pool and a bounce costs ~1 kFLOP.  Large BVH scenes (BASELINE
// configs 4 and 5) are bound by the latency of dependent node fetches instead, and want (a) far more
// resident warps on the traversal than a 126-register shading kernel allows and (b) shading warps that
// are full and lobe-coherent although neighbouring rays hit unrelated materials.  Separate kernels give
// both: the traversal kernels carry only a ray and a stack, the shade kernel runs at its own register
// budget over per-lobe-class queues.
//
// A WAVE is a set of P paths (a range of pixels x S consecutive samples each, at most `wave_paths`),
// carried through the bounces together; every path owns one slot of the SoA state for the whole wave:
//
//   k_stream_generate   camera rays (tracer.rs:34-59)                                      -> ray queue 0
//   per bounce b:
//     k_stream_closest  Scene::closest_hit (+ sample_lights) per queued ray; emitter hits are finished
//                       here; everything else is queued by key = lobe class of the hit material, or
//                       ST_MISS                                                            -> shade queues
//     k_stream_shade    per key, full warps: background for ST_MISS; else finalize, light sample,
//                       Disney eval towards the light (contribution DEFERRED), Disney sample,
//                       throughput, next ray                                  -> ray queue b+1, shadow queue
//     k_stream_shadow   Scene::any_hit per queued shadow ray; an unoccluded ray adds its deferred
//                       contribution to the path's radiance (tracer.rs:150-164)
//   k_stream_accumulate one thread per pixel sums its S radiances in sample order into the accumulator
//
// Queue lengths live in device memory (one BounceCtr per bounce), the host never reads them: every stage
// is launched with a persistent grid whose warps pull chunks until the queue is drained.  Path results do
// not depend on queue order and every path / pixel is summed by exactly one thread, so the image is
// bit-reproducible run to run like the other two integrators'.
//
// HBM traffic per path-bounce: closest reads 32 B of ray and writes a 16 B hit record and a 4 B queue
// entry; shade reads 80 B and writes 48 B of state, 4 B of queue and a 48 B shadow entry; shadow reads
// 48 B and read-modify-writes 32 B — about 310 B, against the 32 B per PIXEL of the on-chip integrators.
// That is what caps this design near 6 Gsamples/s on the demo scene and why AUTO uses it
// only where traversal latency, not arithmetic, is the bound.
#pragma once
#include <string>

#include "ptb_kernels.cuh"

namespace ptb {

constexpr int ST_THREADS = 256;
constexpr uint32_t ST_KEYS = 9;          // 8 lobe classes (lobe_class_of) + ST_MISS
constexpr uint32_t ST_MISS = 8;          // the path left the scene: background lookup (tracer.rs:66-69)
constexpr uint32_t ST_DEFAULT_WAVE = 1u << 23;
#ifndef PTB_ST_SPLIT_MIN
#define PTB_ST_SPLIT_MIN 16384u
#endif
constexpr uint32_t ST_SPLIT_MIN_SPHERES = PTB_ST_SPLIT_MIN;   // BVH scenes above this use the persistent-lane traversal kernels from bounce 1 on
inline bool stream_uses_split(const DScene<float>& d) { return d.use_bvh && d.n_spheres > ST_SPLIT_MIN_SPHERES; }

struct BounceCtr {                       // 128 bytes per bounce, zeroed at the start of a wave
    uint32_t n_ray;                      // rays entering this bounce
    uint32_t cur_ray;                    // chunk cursor of k_stream_closest
    uint32_t n_shade[ST_KEYS];           // shade-queue length per key
    uint32_t cur_shade;                  // chunk cursor of k_stream_shade
    uint32_t n_shadow;                   // shadow rays of this bounce
    uint32_t cur_shadow;                 // chunk cursor of k_stream_shadow
    uint32_t cur_finish;                 // chunk cursor of k_stream_finish (BVH scenes)
    uint32_t pad[32 - 15];
};
static_assert(sizeof(BounceCtr) == 128, "BounceCtr layout");

struct StreamArgs {
    // path state, one slot per path of the wave (SoA of float4)
    float4* a0;          // o.x o.y o.z d.x
    float4* a1;          // d.y d.z hit_dist prev_pdf
    float4* a2;          // throughput.xyz, sample index within the wave (bits)
    float4* a3;          // radiance.xyz, pixel index (bits; 0xffffffff = outside the frame)
    uint4* hit;          // prim, accepted lo, accepted hi, hit_dist (bits)
    uint32_t* rayq[2];   // ray queues (slot indices), ping-pong by bounce parity
    uint32_t* shadeq;    // ST_KEYS queues of `cap` slot indices
    float4* s0;          // shadow queue: scatter_pos.xyz, max_dist
    float4* s1;          //               direction.xyz, slot (bits)
    float4* s2;          //               deferred contribution.xyz, flags (bits 0..3 lobes evaluated, bit 4 pdf > 0)
    BounceCtr* ctr;
    float4* accum;
    float4* flush_dst;   // multi-GPU: partial sums of this call go here (the root GPU's slot) instead of into accum
    uint32_t flush_store;// 1: the first wave that touches a pixel in this call stores, later waves add
    uint32_t W, H, tiles_x, n_items;
    uint32_t pix0, npix; // this wave's range of tiled pixel indices
    uint32_t S;          // samples per pixel in this wave
    uint32_t P;          // paths in this wave = npix * S
    uint32_t cap;        // stride of the shade queues
    uint64_t sample0;    // global index of the wave's first sample
    uint64_t seed;
    uint32_t rr_start;
    DeviceCounters* counters;
};

PTB_DEV void pc_clear(PathCounters& pc) {
    pc.closest_hit = pc.any_hit = pc.shade = pc.nee_contrib = pc.eval_calls = 0;
    pc.lobe[0] = pc.lobe[1] = pc.lobe[2] = pc.lobe[3] = 0;
    pc.end_sky = pc.end_emitter = pc.end_pdf = pc.end_depth = pc.end_rr = 0;
    pc.ev[0] = pc.ev[1] = pc.ev[2] = pc.ev[3] = 0;
    pc.bvh[0] = pc.bvh[1] = 0;
}
PTB_DEV void pc_flush(const PathCounters& pc, uint32_t n_samples, DeviceCounters* c) {
    auto add = [](unsigned long long* p, uint32_t v) { if (v) atomicAdd(p, (unsigned long long)v); };
    add(&c->samples, n_samples);
    add(&c->closest_hit, pc.closest_hit); add(&c->any_hit, pc.any_hit); add(&c->shade, pc.shade);
    add(&c->nee_contrib, pc.nee_contrib); add(&c->eval_calls, pc.eval_calls);
    for (int i = 0; i < 4; ++i) { add(&c->lobe[i], pc.lobe[i]); add(&c->ev[i], pc.ev[i]); }
    add(&c->end_sky, pc.end_sky); add(&c->end_emitter, pc.end_emitter); add(&c->end_pdf, pc.end_pdf);
    add(&c->end_depth, pc.end_depth); add(&c->end_rr, pc.end_rr);
    add(&c->bvh_nodes, pc.bvh[0]); add(&c->bvh_leaf_tests, pc.bvh[1]);
}

// warp-aggregated queue push for the lanes that are converged here (any subset of the warp): one atomic per group
PTB_DEV uint32_t queue_reserve(uint32_t* counter) {
    const unsigned m = __activemask();
    const uint32_t lane = threadIdx.x & 31u;
    const int leader = __ffs(m) - 1;
    uint32_t base = 0;
    if ((int)lane == leader) base = atomicAdd(counter, (uint32_t)__popc(m));
    base = __shfl_sync(m, base, leader);
    return base + (uint32_t)__popc(m & ((1u << lane) - 1u));
}

// A stage is launched with a persistent grid sized for a full queue; deep bounces leave it a few hundred rays.  CTAs that the
// queue cannot feed leave before they copy the scene into shared memory (the floor of a nearly empty stage was 30 us — 740 CTAs
// staging 12 KB each — against 5 us for the launch itself; config 5 runs 39 such stages per wave).
#ifdef PTB_ST_NO_EARLY_EXIT
#define PTB_ST_LEAVE_IF_IDLE(n_entries) do {} while (0)
#else
#define PTB_ST_LEAVE_IF_IDLE(n_entries) do { if ((uint64_t)blockIdx.x * (32u * (ST_THREADS / 32u)) >= (uint64_t)(n_entries)) return; } while (0)
#endif

// tiled pixel index -> pixel coordinates (16x16 tiles, as the other integrators hand pixels out)
PTB_DEV bool tiled_pixel(const StreamArgs& a, uint32_t ip, uint32_t& px, uint32_t& prow) {
    const uint32_t tile = ip >> 8, within = ip & 255u;
    px = (tile % a.tiles_x) * 16u + (within & 15u);
    prow = (tile / a.tiles_x) * 16u + (within >> 4);
    return px < a.W && prow < a.H;
}

// ------------------------------------------------------------------------------------------------
template <bool COUNT>
__global__ void __launch_bounds__(ST_THREADS) k_stream_generate(const __grid_constant__ DScene<float> s, const StreamArgs a) {
    using R = float;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const R inv_w = R(1) / (R)a.W, inv_h = R(1) / (R)a.H;
    bool live = false;
    if (i < a.P) {
        const uint32_t il = a.S == 1u ? i : i / a.S;
        const uint32_t sl = a.S == 1u ? 0u : i - il * a.S;
        uint32_t px, prow;
        live = tiled_pixel(a, a.pix0 + il, px, prow);
        if (live) {
            const uint32_t pix = prow * a.W + px;
            Rng<R> rng(pix, a.sample0 + sl, a.seed);
            R u4[4];
            rng.block(0, 0, u4);
            PathState<R> p;
            path_begin(s, p, px, prow, a.W, a.H, inv_w, inv_h, u4[0], u4[1]);
            a.a0[i] = make_float4(p.o.x, p.o.y, p.o.z, p.d.x);
            a.a1[i] = make_float4(p.d.y, p.d.z, p.hit_dist, 0.0f);
            a.a2[i] = make_float4(1.0f, 1.0f, 1.0f, __uint_as_float(sl));
            a.a3[i] = make_float4(0.0f, 0.0f, 0.0f, __uint_as_float(pix));
        } else {
            a.a3[i] = make_float4(0.0f, 0.0f, 0.0f, __uint_as_float(0xffffffffu));
        }
    }
    if (COUNT) {
        const unsigned m = __ballot_sync(0xffffffffu, live);
        if ((threadIdx.x & 31u) == 0 && m) {
            atomicAdd(&a.counters->samples, (unsigned long long)__popc(m));
            if (s.depth == 0) atomicAdd(&a.counters->end_depth, (unsigned long long)__popc(m));
        }
    }
    if (live && s.depth > 0) a.rayq[0][queue_reserve(&a.ctr[0].n_ray)] = i;      // recursion_depth() == 0: no bounce runs (tracer.rs:61)
}

// ------------------------------------------------------------------------------------------------
// What follows Scene::closest_hit for one ray (tracer.rs:66-87): a path that hit a light is finished here, everything
// else gets its queue key — the lobe class of the hit material, or ST_MISS.  Returns the key, 0xff for "no shading".
template <bool COUNT, bool BVH>
PTB_DEV uint32_t stream_after_hit(const DScene<float>& s, const SceneView<float>& sv, const StreamArgs& a, uint32_t slot, V3<float> d, float prev_pdf,
                                  const HitCore<float>& h, PathCounters& pc) {
    using R = float;
    if (!h.hit) return ST_MISS;
    if (h.is_emitter) {
        const float4 A2 = a.a2[slot];
        float4 A3 = a.a3[slot];
        PathState<R> p;
        p.d = d; p.prev_pdf = prev_pdf;
        p.thr = V3<R>(A2.x, A2.y, A2.z);
        p.rad = V3<R>(A3.x, A3.y, A3.z);
        path_add_emitter<R, BVH>(s, sv, p, h);
        A3.x = p.rad.x; A3.y = p.rad.y; A3.z = p.rad.z;
        a.a3[slot] = A3;
        if (COUNT) pc.end_emitter++;
        return 0xffu;
    }
    a.hit[slot] = make_uint4((uint32_t)h.prim, (uint32_t)h.accepted, (uint32_t)(h.accepted >> 32), __float_as_uint(h.hit_dist));
    return hit_lobe_class<R, BVH>(s, sv, h.prim, h.accepted);
}
// queue `slot` by key, one atomic per (warp, key); every lane of the warp calls this (key 0xff = nothing to queue)
PTB_DEV void push_by_key(const StreamArgs& a, BounceCtr& ctr, uint32_t key, uint32_t slot) {
    const uint32_t lane = threadIdx.x & 31u;
    const unsigned peers = __match_any_sync(0xffffffffu, key);
    if (key != 0xffu) {
        const int leader = __ffs(peers) - 1;
        uint32_t base = 0;
        if ((int)lane == leader) base = atomicAdd(&ctr.n_shade[key], (uint32_t)__popc(peers));
        base = __shfl_sync(peers, base, leader);
        a.shadeq[(size_t)key * a.cap + base + (uint32_t)__popc(peers & ((1u << lane) - 1u))] = slot;
    }
}

// The whole of Scene::closest_hit in one kernel, one ray per lane per 32-ray chunk: small scenes (no BVH), and BVH scenes
// while the rays of a chunk are coherent (camera rays) or the tree is shallow.
// FX: scenes with media or live extended lights (the quad tests are compiled into the FX instantiations only)
template <bool COUNT, bool BVH, bool FX>
__global__ void __launch_bounds__(ST_THREADS) k_stream_closest(const __grid_constant__ DScene<float> s, const StreamArgs a, const uint32_t bounce) {
    using R = float;
    PTB_ST_LEAVE_IF_IDLE(a.ctr[bounce].n_ray);
    __shared__ SceneSmem<R> sm;
    const SceneView<R> sv = stage_scene(s, sm.words, PTB_SMEM_SCENE_BYTES);
    BounceCtr& ctr = a.ctr[bounce];
    const uint32_t n = ctr.n_ray;
    const uint32_t* __restrict__ q = a.rayq[bounce & 1u];
    const uint32_t lane = threadIdx.x & 31u;
    PathCounters pc;
    if (COUNT) pc_clear(pc);
    while (true) {
        uint32_t chunk = 0;
        if (lane == 0) chunk = atomicAdd(&ctr.cur_ray, 32u);
        chunk = __shfl_sync(0xffffffffu, chunk, 0);
        if (chunk >= n) break;
        const uint32_t j = chunk + lane;
        uint32_t key = 0xffu, slot = 0;
        if (j < n) {
            slot = q[j];
            const float4 A0 = a.a0[slot], A1 = a.a1[slot];
            const V3<R> o(A0.x, A0.y, A0.z), d(A0.w, A1.x, A1.y);
            if (COUNT) pc.closest_hit++;
            const HitCore<R> h = closest_hit_core<R, BVH, false, FX>(s, sv, o, d, A1.z, COUNT ? pc.bvh : nullptr);       // (no signed-distance programs in this integrator)
            key = stream_after_hit<COUNT, BVH>(s, sv, a, slot, d, A1.w, h, pc);
        }
        push_by_key(a, ctr, key, slot);
    }
    if (COUNT) pc_flush(pc, 0, a.counters);
}

// ------------------------------------------------------------------------------------------------
// BVH scenes: sphere-BVH traversal in a kernel of its own, for closest-hit rays (ANY = false: reads the ray queue, writes
// (sphere index, distance) into the hit records) and for shadow rays (ANY = true: reads the shadow queue and adds the
// deferred next-event contribution of every unoccluded ray).
//
// ncu on the first version (one ray per lane per 32-ray chunk, profiles/r01_ncu_stream.md): 77 % of the issue slots busy
// at 9.5 of 32 lanes — the traversal is bound by SIMT divergence, not by memory latency: the rays of a warp need very
// different numbers of steps and the warp runs for the longest.  So lanes are PERSISTENT here: a lane whose ray is done
// writes its result and, as soon as ST_REFILL lanes of the warp are idle, the idle lanes pull new rays from the queue
// (one warp-aggregated atomic) while the others keep their traversal state.
//
// Second finding (same file): with refill the box tests run at 26 lanes, but the leaf sphere tests and the stack pops that
// follow them ran at 3-4 lanes and made up a third of the issue slots.  So a leaf is POSTPONED (Aila & Laine's speculative
// while-while): the lane parks the leaf reference, pops the next node and keeps traversing; parked leaves are intersected
// together once ST_LEAF_MIN lanes hold one, or no lane can advance otherwise.  Node references carry the leaf count in
// their low 3 bits, so neither the stack pop nor the parked leaf needs to touch the node array again.
#ifndef PTB_ST_REFILL
#define PTB_ST_REFILL 12
#endif
#ifndef PTB_ST_LEAF_MIN
#define PTB_ST_LEAF_MIN 8
#endif
constexpr int ST_REFILL = PTB_ST_REFILL;
constexpr int ST_LEAF_MIN = PTB_ST_LEAF_MIN;
constexpr int ST_STACK = 40;
constexpr uint32_t ST_NONE = 0xffffffffu;

PTB_DEV uint32_t node_ref(uint32_t left_or_first, uint32_t count) { return (left_or_first << 3) | count; }   // count <= 7 (builder: <= 4)

// slab test with the origin folded in: t = b * (1/d) - o * (1/d), one FMA per plane.  Conservative like box_entry: entry and
// exit are widened by 1e-4 relative; a NaN slab (d == 0 on that axis: inf - inf) is ignored by min/max, i.e. treated as hit.
struct RayS { float idx, idy, idz, nox, noy, noz; };
PTB_DEV float box_entry_s(float4 lo, float4 hi, const RayS& r, float limit) {
    const float tx0 = fmaf(lo.x, r.idx, r.nox), tx1 = fmaf(hi.x, r.idx, r.nox);
    const float ty0 = fmaf(lo.y, r.idy, r.noy), ty1 = fmaf(hi.y, r.idy, r.noy);
    const float tz0 = fmaf(lo.z, r.idz, r.noz), tz1 = fmaf(hi.z, r.idz, r.noz);
    float tn = fmaxf(fmaxf(fminf(tx0, tx1), fminf(ty0, ty1)), fmaxf(fminf(tz0, tz1), 0.0f));
    const float tf = fminf(fminf(fmaxf(tx0, tx1), fmaxf(ty0, ty1)), fmaxf(tz0, tz1));
    tn *= 0.9999f;
    return (tn <= tf * 1.0001f && tn <= limit) ? tn : 3.0e38f;
}

PTB_DEV bool box_test_s(float4 lo, float4 hi, const RayS& r, float limit, float& tn_out) {
    const float tx0 = fmaf(lo.x, r.idx, r.nox), tx1 = fmaf(hi.x, r.idx, r.nox);
    const float ty0 = fmaf(lo.y, r.idy, r.noy), ty1 = fmaf(hi.y, r.idy, r.noy);
    const float tz0 = fmaf(lo.z, r.idz, r.noz), tz1 = fmaf(hi.z, r.idz, r.noz);
    float tn = fmaxf(fmaxf(fminf(tx0, tx1), fminf(ty0, ty1)), fmaxf(fminf(tz0, tz1), 0.0f));
    const float tf = fminf(fminf(fmaxf(tx0, tx1), fmaxf(ty0, ty1)), fmaxf(tz0, tz1));
    tn *= 0.9999f;
    tn_out = tn;
    return tn <= tf * 1.0001f && tn <= limit;
}

#ifndef PTB_ST_INNER_REPS
#define PTB_ST_INNER_REPS 4
#endif
// (an explicit minimum of 1 block makes nvcc budget 255 registers: the bounds below name a minimum only when a knob sets one)
#ifdef PTB_ST_SHADE_MIN_BLOCKS
#define PTB_ST_SHADE_BOUNDS __launch_bounds__(ST_THREADS, PTB_ST_SHADE_MIN_BLOCKS)
#else
#define PTB_ST_SHADE_BOUNDS __launch_bounds__(ST_THREADS)
#endif
#ifndef PTB_ST_TRACE_MIN_BLOCKS
#define PTB_ST_TRACE_MIN_BLOCKS 5
#endif
#ifdef PTB_ST_TRACE_PLAIN
#define PTB_ST_TRACE_BOUNDS __launch_bounds__(ST_THREADS)
#else
#define PTB_ST_TRACE_BOUNDS __launch_bounds__(ST_THREADS, PTB_ST_TRACE_MIN_BLOCKS)
#endif
template <bool ANY, bool COUNT>
__global__ void PTB_ST_TRACE_BOUNDS k_stream_trace(const __grid_constant__ DScene<float> s, const StreamArgs a, const uint32_t bounce) {
    using R = float;
    PTB_ST_LEAVE_IF_IDLE(ANY ? a.ctr[bounce].n_shadow : a.ctr[bounce].n_ray);
    __shared__ SceneSmem<R> sm;
    const SceneView<R> sv = stage_scene(s, sm.words, PTB_SMEM_SCENE_BYTES);       // planes (shadow rays test them first)
    BounceCtr& ctr = a.ctr[bounce];
    const uint32_t n = ANY ? ctr.n_shadow : ctr.n_ray;
    uint32_t* const cursor = ANY ? &ctr.cur_shadow : &ctr.cur_ray;
    const uint32_t* __restrict__ q = a.rayq[bounce & 1u];
    const float4* __restrict__ nodes = reinterpret_cast<const float4*>(s.bvh);    // node i = nodes[2i] (lo, link), nodes[2i+1] (hi, count)
    const DSphere<R>* __restrict__ leaf_spheres = s.bvh_spheres;
    const uint32_t* __restrict__ leaf_prim = s.bvh_prim;
    const uint32_t lane = threadIdx.x & 31u;
    const unsigned FULL = 0xffffffffu;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const bool ignore_max = (s.flags & PTB_SCENE_ANYHIT_IGNORES_MAX_DIST) != 0;
    const uint32_t root_ref = node_ref(__float_as_uint(__ldg(nodes).w), __float_as_uint(__ldg(nodes + 1).w));
    PathCounters pc;
    if (COUNT) pc_clear(pc);

    bool active = false, exhausted = false;
    uint32_t slot = 0;                       // closest: path slot; any: shadow-queue index
    V3<R> o(0, 0, 0), d(0, 0, 1);
    RayS r{};
    R best_t = 0;
    int best = -1;
    uint32_t cur = ST_NONE;                  // node to visit next (inner or leaf reference)
    uint32_t pend = 0;                       // parked leaf reference (0 = none)
    int sp = 0;
    uint2 stack[ST_STACK];                   // (node reference, entry distance bits): one 8-byte local load per pop

#ifndef PTB_ST_EAGER_FINISH
    bool fin = false;                        // the lane's ray is done, its result not yet written (written with the next refill)
#endif
    while (true) {
        // ---- refill idle lanes
        const unsigned idle = __ballot_sync(FULL, !active);
        if (idle) {
#ifndef PTB_ST_EAGER_FINISH
            // results are written by all idle lanes together, right before they are refilled (or at the very end): the write-back
            // used to run for one or two lanes in almost every iteration
            if ((!exhausted && __popc(idle) >= ST_REFILL) || idle == FULL) {
                if (fin) {
                    fin = false;
                    if (ANY) {
                        if (best < 0) {                                       // unoccluded: tracer.rs:162-164
                            const float4 S1 = a.s1[slot], S2 = a.s2[slot];
                            const uint32_t flags = __float_as_uint(S2.w);
                            if (flags & 16u) {
                                const uint32_t ps = __float_as_uint(S1.w);
                                float4 A3 = a.a3[ps];
                                A3.x += S2.x; A3.y += S2.y; A3.z += S2.z;
                                a.a3[ps] = A3;
                                if (COUNT) pc.nee_contrib++;
                            }
                            if (COUNT && !(flags & 32u)) {
                                pc.eval_calls++;
                                for (int k = 0; k < 4; ++k) pc.ev[k] += (flags >> k) & 1u;
                            }
                        }
                    } else {
                        a.hit[slot] = make_uint4((uint32_t)best, 0u, 0u, __float_as_uint(best_t));
                    }
                }
            }
#endif
            if (!exhausted && (__popc(idle) >= ST_REFILL || idle == FULL)) {
                const uint32_t cnt = (uint32_t)__popc(idle);
                uint32_t base = 0;
                if (lane == 0) base = atomicAdd(cursor, cnt);
                base = __shfl_sync(FULL, base, 0);
                if (base + cnt >= n) exhausted = true;
                const uint32_t my = base + (uint32_t)__popc(idle & lt_mask);
                if (!active && my < n) {
                    bool go = true;
                    if (ANY) {
                        slot = my;
                        const float4 S0 = a.s0[my], S1 = a.s1[my];
                        o = V3<R>(S0.x, S0.y, S0.z); d = V3<R>(S1.x, S1.y, S1.z);
                        best_t = ignore_max ? Const<R>::MAXV : S0.w;
                        if (any_hit_planes<R, false>(s, sv, o, d, S0.w)) go = false;            // occluded by a plane: nothing to add
                    } else {
                        slot = q[my];
                        const float4 A0 = a.a0[slot], A1 = a.a1[slot];
                        o = V3<R>(A0.x, A0.y, A0.z); d = V3<R>(A0.w, A1.x, A1.y);
                        best_t = Const<R>::MAXV;
                        if (COUNT) pc.closest_hit++;
                    }
                    if (go) {
                        best = -1;
                        // (a direction component of exactly 0 would make o * inf - b * inf a NaN or a wrongly signed infinity)
                        r.idx = 1.0f / (fabsf(d.x) > 1e-30f ? d.x : copysignf(1e-30f, d.x));
                        r.idy = 1.0f / (fabsf(d.y) > 1e-30f ? d.y : copysignf(1e-30f, d.y));
                        r.idz = 1.0f / (fabsf(d.z) > 1e-30f ? d.z : copysignf(1e-30f, d.z));
                        r.nox = -o.x * r.idx; r.noy = -o.y * r.idy; r.noz = -o.z * r.idz;
                        cur = root_ref; pend = 0; sp = 0;
                        active = true;
                    }
                }
            } else if (idle == FULL) {
                break;
            }
        }
        // ---- one inner node: both children with one 64-byte read, nearer first.  PTB_ST_INNER_REPS > 1: that many traversal steps
        //      (node + park / pop) per iteration, so that the ballots, the leaf batch test, the write-back and the refill test
        //      around them are paid once per several node visits
#pragma unroll
        for (int rep = 0; rep < PTB_ST_INNER_REPS; ++rep) {
        if (active && cur != ST_NONE && (cur & 7u) == 0u) {
            if (COUNT) pc.bvh[0]++;
            const float4* c = nodes + (size_t)(cur >> 3) * 2u;
            const float4 alo = __ldg(c), ahi = __ldg(c + 1), blo = __ldg(c + 2), bhi = __ldg(c + 3);
#ifdef PTB_ST_BOX_SENTINEL
            const float ta = box_entry_s(alo, ahi, r, best_t), tb = box_entry_s(blo, bhi, r, best_t);
            const bool ha = ta < 3.0e38f, hb = tb < 3.0e38f;
#else
            // (hit flag and entry distance separately: the 3e38 sentinel cost four selects and two compares per visit on the ALU
            //  pipe, which is what this kernel is short of)
            float ta, tb;
            const bool ha = box_test_s(alo, ahi, r, best_t, ta), hb = box_test_s(blo, bhi, r, best_t, tb);
#endif
            const uint32_t ra = node_ref(__float_as_uint(alo.w), __float_as_uint(ahi.w)), rb = node_ref(__float_as_uint(blo.w), __float_as_uint(bhi.w));
            const bool a_near = ha && (!hb || ta <= tb);
            const uint32_t near_ref = a_near ? ra : rb;
            if (ha && hb && sp < ST_STACK) {
                // (far child written as ra ^ rb ^ near and max(ta, tb): nvcc 12.9 compiled the mirrored select
                //  `a_near ? rb : ra` next to `a_near ? ra : rb` into an unconditional store of rb — found with a per-ray
                //  differential check against bvh_traverse)
                stack[sp] = make_uint2(ra ^ rb ^ near_ref, __float_as_uint(fmaxf(ta, tb)));
                ++sp;
            }
            cur = (ha || hb) ? near_ref : ST_NONE;
        }
        // ---- park a leaf, pop the next node
        if (active) {
            if (cur != ST_NONE && (cur & 7u) != 0u && pend == 0u) { pend = cur; cur = ST_NONE; }
            if (cur == ST_NONE) {
                while (sp > 0) {
                    --sp;
                    const uint2 e = stack[sp];
                    if (__uint_as_float(e.y) <= best_t) { cur = e.x; break; }
                }
            }
        }
        }
        // ---- parked leaves, together
        const unsigned m_pend = __ballot_sync(FULL, active && pend != 0u);
        const unsigned m_inner = __ballot_sync(FULL, active && cur != ST_NONE && (cur & 7u) == 0u);
        bool finished = false;
        if (m_pend && (__popc(m_pend) >= ST_LEAF_MIN || m_inner == 0u)) {
            if (active && pend != 0u) {
                const uint32_t first = pend >> 3, cnt = pend & 7u;
                pend = 0u;
                if (COUNT) pc.bvh[1] += cnt;
                for (uint32_t i = 0; i < cnt; ++i) {
                    const DSphere<R> sph = leaf_spheres[first + i];
                    const R t = isect_sphere(o, d, V3<R>(sph.cx, sph.cy, sph.cz), sph.r);
                    if (t >= R(0)) {
                        if (ANY) { if (t < best_t) { best = 0; finished = true; } }
                        else {
                            const int si = (int)leaf_prim[first + i];
                            if (t < best_t || (t == best_t && si < best)) { best_t = t; best = si; }
                        }
                    }
                }
            }
        }
        if (active && cur == ST_NONE && pend == 0u && sp == 0) finished = true;
#ifndef PTB_ST_EAGER_FINISH
        if (active && finished) { active = false; fin = true; }
#else
        if (active && finished) {
            active = false;
            if (ANY) {
                if (best < 0) {                                       // unoccluded: tracer.rs:162-164
                    const float4 S1 = a.s1[slot], S2 = a.s2[slot];
                    const uint32_t flags = __float_as_uint(S2.w);
                    if (flags & 16u) {
                        const uint32_t ps = __float_as_uint(S1.w);
                        float4 A3 = a.a3[ps];
                        A3.x += S2.x; A3.y += S2.y; A3.z += S2.z;
                        a.a3[ps] = A3;
                        if (COUNT) pc.nee_contrib++;
                    }
                    if (COUNT && !(flags & 32u)) {
                        pc.eval_calls++;
                        for (int k = 0; k < 4; ++k) pc.ev[k] += (flags >> k) & 1u;
                    }
                }
            } else {
                a.hit[slot] = make_uint4((uint32_t)best, 0u, 0u, __float_as_uint(best_t));
            }
        }
#endif
    }
    if (COUNT) pc_flush(pc, 0, a.counters);
}

// BVH scenes: the rest of Scene::closest_hit after the sphere traversal — planes, Scene::sample_lights (light BVH when
// there are many lights), emitter hits, queue keys.  Full, coherent warps.
template <bool COUNT, bool FX>
__global__ void __launch_bounds__(ST_THREADS) k_stream_finish(const __grid_constant__ DScene<float> s, const StreamArgs a, const uint32_t bounce) {
    using R = float;
    PTB_ST_LEAVE_IF_IDLE(a.ctr[bounce].n_ray);
    __shared__ SceneSmem<R> sm;
    const SceneView<R> sv = stage_scene(s, sm.words, PTB_SMEM_SCENE_BYTES);
    BounceCtr& ctr = a.ctr[bounce];
    const uint32_t n = ctr.n_ray;
    const uint32_t* __restrict__ q = a.rayq[bounce & 1u];
    const uint32_t lane = threadIdx.x & 31u;
    PathCounters pc;
    if (COUNT) pc_clear(pc);
    while (true) {
        uint32_t chunk = 0;
        if (lane == 0) chunk = atomicAdd(&ctr.cur_finish, 32u);
        chunk = __shfl_sync(0xffffffffu, chunk, 0);
        if (chunk >= n) break;
        const uint32_t j = chunk + lane;
        uint32_t key = 0xffu, slot = 0;
        if (j < n) {
            slot = q[j];
            const float4 A0 = a.a0[slot], A1 = a.a1[slot];
            const uint4 H = a.hit[slot];
            const V3<R> o(A0.x, A0.y, A0.z), d(A0.w, A1.x, A1.y);
            const HitCore<R> h = closest_hit_finish<R, true, false, FX>(s, sv, o, d, A1.z, (int)H.x, __uint_as_float(H.w), 0ull);
            key = stream_after_hit<COUNT, true>(s, sv, a, slot, d, A1.w, h, pc);
        }
        push_by_key(a, ctr, key, slot);
    }
    if (COUNT) pc_flush(pc, 0, a.counters);
}

// ------------------------------------------------------------------------------------------------
// sink of the deferred next-event contribution: writes the shadow-queue entry
struct ShadowSink {
    const StreamArgs* a;
    BounceCtr* ctr;
    uint32_t slot;
    V3<float> pos, dir;
    float max_dist;
    bool keep_all;      // counting builds queue every culled-in sample so that the event counters match the reference's calls
    PTB_DEV void operator()(V3<float> contrib, uint32_t flags) const {
        if (!keep_all && !(flags & 16u)) return;                 // pdf <= 0: the ray could not contribute (tracer.rs:162)
        const uint32_t k = queue_reserve(&ctr->n_shadow);
        a->s0[k] = make_float4(pos.x, pos.y, pos.z, max_dist);
        a->s1[k] = make_float4(dir.x, dir.y, dir.z, __uint_as_float(slot));
        a->s2[k] = make_float4(contrib.x, contrib.y, contrib.z, __uint_as_float(flags));
    }
};

// MEDIA: the scene has materials with a medium (PTB_MEDIUM_*); a separate instantiation because the in-medium bounce inlined next
// to the surface shading would cost every scene 90 registers
template <bool COUNT, bool BVH, bool MEDIA>
__global__ void PTB_ST_SHADE_BOUNDS k_stream_shade(const __grid_constant__ DScene<float> s, const StreamArgs a, const uint32_t bounce) {
    using R = float;
    {
        uint32_t chunks = 0;                                 // chunks of this stage: every key's queue rounded up to whole chunks
#pragma unroll
        for (int k = 0; k < (int)ST_KEYS; ++k) chunks += (a.ctr[bounce].n_shade[k] + 31u) >> 5;
        PTB_ST_LEAVE_IF_IDLE((uint64_t)chunks * 32u);
    }
    __shared__ SceneSmem<R> sm;
    const SceneView<R> sv = stage_scene(s, sm.words, PTB_SMEM_SCENE_BYTES);
    BounceCtr& ctr = a.ctr[bounce];
    BounceCtr& next = a.ctr[bounce + 1];
    const uint32_t lane = threadIdx.x & 31u;
    const unsigned FULL = 0xffffffffu;
    // chunk table: keys in descending shading cost (more lobes first), so the stage ends on cheap chunks
    const int order_by_cost[ST_KEYS] = {7, 3, 5, 6, 1, 2, 4, 0, (int)ST_MISS};
    uint32_t first_chunk[ST_KEYS + 1];
    {
        uint32_t run = 0;
#pragma unroll
        for (int k = 0; k < (int)ST_KEYS; ++k) { first_chunk[k] = run; run += (ctr.n_shade[order_by_cost[k]] + 31u) >> 5; }
        first_chunk[ST_KEYS] = run;
    }
    PathCounters pc;
    if (COUNT) pc_clear(pc);

    while (true) {
        uint32_t ch = 0;
        if (lane == 0) ch = atomicAdd(&ctr.cur_shade, 1u);
        ch = __shfl_sync(FULL, ch, 0);
        if (ch >= first_chunk[ST_KEYS]) break;
        int k = 0;
#pragma unroll
        for (int t = 1; t < (int)ST_KEYS; ++t) k += ch >= first_chunk[t] ? 1 : 0;
        uint32_t key = 0, base_chunk = 0;
#pragma unroll
        for (int t = 0; t < (int)ST_KEYS; ++t) if (t == k) { key = (uint32_t)order_by_cost[t]; base_chunk = first_chunk[t]; }
        const uint32_t idx = (ch - base_chunk) * 32u + lane;
        if (idx >= ctr.n_shade[key]) continue;
        const uint32_t slot = a.shadeq[(size_t)key * a.cap + idx];
        const float4 A0 = a.a0[slot], A1 = a.a1[slot], A2 = a.a2[slot];
        float4 A3 = a.a3[slot];
        PathState<R> p;
        p.o = V3<R>(A0.x, A0.y, A0.z);
        p.d = V3<R>(A0.w, A1.x, A1.y);
        p.thr = V3<R>(A2.x, A2.y, A2.z);
        p.rad = V3<R>(A3.x, A3.y, A3.z);
        p.prev_pdf = 0;
        p.bounce = bounce;
        const uint32_t w2 = __float_as_uint(A2.w);               // sample index within the wave | PathState::medium << 24 (scenes with media)
        const uint32_t sidx = w2 & 0xffffffu;
        p.medium = MEDIA ? w2 >> 24 : 0u;
        if (key == ST_MISS) {                                    // tracer.rs:66-69
            path_add_sky(s, p);
            if (COUNT) pc.end_sky++;
            A3.x = p.rad.x; A3.y = p.rad.y; A3.z = p.rad.z;
            a.a3[slot] = A3;
            continue;
        }
        const uint4 H = a.hit[slot];
        const int prim = (int)H.x;
        const uint64_t accepted = (uint64_t)H.y | ((uint64_t)H.z << 32);
        p.hit_dist = __uint_as_float(H.w);
        Rng<R> rng(__float_as_uint(A3.w), a.sample0 + sidx, a.seed);
        R u[8];
        shade_draws(rng, bounce, s.n_lights > 1u || (MEDIA && p.medium != 0u) || (key & 4u) != 0u, u);      // the queue key is the lobe class: warp-uniform
        bool cont0;
        int med = MED_SURFACE;
        if (MEDIA && p.medium) {                                 // media (PTB_MEDIUM_*): the in-medium light sample is deferred like a surface's
            MediumNee<R> mn;
            med = path_medium<R, COUNT, BVH, false>(s, sv, p, u, &pc, &mn);
            if (mn.wants) {
                ShadowSink sink{&a, &ctr, slot, mn.pos, mn.dir, mn.max_dist, COUNT};
                sink(mn.contrib, 16u | 32u);                     // bit 5: not a Disney evaluation (event counters)
            }
            A3.x = p.rad.x; A3.y = p.rad.y; A3.z = p.rad.z; a.a3[slot] = A3;      // (an emissive medium has added to the radiance)
        }
        if (med != MED_SURFACE) {
            cont0 = med == MED_SCATTERED;
        } else {
            Mat<R> mat;
            const uint32_t mi = hit_material<R, BVH>(s, sv, prim, accepted, p.d, mat);
            const V3<R> normal = hit_normal<R, BVH, false>(s, sv, prim, p.o, p.d, p.hit_dist);
            ShadeSetup<R> su;
            shade_setup<R, COUNT>(s, p, normal, mat, su, &pc);
            if (s.has_emissive) { A3.x = p.rad.x; A3.y = p.rad.y; A3.z = p.rad.z; a.a3[slot] = A3; }     // tracer.rs:74
            NeeSample<R> ns;
            shade_nee_sample<R, MEDIA>(s, sv, su, u, ns);            // (MEDIA = the scene's FX flag: media or live extended lights)
            if (COUNT && ns.wants_shadow_ray) pc.any_hit++;
            ShadowSink sink{&a, &ctr, slot, ns.scatter_pos, ns.ls.direction, ns.ls.dist - s.eps, COUNT};
            cont0 = shade_finish<R, COUNT, true, ShadowSink>(s, p, mat, su, ns.wants_shadow_ray, ns.ls, ns.light_area, u, &pc, sink);
            if (MEDIA) path_medium_update<R, BVH>(s, sv, p, normal, mi);
        }
        bool cont = cont0;
        if (cont && a.rr_start != 0 && p.bounce >= a.rr_start) {     // RR extension at the start of bounce p.bounce (slot 0 of that bounce)
            R u4[4];
            rng.block(p.bounce, 0, u4);
            if (!russian_roulette_survives(p, u4[0])) { cont = false; if (COUNT) pc.end_rr++; }
        }
        if (cont) {
            a.a0[slot] = make_float4(p.o.x, p.o.y, p.o.z, p.d.x);
            a.a1[slot] = make_float4(p.d.y, p.d.z, p.hit_dist, p.prev_pdf);
            a.a2[slot] = make_float4(p.thr.x, p.thr.y, p.thr.z, __uint_as_float(sidx | (p.medium << 24)));
            a.rayq[(bounce + 1u) & 1u][queue_reserve(&next.n_ray)] = slot;
        }
    }
    if (COUNT) pc_flush(pc, 0, a.counters);
}

// ------------------------------------------------------------------------------------------------
// Scene::any_hit over the shadow queue, one ray per lane per 32-ray chunk (see k_stream_closest).
template <bool COUNT, bool BVH>
__global__ void __launch_bounds__(ST_THREADS) k_stream_shadow(const __grid_constant__ DScene<float> s, const StreamArgs a, const uint32_t bounce) {
    using R = float;
    PTB_ST_LEAVE_IF_IDLE(a.ctr[bounce].n_shadow);
    __shared__ SceneSmem<R> sm;
    const SceneView<R> sv = stage_scene(s, sm.words, PTB_SMEM_SCENE_BYTES);
    BounceCtr& ctr = a.ctr[bounce];
    const uint32_t n = ctr.n_shadow;
    const uint32_t lane = threadIdx.x & 31u;
    PathCounters pc;
    if (COUNT) pc_clear(pc);
    while (true) {
        uint32_t chunk = 0;
        if (lane == 0) chunk = atomicAdd(&ctr.cur_shadow, 32u);
        chunk = __shfl_sync(0xffffffffu, chunk, 0);
        if (chunk >= n) break;
        const uint32_t j = chunk + lane;
        if (j >= n) continue;
        const float4 S0 = a.s0[j], S1 = a.s1[j];
        const bool occluded = any_hit<R, BVH, false>(s, sv, V3<R>(S0.x, S0.y, S0.z), V3<R>(S1.x, S1.y, S1.z), S0.w, COUNT ? pc.bvh : nullptr);     // tracer.rs:150-154
        if (!occluded) {
            const float4 S2 = a.s2[j];
            const uint32_t flags = __float_as_uint(S2.w);
            if (flags & 16u) {                                   // tracer.rs:162-164
                const uint32_t slot = __float_as_uint(S1.w);
                float4 A3 = a.a3[slot];
                A3.x += S2.x; A3.y += S2.y; A3.z += S2.z;
                a.a3[slot] = A3;
                if (COUNT) pc.nee_contrib++;
            }
            if (COUNT && !(flags & 32u)) {
                pc.eval_calls++;
                for (int k = 0; k < 4; ++k) pc.ev[k] += (flags >> k) & 1u;
            }
        }
    }
    if (COUNT) pc_flush(pc, 0, a.counters);
}

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(ST_THREADS) k_stream_accumulate(const StreamArgs a) {
    const uint32_t il = blockIdx.x * blockDim.x + threadIdx.x;
    if (il >= a.npix) return;
    uint32_t px, prow;
    if (!tiled_pixel(a, a.pix0 + il, px, prow)) return;
    float sx = 0, sy = 0, sz = 0;
    const float4* r = a.a3 + (size_t)il * a.S;
    for (uint32_t k = 0; k < a.S; ++k) { const float4 v = r[k]; sx += v.x; sy += v.y; sz += v.z; }
    const uint32_t pix = prow * a.W + px;
    float4* dst = a.flush_dst ? a.flush_dst : a.accum;
    float4 v = (a.flush_dst && a.flush_store) ? make_float4(0.f, 0.f, 0.f, 0.f) : dst[pix];
    v.x += sx; v.y += sy; v.z += sz; v.w += (float)a.S;
    dst[pix] = v;
}

// ------------------------------------------------------------------------------------------------
struct StreamState {
    void* mem = nullptr;         // one allocation holding every queue
    size_t mem_bytes = 0;
    uint32_t cap = 0;            // paths per wave the allocation was sized for
    BounceCtr* ctr = nullptr;
    uint32_t ctr_bounces = 0;
    int grid[8][7] = {};         // persistent grid per [MEDIA*4 + COUNT*2 + BVH][plain stage 0..2, split stage 0..3]
    void release() {
        if (mem) cudaFree(mem);
        if (ctr) cudaFree(ctr);
        mem = nullptr; ctr = nullptr; mem_bytes = 0; cap = 0; ctr_bounces = 0;
    }
};

inline int stream_render(StreamState& st, const DScene<float>& d, void* accum, void* flush_dst, uint32_t W, uint32_t H, uint32_t spp, uint64_t sample_base,
                         const ptb_config& cfg, cudaStream_t stream, int sm_count, DeviceCounters* counters, cudaEvent_t ev0, cudaEvent_t ev1,
                         uint64_t* launches, std::string& err) {
    cudaError_t e;
    auto cuda_fail = [&](const char* what) { err = std::string(what) + ": " + cudaGetErrorString(e); return PTB_E_CUDA; };
    StreamArgs a{};
    a.W = W; a.H = H;
    a.tiles_x = (W + 15u) / 16u;
    a.n_items = a.tiles_x * ((H + 15u) / 16u) * 256u;
    a.seed = cfg.seed; a.rr_start = cfg.rr_start; a.counters = counters; a.accum = (float4*)accum; a.flush_dst = (float4*)flush_dst;

    // wave shape: all pixels x S samples when the frame is small, else pixel ranges x 1 sample
    uint32_t want = cfg.wave_paths ? cfg.wave_paths : ST_DEFAULT_WAVE;
    want = std::max<uint32_t>(256u, want & ~255u);
    uint32_t pix_per_wave, S;
    if (a.n_items >= want) { pix_per_wave = want; S = 1; }
    else { pix_per_wave = a.n_items; S = std::max<uint32_t>(1u, std::min<uint32_t>(std::min<uint32_t>(spp, want / a.n_items), 4096u)); }
    const uint32_t cap = pix_per_wave * S;

    if (st.cap < cap) {
        if (st.mem) { cudaStreamSynchronize(stream); cudaFree(st.mem); st.mem = nullptr; st.cap = 0; }
        const size_t per_path = 4 * 16 + 16 + 2 * 4 + ST_KEYS * 4 + 3 * 16;
        st.mem_bytes = (size_t)cap * per_path;
        if ((e = cudaMalloc(&st.mem, st.mem_bytes)) != cudaSuccess) return cuda_fail("cudaMalloc(stream queues)");
        st.cap = cap;
    }
    const uint32_t n_ctr = d.depth + 2u;
    if (st.ctr_bounces < n_ctr) {
        if (st.ctr) { cudaStreamSynchronize(stream); cudaFree(st.ctr); st.ctr = nullptr; }
        if ((e = cudaMalloc((void**)&st.ctr, (size_t)n_ctr * sizeof(BounceCtr))) != cudaSuccess) return cuda_fail("cudaMalloc(stream counters)");
        st.ctr_bounces = n_ctr;
    }
    {
        char* p = (char*)st.mem;
        const size_t c = st.cap;
        a.a0 = (float4*)p; p += c * 16; a.a1 = (float4*)p; p += c * 16; a.a2 = (float4*)p; p += c * 16; a.a3 = (float4*)p; p += c * 16;
        a.hit = (uint4*)p; p += c * 16;
        a.s0 = (float4*)p; p += c * 16; a.s1 = (float4*)p; p += c * 16; a.s2 = (float4*)p; p += c * 16;
        a.rayq[0] = (uint32_t*)p; p += c * 4; a.rayq[1] = (uint32_t*)p; p += c * 4;
        a.shadeq = (uint32_t*)p;
        a.cap = st.cap;
        a.ctr = st.ctr;
    }

    const bool count = cfg.collect_counters != 0;
    const bool bvh = d.use_bvh != 0;
    const bool media = d.has_fx != 0;           // media or live extended lights: the FX instantiations of closest / finish / shade
    const int vi = (media ? 4 : 0) + (count ? 2 : 0) + (bvh ? 1 : 0);
    using StageKernel = void (*)(const DScene<float>, const StreamArgs, const uint32_t);
    // Stage kernels of a bounce, in launch order.  Plain form: closest, shade, shadow (one ray per lane per chunk).
    // Large BVH scenes from bounce 1 on (incoherent rays, deep tree): trace, finish, shade, trace<ANY> with persistent lanes —
    // measured on 100k spheres: bounce 0 0.83 ms plain vs 1.85 ms split, bounce 1 1.72 ms plain vs 1.33 + 0.44 ms split
    // (wash) ... while 4096 spheres run 1.3x faster in the plain form on every bounce (profiles/r01_ncu_stream.md).
    StageKernel plain[3], split[4];
    if (bvh) {
        plain[0] = media ? (count ? k_stream_closest<true, true, true> : k_stream_closest<false, true, true>)
                         : (count ? k_stream_closest<true, true, false> : k_stream_closest<false, true, false>);
        plain[1] = media ? (count ? k_stream_shade<true, true, true> : k_stream_shade<false, true, true>)
                         : (count ? k_stream_shade<true, true, false> : k_stream_shade<false, true, false>);
        plain[2] = count ? k_stream_shadow<true, true> : k_stream_shadow<false, true>;
    } else {
        plain[0] = media ? (count ? k_stream_closest<true, false, true> : k_stream_closest<false, false, true>)
                         : (count ? k_stream_closest<true, false, false> : k_stream_closest<false, false, false>);
        plain[1] = media ? (count ? k_stream_shade<true, false, true> : k_stream_shade<false, false, true>)
                         : (count ? k_stream_shade<true, false, false> : k_stream_shade<false, false, false>);
        plain[2] = count ? k_stream_shadow<true, false> : k_stream_shadow<false, false>;
    }
    split[0] = count ? k_stream_trace<false, true> : k_stream_trace<false, false>;
    split[1] = media ? (count ? k_stream_finish<true, true> : k_stream_finish<false, true>) : (count ? k_stream_finish<true, false> : k_stream_finish<false, false>);
    split[2] = plain[1];
    split[3] = count ? k_stream_trace<true, true> : k_stream_trace<true, false>;
    const bool use_split = stream_uses_split(d);
    int g_plain[3], g_split[4];
    auto grid_of = [&](int& slot, StageKernel k) {
        if (slot == 0) {
            int per_sm = 0;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, ST_THREADS, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
            slot = per_sm * sm_count;
        }
        return slot;
    };
    for (int k = 0; k < 3; ++k) g_plain[k] = grid_of(st.grid[vi][k], plain[k]);
    for (int k = 0; k < 4; ++k) g_split[k] = use_split ? grid_of(st.grid[vi][3 + k], split[k]) : 0;

    if ((e = cudaEventRecord(ev0, stream)) != cudaSuccess) return cuda_fail("cudaEventRecord");
    for (uint32_t s0 = 0; s0 < spp; s0 += S) {
        const uint32_t Sw = std::min(S, spp - s0);
        for (uint32_t pix0 = 0; pix0 < a.n_items; pix0 += pix_per_wave) {
            a.pix0 = pix0; a.npix = std::min(pix_per_wave, a.n_items - pix0); a.S = Sw; a.P = a.npix * Sw;
            a.sample0 = sample_base + s0;
            a.flush_store = s0 == 0 ? 1u : 0u;
            if ((e = cudaMemsetAsync(st.ctr, 0, (size_t)n_ctr * sizeof(BounceCtr), stream)) != cudaSuccess) return cuda_fail("cudaMemsetAsync");
            const unsigned g_gen = (a.P + ST_THREADS - 1) / ST_THREADS;
            if (count) k_stream_generate<true><<<g_gen, ST_THREADS, 0, stream>>>(d, a);
            else k_stream_generate<false><<<g_gen, ST_THREADS, 0, stream>>>(d, a);
            (*launches)++;
            const int cap_grid = (int)std::max<uint32_t>(1u, (a.P + ST_THREADS - 1) / ST_THREADS);
            for (uint32_t b = 0; b < d.depth; ++b) {
                if (use_split && b >= 1u) {
                    for (int k = 0; k < 4; ++k) split[k]<<<std::min(g_split[k], cap_grid), ST_THREADS, 0, stream>>>(d, a, b);
                    (*launches) += 4;
                } else {
                    for (int k = 0; k < 2; ++k) plain[k]<<<std::min(g_plain[k], cap_grid), ST_THREADS, 0, stream>>>(d, a, b);
                    // shadow rays are incoherent from the first bounce on (64 lights): persistent lanes already pay at bounce 0 (1.21 -> 1.06 ms)
                    if (use_split) split[3]<<<std::min(g_split[3], cap_grid), ST_THREADS, 0, stream>>>(d, a, b);
                    else plain[2]<<<std::min(g_plain[2], cap_grid), ST_THREADS, 0, stream>>>(d, a, b);
                    (*launches) += 3;
                }
            }
            k_stream_accumulate<<<(a.npix + ST_THREADS - 1) / ST_THREADS, ST_THREADS, 0, stream>>>(a);
            (*launches)++;
            if ((e = cudaGetLastError()) != cudaSuccess) return cuda_fail("stream wavefront launch");
        }
    }
    if ((e = cudaEventRecord(ev1, stream)) != cudaSuccess) return cuda_fail("cudaEventRecord");
    return PTB_OK;
}

}  // namespace ptb
