// rt_path.cuh — the wavefront path tracer (replaces cuda_path_trace_kernel + cuda_trace_path,
// renderers/cuda_path_tracer.py:17-471).
//
// One WAVE = spp_per_wave samples of every pixel.  Path state lives in compacted HBM queues of real4 record
// streams; all kernels are persistent grid-stride launches.  Per bounce:
//     shade_kernel<MODE>   closest hit + shading FUSED (camera-ray generation too at bounce 0): texture, NEE shadow-ray
//                          emission, Russian roulette, BSDF sample -> next ray queue + shadow queue
//                          (warp-ballot / prefix-popcount compaction, one packed 64-bit atomic per warp)
//                          small float32 scenes: scan of box + planar records, surface records, all in shared memory
//                          large scenes, bounce 0: LBVH walk with the top levels in shared memory
//     extend_walk_kernel + shade_kernel<0>   large scenes, bounce >= 1: persistent walk (dynamic ray fetch, majority
//                          scheduling) -> hit stream -> wavefront shade stage
//     shadow_kernel        shadow queue -> occlusion query -> per-path radiance
// and once per wave accumulate_kernel (per-pixel sums).  extend_kernel + shade_kernel<0> is the textbook unfused
// form, kept as the measured baseline (B2RT_PATH_UNFUSED).
// Queue record streams (16 B * 3 per path for float):
//     ro = (origin.xyz, slot)   rd = (direction.xyz, rng state)   th = (throughput.rgb, depth)
// Shadow records: so = (origin.xyz, slot)  sd = (direction.xyz, -)  sc = (throughput*contribution.rgb, -)
// slot = s_local * n_pixels + pixel indexes the per-path radiance L[slot] (never compacted), which makes
// every radiance update a race-free plain read-modify-write and the per-pixel sum order deterministic.
#pragma once
#include <type_traits>

#include "rt_scene.cuh"

namespace b2rt {

// ------------------------------------------------------------------------------------ RNG policies
// random(state) is a pure function of the state and advance(state) steps it — the reference's
// calling convention (cuda_random / cuda_xorshift, cuda_path_tracer.py:61-71), which the integrator
// below follows draw for draw.
struct RefRng {                       // the reference's generator, int64 arithmetic
    static __device__ __forceinline__ uint64_t advance(uint64_t s) {
        long long x = (long long)s;
        x ^= (long long)((unsigned long long)x << 13);
        x ^= x >> 17;                                          // arithmetic shift of the 64-bit value
        x ^= (long long)((unsigned long long)x << 5);
        return (uint64_t)(x & 0xffffffffLL);
    }
    template <typename R> static __device__ __forceinline__ R random(uint64_t s) {
        return R((double)(s & 0xffffffULL) / 16777216.0);
    }
};
struct PcgRng {                       // 32-bit PCG-RXS-M-XS stream, seeded by hashing (pixel, sample, seed)
    static __device__ __forceinline__ uint64_t advance(uint64_t s) {
        return (uint64_t)((uint32_t)s * 747796405u + 2891336453u);
    }
    static __device__ __forceinline__ uint32_t word(uint64_t s64) {          // the full 32-bit output of the state
        uint32_t s = (uint32_t)s64;
        uint32_t w = ((s >> ((s >> 28u) + 4u)) ^ s) * 277803737u;
        return (w >> 22u) ^ w;
    }
    template <typename R> static __device__ __forceinline__ R random(uint64_t s64) {
        return R(word(s64) >> 8) * R(1.0 / 16777216.0);
    }
    static __device__ __forceinline__ uint32_t mix(uint32_t h) {
        h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
        return h;
    }
    static __device__ __forceinline__ uint64_t seed(uint32_t pixel, uint64_t sample, uint64_t seed) {
        uint32_t h = mix(pixel * 0x9e3779b1u + (uint32_t)seed);
        h = mix(h ^ ((uint32_t)sample * 0x85ebca77u + (uint32_t)(seed >> 32)));
        // sample indices beyond 2^32 get a third round (warp-uniform: never taken below 4 G samples per pixel)
        if (sample >> 32) h = mix(h + (uint32_t)(sample >> 32) * 0xc2b2ae3du + 0x27d4eb2fu);
        return (uint64_t)h;
    }
};

template <typename R> struct PathQueues {
    real4<R> *ro[2], *rd[2], *th[2];      // double-buffered ray queue streams
    real4<R> *hit;                         // (t, prim, a, b)
    // hit queue of the split small-scene bounce (scan_hits_kernel -> shade_kernel<7>): compacted, hits only
    //     ha = (hit point.xyz, slot)  hb = (direction.xyz, rng)  hc = (throughput.rgb, prim)  hd = (a, b, -, -)
    real4<R> *ha, *hb, *hc, *hd;
    unsigned long long *hit_tail;          // [max_depth] tail of the hit queue of bounce k (low word)
    real4<R> *so, *sd, *sc;                // shadow queue streams
    real4<R> *L;                           // per-path radiance
    // counts[k] = (shadow rays emitted at bounce k-1) << 32 | (rays queued for bounce k): both queue tails move
    // with ONE 64-bit atomic per warp (the L2 serialises same-line atomics: they were a first-order cost)
    unsigned long long *counts;            // [max_depth + 1]
    unsigned long long *unshadowed;        // [1]
    unsigned long long *culled;            // [1] shadow rays answered by the occluder hint (never queued)
    unsigned long long *dead;              // [max_depth + 1] unused (dead) entries inside the queues of bounce k: rays | shadow records << 32
    unsigned long long *clk;               // [2] sum of SM cycles / nanoseconds that CTA 0 of every bounce kernel ran (effective SM clock)
    unsigned long long *tally;             // [8] bounds-culled camera rays, shaded hits, walk box / leaf steps, sky records
    unsigned *keys;                        // sort key of every ray appended to the next queue (or nullptr)
    const int *perm;                       // permutation the current queue is read through (or nullptr)
    int capacity;                          // entries per queue stream (B2RT_CHECK builds verify every append against it)
};

// Ray-reordering key.  Layout 2 (default): 2 bits per direction component (6 bits) above a 24-bit Morton code of
// the origin; layout 0: octant | 27-bit Morton; layout 1: Morton | octant.
__device__ __forceinline__ unsigned spread9(unsigned v) {
    v &= 0x1ffu;
    v = (v | (v << 16)) & 0x030000ffu;
    v = (v | (v << 8)) & 0x0300f00fu;
    v = (v | (v << 4)) & 0x030c30c3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}
template <typename R> __device__ __forceinline__ unsigned ray_sort_key(V3<R> o, V3<R> d, float inv) {
    float fx = fminf(fmaxf(((float)o.x * inv + 0.5f) * 512.f, 0.f), 511.f);
    float fy = fminf(fmaxf(((float)o.y * inv + 0.5f) * 512.f, 0.f), 511.f);
    float fz = fminf(fmaxf(((float)o.z * inv + 0.5f) * 512.f, 0.f), 511.f);
    unsigned m = (spread9((unsigned)fx) << 2) | (spread9((unsigned)fy) << 1) | spread9((unsigned)fz);
    unsigned oct = (d.x < R(0) ? 4u : 0u) | (d.y < R(0) ? 2u : 0u) | (d.z < R(0) ? 1u : 0u);
#if B2RT_SORT_KEY == 1
    return (m << 3) | oct;
#elif B2RT_SORT_KEY == 2
    unsigned qx = (unsigned)fminf(fmaxf(((float)d.x + 1.f) * 2.f, 0.f), 3.f);
    unsigned qy = (unsigned)fminf(fmaxf(((float)d.y + 1.f) * 2.f, 0.f), 3.f);
    unsigned qz = (unsigned)fminf(fmaxf(((float)d.z + 1.f) * 2.f, 0.f), 3.f);
    return (((qx << 4) | (qy << 2) | qz) << 24) | (m >> 3);
#else
    return (oct << 27) | m;
#endif
}
template <typename R> __device__ __forceinline__ int ray_count(const PathQueues<R> &Q, int bounce) {
    return (int)(Q.counts[bounce] & 0xffffffffULL);
}
template <typename R> __device__ __forceinline__ int shadow_count(const PathQueues<R> &Q, int bounce) {
    return (int)(Q.counts[bounce + 1] >> 32);
}

// ---- chunked queue append (small-scene bounce kernels) ---------------------------------------------------------
// ONE atomic per warp iteration on the queue tail was THE limit of the Cornell bounce kernels: every warp of the grid
// hits the same 8 bytes ~0.8 G times a second, which is all one L2 atomic unit does (measured, profiles/r2m: a second
// no-op atomicAdd on that word doubled the kernel time, 46.6 -> 95.1 ms; the rate differs by ~30 % between GPUs of the
// pool, and so did the whole benchmark).  Here every warp owns a CHUNK of kQueueChunk slots in each queue and goes back
// to the tail only when the chunk is used up (one atomic per ~5 iterations); an append that does not fit fills the old
// chunk and continues in the new one, so the only unused slots are each warp's last chunk remainder, which the warp
// marks dead (slot word -1) before it exits — consumers skip those, the statistics subtract them (Q.dead).
#ifndef B2RT_QCHUNK
#define B2RT_QCHUNK 128
#endif
constexpr int kQueueChunk = B2RT_QCHUNK;
struct WarpCursor { int ray_cur, ray_end, sh_cur, sh_end; };     // shared memory, one per warp, written by lane 0

// refill path, out of line (the bounce kernels live next to the instruction-cache limit): takes a new chunk from the tail
// word (32-bit half of the packed counter) and returns its base
static __device__ __noinline__ int chunk_refill(unsigned *tail_word) {
    unsigned nb = 0;
    if ((threadIdx.x & 31u) == 0) nb = atomicAdd(tail_word, (unsigned)kQueueChunk);
    return (int)__shfl_sync(0xffffffffu, nb, 0);
}
__device__ __forceinline__ int chunk_take(unsigned *tail_word, int &cur, int &end, unsigned m, unsigned lane) {
    const int k = __popc(m), free_ = end - cur, rank = __popc(m & ((1u << lane) - 1u));
    int slot = cur + rank;
    cur += k;
    if (k > free_) {                                             // warp-uniform
        const int nb = chunk_refill(tail_word);
        if (rank >= free_) slot = nb + (rank - free_);
        cur = nb + (k - free_);
        end = nb + kQueueChunk;
    }
    return slot;
}
__device__ __forceinline__ void warp_append_chunked(unsigned long long *tail, WarpCursor *wc, bool want_ray, bool want_shadow,
                                                    int &ray_slot, int &shadow_slot) {
    const unsigned mr = __ballot_sync(0xffffffffu, want_ray), ms = __ballot_sync(0xffffffffu, want_shadow);
    ray_slot = shadow_slot = -1;
    if ((mr | ms) == 0) return;
    const unsigned lane = threadIdx.x & 31u;
    unsigned *words = reinterpret_cast<unsigned *>(tail);        // little endian: [0] ray tail, [1] shadow tail
    if (mr) {
        int cur = wc->ray_cur, end = wc->ray_end;
        const int slot = chunk_take(words, cur, end, mr, lane);
        if (want_ray) ray_slot = slot;
        if (lane == 0) { wc->ray_cur = cur; wc->ray_end = end; }
    }
    if (ms) {
        int cur = wc->sh_cur, end = wc->sh_end;
        const int slot = chunk_take(words + 1, cur, end, ms, lane);
        if (want_shadow) shadow_slot = slot;
        if (lane == 0) { wc->sh_cur = cur; wc->sh_end = end; }
    }
    __syncwarp();
}

// per-thread statistic flushed once per warp at kernel end
__device__ __forceinline__ void warp_flush(unsigned long long *counter, unsigned v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(counter, (unsigned long long)v);
}

// ------------------------------------------------------------------------------------ raygen
// cuda_path_trace_kernel's sample loop (:28-46).  One thread per pixel walks its spp_wave samples so
// the reference generator's per-pixel sequential state can be carried in pixel_rng.
template <typename R, typename Rng>
__global__ void __launch_bounds__(256)
raygen_kernel(Cam<R> cam, int W, int H, int spp_wave, long long first_sample, unsigned long long seed,
              long long *pixel_rng, PathQueues<R> Q) {
    int npix = W * H;
    for (int pix = blockIdx.x * blockDim.x + threadIdx.x; pix < npix; pix += gridDim.x * blockDim.x) {
        int x = pix % W, y = pix / W;
        uint64_t state = 0;
        if constexpr (std::is_same<Rng, RefRng>::value) state = (uint64_t)pixel_rng[pix];
        for (int s = 0; s < spp_wave; ++s) {
            if constexpr (std::is_same<Rng, PcgRng>::value)
                state = PcgRng::seed((uint32_t)pix, (uint64_t)(first_sample + s), seed);
            R rnd = Rng::template random<R>(state);          // :35-36 the same draw jitters u and v
            R u = (R(x) + rnd) / R(W), v = (R(y) + rnd) / R(H);
            state = Rng::advance(state);                      // :37
            Ray<R> r = camera_ray<R>(cam, u, v);
            size_t i = (size_t)s * npix + pix;
            Q.ro[0][i] = Real4<R>::make(r.o.x, r.o.y, r.o.z, pack_int<R>((int64_t)i));
            Q.rd[0][i] = Real4<R>::make(r.d.x, r.d.y, r.d.z, pack_int<R>((int64_t)state));   // by value (:40-41)
            Q.th[0][i] = Real4<R>::make(R(1), R(1), R(1), pack_int<R>(0));
            Q.L[i] = Real4<R>::make(R(0), R(0), R(0), R(0));
            if constexpr (std::is_same<Rng, RefRng>::value) state = Rng::advance(state);     // :46
        }
        if constexpr (std::is_same<Rng, RefRng>::value) pixel_rng[pix] = (long long)state;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) Q.counts[0] = (unsigned long long)(npix * spp_wave);
}

// reference generator: seed every pixel (:28) and skip 2*first_sample steps (two advances per sample)
static __global__ void init_pixel_rng_kernel(int W, int H, long long frame_count, long long first_sample, long long *pixel_rng) {
    int pix = blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= W * H) return;
    long long s = ((long long)pix + frame_count * W * H) * 1103515245LL + 12345LL;
    uint64_t st = (uint64_t)s;
    for (long long k = 0; k < 2 * first_sample; ++k) st = RefRng::advance(st);
    pixel_rng[pix] = (long long)st;
}

// ------------------------------------------------------------------------------------ extend
template <typename R>
__global__ void __launch_bounds__(256)
extend_kernel(SceneDev S, const real4<R> *__restrict__ ro, const real4<R> *__restrict__ rd,
              real4<R> *__restrict__ hit, const unsigned long long *__restrict__ count, int scan) {
    extern __shared__ float4 s_top[];
    if (!scan) stage_top(S, s_top);
    int n = (int)(*count & 0xffffffffULL);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        real4<R> a = ld_stream(ro + i), b = ld_stream(rd + i);
        Ray<R> r; r.o = xyz<R>(a); r.d = xyz<R>(b);
        Hit<R> h;
        if ((int)unpack_u<R>(a.w) < 0) { h.t = R(1000000.0); h.prim = -1; h.a = h.b = R(0); }      // dead queue entry
        else if (scan) scan_all<R, false, false>(S, r, R(0.001), R(1000000.0), h);
        else traverse<R, false, false>(S, s_top, r, R(0.001), R(1000000.0), h);
        st_stream(hit + i, Real4<R>::make(h.t, pack_int<R>((int64_t)h.prim), h.a, h.b));
    }
}

// ------------------------------------------------------------------------------------ extend (persistent walk)
// Closest hit for the incoherent rays of LARGE scenes (bounce >= 1 of LBVH-mode scenes).  Measured on the 1 M-triangle
// scene (profiles/r1d_c4): the per-ray walk fused into the bounce kernel ran with 10 of 32 lanes active — leaf tests
// at 2.3 lanes (a lane in the leaf branch while its neighbours test boxes) and finished rays idling until the
// slowest lane of the warp is done.  This kernel fixes both, after Aila & Laine (HPG 2009 / 2012):
//   * persistent warps with DYNAMIC RAY FETCH: when >= kRefill lanes have finished, they store their hit records and
//     take the next rays of the (sorted) queue through one atomic per warp;
//   * MAJORITY SCHEDULING: per iteration the warp runs either the box step or the leaf step, whichever more lanes
//     are waiting for, so every executed instruction has at least half of the busy lanes on it.
// Results are identical to traverse(): closest t, ties to the lowest packed id.  Rays are FETCHED in sorted order
// (perm) but hit record j is stored at the ray's own queue slot j, so the wavefront shade stage that follows reads
// rays and hits in plain queue order with coalesced loads (gathering 64 B per ray through perm a second time made
// that stage latency-bound: 13.5 % issue-active, ncu profiles/r1f_c4_*).
#ifndef B2RT_WALK_REFILL
#define B2RT_WALK_REFILL 8         // idle lanes that trigger a refill (measured 4 / 8 / 16: within 3 %)
#endif
#ifndef B2RT_WALK_PREFETCH
#define B2RT_WALK_PREFETCH 0      // measured: prefetching the far child (L2) costs 10 % (126 vs 115 ms): the walk is request-bound
#endif
__device__ __forceinline__ void prefetch_line(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void prefetch_tri(const SceneDev &S, int prim) {
    const int t = prim - S.n_rect - S.n_sphere;
    if (t >= 0) prefetch_line(reinterpret_cast<const float4 *>(S.tri) + 3 * (size_t)t);
}
#ifndef B2RT_WALK_NODE_STEPS
#define B2RT_WALK_NODE_STEPS 1     // box steps per vote (measured on 1 M triangles: 1: 115.1, 2: 117.3, 3: 122.4 ms)
#endif
#ifndef B2RT_WALK_MIN_BLOCKS
#define B2RT_WALK_MIN_BLOCKS 4     // resident CTAs/SM (47 registers; 5 measured equal)
#endif
// COUNT (B2RT_PATH_COUNT_TESTS, measurement passes only): per-lane box steps (two slab tests each) and leaf steps
// (one primitive test each) are tallied into tally[0] / tally[1] — the executed-work and bytes-per-ray figures of
// bench.py's 1 M-triangle configuration come from these, never from an estimate.
//
// WIDE: the box step reads one 4-wide node (128 B: the four grandchild boxes of a binary node, lbvh.cu:widen_kernel)
// instead of one 64 B binary node, tests four slabs, continues with the nearest child that is hit and stacks the others
// far to near.  The kernel is bound by the latency of these dependent fetches, so halving their number is what counts;
// the closest hit is the same (the tie rule lives in test_prim, and a stacked child is never culled late in either
// form).  No shared-memory top copy: the wide top levels stay in L1.
#ifndef B2RT_WIDE_MIN_BLOCKS
#define B2RT_WIDE_MIN_BLOCKS 4
#endif
#define B2RT_CSWAP(ta_, ca_, tb_, cb_)                                                  \
    do {                                                                                \
        const bool s_ = (tb_) < (ta_);                                                  \
        const R tl_ = s_ ? (tb_) : (ta_), th_ = s_ ? (ta_) : (tb_);                     \
        const int cl_ = s_ ? (cb_) : (ca_), ch_ = s_ ? (ca_) : (cb_);                   \
        (ta_) = tl_; (tb_) = th_; (ca_) = cl_; (cb_) = ch_;                             \
    } while (0)
// Two switches that cut the L1 data-pipe work of the binary walk, both MEASURED AND OFF (profiles/r2_c4_walk_kernel_analysis.md:
// ncu shows the L1 data pipe at 89.6 % of its wavefront rate on the 1 M-triangle scene, yet every combination below lands
// on the same 65.16 ms per step as the plain kernel, 65.1):
//   B2RT_WALK_LDG256      the 64 B node comes in as two 256-bit loads (sm_100 LDG.E.256) instead of four 128-bit ones;
//   B2RT_WALK_SMEM_STACK  the first N traversal-stack entries live in shared memory as [entry][thread] (every lane has its
//                         own bank whatever its depth: a push or pop is one wavefront and never leaves the SM, where the
//                         per-thread local array misses L1 for 97 % of its sectors).  Deeper entries fall back to the
//                         local array; the shared-memory top copy shrinks to B2RT_WALK_TOP_STAGE nodes to make room (the
//                         rest of the top levels is read from S.top).
#ifndef B2RT_WALK_LDG256
#define B2RT_WALK_LDG256 0
#endif
#ifndef B2RT_WALK_SMEM_STACK
#define B2RT_WALK_SMEM_STACK 0
#endif
#ifndef B2RT_WALK_TOP_STAGE
#define B2RT_WALK_TOP_STAGE 128
#endif
constexpr int kWalkSmemStack = B2RT_WALK_SMEM_STACK < kStackDepth ? B2RT_WALK_SMEM_STACK : kStackDepth;
inline __host__ __device__ int walk_top_staged(int n_top) {
    return (kWalkSmemStack > 0 && n_top > B2RT_WALK_TOP_STAGE) ? B2RT_WALK_TOP_STAGE : n_top;
}
// dynamic shared memory of the binary walk kernel: staged top nodes + the shared part of the stacks (256 threads)
inline size_t walk_smem_bytes(const SceneDev &S) {
    return (size_t)walk_top_staged(S.n_top) * 64 + (size_t)kWalkSmemStack * 256 * sizeof(int);
}
// NODES selects the node format of the box step: 0 = binary 64 B nodes (S.nodes / S.top), 1 = 4-wide 128 B nodes (S.wide),
// 2 = QUANTISED binary nodes of 32 B (S.quant, lbvh.cu:quantize_kernel): both child boxes as 16-bit cell indices on a grid
// over the scene bounds, rounded outward by a whole cell, so one node is TWO 16-byte loads (two L1 sector accesses per lane
// instead of four — the walk is bound by the L1 data pipe, one wavefront per sector access) and a slab plane is one FFMA,
// t = q * (scale / d) + (base - o) / d.  The boxes only grow, so the closest hit is unchanged.
template <typename R, bool COUNT, int NODES>
__global__ void __launch_bounds__(256, sizeof(R) == 4 ? (NODES == 1 ? B2RT_WIDE_MIN_BLOCKS : B2RT_WALK_MIN_BLOCKS) : 1)
extend_walk_kernel(SceneDev S, const real4<R> *__restrict__ ro, const real4<R> *__restrict__ rd,
                   real4<R> *__restrict__ hit, const unsigned long long *__restrict__ count,
                   const int *__restrict__ perm, unsigned *__restrict__ next, unsigned long long *tally) {
    unsigned n_node = 0, n_leaf = 0;
    extern __shared__ float4 s_top[];
    constexpr bool WIDE = NODES == 1, QUANT = NODES == 2;
    const int n_stage = (WIDE || QUANT) ? 0 : walk_top_staged(S.n_top);
    if (!WIDE && !QUANT) {
        for (int i = threadIdx.x; i < 4 * n_stage; i += blockDim.x) s_top[i] = __ldg(S.top + i);
        __syncthreads();
    }
    constexpr int kDone = (int)0x80000000;                       // below every leaf reference (~prim)
    constexpr int kDepth = WIDE ? kWideStackDepth : kStackDepth;
    constexpr int kShared = (WIDE || QUANT) ? 0 : kWalkSmemStack;    // stack entries [0, kShared) live in shared memory
    const int n = (int)(*count & 0xffffffffULL);
    const unsigned lane = threadIdx.x & 31u, lt = (1u << lane) - 1u;
    const R t_min = R(0.001);
    int stack[kDepth - kShared > 0 ? kDepth - kShared : 1];
    int *s_stack = reinterpret_cast<int *>(s_top + 4 * n_stage) + threadIdx.x;     // entry k at s_stack[k * 256]
    int sp = 0, ref = kDone, pos = -1;
    Ray<R> r; r.o = {R(0), R(0), R(0)}; r.d = r.o;
    V3<R> id = r.o, qa = r.o, qb = r.o;
    Hit<R> best; best.t = R(0); best.a = R(0); best.b = R(0); best.prim = -1;
    bool exhausted = false;                                      // warp-uniform: the queue has no rays left
    auto push = [&](int v) {
        if (B2RT_CHECK && sp >= kDepth) { if (S.check) atomicAdd(S.check, 1ULL); return; }
        if (kShared > 0 && sp < kShared) s_stack[sp * 256] = v; else stack[sp - kShared] = v;
        ++sp;
    };
    auto pop = [&]() -> int {
        --sp;
        return (kShared > 0 && sp < kShared) ? s_stack[sp * 256] : stack[sp - kShared];
    };
    for (;;) {
        const bool want_node = ref >= 0, want_leaf = ref < 0 && ref != kDone;
        const unsigned mn = __ballot_sync(0xffffffffu, want_node), ml = __ballot_sync(0xffffffffu, want_leaf);
        const unsigned idle = ~(mn | ml);
        if (!exhausted && (__popc(idle) >= B2RT_WALK_REFILL)) {
            if (ref == kDone && pos >= 0) {
                const real4<R> rec = Real4<R>::make(best.t, pack_int<R>((int64_t)best.prim), best.a, best.b);
                if constexpr ((B2RT_L2_HINT & 4) != 0) st_stream(hit + pos, rec); else hit[pos] = rec;
            }
            // (taking the indices in per-warp chunks of 32 / 128 / 512 instead was measured at 63.3 / 64.4 / 69.7 ms per
            // step against 62.5: this kernel is latency-bound, not bound by the cursor's atomic unit)
            unsigned base = 0;
            if (lane == 0) base = atomicAdd(next, (unsigned)__popc(idle));
            base = __shfl_sync(0xffffffffu, base, 0);
            const unsigned i = base + (unsigned)__popc(idle & lt);
            if (ref == kDone) {
                pos = -1;
                if (i < (unsigned)n) {
                    const int j = perm ? __ldg(perm + i) : (int)i;
                    pos = j;                                     // the hit record goes to the ray's own queue slot
                    const real4<R> a = ld_stream(ro + j), b = ld_stream(rd + j);
                    r.o = xyz<R>(a); r.d = xyz<R>(b);
                    id = {rcp_(r.d.x), rcp_(r.d.y), rcp_(r.d.z)};
                    if constexpr (QUANT) {                       // slab plane of cell index q: t = q * qa + qb
                        const float4 q_base = __ldg(S.quant), q_scale = __ldg(S.quant + 1);     // grid header (uniform, cached)
                        qa = {R(q_scale.x) * id.x, R(q_scale.y) * id.y, R(q_scale.z) * id.z};
                        qb = {(R(q_base.x) - r.o.x) * id.x, (R(q_base.y) - r.o.y) * id.y, (R(q_base.z) - r.o.z) * id.z};
                    }
                    best.t = R(1000000.0); best.prim = -1; best.a = R(0); best.b = R(0);
                    sp = 0; push(kDone);
                    // a dead entry (unused remainder of a producer warp's chunk, slot word -1) is no ray at all
                    const bool dead_entry = (int)unpack_u<R>(a.w) < 0;
                    ref = (S.n_prims > 0 && !dead_entry) ? S.root : kDone;
                    if (dead_entry) pos = -1;
                    if (S.n_outside > 0 && !dead_entry) {        // rectangles outside the hierarchy: leaves visited first
                        push(ref);
                        for (int p = S.n_outside - 1; p >= 1; --p) push(~p);
                        ref = ~0;
                    }
                }
            }
            exhausted = base + (unsigned)__popc(idle) >= (unsigned)n;
            continue;                                            // vote again with the new rays
        }
        if ((mn | ml) == 0u) break;                              // nothing in flight and nothing left to fetch
        if (__popc(mn) >= __popc(ml)) {
            if constexpr (WIDE) {
                if (ref >= 0) {
                    if (COUNT) ++n_node;
                    const float4 *p = S.wide + 8 * (size_t)ref;
                    const float4 ax = __ldg(p), ay = __ldg(p + 1), az = __ldg(p + 2);
                    const float4 bx = __ldg(p + 3), by = __ldg(p + 4), bz = __ldg(p + 5), cf = __ldg(p + 6);
                    constexpr R kMiss = R(3.0e38);
                    // entry distance of one slot, or kMiss (same slab arithmetic as the binary step below)
                    auto enter = [&](float lx, float ly, float lz, float hx, float hy, float hz, int c) -> R {
                        const R x0 = (R(lx) - r.o.x) * id.x, x1 = (R(hx) - r.o.x) * id.x;
                        const R y0 = (R(ly) - r.o.y) * id.y, y1 = (R(hy) - r.o.y) * id.y;
                        const R z0 = (R(lz) - r.o.z) * id.z, z1 = (R(hz) - r.o.z) * id.z;
                        const R tn = max_(max_(min_(x0, x1), min_(y0, y1)), max_(min_(z0, z1), t_min));
                        const R tf = min_(min_(max_(x0, x1), max_(y0, y1)), min_(max_(z0, z1), best.t));
                        return (tn <= tf && c != kDone) ? tn : kMiss;
                    };
                    int c0 = __float_as_int(cf.x), c1 = __float_as_int(cf.y), c2 = __float_as_int(cf.z), c3 = __float_as_int(cf.w);
                    R t0 = enter(ax.x, ay.x, az.x, bx.x, by.x, bz.x, c0), t1 = enter(ax.y, ay.y, az.y, bx.y, by.y, bz.y, c1);
                    R t2 = enter(ax.z, ay.z, az.z, bx.z, by.z, bz.z, c2), t3 = enter(ax.w, ay.w, az.w, bx.w, by.w, bz.w, c3);
                    B2RT_CSWAP(t0, c0, t1, c1); B2RT_CSWAP(t2, c2, t3, c3);      // ascending by entry distance
                    B2RT_CSWAP(t0, c0, t2, c2); B2RT_CSWAP(t1, c1, t3, c3);
                    B2RT_CSWAP(t1, c1, t2, c2);
                    if (t3 < kMiss) push(c3);
                    if (t2 < kMiss) push(c2);
                    if (t1 < kMiss) push(c1);
                    ref = t0 < kMiss ? c0 : pop();
                }
            } else if constexpr (QUANT) {
                if (ref >= 0) {
                    if (COUNT) ++n_node;
                    const uint4 *p = reinterpret_cast<const uint4 *>(S.quant) + 2 + 2 * (size_t)ref;
                    const uint4 w0 = __ldg(p), w1 = __ldg(p + 1);      // (L.x L.y L.z R.x) (R.y R.z refL refR), lo | hi << 16
                    // 16-bit cell index -> float without the conversion unit: 0x4B000000 | q is the float 2^23 + q exactly
                    auto lo = [](unsigned w) { return R(__uint_as_float(__byte_perm(w, 0x4B000000u, 0x7610)) - 8388608.0f); };
                    auto hi = [](unsigned w) { return R(__uint_as_float(__byte_perm(w, 0x4B000000u, 0x7632)) - 8388608.0f); };
                    const R lx0 = fma_(lo(w0.x), qa.x, qb.x), lx1 = fma_(hi(w0.x), qa.x, qb.x);
                    const R ly0 = fma_(lo(w0.y), qa.y, qb.y), ly1 = fma_(hi(w0.y), qa.y, qb.y);
                    const R lz0 = fma_(lo(w0.z), qa.z, qb.z), lz1 = fma_(hi(w0.z), qa.z, qb.z);
                    const R rx0 = fma_(lo(w0.w), qa.x, qb.x), rx1 = fma_(hi(w0.w), qa.x, qb.x);
                    const R ry0 = fma_(lo(w1.x), qa.y, qb.y), ry1 = fma_(hi(w1.x), qa.y, qb.y);
                    const R rz0 = fma_(lo(w1.y), qa.z, qb.z), rz1 = fma_(hi(w1.y), qa.z, qb.z);
                    const R tl = max_(max_(min_(lx0, lx1), min_(ly0, ly1)), max_(min_(lz0, lz1), t_min));
                    const R fl = min_(min_(max_(lx0, lx1), max_(ly0, ly1)), min_(max_(lz0, lz1), best.t));
                    const R tr = max_(max_(min_(rx0, rx1), min_(ry0, ry1)), max_(min_(rz0, rz1), t_min));
                    const R fr = min_(min_(max_(rx0, rx1), max_(ry0, ry1)), min_(max_(rz0, rz1), best.t));
                    const bool hl = tl <= fl, hr = tr <= fr;
                    const int cl = (int)w1.z, cr = (int)w1.w;
                    if (hl && hr) {
                        const bool swap = tr < tl;
                        push(swap ? cl : cr);
                        ref = swap ? cr : cl;
                    } else if (hl) ref = cl;
                    else if (hr) ref = cr;
                    else ref = pop();
                }
            } else {
#pragma unroll
            for (int rep = 0; rep < B2RT_WALK_NODE_STEPS; ++rep) {
            if (ref >= 0) {
                if (COUNT) ++n_node;
                float4 n0, n1, n2, n3;
                if (ref < n_stage) {
                    const float4 *p = s_top + 4 * ref;
                    n0 = p[0]; n1 = p[1]; n2 = p[2]; n3 = p[3];
                } else {
                    const float4 *p = ref < S.n_top ? S.top + 4 * ref : S.nodes + 4 * (size_t)(ref - S.n_top);
                    if constexpr (B2RT_WALK_LDG256 != 0) {
                        ldg8(p, n0, n1); ldg8(p + 2, n2, n3);
                    } else if constexpr ((B2RT_L2_HINT & 1) != 0) {
                        const unsigned long long pol = l2_keep_policy();
                        n0 = ldg4_keep(p, pol); n1 = ldg4_keep(p + 1, pol); n2 = ldg4_keep(p + 2, pol); n3 = ldg4_keep(p + 3, pol);
                    } else {
                        n0 = __ldg(p); n1 = __ldg(p + 1); n2 = __ldg(p + 2); n3 = __ldg(p + 3);
                    }
                }
                // slab distances in the subtraction form the builder's box pad is sized for.  (The one-FMA form
                // b * (1/d) - o * (1/d) is NOT conservative: measured 16 differing closest hits in 72.7 M rays on the
                // 1 M-triangle scene, at the same speed — the walk is not issue-bound.)
                const R lx0 = (R(n0.x) - r.o.x) * id.x, lx1 = (R(n0.w) - r.o.x) * id.x;
                const R ly0 = (R(n0.y) - r.o.y) * id.y, ly1 = (R(n1.x) - r.o.y) * id.y;
                const R lz0 = (R(n0.z) - r.o.z) * id.z, lz1 = (R(n1.y) - r.o.z) * id.z;
                const R rx0 = (R(n1.z) - r.o.x) * id.x, rx1 = (R(n2.y) - r.o.x) * id.x;
                const R ry0 = (R(n1.w) - r.o.y) * id.y, ry1 = (R(n2.z) - r.o.y) * id.y;
                const R rz0 = (R(n2.x) - r.o.z) * id.z, rz1 = (R(n2.w) - r.o.z) * id.z;
                const R tl = max_(max_(min_(lx0, lx1), min_(ly0, ly1)), max_(min_(lz0, lz1), t_min));
                const R fl = min_(min_(max_(lx0, lx1), max_(ly0, ly1)), min_(max_(lz0, lz1), best.t));
                const R tr = max_(max_(min_(rx0, rx1), min_(ry0, ry1)), max_(min_(rz0, rz1), t_min));
                const R fr = min_(min_(max_(rx0, rx1), max_(ry0, ry1)), min_(max_(rz0, rz1), best.t));
                const bool hl = tl <= fl, hr = tr <= fr;
                const int cl = __float_as_int(n3.x), cr = __float_as_int(n3.y);
                if (hl && hr) {
                    const bool swap = tr < tl;
                    push(swap ? cl : cr);
                    ref = swap ? cr : cl;
                } else if (hl) ref = cl;
                else if (hr) ref = cr;
                else ref = pop();
            }
            }
            }
        } else if (want_leaf) {
            if (COUNT) ++n_leaf;
            test_prim<R, false>(S, ~ref, r, t_min, best);
            ref = pop();
        }
    }
    if (pos >= 0) {
        const real4<R> rec = Real4<R>::make(best.t, pack_int<R>((int64_t)best.prim), best.a, best.b);
        if constexpr ((B2RT_L2_HINT & 4) != 0) st_stream(hit + pos, rec); else hit[pos] = rec;
    }
    if (COUNT) { warp_flush(tally, n_node); warp_flush(tally + 1, n_leaf); }
}

// cuda_sample_hemisphere_cosine (:139-180)
template <typename R, typename Rng>
__device__ __forceinline__ V3<R> cos_hemisphere(V3<R> n, uint64_t &rng) {
#ifndef B2RT_OPT_RNG16
#define B2RT_OPT_RNG16 1           // 0: one generator step per random number in the float32 / PCG kernels as well
#endif
    if constexpr (sizeof(R) == 4 && B2RT_OPT_RNG16 && std::is_same<Rng, PcgRng>::value) {
        // float32 production with the counter-based generator: cosine-weighted direction = normalize(n + s), s uniform
        // on the unit sphere (exact; no tangent frame), both coordinates of s from ONE generator step (16 bits each,
        // cell centres: unbiased for anything smoother than 2^-16).  ~35 instructions instead of ~64.
        const uint32_t w = PcgRng::word(rng);
        rng = PcgRng::advance(rng);
        const float z = fmaf((float)(w >> 16), -2.0f / 65536.0f, 1.0f - 1.0f / 65536.0f);       // 1 - 2 (k + 0.5) / 65536
        const float phi = fmaf((float)(w & 0xffffu), 6.2831853071795865f / 65536.0f, -3.14159265358979f + 3.14159265358979f / 65536.0f);
        const float sr = sqrt_(fmaxf(fmaf(-z, z, 1.0f), 0.0f));
        const float dx = fmaf(sr, __cosf(phi), n.x), dy = fmaf(sr, __sinf(phi), n.y), dz = z + n.z;
        const float il = rsqrtf(fmaxf(dx * dx + dy * dy + dz * dz, 1e-12f));
        return {dx * il, dy * il, dz * il};
    }
    R r1 = Rng::template random<R>(rng); rng = Rng::advance(rng);
    R r2 = Rng::template random<R>(rng); rng = Rng::advance(rng);
    R ct = sqrt_(r1), st = sqrt_(R(1) - r1);
    R sp, cp;
#ifndef B2RT_OPT_ONB
#define B2RT_OPT_ONB 1             // 0: reference-order tangent frame + sincospif (measurement switch)
#endif
    if constexpr (sizeof(R) == 4 && B2RT_OPT_ONB) {
        // float32 production: MUFU sin/cos on phi - pi in [-pi, pi) (abs error 2^-21; the azimuth only has to be
        // uniform) and the branch-free orthonormal basis of Duff et al. 2017 instead of cross / normalise / cross:
        // the same cosine-weighted distribution about n with ~45 fewer instructions
        const float phi = fmaf(r2, 6.2831853071795865f, -3.14159265358979f);
        sp = __sinf(phi); cp = __cosf(phi);
        const float x = st * cp, y = st * sp, z = ct;
        const float sg = copysignf(1.0f, n.z);
        const float a = -rcp_approx(sg + n.z), b = n.x * n.y * a;
        const V3<R> u = {1.0f + sg * n.x * n.x * a, sg * b, -sg * n.x}, v = {b, sg + n.y * n.y * a, -n.y};
        return {x * u.x + y * v.x + z * n.x, x * u.y + y * v.y + z * n.y, x * u.z + y * v.z + z * n.z};
    } else if constexpr (sizeof(R) == 4) sincospif(2.0f * r2, &sp, &cp);
    else { R phi = R(2.0) * R(3.141592653589793) * r2; sp = sin(phi); cp = cos(phi); }
    R x = st * cp, y = st * sp, z = ct;
    V3<R> t = abs_(n.z) > R(0.9) ? V3<R>{R(1), R(0), R(0)} : V3<R>{R(0), R(0), R(1)};
    V3<R> u = {t.y * n.z - t.z * n.y, t.z * n.x - t.x * n.z, t.x * n.y - t.y * n.x};
    R ul = length(u);
    u = div3(u, ul);
    V3<R> v = {n.y * u.z - n.z * u.y, n.z * u.x - n.x * u.z, n.x * u.y - n.y * u.x};
    return {x * u.x + y * v.x + z * n.x, x * u.y + y * v.y + z * n.y, x * u.z + y * v.z + z * n.z};
}

// Warp-aggregated append to BOTH queues: one 64-bit atomicAdd per warp moves the ray-queue tail (low word)
// and the shadow-queue tail (high word); slots are handed out by prefix popcount of the ballots.
// (Measured alternatives, 1080p x 128 spp: two 32-bit atomics 55.4 ms, this 53.6 ms, CTA-aggregated with
// three barriers 60.9 ms for the fused kernel — the atomics are not the limiter, barriers cost more.)
__device__ __forceinline__ void warp_append2(unsigned long long *counter, bool want_ray, bool want_shadow,
                                             int &ray_slot, int &shadow_slot) {
    unsigned mr = __ballot_sync(0xffffffffu, want_ray), ms = __ballot_sync(0xffffffffu, want_shadow);
    ray_slot = shadow_slot = -1;
    if ((mr | ms) == 0) return;
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(counter, (unsigned long long)__popc(mr) | ((unsigned long long)__popc(ms) << 32));
#ifdef B2RT_TEST_ATOMIC2            // experiment: a second, no-op atomic on the SAME word (is the queue tail's L2 atomic unit the limit?)
    if (lane == 0) atomicAdd(counter, 0ULL);
#endif
    base = __shfl_sync(0xffffffffu, base, 0);
    if (want_ray) ray_slot = (int)(base & 0xffffffffULL) + __popc(mr & lt);
    if (want_shadow) shadow_slot = (int)(base >> 32) + __popc(ms & lt);
}
// ------------------------------------------------------------------------------------ shade
template <typename R> struct Segment {       // what one loop iteration of cuda_trace_path produces
    bool alive, want_shadow, culled;
    int light;
    V3<R> new_o, new_d, thr, s_o, s_d, s_c;
    uint64_t rng;
};

// One loop iteration of cuda_trace_path (:229-469) for one path: sky / texture / NEE shadow-ray
// emission / Russian roulette / BSDF sampling.  thr and rng come in through g and are updated.
template <typename R, typename Rng, bool FIRST, bool GENERIC_HINT, bool SURF, bool HITS_ONLY = false>
__device__ __forceinline__ void shade_segment(const SceneDev &S, const PathQueues<R> &Q, const float4 *s_scan,
                                              const float4 *s_surf, const Ray<R> &r, const Hit<R> &h, int slot, int bounce, int max_depth,
                                              Segment<R> &g) {
    V3<R> &thr = g.thr, &new_o = g.new_o, &new_d = g.new_d, &s_o = g.s_o, &s_d = g.s_d, &s_c = g.s_c;
    uint64_t &rng = g.rng;
    bool &alive = g.alive, &want_shadow = g.want_shadow;
    constexpr bool RNG16 = B2RT_OPT_RNG16 && sizeof(R) == 4 && std::is_same<Rng, PcgRng>::value;
    float choice16 = 0.f;
    if (FIRST) {                                                            // first touch of L[slot]
        R sky = h.prim < 0 ? R(0.1) : R(0);
        Q.L[slot] = Real4<R>::make(sky, sky, sky, R(0));
    }
    if (!HITS_ONLY && h.prim < 0) {                                         // :234-239 sky
#if B2RT_OPT_SKYQ
        // An escaping path has no shadow ray of its own at this bounce, so its sky term rides in the lane's free
        // shadow-queue slot as a PRE-RESOLVED record (light index -1): the shadow kernel of this bounce adds it to
        // L[slot] — same position in the per-slot add order as before, bit-identical sums.  The load -> add -> store
        // chain on a DRAM-resident line left the bounce kernel (12.9 % of its stall samples, profiles/r2a_*).
        if (!FIRST) { want_shadow = true; g.light = -1; s_o = r.o; s_d = r.d; s_c = {thr.x * R(0.1), thr.y * R(0.1), thr.z * R(0.1)}; }
#else
        if (!FIRST) add_sky(Q.L + slot, thr.x * R(0.1), thr.y * R(0.1), thr.z * R(0.1));
#endif
    } else {
        Surface<R> sf;
        if constexpr (SURF && sizeof(R) == 4) make_surface_small(s_surf, r, h, sf);
        else make_surface<R, false>(S, r, h, sf);
        // the texel load is issued here and decoded after the light-sample geometry and the occluder test below,
        // which do not depend on it: the L2 round trip overlaps ~100 instructions of independent work
        const bool textured = sf.tex >= 0 && sf.tex < S.n_tex;
        uint32_t texel = 0;
        if (textured) texel = fetch_texel<R, false>(S, sf.tex, sf.u, sf.v);
        V3<R> mc = sf.color;
        V3<R> po = sf.p + sf.n * R(0.001);
        if (S.n_lights > 0) {                                               // :265-304
            R nl = R(S.n_lights);
            int li;
            if (RNG16 && S.n_lights <= 4096) {               // (16 bits resolve up to a few thousand light samples evenly)
                // one generator step serves the light pick (low 16 bits) and the lobe choice below (high 16 bits)
                const uint32_t w = PcgRng::word(rng);
                rng = PcgRng::advance(rng);
                li = (int)(((w & 0xffffu) * (uint32_t)S.n_lights) >> 16);
                choice16 = fmaf((float)(w >> 16), 1.0f / 65536.0f, 0.5f / 65536.0f);
            } else {
                li = (int)(Rng::template random<R>(rng) * nl);
                if (li >= S.n_lights) li = S.n_lights - 1;
                rng = Rng::advance(rng);
            }
            V3<R> l = xyz<R>(ldg4(reinterpret_cast<const real4<R> *>(S.lights) + li)) - sf.p;
            R dist = length(l);
            if (dist > R(0.001)) l = div3(l, dist);
            R pdf = R(1) / nl;
            R ct = max_(R(0), l.x * sf.n.x + l.y * sf.n.y + l.z * sf.n.z);
            R li_, lm;
            if (sf.refractive > R(0.5)) { li_ = R(4.0); lm = R(0.6); }
            else if (sf.reflective > R(0.7)) { li_ = R(2.5); lm = R(0.8); }
            else { li_ = R(2.0); lm = R(1.0); }
            s_o = po; s_d = l;
            g.light = li;
            bool blocked = false;
            if constexpr (sizeof(R) == 4) {
                // Occluder hint: test the primitive that blocks most shadow rays to this light sample first.
                // If it blocks this ray the full occlusion query would also say "occluded", so the ray is
                // answered here and never queued (exact, not an approximation).
                if (ct * sf.diffuse != 0.f && S.occl_hint && (GENERIC_HINT || s_scan)) {
                    Ray<float> sr; sr.o = s_o; sr.d = s_d;
                    const int code = __ldg(S.occl_hint + li);
                    blocked = occluder_test<GENERIC_HINT>(S, s_scan, code, sr, 0.001f, 1000000.0f);
                }
            }
            if (textured) mc = decode_texel<R>(texel);
            if constexpr (sizeof(R) == 4) {
                R k = sf.diffuse * ct * li_ * lm * nl;
                s_c = {thr.x * (mc.x * k), thr.y * (mc.y * k), thr.z * (mc.z * k)};
            } else {
                s_c = {thr.x * (mc.x * sf.diffuse * ct * li_ * lm / pdf),
                       thr.y * (mc.y * sf.diffuse * ct * li_ * lm / pdf),
                       thr.z * (mc.z * sf.diffuse * ct * li_ * lm / pdf)};
            }
            // a shadow ray whose payload is exactly zero cannot change the image: not queued
            want_shadow = (s_c.x != R(0)) || (s_c.y != R(0)) || (s_c.z != R(0));
            if (want_shadow && blocked) { want_shadow = false; g.culled = true; }
        } else if (textured) mc = decode_texel<R>(texel);
        bool go = true;
        if (bounce >= 3) {                                                  // :307-314
            R p = max_(R(0.1), R(0.299) * thr.x + R(0.587) * thr.y + R(0.114) * thr.z);
            if (Rng::template random<R>(rng) > p) go = false;
            else { rng = Rng::advance(rng); thr = div3(thr, p); }
        }
        if (go) {
            R choice;                                                       // :317-318
            if (RNG16 && S.n_lights > 0 && S.n_lights <= 4096) choice = R(choice16);
            else { choice = Rng::template random<R>(rng); rng = Rng::advance(rng); }
            R dn = r.d.x * sf.n.x + r.d.y * sf.n.y + r.d.z * sf.n.z;
            V3<R> refl = {r.d.x - R(2) * dn * sf.n.x, r.d.y - R(2) * dn * sf.n.y, r.d.z - R(2) * dn * sf.n.z};
            new_o = po;
            bool lambert = false;                                           // one hemisphere-sample call site
            if (sf.refractive > R(0.1)) {                                   // :320-428 glass
                if (choice < R(0.6)) {
                    R cos_i = max_(R(0), -dn);
                    bool entering = cos_i > R(0);
                    V3<R> on = entering ? sf.n : -sf.n;
                    R eta = entering ? rcp_(sf.ior) : sf.ior;
                    V3<R> rr;
                    if (refract_nb<R>(r.d, on, eta, rr)) {
                        if (entering) new_o = sf.p - sf.n * R(0.001);
                        new_d = rr;
                        thr = thr * (sf.refractive / R(0.6));
                    } else { new_d = refl; thr = thr * R(0.9); }
                } else if (choice < R(0.6) + R(0.25)) {
                    new_d = refl;
                    thr = {thr.x * (mc.x * R(0.9) / R(0.25)), thr.y * (mc.y * R(0.9) / R(0.25)),
                           thr.z * (mc.z * R(0.9) / R(0.25))};
                } else {
                    lambert = true;
                    thr = {thr.x * (mc.x * sf.diffuse * R(3.0) / R(0.15)), thr.y * (mc.y * sf.diffuse * R(3.0) / R(0.15)),
                           thr.z * (mc.z * sf.diffuse * R(3.0) / R(0.15))};
                }
            } else if (sf.reflective > R(0.5)) {                            // :430-449 mirror
                new_d = refl;
                thr = {thr.x * (mc.x * sf.reflective), thr.y * (mc.y * sf.reflective), thr.z * (mc.z * sf.reflective)};
            } else {                                                        // :451-466 diffuse
                lambert = true;
                thr = {thr.x * (mc.x * sf.diffuse), thr.y * (mc.y * sf.diffuse), thr.z * (mc.z * sf.diffuse)};
            }
            if (lambert) new_d = cos_hemisphere<R, Rng>(sf.n, rng);
            alive = !(max_(thr.x, max_(thr.y, thr.z)) < R(0.001))           // :468
                    && (bounce + 1 < max_depth);                            // :229 loop bound
        }
    }
}

#ifndef B2RT_PRIMARY_TILES
#define B2RT_PRIMARY_TILES 1
#endif
#ifndef B2RT_OPT_SKYQ
#define B2RT_OPT_SKYQ 0            // 1: escaping paths hand their sky term to the shadow kernel through the shadow queue instead of a
                                   // read-modify-write of L[slot] in the bounce kernel.  Measured (profiles/r2b): bounce kernels 19.9 ->
                                   // 19.4 ms per 128 spp but the shadow kernels 2.0 -> 4.0 ms: the stall samples on that RMW were hidden
#endif
#ifndef B2RT_OPT_RING
#define B2RT_OPT_RING 0            // 1: the per-warp ring of pending hits below.  MEASURED AND OFF: bounce kernels 34.8 -> 41.4 ms
                                   // per 256 spp (profiles/r2_hit_ring_ab.log) — the kernel grows from 28.8 to 33.3 KB of SASS (past
                                   // the 32 KB instruction cache) and spills 116 B at its 64-register budget, which costs more than
                                   // the full shading lanes win
#endif
// Per-warp ring of PENDING HITS (MODE 3).  A quarter of the rays of bounce >= 1 leave the Cornell box through its open
// front, so shading — 38 % of the instructions — ran with 19 of 32 lanes (profiles/r2_ncu_full_c2.csv).  Each warp now
// pushes the hits of an iteration into a 64-entry ring in shared memory (structure of arrays: one bank per lane, no
// conflicts), answers its misses on the spot, and shades only when 32 hits are pending: every shading pass has all
// lanes on a hit, and every fourth iteration does not shade at all.  Warp-private: __syncwarp, no barrier.
constexpr int kRingWords = 14;     // P.xyz, d.xyz, thr.xyz, prim, a, b, slot, rng
constexpr int kRingSlots = 64;
#ifndef B2RT_OPT_CHUNKED
#define B2RT_OPT_CHUNKED 1         // 0: one queue-tail atomic per warp iteration (warp_append2) in the small-scene kernels too
#endif
#ifndef B2RT_OPT_ASYNC
#define B2RT_OPT_ASYNC 1           // 0: plain streaming loads of the ray records at the top of each iteration
#endif
// Asynchronous double-buffered ray-record fetch (MODE 3): every thread copies the three 16 B records of its NEXT
// grid-stride item global -> shared with cp.async (LDGSTS, no register staging) while it scans and shades the current
// one; it only ever reads back its own slots, so cp.async.wait_group is all the synchronisation there is.
// profiles/r2a_*: 15 % of the bounce kernel's stall samples sat on the first use of the just-issued queue loads.
constexpr int kAsyncStreams = B2RT_OPT_RING ? 2 : 3;  // with the hit ring the throughput record is loaded directly: it is first
                                                      // needed after the scan, and the ring needs the shared memory
constexpr int kAsyncStageF4 = kAsyncStreams * 256;    // float4 per stage
constexpr int kAsyncStageBytes = kAsyncStageF4 * 16;  // per stage: streams x 256 threads x 16 B
constexpr int kRingBytes = B2RT_OPT_RING ? 8 * kRingWords * kRingSlots * 4 : 0;       // 8 warps per CTA
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
#ifndef B2RT_OPT_SURF
#define B2RT_OPT_SURF 1            // 0: generic make_surface in the small-scene kernels (measured 23.8 vs 22.3 ms)
#endif
// Exact unsigned division by a launch-invariant divisor: one multiply-high + shift (the generic 32-bit division of
// i / npix, pix / W costs ~20 instructions each, ~50 of the ~600 a camera-ray warp issues).  Round-up method of
// Granlund & Montgomery / libdivide "branchfree": q = (t + ((n - t) >> 1)) >> (sh - 1) with t = mulhi(n, m).
struct FastDiv {
    unsigned d, m, sh;
    static FastDiv make(unsigned d) {
        FastDiv f; f.d = d; f.m = 0; f.sh = 0;
        if (d <= 1) return f;
        unsigned sh = 0;
        while ((1ull << sh) < d) ++sh;
        f.sh = sh;
        f.m = (unsigned)(((1ull << 32) * ((1ull << sh) - d)) / d + 1);
        return f;
    }
    __device__ __forceinline__ unsigned div(unsigned n) const {
        if (d <= 1) return n;
        const unsigned t = __umulhi(n, m);
        return (t + ((n - t) >> 1)) >> (sh - 1);
    }
};

// ------------------------------------------------------------------------------------ camera-ray candidate masks
// Small scenes, bounce 0: the 32 camera rays of a warp belong to 32 neighbouring pixels of one row and all samples of
// a pixel stay inside its footprint, so most of the scan records cannot be hit at all (a Cornell pixel sees one or two
// of the seven).  Once per render call every 32-pixel tile gets a bit mask of the records whose bounds reach into the
// tile's pyramid (tile widened by half a pixel; a record is dropped only when all its corners lie outside ONE of the
// four side planes by a relative margin: conservative).  Tiles with an empty mask (51 % of the Cornell image) are never
// visited by bounce 0 at all: it iterates over the compact list of non-empty tiles, and accumulate_kernel adds their
// sky samples itself (the same float additions in the same order, without L[slot] ever being written or read).
// stats[0] += canonical flops of the masked record tests of one sample of every pixel, stats[1] += pixels with an
// empty mask (bench.py: executed FP32 work).
#ifndef B2RT_OPT_MASKS
#define B2RT_OPT_MASKS 1
#endif
#ifndef B2RT_OPT_TILE_LIST
#define B2RT_OPT_TILE_LIST 1       // 0: bounce 0 still visits every pixel (row-major) and looks its tile's mask up
#endif
__device__ __forceinline__ bool inv3(const float a[3][3], float inv[3][3]) {
    const float c00 = a[1][1] * a[2][2] - a[1][2] * a[2][1], c01 = a[1][2] * a[2][0] - a[1][0] * a[2][2],
                c02 = a[1][0] * a[2][1] - a[1][1] * a[2][0];
    const float det = a[0][0] * c00 + a[0][1] * c01 + a[0][2] * c02;
    if (!(fabsf(det) > 1e-30f)) return false;
    const float id = 1.0f / det;
    inv[0][0] = c00 * id; inv[0][1] = (a[0][2] * a[2][1] - a[0][1] * a[2][2]) * id; inv[0][2] = (a[0][1] * a[1][2] - a[0][2] * a[1][1]) * id;
    inv[1][0] = c01 * id; inv[1][1] = (a[0][0] * a[2][2] - a[0][2] * a[2][0]) * id; inv[1][2] = (a[0][2] * a[1][0] - a[0][0] * a[1][2]) * id;
    inv[2][0] = c02 * id; inv[2][1] = (a[0][1] * a[2][0] - a[0][0] * a[2][1]) * id; inv[2][2] = (a[0][0] * a[1][1] - a[0][1] * a[1][0]) * id;
    return true;
}
static __global__ void __launch_bounds__(128)
primary_mask_kernel(SceneDev S, Cam<float> cam, int W, int H, unsigned *__restrict__ masks, unsigned long long *stats) {
    const int n_tiles = (W / 32) * H;
    unsigned flops = 0, empty = 0;
    for (int tile = blockIdx.x * blockDim.x + threadIdx.x; tile < n_tiles; tile += gridDim.x * blockDim.x) {
        const int pix0 = tile * 32, y = pix0 / W, x0 = pix0 - y * W;
        const float u0 = (x0 - 0.5f) / W, u1 = (x0 + 32.5f) / W, v0 = (y - 0.5f) / H, v1 = (y + 1.5f) / H;
        const float us[4] = {u0, u1, u1, u0}, vs[4] = {v0, v0, v1, v1};
        float c[4][3], cc[3] = {0.f, 0.f, 0.f};
        for (int k = 0; k < 4; ++k) {
            c[k][0] = cam.llc.x + us[k] * cam.hor.x + vs[k] * cam.ver.x - cam.origin.x;
            c[k][1] = cam.llc.y + us[k] * cam.hor.y + vs[k] * cam.ver.y - cam.origin.y;
            c[k][2] = cam.llc.z + us[k] * cam.hor.z + vs[k] * cam.ver.z - cam.origin.z;
            cc[0] += c[k][0]; cc[1] += c[k][1]; cc[2] += c[k][2];
        }
        float n[4][3];                                   // unit inward normals of the four side planes (through the eye)
        for (int k = 0; k < 4; ++k) {
            const float *a = c[k], *b = c[(k + 1) & 3];
            float nx = a[1] * b[2] - a[2] * b[1], ny = a[2] * b[0] - a[0] * b[2], nz = a[0] * b[1] - a[1] * b[0];
            float sgn = (nx * cc[0] + ny * cc[1] + nz * cc[2]) < 0.f ? -1.f : 1.f;
            float il = sgn * rsqrtf(fmaxf(nx * nx + ny * ny + nz * nz, 1e-30f));
            n[k][0] = nx * il; n[k][1] = ny * il; n[k][2] = nz * il;
        }
        const float ex = cam.origin.x, ey = cam.origin.y, ez = cam.origin.z;
        const float kEps = 1e-4f;
        unsigned mask = 0, bit = 0;
        // box record j: corners C +- h0 +- h1 +- h2 with (h0 h1 h2) = M^-1, C = -M^-1 d
        const float4 *bx = S.scan + 4 * S.n_scan;
        for (int j = 0; j < S.n_box; ++j, ++bit) {
            const float4 q0 = __ldg(bx + 4 * j), q1 = __ldg(bx + 4 * j + 1), q2 = __ldg(bx + 4 * j + 2);
            const float m[3][3] = {{q0.x, q0.y, q0.z}, {q1.x, q1.y, q1.z}, {q2.x, q2.y, q2.z}};
            float h[3][3];
            bool keep = true;
            if (inv3(m, h)) {
                const float cx = -(h[0][0] * q0.w + h[0][1] * q1.w + h[0][2] * q2.w) - ex,
                            cy = -(h[1][0] * q0.w + h[1][1] * q1.w + h[1][2] * q2.w) - ey,
                            cz = -(h[2][0] * q0.w + h[2][1] * q1.w + h[2][2] * q2.w) - ez;
                for (int k = 0; k < 4 && keep; ++k) {
                    float ext = 0.f, scale = fabsf(cx) + fabsf(cy) + fabsf(cz);
                    for (int a = 0; a < 3; ++a) {        // column a of M^-1 is half axis a
                        ext += fabsf(n[k][0] * h[0][a] + n[k][1] * h[1][a] + n[k][2] * h[2][a]);
                        scale += fabsf(h[0][a]) + fabsf(h[1][a]) + fabsf(h[2][a]);
                    }
                    if (n[k][0] * cx + n[k][1] * cy + n[k][2] * cz + ext < -kEps * scale) keep = false;
                }
            }
            if (keep) { mask |= 1u << bit; flops += 42u; }
        }
        // loose planar record k: P(u, v) = A^-1 (cN, u - d1, v - d2), corners (0 | umax) x (0 | vmax)
        for (int k2 = 0; k2 < S.n_loose; ++k2, ++bit) {
            const float4 q0 = __ldg(S.scan + 4 * k2), q1 = __ldg(S.scan + 4 * k2 + 1), q2 = __ldg(S.scan + 4 * k2 + 2),
                         q3 = __ldg(S.scan + 4 * k2 + 3);
            const float m[3][3] = {{q0.x, q0.y, q0.z}, {q1.x, q1.y, q1.z}, {q2.x, q2.y, q2.z}};
            float h[3][3];
            bool keep = true;
            if (inv3(m, h)) {
                const float px = h[0][0] * q0.w - h[0][1] * q1.w - h[0][2] * q2.w - ex,
                            py = h[1][0] * q0.w - h[1][1] * q1.w - h[1][2] * q2.w - ey,
                            pz = h[2][0] * q0.w - h[2][1] * q1.w - h[2][2] * q2.w - ez;
                for (int k = 0; k < 4 && keep; ++k) {
                    const float du = q3.x * (n[k][0] * h[0][1] + n[k][1] * h[1][1] + n[k][2] * h[2][1]);
                    const float dv = q3.y * (n[k][0] * h[0][2] + n[k][1] * h[1][2] + n[k][2] * h[2][2]);
                    const float scale = fabsf(px) + fabsf(py) + fabsf(pz) +
                                        q3.x * (fabsf(h[0][1]) + fabsf(h[1][1]) + fabsf(h[2][1])) +
                                        q3.y * (fabsf(h[0][2]) + fabsf(h[1][2]) + fabsf(h[2][2]));
                    if (n[k][0] * px + n[k][1] * py + n[k][2] * pz + fmaxf(du, 0.f) + fmaxf(dv, 0.f) < -kEps * scale) keep = false;
                }
            }
            if (keep) { mask |= 1u << bit; flops += 33u; }
        }
        const float4 *sph = reinterpret_cast<const float4 *>(S.sphere);
        for (int i = 0; i < S.n_sphere; ++i, ++bit) {
            const float4 s0 = __ldg(sph + 2 * i);
            const float cx = s0.x - ex, cy = s0.y - ey, cz = s0.z - ez;
            bool keep = true;
            for (int k = 0; k < 4 && keep; ++k)
                if (n[k][0] * cx + n[k][1] * cy + n[k][2] * cz + fabsf(s0.w) <
                    -kEps * (fabsf(cx) + fabsf(cy) + fabsf(cz) + fabsf(s0.w))) keep = false;
            if (keep) { mask |= 1u << bit; flops += 28u; }
        }
        masks[tile] = mask;
        empty += mask == 0u ? 32u : 0u;
    }
    // per-pixel figures: every one of the tile's 32 pixels tests the tile's records
    warp_flush(stats, flops * 32u);
    warp_flush(stats + 1, empty);
}

// Ordered compaction of the non-empty tiles (one CTA; 64 800 tiles at 1080p): bounce 0 walks this list in image order, so
// concurrently running warps stay on neighbouring pixels, texels and L[slot] lines exactly as in the untiled sweep.
// (A list in atomic-arrival order ran bounce 1 1.7x slower: profiles/r2e-r2f.)  *count = list length.
static __global__ void __launch_bounds__(1024)
tile_compact_kernel(const unsigned *__restrict__ masks, int n_tiles, int *__restrict__ tiles, unsigned long long *count) {
    __shared__ int s_sum[1024];
    const int t = threadIdx.x, chunk = (n_tiles + 1023) / 1024, lo = min(t * chunk, n_tiles), hi = min(lo + chunk, n_tiles);
    int c = 0;
    for (int i = lo; i < hi; ++i) c += masks[i] != 0u;
    s_sum[t] = c;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {                   // inclusive Hillis-Steele scan
        const int v = t >= off ? s_sum[t - off] : 0;
        __syncthreads();
        s_sum[t] += v;
        __syncthreads();
    }
    int pos = s_sum[t] - c;
    for (int i = lo; i < hi; ++i) if (masks[i] != 0u) tiles[pos++] = i;
    if (t == 1023) {
        const unsigned d = (unsigned)s_sum[1023];
        count[0] = (unsigned long long)d;
    }
}

template <typename R> struct PrimaryArgs {   // MODE 4: camera-ray generation fused into the first bounce
    Cam<R> cam;
    int W, H, spp_wave;
    FastDiv by_npix, by_w, by_tiles;
    int tiles_x;                                 // > 0: 8 x 4 pixel tiles per warp (W % 8 == 0 and H % 4 == 0)
    long long first_sample;
    unsigned long long seed;
    const unsigned *masks;                       // MODE 5: per-32-pixel-tile candidate masks (or nullptr)
    const int *tiles;                            //         compact list of the non-empty tiles ...
    const unsigned long long *n_tiles;           //         ... and its length (device memory: no host sync)
};

// MODE 0: wavefront "shade" stage reading the hit stream written by extend_kernel.
// MODE 4: bounce 0 with the counter-based RNG — the camera ray is generated in-register (no raygen kernel,
// no 64 B/path queue round trip), walked through the LBVH and shaded; L[slot] is initialised here.
// MODE 1/2/3: fused extend+shade — the closest hit is found in-register (1: LBVH walk, 2: warp-uniform scan of
// all primitives with the generic tests, 3: the float32 planar scan records) and shaded at once, so the FP32-issue-bound intersection work overlaps the
// latency-bound shading loads in one kernel and the hit stream (32 B/segment) never touches HBM.
// MODE 5: bounce 0 of a small scene: camera ray generated in-register and intersected with the scan/box records.
// MODE 6: MODE 4 for small float32 scenes (surface records + scan-record occluder hints; MODE 4 itself then only
// carries the generic streams).  Each variant compiles ONE shading flavour: the fused kernels sit close to the
// instruction-cache cliff (measured twice: ~27.6 KB of SASS ran at 44 ms where ~26.6 KB ran at 24 ms).
template <typename R, typename Rng, int MODE>
__global__ void __launch_bounds__(256, sizeof(R) != 4 ? 1 : ((MODE == 3 || MODE == 5) ? B2RT_BOUNCE_MIN_BLOCKS : B2RT_BVH_MIN_BLOCKS))
shade_kernel(SceneDev S, PathQueues<R> Q, int in_buf, int bounce, int max_depth, PrimaryArgs<R> P) {
    // MODE 7: consumer of the hit queue (split small-scene bounce): no intersection at all, every item is a hit
    constexpr bool HITQ = MODE == 7;
    constexpr bool PRIMARY = MODE == 4 || MODE == 5 || MODE == 6, WALK = MODE == 1 || MODE == 4 || MODE == 6,
                   PLANAR = MODE == 3 || MODE == 5, SURF = B2RT_OPT_SURF && sizeof(R) == 4 && (PLANAR || MODE == 6 || HITQ);
    // effective SM clock as the kernels see it: CTA 0's cycle counter against the global nanosecond timer (some GPUs of
    // this pool run sustained FP32 load ~30 % slower at an unchanged nvidia-smi clock reading: bench.py reports both)
    long long clk0 = 0;
    unsigned long long ns0 = 0;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        clk0 = clock64();
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns0));
    }
    // shared memory: [BVH top levels (WALK)] [scan + box records] [surface records]
    extern __shared__ float4 s_top[];
    float4 *cursor = s_top;
    if (WALK) { stage_top(S, s_top); cursor += 4 * S.n_top; }
    const float4 *s_scan = nullptr, *s_surf = nullptr;
    // the scan records are what PLANAR modes intersect; MODE 6 and the generic scan (MODE 2) keep them for the
    // occluder hints.  Only these launches pay for the bytes (rt_api.cuh): MODE 0 / 1 / 4 carry generic hints
    // (codes >= 128) and never stage scan records
    constexpr bool STAGE_SCAN = PLANAR || MODE == 6 || MODE == 2 || HITQ;
    if (STAGE_SCAN && sizeof(R) == 4 && S.n_scan > 0 && S.scan_incoherent && (PLANAR || HITQ || S.occl_hint)) {
        stage_scan(S, cursor); s_scan = cursor; cursor += 4 * (S.n_scan + S.n_box);
    }
    if (SURF) { stage_surf(S, cursor); s_surf = cursor; cursor += 5 * S.n_prims; }
    constexpr bool ASYNC = B2RT_OPT_ASYNC && (MODE == 3 || HITQ) && sizeof(R) == 4;
    float4 *s_ray = cursor;                                                  // [2 stages][3 streams][256 threads]
    // the hit queue has the ray queue's first three record layouts (hit point for origin, primitive for depth)
    const real4<R> *__restrict__ ro = HITQ ? Q.ha : Q.ro[in_buf], *__restrict__ rd = HITQ ? Q.hb : Q.rd[in_buf],
                   *__restrict__ th = HITQ ? Q.hc : Q.th[in_buf];
    real4<R> *__restrict__ no = Q.ro[in_buf ^ 1], *__restrict__ nd = Q.rd[in_buf ^ 1], *__restrict__ nt = Q.th[in_buf ^ 1];
    // tile-list mode (MODE 5 with candidate masks): item i = ((sample * non-empty tiles + tile index) * 32 + lane)
    const bool TILED = PLANAR && PRIMARY && B2RT_OPT_MASKS && B2RT_OPT_TILE_LIST && P.masks != nullptr;
    const unsigned n_act = TILED ? (unsigned)(P.n_tiles[0] & 0xffffffffULL) : 1u;
    // (sample, tile index) of this warp's item, stepped along with the grid-stride loop instead of divided out per item
    unsigned t_s = 0, t_i = 0;
    const unsigned t_step = (gridDim.x * blockDim.x) >> 5;
    if (TILED) { const unsigned g0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; t_s = g0 / n_act; t_i = g0 - t_s * n_act; }
    int n = PRIMARY ? (TILED ? (int)n_act * 32 * P.spp_wave : P.W * P.H * P.spp_wave)
                    : (HITQ ? (int)(Q.hit_tail[bounce] & 0xffffffffULL) : ray_count(Q, bounce));
    // ray statistics: every camera ray counts as answered, whether its tile was visited or not
    if (PRIMARY && blockIdx.x == 0 && threadIdx.x == 0) Q.counts[0] = (unsigned long long)(P.W * P.H * P.spp_wave);
    int n_round = (n + 31) & ~31;                // whole warps iterate together (ballots in warp_append2)
    unsigned n_culled = 0, n_tally = 0;          // n_tally: shaded hits (low 16 bits) | bounds-culled camera rays << 16
    // queue appends go through per-warp chunks (one tail atomic per ~5 iterations instead of one per iteration)
    constexpr bool CHUNKED = B2RT_OPT_CHUNKED != 0;
    __shared__ WarpCursor s_wc[8];
    WarpCursor *wc = s_wc + (threadIdx.x >> 5);
    if (CHUNKED && (threadIdx.x & 31) == 0) { wc->ray_cur = wc->ray_end = wc->sh_cur = wc->sh_end = 0; }
    if (CHUNKED) __syncwarp();
    int stage = 0;
    constexpr bool RING = B2RT_OPT_RING && ASYNC;        // MODE 3, float32
    float *ring = nullptr;
    __shared__ int s_ring_ctl[8][2];                     // per warp: head, pending hits (lane 0 writes)
    int *ring_ctl = s_ring_ctl[threadIdx.x >> 5];
    if constexpr (RING) {
        ring = reinterpret_cast<float *>(s_ray + 2 * kAsyncStageF4) + (threadIdx.x >> 5) * (kRingWords * kRingSlots);
        if ((threadIdx.x & 31) == 0) { ring_ctl[0] = 0; ring_ctl[1] = 0; }
        __syncwarp();
    }
    if constexpr (ASYNC) {
        const int i0 = blockIdx.x * blockDim.x + threadIdx.x;
        if (i0 < n) {
            cp_async16(s_ray + threadIdx.x, ro + i0);
            cp_async16(s_ray + 256 + threadIdx.x, rd + i0);
            if (kAsyncStreams == 3) cp_async16(s_ray + 512 + threadIdx.x, th + i0);
        }
        cp_async_commit();
    }
    for (int i = blockIdx.x * blockDim.x + threadIdx.x;; i += gridDim.x * blockDim.x) {
        const bool more = i < n_round;                   // warp-uniform (n_round is a multiple of 32)
        if constexpr (RING) { if (!more && ring_ctl[1] == 0) break; }      // one extra pass flushes the last pending hits
        else { if (!more) break; }
        bool valid = more && i < n;
        real4<R> qa, qb, qc;
        if constexpr (ASYNC) {
            cp_async_wait_all();
            const float4 *cur = s_ray + stage * kAsyncStageF4 + threadIdx.x;
            if (valid) { qa = cur[0]; qb = cur[256]; qc = kAsyncStreams == 3 ? cur[512] : ld_stream(th + i); }
            stage ^= 1;
            const int nx = i + gridDim.x * blockDim.x;
            if (more && nx < n) {
                float4 *nxt = s_ray + stage * kAsyncStageF4 + threadIdx.x;
                cp_async16(nxt, ro + nx); cp_async16(nxt + 256, rd + nx);
                if (kAsyncStreams == 3) cp_async16(nxt + 512, th + nx);
            }
            cp_async_commit();
        }
        Segment<R> g;
        g.alive = false; g.want_shadow = false; g.culled = false; g.rng = 0; g.light = 0;
        int slot = 0;
        unsigned mask = 0xffffffffu;
        bool dead = false;
        bool shade_now = false;                          // this lane holds a ray + hit record to shade in this iteration
        int hitq_prim = -1;
        Ray<R> r;
        Hit<R> h;
        if (valid) {
            if (PRIMARY) {                       // cuda_path_trace_kernel's sample set-up (:35-41)
                const int npix = P.W * P.H;
                int s, pix, x, y;
                if (TILED) {
                    // sample-major like the untiled order: concurrently running warps touch neighbouring L[slot] lines.
                    // (Tile-major order — the warps of a CTA on the same tile, consecutive samples — put their L lines
                    // a multiple of npix * 16 B apart and ran bounce 0 AND bounce 1 ~60 % slower: profiles/r2e.)
                    s = (int)t_s;
                    const int tile = __ldg(P.tiles + t_i);
                    pix = tile * 32 + (i & 31);
                    mask = __ldg(P.masks + tile);
                } else {
                    s = (int)P.by_npix.div((unsigned)i); pix = i - s * npix;
                    if (PLANAR && B2RT_OPT_MASKS && P.masks) mask = __ldg(P.masks + (pix >> 5));
                }
#if B2RT_PRIMARY_TILES
                if (WALK && P.tiles_x > 0) {
                    // LBVH walk: a warp covers an 8 x 4 pixel tile instead of 32 pixels of one row, so neighbouring
                    // lanes walk the same nodes (same samples, same image; 1 M triangles: first bounce 20.5 -> 18.5 ms.
                    // The record scan of small scenes does the same work in every lane: tiles cost 0.7 % there)
                    const int tile = pix >> 5, in = pix & 31;
                    const int ty = (int)P.by_tiles.div((unsigned)tile), tx = tile - ty * P.tiles_x;
                    x = tx * 8 + (in & 7); y = ty * 4 + (in >> 3);
                    pix = y * P.W + x;
                } else
#endif
                { y = (int)P.by_w.div((unsigned)pix); x = pix - y * P.W; }
                uint64_t state = PcgRng::seed((uint32_t)pix, (uint64_t)(P.first_sample + s), P.seed);
                R rnd = PcgRng::template random<R>(state);
                state = PcgRng::advance(state);
                r = camera_ray<R>(P.cam, (R(x) + rnd) / R(P.W), (R(y) + rnd) / R(P.H));
                slot = s * npix + pix; g.rng = state; g.thr = {R(1), R(1), R(1)};
            } else {
                // fused walk modes read through the sort permutation; the wavefront stage (MODE 0) reads queue order
                real4<R> a, b, c;
                if constexpr (ASYNC) { a = qa; b = qb; c = qc; }
                else {
                    const int j = (MODE != 0 && Q.perm) ? __ldg(Q.perm + i) : i;
                    a = ld_stream(ro + j); b = ld_stream(rd + j); c = ld_stream(th + j);
                }
                r.o = xyz<R>(a); r.d = xyz<R>(b);
                slot = (int)unpack_u<R>(a.w);
                g.rng = unpack_u<R>(b.w);
                g.thr = xyz<R>(c);
                if (HITQ) hitq_prim = (int)unpack_u<R>(c.w);
                prefetch_l2(Q.L + slot);         // a miss adds the sky term to L[slot]: DRAM round trip started now
                dead = slot < 0;                 // unused remainder of a producer warp's last chunk
            }
            if (!dead) {
            if (HITQ) {
                // (the record's .w lane of the third stream was unpacked as "depth" into nothing: it is the primitive)
                const real4<R> ab = ld_stream(Q.hd + i);
                h.t = R(0); h.prim = hitq_prim; h.a = ab.x; h.b = ab.y;       // p = o + 0 d: the stored hit point
            } else if (MODE == 0) {
                real4<R> hrec = ld_stream(Q.hit + i);
                h.t = hrec.x; h.prim = (int)(long long)real_as_int(hrec.y); h.a = hrec.z; h.b = hrec.w;
            } else if (WALK) {
                traverse<R, false, false>(S, s_top, r, R(0.001), R(1000000.0), h);
            } else if (MODE == 2) {
                scan_all<R, false, false>(S, r, R(0.001), R(1000000.0), h);
            } else {
                if constexpr (sizeof(R) == 4) {
                    if constexpr (PRIMARY) {
                        // camera rays: only the records that reach into this warp's 32-pixel tile (warp-uniform mask);
                        // without masks one slab test of the scene bounds answers the (coherent) misses.  ONE inlined
                        // scan either way: the kernel has to stay inside the 32 KB instruction cache.
                        if (!(B2RT_OPT_MASKS && P.masks) && misses_scene(S, r)) { mask = 0u; n_tally += 0x10000u; }
                        if (mask == 0u) { h.t = 1000000.0f; h.prim = -1; h.a = 0.f; h.b = 0.f; }
                        else scan_small<false, true>(S, s_scan, r, 0.001f, 1000000.0f, h, mask);
                    } else scan_small<false>(S, s_scan, r, 0.001f, 1000000.0f, h);
                }
            }
            shade_now = true;
            }
        }
        if constexpr (RING) {
            // misses are answered on the spot (the path ends: sky term into L[slot]); hits go to the ring
            const bool is_hit = shade_now && h.prim >= 0;
            if (shade_now && !is_hit) add_sky(Q.L + slot, g.thr.x * R(0.1), g.thr.y * R(0.1), g.thr.z * R(0.1));
            const unsigned lane = threadIdx.x & 31u, mh = __ballot_sync(0xffffffffu, is_hit);
            int head = ring_ctl[0], cnt = ring_ctl[1];
            if (is_hit) {
                float *e = ring + ((head + cnt + __popc(mh & ((1u << lane) - 1u))) & (kRingSlots - 1));
                e[0 * kRingSlots] = fmaf(h.t, r.d.x, r.o.x); e[1 * kRingSlots] = fmaf(h.t, r.d.y, r.o.y); e[2 * kRingSlots] = fmaf(h.t, r.d.z, r.o.z);
                e[3 * kRingSlots] = r.d.x; e[4 * kRingSlots] = r.d.y; e[5 * kRingSlots] = r.d.z;
                e[6 * kRingSlots] = g.thr.x; e[7 * kRingSlots] = g.thr.y; e[8 * kRingSlots] = g.thr.z;
                e[9 * kRingSlots] = __int_as_float(h.prim); e[10 * kRingSlots] = h.a; e[11 * kRingSlots] = h.b;
                e[12 * kRingSlots] = __int_as_float(slot); e[13 * kRingSlots] = __uint_as_float((unsigned)g.rng);
            }
            cnt += __popc(mh);
            __syncwarp();
            shade_now = false;
            const int take = (cnt >= 32 || !more) ? min(cnt, 32) : 0;          // full groups; the flush pass takes the rest
            if ((int)lane < take) {
                const float *e = ring + ((head + (int)lane) & (kRingSlots - 1));
                // the hit point stands in for the origin with t = 0: shade_segment computes p = o + t d
                r.o = {e[0 * kRingSlots], e[1 * kRingSlots], e[2 * kRingSlots]};
                r.d = {e[3 * kRingSlots], e[4 * kRingSlots], e[5 * kRingSlots]};
                g.thr = {e[6 * kRingSlots], e[7 * kRingSlots], e[8 * kRingSlots]};
                h.t = 0.f; h.prim = __float_as_int(e[9 * kRingSlots]); h.a = e[10 * kRingSlots]; h.b = e[11 * kRingSlots];
                slot = __float_as_int(e[12 * kRingSlots]); g.rng = (uint64_t)__float_as_uint(e[13 * kRingSlots]);
                shade_now = true;
            }
            __syncwarp();
            if (lane == 0) { ring_ctl[0] = (head + take) & (kRingSlots - 1); ring_ctl[1] = cnt - take; }
            __syncwarp();
        }
        if (shade_now) {
            n_tally += h.prim >= 0 ? 1u : 0u;
            shade_segment<R, Rng, PRIMARY, (WALK || MODE == 0) && !(sizeof(R) == 4 && MODE == 6), SURF, RING || HITQ>(S, Q, ((PLANAR || HITQ) && !S.occl_hint) ? nullptr : s_scan, s_surf,
                                                                r, h, slot, bounce, max_depth, g);
        }
        n_culled += g.culled ? 1u : 0u;
        if (TILED) { t_i += t_step; while (t_i >= n_act) { t_i -= n_act; ++t_s; } }
        int si, ni;
        if constexpr (CHUNKED) warp_append_chunked(Q.counts + bounce + 1, wc, g.alive, g.want_shadow, ni, si);
        else warp_append2(Q.counts + bounce + 1, g.alive, g.want_shadow, ni, si);
        if (B2RT_CHECK) {                        // queue overrun: counted (high word), the append is dropped
            if ((g.want_shadow && si >= Q.capacity) || (g.alive && ni >= Q.capacity)) {
                if (S.check) atomicAdd(S.check, 1ULL << 32);
                g.want_shadow = g.alive = false;
            }
        }
        if (g.want_shadow) {
            // bit 31 of the slot word marks a pre-resolved record (sky term of an escaping path: nothing to trace)
            const unsigned slot_word = (unsigned)slot | (g.light < 0 ? 0x80000000u : 0u);
            st_stream(Q.so + si, Real4<R>::make(g.s_o.x, g.s_o.y, g.s_o.z, pack_int<R>((int64_t)slot_word)));
            if (g.light >= 0)
                st_stream(Q.sd + si, Real4<R>::make(g.s_d.x, g.s_d.y, g.s_d.z, pack_int<R>((int64_t)g.light)));
            st_stream(Q.sc + si, Real4<R>::make(g.s_c.x, g.s_c.y, g.s_c.z, R(0)));
        }
        if (g.alive) {
            st_stream(no + ni, Real4<R>::make(g.new_o.x, g.new_o.y, g.new_o.z, pack_int<R>((int64_t)slot)));
            st_stream(nd + ni, Real4<R>::make(g.new_d.x, g.new_d.y, g.new_d.z, pack_int<R>((int64_t)g.rng)));
            st_stream(nt + ni, Real4<R>::make(g.thr.x, g.thr.y, g.thr.z, pack_int<R>((int64_t)(bounce + 1))));
            if ((WALK || MODE == 0) && Q.keys) Q.keys[ni] = ray_sort_key<R>(g.new_o, g.new_d, S.sort_inv);
        }
    }
    if constexpr (CHUNKED) {
        // the unused remainder of this warp's last chunks: marked dead for the consumers, counted for the statistics
        const int lane = threadIdx.x & 31;
        const int rc = wc->ray_cur, re = wc->ray_end, sc = wc->sh_cur, se = wc->sh_end;
#pragma unroll 1
        for (int k = rc + lane; k < re; k += 32) {
            st_stream(no + k, Real4<R>::make(R(0), R(0), R(0), pack_int<R>(-1)));
            if ((WALK || MODE == 0) && Q.keys) Q.keys[k] = 0xffffffffu;          // dead entries sort to the end
        }
#pragma unroll 1
        for (int k = sc + lane; k < se; k += 32) st_stream(Q.so + k, Real4<R>::make(R(0), R(0), R(0), pack_int<R>(-1)));
        if (lane == 0 && (re - rc) + (se - sc) > 0)
            atomicAdd(Q.dead + bounce + 1, (unsigned long long)(re - rc) | ((unsigned long long)(se - sc) << 32));
    }
    warp_flush(Q.culled, n_culled);
    warp_flush(Q.tally, n_tally >> 16);
    warp_flush(Q.tally + 1, n_tally & 0xffffu);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        unsigned long long ns1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns1));
        atomicAdd(Q.clk, (unsigned long long)(clock64() - clk0));
        atomicAdd(Q.clk + 1, ns1 - ns0);
    }
}

// ------------------------------------------------------------------------------------ split bounce: closest hit only
// Small float32 scenes, bounce >= 1 (B2RT_OPT_SPLIT).  The fused bounce kernel shades with 19 of 32 lanes because a
// quarter of these rays leave the box, and at 29 KB of SASS / 64 registers it has neither instruction cache nor
// registers to spare for a fix inside (the per-warp hit ring was measured: slower).  Split in two instead:
//   scan_hits_kernel   ray queue -> record scan; a miss ends the path here (sky term into L[slot]); a hit is appended
//                      to the HIT QUEUE (hit point, direction, throughput, primitive, rng: 64 B), chunked like every queue
//   shade_kernel<7>    hit queue -> shading, every lane on a hit -> next ray queue + shadow queue
// Both kernels are small (the scan has no shading state, the shade no scan), so they fit the instruction cache with room
// and hold more warps.  Costs 64 B written + read per hit; the wavefront is nowhere near HBM-bound.
// MEASURED AND OFF (profiles/r2_split_bounce_ab.log, r2_ncu_split_bounce.csv): bounces >= 1 take 31.5 instead of 27.2 ms per
// 256 spp.  Bounce 1 under ncu: scan 0.73 ms + shade 0.62 ms against 1.17 ms fused; the pair moves 7.1 GB of DRAM traffic
// where the fused kernel moves 4.0 GB and both halves sit at 62-66 % of the DRAM peak with long-scoreboard 4.9 / 7.0 per
// issue — the instructions saved (shading at 26 instead of 19 lanes) are paid for in bandwidth.  Fusion stays.
#ifndef B2RT_OPT_SPLIT
#define B2RT_OPT_SPLIT 0
#endif
#ifndef B2RT_SCAN_MIN_BLOCKS
#define B2RT_SCAN_MIN_BLOCKS 5          // measured 4 / 5 / 6: 17.0 / 17.1 / 18.4 ms per 256 spp
#endif
template <typename R>
__global__ void __launch_bounds__(256, sizeof(R) == 4 ? B2RT_SCAN_MIN_BLOCKS : 1)
scan_hits_kernel(SceneDev S, PathQueues<R> Q, int in_buf, int bounce) {
    extern __shared__ float4 s_dyn[];
    float4 *s_scan = s_dyn;
    stage_scan(S, s_scan);
    float4 *s_ray = s_scan + 4 * (S.n_scan + S.n_box);                       // [2 stages][3 streams][256 threads]
    const real4<R> *__restrict__ ro = Q.ro[in_buf], *__restrict__ rd = Q.rd[in_buf], *__restrict__ th = Q.th[in_buf];
    const int n = ray_count(Q, bounce), n_round = (n + 31) & ~31;
    __shared__ WarpCursor s_wc[8];
    WarpCursor *wc = s_wc + (threadIdx.x >> 5);
    if ((threadIdx.x & 31) == 0) { wc->ray_cur = wc->ray_end = wc->sh_cur = wc->sh_end = 0; }
    __syncwarp();
    unsigned *tail = reinterpret_cast<unsigned *>(Q.hit_tail + bounce);
    int stage = 0;
    if constexpr (sizeof(R) == 4) {
        const int i0 = blockIdx.x * blockDim.x + threadIdx.x;
        if (i0 < n) { cp_async16(s_ray + threadIdx.x, ro + i0); cp_async16(s_ray + 256 + threadIdx.x, rd + i0); cp_async16(s_ray + 512 + threadIdx.x, th + i0); }
        cp_async_commit();
    }
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += gridDim.x * blockDim.x) {
        const bool valid = i < n;
        bool is_hit = false;
        Ray<R> r; Hit<R> h; real4<R> a, b, c;
        if constexpr (sizeof(R) == 4) {
            cp_async_wait_all();
            const float4 *cur = s_ray + stage * 768 + threadIdx.x;
            if (valid) { a = cur[0]; b = cur[256]; c = cur[512]; }
            stage ^= 1;
            const int nx = i + gridDim.x * blockDim.x;
            if (nx < n) { float4 *nxt = s_ray + stage * 768 + threadIdx.x; cp_async16(nxt, ro + nx); cp_async16(nxt + 256, rd + nx); cp_async16(nxt + 512, th + nx); }
            cp_async_commit();
        } else if (valid) { a = ld_stream(ro + i); b = ld_stream(rd + i); c = ld_stream(th + i); }
        int slot = -1;
        if (valid) {
            slot = (int)unpack_u<R>(a.w);
            if (slot >= 0) {                                                 // not a dead chunk remainder
                r.o = xyz<R>(a); r.d = xyz<R>(b);
                if constexpr (sizeof(R) == 4) scan_small<false>(S, s_scan, r, 0.001f, 1000000.0f, h);
                else { h.prim = -1; h.t = R(0); h.a = h.b = R(0); }
                is_hit = h.prim >= 0;
                if (!is_hit) add_sky(Q.L + slot, c.x * R(0.1), c.y * R(0.1), c.z * R(0.1));      // :234-239: the path ends here
            }
        }
        const unsigned mh = __ballot_sync(0xffffffffu, is_hit);
        if (mh) {
            int cur_ = wc->ray_cur, end_ = wc->ray_end;
            const int k = chunk_take(tail, cur_, end_, mh, threadIdx.x & 31u);
            if ((threadIdx.x & 31) == 0) { wc->ray_cur = cur_; wc->ray_end = end_; }
            __syncwarp();
            if (is_hit) {
                st_stream(Q.ha + k, Real4<R>::make(r.o.x + r.d.x * h.t, r.o.y + r.d.y * h.t, r.o.z + r.d.z * h.t, a.w));
                st_stream(Q.hb + k, b);
                st_stream(Q.hc + k, Real4<R>::make(c.x, c.y, c.z, pack_int<R>((int64_t)h.prim)));
                st_stream(Q.hd + k, Real4<R>::make(h.a, h.b, R(0), R(0)));
            }
        }
    }
    {   // dead remainder of this warp's last chunk
        const int lane = threadIdx.x & 31, rc = wc->ray_cur, re = wc->ray_end;
#pragma unroll 1
        for (int k = rc + lane; k < re; k += 32) st_stream(Q.ha + k, Real4<R>::make(R(0), R(0), R(0), pack_int<R>(-1)));
    }
}

// ------------------------------------------------------------------------------------ shadow
template <typename R>
__global__ void __launch_bounds__(256)
shadow_kernel(SceneDev S, PathQueues<R> Q, int bounce) {
    extern __shared__ float4 s_top[];
    const bool planar = S.scan_incoherent && sizeof(R) == 4 && S.n_scan > 0;
    if (!S.scan_incoherent) stage_top(S, s_top);
    else if (planar) stage_scan(S, s_top);
    int n = shadow_count(Q, bounce);
    int n_round = (n + 31) & ~31;
    unsigned n_lit = 0, n_sky = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += gridDim.x * blockDim.x) {
        bool lit = false, sky = false;
        if (i < n) {
            real4<R> a = ld_stream(Q.so + i);
            const unsigned slot_word = (unsigned)unpack_u<R>(a.w);
            const bool dead = slot_word == 0xffffffffu;      // unused remainder of a producer warp's last chunk
            sky = !dead && (slot_word >> 31) != 0u;          // pre-resolved: the sky term of a path that escaped
            if (!sky && !dead) {
                real4<R> b = ld_stream(Q.sd + i);
                Ray<R> r; r.o = xyz<R>(a); r.d = xyz<R>(b);
                Hit<R> h;
                bool occluded;                                                      // :275-277 t_max = 1e6
                if (!S.scan_incoherent) occluded = traverse<R, false, true>(S, s_top, r, R(0.001), R(1000000.0), h);
                else if constexpr (sizeof(R) == 4) {
                    occluded = planar ? scan_small<true>(S, s_top, r, 0.001f, 1000000.0f, h)
                                      : scan_all<R, false, true>(S, r, R(0.001), R(1000000.0), h);
                } else occluded = scan_all<R, false, true>(S, r, R(0.001), R(1000000.0), h);
                lit = !occluded;
            }
            if (lit || sky) {
                const int slot = (int)(slot_word & 0x7fffffffu);
                real4<R> c = ld_stream(Q.sc + i);
                add_stream(Q.L + slot, c.x, c.y, c.z);
            }
        }
        n_lit += lit ? 1u : 0u;
        n_sky += sky ? 1u : 0u;
    }
    warp_flush(Q.unshadowed, n_lit);
    warp_flush(Q.tally + 4, n_sky);
}

// ------------------------------------------------------------------------------------ accumulate / resolve
// per-pixel radiance accumulation: accum[pix] += sum_s L[s][pix], samples in index order (:43-45)
template <typename R>
__global__ void __launch_bounds__(256)
accumulate_kernel(int npix, int spp_wave, const real4<R> *__restrict__ L, real4<R> *__restrict__ accum,
                  real4<R> *__restrict__ accum_sq, const unsigned *__restrict__ masks) {
    for (int pix = blockIdx.x * blockDim.x + threadIdx.x; pix < npix; pix += gridDim.x * blockDim.x) {
        // wave-local partial sum first: the float32 running sum then sees one add per wave, not per sample
        R sx = R(0), sy = R(0), sz = R(0), qx = R(0), qy = R(0), qz = R(0);
        // tiles no camera ray of which can hit anything were never traced (primary_mask_kernel): every sample is the
        // sky value bounce 0 would have stored, added here in the same order
        const bool sky_only = masks != nullptr && __ldg(masks + (pix >> 5)) == 0u;
        for (int s = 0; s < spp_wave; ++s) {
            real4<R> l = sky_only ? Real4<R>::make(R(0.1), R(0.1), R(0.1), R(0)) : L[(size_t)s * npix + pix];
            sx += l.x; sy += l.y; sz += l.z;
            qx += l.x * l.x; qy += l.y * l.y; qz += l.z * l.z;
        }
        real4<R> a = accum[pix];
        a.x += sx; a.y += sy; a.z += sz;
        accum[pix] = a;
        if (accum_sq) {
            real4<R> q = accum_sq[pix];
            q.x += qx; q.y += qy; q.z += qz;
            accum_sq[pix] = q;
        }
    }
}

static __global__ void iota_kernel(int n, int *out) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[i] = i;
}

static __global__ void path_counters_kernel(const unsigned long long *counts, const unsigned long long *unshadowed,
                                     const unsigned long long *culled, const unsigned long long *tally,
                                     const unsigned long long *clk, const unsigned long long *dead, int max_depth,
                                     long long paths, int spp_wave, unsigned long long launches, unsigned long long *out) {
    unsigned long long rays = 0, shadows = 0;
    for (int b = 0; b < max_depth; ++b) {
        rays += (counts[b] & 0xffffffffULL) - (dead[b] & 0xffffffffULL);      // queue tails minus dead chunk remainders
        shadows += (counts[b + 1] >> 32) - (dead[b + 1] >> 32);
    }
    // [2] counts every shadow ray that was answered: queued ones plus those the occluder cache resolved
    shadows -= tally[4];                                     // pre-resolved sky records are not shadow rays
    out[0] += (unsigned long long)paths; out[1] += rays; out[2] += shadows + *culled; out[3] += *unshadowed;
    out[4] += launches; out[5] += *culled; out[6] += tally[0]; out[7] += tally[1];
    out[8] += tally[2]; out[9] += tally[3];
    // [10] canonical flops of this wave's masked camera-ray record tests, [6] also counts rays of empty-mask tiles
    out[10] += tally[5] * (unsigned long long)spp_wave; out[6] += tally[6] * (unsigned long long)spp_wave;
    out[11] += clk[0]; out[12] += clk[1];
}

// mean -> ACES (cuda_tonemap :74-81) -> min(255, max(0, int(c*255))) (:56-58) -> V flip (:807)
template <typename R>
__global__ void __launch_bounds__(256)
resolve_kernel(const real4<R> *__restrict__ accum, int W, int H, R spp, int tonemap, uint8_t *__restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= W * H) return;
    int x = i % W, y = i / W;
    real4<R> a = accum[i];
    R c[3] = {a.x / spp, a.y / spp, a.z / spp};
    uint8_t *o = out + 3 * ((size_t)(H - 1 - y) * W + x);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        R v = c[k];
        if (tonemap) v = (v * (R(2.51) * v + R(0.03))) / (v * (R(2.43) * v + R(0.59)) + R(0.14));
        long long q = (long long)(v * R(255));
        o[k] = (uint8_t)(q < 0 ? 0 : (q > 255 ? 255 : q));
    }
}

// ------------------------------------------------------------------------------------ multi-GPU: fused reduce + resolve
// The float accumulation buffers of all ranks live in peer-accessible (symmetric) memory.  Rank g sums rows
// [row0, row1) of EVERY rank's buffer over NVLink — in rank order, so the result does not depend on timing — applies
// the resolve epilogue (mean, ACES, quantise, V flip) and stores the bytes straight into the root's image buffer:
// the reduce-scatter, the resolve and the gather of SURVEY 8e in one kernel, no float image ever crosses the link
// twice.  Four pixels per thread: 4 x 16 B loads per peer, 12 B of packed output.
constexpr int kMaxPeers = 16;
struct PeerAccum { const float4 *p[kMaxPeers]; };

__device__ __forceinline__ unsigned quant_aces(float v, int tonemap) {
    if (tonemap) v = (v * (2.51f * v + 0.03f)) / (v * (2.43f * v + 0.59f) + 0.14f);
    long long q = (long long)(v * 255.0f);
    return (unsigned)(q < 0 ? 0 : (q > 255 ? 255 : q));
}

static __global__ void __launch_bounds__(256)
reduce_resolve_kernel(PeerAccum peers, int n_peers, int W, int H, int row0, int row1, float spp, int tonemap,
                      uint8_t *__restrict__ root_u8, float4 *__restrict__ root_sum) {
    const int groups_per_row = (W + 3) / 4, n = (row1 - row0) * groups_per_row;
    for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < n; g += gridDim.x * blockDim.x) {
        const int y = row0 + g / groups_per_row, x0 = (g - (g / groups_per_row) * groups_per_row) * 4;
        const int cnt = min(4, W - x0);
        float4 acc[4];
        for (int k = 0; k < 4; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int p = 0; p < n_peers; ++p) {
            const float4 *src = peers.p[p] + (size_t)y * W + x0;
            for (int k = 0; k < cnt; ++k) {
                const float4 v = __ldcs(src + k);
                acc[k].x += v.x; acc[k].y += v.y; acc[k].z += v.z;
            }
        }
        uint8_t b[12];
        for (int k = 0; k < cnt; ++k) {
            if (root_sum) root_sum[(size_t)y * W + x0 + k] = acc[k];
            // the same operations as resolve_kernel<float>: a / spp first
            b[3 * k] = (uint8_t)quant_aces(acc[k].x / spp, tonemap);
            b[3 * k + 1] = (uint8_t)quant_aces(acc[k].y / spp, tonemap);
            b[3 * k + 2] = (uint8_t)quant_aces(acc[k].z / spp, tonemap);
        }
        uint8_t *o = root_u8 + 3 * ((size_t)(H - 1 - y) * W + x0);
        if (cnt == 4 && ((size_t)o & 3) == 0) {
            unsigned w0 = b[0] | (b[1] << 8) | (b[2] << 16) | ((unsigned)b[3] << 24);
            unsigned w1 = b[4] | (b[5] << 8) | (b[6] << 16) | ((unsigned)b[7] << 24);
            unsigned w2 = b[8] | (b[9] << 8) | (b[10] << 16) | ((unsigned)b[11] << 24);
            unsigned *ow = reinterpret_cast<unsigned *>(o);
            ow[0] = w0; ow[1] = w1; ow[2] = w2;
        } else {
            for (int k = 0; k < 3 * cnt; ++k) o[k] = b[k];
        }
    }
}

// RGB8 -> RGBX8 (one 32-bit load per texel in the kernels), four texels per thread
static __global__ void __launch_bounds__(256)
expand_rgb8_kernel(const uint8_t *__restrict__ rgb, long long n_texels, uint32_t *__restrict__ rgbx) {
    const long long n4 = n_texels / 4;
    const uint32_t *in = reinterpret_cast<const uint32_t *>(rgb);
    for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < n4; g += (long long)gridDim.x * blockDim.x) {
        const uint32_t a = __ldcs(in + 3 * g), b = __ldcs(in + 3 * g + 1), c = __ldcs(in + 3 * g + 2);
        uint4 o;
        o.x = (a & 0x00ffffffu) | 0xff000000u;
        o.y = ((a >> 24) | (b << 8)) & 0x00ffffffu | 0xff000000u;
        o.z = ((b >> 16) | (c << 16)) & 0x00ffffffu | 0xff000000u;
        o.w = (c >> 8) | 0xff000000u;
        reinterpret_cast<uint4 *>(rgbx)[g] = o;
    }
    if (blockIdx.x == 0 && threadIdx.x < (int)(n_texels - 4 * n4)) {
        const long long t = 4 * n4 + threadIdx.x;
        rgbx[t] = rgb[3 * t] | (rgb[3 * t + 1] << 8) | (rgb[3 * t + 2] << 16) | 0xff000000u;
    }
}

}  // namespace b2rt
