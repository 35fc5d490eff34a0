// rt_api.cuh — implementation of the launchers declared in rt_api.h (included by rt_f32.cu / rt_f64.cu).
#pragma once
#include <stdlib.h>
#include <string.h>

#include <type_traits>

#include <cub/device/device_radix_sort.cuh>

#include "rt_api.h"
#include "rt_path.cuh"
#include "rt_whitted.cuh"

namespace b2rt {


inline size_t smem_top_bytes(const SceneDev &S) { return (size_t)S.n_top * 64; }

// Dynamic shared memory above the 48 KB default needs an explicit opt-in per kernel (top levels: up to kTopMax * 64 B,
// MODE 6 adds the scan and surface records on top).  Returns an error when the request exceeds the device limit.
inline cudaError_t opt_in_smem(const void *kernel, size_t smem) {
    if (smem <= 40 * 1024) return cudaSuccess;              // (the 48 KB default covers static + dynamic: leave room for the static part)
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}

// persistent grid = resident CTAs per SM x SM count (a multiple of the 148 SMs); a kernel that cannot be resident
// at all (occupancy 0: too much shared memory or registers for the block size) is an error, not a grid of 1 per SM
inline cudaError_t persistent_grid(const void *kernel, int block, size_t smem, int *grid) {
    int dev = 0, sms = 0, per_sm = 0;
    cudaError_t e;
    if ((e = opt_in_smem(kernel, smem))) return e;
    if ((e = cudaGetDevice(&dev))) return e;
    if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev))) return e;
    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, block, smem))) return e;
    if (per_sm < 1) return cudaErrorLaunchOutOfResources;
    *grid = sms * per_sm;
    return cudaSuccess;
}

template <typename R>
cudaError_t Api<R>::primary_hits(const b2rt_scene *s, const double *cam, int W, int H, double du, double dv,
                                 double t_min, double t_max, int use_bvh, int *ids, double *tt, cudaStream_t st) {
    SceneDev S = make_scene_dev(s);
    S.check = check_counter();
    Cam<R> c = make_cam<R>(cam);
    int n = W * H, T = 128;
    if (cudaError_t e = opt_in_smem((const void *)primary_hits_kernel<R, true>, smem_top_bytes(S))) return e;
    if (cudaError_t e = opt_in_smem((const void *)primary_hits_kernel<R, false>, smem_top_bytes(S))) return e;
    if (S.semantics == B2RT_SEM_CPU)
        primary_hits_kernel<R, true><<<(n + T - 1) / T, T, smem_top_bytes(S), st>>>(S, c, W, H, R(du), R(dv), R(t_min), R(t_max), use_bvh, ids, tt);
    else
        primary_hits_kernel<R, false><<<(n + T - 1) / T, T, smem_top_bytes(S), st>>>(S, c, W, H, R(du), R(dv), R(t_min), R(t_max), use_bvh, ids, tt);
    return cudaGetLastError();
}

template <typename R>
cudaError_t Api<R>::trace_rays(const b2rt_scene *s, int n, const double *o, const double *d, double t_min, double t_max,
                               int any_hit, int use_bvh, int *ids, double *rec, cudaStream_t st) {
    SceneDev S = make_scene_dev(s);
    S.check = check_counter();
    int T = 128;
    if (n <= 0) return cudaSuccess;
    const size_t sm = use_bvh == 2 ? smem_scan_bytes(S) + 64 : smem_top_bytes(S);
    if (cudaError_t e = opt_in_smem((const void *)trace_rays_kernel<R, true>, sm)) return e;
    if (cudaError_t e = opt_in_smem((const void *)trace_rays_kernel<R, false>, sm)) return e;
    if (S.semantics == B2RT_SEM_CPU)
        trace_rays_kernel<R, true><<<(n + T - 1) / T, T, sm, st>>>(S, n, o, d, R(t_min), R(t_max), any_hit, use_bvh, ids, rec);
    else
        trace_rays_kernel<R, false><<<(n + T - 1) / T, T, sm, st>>>(S, n, o, d, R(t_min), R(t_max), any_hit, use_bvh, ids, rec);
    return cudaGetLastError();
}

template <typename R>
cudaError_t Api<R>::whitted_cpu(const b2rt_scene *s, const double *cam, int W, int H, const double *jitter,
                                int max_depth, const double *ambient, const double *light_color, double *rgb,
                                cudaStream_t st) {
    SceneDev S = make_scene_dev(s);
    S.check = check_counter();
    Cam<R> c = make_cam<R>(cam);
    V3<R> amb = {R(ambient[0]), R(ambient[1]), R(ambient[2])};
    V3<R> lc = {R(light_color[0]), R(light_color[1]), R(light_color[2])};
    int n = W * H, T = 128;
    if (cudaError_t e = opt_in_smem((const void *)whitted_cpu_kernel<R>, smem_top_bytes(S))) return e;
    whitted_cpu_kernel<R><<<(n + T - 1) / T, T, smem_top_bytes(S), st>>>(S, c, W, H, jitter, max_depth, amb, lc, rgb);
    return cudaGetLastError();
}

template <typename R>
cudaError_t Api<R>::whitted_texture(const b2rt_scene *s, const double *cam, int W, int H, int spp, int max_depth,
                                    double *rgb, uint8_t *u8, cudaStream_t st) {
    SceneDev S = make_scene_dev(s);
    S.check = check_counter();
    Cam<R> c = make_cam<R>(cam);
    int n = W * H, T = 128;
    if (sizeof(R) == 4 && S.scan_incoherent && S.n_scan > 0 && S.surf)
        whitted_texture_kernel<R, true><<<(n + T - 1) / T, T, smem_scan_bytes(S) + smem_surf_bytes(S), st>>>(S, c, W, H, spp, max_depth, rgb, u8);
    else if (cudaError_t e = opt_in_smem((const void *)whitted_texture_kernel<R, false>, smem_top_bytes(S))) return e;
    else
        whitted_texture_kernel<R, false><<<(n + T - 1) / T, T, smem_top_bytes(S), st>>>(S, c, W, H, spp, max_depth, rgb, u8);
    return cudaGetLastError();
}

// ---- wavefront path tracer --------------------------------------------------------------------
inline size_t align256(size_t v) { return (v + 255) & ~size_t(255); }
// the ray re-ordering sorts key bits [kSortBeginBit, 30).  Measured on 1 M triangles: dropping the six finest origin
// bits (three radix passes instead of four) saves 1.35 ms of sorting and costs 1.8 ms of walking, so all 30 are sorted
#ifndef B2RT_SORT_BEGIN_BIT
#define B2RT_SORT_BEGIN_BIT 0
#endif
constexpr int kSortBeginBit = B2RT_SORT_BEGIN_BIT;

constexpr size_t kQueueSlack = (size_t)256 * 64 * kQueueChunk;      // any device with <= 256 SMs

template <typename R> struct PathLayout {
    size_t stream_bytes, counts_off, sort_off, int_bytes, cub_bytes, mask_off, total;
    static PathLayout make(int W, int H, int spp_per_wave, int max_depth) {
        PathLayout L;
        size_t n = (size_t)W * H * spp_per_wave;
        // + the dead remainders of every warp's last chunk (chunked append): <= 148 SMs x 64 resident warps x chunk
        L.stream_bytes = align256((n + (size_t)kQueueSlack) * sizeof(real4<R>));
        // 6 ray + 1 hit + 3 shadow + 1 radiance streams (+ 3 hit-queue streams in a B2RT_OPT_SPLIT build)
        L.counts_off = (B2RT_OPT_SPLIT ? 14 : 11) * L.stream_bytes;
        // per-bounce queue tails, unshadowed, culled + one ray-fetch counter per bounce (extend_walk_kernel)
        L.sort_off = L.counts_off + align256(sizeof(unsigned long long) * (3 * (size_t)max_depth + 18));
        // ray re-ordering (LBVH scenes): keys, sorted keys, iota, permutation + CUB scratch
        const size_t nq = n + (size_t)kQueueSlack;           // queue tails include the dead chunk remainders
        L.int_bytes = align256(nq * sizeof(int));
        L.cub_bytes = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, L.cub_bytes, (const unsigned *)nullptr, (unsigned *)nullptr,
                                        (const int *)nullptr, (int *)nullptr, (int)nq, 0, 30);
        L.cub_bytes = align256(L.cub_bytes);
        // camera-ray candidate masks: one word per 32-pixel tile (small scenes)
        L.mask_off = L.sort_off + 4 * L.int_bytes + L.cub_bytes;
        L.total = L.mask_off + 2 * align256(((size_t)W * H / 32 + 1) * sizeof(unsigned));   // masks + non-empty tile list
        return L;
    }
};

template <typename R> size_t Api<R>::path_workspace_bytes(int W, int H, int spp_per_wave, int max_depth) {
    return PathLayout<R>::make(W, H, spp_per_wave, max_depth).total;
}

template <typename R, int NODES> inline const void *walk_kernel_ptr(bool count) {
    return count ? (const void *)extend_walk_kernel<R, true, NODES> : (const void *)extend_walk_kernel<R, false, NODES>;
}
template <typename R, int NODES>
inline void launch_walk(bool count, int grid, int block, size_t smem, cudaStream_t st, const SceneDev &S, const PathQueues<R> &Q,
                        int buf, int b, unsigned *cursor) {
    if (count)
        extend_walk_kernel<R, true, NODES><<<grid, block, smem, st>>>(S, Q.ro[buf], Q.rd[buf], Q.hit, Q.counts + b, Q.perm, cursor,
                                                                      Q.tally + 2);
    else
        extend_walk_kernel<R, false, NODES><<<grid, block, smem, st>>>(S, Q.ro[buf], Q.rd[buf], Q.hit, Q.counts + b, Q.perm, cursor,
                                                                       nullptr);
}

template <typename R, typename Rng>
cudaError_t render_path_impl(const b2rt_scene *s, const double *cam, const PathArgs &a, cudaStream_t st) {
    SceneDev S = make_scene_dev(s);
    S.check = check_counter();
    Cam<R> c = make_cam<R>(cam);
    const int W = a.width, H = a.height, npix = W * H;
    int wave = a.spp_per_wave < 1 ? 1 : a.spp_per_wave;
    if (wave > a.spp_local) wave = a.spp_local;
    if (a.spp_local <= 0) return cudaSuccess;
    PathLayout<R> L = PathLayout<R>::make(W, H, wave, a.max_depth);
    if (a.workspace_bytes < L.total) return cudaErrorInvalidValue;
    char *base = (char *)a.workspace;
    PathQueues<R> Q;
    auto stream_at = [&](int k) { return (real4<R> *)(base + (size_t)k * L.stream_bytes); };
    Q.ro[0] = stream_at(0); Q.rd[0] = stream_at(1); Q.th[0] = stream_at(2);
    Q.ro[1] = stream_at(3); Q.rd[1] = stream_at(4); Q.th[1] = stream_at(5);
    Q.hit = stream_at(6); Q.so = stream_at(7); Q.sd = stream_at(8); Q.sc = stream_at(9); Q.L = stream_at(10);
    Q.ha = Q.hb = Q.hc = Q.hd = nullptr;
    if (B2RT_OPT_SPLIT) { Q.ha = stream_at(11); Q.hb = stream_at(12); Q.hc = stream_at(13); Q.hd = Q.hit; }      // hit queue (split bounce)
    unsigned long long *counts = (unsigned long long *)(base + L.counts_off);
    Q.counts = counts;
    Q.unshadowed = counts + a.max_depth + 1;
    Q.culled = counts + a.max_depth + 2;
    // cleared per wave: queue tails, unshadowed, culled, fetch cursors, tally[0..4]; tally[5..7] live for the whole call
    size_t counts_bytes_wave = sizeof(unsigned long long) * (2 * (size_t)a.max_depth + 4 + 5);
    Q.clk = counts + 2 * a.max_depth + 12;                                  // [2] SM cycles / ns of CTA 0, cleared per wave
    Q.dead = counts + 2 * a.max_depth + 14;                                 // [max_depth + 1] dead queue entries, cleared per wave
    unsigned long long *tile_info = counts + 3 * a.max_depth + 15;          // [2] non-empty tiles, division constants (per call)
    Q.tally = counts + 2 * a.max_depth + 4;                                 // [8] bounds-culled, hits, walk box / leaf steps, sky records
    unsigned long long *fetch = counts + a.max_depth + 3;                   // [max_depth] dynamic-fetch cursors
    Q.hit_tail = fetch;                                                     // small scenes: hit-queue tails (the walk kernel is unused there)
    const bool sort_rays = S.sort_inv > 0.f && !(a.flags & 2);
    unsigned *keys = (unsigned *)(base + L.sort_off), *keys_sorted = (unsigned *)(base + L.sort_off + L.int_bytes);
    int *iota = (int *)(base + L.sort_off + 2 * L.int_bytes), *perm = (int *)(base + L.sort_off + 3 * L.int_bytes);
    void *cub_tmp = base + L.sort_off + 4 * L.int_bytes;
    Q.keys = sort_rays ? keys : nullptr;
    Q.capacity = (int)(L.stream_bytes / sizeof(real4<R>));
    Q.perm = nullptr;

    const size_t smem = smem_top_bytes(S);
    const size_t smem_scan_only = smem_scan_bytes(S);
    const size_t smem_shadow = S.scan_incoherent ? smem_scan_only : smem;
    const size_t smem_scan = smem_scan_only + smem_surf_bytes(S);           // bounce kernels also stage the surface records
    const int T = 256;
    cudaError_t e;
    // persistent grids: resident CTAs per SM x SM count (a multiple of the 148 SMs)
    int g_extend = 0, g_shade = 0, g_shadow = 0, g_simple = 0, g_fuse_bvh = 0, g_fuse_scan = 0, g_walk = 0;
    int g_primary = 0, g_primary_small = 0, g_primary_scan = 0;
    const bool fused = !(a.flags & 1);
    const size_t smem_bvh = smem;                                           // MODE 1 / 4: generic streams only
    const size_t smem_bvh_small = smem + smem_scan;                         // MODE 6: + scan and surface records
    const bool planar = sizeof(R) == 4 && S.n_scan > 0 && S.surf != nullptr;
    // MODE 2 (generic scan) stages the scan records only when it uses them: for the occluder hints of float32 scenes
    const size_t smem_generic = (sizeof(R) == 4 && S.n_scan > 0 && S.scan_incoherent && S.occl_hint) ? smem_scan_only : 0;
    if ((e = persistent_grid((const void *)extend_kernel<R>, T, smem, &g_extend))) return e;
    if ((e = persistent_grid((const void *)shade_kernel<R, Rng, 0>, T, 0, &g_shade))) return e;
    if ((e = persistent_grid((const void *)shade_kernel<R, Rng, 1>, T, smem_bvh, &g_fuse_bvh))) return e;
    // MODE 3 double-buffers its ray records in shared memory (cp.async)
    const size_t smem_mode3 = smem_scan + ((B2RT_OPT_ASYNC && sizeof(R) == 4) ? 2 * (size_t)kAsyncStageBytes + (size_t)kRingBytes : 0);
    // split bounce (small float32 scenes, bounce >= 1): closest-hit scan -> hit queue -> shading with every lane on a hit
    const bool split = B2RT_OPT_SPLIT && planar && sizeof(R) == 4 && !(a.flags & 256);
    const size_t smem_scan_k = smem_scan_only + 2 * (size_t)(3 * 256 * 16);
    const size_t smem_mode7 = smem_scan + 2 * (size_t)kAsyncStageBytes;
    int g_scan_hits = 0, g_mode7 = 0;
    if (split) {
        if ((e = persistent_grid((const void *)scan_hits_kernel<R>, T, smem_scan_k, &g_scan_hits))) return e;
        if ((e = persistent_grid((const void *)shade_kernel<R, Rng, 7>, T, smem_mode7, &g_mode7))) return e;
    }
    if (planar) e = persistent_grid((const void *)shade_kernel<R, Rng, 3>, T, smem_mode3, &g_fuse_scan);
    else e = persistent_grid((const void *)shade_kernel<R, Rng, 2>, T, smem_generic, &g_fuse_scan);
    if (e) return e;
    if ((e = persistent_grid((const void *)shadow_kernel<R>, T, smem_shadow, &g_shadow))) return e;
    // large scenes: incoherent bounces run the persistent walk kernel + the wavefront shade stage
    const bool walk_kernel = fused && sizeof(R) == 4 && !S.scan_incoherent && !(a.flags & 8);
    const bool count_tests = (a.flags & 64) != 0;
    // node format of the walk kernel: float32 scenes that carry quantised 32 B nodes (b2rt_lbvh_quantize) or 4-wide nodes
    // (b2rt_lbvh_widen) walk those unless B2RT_PATH_BINARY_WALK asks for the plain 64 B nodes
    int walk_nodes = 0;
    if (walk_kernel && sizeof(R) == 4 && !(a.flags & 512)) walk_nodes = S.quant ? 2 : (S.wide ? 1 : 0);
    const size_t smem_walk = walk_nodes ? 0 : walk_smem_bytes(S);
    const void *k_walk = walk_kernel_ptr<R, 0>(count_tests);
    if constexpr (sizeof(R) == 4) {
        if (walk_nodes == 1) k_walk = walk_kernel_ptr<R, 1>(count_tests);
        if (walk_nodes == 2) k_walk = walk_kernel_ptr<R, 2>(count_tests);
    }
    if ((e = persistent_grid(k_walk, T, smem_walk, &g_walk))) return e;
    if ((e = persistent_grid((const void *)accumulate_kernel<R>, T, 0, &g_simple))) return e;

    // Measured and OFF by default (B2RT_L2_PERSIST=1 turns it on): pinning the hierarchy into the persisting part of the L2
    // (access-policy window over the node array) while the waves stream gigabytes of queue records past it.  On the
    // 1 M-triangle scene the step went from 89.3 to 109.2 ms (walk kernel 62.5 -> 75.9 ms): the set-aside takes more from
    // the triangle records and the queues than it gives the nodes (profiles/r2_c4_l2_persist_ab.log).
    bool l2_window = false;
    if (!S.scan_incoherent && S.n_prims > 4096 && S.nodes) {
        const char *ev = getenv("B2RT_L2_PERSIST");
        if (ev && atoi(ev) != 0) {
            int dev = 0, max_persist = 0, max_window = 0;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev);
            cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, dev);
            const size_t node_bytes = (size_t)(S.n_prims > 1 ? S.n_prims - 1 : 1) * 64;
            if (max_persist > 0 && max_window > 0) {
                const size_t win = node_bytes < (size_t)max_window ? node_bytes : (size_t)max_window;
                cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)max_persist);
                cudaStreamAttrValue av;
                memset(&av, 0, sizeof av);
                av.accessPolicyWindow.base_ptr = const_cast<float4 *>(S.nodes);
                av.accessPolicyWindow.num_bytes = win;
                av.accessPolicyWindow.hitRatio = (float)((double)max_persist / (double)win < 1.0 ? (double)max_persist / (double)win : 1.0);
                av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
                av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
                l2_window = cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &av) == cudaSuccess;
                if (!l2_window) cudaGetLastError();
            }
        }
    }
    struct L2Reset {                                            // the stream belongs to the caller: give it back unchanged
        cudaStream_t st; bool on;
        ~L2Reset() {
            if (!on) return;
            cudaStreamAttrValue av;
            memset(&av, 0, sizeof av);
            cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &av);     // later launches: no window
        }
    } l2_reset{st, l2_window};
    if (std::is_same<Rng, RefRng>::value) {
        if (!a.pixel_rng) return cudaErrorInvalidValue;
        init_pixel_rng_kernel<<<(npix + 255) / 256, 256, 0, st>>>(W, H, (long long)a.seed, a.sample_offset, a.pixel_rng);
    }
    // camera rays are generated inside the first bounce kernel when the RNG is counter-based
    // (large scenes, flag 32: the primary rays go through raygen + the persistent walk kernel as well — measured
    // 172.4 vs 169.3 ms per step on the 1 M-triangle scene, so the fused first bounce stays the default)
    const bool fuse_primary = fused && std::is_same<Rng, PcgRng>::value && !(walk_kernel && (a.flags & 32));
    // small scenes: primary rays use the scan/box records too unless B2RT_PATH_PRIMARY_WALK asks for the LBVH walk
    const bool primary_scan = planar && S.scan_incoherent && !(a.flags & 4);
    if ((e = persistent_grid((const void *)shade_kernel<R, PcgRng, 4>, T, smem_bvh, &g_primary))) return e;
    if (planar) {
        if ((e = persistent_grid((const void *)shade_kernel<R, PcgRng, 6>, T, smem_bvh_small, &g_primary_small))) return e;
        if ((e = persistent_grid((const void *)shade_kernel<R, PcgRng, 5>, T, smem_scan, &g_primary_scan))) return e;
    }
    // small float32 scenes: per-tile candidate masks for the camera rays (once per call; the masked-test statistics
    // land in tally[5..6], which the per-wave memset below leaves alone)
    unsigned *masks = nullptr;
    int *tile_list = nullptr;
    if ((e = cudaMemsetAsync(Q.tally + 5, 0, 3 * sizeof(unsigned long long), st))) return e;
    if ((e = cudaMemsetAsync(tile_info, 0, 2 * sizeof(unsigned long long), st))) return e;
    if constexpr (sizeof(R) == 4) {
        if (B2RT_OPT_MASKS && fuse_primary && primary_scan && W % 32 == 0 && !(a.flags & 128) &&
            S.n_box + S.n_loose + S.n_sphere <= 32) {
            masks = (unsigned *)(base + L.mask_off);
            tile_list = (int *)(base + L.mask_off + align256(((size_t)W * H / 32 + 1) * sizeof(unsigned)));
            primary_mask_kernel<<<(npix / 32 + 127) / 128, 128, 0, st>>>(S, make_cam<float>(cam), W, H, masks, Q.tally + 5);
            tile_compact_kernel<<<1, 1024, 0, st>>>(masks, npix / 32, tile_list, tile_info);
        }
    }
    if (sort_rays) iota_kernel<<<g_simple, T, 0, st>>>((int)((size_t)npix * wave + kQueueSlack), iota);
    for (int done = 0; done < a.spp_local; done += wave) {
        int k = a.spp_local - done < wave ? a.spp_local - done : wave;
        unsigned long long launches = 0;
        Q.perm = nullptr;
        PrimaryArgs<R> PA;
        PA.cam = c; PA.W = W; PA.H = H; PA.spp_wave = k; PA.first_sample = a.sample_offset + done; PA.seed = a.seed;
        PA.by_npix = FastDiv::make((unsigned)npix); PA.by_w = FastDiv::make((unsigned)W);
        PA.tiles_x = (W % 8 == 0 && H % 4 == 0) ? W / 8 : 0;
        PA.by_tiles = FastDiv::make((unsigned)(PA.tiles_x > 0 ? PA.tiles_x : 1));
        PA.masks = masks; PA.tiles = tile_list; PA.n_tiles = tile_info;
        if ((e = cudaMemsetAsync(counts, 0, counts_bytes_wave, st))) return e;
        if ((e = cudaMemsetAsync(Q.clk, 0, (2 + (size_t)a.max_depth + 1) * sizeof(unsigned long long), st))) return e;
        if (!fuse_primary) {
            prof_begin(kRaygen, st);
            raygen_kernel<R, Rng><<<g_simple, T, 0, st>>>(c, W, H, k, a.sample_offset + done, a.seed, a.pixel_rng, Q);
            prof_end(st);
            ++launches;
        }
        int buf = 0;
        for (int b = 0; b < a.max_depth; ++b) {
            const bool scan = b > 0 && S.scan_incoherent;
            if (b == 0 && fuse_primary) {
                prof_begin(kShade, st);
                if (primary_scan) shade_kernel<R, PcgRng, 5><<<g_primary_scan, T, smem_scan, st>>>(S, Q, buf, b, a.max_depth, PA);
                else if (planar && S.scan_incoherent) shade_kernel<R, PcgRng, 6><<<g_primary_small, T, smem_bvh_small, st>>>(S, Q, buf, b, a.max_depth, PA);
                else shade_kernel<R, PcgRng, 4><<<g_primary, T, smem_bvh, st>>>(S, Q, buf, b, a.max_depth, PA);
                prof_end(st);
                launches -= 1;
            } else if (walk_kernel) {
                prof_begin(kExtend, st);
                unsigned *cursor = (unsigned *)(fetch + b);
                bool launched = false;
                if constexpr (sizeof(R) == 4) {
                    if (walk_nodes == 1) { launch_walk<R, 1>(count_tests, g_walk, T, 0, st, S, Q, buf, b, cursor); launched = true; }
                    if (walk_nodes == 2) { launch_walk<R, 2>(count_tests, g_walk, T, 0, st, S, Q, buf, b, cursor); launched = true; }
                }
                if (!launched) launch_walk<R, 0>(count_tests, g_walk, T, smem_walk, st, S, Q, buf, b, cursor);
                prof_end(st);
                prof_begin(kShade, st);
                shade_kernel<R, Rng, 0><<<g_shade, T, 0, st>>>(S, Q, buf, b, a.max_depth, PA);
                prof_end(st);
            } else if (fused && scan && split) {
                prof_begin(kExtend, st);
                scan_hits_kernel<R><<<g_scan_hits, T, smem_scan_k, st>>>(S, Q, buf, b);
                prof_end(st);
                prof_begin(kShade, st);
                shade_kernel<R, Rng, 7><<<g_mode7, T, smem_mode7, st>>>(S, Q, buf, b, a.max_depth, PA);
                prof_end(st);
            } else if (fused) {
                prof_begin(kShade, st);
                if (scan && planar) shade_kernel<R, Rng, 3><<<g_fuse_scan, T, smem_mode3, st>>>(S, Q, buf, b, a.max_depth, PA);
                else if (scan) shade_kernel<R, Rng, 2><<<g_fuse_scan, T, smem_generic, st>>>(S, Q, buf, b, a.max_depth, PA);
                else shade_kernel<R, Rng, 1><<<g_fuse_bvh, T, smem_bvh, st>>>(S, Q, buf, b, a.max_depth, PA);
                prof_end(st);
                launches -= 1;
            } else {
                prof_begin(kExtend, st);
                extend_kernel<R><<<g_extend, T, smem, st>>>(S, Q.ro[buf], Q.rd[buf], Q.hit, Q.counts + b, scan ? 1 : 0);
                prof_end(st);
                prof_begin(kShade, st);
                shade_kernel<R, Rng, 0><<<g_shade, T, 0, st>>>(S, Q, buf, b, a.max_depth, PA);
                prof_end(st);
            }
            prof_begin(kShadow, st);
            shadow_kernel<R><<<g_shadow, T, smem_shadow, st>>>(S, Q, b);
            prof_end(st);
            launches += 3;
            buf ^= 1;
            Q.perm = nullptr;
            if (sort_rays && b + 1 < a.max_depth) {
                // the queue length lives on the device; CUB wants it on the host (one small sync per bounce,
                // negligible next to a multi-millisecond LBVH bounce)
                unsigned long long h_cnt = 0;
                if ((e = cudaMemcpyAsync(&h_cnt, counts + b + 1, sizeof h_cnt, cudaMemcpyDeviceToHost, st))) return e;
                if ((e = cudaStreamSynchronize(st))) return e;
                int n_next = (int)(h_cnt & 0xffffffffULL);
                if (n_next >= 65536) {
                    prof_begin(kMisc, st);
                    size_t tmp_bytes = L.cub_bytes;
                    if ((e = cub::DeviceRadixSort::SortPairs(cub_tmp, tmp_bytes, keys, keys_sorted, iota, perm, n_next, kSortBeginBit, 30, st)))
                        return e;
                    prof_end(st);
                    ++launches;
                    Q.perm = perm;
                }
            }
        }
        prof_begin(kAccumulate, st);
        accumulate_kernel<R><<<g_simple, T, 0, st>>>(npix, k, Q.L, (real4<R> *)a.accum, (real4<R> *)a.accum_sq, B2RT_OPT_TILE_LIST ? masks : nullptr);
        prof_end(st);
        ++launches;
        if (a.counters) {
            path_counters_kernel<<<1, 1, 0, st>>>(Q.counts, Q.unshadowed, Q.culled, Q.tally, Q.clk, Q.dead, a.max_depth,
                                                  (long long)npix * k, k, launches + 1, a.counters);
        }
        if ((e = cudaGetLastError())) return e;
    }
    return cudaSuccess;
}

template <typename R>
cudaError_t Api<R>::render_path(const b2rt_scene *s, const double *cam, const PathArgs &a, cudaStream_t st) {
    if (a.rng_mode == B2RT_RNG_REFERENCE) return render_path_impl<R, RefRng>(s, cam, a, st);
    return render_path_impl<R, PcgRng>(s, cam, a, st);
}

template <typename R>
cudaError_t Api<R>::resolve(const void *accum, int W, int H, double spp, int tonemap, uint8_t *u8, cudaStream_t st) {
    int n = W * H;
    resolve_kernel<R><<<(n + 255) / 256, 256, 0, st>>>((const real4<R> *)accum, W, H, R(spp), tonemap, u8);
    return cudaGetLastError();
}

}  // namespace b2rt
