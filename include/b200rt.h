/*
 * b200rt.h — C ABI of libb200rt.so, the B200 (sm_100a) path-tracing core.
 *
 * This is the drop-in boundary below the reference's renderer plug-in API
 * (renderers/base_renderer.py:13-16, BaseRenderer.render(scene, camera, settings)).
 * The reference has no FFI of its own: its device boundary is Numba's JIT
 * (cuda.to_device / kernel[grid, block](...) / copy_to_host,
 * renderers/cuda_path_tracer.py:785-805).  Each entry point below names the reference
 * code it replaces.  INTEGRATION.md shows the ctypes binding a reference maintainer adds.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; b2rt_last_error() has text;
 *   - pointers named d_* are DEVICE pointers owned by the caller (e.g. torch tensor
 *     data_ptr()); the library never allocates caller-visible memory; h_* are host pointers
 *     to small parameter blocks read during the call;
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream); calls are
 *     asynchronous with respect to the host unless stated otherwise;
 *   - `precision`: 0 = float32 kernels (production), 1 = float64 kernels compiled with
 *     -fmad=false (parity instantiation of the SAME templated source); it selects the element
 *     type "real" of every `real4` array in b2rt_scene;
 *   - images are in DEVICE ROW ORDER (row 0 = bottom, like the reference kernels,
 *     cuda_path_tracer.py:27) unless the function says "flipped".
 */
#ifndef B200RT_H
#define B200RT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2RT_PRECISION_F32 0
#define B2RT_PRECISION_F64 1

/* semantics flags (b2rt_scene.semantics): which reference arithmetic regime the kernels follow */
#define B2RT_SEM_NUMBA 0 /* cuda_path_tracer.py / cuda_texture_renderer.py: strict t range, `dot > 0` flips   */
#define B2RT_SEM_CPU   1 /* core/geometry.py via cpu_renderer.py: Plane accepts t == t_max, `dot >= 0` flips */

/* flags for b2rt_render_path */
#define B2RT_PATH_UNFUSED 1
#define B2RT_PATH_NO_RAY_SORT 2   /* keep queue order even if the scene asks for ray re-ordering */
#define B2RT_PATH_PRIMARY_WALK 4  /* small scenes: primary rays walk the LBVH instead of using the scan/box records
                                     (measured on the Cornell box: 22.3 vs 21.8 ms per 128 spp, so the records are the default) */

#define B2RT_PATH_FUSED_WALK 8    /* large scenes: keep the per-ray LBVH walk fused into the bounce kernel for bounces >= 1
                                     instead of the persistent walk kernel (dynamic ray fetch) + wavefront shade stage */

#define B2RT_PATH_WALK_PRIMARY 32 /* large scenes: primary rays too go through raygen + the persistent walk kernel instead of
                                     the fused first-bounce kernel (measured slightly slower: coherent rays need no re-fetch) */

#define B2RT_PATH_NO_PRIMARY_MASKS 128 /* small scenes: camera rays test every scan record behind one scene-bounds slab test
                                     instead of the per-32-pixel-tile candidate masks (measurement / validation switch) */
#define B2RT_PATH_NO_SPLIT 256    /* small float32 scenes: bounces >= 1 run the fused scan + shade kernel instead of the split pair
                                     (closest-hit scan -> hit queue -> shading with every lane on a hit); measurement switch */
#define B2RT_PATH_BINARY_WALK 512  /* large scenes: the persistent walk kernel reads the two-wide nodes even when the scene
                                     carries d_bvh_wide (measurement / validation switch: results are identical) */
#define B2RT_PATH_COUNT_TESTS 64  /* measurement passes: the persistent walk kernel tallies its box and leaf steps into
                                     d_counters[8] / [9] (a separate kernel instantiation: the timed kernels carry no counters) */

/* rng modes for b2rt_render_path */
#define B2RT_RNG_PCG       0 /* counter-based: stream keyed by (pixel, global sample index, seed)        */
#define B2RT_RNG_REFERENCE 1 /* the reference's int64 xorshift, per-pixel sequential (cuda_path_tracer.py:28,61-71) */

/*
 * Packed scene in device memory.  Primitive ids ("packed order") are: rectangles [0, n_rect),
 * spheres [n_rect, n_rect + n_sphere), triangles [.., n_prims) — the scan order of the
 * reference's cuda_scene_hit (cuda_path_tracer.py:511,582,639); ties resolve to the lowest id.
 *
 * real4 record streams (hot = read by intersection, cold = read by shading only):
 *   d_rect   [4*n_rect]   hot : (anchor.xyz, u_len) (normal.xyz, v_len) (u_unit.xyz, 0) (v_unit.xyz, 0)
 *   d_sphere [2*n_sphere] hot : (center.xyz, radius) (radius^2, 0, 0, 0)
 *   d_tri    [3*n_tri]    hot : (v0.xyz, 0) (e1.xyz, 0) (e2.xyz, 0)         e1 = v1 - v0, e2 = v2 - v0
 *   d_shade  [3*n_prims]  cold: (normal.xyz, 0) (uv0.u, uv0.v, uv1.u, uv1.v) (uv2.u, uv2.v, has_uv, 0)
 *   d_mat    [2*n_mat]    cold: (color.rgb, diffuse) (specular, reflective, refractive, ior)
 *   d_lights [n_lights]       : (pos.xyz, 0)
 * replaces the AoS float32 block of _prepare_scene_data / _prepare_light_data
 * (cuda_path_tracer.py:819-899,942-946).
 */
#define B2RT_ABI_VERSION 3       /* layout of b2rt_scene below; bumped whenever a field is added, moved or re-interpreted */
#define B2RT_SCAN_MAX_PRIMS 64   /* scenes up to this size scan all primitives for incoherent rays (scan_incoherent) */

typedef struct b2rt_scene {
    /* ABI guard: every entry point that takes a scene refuses one whose two first words differ from the library's own
     * sizeof(b2rt_scene) / B2RT_ABI_VERSION (rc 3, text in b2rt_last_error), so a caller built against another header
     * fails loudly instead of being read at wrong offsets. */
    uint32_t struct_size;    /* = sizeof(b2rt_scene) */
    uint32_t abi_version;    /* = B2RT_ABI_VERSION */
    int32_t precision;       /* element type of the real4 arrays below */
    int32_t semantics;       /* B2RT_SEM_* */
    int32_t n_rect, n_sphere, n_tri;
    int32_t n_mat, n_tex, n_lights;
    const void *d_rect, *d_sphere, *d_tri, *d_shade;
    const int32_t *d_prim_mat;   /* [n_prims] material index                                              */
    const void *d_mat;
    const int32_t *d_mat_tex;    /* [n_mat] texture id or -1                                              */
    const uint32_t *d_texels;    /* RGBX8, one 32-bit load per texel (replaces the flat RGB8 array of
                                    _prepare_texture_data, cuda_path_tracer.py:901-932)                   */
    const int32_t *d_tex_info;   /* [4*n_tex]: offset (texels), width, height, 0                          */
    const void *d_lights;
    /* LBVH produced by b2rt_lbvh_build (always float32, boxes padded outward) */
    const void *d_bvh_nodes;     /* float4[4*n_internal]                                                  */
    const void *d_bvh_top;       /* float4[4*n_top]: breadth-first copy of the top levels (staged in smem)*/
    int32_t n_bvh_top;
    int32_t bvh_root;            /* child reference of the root: >= 0 node, < 0 means ~prim               */
    int32_t scan_incoherent;     /* 1: rays after the first bounce and shadow rays scan ALL primitives in
                                    packed order (every lane of a warp tests the same primitive: no SIMT
                                    divergence) instead of walking the LBVH — faster when n_prims is a few
                                    dozen (profiles/r1a: LBVH walk 3.8-11 of 32 lanes active)             */
    int32_t n_scan_prims;        /* planar scan records below (0: the generic per-type tests are used)      */
    /* Small-scene scan records, float32 only: every rectangle, triangle and coplanar triangle PAIR forming a
     * parallelogram as one "plane + two edge planes" record of 4 float4:
     *   (N.xyz, cN)  (n1.xyz, d1)  (n2.xyz, d2)  (umax, vmax, bits(kind<<28 | idA), bits(idB))
     * t = (cN - N.o)/(N.d), P = o + t d, u = n1.P + d1, v = n2.P + d2;  kind 0 rectangle (u<=umax, v<=vmax),
     * 1 triangle (u+v<=1), 2/3 parallelogram of triangles idA (u>=v) / idB, diagonal ties to A (2) or B (3). */
    const void *d_scan_prims;
    /* Optional int32[n_lights]: for each light sample the scan record (k) or sphere (64 + i) that blocks the
     * most shadow rays towards it, precomputed on the host (packer.build_occluder_hints).  The shade stage
     * tests this one primitive before queueing a shadow ray: a hit answers the occlusion query exactly. */
    const int32_t *d_occluder_hint;
    /* > 0: scenes that walk the LBVH re-order every ray queue before it is traced (bounce >= 1): key =
     * quantised direction (6 bits) | 24-bit Morton code of the origin quantised over [-extent, extent]^3,
     * radix-sorted; the next bounce reads its rays through the permutation.  0 disables the sort. */
    float ray_sort_extent;
    /* Box records (float32 small-scene scan, packer.group_scan_boxes): planar records that are faces of a common
     * parallelepiped (the Cornell walls, a cube) are tested together with ONE three-slab test in the box's own
     * coordinates.  d_scan_prims then holds n_scan_prims planar records — the first n_scan_loose belong to no
     * box and are scanned one by one, the rest are box faces kept for (u, v) / triangle-id recovery and as
     * occluder hints — followed by n_scan_boxes box records of 4 float4:
     *   (m0.xyz, d0) (m1.xyz, d1) (m2.xyz, d2)   l_k = m_k.P + d_k in [-1, 1] inside the box
     *   (bits(f0|f1<<8|f2<<16|f3<<24), bits(f4|f5<<8), bits(flags), -)   f[2k + (l_k == +1)] = planar record of that
     *   face, 255 = none; the two upper bytes of the second word must be 0; flags bit 0 = CLOSED (all six faces exist:
     *   the kernels then use a three-slab test without per-face bookkeeping and recover the face from the hit point)
     * With n_scan_boxes == 0 set n_scan_loose = n_scan_prims. */
    int32_t n_scan_loose;
    int32_t n_scan_boxes;
    int32_t bvh_rects_outside;   /* 1: the hierarchy was built with B2RT_LBVH_RECTS_OUTSIDE */
    /* Optional (small float32 scenes; NULL otherwise): float4[5*n_prims] per-primitive shading records, staged in
     * shared memory by the bounce kernels (packer.build_surface_records) — one branch-free, one-hop record instead
     * of the per-type branches and the prim -> material -> texture chain of dependent loads:
     *   (n.xyz | sphere centre.xyz, 1/radius or 0) (color.rgb, diffuse) (specular, reflective, refractive, ior)
     *   (u0, du_a, du_b, bits(texture id)) (v0, dv_a, dv_b, bits(flags))   uv = uv0 + a*d_a + b*d_b;
     *   flags bit 0 = flip the normal to face the ray (triangles) */
    const void *d_surface_records;
    /* Outward-padded bounds of all primitives (lo > hi: unknown).  Camera rays of small scenes are tested against
     * them first: on the Cornell box 51 % of the primary rays miss the scene and skip the record scan. */
    float bounds_lo[3], bounds_hi[3];
    /* Optional (ABI 3; NULL: the two-wide nodes are walked): 4-wide nodes written by b2rt_lbvh_widen, float4[8 * (n_bvh_top +
     * n_internal)].  The persistent walk kernel that traces the incoherent rays (bounce >= 1) of float32 scenes too large
     * for the record scan reads these instead of d_bvh_nodes: half the dependent fetches per ray. */
    const void *d_bvh_wide;
    /* Optional (ABI 3; NULL: not used): quantised binary nodes written by b2rt_lbvh_quantize, 32 B header + 32 B per child
     * reference.  When present the persistent walk kernel reads these in preference to d_bvh_wide / d_bvh_nodes: a node is
     * two 16-byte loads instead of four. */
    const void *d_bvh_quant;
} b2rt_scene;

const char *b2rt_last_error(void);

/* ---- scene preparation: the optional small-scene acceleration data, derived INSIDE the library ---------------------
 * b2rt_scene's base streams (d_rect .. d_lights, d_prim_mat, d_mat, d_mat_tex) are all a caller has to fill.  The
 * optional fields that make small scenes fast — scan_incoherent, d_scan_prims / n_scan_prims / n_scan_loose /
 * n_scan_boxes (planar + box records), d_surface_records, d_occluder_hint, bounds_lo / bounds_hi — are derived from
 * them by b2rt_scene_prepare (the reference has no counterpart: its cuda_scene_hit tests every primitive in turn,
 * cuda_path_tracer.py:496-730).  Python's b200rt.packer keeps an independent numpy implementation of the same
 * derivations; the CPU tests compare the two. */
#define B2RT_PREPARE_NO_BOXES   1   /* planar records only (no three-slab box records) */
#define B2RT_PREPARE_NO_SURFACE 2   /* no surface records: the bounce kernels use the generic shade / material streams */
#define B2RT_PREPARE_NO_HINTS   4   /* no occluder hints */
typedef struct b2rt_prepare_layout {   /* where b2rt_scene_prepare_host put what (byte offsets into its output buffer) */
    size_t scan_offset, surface_offset, hint_offset;      /* (size_t)-1: not produced */
    size_t bytes_used;
    int32_t n_scan_prims, n_scan_loose, n_scan_boxes, scan_incoherent;
    float bounds_lo[3], bounds_hi[3];
} b2rt_prepare_layout;
/* upper bound on the bytes b2rt_scene_prepare / _host write for a scene with these counts */
int b2rt_scene_prepare_bytes(int32_t n_rect, int32_t n_sphere, int32_t n_tri, int32_t n_lights, size_t *h_bytes);
/* Reads the float32 base streams of `scene` back from the device (a few KB: only scenes of <= B2RT_SCAN_MAX_PRIMS
 * primitives get records), derives the records on the host, uploads them into d_buffer (caller-owned, at least
 * b2rt_scene_prepare_bytes) and points the optional fields of `scene` at them.  Larger, float64 or CPU-semantics
 * scenes get scan_incoherent / bounds set and nothing else.  Synchronises the stream. */
int b2rt_scene_prepare(struct b2rt_scene *scene, void *d_buffer, size_t buffer_bytes, int32_t flags, void *stream);
/* The same derivation on HOST copies of the float32 streams (layouts as documented at b2rt_scene) into a host buffer:
 * for binders that pack on the host and upload everything in one copy, and for tests without a GPU. */
int b2rt_scene_prepare_host(int32_t n_rect, int32_t n_sphere, int32_t n_tri, int32_t n_mat, int32_t n_lights,
                            const float *h_rect, const float *h_sphere, const float *h_tri, const float *h_shade,
                            const float *h_mat, const int32_t *h_prim_mat, const int32_t *h_mat_tex,
                            const float *h_lights, int32_t flags, void *h_out, size_t out_bytes,
                            b2rt_prepare_layout *h_layout);
int b2rt_version(void);
/* h_out[0..5] = SM count, max smem per block (opt-in), L2 bytes, SM clock kHz, cc major, cc minor */
int b2rt_device_info(int device, int64_t *h_out);

/* ---- LBVH (replaces BVHNode.__init__, core/acceleration.py:8-30, with a Morton-code LBVH) -------------- */
/* bytes of scratch b2rt_lbvh_build needs for n_prims primitives */
int b2rt_lbvh_temp_bytes(int32_t n_prims, size_t *h_bytes);
/*
 * Builds the hierarchy over FLOAT32 geometry streams (same layouts as d_rect/d_sphere/d_tri with
 * real = float; for an f64 scene pass float copies).  box_pad is added to every box face.
 * d_nodes_out: float4[4*(n_prims-1)], d_top_out: float4[4*top_capacity].
 * h_meta_out[0] = n_top, [1] = root reference, [2] = number of internal nodes.
 * Synchronises the stream before returning (h_meta_out is host memory).
 */
int b2rt_lbvh_build(int32_t n_rect, int32_t n_sphere, int32_t n_tri,
                    const void *d_rect_f32, const void *d_sphere_f32, const void *d_tri_f32,
                    float box_pad, void *d_nodes_out, void *d_top_out, int32_t top_capacity,
                    int32_t *h_meta_out, void *d_temp, size_t temp_bytes, void *stream, int32_t flags);
/* 4-wide nodes from a finished hierarchy (d_nodes / d_top / n_top / n_internal exactly as b2rt_lbvh_build left them):
 * entry `ref` (a child reference: < n_top top copy, else n_top + node index) holds the boxes and references of that
 * node's GRANDCHILDREN — 8 float4 = one 128 B line: lo.x[4] lo.y[4] lo.z[4] hi.x[4] hi.y[4] hi.z[4] bits(ref[4]) unused;
 * a leaf child stays one slot, an empty slot has ref 0x80000000.  Asynchronous on the stream.  (The reference has no
 * counterpart: its BVHNode is a binary tree of Python objects, core/acceleration.py:8-30.) */
int b2rt_lbvh_wide_bytes(int32_t n_top, int32_t n_internal, size_t *h_bytes);
int b2rt_lbvh_widen(const void *d_nodes, const void *d_top, int32_t n_top, int32_t n_internal, void *d_wide_out,
                    size_t wide_bytes, void *stream);
/* Quantised binary nodes from a finished hierarchy, for the persistent walk kernel: both child boxes of node `ref` as
 * 16-bit cell indices on a 65536^3 grid over [h_lo, h_hi] (the bounds of ALL primitives incl. the box pad; boxes outside are
 * clamped to the last cell, which is right for the far-away placeholders of B2RT_LBVH_RECTS_OUTSIDE and wrong for
 * anything else), every face moved outward by one whole cell on top of the outward rounding so that the float32
 * dequantisation in the kernel stays conservative.  Layout: 2 float4 header (base.xyz, 0) (cell.xyz, 0), then per
 * reference 8 words: L.x L.y L.z R.x R.y R.z as lo | hi << 16, ref L, ref R.  Asynchronous on the stream. */
int b2rt_lbvh_quant_bytes(int32_t n_top, int32_t n_internal, size_t *h_bytes);
int b2rt_lbvh_quantize(const void *d_nodes, const void *d_top, int32_t n_top, int32_t n_internal, const float *h_lo,
                       const float *h_hi, void *d_quant_out, size_t quant_bytes, void *stream);
/* flags for b2rt_lbvh_build */
#define B2RT_LBVH_NO_ROTATIONS 2  /* skip the bottom-up tree-rotation pass (measurement switch) */
#define B2RT_LBVH_RECTS_OUTSIDE 1 /* the rectangles get no place in the hierarchy (set b2rt_scene.bvh_rects_outside too: every
                                     walk then tests them directly first).  For a few room-sized rectangles around a fine
                                     mesh: inside the tree they widen every ancestor box of their leaves. */

/* ---- closest hit (replaces cuda_scene_hit, cuda_path_tracer.py:496-730, and Scene.hit, core/scene.py:45) */
/* One primary ray per pixel at sub-pixel offset (du, dv): u = (x+du)/W, v = (y+dv)/H
 * (cuda_get_ray, cuda_path_tracer.py:84-112 / Camera.get_ray, core/camera.py:26).
 * h_cam: 12 doubles origin, lower_left, horizontal, vertical.  d_ids: int32[H*W] packed prim id or -1;
 * d_t: float64[H*W] hit distance or -1 (may be NULL).  use_bvh = 0 scans all primitives (debug/validation). */
int b2rt_primary_hits(const b2rt_scene *scene, const double *h_cam, int32_t width, int32_t height,
                      double du, double dv, double t_min, double t_max, int32_t use_bvh,
                      int32_t *d_ids, double *d_t, void *stream);
/* Explicit rays: d_o, d_d float64[3*n].  d_rec (optional) float64[9*n]: t, point(3), normal(3), uv(2).
 * use_bvh: 1 LBVH walk, 0 generic scan of all primitives, 2 small-scene scan records (d_scan_prims). */
int b2rt_trace_rays(const b2rt_scene *scene, int32_t n, const double *d_o, const double *d_d,
                    double t_min, double t_max, int32_t any_hit, int32_t use_bvh,
                    int32_t *d_ids, double *d_rec, void *stream);

/* ---- Whitted integrators ------------------------------------------------------------------------------- */
/* CPURenderer._trace semantics (renderers/cpu_renderer.py:75-151), one sample per pixel.
 * d_jitter: float64[2*H*W] (du, dv) per pixel or NULL for pixel centres.  h_ambient / h_light_color: 3 doubles.
 * d_rgb: float64[3*H*W] pre-quantisation colour. */
int b2rt_render_whitted_cpu(const b2rt_scene *scene, const double *h_cam, int32_t width, int32_t height,
                            const double *d_jitter, int32_t max_depth, const double *h_ambient,
                            const double *h_light_color, double *d_rgb, void *stream);
/* cuda_trace_kernel + cuda_trace_ray semantics (renderers/cuda_texture_renderer.py:17-73,173-430):
 * int(sqrt(spp))^2 jittered samples with the reference's LCG, divided by spp.
 * d_rgb (optional): float64[3*H*W] mean; d_u8 (optional): uint8[3*H*W] min(255,max(0,int(c*255))). */
int b2rt_render_whitted_texture(const b2rt_scene *scene, const double *h_cam, int32_t width, int32_t height,
                                int32_t spp, int32_t max_depth, double *d_rgb, uint8_t *d_u8, void *stream);

/* ---- wavefront path tracer (replaces cuda_path_trace_kernel + cuda_trace_path, cuda_path_tracer.py:17-471) */
/* scratch bytes for a wave of `spp_per_wave` samples per pixel */
int b2rt_path_workspace_bytes(int32_t precision, int32_t width, int32_t height, int32_t spp_per_wave,
                              int32_t max_depth, size_t *h_bytes);
/*
 * Adds `spp_local` samples per pixel (global sample indices sample_offset .. sample_offset+spp_local-1)
 * to d_accum (real[4*H*W]: sum r, g, b, unused), in waves of spp_per_wave.  d_accum_sq (optional, same
 * shape) receives the per-pixel sum of SQUARED per-sample radiance, for Monte-Carlo variance estimates.
 * rng_mode PCG: seed keys the streams.  rng_mode REFERENCE: seed is the reference's frame_count and
 * d_pixel_rng (int64[H*W], caller-zeroed before sample 0... see DESIGN.md) carries the per-pixel state.
 * flags: B2RT_PATH_UNFUSED runs extend and shade as separate kernels through the hit stream (the default
 * fuses them: closest hit and shading in one kernel per bounce).
 * d_counters (optional) uint64[16], accumulated: [0] paths, [1] closest-hit rays, [2] shadow rays answered
 * (queued + resolved by the occluder cache), [3] unshadowed light samples, [4] kernel launches made by this
 * call, [5] shadow rays resolved by the occluder cache without being queued, [6] camera rays answered by the
 * scene-bounds slab test alone (no record scan), [7] shaded path segments (closest hits that were shaded),
 * [8] / [9] per-lane box steps / leaf steps of the persistent walk kernel (only with B2RT_PATH_COUNT_TESTS),
 * [10] canonical flops (box record 42, planar record 33, sphere 28) of the camera rays' masked record tests,
 * [11] / [12] SM cycles / nanoseconds that CTA 0 of every bounce kernel was resident (their ratio is the effective SM
 * clock the kernels saw), [13..15] reserved (zero).
 */
int b2rt_render_path(const b2rt_scene *scene, const double *h_cam, int32_t width, int32_t height,
                     int32_t spp_local, int64_t sample_offset, int32_t spp_per_wave, int32_t max_depth,
                     int32_t rng_mode, uint64_t seed, int32_t flags, void *d_accum, void *d_accum_sq,
                     int64_t *d_pixel_rng,
                     void *d_workspace, size_t workspace_bytes, uint64_t *d_counters, void *stream);
/* mean = accum / spp_total, optional ACES tonemap (cuda_tonemap, :74-81), quantise (:56-58), and write the
 * FLIPPED image (row 0 = top, replaces np.flip, :807) into d_u8 uint8[3*H*W]. */
int b2rt_resolve(int32_t precision, const void *d_accum, int32_t width, int32_t height, double spp_total,
                 int32_t tonemap, uint8_t *d_u8, void *stream);

/* ---- multi-GPU: samples split across ranks, float buffers combined over NVLink (the reference is
 *      single-GPU, cuda.select_device(0), cuda_path_tracer.py:743) ------------------------------------------- */
/*
 * Fused reduce + resolve over PEER memory.  h_peer_accum[p] = device pointer to rank p's float32 accumulation buffer
 * (float4[H*W], device row order), all of them mapped into this process (symmetric / IPC memory).  The caller's rank
 * sums rows [row0, row1) of every peer IN RANK ORDER, divides by spp_total, tone-maps, quantises and writes the FLIPPED
 * bytes into d_u8_root (uint8[3*H*W], normally the root's image buffer: a peer pointer) and, when d_sum_root is not
 * NULL, the float sums into d_sum_root (float4[H*W]).  With every rank taking H / n_peers rows this is reduce-scatter,
 * resolve and gather in one kernel.  The caller orders it after all ranks' b2rt_render_path (a cross-rank barrier on
 * the stream) and before the root reads the image (another one).  float32 only.
 */
int b2rt_reduce_resolve(const void *const *h_peer_accum, int32_t n_peers, int32_t width, int32_t height, int32_t row0,
                        int32_t row1, double spp_total, int32_t tonemap, uint8_t *d_u8_root, void *d_sum_root, void *stream);

/* ---- texture upload helper (replaces the RGB flattening of _prepare_texture_data, cuda_path_tracer.py:901-932) */
/* d_rgb: n_texels packed RGB8 triples (4-byte aligned) -> d_rgbx: n_texels RGBX8 words (16-byte aligned) */
int b2rt_expand_rgb8(const uint8_t *d_rgb, int64_t n_texels, uint32_t *d_rgbx, void *stream);

/* ---- checked build (the compute-sanitizer substitute) ------------------------------------------------------------------
 * A library compiled with -DB2RT_CHECK=1 bounds-checks every traversal-stack push and every queue append inside the
 * kernels; violations are counted and the offending write is dropped (the context stays usable).  b2rt_check_read
 * returns and clears the counts of the current device (both 0 in a normal build, where b2rt_check_enabled() is 0). */
int b2rt_check_enabled(void);
int b2rt_check_read(uint64_t *h_stack_overflows, uint64_t *h_queue_overruns);

/* ---- measurement helpers (no reference counterpart: the reference times render() with time.time(),
 *      main.py:89-91) ----------------------------------------------------------------------------------- */
/* When on, b2rt_render_path brackets every kernel launch with CUDA events on its stream. */
int b2rt_profile_enable(int32_t on);
/* Synchronises the recorded events; h_ms[8] / h_launches[8] = device milliseconds and launch counts since
 * the last read, per kernel class: 0 raygen, 1 extend, 2 shade, 3 shadow, 4 accumulate. */
int b2rt_profile_read(double *h_ms, int64_t *h_launches);
/* FP32 FMA micro-benchmark (8 independent chains per thread, SMs x 8 CTAs x 256 threads): the measured
 * non-tensor FP32 peak that the FP32 roofline fraction is quoted against. */
int b2rt_fp32_peak(int32_t iters, double *h_tflops, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* B200RT_H */
