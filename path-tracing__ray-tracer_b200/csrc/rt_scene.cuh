// rt_scene.cuh — device scene view, primitive intersection and stack-based LBVH traversal.
//
// Reference functions restated here (operation order kept so that the float64 instantiation,
// compiled with -fmad=false, reproduces the reference's float64 results bit for bit):
//   rectangle : Plane.hit      core/geometry.py:50-72     / cuda_scene_hit :511-574
//   sphere    : Sphere.hit     core/geometry.py:85-111    / cuda_scene_hit :582-631
//   triangle  : Triangle.hit   core/geometry.py:139-171   / cuda_scene_hit :639-728
//   closest   : BVHNode.hit / Scene.hit (core/acceleration.py:32-40) replaced by an LBVH walk whose
//               result equals the brute-force scan's: min t, ties -> lowest packed id.
#pragma once
#include "rt_math.cuh"
#include "rt_api.h"

namespace b2rt {

#ifndef B2RT_SCAN_UNROLL
#define B2RT_SCAN_UNROLL 1            // loose planar records per loop trip (box records left few of them)
#endif
#ifndef B2RT_BOUNCE_MIN_BLOCKS
#define B2RT_BOUNCE_MIN_BLOCKS 4      // resident CTAs/SM requested for the float32 planar-scan bounce kernel (64 regs)
#endif
#ifndef B2RT_WHILE_WHILE
#define B2RT_WHILE_WHILE 0            // LBVH walk style (see traverse); measured per round in profiles/
#endif
#ifndef B2RT_SORT_KEY
#define B2RT_SORT_KEY 2               // ray re-ordering key layout (see ray_sort_key; measured on C4: 0: 624, 1: 586, 2: 656 Mpaths/s)
#endif
#ifndef B2RT_BVH_MIN_BLOCKS
#define B2RT_BVH_MIN_BLOCKS 4         // ... for the float32 LBVH-walk bounce kernels (measured 6729 vs 6600 Mpaths/s with 3)
#endif
// B2RT_CHECK build (the compute-sanitizer substitute: this pool refuses compute-sanitizer): every traversal-stack push and
// every queue append is bounds-checked in the kernel; a violation is COUNTED (SceneDev.check: stack overflows in the low
// word, queue overruns in the high word) and the offending write is skipped, so the context survives and the host can
// assert on b2rt_check_read().  B2RT_STACK_DEPTH shrinks the stack to provoke the check in the negative test.
#ifndef B2RT_CHECK
#define B2RT_CHECK 0
#endif
#ifndef B2RT_STACK_DEPTH
#define B2RT_STACK_DEPTH 64
#endif
constexpr int kStackDepth = B2RT_STACK_DEPTH;
constexpr int kWideStackDepth = B2RT_STACK_DEPTH + B2RT_STACK_DEPTH / 2;   // 4-wide walk: up to three pushes per level, half the levels
constexpr int kScanUnroll = B2RT_SCAN_UNROLL;       // LBVH depth bound: 30 Morton bits + log2(duplicates)

struct SceneDev {
    int n_rect, n_sphere, n_tri, n_prims;
    int n_mat, n_tex, n_lights;
    int semantics;
    const void *rect, *sphere, *tri, *shade, *mat, *lights;
    const int *prim_mat, *mat_tex;
    const uint32_t *texels;
    const int4 *tex_info;
    const float4 *nodes, *top;
    const float4 *wide;                // 4-wide nodes (b2rt_lbvh_widen) or nullptr
    const float4 *quant;               // quantised 32 B nodes behind a 32 B header (b2rt_lbvh_quantize) or nullptr
    int n_top, root;
    int scan_incoherent;
    int n_outside;                     // rectangles [0, n_outside) are not in the hierarchy: tested before every walk
    int n_scan, n_loose, n_box;        // planar scan records (loose ones first), box records behind them
    const float4 *scan;
    const int *occl_hint;
    const float4 *surf;                // per-primitive shading records of small float32 scenes (or nullptr)
    float sort_inv;                    // 0.5 / ray_sort_extent, or 0 when ray sorting is off
    float blo[3], bhi[3];              // padded scene bounds (blo > bhi: unknown)
    unsigned long long *check;         // B2RT_CHECK builds: violation counter (library-owned), else nullptr
};

inline SceneDev make_scene_dev(const b2rt_scene *s) {
    SceneDev d;
    d.n_rect = s->n_rect; d.n_sphere = s->n_sphere; d.n_tri = s->n_tri;
    d.n_prims = s->n_rect + s->n_sphere + s->n_tri;
    d.n_mat = s->n_mat; d.n_tex = s->n_tex; d.n_lights = s->n_lights;
    d.semantics = s->semantics;
    d.rect = s->d_rect; d.sphere = s->d_sphere; d.tri = s->d_tri; d.shade = s->d_shade;
    d.mat = s->d_mat; d.lights = s->d_lights;
    d.prim_mat = s->d_prim_mat; d.mat_tex = s->d_mat_tex;
    d.texels = s->d_texels; d.tex_info = reinterpret_cast<const int4 *>(s->d_tex_info);
    d.nodes = reinterpret_cast<const float4 *>(s->d_bvh_nodes);
    d.top = reinterpret_cast<const float4 *>(s->d_bvh_top);
    d.wide = reinterpret_cast<const float4 *>(s->d_bvh_wide);
    d.quant = reinterpret_cast<const float4 *>(s->d_bvh_quant);
    d.n_top = s->n_bvh_top; d.root = s->bvh_root;
    d.scan_incoherent = s->scan_incoherent;
    d.n_outside = s->bvh_rects_outside ? s->n_rect : 0;
    d.n_scan = s->precision == B2RT_PRECISION_F32 ? s->n_scan_prims : 0;
    d.n_box = d.n_scan > 0 ? s->n_scan_boxes : 0;
    d.n_loose = d.n_box > 0 ? s->n_scan_loose : d.n_scan;
    d.scan = reinterpret_cast<const float4 *>(s->d_scan_prims);
    d.occl_hint = s->precision == B2RT_PRECISION_F32 ? s->d_occluder_hint : nullptr;
    d.surf = s->precision == B2RT_PRECISION_F32 ? reinterpret_cast<const float4 *>(s->d_surface_records) : nullptr;
    for (int k = 0; k < 3; ++k) { d.blo[k] = s->bounds_lo[k]; d.bhi[k] = s->bounds_hi[k]; }
    d.check = nullptr;
    d.sort_inv = (s->ray_sort_extent > 0.f && !s->scan_incoherent) ? 0.5f / s->ray_sort_extent : 0.f;
    return d;
}

// closest-hit record: a/b are (u_hit, v_hit) in world units for a rectangle, the barycentrics
// (u, v) for a triangle, unused for a sphere
// bounds-checked traversal-stack push (plain store unless B2RT_CHECK)
#define B2RT_PUSH_N(S_, stack_, sp_, v_, depth_)                                                    \
    do {                                                                                            \
        if (B2RT_CHECK && (sp_) >= (depth_)) { if ((S_).check) atomicAdd((S_).check, 1ULL); }       \
        else (stack_)[(sp_)++] = (v_);                                                              \
    } while (0)
#define B2RT_PUSH(S_, stack_, sp_, v_) B2RT_PUSH_N(S_, stack_, sp_, v_, kStackDepth)

template <typename R> struct Hit {
    R t, a, b;
    int prim;
};

template <typename R> struct Ray {
    V3<R> o, d;
};

// ---------------------------------------------------------------------------------- primitives
// Every test returns true and fills (t, a, b) when the primitive is hit inside the open interval
// (t_min, t_far); the caller applies the tie rule.  CPU semantics: Plane accepts the closed range.

template <typename R, bool CpuSem>
__device__ __forceinline__ bool hit_rect(const SceneDev &S, int i, const Ray<R> &r, R t_min, R t_far, bool allow_eq,
                                         R &t_out, R &a_out, R &b_out) {
    const real4<R> *q = reinterpret_cast<const real4<R> *>(S.rect) + 4 * i;
    real4<R> r0 = ldg4(q), r1 = ldg4(q + 1);
    V3<R> anchor = xyz<R>(r0), n = xyz<R>(r1);
    R denom = dot(n, r.d);
    if (CpuSem ? (abs_(denom) < R(1e-6)) : !(abs_(denom) > R(1e-6))) return false;
    R t = div_(dot(anchor - r.o, n), denom);
    if (CpuSem) {                                            // closed range (core/geometry.py:56)
        if (t < t_min || !(t < t_far || (allow_eq && t == t_far))) return false;
    } else {
        if (!(t_min < t && (t < t_far || (allow_eq && t == t_far)))) return false;
    }
    real4<R> r2 = ldg4(q + 2), r3 = ldg4(q + 3);
    V3<R> p = r.o + r.d * t;
    V3<R> rel = p - anchor;
    R uh = dot(rel, xyz<R>(r2)), vh = dot(rel, xyz<R>(r3));
    if (!(R(0) <= uh && uh <= r0.w && R(0) <= vh && vh <= r1.w)) return false;
    t_out = t; a_out = uh; b_out = vh;
    return true;
}

template <typename R>
__device__ __forceinline__ bool hit_sphere(const SceneDev &S, int i, const Ray<R> &r, R t_min, R t_far, bool allow_eq,
                                           R &t_out) {
    const real4<R> *q = reinterpret_cast<const real4<R> *>(S.sphere) + 2 * i;
    real4<R> s0 = ldg4(q);
    V3<R> oc = r.o - xyz<R>(s0);
    R a = dot(r.d, r.d);
    R b = dot(oc, r.d);
    R disc, r2;
    if constexpr (sizeof(R) == 8) {
        r2 = ldg4(q + 1).x;                                   // radius^2 as the reference rounds it
        R c = dot(oc, oc) - r2;
        disc = b * b - a * c;                                 // textbook half-b form (:601-606)
    } else {
        // float32: c = |oc|^2 - r^2 cancels catastrophically from 50 units away;
        // b^2 - a*c == a * (r^2 - |oc - (b/a) d|^2) is the same quantity without the cancellation.
        r2 = s0.w * s0.w;
        V3<R> perp = oc - r.d * div_(b, a);
        disc = a * (r2 - dot(perp, perp));
    }
    if (!(disc > R(0))) return false;
    R sq = sqrt_(disc);
    R t1, t2;
    if constexpr (sizeof(R) == 4) { R ia = rcp_(a); t1 = (-b - sq) * ia; t2 = (-b + sq) * ia; }
    else { t1 = (-b - sq) / a; t2 = (-b + sq) / a; }
    // nearest root beyond t_min, then the (t_min, t_far) range test — equals the reference's
    // "t1 if in range else t2 if in range" because t2 > t1
    R t;
    if (t_min < t1) t = t1; else if (t_min < t2) t = t2; else return false;
    if (!(t < t_far || (allow_eq && t == t_far))) {
        // the reference falls through to t2 when t1 >= closest; t2 > t1 so it fails too
        return false;
    }
    t_out = t;
    return true;
}

template <typename R>
__device__ __forceinline__ bool hit_tri(const SceneDev &S, int i, const Ray<R> &r, R t_min, R t_far, bool allow_eq,
                                        R &t_out, R &u_out, R &v_out) {
    const real4<R> *q = reinterpret_cast<const real4<R> *>(S.tri) + 3 * i;
    V3<R> v0, e1, e2;
    if constexpr ((B2RT_L2_HINT & 2) != 0 && sizeof(R) == 4) {
        const unsigned long long pol = l2_keep_policy();
        v0 = xyz<R>(ldg4_keep(q, pol)); e1 = xyz<R>(ldg4_keep(q + 1, pol)); e2 = xyz<R>(ldg4_keep(q + 2, pol));
    } else {
        v0 = xyz<R>(ldg4(q)); e1 = xyz<R>(ldg4(q + 1)); e2 = xyz<R>(ldg4(q + 2));
    }
    V3<R> h = cross(r.d, e2);
    R a = dot(e1, h);
    if constexpr (sizeof(R) == 4) {
        // float32: every rejection test is done on the un-divided determinants (u = U/a, v = V/a,
        // t = T/a; multiply the inequalities by |a|), so the division runs only for accepted hits and
        // the common path is branch-free: ~30 FMA-class instructions per triangle.
        V3<R> s = r.o - v0;
        R U = dot(s, h);
        V3<R> qq = cross(s, e1);
        R V = dot(r.d, qq), T = dot(e2, qq);
        R aa = abs_(a);
        int sg = __float_as_int(a) & 0x80000000;
        R Us = __int_as_float(__float_as_int(U) ^ sg), Vs = __int_as_float(__float_as_int(V) ^ sg),
          Ts = __int_as_float(__float_as_int(T) ^ sg);
        // the range pre-check is a hair wide (2^-20 relative): T <= t_far * |a| in float32 can reject a hit whose divided
        // distance f * T EQUALS t_far, which the exact tie rule below has to see (coplanar overlapping triangles:
        // the walk reaches them in tree order, the scan in id order, and both must keep the lowest id)
        bool ok = !(aa < R(1e-6)) && Us >= R(0) && Us <= aa && Vs >= R(0) && Us + Vs <= aa &&
                  Ts > t_min * aa * R(0.999999) && Ts <= t_far * aa * R(1.000001);
        if (!ok) return false;
        R f = rcp_(a);
        R t = f * T;
        if (!(t_min < t && (t < t_far || (allow_eq && t == t_far)))) return false;
        t_out = t; u_out = f * U; v_out = f * V;
        return true;
    }
    if (abs_(a) < R(1e-6)) return false;
    R f = R(1) / a;
    V3<R> s = r.o - v0;
    R u = f * dot(s, h);
    if (u < R(0) || u > R(1)) return false;
    V3<R> qq = cross(s, e1);
    R v = f * dot(r.d, qq);
    if (v < R(0) || u + v > R(1)) return false;
    R t = f * dot(e2, qq);
    if (!(t_min < t && (t < t_far || (allow_eq && t == t_far)))) return false;
    t_out = t; u_out = u; v_out = v;
    return true;
}

// Tests packed primitive `prim` and folds it into the running closest hit with the tie rule
// "smaller t wins; equal t -> lower packed id" (the brute-force scan order of cuda_scene_hit).
template <typename R, bool CpuSem>
__device__ __forceinline__ void test_prim(const SceneDev &S, int prim, const Ray<R> &r, R t_min, Hit<R> &best) {
    R t, a = R(0), b = R(0);
    bool allow_eq;
    if (CpuSem) {
        // The reference BVH visits its leaves in scene.objects order (core/acceleration.py:20-26) and
        // Plane.hit accepts t == t_max while Sphere/Triangle are strict (core/geometry.py:56,94,159):
        // at equal t a later rectangle replaces anything, a later sphere/triangle replaces nothing.
        bool cand_rect = prim < S.n_rect, best_rect = best.prim >= 0 && best.prim < S.n_rect;
        if (best.prim < 0) allow_eq = cand_rect;             // t == t_max itself is inside a Plane's range
        else allow_eq = cand_rect ? (!best_rect || prim > best.prim) : (!best_rect && prim < best.prim);
    } else {
        allow_eq = best.prim >= 0 && prim < best.prim;       // an equal-t hit may replace a higher id
    }
    bool ok;
    if (prim < S.n_rect) ok = hit_rect<R, CpuSem>(S, prim, r, t_min, best.t, allow_eq, t, a, b);
    else if (prim < S.n_rect + S.n_sphere) ok = hit_sphere<R>(S, prim - S.n_rect, r, t_min, best.t, allow_eq, t);
    else ok = hit_tri<R>(S, prim - S.n_rect - S.n_sphere, r, t_min, best.t, allow_eq, t, a, b);
    if (ok) { best.t = t; best.a = a; best.b = b; best.prim = prim; }
}

// ---------------------------------------------------------------------------------- traversal
// LBVH node = 4 float4: child boxes + child references.
//   n0 = (L.min.xyz, L.max.x)  n1 = (L.max.yz, R.min.xy)  n2 = (R.min.z, R.max.xyz)
//   n3 = (bits(left ref), bits(right ref), -, -)     ref >= 0: node index (< n_top: in the smem copy),
//                                                    ref <  0: leaf holding packed primitive ~ref
// Boxes are float32, padded outward at build time; the slab test runs in R on the widened values.

template <typename R>
__device__ __forceinline__ bool slab(R bx0, R by0, R bz0, R bx1, R by1, R bz1, const V3<R> &o, const V3<R> &id,
                                     R t_min, R t_far, R &t_enter) {
    R tx0 = (bx0 - o.x) * id.x, tx1 = (bx1 - o.x) * id.x;
    R ty0 = (by0 - o.y) * id.y, ty1 = (by1 - o.y) * id.y;
    R tz0 = (bz0 - o.z) * id.z, tz1 = (bz1 - o.z) * id.z;
    R tn = max_(max_(min_(tx0, tx1), min_(ty0, ty1)), max_(min_(tz0, tz1), t_min));
    R tf = min_(min_(max_(tx0, tx1), max_(ty0, ty1)), min_(max_(tz0, tz1), t_far));
    t_enter = tn;
    return tn <= tf;
}

// Closest hit (AnyHit = false) or occlusion query (AnyHit = true) over (t_min, t_max).
// s_top: shared-memory copy of S.top (may be nullptr when S.n_top == 0).
template <typename R, bool CpuSem, bool AnyHit>
__device__ __forceinline__ bool traverse(const SceneDev &S, const float4 *s_top, const Ray<R> &r, R t_min, R t_max,
                                         Hit<R> &best) {
    best.t = t_max; best.prim = -1; best.a = R(0); best.b = R(0);
    if (S.n_prims == 0) return false;
    V3<R> id = {rcp_(r.d.x), rcp_(r.d.y), rcp_(r.d.z)};
#if B2RT_WHILE_WHILE
    // "while-while" walk: every lane first descends through internal nodes until it holds a leaf, then the
    // leaves are intersected together.
    constexpr int kDone = (int)0x80000000;                       // sentinel below every leaf reference
    int stack[kStackDepth];
    stack[0] = kDone;
    int sp = 1;
    int ref = S.root;
    if (S.n_outside > 0) {                                       // see the if-if variant below
        B2RT_PUSH(S, stack, sp, S.root);
        for (int p = S.n_outside - 1; p >= 1; --p) B2RT_PUSH(S, stack, sp, ~p);
        ref = ~0;
    }
    while (ref != kDone) {
        while (ref >= 0) {
            float4 n0, n1, n2, n3;
            if (ref < S.n_top) {
                const float4 *p = s_top + 4 * ref;
                n0 = p[0]; n1 = p[1]; n2 = p[2]; n3 = p[3];
            } else {
                const float4 *p = S.nodes + 4 * (size_t)(ref - S.n_top);
                n0 = __ldg(p); n1 = __ldg(p + 1); n2 = __ldg(p + 2); n3 = __ldg(p + 3);
            }
            R tl, tr;
            bool hl = slab<R>(n0.x, n0.y, n0.z, n0.w, n1.x, n1.y, r.o, id, t_min, best.t, tl);
            bool hr = slab<R>(n1.z, n1.w, n2.x, n2.y, n2.z, n2.w, r.o, id, t_min, best.t, tr);
            int cl = __float_as_int(n3.x), cr = __float_as_int(n3.y);
            if (hl && hr) {
                bool swap = tr < tl;
                B2RT_PUSH(S, stack, sp, swap ? cl : cr);
                ref = swap ? cr : cl;
            } else if (hl) {
                ref = cl;
            } else if (hr) {
                ref = cr;
            } else {
                ref = stack[--sp];
            }
        }
        while (ref < 0 && ref != kDone) {
            test_prim<R, CpuSem>(S, ~ref, r, t_min, best);
            if (AnyHit && best.prim >= 0) return true;
            ref = stack[--sp];
        }
    }
    return best.prim >= 0;
#else
    // one node-or-leaf step per iteration; boxes are tested against the closed range [t_min, best.t]
    // (ties, and CPU-semantics rectangles that accept t == t_far)
    int stack[kStackDepth];
    int sp = 0;
    int ref = S.root;
    // rectangles kept outside the hierarchy (B2RT_LBVH_RECTS_OUTSIDE) are visited FIRST, as leaves stacked above the
    // root, so a wall hit already bounds the walk (no second inlined copy of the primitive tests: that cost 10 %)
    if (S.n_outside > 0) {
        B2RT_PUSH(S, stack, sp, S.root);
        for (int p = S.n_outside - 1; p >= 1; --p) B2RT_PUSH(S, stack, sp, ~p);
        ref = ~0;
    }
    while (true) {
        if (ref >= 0) {
            float4 n0, n1, n2, n3;
            if (ref < S.n_top) {
                const float4 *p = s_top + 4 * ref;
                n0 = p[0]; n1 = p[1]; n2 = p[2]; n3 = p[3];
            } else {
                const float4 *p = S.nodes + 4 * (size_t)(ref - S.n_top);
                n0 = __ldg(p); n1 = __ldg(p + 1); n2 = __ldg(p + 2); n3 = __ldg(p + 3);
            }
            R tl, tr;
            bool hl = slab<R>(n0.x, n0.y, n0.z, n0.w, n1.x, n1.y, r.o, id, t_min, best.t, tl);
            bool hr = slab<R>(n1.z, n1.w, n2.x, n2.y, n2.z, n2.w, r.o, id, t_min, best.t, tr);
            int cl = __float_as_int(n3.x), cr = __float_as_int(n3.y);
            if (hl && hr) {
                bool swap = tr < tl;
                B2RT_PUSH(S, stack, sp, swap ? cl : cr);
                ref = swap ? cr : cl;
            } else if (hl) {
                ref = cl;
            } else if (hr) {
                ref = cr;
            } else {
                if (sp == 0) break;
                ref = stack[--sp];
            }
        } else {
            test_prim<R, CpuSem>(S, ~ref, r, t_min, best);
            if (AnyHit && best.prim >= 0) return true;
            if (sp == 0) break;
            ref = stack[--sp];
        }
    }
    return best.prim >= 0;
#endif
}

// brute-force scan in packed order (validation path; equals cuda_scene_hit's loop structure)
template <typename R, bool CpuSem, bool AnyHit>
__device__ __forceinline__ bool scan_all(const SceneDev &S, const Ray<R> &r, R t_min, R t_max, Hit<R> &best) {
    best.t = t_max; best.prim = -1; best.a = R(0); best.b = R(0);
    for (int p = 0; p < S.n_prims; ++p) {
        test_prim<R, CpuSem>(S, p, r, t_min, best);
        if (AnyHit && best.prim >= 0) return true;
    }
    return best.prim >= 0;
}

// Small-scene scan over the planar records of b2rt_scene.d_scan_prims (float32 only) followed by the spheres.
// Every lane of a warp tests the same record (uniform shared-memory reads, no stack, no per-type branch
// inside the loop), and a parallelogram record covers two triangles: ~40 instructions per record instead
// of ~65 per triangle for the generic Moeller-Trumbore test.  Same result contract as traverse():
// closest t, ties to the lowest packed id.
// Occluder-hint codes: k in [0, n_scan) = planar scan record k, 64 + i = sphere i, 128 + p = packed primitive p
// tested with the generic per-type test (scenes without scan records).
template <bool GENERIC>
__device__ __forceinline__ bool occluder_test(const SceneDev &S, const float4 *sp, int code, const Ray<float> &r,
                                              float t_min, float t_max) {
    if (code < 0) return false;
    if (GENERIC) {
        if (code < 128) return false;
        Hit<float> h; h.t = t_max; h.prim = -1; h.a = 0.f; h.b = 0.f;
        test_prim<float, false>(S, code - 128, r, t_min, h);
        return h.prim >= 0;
    }
    if (sp == nullptr || code >= 128) return false;
    if (code >= 64) {
        float t;
        return hit_sphere<float>(S, code - 64, r, t_min, t_max, false, t);
    }
    const float4 q0 = sp[4 * code], q1 = sp[4 * code + 1], q2 = sp[4 * code + 2], q3 = sp[4 * code + 3];
    float dn = q0.x * r.d.x + q0.y * r.d.y + q0.z * r.d.z;
    float T = q0.w - (q0.x * r.o.x + q0.y * r.o.y + q0.z * r.o.z);
    float t = __fdividef(T, dn);
    float px = fmaf(t, r.d.x, r.o.x), py = fmaf(t, r.d.y, r.o.y), pz = fmaf(t, r.d.z, r.o.z);
    float u = q1.x * px + q1.y * py + q1.z * pz + q1.w;
    float v = q2.x * px + q2.y * py + q2.z * pz + q2.w;
    const int kind = __float_as_int(q3.z) >> 28;
    bool inside = u >= 0.f && v >= 0.f && (kind == 1 ? (u + v <= 1.f) : (u <= q3.x && v <= q3.y));
    return inside && fabsf(dn) > 1e-6f && t > t_min && t < t_max;
}

// Folds planar record k into the running closest hit (the loose-record loop).  The distance test comes first and
// branches: with box records most scenes keep only a few loose records and most rays miss each of them, so the
// (u, v) work runs for the few lanes that are in range.
template <bool AnyHit>
__device__ __forceinline__ bool scan_planar(const float4 *sp, int k, float ox, float oy, float oz, float dx, float dy,
                                            float dz, float t_min, Hit<float> &best) {
    const float4 q0 = sp[4 * k];
    const float dn = q0.x * dx + q0.y * dy + q0.z * dz;
    const float t = (q0.w - (q0.x * ox + q0.y * oy + q0.z * oz)) * rcp_approx(dn);
    if (!(t > t_min && t <= best.t && fabsf(dn) > 1e-6f)) return false;
    const float4 q1 = sp[4 * k + 1], q2 = sp[4 * k + 2], q3 = sp[4 * k + 3];
    const float px = fmaf(t, dx, ox), py = fmaf(t, dy, oy), pz = fmaf(t, dz, oz);
    const float u = q1.x * px + q1.y * py + q1.z * pz + q1.w;
    const float v = q2.x * px + q2.y * py + q2.z * pz + q2.w;
    const int w = __float_as_int(q3.z), kind = w >> 28;          // warp-uniform
    int id = w & 0x0fffffff;
    bool inside = u >= 0.f && v >= 0.f;
    float a = u, b = v;
    if (kind == 1) inside = inside && (u + v <= 1.f);
    else inside = inside && u <= q3.x && v <= q3.y;
    if (kind >= 2) {
        const bool first = kind == 2 ? (u >= v) : (u > v);
        id = first ? id : __float_as_int(q3.w);
        a = first ? u - v : u;
        b = first ? v : v - u;
    }
    const bool ok = inside && (t < best.t || id < best.prim);
    if (ok) { best.t = t; best.prim = id; best.a = a; best.b = b; }
    return ok;
}

// Box record j (see b2rt_scene.n_scan_boxes): three slabs in the box's own coordinates.  The candidate hit is the
// first boundary crossing beyond t_min whose face exists; returns the planar record index of that face or -1.
// The face slot (2 * axis + (l == +1)) rides in the three low mantissa bits of each crossing distance, so the
// min / max reductions that find t_enter / t_exit carry "which face" along for free (t moves by <= 7 ulp).
__device__ __forceinline__ int scan_box(const float4 *bx, int j, float ox, float oy, float oz, float dx, float dy,
                                        float dz, float t_min, float t_far, float &t_out) {
    const float4 q0 = bx[4 * j], q1 = bx[4 * j + 1], q2 = bx[4 * j + 2], q3 = bx[4 * j + 3];
    const float lo0 = fmaf(q0.x, ox, fmaf(q0.y, oy, fmaf(q0.z, oz, q0.w)));
    const float lo1 = fmaf(q1.x, ox, fmaf(q1.y, oy, fmaf(q1.z, oz, q1.w)));
    const float lo2 = fmaf(q2.x, ox, fmaf(q2.y, oy, fmaf(q2.z, oz, q2.w)));
    // + 1e-30: a direction parallel to a slab gives huge finite crossings (never inf: tagging inf would make NaN)
    const float r0 = rcp_approx(fmaf(q0.x, dx, fmaf(q0.y, dy, fmaf(q0.z, dz, 1e-30f))));
    const float r1 = rcp_approx(fmaf(q1.x, dx, fmaf(q1.y, dy, fmaf(q1.z, dz, 1e-30f))));
    const float r2 = rcp_approx(fmaf(q2.x, dx, fmaf(q2.y, dy, fmaf(q2.z, dz, 1e-30f))));
    auto tag = [](float t, unsigned slot) { return __uint_as_float((__float_as_uint(t) & 0xfffffff8u) | slot); };
    // crossings of the faces l_k = -1 / +1: (-+1 - lo) * r as ONE fma each (single rounding: exact to an ulp even
    // for a ray that starts on a face, where lo -> +-1 cancels)
    const float a0 = tag(fmaf(-lo0, r0, -r0), 0u), b0 = tag(fmaf(-lo0, r0, r0), 1u);
    const float a1 = tag(fmaf(-lo1, r1, -r1), 2u), b1 = tag(fmaf(-lo1, r1, r1), 3u);
    const float a2 = tag(fmaf(-lo2, r2, -r2), 4u), b2 = tag(fmaf(-lo2, r2, r2), 5u);
    const float te = fmaxf(fmaxf(fminf(a0, b0), fminf(a1, b1)), fminf(a2, b2));
    const float tx = fminf(fminf(fmaxf(a0, b0), fmaxf(a1, b1)), fmaxf(a2, b2));
    // bytes 6 and 7 of (w0, w1) are zero: selector nibbles 7 clear the upper bytes of the result
    const unsigned w0 = __float_as_uint(q3.x), w1 = __float_as_uint(q3.y);
    const int ce = (int)__byte_perm(w0, w1, (__float_as_uint(te) & 7u) | 0x7770u);
    const int cx = (int)__byte_perm(w0, w1, (__float_as_uint(tx) & 7u) | 0x7770u);
    const bool use_e = te > t_min && ce != 255;
    const float t = use_e ? te : tx;
    const int c = use_e ? ce : cx;
    t_out = t;
    return (te <= tx && t > t_min && t < t_far && c != 255) ? c : -1;
}

// CLOSED box record (all six faces exist: a cube; flag bit 0 of the record's third info word): the hit is t_enter if
// that lies beyond t_min, else t_exit — no face lookup is needed to decide, so the face tags, the two byte
// permutes and the `face exists` tests of scan_box leave the loop (74 -> ~45 instructions per box); the winning
// box's face is recovered once, after the loop, from the hit point (box_face_at).  Same crossing distances as
// scan_box bit for bit, minus the <= 7 ulp of the tags: fma(-lo, r, -+|r|) IS min / max(fma(-lo, r, -r), fma(-lo, r, r)).
#ifndef B2RT_OPT_CLOSED_BOX
#define B2RT_OPT_CLOSED_BOX 1
#endif
__device__ __forceinline__ bool scan_box_closed(const float4 *bx, int j, float ox, float oy, float oz, float dx, float dy,
                                                float dz, float t_min, float t_far, float &t_out) {
    const float4 q0 = bx[4 * j], q1 = bx[4 * j + 1], q2 = bx[4 * j + 2];
    const float lo0 = fmaf(q0.x, ox, fmaf(q0.y, oy, fmaf(q0.z, oz, q0.w)));
    const float lo1 = fmaf(q1.x, ox, fmaf(q1.y, oy, fmaf(q1.z, oz, q1.w)));
    const float lo2 = fmaf(q2.x, ox, fmaf(q2.y, oy, fmaf(q2.z, oz, q2.w)));
    const float r0 = rcp_approx(fmaf(q0.x, dx, fmaf(q0.y, dy, fmaf(q0.z, dz, 1e-30f))));
    const float r1 = rcp_approx(fmaf(q1.x, dx, fmaf(q1.y, dy, fmaf(q1.z, dz, 1e-30f))));
    const float r2 = rcp_approx(fmaf(q2.x, dx, fmaf(q2.y, dy, fmaf(q2.z, dz, 1e-30f))));
    // FFMA takes |r| as an operand modifier: the near / far crossing of each slab without a min / max pair
    const float te = fmaxf(fmaxf(fmaf(-lo0, r0, -fabsf(r0)), fmaf(-lo1, r1, -fabsf(r1))), fmaf(-lo2, r2, -fabsf(r2)));
    const float tx = fminf(fminf(fmaf(-lo0, r0, fabsf(r0)), fmaf(-lo1, r1, fabsf(r1))), fmaf(-lo2, r2, fabsf(r2)));
    const float t = te > t_min ? te : tx;
    t_out = t;
    return te <= tx && t > t_min && t < t_far;
}

// planar record index of the face of box j that the point o + t d lies on: the axis whose box coordinate is
// closest to +-1 (at an edge either adjacent face is a correct answer: documented tie)
__device__ __forceinline__ int box_face_at(const float4 *bx, int j, float ox, float oy, float oz, float dx, float dy,
                                           float dz, float t) {
    const float4 q0 = bx[4 * j], q1 = bx[4 * j + 1], q2 = bx[4 * j + 2], q3 = bx[4 * j + 3];
    const float px = fmaf(t, dx, ox), py = fmaf(t, dy, oy), pz = fmaf(t, dz, oz);
    const float l0 = fmaf(q0.x, px, fmaf(q0.y, py, fmaf(q0.z, pz, q0.w)));
    const float l1 = fmaf(q1.x, px, fmaf(q1.y, py, fmaf(q1.z, pz, q1.w)));
    const float l2 = fmaf(q2.x, px, fmaf(q2.y, py, fmaf(q2.z, pz, q2.w)));
    const float a0 = fabsf(l0), a1 = fabsf(l1), a2 = fabsf(l2);
    unsigned slot = l0 > 0.f ? 1u : 0u;
    float am = a0;
    if (a1 > am) { am = a1; slot = 2u | (l1 > 0.f ? 1u : 0u); }
    if (a2 > am) { slot = 4u | (l2 > 0.f ? 1u : 0u); }
    return (int)__byte_perm(__float_as_uint(q3.x), __float_as_uint(q3.y), slot | 0x7770u);
}

// (u, v) -> triangle id / barycentrics of the hit on planar record k at distance best.t (box faces)
__device__ __forceinline__ void scan_resolve_face(const float4 *sp, int k, float ox, float oy, float oz, float dx,
                                                  float dy, float dz, Hit<float> &best) {
    const float4 q1 = sp[4 * k + 1], q2 = sp[4 * k + 2], q3 = sp[4 * k + 3];
    const float t = best.t;
    float px = fmaf(t, dx, ox), py = fmaf(t, dy, oy), pz = fmaf(t, dz, oz);
    float u = q1.x * px + q1.y * py + q1.z * pz + q1.w;
    float v = q2.x * px + q2.y * py + q2.z * pz + q2.w;
    const int w = __float_as_int(q3.z), kind = w >> 28;
    int id = w & 0x0fffffff;
    if (kind >= 2) {
        bool first = kind == 2 ? (u >= v) : (u > v);
        id = first ? id : __float_as_int(q3.w);
        float a = first ? u - v : u, b = first ? v : v - u;
        u = a; v = b;
    }
    best.prim = id; best.a = u; best.b = v;
}

// slab test of the whole scene (camera rays start outside it)
__device__ __forceinline__ bool misses_scene(const SceneDev &S, const Ray<float> &r) {
    if (S.blo[0] > S.bhi[0]) return false;
    const float ix = rcp_approx(r.d.x), iy = rcp_approx(r.d.y), iz = rcp_approx(r.d.z);
    const float x0 = (S.blo[0] - r.o.x) * ix, x1 = (S.bhi[0] - r.o.x) * ix;
    const float y0 = (S.blo[1] - r.o.y) * iy, y1 = (S.bhi[1] - r.o.y) * iy;
    const float z0 = (S.blo[2] - r.o.z) * iz, z1 = (S.bhi[2] - r.o.z) * iz;
    const float tn = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fmaxf(fminf(z0, z1), 0.f));
    const float tf = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fmaxf(z0, z1));
    return tn > tf;
}

// MASKED: `mask` (warp-uniform) says which records can be hit at all — bit j box j, bit n_box + k loose planar record k,
// bit n_box + n_loose + i sphere i (camera rays: per-tile candidate masks, rt_path.cuh:primary_mask_kernel).
// CLOSED_PATH: closed boxes take scan_box_closed (a second inlined box test: left out of the masked camera-ray scan,
// whose kernel also carries the ray generation and sits closest to the instruction-cache limit).
template <bool AnyHit, bool MASKED = false, bool CLOSED_PATH = !MASKED>
__device__ __forceinline__ bool scan_small(const SceneDev &S, const float4 *sp, const Ray<float> &r, float t_min,
                                           float t_max, Hit<float> &best, unsigned mask = 0xffffffffu) {
    best.t = t_max; best.prim = -1; best.a = 0.f; best.b = 0.f;
    const float ox = r.o.x, oy = r.o.y, oz = r.o.z, dx = r.d.x, dy = r.d.y, dz = r.d.z;
    if (S.n_box > 0) {
        const float4 *bx = sp + 4 * S.n_scan;
        int face = -1;                                           // planar record of the winning face, or 256 + j: closed box j
        for (int j = 0; j < S.n_box; ++j) {
            float t;
            if (MASKED && !((mask >> j) & 1u)) continue;
            if (B2RT_OPT_CLOSED_BOX && CLOSED_PATH && (__float_as_uint(bx[4 * j + 3].z) & 1u)) {        // warp-uniform
                if (scan_box_closed(bx, j, ox, oy, oz, dx, dy, dz, t_min, best.t, t)) {
                    best.t = t; face = 256 + j;
                    if (AnyHit) { best.prim = 0; return true; }
                }
                continue;
            }
            int c = scan_box(bx, j, ox, oy, oz, dx, dy, dz, t_min, best.t, t);
            if (c >= 0) {
                best.t = t; face = c;
                if (AnyHit) { best.prim = 0; return true; }
            }
        }
        if (face >= 256) face = box_face_at(bx, face - 256, ox, oy, oz, dx, dy, dz, best.t);
        if (face >= 0) scan_resolve_face(sp, face, ox, oy, oz, dx, dy, dz, best);
    }
    if (MASKED) mask >>= S.n_box;
#pragma unroll kScanUnroll
    for (int k = 0; k < S.n_loose; ++k) {
        if (MASKED && !((mask >> k) & 1u)) continue;
        bool ok = scan_planar<AnyHit>(sp, k, ox, oy, oz, dx, dy, dz, t_min, best);
        if (AnyHit && ok) return true;
    }
#ifndef B2RT_OPT_SPH
#define B2RT_OPT_SPH 1             // 0: per-sphere hit_sphere() calls (measurement switch)
#endif
    if (!B2RT_OPT_SPH) {
        for (int i = 0; i < S.n_sphere; ++i) {
            int prim = S.n_rect + i;
            float t;
            bool allow_eq = best.prim >= 0 && prim < best.prim;
            if (hit_sphere<float>(S, i, r, t_min, best.t, allow_eq, t)) {
                best.t = t; best.prim = prim; best.a = 0.f; best.b = 0.f;
                if (AnyHit) return true;
            }
        }
        return best.prim >= 0;
    }
    // spheres: the cancellation-free form of hit_sphere with 1 / (d.d) hoisted out of the loop
    if (MASKED) { mask >>= S.n_loose; if (mask == 0u) return best.prim >= 0; }
    const float ia = rcp_approx(dx * dx + dy * dy + dz * dz);
    const float4 *sph = reinterpret_cast<const float4 *>(S.sphere);
    for (int i = 0; i < S.n_sphere; ++i) {
        if (MASKED && !((mask >> i) & 1u)) continue;
        const float4 s0 = __ldg(sph + 2 * i);
        const float cx = ox - s0.x, cy = oy - s0.y, cz = oz - s0.z;
        const float b = (cx * dx + cy * dy + cz * dz) * ia;              // b / a
        const float px = fmaf(-b, dx, cx), py = fmaf(-b, dy, cy), pz = fmaf(-b, dz, cz);
        const float disc = fmaf(s0.w, s0.w, -(px * px + py * py + pz * pz));     // r^2 - |centre to line|^2
        if (disc > 0.f) {
            const float sq = sqrt_(disc * ia);
            const float t1 = -b - sq, t2 = -b + sq;
            const float t = t_min < t1 ? t1 : t2;
            const int prim = S.n_rect + i;
            if (t_min < t && (t < best.t || (t == best.t && prim < best.prim))) {
                best.t = t; best.prim = prim; best.a = 0.f; best.b = 0.f;
                if (AnyHit) return true;
            }
        }
    }
    return best.prim >= 0;
}

inline size_t smem_scan_bytes(const SceneDev &S) { return (size_t)(S.n_scan + S.n_box) * 64; }
inline size_t smem_surf_bytes(const SceneDev &S) { return S.surf ? (size_t)S.n_prims * 80 : 0; }

__device__ __forceinline__ void stage_surf(const SceneDev &S, float4 *s_surf) {
    for (int i = threadIdx.x; i < 5 * S.n_prims; i += blockDim.x) s_surf[i] = __ldg(S.surf + i);
    __syncthreads();
}

__device__ __forceinline__ void stage_scan(const SceneDev &S, float4 *s_scan) {
    for (int i = threadIdx.x; i < 4 * (S.n_scan + S.n_box); i += blockDim.x) s_scan[i] = __ldg(S.scan + i);
    __syncthreads();
}

// cooperative copy of the BVH top levels into shared memory (call from every thread of the CTA)
__device__ __forceinline__ void stage_top(const SceneDev &S, float4 *s_top) {
    for (int i = threadIdx.x; i < 4 * S.n_top; i += blockDim.x) s_top[i] = __ldg(S.top + i);
    __syncthreads();
}

// ---------------------------------------------------------------------------------- surface
template <typename R> struct Surface {
    V3<R> p, n;            // hit point, shading normal as the reference defines it per primitive type
    R u, v;                // texture coordinates
    V3<R> color;           // material colour (before texturing)
    R diffuse, specular, reflective, refractive, ior;
    int tex;               // texture id or -1
};

// Rebuilds what cuda_scene_hit returns beside t (point, normal, uv, material; :568-574,:616-631,:706-728)
template <typename R, bool CpuSem>
__device__ __forceinline__ void make_surface(const SceneDev &S, const Ray<R> &r, const Hit<R> &h, Surface<R> &sf) {
    sf.p = r.o + r.d * h.t;
    const real4<R> *sh = reinterpret_cast<const real4<R> *>(S.shade) + 3 * h.prim;
    if (h.prim < S.n_rect) {
        const real4<R> *q = reinterpret_cast<const real4<R> *>(S.rect) + 4 * h.prim;
        real4<R> r0 = ldg4(q), r1 = ldg4(q + 1);
        sf.n = xyz<R>(r1);
        sf.u = div_(h.a, r0.w); sf.v = div_(h.b, r1.w);
    } else if (h.prim < S.n_rect + S.n_sphere) {
        const real4<R> *q = reinterpret_cast<const real4<R> *>(S.sphere) + 2 * (h.prim - S.n_rect);
        real4<R> s0 = ldg4(q);
        sf.n = div3(sf.p - xyz<R>(s0), s0.w);
        sf.u = R(0); sf.v = R(0);
    } else {
        real4<R> nn = ldg4(sh), uva = ldg4(sh + 1), uvb = ldg4(sh + 2);
        V3<R> n = xyz<R>(nn);
        R dp = dot(n, r.d);
        bool flip = CpuSem ? !(dp < R(0)) : (dp > R(0));
        sf.n = flip ? -n : n;
        if (uvb.z != R(0)) {
            R w = R(1) - h.a - h.b;
            if (CpuSem) {      // u*uv1 + v*uv2 + w*uv0   (core/geometry.py:166-168)
                sf.u = h.a * uva.z + h.b * uvb.x + w * uva.x;
                sf.v = h.a * uva.w + h.b * uvb.y + w * uva.y;
            } else {           // w*uv0 + u*uv1 + v*uv2   (cuda_path_tracer.py:722-724)
                sf.u = w * uva.x + h.a * uva.z + h.b * uvb.x;
                sf.v = w * uva.y + h.a * uva.w + h.b * uvb.y;
            }
        } else { sf.u = R(0); sf.v = R(0); }
    }
    int m = __ldg(S.prim_mat + h.prim);
    const real4<R> *mq = reinterpret_cast<const real4<R> *>(S.mat) + 2 * m;
    real4<R> m0 = ldg4(mq), m1 = ldg4(mq + 1);
    sf.color = xyz<R>(m0); sf.diffuse = m0.w;
    sf.specular = m1.x; sf.reflective = m1.y; sf.refractive = m1.z; sf.ior = m1.w;
    sf.tex = __ldg(S.mat_tex + m);
}

// make_surface from the shared-memory surface records (b2rt_scene.d_surface_records): no per-type branch, one hop
__device__ __forceinline__ void make_surface_small(const float4 *s_surf, const Ray<float> &r, const Hit<float> &h,
                                                   Surface<float> &sf) {
    const float4 *q = s_surf + 5 * h.prim;
    const float4 s0 = q[0], s1 = q[1], s2 = q[2], s3 = q[3], s4 = q[4];
    sf.p = r.o + r.d * h.t;
    const bool sph = s0.w != 0.f;
    float nx = sph ? (sf.p.x - s0.x) * s0.w : s0.x, ny = sph ? (sf.p.y - s0.y) * s0.w : s0.y,
          nz = sph ? (sf.p.z - s0.z) * s0.w : s0.z;
    const bool flip = (__float_as_int(s4.w) & 1) && (nx * r.d.x + ny * r.d.y + nz * r.d.z > 0.f);
    sf.n = flip ? V3<float>{-nx, -ny, -nz} : V3<float>{nx, ny, nz};
    sf.u = fmaf(h.b, s3.z, fmaf(h.a, s3.y, s3.x));
    sf.v = fmaf(h.b, s4.z, fmaf(h.a, s4.y, s4.x));
    sf.color = {s1.x, s1.y, s1.z}; sf.diffuse = s1.w;
    sf.specular = s2.x; sf.reflective = s2.y; sf.refractive = s2.z; sf.ior = s2.w;
    sf.tex = __float_as_int(s3.w);
}

// nearest-texel fetch with V flip: cuda_sample_texture (cuda_path_tracer.py:473-493) /
// Texture.sample (core/material.py:13-21).  RGBX8 texel -> one 32-bit load.
template <typename R> __device__ __forceinline__ V3<R> decode_texel(uint32_t px) {
    if constexpr (sizeof(R) == 4) {          // one multiply instead of an IEEE division per channel
        const float k = 1.0f / 255.0f;
        return V3<R>{R(px & 255u) * k, R((px >> 8) & 255u) * k, R((px >> 16) & 255u) * k};
    }
    return V3<R>{R(px & 255u) / R(255), R((px >> 8) & 255u) / R(255), R((px >> 16) & 255u) / R(255)};
}

template <typename R, bool CpuSem>
__device__ __forceinline__ uint32_t fetch_texel(const SceneDev &S, int tex, R u, R v) {
    int4 info = __ldg(S.tex_info + tex);       // offset, w, h
    int w = info.y, h = info.z;
    int iu, iv;
    if (CpuSem) {
        iu = (int)max_(R(0), min_(R(w - 1), u * R(w - 1)));
        iv = (int)max_(R(0), min_(R(h - 1), (R(1) - v) * R(h - 1)));
    } else {
        u = max_(R(0), min_(R(1), u));
        v = max_(R(0), min_(R(1), v));
        iu = (int)(u * R(w - 1));
        iv = (int)((R(1) - v) * R(h - 1));
        iu = max(0, min(w - 1, iu));
        iv = max(0, min(h - 1, iv));
    }
    return __ldg(S.texels + (size_t)info.x + (size_t)iv * w + iu);
}

template <typename R, bool CpuSem>
__device__ __forceinline__ V3<R> sample_texture(const SceneDev &S, int tex, R u, R v) {
    return decode_texel<R>(fetch_texel<R, CpuSem>(S, tex, u, v));
}

template <typename R, bool CpuSem>
__device__ __forceinline__ V3<R> base_color(const SceneDev &S, const Surface<R> &sf) {
    if (sf.tex >= 0 && sf.tex < S.n_tex) return sample_texture<R, CpuSem>(S, sf.tex, sf.u, sf.v);
    return sf.color;
}

// pinhole ray: cuda_get_ray (cuda_path_tracer.py:84-112) == Camera.get_ray + Ray() (core/camera.py:26-31)
template <typename R>
__device__ __forceinline__ Ray<R> camera_ray(const Cam<R> &c, R u, R v) {
    Ray<R> r;
    r.o = c.origin;
    V3<R> d = {c.llc.x + u * c.hor.x + v * c.ver.x - c.origin.x,
               c.llc.y + u * c.hor.y + v * c.ver.y - c.origin.y,
               c.llc.z + u * c.hor.z + v * c.ver.z - c.origin.z};
    R l = length(d);
    if (l > R(0)) d = div3(d, l);
    r.d = d;
    return r;
}

// Snell refraction: cuda_refract_path (cuda_path_tracer.py:115-131)
template <typename R>
__device__ __forceinline__ bool refract_nb(V3<R> in, V3<R> n, R eta, V3<R> &out) {
    R cos_i = -(in.x * n.x + in.y * n.y + in.z * n.z);
    R sin2_t = eta * eta * (R(1) - cos_i * cos_i);
    if (sin2_t > R(1)) return false;
    R cos_t = sqrt_(R(1) - sin2_t);
    R f2 = eta * cos_i - cos_t;
    out = {eta * in.x + f2 * n.x, eta * in.y + f2 * n.y, eta * in.z + f2 * n.z};
    return true;
}

}  // namespace b2rt
