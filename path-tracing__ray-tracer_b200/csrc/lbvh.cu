// lbvh.cu — GPU LBVH build: primitive bounds -> 30-bit Morton codes -> radix sort -> Karras (2012)
// hierarchy -> bottom-up refit with atomic arrival flags -> breadth-first copy of the top levels.
//
// Replaces the reference's recursive random-axis median split (BVHNode.__init__,
// core/acceleration.py:8-30; per-primitive boxes core/geometry.py:40-48,83,131-137;
// AABB.surrounding_box core/math.py:90-102).  Closest-hit results do not depend on the hierarchy
// (rt_scene.cuh applies a fixed tie rule), so parity is preserved while the build becomes O(n) sorts
// and scans.  Leaves hold exactly one primitive (like the reference's leaves).
#include <cub/cub.cuh>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "lbvh.h"

namespace b2rt {

namespace {

__device__ __forceinline__ int float_to_ordered(float f) {
    int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float ordered_to_float(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

struct Box {
    float3 lo, hi;
};

__device__ __forceinline__ void grow(Box &b, float x, float y, float z) {
    b.lo.x = fminf(b.lo.x, x); b.lo.y = fminf(b.lo.y, y); b.lo.z = fminf(b.lo.z, z);
    b.hi.x = fmaxf(b.hi.x, x); b.hi.y = fmaxf(b.hi.y, y); b.hi.z = fmaxf(b.hi.z, z);
}

// boxes of packed primitive i from the float32 hot streams (layouts in include/b200rt.h)
__device__ Box prim_box(int i, int n_rect, int n_sphere, const float4 *rect, const float4 *sphere, const float4 *tri,
                        float pad) {
    Box b;
    b.lo = make_float3(3.0e38f, 3.0e38f, 3.0e38f);
    b.hi = make_float3(-3.0e38f, -3.0e38f, -3.0e38f);
    if (i < n_rect) {
        float4 r0 = rect[4 * i], r1 = rect[4 * i + 1], r2 = rect[4 * i + 2], r3 = rect[4 * i + 3];
        float ux = r2.x * r0.w, uy = r2.y * r0.w, uz = r2.z * r0.w;
        float vx = r3.x * r1.w, vy = r3.y * r1.w, vz = r3.z * r1.w;
        grow(b, r0.x, r0.y, r0.z);
        grow(b, r0.x + ux, r0.y + uy, r0.z + uz);
        grow(b, r0.x + vx, r0.y + vy, r0.z + vz);
        grow(b, r0.x + ux + vx, r0.y + uy + vy, r0.z + uz + vz);
    } else if (i < n_rect + n_sphere) {
        float4 s = sphere[2 * (i - n_rect)];
        grow(b, s.x - s.w, s.y - s.w, s.z - s.w);
        grow(b, s.x + s.w, s.y + s.w, s.z + s.w);
    } else {
        int k = i - n_rect - n_sphere;
        float4 v0 = tri[3 * k], e1 = tri[3 * k + 1], e2 = tri[3 * k + 2];
        grow(b, v0.x, v0.y, v0.z);
        grow(b, v0.x + e1.x, v0.y + e1.y, v0.z + e1.z);
        grow(b, v0.x + e2.x, v0.y + e2.y, v0.z + e2.z);
    }
    // outward pad: absolute pad plus a relative term for large coordinates
    float ax = fmaxf(fabsf(b.lo.x), fabsf(b.hi.x)), ay = fmaxf(fabsf(b.lo.y), fabsf(b.hi.y)),
          az = fmaxf(fabsf(b.lo.z), fabsf(b.hi.z));
    float px = pad + 4e-7f * ax, py = pad + 4e-7f * ay, pz = pad + 4e-7f * az;
    b.lo.x -= px; b.lo.y -= py; b.lo.z -= pz;
    b.hi.x += px; b.hi.y += py; b.hi.z += pz;
    return b;
}

__global__ void bounds_kernel(int n, int n_rect, int n_sphere, const float4 *rect, const float4 *sphere,
                              const float4 *tri, float pad, float4 *box_lo, float4 *box_hi, int *scene_bounds,
                              double *moments, int n_outside) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    float c[6] = {3.0e38f, 3.0e38f, 3.0e38f, -3.0e38f, -3.0e38f, -3.0e38f};   // centroid bounds
    float ext[3] = {0.f, 0.f, 0.f};                          // box extent of this primitive
    const bool outside = i < n_outside;                      // kept out of the hierarchy (see lbvh_build)
    if (outside) {
        // a far-away point box no ray reaches; its 31-bit key sorts it behind every real primitive, so the root
        // separates the "outside" leaves from the real tree and no real node's box is widened by them
        box_lo[i] = box_hi[i] = make_float4(3.0e37f, 3.0e37f, 3.0e37f, 0.f);
    } else if (i < n) {
        Box b = prim_box(i, n_rect, n_sphere, rect, sphere, tri, pad);
        box_lo[i] = make_float4(b.lo.x, b.lo.y, b.lo.z, 0.f);
        box_hi[i] = make_float4(b.hi.x, b.hi.y, b.hi.z, 0.f);
        c[0] = c[3] = 0.5f * (b.lo.x + b.hi.x);
        c[1] = c[4] = 0.5f * (b.lo.y + b.hi.y);
        c[2] = c[5] = 0.5f * (b.lo.z + b.hi.z);
        ext[0] = b.hi.x - b.lo.x; ext[1] = b.hi.y - b.lo.y; ext[2] = b.hi.z - b.lo.z;
    }
    typedef cub::BlockReduce<float, 256> BR;
    __shared__ typename BR::TempStorage tmp;
    for (int k = 0; k < 6; ++k) {
        float r = k < 3 ? BR(tmp).Reduce(c[k], cub::Min()) : BR(tmp).Reduce(c[k], cub::Max());
        __syncthreads();
        if (threadIdx.x == 0) {
            if (k < 3) atomicMin(scene_bounds + k, float_to_ordered(r));
            else atomicMax(scene_bounds + k, float_to_ordered(r));
        }
    }
    // first and second moments of the centroids and the summed box extents: the Morton bit allocation follows the
    // number of primitives that fit across the scene along each axis
    for (int k = 0; k < 9; ++k) {
        float v = (i < n && !outside) ? (k < 3 ? c[k] : k < 6 ? c[k - 3] * c[k - 3] : ext[k - 6]) : 0.f;
        float r = BR(tmp).Sum(v);
        __syncthreads();
        if (threadIdx.x == 0) atomicAdd(moments + k, (double)r);
    }
}

// Morton bits per axis (sum 30) and the world-space spread the cells are measured in.
// Two facts bound the useful splits along an axis: (1) cells should stay roughly CUBIC in world space (the surface-area
// heuristic), so the next bit goes to the axis whose cell is currently the longest (spread_k / 2^bits_k, spread = 4 sigma
// of the centroids); (2) key bits finer than the primitives cannot separate them any more, so an axis stops at
// cap_k = log2(spread_k / mean primitive extent_k) + 1 bits.  Equal spreads and extents give the classic 10/10/10.
// On 2.5-D data (a terrain, a city) the classic x,y,z cycle spends every third split on the flat axis and cuts the mesh
// into bands whose boxes overlap in the other two axes.  Measured on the 1 M-triangle height field (rough: triangles
// of 0.03 x 0.85 x 0.06), ms per step with three rotation sweeps: 10/10/10 194.9, 11/8/11 110.1, 12/7/11 98.2,
// 14/4/12 97.0, 12/6/12 90.5, 13/4/13 88.4; this rule: 91.0.
__device__ __forceinline__ void morton_bits(int n, const double *moments, int *bits, float *spread) {
    float cap[3];
    bool any = false;
    for (int k = 0; k < 3; ++k) {
        const double m = moments[k] / n, v = moments[3 + k] / n - m * m;
        const float ext = (float)(moments[6 + k] / n);
        spread[k] = 4.f * sqrtf(fmaxf((float)v, 0.f));
        cap[k] = spread[k] > 0.f ? fminf(fmaxf(log2f(spread[k] / fmaxf(ext, 1e-30f)) + 1.f, 0.f), 20.f) : 0.f;
        any = any || cap[k] >= 1.f;
        bits[k] = 0;
    }
    if (!any) { bits[0] = bits[1] = bits[2] = 10; spread[0] = spread[1] = spread[2] = 1.f; return; }
    for (int t = 0; t < 30; ++t) {
        int a = -1;
        float best = -1.f;
        for (int k = 0; k < 3; ++k) {                        // longest cell among the axes that still have room
            const float cell = spread[k] / (float)(1u << bits[k]);
            if ((float)bits[k] + 1.f <= cap[k] && cell > best) { best = cell; a = k; }
        }
        if (a < 0) {                                         // every axis at its cap: the remaining bits only break ties
            best = -1e30f;                                   // between primitives; they go where the cap is exceeded least
            for (int k = 0; k < 3; ++k)
                if (bits[k] < 20 && cap[k] - (float)bits[k] > best) { best = cap[k] - (float)bits[k]; a = k; }
        }
        ++bits[a];
    }
}

__device__ __forceinline__ uint32_t expand10(uint32_t v) {
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}

__global__ void morton_kernel(int n, const float4 *box_lo, const float4 *box_hi, const int *scene_bounds,
                              uint32_t *keys, int *vals, const double *moments, int bx, int by, int bz, int n_outside) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (i < n_outside) { keys[i] = 0x7fffffffu; vals[i] = i; return; }
    float lo[3], ext[3];
    for (int k = 0; k < 3; ++k) {
        lo[k] = ordered_to_float(scene_bounds[k]);
        ext[k] = fmaxf(ordered_to_float(scene_bounds[3 + k]) - lo[k], 1e-20f);
    }
    float4 a = box_lo[i], b = box_hi[i];
    float cx = (0.5f * (a.x + b.x) - lo[0]) / ext[0];
    float cy = (0.5f * (a.y + b.y) - lo[1]) / ext[1];
    float cz = (0.5f * (a.z + b.z) - lo[2]) / ext[2];
    float spread[3] = {1.f, 1.f, 1.f};
    if (bx <= 0) {                                           // automatic allocation (morton_bits)
        int bits[3];
        morton_bits(n - n_outside, moments, bits, spread);
        bx = bits[0]; by = bits[1]; bz = bits[2];
    }
    if (bx == 10 && by == 10 && bz == 10) {
        uint32_t qx = (uint32_t)fminf(fmaxf(cx * 1024.f, 0.f), 1023.f);
        uint32_t qy = (uint32_t)fminf(fmaxf(cy * 1024.f, 0.f), 1023.f);
        uint32_t qz = (uint32_t)fminf(fmaxf(cz * 1024.f, 0.f), 1023.f);
        keys[i] = (expand10(qx) << 2) | (expand10(qy) << 1) | expand10(qz);
    } else {
        // uneven bit allocation (bx + by + bz <= 30): interleave from the top, always taking the next bit of the
        // axis that has the most bits left (ties x, y, z), so every axis reaches its last bit at the bottom
        // (taking it from the axis whose world-space cell is currently the longest instead measured 118 vs 91 ms)
        int rem[3] = {bx, by, bz};
        const float c[3] = {cx, cy, cz};
        uint32_t q[3];
        for (int k = 0; k < 3; ++k) {
            float s = (float)(1u << rem[k]);
            q[k] = (uint32_t)fminf(fmaxf(c[k] * s, 0.f), s - 1.f);
        }
        uint32_t key = 0;
        for (int t = bx + by + bz; t > 0; --t) {
            int a = rem[0] >= rem[1] ? (rem[0] >= rem[2] ? 0 : 2) : (rem[1] >= rem[2] ? 1 : 2);
            --rem[a];
            key = (key << 1) | ((q[a] >> rem[a]) & 1u);
        }
        keys[i] = key;
    }
    vals[i] = i;
}

// common-prefix length of sorted keys i and j (Karras 2012, with the index as tie-break bits)
__device__ __forceinline__ int delta(const uint32_t *keys, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    uint32_t a = keys[i], b = keys[j];
    if (a == b) return 32 + __clz((uint32_t)i ^ (uint32_t)j);
    return __clz(a ^ b);
}

__global__ void hierarchy_kernel(int n, const uint32_t *keys, const int *vals, int2 *children, int *parent_node,
                                 int *parent_leaf, int *pos_of) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) pos_of[vals[i]] = i;                          // primitive id -> sorted position (parent_leaf is indexed by it)
    if (i >= n - 1) return;
    int d = (delta(keys, n, i, i + 1) - delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
    int dmin = delta(keys, n, i, i - d);
    int lmax = 2;
    while (delta(keys, n, i, i + lmax * d) > dmin) lmax <<= 1;
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1)
        if (delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
    int j = i + l * d;
    int dnode = delta(keys, n, i, j);
    int s = 0, t = l;
    do {
        t = (t + 1) >> 1;
        if (delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
    } while (t > 1);
    int gamma = i + s * d + min(d, 0);
    int lo = min(i, j), hi = max(i, j);
    int2 c;
    if (lo == gamma) { c.x = ~vals[gamma]; parent_leaf[gamma] = i; } else { c.x = gamma; parent_node[gamma] = i; }
    if (hi == gamma + 1) { c.y = ~vals[gamma + 1]; parent_leaf[gamma + 1] = i; } else { c.y = gamma + 1; parent_node[gamma + 1] = i; }
    children[i] = c;
    if (i == 0) parent_node[0] = -1;
}

__device__ __forceinline__ float box_area(const float4 &lo, const float4 &hi) {
    float dx = hi.x - lo.x, dy = hi.y - lo.y, dz = hi.z - lo.z;
    return dx * dy + dy * dz + dz * dx;
}
__device__ __forceinline__ void box_union(const float4 &alo, const float4 &ahi, const float4 &blo, const float4 &bhi,
                                          float4 &lo, float4 &hi) {
    lo = make_float4(fminf(alo.x, blo.x), fminf(alo.y, blo.y), fminf(alo.z, blo.z), 0.f);
    hi = make_float4(fmaxf(ahi.x, bhi.x), fmaxf(ahi.y, bhi.y), fmaxf(ahi.z, bhi.z), 0.f);
}
__device__ __forceinline__ void write_node(float4 *nodes, int i, const float4 &llo, const float4 &lhi, const float4 &rlo,
                                           const float4 &rhi, int cl, int cr) {
    nodes[4 * (size_t)i + 0] = make_float4(llo.x, llo.y, llo.z, lhi.x);
    nodes[4 * (size_t)i + 1] = make_float4(lhi.y, lhi.z, rlo.x, rlo.y);
    nodes[4 * (size_t)i + 2] = make_float4(rlo.z, rhi.x, rhi.y, rhi.z);
    nodes[4 * (size_t)i + 3] = make_float4(__int_as_float(cl), __int_as_float(cr), 0.f, 0.f);
}
// child boxes and references of the (already finished) internal node i, read past the L1
__device__ __forceinline__ void read_node(const float4 *nodes, int i, float4 &llo, float4 &lhi, float4 &rlo, float4 &rhi,
                                          int &cl, int &cr) {
    float4 n0 = __ldcg(nodes + 4 * (size_t)i), n1 = __ldcg(nodes + 4 * (size_t)i + 1),
           n2 = __ldcg(nodes + 4 * (size_t)i + 2), n3 = __ldcg(nodes + 4 * (size_t)i + 3);
    llo = make_float4(n0.x, n0.y, n0.z, 0.f); lhi = make_float4(n0.w, n1.x, n1.y, 0.f);
    rlo = make_float4(n1.z, n1.w, n2.x, 0.f); rhi = make_float4(n2.y, n2.z, n2.w, 0.f);
    cl = __float_as_int(n3.x); cr = __float_as_int(n3.y);
}

// Bottom-up refit.  One thread per leaf climbs; the second arrival at a node (atomic flag) owns it.
// With `rotate`, the owner also tries the four tree rotations of Kensler (2008) — exchange one child with a grandchild
// on the other side — and applies the one that shrinks the surface area of the re-formed child the most.  Both
// subtrees are finished and the parent is still waiting on its flag, so the owner is the only thread touching these
// nodes: one pass, no locks.  (The Morton hierarchy splits space blindly; rotations repair the worst overlaps.)
__global__ void refit_kernel(int n, int2 *children, int *parent_node, int *parent_leaf, const int *pos_of,
                             const float4 *box_lo, const float4 *box_hi, int *flags, float4 *node_lo, float4 *node_hi,
                             float4 *nodes, int rotate) {
    // a moved subtree gets its new parent recorded, so that a further sweep can climb the rotated tree
    auto set_parent = [&](int ref, int parent) {
        if (ref >= 0) parent_node[ref] = parent; else parent_leaf[pos_of[~ref]] = parent;
    };
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    int cur = parent_leaf[k];
    while (cur >= 0) {
        __threadfence();
        if (atomicAdd(flags + cur, 1) == 0) return;
        int2 c = children[cur];
        float4 llo, lhi, rlo, rhi;
        if (c.x < 0) { llo = box_lo[~c.x]; lhi = box_hi[~c.x]; }
        else { llo = __ldcg(node_lo + c.x); lhi = __ldcg(node_hi + c.x); }
        if (c.y < 0) { rlo = box_lo[~c.y]; rhi = box_hi[~c.y]; }
        else { rlo = __ldcg(node_lo + c.y); rhi = __ldcg(node_hi + c.y); }
        if (rotate) {
            float best = 0.f;
            int which = -1;                                   // 0: L<->RL  1: L<->RR  2: R<->LL  3: R<->LR
            float4 RLlo, RLhi, RRlo, RRhi, LLlo, LLhi, LRlo, LRhi;   // grandchild boxes (named: no local-memory arrays)
            int cRL = 0, cRR = 0, cLL = 0, cLR = 0;
            float4 ulo, uhi;
            if (c.y >= 0) {
                read_node(nodes, c.y, RLlo, RLhi, RRlo, RRhi, cRL, cRR);
                const float area = box_area(rlo, rhi);
                box_union(llo, lhi, RRlo, RRhi, ulo, uhi);
                float g = area - box_area(ulo, uhi);
                if (g > best) { best = g; which = 0; }
                box_union(RLlo, RLhi, llo, lhi, ulo, uhi);
                g = area - box_area(ulo, uhi);
                if (g > best) { best = g; which = 1; }
            }
            if (c.x >= 0) {
                read_node(nodes, c.x, LLlo, LLhi, LRlo, LRhi, cLL, cLR);
                const float area = box_area(llo, lhi);
                box_union(rlo, rhi, LRlo, LRhi, ulo, uhi);
                float g = area - box_area(ulo, uhi);
                if (g > best) { best = g; which = 2; }
                box_union(LLlo, LLhi, rlo, rhi, ulo, uhi);
                g = area - box_area(ulo, uhi);
                if (g > best) { best = g; which = 3; }
            }
            if (which == 0 || which == 1) {                   // L goes below R; a grandchild of R comes up as the new L
                const bool a0 = which == 0;
                const int R = c.y;
                const float4 slo = a0 ? RRlo : RLlo, shi = a0 ? RRhi : RLhi;      // the grandchild that stays
                const float4 plo = a0 ? RLlo : RRlo, phi = a0 ? RLhi : RRhi;      // the one that comes up
                box_union(llo, lhi, slo, shi, ulo, uhi);
                if (a0) write_node(nodes, R, llo, lhi, RRlo, RRhi, c.x, cRR);
                else write_node(nodes, R, RLlo, RLhi, llo, lhi, cRL, c.x);
                children[R] = a0 ? make_int2(c.x, cRR) : make_int2(cRL, c.x);
                node_lo[R] = ulo; node_hi[R] = uhi;
                set_parent(c.x, R);
                c.x = a0 ? cRL : cRR; llo = plo; lhi = phi;
                set_parent(c.x, cur);
                rlo = ulo; rhi = uhi;
                children[cur] = c;
            } else if (which == 2 || which == 3) {            // R goes below L; a grandchild of L comes up as the new R
                const bool a2 = which == 2;
                const int L = c.x;
                const float4 slo = a2 ? LRlo : LLlo, shi = a2 ? LRhi : LLhi;
                const float4 plo = a2 ? LLlo : LRlo, phi = a2 ? LLhi : LRhi;
                box_union(rlo, rhi, slo, shi, ulo, uhi);
                if (a2) write_node(nodes, L, rlo, rhi, LRlo, LRhi, c.y, cLR);
                else write_node(nodes, L, LLlo, LLhi, rlo, rhi, cLL, c.y);
                children[L] = a2 ? make_int2(c.y, cLR) : make_int2(cLL, c.y);
                node_lo[L] = ulo; node_hi[L] = uhi;
                set_parent(c.y, L);
                c.y = a2 ? cLL : cLR; rlo = plo; rhi = phi;
                set_parent(c.y, cur);
                llo = ulo; lhi = uhi;
                children[cur] = c;
            }
        }
        write_node(nodes, cur, llo, lhi, rlo, rhi, c.x, c.y);
        node_lo[cur] = make_float4(fminf(llo.x, rlo.x), fminf(llo.y, rlo.y), fminf(llo.z, rlo.z), 0.f);
        node_hi[cur] = make_float4(fmaxf(lhi.x, rhi.x), fmaxf(lhi.y, rhi.y), fmaxf(lhi.z, rhi.z), 0.f);
        cur = parent_node[cur];
    }
}

// Level-synchronous breadth-first numbering of the first `cap` internal nodes (one CTA).
// top_id[node] = position in the smem-staged copy, or -1.
__global__ void __launch_bounds__(1024)
top_order_kernel(int n_internal, const int2 *children, int cap, int *top_id, int *order, int *meta) {
    typedef cub::BlockScan<int, 1024> Scan;
    __shared__ typename Scan::TempStorage tmp;
    __shared__ int s_begin, s_end;
    if (threadIdx.x == 0) {
        s_begin = 0; s_end = 0;
        if (n_internal > 0 && cap > 0) { order[0] = 0; top_id[0] = 0; s_end = 1; }
    }
    __syncthreads();
    while (true) {
        int begin = s_begin, end = s_end;
        if (begin >= end || end >= cap) break;
        // a level can be wider than the CTA: walk it in chunks, appending in order
        for (int base = begin; base < end; base += blockDim.x) {
            int t = base + threadIdx.x;
            int c0 = -1, c1 = -1;
            if (t < end) {
                int2 c = children[order[t]];
                if (c.x >= 0) c0 = c.x;
                if (c.y >= 0) c1 = c.y;
            }
            int cnt = (c0 >= 0) + (c1 >= 0), pos, total;
            Scan(tmp).ExclusiveSum(cnt, pos, total);
            int tail = s_end;
            __syncthreads();
            if (c0 >= 0) { int p = tail + pos; if (p < cap) { order[p] = c0; top_id[c0] = p; } ++pos; }
            if (c1 >= 0) { int p = tail + pos; if (p < cap) { order[p] = c1; top_id[c1] = p; } }
            if (threadIdx.x == 0) s_end = min(cap, tail + total);
            __syncthreads();
        }
        if (threadIdx.x == 0) s_begin = end;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        meta[0] = s_end;                                   // n_top
        meta[2] = n_internal;
    }
}

// Rewrites child references into the final encoding (node in top: its top index; other node:
// n_top + index; leaf: ~prim) and emits the breadth-first top copy.
__global__ void finalize_kernel(int n_internal, const int *top_id, int *meta, float4 *nodes, float4 *top) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int n_top = meta[0];
    if (i == 0) meta[1] = n_internal > 0 ? (top_id[0] >= 0 ? top_id[0] : n_top) : -1;
    if (i >= n_internal) return;
    float4 n3 = nodes[4 * (size_t)i + 3];
    int cl = __float_as_int(n3.x), cr = __float_as_int(n3.y);
    if (cl >= 0) cl = top_id[cl] >= 0 ? top_id[cl] : n_top + cl;
    if (cr >= 0) cr = top_id[cr] >= 0 ? top_id[cr] : n_top + cr;
    n3 = make_float4(__int_as_float(cl), __int_as_float(cr), 0.f, 0.f);
    nodes[4 * (size_t)i + 3] = n3;
    int t = top_id[i];
    if (t >= 0) {
        top[4 * t + 0] = nodes[4 * (size_t)i + 0];
        top[4 * t + 1] = nodes[4 * (size_t)i + 1];
        top[4 * t + 2] = nodes[4 * (size_t)i + 2];
        top[4 * t + 3] = n3;
    }
}

// 4-wide nodes for the persistent walk kernel of large scenes (rt_path.cuh:extend_walk_kernel<.., WIDE>).
// That kernel is bound by the latency of its dependent node fetches (one 64 B node per box step, 28.8 steps per ray on
// the 1 M-triangle scene), so the remedy is fewer, fatter steps: wide[ref] holds the boxes and references of the
// GRANDCHILDREN of binary node `ref` (a leaf child stays one slot), one 128 B line, and a walk that starts at the root
// only ever lands on every other level of the binary tree.  Every reference gets a wide node — half of them are never
// visited, which costs build bandwidth (< 0.1 ms per million) but no cache footprint, and needs no depth pass.
// References keep the final encoding of finalize_kernel (< n_top: top copy, >= n_top: n_top + node index, < 0: ~prim),
// so wide[] is indexed by the reference itself.  Layout (8 float4):
//   lo.x[4] lo.y[4] lo.z[4] hi.x[4] hi.y[4] hi.z[4] bits(ref[4]) -      empty slot: ref = 0x80000000
__global__ void widen_kernel(int n_refs, int n_top, const float4 *__restrict__ nodes, const float4 *__restrict__ top,
                             float4 *__restrict__ wide) {
    const int ref = blockIdx.x * blockDim.x + threadIdx.x;
    if (ref >= n_refs) return;
    auto record = [&](int r) { return r < n_top ? top + 4 * (size_t)r : nodes + 4 * (size_t)(r - n_top); };
    float lo[3][4], hi[3][4];
    int c[4];
    for (int k = 0; k < 4; ++k) {
        c[k] = (int)0x80000000;
        for (int a = 0; a < 3; ++a) lo[a][k] = hi[a][k] = 0.f;
    }
    int k = 0;
    auto put = [&](float x0, float y0, float z0, float x1, float y1, float z1, int r) {
        lo[0][k] = x0; lo[1][k] = y0; lo[2][k] = z0; hi[0][k] = x1; hi[1][k] = y1; hi[2][k] = z1; c[k] = r;
        ++k;
    };
    auto expand = [&](float x0, float y0, float z0, float x1, float y1, float z1, int r) {
        if (r < 0) { put(x0, y0, z0, x1, y1, z1, r); return; }           // a leaf child is one slot
        const float4 *q = record(r);
        const float4 m0 = q[0], m1 = q[1], m2 = q[2], m3 = q[3];
        put(m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, __float_as_int(m3.x));
        put(m1.z, m1.w, m2.x, m2.y, m2.z, m2.w, __float_as_int(m3.y));
    };
    const float4 *p = record(ref);
    const float4 n0 = p[0], n1 = p[1], n2 = p[2], n3 = p[3];
    expand(n0.x, n0.y, n0.z, n0.w, n1.x, n1.y, __float_as_int(n3.x));
    expand(n1.z, n1.w, n2.x, n2.y, n2.z, n2.w, __float_as_int(n3.y));
    float4 *w = wide + 8 * (size_t)ref;
    for (int a = 0; a < 3; ++a) {
        w[a] = make_float4(lo[a][0], lo[a][1], lo[a][2], lo[a][3]);
        w[3 + a] = make_float4(hi[a][0], hi[a][1], hi[a][2], hi[a][3]);
    }
    w[6] = make_float4(__int_as_float(c[0]), __int_as_float(c[1]), __int_as_float(c[2]), __int_as_float(c[3]));
    w[7] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// Quantised binary nodes for the persistent walk kernel (rt_path.cuh:extend_walk_kernel<.., 2>): entry `ref` holds the two
// child boxes of node `ref` as 16-bit cell indices on a grid (base, cell) over the scene bounds + the two references,
// 32 B instead of 64.  lo faces round down and move one more cell out, hi faces round up and one more cell out: the
// kernel dequantises with one float32 FFMA per plane (error ~0.02 cells), so the box it tests always contains the
// float32 box of the tree.  Indices clamp to [0, 65535]: only the far-away placeholder boxes of rectangles kept outside
// the hierarchy lie beyond the grid, and they collapse into its last cell.
__global__ void quantize_kernel(int n_refs, int n_top, const float4 *__restrict__ nodes, const float4 *__restrict__ top,
                                float3 base, float3 cell, uint4 *__restrict__ out) {
    const int ref = blockIdx.x * blockDim.x + threadIdx.x;
    if (ref == 0) {
        reinterpret_cast<float4 *>(out)[0] = make_float4(base.x, base.y, base.z, 0.f);
        reinterpret_cast<float4 *>(out)[1] = make_float4(cell.x, cell.y, cell.z, 0.f);
    }
    if (ref >= n_refs) return;
    const float4 *p = ref < n_top ? top + 4 * (size_t)ref : nodes + 4 * (size_t)(ref - n_top);
    const float4 n0 = p[0], n1 = p[1], n2 = p[2], n3 = p[3];
    auto pack = [](float lo, float hi, float b, float c) -> unsigned {
        const float ql = fminf(fmaxf(floorf((lo - b) / c) - 1.f, 0.f), 65535.f);
        const float qh = fminf(fmaxf(ceilf((hi - b) / c) + 1.f, 0.f), 65535.f);
        return (unsigned)ql | ((unsigned)qh << 16);
    };
    uint4 a, b;
    a.x = pack(n0.x, n0.w, base.x, cell.x); a.y = pack(n0.y, n1.x, base.y, cell.y); a.z = pack(n0.z, n1.y, base.z, cell.z);
    a.w = pack(n1.z, n2.y, base.x, cell.x); b.x = pack(n1.w, n2.z, base.y, cell.y); b.y = pack(n2.x, n2.w, base.z, cell.z);
    b.z = (unsigned)__float_as_int(n3.x); b.w = (unsigned)__float_as_int(n3.y);
    out[2 + 2 * (size_t)ref] = a;
    out[3 + 2 * (size_t)ref] = b;
}

inline size_t align_up(size_t v) { return (v + 255) & ~size_t(255); }

struct TempLayout {
    size_t box_lo, box_hi, keys_a, keys_b, vals_a, vals_b, children, parent_node, parent_leaf, flags, node_lo, node_hi,
        top_id, order, bounds, meta, cub, total, cub_bytes, moments, pos_of;
};

TempLayout layout(int n) {
    TempLayout L;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes); return o; };
    size_t m = (size_t)(n > 1 ? n - 1 : 1);
    L.box_lo = take(16 * (size_t)n); L.box_hi = take(16 * (size_t)n);
    L.keys_a = take(4 * (size_t)n); L.keys_b = take(4 * (size_t)n);
    L.vals_a = take(4 * (size_t)n); L.vals_b = take(4 * (size_t)n);
    L.children = take(8 * m); L.parent_node = take(4 * m); L.parent_leaf = take(4 * (size_t)n);
    L.flags = take(4 * m); L.node_lo = take(16 * m); L.node_hi = take(16 * m);
    L.top_id = take(4 * m); L.order = take(4 * 4096); L.bounds = take(32); L.meta = take(32);
    L.moments = take(96);
    L.pos_of = take(4 * (size_t)n);
    size_t cub_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, (const uint32_t *)nullptr, (uint32_t *)nullptr,
                                    (const int *)nullptr, (int *)nullptr, n, 0, 31);
    L.cub_bytes = cub_bytes;
    L.cub = take(cub_bytes);
    L.total = off;
    return L;
}

}  // namespace

size_t lbvh_temp_bytes(int n_prims) { return layout(n_prims > 0 ? n_prims : 1).total; }

cudaError_t lbvh_build(int n_rect, int n_sphere, int n_tri, const float4 *rect, const float4 *sphere, const float4 *tri,
                       float pad, float4 *nodes, float4 *top, int top_capacity, int *h_meta, void *temp,
                       size_t temp_bytes, cudaStream_t stream, int build_flags) {
    int n = n_rect + n_sphere + n_tri;
    // build_flags bit 0: the rectangles stay OUTSIDE the hierarchy (the caller tests them directly before every walk).
    // A few room-sized rectangles inside a fine mesh would otherwise widen every ancestor box of their leaves.
    const int n_outside = ((build_flags & 1) && n_rect < n - 1) ? n_rect : 0;
    h_meta[0] = 0; h_meta[1] = -1; h_meta[2] = 0;
    if (n <= 0) { h_meta[1] = 0; return cudaSuccess; }
    if (n == 1) { h_meta[1] = ~0; return cudaSuccess; }        // a single leaf: root = ~prim 0
    TempLayout L = layout(n);
    if (temp_bytes < L.total) return cudaErrorInvalidValue;
    if (top_capacity > 4096) top_capacity = 4096;
    char *base = (char *)temp;
    float4 *box_lo = (float4 *)(base + L.box_lo), *box_hi = (float4 *)(base + L.box_hi);
    uint32_t *keys_a = (uint32_t *)(base + L.keys_a), *keys_b = (uint32_t *)(base + L.keys_b);
    int *vals_a = (int *)(base + L.vals_a), *vals_b = (int *)(base + L.vals_b);
    int2 *children = (int2 *)(base + L.children);
    int *parent_node = (int *)(base + L.parent_node), *parent_leaf = (int *)(base + L.parent_leaf);
    int *flags = (int *)(base + L.flags);
    float4 *node_lo = (float4 *)(base + L.node_lo), *node_hi = (float4 *)(base + L.node_hi);
    int *top_id = (int *)(base + L.top_id), *order = (int *)(base + L.order);
    int *bounds = (int *)(base + L.bounds), *meta = (int *)(base + L.meta);
    double *moments = (double *)(base + L.moments);
    int *pos_of = (int *)(base + L.pos_of);

    const int T = 256, G = (n + T - 1) / T;
    // scene centroid bounds start at (+max, -max) in the ordered-int encoding
    int init[8] = {0x7f7fffff, 0x7f7fffff, 0x7f7fffff, (int)0x80800000, (int)0x80800000, (int)0x80800000, 0, 0};
    cudaError_t e;
    if ((e = cudaMemcpyAsync(bounds, init, sizeof init, cudaMemcpyHostToDevice, stream))) return e;
    if ((e = cudaMemsetAsync(flags, 0, 4 * (size_t)(n - 1), stream))) return e;
    if ((e = cudaMemsetAsync(top_id, 0xff, 4 * (size_t)(n - 1), stream))) return e;
    if ((e = cudaMemsetAsync(meta, 0, 32, stream))) return e;
    if ((e = cudaMemsetAsync(moments, 0, 96, stream))) return e;
    bounds_kernel<<<G, T, 0, stream>>>(n, n_rect, n_sphere, rect, sphere, tri, pad, box_lo, box_hi, bounds, moments, n_outside);
    int mb[3] = {0, 0, 0};                                   // 0: automatic (morton_bits)
    if (const char *ev = getenv("B2RT_MORTON_BITS")) {       // measurement hook: "10,10,10" forces an allocation
        if (sscanf(ev, "%d,%d,%d", &mb[0], &mb[1], &mb[2]) != 3 || mb[0] + mb[1] + mb[2] > 30 || mb[0] < 0 || mb[1] < 0 || mb[2] < 0)
            mb[0] = mb[1] = mb[2] = 0;
    }
    morton_kernel<<<G, T, 0, stream>>>(n, box_lo, box_hi, bounds, keys_a, vals_a, moments, mb[0], mb[1], mb[2], n_outside);
    size_t cub_bytes = L.cub_bytes;
    if ((e = cub::DeviceRadixSort::SortPairs(base + L.cub, cub_bytes, keys_a, keys_b, vals_a, vals_b, n, 0, 31, stream)))
        return e;
    hierarchy_kernel<<<G, T, 0, stream>>>(n, keys_b, vals_b, children, parent_node, parent_leaf, pos_of);
    // refit + rotation sweeps (each sweep climbs the tree the previous one left behind)
    // measured on 1 M triangles with the bit allocation above: 90.45 / 89.87 / 89.19 / 89.10 ms per step for 0..3 sweeps,
    // +0.28 ms of build each: two sweeps minimise build + render for a 64-spp frame
    int sweeps = (build_flags & 2) ? 0 : 2;
    if (const char *ev = getenv("B2RT_LBVH_SWEEPS")) sweeps = atoi(ev);      // measurement hook
    refit_kernel<<<G, T, 0, stream>>>(n, children, parent_node, parent_leaf, pos_of, box_lo, box_hi, flags, node_lo, node_hi,
                                      nodes, sweeps > 0 ? 1 : 0);
    for (int sw = 1; sw < sweeps; ++sw) {
        if ((e = cudaMemsetAsync(flags, 0, 4 * (size_t)(n - 1), stream))) return e;
        refit_kernel<<<G, T, 0, stream>>>(n, children, parent_node, parent_leaf, pos_of, box_lo, box_hi, flags, node_lo,
                                          node_hi, nodes, 1);
    }
    top_order_kernel<<<1, 1024, 0, stream>>>(n - 1, children, top_capacity, top_id, order, meta);
    finalize_kernel<<<(n - 1 + T - 1) / T, T, 0, stream>>>(n - 1, top_id, meta, nodes, top);
    if ((e = cudaGetLastError())) return e;
    int hm[8];
    if ((e = cudaMemcpyAsync(hm, meta, 32, cudaMemcpyDeviceToHost, stream))) return e;
    if ((e = cudaStreamSynchronize(stream))) return e;
    h_meta[0] = hm[0]; h_meta[1] = hm[1]; h_meta[2] = hm[2];
    return cudaSuccess;
}

size_t lbvh_wide_bytes(int n_top, int n_internal) {
    const size_t n = (size_t)(n_top > 0 ? n_top : 0) + (size_t)(n_internal > 0 ? n_internal : 0);
    return (n > 0 ? n : 1) * 128;
}

cudaError_t lbvh_widen(const float4 *nodes, const float4 *top, int n_top, int n_internal, float4 *wide, size_t wide_bytes,
                       cudaStream_t stream) {
    if (n_top < 0 || n_internal < 0 || (n_internal > 0 && (!nodes || !wide)) || (n_top > 0 && !top)) return cudaErrorInvalidValue;
    if (wide_bytes < lbvh_wide_bytes(n_top, n_internal)) return cudaErrorInvalidValue;
    const int n_refs = n_top + n_internal;
    if (n_internal == 0) return cudaSuccess;                  // no internal node: the root reference is a leaf (or nothing)
    widen_kernel<<<(n_refs + 255) / 256, 256, 0, stream>>>(n_refs, n_top, nodes, top, wide);
    return cudaGetLastError();
}

size_t lbvh_quant_bytes(int n_top, int n_internal) {
    const size_t n = (size_t)(n_top > 0 ? n_top : 0) + (size_t)(n_internal > 0 ? n_internal : 0);
    return 32 + (n > 0 ? n : 1) * 32;
}

cudaError_t lbvh_quantize(const float4 *nodes, const float4 *top, int n_top, int n_internal, const float *lo, const float *hi,
                          void *quant, size_t quant_bytes, cudaStream_t stream) {
    if (n_top < 0 || n_internal < 0 || !lo || !hi || !quant || (n_internal > 0 && !nodes) || (n_top > 0 && !top))
        return cudaErrorInvalidValue;
    if (quant_bytes < lbvh_quant_bytes(n_top, n_internal)) return cudaErrorInvalidValue;
    // grid: 65536 cells per axis, the bounds occupy cells [2, 65533] so that the outward cell never clamps
    float ext_max = 1e-3f;
    for (int k = 0; k < 3; ++k) {
        if (!(hi[k] >= lo[k])) return cudaErrorInvalidValue;
        ext_max = fmaxf(ext_max, fmaxf(hi[k] - lo[k], fmaxf(fabsf(lo[k]), fabsf(hi[k])) * 1e-6f));
    }
    float b[3], c[3];
    for (int k = 0; k < 3; ++k) {
        const float ext = fmaxf(hi[k] - lo[k], ext_max * 1e-6f);          // a flat axis still gets a non-zero cell
        c[k] = ext / 65531.f;
        b[k] = lo[k] - 2.f * c[k];
    }
    const int n_refs = n_top + n_internal;
    quantize_kernel<<<(n_refs > 0 ? n_refs + 255 : 256) / 256, 256, 0, stream>>>(n_internal > 0 ? n_refs : 0, n_top, nodes, top,
                                                                                make_float3(b[0], b[1], b[2]),
                                                                                make_float3(c[0], c[1], c[2]), (uint4 *)quant);
    return cudaGetLastError();
}

}  // namespace b2rt
