// scene_prepare.cu — derives the optional small-scene acceleration data of b2rt_scene INSIDE the library, from the
// base streams alone (b2rt_scene_prepare / b2rt_scene_prepare_host, include/b200rt.h):
//   planar scan records ("plane + two edge planes", coplanar triangle pairs merged into parallelograms),
//   box records (parallelogram faces of a common parallelepiped behind one three-slab test),
//   per-primitive surface records, padded scene bounds, per-light occluder hints.
// A binder that fills only the six base streams gets the same fast kernels as the Python packer (which keeps an
// independent numpy implementation of the same derivations, b200rt/packer.py — the CPU tests compare the two).
// Everything here is small-n host arithmetic in double precision (n_prims <= 64); nothing on the per-ray path.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "../../include/b200rt.h"

namespace b2rt {
namespace prep {

struct D3 { double x, y, z; };
static inline D3 operator+(D3 a, D3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
static inline D3 operator-(D3 a, D3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
static inline D3 operator*(D3 a, double k) { return {a.x * k, a.y * k, a.z * k}; }
static inline double dot(D3 a, D3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline D3 cross(D3 a, D3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
static inline double maxabs(D3 a) { return std::max(fabs(a.x), std::max(fabs(a.y), fabs(a.z))); }
static inline double norm(D3 a) { return sqrt(dot(a, a)); }

struct Rec {                   // one planar scan record before packing
    double q0[4], q1[4], q2[4], umax, vmax;
    int kind, ida, idb;
    bool quad;                 // a parallelogram (rectangle or triangle pair): can be a box face
    D3 p0, ea, eb;             // corner and the two edges (quad only)
};

static inline float bits_f(uint32_t v) { float f; memcpy(&f, &v, 4); return f; }

static void edge_planes(D3 v0, D3 e1, D3 e2, Rec &r) {
    const D3 N = cross(e1, e2);
    D3 a1 = cross(e2, N); a1 = a1 * (1.0 / dot(e1, a1));
    D3 a2 = cross(N, e1); a2 = a2 * (1.0 / dot(e2, a2));
    r.q0[0] = N.x; r.q0[1] = N.y; r.q0[2] = N.z; r.q0[3] = dot(N, v0);
    r.q1[0] = a1.x; r.q1[1] = a1.y; r.q1[2] = a1.z; r.q1[3] = -dot(a1, v0);
    r.q2[0] = a2.x; r.q2[1] = a2.y; r.q2[2] = a2.z; r.q2[3] = -dot(a2, v0);
    r.umax = r.vmax = 1.0;
}

// packer.build_scan_prims
static std::vector<Rec> build_scan_prims(int n_rect, int n_sphere, int n_tri, const float *rect, const float *tri) {
    const double pair_tol = 1e-5;
    std::vector<Rec> recs;
    for (int i = 0; i < n_rect; ++i) {
        const float *q = rect + 16 * i;
        const D3 anchor = {q[0], q[1], q[2]}, n = {q[4], q[5], q[6]}, uu = {q[8], q[9], q[10]}, vv = {q[12], q[13], q[14]};
        const double ul = q[3], vl = q[7];
        Rec r{};
        r.q0[0] = n.x; r.q0[1] = n.y; r.q0[2] = n.z; r.q0[3] = dot(n, anchor);
        r.q1[0] = uu.x; r.q1[1] = uu.y; r.q1[2] = uu.z; r.q1[3] = -dot(uu, anchor);
        r.q2[0] = vv.x; r.q2[1] = vv.y; r.q2[2] = vv.z; r.q2[3] = -dot(vv, anchor);
        r.umax = ul; r.vmax = vl; r.kind = 0; r.ida = i; r.idb = 0;
        r.quad = true; r.p0 = anchor; r.ea = uu * ul; r.eb = vv * vl;
        recs.push_back(r);
    }
    const int base = n_rect + n_sphere;
    std::vector<char> used(n_tri, 0);
    auto V = [&](int t, int row) { const float *q = tri + 12 * t + 4 * row; return D3{q[0], q[1], q[2]}; };
    for (int i = 0; i < n_tri; ++i) {
        if (used[i]) continue;
        const D3 v0 = V(i, 0), e1 = V(i, 1), e2 = V(i, 2);
        const double scale = std::max(std::max(maxabs(v0), maxabs(e1)), std::max(maxabs(e2), 1e-30));
        int mate = -1;
        for (int j = 0; j < n_tri && mate < 0; ++j) {
            if (j == i || used[j]) continue;
            const D3 w0 = V(j, 0), f1 = V(j, 1), f2 = V(j, 2);
            // j = (p0, p2, p3) with p2 = i.v2 and p3 = p0 + (e2_i - e1_i): the other half of a parallelogram
            if (maxabs(w0 - v0) <= pair_tol * scale && maxabs(f1 - e2) <= pair_tol * scale &&
                maxabs(f2 - (e2 - e1)) <= pair_tol * scale)
                mate = j;
        }
        used[i] = 1;
        Rec r{};
        if (mate >= 0) {
            used[mate] = 1;
            const D3 eq1 = e1, eq2 = V(mate, 2);
            edge_planes(v0, eq1, eq2, r);
            r.kind = i < mate ? 2 : 3;             // diagonal ties go to the lower packed id
            r.ida = base + i; r.idb = base + mate;
            r.quad = true; r.p0 = v0; r.ea = eq1; r.eb = eq2;
        } else {
            edge_planes(v0, e1, e2, r);
            r.kind = 1; r.ida = base + i; r.idb = 0; r.quad = false;
        }
        recs.push_back(r);
    }
    return recs;
}

struct Box { D3 C; double H[3][3]; int slot[6]; int count; };      // H columns = half axes; slot[f] = record or -1

static bool inv3(const double a[3][3], double inv[3][3], double *det_out) {
    const double c00 = a[1][1] * a[2][2] - a[1][2] * a[2][1], c01 = a[1][2] * a[2][0] - a[1][0] * a[2][2],
                 c02 = a[1][0] * a[2][1] - a[1][1] * a[2][0];
    const double det = a[0][0] * c00 + a[0][1] * c01 + a[0][2] * c02;
    if (det_out) *det_out = det;
    if (det == 0.0) return false;
    const double id = 1.0 / det;
    inv[0][0] = c00 * id; inv[0][1] = (a[0][2] * a[2][1] - a[0][1] * a[2][2]) * id; inv[0][2] = (a[0][1] * a[1][2] - a[0][2] * a[1][1]) * id;
    inv[1][0] = c01 * id; inv[1][1] = (a[0][0] * a[2][2] - a[0][2] * a[2][0]) * id; inv[1][2] = (a[0][2] * a[1][0] - a[0][0] * a[1][2]) * id;
    inv[2][0] = c02 * id; inv[2][1] = (a[0][1] * a[2][0] - a[0][0] * a[2][1]) * id; inv[2][2] = (a[0][0] * a[1][1] - a[0][1] * a[1][0]) * id;
    return true;
}

// packer.group_scan_boxes -> order (loose first), n_loose, box records float[16 * n_box]
static void group_scan_boxes(const std::vector<Rec> &recs, std::vector<int> &order, int &n_loose, std::vector<float> &boxes) {
    const double tol = 1e-5;
    const int n = (int)recs.size();
    struct Face { D3 c, a, b; };
    std::vector<int> keys;
    std::vector<Face> face(n);
    for (int k = 0; k < n; ++k)
        if (recs[k].quad) {
            keys.push_back(k);
            face[k] = {recs[k].p0 + (recs[k].ea + recs[k].eb) * 0.5, recs[k].ea * 0.5, recs[k].eb * 0.5};
        }
    auto same_dir = [&](D3 x, D3 y, double scale) { return std::min(maxabs(x - y), maxabs(x + y)) <= tol * scale; };
    auto same_edges = [&](D3 a, D3 b, D3 x, D3 y, double scale) {
        return (same_dir(a, x, scale) && same_dir(b, y, scale)) || (same_dir(a, y, scale) && same_dir(b, x, scale));
    };
    std::vector<Box> cands;
    for (size_t ii = 0; ii < keys.size(); ++ii) {
        const int i = keys[ii];
        const Face &fi = face[i];
        for (size_t jj = ii + 1; jj < keys.size(); ++jj) {
            const int j = keys[jj];
            const Face &fj = face[j];
            const double scale = std::max(std::max(maxabs(fi.a), maxabs(fi.b)), std::max(maxabs(fj.c - fi.c), 1e-30));
            if (!same_edges(fi.a, fi.b, fj.a, fj.b, scale)) continue;
            const D3 h = (fj.c - fi.c) * 0.5;
            Box bx{};
            const D3 cols[3] = {fi.a, fi.b, h};
            for (int c = 0; c < 3; ++c) { bx.H[0][c] = cols[c].x; bx.H[1][c] = cols[c].y; bx.H[2][c] = cols[c].z; }
            double inv[3][3], det;
            inv3(bx.H, inv, &det);
            if (fabs(det) <= 1e-9 * scale * scale * scale) continue;
            bx.C = (fi.c + fj.c) * 0.5;
            for (int f = 0; f < 6; ++f) bx.slot[f] = -1;
            bx.slot[4] = i; bx.slot[5] = j;
            for (int f : keys) {
                if (f == i || f == j) continue;
                const Face &ff = face[f];
                for (int k = 0; k < 2; ++k) {
                    const D3 other0 = cols[1 - k], other1 = h;
                    for (int sgn = -1; sgn <= 1; sgn += 2) {
                        if (maxabs(ff.c - (bx.C + cols[k] * (double)sgn)) <= tol * scale &&
                            same_edges(ff.a, ff.b, other0, other1, scale)) {
                            const int s = 2 * k + (sgn > 0 ? 1 : 0);
                            if (bx.slot[s] < 0) bx.slot[s] = f;
                        }
                    }
                }
            }
            bx.count = 0;
            for (int f = 0; f < 6; ++f) bx.count += bx.slot[f] >= 0;
            cands.push_back(bx);
        }
    }
    std::stable_sort(cands.begin(), cands.end(), [](const Box &a, const Box &b) { return a.count > b.count; });
    std::vector<char> taken(n, 0);
    std::vector<Box> chosen;
    for (const Box &c : cands) {
        if (c.count < 3) continue;
        bool clash = false;
        for (int f = 0; f < 6; ++f) clash |= c.slot[f] >= 0 && taken[c.slot[f]];
        if (clash) continue;
        for (int f = 0; f < 6; ++f) if (c.slot[f] >= 0) taken[c.slot[f]] = 1;
        chosen.push_back(c);
    }
    order.clear();
    for (int k = 0; k < n; ++k) if (!taken[k]) order.push_back(k);
    n_loose = (int)order.size();
    for (int k = 0; k < n; ++k) if (taken[k]) order.push_back(k);
    std::vector<int> new_index(n, 0);
    for (int k = 0; k < n; ++k) new_index[order[k]] = k;
    boxes.assign(16 * chosen.size(), 0.f);
    for (size_t b = 0; b < chosen.size(); ++b) {
        const Box &c = chosen[b];
        double M[3][3];
        inv3(c.H, M, nullptr);                          // rows m_k: l = M (P - C)
        float *o = boxes.data() + 16 * b;
        for (int k = 0; k < 3; ++k) {
            o[4 * k] = (float)M[k][0]; o[4 * k + 1] = (float)M[k][1]; o[4 * k + 2] = (float)M[k][2];
            o[4 * k + 3] = (float)(-(M[k][0] * c.C.x + M[k][1] * c.C.y + M[k][2] * c.C.z));
        }
        uint32_t code[6];
        bool closed = true;
        for (int f = 0; f < 6; ++f) { code[f] = c.slot[f] >= 0 ? (uint32_t)new_index[c.slot[f]] : 255u; closed &= c.slot[f] >= 0; }
        o[12] = bits_f(code[0] | code[1] << 8 | code[2] << 16 | code[3] << 24);
        o[13] = bits_f(code[4] | code[5] << 8);
        o[14] = bits_f(closed ? 1u : 0u);
        o[15] = 0.f;
    }
}

static void pack_records(const std::vector<Rec> &recs, const std::vector<int> &order, float *out) {
    for (size_t k = 0; k < order.size(); ++k) {
        const Rec &r = recs[order[k]];
        float *o = out + 16 * k;
        for (int c = 0; c < 4; ++c) { o[c] = (float)r.q0[c]; o[4 + c] = (float)r.q1[c]; o[8 + c] = (float)r.q2[c]; }
        o[12] = (float)r.umax; o[13] = (float)r.vmax;
        o[14] = bits_f(((uint32_t)r.kind << 28) | (uint32_t)r.ida);
        o[15] = bits_f((uint32_t)r.idb);
    }
}

// packer.build_surface_records: float[20 * n_prims]
static void build_surface_records(int n_rect, int n_sphere, int n_tri, const float *rect, const float *sphere,
                                  const float *shade, const float *mat, const int32_t *prim_mat, const int32_t *mat_tex,
                                  float *out) {
    const int n = n_rect + n_sphere + n_tri;
    memset(out, 0, sizeof(float) * 20 * (size_t)n);
    for (int i = 0; i < n_rect; ++i) {
        float *o = out + 20 * i;
        const float *q = rect + 16 * i;
        o[0] = q[4]; o[1] = q[5]; o[2] = q[6];
        o[13] = (float)(1.0 / (double)q[3]);           // du_a = 1 / u_len
        o[18] = (float)(1.0 / (double)q[7]);           // dv_b = 1 / v_len
    }
    for (int i = 0; i < n_sphere; ++i) {
        float *o = out + 20 * (n_rect + i);
        const float *q = sphere + 8 * i;
        o[0] = q[0]; o[1] = q[1]; o[2] = q[2]; o[3] = (float)(1.0 / (double)q[3]);
    }
    const int base = n_rect + n_sphere;
    for (int i = 0; i < n_tri; ++i) {
        const int k = base + i;
        float *o = out + 20 * k;
        const float *sh = shade + 12 * k;
        o[0] = sh[0]; o[1] = sh[1]; o[2] = sh[2];
        uint32_t flags = 1u;
        if (sh[10] != 0.f) {
            const double u0 = sh[4], v0 = sh[5], u1 = sh[6], v1 = sh[7], u2 = sh[8], v2 = sh[9];
            o[12] = (float)u0; o[13] = (float)(u1 - u0); o[14] = (float)(u2 - u0);
            o[16] = (float)v0; o[17] = (float)(v1 - v0); o[18] = (float)(v2 - v0);
        }
        o[19] = bits_f(flags);
    }
    for (int k = 0; k < n; ++k) {
        float *o = out + 20 * k;
        const int m = prim_mat[k];
        for (int c = 0; c < 4; ++c) { o[4 + c] = mat[8 * m + c]; o[8 + c] = mat[8 * m + 4 + c]; }
        o[15] = bits_f((uint32_t)mat_tex[m]);
        if (k < base) o[19] = bits_f(0u);
    }
}

static void scene_bounds(int n_rect, int n_sphere, int n_tri, const float *rect, const float *sphere, const float *tri,
                         float lo_out[3], float hi_out[3]) {
    double lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    auto add = [&](D3 p) {
        lo[0] = std::min(lo[0], p.x); lo[1] = std::min(lo[1], p.y); lo[2] = std::min(lo[2], p.z);
        hi[0] = std::max(hi[0], p.x); hi[1] = std::max(hi[1], p.y); hi[2] = std::max(hi[2], p.z);
    };
    for (int i = 0; i < n_rect; ++i) {
        const float *q = rect + 16 * i;
        const D3 a = {q[0], q[1], q[2]}, u = D3{q[8], q[9], q[10]} * (double)q[3], v = D3{q[12], q[13], q[14]} * (double)q[7];
        add(a); add(a + u); add(a + v); add(a + u + v);
    }
    for (int i = 0; i < n_sphere; ++i) {
        const float *q = sphere + 8 * i;
        const double r = q[3];
        add({q[0] - r, q[1] - r, q[2] - r}); add({q[0] + r, q[1] + r, q[2] + r});
    }
    for (int i = 0; i < n_tri; ++i) {
        const float *q = tri + 12 * i;
        const D3 v0 = {q[0], q[1], q[2]}, e1 = {q[4], q[5], q[6]}, e2 = {q[8], q[9], q[10]};
        add(v0); add(v0 + e1); add(v0 + e2);
    }
    if (!(lo[0] <= hi[0])) { for (int k = 0; k < 3; ++k) { lo_out[k] = 1.f; hi_out[k] = -1.f; } return; }
    double m = 1e-3;
    for (int k = 0; k < 3; ++k) m = std::max(m, std::max(fabs(lo[k]), fabs(hi[k])));
    const double pad = 1e-4 * m;
    for (int k = 0; k < 3; ++k) { lo_out[k] = (float)(lo[k] - pad); hi_out[k] = (float)(hi[k] + pad); }
}

// packer.build_occluder_hints (scan-record codes): per light sample the record (k) or sphere (64 + i) that blocks the
// most next-event shadow rays from area-weighted random surface points.  Performance hint only: any deterministic
// estimate is valid; the points come from a fixed LCG.
static void build_occluder_hints(int n_rect, int n_sphere, int n_tri, const float *rect, const float *sphere, const float *tri,
                                 const float *lights, int n_lights, const float *scan, int n_scan, int32_t *hints) {
    for (int j = 0; j < std::max(1, n_lights); ++j) hints[j] = -1;
    if (n_lights == 0 || n_scan == 0) return;
    uint64_t state = 0x9E3779B97F4A7C15ull;
    auto rnd = [&]() { state = state * 6364136223846793005ull + 1442695040888963407ull; return (double)(state >> 11) * (1.0 / 9007199254740992.0); };
    struct Pt { D3 p, n; };
    std::vector<Pt> pts;
    std::vector<double> area;
    double total = 0.0;
    for (int i = 0; i < n_rect; ++i) { area.push_back((double)rect[16 * i + 3] * rect[16 * i + 7]); total += area.back(); }
    for (int i = 0; i < n_sphere; ++i) { const double r = sphere[8 * i + 3]; area.push_back(4 * M_PI * r * r); total += area.back(); }
    for (int i = 0; i < n_tri; ++i) {
        const float *q = tri + 12 * i;
        area.push_back(0.5 * norm(cross(D3{q[4], q[5], q[6]}, D3{q[8], q[9], q[10]}))); total += area.back();
    }
    if (!(total > 0.0)) total = 1.0;
    const int n_points = 4096;
    int a_idx = 0;
    for (int i = 0; i < n_rect; ++i, ++a_idx) {
        const float *q = rect + 16 * i;
        const int m = std::max(4, (int)lround(n_points * area[a_idx] / total));
        const D3 a = {q[0], q[1], q[2]}, n = {q[4], q[5], q[6]}, u = {q[8], q[9], q[10]}, v = {q[12], q[13], q[14]};
        for (int k = 0; k < m; ++k) pts.push_back({a + u * (rnd() * q[3]) + v * (rnd() * q[7]), n});
    }
    for (int i = 0; i < n_sphere; ++i, ++a_idx) {
        const float *q = sphere + 8 * i;
        const int m = std::max(4, (int)lround(n_points * area[a_idx] / total));
        for (int k = 0; k < m; ++k) {
            const double z = 1 - 2 * rnd(), ph = 2 * M_PI * rnd(), s = sqrt(std::max(0.0, 1 - z * z));
            const D3 d = {s * cos(ph), s * sin(ph), z};
            pts.push_back({D3{q[0], q[1], q[2]} + d * (double)q[3], d});
        }
    }
    for (int i = 0; i < n_tri; ++i, ++a_idx) {
        const float *q = tri + 12 * i;
        const int m = std::max(4, (int)lround(n_points * area[a_idx] / total));
        const D3 v0 = {q[0], q[1], q[2]}, e1 = {q[4], q[5], q[6]}, e2 = {q[8], q[9], q[10]};
        D3 n = cross(e1, e2);
        const double ln = norm(n);
        n = n * (1.0 / (ln > 0 ? ln : 1.0));
        for (int k = 0; k < m; ++k) {
            double u = rnd(), v = rnd();
            if (u + v > 1) { u = 1 - u; v = 1 - v; }
            pts.push_back({v0 + e1 * u + e2 * v, n * (rnd() < 0.5 ? 1.0 : -1.0)});       // both sides of the surface
        }
    }
    std::vector<int> counts(64 + std::max(n_sphere, 1));
    for (int j = 0; j < n_lights; ++j) {
        const D3 L = {lights[4 * j], lights[4 * j + 1], lights[4 * j + 2]};
        std::fill(counts.begin(), counts.end(), 0);
        for (const Pt &pt : pts) {
            D3 d = L - pt.p;
            const double dist = norm(d);
            if (!(dist > 1e-3)) continue;
            d = d * (1.0 / dist);
            if (!(dot(d, pt.n) > 0)) continue;           // zero-payload shadow rays are never queued
            const D3 o = pt.p + pt.n * 1e-3;
            for (int k = 0; k < n_scan && k < 64; ++k) {
                const float *q = scan + 16 * k;
                const D3 N = {q[0], q[1], q[2]};
                const double dn = dot(d, N);
                if (!(fabs(dn) > 1e-6)) continue;
                const double t = (q[3] - dot(o, N)) / dn;
                if (!(t > 1e-3 && t < 1e6)) continue;
                const D3 X = o + d * t;
                const double u = dot(X, D3{q[4], q[5], q[6]}) + q[7], v = dot(X, D3{q[8], q[9], q[10]}) + q[11];
                uint32_t w; memcpy(&w, q + 14, 4);
                const int kind = (int)(w >> 28);
                const bool inside = u >= 0 && v >= 0 && (kind == 1 ? (u + v <= 1) : (u <= q[12] && v <= q[13]));
                counts[k] += inside;
            }
            for (int i = 0; i < n_sphere; ++i) {
                const float *q = sphere + 8 * i;
                const D3 oc = o - D3{q[0], q[1], q[2]};
                const double b = dot(oc, d), disc = b * b - (dot(oc, oc) - (double)q[4]);
                if (disc > 0) { const double sq = sqrt(disc); counts[64 + i] += ((-b - sq) > 1e-3) || ((-b + sq) > 1e-3); }
            }
        }
        int best = 0;
        for (int k = 1; k < (int)counts.size(); ++k) if (counts[k] > counts[best]) best = k;
        hints[j] = counts[best] > 0 ? best : -1;
    }
}

}  // namespace prep
}  // namespace b2rt

namespace b2rt { int set_error(int code, const char *fmt, ...); }      // c_api.cu: text for b2rt_last_error()
using b2rt::set_error;

static size_t align256(size_t v) { return (v + 255) & ~size_t(255); }

extern "C" int b2rt_scene_prepare_bytes(int32_t n_rect, int32_t n_sphere, int32_t n_tri, int32_t n_lights, size_t *h_bytes) {
    const size_t n = (size_t)std::max(0, n_rect) + (size_t)std::max(0, n_sphere) + (size_t)std::max(0, n_tri);
    // scan records (<= one planar per rectangle / triangle) + box records (<= n / 3) + surface records + hints
    *h_bytes = align256(64 * (n + n / 3 + 1)) + align256(80 * (n + 1)) + align256(4 * (size_t)std::max(1, n_lights)) + 256;
    return 0;
}

extern "C" int b2rt_scene_prepare_host(int32_t n_rect, int32_t n_sphere, int32_t n_tri, int32_t n_mat, int32_t n_lights,
                                       const float *h_rect, const float *h_sphere, const float *h_tri, const float *h_shade,
                                       const float *h_mat, const int32_t *h_prim_mat, const int32_t *h_mat_tex,
                                       const float *h_lights, int32_t flags, void *h_out, size_t out_bytes,
                                       b2rt_prepare_layout *h_layout) {
    using namespace b2rt::prep;
    (void)n_mat;
    if (!h_layout || !h_out) return set_error(2, "scene_prepare_host: NULL output");
    memset(h_layout, 0, sizeof *h_layout);
    const int n = n_rect + n_sphere + n_tri;
    scene_bounds(n_rect, n_sphere, n_tri, h_rect, h_sphere, h_tri, h_layout->bounds_lo, h_layout->bounds_hi);
    h_layout->scan_offset = h_layout->surface_offset = h_layout->hint_offset = (size_t)-1;
    if (n <= 0 || n > B2RT_SCAN_MAX_PRIMS) return 0;                 // large scenes walk the LBVH: bounds only
    size_t need = 0;
    b2rt_scene_prepare_bytes(n_rect, n_sphere, n_tri, n_lights, &need);
    if (out_bytes < need) return set_error(2, "scene_prepare_host: buffer of %zu bytes, need %zu", out_bytes, need);
    std::vector<Rec> recs = build_scan_prims(n_rect, n_sphere, n_tri, h_rect, h_tri);
    if (recs.empty() || recs.size() > 64) return 0;
    std::vector<int> order(recs.size());
    for (size_t k = 0; k < recs.size(); ++k) order[k] = (int)k;
    int n_loose = (int)recs.size();
    std::vector<float> boxes;
    if (!(flags & B2RT_PREPARE_NO_BOXES)) group_scan_boxes(recs, order, n_loose, boxes);
    char *out = (char *)h_out;
    size_t off = 0;
    float *scan = (float *)(out + off);
    pack_records(recs, order, scan);
    memcpy(scan + 16 * recs.size(), boxes.data(), boxes.size() * sizeof(float));
    h_layout->scan_offset = off;
    h_layout->n_scan_prims = (int32_t)recs.size();
    h_layout->n_scan_loose = n_loose;
    h_layout->n_scan_boxes = (int32_t)(boxes.size() / 16);
    off = align256(off + 64 * (recs.size() + boxes.size() / 16));
    if (!(flags & B2RT_PREPARE_NO_SURFACE)) {
        build_surface_records(n_rect, n_sphere, n_tri, h_rect, h_sphere, h_shade, h_mat, h_prim_mat, h_mat_tex, (float *)(out + off));
        h_layout->surface_offset = off;
        off = align256(off + 80 * (size_t)n);
    }
    if (!(flags & B2RT_PREPARE_NO_HINTS) && n_lights > 0 && n_lights <= 4096) {
        build_occluder_hints(n_rect, n_sphere, n_tri, h_rect, h_sphere, h_tri, h_lights, n_lights, scan, (int)recs.size(),
                             (int32_t *)(out + off));
        h_layout->hint_offset = off;
        off = align256(off + 4 * (size_t)n_lights);
    }
    h_layout->bytes_used = off;
    h_layout->scan_incoherent = 1;
    return 0;
}

extern "C" int b2rt_scene_prepare(b2rt_scene *scene, void *d_buffer, size_t buffer_bytes, int32_t flags, void *stream) {
    if (!scene) return set_error(2, "scene_prepare: NULL scene");
    if (scene->struct_size != sizeof(b2rt_scene) || scene->abi_version != B2RT_ABI_VERSION) {
        return set_error(2, "scene_prepare: b2rt_scene ABI mismatch");
    }
    const int n_rect = scene->n_rect, n_sphere = scene->n_sphere, n_tri = scene->n_tri;
    const int n = n_rect + n_sphere + n_tri;
    scene->n_scan_prims = scene->n_scan_loose = scene->n_scan_boxes = 0;
    scene->d_scan_prims = scene->d_surface_records = nullptr;
    scene->d_occluder_hint = nullptr;
    scene->scan_incoherent = 0;
    if (scene->precision != B2RT_PRECISION_F32 || scene->semantics != B2RT_SEM_NUMBA || n <= 0 || n > B2RT_SCAN_MAX_PRIMS) {
        // float64 parity scenes and large scenes use the generic streams / the LBVH walk; small float64 scenes still scan
        scene->scan_incoherent = (n > 0 && n <= B2RT_SCAN_MAX_PRIMS) ? 1 : 0;
        scene->bounds_lo[0] = 1.f; scene->bounds_hi[0] = -1.f;
        return 0;
    }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    std::vector<float> rect(16 * (size_t)n_rect + 1), sphere(8 * (size_t)n_sphere + 1), tri(12 * (size_t)n_tri + 1),
        shade(12 * (size_t)n + 1), mat(8 * (size_t)std::max(1, scene->n_mat)), lights(4 * (size_t)std::max(1, scene->n_lights));
    std::vector<int32_t> prim_mat(n), mat_tex(std::max(1, scene->n_mat));
    cudaError_t e = cudaSuccess;
    auto get = [&](void *dst, const void *src, size_t bytes) { if (!e && bytes && src) e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, st); };
    get(rect.data(), scene->d_rect, 64 * (size_t)n_rect); get(sphere.data(), scene->d_sphere, 32 * (size_t)n_sphere);
    get(tri.data(), scene->d_tri, 48 * (size_t)n_tri); get(shade.data(), scene->d_shade, 48 * (size_t)n);
    get(mat.data(), scene->d_mat, 32 * (size_t)scene->n_mat); get(lights.data(), scene->d_lights, 16 * (size_t)scene->n_lights);
    get(prim_mat.data(), scene->d_prim_mat, 4 * (size_t)n); get(mat_tex.data(), scene->d_mat_tex, 4 * (size_t)scene->n_mat);
    if (!e) e = cudaStreamSynchronize(st);
    if (e) return set_error(1, "scene_prepare: %s", cudaGetErrorString(e));
    size_t need = 0;
    b2rt_scene_prepare_bytes(n_rect, n_sphere, n_tri, scene->n_lights, &need);
    if (!d_buffer || buffer_bytes < need) return set_error(2, "scene_prepare: buffer of %zu bytes, need %zu", buffer_bytes, need);
    std::vector<char> host(need);
    b2rt_prepare_layout lay;
    if (int rc = b2rt_scene_prepare_host(n_rect, n_sphere, n_tri, scene->n_mat, scene->n_lights, rect.data(), sphere.data(), tri.data(),
                                         shade.data(), mat.data(), prim_mat.data(), mat_tex.data(), lights.data(), flags,
                                         host.data(), host.size(), &lay))
        return rc;
    for (int k = 0; k < 3; ++k) { scene->bounds_lo[k] = lay.bounds_lo[k]; scene->bounds_hi[k] = lay.bounds_hi[k]; }
    scene->scan_incoherent = 1;
    if (lay.scan_offset == (size_t)-1) return 0;
    e = cudaMemcpyAsync(d_buffer, host.data(), lay.bytes_used, cudaMemcpyHostToDevice, st);
    if (!e) e = cudaStreamSynchronize(st);                          // `host` goes out of scope
    if (e) return set_error(1, "scene_prepare: %s", cudaGetErrorString(e));
    char *d = (char *)d_buffer;
    scene->d_scan_prims = d + lay.scan_offset;
    scene->n_scan_prims = lay.n_scan_prims; scene->n_scan_loose = lay.n_scan_loose; scene->n_scan_boxes = lay.n_scan_boxes;
    if (lay.surface_offset != (size_t)-1) scene->d_surface_records = d + lay.surface_offset;
    if (lay.hint_offset != (size_t)-1) scene->d_occluder_hint = (const int32_t *)(d + lay.hint_offset);
    return 0;
}
