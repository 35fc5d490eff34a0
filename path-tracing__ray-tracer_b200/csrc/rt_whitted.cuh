// rt_whitted.cuh — primary-hit query and the two deterministic Whitted integrators.
//   primary_hits_kernel      : check (a) of the north star (primary-ray primitive ids)
//   trace_rays_kernel        : explicit-ray closest/any hit (unit parity against cuda_scene_hit)
//   whitted_cpu_kernel       : CPURenderer._trace, renderers/cpu_renderer.py:75-151 (recursion tree
//                              unrolled onto a per-thread stack of weighted rays)
//   whitted_texture_kernel   : cuda_trace_kernel + cuda_trace_ray, renderers/cuda_texture_renderer.py:17-73,173-430
// One thread per pixel: these images are small/deterministic parity paths; the throughput path is
// the wavefront path tracer in rt_path.cuh.
#pragma once
#include "rt_scene.cuh"

namespace b2rt {

extern __shared__ float4 smem_top[];

template <typename R, bool CpuSem>
__global__ void __launch_bounds__(128)
primary_hits_kernel(SceneDev S, Cam<R> cam, int W, int H, R du, R dv, R t_min, R t_max, int use_bvh,
                    int *ids, double *tt) {
    stage_top(S, smem_top);
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= W * H) return;
    int x = i % W, y = i / W;
    Ray<R> r = camera_ray<R>(cam, (R(x) + du) / R(W), (R(y) + dv) / R(H));
    Hit<R> h;
    bool ok = use_bvh ? traverse<R, CpuSem, false>(S, smem_top, r, t_min, t_max, h)
                      : scan_all<R, CpuSem, false>(S, r, t_min, t_max, h);
    ids[i] = ok ? h.prim : -1;
    if (tt) tt[i] = ok ? (double)h.t : -1.0;
}

template <typename R, bool CpuSem>
__global__ void __launch_bounds__(128)
trace_rays_kernel(SceneDev S, int n, const double *o, const double *d, R t_min, R t_max, int any_hit, int use_bvh,
                  int *ids, double *rec) {
    const bool planar = use_bvh == 2 && sizeof(R) == 4 && S.n_scan > 0;
    if (planar) stage_scan(S, smem_top); else stage_top(S, smem_top);
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Ray<R> r;
    r.o = {R(o[3 * i]), R(o[3 * i + 1]), R(o[3 * i + 2])};
    r.d = {R(d[3 * i]), R(d[3 * i + 1]), R(d[3 * i + 2])};
    Hit<R> h;
    bool ok;
    if constexpr (sizeof(R) == 4) {
        if (planar) {
            ok = any_hit ? scan_small<true>(S, smem_top, r, t_min, t_max, h) : scan_small<false>(S, smem_top, r, t_min, t_max, h);
            use_bvh = -1;
        }
    }
    if (use_bvh < 0) {}
    else if (any_hit) ok = use_bvh ? traverse<R, CpuSem, true>(S, smem_top, r, t_min, t_max, h)
                              : scan_all<R, CpuSem, true>(S, r, t_min, t_max, h);
    else ok = use_bvh ? traverse<R, CpuSem, false>(S, smem_top, r, t_min, t_max, h)
                      : scan_all<R, CpuSem, false>(S, r, t_min, t_max, h);
    ids[i] = ok ? h.prim : -1;
    if (rec) {
        double *q = rec + 9 * (size_t)i;
        if (ok && !any_hit) {
            Surface<R> sf;
            make_surface<R, CpuSem>(S, r, h, sf);
            q[0] = h.t; q[1] = sf.p.x; q[2] = sf.p.y; q[3] = sf.p.z;
            q[4] = sf.n.x; q[5] = sf.n.y; q[6] = sf.n.z; q[7] = sf.u; q[8] = sf.v;
        } else {
            q[0] = ok ? (double)h.t : -1.0;
            for (int k = 1; k < 9; ++k) q[k] = 0.0;
        }
    }
}

// ------------------------------------------------------------------ CPURenderer._trace
constexpr int kWhittedStack = 24;      // >= max_depth + 2 pending children; host rejects max_depth > 20

template <typename R>
__global__ void __launch_bounds__(128)
whitted_cpu_kernel(SceneDev S, Cam<R> cam, int W, int H, const double *jitter, int max_depth,
                   V3<R> ambient, V3<R> light_color, double *rgb) {
    stage_top(S, smem_top);
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= W * H) return;
    int x = i % W, y = i / W;
    R du = jitter ? R(jitter[2 * i]) : R(0.5), dv = jitter ? R(jitter[2 * i + 1]) : R(0.5);

    // pending rays: (origin, direction, weight, depth); colour = sum over tree nodes of
    // weight * local(node) * (1 - refl - refr)   (cpu_renderer.py:144-147 is linear in its children)
    V3<R> so[kWhittedStack], sd[kWhittedStack];
    R sw[kWhittedStack];
    int sdep[kWhittedStack];
    int sp = 0;
    Ray<R> r0 = camera_ray<R>(cam, (R(x) + du) / R(W), (R(y) + dv) / R(H));
    so[0] = r0.o; sd[0] = r0.d; sw[0] = R(1); sdep[0] = 0; sp = 1;
    V3<R> col = {R(0), R(0), R(0)};
    const R inf = R(1) / R(0);
    const R nl = R(S.n_lights);

    while (sp > 0) {
        --sp;
        Ray<R> r; r.o = so[sp]; r.d = sd[sp];
        R wgt = sw[sp];
        int depth = sdep[sp];
        Hit<R> h;
        if (!traverse<R, true, false>(S, smem_top, r, R(1e-3), inf, h)) continue;   // background: black (:150)
        Surface<R> sf;
        make_surface<R, true>(S, r, h, sf);
        V3<R> base = base_color<R, true>(S, sf);
        V3<R> local = had(base * sf.diffuse, ambient);                               // :88
        for (int li = 0; li < S.n_lights; ++li) {                                    // :92-111
            V3<R> L = xyz<R>(ldg4(reinterpret_cast<const real4<R> *>(S.lights) + li));
            V3<R> to_l = normalize(L - sf.p);
            Ray<R> sr; sr.o = sf.p + sf.n * R(1e-3); sr.d = normalize(to_l);         // Ray() renormalises
            R dist = length(L - sf.p);
            Hit<R> sh;
            if (!traverse<R, true, true>(S, smem_top, sr, R(1e-3), dist, sh)) {
                R diff = max_(dot(sf.n, to_l), R(0));
                local = local + (had(base * sf.diffuse, light_color) * diff) / nl;
                V3<R> view = normalize(r.o - sf.p);
                V3<R> rdir = to_l - sf.n * (R(2) * dot(to_l, sf.n));
                R spec = max_(dot(view, rdir), R(0));
                local = local + (light_color * (sf.specular * pow_(spec, R(32)))) / nl;
            }
        }
        col = col + local * (wgt * (R(1) - sf.reflective - sf.refractive));

        if (depth < max_depth) {
            R dn = dot(r.d, sf.n);
            V3<R> refl_d = normalize(r.d - sf.n * (R(2) * dn));
            V3<R> off_p = sf.p + sf.n * R(1e-3);
            if (sf.refractive > R(0)) {                                              // :121-142
                V3<R> on; R eta;
                if (dn > R(0)) { on = -sf.n; eta = sf.ior; } else { on = sf.n; eta = R(1) / sf.ior; }
                V3<R> uv = normalize(r.d);
                R dt = dot(uv, on);
                R disc = R(1) - eta * eta * (R(1) - dt * dt);
                if (sp < kWhittedStack) {
                    if (disc > R(0)) {
                        V3<R> rd = (uv - on * dt) * eta - on * sqrt_(disc);
                        so[sp] = sf.p - sf.n * R(1e-3); sd[sp] = normalize(rd);
                    } else { so[sp] = off_p; sd[sp] = refl_d; }
                    sw[sp] = wgt * sf.refractive; sdep[sp] = depth + 1; ++sp;
                }
            }
            if (sf.reflective > R(0) && sp < kWhittedStack) {                        // :114-117
                so[sp] = off_p; sd[sp] = refl_d; sw[sp] = wgt * sf.reflective; sdep[sp] = depth + 1; ++sp;
            }
        }
    }
    rgb[3 * (size_t)i] = col.x; rgb[3 * (size_t)i + 1] = col.y; rgb[3 * (size_t)i + 2] = col.z;
}

// ------------------------------------------------------------------ cuda_trace_ray (textured Whitted)
// SMALL (float32, small scenes): closest hits and the 16 shadow rays per hit scan the box / planar records and
// shading reads the surface records, all staged in shared memory (s_top then points at the scan records);
// otherwise the LBVH is walked with its top levels in s_top.
template <typename R, bool SMALL>
__device__ __forceinline__ V3<R> trace_ray_texture(const SceneDev &S, const float4 *s_top, const float4 *s_surf, Ray<R> r,
                                                   int max_depth) {
    V3<R> col = {R(0), R(0), R(0)}, att = {R(1), R(1), R(1)};
    const R nl = R(S.n_lights);
    for (int depth = 0; depth < max_depth; ++depth) {
        Hit<R> h;
        Surface<R> sf;
        if constexpr (SMALL && sizeof(R) == 4) {
            if (!scan_small<false>(S, s_top, r, 0.001f, 1000000.0f, h)) break;
            make_surface_small(s_surf, r, h, sf);
        } else {
            if (!traverse<R, false, false>(S, s_top, r, R(0.001), R(1000000.0), h)) break;
            make_surface<R, false>(S, r, h, sf);
        }
        V3<R> mc = base_color<R, false>(S, sf);
        V3<R> local = mc * R(0.4);                                                    // :222-225
        if (S.n_lights > 0) {
            V3<R> dc = {R(0), R(0), R(0)}, sc = {R(0), R(0), R(0)};
            for (int li = 0; li < S.n_lights; ++li) {                                 // :238-330
                V3<R> L = xyz<R>(ldg4(reinterpret_cast<const real4<R> *>(S.lights) + li));
                V3<R> l = L - sf.p;
                R ld = length(l);
                if (!(ld > R(0.001))) continue;
                l = l / ld;
                Ray<R> sr; sr.o = sf.p + sf.n * R(0.001); sr.d = l;
                Hit<R> sh;
                if constexpr (SMALL && sizeof(R) == 4) {
                    if (scan_small<true>(S, s_top, sr, 0.001f, ld - 0.001f, sh)) continue;
                } else if (traverse<R, false, true>(S, s_top, sr, R(0.001), ld - R(0.001), sh)) continue;
                R ndl = sf.n.x * l.x + sf.n.y * l.y + sf.n.z * l.z;
                R df = max_(R(0), ndl);
                R atten = R(1.5) / (R(1.0) + R(0.001) * ld + R(0.0001) * ld * ld);
                R di = df * atten / nl;
                dc.x += mc.x * di * sf.diffuse * R(0.6);
                dc.y += mc.y * di * sf.diffuse * R(0.6);
                dc.z += mc.z * di * sf.diffuse * R(0.6);
                if (sf.specular > R(0.01) && df > R(0)) {
                    V3<R> rf = {R(2) * ndl * sf.n.x - l.x, R(2) * ndl * sf.n.y - l.y, R(2) * ndl * sf.n.z - l.z};
                    V3<R> vw = -r.d;
                    R rv = max_(R(0), rf.x * vw.x + rf.y * vw.y + rf.z * vw.z);
                    R shin = R(32), smul = R(1);
                    if (sf.reflective > R(0.9) && sf.specular > R(0.9)) { shin = R(256); smul = R(1.5); }
                    else if (sf.reflective > R(0.7)) { shin = R(128); smul = R(1.2); }
                    else if (sf.specular > R(0.5)) { shin = R(64); }
                    R si = pow_(rv, shin) * atten * smul / nl;
                    if (sf.reflective > R(0.7)) {
                        sc.x += si * sf.specular * mc.x; sc.y += si * sf.specular * mc.y; sc.z += si * sf.specular * mc.z;
                    } else {
                        sc.x += si * sf.specular; sc.y += si * sf.specular; sc.z += si * sf.specular;
                    }
                }
            }
            local = {local.x + (dc.x + sc.x), local.y + (dc.y + sc.y), local.z + (dc.z + sc.z)};
        }
        R base = max_(R(0.1), R(1) - sf.reflective - sf.refractive);                  // :338
        col.x += local.x * att.x * base; col.y += local.y * att.y * base; col.z += local.z * att.z * base;

        if (!((sf.reflective > R(0.01) || sf.refractive > R(0.01)) && depth < max_depth - 1)) break;   // :344
        bool use_refr = sf.refractive > sf.reflective && sf.refractive > R(0.1);
        R dn = r.d.x * sf.n.x + r.d.y * sf.n.y + r.d.z * sf.n.z;
        bool done = false;
        if (use_refr) {
            V3<R> on, offd, rd; R eta;
            if (dn > R(0)) { on = -sf.n; eta = sf.ior; offd = sf.n; }
            else { on = sf.n; eta = R(1) / sf.ior; offd = -sf.n; }
            if (refract_nb<R>(r.d, on, eta, rd)) {
                r.o = sf.p + offd * R(0.001); r.d = rd;
                R k = sf.refractive * R(0.95);
                att = att * k;
                done = true;
            }
        }
        if (!done) {                     // total internal reflection (:384-403) == reflection (:404-423)
            V3<R> rd = {r.d.x - R(2) * dn * sf.n.x, r.d.y - R(2) * dn * sf.n.y, r.d.z - R(2) * dn * sf.n.z};
            r.o = sf.p + sf.n * R(0.001); r.d = rd;
            att = att * sf.reflective;
        }
    }
    return col;
}

__device__ __forceinline__ uint8_t quant8(double c) {      // min(255, max(0, int(c * 255)))
    long long q = (long long)(c * 255.0);
    return (uint8_t)(q < 0 ? 0 : (q > 255 ? 255 : q));
}

template <typename R, bool SMALL>
__global__ void __launch_bounds__(128)
whitted_texture_kernel(SceneDev S, Cam<R> cam, int W, int H, int spp, int max_depth, double *rgb, uint8_t *u8) {
    const float4 *s_surf = nullptr;
    if (SMALL) {
        stage_scan(S, smem_top);
        stage_surf(S, smem_top + 4 * (S.n_scan + S.n_box));
        s_surf = smem_top + 4 * (S.n_scan + S.n_box);
    } else stage_top(S, smem_top);
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= W * H) return;
    int x = i % W, y = i / W;
    long long rng = (long long)x + (long long)y * W + 1;                               // :32
    int grid_n = (int)sqrt((double)spp);
    V3<R> c = {R(0), R(0), R(0)};
    for (int a = 0; a < grid_n; ++a)
        for (int b = 0; b < grid_n; ++b) {
            // cuda_random (:76-80) leaves the caller's state alone, so du and dv share one draw
            long long rr = (rng * 1103515245LL + 12345LL) & 0x7fffffffLL;
            R rnd = R((double)rr / 2147483647.0);
            R du = (R(a) + rnd) / R(grid_n), dv = (R(b) + rnd) / R(grid_n);
            rng = (rng * 1103515245LL + 12345LL) & 0x7fffffffLL;
            Ray<R> r = camera_ray<R>(cam, (R(x) + du) / R(W), (R(y) + dv) / R(H));
            V3<R> s = trace_ray_texture<R, SMALL>(S, smem_top, s_surf, r, max_depth);
            c = c + s;
            rng = (rng * 1103515245LL + 12345LL) & 0x7fffffffLL;
        }
    c = c / R(spp);
    if (rgb) { rgb[3 * (size_t)i] = c.x; rgb[3 * (size_t)i + 1] = c.y; rgb[3 * (size_t)i + 2] = c.z; }
    if (u8) { u8[3 * (size_t)i] = quant8(c.x); u8[3 * (size_t)i + 1] = quant8(c.y); u8[3 * (size_t)i + 2] = quant8(c.z); }
}

}  // namespace b2rt
