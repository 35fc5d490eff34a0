/*
 * rt_oracle.c — CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C, double-precision restatement of the reference renderers'
 * algorithms (enginism/Path-Tracing__ray-tracer), used ONLY as the checker in
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs.  Nothing under path-tracing__ray-tracer_b200/ may link, import or call it.
 *
 * Parity pin: every entry point below is checked against outputs of the
 * reference's own Python code run in the build container (oracle/make_golden.py
 * -> tests/golden/, tests/test_oracle_golden.py), including the reference's only
 * golden vector, output_RayTracer.png.  See oracle/README.md.
 *
 * Two families, mirroring the reference's two arithmetic regimes:
 *   "nb_*"  : the numba renderers — float64 math on the float32-packed AoS scene
 *             block built by _prepare_scene_data (cuda_path_tracer.py:819-899).
 *   "cpu_*" : renderers/cpu_renderer.py — float64 on un-rounded object data,
 *             closest hit through the random-axis median BVH (core/acceleration.py).
 *
 * Build: gcc -O2 -fno-fast-math -ffp-contract=off -fopenmp -shared -fPIC (oracle/build.sh)
 *        -ffp-contract=off matters: the reference never fuses a*b+c.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------- */
/* shared helpers                                                              */
/* ------------------------------------------------------------------------- */
typedef struct { double x, y, z; } v3;

static inline v3 V(double x, double y, double z) { v3 r = {x, y, z}; return r; }
static inline v3 vadd(v3 a, v3 b) { return V(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 vsub(v3 a, v3 b) { return V(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 vmul(v3 a, double k) { return V(a.x * k, a.y * k, a.z * k); }
static inline v3 vhad(v3 a, v3 b) { return V(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline v3 vdiv(v3 a, double k) { return V(a.x / k, a.y / k, a.z / k); }
static inline v3 vneg(v3 a) { return V(-a.x, -a.y, -a.z); }
static inline double vdot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline v3 vcross(v3 a, v3 b) {
    return V(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
static inline double vlen(v3 a) { return sqrt(a.x * a.x + a.y * a.y + a.z * a.z); }
/* Vec3.normalize, core/math.py:49-53: divide by the length, zero vector stays zero */
static inline v3 vnorm(v3 a) { double l = vlen(a); return l == 0 ? V(0, 0, 0) : vdiv(a, l); }
static inline double dmax(double a, double b) { return a > b ? a : b; }
static inline double dmin(double a, double b) { return a < b ? a : b; }

/* int(x) of a Python/numba float: truncation toward zero */
static inline long trunc_l(double x) { return (long)x; }

/* ------------------------------------------------------------------------- */
/* numba family: packed float32 scene block                                    */
/* ------------------------------------------------------------------------- */
typedef struct {
    int hit;
    double t;
    double p[3], n[3], mat[10], uv[2];
    int prim;            /* packed index: planes, then spheres, then triangles (not in the reference) */
} nb_hit;

typedef struct {
    const float *scene;  /* [nP, 20*nP, nS, 12*nS, nT, 26*nT] */
    const float *cam;    /* 12 */
    const float *lights; /* [n, xyz*n] */
    int n_light_floats;
    const uint8_t *tex;
    long n_tex_bytes;
    const int32_t *tex_info;
    int n_tex_info;
} nb_scene;

/* cuda_scene_hit, cuda_path_tracer.py:496-730 (== cuda_texture_renderer.py:433-704) */
static void nb_scene_hit(const float *sd, const double o[3], const double d[3],
                         double t_min, double t_max, nb_hit *h)
{
    static const double dflt_mat[10] = {0.5, 0.5, 0.5, 0.8, 0.2, 0.0, 0.0, 1.0, 0.0, -1.0};
    double closest = t_max;
    h->hit = 0; h->prim = -1;
    h->p[0] = h->p[1] = h->p[2] = 0.0;
    h->n[0] = 0.0; h->n[1] = 1.0; h->n[2] = 0.0;
    memcpy(h->mat, dflt_mat, sizeof dflt_mat);
    h->uv[0] = h->uv[1] = 0.0;

    int off = 0, prim = 0;
    int nP = (int)sd[off]; off += 1;
    for (int i = 0; i < nP; ++i, ++prim) {                      /* :511-574 */
        const float *q = sd + off + i * 20;
        double ax = q[0], ay = q[1], az = q[2], nx = q[3], ny = q[4], nz = q[5];
        double ul = q[12], vl = q[13];
        double denom = nx * d[0] + ny * d[1] + nz * d[2];
        if (fabs(denom) > 1e-6) {
            double dx = ax - o[0], dy = ay - o[1], dz = az - o[2];
            double t = (dx * nx + dy * ny + dz * nz) / denom;
            if (t_min < t && t < closest) {
                double hx = o[0] + t * d[0], hy = o[1] + t * d[1], hz = o[2] + t * d[2];
                double rx = hx - ax, ry = hy - ay, rz = hz - az;
                /* numba types f32 (+,-,*,/) f32 and math.sqrt(f32) as FLOAT32: the axis
                 * normalisation below runs in single precision in the reference (:552-562) */
                float ulen = sqrtf(q[6] * q[6] + q[7] * q[7] + q[8] * q[8]);
                float vlen_ = sqrtf(q[9] * q[9] + q[10] * q[10] + q[11] * q[11]);
                if (ulen > 0 && vlen_ > 0) {
                    double uux = q[6] / ulen, uuy = q[7] / ulen, uuz = q[8] / ulen;
                    double vvx = q[9] / vlen_, vvy = q[10] / vlen_, vvz = q[11] / vlen_;
                    double uh = rx * uux + ry * uuy + rz * uuz;
                    double vh = rx * vvx + ry * vvy + rz * vvz;
                    if (0 <= uh && uh <= ul && 0 <= vh && vh <= vl) {
                        closest = t; h->hit = 1; h->prim = prim;
                        h->p[0] = hx; h->p[1] = hy; h->p[2] = hz;
                        h->n[0] = nx; h->n[1] = ny; h->n[2] = nz;
                        h->uv[0] = uh / ul; h->uv[1] = vh / vl;
                        h->mat[0] = q[14]; h->mat[1] = q[15]; h->mat[2] = q[16];
                        h->mat[3] = q[17]; h->mat[4] = q[18]; h->mat[5] = q[19];
                        h->mat[6] = 0.0; h->mat[7] = 1.0; h->mat[8] = 0.0; h->mat[9] = -1.0;
                    }
                }
            }
        }
    }
    off += nP * 20;

    int nS = (int)sd[off]; off += 1;
    for (int i = 0; i < nS; ++i, ++prim) {                      /* :582-631 */
        const float *q = sd + off + i * 12;
        double cx = q[0], cy = q[1], cz = q[2], r = q[3];
        double ocx = o[0] - cx, ocy = o[1] - cy, ocz = o[2] - cz;
        double a = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
        double b = ocx * d[0] + ocy * d[1] + ocz * d[2];
        float r2 = q[3] * q[3];                    /* radius * radius is an f32*f32 product (:604) */
        double c = (ocx * ocx + ocy * ocy + ocz * ocz) - r2;
        double disc = b * b - a * c;
        if (disc > 0) {
            double s = sqrt(disc);
            double t1 = (-b - s) / a, t2 = (-b + s) / a;
            double t = (t_min < t1 && t1 < closest) ? t1 : ((t_min < t2 && t2 < closest) ? t2 : -1);
            if (t > 0) {
                closest = t; h->hit = 1; h->prim = prim;
                double hx = o[0] + t * d[0], hy = o[1] + t * d[1], hz = o[2] + t * d[2];
                h->p[0] = hx; h->p[1] = hy; h->p[2] = hz;
                h->n[0] = (hx - cx) / r; h->n[1] = (hy - cy) / r; h->n[2] = (hz - cz) / r;
                h->uv[0] = h->uv[1] = 0.0;
                h->mat[0] = q[4]; h->mat[1] = q[5]; h->mat[2] = q[6]; h->mat[3] = q[7];
                h->mat[4] = q[8]; h->mat[5] = q[9]; h->mat[6] = q[10]; h->mat[7] = q[11];
                h->mat[8] = 0.0; h->mat[9] = -1.0;
            }
        }
    }
    off += nS * 12;

    int nT = (int)sd[off]; off += 1;
    for (int i = 0; i < nT; ++i, ++prim) {                      /* :639-728 */
        const float *q = sd + off + i * 26;
        double v0x = q[0], v0y = q[1], v0z = q[2];
        /* edge = v1 - v0 on two f32 values is an f32 subtraction in numba (:669-675) */
        float e1xf = q[3] - q[0], e1yf = q[4] - q[1], e1zf = q[5] - q[2];
        float e2xf = q[6] - q[0], e2yf = q[7] - q[1], e2zf = q[8] - q[2];
        double e1x = e1xf, e1y = e1yf, e1z = e1zf, e2x = e2xf, e2y = e2yf, e2z = e2zf;
        double nx = q[9], ny = q[10], nz = q[11];
        double hx_ = d[1] * e2z - d[2] * e2y;
        double hy_ = d[2] * e2x - d[0] * e2z;
        double hz_ = d[0] * e2y - d[1] * e2x;
        double a = e1x * hx_ + e1y * hy_ + e1z * hz_;
        if (fabs(a) < 1e-6) continue;
        double f = 1.0 / a;
        double sx = o[0] - v0x, sy = o[1] - v0y, sz = o[2] - v0z;
        double u = f * (sx * hx_ + sy * hy_ + sz * hz_);
        if (u < 0.0 || u > 1.0) continue;
        double qx = sy * e1z - sz * e1y, qy = sz * e1x - sx * e1z, qz = sx * e1y - sy * e1x;
        double v = f * (d[0] * qx + d[1] * qy + d[2] * qz);
        if (v < 0.0 || u + v > 1.0) continue;
        double t = f * (e2x * qx + e2y * qy + e2z * qz);
        if (t_min < t && t < closest) {
            closest = t; h->hit = 1; h->prim = prim;
            h->p[0] = o[0] + t * d[0]; h->p[1] = o[1] + t * d[1]; h->p[2] = o[2] + t * d[2];
            double dp = nx * d[0] + ny * d[1] + nz * d[2];
            if (dp > 0) { nx = -nx; ny = -ny; nz = -nz; }
            h->n[0] = nx; h->n[1] = ny; h->n[2] = nz;
            double w = 1.0 - u - v;
            h->uv[0] = w * q[20] + u * q[22] + v * q[24];
            h->uv[1] = w * q[21] + u * q[23] + v * q[25];
            h->mat[0] = q[12]; h->mat[1] = q[13]; h->mat[2] = q[14]; h->mat[3] = q[15];
            h->mat[4] = q[16]; h->mat[5] = q[17]; h->mat[6] = 0.0; h->mat[7] = 1.0;
            h->mat[8] = q[18]; h->mat[9] = q[19];
        }
    }
    h->t = closest;
}

/* cuda_get_ray, cuda_path_tracer.py:84-112 */
static void nb_get_ray(const float *cam, double u, double v, double o[3], double d[3])
{
    o[0] = cam[0]; o[1] = cam[1]; o[2] = cam[2];
    d[0] = cam[3] + u * cam[6] + v * cam[9] - o[0];
    d[1] = cam[4] + u * cam[7] + v * cam[10] - o[1];
    d[2] = cam[5] + u * cam[8] + v * cam[11] - o[2];
    double l = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    if (l > 0) { d[0] /= l; d[1] /= l; d[2] /= l; }
}

/* cuda_sample_texture, cuda_path_tracer.py:473-493 */
static void nb_sample_texture(const nb_scene *s, long start, long w, long h_, double u, double v, double rgb[3])
{
    u = dmax(0.0, dmin(1.0, u));
    v = dmax(0.0, dmin(1.0, v));
    long iu = trunc_l(u * (double)(w - 1));
    long iv = trunc_l((1.0 - v) * (double)(h_ - 1));
    if (iu > w - 1) iu = w - 1;
    if (iu < 0) iu = 0;
    if (iv > h_ - 1) iv = h_ - 1;
    if (iv < 0) iv = 0;
    long base = start + (iv * w + iu) * 3;
    if (base + 2 < s->n_tex_bytes) {
        rgb[0] = s->tex[base] / 255.0; rgb[1] = s->tex[base + 1] / 255.0; rgb[2] = s->tex[base + 2] / 255.0;
    } else {
        rgb[0] = rgb[1] = rgb[2] = 1.0;
    }
}

static void nb_apply_texture(const nb_scene *s, const nb_hit *h, double col[3])
{
    col[0] = h->mat[0]; col[1] = h->mat[1]; col[2] = h->mat[2];
    int has_tex = h->mat[8] > 0.5;
    long tid = trunc_l(h->mat[9]);
    if (has_tex && tid >= 0 && tid < s->n_tex_info / 3)
        nb_sample_texture(s, s->tex_info[tid * 3], s->tex_info[tid * 3 + 1], s->tex_info[tid * 3 + 2],
                          h->uv[0], h->uv[1], col);
}

/* cuda_refract / cuda_refract_path, cuda_path_tracer.py:115-131 */
static int nb_refract(const double in[3], const double n[3], double eta, double out[3])
{
    double cos_i = -(in[0] * n[0] + in[1] * n[1] + in[2] * n[2]);
    double sin2_t = eta * eta * (1.0 - cos_i * cos_i);
    if (sin2_t > 1.0) return 0;
    double cos_t = sqrt(1.0 - sin2_t);
    double f2 = eta * cos_i - cos_t;
    out[0] = eta * in[0] + f2 * n[0];
    out[1] = eta * in[1] + f2 * n[1];
    out[2] = eta * in[2] + f2 * n[2];
    return 1;
}

/* ---- textured Whitted: cuda_texture_renderer.py ---- */

/* cuda_trace_ray, cuda_texture_renderer.py:173-430 */
static void nb_trace_ray(const nb_scene *s, const double o_in[3], const double d_in[3], int max_depth,
                         double rgb[3], uint64_t *n_hit_calls)
{
    double col[3] = {0, 0, 0}, att[3] = {1, 1, 1};
    double o[3] = {o_in[0], o_in[1], o_in[2]}, d[3] = {d_in[0], d_in[1], d_in[2]};
    uint64_t calls = 0;
    for (int depth = 0; depth < max_depth; ++depth) {
        nb_hit h;
        nb_scene_hit(s->scene, o, d, 0.001, 1000000.0, &h); ++calls;
        if (!h.hit) break;
        double mc[3];
        nb_apply_texture(s, &h, mc);
        double m_diff = h.mat[3], m_spec = h.mat[4], m_refl = h.mat[5], m_refr = h.mat[6], m_ior = h.mat[7];
        double local[3] = {mc[0] * 0.4, mc[1] * 0.4, mc[2] * 0.4};              /* :222-225 */
        long nl = trunc_l(s->lights[0]);
        if (nl > 0) {
            double dc[3] = {0, 0, 0}, sc[3] = {0, 0, 0};
            for (long i = 0; i < nl; ++i) {                                      /* :238-330 */
                const float *L = s->lights + 1 + i * 3;
                double lx = L[0] - h.p[0], ly = L[1] - h.p[1], lz = L[2] - h.p[2];
                double ld = sqrt(lx * lx + ly * ly + lz * lz);
                if (ld > 0.001) {
                    lx /= ld; ly /= ld; lz /= ld;
                    double so[3] = {h.p[0] + h.n[0] * 0.001, h.p[1] + h.n[1] * 0.001, h.p[2] + h.n[2] * 0.001};
                    double sdv[3] = {lx, ly, lz};
                    nb_hit sh;
                    nb_scene_hit(s->scene, so, sdv, 0.001, ld - 0.001, &sh); ++calls;
                    if (!sh.hit) {
                        double df = dmax(0.0, h.n[0] * lx + h.n[1] * ly + h.n[2] * lz);
                        double atten = 1.5 / (1.0 + 0.001 * ld + 0.0001 * ld * ld);
                        double di = df * atten / (double)nl;
                        dc[0] += mc[0] * di * m_diff * 0.6;
                        dc[1] += mc[1] * di * m_diff * 0.6;
                        dc[2] += mc[2] * di * m_diff * 0.6;
                        if (m_spec > 0.01 && df > 0.0) {
                            double nl_ = h.n[0] * lx + h.n[1] * ly + h.n[2] * lz;
                            double rx = 2.0 * nl_ * h.n[0] - lx, ry = 2.0 * nl_ * h.n[1] - ly, rz = 2.0 * nl_ * h.n[2] - lz;
                            double vx = -d[0], vy = -d[1], vz = -d[2];
                            double rv = dmax(0.0, rx * vx + ry * vy + rz * vz);
                            double shin = 32.0, smul = 1.0;
                            if (m_refl > 0.9 && m_spec > 0.9) { shin = 256.0; smul = 1.5; }
                            else if (m_refl > 0.7) { shin = 128.0; smul = 1.2; }
                            else if (m_spec > 0.5) { shin = 64.0; }
                            double sf = pow(rv, shin);
                            double si = sf * atten * smul / (double)nl;
                            if (m_refl > 0.7) {
                                sc[0] += si * m_spec * mc[0]; sc[1] += si * m_spec * mc[1]; sc[2] += si * m_spec * mc[2];
                            } else {
                                sc[0] += si * m_spec; sc[1] += si * m_spec; sc[2] += si * m_spec;
                            }
                        }
                    }
                }
            }
            local[0] += dc[0] + sc[0]; local[1] += dc[1] + sc[1]; local[2] += dc[2] + sc[2];
        }
        double base = dmax(0.1, 1.0 - m_refl - m_refr);                          /* :338 */
        col[0] += local[0] * att[0] * base; col[1] += local[1] * att[1] * base; col[2] += local[2] * att[2] * base;

        if ((m_refl > 0.01 || m_refr > 0.01) && depth < max_depth - 1) {          /* :344 */
            int use_refr = (m_refr > m_refl) && (m_refr > 0.1);
            double dn = d[0] * h.n[0] + d[1] * h.n[1] + d[2] * h.n[2];
            int refracted_ok = 0;
            if (use_refr) {
                double on[3], offd[3], eta, r[3];
                if (dn > 0) { on[0] = -h.n[0]; on[1] = -h.n[1]; on[2] = -h.n[2]; eta = m_ior;
                              offd[0] = h.n[0]; offd[1] = h.n[1]; offd[2] = h.n[2]; }
                else { on[0] = h.n[0]; on[1] = h.n[1]; on[2] = h.n[2]; eta = 1.0 / m_ior;
                       offd[0] = -h.n[0]; offd[1] = -h.n[1]; offd[2] = -h.n[2]; }
                if (nb_refract(d, on, eta, r)) {
                    refracted_ok = 1;
                    o[0] = h.p[0] + offd[0] * 0.001; o[1] = h.p[1] + offd[1] * 0.001; o[2] = h.p[2] + offd[2] * 0.001;
                    d[0] = r[0]; d[1] = r[1]; d[2] = r[2];
                    double k = m_refr * 0.95;
                    att[0] *= k; att[1] *= k; att[2] *= k;
                }
            }
            if (!refracted_ok) {      /* TIR (:384-403) and plain reflection (:404-423) are the same code */
                double rx = d[0] - 2.0 * dn * h.n[0], ry = d[1] - 2.0 * dn * h.n[1], rz = d[2] - 2.0 * dn * h.n[2];
                o[0] = h.p[0] + h.n[0] * 0.001; o[1] = h.p[1] + h.n[1] * 0.001; o[2] = h.p[2] + h.n[2] * 0.001;
                d[0] = rx; d[1] = ry; d[2] = rz;
                att[0] *= m_refl; att[1] *= m_refl; att[2] *= m_refl;
            }
        } else break;
    }
    rgb[0] = col[0]; rgb[1] = col[1]; rgb[2] = col[2];
    if (n_hit_calls) *n_hit_calls += calls;
}

static inline uint8_t quant8(double c) {           /* min(255, max(0, int(c*255))) */
    long q = trunc_l(c * 255);
    if (q < 0) q = 0;
    if (q > 255) q = 255;
    return (uint8_t)q;
}

/* cuda_trace_kernel, cuda_texture_renderer.py:17-73.  Row 0 of the outputs is the BOTTOM row
 * (device order); the host flips (:782).  out_f (optional) receives the pre-quantisation mean. */
ORC_API void orc_nb_whitted_texture(const float *scene, const float *cam, const float *lights, int n_light_floats,
                                    const uint8_t *tex, long n_tex_bytes, const int32_t *tex_info, int n_tex_info,
                                    int width, int height, int spp, int max_depth,
                                    uint8_t *out_u8, double *out_f, uint64_t *n_hit_calls)
{
    nb_scene s = {scene, cam, lights, n_light_floats, tex, n_tex_bytes, tex_info, n_tex_info};
    int grid_n = (int)sqrt((double)spp);
    uint64_t total_calls = 0;
#pragma omp parallel for schedule(dynamic, 4) reduction(+ : total_calls)
    for (int y = 0; y < height; ++y)
        for (int x = 0; x < width; ++x) {
            int64_t rng = (int64_t)x + (int64_t)y * width + 1;                   /* :32 */
            double c[3] = {0, 0, 0};
            for (int a = 0; a < grid_n; ++a)
                for (int b = 0; b < grid_n; ++b) {
                    /* cuda_random (:76-80) does not advance the caller's state: du and dv share one draw */
                    int64_t r = (rng * 1103515245LL + 12345LL) & 0x7fffffffLL;
                    double rnd = (double)r / 2147483647.0;
                    double du = (a + rnd) / grid_n, dv = (b + rnd) / grid_n;
                    rng = (rng * 1103515245LL + 12345LL) & 0x7fffffffLL;
                    double u = (x + du) / width, v = (y + dv) / height;
                    double o[3], d[3], rgb[3];
                    nb_get_ray(cam, u, v, o, d);
                    nb_trace_ray(&s, o, d, max_depth, rgb, &total_calls);
                    c[0] += rgb[0]; c[1] += rgb[1]; c[2] += rgb[2];
                    rng = (rng * 1103515245LL + 12345LL) & 0x7fffffffLL;
                }
            c[0] /= spp; c[1] /= spp; c[2] /= spp;
            size_t pi = ((size_t)y * width + x) * 3;
            if (out_u8) { out_u8[pi] = quant8(c[0]); out_u8[pi + 1] = quant8(c[1]); out_u8[pi + 2] = quant8(c[2]); }
            if (out_f) { out_f[pi] = c[0]; out_f[pi + 1] = c[1]; out_f[pi + 2] = c[2]; }
        }
    if (n_hit_calls) *n_hit_calls = total_calls;
}

/* ---- path tracer: cuda_path_tracer.py ---- */

/* cuda_xorshift, :61-66 — int64 arithmetic: '<<' wraps, '>>' is an arithmetic shift, mask at the end */
static inline int64_t nb_xorshift(int64_t s)
{
    s ^= (int64_t)((uint64_t)s << 13);
    s ^= s >> 17;
    s ^= (int64_t)((uint64_t)s << 5);
    return s & 0xffffffffLL;
}
/* cuda_random, :69-71 */
static inline double nb_random(int64_t s) { return (double)(s & 0xffffffLL) / 16777216.0; }

ORC_API int64_t orc_nb_xorshift(int64_t s) { return nb_xorshift(s); }
ORC_API double orc_nb_random(int64_t s) { return nb_random(s); }

/* cuda_sample_hemisphere_cosine, :139-180 */
static void nb_cos_hemisphere(const double n[3], int64_t *rng, double out[3])
{
    double r1 = nb_random(*rng); *rng = nb_xorshift(*rng);
    double r2 = nb_random(*rng); *rng = nb_xorshift(*rng);
    double ct = sqrt(r1), st = sqrt(1.0 - r1), phi = 2.0 * M_PI * r2;
    double x = st * cos(phi), y = st * sin(phi), z = ct;
    double tx, ty, tz;
    if (fabs(n[2]) > 0.9) { tx = 1.0; ty = 0.0; tz = 0.0; } else { tx = 0.0; ty = 0.0; tz = 1.0; }
    double ux = ty * n[2] - tz * n[1], uy = tz * n[0] - tx * n[2], uz = tx * n[1] - ty * n[0];
    double ul = sqrt(ux * ux + uy * uy + uz * uz);
    ux /= ul; uy /= ul; uz /= ul;
    double vx = n[1] * uz - n[2] * uy, vy = n[2] * ux - n[0] * uz, vz = n[0] * uy - n[1] * ux;
    out[0] = x * ux + y * vx + z * n[0];
    out[1] = x * uy + y * vy + z * n[1];
    out[2] = x * uz + y * vz + z * n[2];
}

typedef struct { uint64_t closest_rays, shadow_rays, segments, nee_unshadowed; } path_counters;

/* cuda_trace_path, :215-471.  rng is by value (the caller's copy is not advanced, :40-41). */
static void nb_trace_path(const nb_scene *s, const double o_in[3], const double d_in[3], int max_depth,
                          int64_t rng, double rgb[3], path_counters *pc)
{
    double col[3] = {0, 0, 0}, thr[3] = {1, 1, 1};
    double o[3] = {o_in[0], o_in[1], o_in[2]}, d[3] = {d_in[0], d_in[1], d_in[2]};
    for (int depth = 0; depth < max_depth; ++depth) {
        nb_hit h;
        nb_scene_hit(s->scene, o, d, 0.001, 1000000.0, &h);
        if (pc) pc->closest_rays++;
        if (!h.hit) {                                                            /* :234-239 */
            col[0] += thr[0] * 0.1; col[1] += thr[1] * 0.1; col[2] += thr[2] * 0.1;
            break;
        }
        if (pc) pc->segments++;
        double mc[3];
        nb_apply_texture(s, &h, mc);
        double m_diff = h.mat[3], m_refl = h.mat[5], m_refr = h.mat[6], m_ior = h.mat[7];

        if (s->n_light_floats > 1) {                                             /* :265-304 */
            long nl = trunc_l(s->lights[0]);
            double ld[3] = {0, 0, 0}, pdf = 0.0;
            if (nl != 0) {                                                       /* :183-210 */
                long li = trunc_l(nb_random(rng) * (double)nl);
                if (li >= nl) li = nl - 1;
                rng = nb_xorshift(rng);
                const float *L = s->lights + 1 + li * 3;
                ld[0] = L[0] - h.p[0]; ld[1] = L[1] - h.p[1]; ld[2] = L[2] - h.p[2];
                double dist = sqrt(ld[0] * ld[0] + ld[1] * ld[1] + ld[2] * ld[2]);
                if (dist > 0.001) { ld[0] /= dist; ld[1] /= dist; ld[2] /= dist; }
                pdf = 1.0 / (double)nl;
            }
            if (pdf > 0.0) {
                double so[3] = {h.p[0] + h.n[0] * 0.001, h.p[1] + h.n[1] * 0.001, h.p[2] + h.n[2] * 0.001};
                nb_hit sh;
                nb_scene_hit(s->scene, so, ld, 0.001, 1000000.0, &sh);
                if (pc) pc->shadow_rays++;
                if (!sh.hit) {
                    if (pc) pc->nee_unshadowed++;
                    double ct = dmax(0.0, ld[0] * h.n[0] + ld[1] * h.n[1] + ld[2] * h.n[2]);
                    double li_, lm;
                    if (m_refr > 0.5) { li_ = 4.0; lm = 0.6; }
                    else if (m_refl > 0.7) { li_ = 2.5; lm = 0.8; }
                    else { li_ = 2.0; lm = 1.0; }
                    for (int k = 0; k < 3; ++k) {
                        double contrib = mc[k] * m_diff * ct * li_ * lm / pdf;
                        col[k] += thr[k] * contrib;
                    }
                }
            }
        }

        if (depth >= 3) {                                                        /* :307-314 */
            double p = dmax(0.1, 0.299 * thr[0] + 0.587 * thr[1] + 0.114 * thr[2]);
            if (nb_random(rng) > p) break;
            rng = nb_xorshift(rng);
            thr[0] /= p; thr[1] /= p; thr[2] /= p;
        }

        double choice = nb_random(rng);                                          /* :317-318 */
        rng = nb_xorshift(rng);
        double dn = d[0] * h.n[0] + d[1] * h.n[1] + d[2] * h.n[2];
        double po[3] = {h.p[0] + h.n[0] * 0.001, h.p[1] + h.n[1] * 0.001, h.p[2] + h.n[2] * 0.001};
        double refl[3] = {d[0] - 2.0 * dn * h.n[0], d[1] - 2.0 * dn * h.n[1], d[2] - 2.0 * dn * h.n[2]};

        if (m_refr > 0.1) {                                                      /* :320-428 */
            if (choice < 0.6) {
                double cos_i = dmax(0.0, -dn);
                int entering = cos_i > 0.0;
                double eta, on[3], r[3];
                if (entering) { eta = 1.0 / m_ior; on[0] = h.n[0]; on[1] = h.n[1]; on[2] = h.n[2]; }
                else { eta = m_ior; on[0] = -h.n[0]; on[1] = -h.n[1]; on[2] = -h.n[2]; }
                if (nb_refract(d, on, eta, r)) {
                    if (entering) {                                              /* :350-355 */
                        o[0] = h.p[0] - h.n[0] * 0.001; o[1] = h.p[1] - h.n[1] * 0.001; o[2] = h.p[2] - h.n[2] * 0.001;
                    } else {                                                     /* :356-361 */
                        o[0] = po[0]; o[1] = po[1]; o[2] = po[2];
                    }
                    d[0] = r[0]; d[1] = r[1]; d[2] = r[2];
                    double k = m_refr / 0.6;
                    thr[0] *= k; thr[1] *= k; thr[2] *= k;
                } else {
                    memcpy(o, po, sizeof po); memcpy(d, refl, sizeof refl);
                    thr[0] *= 0.9; thr[1] *= 0.9; thr[2] *= 0.9;
                }
            } else if (choice < 0.6 + 0.25) {
                memcpy(o, po, sizeof po); memcpy(d, refl, sizeof refl);
                thr[0] *= mc[0] * 0.9 / 0.25; thr[1] *= mc[1] * 0.9 / 0.25; thr[2] *= mc[2] * 0.9 / 0.25;
            } else {
                double nd[3];
                nb_cos_hemisphere(h.n, &rng, nd);
                memcpy(o, po, sizeof po); memcpy(d, nd, sizeof nd);
                thr[0] *= mc[0] * m_diff * 3.0 / 0.15; thr[1] *= mc[1] * m_diff * 3.0 / 0.15;
                thr[2] *= mc[2] * m_diff * 3.0 / 0.15;
            }
        } else if (m_refl > 0.5) {                                               /* :430-449 */
            memcpy(o, po, sizeof po); memcpy(d, refl, sizeof refl);
            thr[0] *= mc[0] * m_refl; thr[1] *= mc[1] * m_refl; thr[2] *= mc[2] * m_refl;
        } else {                                                                 /* :451-466 */
            double nd[3];
            nb_cos_hemisphere(h.n, &rng, nd);
            memcpy(o, po, sizeof po); memcpy(d, nd, sizeof nd);
            thr[0] *= mc[0] * m_diff; thr[1] *= mc[1] * m_diff; thr[2] *= mc[2] * m_diff;
        }
        if (dmax(thr[0], dmax(thr[1], thr[2])) < 0.001) break;                   /* :468 */
    }
    rgb[0] = col[0]; rgb[1] = col[1]; rgb[2] = col[2];
}

/* cuda_tonemap, :74-81 */
static inline double nb_tonemap(double x) { return (x * (2.51 * x + 0.03)) / (x * (2.43 * x + 0.59) + 0.14); }
ORC_API double orc_nb_tonemap(double x) { return nb_tonemap(x); }

/* cuda_path_trace_kernel, :17-58.  Outputs in device row order (row 0 = bottom).
 * sum / sumsq (optional, [H*W*3]) receive the per-pixel sum and sum of squares of the per-sample
 * radiance (pre-tonemap) so callers can form mean and variance.  counters (optional, 4 x u64):
 * closest-hit rays, shadow rays, segments (hits), unshadowed NEE events. */
ORC_API void orc_nb_path_trace(const float *scene, const float *cam, const float *lights, int n_light_floats,
                               const uint8_t *tex, long n_tex_bytes, const int32_t *tex_info, int n_tex_info,
                               int width, int height, int spp, int max_depth, long frame_count,
                               uint8_t *out_u8, double *sum, double *sumsq, uint64_t *counters)
{
    nb_scene s = {scene, cam, lights, n_light_floats, tex, n_tex_bytes, tex_info, n_tex_info};
    uint64_t c0 = 0, c1 = 0, c2 = 0, c3 = 0;
#pragma omp parallel for schedule(dynamic, 2) reduction(+ : c0, c1, c2, c3)
    for (int y = 0; y < height; ++y)
        for (int x = 0; x < width; ++x) {
            path_counters pc = {0, 0, 0, 0};
            int64_t rng = ((int64_t)x + (int64_t)y * width + (int64_t)frame_count * width * height)
                          * 1103515245LL + 12345LL;                              /* :28 */
            double c[3] = {0, 0, 0}, q[3] = {0, 0, 0};
            for (int smp = 0; smp < spp; ++smp) {
                double u = (x + nb_random(rng)) / width;                         /* :35-36: same draw twice */
                double v = (y + nb_random(rng)) / height;
                rng = nb_xorshift(rng);
                double o[3], d[3], rgb[3];
                nb_get_ray(cam, u, v, o, d);
                nb_trace_path(&s, o, d, max_depth, rng, rgb, &pc);
                for (int k = 0; k < 3; ++k) { c[k] += rgb[k]; q[k] += rgb[k] * rgb[k]; }
                rng = nb_xorshift(rng);
            }
            size_t pi = ((size_t)y * width + x) * 3;
            for (int k = 0; k < 3; ++k) {
                if (sum) sum[pi + k] = c[k];
                if (sumsq) sumsq[pi + k] = q[k];
                if (out_u8) out_u8[pi + k] = quant8(nb_tonemap(c[k] / spp));
            }
            c0 += pc.closest_rays; c1 += pc.shadow_rays; c2 += pc.segments; c3 += pc.nee_unshadowed;
        }
    if (counters) { counters[0] = c0; counters[1] = c1; counters[2] = c2; counters[3] = c3; }
}

/* one path with an explicit ray and rng state (unit-level pin against cuda_trace_path) */
ORC_API void orc_nb_trace_path_one(const float *scene, const float *cam, const float *lights, int n_light_floats,
                                   const uint8_t *tex, long n_tex_bytes, const int32_t *tex_info, int n_tex_info,
                                   const double *o, const double *d, int max_depth, int64_t rng, double *rgb)
{
    nb_scene s = {scene, cam, lights, n_light_floats, tex, n_tex_bytes, tex_info, n_tex_info};
    nb_trace_path(&s, o, d, max_depth, rng, rgb, NULL);
}

/* closest hit for explicit rays (unit-level pin against cuda_scene_hit).
 * out per ray: hit flag + packed prim id (ids), t, point(3), normal(3), uv(2), mat(10) = 19 doubles */
ORC_API void orc_nb_scene_hit_rays(const float *scene, int n_rays, const double *o, const double *d,
                                   double t_min, double t_max, int32_t *ids, double *rec)
{
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n_rays; ++i) {
        nb_hit h;
        nb_scene_hit(scene, o + 3 * i, d + 3 * i, t_min, t_max, &h);
        ids[i] = h.hit ? h.prim : -1;
        if (rec) {
            double *r = rec + (size_t)i * 19;
            r[0] = h.t;
            memcpy(r + 1, h.p, 3 * sizeof(double)); memcpy(r + 4, h.n, 3 * sizeof(double));
            memcpy(r + 7, h.uv, 2 * sizeof(double)); memcpy(r + 9, h.mat, 10 * sizeof(double));
        }
    }
}

/* primary hits at a fixed sub-pixel offset (du, dv): u = (x+du)/W, v = (y+dv)/H; row 0 = bottom */
ORC_API void orc_nb_primary_hits(const float *scene, const float *cam, int width, int height,
                                 double du, double dv, int32_t *ids, double *t)
{
#pragma omp parallel for schedule(static)
    for (int y = 0; y < height; ++y)
        for (int x = 0; x < width; ++x) {
            double o[3], d[3];
            nb_get_ray(cam, (x + du) / width, (y + dv) / height, o, d);
            nb_hit h;
            nb_scene_hit(scene, o, d, 0.001, 1000000.0, &h);
            ids[(size_t)y * width + x] = h.hit ? h.prim : -1;
            if (t) t[(size_t)y * width + x] = h.hit ? h.t : -1.0;
        }
}

/* ------------------------------------------------------------------------- */
/* CPU-renderer family: un-rounded float64 objects + the reference's BVH       */
/* ------------------------------------------------------------------------- */
/* Object table (float64, 24 doubles per object), in scene.objects order:
 *   type 0 Plane   : anchor(3) normal(3) u_unit(3) v_unit(3) u_extent v_extent
 *   type 1 Sphere  : center(3) radius
 *   type 2 Triangle: v0(3) v1(3) v2(3) normal(3) uv0(2) uv1(2) uv2(2) has_uv
 * Material table (9 doubles): color(3) diffuse specular reflective refractive ior tex_id(-1 = none)
 * BVH node table: box min(3) max(3) as doubles; children as int32 pairs, >= 0 node index,
 *   < 0 means object ~child.  Node 0 is the root.                              */
typedef struct {
    int n_obj;
    const int32_t *type, *mat_id;
    const double *obj;      /* [n_obj][24] */
    const double *mat;      /* [n_mat][9] */
    int n_node;
    const double *box;      /* [n_node][6] */
    const int32_t *child;   /* [n_node][2] */
    const double *lights; int n_lights;
    v3 light_color, ambient;
    const uint8_t *tex; const int32_t *tex_info; int n_tex;   /* tex_info: offset(bytes), w, h */
} cpu_scene;

typedef struct { double t; v3 p, n; int obj; double u, v; } cpu_rec;

#define OBJ_STRIDE 24
#define MAT_STRIDE 9

/* Plane.hit, core/geometry.py:50-72 */
static int cpu_plane_hit(const double *q, v3 o, v3 d, double t_min, double t_max, cpu_rec *r)
{
    v3 a = V(q[0], q[1], q[2]), n = V(q[3], q[4], q[5]), uu = V(q[6], q[7], q[8]), vv = V(q[9], q[10], q[11]);
    double denom = vdot(n, d);
    if (fabs(denom) < 1e-6) return 0;
    double t = vdot(vsub(a, o), n) / denom;
    if (t < t_min || t > t_max) return 0;
    v3 p = vadd(o, vmul(d, t));
    v3 rel = vsub(p, a);
    double uh = vdot(rel, uu), vh = vdot(rel, vv);
    if (uh < 0 || uh > q[12] || vh < 0 || vh > q[13]) return 0;
    r->t = t; r->p = p; r->n = n; r->u = uh / q[12]; r->v = vh / q[13];
    return 1;
}

/* Sphere.hit, core/geometry.py:85-111 */
static int cpu_sphere_hit(const double *q, v3 o, v3 d, double t_min, double t_max, cpu_rec *r)
{
    v3 c = V(q[0], q[1], q[2]);
    double rad = q[3];
    v3 oc = vsub(o, c);
    double a = vdot(d, d), b = vdot(oc, d), cc = vdot(oc, oc) - rad * rad;
    double disc = b * b - a * cc;
    if (disc > 0) {
        double s = sqrt(disc);
        double cand[2] = {(-b - s) / a, (-b + s) / a};
        for (int k = 0; k < 2; ++k) {
            double t = cand[k];
            if (t_min < t && t < t_max) {
                r->t = t; r->p = vadd(o, vmul(d, t)); r->n = vdiv(vsub(r->p, c), rad); r->u = r->v = 0.0;
                return 1;
            }
        }
    }
    return 0;
}

/* Triangle.hit, core/geometry.py:139-171 */
static int cpu_tri_hit(const double *q, v3 o, v3 d, double t_min, double t_max, cpu_rec *r)
{
    v3 v0 = V(q[0], q[1], q[2]), v1 = V(q[3], q[4], q[5]), v2 = V(q[6], q[7], q[8]), n = V(q[9], q[10], q[11]);
    v3 e1 = vsub(v1, v0), e2 = vsub(v2, v0);
    v3 h = vcross(d, e2);
    double a = vdot(e1, h);
    if (fabs(a) < 1e-6) return 0;
    double f = 1.0 / a;
    v3 s = vsub(o, v0);
    double u = f * vdot(s, h);
    if (u < 0.0 || u > 1.0) return 0;
    v3 qq = vcross(s, e1);
    double v = f * vdot(d, qq);
    if (v < 0.0 || u + v > 1.0) return 0;
    double t = f * vdot(e2, qq);
    if (!(t_min < t && t < t_max)) return 0;
    r->t = t; r->p = vadd(o, vmul(d, t));
    r->n = vdot(n, d) < 0 ? n : vneg(n);
    if (q[18] != 0.0) {
        double w = 1 - u - v;
        r->u = u * q[14] + v * q[16] + w * q[12];
        r->v = u * q[15] + v * q[17] + w * q[13];
    } else { r->u = r->v = 0.0; }
    return 1;
}

static int cpu_obj_hit(const cpu_scene *s, int i, v3 o, v3 d, double t_min, double t_max, cpu_rec *r)
{
    const double *q = s->obj + (size_t)i * OBJ_STRIDE;
    int ok;
    switch (s->type[i]) {
    case 0: ok = cpu_plane_hit(q, o, d, t_min, t_max, r); break;
    case 1: ok = cpu_sphere_hit(q, o, d, t_min, t_max, r); break;
    default: ok = cpu_tri_hit(q, o, d, t_min, t_max, r); break;
    }
    if (ok) r->obj = i;
    return ok;
}

/* AABB.hit, core/math.py:104-117 (inclusive: rejects only when t_max < t_min) */
static int cpu_aabb_hit(const double *b, v3 o, v3 d, double t_min, double t_max)
{
    const double oo[3] = {o.x, o.y, o.z}, dd[3] = {d.x, d.y, d.z};
    for (int a = 0; a < 3; ++a) {
        double inv = 1.0 / dd[a];      /* the reference raises ZeroDivisionError here; C yields +-inf */
        double t0 = (b[a] - oo[a]) * inv, t1 = (b[3 + a] - oo[a]) * inv;
        if (inv < 0.0) { double tmp = t0; t0 = t1; t1 = tmp; }
        t_min = t0 > t_min ? t0 : t_min;
        t_max = t1 < t_max ? t1 : t_max;
        if (t_max < t_min) return 0;
    }
    return 1;
}

/* BVHNode.hit, core/acceleration.py:32-40 */
static int cpu_bvh_hit(const cpu_scene *s, int node, v3 o, v3 d, double t_min, double t_max, cpu_rec *r)
{
    if (node < 0) return cpu_obj_hit(s, ~node, o, d, t_min, t_max, r);
    if (!cpu_aabb_hit(s->box + (size_t)node * 6, o, d, t_min, t_max)) return 0;
    int hl = cpu_bvh_hit(s, s->child[2 * node], o, d, t_min, t_max, r);
    if (hl) t_max = r->t;
    int hr = cpu_bvh_hit(s, s->child[2 * node + 1], o, d, t_min, t_max, r);
    return hl || hr;
}

/* Scene.hit, core/scene.py:45-64 */
static int cpu_scene_hit(const cpu_scene *s, v3 o, v3 d, double t_min, double t_max, cpu_rec *r)
{
    if (s->n_node > 0) return cpu_bvh_hit(s, 0, o, d, t_min, t_max, r);
    int any = 0; double closest = t_max; cpu_rec tmp;
    for (int i = 0; i < s->n_obj; ++i)
        if (cpu_obj_hit(s, i, o, d, t_min, closest, &tmp)) { any = 1; closest = tmp.t; *r = tmp; }
    return any;
}

/* Texture.sample, core/material.py:13-21 */
static v3 cpu_tex_sample(const cpu_scene *s, int tid, double u, double v)
{
    long off = s->tex_info[3 * tid], w = s->tex_info[3 * tid + 1], h = s->tex_info[3 * tid + 2];
    long iu = trunc_l(dmax(0, dmin((double)(w - 1), u * (double)(w - 1))));
    long iv = trunc_l(dmax(0, dmin((double)(h - 1), (1.0 - v) * (double)(h - 1))));
    const uint8_t *px = s->tex + off + (iv * w + iu) * 3;
    return V(px[0] / 255.0, px[1] / 255.0, px[2] / 255.0);
}

/* Vec3.refract, core/math.py:59-67 */
static int cpu_refract(v3 dir, v3 n, double ni_over_nt, v3 *out)
{
    v3 uv = vnorm(dir);
    double dt = vdot(uv, n);
    double disc = 1.0 - ni_over_nt * ni_over_nt * (1 - dt * dt);
    if (disc > 0) {
        *out = vsub(vmul(vsub(uv, vmul(n, dt)), ni_over_nt), vmul(n, sqrt(disc)));
        return 1;
    }
    return 0;
}

/* CPURenderer._trace, renderers/cpu_renderer.py:75-151.  d must be normalised (Ray.__init__). */
static v3 cpu_trace(const cpu_scene *s, v3 o, v3 d, int depth, int max_depth, uint64_t *calls)
{
    cpu_rec rec; rec.t = INFINITY;
    ++*calls;
    if (!cpu_scene_hit(s, o, d, 1e-3, INFINITY, &rec)) return V(0, 0, 0);
    const double *m = s->mat + (size_t)s->mat_id[rec.obj] * MAT_STRIDE;
    double m_diff = m[3], m_spec = m[4], m_refl = m[5], m_refr = m[6], m_ior = m[7];
    int tid = (int)m[8];
    v3 base = tid >= 0 ? cpu_tex_sample(s, tid, rec.u, rec.v) : V(m[0], m[1], m[2]);

    v3 local = vhad(vmul(base, m_diff), s->ambient);                             /* :88 */
    int n = s->n_lights;
    for (int i = 0; i < n; ++i) {                                                /* :92-111 */
        v3 L = V(s->lights[3 * i], s->lights[3 * i + 1], s->lights[3 * i + 2]);
        v3 to_l = vnorm(vsub(L, rec.p));
        v3 so = vadd(rec.p, vmul(rec.n, 1e-3));
        v3 sd = vnorm(to_l);                       /* Ray() normalises again */
        double dist = vlen(vsub(L, rec.p));
        cpu_rec srec; srec.t = INFINITY;
        ++*calls;
        if (!cpu_scene_hit(s, so, sd, 1e-3, dist, &srec)) {
            double diff = dmax(vdot(rec.n, to_l), 0.0);
            local = vadd(local, vdiv(vmul(vhad(vmul(base, m_diff), s->light_color), diff), n));
            v3 view = vnorm(vsub(o, rec.p));
            v3 rdir = vsub(to_l, vmul(rec.n, 2 * vdot(to_l, rec.n)));
            double spec = dmax(vdot(view, rdir), 0.0);
            local = vadd(local, vdiv(vmul(s->light_color, m_spec * pow(spec, 32)), n));
        }
    }

    v3 refl_c = V(0, 0, 0), refr_c = V(0, 0, 0);
    if (m_refl > 0 && depth < max_depth) {                                       /* :114-117 */
        v3 rd = vsub(d, vmul(rec.n, 2 * vdot(d, rec.n)));
        refl_c = cpu_trace(s, vadd(rec.p, vmul(rec.n, 1e-3)), vnorm(rd), depth + 1, max_depth, calls);
    }
    if (m_refr > 0 && depth < max_depth) {                                       /* :121-142 */
        v3 on; double eta;
        if (vdot(d, rec.n) > 0) { on = vneg(rec.n); eta = m_ior; }
        else { on = rec.n; eta = 1.0 / m_ior; }
        v3 rdir;
        if (cpu_refract(d, on, eta, &rdir))
            refr_c = cpu_trace(s, vsub(rec.p, vmul(rec.n, 1e-3)), vnorm(rdir), depth + 1, max_depth, calls);
        else {
            v3 rd = vsub(d, vmul(rec.n, 2 * vdot(d, rec.n)));
            refr_c = cpu_trace(s, vadd(rec.p, vmul(rec.n, 1e-3)), vnorm(rd), depth + 1, max_depth, calls);
        }
    }
    v3 c = V(0, 0, 0);                                                           /* :144-147 */
    c = vadd(c, vmul(local, 1.0 - m_refl - m_refr));
    c = vadd(c, vmul(refl_c, m_refl));
    c = vadd(c, vmul(refr_c, m_refr));
    return c;
}

typedef struct {
    int n_obj; const int32_t *type, *mat_id; const double *obj; const double *mat;
    int n_node; const double *box; const int32_t *child;
    const double *lights; int n_lights; const double *light_color, *ambient;
    const uint8_t *tex; const int32_t *tex_info; int n_tex;
    const double *cam;   /* origin(3) lower_left(3) horizontal(3) vertical(3) */
} orc_cpu_scene_desc;

static cpu_scene cpu_from_desc(const orc_cpu_scene_desc *q)
{
    cpu_scene s;
    s.n_obj = q->n_obj; s.type = q->type; s.mat_id = q->mat_id; s.obj = q->obj; s.mat = q->mat;
    s.n_node = q->n_node; s.box = q->box; s.child = q->child;
    s.lights = q->lights; s.n_lights = q->n_lights;
    s.light_color = V(q->light_color[0], q->light_color[1], q->light_color[2]);
    s.ambient = V(q->ambient[0], q->ambient[1], q->ambient[2]);
    s.tex = q->tex; s.tex_info = q->tex_info; s.n_tex = q->n_tex;
    return s;
}

/* Camera.get_ray + Ray(), core/camera.py:26-31, core/math.py:76-82 */
static void cpu_get_ray(const double *cam, double su, double sv, v3 *o, v3 *d)
{
    v3 org = V(cam[0], cam[1], cam[2]), llc = V(cam[3], cam[4], cam[5]);
    v3 hor = V(cam[6], cam[7], cam[8]), ver = V(cam[9], cam[10], cam[11]);
    v3 dir = vsub(vadd(vadd(llc, vmul(hor, su)), vmul(ver, sv)), org);
    *o = org; *d = vnorm(dir);
}

/* CPURenderer.render's per-sample body (cpu_renderer.py:46-56) with the jitter supplied by the
 * caller: jitter[(y*W+x)*2 + {0,1}] = (du, dv) in [0,1), or NULL for pixel centres (0.5, 0.5).
 * One sample per pixel.  rgb: [H*W*3] float64 pre-quantisation, row 0 = bottom (j = 0).
 * ids (optional): scene.objects index of the primary hit or -1; tt (optional): its t. */
ORC_API void orc_cpu_whitted(const orc_cpu_scene_desc *desc, int width, int height, const double *jitter,
                             int max_depth, double *rgb, int32_t *ids, double *tt, uint64_t *n_hit_calls)
{
    cpu_scene s = cpu_from_desc(desc);
    uint64_t total = 0;
#pragma omp parallel for schedule(dynamic, 2) reduction(+ : total)
    for (int y = 0; y < height; ++y)
        for (int x = 0; x < width; ++x) {
            size_t pix = (size_t)y * width + x;
            double du = jitter ? jitter[pix * 2] : 0.5, dv = jitter ? jitter[pix * 2 + 1] : 0.5;
            v3 o, d;
            cpu_get_ray(desc->cam, (x + du) / width, (y + dv) / height, &o, &d);
            uint64_t calls = 0;
            if (rgb) {
                v3 c = cpu_trace(&s, o, d, 0, max_depth, &calls);
                rgb[pix * 3] = c.x; rgb[pix * 3 + 1] = c.y; rgb[pix * 3 + 2] = c.z;
            }
            if (ids || tt) {
                cpu_rec rec; rec.t = INFINITY;
                int ok = cpu_scene_hit(&s, o, d, 1e-3, INFINITY, &rec);
                if (ids) ids[pix] = ok ? rec.obj : -1;
                if (tt) tt[pix] = ok ? rec.t : -1.0;
            }
            total += calls;
        }
    if (n_hit_calls) *n_hit_calls = total;
}

/* Per-object closest hit WITHOUT the BVH: evaluate every object alone over [1e-3, inf) and take the
 * first arg-min in scene.objects order (the id oracle must not go through AABB.hit). */
ORC_API void orc_cpu_primary_ids_bruteforce(const orc_cpu_scene_desc *desc, int width, int height,
                                            double du, double dv, int32_t *ids, double *tt)
{
    cpu_scene s = cpu_from_desc(desc);
#pragma omp parallel for schedule(static)
    for (int y = 0; y < height; ++y)
        for (int x = 0; x < width; ++x) {
            v3 o, d;
            cpu_get_ray(desc->cam, (x + du) / width, (y + dv) / height, &o, &d);
            int best = -1; double bt = INFINITY;
            for (int i = 0; i < s.n_obj; ++i) {
                cpu_rec r;
                if (cpu_obj_hit(&s, i, o, d, 1e-3, INFINITY, &r) && r.t < bt) { bt = r.t; best = i; }
            }
            ids[(size_t)y * width + x] = best;
            if (tt) tt[(size_t)y * width + x] = best >= 0 ? bt : -1.0;
        }
}

ORC_API int orc_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

ORC_API void orc_set_num_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}
