// rt_math.cuh — scalar/vector helpers templated on the arithmetic type (float = production,
// double = parity instantiation compiled with -fmad=false).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b2rt {

template <typename R> struct Real4;
template <> struct Real4<float> {
    using type = float4;
    static __host__ __device__ __forceinline__ float4 make(float a, float b, float c, float d) { return make_float4(a, b, c, d); }
};
template <> struct Real4<double> {
    using type = double4;
    static __host__ __device__ __forceinline__ double4 make(double a, double b, double c, double d) { return make_double4(a, b, c, d); }
};
template <typename R> using real4 = typename Real4<R>::type;

// read-only 4-vector load (LDG.E.128.CONSTANT for float4, two of them for double4)
__device__ __forceinline__ float4 ldg4(const float4 *p) { return __ldg(p); }
__device__ __forceinline__ double4 ldg4(const double4 *p) {
    const double2 *q = reinterpret_cast<const double2 *>(p);
    double2 a = __ldg(q), b = __ldg(q + 1);
    return make_double4(a.x, a.y, b.x, b.y);
}

// L2 residency hints for the hierarchy of large scenes (measurement switch): bit 0 = the node loads
// of the persistent walk kernel, bit 1 = the triangle records, bit 2 = the walk kernel's hit records leave with a
// streaming store.  The loads carry an L2::evict_last policy so that the gigabytes of queue records streaming past do not
// push nodes and triangles out.
#ifndef B2RT_L2_HINT
#define B2RT_L2_HINT 0
#endif
__device__ __forceinline__ unsigned long long l2_keep_policy() {
    unsigned long long pol;
    asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ float4 ldg4_keep(const float4 *p, unsigned long long pol) {
    float4 v;
    asm volatile("ld.global.nc.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ double4 ldg4_keep(const double4 *p, unsigned long long) { return ldg4(p); }

// 256-bit read-only load (sm_100: LDG.E.256): one 32 B sector per lane in ONE L1 data-pipe wavefront where two
// LDG.128 to the same sector cost two.  p must be 32-byte aligned.
__device__ __forceinline__ void ldg8(const float4 *p, float4 &a, float4 &b) {
    asm volatile("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p));
}

// streaming (evict-first) access for queue records that are touched exactly once per kernel, so they do
// not push the small scene tables out of L1/L2
__device__ __forceinline__ float4 ld_stream(const float4 *p) { return __ldcs(p); }
__device__ __forceinline__ double4 ld_stream(const double4 *p) {
    const double2 *q = reinterpret_cast<const double2 *>(p);
    double2 a = __ldcs(q), b = __ldcs(q + 1);
    return make_double4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ void st_stream(float4 *p, float4 v) { __stcs(p, v); }
__device__ __forceinline__ void st_stream(double4 *p, double4 v) {
    double2 *q = reinterpret_cast<double2 *>(p);
    __stcs(q, make_double2(v.x, v.y));
    __stcs(q + 1, make_double2(v.z, v.w));
}

// L[slot] += (x, y, z): every radiance slot belongs to one path and is touched by one thread at a time, so a plain
// load + add + store is race-free.  (Measured: the fire-and-forget vector reduction red.global.add.v4.f32 removes
// the scoreboard wait but runs the whole bounce kernel 2x SLOWER, 43.3 vs 22.4 ms per 128 spp: B2RT_OPT_RED=1.)
#ifndef B2RT_OPT_RED
#define B2RT_OPT_RED 0
#endif
__device__ __forceinline__ void add_stream(float4 *p, float x, float y, float z) {
#if B2RT_OPT_RED
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(x), "f"(y), "f"(z), "f"(0.f) : "memory");
#else
    float4 l = *p;
    *p = make_float4(l.x + x, l.y + y, l.z + z, l.w);
#endif
}
// the escaping path's sky term inside the bounce kernel: B2RT_OPT_RED_SKY=1 uses three scalar fire-and-forget
// reductions (no scoreboard wait on a DRAM-resident line) instead of load + add + store
#ifndef B2RT_OPT_RED_SKY
#define B2RT_OPT_RED_SKY 0
#endif
__device__ __forceinline__ void add_sky(float4 *p, float x, float y, float z) {
#if B2RT_OPT_RED_SKY
    float *q = reinterpret_cast<float *>(p);
    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(q), "f"(x) : "memory");
    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(q + 1), "f"(y) : "memory");
    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(q + 2), "f"(z) : "memory");
#else
    add_stream(p, x, y, z);
#endif
}
__device__ __forceinline__ void add_stream(double4 *p, double x, double y, double z) {
    double4 l = *p;
    *p = make_double4(l.x + x, l.y + y, l.z + z, l.w);
}
// L2 prefetch of a record that a later divergent branch may read-modify-write
#ifndef B2RT_OPT_PREFETCH_L
#define B2RT_OPT_PREFETCH_L 0      // measured: no gain in the bounce kernel (22.3 ms either way), shadow kernel 1.8 -> 2.2 ms
#endif
__device__ __forceinline__ void prefetch_l2(const void *p) {
#if B2RT_OPT_PREFETCH_L
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#endif
}

__device__ __forceinline__ void add_sky(double4 *p, double x, double y, double z) { add_stream(p, x, y, z); }

template <typename R> struct V3 {
    R x, y, z;
};
template <typename R> __device__ __forceinline__ V3<R> mk3(R x, R y, R z) { return V3<R>{x, y, z}; }
template <typename R> __device__ __forceinline__ V3<R> xyz(const real4<R> &v) { return V3<R>{v.x, v.y, v.z}; }
template <typename R> __device__ __forceinline__ V3<R> operator+(V3<R> a, V3<R> b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
template <typename R> __device__ __forceinline__ V3<R> operator-(V3<R> a, V3<R> b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
template <typename R> __device__ __forceinline__ V3<R> operator-(V3<R> a) { return {-a.x, -a.y, -a.z}; }
template <typename R> __device__ __forceinline__ V3<R> operator*(V3<R> a, R k) { return {a.x * k, a.y * k, a.z * k}; }
template <typename R> __device__ __forceinline__ V3<R> operator*(R k, V3<R> a) { return {a.x * k, a.y * k, a.z * k}; }
template <typename R> __device__ __forceinline__ V3<R> operator/(V3<R> a, R k) { return {a.x / k, a.y / k, a.z / k}; }
template <typename R> __device__ __forceinline__ V3<R> had(V3<R> a, V3<R> b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
// left-to-right dot, the association every reference expression uses (x*x' + y*y' + z*z')
template <typename R> __device__ __forceinline__ R dot(V3<R> a, V3<R> b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
template <typename R> __device__ __forceinline__ V3<R> cross(V3<R> a, V3<R> b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
// float32 production kernels use the single-instruction MUFU approximations (<= 1 ulp, no slow-path call:
// the IEEE sequences cost ~10 instructions and a CALL each and pushed the fused kernel past the 32 KB
// instruction cache); the float64 parity kernels keep IEEE sqrt and division.
__device__ __forceinline__ float sqrt_(float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ double sqrt_(double x) { return sqrt(x); }
__device__ __forceinline__ float abs_(float x) { return fabsf(x); }
__device__ __forceinline__ double abs_(double x) { return fabs(x); }
__device__ __forceinline__ float max_(float a, float b) { return fmaxf(a, b); }
__device__ __forceinline__ double max_(double a, double b) { return fmax(a, b); }
__device__ __forceinline__ float fma_(float a, float b, float c) { return fmaf(a, b, c); }
__device__ __forceinline__ double fma_(double a, double b, double c) { return fma(a, b, c); }
__device__ __forceinline__ float min_(float a, float b) { return fminf(a, b); }
__device__ __forceinline__ double min_(double a, double b) { return fmin(a, b); }
__device__ __forceinline__ float pow_(float a, float b) { return powf(a, b); }
__device__ __forceinline__ double pow_(double a, double b) { return pow(a, b); }
// Division policy: the float64 parity instantiation divides exactly where the reference divides;
// the float32 production instantiation multiplies by one MUFU.RCP reciprocal.
template <typename R> __device__ __forceinline__ R rcp_(R x) {
    if constexpr (sizeof(R) == 4) return rcp_approx(x);
    else return R(1) / x;
}
template <typename R> __device__ __forceinline__ V3<R> div3(V3<R> a, R k) {
    if constexpr (sizeof(R) == 4) { R r = rcp_approx(k); return {a.x * r, a.y * r, a.z * r}; }
    else return {a.x / k, a.y / k, a.z / k};
}
template <typename R> __device__ __forceinline__ R div_(R a, R b) {
    if constexpr (sizeof(R) == 4) return a * rcp_approx(b);
    else return a / b;
}
template <typename R> __device__ __forceinline__ R length(V3<R> a) { return sqrt_(a.x * a.x + a.y * a.y + a.z * a.z); }
// Vec3.normalize (core/math.py:49-53): divide by the length; the zero vector stays zero
template <typename R> __device__ __forceinline__ V3<R> normalize(V3<R> a) {
    R l = length(a);
    return l == R(0) ? V3<R>{R(0), R(0), R(0)} : div3(a, l);
}
template <typename R> __device__ __forceinline__ V3<R> cvt3(V3<double> a) { return {R(a.x), R(a.y), R(a.z)}; }

template <typename R> struct Cam {
    V3<R> origin, llc, hor, ver;
};
template <typename R> __host__ inline Cam<R> make_cam(const double *c) {
    Cam<R> k;
    k.origin = {R(c[0]), R(c[1]), R(c[2])};
    k.llc = {R(c[3]), R(c[4]), R(c[5])};
    k.hor = {R(c[6]), R(c[7]), R(c[8])};
    k.ver = {R(c[9]), R(c[10]), R(c[11])};
    return k;
}

// bit casts between the real type and a same-width integer (queue records carry ints in .w lanes)
__device__ __forceinline__ float int_as_real(int32_t v, float) { return __int_as_float(v); }
__device__ __forceinline__ double int_as_real(int64_t v, double) { return __longlong_as_double(v); }
__device__ __forceinline__ int64_t real_as_int(float v) { return (int64_t)__float_as_int(v); }
__device__ __forceinline__ int64_t real_as_int(double v) { return __double_as_longlong(v); }
template <typename R> __device__ __forceinline__ R pack_int(int64_t v) {
    if constexpr (sizeof(R) == 4) return __int_as_float((int32_t)v);
    else return __longlong_as_double(v);
}
template <typename R> __device__ __forceinline__ uint64_t unpack_u(R v) {
    if constexpr (sizeof(R) == 4) return (uint64_t)(uint32_t)__float_as_int(v);
    else return (uint64_t)__double_as_longlong(v);
}

}  // namespace b2rt
