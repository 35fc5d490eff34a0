// c_api.cu — the extern "C" surface declared in include/b200rt.h.
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <mutex>
#include <vector>

#include "../../include/b200rt.h"
#include "lbvh.h"
#include "rt_api.h"

namespace {
thread_local char g_err[512] = "";

int fail(const char *where, cudaError_t e) {
    snprintf(g_err, sizeof g_err, "%s: %s", where, cudaGetErrorString(e));
    return 1;
}
int fail_msg(const char *msg) {
    snprintf(g_err, sizeof g_err, "%s", msg);
    return 2;
}
inline cudaStream_t S(void *s) { return reinterpret_cast<cudaStream_t>(s); }
}  // namespace
namespace b2rt {
int set_error(int code, const char *fmt, ...) {          // used by the other translation units (scene_prepare.cu)
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}
}  // namespace b2rt
namespace {

#define DISPATCH(scene_precision, call_f32, call_f64)                                   \
    ((scene_precision) == B2RT_PRECISION_F64 ? (call_f64) : (call_f32))
}  // namespace

using b2rt::Api;

// ---- per-kernel-class event timing ----------------------------------------------------------------
// State is per DEVICE (events belong to the device they were created on) and guarded by a mutex: host threads that
// drive different GPUs, or the same GPU, never share a vector unsynchronised.  Records are bounded.
namespace b2rt {
namespace {
struct ProfRec { int cls; cudaEvent_t a, b; };
struct ProfState {
    std::mutex mu;
    bool on = false, open = false;          // open: the last record still waits for its end event
    std::vector<ProfRec> recs;
    std::vector<cudaEvent_t> pool;
};
constexpr int kMaxDevices = 64;
constexpr size_t kMaxProfRecs = size_t(1) << 20;      // ~90 frames of the headline bench; further launches are not recorded
ProfState g_prof[kMaxDevices];
ProfState *prof_state() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return nullptr;
    return &g_prof[dev];
}
cudaEvent_t prof_event(ProfState &p) {
    cudaEvent_t e;
    if (!p.pool.empty()) { e = p.pool.back(); p.pool.pop_back(); }
    else cudaEventCreate(&e);
    return e;
}
}  // namespace
void prof_begin(int cls, cudaStream_t st) {
    ProfState *p = prof_state();
    if (!p || !p->on) return;
    std::lock_guard<std::mutex> lock(p->mu);
    p->open = false;
    if (p->recs.size() >= kMaxProfRecs) return;
    ProfRec r{cls, prof_event(*p), prof_event(*p)};
    cudaEventRecord(r.a, st);
    cudaEventRecord(r.b, st);          // re-recorded by prof_end; never left unrecorded
    p->recs.push_back(r);
    p->open = true;
}
void prof_end(cudaStream_t st) {
    ProfState *p = prof_state();
    if (!p || !p->on) return;
    std::lock_guard<std::mutex> lock(p->mu);
    if (!p->open || p->recs.empty()) return;
    cudaEventRecord(p->recs.back().b, st);
    p->open = false;
}
}  // namespace b2rt

// ---- B2RT_CHECK: in-kernel bounds checks counted into a library-owned device word (one per device) -------------------
#ifndef B2RT_CHECK
#define B2RT_CHECK 0
#endif
namespace b2rt {
unsigned long long *check_counter() {
#if B2RT_CHECK
    static unsigned long long *ptr[64] = {nullptr};
    static std::mutex mu;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    std::lock_guard<std::mutex> lock(mu);
    if (!ptr[dev] && cudaMalloc(&ptr[dev], sizeof(unsigned long long)) == cudaSuccess) cudaMemset(ptr[dev], 0, sizeof(unsigned long long));
    return ptr[dev];
#else
    return nullptr;
#endif
}
}  // namespace b2rt

static __global__ void fp32_peak_kernel(float *out, int iters, float a, float b) {
    float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
        x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
        x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
    }
    float s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
    if (s == 12345.678f) out[0] = s;
}

extern "C" {

const char *b2rt_last_error(void) { return g_err; }
int b2rt_version(void) { return 200 + B2RT_ABI_VERSION; }

int b2rt_device_info(int device, int64_t *h_out) {
    cudaDeviceProp p;
    cudaError_t e = cudaGetDeviceProperties(&p, device);
    if (e) return fail("b2rt_device_info", e);
    int clock_khz = 0;
    cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, device);
    h_out[0] = p.multiProcessorCount;
    h_out[1] = (int64_t)p.sharedMemPerBlockOptin;
    h_out[2] = p.l2CacheSize;
    h_out[3] = clock_khz;
    h_out[4] = p.major;
    h_out[5] = p.minor;
    return 0;
}

int b2rt_lbvh_temp_bytes(int32_t n_prims, size_t *h_bytes) {
    *h_bytes = b2rt::lbvh_temp_bytes(n_prims);
    return 0;
}

int b2rt_lbvh_build(int32_t n_rect, int32_t n_sphere, int32_t n_tri, const void *d_rect, const void *d_sphere,
                    const void *d_tri, float box_pad, void *d_nodes_out, void *d_top_out, int32_t top_capacity,
                    int32_t *h_meta_out, void *d_temp, size_t temp_bytes, void *stream, int32_t flags) {
    if (top_capacity > b2rt::kTopMax) top_capacity = b2rt::kTopMax;
    if (top_capacity < 0) top_capacity = 0;
    int meta[3];
    cudaError_t e = b2rt::lbvh_build(n_rect, n_sphere, n_tri, (const float4 *)d_rect, (const float4 *)d_sphere,
                                     (const float4 *)d_tri, box_pad, (float4 *)d_nodes_out, (float4 *)d_top_out,
                                     top_capacity, meta, d_temp, temp_bytes, S(stream), flags);
    if (e) return fail("b2rt_lbvh_build", e);
    h_meta_out[0] = meta[0]; h_meta_out[1] = meta[1]; h_meta_out[2] = meta[2];
    return 0;
}

int b2rt_lbvh_wide_bytes(int32_t n_top, int32_t n_internal, size_t *h_bytes) {
    *h_bytes = b2rt::lbvh_wide_bytes(n_top, n_internal);
    return 0;
}

int b2rt_lbvh_widen(const void *d_nodes, const void *d_top, int32_t n_top, int32_t n_internal, void *d_wide_out,
                    size_t wide_bytes, void *stream) {
    cudaError_t e = b2rt::lbvh_widen((const float4 *)d_nodes, (const float4 *)d_top, n_top, n_internal, (float4 *)d_wide_out,
                                     wide_bytes, S(stream));
    if (e) return fail("b2rt_lbvh_widen", e);
    return 0;
}

int b2rt_lbvh_quant_bytes(int32_t n_top, int32_t n_internal, size_t *h_bytes) {
    *h_bytes = b2rt::lbvh_quant_bytes(n_top, n_internal);
    return 0;
}

int b2rt_lbvh_quantize(const void *d_nodes, const void *d_top, int32_t n_top, int32_t n_internal, const float *h_lo,
                       const float *h_hi, void *d_quant_out, size_t quant_bytes, void *stream) {
    cudaError_t e = b2rt::lbvh_quantize((const float4 *)d_nodes, (const float4 *)d_top, n_top, n_internal, h_lo, h_hi,
                                        d_quant_out, quant_bytes, S(stream));
    if (e) return fail("b2rt_lbvh_quantize", e);
    return 0;
}

static int check_scene(const b2rt_scene *s) {
    if (!s) return fail_msg("scene is NULL");
    if (s->struct_size != sizeof(b2rt_scene) || s->abi_version != B2RT_ABI_VERSION) {
        snprintf(g_err, sizeof g_err, "b2rt_scene ABI mismatch: caller has struct_size %u / abi_version %u, library has %zu / %d "
                                      "(set both from the header the caller was built against)",
                 s->struct_size, s->abi_version, sizeof(b2rt_scene), B2RT_ABI_VERSION);
        return 3;
    }
    if (s->precision != B2RT_PRECISION_F32 && s->precision != B2RT_PRECISION_F64) return fail_msg("bad precision");
    if (s->n_bvh_top < 0 || s->n_bvh_top > b2rt::kTopMax) return fail_msg("n_bvh_top out of range");
    return 0;
}

int b2rt_primary_hits(const b2rt_scene *scene, const double *h_cam, int32_t width, int32_t height, double du,
                      double dv, double t_min, double t_max, int32_t use_bvh, int32_t *d_ids, double *d_t,
                      void *stream) {
    if (int r = check_scene(scene)) return r;
    cudaError_t e = DISPATCH(scene->precision,
        Api<float>::primary_hits(scene, h_cam, width, height, du, dv, t_min, t_max, use_bvh, d_ids, d_t, S(stream)),
        Api<double>::primary_hits(scene, h_cam, width, height, du, dv, t_min, t_max, use_bvh, d_ids, d_t, S(stream)));
    return e ? fail("b2rt_primary_hits", e) : 0;
}

int b2rt_trace_rays(const b2rt_scene *scene, int32_t n, const double *d_o, const double *d_d, double t_min,
                    double t_max, int32_t any_hit, int32_t use_bvh, int32_t *d_ids, double *d_rec, void *stream) {
    if (int r = check_scene(scene)) return r;
    cudaError_t e = DISPATCH(scene->precision,
        Api<float>::trace_rays(scene, n, d_o, d_d, t_min, t_max, any_hit, use_bvh, d_ids, d_rec, S(stream)),
        Api<double>::trace_rays(scene, n, d_o, d_d, t_min, t_max, any_hit, use_bvh, d_ids, d_rec, S(stream)));
    return e ? fail("b2rt_trace_rays", e) : 0;
}

int b2rt_render_whitted_cpu(const b2rt_scene *scene, const double *h_cam, int32_t width, int32_t height,
                            const double *d_jitter, int32_t max_depth, const double *h_ambient,
                            const double *h_light_color, double *d_rgb, void *stream) {
    if (int r = check_scene(scene)) return r;
    if (max_depth < 0 || max_depth > 20) return fail_msg("whitted_cpu: max_depth must be in [0, 20]");
    cudaError_t e = DISPATCH(scene->precision,
        Api<float>::whitted_cpu(scene, h_cam, width, height, d_jitter, max_depth, h_ambient, h_light_color, d_rgb, S(stream)),
        Api<double>::whitted_cpu(scene, h_cam, width, height, d_jitter, max_depth, h_ambient, h_light_color, d_rgb, S(stream)));
    return e ? fail("b2rt_render_whitted_cpu", e) : 0;
}

int b2rt_render_whitted_texture(const b2rt_scene *scene, const double *h_cam, int32_t width, int32_t height,
                                int32_t spp, int32_t max_depth, double *d_rgb, uint8_t *d_u8, void *stream) {
    if (int r = check_scene(scene)) return r;
    if (spp < 1) return fail_msg("whitted_texture: spp must be >= 1");
    cudaError_t e = DISPATCH(scene->precision,
        Api<float>::whitted_texture(scene, h_cam, width, height, spp, max_depth, d_rgb, d_u8, S(stream)),
        Api<double>::whitted_texture(scene, h_cam, width, height, spp, max_depth, d_rgb, d_u8, S(stream)));
    return e ? fail("b2rt_render_whitted_texture", e) : 0;
}

int b2rt_path_workspace_bytes(int32_t precision, int32_t width, int32_t height, int32_t spp_per_wave,
                              int32_t max_depth, size_t *h_bytes) {
    if (width < 1 || height < 1 || spp_per_wave < 1 || max_depth < 1) return fail_msg("path_workspace_bytes: bad size");
    *h_bytes = precision == B2RT_PRECISION_F64 ? Api<double>::path_workspace_bytes(width, height, spp_per_wave, max_depth)
                                               : Api<float>::path_workspace_bytes(width, height, spp_per_wave, max_depth);
    return 0;
}

int b2rt_render_path(const b2rt_scene *scene, const double *h_cam, int32_t width, int32_t height, int32_t spp_local,
                     int64_t sample_offset, int32_t spp_per_wave, int32_t max_depth, int32_t rng_mode, uint64_t seed,
                     int32_t flags,
                     void *d_accum, void *d_accum_sq, int64_t *d_pixel_rng, void *d_workspace, size_t workspace_bytes,
                     uint64_t *d_counters, void *stream) {
    if (int r = check_scene(scene)) return r;
    if (max_depth < 1) return fail_msg("render_path: max_depth must be >= 1");
    if ((long long)width * height * (long long)(spp_per_wave < 1 ? 1 : spp_per_wave) > 0x7fffffffLL - (1LL << 22))
        return fail_msg("render_path: wave larger than 2^31 paths");
    b2rt::PathArgs a;
    a.width = width; a.height = height; a.spp_local = spp_local; a.spp_per_wave = spp_per_wave;
    a.max_depth = max_depth; a.rng_mode = rng_mode; a.flags = flags; a.sample_offset = sample_offset; a.seed = seed;
    a.accum = d_accum; a.accum_sq = d_accum_sq; a.pixel_rng = (long long *)d_pixel_rng; a.workspace = d_workspace;
    a.workspace_bytes = workspace_bytes; a.counters = (unsigned long long *)d_counters;
    cudaError_t e = DISPATCH(scene->precision, Api<float>::render_path(scene, h_cam, a, S(stream)),
                             Api<double>::render_path(scene, h_cam, a, S(stream)));
    return e ? fail("b2rt_render_path", e) : 0;
}

int b2rt_resolve(int32_t precision, const void *d_accum, int32_t width, int32_t height, double spp_total,
                 int32_t tonemap, uint8_t *d_u8, void *stream) {
    cudaError_t e = DISPATCH(precision, Api<float>::resolve(d_accum, width, height, spp_total, tonemap, d_u8, S(stream)),
                             Api<double>::resolve(d_accum, width, height, spp_total, tonemap, d_u8, S(stream)));
    return e ? fail("b2rt_resolve", e) : 0;
}

int b2rt_reduce_resolve(const void *const *h_peer_accum, int32_t n_peers, int32_t width, int32_t height, int32_t row0,
                        int32_t row1, double spp_total, int32_t tonemap, uint8_t *d_u8_root, void *d_sum_root, void *stream) {
    if (!h_peer_accum || !d_u8_root) return fail_msg("reduce_resolve: NULL buffer");
    if (row0 < 0 || row1 > height || width < 1) return fail_msg("reduce_resolve: bad row range");
    cudaError_t e = b2rt::reduce_resolve_f32(h_peer_accum, n_peers, width, height, row0, row1, spp_total, tonemap,
                                             d_u8_root, d_sum_root, S(stream));
    return e ? fail("b2rt_reduce_resolve", e) : 0;
}

int b2rt_expand_rgb8(const uint8_t *d_rgb, int64_t n_texels, uint32_t *d_rgbx, void *stream) {
    if (((size_t)d_rgb & 3) || ((size_t)d_rgbx & 15)) return fail_msg("expand_rgb8: d_rgb must be 4-byte, d_rgbx 16-byte aligned");
    cudaError_t e = b2rt::expand_rgb8(d_rgb, n_texels, d_rgbx, S(stream));
    return e ? fail("b2rt_expand_rgb8", e) : 0;
}

int b2rt_check_enabled(void) { return B2RT_CHECK ? 1 : 0; }

int b2rt_check_read(uint64_t *h_stack_overflows, uint64_t *h_queue_overruns) {
    *h_stack_overflows = *h_queue_overruns = 0;
    unsigned long long *p = b2rt::check_counter();
    if (!p) return 0;
    unsigned long long v = 0;
    cudaError_t e = cudaMemcpy(&v, p, sizeof v, cudaMemcpyDeviceToHost);        // synchronises with the kernels before it
    if (!e) e = cudaMemset(p, 0, sizeof v);
    if (e) return fail("b2rt_check_read", e);
    *h_stack_overflows = v & 0xffffffffULL;
    *h_queue_overruns = v >> 32;
    return 0;
}

int b2rt_profile_enable(int32_t on) {
    b2rt::ProfState *p = b2rt::prof_state();
    if (!p) return fail_msg("b2rt_profile_enable: no current device");
    std::lock_guard<std::mutex> lock(p->mu);
    p->on = on != 0;
    return 0;
}

int b2rt_profile_read(double *h_ms, int64_t *h_launches) {
    for (int k = 0; k < b2rt::kNumClasses; ++k) { h_ms[k] = 0.0; h_launches[k] = 0; }
    b2rt::ProfState *p = b2rt::prof_state();
    if (!p) return fail_msg("b2rt_profile_read: no current device");
    std::lock_guard<std::mutex> lock(p->mu);
    for (auto &r : p->recs) {
        cudaError_t e = cudaEventSynchronize(r.b);
        if (e) return fail("b2rt_profile_read", e);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, r.a, r.b);
        h_ms[r.cls] += ms;
        h_launches[r.cls] += 1;
        p->pool.push_back(r.a);
        p->pool.push_back(r.b);
    }
    p->recs.clear();
    return 0;
}

int b2rt_fp32_peak(int32_t iters, double *h_tflops, void *stream) {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    float *out = nullptr;
    cudaError_t e = cudaMalloc(&out, 16);
    if (e) return fail("b2rt_fp32_peak", e);
    const int T = 256, G = sms * 8;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    fp32_peak_kernel<<<G, T, 0, S(stream)>>>(out, 1000, 1.0000001f, 1e-7f);       // warm-up
    cudaEventRecord(a, S(stream));
    fp32_peak_kernel<<<G, T, 0, S(stream)>>>(out, iters, 1.0000001f, 1e-7f);
    cudaEventRecord(b, S(stream));
    e = cudaEventSynchronize(b);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, a, b);
    cudaEventDestroy(a); cudaEventDestroy(b); cudaFree(out);
    if (e) return fail("b2rt_fp32_peak", e);
    *h_tflops = (double)G * T * 8.0 * 2.0 * iters / (ms * 1e-3) / 1e12;
    return 0;
}

}  // extern "C"
