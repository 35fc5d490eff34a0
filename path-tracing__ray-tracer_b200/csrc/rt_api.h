// rt_api.h — host-side launcher declarations, templated on the arithmetic type.  rt_f32.cu
// instantiates Api<float> (default nvcc flags), rt_f64.cu instantiates Api<double> with -fmad=false.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "../../include/b200rt.h"

namespace b2rt {

constexpr int kTopMax = 1024;         // upper bound on smem-staged BVH nodes (64 B each)

// optional per-kernel-class CUDA-event timing (b2rt_profile_enable / b2rt_profile_read, c_api.cu)
enum KernelClass { kRaygen = 0, kExtend = 1, kShade = 2, kShadow = 3, kAccumulate = 4, kMisc = 5, kNumClasses = 8 };
void prof_begin(int cls, cudaStream_t st);
void prof_end(cudaStream_t st);

struct PathArgs {
    int width, height, spp_local, spp_per_wave, max_depth, rng_mode, flags;
    long long sample_offset;
    unsigned long long seed;
    void *accum, *accum_sq;
    long long *pixel_rng;
    void *workspace;
    size_t workspace_bytes;
    unsigned long long *counters;
};

// B2RT_CHECK builds: the library-owned violation counter of the current device (nullptr in normal builds)
unsigned long long *check_counter();

// float32-only helpers (rt_f32.cu)
cudaError_t reduce_resolve_f32(const void *const *peer_accum, int n_peers, int W, int H, int row0, int row1, double spp,
                               int tonemap, uint8_t *root_u8, void *root_sum, cudaStream_t st);
cudaError_t expand_rgb8(const uint8_t *rgb, long long n_texels, uint32_t *rgbx, cudaStream_t st);

template <typename R> struct Api {
    static cudaError_t primary_hits(const b2rt_scene *s, const double *cam, int W, int H, double du, double dv,
                                    double t_min, double t_max, int use_bvh, int *ids, double *tt, cudaStream_t st);
    static cudaError_t trace_rays(const b2rt_scene *s, int n, const double *o, const double *d, double t_min,
                                  double t_max, int any_hit, int use_bvh, int *ids, double *rec, cudaStream_t st);
    static cudaError_t whitted_cpu(const b2rt_scene *s, const double *cam, int W, int H, const double *jitter,
                                   int max_depth, const double *ambient, const double *light_color, double *rgb,
                                   cudaStream_t st);
    static cudaError_t whitted_texture(const b2rt_scene *s, const double *cam, int W, int H, int spp, int max_depth,
                                       double *rgb, uint8_t *u8, cudaStream_t st);
    static size_t path_workspace_bytes(int W, int H, int spp_per_wave, int max_depth);
    static cudaError_t render_path(const b2rt_scene *s, const double *cam, const PathArgs &a, cudaStream_t st);
    static cudaError_t resolve(const void *accum, int W, int H, double spp, int tonemap, uint8_t *u8, cudaStream_t st);
};

}  // namespace b2rt
