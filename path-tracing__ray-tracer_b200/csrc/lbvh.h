// lbvh.h — host entry points of the LBVH builder (lbvh.cu).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace b2rt {
size_t lbvh_temp_bytes(int n_prims);
// h_meta: [0] n_top, [1] root reference, [2] internal node count.  Synchronises `stream`.
cudaError_t lbvh_build(int n_rect, int n_sphere, int n_tri, const float4 *rect, const float4 *sphere, const float4 *tri,
                       float pad, float4 *nodes, float4 *top, int top_capacity, int *h_meta, void *temp,
                       size_t temp_bytes, cudaStream_t stream, int build_flags = 0);
// 4-wide nodes (128 B each, indexed by the child reference) derived from a finished hierarchy; see lbvh.cu:widen_kernel
size_t lbvh_wide_bytes(int n_top, int n_internal);
cudaError_t lbvh_widen(const float4 *nodes, const float4 *top, int n_top, int n_internal, float4 *wide, size_t wide_bytes,
                       cudaStream_t stream);
// quantised 32 B binary nodes on a 16-bit grid over [lo, hi] (host float[3] each); see lbvh.cu:quantize_kernel
size_t lbvh_quant_bytes(int n_top, int n_internal);
cudaError_t lbvh_quantize(const float4 *nodes, const float4 *top, int n_top, int n_internal, const float *lo, const float *hi,
                          void *quant, size_t quant_bytes, cudaStream_t stream);
}  // namespace b2rt
