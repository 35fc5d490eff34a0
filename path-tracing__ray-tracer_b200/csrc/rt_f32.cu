// rt_f32.cu — float32 (production) instantiation of the templated kernels, plus the float32-only launchers.
#include "rt_api.cuh"
namespace b2rt {
template struct Api<float>;

cudaError_t reduce_resolve_f32(const void *const *peer_accum, int n_peers, int W, int H, int row0, int row1, double spp,
                               int tonemap, uint8_t *root_u8, void *root_sum, cudaStream_t st) {
    if (n_peers < 1 || n_peers > kMaxPeers) return cudaErrorInvalidValue;
    if (row1 <= row0) return cudaSuccess;
    PeerAccum pa;
    for (int p = 0; p < kMaxPeers; ++p) pa.p[p] = reinterpret_cast<const float4 *>(peer_accum[p < n_peers ? p : 0]);
    const int n = (row1 - row0) * ((W + 3) / 4);
    int grid = 0;
    if (cudaError_t e = persistent_grid((const void *)reduce_resolve_kernel, 256, 0, &grid)) return e;
    if (grid > (n + 255) / 256) grid = (n + 255) / 256;
    reduce_resolve_kernel<<<grid, 256, 0, st>>>(pa, n_peers, W, H, row0, row1, (float)spp, tonemap, root_u8, (float4 *)root_sum);
    return cudaGetLastError();
}

cudaError_t expand_rgb8(const uint8_t *rgb, long long n_texels, uint32_t *rgbx, cudaStream_t st) {
    if (n_texels <= 0) return cudaSuccess;
    int grid = 0;
    if (cudaError_t e = persistent_grid((const void *)expand_rgb8_kernel, 256, 0, &grid)) return e;
    expand_rgb8_kernel<<<grid, 256, 0, st>>>(rgb, n_texels, rgbx);
    return cudaGetLastError();
}
}  // namespace b2rt
