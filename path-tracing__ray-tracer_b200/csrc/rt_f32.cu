// rt_f32.cu — float32 (production) instantiation of the templated kernels.
#include "rt_api.cuh"
namespace b2rt { template struct Api<float>; }
