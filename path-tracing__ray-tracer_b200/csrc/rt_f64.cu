// rt_f64.cu — float64 parity instantiation of the SAME templated kernels; built with -fmad=false so
// that a*b+c is never contracted (the reference's float64 arithmetic is unfused).
#include "rt_api.cuh"
namespace b2rt { template struct Api<double>; }
