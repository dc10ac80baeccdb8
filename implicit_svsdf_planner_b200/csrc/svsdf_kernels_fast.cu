// Kernel translation unit, default build: FMA contraction enabled (nvcc default -fmad=true).
#define SVSDF_NS fast
#include "svsdf_kernels.cuh"
