// Kernel translation unit, parity-debug build: compiled with -fmad=false so every a*b+c rounds twice like the
// CPU oracle / the reference's x86-64 build (SURVEY.md §7 "Hard parts").
#define SVSDF_NS strict
#include "svsdf_kernels.cuh"
