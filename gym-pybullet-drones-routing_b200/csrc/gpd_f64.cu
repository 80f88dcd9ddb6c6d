// FP64 instantiation of the kernels (parity mode).  Compiled with -fmad=false so that every expression keeps
// the reference's (numpy's) rounding: one IEEE operation per source-level operation.
#define GPD_REAL double
#include "gpd_launch.inl"
