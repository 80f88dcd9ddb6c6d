// FP32 instantiation of the kernels (throughput mode; FMA contraction on).
#define GPD_REAL float
#include "gpd_launch.inl"

namespace gpd {
// precision-independent helpers live in this translation unit
// ---- episode statistics: reduce the per-block slots ----
__device__ __forceinline__ float ordered_to_float_h(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

__global__ void stats_reduce_kernel(const StatSlot* __restrict__ slots, int64_t nslots, double* __restrict__ out8)
{
    __shared__ double sh[8][32];
    double acc[8] = { 0, 0, 0, 0, 1e300, -1e300, 0, 0 };   // episodes, ret, len, ret^2, min, max, env_steps, terminated
    for (int64_t k = threadIdx.x; k < nslots; k += blockDim.x) {
        const StatSlot& s = slots[k];
        acc[0] += s.s[0]; acc[1] += s.s[1]; acc[2] += s.s[2]; acc[3] += s.s[3];
        if (s.s[0] > 0) { acc[4] = fmin(acc[4], (double)ordered_to_float_h(s.mn)); acc[5] = fmax(acc[5], (double)ordered_to_float_h(s.mx)); }
        acc[6] += s.s[4]; acc[7] += s.s[5];
    }
    for (int off = 16; off > 0; off >>= 1)
        for (int j = 0; j < 8; ++j) {
            double o = __shfl_down_sync(0xffffffffu, acc[j], off);
            if (j == 4) acc[j] = fmin(acc[j], o); else if (j == 5) acc[j] = fmax(acc[j], o); else acc[j] += o;
        }
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) for (int j = 0; j < 8; ++j) sh[j][w] = acc[j];
    __syncthreads();
    if (threadIdx.x == 0) {
        int nw = blockDim.x >> 5;
        for (int j = 0; j < 8; ++j) {
            double v = sh[j][0];
            for (int k = 1; k < nw; ++k) { if (j == 4) v = fmin(v, sh[j][k]); else if (j == 5) v = fmax(v, sh[j][k]); else v += sh[j][k]; }
            out8[j] = v;
        }
    }
}

__global__ void stats_clear_kernel(StatSlot* __restrict__ slots, int64_t nslots)
{
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nslots) return;
    StatSlot& s = slots[k];
    for (int j = 0; j < 6; ++j) s.s[j] = 0.0;
    s.mn = 0x7fffffff; s.mx = (int32_t)0x80000000;
}

cudaError_t launch_stats(const StatSlot* slots, int64_t nslots, double* out8, int clear, StatSlot* slots_mut, cudaStream_t st)
{
    stats_reduce_kernel<<<1, 256, 0, st>>>(slots, nslots, out8);
    if (clear) stats_clear_kernel<<<(unsigned)((nslots + 127) / 128), 128, 0, st>>>(slots_mut, nslots);
    return cudaGetLastError();
}
}  // namespace gpd
