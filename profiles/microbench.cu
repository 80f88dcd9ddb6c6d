// Peak issue rates the roofline of the compute-bound shapes needs and MEASURED_PEAKS.json lacks:
// FP32 FFMA, FP64 DFMA and MUFU (ex2 / rcp) throughput of one B200, all SMs busy, 8 independent chains per thread.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/microbench profiles/microbench.cu && /tmp/microbench
#include <cstdio>
#include <cuda_runtime.h>

template <int OP>
__global__ void __launch_bounds__(256) kern(float* out, int iters, float seed)
{
    float a[8];
    double d[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { a[k] = seed + k * 0.001f + threadIdx.x * 1e-6f; d[k] = a[k]; }
    const float m = 0.9999f, c = 1e-4f;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (OP == 0) a[k] = fmaf(a[k], m, c);
            else if (OP == 1) d[k] = fma(d[k], (double)m, (double)c);
            else if (OP == 2) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[k]));
            else asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a[k]));
        }
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += a[k] + (float)d[k];
    if (s == 123.456f) out[0] = s;
}

template <int OP>
double run(const char* name, double ops_per_inst)
{
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float* out;
    cudaMalloc(&out, 4);
    const int iters = 20000, blocks = sms * 8, threads = 256;
    kern<OP><<<blocks, threads>>>(out, 100, 0.5f);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    double best = 0;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        kern<OP><<<blocks, threads>>>(out, iters, 0.5f);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        double rate = (double)blocks * threads * iters * 8 / (ms * 1e-3);
        if (rate > best) best = rate;
    }
    printf("{\"op\": \"%s\", \"lane_ops_per_s\": %.4g, \"flops_per_s\": %.4g}\n", name, best, best * ops_per_inst);
    cudaFree(out);
    return best;
}

int main()
{
    run<0>("FFMA (fp32)", 2);
    run<1>("DFMA (fp64)", 2);
    run<2>("MUFU.EX2", 1);
    run<3>("MUFU.RCP", 1);
    return 0;
}
