// stream_microbench.cu — what does a PLAIN streaming kernel reach with the step kernel's launch structure?
// 8 rotating buffer sets (> L2), one launch = 65,536 "envs" x (300 B read + 354 B written) by 1024 CTAs, launches back to back
// in a CUDA graph, with and without programmatic dependent launch.  No dependencies, no compute: the ceiling of the
// memory system for this traffic volume and grid shape (profiles/README.md, round 2).
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o stream_microbench stream_microbench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

template <int PDL>
__global__ void __launch_bounds__(256) stream_kernel(const float4* __restrict__ src, float4* __restrict__ dst, int rd4, int wr4)
{
    if (PDL) asm volatile("griddepcontrol.launch_dependents;");
    const float4* s = src + (size_t)blockIdx.x * rd4;
    float4* d = dst + (size_t)blockIdx.x * wr4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = threadIdx.x; i < rd4; i += blockDim.x * 4) {
        float4 v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) { int k = i + j * blockDim.x; v[j] = k < rd4 ? __ldg(s + k) : make_float4(0, 0, 0, 0); }
#pragma unroll
        for (int j = 0; j < 4; ++j) { acc.x += v[j].x; acc.y += v[j].y; acc.z += v[j].z; acc.w += v[j].w; }
    }
    for (int i = threadIdx.x; i < wr4; i += blockDim.x) d[i] = acc;
}

// The step kernel's ACCESS PATTERN without its dependencies and arithmetic: per 64-env tile the same 14 address streams
// (state SoA 16+16+16+4 B, counter, episode return, action, 224 of every 288-byte history row; written back: the same state,
// a 288-byte observation row, reward, two flags), physics threads + one copy warp.  Tells a pattern limit from a latency limit.
struct Pat {
    float4 *sP, *sQ, *sV; float *sW; int* cnt; float* ep; const float4* act; const float4* prev; float4* out; float* rew;
    unsigned char *te, *tr;
};
template <int PDL>
__global__ void __launch_bounds__(96) pattern_kernel(Pat p, int E)
{
    if (PDL) asm volatile("griddepcontrol.launch_dependents;");
    const int t = threadIdx.x, row0 = blockIdx.x * 64;
    if (t < 64) {
        const int e = row0 + t;
        if (e >= E) return;
        float4 a = p.sP[e], b = p.sQ[e], c = p.sV[e], u = __ldg(p.act + e);
        float w = p.sW[e], r = p.ep[e];
        int k = p.cnt[e];
        a.x += u.x; b.y += u.y; c.z += u.z; w += u.w; r += a.x; k += 8;
        p.sP[e] = a; p.sQ[e] = b; p.sV[e] = c; p.sW[e] = w; p.ep[e] = r; p.cnt[e] = k;
        float4* o = p.out + (size_t)e * 18;
        o[0] = a; o[1] = b; o[2] = c; o[17] = u;
        p.rew[e] = r; p.te[e] = (unsigned char)(k & 1); p.tr[e] = (unsigned char)(k & 2);
    } else {
        const int l = t - 64;                       // 32 lanes: 8 lanes per row, 4 rows per pass
        for (int r = l >> 3; r < 64; r += 4) {
            const int e = row0 + r;
            if (e >= E) break;
            const float4* s = p.prev + (size_t)e * 18 + 4;      // old slots 1..14
            float4* d = p.out + (size_t)e * 18 + 3;             // new slots 0..13
            float4 v0 = s[l & 7], v1 = (l & 7) + 8 < 14 ? s[(l & 7) + 8] : make_float4(0, 0, 0, 0);
            d[l & 7] = v0;
            if ((l & 7) + 8 < 14) d[(l & 7) + 8] = v1;
        }
    }
}

// The same traffic moved ONLY by bulk asynchronous copies (cp.async.bulk, 1-D): per 64-env tile one thread loads the
// shifted observation tile (18 KB, contiguous) and the state / action tiles into shared memory, the threads patch the rows
// (kin + newest action) in shared memory, one thread stores everything back with bulk copies.
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, unsigned bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}
template <int PDL>
__global__ void __launch_bounds__(64) bulk_pattern_kernel(Pat p, int E)
{
    extern __shared__ __align__(128) unsigned char sm[];
    __shared__ unsigned long long bar;
    float4* obs = reinterpret_cast<float4*>(sm);                    // [64][18] float4
    float4* sP = obs + 64 * 18; float4* sQ = sP + 64; float4* sV = sQ + 64; float4* act = sV + 64;
    float* sW = reinterpret_cast<float*>(act + 64); float* ep = sW + 64; int* cnt = reinterpret_cast<int*>(ep + 64);
    float* rew = reinterpret_cast<float*>(cnt + 64);
    unsigned char* te = reinterpret_cast<unsigned char*>(rew + 64); unsigned char* tr = te + 64;
    if (PDL) asm volatile("griddepcontrol.launch_dependents;");
    const int t = threadIdx.x, row0 = blockIdx.x * 64;
    const bool last = row0 + 64 >= E;
    if (t == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const unsigned ob = 64 * 288 - (last ? 16 : 0);
        const unsigned total = ob + 4 * 1024 + 3 * 256;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(&bar)), "r"(total) : "memory");
        bulk_g2s(obs, reinterpret_cast<const char*>(p.prev + (size_t)row0 * 18) + 16, ob, &bar);
        bulk_g2s(sP, p.sP + row0, 1024, &bar); bulk_g2s(sQ, p.sQ + row0, 1024, &bar); bulk_g2s(sV, p.sV + row0, 1024, &bar);
        bulk_g2s(act, p.act + row0, 1024, &bar);
        bulk_g2s(sW, p.sW + row0, 256, &bar); bulk_g2s(ep, p.ep + row0, 256, &bar); bulk_g2s(cnt, p.cnt + row0, 256, &bar);
    }
    __syncthreads();
    unsigned ok = 0;
    while (!ok)
        asm volatile("{ .reg .pred q; mbarrier.try_wait.parity.shared::cta.b64 q, [%1], 0; selp.u32 %0, 1, 0, q; }" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
    float4 a = sP[t], b = sQ[t], c = sV[t], u = act[t];
    a.x += u.x; b.y += u.y; c.z += u.z;
    sP[t] = a; sQ[t] = b; sV[t] = c; sW[t] += u.w; ep[t] += a.x; cnt[t] += 8;
    obs[t * 18] = a; obs[t * 18 + 1] = b; obs[t * 18 + 2] = c; obs[t * 18 + 17] = u;
    rew[t] = a.x; te[t] = (unsigned char)(cnt[t] & 1); tr[t] = (unsigned char)(cnt[t] & 2);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (t == 0) {
        bulk_s2g(p.out + (size_t)row0 * 18, obs, 64 * 288);
        bulk_s2g(p.sP + row0, sP, 1024); bulk_s2g(p.sQ + row0, sQ, 1024); bulk_s2g(p.sV + row0, sV, 1024);
        bulk_s2g(p.sW + row0, sW, 256); bulk_s2g(p.ep + row0, ep, 256); bulk_s2g(p.cnt + row0, cnt, 256);
        bulk_s2g(p.rew + row0, rew, 256); bulk_s2g(p.te + row0, te, 64); bulk_s2g(p.tr + row0, tr, 64);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
}

static void run_pattern(cudaStream_t st, int E)
{
    const int NS = 8;
    std::vector<Pat> P(NS), Q(NS);
    auto mk = [&](Pat& p, float4* other_obs) {
        CK(cudaMalloc(&p.sP, (size_t)E * 16)); CK(cudaMalloc(&p.sQ, (size_t)E * 16)); CK(cudaMalloc(&p.sV, (size_t)E * 16));
        CK(cudaMalloc(&p.sW, (size_t)E * 4)); CK(cudaMalloc(&p.cnt, (size_t)E * 4)); CK(cudaMalloc(&p.ep, (size_t)E * 4));
        float4* a; CK(cudaMalloc(&a, (size_t)E * 16)); p.act = a;
        CK(cudaMalloc(&p.rew, (size_t)E * 4)); CK(cudaMalloc(&p.te, E)); CK(cudaMalloc(&p.tr, E));
        (void)other_obs;
    };
    std::vector<float4*> obs(2 * NS);
    for (auto& o : obs) { CK(cudaMalloc(&o, (size_t)E * 288)); CK(cudaMemset(o, 0, (size_t)E * 288)); }
    for (int k = 0; k < NS; ++k) { mk(P[k], nullptr); CK(cudaMemset(P[k].sP, 0, (size_t)E * 16)); CK(cudaMemset(P[k].sQ, 0, (size_t)E * 16)); CK(cudaMemset(P[k].sV, 0, (size_t)E * 16)); CK(cudaMemset(P[k].sW, 0, (size_t)E * 4)); CK(cudaMemset(P[k].cnt, 0, (size_t)E * 4)); CK(cudaMemset(P[k].ep, 0, (size_t)E * 4)); CK(cudaMemset((void*)P[k].act, 0, (size_t)E * 16)); }
    const size_t bulk_smem = 64 * 288 + 4 * 1024 + 4 * 256 + 128;
    CK(cudaFuncSetAttribute(bulk_pattern_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bulk_smem));
    CK(cudaFuncSetAttribute(bulk_pattern_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bulk_smem));
    for (int variant = 0; variant < 4; ++variant) {
        const int pdl = variant & 1, bulk = variant >> 1;
        cudaGraph_t g; cudaGraphExec_t ge;
        CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
        for (int k = 0; k < 2 * NS; ++k) {
            Pat p = P[k % NS];
            p.prev = obs[2 * (k % NS) + (k / NS) % 2]; p.out = obs[2 * (k % NS) + 1 - (k / NS) % 2];
            cudaLaunchConfig_t cfg{}; cfg.gridDim = dim3((E + 63) / 64); cfg.blockDim = dim3(bulk ? 64 : 96); cfg.stream = st;
            cfg.dynamicSmemBytes = bulk ? bulk_smem : 0;
            cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = pdl;
            cfg.attrs = at; cfg.numAttrs = 1;
            if (bulk) { if (pdl) CK(cudaLaunchKernelEx(&cfg, bulk_pattern_kernel<1>, p, E)); else CK(cudaLaunchKernelEx(&cfg, bulk_pattern_kernel<0>, p, E)); }
            else if (pdl) CK(cudaLaunchKernelEx(&cfg, pattern_kernel<1>, p, E)); else CK(cudaLaunchKernelEx(&cfg, pattern_kernel<0>, p, E));
        }
        CK(cudaStreamEndCapture(st, &g));
        CK(cudaGraphInstantiate(&ge, g, 0));
        for (int r = 0; r < 20; ++r) CK(cudaGraphLaunch(ge, st));
        CK(cudaStreamSynchronize(st));
        cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        const int REP = 200;
        CK(cudaEventRecord(e0, st));
        for (int r = 0; r < REP; ++r) CK(cudaGraphLaunch(ge, st));
        CK(cudaEventRecord(e1, st));
        CK(cudaStreamSynchronize(st));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        const double us = ms * 1e3 / (REP * 2 * NS);
        printf("{\"test\": \"%s\", \"envs\": %d, \"pdl\": %d, \"us_per_launch\": %.3f, \"GBps_algorithmic_646B\": %.0f}\n",
               bulk ? "step_access_pattern_bulk_copies" : "step_access_pattern_no_deps", E, pdl, us, (double)E * 646 / (us * 1e-6) / 1e9);
        fflush(stdout);
    }
}

int main(int argc, char** argv)
{
    const int E = argc > 1 ? atoi(argv[1]) : 65536, NS = 8, ROWS = 64;
    const int rdB = 300, wrB = 354;
    CK(cudaSetDevice(0));
    cudaStream_t st; CK(cudaStreamCreate(&st));
    run_pattern(st, E);
    for (int threads : {96, 128, 256}) {
        const int grid = E / ROWS;
        const int rd4 = ROWS * rdB / 16, wr4 = (ROWS * wrB + 15) / 16;
        std::vector<float4*> S(NS), D(NS);
        for (int k = 0; k < NS; ++k) { CK(cudaMalloc(&S[k], (size_t)grid * rd4 * 16)); CK(cudaMalloc(&D[k], (size_t)grid * wr4 * 16)); CK(cudaMemset(S[k], 0, (size_t)grid * rd4 * 16)); }
        for (int pdl = 0; pdl < 2; ++pdl) {
            cudaGraph_t g; cudaGraphExec_t ge;
            CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
            for (int k = 0; k < 2 * NS; ++k) {
                cudaLaunchConfig_t cfg{}; cfg.gridDim = dim3(grid); cfg.blockDim = dim3(threads); cfg.stream = st;
                cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = pdl;
                cfg.attrs = at; cfg.numAttrs = 1;
                if (pdl) CK(cudaLaunchKernelEx(&cfg, stream_kernel<1>, (const float4*)S[k % NS], D[k % NS], rd4, wr4));
                else CK(cudaLaunchKernelEx(&cfg, stream_kernel<0>, (const float4*)S[k % NS], D[k % NS], rd4, wr4));
            }
            CK(cudaStreamEndCapture(st, &g));
            CK(cudaGraphInstantiate(&ge, g, 0));
            for (int r = 0; r < 20; ++r) CK(cudaGraphLaunch(ge, st));
            CK(cudaStreamSynchronize(st));
            cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
            const int REP = 200;
            CK(cudaEventRecord(e0, st));
            for (int r = 0; r < REP; ++r) CK(cudaGraphLaunch(ge, st));
            CK(cudaEventRecord(e1, st));
            CK(cudaStreamSynchronize(st));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            const double us = ms * 1e3 / (REP * 2 * NS);
            printf("{\"test\": \"plain_stream\", \"envs\": %d, \"threads\": %d, \"grid\": %d, \"pdl\": %d, \"us_per_launch\": %.3f, \"GBps\": %.0f}\n",
                   E, threads, grid, pdl, us, (double)E * (rdB + wrB) / (us * 1e-6) / 1e9);
            fflush(stdout);
        }
        for (int k = 0; k < NS; ++k) { cudaFree(S[k]); cudaFree(D[k]); }
    }
    return 0;
}
