// e2e_microbench.cu — which way should the per-step result of a 65,536-env HoverAviary step reach a HOST observation
// whose row e is the sliding window log[e][4t : 4t + 12 + 4B]?  (round 2)
//   A  contiguous D2H of [E][12] kin (+ reward/flags) then a T-thread CPU scatter into the strided rows
//   B  a GPU kernel storing kin (48 B) + newest action (16 B) straight into the mapped pinned log (zero-copy over PCIe)
//   C  cudaMemcpy2DAsync D2H with 48-byte-wide rows
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o e2e_microbench e2e_microbench.cu -lpthread
#include <cuda_runtime.h>
#include <pthread.h>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

static double now_us() { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

__global__ void busy_kernel(float* x, int n, int iters)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float v = x[i];
    for (int k = 0; k < iters; ++k) v = v * 1.0001f + 0.5f;
    x[i] = v;
}

// B1: one thread per env: 3 x float4 kin + 1 x float4 action
__global__ void export_thread_per_env(const float4* __restrict__ kin, const float4* __restrict__ act, float* __restrict__ log,
                                      long long stride_f, int off_f, int B, int E, float* __restrict__ rew_h, const float* __restrict__ rew_d)
{
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    float4* row = reinterpret_cast<float4*>(log + (long long)e * stride_f + off_f);
    row[0] = kin[3 * e]; row[1] = kin[3 * e + 1]; row[2] = kin[3 * e + 2];
    row[3 + B - 1] = act[e];
    rew_h[e] = rew_d[e];
}

// B2: four lanes per env (lanes 0..2: the three kin float4s, contiguous 48 B per env; lane 3: the action)
__global__ void export_4lanes(const float4* __restrict__ kin, const float4* __restrict__ act, float* __restrict__ log,
                              long long stride_f, int off_f, int B, int E, float* __restrict__ rew_h, const float* __restrict__ rew_d)
{
    long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int e = (int)(g >> 2), part = (int)(g & 3);
    if (e >= E) return;
    float4* row = reinterpret_cast<float4*>(log + (long long)e * stride_f + off_f);
    if (part < 3) row[part] = kin[3 * e + part];
    else row[3 + B - 1] = act[e];
    if (g < E) rew_h[g] = rew_d[g];
}

struct Pool {
    int T;
    std::vector<pthread_t> th;
    std::atomic<int> gen{0}, done{0};
    std::atomic<bool> quit{false};
    void (*fn)(int, int, void*) = nullptr;
    void* arg = nullptr;
};
struct WArg { Pool* p; int id; };
static void* worker(void* a_)
{
    WArg* a = (WArg*)a_;
    Pool* p = a->p;
    int seen = 0;
    while (true) {
        while (p->gen.load(std::memory_order_acquire) == seen) { if (p->quit.load()) return nullptr; __builtin_ia32_pause(); }
        seen = p->gen.load(std::memory_order_acquire);
        p->fn(a->id, p->T, p->arg);
        p->done.fetch_add(1, std::memory_order_release);
    }
}
static void pool_run(Pool& p, void (*fn)(int, int, void*), void* arg)
{
    p.fn = fn; p.arg = arg; p.done.store(0);
    p.gen.fetch_add(1, std::memory_order_release);
    fn(0, p.T, arg);
    while (p.done.load(std::memory_order_acquire) < p.T - 1) __builtin_ia32_pause();
}

struct Scat { const float* stage; const float* act; float* log; long long stride_f; int off_f, B, E; };
static void scatter_fn(int id, int T, void* a_)
{
    Scat* s = (Scat*)a_;
    int e0 = (int)((long long)s->E * id / T), e1 = (int)((long long)s->E * (id + 1) / T);
    for (int e = e0; e < e1; ++e) {
        float* row = s->log + (long long)e * s->stride_f + s->off_f;
        memcpy(row, s->stage + 12 * e, 48);
        memcpy(row + 12 + 4 * (s->B - 1), s->act + 4 * e, 16);
    }
}
struct Rebuild { const float* stage; const float* act; float* obs; int W, B, E; };
static void rebuild_fn(int id, int T, void* a_)
{
    Rebuild* s = (Rebuild*)a_;
    int e0 = (int)((long long)s->E * id / T), e1 = (int)((long long)s->E * (id + 1) / T);
    for (int e = e0; e < e1; ++e) {
        float* row = s->obs + (long long)e * s->W;
        memmove(row + 12, row + 16, (size_t)(s->B - 1) * 16);
        memcpy(row, s->stage + 12 * e, 48);
        memcpy(row + s->W - 4, s->act + 4 * e, 16);
    }
}

int main(int argc, char** argv)
{
    const int E = argc > 1 ? atoi(argv[1]) : 65536, B = 15, W = 12 + 4 * B, TS = 64;
    const long long stride_f = W + 4 * TS;            // 328 floats = 1312 B
    const int REP = 200;
    CK(cudaSetDevice(0));
    cudaStream_t st; CK(cudaStreamCreate(&st));
    float *d_kin, *d_act, *d_rew, *d_big, *d_busy;
    CK(cudaMalloc(&d_kin, (size_t)E * 64)); CK(cudaMalloc(&d_act, (size_t)E * 16)); CK(cudaMalloc(&d_rew, (size_t)E * 8));
    CK(cudaMalloc(&d_big, (size_t)E * W * 4 + (size_t)E * 8)); CK(cudaMalloc(&d_busy, (size_t)E * 4));
    CK(cudaMemset(d_kin, 0, (size_t)E * 64)); CK(cudaMemset(d_act, 0, (size_t)E * 16)); CK(cudaMemset(d_rew, 0, (size_t)E * 8));
    float *h_act, *h_stage, *h_log, *h_rew, *h_big;
    CK(cudaHostAlloc(&h_act, (size_t)E * 16, cudaHostAllocDefault));
    CK(cudaHostAlloc(&h_stage, (size_t)E * 64, cudaHostAllocDefault));
    CK(cudaHostAlloc(&h_log, (size_t)E * stride_f * 4, cudaHostAllocMapped));
    CK(cudaHostAlloc(&h_rew, (size_t)E * 8, cudaHostAllocMapped));
    CK(cudaHostAlloc(&h_big, (size_t)E * W * 4 + (size_t)E * 8, cudaHostAllocDefault));
    memset(h_log, 0, (size_t)E * stride_f * 4); memset(h_act, 0, (size_t)E * 16); memset(h_big, 0, (size_t)E * W * 4);
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    auto dev_time = [&](auto&& f, const char* name, double bytes) {
        for (int k = 0; k < 5; ++k) f(k);
        CK(cudaStreamSynchronize(st));
        CK(cudaEventRecord(e0, st));
        for (int k = 0; k < REP; ++k) f(k);
        CK(cudaEventRecord(e1, st));
        CK(cudaStreamSynchronize(st));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        double w0 = now_us();
        for (int k = 0; k < REP; ++k) { f(k); CK(cudaStreamSynchronize(st)); }
        double wall = (now_us() - w0) / REP;
        printf("{\"test\": \"%s\", \"dev_us\": %.2f, \"sync_wall_us\": %.2f, \"GBps_dev\": %.1f}\n", name, ms * 1e3 / REP, wall, bytes / (ms * 1e-3 / REP) / 1e9);
        fflush(stdout);
    };
    dev_time([&](int) { CK(cudaMemcpyAsync(d_act, h_act, (size_t)E * 16, cudaMemcpyHostToDevice, st)); }, "h2d_actions_1MB", E * 16.0);
    dev_time([&](int) { CK(cudaMemcpyAsync(h_stage, d_kin, (size_t)E * 54, cudaMemcpyDeviceToHost, st)); }, "d2h_contig_54B", E * 54.0);
    dev_time([&](int) { CK(cudaMemcpyAsync(h_stage, d_kin, (size_t)E * 64, cudaMemcpyDeviceToHost, st)); }, "d2h_contig_64B", E * 64.0);
    dev_time([&](int) { CK(cudaMemcpyAsync(h_big, d_big, (size_t)E * (W * 4 + 6), cudaMemcpyDeviceToHost, st)); }, "d2h_full_rows_294B", E * (W * 4 + 6.0));
    dev_time([&](int k) { CK(cudaMemcpy2DAsync(h_log + 4 * (k % TS), stride_f * 4, d_kin, 48, 48, E, cudaMemcpyDeviceToHost, st)); }, "d2h_2d_width48", E * 48.0);
    dev_time([&](int k) { CK(cudaMemcpy2DAsync(h_log + 4 * (k % TS), stride_f * 4, d_kin, 64, 64, E, cudaMemcpyDeviceToHost, st)); }, "d2h_2d_width64", E * 64.0);
    dev_time([&](int k) { CK(cudaMemcpy2DAsync(h_log, stride_f * 4, d_big, W * 4, W * 4, E, cudaMemcpyDeviceToHost, st)); }, "d2h_2d_width288_compaction", E * W * 4.0);
    float* dl; CK(cudaHostGetDevicePointer(&dl, h_log, 0));
    float* dr; CK(cudaHostGetDevicePointer(&dr, h_rew, 0));
    dev_time([&](int k) { export_thread_per_env<<<(E + 127) / 128, 128, 0, st>>>((const float4*)d_kin, (const float4*)d_act, dl, stride_f, 4 * (k % TS), B, E, dr, d_rew); }, "zerocopy_thread_per_env", E * 68.0);
    dev_time([&](int k) { export_4lanes<<<(4 * E + 127) / 128, 128, 0, st>>>((const float4*)d_kin, (const float4*)d_act, dl, stride_f, 4 * (k % TS), B, E, dr, d_rew); }, "zerocopy_4lanes", E * 68.0);
    dev_time([&](int k) { export_4lanes<<<(4 * E + 255) / 256, 256, 0, st>>>((const float4*)d_kin, (const float4*)d_act, dl, stride_f, 4 * ((2 * k) % TS), B, E, dr, d_rew); }, "zerocopy_4lanes_32Baligned", E * 68.0);
    // full per-step pipelines, wall clock: H2D actions -> ~11 us kernel -> export -> sync
    auto pipe = [&](auto&& f, const char* name) {
        for (int k = 0; k < 5; ++k) f(k);
        double w0 = now_us();
        for (int k = 0; k < REP; ++k) f(k);
        printf("{\"test\": \"%s\", \"wall_us_per_step\": %.2f}\n", name, (now_us() - w0) / REP);
        fflush(stdout);
    };
    pipe([&](int k) {
        CK(cudaMemcpyAsync(d_act, h_act, (size_t)E * 16, cudaMemcpyHostToDevice, st));
        busy_kernel<<<(E + 127) / 128, 128, 0, st>>>(d_busy, E, 3000);
        export_4lanes<<<(4 * E + 127) / 128, 128, 0, st>>>((const float4*)d_kin, (const float4*)d_act, dl, stride_f, 4 * (k % TS), B, E, dr, d_rew);
        CK(cudaStreamSynchronize(st));
    }, "pipeline_B_zerocopy");
    pipe([&](int k) {
        CK(cudaMemcpyAsync(d_act, h_act, (size_t)E * 16, cudaMemcpyHostToDevice, st));
        busy_kernel<<<(E + 127) / 128, 128, 0, st>>>(d_busy, E, 3000);
        CK(cudaMemcpyAsync(h_big, d_big, (size_t)E * (W * 4 + 6), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
    }, "pipeline_round1_full_rows");
    pipe([&](int k) {
        busy_kernel<<<(E + 127) / 128, 128, 0, st>>>(d_busy, E, 3000);
        CK(cudaStreamSynchronize(st));
    }, "pipeline_kernel_only");
    for (int T : {1, 2, 4, 8, 16}) {
        Pool p; p.T = T;
        std::vector<WArg> wa(T);
        p.th.resize(T);
        for (int i = 1; i < T; ++i) { wa[i] = { &p, i }; pthread_create(&p.th[i], nullptr, worker, &wa[i]); }
        Scat sc{ h_stage, h_act, h_log, stride_f, 0, B, E };
        Rebuild rb{ h_stage, h_act, h_big, W, B, E };
        char name[96];
        double w0 = now_us();
        for (int k = 0; k < REP; ++k) { sc.off_f = 4 * (k % TS); pool_run(p, scatter_fn, &sc); }
        printf("{\"test\": \"cpu_scatter_window\", \"threads\": %d, \"wall_us\": %.2f}\n", T, (now_us() - w0) / REP);
        w0 = now_us();
        for (int k = 0; k < 50; ++k) pool_run(p, rebuild_fn, &rb);
        printf("{\"test\": \"cpu_rebuild_rows\", \"threads\": %d, \"wall_us\": %.2f}\n", T, (now_us() - w0) / 50);
        snprintf(name, sizeof name, "pipeline_A_d2h54_plus_scatter_T%d", T);
        pipe([&](int k) {
            CK(cudaMemcpyAsync(d_act, h_act, (size_t)E * 16, cudaMemcpyHostToDevice, st));
            busy_kernel<<<(E + 127) / 128, 128, 0, st>>>(d_busy, E, 3000);
            CK(cudaMemcpyAsync(h_stage, d_kin, (size_t)E * 54, cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            sc.off_f = 4 * (k % TS); pool_run(p, scatter_fn, &sc);
        }, name);
        p.quit.store(true);
        for (int i = 1; i < T; ++i) pthread_join(p.th[i], nullptr);
    }
    // does the CPU see what the GPU wrote (coherence sanity)?
    CK(cudaMemset(d_kin, 0x3f, (size_t)E * 64));
    export_4lanes<<<(4 * E + 127) / 128, 128, 0, st>>>((const float4*)d_kin, (const float4*)d_act, dl, stride_f, 8, B, E, dr, d_rew);
    CK(cudaStreamSynchronize(st));
    unsigned u; memcpy(&u, h_log + (long long)(E - 1) * stride_f + 8, 4);
    printf("{\"test\": \"coherence\", \"last_row_word\": \"0x%08x\", \"expect\": \"0x3f3f3f3f\"}\n", u);
    return 0;
}
