// gpd_launch.inl — launchers, compiled once per precision with GPD_REAL defined (gpd_f32.cu / gpd_f64.cu).
#include <atomic>
#include <mutex>

#include "gpd_kernels.cuh"
#include "gpd_step_bulk.cuh"

namespace gpd {

using Real = GPD_REAL;

// The opt-in dynamic shared-memory limit of a kernel is a per-device (per-context) attribute that only ever grows here
// (handles of different block sizes share it).  Tracked per device ordinal; the slow path is serialised.
struct SmemLimit {
    std::atomic<size_t> cur[GPD_MAX_DEVICES];
    std::mutex mu;
    SmemLimit() { for (auto& c : cur) c.store(48 * 1024); }
    template <typename F> cudaError_t ensure(size_t need, F kernel)
    {
        int dev = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        if (dev < 0 || dev >= GPD_MAX_DEVICES) return cudaErrorInvalidDevice;
        if (need <= cur[dev].load(std::memory_order_acquire)) return cudaSuccess;
        std::lock_guard<std::mutex> lock(mu);
        if (need <= cur[dev].load(std::memory_order_relaxed)) return cudaSuccess;
        e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need);
        if (e == cudaSuccess) cur[dev].store(need, std::memory_order_release);
        return e;
    }
};

template <int KIND, bool MULTI, bool VEC>
static cudaError_t ensure_step_smem(size_t need)
{
    static SmemLimit lim;
    return lim.ensure(need, step_kernel<Real, KIND, MULTI, VEC>);
}

template <int KIND, bool MULTI, bool VEC>
static cudaError_t launch_step_t(const StepArgs<Real>& a, const LaunchCfg& lc, const CUtensorMap& tp, const CUtensorMap& to,
                                 const CUtensorMap& te, cudaStream_t st)
{
    cudaError_t e = ensure_step_smem<KIND, MULTI, VEC>(lc.smem);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)lc.grid);
    cfg.blockDim = dim3((unsigned)lc.threads);
    cfg.dynamicSmemBytes = lc.smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = lc.pdl ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, step_kernel<Real, KIND, MULTI, VEC>, a, tp, to, te);
}

static int step_variant(int action_type, int phy, int N, int A, int W, int env_kind)
{
    const bool rpm_like = action_type == GPD_ACT_RPM || action_type == GPD_ACT_ONE_D_RPM || action_type == GPD_ACT_CTRL_RPM;
    const int kind = !rpm_like ? GPD_K_PID : (phy == 0 ? GPD_K_LEAN : GPD_K_FORCES);
    const bool multi = N > 1;
    const bool vec = A == 4 && env_kind != GPD_ENV_CTRL && (W % 4 == 0);
    return kind * 4 + (multi ? 2 : 0) + (vec ? 1 : 0);
}

template <int KIND, bool MULTI, bool VEC>
static int occupancy_t(int threads, size_t smem)
{
    if (ensure_step_smem<KIND, MULTI, VEC>(smem) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, step_kernel<Real, KIND, MULTI, VEC>, threads, smem) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

// Resident CTAs per SM of the step-kernel variant this configuration dispatches to (registers, threads, shared memory).
template <>
int step_blocks_per_sm<Real>(int action_type, int phy, int N, int A, int W, int env_kind, int threads, size_t smem)
{
    switch (step_variant(action_type, phy, N, A, W, env_kind)) {
    case 0: return occupancy_t<GPD_K_FORCES, false, false>(threads, smem);
    case 1: return occupancy_t<GPD_K_FORCES, false, true>(threads, smem);
    case 2: return occupancy_t<GPD_K_FORCES, true, false>(threads, smem);
    case 3: return occupancy_t<GPD_K_FORCES, true, true>(threads, smem);
    case 4: return occupancy_t<GPD_K_LEAN, false, false>(threads, smem);
    case 5: return occupancy_t<GPD_K_LEAN, false, true>(threads, smem);
    case 6: return occupancy_t<GPD_K_LEAN, true, false>(threads, smem);
    case 7: return occupancy_t<GPD_K_LEAN, true, true>(threads, smem);
    case 8: return occupancy_t<GPD_K_PID, false, false>(threads, smem);
    case 9: return occupancy_t<GPD_K_PID, false, true>(threads, smem);
    case 10: return occupancy_t<GPD_K_PID, true, false>(threads, smem);
    case 11: return occupancy_t<GPD_K_PID, true, true>(threads, smem);
    default: return 0;
    }
}

template <>
cudaError_t launch_step<Real>(const StepArgs<Real>& a, const LaunchCfg& lc, const CUtensorMap* tm_prev,
                              const CUtensorMap* tm_out, const CUtensorMap* tm_edge, cudaStream_t st)
{
    static const CUtensorMap dummy{};
    const CUtensorMap& tp = tm_prev ? *tm_prev : dummy;
    const CUtensorMap& to = tm_out ? *tm_out : dummy;
    const CUtensorMap& te = tm_edge ? *tm_edge : dummy;
    switch (step_variant(a.action_type, a.phy, a.N, a.A, a.W, a.env_kind)) {
    case 0: return launch_step_t<GPD_K_FORCES, false, false>(a, lc, tp, to, te, st);
    case 1: return launch_step_t<GPD_K_FORCES, false, true>(a, lc, tp, to, te, st);
    case 2: return launch_step_t<GPD_K_FORCES, true, false>(a, lc, tp, to, te, st);
    case 3: return launch_step_t<GPD_K_FORCES, true, true>(a, lc, tp, to, te, st);
    case 4: return launch_step_t<GPD_K_LEAN, false, false>(a, lc, tp, to, te, st);
    case 5: return launch_step_t<GPD_K_LEAN, false, true>(a, lc, tp, to, te, st);
    case 6: return launch_step_t<GPD_K_LEAN, true, false>(a, lc, tp, to, te, st);
    case 7: return launch_step_t<GPD_K_LEAN, true, true>(a, lc, tp, to, te, st);
    case 8: return launch_step_t<GPD_K_PID, false, false>(a, lc, tp, to, te, st);
    case 9: return launch_step_t<GPD_K_PID, false, true>(a, lc, tp, to, te, st);
    case 10: return launch_step_t<GPD_K_PID, true, false>(a, lc, tp, to, te, st);
    case 11: return launch_step_t<GPD_K_PID, true, true>(a, lc, tp, to, te, st);
    default: return cudaErrorInvalidValue;
    }
}

// ---- bulk-copy data path of the RL envs without downwash (gpd_step_bulk.cuh) ----
template <int KIND, bool MULTI>
static cudaError_t ensure_bulk_smem(size_t need)
{
    static SmemLimit lim;
    return lim.ensure(need, step_kernel_bulk<Real, KIND, MULTI>);
}

static int bulk_kind(int action_type, int phy)
{
    const bool rpm_like = action_type == GPD_ACT_RPM || action_type == GPD_ACT_ONE_D_RPM;
    return !rpm_like ? GPD_K_PID : (phy == 0 ? GPD_K_LEAN : GPD_K_FORCES);
}

template <int KIND, bool MULTI>
static int bulk_occupancy_t(int threads, size_t smem)
{
    if (ensure_bulk_smem<KIND, MULTI>(smem) != cudaSuccess) { cudaGetLastError(); return 0; }
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, step_kernel_bulk<Real, KIND, MULTI>, threads, smem) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

template <>
int step_bulk_blocks_per_sm<Real>(int action_type, int phy, int N, int threads, size_t smem)
{
    const int k = bulk_kind(action_type, phy);
    if (N > 1) {
        if (k == GPD_K_FORCES) return bulk_occupancy_t<GPD_K_FORCES, true>(threads, smem);
        if (k == GPD_K_LEAN) return bulk_occupancy_t<GPD_K_LEAN, true>(threads, smem);
        return bulk_occupancy_t<GPD_K_PID, true>(threads, smem);
    }
    if (k == GPD_K_FORCES) return bulk_occupancy_t<GPD_K_FORCES, false>(threads, smem);
    if (k == GPD_K_LEAN) return bulk_occupancy_t<GPD_K_LEAN, false>(threads, smem);
    return bulk_occupancy_t<GPD_K_PID, false>(threads, smem);
}

template <int KIND, bool MULTI>
static cudaError_t launch_step_bulk_t(const StepArgs<Real>& a, const LaunchCfg& lc, cudaStream_t st)
{
    cudaError_t e = ensure_bulk_smem<KIND, MULTI>(lc.smem);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)lc.grid);
    cfg.blockDim = dim3((unsigned)lc.threads);
    cfg.dynamicSmemBytes = lc.smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = lc.pdl ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, step_kernel_bulk<Real, KIND, MULTI>, a);
}

template <>
cudaError_t launch_step_bulk<Real>(const StepArgs<Real>& a, const LaunchCfg& lc, cudaStream_t st)
{
    const int k = bulk_kind(a.action_type, a.phy);
    if (a.N > 1) {
        if (k == GPD_K_FORCES) return launch_step_bulk_t<GPD_K_FORCES, true>(a, lc, st);
        if (k == GPD_K_LEAN) return launch_step_bulk_t<GPD_K_LEAN, true>(a, lc, st);
        return launch_step_bulk_t<GPD_K_PID, true>(a, lc, st);
    }
    if (k == GPD_K_FORCES) return launch_step_bulk_t<GPD_K_FORCES, false>(a, lc, st);
    if (k == GPD_K_LEAN) return launch_step_bulk_t<GPD_K_LEAN, false>(a, lc, st);
    return launch_step_bulk_t<GPD_K_PID, false>(a, lc, st);
}

template <>
cudaError_t launch_reset<Real>(const StepArgs<Real>& a, const LaunchCfg& lc, cudaStream_t st)
{
    const bool vec = a.A == 4 && a.env_kind != GPD_ENV_CTRL && (a.W % 4 == 0);
    static SmemLimit lim[2];
    cudaError_t e = vec ? lim[1].ensure(lc.smem, reset_kernel<Real, true>) : lim[0].ensure(lc.smem, reset_kernel<Real, false>);
    if (e != cudaSuccess) return e;
    if (vec) reset_kernel<Real, true><<<(unsigned)lc.grid, lc.threads, lc.smem, st>>>(a);
    else reset_kernel<Real, false><<<(unsigned)lc.grid, lc.threads, lc.smem, st>>>(a);
    return cudaGetLastError();
}

static inline unsigned blocks_for(int64_t n, int t) { return (unsigned)((n + t - 1) / t); }

template <>
cudaError_t launch_get_state<Real>(const StepArgs<Real>& a, const float* obs_latest, Real* state20, Real* rpy_rates,
                                   Real* pid_state, int32_t* counter, cudaStream_t st)
{
    get_state_kernel<Real><<<blocks_for(a.D, 128), 128, 0, st>>>(a, obs_latest, state20, rpy_rates, pid_state, counter);
    return cudaGetLastError();
}

template <>
cudaError_t launch_set_state<Real>(const StepArgs<Real>& a, const Real* state20, const Real* rpy_rates,
                                   const Real* pid_state, const int32_t* counter, cudaStream_t st)
{
    set_state_kernel<Real><<<blocks_for(a.D, 128), 128, 0, st>>>(a, state20, rpy_rates, pid_state, counter);
    return cudaGetLastError();
}

template <>
cudaError_t launch_pid<Real>(const DevPid<Real>& c, int64_t n, Real dt, const Real* cur_pos, const Real* cur_quat,
                             const Real* cur_vel, const Real* target_pos, const Real* target_rpy, const Real* target_vel,
                             const Real* target_rates, Real* pid_state, Real* rpm_out, Real* pos_e_out, Real* yaw_e_out,
                             cudaStream_t st)
{
    pid_kernel<Real><<<blocks_for(n, 128), 128, 0, st>>>(c, n, dt, cur_pos, cur_quat, cur_vel, target_pos, target_rpy,
                                                         target_vel, target_rates, pid_state, rpm_out, pos_e_out, yaw_e_out);
    return cudaGetLastError();
}

template <>
cudaError_t launch_ground_effect<Real>(const DevDrone<Real>& d, int64_t n, const Real* rpm, const Real* pos,
                                       const Real* quat, Real* out, uint8_t* applied, cudaStream_t st)
{
    ground_effect_kernel<Real><<<blocks_for(n, 128), 128, 0, st>>>(d, n, rpm, pos, quat, out, applied);
    return cudaGetLastError();
}

template <>
cudaError_t launch_drag<Real>(const DevDrone<Real>& d, int64_t n, const Real* rpm, const Real* quat, const Real* vel,
                              Real* out, cudaStream_t st)
{
    drag_kernel<Real><<<blocks_for(n, 128), 128, 0, st>>>(d, n, rpm, quat, vel, out);
    return cudaGetLastError();
}

template <>
cudaError_t launch_downwash<Real>(const DevDrone<Real>& d, int64_t E, int N, const Real* pos, Real* out, cudaStream_t st)
{
    downwash_kernel<Real><<<blocks_for(E * N, 128), 128, 0, st>>>(d, E, N, pos, out);
    return cudaGetLastError();
}

template <>
cudaError_t launch_adjacency<Real>(const StepArgs<Real>& a, Real radius, Real* out, cudaStream_t st)
{
    const int N = a.N;
    const int EPC = N * N >= 256 ? 1 : 256 / (N * N);       // whole envs per CTA
    const size_t smem = (size_t)EPC * N * sizeof(typename Vec4<Real>::type);
    const unsigned grid = (unsigned)((a.E + EPC - 1) / EPC);
    adjacency_kernel<Real><<<grid, 256, smem, st>>>(a.p.sP, a.E, N, EPC, radius, out);
    return cudaGetLastError();
}

template <>
cudaError_t launch_rollout_pid<Real>(const StepArgs<Real>& a, int n_steps, const Real* waypoints, int n_wp,
                                     int32_t* wp_counters, Real* action, cudaStream_t st)
{
    rollout_pid_kernel<Real><<<blocks_for(a.D, 128), 128, 0, st>>>(a, n_steps, waypoints, n_wp, wp_counters, action);
    return cudaGetLastError();
}

template <>
cudaError_t launch_nonfinite<Real>(const StepArgs<Real>& a, unsigned long long* out, cudaStream_t st)
{
    nonfinite_kernel<Real><<<blocks_for(a.D, 256), 256, 0, st>>>(a, out);
    return cudaGetLastError();
}

}  // namespace gpd
