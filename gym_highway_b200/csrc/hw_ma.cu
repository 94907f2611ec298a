// Multi-agent highway path for sm_100a: C-ABI entry points, argument checks, kernel selection, and the reset / observe kernels
// (A <= 5 cars - 8 with the extension -, K <= 16 obstacle slots).  The step kernels themselves are hw_ma_tpe.cu (a thread per
// environment, cars in registers) and hw_ma_grp.cu (a lane per car).
//
// Replaces, for a whole batch per launch, the reference's
//   MultiAgentEnv.step / MultiAgentEnvContinuous.step  gym_highway/multiagent_envs/highway_env.py:63-99, highway_ma_c_env.py:56-82
//   HighwaySimulator.act (tick, spawn, retire)         gym_highway/multiagent_envs/highway_core.py:245-399
//   executeAction*, maintain, turn_*                   highway_core.py:110-188
//   Car.update_d / update_c, Obstacle.update           gym_highway/multiagent_envs/agent.py:79-165, :223-260
//   check_collisions + Scenario.rewards                highway_core.py:401-422, simple_base.py:65-80
//   Scenario.observations + _get_obs_norm              simple_base.py:82-154, highway_env.py:139-149
//   reset + DummyVecMAEnv auto-reset + Monitor         highway_core.py:190-243, dummy_vec_ma_env.py:66-79, bench/monitor.py:58-79
//
// The reset / observe kernels below run one thread per environment with the fp32 state promoted to fp64 in shared memory
// ([field][thread] layout: every access is conflict free); they are not on the per-step path.
#include "hw_ma_common.cuh"

namespace hw {

#define CAR(a, f) sm_car[((a) * C_NF + (f)) * kMaBlock + threadIdx.x]
#define OB(k, f) sm_ob[((k) * O_NF + (f)) * kMaBlock + threadIdx.x]

// start grid: simple_base.py:32-43 ("X" formation, three obstacles behind the cars); counters of the episode cleared
__device__ __forceinline__ void ma_reset_env(MaHead &h, double *sm_car, double *sm_ob, uint32_t *cbits, int A)
{
    for (int a = 0; a < A; ++a) {
        const double cx = ma_start_x(a);
        const int cl = ma_start_lane(a);
        CAR(a, C_PX) = cx; CAR(a, C_PY) = lane_centre(cl - 1); CAR(a, C_RX) = cx;
        CAR(a, C_VX) = 0.0; CAR(a, C_VY) = 0.0; CAR(a, C_ACC) = 0.0; CAR(a, C_STEER) = 0.0; CAR(a, C_CRUISE) = 0.0;
        cbits[a] = (uint32_t)cl << HW_BITS_LANE_SHIFT;
    }
    const double ox[3] = {-20.0, -25.0, -40.0}, ov[3] = {7.0, 5.0, 6.0};
    for (int k = 0; k < 3; ++k) { OB(k, O_PX) = ox[k]; OB(k, O_RX) = ox[k]; OB(k, O_VX) = ov[k]; OB(k, O_IV) = ov[k]; }
    h.nobs = 3; h.olane = 1u | (2u << 2) | (3u << 4); h.ofresh = 7u;
    h.newest[0] = 0; h.newest[1] = 1; h.newest[2] = 2;
    h.tick = 0; h.leader = 0; h.nc_obs = 0; h.nc_agent = 0; h.ep_len = 0; h.ep_ret = 0.0; h.flags = 0;
}

// Scenario.observations for every car, written straight to global memory (row = env, car)
__device__ __forceinline__ void ma_observe(MaHead &h, const double *sm_car, const double *sm_ob, const int A,
                                           float *__restrict__ out /* this env's [A][4A+14] block */)
{
    const int S = A + 2, D = 4 * A + 14;
    for (int a = 0; a < A; ++a) {
        float *ob = out + a * D;
        const double rx = CAR(a, C_RX), py = CAR(a, C_PY), vx = CAR(a, C_VX), vy = CAR(a, C_VY);
        ob[0] = norm_f64(CAR(a, C_ACC), -5.0, 5.0);
        ob[1] = norm_f64(CAR(a, C_STEER), -30.0, 30.0);
        ob[2] = norm_f64(rx, 0.0, 1200.0);
        ob[3] = norm_f64(py, 0.0, 8.0);
        ob[4 + 2 * S] = norm_f64(vx, -20.0, 20.0);
        ob[5 + 2 * S] = norm_f64(vy, -20.0, 20.0);
        // other cars: slot j - (j > a)
        for (int j = 0; j < A; ++j) {
            if (j == a) continue;
            const int s = j - (j > a ? 1 : 0);
            ob[4 + 2 * s] = norm_f64(dsub(CAR(j, C_RX), rx), -60.0, 60.0);
            ob[5 + 2 * s] = norm_f64(dsub(CAR(j, C_PY), py), -8.0, 8.0);
            ob[6 + 2 * S + 2 * s] = norm_f64(dsub(CAR(j, C_VX), vx), -20.0, 20.0);
            ob[7 + 2 * S + 2 * s] = norm_f64(dsub(CAR(j, C_VY), vy), -20.0, 20.0);
        }
        // per lane the obstacle nearest in raw x (first wins ties, insertion order): slot A-1+lane
        int near[3] = {-1, -1, -1};
        double nd[3] = {0.0, 0.0, 0.0};
        for (int k = 0; k < (int)h.nobs; ++k) {
            const int l = ob_lane(h, k) - 1;
            const double d = fabs(dsub(rx, OB(k, O_RX)));
            const bool take = (l == 0) ? (near[0] < 0 || d < nd[0]) : (l == 1) ? (near[1] < 0 || d < nd[1]) : (near[2] < 0 || d < nd[2]);
            if (take) { if (l == 0) { near[0] = k; nd[0] = d; } else if (l == 1) { near[1] = k; nd[1] = d; } else { near[2] = k; nd[2] = d; } }
        }
#pragma unroll
        for (int l = 0; l < 3; ++l) {
            const int s = A - 1 + l;
            const int k = near[l];
            if (k < 0) {  // the reference raises here; flag it and emit the centre of the range
                h.flags |= 2u;
                ob[4 + 2 * s] = 0.5f; ob[5 + 2 * s] = 0.5f; ob[6 + 2 * S + 2 * s] = 0.5f; ob[7 + 2 * S + 2 * s] = 0.5f;
                continue;
            }
            ob[4 + 2 * s] = norm_f64(dsub(OB(k, O_RX), rx), -60.0, 60.0);
            ob[5 + 2 * s] = norm_f64(dsub(lane_centre(l), py), -8.0, 8.0);
            ob[6 + 2 * S + 2 * s] = norm_f64(dsub(OB(k, O_VX), vx), -20.0, 20.0);
            ob[7 + 2 * S + 2 * s] = norm_f64(dsub(0.0, vy), -20.0, 20.0);
        }
    }
}

__device__ __forceinline__ void ma_load(MaHead &h, double *sm_car, double *sm_ob, uint32_t *cbits, const MaPtrs &S, int64_t i, int64_t n)
{
    ma_load_head(h, S, i, n);
    for (int a = 0; a < S.A; ++a) {
#pragma unroll
        for (int f = 0; f < C_NF; ++f) CAR(a, f) = (double)S.cars[((int64_t)(a * 9 + f)) * n + i];
        cbits[a] = __float_as_uint(S.cars[((int64_t)(a * 9 + 8)) * n + i]);
    }
    for (int k = 0; k < (int)h.nobs; ++k) {
        const float4 o = S.obst[(int64_t)k * n + i];
        OB(k, O_PX) = o.x; OB(k, O_RX) = o.y; OB(k, O_VX) = o.z; OB(k, O_IV) = o.w;
    }
}

__device__ __forceinline__ void ma_store(const MaHead &h, const double *sm_car, const double *sm_ob, const uint32_t *cbits,
                                         const MaPtrs &S, int64_t i, int64_t n)
{
    ma_store_head(h, S, i, n);
    for (int a = 0; a < S.A; ++a) {
#pragma unroll
        for (int f = 0; f < C_NF; ++f) S.cars[((int64_t)(a * 9 + f)) * n + i] = (float)CAR(a, f);
        S.cars[((int64_t)(a * 9 + 8)) * n + i] = __uint_as_float(cbits[a]);
    }
    for (int k = 0; k < (int)h.nobs; ++k)
        S.obst[(int64_t)k * n + i] = make_float4((float)OB(k, O_PX), (float)OB(k, O_RX), (float)OB(k, O_VX), (float)OB(k, O_IV));
}

__global__ void __launch_bounds__(kMaBlock)
ma_reset_kernel(MaPtrs S, const uint8_t *__restrict__ mask, float *__restrict__ obs, const int64_t n)
{
    extern __shared__ double sm_dyn[];
    double *sm_car = sm_dyn;
    double *sm_ob = sm_dyn + (size_t)S.A * C_NF * kMaBlock;
    const int64_t i = (int64_t)blockIdx.x * kMaBlock + threadIdx.x;
    if (i >= n) return;
    if (mask && !mask[i]) return;
    MaHead h;
    uint32_t cbits[kMaxAExt];
    ma_load_head(h, S, i, n);  // spawn_ctr survives
    ma_reset_env(h, sm_car, sm_ob, cbits, S.A);
    if (obs) ma_observe(h, sm_car, sm_ob, S.A, obs + i * (int64_t)S.A * (4 * S.A + 14));
    ma_store(h, sm_car, sm_ob, cbits, S, i, n);
}

__global__ void __launch_bounds__(kMaBlock)
ma_observe_kernel(MaPtrs S, float *__restrict__ obs, const int64_t n)
{
    extern __shared__ double sm_dyn[];
    double *sm_car = sm_dyn;
    double *sm_ob = sm_dyn + (size_t)S.A * C_NF * kMaBlock;
    const int64_t i = (int64_t)blockIdx.x * kMaBlock + threadIdx.x;
    if (i >= n) return;
    MaHead h;
    uint32_t cbits[kMaxAExt];
    ma_load(h, sm_car, sm_ob, cbits, S, i, n);
    ma_observe(h, sm_car, sm_ob, S.A, obs + i * (int64_t)S.A * (4 * S.A + 14));
}

static inline bool ma_aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

static int ma_check(const HwMaState *st, int64_t n)
{
    if (!st || n <= 0 || !st->cars || !st->obst || !st->head) return HW_E_BADARG;
    if (st->num_agents < 1 || st->num_agents > kMaxAExt || st->num_slots < 3 || st->num_slots > kMaxK) return HW_E_BADARG;
    if (!ma_aligned16(st->obst)) return HW_E_ALIGN;
    return 0;
}

static MaPtrs ma_ptrs(const HwMaState *st)
{
    MaPtrs S;
    S.cars = st->cars; S.obst = reinterpret_cast<float4 *>(st->obst); S.head = st->head;
    S.A = st->num_agents; S.K = st->num_slots;
    return S;
}

static size_t ma_smem(const HwMaState *st)
{
    return ((size_t)st->num_agents * C_NF + (size_t)st->num_slots * O_NF) * kMaBlock * sizeof(double);
}

template <typename Kern>
static int ma_prepare(Kern kern, size_t smem)
{
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    return 0;
}

static int ma_launch_step(const HwMaState *st, const void *actions, const HwMaOut *out, const HwSimParams *P,
                          uint64_t seed, uint64_t env_id0, int64_t n, int flags, cudaStream_t stream)
{
    int rc = ma_check(st, n);
    if (rc) return rc;
    if (!actions || !out || !out->obs || !out->rew || !out->done || !P) return HW_E_BADARG;
    if (P->ticks_per_step < 1 || P->timeout_tick < 1 || P->timeout_tick > 0xFFFF || !(P->dt > 0.0)) return HW_E_BADARG;
    MaConsts C_;
    C_.dt = P->dt; C_.k30dt = 30.0 * P->dt; C_.ticks_per_step = P->ticks_per_step; C_.timeout_tick = P->timeout_tick;
    // two layouts of the same step (identical results): a thread per environment with the cars in registers (hw_ma_tpe.cu: fewest
    // instructions per environment, 252 us against 379 us at 262 144 environments) or a lane per car (hw_ma_grp.cu: five times the
    // threads, the better fit for the smallest batches; the two meet at 16 384 environments - 41 us each - and at 32 768 it is 50 us
    // against 66 us)
    if (st->num_agents > kMaxA && (flags & HW_FLAG_MA_LANE_PER_CAR)) return HW_E_BADARG;  // the extension (6..8 cars) has one kernel
    // measured crossover (B200): the lane-per-car kernel always spends five lanes per environment, so the fewer the cars the
    // earlier the thread-per-environment kernel wins - at every size for one or two cars, from HW_MA_TPE_MIN_ENVS
    // environments for five cars, a third / two thirds of that for three / four
    const int64_t tpe_min = (st->num_agents <= 2) ? 0 : (int64_t)(HW_MA_TPE_MIN_ENVS / 3) * (st->num_agents - 2);
    const bool tpe = (flags & HW_FLAG_MA_THREAD_PER_ENV) || st->num_agents > kMaxA
                     || (!(flags & HW_FLAG_MA_LANE_PER_CAR) && n >= tpe_min);
    if (tpe) return ma_tpe_launch_step(st, actions, out, C_, seed, env_id0, n, flags, stream);
    return ma_grp_launch_step(st, actions, out, C_, seed, env_id0, n, flags, stream);
}

}  // namespace hw

extern "C" {

int hw_ma_reset(const HwMaState *state, const uint8_t *mask, float *obs, int64_t n, void *stream)
{
    int rc = hw::ma_check(state, n);
    if (rc) return rc;
    const size_t smem = hw::ma_smem(state);
    if ((rc = hw::ma_prepare(hw::ma_reset_kernel, smem))) return rc;
    const unsigned grid = (unsigned)((n + hw::kMaBlock - 1) / hw::kMaBlock);
    hw::ma_reset_kernel<<<grid, hw::kMaBlock, smem, (cudaStream_t)stream>>>(hw::ma_ptrs(state), mask, obs, n);
    return (int)cudaGetLastError();
}

int hw_ma_step(const HwMaState *state, const void *actions, const HwMaOut *out, const HwSimParams *params,
               uint64_t seed, uint64_t env_id0, int64_t n, int flags, void *stream)
{
    return hw::ma_launch_step(state, actions, out, params, seed, env_id0, n, flags, (cudaStream_t)stream);
}

int hw_ma_observe(const HwMaState *state, float *obs, int64_t n, void *stream)
{
    int rc = hw::ma_check(state, n);
    if (rc) return rc;
    if (!obs) return HW_E_BADARG;
    const size_t smem = hw::ma_smem(state);
    if ((rc = hw::ma_prepare(hw::ma_observe_kernel, smem))) return rc;
    const unsigned grid = (unsigned)((n + hw::kMaBlock - 1) / hw::kMaBlock);
    hw::ma_observe_kernel<<<grid, hw::kMaBlock, smem, (cudaStream_t)stream>>>(hw::ma_ptrs(state), obs, n);
    return (int)cudaGetLastError();
}

int hw_ma_step_host(const HwMaState *state, const void *h_actions, void *d_actions, const HwMaOut *d_out,
                    float *h_obs, float *h_rew, uint8_t *h_done, const HwSimParams *params,
                    uint64_t seed, uint64_t env_id0, int64_t n, int flags, void *stream)
{
    if (!state || !h_actions || !d_actions || !d_out || !h_obs || !h_rew || !h_done || n <= 0) return HW_E_BADARG;
    cudaStream_t s = (cudaStream_t)stream;
    const int A = state->num_agents;
    const size_t abytes = (size_t)n * A * ((flags & HW_FLAG_CONTINUOUS) ? 2 * sizeof(float) : sizeof(int32_t));
    cudaError_t ce = cudaMemcpyAsync(d_actions, h_actions, abytes, cudaMemcpyHostToDevice, s);
    if (ce != cudaSuccess) return (int)ce;
    int rc = hw::ma_launch_step(state, d_actions, d_out, params, seed, env_id0, n, flags, s);
    if (rc) return rc;
    ce = cudaMemcpyAsync(h_obs, d_out->obs, (size_t)n * A * (4 * A + 14) * sizeof(float), cudaMemcpyDeviceToHost, s);
    if (ce != cudaSuccess) return (int)ce;
    ce = cudaMemcpyAsync(h_rew, d_out->rew, (size_t)n * A * sizeof(float), cudaMemcpyDeviceToHost, s);
    if (ce != cudaSuccess) return (int)ce;
    ce = cudaMemcpyAsync(h_done, d_out->done, (size_t)n * A, cudaMemcpyDeviceToHost, s);
    if (ce != cudaSuccess) return (int)ce;
    return (int)cudaStreamSynchronize(s);
}

}  // extern "C"
