// Single-agent highway step path for sm_100a: one thread per environment.
//
// Replaces, for a whole batch per launch, the reference's
//   HighwayEnv.step / HighwayEnvContinuous.step   gym_highway/envs/highway_env.py:70-114, highway_env_cont.py:42-55
//   HighwaySimulator.act (one tick)               gym_highway/envs/multi_lane_sim.py:879-1049
//   executeActionDiscrete / Continuous            multi_lane_sim.py:502-572
//   Car.update_d / update_c, Obstacle.update      multi_lane_sim.py:140-272, :324-347
//   get_state + clip/normalise                    multi_lane_sim.py:1051-1092, highway_env.py:131-140
//   reset + DummyVecEnv auto-reset + Monitor      multi_lane_sim.py:826-877, dummy_vec_env.py:52-55, bench/monitor.py:58-79
//
// Data movement per env-step (algorithmic bytes): 6 x 16 B state loads + 6 x 16 B state
// stores (coalesced float4 SoA), 4 B action, 72 B observation (staged through shared memory so the
// global stores are full 128-bit coalesced rows), 4 B reward, 1 B done  = 273 B.
#include "hw_common.cuh"
#include "../../include/highway_b200.h"

namespace hw {

constexpr int kSaBlock = 128;

struct SaEnv {
    double px, py, rx, vx, acc, steer, cruise;
    double opx[3], orx[3], ovx[3], oiv[3];
    uint32_t bits;  // tick | lane << 16 | mode flags
    uint32_t ncoll, spawn_ctr;
    int ep_ret_milli;
    uint32_t ep_len;
};

struct SaPtrs {
    float4 *car_a, *car_b, *obst;
    uint4 *stat;
};

__device__ __forceinline__ void sa_load(SaEnv &e, const SaPtrs &S, int64_t i, int64_t n)
{
    const float4 a = S.car_a[i], b = S.car_b[i];
    const float4 o0 = S.obst[i], o1 = S.obst[n + i], o2 = S.obst[2 * n + i];
    const uint4 s = S.stat[i];
    e.px = a.x; e.py = a.y; e.rx = a.z; e.vx = a.w;
    e.acc = b.x; e.steer = b.y; e.cruise = b.z; e.bits = __float_as_uint(b.w);
    e.opx[0] = o0.x; e.orx[0] = o0.y; e.ovx[0] = o0.z; e.oiv[0] = o0.w;
    e.opx[1] = o1.x; e.orx[1] = o1.y; e.ovx[1] = o1.z; e.oiv[1] = o1.w;
    e.opx[2] = o2.x; e.orx[2] = o2.y; e.ovx[2] = o2.z; e.oiv[2] = o2.w;
    e.ncoll = s.x; e.spawn_ctr = s.y; e.ep_ret_milli = (int)s.z; e.ep_len = s.w;
}

__device__ __forceinline__ void sa_store(const SaEnv &e, const SaPtrs &S, int64_t i, int64_t n)
{
    S.car_a[i] = make_float4((float)e.px, (float)e.py, (float)e.rx, (float)e.vx);
    S.car_b[i] = make_float4((float)e.acc, (float)e.steer, (float)e.cruise, __uint_as_float(e.bits));
    S.obst[i] = make_float4((float)e.opx[0], (float)e.orx[0], (float)e.ovx[0], (float)e.oiv[0]);
    S.obst[n + i] = make_float4((float)e.opx[1], (float)e.orx[1], (float)e.ovx[1], (float)e.oiv[1]);
    S.obst[2 * n + i] = make_float4((float)e.opx[2], (float)e.orx[2], (float)e.ovx[2], (float)e.oiv[2]);
    S.stat[i] = make_uint4(e.ncoll, e.spawn_ctr, (uint32_t)e.ep_ret_milli, e.ep_len);
}

// start grid, highway_env.py:55-65; spawn_ctr survives so that every episode sees fresh obstacles
__device__ __forceinline__ void sa_reset_env(SaEnv &e)
{
    e.px = 20.0; e.py = 3.75; e.rx = 20.0; e.vx = 0.0;
    e.acc = 0.0; e.steer = 0.0; e.cruise = 0.0;
    e.bits = 2u << HW_BITS_LANE_SHIFT;
    e.opx[0] = -20.0; e.orx[0] = -20.0; e.ovx[0] = 7.0; e.oiv[0] = 7.0;
    e.opx[1] = -25.0; e.orx[1] = -25.0; e.ovx[1] = 5.0; e.oiv[1] = 5.0;
    e.opx[2] = -40.0; e.orx[2] = -40.0; e.ovx[2] = 6.0; e.oiv[2] = 6.0;
    e.ncoll = 0; e.ep_ret_milli = 0; e.ep_len = 0;
}

// One simulator tick (multi_lane_sim.py:879-1049).  Returns the number of obstacle rects the car
// overlapped (0 when the collision lock is on); the tick reward is -250*n, or vx/20-1 when n == 0.
template <bool CONT>
__device__ __forceinline__ int sa_tick(SaEnv &e, const double dt, const double k30dt, const bool inf_obs,
                                       const int act_d, const double a0_5dt, const double a1_30dt,
                                       const uint64_t seed, const uint64_t env_id)
{
    const uint32_t tick0 = e.bits & HW_BITS_TICK_MASK;
    e.bits += 1;  // tick field lives in the low 16 bits; 2160 (or 3600) never carries out
    int lane = (int)((e.bits >> HW_BITS_LANE_SHIFT) & 3u);

    // ---- apply action (multi_lane_sim.py:502-526) ----
    if (CONT) {
        e.acc = py_clamp(dadd(e.acc, a0_5dt), -5.0, 5.0);
        e.steer = py_clamp(dadd(e.steer, a1_30dt), -30.0, 30.0);
    } else {
        if (act_d == HW_ACT_ACCELERATE && !(e.bits & kBitDoAcc)) {
            e.bits = (e.bits | kBitDoAcc) & ~kBitDoMaint;
        } else if (act_d == HW_ACT_MAINTAIN && !(e.bits & kBitDoMaint)) {
            // nearest vehicle strictly ahead in the car's lane; each lane holds exactly one obstacle
            const double fpx = (lane == 1) ? e.opx[0] : (lane == 2) ? e.opx[1] : e.opx[2];
            const double fvx = (lane == 1) ? e.ovx[0] : (lane == 2) ? e.ovx[1] : e.ovx[2];
            e.cruise = (fpx > e.px) ? fvx : e.vx;
            e.bits = (e.bits | kBitDoMaint) & ~kBitDoAcc;
        }
        e.acc = py_clamp(e.acc, -5.0, 5.0);
        if (act_d == HW_ACT_RIGHT && !(e.bits & kBitRight)) {
            lane = min(lane + 1, 3);
            e.bits = (e.bits | kBitRight) & ~kBitLeft;
        } else if (act_d == HW_ACT_LEFT && !(e.bits & kBitLeft)) {
            lane = max(lane - 1, 1);
            e.bits = (e.bits | kBitLeft) & ~kBitRight;
        }
        e.steer = py_clamp(e.steer, -30.0, 30.0);
    }

    // ---- collision test on the rects the previous tick left (:920-937); skipped on the first tick ----
    int ncol = 0;
    if (tick0 != 0) {
        const int ax = rect_x(e.px), ay = rect_y(e.py);
#pragma unroll
        for (int k = 0; k < 3; ++k) ncol += rects_overlap(ax, ay, rect_x(e.opx[k]), 21 + 80 * k) ? 1 : 0;
        e.ncoll += ncol;
    }

    // ---- car update (:140-229) ----
    if (!CONT) {
        if (e.bits & kBitDoAcc) {
            e.acc = (e.acc < 0.0) ? 0.0 : dadd(e.acc, dt);
            if (e.acc == 5.0) e.bits &= ~kBitDoAcc;
        } else if (e.bits & kBitDoMaint) {
            const double vc = ceil(e.vx), cc = ceil(e.cruise);
            e.acc = (vc <= cc) ? 5.0 : -5.0;
            if (vc == cc) { e.vx = e.cruise; e.acc = 0.0; e.bits &= ~kBitDoMaint; }
        }
    }
    e.vx = py_clamp(dadd(e.vx, dmul(e.acc, dt)), 0.0, 20.0);
    if (!CONT) {
        const double ty = lane_centre(lane - 1);
        if (e.bits & kBitLeft) {
            e.steer = dadd(e.steer, k30dt);
            if (e.py <= ty) { e.steer = 0.0; e.bits &= ~kBitLeft; }
        } else if (e.bits & kBitRight) {
            e.steer = dsub(e.steer, k30dt);
            if (e.py >= ty) { e.steer = 0.0; e.bits &= ~kBitRight; }
        }
    }
    double ang = 0.0;
    if (e.steer != 0.0) ang = ddiv(e.vx, ddiv(4.0, tan(dmul(e.steer, kDeg2Rad))));
    const double vdt = dmul(e.vx, dt);
    // position += velocity * dt with velocity.y == 0: py + 0.0 == py for every py the clamps can produce
    if (CONT) e.py = dsub(e.py, dmul(ang, dt));
    else      e.py = dsub(e.py, dmul(dmul(dmul(ang, kRad2Deg), dt), dt));
    e.px = 10.0;  // the only car is always the leader
    if (CONT) { if (e.py < 1.0) e.py = 1.0; else if (e.py > 6.0) e.py = (7.0 < e.py) ? 7.0 : e.py; }
    else      { if (e.py < 0.0) e.py = 0.0; else if (e.py > 6.0) e.py = (7.0 < e.py) ? 7.0 : e.py; }
    e.rx = dadd(e.rx, vdt);
    if (CONT) {  // lane = nearest lane centre, first minimum wins
        const double d0 = fabs(dsub(e.py, 1.25)), d1 = fabs(dsub(e.py, 3.75)), d2 = fabs(dsub(e.py, 6.25));
        const double m = fmin(d0, fmin(d1, d2));
        lane = (d0 == m) ? 1 : (d1 == m) ? 2 : 3;
    }
    e.bits = (e.bits & ~(3u << HW_BITS_LANE_SHIFT)) | ((uint32_t)lane << HW_BITS_LANE_SHIFT);

    // ---- obstacle update (:324-347): sees the leader's post-update lane, position and velocity ----
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        double v = py_clamp(e.oiv[k], -20.0, 20.0);
        if (k + 1 == lane && e.opx[k] < 10.0) v = (e.vx < e.oiv[k]) ? e.vx : e.oiv[k];
        e.ovx[k] = v;
        const double odt = dmul(v, dt);
        e.opx[k] = dsub(dadd(e.opx[k], odt), vdt);
        e.orx[k] = dadd(e.orx[k], odt);
    }

    // ---- respawn (:963-991) ----
    if (inf_obs) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            if (e.opx[k] < -2.375) {
                double u1, u2;
                spawn_uniforms(seed, env_id, e.spawn_ctr, u1, u2);
                e.spawn_ctr += 1;
                const double x = dadd(70.0, dmul(30.0, u1));
                const double v = dadd(5.0, dmul(2.0, u2));
                e.opx[k] = x;
                e.orx[k] = dadd(e.rx, dsub(x, 10.0));
                e.ovx[k] = v;
                e.oiv[k] = v;
            }
        }
    }
    return ncol;
}

// observation pairs (x, y) j = 0..8 -> obs[2j], obs[2j+1]; float32 clip + normalise
__device__ __forceinline__ void sa_observe(const SaEnv &e, float2 *row)
{
    row[0] = make_float2(norm_f32(e.acc, -5.f, 5.f), norm_f32(e.steer, -30.f, 30.f));
    row[1] = make_float2(norm_f32(e.rx, 0.f, 1200.f), norm_f32(e.py, 0.f, 8.f));
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        row[2 + k] = make_float2(norm_f32(dsub(e.orx[k], e.rx), -60.f, 60.f),
                                 norm_f32(dsub(lane_centre(k), e.py), -8.f, 8.f));
        row[6 + k] = make_float2(norm_f32(dsub(e.ovx[k], e.vx), -20.f, 20.f), 0.5f);
    }
    row[5] = make_float2(norm_f32(e.vx, -20.f, 20.f), 0.5f);
}

// shared-memory staged, fully coalesced store of a block's observations ([kSaBlock][18] floats)
__device__ __forceinline__ void store_obs_block(const float *sm, float *obs, int64_t block_base, int64_t n)
{
    const int64_t rows = min((int64_t)kSaBlock, n - block_base);
    const int nvec = (int)(rows * HW_SA_OBS_DIM / 4);  // rows*18 floats; rows*18*4 bytes, 16 B multiple iff rows even
    float4 *dst = reinterpret_cast<float4 *>(obs + block_base * HW_SA_OBS_DIM);
    const float4 *src = reinterpret_cast<const float4 *>(sm);
    for (int v = threadIdx.x; v < nvec; v += kSaBlock) dst[v] = src[v];
    const int tail0 = nvec * 4, total = (int)(rows * HW_SA_OBS_DIM);
    for (int f = tail0 + threadIdx.x; f < total; f += kSaBlock) obs[block_base * HW_SA_OBS_DIM + f] = sm[f];
}

template <bool CONT, bool RANDOM_ACTIONS>
__global__ void __launch_bounds__(kSaBlock)
sa_step_kernel(SaPtrs S, const void *__restrict__ actions, float *__restrict__ obs, float *__restrict__ rew,
               uint8_t *__restrict__ done, int4 *__restrict__ term, unsigned long long *__restrict__ stats,
               const HwSimParams P, const uint64_t seed, const uint64_t env_id0, const uint64_t action_ctr0,
               const int steps, const int64_t n, const int flags)
{
    __shared__ __align__(16) float sm_obs[kSaBlock * HW_SA_OBS_DIM];
    const int64_t block_base = (int64_t)blockIdx.x * kSaBlock;
    const int64_t i = block_base + threadIdx.x;
    const bool valid = i < n;
    const bool inf_obs = !(flags & HW_FLAG_NO_INF_OBS);
    const bool autoreset = !(flags & HW_FLAG_NO_AUTORESET);
    const double dt = P.dt;
    const double k30dt = dmul(30.0, dt);

    if (valid) {
        SaEnv e;
        sa_load(e, S, i, n);
        const uint64_t env_id = env_id0 + (uint64_t)i;

        double reward = 0.0;
        int rew_milli = 0;
        bool is_done = false;
        for (int s = 0; s < steps; ++s) {
            int act_d = 0;
            double a0_5dt = 0.0, a1_30dt = 0.0;
            if (RANDOM_ACTIONS) {
                const uint64_t aseed = seed ^ 0x9E3779B97F4A7C15ull;
                const uint64_t c = action_ctr0 + (uint64_t)s;
                const uint4 w = philox4x32_10(make_uint4((uint32_t)c, (uint32_t)env_id, (uint32_t)(env_id >> 32), kActionTag),
                                              make_uint2((uint32_t)aseed, (uint32_t)(aseed >> 32)));
                if (CONT) {
                    // U(-1,1) on the float32 grid: 24-bit mantissa uniform, exactly representable
                    const float f0 = (float)(w.x >> 8) * (1.0f / 8388608.0f) - 1.0f;
                    const float f1 = (float)(w.y >> 8) * (1.0f / 8388608.0f) - 1.0f;
                    a0_5dt = dmul(dmul((double)f0, 5.0), dt);
                    a1_30dt = dmul(dmul((double)f1, 30.0), dt);
                } else {
                    act_d = (int)(w.x >> 30);
                }
            } else if (CONT) {
                const float2 a = reinterpret_cast<const float2 *>(actions)[i];
                a0_5dt = dmul(dmul((double)a.x, 5.0), dt);   // selected_action[0] * max_acceleration * dt
                a1_30dt = dmul(dmul((double)a.y, 30.0), dt); // selected_action[1] * max_steering * dt
            } else {
                act_d = reinterpret_cast<const int *>(actions)[i];
            }

            // env.step: up to ticks_per_step ticks, stop at the first crash (highway_env.py:99-108)
            int ncol = 0;
            for (int t = 0; t < P.ticks_per_step; ++t) {
                ncol = sa_tick<CONT>(e, dt, k30dt, inf_obs, act_d, a0_5dt, a1_30dt, seed, env_id);
                if (ncol > 0) break;
            }
            const uint32_t tick = e.bits & HW_BITS_TICK_MASK;
            const bool timeout = tick >= (uint32_t)P.timeout_tick;
            is_done = timeout || ncol > 0;
            reward = (ncol > 0) ? -250.0 * (double)ncol : dsub(ddiv(e.vx, 20.0), 1.0);
            rew_milli = round_milli(reward);
            e.ep_ret_milli += rew_milli;
            e.ep_len += 1;

            if (is_done) {
                if (term) term[i] = make_int4((int)tick, (int)e.ncoll, e.ep_ret_milli, (int)e.ep_len);
                if (stats) {
                    atomicAdd(&stats[0], 1ull);
                    atomicAdd(&stats[1], (unsigned long long)e.ep_len);
                    atomicAdd(&stats[2], (unsigned long long)(long long)e.ep_ret_milli);
                    atomicAdd(&stats[3], (unsigned long long)e.ncoll);
                    if (ncol == 0) atomicAdd(&stats[4], 1ull);
                }
                if (autoreset) sa_reset_env(e);
            }
        }
        rew[i] = (float)milli_to_reward(rew_milli, reward);
        done[i] = is_done ? 1 : 0;
        sa_observe(e, reinterpret_cast<float2 *>(sm_obs) + threadIdx.x * (HW_SA_OBS_DIM / 2));
        sa_store(e, S, i, n);
    }
    __syncthreads();
    if (block_base < n) store_obs_block(sm_obs, obs, block_base, n);
}

__global__ void __launch_bounds__(kSaBlock)
sa_reset_kernel(SaPtrs S, const uint8_t *__restrict__ mask, float *__restrict__ obs, const int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * kSaBlock + threadIdx.x;
    if (i >= n) return;
    if (mask && !mask[i]) return;
    SaEnv e;
    e.spawn_ctr = S.stat[i].y;
    sa_reset_env(e);
    sa_store(e, S, i, n);
    if (obs) {
        float2 row[9];
        sa_observe(e, row);
        float2 *dst = reinterpret_cast<float2 *>(obs + i * HW_SA_OBS_DIM);
#pragma unroll
        for (int j = 0; j < 9; ++j) dst[j] = row[j];
    }
}

__global__ void __launch_bounds__(kSaBlock)
sa_observe_kernel(SaPtrs S, float *__restrict__ obs, const int64_t n)
{
    __shared__ __align__(16) float sm_obs[kSaBlock * HW_SA_OBS_DIM];
    const int64_t block_base = (int64_t)blockIdx.x * kSaBlock;
    const int64_t i = block_base + threadIdx.x;
    if (i < n) {
        SaEnv e;
        sa_load(e, S, i, n);
        sa_observe(e, reinterpret_cast<float2 *>(sm_obs) + threadIdx.x * (HW_SA_OBS_DIM / 2));
    }
    __syncthreads();
    if (block_base < n) store_obs_block(sm_obs, obs, block_base, n);
}

static inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

static int check_state(const HwSaState *st, int64_t n)
{
    if (!st || n <= 0 || !st->car_a || !st->car_b || !st->obst || !st->stat) return HW_E_BADARG;
    if (!aligned16(st->car_a) || !aligned16(st->car_b) || !aligned16(st->obst) || !aligned16(st->stat)) return HW_E_ALIGN;
    return 0;
}

static SaPtrs ptrs(const HwSaState *st)
{
    SaPtrs S;
    S.car_a = reinterpret_cast<float4 *>(st->car_a);
    S.car_b = reinterpret_cast<float4 *>(st->car_b);
    S.obst = reinterpret_cast<float4 *>(st->obst);
    S.stat = reinterpret_cast<uint4 *>(st->stat);
    return S;
}

static int launch_step(const HwSaState *st, const void *actions, const HwSaOut *out, const HwSimParams *P,
                       uint64_t seed, uint64_t env_id0, uint64_t action_ctr0, int steps, bool random_actions,
                       int64_t n, int flags, cudaStream_t stream)
{
    int rc = check_state(st, n);
    if (rc) return rc;
    if (!out || !out->obs || !out->rew || !out->done || !P) return HW_E_BADARG;
    if (!random_actions && !actions) return HW_E_BADARG;
    if (steps < 1 || P->ticks_per_step < 1 || P->timeout_tick < 1 || P->timeout_tick > 0xFFFF || !(P->dt > 0.0)) return HW_E_BADARG;
    if (!aligned16(out->obs) || (out->term && !aligned16(out->term))) return HW_E_ALIGN;
    const SaPtrs S = ptrs(st);
    const unsigned grid = (unsigned)((n + kSaBlock - 1) / kSaBlock);
    int4 *term = reinterpret_cast<int4 *>(out->term);
    const bool cont = (flags & HW_FLAG_CONTINUOUS) != 0;
#define HW_LAUNCH(C, R) sa_step_kernel<C, R><<<grid, kSaBlock, 0, stream>>>(S, actions, out->obs, out->rew, out->done, term, out->stats, *P, seed, env_id0, action_ctr0, steps, n, flags)
    if (cont) { if (random_actions) HW_LAUNCH(true, true); else HW_LAUNCH(true, false); }
    else      { if (random_actions) HW_LAUNCH(false, true); else HW_LAUNCH(false, false); }
#undef HW_LAUNCH
    return (int)cudaGetLastError();
}

}  // namespace hw

extern "C" {

int hw_abi_version(void) { return HW_ABI_VERSION; }

const char *hw_error_string(int code)
{
    if (code == 0) return "success";
    if (code == HW_E_BADARG) return "highway_b200: bad argument (null pointer, non-positive size or unsupported combination)";
    if (code == HW_E_ALIGN) return "highway_b200: pointer must be 16-byte aligned";
    if (code > 0) return cudaGetErrorString((cudaError_t)code);
    return "highway_b200: unknown error";
}

int hw_sa_reset(const HwSaState *state, const uint8_t *mask, float *obs, int64_t n, void *stream)
{
    int rc = hw::check_state(state, n);
    if (rc) return rc;
    const unsigned grid = (unsigned)((n + hw::kSaBlock - 1) / hw::kSaBlock);
    hw::sa_reset_kernel<<<grid, hw::kSaBlock, 0, (cudaStream_t)stream>>>(hw::ptrs(state), mask, obs, n);
    return (int)cudaGetLastError();
}

int hw_sa_step(const HwSaState *state, const void *actions, const HwSaOut *out, const HwSimParams *params,
               uint64_t seed, uint64_t env_id0, int64_t n, int flags, void *stream)
{
    return hw::launch_step(state, actions, out, params, seed, env_id0, 0, 1, false, n, flags, (cudaStream_t)stream);
}

int hw_sa_rollout_random(const HwSaState *state, const HwSaOut *out, const HwSimParams *params,
                         uint64_t seed, uint64_t env_id0, uint64_t action_ctr0, int steps,
                         int64_t n, int flags, void *stream)
{
    return hw::launch_step(state, nullptr, out, params, seed, env_id0, action_ctr0, steps, true, n, flags, (cudaStream_t)stream);
}

int hw_sa_observe(const HwSaState *state, float *obs, int64_t n, void *stream)
{
    int rc = hw::check_state(state, n);
    if (rc) return rc;
    if (!obs) return HW_E_BADARG;
    if (!hw::aligned16(obs)) return HW_E_ALIGN;
    const unsigned grid = (unsigned)((n + hw::kSaBlock - 1) / hw::kSaBlock);
    hw::sa_observe_kernel<<<grid, hw::kSaBlock, 0, (cudaStream_t)stream>>>(hw::ptrs(state), obs, n);
    return (int)cudaGetLastError();
}

int hw_sa_step_host(const HwSaState *state, const void *h_actions, void *d_actions, const HwSaOut *d_out,
                    float *h_obs, float *h_rew, uint8_t *h_done, const HwSimParams *params,
                    uint64_t seed, uint64_t env_id0, int64_t n, int flags, void *stream)
{
    if (!h_actions || !d_actions || !d_out || !h_obs || !h_rew || !h_done || n <= 0) return HW_E_BADARG;
    cudaStream_t s = (cudaStream_t)stream;
    const size_t abytes = (size_t)n * ((flags & HW_FLAG_CONTINUOUS) ? 2 * sizeof(float) : sizeof(int32_t));
    cudaError_t ce = cudaMemcpyAsync(d_actions, h_actions, abytes, cudaMemcpyHostToDevice, s);
    if (ce != cudaSuccess) return (int)ce;
    int rc = hw::launch_step(state, d_actions, d_out, params, seed, env_id0, 0, 1, false, n, flags, s);
    if (rc) return rc;
    ce = cudaMemcpyAsync(h_obs, d_out->obs, (size_t)n * HW_SA_OBS_DIM * sizeof(float), cudaMemcpyDeviceToHost, s);
    if (ce != cudaSuccess) return (int)ce;
    ce = cudaMemcpyAsync(h_rew, d_out->rew, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, s);
    if (ce != cudaSuccess) return (int)ce;
    ce = cudaMemcpyAsync(h_done, d_out->done, (size_t)n, cudaMemcpyDeviceToHost, s);
    if (ce != cudaSuccess) return (int)ce;
    return (int)cudaStreamSynchronize(s);
}

}  // extern "C"
