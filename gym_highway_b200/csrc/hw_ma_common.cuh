// Declarations shared by the multi-agent kernels (entry points, reset / observe: hw_ma.cu; thread per environment: hw_ma_tpe.cu; lanes per
// environment: hw_ma_grp.cu).
#pragma once
#include "hw_common.cuh"
#include "../../include/highway_b200.h"

namespace hw {

constexpr int kMaBlock = 64;
constexpr int kMaxA = HW_MA_MAX_AGENTS;
constexpr int kMaxAExt = HW_MA_MAX_AGENTS_EXT;

// start grid (simple_base.py:32-38): the "X" formation (10, L1), (10, L3), (7.5, L2), (5, L1), (5, L3); cars 5..7 are the
// extension beyond the reference's five slots: (2.5, L2), (0, L3), (-2.5, L1) - staggered backwards, and keeping lane 1 clear of
// x in (-1.2, 3.6): the three start obstacles respawn on the first tick, and a car whose rect covers the pixel origin is hit by
// their default rects (0, 0, 76, 38) (highway_core.py:314-316 then :407)
__host__ __device__ __forceinline__ double ma_start_x(int a) { return (a < 2) ? 10.0 : (a == 2) ? 7.5 : (a < 5) ? 5.0 : 7.5 - 2.5 * (a - 3); }
__host__ __device__ __forceinline__ int ma_start_lane(int a) { return (a == 2 || a == 5) ? 2 : (a == 0 || a == 3 || a == 7) ? 1 : 3; }
constexpr int kMaxK = HW_MA_MAX_SLOTS;

// car fields in shared memory
enum { C_PX = 0, C_PY, C_RX, C_VX, C_VY, C_ACC, C_STEER, C_CRUISE, C_NF };
enum { O_PX = 0, O_RX, O_VX, O_IV, O_NF };

struct MaPtrs {
    float *cars;
    float4 *obst;
    uint32_t *head;
    int A, K;
};

struct MaConsts {
    double dt, k30dt;
    int ticks_per_step, timeout_tick;
};

struct MaHead {
    uint32_t tick, leader, nobs;
    uint32_t newest[3];  // slot index, 31 = retired
    uint32_t olane;      // 2 bits per slot
    uint32_t ofresh;     // bit k: slot k still carries pygame's default rect (0,0,76,38)
    uint32_t flags;      // bit 0: a spawn was dropped for lack of slots, bit 1: a lane had no obstacle to observe;
                         // in memory bits 8..12 hold the per-car done flags that persist until reset (no auto-reset only)
    uint32_t nc_obs, nc_agent, spawn_ctr, ep_len;
    double ep_ret;
};

__device__ __forceinline__ void ma_load_head(MaHead &h, const MaPtrs &S, int64_t i, int64_t n)
{
    const uint32_t w0 = S.head[i], w1 = S.head[n + i];
    h.tick = w0 & 0xFFFFu; h.leader = (w0 >> 16) & 7u; h.nobs = (w0 >> 20) & 31u;
    h.newest[0] = w1 & 31u; h.newest[1] = (w1 >> 5) & 31u; h.newest[2] = (w1 >> 10) & 31u;
    h.olane = S.head[2 * n + i];
    const uint32_t w3 = S.head[3 * n + i];
    h.ofresh = w3 & 0xFFFFu; h.flags = w3 >> 16;
    h.nc_obs = S.head[4 * n + i]; h.nc_agent = S.head[5 * n + i]; h.spawn_ctr = S.head[6 * n + i]; h.ep_len = S.head[7 * n + i];
    h.ep_ret = __hiloint2double((int)S.head[9 * n + i], (int)S.head[8 * n + i]);
}

__device__ __forceinline__ void ma_store_head(const MaHead &h, const MaPtrs &S, int64_t i, int64_t n)
{
    S.head[i] = h.tick | (h.leader << 16) | (h.nobs << 20);
    S.head[n + i] = h.newest[0] | (h.newest[1] << 5) | (h.newest[2] << 10);
    S.head[2 * n + i] = h.olane;
    S.head[3 * n + i] = h.ofresh | (h.flags << 16);
    S.head[4 * n + i] = h.nc_obs; S.head[5 * n + i] = h.nc_agent; S.head[6 * n + i] = h.spawn_ctr; S.head[7 * n + i] = h.ep_len;
    S.head[8 * n + i] = (uint32_t)__double2loint(h.ep_ret); S.head[9 * n + i] = (uint32_t)__double2hiint(h.ep_ret);
}

// 2-bit lane fields of the live slots; nobs == 16 would be a shift by 32 (undefined in C++)
__device__ __forceinline__ uint32_t live_mask(uint32_t nobs) { return nobs >= 16u ? 0xFFFFFFFFu : (1u << (2 * nobs)) - 1u; }

__device__ __forceinline__ int ob_lane(const MaHead &h, int k) { return (int)((h.olane >> (2 * k)) & 3u); }

// (v - lo) / (hi - lo) in float64, cast to float32 (MultiAgentEnv._get_obs_norm normalises in float64; the
// VecEnv buffer is float32).  The division is q = fma(fma(-q0, c, x), rc, q0), q0 = x*rc, rc = RN(1/c):
// Markstein's correction, which returns the correctly rounded quotient (checked against IEEE division on 10^9
// random operands per constant, scripts/verify_div_by_const.c).
__device__ __forceinline__ float norm_f64(double raw, const double lo, const double hi)
{
    // np.array(..., dtype=float32), then np.clip against bounds that are all exact in float32 (5, 30, 1200, 8, 60, 20): the
    // clip of the float32 value is the same number in either precision, and two FMNMX are a third of the fp64 compare-selects
    const float v32 = fminf(fmaxf((float)raw, (float)lo), (float)hi);
    const double v = (double)v32;
    const double x = dsub(v, lo), c = hi - lo, rc = 1.0 / c;
    const double q0 = dmul(x, rc);
    return (float)__fma_rn(__fma_rn(-q0, c, x), rc, q0);
}

// N normalisations (norm_f64: float32 cast, clip, (v - lo) / (hi - lo) in float64, float32 cast) evaluated stage by stage.  One value
// is a chain of nine dependent instructions (~95 cycles with the three conversions); written one value after the other ptxas
// emitted the 170 chains of a step back to back (a third of the kernel's stall samples sat in the epilogue).  Side by side the
// N chains of a batch overlap.  BND: 0 = +-5, 1 = +-30, 2 = 0..1200, 3 = 0..8, 4 = +-60, 5 = +-8, 6 = +-20.
__device__ __forceinline__ constexpr double ma_bnd_lo(int b) { return b == 0 ? -5.0 : b == 1 ? -30.0 : b == 2 ? 0.0 : b == 3 ? 0.0 : b == 4 ? -60.0 : b == 5 ? -8.0 : -20.0; }
__device__ __forceinline__ constexpr double ma_bnd_hi(int b) { return b == 0 ? 5.0 : b == 1 ? 30.0 : b == 2 ? 1200.0 : b == 3 ? 8.0 : b == 4 ? 60.0 : b == 5 ? 8.0 : 20.0; }
template <int N, int B0, int B1 = 0, int B2 = 0, int B3 = 0, int B4 = 0, int B5 = 0>
__device__ __forceinline__ void ma_norm_n(const double (&raw)[N], float (&out)[N])
{
    constexpr int B[6] = {B0, B1, B2, B3, B4, B5};
    static_assert(N <= 6, "batch size");
    float v[N];
    double x[N], q0[N], e[N];
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = (float)raw[i];
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = fminf(fmaxf(v[i], (float)ma_bnd_lo(B[i])), (float)ma_bnd_hi(B[i]));
#pragma unroll
    for (int i = 0; i < N; ++i) x[i] = dsub((double)v[i], ma_bnd_lo(B[i]));
#pragma unroll
    for (int i = 0; i < N; ++i) q0[i] = dmul(x[i], 1.0 / (ma_bnd_hi(B[i]) - ma_bnd_lo(B[i])));
    // spans of 8 and 16 (the lateral quantities) divide exactly by the multiplication above: no correction step
#pragma unroll
    for (int i = 0; i < N; ++i) e[i] = (B[i] == 3 || B[i] == 5) ? 0.0 : __fma_rn(-q0[i], ma_bnd_hi(B[i]) - ma_bnd_lo(B[i]), x[i]);
#pragma unroll
    for (int i = 0; i < N; ++i)
        out[i] = (float)((B[i] == 3 || B[i] == 5) ? q0[i] : __fma_rn(e[i], 1.0 / (ma_bnd_hi(B[i]) - ma_bnd_lo(B[i])), q0[i]));
}


// G lanes per environment (hw_ma_grp.cu); arguments already validated by the caller
int ma_grp_launch_step(const HwMaState *st, const void *actions, const HwMaOut *out, const MaConsts &C,
                       uint64_t seed, uint64_t env_id0, int64_t n, int flags, cudaStream_t stream);

// one thread per environment, cars in registers (hw_ma_tpe.cu)
int ma_tpe_launch_step(const HwMaState *st, const void *actions, const HwMaOut *out, const MaConsts &C,
                       uint64_t seed, uint64_t env_id0, int64_t n, int flags, cudaStream_t stream);

}  // namespace hw
