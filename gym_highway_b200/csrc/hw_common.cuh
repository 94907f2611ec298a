// Shared device helpers for the highway step kernels (sm_100a).
//
// Arithmetic contract: the reference computes every tick in Python floats (fp64), one rounding per
// binary operation.  The kernels therefore promote the fp32 state to fp64 registers, use the
// explicitly rounded intrinsics (__dadd_rn/__dmul_rn/__ddiv_rn are never contracted into FMAs) and
// round once when the state is stored back.  The file is additionally compiled with -fmad=false.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace hw {

constexpr double kPi = 3.14159265358979323846;
constexpr double kDeg2Rad = kPi / 180.0;  // math.radians(x) = x * (pi/180)
constexpr double kRad2Deg = 180.0 / kPi;  // math.degrees(x) = x * (180/pi)

constexpr uint32_t kBitLeft = 1u << 18, kBitRight = 1u << 19, kBitDoAcc = 1u << 20, kBitDoMaint = 1u << 21;

__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double ddiv(double a, double b) { return __ddiv_rn(a, b); }

// x / c for a literal c, correctly rounded, in three instructions: q0 = x*rc, q = fma(fma(-q0, c, x), rc, q0) with
// rc = RN(1/c) folded at compile time (Markstein: rc correctly rounded and q0 faithful give the IEEE quotient; checked
// against the division for c = 20 on 2e9 velocities and for c = 1000 on every integer numerator the reward can
// produce, scripts/verify_div_by_const_f64.c).  x must not be -0.0 (the quotient would come out as +0.0).
__device__ __forceinline__ double ddiv_const(double x, const double c)
{
    const double rc = 1.0 / c;
    const double q0 = dmul(x, rc);
    return __fma_rn(__fma_rn(-q0, c, x), rc, q0);
}

// Conditional moves of doubles, spelled in PTX as a branch around one instruction so that ptxas predicates that
// instruction instead of emitting the FSEL pair a C++ ternary gives.  Why: the step kernels are bound by
// instruction issue, and within that by the 16-lane ALU pipe that executes FSEL / ISETP / LOP3 (it was ~70 % busy
// against 36 % for the fp64 pipe).
//   dmov_if: if (p) x = y for y in a register: ONE predicated DMUL by 1.0 (exact for every y, -0.0 included).
//   dset_if: if (p) x = c for a literal c: predicated moves (a single CS2R / DMUL for zero).
__device__ __forceinline__ void dmov_if(bool p, double &x, double y)
{
    asm("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %1, 0;\n\t@!q bra HW_SKIP;\n\tmul.rn.f64 %0, %2, 0d3FF0000000000000;\nHW_SKIP:\n\t}" : "+d"(x) : "r"((int)p), "d"(y));
}

// same with the condition given as a word that is non-zero when true (e.g. bits & HW_BIT_x): the test is then one LOP3
__device__ __forceinline__ void dmov_if_nz(uint32_t w, double &x, double y)
{
    asm("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %1, 0;\n\t@!q bra HW_SKIP;\n\tmul.rn.f64 %0, %2, 0d3FF0000000000000;\nHW_SKIP:\n\t}" : "+d"(x) : "r"(w), "d"(y));
}


// a copy that costs ONE instruction (DMUL by 1.0) instead of the two 32-bit moves of a plain assignment; for values that
// seed a chain of dmov_if's
__device__ __forceinline__ double dcopy(double y)
{
    double r;
    asm("mul.rn.f64 %0, %1, 0d3FF0000000000000;" : "=d"(r) : "d"(y));
    return r;
}

// a, b or c for sel = 1, 2, anything else: three DMULs, the last two predicated
__device__ __forceinline__ double dsel3(int sel, double a, double b, double c)
{
    double r;
    asm("{\n\t.reg .pred p1, p2;\n\tsetp.eq.s32 p1, %4, 1;\n\tsetp.eq.s32 p2, %4, 2;\n\t"
        "mul.rn.f64 %0, %3, 0d3FF0000000000000;\n\t"
        "@!p2 bra HW_S2;\n\tmul.rn.f64 %0, %2, 0d3FF0000000000000;\nHW_S2:\n\t"
        "@!p1 bra HW_S1;\n\tmul.rn.f64 %0, %1, 0d3FF0000000000000;\nHW_S1:\n\t}"
        : "=&d"(r) : "d"(a), "d"(b), "d"(c), "r"(sel));
    return r;
}

// two three-way selects on the same selector (shares the two compares)
__device__ __forceinline__ void dsel3x2(int sel, double a0, double a1, double a2, double b0, double b1, double b2, double &ra, double &rb)
{
    asm("{\n\t.reg .pred p1, p2;\n\tsetp.eq.s32 p1, %8, 1;\n\tsetp.eq.s32 p2, %8, 2;\n\t"
        "mul.rn.f64 %0, %4, 0d3FF0000000000000;\n\tmul.rn.f64 %1, %7, 0d3FF0000000000000;\n\t"
        "@!p2 bra HW_T2;\n\tmul.rn.f64 %0, %3, 0d3FF0000000000000;\n\tmul.rn.f64 %1, %6, 0d3FF0000000000000;\nHW_T2:\n\t"
        "@!p1 bra HW_T1;\n\tmul.rn.f64 %0, %2, 0d3FF0000000000000;\n\tmul.rn.f64 %1, %5, 0d3FF0000000000000;\nHW_T1:\n\t}"
        : "=&d"(ra), "=&d"(rb) : "d"(a0), "d"(a1), "d"(a2), "d"(b0), "d"(b1), "d"(b2), "r"(sel));
}

__device__ __forceinline__ void dset_if(bool p, double &x, double c)
{
    asm("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %1, 0;\n\t@!q bra HW_SKIP;\n\tmov.f64 %0, %2;\nHW_SKIP:\n\t}" : "+d"(x) : "r"((int)p), "d"(c));
}

// python max(lo, min(x, hi)): min keeps x unless hi < x; max keeps lo unless the inner value is > lo.  (Spelled with
// conditional moves: NVVM turns the C++ ternary `(x > lo) ? x : lo` into max.f64, whose NaN-aware expansion on sm_100a is
// six instructions.)
// Both tests read the ORIGINAL x (x > hi implies x > lo, so the second test cannot fire after the first moved x to hi; a NaN fails
// the first and passes the second either way): the two DSETP issue back to back instead of the second one waiting for the first
// conditional move - a DSETP delivers its predicate after ~20 cycles on B200 (scripts/micro/fp64_latency.cu).
__device__ __forceinline__ double py_clamp(double x, const double lo, const double hi)
{
    const bool over = hi < x, under = !(x > lo);
    dset_if(over, x, hi);
    dset_if(under, x, lo);
    return x;
}

__device__ __forceinline__ double lane_centre(int k)  // k = lane-1: 1.25, 3.75, 6.25
{
    return 1.25 + 2.5 * (double)k;  // exact in binary
}

// pygame Rect stores ints; assigning a float truncates toward zero
// (x * 32 is exact - a power of two - so the product and the subtraction round once, and one FMA returns the same bits)
__device__ __forceinline__ int rect_x(double px) { return __double2int_rz(__fma_rn(px, 32.0, -38.0)); }
__device__ __forceinline__ int rect_y(double py) { return __double2int_rz(__fma_rn(py, 32.0, -19.0)); }

__device__ __forceinline__ bool rects_overlap(int ax, int ay, int bx, int by)
{
    return ax < bx + 76 && ay < by + 38 && ax + 76 > bx && ay + 38 > by;
}

// ---- Philox4x32-10 (Salmon et al., SC'11) -------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u;
        k.y += 0xBB67AE85u;
    }
    return c;
}

// CPython random(): (a>>5, b>>6) -> (a*2^26 + b) / 2^53
__device__ __forceinline__ double u53(uint32_t a, uint32_t b)
{
    return dmul(dadd(dmul((double)(a >> 5), 67108864.0), (double)(b >> 6)), 1.0 / 9007199254740992.0);
}

constexpr uint32_t kSpawnTag = 0x48574159u;   // "HWAY": obstacle respawn stream
constexpr uint32_t kActionTag = 0x41435421u;  // "ACT!": in-kernel random-action stream

// out of line on purpose: a respawn is rare, and the ten Philox rounds would otherwise sit in the tick loop
#ifndef HW_SPAWN_ATTR
#define HW_SPAWN_ATTR __noinline__
#endif
static __device__ HW_SPAWN_ATTR double2 spawn_uniform_pair(uint64_t seed, uint64_t env_id, uint32_t spawn_ctr)
{
    const uint4 w = philox4x32_10(make_uint4(spawn_ctr, (uint32_t)env_id, (uint32_t)(env_id >> 32), kSpawnTag),
                                  make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    return make_double2(u53(w.x, w.y), u53(w.z, w.w));
}

__device__ __forceinline__ void spawn_uniforms(uint64_t seed, uint64_t env_id, uint32_t spawn_ctr, double &u_pos, double &u_vel)
{
    const double2 u = spawn_uniform_pair(seed, env_id, spawn_ctr);
    u_pos = u.x;
    u_vel = u.y;
}

// python round(x, 3) as an integer count of 1/1000: nearest integer to the EXACT product x*1000,
// ties (on the exact value) to even.  p + e == x*1000 exactly (e from the FMA residual).
__device__ __forceinline__ int round_milli(double x)
{
    const double p = dmul(x, 1000.0);
    const double e = __fma_rn(x, 1000.0, -p);
    double k = rint(p);
    if (fabs(dsub(p, k)) == 0.5 && e != 0.0) k = (e > 0.0) ? ceil(p) : floor(p);
    return (int)k;
}

// value python would return for round(x, 3), given k = round_milli(x)
__device__ __forceinline__ double milli_to_reward(int k, double x)
{
    return (k == 0) ? copysign(0.0, x) : ddiv_const((double)k, 1000.0);
}

// np.clip(ob, low, high) then (ob - low) / (high - low), all in float32 (gym Box stores float32 bounds).
// The span is a compile-time constant.  Power-of-two spans divide exactly by multiplication; for the
// others (10, 40, 60, 120, 1200) the quotient is computed as q0 = x*rc, q = fma(fma(-q0, c, x), rc, q0)
// with rc = RN(1/c), which equals the IEEE-rounded x/c for x == 0 and for every float x in [1e-30, c]
// (verified exhaustively on the host for each constant, scripts/verify_div_by_const.c); x = v - lo
// cannot fall into (0, 1e-30) unless lo == 0, where GUARD_TINY keeps the IEEE division for that range.
// CLIP = false: the caller guarantees lo <= raw <= hi (np.clip is then the identity and its two compare-selects are dropped).
template <bool GUARD_TINY = false, bool CLIP = true>
__device__ __forceinline__ float norm_f32(double raw, const float lo, const float hi)
{
    float v = (float)raw;  // numpy.array(..., dtype=float32): round to nearest
    if (CLIP) {
        v = (v < lo) ? lo : v;
        v = (v > hi) ? hi : v;
    }
    const float x = __fsub_rn(v, lo);
    const float c = hi - lo;  // exact for every bound pair of the observation space
    if (c == 8.f || c == 16.f) return __fmul_rn(x, 1.0f / c);
    const float rc = 1.0f / c;  // folded at compile time: RN(1/c)
    if (GUARD_TINY && x < 1e-30f && x != 0.f) return __fdiv_rn(x, c);
    const float q0 = __fmul_rn(x, rc);
    return __fmaf_rn(__fmaf_rn(-q0, c, x), rc, q0);
}

// tan(x) = x + x*z*p(z), z = x*x, for |x| <= 0.55 rad (python clamps |steering| to 30 deg = 0.5236 rad).
// p = near-minimax fit of (tan x - x)/x^3 (mpmath chebyfit, 12 terms, |error| < 4e-19 relative to tan),
// Horner in FMAs: no range reduction, no special cases; within 1 ulp of libm's tan.
static __constant__ double kTanC[12] = {
    0x1.20f68de6f5b19p-15, 0x1.790a294731d86p-16, 0x1.b8cea70623061p-14, 0x1.f054fa543c93bp-13,
    0x1.3598d277c0b71p-11, 0x1.7d9f32bbba0bcp-10, 0x1.d6d3fe7809e9dp-9,  0x1.226e34c1e860cp-7,
    0x1.664f48853966bp-6,  0x1.ba1ba1ba167cdp-5,  0x1.1111111111133p-3,  0x1.5555555555555p-2};

__device__ __forceinline__ double tan_small(double x)
{
    const double z = x * x;
    double p = kTanC[0];
#pragma unroll
    for (int i = 1; i < 12; ++i) p = __fma_rn(p, z, kTanC[i]);
    return __fma_rn(x * z, p, x);
}

// angular_velocity = velocity.x / (length / tan(radians(steering))), length = 4 (multi_lane_sim.py:204-208).
//
// The reference rounds three times (tan, 4/t, vx/R).  libm's tan cannot be reproduced bit for bit on the
// GPU anyway (tan_small is within 1 ulp of it), so the default (EXACT = false) evaluates the mathematically
// identical vx * tan / 4 with ONE rounding (the scaling by 0.25 is exact): the result is within 3 ulp(fp64)
// of the reference's value instead of within 1-2, and the two fp64 divisions (~30 instructions per tick)
// are gone (+15 % discrete, +40 % continuous throughput on B200).  Measured effect: py differs from the
// oracle by one fp64 ulp on a fraction of the steering ticks, which reaches a stored or observed fp32
// value about once per 10^7 values; discrete decisions (lane arrival, rect truncation) are affected with
// probability ~1e-14 per tick.  EXACT = true (step flag HW_FLAG_EXACT_STEERING) keeps the two IEEE
// divisions (4/t as 4*(1/t), exact because 4 is a power of two) and is what the bit-exact trajectory tests run.
template <bool EXACT>
__device__ __forceinline__ double angular_velocity_t(double vx, double t /* tan(radians(steer)) */)
{
    if (EXACT) return ddiv(vx, dmul(4.0, __drcp_rn(t)));
    return dmul(dmul(vx, t), 0.25);
}

template <bool EXACT>
__device__ __forceinline__ double angular_velocity(double vx, double steer_deg)
{
    return angular_velocity_t<EXACT>(vx, tan_small(dmul(steer_deg, kDeg2Rad)));
}

// Sum over the 32 lanes of a fully converged warp.  Episode counters are accumulated per thread and
// published with ONE atomic per warp and counter: when a whole batch times out on the same step (every env that
// never crashed started together), per-thread atomics on eight addresses serialise into milliseconds.
__device__ __forceinline__ long long warp_sum(long long v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace hw
