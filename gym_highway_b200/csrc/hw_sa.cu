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

#ifndef HW_SA_BLOCK
#define HW_SA_BLOCK 64   // two warps per block: warps never synchronise with each other, small blocks only shorten the tail
#endif
constexpr int kSaBlock = HW_SA_BLOCK;

// Per-environment working set (registers).  ovx[] (each obstacle's velocity after its last update) is
// only materialised at the step boundaries: inside a step it is init_vx, or the car's vx for the one
// obstacle that is being held back (`slow`), so the tick loop carries a flag instead of three doubles.
struct SaEnv {
    double px, py, rx, vx, acc, steer, cruise;
    double opx[3], orx[3], ovx[3], oiv[3];
    uint32_t bits;  // tick | lane << 16 | mode flags
    uint32_t spawn_ctr;
    bool slow;      // the obstacle in the car's lane moved with min(init_vx, car vx) == car vx in the last tick
};

struct SaPtrs {
    float4 *car_a, *car_b;
    float4 *obst0, *obst1, *obst2;  // the three lanes of `obst` (batch-size apart; a launch may cover a sub-range)
    uint4 *stat;
};

// environments are indexed with 32 bits (a launch covers < 2^31 of them): one IMAD.WIDE per address instead of a 64-bit add chain
__device__ __forceinline__ void sa_load(SaEnv &e, const SaPtrs &S, uint32_t i)
{
    const float4 a = S.car_a[i], b = S.car_b[i];
    const float4 o0 = S.obst0[i], o1 = S.obst1[i], o2 = S.obst2[i];
    e.px = a.x; e.py = a.y; e.rx = a.z; e.vx = a.w;
    e.acc = b.x; e.steer = b.y; e.cruise = b.z; e.bits = __float_as_uint(b.w);
    e.opx[0] = o0.x; e.orx[0] = o0.y; e.ovx[0] = o0.z; e.oiv[0] = o0.w;
    e.opx[1] = o1.x; e.orx[1] = o1.y; e.ovx[1] = o1.z; e.oiv[1] = o1.w;
    e.opx[2] = o2.x; e.orx[2] = o2.y; e.ovx[2] = o2.z; e.oiv[2] = o2.w;
    e.spawn_ctr = reinterpret_cast<const uint32_t *>(S.stat + i)[1];  // the other counters are touched once, at the end
    e.slow = false;
}

__device__ __forceinline__ void sa_store(const SaEnv &e, const SaPtrs &S, uint32_t i)
{
    S.car_a[i] = make_float4((float)e.px, (float)e.py, (float)e.rx, (float)e.vx);
    S.car_b[i] = make_float4((float)e.acc, (float)e.steer, (float)e.cruise, __uint_as_float(e.bits));
    S.obst0[i] = make_float4((float)e.opx[0], (float)e.orx[0], (float)e.ovx[0], (float)e.oiv[0]);
    S.obst1[i] = make_float4((float)e.opx[1], (float)e.orx[1], (float)e.ovx[1], (float)e.oiv[1]);
    S.obst2[i] = make_float4((float)e.opx[2], (float)e.orx[2], (float)e.ovx[2], (float)e.oiv[2]);
}

// start grid, highway_env.py:55-65
__device__ __forceinline__ void sa_reset_env(SaEnv &e)
{
    e.px = 20.0; e.py = 3.75; e.rx = 20.0; e.vx = 0.0;
    e.acc = 0.0; e.steer = 0.0; e.cruise = 0.0;
    e.bits = 2u << HW_BITS_LANE_SHIFT;
    e.opx[0] = -20.0; e.orx[0] = -20.0; e.ovx[0] = 7.0; e.oiv[0] = 7.0;
    e.opx[1] = -25.0; e.orx[1] = -25.0; e.ovx[1] = 5.0; e.oiv[1] = 5.0;
    e.opx[2] = -40.0; e.orx[2] = -40.0; e.ovx[2] = 6.0; e.oiv[2] = 6.0;
    e.slow = false;
}

// exact slow path of the collision test (generic car rect); only reached when an obstacle is within
// about one car length in x, or for states the reference cannot produce (car px != 10 after the first tick)
__device__ __noinline__ int sa_collide_exact(double px, double py, double o0, double o1, double o2)
{
    const int ax = rect_x(px), ay = rect_y(py);
    return (rects_overlap(ax, ay, rect_x(o0), 21) ? 1 : 0) + (rects_overlap(ax, ay, rect_x(o1), 101) ? 1 : 0)
         + (rects_overlap(ax, ay, rect_x(o2), 181) ? 1 : 0);
}

// road clamp of Car.update_d (:219-222, lower bound 0) / update_c (:169-172, lower bound 1)
template <bool CONT>
__device__ __forceinline__ double road_clamp(double y)
{
    // `if y < lo: y = lo  elif y > 6: y = min(y, 7)`: for lo <= y <= 6 the min is the identity.  Both tests read the value before
    // either move (they exclude each other), so the two DSETP issue back to back (a DSETP delivers its predicate after ~20 cycles)
    const bool over = 7.0 < y, under = y < (CONT ? 1.0 : 0.0);
    dset_if(over, y, 7.0);
    dset_if(under, y, CONT ? 1.0 : 0.0);
    return y;
}

// hi word of a double; for non-negative values the unsigned order of hi words is the numeric order
__device__ __forceinline__ uint32_t hi32(double x) { return (uint32_t)__double2hiint(x); }

// a, b or c for lane 1, 2, 3: plain 32-bit selects on the two halves (keeps ptxas from branching)
__device__ __forceinline__ double sel_lane(const bool l1, const bool l2, const double a, const double b, const double c)
{
    const int hi = l1 ? __double2hiint(a) : (l2 ? __double2hiint(b) : __double2hiint(c));
    const int lo = l1 ? __double2loint(a) : (l2 ? __double2loint(b) : __double2loint(c));
    return __hiloint2double(hi, lo);
}

// per-launch constants, computed on the host with the same IEEE operations python performs
struct SaConsts {
    double dt;       // 1/ticks
    double k30dt;    // 30.0 * dt   (moveLeft / moveRight steering increment)
    int ticks_per_step;
    int timeout_tick;
    double five_dt;  // 5.0 + dt: accelerate() on a clamped acceleration
    uint32_t respawn_hi;  // hi word of the respawn threshold -2.375, or 0xFFFFFFFF (never) with inf_obs=False: the flag costs nothing per tick
    double deg2rad, rad2deg_q;  // pi/180 and 0.25 * 180/pi as launch constants: one LDCU.64 each instead of a UMOV pair per use
};

// the car's lateral displacement of one tick: degrees(angular_velocity) * dt * dt (multi_lane_sim.py:204-211).
// Fast mode: angular_velocity = RN(vx*t) * 0.25 and the scaling by 0.25 commutes with the next rounding, so it is folded
// into the degrees() constant (one DMUL less, same bits).
template <bool EXACT>
__device__ __forceinline__ double lateral_step_d(double vx, double t, double dt, const double rad2deg_q)
{
    if (EXACT) return dmul(dmul(dmul(angular_velocity_t<true>(vx, t), kRad2Deg), dt), dt);
    return dmul(dmul(dmul(dmul(vx, t), rad2deg_q), dt), dt);
}

// respawn (multi_lane_sim.py:963-991): opx < -2.375, pre-filtered on the hi words (negative doubles order reversed)
__device__ __forceinline__ void sa_respawn(SaEnv &e, const uint32_t kNeg, const bool l1, const bool l2,
                                           const uint64_t seed, const uint64_t env_id)
{
    // kNeg = 0xC0030000 = hi(-2.375); hi >= kNeg  <=>  opx <= -2.375 (- a few ulps of slack)
    if (max(max(hi32(e.opx[0]), hi32(e.opx[1])), hi32(e.opx[2])) >= kNeg) {
        // lanes in ascending order; x = U(70,100) then v = U(5,7) from one Philox block per spawn
#pragma unroll 1
        for (int k = 0; k < 3; ++k) {
            const double o = (k == 0) ? e.opx[0] : (k == 1) ? e.opx[1] : e.opx[2];
            if (o < -2.375) {
                double u1, u2;
                spawn_uniforms(seed, env_id, e.spawn_ctr, u1, u2);
                e.spawn_ctr += 1;
                const double x = dadd(70.0, dmul(30.0, u1));
                const double v = dadd(5.0, dmul(2.0, u2));
                const double r = dadd(e.rx, dsub(x, 10.0));  // leader.raw_x + (x - leader.x), the leader sits at x = 10
                if (k == 0) { e.opx[0] = x; e.orx[0] = r; e.oiv[0] = v; }
                else if (k == 1) { e.opx[1] = x; e.orx[1] = r; e.oiv[1] = v; }
                else { e.opx[2] = x; e.orx[2] = r; e.oiv[2] = v; }
            }
        }
        // a fresh obstacle is far ahead (x >= 70): it cannot be the held-back one
        e.slow = e.slow && sel_lane(l1, l2, e.opx[0], e.opx[1], e.opx[2]) < 10.0;
    }
}

// One simulator tick (multi_lane_sim.py:879-1049).  Returns the number of obstacle rects the car
// overlapped (0 when the collision lock is on); the tick reward is -250*n, or vx/20-1 when n == 0.
//
// Written for instruction count: with i.i.d. random actions practically every warp holds lanes in every
// mode, so divergent branches cost the sum of both sides plus BSSY/BSYNC; the modes are therefore
// evaluated with selects, and only genuinely rare events (an obstacle within a car length, a respawn) or
// long ones (the steering path with its tan and two divisions) stay behind branches.  FIRST = first tick
// of the step, the only one that can see the collision lock or a car that is not yet pinned at x = 10.
template <bool CONT, bool FIRST, bool EXACT>
__device__ __forceinline__ int sa_tick(SaEnv &e, uint32_t &ncoll_acc, const SaConsts &K, const bool inf_obs,
                                       const uint32_t amask, const uint32_t apair, const double a0_5dt, const double a1_30dt,
                                       const uint64_t seed, const uint64_t env_id, const double dt)
{
    const bool lock = FIRST && (e.bits & HW_BITS_TICK_MASK) == 0;  // collision_count_lock: first tick after reset
    const double px = FIRST ? e.px : 10.0;
    uint32_t b = e.bits + 1;  // tick field lives in the low 16 bits; 2160 (or 3601) never carries out
    int lane = (int)((b >> HW_BITS_LANE_SHIFT) & 3u);

    // ---- apply action (multi_lane_sim.py:502-526) ----
    uint32_t fire = 0;
    if (CONT) {
        e.acc = py_clamp(dadd(e.acc, a0_5dt), -5.0, 5.0);
        e.steer = py_clamp(dadd(e.steer, a1_30dt), -30.0, 30.0);
    } else {
        // amask = the mode bit the action wants (LEFT/RIGHT/ACCELERATE/MAINTAIN); it "fires" when not already set,
        // which sets the bit and clears the other bit of its pair (apair = both bits of the pair)
        fire = amask & ~b;
        b = (b & ~(fire ? apair : 0u)) | fire;
        lane += ((fire & kBitRight) && lane < 3) ? 1 : 0;
        lane -= ((fire & kBitLeft) && lane > 1) ? 1 : 0;
        e.acc = py_clamp(e.acc, -5.0, 5.0);
        e.steer = py_clamp(e.steer, -30.0, 30.0);
    }
    // the obstacle that shares the car's lane (each lane holds exactly one).  A tick never fires both a lane
    // change and MAINTAIN, so the post-action lane is also the lane maintain() looks at.
    bool l1 = lane == 1, l2 = lane == 2;
    double fpx = e.opx[2], fiv = e.oiv[2];
    dmov_if(l2, fpx, e.opx[1]); dmov_if(l2, fiv, e.oiv[1]);
    dmov_if(l1, fpx, e.opx[0]); dmov_if(l1, fiv, e.oiv[0]);
    if (!CONT) {
        // maintain(): cruise = velocity of the nearest vehicle strictly ahead in the lane, else own velocity.  That
        // velocity is what its last update left: the stored one on the first tick, else init_vx or the car's vx.
        double fvx = FIRST ? e.ovx[2] : fiv;
        if (FIRST) { dmov_if(l2, fvx, e.ovx[1]); dmov_if(l1, fvx, e.ovx[0]); }
        else       dmov_if(e.slow, fvx, e.vx);
        double cr = e.vx;
        dmov_if(fpx > px, cr, fvx);
        dmov_if_nz(fire & kBitDoMaint, e.cruise, cr);
    }

    // ---- collision test on the rects the previous tick left (:920-937) ----
    // The car's rect x is 282 whenever px == 10 (always, after its first update), so an obstacle overlaps in x
    // iff 207 <= trunc(opx*32-38) <= 357 iff 7.65625 <= opx < 12.375.  The filter below is a superset of
    // that interval taken on the hi words (positive doubles order like their bit patterns); the exact integer
    // rect test runs only for the few lanes that pass it.
    int ncol = 0;
    if (!lock) {
        constexpr uint32_t kLo = 0x401EA000u, kSpan = 0x4028C000u - 0x401EA000u;  // hi(7.65625), hi(12.375)
        const bool n0 = hi32(e.opx[0]) - kLo <= kSpan, n1 = hi32(e.opx[1]) - kLo <= kSpan, n2 = hi32(e.opx[2]) - kLo <= kSpan;
        if (FIRST && px != 10.0) {
            // only a state the reference cannot produce (the car leaves x = 10 only at reset, under the collision lock)
            ncol = sa_collide_exact(px, e.py, e.opx[0], e.opx[1], e.opx[2]);
            ncoll_acc += ncol;
        } else if (n0 | n1 | n2) {
            // the car is pinned at x = 10: the x overlap is the interval itself, the y overlap of rect k (y = 21 + 80k,
            // height 38) with the car's rect is |ay - by| <= 37 (see sa_tick_later)
            const int u = rect_y(e.py) + 16;
            const int k = (int)(((uint32_t)u * 205u) >> 14);
            const double ox = dsel3(k + 1, e.opx[0], e.opx[1], e.opx[2]);
            ncol = ((uint32_t)u < 240u) & ((uint32_t)(u - 80 * k) <= 74u) & (ox >= 7.65625) & (ox < 12.375);
            ncoll_acc += ncol;
        }
    }

    // ---- car update (:140-229) ----
    if (!CONT) {
        const bool m_acc = (b & kBitDoAcc) != 0;
        const bool m_maint = (b & kBitDoMaint) != 0;  // never set together with do_accelerate
        // accelerate(): acc = 0 if acc < 0 else acc + dt ; `acc == 5.0` clears the mode
        double acc_a = dadd(e.acc, dt);
        dset_if(e.acc < 0.0, acc_a, 0.0);
        // maintain(): bang-bang towards ceil(cruise), snap when the ceilings agree
        const double vc = ceil(e.vx), cc = ceil(e.cruise);
        const bool snap = m_maint && vc == cc;
        const double acc_m = __hiloint2double(snap ? 0 : ((vc <= cc) ? 0x40140000 : (int)0xC0140000), 0);  // 0.0, 5.0, -5.0
        dmov_if_nz(b & kBitDoAcc, e.acc, acc_a);
        dmov_if_nz(b & kBitDoMaint, e.acc, acc_m);
        dmov_if(snap, e.vx, e.cruise);
        if (m_acc && e.acc == 5.0) b &= ~kBitDoAcc;
        if (snap) b &= ~kBitDoMaint;
    }
    e.vx = py_clamp(dadd(e.vx, dmul(e.acc, dt)), 0.0, 20.0);
    if (!CONT) {
        const double ty = __hiloint2double(l1 ? 0x3FF40000 : (l2 ? 0x400E0000 : 0x40190000), 0);  // 1.25, 3.75, 6.25
        const bool m_left = (b & kBitLeft) != 0;
        const bool m_right = (b & kBitRight) != 0;  // never set together with left_mode
        // adding +0.0 is the identity on every value steer can hold (it is never -0.0)
        double inc = 0.0;
        dmov_if(m_left, inc, K.k30dt); dmov_if(m_right, inc, -K.k30dt);
        e.steer = dadd(e.steer, inc);
        const bool reached = (m_left && e.py <= ty) || (m_right && e.py >= ty);
        dset_if(reached, e.steer, 0.0);
        if (reached) b &= ~(kBitLeft | kBitRight);
    }
    const double vdt = dmul(e.vx, dt);
    if (e.steer != 0.0) {
        // position += velocity * dt has velocity.y == 0: py + 0.0 == py for every py the clamps can produce
        if (CONT) e.py = dsub(e.py, dmul(angular_velocity<EXACT>(e.vx, e.steer), dt));
        else      e.py = dsub(e.py, lateral_step_d<EXACT>(e.vx, tan_small(dmul(e.steer, K.deg2rad)), dt, K.rad2deg_q));
    }
    e.py = road_clamp<CONT>(e.py);  // with ang == 0, py - 0.0 == py: only the clamp can move it
    e.rx = dadd(e.rx, vdt);
    if (CONT) {  // lane = nearest lane centre, first minimum wins
        const double d0 = fabs(dsub(e.py, 1.25)), d1 = fabs(dsub(e.py, 3.75)), d2 = fabs(dsub(e.py, 6.25));
        lane = (d0 <= d1 && d0 <= d2) ? 1 : (d1 <= d2) ? 2 : 3;
        l1 = lane == 1; l2 = lane == 2;
        fpx = e.opx[2]; fiv = e.oiv[2];
        dmov_if(l2, fpx, e.opx[1]); dmov_if(l2, fiv, e.oiv[1]);
        dmov_if(l1, fpx, e.opx[0]); dmov_if(l1, fiv, e.oiv[0]);
    }
    e.bits = (b & ~(3u << HW_BITS_LANE_SHIFT)) | ((uint32_t)lane << HW_BITS_LANE_SHIFT);

    // ---- obstacle update (:324-347): sees the leader's post-update lane, position (10) and velocity ----
    // velocity.x = max(-20, min(init_vx, 20)) is the identity: init_vx is 5..7 by construction (start grid, U[5,7)).
    // Only the obstacle in the car's lane can be held back: v = min(init_vx, car vx) when it is behind the car.
    e.slow = fpx < 10.0 && e.vx < fiv;
    {
        const int held = e.slow ? lane : 0;  // lane of the held-back obstacle, 0 = none
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            double d = dmul(e.oiv[k], dt);  // v * dt
            dmov_if(held == k + 1, d, vdt);
            e.opx[k] = dsub(dadd(e.opx[k], d), vdt);
            e.orx[k] = dadd(e.orx[k], d);
        }
    }

    // ---- respawn (:963-991) ----
    sa_respawn(e, K.respawn_hi, l1, l2, seed, env_id);
    return ncol;
}

// Ticks 2..T of a DISCRETE step.  Same arithmetic as sa_tick<false, false, .>, restated around what the first tick
// of the step has already established, to shrink the instruction stream the kernel is bound by:
//   * the car is pinned at px = 10; acc is in [-5, 5 + dt] and only the accelerate() branch can see a value
//     above 5, so the action stage's clamp of acc folds into that branch; steer is non-zero only while a
//     lane-change mode is set, and py only moves (and only then needs the road clamp) on a steering tick;
//   * the action of the step has fired on its first tick, so it can only fire again after the simulator
//     cleared the mode it asks for (lane reached, maintain snapped): that whole stage sits behind a branch;
//   * signs and "< 10.0" are read from the hi words (integer pipe) instead of fp64 compares.
//   * a collision ends the step after this tick (highway_env.py:99-108): the rare branch that finds one also sets the loop's
//     `remaining` tick count to 1, so the loop test is one decrement and one compare and knows nothing about collisions.
template <bool EXACT>
__device__ __forceinline__ void sa_tick_later(SaEnv &e, int &lane, int &ncol, int &remaining, uint32_t &ncoll_acc, const SaConsts &K, const bool inf_obs,
                                             const uint32_t amask, const uint32_t apair, const int dlane,
                                             const uint64_t seed, const uint64_t env_id, const double dt)
{
    uint32_t b = e.bits + 1;

    // ---- apply action (multi_lane_sim.py:502-517, :529-572) ----
    // Not rare: an env whose action is MAINTAIN with nobody ahead snaps and re-fires on every tick, and so does a
    // LEFT/RIGHT against the edge of the road, so the stage stays branch-free.  The two modes of a pair exclude
    // each other, hence setting the requested bit and clearing its partner is the identity when nothing fires.
    const uint32_t fire = amask & ~b;
    b = (b & ~apair) | amask;
    // lane id in its own register for the duration of the loop (the field in `bits` is refreshed after it);
    // dlane = -1 / +1 / 0 for LEFT / RIGHT / the other actions is fixed for the step
    lane = (fire & (kBitLeft | kBitRight)) ? min(max(lane + dlane, 1), 3) : lane;
    const bool l1 = lane == 1, l2 = lane == 2, l3 = lane == 3;
    double fpx, fiv;
    dsel3x2(lane, e.opx[0], e.opx[1], e.opx[2], e.oiv[0], e.oiv[1], e.oiv[2], fpx, fiv);
    {
        // maintain(): cruise = velocity of the vehicle ahead in the lane, else the car's own.  An obstacle that was held
        // back in the previous tick (e.slow) moved by fl(fl(x + vdt) - vdt) <= 10 from x < 10, so one that is ahead
        // (fpx > 10) moved with its own init_vx.
        double cr = dcopy(e.vx);
        dmov_if(fpx > 10.0, cr, fiv);
        dmov_if_nz(fire & kBitDoMaint, e.cruise, cr);
    }

    // ---- collision test on the rects the previous tick left (:920-937) ----
    // The car's rect x is 282, so obstacle k overlaps in x iff 207 <= trunc(opx*32-38) <= 357 iff 7.65625 <= opx < 12.375
    // (the scaling and the subtraction are exact there); the hi-word window below is a superset, and roughly every
    // second tick some lane of a warp passes it (obstacles in the other lanes drift past the car all the time), so
    // the exact test behind it is kept short: the y overlap of rect k (y = 21 + 80k, height 38) with the car's rect
    // is |ay - by| <= 37.
    {
        constexpr uint32_t kLo = 0x401EA000u, kSpan = 0x4028C000u - 0x401EA000u;  // hi(7.65625), hi(12.375)
        const bool n0 = hi32(e.opx[0]) - kLo <= kSpan, n1 = hi32(e.opx[1]) - kLo <= kSpan, n2 = hi32(e.opx[2]) - kLo <= kSpan;
        if (n0 | n1 | n2) {
            // rect k (y = 21 + 80k, height 38) overlaps the car's rect in y iff |ay - (21 + 80k)| <= 37; the three windows are
            // disjoint, so at most ONE lane can collide: pick it from ay and test that obstacle alone
            const int u = rect_y(e.py) + 16;                        // window k is u in [80k, 80k + 74]
            const int k = (int)(((uint32_t)u * 205u) >> 14);        // u / 80 for 0 <= u < 240
            const double ox = dsel3(k + 1, e.opx[0], e.opx[1], e.opx[2]);
            if (((uint32_t)u < 240u) & ((uint32_t)(u - 80 * k) <= 74u) & (ox >= 7.65625) & (ox < 12.375)) {
                ncol = 1; remaining = 1; ncoll_acc += 1;
            }
        }
    }

    // ---- car update (:185-229) ----
    const bool m_acc = (b & kBitDoAcc) != 0;
    const bool m_maint = (b & kBitDoMaint) != 0;  // never set together with do_accelerate
    {
        // accelerate(): acc = 0 if acc < 0 else acc + dt, on the acc the action stage clamped to <= 5
        double acc_a = dadd(e.acc, dt);
        dmov_if(5.0 < e.acc, acc_a, K.five_dt);  // min(acc, 5) + dt
        dset_if(e.acc < 0.0, acc_a, 0.0);
        // maintain(): bang-bang towards ceil(cruise), snap when the ceilings agree
        const double vc = ceil(e.vx), cc = ceil(e.cruise);
        const bool snap = m_maint & (vc == cc);
        const double acc_m = __hiloint2double(snap ? 0 : ((vc <= cc) ? 0x40140000 : (int)0xC0140000), 0);  // 0.0, 5.0, -5.0
        dmov_if_nz(b & kBitDoAcc, e.acc, acc_a);
        dmov_if_nz(b & kBitDoMaint, e.acc, acc_m);
        dmov_if(snap, e.vx, e.cruise);
        if (m_acc & (e.acc == 5.0)) b &= ~kBitDoAcc;
        if (snap) b &= ~kBitDoMaint;
    }
    {
        double v = dadd(e.vx, dmul(e.acc, dt));
        const bool over = 20.0 < v, under = !(v > 0.0);  // max(0.0, v) is v only when v > 0; both tests on the value before either move
        dset_if(over, v, 20.0);
        dset_if(under, v, 0.0);
        e.vx = v;
    }
    const double vdt = dmul(e.vx, dt);
    e.rx = dadd(e.rx, vdt);
    if (b & (kBitLeft | kBitRight)) {  // the two never come together
        // steering clamp of the action stage (only a lane-change mode leaves steer != 0), then moveLeft/moveRight
        double sc = e.steer;
        dmov_if(30.0 < fabs(e.steer), sc, copysign(30.0, e.steer));
        // +30*dt in left mode, -30*dt in right mode: the right-mode bit (bit 19) moved onto the sign bit of the increment
        const double k30 = K.k30dt;
        const double inc = __hiloint2double(__double2hiint(k30) ^ (int)((b << 12) & 0x80000000u), __double2loint(k30));
        double st = dadd(sc, inc);
        const double ty = __hiloint2double(l1 ? 0x3FF40000 : (l2 ? 0x400E0000 : 0x40190000), 0);  // 1.25, 3.75, 6.25
        int reached_i;  // (left_mode and y <= target) or (right_mode and y >= target): two compares and one predicate op
        asm("{\n\t.reg .pred a, c, l, nl, r;\n\tsetp.ne.b32 l, %3, 0;\n\tnot.pred nl, l;\n\t"
            "setp.le.and.f64 a, %1, %2, l;\n\tsetp.ge.and.f64 c, %1, %2, nl;\n\tor.pred r, a, c;\n\tselp.s32 %0, 1, 0, r;\n\t}"
            : "=r"(reached_i) : "d"(e.py), "d"(ty), "r"(b & kBitLeft));
        const bool reached = reached_i != 0;
        dset_if(reached, st, 0.0);
        if (reached) b &= ~(kBitLeft | kBitRight);
        e.steer = st;
        // when the lane has just been reached st is 0: tan(0) = 0, the step is 0 (also through the exact path's vx / inf) and
        // py - 0 = py lies inside the road clamp already - no second branch around the steering arithmetic
        e.py = road_clamp<false>(dsub(e.py, lateral_step_d<EXACT>(e.vx, tan_small(dmul(st, K.deg2rad)), dt, K.rad2deg_q)));
    }
    e.bits = b;

    // ---- obstacle update (:324-347): only the obstacle in the car's lane can be held back ----
    {
        e.slow = (fpx < 10.0) & (e.vx < fiv);
        double d0 = dmul(e.oiv[0], dt), d1 = dmul(e.oiv[1], dt), d2 = dmul(e.oiv[2], dt);
        dmov_if(l1 & e.slow, d0, vdt); dmov_if(l2 & e.slow, d1, vdt); dmov_if(l3 & e.slow, d2, vdt);
        e.opx[0] = dsub(dadd(e.opx[0], d0), vdt); e.orx[0] = dadd(e.orx[0], d0);
        e.opx[1] = dsub(dadd(e.opx[1], d1), vdt); e.orx[1] = dadd(e.orx[1], d1);
        e.opx[2] = dsub(dadd(e.opx[2], d2), vdt); e.orx[2] = dadd(e.orx[2], d2);
    }

    // ---- respawn (:963-991) ----
    sa_respawn(e, K.respawn_hi, l1, l2, seed, env_id);
}

// velocities of the obstacles after their last update, materialised for the observation and the stored state
__device__ __forceinline__ void sa_materialise_ovx(SaEnv &e)
{
    const int lane = (int)((e.bits >> HW_BITS_LANE_SHIFT) & 3u);
#pragma unroll
    for (int k = 0; k < 3; ++k) e.ovx[k] = (e.slow && lane == k + 1) ? e.vx : e.oiv[k];
}

// Where an observation row goes.  SmRow: the lane's row of the warp-private shared-memory tile, addressed in the shared
// state space with a 32-bit address (a generic float2* made ptxas rebuild the shared window base - S2R CgaCtaId, LEA, IMAD -
// at every group of stores); PtrRow: plain memory (reset kernel).
struct SmRow {
    uint32_t addr;
    __device__ __forceinline__ void put(int j, float2 v) const
    {
        asm volatile("st.shared.v2.f32 [%0], {%1, %2};" :: "r"(addr + 8u * (uint32_t)j), "f"(v.x), "f"(v.y) : "memory");
    }
};
struct PtrRow {
    float2 *p;
    __device__ __forceinline__ void put(int j, float2 v) const { p[j] = v; }
};

// observation pairs (x, y) j = 0..8 -> obs[2j], obs[2j+1]; float32 clip + normalise.
// STEPPED = true (the state has just gone through at least one tick, or is the start grid): py is in [0, 7], vx in [0, 20]
// and every obstacle velocity is its init_vx (5..7, DESIGN.md deviation 5) or the car's vx, so np.clip is the identity on
// py, the three lane offsets, vx and the three relative velocities - 16 compare-selects per step the kernel does not issue.
// acc (<= 5 + dt), steer (<= 30 + 30*dt), raw x (<= 1220) and the obstacle distances (spawns 70..100 ahead) do clip.
// COMPACT (HW_FLAG_COMPACT_OBS): rows of 14 floats - columns 11, 13, 15, 17 (the lateral velocities, always zero -> the constant 0.5) are
// dropped, i.e. the row is the first ten columns followed by {vx, dv0, dv1, dv2}.
template <bool STEPPED, class Row, bool COMPACT = false>
__device__ __forceinline__ void sa_observe(const SaEnv &e, const Row row)
{
    constexpr bool C = !STEPPED;
    row.put(0, make_float2(norm_f32(e.acc, -5.f, 5.f), norm_f32(e.steer, -30.f, 30.f)));
    row.put(1, make_float2(norm_f32<true>(e.rx, 0.f, 1200.f), norm_f32<false, C>(e.py, 0.f, 8.f)));
    float dv[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        row.put(2 + k, make_float2(norm_f32(dsub(e.orx[k], e.rx), -60.f, 60.f),
                                   norm_f32<false, C>(dsub(lane_centre(k), e.py), -8.f, 8.f)));
        dv[k] = norm_f32<false, C>(dsub(e.ovx[k], e.vx), -20.f, 20.f);
    }
    const float vxn = norm_f32<false, C>(e.vx, -20.f, 20.f);
    if (COMPACT) {
        row.put(5, make_float2(vxn, dv[0]));
        row.put(6, make_float2(dv[1], dv[2]));
    } else {
        row.put(5, make_float2(vxn, 0.5f));
#pragma unroll
        for (int k = 0; k < 3; ++k) row.put(6 + k, make_float2(dv[k], 0.5f));
    }
}

// Observation + state store of an environment that has just been stepped (no reset), without materialising ovx[] in
// fp64: the velocity an obstacle had after its last update is its init_vx, except for the one that is being held back
// behind the car (`slow`, in the car's lane), which moved with the car's vx - its relative velocity is exactly 0 -> 0.5.
template <bool COMPACT>
__device__ __forceinline__ void sa_finish_stepped(const SaEnv &e, const SmRow row, const SaPtrs &S, uint32_t i)
{
    const int held = e.slow ? (int)((e.bits >> HW_BITS_LANE_SHIFT) & 3u) : 0;
    const float vxf = (float)e.vx, pyf = (float)e.py, rxf = (float)e.rx, accf = (float)e.acc, steerf = (float)e.steer;
    row.put(0, make_float2(norm_f32(e.acc, -5.f, 5.f), norm_f32(e.steer, -30.f, 30.f)));
    row.put(1, make_float2(norm_f32<true>(e.rx, 0.f, 1200.f), norm_f32<false, false>(e.py, 0.f, 8.f)));
    const float vxn = norm_f32<false, false>(e.vx, -20.f, 20.f);
    float ovf[3], oif[3], dvn[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        row.put(2 + k, make_float2(norm_f32(dsub(e.orx[k], e.rx), -60.f, 60.f),
                                   norm_f32<false, false>(dsub(lane_centre(k), e.py), -8.f, 8.f)));
        const float dv = norm_f32<false, false>(dsub(e.oiv[k], e.vx), -20.f, 20.f);
        dvn[k] = held == k + 1 ? 0.5f : dv;
        oif[k] = (float)e.oiv[k];
        ovf[k] = held == k + 1 ? vxf : oif[k];
    }
    if (COMPACT) {
        row.put(5, make_float2(vxn, dvn[0]));
        row.put(6, make_float2(dvn[1], dvn[2]));
    } else {
        row.put(5, make_float2(vxn, 0.5f));
#pragma unroll
        for (int k = 0; k < 3; ++k) row.put(6 + k, make_float2(dvn[k], 0.5f));
    }
    S.car_a[i] = make_float4((float)e.px, pyf, rxf, vxf);
    S.car_b[i] = make_float4(accf, steerf, (float)e.cruise, __uint_as_float(e.bits));
    S.obst0[i] = make_float4((float)e.opx[0], (float)e.orx[0], ovf[0], oif[0]);
    S.obst1[i] = make_float4((float)e.opx[1], (float)e.orx[1], ovf[1], oif[1]);
    S.obst2[i] = make_float4((float)e.opx[2], (float)e.orx[2], ovf[2], oif[2]);
}

// One warp's observations (32 rows x 18 floats = 2304 contiguous bytes in global memory) go through a
// warp-private shared-memory tile: each lane writes its row as nine 8-byte words (conflict-free: the row
// stride is 9 words of 8 bytes, odd), then the warp streams the tile out as 144 coalesced 16-byte stores.
// Only __syncwarp() is needed, so warps never wait for each other.
__device__ __forceinline__ float4 lds128(uint32_t addr)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}

// tile_s: shared-state-space address of the warp's tile; D = floats per row (18, or 14 with HW_FLAG_COMPACT_OBS: 32 x 56 bytes = 112 stores)
template <int D = HW_SA_OBS_DIM>
__device__ __forceinline__ void store_obs_warp(const uint32_t tile_s, float *__restrict__ obs, int64_t warp_base, int64_t n)
{
    __syncwarp();
    const uint32_t lane = threadIdx.x & 31;
    const int64_t rows = min((int64_t)32, n - warp_base);
    float *dst = obs + warp_base * D;
    if (rows == 32) {
        float4 *d4 = reinterpret_cast<float4 *>(dst);
        constexpr int kVec = 32 * D / 4, kFull = kVec / 32, kRem = kVec % 32;  // 144 = 4 x 32 + 16, 112 = 3 x 32 + 16
#pragma unroll
        for (int v = 0; v < kFull; ++v) d4[lane + 32 * v] = lds128(tile_s + 16u * (lane + 32 * v));
        if (kRem && lane < kRem) d4[lane + 32 * kFull] = lds128(tile_s + 16u * (lane + 32 * kFull));
    } else {
        for (uint32_t f = lane; f < (uint32_t)rows * D; f += 32) {
            float v;
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(tile_s + 4u * f) : "memory");
            dst[f] = v;
        }
    }
}

// Resident blocks per SM the register allocator is asked to make room for: 16 blocks of 64 threads = 64 registers per
// thread = 32 warps/SM.  Measured on B200 at 4M envs (discrete): round 1, 128-thread blocks: 7 blocks (72 registers) 316 us,
// 8 blocks 307 us, 9 blocks (56 registers, spills) 374 us, 10 blocks 399 us; round 2 (spill-free kernel): 256-thread blocks
// 298.9 us, 128-thread 290.8 - 294.8 us, 64-thread 288.8 - 292.8 us, 72 registers 292.9 us (continuous 268 vs 258 us).
#ifndef HW_SA_MIN_BLOCKS_D
#define HW_SA_MIN_BLOCKS_D (2048 / HW_SA_BLOCK / 2)
#endif
#ifndef HW_SA_MIN_BLOCKS_C
#define HW_SA_MIN_BLOCKS_C (2048 / HW_SA_BLOCK / 2)
#endif

template <bool CONT, bool RANDOM_ACTIONS, bool EXACT, bool COMPACT = false>
__global__ void __launch_bounds__(kSaBlock, CONT ? HW_SA_MIN_BLOCKS_C : HW_SA_MIN_BLOCKS_D)
sa_step_kernel(SaPtrs S, const void *__restrict__ actions, float *__restrict__ obs, float *__restrict__ rew,
               uint8_t *__restrict__ done, int4 *__restrict__ term, unsigned long long *__restrict__ stats,
               const SaConsts K, const uint64_t seed, const uint64_t env_id0, const uint64_t action_ctr0,
               const int steps_arg, const int64_t n, const int flags)
{
    const int steps = RANDOM_ACTIONS ? steps_arg : 1;  // step() proper is one step per launch: lets the episode bookkeeping fold
    constexpr int D = COMPACT ? HW_SA_OBS_DIM_COMPACT : HW_SA_OBS_DIM;  // floats per observation row

    __shared__ __align__(16) float sm_obs[kSaBlock * D];
    const uint32_t i = blockIdx.x * kSaBlock + threadIdx.x;  // < 2^31, checked by the launcher
    const int64_t warp_base = (int64_t)(i - (threadIdx.x & 31));
    const uint32_t tile_s = (uint32_t)__cvta_generic_to_shared(sm_obs) + (threadIdx.x >> 5) * (32 * D * 4);
    if (warp_base >= n) return;  // whole warp out of range
    const bool inf_obs = !(flags & HW_FLAG_NO_INF_OBS);
    const bool autoreset = !(flags & HW_FLAG_NO_AUTORESET);
    // dt is used by a dozen instructions of every tick.  Left to itself ptxas re-loads it from the constant bank at the top of
    // every tick (LDC.64 + a scoreboard wait in front of the first multiply: 2.8 % of the step, measured); xor-ing the high
    // word with blockIdx.y (always 0) makes it a value that cannot be rematerialised, so it stays in a register pair.
    // (Doing the same for 30*dt and 5+dt costs more in registers than the loads it saves: 288.8 us against 280.7 us.)
    const double dt = __hiloint2double(__double2hiint(K.dt) ^ (int)blockIdx.y, __double2loint(K.dt));
    long long acc_eps = 0, acc_len = 0, acc_ret = 0, acc_col = 0, acc_to = 0;  // finished episodes of this thread

    if (i < n) {
        SaEnv e;
        sa_load(e, S, i);
        const uint64_t env_id = env_id0 + (uint64_t)i;

        // episode accounting (Monitor) lives outside the tick loop
        uint32_t ncoll_step = 0;       // obstacle collisions counted during this launch (since the last reset)
        int ret_milli_step = 0;        // rounded rewards since the last reset inside this launch, in 1/1000
        uint32_t len_step = 0;         // steps since the last reset inside this launch
        bool was_reset = false;        // an auto-reset happened inside this launch
        double reward = 0.0;
        int rew_milli = 0;
        bool is_done = false;
        for (int s = 0; s < steps; ++s) {
            int act_d = 0;
            double a0_5dt = 0.0, a1_30dt = 0.0;
            uint32_t amask = 0, apair = 0;
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
            if (!CONT) {  // ACTION_LOOKUP (highway_env.py:14-20) as the mode bit it requests
                // LEFT, RIGHT, ACCELERATE, MAINTAIN = 0..3 ask for the mode bits 18..21 in that order: shifts instead of a
                // four-way select (which nvcc turns into a divergent jump table); any other value requests nothing
                static_assert(HW_ACT_LEFT == 0 && HW_ACT_RIGHT == 1 && HW_ACT_ACCELERATE == 2 && HW_ACT_MAINTAIN == 3, "action codes");
                static_assert(kBitRight == kBitLeft << 1 && kBitDoAcc == kBitLeft << 2 && kBitDoMaint == kBitLeft << 3, "mode bits");
                const bool valid = (uint32_t)act_d < 4u;
                amask = valid ? (kBitLeft << (act_d & 3)) : 0u;
                apair = valid ? ((kBitLeft | kBitRight) << (act_d & 2)) : 0u;
            }
            // env.step: up to ticks_per_step ticks, stop at the first crash (highway_env.py:99-108)
            int ncol = sa_tick<CONT, true, EXACT>(e, ncoll_step, K, inf_obs, amask, apair, a0_5dt, a1_30dt, seed, env_id, dt);
            e.px = 10.0;  // the only car is always the leader (:212-213)
            if (CONT) {
                for (int t = 1; t < K.ticks_per_step && ncol == 0; ++t)
                    ncol = sa_tick<CONT, false, EXACT>(e, ncoll_step, K, inf_obs, amask, apair, a0_5dt, a1_30dt, seed, env_id, dt);
            } else {
                int lane = (int)((e.bits >> HW_BITS_LANE_SHIFT) & 3u);
                const int dlane = ((uint32_t)act_d < 2u) ? 2 * act_d - 1 : 0;  // LEFT -1, RIGHT +1
                for (int remaining = (ncol == 0) ? K.ticks_per_step - 1 : 0; remaining > 0; --remaining)
                    sa_tick_later<EXACT>(e, lane, ncol, remaining, ncoll_step, K, inf_obs, amask, apair, dlane, seed, env_id, dt);
                e.bits = (e.bits & ~(3u << HW_BITS_LANE_SHIFT)) | ((uint32_t)lane << HW_BITS_LANE_SHIFT);
            }
            const uint32_t tick = e.bits & HW_BITS_TICK_MASK;
            const bool timeout = tick >= (uint32_t)K.timeout_tick;
            is_done = timeout || ncol > 0;
            // the three lanes' rect windows are disjoint in y: a tick counts at most one overlapping obstacle, so -250 * n is -250
            reward = (ncol > 0) ? -250.0 : dsub(ddiv_const(e.vx, 20.0), 1.0);
            rew_milli = round_milli(reward);
            if (!RANDOM_ACTIONS) {
                // ---- epilogue of step() proper (one step per launch) ----
                // The finished-episode path never merges with the live state: an auto-reset environment stores the start grid
                // and its observation (constants), every other one stores what the ticks left.  (With the reset written as
                // `e = start grid` ahead of a common store, ptxas spilled the whole 19-double state around the branch.)
                rew[i] = (float)milli_to_reward(rew_milli, reward);
                done[i] = is_done ? 1 : 0;
                const SmRow row{tile_s + (threadIdx.x & 31) * (D * 4)};
                uint4 st = S.stat[i];  // per-env counters: read-modify-write once per launch
                st.x += ncoll_step; st.y = e.spawn_ctr; st.z = (uint32_t)((int)st.z + rew_milli); st.w += 1;
                if (is_done) {  // Monitor / info of the finished episode
                    if (term) term[i] = make_int4((int)tick, (int)st.x, (int)st.z, (int)st.w);
                    acc_eps += 1; acc_len += st.w; acc_ret += (int)st.z; acc_col += st.x; acc_to += (ncol == 0) ? 1 : 0;
                }
                if (is_done && autoreset) {
                    SaEnv r;
                    sa_reset_env(r);
                    r.spawn_ctr = 0;
                    sa_observe<false, SmRow, COMPACT>(r, row);
                    sa_store(r, S, i);
                    st.x = 0; st.z = 0; st.w = 0;
                } else {
                    // gym per-env semantics (no auto-reset) stepped far past the time-out: the 16-bit tick field saturates
                    // instead of carrying into the lane bits (the episode has been over for 60000 ticks; only `done` matters)
                    if (tick >= 0xFF00u) e.bits = (e.bits & ~HW_BITS_TICK_MASK) | 0xFF00u;
                    sa_finish_stepped<COMPACT>(e, row, S, i);
                }
                S.stat[i] = st;
            } else {
                ret_milli_step += rew_milli;
                len_step += 1;
                sa_materialise_ovx(e);
                if (is_done) {
                    // Monitor / info of the finished episode: state counters + what this launch added
                    const uint4 st = S.stat[i];
                    const uint32_t nc = (was_reset ? 0u : st.x) + ncoll_step;
                    const int er = (was_reset ? 0 : (int)st.z) + ret_milli_step;
                    const uint32_t el = (was_reset ? 0u : st.w) + len_step;
                    if (term) term[i] = make_int4((int)tick, (int)nc, er, (int)el);
                    acc_eps += 1; acc_len += el; acc_ret += er; acc_col += nc; acc_to += (ncol == 0) ? 1 : 0;
                    if (autoreset) {
                        sa_reset_env(e);
                        was_reset = true; ncoll_step = 0; ret_milli_step = 0; len_step = 0;
                    } else if (tick >= 0xFF00u) {
                        e.bits = (e.bits & ~HW_BITS_TICK_MASK) | 0xFF00u;
                    }
                }
            }
        }
        if (RANDOM_ACTIONS) {
            rew[i] = (float)milli_to_reward(rew_milli, reward);
            done[i] = is_done ? 1 : 0;
            sa_observe<true, SmRow, COMPACT>(e, SmRow{tile_s + (threadIdx.x & 31) * (D * 4)});
            sa_store(e, S, i);
            {   // per-env counters: read-modify-write once per launch
                uint4 st = S.stat[i];
                if (was_reset) { st.x = 0; st.z = 0; st.w = 0; }
                st.x += ncoll_step; st.y = e.spawn_ctr; st.z = (uint32_t)((int)st.z + ret_milli_step); st.w += len_step;
                S.stat[i] = st;
            }
        }
    }
    store_obs_warp<D>(tile_s, obs, warp_base, n);
    if (stats && __any_sync(0xffffffffu, acc_eps != 0)) {  // all 32 lanes are here (warps entirely out of range returned at the top)
        {
            const long long eps = warp_sum(acc_eps);
            const long long len = warp_sum(acc_len), ret = warp_sum(acc_ret), col = warp_sum(acc_col), to = warp_sum(acc_to);
            if ((threadIdx.x & 31) == 0) {
                atomicAdd(&stats[0], (unsigned long long)eps);
                atomicAdd(&stats[1], (unsigned long long)len);
                atomicAdd(&stats[2], (unsigned long long)ret);
                atomicAdd(&stats[3], (unsigned long long)col);
                atomicAdd(&stats[4], (unsigned long long)to);
            }
        }
    }
}

__global__ void __launch_bounds__(kSaBlock)
sa_reset_kernel(SaPtrs S, const uint8_t *__restrict__ mask, float *__restrict__ obs, const int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * kSaBlock + threadIdx.x;
    if (i >= n) return;
    if (mask && !mask[i]) return;
    SaEnv e;
    sa_reset_env(e);
    sa_store(e, S, i);
    S.stat[i] = make_uint4(0u, S.stat[i].y, 0u, 0u);  // spawn_ctr survives so that every episode sees fresh obstacles
    if (obs) {
        sa_observe<false>(e, PtrRow{reinterpret_cast<float2 *>(obs + i * HW_SA_OBS_DIM)});
    }
}

__global__ void __launch_bounds__(kSaBlock)
sa_observe_kernel(SaPtrs S, float *__restrict__ obs, const int64_t n)
{
    __shared__ __align__(16) float sm_obs[kSaBlock * HW_SA_OBS_DIM];
    const int64_t i = (int64_t)blockIdx.x * kSaBlock + threadIdx.x;
    const int64_t warp_base = i - (threadIdx.x & 31);
    const uint32_t tile_s = (uint32_t)__cvta_generic_to_shared(sm_obs) + (threadIdx.x >> 5) * (32 * HW_SA_OBS_DIM * 4);
    if (warp_base >= n) return;
    if (i < n) {
        SaEnv e;
        sa_load(e, S, i);
        sa_observe<false>(e, SmRow{tile_s + (threadIdx.x & 31) * (HW_SA_OBS_DIM * 4)});
    }
    store_obs_warp(tile_s, obs, warp_base, n);
}

static inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

static int check_state(const HwSaState *st, int64_t n)
{
    if (!st || n <= 0 || !st->car_a || !st->car_b || !st->obst || !st->stat) return HW_E_BADARG;
    if (!aligned16(st->car_a) || !aligned16(st->car_b) || !aligned16(st->obst) || !aligned16(st->stat)) return HW_E_ALIGN;
    return 0;
}

// pointers of the environments [i0, n) of a batch of n
static SaPtrs ptrs(const HwSaState *st, int64_t n, int64_t i0 = 0)
{
    SaPtrs S;
    S.car_a = reinterpret_cast<float4 *>(st->car_a) + i0;
    S.car_b = reinterpret_cast<float4 *>(st->car_b) + i0;
    S.obst0 = reinterpret_cast<float4 *>(st->obst) + i0;
    S.obst1 = S.obst0 + n;
    S.obst2 = S.obst1 + n;
    S.stat = reinterpret_cast<uint4 *>(st->stat) + i0;
    return S;
}

// step launch over the environments [i0, i0 + count) of a batch of n (count == n, i0 == 0: the whole batch)
static int launch_step_range(const HwSaState *st, const void *actions, const HwSaOut *out, const HwSimParams *P,
                             uint64_t seed, uint64_t env_id0, uint64_t action_ctr0, int steps, bool random_actions,
                             int64_t n, int64_t i0, int64_t count, int flags, cudaStream_t stream)
{
    int rc = check_state(st, n);
    if (rc) return rc;
    if (!out || !out->obs || !out->rew || !out->done || !P) return HW_E_BADARG;
    if (!random_actions && !actions) return HW_E_BADARG;
    if (steps < 1 || P->ticks_per_step < 1 || P->timeout_tick < 1 || P->timeout_tick > 0xFFFF || !(P->dt > 0.0)) return HW_E_BADARG;
    if (!aligned16(out->obs) || (out->term && !aligned16(out->term))) return HW_E_ALIGN;
    if (i0 < 0 || count < 1 || i0 + count > n || (i0 & 1)) return HW_E_BADARG;  // an even i0 keeps the observation rows 16-byte aligned
    const bool cont = (flags & HW_FLAG_CONTINUOUS) != 0;
    const bool exact = (flags & HW_FLAG_EXACT_STEERING) != 0;
    const bool compact = (flags & HW_FLAG_COMPACT_OBS) != 0;
    if (compact && random_actions) return HW_E_BADARG;  // the compact rows are an option of step() / step_host() only
    const int obs_dim = compact ? HW_SA_OBS_DIM_COMPACT : HW_SA_OBS_DIM;
    if (count >= ((int64_t)1 << 31)) return HW_E_BADARG;  // environments of one launch are indexed with 32 bits
    const SaPtrs S = ptrs(st, n, i0);
    const void *act = actions ? static_cast<const char *>(actions) + (size_t)i0 * (cont ? 2 * sizeof(float) : sizeof(int32_t)) : nullptr;
    float *obs = out->obs + i0 * obs_dim, *rew = out->rew + i0;
    uint8_t *done = out->done + i0;
    int4 *term = out->term ? reinterpret_cast<int4 *>(out->term) + i0 : nullptr;
    SaConsts K;
    K.dt = P->dt;
    K.k30dt = 30.0 * P->dt;  // python: 30.0 * dt
    K.five_dt = 5.0 + P->dt;
    K.deg2rad = kDeg2Rad;
    K.respawn_hi = (flags & HW_FLAG_NO_INF_OBS) ? 0xFFFFFFFFu : 0xC0030000u;
    K.rad2deg_q = 0.25 * kRad2Deg;  // exact scaling of the constant
    K.ticks_per_step = P->ticks_per_step;
    K.timeout_tick = P->timeout_tick;
    const unsigned grid = (unsigned)((count + kSaBlock - 1) / kSaBlock);
#define HW_LAUNCH(C, R, E, P) sa_step_kernel<C, R, E, P><<<grid, kSaBlock, 0, stream>>>(S, act, obs, rew, done, term, out->stats, K, seed, env_id0 + (uint64_t)i0, action_ctr0, steps, count, flags)
#define HW_LAUNCH_E(C, R, P) do { if (exact) HW_LAUNCH(C, R, true, P); else HW_LAUNCH(C, R, false, P); } while (0)
    if (compact)   { if (cont) HW_LAUNCH_E(true, false, true); else HW_LAUNCH_E(false, false, true); }
    else if (cont) { if (random_actions) HW_LAUNCH_E(true, true, false); else HW_LAUNCH_E(true, false, false); }
    else           { if (random_actions) HW_LAUNCH_E(false, true, false); else HW_LAUNCH_E(false, false, false); }
#undef HW_LAUNCH_E
#undef HW_LAUNCH
    return (int)cudaGetLastError();
}

static int launch_step(const HwSaState *st, const void *actions, const HwSaOut *out, const HwSimParams *P,
                       uint64_t seed, uint64_t env_id0, uint64_t action_ctr0, int steps, bool random_actions,
                       int64_t n, int flags, cudaStream_t stream)
{
    return launch_step_range(st, actions, out, P, seed, env_id0, action_ctr0, steps, random_actions, n, 0, n, flags, stream);
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
    if (n >= ((int64_t)1 << 31)) return HW_E_BADARG;
    const hw::SaPtrs S = hw::ptrs(state, n);
    hw::sa_reset_kernel<<<grid, hw::kSaBlock, 0, (cudaStream_t)stream>>>(S, mask, obs, n);
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
    if (n >= ((int64_t)1 << 31)) return HW_E_BADARG;
    const hw::SaPtrs S = hw::ptrs(state, n);
    hw::sa_observe_kernel<<<grid, hw::kSaBlock, 0, (cudaStream_t)stream>>>(S, obs, n);
    return (int)cudaGetLastError();
}

int hw_sa_step_host(const HwSaState *state, const void *h_actions, void *d_actions, const HwSaOut *d_out,
                    float *h_obs, float *h_rew, uint8_t *h_done, const HwSimParams *params,
                    uint64_t seed, uint64_t env_id0, int64_t n, int flags, void *stream, void *aux_stream)
{
    if (!h_actions || !d_actions || !d_out || !h_obs || !h_rew || !h_done || n <= 0) return HW_E_BADARG;
    cudaStream_t s = (cudaStream_t)stream, s2 = (cudaStream_t)aux_stream;
    const size_t asize = (flags & HW_FLAG_CONTINUOUS) ? 2 * sizeof(float) : sizeof(int32_t);
    // The call is bound by the device-to-host copy of the observations (72 of the 81 bytes per environment that cross PCIe).
    // Large batches are therefore cut into two halves that alternate between the caller's stream and the caller's auxiliary
    // stream (both live as long as the environment: nothing is created per call but one untimed event), so that the action
    // upload and the kernel of the second half run under the download of the first.
    // (Measured on B200 / PCIe Gen5 at 4 M envs: 6.87 ms unsplit, 6.73 ms in two halves, worse again from eight chunks on:
    // the download alone is 6.3 ms at the link's ~51 GB/s, so there is little left to hide.)
    const int obs_dim = (flags & HW_FLAG_COMPACT_OBS) ? HW_SA_OBS_DIM_COMPACT : HW_SA_OBS_DIM;
    const int chunks = (n >= 262144 && s2 != nullptr && s2 != s) ? 2 : 1;
    const int64_t per = (((n + chunks - 1) / chunks) + 127) & ~(int64_t)127;
    cudaEvent_t ev = nullptr;
    cudaError_t ce = cudaSuccess;
    if (chunks > 1) {
        if ((ce = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming)) != cudaSuccess) return (int)ce;
        cudaEventRecord(ev, s);           // the auxiliary stream starts after whatever the caller has queued
        cudaStreamWaitEvent(s2, ev, 0);
    }
    int rc = 0;
    for (int c = 0; c < chunks && rc == 0; ++c) {
        const int64_t i0 = c * per, cnt = (i0 + per <= n) ? per : n - i0;
        if (cnt <= 0) break;
        cudaStream_t q = (c & 1) ? s2 : s;
        ce = cudaMemcpyAsync(static_cast<char *>(d_actions) + i0 * asize, static_cast<const char *>(h_actions) + i0 * asize, cnt * asize, cudaMemcpyHostToDevice, q);
        if (ce != cudaSuccess) { rc = (int)ce; break; }
        rc = hw::launch_step_range(state, d_actions, d_out, params, seed, env_id0, 0, 1, false, n, i0, cnt, flags, q);
        if (rc) break;
        ce = cudaMemcpyAsync(h_obs + i0 * obs_dim, d_out->obs + i0 * obs_dim, (size_t)cnt * obs_dim * sizeof(float), cudaMemcpyDeviceToHost, q);
        if (ce == cudaSuccess) ce = cudaMemcpyAsync(h_rew + i0, d_out->rew + i0, (size_t)cnt * sizeof(float), cudaMemcpyDeviceToHost, q);
        if (ce == cudaSuccess) ce = cudaMemcpyAsync(h_done + i0, d_out->done + i0, (size_t)cnt, cudaMemcpyDeviceToHost, q);
        if (ce != cudaSuccess) rc = (int)ce;
    }
    if (chunks > 1) {
        cudaEventRecord(ev, s2);          // and the caller's stream ends after the auxiliary one
        cudaStreamWaitEvent(s, ev, 0);
        cudaEventDestroy(ev);             // released once the recorded work has completed; the stream wait above is already queued
    }
    ce = cudaStreamSynchronize(s);
    return rc ? rc : (int)ce;
}

}  // extern "C"
