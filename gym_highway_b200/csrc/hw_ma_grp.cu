// Multi-agent highway step for sm_100a, G lanes per environment (G = 8: four environments per warp).
//
// Same path as hw_ma.cu (thread per environment): MultiAgentEnv.step, highway_core.HighwaySimulator.act,
// agent.Car.update_d/update_c, agent.Obstacle.update, check_collisions, Scenario.rewards/observations, the
// auto-reset of DummyVecMAEnv and Monitor's episode record (file:line citations at each stage below).
//
// Why a second layout.  With one thread per environment the fp64 working set of an environment (5 cars x 8 fields +
// 16 obstacle slots x 4 fields = 832 B) has to sit in shared memory, which caps an SM at 8 warps; the kernel
// is bound by the latency of one thread walking cars and obstacles one after the other (issue slots 28 % busy).
// Here lane g of an environment's group IS car g (its state stays in registers for the whole step) and owns the
// obstacle slots g, g + G, ...; what the other lanes need of a car (position, velocity, lane, rect) is published
// in shared memory, and the stages of a tick are separated by __syncwarp on the group's lanes.  The order-
// dependent semantics of the reference survive the parallel evaluation because every cross-car read is of a
// value whose version (before / after this tick's update) is known statically:
//   * actions run for every car in id order (highway_core.py:280-281): maintain() of car a sees the lanes cars j < a
//     have just switched to and the old lanes of cars j > a; a lane change depends only on the car's own action
//     and flags, so all cars switch first (old and new lane are both published) and then search;
//   * cars update in id order and subtract the leader's velocity "as it is now" (agent.py:147-151): its new value
//     for cars behind it in id order, its old value for cars before it; velocities are double-buffered;
//   * obstacles, the leader election, spawn / retire and the collision test each read only finished values.
// Scalars of the environment header (tick, leader, counters, slot bookkeeping) are replicated in the registers
// of all G lanes and evolve identically; the serial slot bookkeeping of a spawn or a retire runs on lane 0 and is
// broadcast with shuffles.
#include "hw_ma_common.cuh"

namespace hw {

constexpr int kGrpWarps = 8;  // warps per block

template <int G>
struct GrpLayout {
    static constexpr int kEnvsPerWarp = 32 / G;
    static constexpr int kEnvs = kEnvsPerWarp * kGrpWarps;  // environments per block
    static constexpr int kRounds = (kMaxK + G - 1) / G;      // obstacle slots per lane
    // published car fields (doubles): px, py, rx, vy, vx[2] (double buffered); ints: lane_old, lane, rect x, rect y
    static constexpr int kCarD = 6, kCarI = 4;
    static constexpr size_t kDoubles = (size_t)(kCarD + 1) * kEnvs * G + (size_t)O_NF * kRounds * kEnvs * G;
    static constexpr size_t kInts = (size_t)kCarI * kEnvs * G + (size_t)kRounds * kEnvs * G;
    static constexpr size_t kBytes = kDoubles * sizeof(double) + kInts * sizeof(int);
};

enum { P_PX = 0, P_PY, P_RX, P_VY, P_VX0, P_VX1 };
enum { I_LANE_OLD = 0, I_LANE, I_RECTX, I_RECTY };

template <int G>
struct GrpSm {
    using L = GrpLayout<G>;
    double *car, *ob, *rew;
    int *ci, *obx;
    int e;  // environment within the block
    __device__ __forceinline__ GrpSm(double *base, int e_) : e(e_)
    {
        car = base;
        ob = car + (size_t)L::kCarD * L::kEnvs * G;
        rew = ob + (size_t)O_NF * L::kRounds * L::kEnvs * G;
        ci = reinterpret_cast<int *>(rew + (size_t)L::kEnvs * G);
        obx = ci + (size_t)L::kCarI * L::kEnvs * G;
    }
    __device__ __forceinline__ double &c(int f, int a) const { return car[((size_t)f * L::kEnvs + e) * G + a]; }
    __device__ __forceinline__ int &i(int f, int a) const { return ci[((size_t)f * L::kEnvs + e) * G + a]; }
    __device__ __forceinline__ double &r(int a) const { return rew[(size_t)e * G + a]; }
    // obstacle slot k lives with lane k % G, round k / G
    __device__ __forceinline__ double &o(int f, int k) const { return ob[(((size_t)f * L::kRounds + k / G) * L::kEnvs + e) * G + (k % G)]; }
    __device__ __forceinline__ int &ox(int k) const { return obx[((size_t)(k / G) * L::kEnvs + e) * G + (k % G)]; }
};

struct GrpCar {  // the lane's own car, registers
    double px, py, rx, vx, vy, acc, steer, cruise;
    uint32_t b;
};

__device__ __forceinline__ int car_lane(uint32_t b) { return (int)((b >> HW_BITS_LANE_SHIFT) & 3u); }
__device__ __forceinline__ uint32_t with_lane(uint32_t b, int lane) { return (b & ~(3u << HW_BITS_LANE_SHIFT)) | ((uint32_t)lane << HW_BITS_LANE_SHIFT); }

// start grid: simple_base.py:32-43 ("X" formation, three obstacles behind the cars); counters of the episode cleared
template <int G>
__device__ __forceinline__ void grp_reset_env(MaHead &h, GrpCar &c, const GrpSm<G> &sm, const int g, const bool is_car)
{
    if (is_car) {
        const double cx = (g < 2) ? 10.0 : (g == 2) ? 7.5 : 5.0;
        const int cl = (g == 0 || g == 3) ? 1 : (g == 2) ? 2 : 3;
        c.px = cx; c.py = lane_centre(cl - 1); c.rx = cx; c.vx = 0.0; c.vy = 0.0; c.acc = 0.0; c.steer = 0.0; c.cruise = 0.0;
        c.b = (uint32_t)cl << HW_BITS_LANE_SHIFT;
    }
    if (g < 3) {
        const double ox = (g == 0) ? -20.0 : (g == 1) ? -25.0 : -40.0, ov = (g == 0) ? 7.0 : (g == 1) ? 5.0 : 6.0;
        sm.o(O_PX, g) = ox; sm.o(O_RX, g) = ox; sm.o(O_VX, g) = ov; sm.o(O_IV, g) = ov;
    }
    h.nobs = 3; h.olane = 1u | (2u << 2) | (3u << 4); h.ofresh = 7u;
    h.newest[0] = 0; h.newest[1] = 1; h.newest[2] = 2;
    h.tick = 0; h.leader = 0; h.nc_obs = 0; h.nc_agent = 0; h.ep_len = 0; h.ep_ret = 0.0; h.flags = 0;
}

// what the other lanes read of this lane's car; `cur` = velocity buffer in use
template <int G>
__device__ __forceinline__ void grp_publish(const GrpCar &c, const GrpSm<G> &sm, const int g, const int cur)
{
    sm.c(P_PX, g) = c.px; sm.c(P_PY, g) = c.py; sm.c(P_RX, g) = c.rx; sm.c(P_VY, g) = c.vy; sm.c(P_VX0 + cur, g) = c.vx;
    sm.i(I_LANE, g) = car_lane(c.b);
    sm.i(I_RECTX, g) = rect_x(c.px); sm.i(I_RECTY, g) = rect_y(c.py);
}

__device__ __forceinline__ double dinf() { return __longlong_as_double(0x7FF0000000000000ll); }

// one world.act() followed by Scenario.rewards(); `reward` is the lane's own, done bits are or-ed into `done`.
// AT = number of cars when known at compile time (5), 0 = read it from the state.  Lanes g >= A run the car code on a
// dummy car (their slots in shared memory exist and are never read) so that the group stays converged; only memory
// traffic to global memory and the counters are masked.
template <int G, int AT, bool CONT, bool EXACT>
__device__ __forceinline__ void grp_tick(MaHead &h, GrpCar &c, const GrpSm<G> &sm, int &cur, const int g, const int A_rt, const int K,
                                         const unsigned gmask, const MaConsts &C, const bool inf_obs,
                                         const uint32_t amask, const uint32_t apair, const int dlane,
                                         const double a0_5dt, const double a1_30dt, const uint64_t seed, const uint64_t env_id,
                                         double &reward, uint32_t &done)
{
    using LY = GrpLayout<G>;
    constexpr int CS = LY::kEnvs * G;  // stride between fields
    const int A = AT ? AT : A_rt;
    const double dt = C.dt;
    const bool is_car = g < A;
    double *pc = sm.car + (size_t)sm.e * G;
    int *pi = sm.ci + (size_t)sm.e * G;
    h.tick += 1;
    if (h.tick >= (uint32_t)C.timeout_tick) done |= (1u << A) - 1u;  // highway_core.py:276-277

    // ---- actions for every car in id order (highway_core.py:280-281, :110-188) ----
    if (CONT) {
        c.acc = py_clamp(dadd(c.acc, a0_5dt), -5.0, 5.0);
        c.steer = py_clamp(dadd(c.steer, a1_30dt), -30.0, 30.0);
    } else {
        // amask = the mode bit the action asks for; it fires when not yet set.  The two modes of a pair exclude each
        // other, so setting the bit and clearing its partner is the identity when nothing fires.
        uint32_t b = c.b;
        int lane = car_lane(b);
        pi[I_LANE_OLD * CS + g] = lane;
        const uint32_t fire = amask & ~b;
        b = (b & ~apair) | amask;
        lane = (fire & (kBitLeft | kBitRight)) ? min(max(lane + dlane, 1), 3) : lane;
        c.acc = py_clamp(c.acc, -5.0, 5.0);
        c.steer = py_clamp(c.steer, -30.0, 30.0);
        c.b = with_lane(b, lane);
        pi[I_LANE * CS + g] = lane;
        __syncwarp(gmask);
        if (fire & kBitDoMaint) {
            // all_obstacles iterates the cars (id order) and then the obstacles (insertion order): nearest strictly
            // ahead in the lane, first wins ties; cars before this one have already switched lanes, cars after it have not
            double bpx = dinf(), bvx = c.vx;
#pragma unroll
            for (int j = 0; j < kMaxA; ++j) {
                if (AT ? (j >= AT) : false) break;
                const int lj = (j < g) ? pi[I_LANE * CS + j] : pi[I_LANE_OLD * CS + j];
                const double opx = pc[P_PX * CS + j], ovx = pc[(P_VX0 + cur) * CS + j];
                const bool take = (j < A) & (j != g) & (lj == lane) & (opx > c.px) & (opx < bpx);
                dmov_if(take, bpx, opx); dmov_if(take, bvx, ovx);
            }
            for (int k = 0; k < (int)h.nobs; ++k) {
                const double opx = sm.o(O_PX, k), ovx = sm.o(O_VX, k);
                const bool take = (ob_lane(h, k) == lane) & (opx > c.px) & (opx < bpx);
                dmov_if(take, bpx, opx); dmov_if(take, bvx, ovx);
            }
            c.cruise = bvx;
        }
    }

    // ---- car updates in id order against last tick's leader (agent.py:79-165) ----
    const int L = (int)h.leader;
    const int nxt = cur ^ 1;
    double ang = 0.0;
    {
        uint32_t b = c.b;
        if (!CONT) {
            const bool m_acc = (b & kBitDoAcc) != 0, m_maint = (b & kBitDoMaint) != 0;  // never both
            double acc_a = dadd(c.acc, dt);
            dset_if(c.acc < 0.0, acc_a, 0.0);
            const double vc = ceil(c.vx), cc = ceil(c.cruise);
            const bool snap = m_maint && vc == cc;
            const double acc_m = __hiloint2double(snap ? 0 : ((vc <= cc) ? 0x40140000 : (int)0xC0140000), 0);  // 0.0, 5.0, -5.0
            dmov_if_nz(b & kBitDoAcc, c.acc, acc_a);
            dmov_if_nz(b & kBitDoMaint, c.acc, acc_m);
            dmov_if(snap, c.vx, c.cruise);
            if (m_acc && c.acc == 5.0) b &= ~kBitDoAcc;
            if (snap) b &= ~kBitDoMaint;
        }
        c.vx = py_clamp(dadd(c.vx, dmul(c.acc, dt)), 0.0, 20.0);
        if (!CONT) {
            const int lane = car_lane(b);
            const double ty = __hiloint2double(lane == 1 ? 0x3FF40000 : (lane == 2 ? 0x400E0000 : 0x40190000), 0);  // 1.25, 3.75, 6.25
            const bool m_left = (b & kBitLeft) != 0, m_right = (b & kBitRight) != 0;  // never both
            double inc = 0.0;
            dmov_if(m_left, inc, C.k30dt); dmov_if(m_right, inc, -C.k30dt);
            c.steer = dadd(c.steer, inc);  // + 0.0 is the identity on every value steer can hold
            const bool reached = (m_left && c.py <= ty) || (m_right && c.py >= ty);
            dset_if(reached, c.steer, 0.0);
            if (reached) b &= ~(kBitLeft | kBitRight);
        }
        if (c.steer != 0.0) ang = angular_velocity<EXACT>(c.vx, c.steer);
        if (!CONT) c.vy = ang;  // agent.py:144: velocity.y = angular_velocity (discrete update only)
        c.b = b;
        pc[(P_VX0 + nxt) * CS + g] = c.vx;  // the leader's velocity "as it is now" for the cars after it
    }
    __syncwarp(gmask);
    {
        const double vdt = dmul(c.vx, dt);
        double px = dadd(c.px, vdt);
        double py = dadd(c.py, dmul(c.vy, dt));
        if (CONT) py = dsub(py, dmul(ang, dt));
        else      py = dsub(py, dmul(dmul(dmul(ang, kRad2Deg), dt), dt));
        const double lv = pc[(P_VX0 + ((L < g) ? nxt : cur)) * CS + L];
        px = dsub(px, dmul(lv, dt));
        dset_if(g == L, px, 10.0);
        dset_if(7.0 < py, py, 7.0);   // if y < 1: y = 1 elif y > 6: y = min(y, 7)
        dset_if(py < 1.0, py, 1.0);
        c.rx = dadd(c.rx, vdt);
        c.px = px; c.py = py;
        if (CONT) {
            const double d0 = fabs(dsub(py, 1.25)), d1 = fabs(dsub(py, 3.75)), d2 = fabs(dsub(py, 6.25));
            c.b = with_lane(c.b, (d0 <= d1 && d0 <= d2) ? 1 : (d1 <= d2) ? 2 : 3);
        }
        grp_publish<G>(c, sm, g, nxt);
    }
    cur = nxt;
    __syncwarp(gmask);

    // ---- obstacle updates (agent.py:223-260): yield to the nearest car ahead in the lane, by raw x ----
    const double lvdt = dmul(pc[(P_VX0 + cur) * CS + L], dt);
    for (int k = g; k < (int)h.nobs; k += G) {
        const int lane = ob_lane(h, k);
        const double orx = sm.o(O_RX, k), iv = sm.o(O_IV, k);
        double bd = dinf(), bv = iv;
#pragma unroll
        for (int j = 0; j < kMaxA; ++j) {
            if (AT ? (j >= AT) : false) break;
            const double crx = pc[P_RX * CS + j], cvx = pc[(P_VX0 + cur) * CS + j];
            const double d = dsub(crx, orx);
            const bool take = (j < A) & (pi[I_LANE * CS + j] == lane) & (orx < crx) & (d < bd);  // first wins ties
            dmov_if(take, bd, d); dmov_if(take, bv, cvx);
        }
        double v = py_clamp(iv, -20.0, 20.0);  // velocity.x = max(-20, min(init_vx, 20)) when nobody is ahead
        dmov_if(bd < dinf(), v, iv);            // else min(init_vx, that car's vx)
        dmov_if(bv < iv, v, bv);
        const double odt = dmul(v, dt);
        const double opx = dsub(dadd(sm.o(O_PX, k), odt), lvdt);
        sm.o(O_VX, k) = v; sm.o(O_PX, k) = opx; sm.o(O_RX, k) = dadd(orx, odt);
        sm.ox(k) = rect_x(opx);
    }
    h.ofresh = 0;  // every live obstacle has now written its rect

    // ---- leader / last election: stable sort by x, descending (highway_core.py:294-297) ----
    int lead = 0, last = 0;
    {
        double plead = pc[P_PX * CS], plast = plead;
#pragma unroll
        for (int j = 1; j < kMaxA; ++j) {
            if (AT ? (j >= AT) : false) break;
            const double p = pc[P_PX * CS + j];
            const bool up = (j < A) & (p > plead), dn = (j < A) & (p <= plast);  // first of the maxima, last of the minima
            lead = up ? j : lead; dmov_if(up, plead, p);
            last = dn ? j : last; dmov_if(dn, plast, p);
        }
    }
    h.leader = (uint32_t)lead;
    __syncwarp(gmask);

    if (inf_obs) {
        // ---- spawn (highway_core.py:300-323) and retire (:327-342): rare, so detected by all lanes and carried out by lane 0 ----
        bool want = false;
#pragma unroll
        for (int lane = 0; lane < 3; ++lane) {
            const uint32_t s = h.newest[lane];
            want = want | ((s < 31u) & (sm.o(O_PX, (int)(s & 15u)) < -2.375));
        }
        const double thr = dsub(dsub(-2.375, pc[P_PX * CS + lead]), 2.0);
        const double last_rx = pc[P_RX * CS + last];
        for (int k = g; k < (int)h.nobs; k += G) want = want | (dsub(sm.o(O_RX, k), last_rx) < thr);
        if (__any_sync(gmask, want)) {
            if (g == 0) {
                for (int lane = 0; lane < 3; ++lane) {
                    const uint32_t s = h.newest[lane];
                    if (s >= 31u) continue;  // the lane's newest obstacle was retired: the reference never spawns there again
                    if (sm.o(O_PX, (int)s) < -2.375) {
                        double u1, u2;
                        spawn_uniforms(seed, env_id, h.spawn_ctr, u1, u2);
                        h.spawn_ctr += 1;
                        const double x = dadd(70.0, dmul(30.0, u1)), v = dadd(5.0, dmul(2.0, u2));
                        const double ovx = sm.o(O_VX, (int)s);
                        const double m = (ovx < v) ? ovx : v;  // min(rand_vel_x, obstacle.velocity.x): the old obstacle stays, slowed down
                        sm.o(O_VX, (int)s) = m; sm.o(O_IV, (int)s) = m;
                        if ((int)h.nobs < K) {
                            const int nk = (int)h.nobs;
                            sm.o(O_PX, nk) = x; sm.o(O_RX, nk) = dadd(pc[P_RX * CS + lead], dsub(x, pc[P_PX * CS + lead]));
                            sm.o(O_VX, nk) = v; sm.o(O_IV, nk) = v;
                            h.olane = (h.olane & ~(3u << (2 * nk))) | ((uint32_t)(lane + 1) << (2 * nk));
                            h.ofresh |= 1u << nk;
                            h.newest[lane] = (uint32_t)nk;
                            h.nobs += 1;
                        } else {
                            h.flags |= 1u;  // out of slots: the spawn is dropped (tests require this never to happen)
                        }
                    }
                }
                // order preserving compaction of what the last car has cleared
                int w = 0;
                uint32_t nl = 0, nf = 0, nn0 = 31u, nn1 = 31u, nn2 = 31u;
                for (int k = 0; k < (int)h.nobs; ++k) {
                    if (dsub(sm.o(O_RX, k), last_rx) < thr) continue;
                    if (w != k) {
                        sm.o(O_PX, w) = sm.o(O_PX, k); sm.o(O_RX, w) = sm.o(O_RX, k); sm.o(O_VX, w) = sm.o(O_VX, k); sm.o(O_IV, w) = sm.o(O_IV, k);
                        sm.ox(w) = sm.ox(k);
                    }
                    nl |= (uint32_t)ob_lane(h, k) << (2 * w);
                    nf |= ((h.ofresh >> k) & 1u) << w;
                    if (h.newest[0] == (uint32_t)k) nn0 = (uint32_t)w;
                    if (h.newest[1] == (uint32_t)k) nn1 = (uint32_t)w;
                    if (h.newest[2] == (uint32_t)k) nn2 = (uint32_t)w;
                    ++w;
                }
                h.nobs = (uint32_t)w; h.olane = nl; h.ofresh = nf; h.newest[0] = nn0; h.newest[1] = nn1; h.newest[2] = nn2;
            }
            __syncwarp(gmask);
            const int src = (int)(threadIdx.x & 31u) - g;  // lane 0 of the group
            const uint32_t packed = h.nobs | (h.newest[0] << 8) | (h.newest[1] << 13) | (h.newest[2] << 18) | (h.flags << 24);
            const uint32_t p = __shfl_sync(gmask, packed, src);
            h.nobs = p & 31u; h.newest[0] = (p >> 8) & 31u; h.newest[1] = (p >> 13) & 31u; h.newest[2] = (p >> 18) & 31u; h.flags = p >> 24;
            h.olane = __shfl_sync(gmask, h.olane, src);
            h.ofresh = __shfl_sync(gmask, h.ofresh, src);
            h.spawn_ctr = __shfl_sync(gmask, h.spawn_ctr, src);
        }
    }

    // ---- Scenario.rewards -> check_collisions, after the updates and spawns of this tick (highway_core.py:401-422) ----
    int no = 0, na = 0;
    {
        const int ax = pi[I_RECTX * CS + g], ay = pi[I_RECTY * CS + g];
        for (int k = 0; k < (int)h.nobs; ++k) {
            const bool fresh = (h.ofresh >> k) & 1u;  // spawned in this tick: still pygame's default rect (0, 0, 76, 38)
            const int bx = fresh ? 0 : sm.ox(k);
            const int by = fresh ? 0 : (80 * ob_lane(h, k) - 59);  // 21 + 80 * (lane - 1)
            no += rects_overlap(ax, ay, bx, by) ? 1 : 0;
        }
#pragma unroll
        for (int j = 0; j < kMaxA; ++j) {
            if (AT ? (j >= AT) : false) break;
            na += ((j < A) & (j != g) & rects_overlap(ax, ay, pi[I_RECTX * CS + j], pi[I_RECTY * CS + j])) ? 1 : 0;
        }
        no = is_car ? no : 0; na = is_car ? na : 0;
        reward = dsub(ddiv_const(c.vx, 20.0), 1.0);
        dset_if(no + na > 0, reward, -250.0);
    }
    h.nc_obs += (uint32_t)__reduce_add_sync(gmask, no);
    h.nc_agent += (uint32_t)__reduce_add_sync(gmask, na);
    const unsigned hit = __ballot_sync(gmask, no + na > 0);
    done |= (hit >> ((threadIdx.x & 31u) - g)) & ((1u << A) - 1u);
}

// Scenario.observations (simple_base.py:82-154) + MultiAgentEnv._get_obs_norm (highway_env.py:139-149) of the lane's car
template <int G>
__device__ __forceinline__ void grp_observe(MaHead &h, const GrpCar &c, const GrpSm<G> &sm, const int cur, const int g, const int A,
                                            float *__restrict__ ob /* this car's row of 4A+14 floats */)
{
    const int S = A + 2;
    float2 *o2 = reinterpret_cast<float2 *>(ob);  // rows are 8-byte aligned: (4A+14)*4 bytes is a multiple of 8
    o2[0] = make_float2(norm_f64(c.acc, -5.0, 5.0), norm_f64(c.steer, -30.0, 30.0));
    o2[1] = make_float2(norm_f64(c.rx, 0.0, 1200.0), norm_f64(c.py, 0.0, 8.0));
    o2[2 + S] = make_float2(norm_f64(c.vx, -20.0, 20.0), norm_f64(c.vy, -20.0, 20.0));
    // other cars: slot j - (j > a)
    for (int j = 0; j < A; ++j) {
        if (j == g) continue;
        const int s = j - (j > g ? 1 : 0);
        o2[2 + s] = make_float2(norm_f64(dsub(sm.c(P_RX, j), c.rx), -60.0, 60.0), norm_f64(dsub(sm.c(P_PY, j), c.py), -8.0, 8.0));
        o2[3 + S + s] = make_float2(norm_f64(dsub(sm.c(P_VX0 + cur, j), c.vx), -20.0, 20.0), norm_f64(dsub(sm.c(P_VY, j), c.vy), -20.0, 20.0));
    }
    // per lane the obstacle nearest in raw x (first wins ties, insertion order): slot A-1+lane
    int near0 = -1, near1 = -1, near2 = -1;
    double nd0 = 0.0, nd1 = 0.0, nd2 = 0.0;
    for (int k = 0; k < (int)h.nobs; ++k) {
        const int l = ob_lane(h, k);
        const double d = fabs(dsub(c.rx, sm.o(O_RX, k)));
        if (l == 1) { if (near0 < 0 || d < nd0) { near0 = k; nd0 = d; } }
        else if (l == 2) { if (near1 < 0 || d < nd1) { near1 = k; nd1 = d; } }
        else { if (near2 < 0 || d < nd2) { near2 = k; nd2 = d; } }
    }
#pragma unroll
    for (int l = 0; l < 3; ++l) {
        const int s = A - 1 + l;
        const int k = (l == 0) ? near0 : (l == 1) ? near1 : near2;
        if (k < 0) {  // the reference raises here; flag it and emit the centre of the range
            h.flags |= 2u;
            o2[2 + s] = make_float2(0.5f, 0.5f); o2[3 + S + s] = make_float2(0.5f, 0.5f);
            continue;
        }
        o2[2 + s] = make_float2(norm_f64(dsub(sm.o(O_RX, k), c.rx), -60.0, 60.0), norm_f64(dsub(lane_centre(l), c.py), -8.0, 8.0));
        o2[3 + S + s] = make_float2(norm_f64(dsub(sm.o(O_VX, k), c.vx), -20.0, 20.0), norm_f64(dsub(0.0, c.vy), -20.0, 20.0));
    }
}

template <int G>
__device__ __forceinline__ void grp_load(MaHead &h, GrpCar &c, const GrpSm<G> &sm, const MaPtrs &S, const int g, const bool is_car,
                                         const int64_t i, const int64_t n)
{
    ma_load_head(h, S, i, n);
    if (is_car) {
        const float *p = S.cars + (int64_t)(g * 9) * n + i;
        c.px = (double)p[0]; c.py = (double)p[n]; c.rx = (double)p[2 * n]; c.vx = (double)p[3 * n]; c.vy = (double)p[4 * n];
        c.acc = (double)p[5 * n]; c.steer = (double)p[6 * n]; c.cruise = (double)p[7 * n];
        c.b = __float_as_uint(p[8 * n]);
    }
    for (int k = g; k < (int)h.nobs; k += G) {
        const float4 o = S.obst[(int64_t)k * n + i];
        sm.o(O_PX, k) = o.x; sm.o(O_RX, k) = o.y; sm.o(O_VX, k) = o.z; sm.o(O_IV, k) = o.w;
        sm.ox(k) = rect_x((double)o.x);
    }
}

template <int G>
__device__ __forceinline__ void grp_store(const MaHead &h, const GrpCar &c, const GrpSm<G> &sm, const MaPtrs &S, const int g,
                                          const bool is_car, const int64_t i, const int64_t n)
{
    if (g == 0) ma_store_head(h, S, i, n);
    if (is_car) {
        float *p = S.cars + (int64_t)(g * 9) * n + i;
        p[0] = (float)c.px; p[n] = (float)c.py; p[2 * n] = (float)c.rx; p[3 * n] = (float)c.vx; p[4 * n] = (float)c.vy;
        p[5 * n] = (float)c.acc; p[6 * n] = (float)c.steer; p[7 * n] = (float)c.cruise;
        p[8 * n] = __uint_as_float(c.b);
    }
    for (int k = g; k < (int)h.nobs; k += G)
        S.obst[(int64_t)k * n + i] = make_float4((float)sm.o(O_PX, k), (float)sm.o(O_RX, k), (float)sm.o(O_VX, k), (float)sm.o(O_IV, k));
}

template <int G>
struct GrpIds {
    int g, e;
    int64_t i;
    bool live;
    unsigned gmask;
    __device__ __forceinline__ GrpIds(int64_t n)
    {
        using L = GrpLayout<G>;
        const int lane_id = threadIdx.x & 31, warp = threadIdx.x >> 5;
        const int ew = lane_id / G;
        g = lane_id - ew * G;
        e = warp * L::kEnvsPerWarp + ew;
        i = (int64_t)blockIdx.x * L::kEnvs + e;
        live = ew < L::kEnvsPerWarp && i < n;
        gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (ew * G));
    }
};

template <int G, int AT, bool CONT, bool EXACT>
__global__ void __launch_bounds__(kGrpWarps * 32, 3)
ma_grp_step_kernel(MaPtrs S, const void *__restrict__ actions, float *__restrict__ obs, float *__restrict__ rew,
                   uint8_t *__restrict__ done_out, double *__restrict__ term, unsigned long long *__restrict__ stats,
                   const MaConsts C, const uint64_t seed, const uint64_t env_id0, const int64_t n, const int flags)
{
    extern __shared__ double sm_dyn[];
    const GrpIds<G> id(n);
    const int g = id.g, A = AT ? AT : S.A, K = S.K, D = 4 * A + 14;
    const unsigned gmask = id.gmask;
    const int64_t i = id.i;
    long long acc[7] = {0, 0, 0, 0, 0, 0, 0};  // finished episode of this group (lane 0), summed per warp at the end
    if (id.live) {
        const GrpSm<G> sm(sm_dyn, id.e);
        const bool is_car = g < A;
        const bool inf_obs = !(flags & HW_FLAG_NO_INF_OBS);
        const bool autoreset = !(flags & HW_FLAG_NO_AUTORESET);
        const bool shared_reward = (flags & HW_FLAG_SHARED_REWARD) != 0;
        MaHead h;
        GrpCar c;
        c.px = c.py = c.rx = c.vx = c.vy = c.acc = c.steer = c.cruise = 0.0; c.b = 0;
        grp_load<G>(h, c, sm, S, g, is_car, i, n);
        int cur = 0;
        if (is_car) grp_publish<G>(c, sm, g, cur);
        uint32_t amask = 0, apair = 0;
        int dlane = 0;
        double a0_5dt = 0.0, a1_30dt = 0.0;
        if (is_car) {
            if (CONT) {
                const float2 v = reinterpret_cast<const float2 *>(actions)[i * A + g];
                a0_5dt = dmul(dmul((double)v.x, 5.0), C.dt);
                a1_30dt = dmul(dmul((double)v.y, 30.0), C.dt);
            } else {
                const int act = reinterpret_cast<const int *>(actions)[i * A + g];
                amask = (act == HW_ACT_LEFT) ? kBitLeft : (act == HW_ACT_RIGHT) ? kBitRight
                      : (act == HW_ACT_ACCELERATE) ? kBitDoAcc : (act == HW_ACT_MAINTAIN) ? kBitDoMaint : 0u;
                apair = (amask & (kBitLeft | kBitRight)) ? (kBitLeft | kBitRight) : amask ? (kBitDoAcc | kBitDoMaint) : 0u;
                dlane = (act == HW_ACT_RIGHT) ? 1 : (act == HW_ACT_LEFT) ? -1 : 0;
            }
        }
        __syncwarp(gmask);

        double r = 0.0;
        uint32_t done = 0;
        for (int t = 0; t < C.ticks_per_step; ++t) {
            grp_tick<G, AT, CONT, EXACT>(h, c, sm, cur, g, A, K, gmask, C, inf_obs, amask, apair, dlane, a0_5dt, a1_30dt, seed, env_id0 + (uint64_t)i, r, done);
            if (done) break;  // highway_env.py:85-86
        }
        // rewards of the step: summed in id order for shared_reward (highway_env.py:92-94) and for Monitor
        if (is_car) sm.r(g) = r;
        __syncwarp(gmask);
        double rsum = 0.0;
        for (int a = 0; a < A; ++a) { const double ra = sm.r(a); rsum = dadd(rsum, ra); }
        if (shared_reward) {
            r = rsum;
            for (int a = 0; a < A; ++a) h.ep_ret = dadd(h.ep_ret, rsum);
        } else {
            for (int a = 0; a < A; ++a) h.ep_ret = dadd(h.ep_ret, sm.r(a));  // Monitor sums the python floats of every agent and step
        }
        if (is_car) {
            rew[i * A + g] = (float)r;
            done_out[i * A + g] = (done >> g) & 1u;
        }
        h.ep_len += 1;

        if (done) {
            if (g == 0) {
                if (term) {
                    double *tm = term + 5 * i;
                    tm[0] = (double)h.tick; tm[1] = (double)h.nc_obs; tm[2] = (double)h.nc_agent; tm[3] = h.ep_ret; tm[4] = (double)h.ep_len;
                }
                acc[0] = 1; acc[1] = h.ep_len; acc[2] = __double2ll_rn(h.ep_ret * 1000.0); acc[3] = h.nc_obs;
                acc[4] = (h.tick >= (uint32_t)C.timeout_tick) ? 1 : 0; acc[5] = h.nc_agent; acc[6] = h.flags & 1u;
            }
            if (autoreset) {
                const uint32_t keep = h.flags;
                __syncwarp(gmask);  // every lane has finished reading the terminal state
                grp_reset_env<G>(h, c, sm, g, is_car);
                h.flags = keep;  // sticky diagnostics survive the reset
                cur = 0;
                if (is_car) grp_publish<G>(c, sm, g, cur);
                for (int k = g; k < 3; k += G) sm.ox(k) = 0;
                __syncwarp(gmask);
            }
        }
        if (is_car) grp_observe<G>(h, c, sm, cur, g, A, obs + (i * A + g) * (int64_t)D);
        // a lane without an obstacle to observe is the same finding for every car; the header is stored by lane 0 (always a car)
        grp_store<G>(h, c, sm, S, g, is_car, i, n);
    }
    if (stats) {
        if (__any_sync(0xffffffffu, acc[0] != 0)) {
#pragma unroll
            for (int j = 0; j < 7; ++j) acc[j] = warp_sum(acc[j]);
            if ((threadIdx.x & 31) == 0) {
#pragma unroll
                for (int j = 0; j < 7; ++j) atomicAdd(&stats[j], (unsigned long long)acc[j]);
            }
        }
    }
}

int ma_grp_launch_step(const HwMaState *st, const void *actions, const HwMaOut *out, const MaConsts &C_,
                       uint64_t seed, uint64_t env_id0, int64_t n, int flags, cudaStream_t stream)
{
    constexpr int G = 8;
    using L = GrpLayout<G>;
    MaPtrs S;
    S.cars = st->cars; S.obst = reinterpret_cast<float4 *>(st->obst); S.head = st->head; S.A = st->num_agents; S.K = st->num_slots;
    const size_t smem = L::kBytes;
    const unsigned grid = (unsigned)((n + L::kEnvs - 1) / L::kEnvs);
    cudaError_t e = cudaSuccess;
#define HW_GRP_LAUNCH_A(AT, C, E) do { \
        if (smem > 48 * 1024) e = cudaFuncSetAttribute(ma_grp_step_kernel<G, AT, C, E>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        if (e != cudaSuccess) return (int)e; \
        ma_grp_step_kernel<G, AT, C, E><<<grid, kGrpWarps * 32, smem, stream>>>(S, actions, out->obs, out->rew, out->done, out->term, out->stats, C_, seed, env_id0, n, flags); \
    } while (0)
#define HW_GRP_LAUNCH(C, E) do { if (S.A == kMaxA) HW_GRP_LAUNCH_A(kMaxA, C, E); else HW_GRP_LAUNCH_A(0, C, E); } while (0)
    const bool exact = (flags & HW_FLAG_EXACT_STEERING) != 0;
    if (flags & HW_FLAG_CONTINUOUS) { if (exact) HW_GRP_LAUNCH(true, true); else HW_GRP_LAUNCH(true, false); }
    else                            { if (exact) HW_GRP_LAUNCH(false, true); else HW_GRP_LAUNCH(false, false); }
#undef HW_GRP_LAUNCH_A
#undef HW_GRP_LAUNCH
    return (int)cudaGetLastError();
}

}  // namespace hw
