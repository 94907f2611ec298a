// Multi-agent highway step for sm_100a, G lanes per environment (G = 5: six environments per warp, two idle lanes).
//
// Same path as hw_ma_tpe.cu (thread per environment; entry points in hw_ma.cu): MultiAgentEnv.step, highway_core.HighwaySimulator.act,
// agent.Car.update_d/update_c, agent.Obstacle.update, check_collisions, Scenario.rewards/observations, the
// auto-reset of DummyVecMAEnv and Monitor's episode record (file:line citations at each stage below).
//
// Why this layout (historically the second one).  The first kernel ran a thread per environment with ALL of the environment's fp64
// working set (5 cars x 8 fields + 16 obstacle slots x 4 fields = 832 B) in shared memory, which capped an SM at 8 warps and left it
// bound by the latency of one thread walking cars and obstacles one after the other (issue slots 28 % busy).
// Here lane g of an environment's group IS car g (its state stays in registers for the whole step) and owns the
// obstacle slots g, g + G, ...; what the other lanes need of a car (position, velocity, lane, rect) is published
// in shared memory, and the stages of a tick are separated by full-mask __syncwarp (control flow is uniform over the
// warp; a group that has finished its step just goes inactive).  The order-
// dependent semantics of the reference survive the parallel evaluation because every cross-car read is of a
// value whose version (before / after this tick's update) is known statically:
//   * actions run for every car in id order (highway_core.py:280-281): maintain() of car a sees the lanes cars j < a
//     have just switched to and the old lanes of cars j > a; a lane change depends only on the car's own action
//     and flags, so all cars switch first (old and new lane are both published) and then search;
//   * cars update in id order and subtract the leader's velocity "as it is now" (agent.py:147-151): its new value
//     for cars behind it in id order, its old value for cars before it; velocities are double-buffered;
//   * the leader election (moved ahead of the obstacle updates: it only reads the cars), the obstacle updates, spawn /
//     retire and the collision test each read only finished values.
// Scalars of the environment header (tick, leader, counters, slot bookkeeping) are replicated in the registers
// of all G lanes and evolve identically; the serial slot bookkeeping of a spawn or a retire runs on lane 0 and is
// broadcast with shuffles.
#include "hw_ma_common.cuh"

namespace hw {

#ifndef HW_MA_WARPS
#define HW_MA_WARPS 8
#endif
constexpr int kGrpWarps = HW_MA_WARPS;  // warps per block

template <int G>
struct GrpLayout {
    static constexpr int kEnvsPerWarp = 32 / G;
    static constexpr int kEnvs = kEnvsPerWarp * kGrpWarps;  // environments per block
    // published car fields (doubles): px, py, rx, vy, vx[2] (double buffered); ints: lane (old << 8 | new), rect {x, y}
    static constexpr int kCarD = 6, kCarI = 3;
    // obstacle slots of one environment are contiguous (slot k of field f at ob[f][e * kObStride + k]): the loops that
    // walk the list index with one add; the odd stride keeps the groups of a warp on different banks
    static constexpr int kObStride = kMaxK + 1;
    static constexpr size_t kDoubles = (size_t)(kCarD + 1) * kEnvs * G + (size_t)O_NF * kEnvs * kObStride;
    static constexpr size_t kInts = (size_t)kCarI * kEnvs * G + (size_t)2 * kEnvs * kObStride;  // + one rect {x, y} per obstacle slot
    static constexpr size_t kBytes = kDoubles * sizeof(double) + kInts * sizeof(int);
};

enum { P_PX = 0, P_PY, P_RX, P_VY, P_VX0, P_VX1 };
enum { I_LANE = 0, I_RECT };  // I_RECT: int2 per car, two ints wide

template <int G>
struct GrpSm {
    using L = GrpLayout<G>;
    static constexpr int OS = L::kEnvs * L::kObStride;  // stride between obstacle fields
    double *car, *ob, *rew;
    int *ci;
    int2 *obr;  // rect {x, y} per obstacle slot ({0, 0} = pygame's default rect of a slot spawned in this tick)
    int e;  // environment within the block
    __device__ __forceinline__ GrpSm(double *base, int e_) : e(e_)
    {
        car = base;
        double *ob0 = car + (size_t)L::kCarD * L::kEnvs * G;
        rew = ob0 + (size_t)O_NF * OS;
        ci = reinterpret_cast<int *>(rew + (size_t)L::kEnvs * G);
        int2 *obr0 = reinterpret_cast<int2 *>(ci + (size_t)L::kCarI * L::kEnvs * G);
        ob = ob0 + e * L::kObStride;
        obr = obr0 + e * L::kObStride;
    }
    __device__ __forceinline__ double &c(int f, int a) const { return car[((size_t)f * L::kEnvs + e) * G + a]; }
    __device__ __forceinline__ int &i(int f, int a) const { return ci[((size_t)f * L::kEnvs + e) * G + a]; }
    __device__ __forceinline__ double &r(int a) const { return rew[(size_t)e * G + a]; }
    __device__ __forceinline__ double &o(int f, int k) const { return ob[f * OS + k]; }
    __device__ __forceinline__ int2 &orect(int k) const { return obr[k]; }
    __device__ __forceinline__ int2 &crect(int a) const { return reinterpret_cast<int2 *>(ci + (size_t)I_RECT * L::kEnvs * G)[(size_t)e * G + a]; }
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
        sm.orect(g) = make_int2(0, 0);  // not yet updated: pygame's default rect
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
    sm.crect(g) = make_int2(rect_x(c.px), rect_y(c.py));
}

__device__ __forceinline__ double dinf() { return __longlong_as_double(0x7FF0000000000000ll); }

// State of one lane during a step, and the stages of a tick (one world.act() followed by Scenario.rewards()).
//
// Control flow is uniform over the WARP: every lane walks through the same stages and the same full-mask
// __syncwarp / __ballot_sync calls, and a group that has finished its step (or has no environment) simply has
// `active` false.  (Group-mask collectives compile to a divergence check plus a slow path that the hardware takes
// whenever the warp is not fully converged - always, with G = 5, where lanes 30 and 31 idle.)
// AT = number of cars when known at compile time (5), 0 = read it from the state.  Lanes g >= A run the car code on a
// dummy car (their slots in shared memory exist and are never read) so that a group stays converged; only traffic
// to global memory and the collision counters are masked.
template <int G, int AT, bool CONT, bool EXACT>
struct GrpStep {
    using LY = GrpLayout<G>;
    static constexpr int CS = LY::kEnvs * G;  // stride between fields of the published arrays
    MaHead h;
    GrpCar c;
    GrpSm<G> sm;
    double *pc;
    int *pi;
    int g, base, A, K, cur;
    bool active;
    uint32_t done;
    int no_acc, na_acc;   // this car's collisions during the step
    double ang;
    bool last_hit;        // this car collided in the last tick that ran (its reward is then -250)
    int lead, last;
    uint32_t amask, apair, fire;
    int dlane;
    double a0_5dt, a1_30dt;

    __device__ __forceinline__ GrpStep(double *smem, int e, int g_, int base_, int A_, int K_)
        : sm(smem, e), g(g_), base(base_), A(AT ? AT : A_), K(K_), cur(0), active(false), done(0), no_acc(0), na_acc(0),
          ang(0.0), last_hit(false), lead(0), last(0), amask(0), apair(0), fire(0), dlane(0), a0_5dt(0.0), a1_30dt(0.0)
    {
        pc = sm.car + (size_t)e * G;
        pi = sm.ci + (size_t)e * G;
    }

    __device__ __forceinline__ unsigned group_bits(unsigned ballot) const { return (ballot >> base) & ((1u << G) - 1u); }

    // ---- actions for every car in id order (highway_core.py:280-281, :110-188), part 1: everything but maintain()'s search ----
    __device__ __forceinline__ void actions(const MaConsts &C)
    {
        h.tick += 1;
        if (h.tick >= (uint32_t)C.timeout_tick) done |= (1u << A) - 1u;  // highway_core.py:276-277
        if (CONT) {
            c.acc = py_clamp(dadd(c.acc, a0_5dt), -5.0, 5.0);
            c.steer = py_clamp(dadd(c.steer, a1_30dt), -30.0, 30.0);
            return;
        }
        // amask = the mode bit the action asks for; it fires when not yet set.  The two modes of a pair exclude each
        // other, so setting the bit and clearing its partner is the identity when nothing fires.
        uint32_t b = c.b;
        int lane = car_lane(b);
        const int lane_old = lane;
        fire = amask & ~b;
        b = (b & ~apair) | amask;
        lane = (fire & (kBitLeft | kBitRight)) ? min(max(lane + dlane, 1), 3) : lane;
        c.acc = py_clamp(c.acc, -5.0, 5.0);
        c.steer = py_clamp(c.steer, -30.0, 30.0);
        c.b = with_lane(b, lane);
        pi[I_LANE * CS + g] = lane | (lane_old << 8);  // both are needed by the other cars' maintain()
    }

    // part 2, for the cars whose MAINTAIN fired: all_obstacles iterates the cars (id order) and then the obstacles
    // (insertion order): nearest strictly ahead in the lane, first wins ties; cars before this one have already
    // switched lanes, cars after it have not
    __device__ __forceinline__ void maintain_search()
    {
        const int lane = car_lane(c.b);
        double bpx = dinf(), bvx = c.vx;
#pragma unroll
        for (int j = 0; j < kMaxA; ++j) {
            if (AT ? (j >= AT) : false) break;
            const int lw = pi[I_LANE * CS + j];
            const int lj = (j < g) ? (lw & 3) : (lw >> 8);
            const double opx = pc[P_PX * CS + j], ovx = pc[(P_VX0 + cur) * CS + j];
            const bool take = (j < A) & (j != g) & (lj == lane) & (opx > c.px) & (opx < bpx);
            dmov_if(take, bpx, opx); dmov_if(take, bvx, ovx);
        }
        // only the slots that drive in this lane (usually one): 2-bit lane fields of `olane` equal to `lane`, in slot order
        const uint32_t diff = h.olane ^ ((uint32_t)lane * 0x55555555u);
        uint32_t same = ~(diff | (diff >> 1)) & 0x55555555u & live_mask(h.nobs);
        while (same) {
            const int k = (__ffs((int)same) - 1) >> 1;
            same &= same - 1u;
            const double opx = sm.o(O_PX, k), ovx = sm.o(O_VX, k);
            const bool take = (opx > c.px) & (opx < bpx);
            dmov_if(take, bpx, opx); dmov_if(take, bvx, ovx);
        }
        c.cruise = bvx;
    }

    // ---- car updates in id order against last tick's leader (agent.py:79-165), part 1: velocity and steering ----
    __device__ __forceinline__ void update_velocity(const MaConsts &C)
    {
        const double dt = C.dt;
        uint32_t b = c.b;
        if (!CONT) {
            pi[I_LANE * CS + g] = car_lane(b);  // from here on only the current lane is read
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
        ang = 0.0;
        if (c.steer != 0.0) ang = angular_velocity<EXACT>(c.vx, c.steer);
        if (!CONT) c.vy = ang;  // agent.py:144: velocity.y = angular_velocity (discrete update only)
        c.b = b;
        pc[(P_VX0 + (cur ^ 1)) * CS + g] = c.vx;  // the leader's velocity "as it is now" for the cars after it
    }

    // part 2: positions, against the leader's new velocity for the cars after it in id order, its old one for those before
    __device__ __forceinline__ void update_position(const MaConsts &C)
    {
        const double dt = C.dt;
        const int L = (int)h.leader, nxt = cur ^ 1;
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
        cur = nxt;
    }

    // ---- leader / last election: stable sort of the cars by x, descending (highway_core.py:294-297), then the obstacle
    //      updates (agent.py:223-260): yield to the nearest car ahead in the lane, by raw x.  In the reference the
    //      election follows the obstacle updates, but it only reads the cars, which are final by now; doing it first
    //      lets the lane that updates a slot also test it for a spawn (highway_core.py:300-323: the lane's newest
    //      obstacle has fallen behind x = -2.375) or a retire (:327-342), both rare, while the values are in registers.
    //      Returns whether this lane saw either condition.
    __device__ __forceinline__ bool obstacles_and_leader(const MaConsts &C, const bool inf_obs, double &thr, double &last_rx)
    {
        const double dt = C.dt;
        const double lvdt = dmul(pc[(P_VX0 + cur) * CS + (int)h.leader], dt);  // last tick's leader, this tick's velocity
        lead = 0; last = 0;
        double plead = pc[P_PX * CS], plast = plead;
#pragma unroll
        for (int j = 1; j < kMaxA; ++j) {
            if (AT ? (j >= AT) : false) break;
            const double p = pc[P_PX * CS + j];
            const bool up = (j < A) & (p > plead), dn = (j < A) & (p <= plast);  // first of the maxima, last of the minima
            lead = up ? j : lead; dmov_if(up, plead, p);
            last = dn ? j : last; dmov_if(dn, plast, p);
        }
        h.leader = (uint32_t)lead;
        thr = dsub(dsub(-2.375, plead), 2.0);
        last_rx = pc[P_RX * CS + last];
        bool want = false;
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
            const double orx2 = dadd(orx, odt);
            sm.o(O_VX, k) = v; sm.o(O_PX, k) = opx; sm.o(O_RX, k) = orx2;
            sm.orect(k) = make_int2(rect_x(opx), 80 * lane - 59);  // y = 21 + 80 * (lane - 1)
            const uint32_t newest = (lane == 1) ? h.newest[0] : (lane == 2) ? h.newest[1] : h.newest[2];
            want = want | ((newest == (uint32_t)k) & (opx < -2.375)) | (dsub(orx2, last_rx) < thr);
        }
        h.ofresh = 0;  // every live obstacle has now written its rect
        return inf_obs && want;
    }

    // ... and lane 0 of the group does the serial slot bookkeeping
    __device__ __forceinline__ void spawn_retire(const double thr, const double last_rx, const uint64_t seed, const uint64_t env_id)
    {
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
                    sm.orect(nk) = make_int2(0, 0);
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
                sm.orect(w) = sm.orect(k);
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

    // header words lane 0 may have changed, as one 64-bit + two 32-bit values for the broadcast
    __device__ __forceinline__ uint32_t pack_slots() const { return h.nobs | (h.newest[0] << 8) | (h.newest[1] << 13) | (h.newest[2] << 18) | (h.flags << 24); }
    __device__ __forceinline__ void unpack_slots(uint32_t p)
    {
        h.nobs = p & 31u; h.newest[0] = (p >> 8) & 31u; h.newest[1] = (p >> 13) & 31u; h.newest[2] = (p >> 18) & 31u; h.flags = p >> 24;
    }

    // ---- Scenario.rewards -> check_collisions, after the updates and spawns of this tick (highway_core.py:401-422) ----
    __device__ __forceinline__ bool collide()
    {
        // rects are 76 x 38: A overlaps B iff |ax - bx| < 76 and |ay - by| < 38, each as one unsigned compare
        const int2 ar = sm.crect(g);
        const int ax75 = ar.x + 75, ay37 = ar.y + 37;
        int no = 0, na = 0;
        for (int k = 0; k < (int)h.nobs; ++k) {
            const int2 br = sm.orect(k);
            no += ((uint32_t)(ax75 - br.x) <= 150u) & ((uint32_t)(ay37 - br.y) <= 74u);
        }
#pragma unroll
        for (int j = 0; j < kMaxA; ++j) {
            if (AT ? (j >= AT) : false) break;
            const int2 br = sm.crect(j);
            na += (j < A) & (j != g) & ((uint32_t)(ax75 - br.x) <= 150u) & ((uint32_t)(ay37 - br.y) <= 74u);
        }
        const bool is_car = g < A;
        no = is_car ? no : 0; na = is_car ? na : 0;
        no_acc += no; na_acc += na;
        last_hit = no + na > 0;
        return last_hit;
    }

    // Scenario.observations (simple_base.py:82-154) + MultiAgentEnv._get_obs_norm (highway_env.py:139-149) of the lane's car
    __device__ __forceinline__ void observe(float *__restrict__ ob /* this car's row of 4A+14 floats */)
    {
        const int S = A + 2;
        float2 *o2 = reinterpret_cast<float2 *>(ob);  // rows are 8-byte aligned: (4A+14)*4 bytes is a multiple of 8
        {   // the car itself; the normalisations go side by side in batches (ma_norm_n) so that their dependency chains overlap
            const double raw[6] = {c.acc, c.steer, c.rx, c.py, c.vx, c.vy};
            float o[6];
            ma_norm_n<6, 0, 1, 2, 3, 6, 6>(raw, o);
            o2[0] = make_float2(o[0], o[1]);
            o2[1] = make_float2(o[2], o[3]);
            o2[2 + S] = make_float2(o[4], o[5]);
        }
        // other cars: car j goes to slot j - (j > a).  Walked by slot (j = s + (s >= a)) so that the lanes of a warp stay
        // converged: skipping "j == own index" inside a loop over j makes every lane skip a different iteration
#pragma unroll
        for (int sl = 0; sl < kMaxA - 1; ++sl) {
            if (AT ? (sl >= AT - 1) : false) break;
            if (sl >= A - 1) break;
            const int j = sl + (sl >= g ? 1 : 0);
            const double raw[4] = {dsub(pc[P_RX * CS + j], c.rx), dsub(pc[P_PY * CS + j], c.py), dsub(pc[(P_VX0 + cur) * CS + j], c.vx), dsub(pc[P_VY * CS + j], c.vy)};
            float o[4];
            ma_norm_n<4, 4, 5, 6, 6>(raw, o);
            o2[2 + sl] = make_float2(o[0], o[1]);
            o2[3 + S + sl] = make_float2(o[2], o[3]);
        }
        // per lane the obstacle nearest in raw x (first wins ties, insertion order): slot A-1+lane.  The slots of a lane are
        // picked out of the 2-bit lane fields of `olane`, so each slot is visited once, by its own lane's search.
        const uint32_t live = live_mask(h.nobs);
#pragma unroll
        for (int l = 0; l < 3; ++l) {
            const int s = A - 1 + l;
            const uint32_t diff = h.olane ^ ((uint32_t)(l + 1) * 0x55555555u);
            uint32_t same = ~(diff | (diff >> 1)) & 0x55555555u & live;
            double nd = dinf(), orx = 0.0, ovx = 0.0;
            while (same) {
                const int k = (__ffs((int)same) - 1) >> 1;
                same &= same - 1u;
                const double krx = sm.o(O_RX, k), kvx = sm.o(O_VX, k);
                const double d = fabs(dsub(c.rx, krx));
                const bool take = d < nd;
                dmov_if(take, nd, d); dmov_if(take, orx, krx); dmov_if(take, ovx, kvx);
            }
            const double raw[4] = {dsub(orx, c.rx), dsub(lane_centre(l), c.py), dsub(ovx, c.vx), dsub(0.0, c.vy)};
            float o[4];
            ma_norm_n<4, 4, 5, 6, 6>(raw, o);
            float2 pos = make_float2(o[0], o[1]);
            float2 vel = make_float2(o[2], o[3]);
            if (!(nd < dinf())) {  // the reference raises here; flag it and emit the centre of the range
                h.flags |= 2u;
                pos = make_float2(0.5f, 0.5f); vel = make_float2(0.5f, 0.5f);
            }
            o2[2 + s] = pos; o2[3 + S + s] = vel;
        }
    }

    __device__ __forceinline__ void load(const MaPtrs &S, const int64_t i, const int64_t n)
    {
        ma_load_head(h, S, i, n);
        // is_done persists until reset() (highway_core.py:239, :420): with HW_FLAG_NO_AUTORESET a car that finished in an
        // earlier step is still done, and the step then ends after its first tick (highway_env.py:85-86)
        done = (h.flags >> 8) & 31u; h.flags &= 0xFFu;
        c.px = c.py = c.rx = c.vx = c.vy = c.acc = c.steer = c.cruise = 0.0; c.b = 0;
        if (g < A) {
            const float *p = S.cars + (int64_t)(g * 9) * n + i;
            c.px = (double)p[0]; c.py = (double)p[n]; c.rx = (double)p[2 * n]; c.vx = (double)p[3 * n]; c.vy = (double)p[4 * n];
            c.acc = (double)p[5 * n]; c.steer = (double)p[6 * n]; c.cruise = (double)p[7 * n];
            c.b = __float_as_uint(p[8 * n]);
        }
        for (int k = g; k < (int)h.nobs; k += G) {
            const float4 o = S.obst[(int64_t)k * n + i];
            sm.o(O_PX, k) = o.x; sm.o(O_RX, k) = o.y; sm.o(O_VX, k) = o.z; sm.o(O_IV, k) = o.w;
            sm.orect(k) = ((h.ofresh >> k) & 1u) ? make_int2(0, 0) : make_int2(rect_x((double)o.x), 80 * ob_lane(h, k) - 59);
        }
        cur = 0;
        grp_publish<G>(c, sm, g, cur);
    }

    __device__ __forceinline__ void store(const MaPtrs &S, const int64_t i, const int64_t n) const
    {
        if (g == 0) ma_store_head(h, S, i, n);
        if (g < A) {
            float *p = S.cars + (int64_t)(g * 9) * n + i;
            p[0] = (float)c.px; p[n] = (float)c.py; p[2 * n] = (float)c.rx; p[3 * n] = (float)c.vx; p[4 * n] = (float)c.vy;
            p[5 * n] = (float)c.acc; p[6 * n] = (float)c.steer; p[7 * n] = (float)c.cruise;
            p[8 * n] = __uint_as_float(c.b);
        }
        for (int k = g; k < (int)h.nobs; k += G)
            S.obst[(int64_t)k * n + i] = make_float4((float)sm.o(O_PX, k), (float)sm.o(O_RX, k), (float)sm.o(O_VX, k), (float)sm.o(O_IV, k));
    }
};

#ifndef HW_MA_MIN_BLOCKS
#define HW_MA_MIN_BLOCKS 3
#endif

template <int G, int AT, bool CONT, bool EXACT>
__global__ void __launch_bounds__(kGrpWarps * 32, HW_MA_MIN_BLOCKS)
ma_grp_step_kernel(MaPtrs S, const void *__restrict__ actions, float *__restrict__ obs, float *__restrict__ rew,
                   uint8_t *__restrict__ done_out, double *__restrict__ term, unsigned long long *__restrict__ stats,
                   const MaConsts C, const uint64_t seed, const uint64_t env_id0, const int64_t n, const int flags)
{
    using LY = GrpLayout<G>;
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ double sm_dyn[];
    const int lane_id = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ew = lane_id / G;                      // environment within the warp (ew == kEnvsPerWarp: idle lane)
    const int g = lane_id - ew * G;
    const bool has_env = ew < LY::kEnvsPerWarp;
    const int e = has_env ? warp * LY::kEnvsPerWarp + ew : 0;
    const int64_t i = (int64_t)blockIdx.x * LY::kEnvs + e;
    const bool live = has_env && i < n;
    GrpStep<G, AT, CONT, EXACT> st(sm_dyn, e, g, ew * G, S.A, S.K);
    const int A = st.A, D = 4 * A + 14;
    const bool is_car = g < A;
    const bool inf_obs = !(flags & HW_FLAG_NO_INF_OBS);
    const bool autoreset = !(flags & HW_FLAG_NO_AUTORESET);
    const bool shared_reward = (flags & HW_FLAG_SHARED_REWARD) != 0;
    const uint64_t env_id = env_id0 + (uint64_t)i;
    long long acc[7] = {0, 0, 0, 0, 0, 0, 0};  // finished episode of this group (lane 0), summed per warp at the end

    if (live) {
        st.load(S, i, n);
        if (is_car) {
            if (CONT) {
                const float2 v = reinterpret_cast<const float2 *>(actions)[i * A + g];
                st.a0_5dt = dmul(dmul((double)v.x, 5.0), C.dt);
                st.a1_30dt = dmul(dmul((double)v.y, 30.0), C.dt);
            } else {
                const int act = reinterpret_cast<const int *>(actions)[i * A + g];
                st.amask = (act == HW_ACT_LEFT) ? kBitLeft : (act == HW_ACT_RIGHT) ? kBitRight
                         : (act == HW_ACT_ACCELERATE) ? kBitDoAcc : (act == HW_ACT_MAINTAIN) ? kBitDoMaint : 0u;
                st.apair = (st.amask & (kBitLeft | kBitRight)) ? (kBitLeft | kBitRight) : st.amask ? (kBitDoAcc | kBitDoMaint) : 0u;
                st.dlane = (act == HW_ACT_RIGHT) ? 1 : (act == HW_ACT_LEFT) ? -1 : 0;
            }
        }
    }
    st.active = live;
    __syncwarp();

    for (int t = 0; t < C.ticks_per_step; ++t) {
        if (__ballot_sync(FULL, st.active) == 0u) break;
        if (st.active) st.actions(C);
        if (!CONT) {
            __syncwarp();
            if (st.active && (st.fire & kBitDoMaint)) st.maintain_search();
            __syncwarp();  // the searches read the packed old / new lanes, which update_velocity overwrites
        }
        if (st.active) st.update_velocity(C);
        __syncwarp();
        if (st.active) st.update_position(C);
        __syncwarp();
        double thr = 0.0, last_rx = 0.0;
        const bool want = st.active && st.obstacles_and_leader(C, inf_obs, thr, last_rx);
        __syncwarp();
        if (inf_obs) {
            const unsigned wanted = __ballot_sync(FULL, want);
            if (wanted) {  // some group of the warp has an obstacle to spawn or to retire
                const bool mine = st.group_bits(wanted) != 0u;
                if (mine && g == 0) st.spawn_retire(thr, last_rx, seed, env_id);
                __syncwarp();
                const uint32_t p = __shfl_sync(FULL, st.pack_slots(), ew * G);
                const uint32_t ol = __shfl_sync(FULL, st.h.olane, ew * G), of = __shfl_sync(FULL, st.h.ofresh, ew * G);
                const uint32_t sc = __shfl_sync(FULL, st.h.spawn_ctr, ew * G);
                if (mine) { st.unpack_slots(p); st.h.olane = ol; st.h.ofresh = of; st.h.spawn_ctr = sc; }
            }
        }
        const bool hit = st.active && st.collide();
        const unsigned hits = st.group_bits(__ballot_sync(FULL, hit));
        if (st.active) {
            st.done |= hits & ((1u << A) - 1u);
            if (st.done) st.active = false;  // highway_env.py:85-86: the step ends with the first tick that finishes a car
        }
    }

    // rewards of the step: summed in id order for shared_reward (highway_env.py:92-94) and for Monitor; collision counters
    // Scenario.rewards of the last tick that ran (simple_base.py:65-80): -250 for a car that collided in it, else v.x / 20 - 1
    double my_reward = dsub(ddiv_const(st.c.vx, 20.0), 1.0);
    dset_if(st.last_hit, my_reward, -250.0);
    if (live) { st.sm.r(g) = my_reward; st.sm.crect(g) = make_int2(st.no_acc, st.na_acc); }  // the rect slot is free now
    __syncwarp();
    if (live) {
        double r = my_reward, rsum = 0.0;
        for (int a = 0; a < A; ++a) {
            rsum = dadd(rsum, st.sm.r(a));
            const int2 cnt = st.sm.crect(a);
            st.h.nc_obs += (uint32_t)cnt.x; st.h.nc_agent += (uint32_t)cnt.y;
        }
        if (shared_reward) {
            r = rsum;
            for (int a = 0; a < A; ++a) st.h.ep_ret = dadd(st.h.ep_ret, rsum);
        } else {
            for (int a = 0; a < A; ++a) st.h.ep_ret = dadd(st.h.ep_ret, st.sm.r(a));  // Monitor sums the python floats of every agent and step
        }
        if (is_car) {
            rew[i * A + g] = (float)r;
            done_out[i * A + g] = (st.done >> g) & 1u;
        }
        st.h.ep_len += 1;
        if (st.done && g == 0) {
            if (term) {
                double *tm = term + 5 * i;
                tm[0] = (double)st.h.tick; tm[1] = (double)st.h.nc_obs; tm[2] = (double)st.h.nc_agent; tm[3] = st.h.ep_ret; tm[4] = (double)st.h.ep_len;
            }
            acc[0] = 1; acc[1] = st.h.ep_len; acc[2] = __double2ll_rn(st.h.ep_ret * 1000.0); acc[3] = st.h.nc_obs;
            acc[4] = (st.h.tick >= (uint32_t)C.timeout_tick) ? 1 : 0; acc[5] = st.h.nc_agent; acc[6] = st.h.flags & 1u;
        }
    }
    __syncwarp();  // every lane has finished reading the terminal state
    if (live && st.done && autoreset) {
        const uint32_t keep = st.h.flags;
        grp_reset_env<G>(st.h, st.c, st.sm, g, is_car);
        st.h.flags = keep;  // sticky diagnostics survive the reset
        st.cur = 0;
        grp_publish<G>(st.c, st.sm, g, st.cur);
    }
    __syncwarp();
    if (live) {
        if (is_car) st.observe(obs + (i * A + g) * (int64_t)D);
        // a lane without an obstacle to observe is the same finding for every car; the header is stored by lane 0 (always a car)
        if (!(st.done && autoreset)) st.h.flags |= st.done << 8;
        st.store(S, i, n);
    }
    if (stats) {
        if (__any_sync(FULL, acc[0] != 0)) {
#pragma unroll
            for (int j = 0; j < 7; ++j) acc[j] = warp_sum(acc[j]);
            if (lane_id == 0) {
#pragma unroll
                for (int j = 0; j < 7; ++j) atomicAdd(&stats[j], (unsigned long long)acc[j]);
            }
        }
    }
}

int ma_grp_launch_step(const HwMaState *st, const void *actions, const HwMaOut *out, const MaConsts &C_,
                       uint64_t seed, uint64_t env_id0, int64_t n, int flags, cudaStream_t stream)
{
#ifndef HW_MA_GROUP
#define HW_MA_GROUP 5
#endif
    constexpr int G = HW_MA_GROUP;
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
