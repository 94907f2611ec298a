// Multi-agent highway step for sm_100a, one THREAD per environment with the cars in registers.
//
// Same path as hw_ma_grp.cu (MultiAgentEnv.step, highway_core.HighwaySimulator.act, agent.Car.update_d/update_c,
// agent.Obstacle.update, check_collisions, Scenario.rewards/observations, DummyVecMAEnv's auto-reset and Monitor's
// episode record; file:line citations at each stage below).
//
// Why a third layout.  With a lane per car (hw_ma_grp.cu) every lane walks through the whole tick - the per-environment
// stages (leader election, obstacle list, spawn / retire, the tick's control flow) are executed by all five lanes of a
// group, and the pairwise stages pay shared-memory traffic and warp synchronisation for values that sit in another lane:
// 7 700 warp-instructions per six env-steps, i.e. 3.5x the thread-instructions per car of the single-agent kernel.  Here
// a thread owns an environment: the A cars are A-fold unrolled register arrays (A is a template parameter, all indexing
// is static), so the order-dependent semantics of the reference - cars act and move in id order and see each other's
// partially updated state - are simply the program order, nothing is published, nothing is synchronised, and the
// per-environment work is done once.  Only the obstacle list (up to 16 slots x 4 doubles, dynamically indexed) lives in
// shared memory, in a warp-private [slot][field][lane] layout (conflict free); after the ticks the same region stages the
// observation rows so that they leave as contiguous 8-byte-per-lane streams (an environment's A rows of 4A+14 floats are
// contiguous in the output, and so are the 32 environments of a warp).  State loads and stores are coalesced by
// construction (structure of arrays, consecutive threads = consecutive environments).
// Cost of the layout: ~165 registers per thread -> 12 warps per SM (3 per scheduler), so the kernel lives on the instruction-level
// parallelism inside a thread.  Measured on B200 (scripts/micro/fp64_latency.cu): DADD / DMUL / DFMA 8 cycles, DSETP -> its predicate
// ~20, F2F / F2I / FRND ~18-21, LDS 29; a "compare -> conditional move -> compare" link therefore costs ~28 cycles, and ptxas
// serialises independent chains when it is short of registers.  Hence: stages are written over all cars side by side (searches,
// clamps on the original value, the 170 normalisations of the observation in batches), and rare work (spawn / retire, the exact
// collision count behind an EXACT packed-rect filter) stays out of line.
#include "hw_ma_common.cuh"

namespace hw {

#ifndef HW_TPE_BLOCK
#define HW_TPE_BLOCK 128
#endif
#ifndef HW_TPE_MIN_BLOCKS
#define HW_TPE_MIN_BLOCKS 3
#endif
constexpr int kTpeBlock = HW_TPE_BLOCK;
#ifndef HW_TPE_SYNC
#define HW_TPE_SYNC 1   // 1: block-synchronous ticks (252 us at 262 144 envs), 2: same loop without the barrier (301), 0: break out of the loop
#endif
#ifndef HW_TPE_SLOTS
#define HW_TPE_SLOTS kMaxK
#endif
constexpr int kTpeWarpDoubles = HW_TPE_SLOTS * O_NF * 32;  // one warp's obstacle region: 16 KB for 16 slots

// Conditional moves of doubles: the shared helpers (one predicated DMUL / predicated moves).  selp.f64 - a pair of 32-bit selects on
// the ALU pipe - measured slower here as well (268 vs 252 us, profiles/r4_variants.md).
__device__ __forceinline__ void tmov_if(bool p, double &x, double y) { dmov_if(p, x, y); }
__device__ __forceinline__ void tmov_if_nz(uint32_t w, double &x, double y) { dmov_if_nz(w, x, y); }
__device__ __forceinline__ void tset_if(bool p, double &x, double c) { dset_if(p, x, c); }
__device__ __forceinline__ double tclamp(double x, const double lo, const double hi) { return py_clamp(x, lo, hi); }

struct TpeOb {  // this thread's column of the warp-private obstacle region
    double *base;
    __device__ __forceinline__ double &operator()(int k, int f) const { return base[(k * O_NF + f) * 32]; }
};

template <int A>
struct TpeCars {
    double px[A], py[A], rx[A], vx[A], vy[A], acc[A], steer[A], cruise[A];
    uint32_t b[A];  // mode bits; the lane field of the packed word is kept apart in ln[] while the state is in registers
    int ln[A];
};

__device__ __forceinline__ int tpe_lane(uint32_t b) { return (int)((b >> HW_BITS_LANE_SHIFT) & 3u); }
__device__ __forceinline__ uint32_t tpe_with_lane(uint32_t b, int lane) { return (b & ~(3u << HW_BITS_LANE_SHIFT)) | ((uint32_t)lane << HW_BITS_LANE_SHIFT); }
__device__ __forceinline__ double tpe_inf() { return __longlong_as_double(0x7FF0000000000000ll); }
__device__ __forceinline__ int tpe_overlap(int ax75, int ay37, int bx, int by)
{
    // rects are 76 x 38: A overlaps B iff |ax - bx| < 76 and |ay - by| < 38, each as one unsigned compare
    return (int)(((uint32_t)(ax75 - bx) <= 150u) & ((uint32_t)(ay37 - by) <= 74u));
}

// start grid: simple_base.py:32-43 ("X" formation, three obstacles behind the cars); counters of the episode cleared
template <int A>
__device__ __forceinline__ void tpe_reset(MaHead &h, TpeCars<A> &c, const TpeOb ob)
{
#pragma unroll
    for (int a = 0; a < A; ++a) {
        const double cx = ma_start_x(a);
        const int cl = ma_start_lane(a);
        c.px[a] = cx; c.py[a] = lane_centre(cl - 1); c.rx[a] = cx; c.vx[a] = 0.0; c.vy[a] = 0.0; c.acc[a] = 0.0; c.steer[a] = 0.0; c.cruise[a] = 0.0;
        c.b[a] = 0; c.ln[a] = cl;
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const double ox = (k == 0) ? -20.0 : (k == 1) ? -25.0 : -40.0, ov = (k == 0) ? 7.0 : (k == 1) ? 5.0 : 6.0;
        ob(k, O_PX) = ox; ob(k, O_RX) = ox; ob(k, O_VX) = ov; ob(k, O_IV) = ov;
    }
    h.nobs = 3; h.olane = 1u | (2u << 2) | (3u << 4); h.ofresh = 7u;
    h.newest[0] = 0; h.newest[1] = 1; h.newest[2] = 2;
    h.tick = 0; h.leader = 0; h.nc_obs = 0; h.nc_agent = 0; h.ep_len = 0; h.ep_ret = 0.0; h.flags = 0;
}

// spawn (the lane's newest obstacle has fallen behind x = -2.375: highway_core.py:300-323, the old obstacle stays and is slowed
// down) and retire (what the last car has cleared: :327-342, order preserving compaction).  Rare; kept out of line.
struct TpeSlots {  // the header words spawn / retire change, passed and returned by value (registers) across the out-of-line call
    uint32_t nobs, olane, ofresh, n0, n1, n2, spawn_ctr, flags;
};

#ifndef HW_TPE_RARE_ATTR
#define HW_TPE_RARE_ATTR __noinline__
#endif
static __device__ HW_TPE_RARE_ATTR TpeSlots tpe_spawn_retire(TpeSlots h, double *ob_base, const int K, const double plead, const double lead_rx,
                                                            const double last_rx, const uint64_t seed, const uint64_t env_id)
{
    const TpeOb ob{ob_base};
    uint32_t newest[3] = {h.n0, h.n1, h.n2};
#pragma unroll
    for (int lane = 0; lane < 3; ++lane) {
        const uint32_t s = newest[lane];
        if (s >= 31u) continue;  // the lane's newest obstacle was retired: the reference never spawns there again
        if (ob((int)s, O_PX) < -2.375) {
            double u1, u2;
            spawn_uniforms(seed, env_id, h.spawn_ctr, u1, u2);
            h.spawn_ctr += 1;
            const double x = dadd(70.0, dmul(30.0, u1)), v = dadd(5.0, dmul(2.0, u2));
            const double ovx = ob((int)s, O_VX);
            const double m = (ovx < v) ? ovx : v;  // min(rand_vel_x, obstacle.velocity.x)
            ob((int)s, O_VX) = m; ob((int)s, O_IV) = m;
            if ((int)h.nobs < K) {
                const int nk = (int)h.nobs;
                ob(nk, O_PX) = x; ob(nk, O_RX) = dadd(lead_rx, dsub(x, plead));
                ob(nk, O_VX) = v; ob(nk, O_IV) = v;
                h.olane = (h.olane & ~(3u << (2 * nk))) | ((uint32_t)(lane + 1) << (2 * nk));
                h.ofresh |= 1u << nk;
                newest[lane] = (uint32_t)nk;
                h.nobs += 1;
            } else {
                h.flags |= 1u;  // out of slots: the spawn is dropped (tests require this never to happen)
            }
        }
    }
    const double thr = dsub(dsub(-2.375, plead), 2.0);
    int w = 0;
    uint32_t nl = 0, nf = 0, nn0 = 31u, nn1 = 31u, nn2 = 31u;
    for (int k = 0; k < (int)h.nobs; ++k) {
        if (dsub(ob(k, O_RX), last_rx) < thr) continue;
        if (w != k) { ob(w, O_PX) = ob(k, O_PX); ob(w, O_RX) = ob(k, O_RX); ob(w, O_VX) = ob(k, O_VX); ob(w, O_IV) = ob(k, O_IV); }
        nl |= ((h.olane >> (2 * k)) & 3u) << (2 * w);
        nf |= ((h.ofresh >> k) & 1u) << w;
        if (newest[0] == (uint32_t)k) nn0 = (uint32_t)w;
        if (newest[1] == (uint32_t)k) nn1 = (uint32_t)w;
        if (newest[2] == (uint32_t)k) nn2 = (uint32_t)w;
        ++w;
    }
    h.nobs = (uint32_t)w; h.olane = nl; h.ofresh = nf; h.n0 = nn0; h.n1 = nn1; h.n2 = nn2;
    return h;
}

// exact check_collisions over the current obstacle list (fresh slots carry pygame's default rect) and the other cars: counters,
// and the mask of the cars that collided.  Only reached when the tick's filter saw an overlap - with the exact filter that is a real
// collision, once per episode - so it lives out of line (HW_TPE_COLLIDE_OOL): the cars' rects go in by value, the counts come back
// packed in one word (hit mask | obstacle count << 8 | car count << 16), and the tick body is ~350 instructions shorter.
template <int A> struct TpeRects { int ax75[A], ay37[A]; };
template <int A>
static __device__ __noinline__
uint32_t tpe_collide_count(const TpeRects<A> r, const uint32_t nobs, const uint32_t ofresh, const uint32_t olane, const double *ob_base)
{
    const TpeOb ob{const_cast<double *>(ob_base)};
    uint32_t hit = 0;
    int no = 0, na = 0;
    for (int k = 0; k < (int)nobs; ++k) {
        const bool fresh = ((ofresh >> k) & 1u) != 0u;  // spawned in this tick: still pygame's default rect (0, 0, 76, 38)
        const int bx = fresh ? 0 : rect_x(ob(k, O_PX));
        const int by = fresh ? 0 : 80 * (int)((olane >> (2 * k)) & 3u) - 59;  // y = 21 + 80 * (lane - 1)
#pragma unroll
        for (int a = 0; a < A; ++a) {
            const int o = tpe_overlap(r.ax75[a], r.ay37[a], bx, by);
            no += o; hit |= (uint32_t)o << a;
        }
    }
#pragma unroll
    for (int a = 0; a < A; ++a) {
#pragma unroll
        for (int j = a + 1; j < A; ++j) {
            const int o = tpe_overlap(r.ax75[a], r.ay37[a], r.ax75[j] - 75, r.ay37[j] - 37);
            na += 2 * o; hit |= ((uint32_t)o << a) | ((uint32_t)o << j);
        }
    }
    return hit | ((uint32_t)no << 8) | ((uint32_t)na << 16);  // at most 16 * 8 obstacle and 8 * 7 car overlaps
}

template <int A>
__device__ __forceinline__ uint32_t tpe_collide_exact(MaHead &h, const TpeCars<A> &c, const TpeOb ob)
{
    TpeRects<A> r;
#pragma unroll
    for (int a = 0; a < A; ++a) { r.ax75[a] = rect_x(c.px[a]) + 75; r.ay37[a] = rect_y(c.py[a]) + 37; }
    const uint32_t w = tpe_collide_count<A>(r, h.nobs, h.ofresh, h.olane, ob.base);
    h.nc_obs += (w >> 8) & 0xFFu; h.nc_agent += w >> 16;
    return w & 0xFFu;
}

// The per-tick overlap FILTER works on rects packed into one word, x in the upper and y in the lower 16 bits.  With
// dx = ax - bx + 75 and dy = ay - by + 37, A overlaps B iff 0 <= dx <= 150 and 0 <= dy <= 74, i.e. iff dx <= 150, dy <= 74 AND the same
// two bounds with the roles of A and B exchanged (dx' = 150 - dx, dy' = 74 - dy).  Adding 105 to dx and 53 to dy makes the upper
// bounds 2^k - 1 (255, 127), so T = ((dx + 105) << 16) + (dy + 53) has no bit under the mask 0xFF00FF80 iff both upper bounds hold
// (and the value is not below -105 / -53); T | T' has none iff all four hold: the filter is EXACT.  (One word alone also passes a
// vehicle up to 180 px - 5.6 m - ahead in the same lane, which the start grid's (10, L) / (5, L) pairs are: the exact count then
// ran for 13 of 32 lanes in a third of all ticks - 11 % of the kernel's instructions.)  A negative field borrows into its own high
// bits (and from the x field, which then fails or stays valid - the pair has failed already).  Fields cannot alias while
// |dx| < 65 000 px, i.e. within 2 000 m: an episode moves a vehicle by 1 200 m at most.  min() over the masked words of all pairs
// is zero iff some pair overlaps.
constexpr uint32_t kRectMask = 0xFF00FF80u;
constexpr uint32_t kRectBias = (105u << 16) + 53u;
constexpr uint32_t kRectSize = (75u << 16) + 37u;
// pa = pack(ax + 75, ay + 37) of a car, pb = pack(bx, by) of the other rect: the masked test word
__device__ __forceinline__ uint32_t tpe_filter_word(uint32_t pa, uint32_t pb)
{
    return ((pa - pb + kRectBias) | (pb - pa + (2u * kRectSize + kRectBias))) & kRectMask;
}
__device__ __forceinline__ uint32_t tpe_pack_rect(int x, int y) { return ((uint32_t)x << 16) + (uint32_t)y; }

// one world.act() followed by Scenario.rewards() (highway_core.py:245-399, simple_base.py:65-80); returns the cars that collided
template <int A, bool CONT, bool EXACT>
__device__ __forceinline__ uint32_t tpe_tick(MaHead &h, TpeCars<A> &c, const TpeOb ob, const int K, const MaConsts &C, const double dt,
                                             const bool inf_obs, const uint32_t (&amask)[A], const double (&a0)[A], const double (&a1)[A],
                                             const uint64_t seed, const uint64_t env_id, uint32_t &done)
{
    h.tick += 1;
    if (h.tick >= (uint32_t)C.timeout_tick) done |= (1u << A) - 1u;  // highway_core.py:276-277

    // ---- actions for every car in id order (highway_core.py:280-281, :110-188): maintain() of car a sees the lanes the cars
    //      before it have just switched to and the old lanes of the cars after it - program order ----
    if (CONT) {
#pragma unroll
        for (int a = 0; a < A; ++a) {
            c.acc[a] = tclamp(dadd(c.acc[a], a0[a]), -5.0, 5.0);
            c.steer[a] = tclamp(dadd(c.steer[a], a1[a]), -30.0, 30.0);
        }
    } else {
        // amask = the mode bit the action asks for; it fires when not yet set.  The two modes of a pair exclude each other, so
        // setting the bit and clearing its partner is the identity when nothing fires.  Lane changes first (integers only): the
        // maintain() of car a below wants the NEW lane of the cars before it and the OLD lane of the cars after it (its own lane
        // does not change in a tick in which its MAINTAIN fires).
        uint32_t fire[A], anyfire = 0;
        int ln_old[A];
#pragma unroll
        for (int a = 0; a < A; ++a) {
            const uint32_t am = amask[a];
            const uint32_t apair = am | ((am & (kBitLeft | kBitDoAcc)) << 1) | ((am & (kBitRight | kBitDoMaint)) >> 1);
            fire[a] = am & ~c.b[a];
            anyfire |= fire[a];
            c.b[a] = (c.b[a] & ~apair) | am;
            ln_old[a] = c.ln[a];
            const int dl = (int)((fire[a] >> 19) & 1u) - (int)((fire[a] >> 18) & 1u);  // RIGHT +1, LEFT -1, only when it fires
            c.ln[a] = min(max(c.ln[a] + dl, 1), 3);
            c.acc[a] = tclamp(c.acc[a], -5.0, 5.0);
            c.steer[a] = tclamp(c.steer[a], -30.0, 30.0);
        }
        if (anyfire & kBitDoMaint) {
            // maintain(): cruise = velocity of the nearest vehicle strictly ahead in the lane, else the car's own.  all_obstacles
            // iterates the cars (id order) and then the obstacles (insertion order); the first wins ties.  The searches of the A cars
            // are evaluated side by side - A independent dependency chains, one pass over the obstacle list - and only the cars whose
            // MAINTAIN fired keep the result.  (One search after the other, each behind its own branch, was a chain of
            // compare -> conditional move links of ~28 cycles each that nothing overlapped.)
            double bpx[A], bvx[A];
#pragma unroll
            for (int a = 0; a < A; ++a) { bpx[a] = tpe_inf(); bvx[a] = dcopy(c.vx[a]); }
#pragma unroll
            for (int j = 0; j < A; ++j) {
#pragma unroll
                for (int a = 0; a < A; ++a) {
                    if (j == a) continue;
                    const int lane_j = (j < a) ? c.ln[j] : ln_old[j];
                    const bool take = (lane_j == c.ln[a]) & (c.px[j] > c.px[a]) & (c.px[j] < bpx[a]);
                    tmov_if(take, bpx[a], c.px[j]); tmov_if(take, bvx[a], c.vx[j]);
                }
            }
            for (int k = 0; k < (int)h.nobs; ++k) {
                const int lane_k = ob_lane(h, k);
                const double opx = ob(k, O_PX), ovx = ob(k, O_VX);
#pragma unroll
                for (int a = 0; a < A; ++a) {
                    const bool take = (lane_k == c.ln[a]) & (opx > c.px[a]) & (opx < bpx[a]);
                    tmov_if(take, bpx[a], opx); tmov_if(take, bvx[a], ovx);
                }
            }
#pragma unroll
            for (int a = 0; a < A; ++a) tmov_if_nz(fire[a] & kBitDoMaint, c.cruise[a], bvx[a]);
        }
    }
    // ---- car updates in id order against last tick's leader (agent.py:79-165): a car subtracts the leader's velocity "as it is
    //      now" (agent.py:147-151) - its new value for the cars after it in id order, its old one for those before ----
    // Written as three passes over the cars (velocity and steering, angular velocity, position), each a straight line of A
    // independent dependency chains: the instruction-level parallelism that 12 warps per SM need.
    const int L = (int)h.leader;
    double lv_old = c.vx[0];
#pragma unroll
    for (int j = 1; j < A; ++j) tmov_if(L == j, lv_old, c.vx[j]);
#pragma unroll
    for (int a = 0; a < A; ++a) {
        uint32_t b = c.b[a];
        if (!CONT) {
            const bool m_acc = (b & kBitDoAcc) != 0, m_maint = (b & kBitDoMaint) != 0;  // never both
            double acc_a = dadd(c.acc[a], dt);
            tset_if(c.acc[a] < 0.0, acc_a, 0.0);
            const double vc = ceil(c.vx[a]), cc = ceil(c.cruise[a]);
            const bool snap = m_maint && vc == cc;
            const double acc_m = __hiloint2double(snap ? 0 : ((vc <= cc) ? 0x40140000 : (int)0xC0140000), 0);  // 0.0, 5.0, -5.0
            tmov_if_nz(b & kBitDoAcc, c.acc[a], acc_a);
            tmov_if_nz(b & kBitDoMaint, c.acc[a], acc_m);
            tmov_if(snap, c.vx[a], c.cruise[a]);
            if (m_acc && c.acc[a] == 5.0) b &= ~kBitDoAcc;
            if (snap) b &= ~kBitDoMaint;
        }
        c.vx[a] = tclamp(dadd(c.vx[a], dmul(c.acc[a], dt)), 0.0, 20.0);
        if (!CONT) {
            const double ty = __fma_rn((double)c.ln[a], 2.5, -1.25);  // lane 1, 2, 3: 1.25, 3.75, 6.25 (exact)
            const bool m_left = (b & kBitLeft) != 0, m_right = (b & kBitRight) != 0;  // never both
            double inc = 0.0;
            tmov_if(m_left, inc, C.k30dt); tmov_if(m_right, inc, -C.k30dt);
            c.steer[a] = dadd(c.steer[a], inc);  // + 0.0 is the identity on every value steer can hold
            const bool reached = (m_left && c.py[a] <= ty) || (m_right && c.py[a] >= ty);
            tset_if(reached, c.steer[a], 0.0);
            if (reached) b &= ~(kBitLeft | kBitRight);
        }
        c.b[a] = b;
    }
    double lv = c.vx[0];  // the leader's velocity "as it is now" for the cars after it: its new value
#pragma unroll
    for (int j = 1; j < A; ++j) tmov_if(L == j, lv, c.vx[j]);
    double ang[A];
#pragma unroll
    for (int a = 0; a < A; ++a) {
        // evaluated unconditionally (tan(0) = 0 gives 0 as well; a division by 4 / tan(0) = inf gives 0): no branch between the cars
        ang[a] = angular_velocity<EXACT>(c.vx[a], c.steer[a]);
        tset_if(!(c.steer[a] != 0.0), ang[a], 0.0);
        if (!CONT) c.vy[a] = ang[a];  // agent.py:144: velocity.y = angular_velocity (discrete update only)
    }
#pragma unroll
    for (int a = 0; a < A; ++a) {
        double lva = lv_old;
        tmov_if(L < a, lva, lv);
        const double vdt = dmul(c.vx[a], dt);
        double px = dadd(c.px[a], vdt);
        double py = dadd(c.py[a], dmul(c.vy[a], dt));
        if (CONT) py = dsub(py, dmul(ang[a], dt));
        else      py = dsub(py, dmul(dmul(dmul(ang[a], kRad2Deg), dt), dt));
        px = dsub(px, dmul(lva, dt));
        tset_if(L == a, px, 10.0);
        {   // if y < 1: y = 1 elif y > 6: y = min(y, 7); both tests on the value before either move (they exclude each other)
            const bool over = 7.0 < py, under = py < 1.0;
            tset_if(over, py, 7.0);
            tset_if(under, py, 1.0);
        }
        c.rx[a] = dadd(c.rx[a], vdt);
        c.px[a] = px; c.py[a] = py;
        if (CONT) {
            const double d0 = fabs(dsub(py, 1.25)), d1 = fabs(dsub(py, 3.75)), d2 = fabs(dsub(py, 6.25));
            c.ln[a] = (d0 <= d1 && d0 <= d2) ? 1 : (d1 <= d2) ? 2 : 3;
        }
    }
    const double lvdt = dmul(lv, dt);  // last tick's leader, this tick's velocity

    // ---- leader / last election: stable sort of the cars by x, descending (highway_core.py:294-297).  In the reference it follows
    //      the obstacle updates, but it only reads the cars, which are final by now ----
    int lead = 0;
    double plead = c.px[0], plast = c.px[0], lead_rx = c.rx[0], last_rx = c.rx[0];
#pragma unroll
    for (int j = 1; j < A; ++j) {
        const bool up = c.px[j] > plead, dn = c.px[j] <= plast;  // first of the maxima, last of the minima
        lead = up ? j : lead; tmov_if(up, plead, c.px[j]); tmov_if(up, lead_rx, c.rx[j]);
        tmov_if(dn, plast, c.px[j]); tmov_if(dn, last_rx, c.rx[j]);
    }
    h.leader = (uint32_t)lead;

    // the cars' rects for check_collisions (highway_core.py:401-422), packed for the overlap filter (see kRectMask)
    uint32_t pa[A];
#pragma unroll
    for (int a = 0; a < A; ++a) pa[a] = tpe_pack_rect(rect_x(c.px[a]) + 75, rect_y(c.py[a]) + 37);
    uint32_t nohit = kRectMask;  // min over all rect pairs of the masked test word: 0 = some pair overlaps -> the exact count runs (rare)

    // ---- obstacle updates (agent.py:223-260): yield to the nearest car ahead in the lane, by raw x.  While a slot's new values
    //      are in registers it is tested for a spawn / retire and against the cars' rects (an obstacle that is retired below is
    //      more than 14 m behind the last car, so testing it too cannot add an overlap in a consistent frame; the exact count
    //      walks the final list in any case) ----
    const double thr = dsub(dsub(-2.375, plead), 2.0);
    const uint32_t newest_mask = (1u << h.newest[0]) | (1u << h.newest[1]) | (1u << h.newest[2]);  // 31 = retired: matches no slot
    bool want = false;
    for (int k = 0; k < (int)h.nobs; ++k) {
        const int lane = ob_lane(h, k);
        const double orx = ob(k, O_RX), iv = ob(k, O_IV);
        double bd = tpe_inf(), bv = iv;
#pragma unroll
        for (int j = 0; j < A; ++j) {
            const double d = dsub(c.rx[j], orx);
            const bool take = (c.ln[j] == lane) & (orx < c.rx[j]) & (d < bd);  // first wins ties
            tmov_if(take, bd, d); tmov_if(take, bv, c.vx[j]);
        }
        double v = tclamp(iv, -20.0, 20.0);  // velocity.x = max(-20, min(init_vx, 20)) when nobody is ahead
        tmov_if(bd < tpe_inf(), v, iv);        // else min(init_vx, that car's vx)
        tmov_if(bv < iv, v, bv);
        const double odt = dmul(v, dt);
        const double opx = dsub(dadd(ob(k, O_PX), odt), lvdt);
        const double orx2 = dadd(orx, odt);
        ob(k, O_VX) = v; ob(k, O_PX) = opx; ob(k, O_RX) = orx2;
        want = want | ((((newest_mask >> k) & 1u) != 0u) & (opx < -2.375)) | (dsub(orx2, last_rx) < thr);
        const uint32_t pb = tpe_pack_rect(rect_x(opx), 80 * lane - 59);  // y = 21 + 80 * (lane - 1)
#pragma unroll
        for (int a = 0; a < A; ++a) nohit = min(nohit, tpe_filter_word(pa[a], pb));
    }
    h.ofresh = 0;  // every live obstacle has now written its rect
    if (inf_obs && want) {
        const TpeSlots r = tpe_spawn_retire(TpeSlots{h.nobs, h.olane, h.ofresh, h.newest[0], h.newest[1], h.newest[2], h.spawn_ctr, h.flags},
                                            ob.base, K, plead, lead_rx, last_rx, seed, env_id);
        h.nobs = r.nobs; h.olane = r.olane; h.ofresh = r.ofresh; h.newest[0] = r.n0; h.newest[1] = r.n1; h.newest[2] = r.n2;
        h.spawn_ctr = r.spawn_ctr; h.flags = r.flags;
        if (h.ofresh) {  // a slot spawned in this tick carries pygame's default rect (0, 0, 76, 38) into check_collisions
#pragma unroll
            for (int a = 0; a < A; ++a) nohit = min(nohit, tpe_filter_word(pa[a], 0u));
        }
    }

    // ---- Scenario.rewards -> check_collisions, after the updates and spawns of this tick (highway_core.py:401-422) ----
#pragma unroll
    for (int a = 0; a < A; ++a) {
#pragma unroll
        for (int j = a + 1; j < A; ++j) nohit = min(nohit, tpe_filter_word(pa[a], pa[j] - kRectSize));
    }
    uint32_t hit = 0;
    if (nohit == 0u) hit = tpe_collide_exact<A>(h, c, ob);
    done |= hit;
    return hit;
}

// MultiAgentEnv._get_obs_norm (highway_env.py:139-149) of one value: see norm_f64
// Scenario.observations (simple_base.py:82-154): the part that needs the obstacle list - per car and lane the obstacle nearest in
// raw x (first wins ties, insertion order) as two normalised floats {dx, dvx}; a lane without an obstacle (the reference raises
// there) is flagged and reads 0.5
template <int A>
__device__ __forceinline__ uint32_t tpe_observe_obstacles(MaHead &h, const TpeCars<A> &c, const TpeOb ob, float (&odx)[A][3], float (&odv)[A][3])
{
    const uint32_t live = live_mask(h.nobs);
    uint32_t nolane = 0;
#pragma unroll
    for (int l = 0; l < 3; ++l) {
        const uint32_t diff = h.olane ^ ((uint32_t)(l + 1) * 0x55555555u);
        uint32_t same = ~(diff | (diff >> 1)) & 0x55555555u & live;
        double nd[A], brx[A], bvx[A];
#pragma unroll
        for (int a = 0; a < A; ++a) { nd[a] = tpe_inf(); brx[a] = 0.0; bvx[a] = 0.0; }
        const bool none = same == 0u;
        while (same) {
            const int k = (__ffs((int)same) - 1) >> 1;
            same &= same - 1u;
            const double krx = ob(k, O_RX), kvx = ob(k, O_VX);
#pragma unroll
            for (int a = 0; a < A; ++a) {
                const double d = fabs(dsub(c.rx[a], krx));
                const bool take = d < nd[a];
                tmov_if(take, nd[a], d); tmov_if(take, brx[a], krx); tmov_if(take, bvx[a], kvx);
            }
        }
        if (none) { h.flags |= 2u; nolane |= 1u << l; }
#pragma unroll
        for (int a0 = 0; a0 < A; a0 += 2) {  // two cars = four chains side by side
            if (a0 + 1 < A) {
                const double raw[4] = {dsub(brx[a0], c.rx[a0]), dsub(bvx[a0], c.vx[a0]), dsub(brx[a0 + 1], c.rx[a0 + 1]), dsub(bvx[a0 + 1], c.vx[a0 + 1])};
                float o[4];
                ma_norm_n<4, 4, 6, 4, 6>(raw, o);
                odx[a0][l] = none ? 0.5f : o[0]; odv[a0][l] = none ? 0.5f : o[1];
                odx[a0 + 1][l] = none ? 0.5f : o[2]; odv[a0 + 1][l] = none ? 0.5f : o[3];
            } else {
                const double raw[2] = {dsub(brx[a0], c.rx[a0]), dsub(bvx[a0], c.vx[a0])};
                float o[2];
                ma_norm_n<2, 4, 6>(raw, o);
                odx[a0][l] = none ? 0.5f : o[0]; odv[a0][l] = none ? 0.5f : o[1];
            }
        }
    }
    return nolane;
}

// the 4A+14 floats of car a as 2A+7 float2 (simple_base.py:94-154 + highway_env.py:139-149), written through `put(q, value)`
template <int A, typename Put>
__device__ __forceinline__ void tpe_observe_row(const TpeCars<A> &c, const int a, const float (&odx)[A][3], const float (&odv)[A][3],
                                                const uint32_t nolane, Put put)
{
    constexpr int S = A + 2;
    {   // the car itself: acc, steer, raw x, y, vx, vy
        const double raw[6] = {c.acc[a], c.steer[a], c.rx[a], c.py[a], c.vx[a], c.vy[a]};
        float o[6];
        ma_norm_n<6, 0, 1, 2, 3, 6, 6>(raw, o);
        put(0, make_float2(o[0], o[1]));
        put(1, make_float2(o[2], o[3]));
        put(2 + S, make_float2(o[4], o[5]));
    }
#pragma unroll
    for (int j = 0; j < A; ++j) {  // other cars: car j goes to slot j - (j > a)
        if (j == a) continue;
        const int sl = j - (j > a ? 1 : 0);
        const double raw[4] = {dsub(c.rx[j], c.rx[a]), dsub(c.py[j], c.py[a]), dsub(c.vx[j], c.vx[a]), dsub(c.vy[j], c.vy[a])};
        float o[4];
        ma_norm_n<4, 4, 5, 6, 6>(raw, o);
        put(2 + sl, make_float2(o[0], o[1]));
        put(3 + S + sl, make_float2(o[2], o[3]));
    }
    {   // obstacles: slot A - 1 + lane; the lateral offsets to the three lane centres and the (common) relative lateral velocity
        const double raw[4] = {dsub(lane_centre(0), c.py[a]), dsub(lane_centre(1), c.py[a]), dsub(lane_centre(2), c.py[a]), dsub(0.0, c.vy[a])};
        float o[4];
        ma_norm_n<4, 5, 5, 5, 6>(raw, o);
#pragma unroll
        for (int l = 0; l < 3; ++l) {
            const int s = A - 1 + l;
            const bool none = ((nolane >> l) & 1u) != 0u;
            put(2 + s, make_float2(odx[a][l], none ? 0.5f : o[l]));
            put(3 + S + s, make_float2(odv[a][l], none ? 0.5f : o[3]));
        }
    }
}

template <int A, bool CONT, bool EXACT>
__global__ void __launch_bounds__(kTpeBlock, (A <= kMaxA) ? HW_TPE_MIN_BLOCKS : 2)  // the 6..8-car extension needs ~250 registers
ma_tpe_step_kernel(MaPtrs S, const void *__restrict__ actions, float *__restrict__ obs, float *__restrict__ rew,
                   uint8_t *__restrict__ done_out, double *__restrict__ term, unsigned long long *__restrict__ stats,
                   const MaConsts C, const uint64_t seed, const uint64_t env_id0, const int64_t n, const int flags)
{
    extern __shared__ double sm_dyn[];
    const int lane_id = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double *wbase = sm_dyn + (size_t)warp * kTpeWarpDoubles;
    const TpeOb ob{wbase + lane_id};
    const int64_t i0 = (int64_t)blockIdx.x * kTpeBlock + warp * 32;  // first environment of this warp
    const int64_t i = i0 + lane_id;
    const bool live = i < n;
    const bool inf_obs = !(flags & HW_FLAG_NO_INF_OBS);
    const bool autoreset = !(flags & HW_FLAG_NO_AUTORESET);
    const bool shared_reward = (flags & HW_FLAG_SHARED_REWARD) != 0;
    const int K = S.K;
    const double dt = C.dt;
    constexpr int D2 = 2 * A + 7;  // float2 per observation row
    long long acc[7] = {0, 0, 0, 0, 0, 0, 0};  // this thread's finished episode, summed per warp at the end

    MaHead h;
    TpeCars<A> c;
    uint32_t done = 0, hit = 0;
    float odx[A][3], odv[A][3];
    uint32_t nolane = 0;
    uint32_t amask[A];
    double a0[A], a1[A];
    const uint64_t env_id = env_id0 + (uint64_t)i;
    if (live) {
        ma_load_head(h, S, i, n);
        // is_done persists until reset() (highway_core.py:239, :420): with HW_FLAG_NO_AUTORESET a car that finished in an earlier
        // step is still done, and the step then ends after its first tick (highway_env.py:85-86)
        done = (h.flags >> 8) & 0xFFu; h.flags &= 0xFFu;
#pragma unroll
        for (int a = 0; a < A; ++a) {
            const float *p = S.cars + (int64_t)(a * 9) * n + i;
            c.px[a] = (double)p[0]; c.py[a] = (double)p[n]; c.rx[a] = (double)p[2 * n]; c.vx[a] = (double)p[3 * n]; c.vy[a] = (double)p[4 * n];
            c.acc[a] = (double)p[5 * n]; c.steer[a] = (double)p[6 * n]; c.cruise[a] = (double)p[7 * n];
            const uint32_t w = __float_as_uint(p[8 * n]);
            c.b[a] = w & ~(3u << HW_BITS_LANE_SHIFT); c.ln[a] = tpe_lane(w);
        }
        for (int k = 0; k < (int)h.nobs; ++k) {
            const float4 o = S.obst[(int64_t)k * n + i];
            ob(k, O_PX) = o.x; ob(k, O_RX) = o.y; ob(k, O_VX) = o.z; ob(k, O_IV) = o.w;
        }
#pragma unroll
        for (int a = 0; a < A; ++a) {
            amask[a] = 0; a0[a] = 0.0; a1[a] = 0.0;
            if (CONT) {
                const float2 v = reinterpret_cast<const float2 *>(actions)[i * A + a];
                a0[a] = dmul(dmul((double)v.x, 5.0), dt);
                a1[a] = dmul(dmul((double)v.y, 30.0), dt);
            } else {
                const int act = reinterpret_cast<const int *>(actions)[i * A + a];
                amask[a] = ((uint32_t)act < 4u) ? (kBitLeft << act) : 0u;  // LEFT, RIGHT, ACCELERATE, MAINTAIN = bits 18..21
            }
        }
    }
#if HW_TPE_SYNC
    // block-synchronous ticks: every warp of the block enters tick t together, so the warps that share a scheduler walk through the
    // same stretch of the (large) tick body at the same time and share its instruction-cache lines
    {
        bool run = live;
        for (int t = 0; t < C.ticks_per_step; ++t) {
            if (HW_TPE_SYNC == 1) __syncthreads();
            if (run) {
                hit = tpe_tick<A, CONT, EXACT>(h, c, ob, K, C, dt, inf_obs, amask, a0, a1, seed, env_id, done);
                if (done) run = false;  // highway_env.py:85-86: the step ends with the first tick that finishes a car
            }
        }
    }
#else
    if (live) {
        for (int t = 0; t < C.ticks_per_step; ++t) {
            hit = tpe_tick<A, CONT, EXACT>(h, c, ob, K, C, dt, inf_obs, amask, a0, a1, seed, env_id, done);
            if (done) break;  // highway_env.py:85-86: the step ends with the first tick that finishes a car
        }
    }
#endif
    if (live) {

        // rewards of the step (Scenario.rewards of the last tick that ran, simple_base.py:65-80): -250 for a car that collided in it,
        // else v.x / 20 - 1; summed in id order for shared_reward (highway_env.py:92-94) and for Monitor
        double r[A], rsum = 0.0;
#pragma unroll
        for (int a = 0; a < A; ++a) {
            r[a] = dsub(ddiv_const(c.vx[a], 20.0), 1.0);
            tset_if(((hit >> a) & 1u) != 0u, r[a], -250.0);
            rsum = dadd(rsum, r[a]);
        }
#pragma unroll
        for (int a = 0; a < A; ++a) {
            const double ra = shared_reward ? rsum : r[a];
            h.ep_ret = dadd(h.ep_ret, ra);  // Monitor sums the python floats of every agent and step
            rew[i * A + a] = (float)ra;
            done_out[i * A + a] = (done >> a) & 1u;
        }
        h.ep_len += 1;
        if (done) {
            if (term) {
                double *tm = term + 5 * i;
                tm[0] = (double)h.tick; tm[1] = (double)h.nc_obs; tm[2] = (double)h.nc_agent; tm[3] = h.ep_ret; tm[4] = (double)h.ep_len;
            }
            acc[0] = 1; acc[1] = h.ep_len; acc[2] = __double2ll_rn(h.ep_ret * 1000.0); acc[3] = h.nc_obs;
            acc[4] = (h.tick >= (uint32_t)C.timeout_tick) ? 1 : 0; acc[5] = h.nc_agent; acc[6] = h.flags & 1u;
            if (autoreset) {
                const uint32_t keep = h.flags;
                tpe_reset<A>(h, c, ob);
                h.flags = keep;  // sticky diagnostics survive the reset
            }
        }
        if (!(done && autoreset)) h.flags |= done << 8;

        // the part of the observation that reads the obstacle list, then the state goes back to HBM (fp32) and the list's
        // shared memory is free for the staging of the rows
        nolane = tpe_observe_obstacles<A>(h, c, ob, odx, odv);
        ma_store_head(h, S, i, n);
#pragma unroll
        for (int a = 0; a < A; ++a) {
            float *p = S.cars + (int64_t)(a * 9) * n + i;
            p[0] = (float)c.px[a]; p[n] = (float)c.py[a]; p[2 * n] = (float)c.rx[a]; p[3 * n] = (float)c.vx[a]; p[4 * n] = (float)c.vy[a];
            p[5 * n] = (float)c.acc[a]; p[6 * n] = (float)c.steer[a]; p[7 * n] = (float)c.cruise[a];
            p[8 * n] = __uint_as_float(tpe_with_lane(c.b[a], c.ln[a]));
        }
        for (int k = 0; k < (int)h.nobs; ++k)
            S.obst[(int64_t)k * n + i] = make_float4((float)ob(k, O_PX), (float)ob(k, O_RX), (float)ob(k, O_VX), (float)ob(k, O_IV));
    }

    // observation rows: staged in the warp's (now free) shared region, kPass cars at a time, and streamed out 8 bytes per lane -
    // consecutive lanes write consecutive addresses of the environments' contiguous [A][4A+14] blocks
    constexpr int kPass = (kTpeWarpDoubles / 32) / D2 < A ? (kTpeWarpDoubles / 32) / D2 : A;  // cars per pass that fit the region
    float2 *stage = reinterpret_cast<float2 *>(wbase);
    float2 *out2 = reinterpret_cast<float2 *>(obs);
#pragma unroll
    for (int c0 = 0; c0 < A; c0 += kPass) {
        const int cn = (A - c0 < kPass) ? A - c0 : kPass;  // cars in this pass
        __syncwarp();  // every lane is done with the region (obstacle reads / the previous pass's copy)
        if (live) {
#pragma unroll
            for (int a = c0; a < c0 + kPass; ++a) {
                if (a >= A) break;
                float2 *row = stage + (size_t)lane_id * (cn * D2) + (a - c0) * D2;
                tpe_observe_row<A>(c, a, odx, odv, nolane, [&](int q, float2 v) { row[q] = v; });
            }
        }
        __syncwarp();
        const int per_env = cn * D2;                  // compile-time after unrolling: the division below is a multiply and a shift
        const int nv = (n - i0 < 32) ? (int)(n - i0) : 32;   // environments of this warp that exist
        float2 *dst = out2 + i0 * (int64_t)(A * D2) + c0 * D2;
        if (nv == 32) {  // independent iterations, unrolled so that their LDS -> STG chains overlap
#pragma unroll 6
            for (int idx = lane_id; idx < 32 * per_env; idx += 32) {
                const int e = idx / per_env, q = idx - e * per_env;
                dst[e * (A * D2) + q] = stage[idx];
            }
        } else {
            for (int idx = lane_id; idx < 32 * per_env; idx += 32) {
                const int e = idx / per_env, q = idx - e * per_env;
                if (e < nv) dst[e * (A * D2) + q] = stage[idx];
            }
        }
    }

    if (stats) {
        if (__any_sync(0xffffffffu, acc[0] != 0)) {
#pragma unroll
            for (int j = 0; j < 7; ++j) acc[j] = warp_sum(acc[j]);
            if (lane_id == 0) {
#pragma unroll
                for (int j = 0; j < 7; ++j) atomicAdd(&stats[j], (unsigned long long)acc[j]);
            }
        }
    }
}

template <int A>
static int ma_tpe_launch_a(const MaPtrs &S, const void *actions, const HwMaOut *out, const MaConsts &C_, uint64_t seed, uint64_t env_id0,
                           int64_t n, int flags, cudaStream_t stream)
{
    const size_t smem = (size_t)(kTpeBlock / 32) * kTpeWarpDoubles * sizeof(double);
    const unsigned grid = (unsigned)((n + kTpeBlock - 1) / kTpeBlock);
    cudaError_t e = cudaSuccess;
#define HW_TPE_LAUNCH(CT, EX) do { \
        e = cudaFuncSetAttribute(ma_tpe_step_kernel<A, CT, EX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        if (e != cudaSuccess) return (int)e; \
        ma_tpe_step_kernel<A, CT, EX><<<grid, kTpeBlock, smem, stream>>>(S, actions, out->obs, out->rew, out->done, out->term, out->stats, C_, seed, env_id0, n, flags); \
    } while (0)
    const bool exact = (flags & HW_FLAG_EXACT_STEERING) != 0;
    if (flags & HW_FLAG_CONTINUOUS) { if (exact) HW_TPE_LAUNCH(true, true); else HW_TPE_LAUNCH(true, false); }
    else                            { if (exact) HW_TPE_LAUNCH(false, true); else HW_TPE_LAUNCH(false, false); }
#undef HW_TPE_LAUNCH
    return (int)cudaGetLastError();
}

int ma_tpe_launch_step(const HwMaState *st, const void *actions, const HwMaOut *out, const MaConsts &C_,
                       uint64_t seed, uint64_t env_id0, int64_t n, int flags, cudaStream_t stream)
{
    MaPtrs S;
    S.cars = st->cars; S.obst = reinterpret_cast<float4 *>(st->obst); S.head = st->head; S.A = st->num_agents; S.K = st->num_slots;
#ifdef HW_TPE_ONLY   // experiment builds: one instantiation instead of eight (compile time)
    if (S.A == HW_TPE_ONLY) return ma_tpe_launch_a<HW_TPE_ONLY>(S, actions, out, C_, seed, env_id0, n, flags, stream);
#else
    switch (S.A) {
    case 1: return ma_tpe_launch_a<1>(S, actions, out, C_, seed, env_id0, n, flags, stream);
    case 2: return ma_tpe_launch_a<2>(S, actions, out, C_, seed, env_id0, n, flags, stream);
    case 3: return ma_tpe_launch_a<3>(S, actions, out, C_, seed, env_id0, n, flags, stream);
    case 4: return ma_tpe_launch_a<4>(S, actions, out, C_, seed, env_id0, n, flags, stream);
    case 5: return ma_tpe_launch_a<5>(S, actions, out, C_, seed, env_id0, n, flags, stream);
    case 6: return ma_tpe_launch_a<6>(S, actions, out, C_, seed, env_id0, n, flags, stream);
    case 7: return ma_tpe_launch_a<7>(S, actions, out, C_, seed, env_id0, n, flags, stream);
    case 8: return ma_tpe_launch_a<8>(S, actions, out, C_, seed, env_id0, n, flags, stream);
    }
#endif
    return HW_E_BADARG;
}

}  // namespace hw

