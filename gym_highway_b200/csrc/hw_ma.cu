// Multi-agent highway step path for sm_100a: one thread per environment (A <= 5 cars, K <= 16 obstacle slots).
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
// The simulator is order dependent (cars act and move in id order and see each other's partially updated
// state; obstacles are an insertion-ordered list that grows and shrinks), which one thread per environment
// reproduces by construction.  The fp32 state is promoted to fp64 in SHARED memory for the duration of the
// step ([field][thread] layout: every access is conflict free), the ticks run op-for-op like the reference,
// and the state is rounded back to fp32 once at the end.
#include "hw_ma_common.cuh"

namespace hw {

#define CAR(a, f) sm_car[((a) * C_NF + (f)) * kMaBlock + threadIdx.x]
#define OB(k, f) sm_ob[((k) * O_NF + (f)) * kMaBlock + threadIdx.x]

// start grid: simple_base.py:32-43 ("X" formation, three obstacles behind the cars); counters of the episode cleared
__device__ __forceinline__ void ma_reset_env(MaHead &h, double *sm_car, double *sm_ob, uint32_t *cbits, int A)
{
    const double cx[5] = {10.0, 10.0, 7.5, 5.0, 5.0};
    const int cl[5] = {1, 3, 2, 1, 3};
    for (int a = 0; a < A; ++a) {
        CAR(a, C_PX) = cx[a]; CAR(a, C_PY) = lane_centre(cl[a] - 1); CAR(a, C_RX) = cx[a];
        CAR(a, C_VX) = 0.0; CAR(a, C_VY) = 0.0; CAR(a, C_ACC) = 0.0; CAR(a, C_STEER) = 0.0; CAR(a, C_CRUISE) = 0.0;
        cbits[a] = (uint32_t)cl[a] << HW_BITS_LANE_SHIFT;
    }
    const double ox[3] = {-20.0, -25.0, -40.0}, ov[3] = {7.0, 5.0, 6.0};
    for (int k = 0; k < 3; ++k) { OB(k, O_PX) = ox[k]; OB(k, O_RX) = ox[k]; OB(k, O_VX) = ov[k]; OB(k, O_IV) = ov[k]; }
    h.nobs = 3; h.olane = 1u | (2u << 2) | (3u << 4); h.ofresh = 7u;
    h.newest[0] = 0; h.newest[1] = 1; h.newest[2] = 2;
    h.tick = 0; h.leader = 0; h.nc_obs = 0; h.nc_agent = 0; h.ep_len = 0; h.ep_ret = 0.0; h.flags = 0;
}

// one world.act() followed by Scenario.rewards(); rewards[a] is overwritten, done bits are or-ed into `done`
template <bool CONT, bool EXACT>
__device__ __forceinline__ void ma_tick(MaHead &h, double *sm_car, double *sm_ob, uint32_t *cbits, const int A, const int K,
                                        const MaConsts &C, const bool inf_obs, const int *act_d, const double *act_c,
                                        const uint64_t seed, const uint64_t env_id, double *rewards, uint32_t &done)
{
    const double dt = C.dt;
    h.tick += 1;
    if (h.tick >= (uint32_t)C.timeout_tick) done |= (1u << A) - 1u;  // highway_core.py:276-277

    // ---- actions for every car in id order (highway_core.py:280-281): later cars see earlier cars' new lanes ----
    for (int a = 0; a < A; ++a) {
        uint32_t b = cbits[a];
        if (CONT) {
            CAR(a, C_ACC) = py_clamp(dadd(CAR(a, C_ACC), dmul(dmul(act_c[2 * a], 5.0), dt)), -5.0, 5.0);
            CAR(a, C_STEER) = py_clamp(dadd(CAR(a, C_STEER), dmul(dmul(act_c[2 * a + 1], 30.0), dt)), -30.0, 30.0);
            continue;
        }
        const int act = act_d[a];
        int lane = (int)((b >> HW_BITS_LANE_SHIFT) & 3u);
        if (act == HW_ACT_ACCELERATE && !(b & kBitDoAcc)) {
            b = (b | kBitDoAcc) & ~kBitDoMaint;
        } else if (act == HW_ACT_MAINTAIN && !(b & kBitDoMaint)) {
            // all_obstacles iterates the cars (id order) and then the obstacles (insertion order): nearest strictly ahead
            const double mypx = CAR(a, C_PX);
            bool found = false;
            double bpx = 0.0, bvx = 0.0;
            for (int j = 0; j < A; ++j) {
                if (j == a) continue;
                const double opx = CAR(j, C_PX);
                if ((int)((cbits[j] >> HW_BITS_LANE_SHIFT) & 3u) == lane && opx > mypx && (!found || opx < bpx)) { found = true; bpx = opx; bvx = CAR(j, C_VX); }
            }
            for (int k = 0; k < (int)h.nobs; ++k) {
                const double opx = OB(k, O_PX);
                if (ob_lane(h, k) == lane && opx > mypx && (!found || opx < bpx)) { found = true; bpx = opx; bvx = OB(k, O_VX); }
            }
            CAR(a, C_CRUISE) = found ? bvx : CAR(a, C_VX);
            b = (b | kBitDoMaint) & ~kBitDoAcc;
        }
        CAR(a, C_ACC) = py_clamp(CAR(a, C_ACC), -5.0, 5.0);
        if (act == HW_ACT_RIGHT && !(b & kBitRight)) {
            lane = min(lane + 1, 3);
            b = (b | kBitRight) & ~kBitLeft;
        } else if (act == HW_ACT_LEFT && !(b & kBitLeft)) {
            lane = max(lane - 1, 1);
            b = (b | kBitLeft) & ~kBitRight;
        }
        CAR(a, C_STEER) = py_clamp(CAR(a, C_STEER), -30.0, 30.0);
        cbits[a] = (b & ~(3u << HW_BITS_LANE_SHIFT)) | ((uint32_t)lane << HW_BITS_LANE_SHIFT);
    }

    // ---- car updates in id order against last tick's leader (agent.py:79-165) ----
    const int L = (int)h.leader;
    for (int a = 0; a < A; ++a) {
        uint32_t b = cbits[a];
        int lane = (int)((b >> HW_BITS_LANE_SHIFT) & 3u);
        double acc = CAR(a, C_ACC), vx = CAR(a, C_VX), steer = CAR(a, C_STEER), py = CAR(a, C_PY);
        if (!CONT) {
            if (b & kBitDoAcc) {
                acc = (acc < 0.0) ? 0.0 : dadd(acc, dt);
                if (acc == 5.0) b &= ~kBitDoAcc;
            } else if (b & kBitDoMaint) {
                const double cruise = CAR(a, C_CRUISE);
                const double vc = ceil(vx), cc = ceil(cruise);
                acc = (vc <= cc) ? 5.0 : -5.0;
                if (vc == cc) { vx = cruise; acc = 0.0; b &= ~kBitDoMaint; }
            }
        }
        vx = py_clamp(dadd(vx, dmul(acc, dt)), 0.0, 20.0);
        if (!CONT) {
            const double ty = lane_centre(lane - 1);
            if (b & kBitLeft) {
                steer = dadd(steer, C.k30dt);
                if (py <= ty) { steer = 0.0; b &= ~kBitLeft; }
            } else if (b & kBitRight) {
                steer = dsub(steer, C.k30dt);
                if (py >= ty) { steer = 0.0; b &= ~kBitRight; }
            }
        }
        double ang = 0.0;
        if (steer != 0.0) ang = angular_velocity<EXACT>(vx, steer);
        double vy = CAR(a, C_VY);
        if (!CONT) vy = ang;  // agent.py:144: velocity.y = angular_velocity (discrete update only)
        const double vdt = dmul(vx, dt);
        double px = dadd(CAR(a, C_PX), vdt);
        py = dadd(py, dmul(vy, dt));
        if (CONT) py = dsub(py, dmul(ang, dt));
        else      py = dsub(py, dmul(dmul(dmul(ang, kRad2Deg), dt), dt));
        CAR(a, C_VX) = vx;  // the leader's velocity "as it is now" for the cars after it
        if (a == L) px = 10.0;
        else        px = dsub(px, dmul(CAR(L, C_VX), dt));
        if (py < 1.0) py = 1.0; else if (py > 6.0) py = (7.0 < py) ? 7.0 : py;
        CAR(a, C_RX) = dadd(CAR(a, C_RX), vdt);
        if (CONT) {
            const double d0 = fabs(dsub(py, 1.25)), d1 = fabs(dsub(py, 3.75)), d2 = fabs(dsub(py, 6.25));
            lane = (d0 <= d1 && d0 <= d2) ? 1 : (d1 <= d2) ? 2 : 3;
        }
        CAR(a, C_ACC) = acc; CAR(a, C_STEER) = steer; CAR(a, C_VY) = vy; CAR(a, C_PX) = px; CAR(a, C_PY) = py;
        cbits[a] = (b & ~(3u << HW_BITS_LANE_SHIFT)) | ((uint32_t)lane << HW_BITS_LANE_SHIFT);
    }

    // ---- obstacle updates (agent.py:223-260): yield to the nearest car ahead in the lane, by raw x ----
    const double lvdt = dmul(CAR(L, C_VX), dt);
    for (int k = 0; k < (int)h.nobs; ++k) {
        const int lane = ob_lane(h, k);
        const double orx = OB(k, O_RX), iv = OB(k, O_IV);
        double v = py_clamp(iv, -20.0, 20.0);
        bool found = false;
        double bd = 0.0, bv = 0.0;
        for (int j = 0; j < A; ++j) {
            const double crx = CAR(j, C_RX);
            if ((int)((cbits[j] >> HW_BITS_LANE_SHIFT) & 3u) == lane && orx < crx) {
                const double d = dsub(crx, orx);
                if (!found || d < bd) { found = true; bd = d; bv = CAR(j, C_VX); }
            }
        }
        if (found) v = (bv < iv) ? bv : iv;
        OB(k, O_VX) = v;
        const double odt = dmul(v, dt);
        OB(k, O_PX) = dsub(dadd(OB(k, O_PX), odt), lvdt);
        OB(k, O_RX) = dadd(orx, odt);
    }
    h.ofresh = 0;  // every live obstacle has now written its rect

    // ---- leader / last election: stable sort by x, descending (highway_core.py:294-297) ----
    int lead = 0, last = 0;
    for (int j = 1; j < A; ++j) {
        if (CAR(j, C_PX) > CAR(lead, C_PX)) lead = j;    // first of the maxima
        if (CAR(j, C_PX) <= CAR(last, C_PX)) last = j;   // last of the minima
    }
    h.leader = (uint32_t)lead;

    if (inf_obs) {
        // ---- spawn; the old obstacle stays and is slowed down (highway_core.py:300-323) ----
        for (int lane = 0; lane < 3; ++lane) {
            const uint32_t s = h.newest[lane];
            if (s >= 31u) continue;  // the lane's newest obstacle was retired: the reference never spawns there again
            if (OB(s, O_PX) < -2.375) {
                double u1, u2;
                spawn_uniforms(seed, env_id, h.spawn_ctr, u1, u2);
                h.spawn_ctr += 1;
                const double x = dadd(70.0, dmul(30.0, u1)), v = dadd(5.0, dmul(2.0, u2));
                const double ovx = OB(s, O_VX);
                const double m = (ovx < v) ? ovx : v;  // min(rand_vel_x, obstacle.velocity.x)
                OB(s, O_VX) = m; OB(s, O_IV) = m;
                if ((int)h.nobs < K) {
                    const int nk = (int)h.nobs;
                    OB(nk, O_PX) = x; OB(nk, O_RX) = dadd(CAR(lead, C_RX), dsub(x, CAR(lead, C_PX)));
                    OB(nk, O_VX) = v; OB(nk, O_IV) = v;
                    h.olane = (h.olane & ~(3u << (2 * nk))) | ((uint32_t)(lane + 1) << (2 * nk));
                    h.ofresh |= 1u << nk;
                    h.newest[lane] = (uint32_t)nk;
                    h.nobs += 1;
                } else {
                    h.flags |= 1u;  // out of slots: the spawn is dropped (tests require this never to happen)
                }
            }
        }
        // ---- retire what the last car has cleared (highway_core.py:327-342); order preserving compaction ----
        const double thr = dsub(dsub(-2.375, CAR(lead, C_PX)), 2.0);
        const double last_rx = CAR(last, C_RX);
        int w = 0;
        uint32_t nl = 0, nf = 0, nn0 = 31u, nn1 = 31u, nn2 = 31u;
        for (int k = 0; k < (int)h.nobs; ++k) {
            if (dsub(OB(k, O_RX), last_rx) < thr) continue;
            if (w != k) { OB(w, O_PX) = OB(k, O_PX); OB(w, O_RX) = OB(k, O_RX); OB(w, O_VX) = OB(k, O_VX); OB(w, O_IV) = OB(k, O_IV); }
            nl |= (uint32_t)ob_lane(h, k) << (2 * w);
            nf |= ((h.ofresh >> k) & 1u) << w;
            if (h.newest[0] == (uint32_t)k) nn0 = (uint32_t)w;
            if (h.newest[1] == (uint32_t)k) nn1 = (uint32_t)w;
            if (h.newest[2] == (uint32_t)k) nn2 = (uint32_t)w;
            ++w;
        }
        h.nobs = (uint32_t)w; h.olane = nl; h.ofresh = nf; h.newest[0] = nn0; h.newest[1] = nn1; h.newest[2] = nn2;
    }

    // ---- Scenario.rewards -> check_collisions, after the updates and spawns of this tick ----
    int cax[kMaxA], cay[kMaxA];
#pragma unroll
    for (int a = 0; a < kMaxA; ++a) if (a < A) { cax[a] = rect_x(CAR(a, C_PX)); cay[a] = rect_y(CAR(a, C_PY)); }
#pragma unroll
    for (int a = 0; a < kMaxA; ++a) {
        if (a >= A) break;
        int no = 0, na = 0;
        for (int k = 0; k < (int)h.nobs; ++k) {
            const bool fresh = (h.ofresh >> k) & 1u;
            const int bx = fresh ? 0 : rect_x(OB(k, O_PX));
            const int by = fresh ? 0 : (21 + 80 * (ob_lane(h, k) - 1));
            no += rects_overlap(cax[a], cay[a], bx, by) ? 1 : 0;
        }
#pragma unroll
        for (int j = 0; j < kMaxA; ++j) if (j < A && j != a) na += rects_overlap(cax[a], cay[a], cax[j], cay[j]) ? 1 : 0;
        h.nc_obs += no; h.nc_agent += na;
        if (no + na > 0) { done |= 1u << a; rewards[a] = -250.0; }
        else rewards[a] = dsub(ddiv_const(CAR(a, C_VX), 20.0), 1.0);
    }
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

template <bool CONT, bool EXACT>
__device__ __forceinline__ void ma_step_env(const MaPtrs &S, double *sm_car, double *sm_ob, const void *__restrict__ actions,
                                            float *__restrict__ obs, float *__restrict__ rew, uint8_t *__restrict__ done_out,
                                            double *__restrict__ term, const MaConsts &C, const uint64_t seed, const uint64_t env_id0,
                                            const int64_t n, const int flags, const int64_t i, long long *acc)
{
    const int A = S.A, K = S.K, D = 4 * A + 14;
    const bool inf_obs = !(flags & HW_FLAG_NO_INF_OBS);
    const bool autoreset = !(flags & HW_FLAG_NO_AUTORESET);
    const bool shared_reward = (flags & HW_FLAG_SHARED_REWARD) != 0;

    MaHead h;
    uint32_t cbits[kMaxA];
    ma_load(h, sm_car, sm_ob, cbits, S, i, n);

    int act_d[kMaxA];
    double act_c[2 * kMaxA];
#pragma unroll
    for (int a = 0; a < kMaxA; ++a) {
        if (a >= A) break;
        if (CONT) {
            const float2 v = reinterpret_cast<const float2 *>(actions)[i * A + a];
            act_c[2 * a] = (double)v.x; act_c[2 * a + 1] = (double)v.y;
        } else {
            act_d[a] = reinterpret_cast<const int *>(actions)[i * A + a];
        }
    }

    double r[kMaxA];
#pragma unroll
    for (int a = 0; a < kMaxA; ++a) r[a] = 0.0;
    uint32_t done = (h.flags >> 8) & 31u;  // is_done persists until reset() (highway_core.py:239, :420)
    h.flags &= 0xFFu;
    for (int t = 0; t < C.ticks_per_step; ++t) {
        ma_tick<CONT, EXACT>(h, sm_car, sm_ob, cbits, A, K, C, inf_obs, act_d, act_c, seed, env_id0 + (uint64_t)i, r, done);
        if (done) break;  // highway_env.py:85-86
    }
    if (shared_reward) {  // highway_env.py:92-94: every agent receives np.sum(reward_n)
        double s = 0.0;
#pragma unroll
        for (int a = 0; a < kMaxA; ++a) if (a < A) s = dadd(s, r[a]);
#pragma unroll
        for (int a = 0; a < kMaxA; ++a) r[a] = s;
    }
#pragma unroll
    for (int a = 0; a < kMaxA; ++a) {
        if (a >= A) break;
        rew[i * A + a] = (float)r[a];
        done_out[i * A + a] = (done >> a) & 1u;
        h.ep_ret = dadd(h.ep_ret, r[a]);  // Monitor sums the python floats of every agent and step
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
            ma_reset_env(h, sm_car, sm_ob, cbits, A);
            h.flags = keep;  // sticky diagnostics survive the reset
        }
    }
    if (!(done && autoreset)) h.flags |= done << 8;
    ma_observe(h, sm_car, sm_ob, A, obs + i * (int64_t)A * D);
    ma_store(h, sm_car, sm_ob, cbits, S, i, n);
}

template <bool CONT, bool EXACT>
__global__ void __launch_bounds__(kMaBlock)
ma_step_kernel(MaPtrs S, const void *__restrict__ actions, float *__restrict__ obs, float *__restrict__ rew,
               uint8_t *__restrict__ done_out, double *__restrict__ term, unsigned long long *__restrict__ stats,
               const MaConsts C, const uint64_t seed, const uint64_t env_id0, const int64_t n, const int flags)
{
    extern __shared__ double sm_dyn[];
    double *sm_car = sm_dyn;                                   // [A][C_NF][kMaBlock]
    double *sm_ob = sm_dyn + (size_t)S.A * C_NF * kMaBlock;    // [K][O_NF][kMaBlock]
    const int64_t i = (int64_t)blockIdx.x * kMaBlock + threadIdx.x;
    long long acc[7] = {0, 0, 0, 0, 0, 0, 0};  // this thread's finished episode, summed per warp at the end
    if (i < n) ma_step_env<CONT, EXACT>(S, sm_car, sm_ob, actions, obs, rew, done_out, term, C, seed, env_id0, n, flags, i, acc);
    if (stats) {
        const long long eps = warp_sum(acc[0]);
        if (eps != 0) {
#pragma unroll
            for (int j = 1; j < 7; ++j) acc[j] = warp_sum(acc[j]);
            if ((threadIdx.x & 31) == 0) {
                atomicAdd(&stats[0], (unsigned long long)eps);
#pragma unroll
                for (int j = 1; j < 7; ++j) atomicAdd(&stats[j], (unsigned long long)acc[j]);
            }
        }
    }
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
    uint32_t cbits[kMaxA];
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
    uint32_t cbits[kMaxA];
    ma_load(h, sm_car, sm_ob, cbits, S, i, n);
    ma_observe(h, sm_car, sm_ob, S.A, obs + i * (int64_t)S.A * (4 * S.A + 14));
}

static inline bool ma_aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

static int ma_check(const HwMaState *st, int64_t n)
{
    if (!st || n <= 0 || !st->cars || !st->obst || !st->head) return HW_E_BADARG;
    if (st->num_agents < 1 || st->num_agents > kMaxA || st->num_slots < 3 || st->num_slots > kMaxK) return HW_E_BADARG;
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
    if (!(flags & HW_FLAG_MA_THREAD_PER_ENV)) return ma_grp_launch_step(st, actions, out, C_, seed, env_id0, n, flags, stream);
    const size_t smem = ma_smem(st);
    const unsigned grid = (unsigned)((n + kMaBlock - 1) / kMaBlock);
#define HW_MA_LAUNCH(C, E) do { \
        if ((rc = ma_prepare(ma_step_kernel<C, E>, smem))) return rc; \
        ma_step_kernel<C, E><<<grid, kMaBlock, smem, stream>>>(ma_ptrs(st), actions, out->obs, out->rew, out->done, out->term, out->stats, C_, seed, env_id0, n, flags); \
    } while (0)
    const bool exact = (flags & HW_FLAG_EXACT_STEERING) != 0;
    if (flags & HW_FLAG_CONTINUOUS) { if (exact) HW_MA_LAUNCH(true, true); else HW_MA_LAUNCH(true, false); }
    else                            { if (exact) HW_MA_LAUNCH(false, true); else HW_MA_LAUNCH(false, false); }
#undef HW_MA_LAUNCH
    return (int)cudaGetLastError();
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
