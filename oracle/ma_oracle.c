/* CPU restatement (fp64, scalar C) of gym-highway's MULTI-AGENT step path.
 *
 * TEST INFRASTRUCTURE ONLY (see sa_oracle.c for the rules: only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may execute this).
 *
 * Parity pin: unpinned by the reference (it has no tests); pinned here by tests/golden/ma_*.npz, which
 * oracle/gen_golden_ma.py records from the reference's own unmodified code under oracle/shim.
 *
 * Reference lines followed (all under /root/reference/gym_highway/multiagent_envs/):
 *   start grid, rewards, observations, dones   simple_base.py:29-46, :62-80, :82-154
 *   reset                                      highway_core.py:190-243
 *   apply action                               highway_core.py:110-188
 *   tick driver, spawn, retire                 highway_core.py:245-399
 *   collisions                                 highway_core.py:401-422
 *   car update                                 agent.py:79-116 (continuous), :118-165 (discrete), :170-208
 *   obstacle update                            agent.py:223-260
 *   env.step / normalisation                   highway_env.py:63-99, :139-149, highway_ma_c_env.py:56-82
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <string.h>
#include "philox.h"
#include "tan_small.h"

#define HW_PI 3.14159265358979323846
static const double DEG2RAD = HW_PI / 180.0;
static const double RAD2DEG = 180.0 / HW_PI;
static const double LANE_C[3] = { 1.25, 3.75, 6.25 };

enum { ACT_LEFT = 0, ACT_RIGHT = 1, ACT_ACCELERATE = 2, ACT_MAINTAIN = 3 };
enum { F_LEFT = 1, F_RIGHT = 2, F_DO_ACC = 4, F_DO_MAINT = 8 };
enum { HW_FLAG_CONTINUOUS = 1, HW_FLAG_NO_INF_OBS = 2, HW_FLAG_NO_AUTORESET = 4, HW_FLAG_QUANTISE_F32 = 8,
       HW_FLAG_SHARED_REWARD = 16 };

#define MA_MAX_A 8   /* 1..5: the reference's start grid; 6..8: the repo's extension (no reference behind it) */
#define MA_MAX_K 64

typedef struct { double px, py, rx, vx, vy, acc, steer, cruise; int lane, flags; } ma_car;
typedef struct { double px, rx, vx, iv; int lane, fresh; } ma_obs;
typedef struct {
    int A, K, nobs, tick, leader, newest[3], nc_obs, nc_agent, overflow;
    uint32_t spawn_ctr;
    ma_car car[MA_MAX_A];
    ma_obs ob[MA_MAX_K];
} ma_env;

static inline double py_clamp(double x, double lo, double hi)
{
    double inner = (hi < x) ? hi : x;
    return (inner > lo) ? inner : lo;
}
static inline int trunc_i(double v) { return (int)v; }
static inline int overlap(int ax, int ay, int bx, int by) { return ax < bx + 76 && ay < by + 38 && ax + 76 > bx && ay + 38 > by; }

static void ma_reset(ma_env *e)
{
    /* simple_base.py:32-43 "X" formation, three obstacles behind the cars */
    /* cars 5..7 are NOT in the reference (its grid ends at five, simple_base.py:32-38): the extension continues the pattern */
    static const double cx[8] = { 10.0, 10.0, 7.5, 5.0, 5.0, 2.5, 0.0, -2.5 };
    static const int cl[8] = { 1, 3, 2, 1, 3, 2, 3, 1 };
    static const double ox[3] = { -20.0, -25.0, -40.0 }, ov[3] = { 7.0, 5.0, 6.0 };
    for (int i = 0; i < e->A; ++i) {
        ma_car *c = &e->car[i];
        c->px = cx[i]; c->py = LANE_C[cl[i] - 1]; c->rx = cx[i]; c->vx = 0.0; c->vy = 0.0;
        c->acc = 0.0; c->steer = 0.0; c->cruise = 0.0; c->lane = cl[i]; c->flags = 0;
    }
    e->nobs = 3;
    for (int k = 0; k < 3; ++k) {
        ma_obs *o = &e->ob[k];
        o->px = ox[k]; o->rx = ox[k]; o->vx = ov[k]; o->iv = ov[k]; o->lane = k + 1; o->fresh = 1; /* rect still (0,0) */
        e->newest[k] = k;
    }
    e->tick = 0; e->leader = 0; e->nc_obs = 0; e->nc_agent = 0; e->overflow = 0;
}

/* one world.act() + Scenario.rewards(); fills reward[A], or-s into done[A] */
static void ma_tick(ma_env *e, double dt, int timeout_tick, int cont, int inf_obs, const int *act_d, const double *act_c,
                    uint64_t seed, uint64_t env_id, const double *u_override, int *u_used, double *reward, uint8_t *done)
{
    const int A = e->A;
    e->tick += 1;
    if (e->tick >= timeout_tick) for (int i = 0; i < A; ++i) done[i] = 1; /* highway_core.py:276-277 */

    /* --- actions for every car in id order (highway_core.py:280-281) --- */
    for (int i = 0; i < A; ++i) {
        ma_car *c = &e->car[i];
        if (cont) {
            c->acc = py_clamp(c->acc + act_c[2 * i] * 5.0 * dt, -5.0, 5.0);
            c->steer = py_clamp(c->steer + act_c[2 * i + 1] * 30.0 * dt, -30.0, 30.0);
            continue;
        }
        const int a = act_d[i];
        if (a == ACT_ACCELERATE && !(c->flags & F_DO_ACC)) {
            c->flags = (c->flags | F_DO_ACC) & ~F_DO_MAINT;
        } else if (a == ACT_MAINTAIN && !(c->flags & F_DO_MAINT)) {
            /* all_obstacles iterates cars (id order) then obstacles (insertion order); nearest strictly ahead */
            int found = 0; double bpx = 0, bvx = 0;
            for (int j = 0; j < A; ++j) {
                if (j == i) continue;
                const ma_car *o = &e->car[j];
                if (o->lane == c->lane && o->px > c->px) {
                    if (!found) { found = 1; bpx = o->px; bvx = o->vx; }
                    else if (o->px < bpx) { bpx = o->px; bvx = o->vx; }
                }
            }
            for (int k = 0; k < e->nobs; ++k) {
                const ma_obs *o = &e->ob[k];
                if (o->lane == c->lane && o->px > c->px) {
                    if (!found) { found = 1; bpx = o->px; bvx = o->vx; }
                    else if (o->px < bpx) { bpx = o->px; bvx = o->vx; }
                }
            }
            c->cruise = found ? bvx : c->vx;
            c->flags = (c->flags | F_DO_MAINT) & ~F_DO_ACC;
        }
        c->acc = py_clamp(c->acc, -5.0, 5.0);
        if (a == ACT_RIGHT && !(c->flags & F_RIGHT)) {
            c->lane = c->lane + 1 < 3 ? c->lane + 1 : 3;
            c->flags = (c->flags | F_RIGHT) & ~F_LEFT;
        } else if (a == ACT_LEFT && !(c->flags & F_LEFT)) {
            c->lane = c->lane - 1 > 1 ? c->lane - 1 : 1;
            c->flags = (c->flags | F_LEFT) & ~F_RIGHT;
        }
        c->steer = py_clamp(c->steer, -30.0, 30.0);
    }

    /* --- car updates in id order against last tick's leader (agent.py:79-165) --- */
    const int L = e->leader;
    for (int i = 0; i < A; ++i) {
        ma_car *c = &e->car[i];
        if (!cont) {
            if (c->flags & F_DO_ACC) {
                if (c->acc < 0.0) c->acc = 0.0; else c->acc = c->acc + 1 * dt;
                if (c->acc == 5.0) c->flags &= ~F_DO_ACC;
            } else if (c->flags & F_DO_MAINT) {
                double vc = ceil(c->vx), cc = ceil(c->cruise);
                c->acc = (vc <= cc) ? 5.0 : -5.0;
                if (vc == cc) { c->vx = c->cruise; c->acc = 0.0; c->flags &= ~F_DO_MAINT; }
            }
        }
        c->vx = py_clamp(c->vx + c->acc * dt, 0.0, 20.0);
        if (!cont) {
            double ty = (80 * c->lane - 40.0) / 32.0;
            if (c->flags & F_LEFT) {
                c->steer = c->steer + 30.0 * dt;
                if (c->py <= ty) { c->steer = 0.0; c->flags &= ~F_LEFT; }
            } else if (c->flags & F_RIGHT) {
                c->steer = c->steer - 30.0 * dt;
                if (c->py >= ty) { c->steer = 0.0; c->flags &= ~F_RIGHT; }
            }
        }
        double ang = 0.0;
        if (c->steer != 0.0) ang = c->vx / (4.0 / hw_tan(c->steer * DEG2RAD));
        if (!cont) c->vy = ang; /* agent.py:144: velocity.y = angular_velocity (discrete only) */
        c->px = c->px + c->vx * dt;
        c->py = c->py + c->vy * dt;
        if (cont) c->py = c->py - ang * dt;
        else      c->py = c->py - (ang * RAD2DEG) * dt * dt;
        if (i == L) c->px = 10.0;
        else        c->px = c->px - e->car[L].vx * dt; /* leader's velocity as it is now */
        if (c->py < 1.0) c->py = 1.0;
        else if (c->py > 6.0) c->py = (7.0 < c->py) ? 7.0 : c->py;
        c->rx = c->rx + c->vx * dt;
        if (cont) {
            double d0 = fabs(c->py - LANE_C[0]), d1 = fabs(c->py - LANE_C[1]), d2 = fabs(c->py - LANE_C[2]);
            double m = d0; if (d1 < m) m = d1; if (d2 < m) m = d2;
            c->lane = (d0 == m) ? 1 : (d1 == m) ? 2 : 3;
        }
    }

    /* --- obstacle updates (agent.py:223-260) --- */
    for (int k = 0; k < e->nobs; ++k) {
        ma_obs *o = &e->ob[k];
        double v = py_clamp(o->iv, -20.0, 20.0);
        int found = 0; double bd = 0, bv = 0;
        for (int j = 0; j < A; ++j) {
            const ma_car *c = &e->car[j];
            if (c->lane == o->lane && o->rx < c->rx) {
                double d = c->rx - o->rx;
                if (!found) { found = 1; bd = d; bv = c->vx; }
                else if (d < bd) { bd = d; bv = c->vx; }
            }
        }
        if (found) v = (bv < o->iv) ? bv : o->iv; /* min(init_vx, closest.vx) */
        o->vx = v;
        o->px = o->px + v * dt;
        o->px = o->px - e->car[L].vx * dt;
        o->rx = o->rx + v * dt;
        o->fresh = 0;
    }

    /* --- leader / last election: stable sort by px descending (highway_core.py:294-297) --- */
    int lead = 0, last = 0;
    for (int j = 1; j < A; ++j) {
        if (e->car[j].px > e->car[lead].px) lead = j;   /* first of the maxima */
        if (e->car[j].px <= e->car[last].px) last = j;  /* last of the minima */
    }
    e->leader = lead;

    if (inf_obs) {
        /* --- spawn without removing the old obstacle (highway_core.py:300-323) --- */
        for (int lane = 0; lane < 3; ++lane) {
            const int s = e->newest[lane];
            if (s < 0) continue; /* the lane's newest obstacle was retired: the reference never spawns there again */
            ma_obs *o = &e->ob[s];
            if (o->px < -76.0 / 32.0) {
                double u1, u2;
                if (u_override) { u1 = u_override[*u_used]; u2 = u_override[*u_used + 1]; *u_used += 2; }
                else hw_spawn_uniforms(seed, env_id, e->spawn_ctr, &u1, &u2);
                e->spawn_ctr += 1;
                double x = 70.0 + 30.0 * u1, v = 5.0 + 2.0 * u2;
                double m = (o->vx < v) ? o->vx : v; /* min(rand_vel_x, obstacle.velocity.x) */
                o->vx = m; o->iv = m;
                if (e->nobs < e->K) {
                    ma_obs *nw = &e->ob[e->nobs];
                    nw->px = x; nw->rx = e->car[lead].rx + (x - e->car[lead].px);
                    nw->vx = v; nw->iv = v; nw->lane = lane + 1; nw->fresh = 1;
                    e->newest[lane] = e->nobs;
                    e->nobs += 1;
                } else {
                    e->overflow += 1; /* out of slots: the spawn is dropped (tests require this to stay 0) */
                }
            }
        }
        /* --- retire obstacles the last car has cleared (highway_core.py:327-342), stable --- */
        const double thr = (-76.0 / 32.0 - e->car[lead].px) - 2;
        int w = 0, map[MA_MAX_K];
        for (int k = 0; k < e->nobs; ++k) {
            if ((e->ob[k].rx - e->car[last].rx) < thr) { map[k] = -1; continue; }
            map[k] = w;
            if (w != k) e->ob[w] = e->ob[k];
            ++w;
        }
        for (int lane = 0; lane < 3; ++lane) if (e->newest[lane] >= 0) e->newest[lane] = map[e->newest[lane]];
        e->nobs = w;
    }

    /* --- Scenario.rewards -> check_collisions (after the updates and spawns of this tick) --- */
    int cax[MA_MAX_A], cay[MA_MAX_A];
    for (int i = 0; i < A; ++i) { cax[i] = trunc_i(e->car[i].px * 32.0 - 38.0); cay[i] = trunc_i(e->car[i].py * 32.0 - 19.0); }
    for (int i = 0; i < A; ++i) {
        int no = 0, na = 0;
        for (int k = 0; k < e->nobs; ++k) {
            const ma_obs *o = &e->ob[k];
            int bx = o->fresh ? 0 : trunc_i(o->px * 32.0 - 38.0);
            int by = o->fresh ? 0 : trunc_i(LANE_C[o->lane - 1] * 32.0 - 19.0);
            no += overlap(cax[i], cay[i], bx, by);
        }
        for (int j = 0; j < A; ++j) if (j != i) na += overlap(cax[i], cay[i], cax[j], cay[j]);
        e->nc_obs += no; e->nc_agent += na;
        if (no + na > 0) { done[i] = 1; reward[i] = -250.0; }
        else reward[i] = e->car[i].vx / 20.0 - 1.0;
    }
}

/* Scenario.observations + MultiAgentEnv._get_obs_norm for car i: float32 raw, float64 clip/normalise, float32 out */
static int ma_observe(const ma_env *e, float *obs /* [A][4A+14] */)
{
    const int A = e->A, D = 4 * A + 14, S = A + 2;
    int bad = 0;
    for (int i = 0; i < A; ++i) {
        const ma_car *c = &e->car[i];
        double pos[2 * (MA_MAX_A + 2)], vel[2 * (MA_MAX_A + 2)];
        int near[3] = { -1, -1, -1 };
        for (int k = 0; k < e->nobs; ++k) {
            const int l = e->ob[k].lane - 1;
            if (near[l] < 0) near[l] = k;
            else if (fabs(c->rx - e->ob[k].rx) < fabs(c->rx - e->ob[near[l]].rx)) near[l] = k;
        }
        for (int l = 0; l < 3; ++l) {
            const int s = A - 1 + l; /* obstacle id A+l is larger than every car id: slot id-1 */
            if (near[l] < 0) { bad = 1; pos[2 * s] = pos[2 * s + 1] = vel[2 * s] = vel[2 * s + 1] = 0.0; continue; }
            const ma_obs *o = &e->ob[near[l]];
            pos[2 * s] = o->rx - c->rx; pos[2 * s + 1] = LANE_C[l] - c->py;
            vel[2 * s] = o->vx - c->vx; vel[2 * s + 1] = 0.0 - c->vy;
        }
        for (int j = 0; j < A; ++j) {
            if (j == i) continue;
            const int s = j - (j > i);
            pos[2 * s] = e->car[j].rx - c->rx; pos[2 * s + 1] = e->car[j].py - c->py;
            vel[2 * s] = e->car[j].vx - c->vx; vel[2 * s + 1] = e->car[j].vy - c->vy;
        }
        float *ob = obs + (size_t)i * D;
        for (int j = 0; j < D; ++j) {
            double raw, lo, hi;
            if (j == 0) { raw = c->acc; lo = -5; hi = 5; }
            else if (j == 1) { raw = c->steer; lo = -30; hi = 30; }
            else if (j == 2) { raw = c->rx; lo = 0; hi = 1200; }
            else if (j == 3) { raw = c->py; lo = 0; hi = 8; }
            else if (j < 4 + 2 * S) { raw = pos[j - 4]; if ((j - 4) & 1) { lo = -8; hi = 8; } else { lo = -60; hi = 60; } }
            else if (j < 6 + 2 * S) { raw = (j == 4 + 2 * S) ? c->vx : c->vy; lo = -20; hi = 20; }
            else { raw = vel[j - 6 - 2 * S]; lo = -20; hi = 20; }
            double v = (double)(float)raw;      /* np.array(..., dtype=float32) */
            v = v < lo ? lo : v; v = v > hi ? hi : v; /* np.clip against the float64 bounds */
            ob[j] = (float)((v - lo) / (hi - lo)); /* float64 arithmetic, float32 VecEnv buffer */
        }
    }
    return bad;
}

static inline double q32(double v) { return (double)(float)v; }

typedef struct {
    int n, A, K;
    double *car; int32_t *lane, *cflags, *tick, *leader, *nobs; double *obst; int32_t *olane, *ofresh, *newest;
    int32_t *nc_obs, *nc_agent; uint32_t *spawn_ctr; double *ep_ret; int32_t *ep_len, *overflow, *done_mask;
    const int32_t *act_d; const double *act_c;
    float *obs; double *rew; uint8_t *done; double *term;
    double dt; int ticks_per_step, timeout_tick; uint64_t seed, env_id0; int hwflags;
    const double *u_override; int u_stride; int32_t *u_used_out;
    double *next_car; int32_t *next_meta; double *next_obst; int32_t *next_ometa;
    int i0, i1;
} ma_job;

static void ma_unpack(const ma_job *J, int i, ma_env *e)
{
    const int A = J->A, K = J->K;
    e->A = A; e->K = K; e->nobs = J->nobs[i]; e->tick = J->tick[i]; e->leader = J->leader[i];
    for (int l = 0; l < 3; ++l) e->newest[l] = J->newest[3 * (size_t)i + l];
    e->nc_obs = J->nc_obs[i]; e->nc_agent = J->nc_agent[i]; e->spawn_ctr = J->spawn_ctr[i]; e->overflow = J->overflow[i];
    for (int a = 0; a < A; ++a) {
        const double *c = J->car + ((size_t)i * A + a) * 8;
        ma_car *m = &e->car[a];
        m->px = c[0]; m->py = c[1]; m->rx = c[2]; m->vx = c[3]; m->vy = c[4]; m->acc = c[5]; m->steer = c[6]; m->cruise = c[7];
        m->lane = J->lane[(size_t)i * A + a]; m->flags = J->cflags[(size_t)i * A + a];
    }
    for (int k = 0; k < e->nobs; ++k) {
        const double *o = J->obst + ((size_t)i * K + k) * 4;
        ma_obs *m = &e->ob[k];
        m->px = o[0]; m->rx = o[1]; m->vx = o[2]; m->iv = o[3];
        m->lane = J->olane[(size_t)i * K + k]; m->fresh = J->ofresh[(size_t)i * K + k];
    }
}

static void ma_pack(const ma_env *e, double *car, int32_t *lane, int32_t *cflags, double *obst, int32_t *olane, int32_t *ofresh,
                    int A, int K, size_t i)
{
    for (int a = 0; a < A; ++a) {
        double *c = car + (i * A + a) * 8;
        const ma_car *m = &e->car[a];
        c[0] = m->px; c[1] = m->py; c[2] = m->rx; c[3] = m->vx; c[4] = m->vy; c[5] = m->acc; c[6] = m->steer; c[7] = m->cruise;
        lane[i * A + a] = m->lane; cflags[i * A + a] = m->flags;
    }
    for (int k = 0; k < K; ++k) {
        double *o = obst + (i * K + k) * 4;
        if (k < e->nobs) {
            const ma_obs *m = &e->ob[k];
            o[0] = m->px; o[1] = m->rx; o[2] = m->vx; o[3] = m->iv; olane[i * K + k] = m->lane; ofresh[i * K + k] = m->fresh;
        } else { o[0] = o[1] = o[2] = o[3] = 0.0; olane[i * K + k] = 0; ofresh[i * K + k] = 0; }
    }
}

static void *ma_step_range(void *arg)
{
    ma_job *J = (ma_job *)arg;
    const int A = J->A, K = J->K, D = 4 * A + 14;
    const int cont = (J->hwflags & HW_FLAG_CONTINUOUS) != 0, inf_obs = !(J->hwflags & HW_FLAG_NO_INF_OBS);
    const int autoreset = !(J->hwflags & HW_FLAG_NO_AUTORESET), quant = (J->hwflags & HW_FLAG_QUANTISE_F32) != 0;
    const int shared = (J->hwflags & HW_FLAG_SHARED_REWARD) != 0;
    for (int i = J->i0; i < J->i1; ++i) {
        ma_env e;
        ma_unpack(J, i, &e);
        double r[MA_MAX_A]; uint8_t d[MA_MAX_A]; int used = 0;
        /* is_done persists until reset() (highway_core.py:239, :420): a car that finished in an earlier step is still done */
        for (int a = 0; a < A; ++a) { r[a] = 0.0; d[a] = (uint8_t)((J->done_mask ? J->done_mask[i] >> a : 0) & 1); }
        const int32_t *ad = J->act_d ? J->act_d + (size_t)i * A : 0;
        const double *ac = J->act_c ? J->act_c + (size_t)i * A * 2 : 0;
        for (int t = 0; t < J->ticks_per_step; ++t) {
            ma_tick(&e, J->dt, J->timeout_tick, cont, inf_obs, ad, ac, J->seed, J->env_id0 + (uint64_t)i,
                    J->u_override ? J->u_override + (size_t)i * J->u_stride : 0, &used, r, d);
            int any = 0; for (int a = 0; a < A; ++a) any |= d[a];
            if (any) break; /* highway_env.py:85-86 */
        }
        if (J->u_used_out) J->u_used_out[i] = used;
        if (shared) { double s = 0; for (int a = 0; a < A; ++a) s += r[a]; for (int a = 0; a < A; ++a) r[a] = s; } /* np.sum, :92-94 */
        int any = 0;
        for (int a = 0; a < A; ++a) { J->rew[(size_t)i * A + a] = r[a]; J->done[(size_t)i * A + a] = d[a]; any |= d[a]; }
        /* Monitor: sum over agents of the float32 VecEnv rewards is not what it sees; it sees the python floats */
        for (int a = 0; a < A; ++a) J->ep_ret[i] += r[a];
        J->ep_len[i] += 1;

        if (J->next_car) ma_pack(&e, J->next_car, J->next_meta, J->next_meta + (size_t)J->n * A, J->next_obst,
                                  J->next_ometa, J->next_ometa + (size_t)J->n * K, A, K, (size_t)i);
        if (any) {
            if (J->term) { double *tm = J->term + 5 * (size_t)i; tm[0] = e.tick; tm[1] = e.nc_obs; tm[2] = e.nc_agent; tm[3] = J->ep_ret[i]; tm[4] = J->ep_len[i]; }
            if (autoreset) { ma_reset(&e); J->ep_ret[i] = 0.0; J->ep_len[i] = 0; }
        }
        if (J->done_mask) { int m = 0; if (!(any && autoreset)) for (int a = 0; a < A; ++a) m |= d[a] << a; J->done_mask[i] = m; }
        e.overflow |= ma_observe(&e, J->obs + (size_t)i * A * D) ? 0x10000 : 0;
        if (quant) {
            for (int a = 0; a < A; ++a) {
                ma_car *m = &e.car[a];
                m->px = q32(m->px); m->py = q32(m->py); m->rx = q32(m->rx); m->vx = q32(m->vx); m->vy = q32(m->vy);
                m->acc = q32(m->acc); m->steer = q32(m->steer); m->cruise = q32(m->cruise);
            }
            for (int k = 0; k < e.nobs; ++k) { ma_obs *m = &e.ob[k]; m->px = q32(m->px); m->rx = q32(m->rx); m->vx = q32(m->vx); m->iv = q32(m->iv); }
        }
        ma_pack(&e, J->car, J->lane, J->cflags, J->obst, J->olane, J->ofresh, A, K, (size_t)i);
        J->nobs[i] = e.nobs; J->tick[i] = e.tick; J->leader[i] = e.leader;
        for (int l = 0; l < 3; ++l) J->newest[3 * (size_t)i + l] = e.newest[l];
        J->nc_obs[i] = e.nc_obs; J->nc_agent[i] = e.nc_agent; J->spawn_ctr[i] = e.spawn_ctr; J->overflow[i] = e.overflow;
    }
    return 0;
}

/* Step n multi-agent environments.  Canonical arrays (row i = environment i):
 *   car f64[n][A][8] = px,py,rx,vx,vy,acc,steer,cruise   lane,cflags i32[n][A]   tick,leader,nobs i32[n]
 *   obst f64[n][K][4] = px,rx,vx,init_vx (slots 0..nobs-1, insertion order)   olane,ofresh i32[n][K]   newest i32[n][3]
 *   nc_obs,nc_agent i32[n]  spawn_ctr u32[n]  ep_ret f64[n]  ep_len i32[n]  overflow i32[n]
 *   act_d i32[n][A] | act_c f64[n][A][2]
 * outputs: obs f32[n][A][4A+14], rew f64[n][A], done u8[n][A], term f64[n][5] = tick, nc_obs, nc_agent, ep_ret, ep_len
 *   next_* (optional): pre-auto-reset successor (car; meta = lane[n][A] then cflags[n][A]; obst; ometa = olane then ofresh)
 */
int hw_oracle_ma_step(int n, int A, int K, double *car, int32_t *lane, int32_t *cflags, int32_t *tick, int32_t *leader,
                      int32_t *nobs, double *obst, int32_t *olane, int32_t *ofresh, int32_t *newest,
                      int32_t *nc_obs, int32_t *nc_agent, uint32_t *spawn_ctr, double *ep_ret, int32_t *ep_len, int32_t *overflow,
                      int32_t *done_mask /* nullable: bit a = car a is done since an earlier step (no auto-reset) */,
                      const int32_t *act_d, const double *act_c, float *obs, double *rew, uint8_t *done, double *term,
                      double dt, int ticks_per_step, int timeout_tick, uint64_t seed, uint64_t env_id0, int hwflags,
                      const double *u_override, int u_stride, int32_t *u_used_out,
                      double *next_car, int32_t *next_meta, double *next_obst, int32_t *next_ometa, int nthreads)
{
    if (A < 1 || A > MA_MAX_A || K < 3 || K > MA_MAX_K) return -1;
    ma_job base = { n, A, K, car, lane, cflags, tick, leader, nobs, obst, olane, ofresh, newest, nc_obs, nc_agent, spawn_ctr,
                    ep_ret, ep_len, overflow, done_mask, act_d, act_c, obs, rew, done, term, dt, ticks_per_step, timeout_tick, seed, env_id0,
                    hwflags, u_override, u_stride, u_used_out, next_car, next_meta, next_obst, next_ometa, 0, n };
    if (nthreads <= 1 || n < 2 * nthreads) { ma_step_range(&base); return 0; }
    if (nthreads > 256) nthreads = 256;
    pthread_t th[256]; ma_job jobs[256];
    for (int t = 0; t < nthreads; ++t) {
        jobs[t] = base;
        jobs[t].i0 = (int)((int64_t)n * t / nthreads);
        jobs[t].i1 = (int)((int64_t)n * (t + 1) / nthreads);
        if (pthread_create(&th[t], 0, ma_step_range, &jobs[t]) != 0) return -1;
    }
    for (int t = 0; t < nthreads; ++t) pthread_join(th[t], 0);
    return 0;
}

int hw_oracle_ma_reset(int n, int A, int K, double *car, int32_t *lane, int32_t *cflags, int32_t *tick, int32_t *leader,
                       int32_t *nobs, double *obst, int32_t *olane, int32_t *ofresh, int32_t *newest,
                       int32_t *nc_obs, int32_t *nc_agent, double *ep_ret, int32_t *ep_len, int32_t *overflow,
                       const uint8_t *mask, float *obs)
{
    if (A < 1 || A > MA_MAX_A || K < 3 || K > MA_MAX_K) return -1;
    const int D = 4 * A + 14;
    for (int i = 0; i < n; ++i) {
        if (mask && !mask[i]) continue;
        ma_env e; memset(&e, 0, sizeof e);
        e.A = A; e.K = K;
        ma_reset(&e);
        ma_pack(&e, car, lane, cflags, obst, olane, ofresh, A, K, (size_t)i);
        nobs[i] = e.nobs; tick[i] = 0; leader[i] = 0;
        for (int l = 0; l < 3; ++l) newest[3 * (size_t)i + l] = e.newest[l];
        nc_obs[i] = 0; nc_agent[i] = 0; ep_ret[i] = 0.0; ep_len[i] = 0; overflow[i] = 0;
        if (obs) ma_observe(&e, obs + (size_t)i * A * D);
    }
    return 0;
}
