/* CPU restatement (fp64, scalar C) of gym-highway's SINGLE-AGENT step path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the checker for the CUDA kernels in
 * gym_highway_b200/csrc; it may be imported/linked/executed only by tests/, by
 * __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs.  The product path
 * never routes through it.
 *
 * Parity pin: the reference holds no tests or golden vectors for this path ("parity unpinned" by
 * the reference itself).  The pin used instead is the reference's own
 * unmodified Python code executed in the build container under oracle/shim (pygame/gym stand-ins):
 * oracle/gen_golden.py writes its transitions to tests/golden/ and tests/test_oracle_golden.py
 * requires this file to reproduce them bit-for-bit.
 *
 * Reference lines followed (all under /root/reference/gym_highway/envs/):
 *   constants            multi_lane_sim.py:20-45
 *   reset                highway_env.py:53-68, multi_lane_sim.py:826-877
 *   apply action         multi_lane_sim.py:502-526 (+ helpers :529-572)
 *   car update           multi_lane_sim.py:140-183 (continuous), :185-229 (discrete), :234-272
 *   obstacle update      multi_lane_sim.py:324-347
 *   tick driver          multi_lane_sim.py:879-1049
 *   env.step / obs       highway_env.py:70-140, highway_env_cont.py:42-55, multi_lane_sim.py:1051-1092
 *
 * Every binary operation is rounded separately (compile with -ffp-contract=off), mirroring CPython
 * float arithmetic.  tan/ceil come from the same libm CPython uses.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <string.h>
#include "philox.h"
#include "tan_small.h"

int hw_oracle_tan_poly = 0; /* 0: libm tan (the reference's); 1: the kernels' polynomial, for diagnosing soak mismatches (tan_small.h) */
void hw_oracle_set_tan(int poly) { hw_oracle_tan_poly = poly; }

#define HW_PI 3.14159265358979323846
static const double DEG2RAD = HW_PI / 180.0; /* math.radians: x * (pi/180) */
static const double RAD2DEG = 180.0 / HW_PI; /* math.degrees: x * (180/pi) */
static const double LANE_C[3] = { 1.25, 3.75, 6.25 };

enum { ACT_LEFT = 0, ACT_RIGHT = 1, ACT_ACCELERATE = 2, ACT_MAINTAIN = 3 };
enum { F_LEFT = 1, F_RIGHT = 2, F_DO_ACC = 4, F_DO_MAINT = 8 };
enum { HW_FLAG_CONTINUOUS = 1, HW_FLAG_NO_INF_OBS = 2, HW_FLAG_NO_AUTORESET = 4, HW_FLAG_QUANTISE_F32 = 8 };

typedef struct {
    double dt;          /* 1/36 (or 1/60 when real_time) */
    int ticks_per_step; /* int(ticks/4): 9 (or 15) */
    int timeout_tick;   /* first tick whose accumulated run_time >= 60: 2160 (or 3600) */
} hw_sim_params;

/* python: max(lo, min(x, hi)) — min returns x unless hi < x; max returns lo unless inner > lo */
static inline double py_clamp(double x, double lo, double hi)
{
    double inner = (hi < x) ? hi : x;
    return (inner > lo) ? inner : lo;
}

typedef struct {
    double px, py, rx, vx, acc, steer, cruise;
    int lane, flags, tick;
    double opx[3], orx[3], ovx[3], oiv[3];
    int ncoll;
    uint32_t spawn_ctr;
} sa_env;

static void sa_reset(sa_env *e)
{
    /* highway_env.py:55-65 start grid */
    e->px = 20.0; e->py = LANE_C[1]; e->rx = 20.0; e->vx = 0.0;
    e->acc = 0.0; e->steer = 0.0; e->cruise = 0.0;
    e->lane = 2; e->flags = 0; e->tick = 0;
    static const double ox[3] = { -20.0, -25.0, -40.0 }, ov[3] = { 7.0, 5.0, 6.0 };
    for (int k = 0; k < 3; ++k) { e->opx[k] = ox[k]; e->orx[k] = ox[k]; e->ovx[k] = ov[k]; e->oiv[k] = ov[k]; }
    e->ncoll = 0;
}

static inline int trunc_i(double v) { return (int)v; } /* pygame Rect setter: C cast, toward zero */

static inline int rects_overlap(int ax, int ay, int bx, int by)
{
    return ax < bx + 76 && ay < by + 38 && ax + 76 > bx && ay + 38 > by;
}

/* one simulator tick; returns the tick reward, sets *done */
static double sa_tick(sa_env *e, const hw_sim_params *P, int cont, int inf_obs, int act_d, double a0, double a1,
                      uint64_t seed, uint64_t env_id, const double *u_override, int *u_used, int *done)
{
    const double dt = P->dt;
    const int lock = (e->tick == 0); /* collision_count_lock is True only on the first tick after reset */
    e->tick += 1;
    *done = e->tick >= P->timeout_tick; /* run_time >= 60, multi_lane_sim.py:913 */

    /* --- apply action (multi_lane_sim.py:502-526) --- */
    if (cont) {
        e->acc = e->acc + a0 * 5.0 * dt;
        e->acc = py_clamp(e->acc, -5.0, 5.0);
        e->steer = e->steer + a1 * 30.0 * dt;
        e->steer = py_clamp(e->steer, -30.0, 30.0);
    } else {
        if (act_d == ACT_ACCELERATE && !(e->flags & F_DO_ACC)) {
            e->flags = (e->flags | F_DO_ACC) & ~F_DO_MAINT;
        } else if (act_d == ACT_MAINTAIN && !(e->flags & F_DO_MAINT)) {
            int f = -1;
            for (int k = 0; k < 3; ++k) { /* obstacle k drives in lane k+1 */
                if (k + 1 == e->lane && e->opx[k] > e->px) {
                    if (f < 0) f = k;
                    else if (e->opx[k] < e->opx[f]) f = k;
                }
            }
            e->cruise = (f >= 0) ? e->ovx[f] : e->vx;
            e->flags = (e->flags | F_DO_MAINT) & ~F_DO_ACC;
        }
        e->acc = py_clamp(e->acc, -5.0, 5.0);
        if (act_d == ACT_RIGHT && !(e->flags & F_RIGHT)) {
            e->lane = e->lane + 1 < 3 ? e->lane + 1 : 3;
            e->flags = (e->flags | F_RIGHT) & ~F_LEFT;
        } else if (act_d == ACT_LEFT && !(e->flags & F_LEFT)) {
            e->lane = e->lane - 1 > 1 ? e->lane - 1 : 1;
            e->flags = (e->flags | F_LEFT) & ~F_RIGHT;
        }
        e->steer = py_clamp(e->steer, -30.0, 30.0);
    }

    /* --- collision test on the rects the previous tick left (multi_lane_sim.py:920-937) --- */
    double reward = 0.0;
    if (!lock) {
        int ax = trunc_i(e->px * 32.0 - 38.0), ay = trunc_i(e->py * 32.0 - 19.0);
        int n = 0;
        for (int k = 0; k < 3; ++k) {
            int bx = trunc_i(e->opx[k] * 32.0 - 38.0), by = trunc_i(LANE_C[k] * 32.0 - 19.0);
            n += rects_overlap(ax, ay, bx, by);
        }
        e->ncoll += n;
        reward -= (double)n * 250.0;
        if (reward < 0) *done = 1;
    }

    /* --- car update --- */
    if (!cont) {
        if (e->flags & F_DO_ACC) { /* :248-255 */
            if (e->acc < 0.0) e->acc = 0.0; else e->acc = e->acc + 1 * dt;
            if (e->acc == 5.0) e->flags &= ~F_DO_ACC;
        } else if (e->flags & F_DO_MAINT) { /* :257-272 */
            double vc = ceil(e->vx), cc = ceil(e->cruise);
            int is_speed = vc <= cc;
            e->acc = is_speed ? 5.0 : -5.0;
            int is_cruise = (is_speed && vc >= cc) || (!is_speed && vc <= cc);
            if (is_cruise) { e->vx = e->cruise; e->acc = 0.0; e->flags &= ~F_DO_MAINT; }
        }
    }
    e->vx = e->vx + e->acc * dt;
    e->vx = py_clamp(e->vx, 0.0, 20.0);
    if (!cont) {
        double ty = (double)(80 * e->lane - 40.0) / 32.0; /* (LANE_WIDTH*lane - LANE_WIDTH/2)/ppu */
        if (e->flags & F_LEFT) {
            e->steer = e->steer + 30.0 * dt;
            if (e->py <= ty) { e->steer = 0.0; e->flags &= ~F_LEFT; }
        } else if (e->flags & F_RIGHT) {
            e->steer = e->steer - 30.0 * dt;
            if (e->py >= ty) { e->steer = 0.0; e->flags &= ~F_RIGHT; }
        }
    }
    double ang = 0.0;
    if (e->steer != 0.0) {
        double radius = 4.0 / hw_tan(e->steer * DEG2RAD);
        ang = e->vx / radius;
    }
    e->px = e->px + e->vx * dt; /* position += velocity*dt; velocity.y == 0 in the single-agent sim */
    e->py = e->py + 0.0 * dt;
    if (cont) e->py = e->py - ang * dt;
    else      e->py = e->py - (ang * RAD2DEG) * dt * dt;
    e->px = 10.0; /* the only car is always the leader (:212-213) */
    if (cont) {
        if (e->py < 1.0) e->py = (e->py > 1.0) ? e->py : 1.0;
        else if (e->py > 6.0) e->py = (7.0 < e->py) ? 7.0 : e->py;
    } else {
        if (e->py < 0.0) e->py = (0.0 > e->py) ? 0.0 : e->py;
        else if (e->py > 6.0) e->py = (7.0 < e->py) ? 7.0 : e->py;
    }
    e->rx = e->rx + e->vx * dt;
    if (cont) { /* lane id = nearest lane centre, first minimum (:178-179) */
        double d0 = fabs(e->py - LANE_C[0]), d1 = fabs(e->py - LANE_C[1]), d2 = fabs(e->py - LANE_C[2]);
        double m = d0; if (d1 < m) m = d1; if (d2 < m) m = d2;
        e->lane = (d0 == m) ? 1 : (d1 == m) ? 2 : 3;
    }

    /* --- obstacle update (:324-347), sees the leader's post-update lane/position/velocity --- */
    for (int k = 0; k < 3; ++k) {
        double v = py_clamp(e->oiv[k], -20.0, 20.0);
        if (k + 1 == e->lane && e->opx[k] < e->px) v = (e->vx < e->oiv[k]) ? e->vx : e->oiv[k]; /* min(init, leader) */
        e->ovx[k] = v;
        e->opx[k] = e->opx[k] + v * dt;
        e->opx[k] = e->opx[k] - e->vx * dt;
        e->orx[k] = e->orx[k] + v * dt;
    }

    if (reward == 0.0) reward = e->vx / 20.0 - 1.0;

    /* --- respawn (:963-991) --- */
    if (inf_obs) {
        for (int k = 0; k < 3; ++k) {
            if (e->opx[k] < -76.0 / 32.0) {
                double u1, u2;
                if (u_override) { u1 = u_override[*u_used]; u2 = u_override[*u_used + 1]; *u_used += 2; }
                else hw_spawn_uniforms(seed, env_id, e->spawn_ctr, &u1, &u2);
                e->spawn_ctr += 1;
                double x = 70.0 + 30.0 * u1; /* random.uniform(70,100) = a + (b-a)*random() */
                double v = 5.0 + 2.0 * u2;   /* random.uniform(5,7) */
                e->opx[k] = x;
                e->orx[k] = e->rx + (x - e->px);
                e->ovx[k] = v;
                e->oiv[k] = v;
            }
        }
    }
    return reward;
}

/* python round(x, 3): correctly rounded to 3 decimals (ties on the exact binary value -> even),
 * then the nearest double of k/1000.  x*1000 is exact in x87 long double (53+10 <= 64 bits). */
static double py_round3(double x, long *k_out)
{
    long double p = (long double)x * 1000.0L;
    long k = (long)rintl(p);
    if (k_out) *k_out = k;
    if (k == 0) return copysign(0.0, x); /* round(-0.0004, 3) == -0.0 in CPython */
    return (double)k / 1000.0;
}

static void sa_observe(const sa_env *e, float *ob)
{
    /* multi_lane_sim.py:1076-1092 -> float32; highway_env.py:131-140 clip + normalise in float32 */
    double raw[18];
    raw[0] = e->acc; raw[1] = e->steer; raw[2] = e->rx; raw[3] = e->py;
    for (int k = 0; k < 3; ++k) {
        raw[4 + 2 * k] = e->orx[k] - e->rx;
        raw[5 + 2 * k] = LANE_C[k] - e->py;
        raw[12 + 2 * k] = e->ovx[k] - e->vx;
        raw[13 + 2 * k] = 0.0 - 0.0;
    }
    raw[10] = e->vx; raw[11] = 0.0;
    for (int j = 0; j < 18; ++j) {
        float lo, hi;
        if (j == 0) { lo = -5.f; hi = 5.f; }
        else if (j == 1) { lo = -30.f; hi = 30.f; }
        else if (j == 2) { lo = 0.f; hi = 1200.f; }
        else if (j == 3) { lo = 0.f; hi = 8.f; }
        else if (j < 10) { if (j & 1) { lo = -8.f; hi = 8.f; } else { lo = -60.f; hi = 60.f; } }
        else { lo = -20.f; hi = 20.f; }
        float v = (float)raw[j];
        v = v < lo ? lo : v;
        v = v > hi ? hi : v;
        volatile float num = v - lo, den = hi - lo;
        ob[j] = num / den;
    }
}

static inline double q32(double v) { return (double)(float)v; }

typedef struct {
    int n; double *car; int32_t *lane, *flags, *tick; double *obst;
    int32_t *ncoll; uint32_t *spawn_ctr; int64_t *ep_ret_milli; int32_t *ep_len;
    const int32_t *act_d; const double *act_c;
    float *obs; double *rew; uint8_t *done; int64_t *term;
    hw_sim_params P; uint64_t seed, env_id0; int hwflags;
    const double *u_override; int u_stride; int32_t *u_used_out;
    double *next_car; int32_t *next_lft; double *next_obst;
    int i0, i1;
} sa_step_job;

static void *sa_step_range(void *arg)
{
    sa_step_job *J = (sa_step_job *)arg;
    double *car = J->car; int32_t *lane = J->lane, *flags = J->flags, *tick = J->tick; double *obst = J->obst;
    int32_t *ncoll = J->ncoll; uint32_t *spawn_ctr = J->spawn_ctr; int64_t *ep_ret_milli = J->ep_ret_milli; int32_t *ep_len = J->ep_len;
    const int32_t *act_d = J->act_d; const double *act_c = J->act_c;
    float *obs = J->obs; double *rew = J->rew; uint8_t *done = J->done; int64_t *term = J->term;
    const hw_sim_params P = J->P; const uint64_t seed = J->seed, env_id0 = J->env_id0; const int hwflags = J->hwflags;
    const double *u_override = J->u_override; const int u_stride = J->u_stride; int32_t *u_used_out = J->u_used_out;
    double *next_car = J->next_car; int32_t *next_lft = J->next_lft; double *next_obst = J->next_obst;
    const int cont = (hwflags & HW_FLAG_CONTINUOUS) != 0;
    const int inf_obs = !(hwflags & HW_FLAG_NO_INF_OBS);
    const int autoreset = !(hwflags & HW_FLAG_NO_AUTORESET);
    const int quant = (hwflags & HW_FLAG_QUANTISE_F32) != 0;
    for (int i = J->i0; i < J->i1; ++i) {
        sa_env e;
        const double *c = car + 7 * (size_t)i;
        e.px = c[0]; e.py = c[1]; e.rx = c[2]; e.vx = c[3]; e.acc = c[4]; e.steer = c[5]; e.cruise = c[6];
        e.lane = lane[i]; e.flags = flags[i]; e.tick = tick[i];
        for (int k = 0; k < 3; ++k) {
            const double *o = obst + ((size_t)i * 3 + k) * 4;
            e.opx[k] = o[0]; e.orx[k] = o[1]; e.ovx[k] = o[2]; e.oiv[k] = o[3];
        }
        e.ncoll = ncoll[i]; e.spawn_ctr = spawn_ctr[i];

        double r = 0.0; int d = 0, used = 0;
        double a0 = 0, a1 = 0; int ad = 0;
        if (cont) { a0 = act_c[2 * (size_t)i]; a1 = act_c[2 * (size_t)i + 1]; } else ad = act_d[i];
        for (int t = 0; t < P.ticks_per_step; ++t) {
            r = sa_tick(&e, &P, cont, inf_obs, ad, a0, a1, seed, env_id0 + (uint64_t)i,
                        u_override ? u_override + (size_t)i * u_stride : 0, &used, &d);
            if (r < -1.0) break; /* highway_env.py:106-108 */
        }
        if (u_used_out) u_used_out[i] = used;
        long k_milli;
        rew[i] = py_round3(r, &k_milli);
        done[i] = (uint8_t)d;
        ep_ret_milli[i] += k_milli;
        ep_len[i] += 1;

        if (next_car) {
            double *nc = next_car + 7 * (size_t)i;
            nc[0] = e.px; nc[1] = e.py; nc[2] = e.rx; nc[3] = e.vx; nc[4] = e.acc; nc[5] = e.steer; nc[6] = e.cruise;
            next_lft[3 * (size_t)i] = e.lane; next_lft[3 * (size_t)i + 1] = e.flags; next_lft[3 * (size_t)i + 2] = e.tick;
            for (int k = 0; k < 3; ++k) {
                double *o = next_obst + ((size_t)i * 3 + k) * 4;
                o[0] = e.opx[k]; o[1] = e.orx[k]; o[2] = e.ovx[k]; o[3] = e.oiv[k];
            }
        }
        if (d) {
            if (term) { int64_t *tm = term + 4 * (size_t)i; tm[0] = e.tick; tm[1] = e.ncoll; tm[2] = ep_ret_milli[i]; tm[3] = ep_len[i]; }
            if (autoreset) { sa_reset(&e); ep_ret_milli[i] = 0; ep_len[i] = 0; }
        }
        sa_observe(&e, obs + 18 * (size_t)i); /* from the fp64 values, like the reference */
        if (quant) { /* fp32 storage between steps (what the kernels keep in HBM) */
            e.px = q32(e.px); e.py = q32(e.py); e.rx = q32(e.rx); e.vx = q32(e.vx);
            e.acc = q32(e.acc); e.steer = q32(e.steer); e.cruise = q32(e.cruise);
            for (int k = 0; k < 3; ++k) { e.opx[k] = q32(e.opx[k]); e.orx[k] = q32(e.orx[k]); e.ovx[k] = q32(e.ovx[k]); e.oiv[k] = q32(e.oiv[k]); }
        }

        double *cw = car + 7 * (size_t)i;
        cw[0] = e.px; cw[1] = e.py; cw[2] = e.rx; cw[3] = e.vx; cw[4] = e.acc; cw[5] = e.steer; cw[6] = e.cruise;
        lane[i] = e.lane; flags[i] = e.flags; tick[i] = e.tick;
        for (int k = 0; k < 3; ++k) {
            double *o = obst + ((size_t)i * 3 + k) * 4;
            o[0] = e.opx[k]; o[1] = e.orx[k]; o[2] = e.ovx[k]; o[3] = e.oiv[k];
        }
        ncoll[i] = e.ncoll; spawn_ctr[i] = e.spawn_ctr;
    }
    return 0;
}

/* Step n environments (nthreads > 1: contiguous ranges on pthreads).  All arrays are canonical
 * (unpacked) host arrays:
 *   car  f64[n][7] = px,py,rx,vx,acc,steer,cruise     lane,flags,tick  i32[n]
 *   obst f64[n][3][4] = px,rx,vx,init_vx (slot k = lane k+1)
 *   ncoll i32[n]  spawn_ctr u32[n]  ep_ret_milli i64[n]  ep_len i32[n]
 *   act_d i32[n] | act_c f32-representable f64[n][2]
 * outputs: obs f32[n][18], rew f64[n] (rounded to 3 dp), done u8[n],
 *   term i64[n][4] = tick, ncoll, ep_ret_milli, ep_len of an episode that ended this step
 *   next_* (optional) = pre-auto-reset successor state (car f64[n][7], lane/flags/tick i32[n][3], obst)
 * u_override (optional): f64[n][u_stride] explicit uniforms consumed in draw order instead of Philox.
 */
int hw_oracle_sa_step(int n, double *car, int32_t *lane, int32_t *flags, int32_t *tick, double *obst,
                      int32_t *ncoll, uint32_t *spawn_ctr, int64_t *ep_ret_milli, int32_t *ep_len,
                      const int32_t *act_d, const double *act_c,
                      float *obs, double *rew, uint8_t *done, int64_t *term,
                      double dt, int ticks_per_step, int timeout_tick,
                      uint64_t seed, uint64_t env_id0, int hwflags,
                      const double *u_override, int u_stride, int32_t *u_used_out,
                      double *next_car, int32_t *next_lft, double *next_obst, int nthreads)
{
    sa_step_job base = { n, car, lane, flags, tick, obst, ncoll, spawn_ctr, ep_ret_milli, ep_len, act_d, act_c,
                         obs, rew, done, term, { dt, ticks_per_step, timeout_tick }, seed, env_id0, hwflags,
                         u_override, u_stride, u_used_out, next_car, next_lft, next_obst, 0, n };
    if (nthreads <= 1 || n < 2 * nthreads) { sa_step_range(&base); return 0; }
    if (nthreads > 256) nthreads = 256;
    pthread_t th[256]; sa_step_job jobs[256];
    for (int t = 0; t < nthreads; ++t) {
        jobs[t] = base;
        jobs[t].i0 = (int)((int64_t)n * t / nthreads);
        jobs[t].i1 = (int)((int64_t)n * (t + 1) / nthreads);
        if (pthread_create(&th[t], 0, sa_step_range, &jobs[t]) != 0) return -1;
    }
    for (int t = 0; t < nthreads; ++t) pthread_join(th[t], 0);
    return 0;
}

/* reset every env whose mask byte is non-zero (mask == NULL: all); writes the reset observation */
int hw_oracle_sa_reset(int n, double *car, int32_t *lane, int32_t *flags, int32_t *tick, double *obst,
                       int32_t *ncoll, int64_t *ep_ret_milli, int32_t *ep_len, const uint8_t *mask, float *obs)
{
    for (int i = 0; i < n; ++i) {
        if (mask && !mask[i]) continue;
        sa_env e; memset(&e, 0, sizeof e);
        sa_reset(&e);
        double *cw = car + 7 * (size_t)i;
        cw[0] = e.px; cw[1] = e.py; cw[2] = e.rx; cw[3] = e.vx; cw[4] = e.acc; cw[5] = e.steer; cw[6] = e.cruise;
        lane[i] = e.lane; flags[i] = e.flags; tick[i] = e.tick;
        for (int k = 0; k < 3; ++k) {
            double *o = obst + ((size_t)i * 3 + k) * 4;
            o[0] = e.opx[k]; o[1] = e.orx[k]; o[2] = e.ovx[k]; o[3] = e.oiv[k];
        }
        ncoll[i] = 0; ep_ret_milli[i] = 0; ep_len[i] = 0;
        if (obs) sa_observe(&e, obs + 18 * (size_t)i);
    }
    return 0;
}

/* observation of the current state only */
int hw_oracle_sa_observe(int n, const double *car, const double *obst, float *obs)
{
    for (int i = 0; i < n; ++i) {
        sa_env e; memset(&e, 0, sizeof e);
        const double *c = car + 7 * (size_t)i;
        e.px = c[0]; e.py = c[1]; e.rx = c[2]; e.vx = c[3]; e.acc = c[4]; e.steer = c[5]; e.cruise = c[6];
        for (int k = 0; k < 3; ++k) {
            const double *o = obst + ((size_t)i * 3 + k) * 4;
            e.opx[k] = o[0]; e.orx[k] = o[1]; e.ovx[k] = o[2]; e.oiv[k] = o[3];
        }
        sa_observe(&e, obs + 18 * (size_t)i);
    }
    return 0;
}

void hw_oracle_spawn_uniforms(uint64_t seed, uint64_t env_id, uint32_t spawn_ctr, double *out2)
{
    hw_spawn_uniforms(seed, env_id, spawn_ctr, &out2[0], &out2[1]);
}
