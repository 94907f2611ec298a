/* highway_b200 — C ABI of the B200-native batched gym-highway step path.
 *
 * The reference (sahilPereira/gym-highway) is pure Python and has no FFI of its own; the seam a
 * maintainer would bind is the per-env gym API and the baselines VecEnv API above it:
 *   reset()      gym_highway/envs/highway_env.py:121-129, multiagent_envs/highway_env.py:107-112,
 *                baselines/common/vec_env/dummy_vec_env.py:59-63
 *   step()       gym_highway/envs/highway_env.py:70-114, highway_env_cont.py:42-55,
 *                multiagent_envs/highway_env.py:63-99, highway_ma_c_env.py:56-82,
 *                baselines/common/vec_env/dummy_vec_env.py:46-57 (auto-reset on done),
 *                dummy_vec_ma_env.py:66-79 (reset when any agent is done)
 *   get_state()  gym_highway/envs/multi_lane_sim.py:1051-1092, multiagent_envs/simple_base.py:82-154
 *   Monitor      baselines/bench/monitor.py:58-79 (episode return / length on the terminal step)
 * Each entry point below replaces one of those calls for a whole batch of environments; the ctypes
 * binding the Python host uses is gym_highway_b200/_lib.py, and INTEGRATION.md shows the stub a
 * reference maintainer would add.
 *
 * Conventions: plain pointers and sizes only.  Unless a parameter says HOST, every pointer is a
 * DEVICE pointer into caller-owned memory (the library allocates nothing and keeps no global
 * state).  Calls are asynchronous with respect to the host and ordered on `stream` (a cudaStream_t
 * passed as void*; NULL = default stream), except the *_host calls which return after the stream
 * has drained.  Return value: 0 on success, a cudaError_t (> 0) if the launch failed, or one of the
 * negative HW_E_* codes for bad arguments.  Re-entrant for disjoint states; no host threads.
 */
#ifndef HIGHWAY_B200_H
#define HIGHWAY_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HW_ABI_VERSION 3

#define HW_E_BADARG (-1)   /* null pointer / non-positive size / unsupported combination */
#define HW_E_ALIGN  (-2)   /* a 16-byte-aligned pointer was required */

/* step flags */
#define HW_FLAG_CONTINUOUS   1  /* actions are float[.][2] in [-1,1] instead of int32 in {0..3} */
#define HW_FLAG_NO_INF_OBS   2  /* inf_obs=False: obstacles are never respawned */
#define HW_FLAG_NO_AUTORESET 4  /* leave terminal states in place (gym per-env semantics) */
#define HW_FLAG_SHARED_REWARD 16 /* multi-agent: every car receives the sum of the rewards (shared_reward=True) */
#define HW_FLAG_EXACT_STEERING 32 /* evaluate vx / (4 / tan) with the reference's two divisions instead of vx*tan/4
                                    * (slower; floats then match the fp64 reference bit for bit after fp32 rounding) */
#define HW_FLAG_MA_THREAD_PER_ENV 64  /* multi-agent step: force the one-thread-per-environment kernel (cars in registers, hw_ma_tpe.cu) */
#define HW_FLAG_MA_LANE_PER_CAR  128  /* multi-agent step: force the lane-per-car kernel (hw_ma_grp.cu).  Neither flag: chosen by batch
                                       * size - the thread-per-environment kernel from HW_MA_TPE_MIN_ENVS environments up for five cars
                                       * (a third / two thirds of that for three / four cars, always for one or two: fewer instructions
                                       * per environment), the lane-per-car kernel below (five times the threads for the smallest
                                       * batches).  Same results either way. */
#define HW_MA_TPE_MIN_ENVS 16384
#define HW_FLAG_COMPACT_OBS 256  /* single-agent hw_sa_step / hw_sa_step_host: observation rows of HW_SA_OBS_DIM_COMPACT = 14 floats instead of 18 -
                                  * columns 11, 13, 15, 17 (the lateral velocities of the car and the three obstacles: always zero, i.e. the
                                  * constant 0.5 after normalisation, multi_lane_sim.py:1062-1090) are dropped: row = columns 0..9, then
                                  * {10, 12, 14, 16}.  22 % fewer bytes over PCIe for a host consumer; HighwayVecEnv.expand_obs() restores
                                  * the reference layout.  `obs` buffers are then [n][14]. */

/* discrete actions, gym_highway/envs/highway_env.py:14-20 */
#define HW_ACT_LEFT 0
#define HW_ACT_RIGHT 1
#define HW_ACT_ACCELERATE 2
#define HW_ACT_MAINTAIN 3

#define HW_SA_OBS_DIM 18
#define HW_SA_OBS_DIM_COMPACT 14
#define HW_MA_MAX_AGENTS 5   /* the reference's start grid holds five cars (simple_base.py:32-38) */
#define HW_MA_MAX_AGENTS_EXT 8 /* EXTENSION beyond the reference (no reference parity, the CPU restatement is its only
                                * specification): 6..8 cars on a start grid that continues the reference's staggered rows
                                * backwards - (2.5, lane 2), (0, lane 3), (-2.5, lane 1).  Stepped by the thread-per-environment kernel only. */
#define HW_MA_MAX_SLOTS 16   /* live obstacles per environment (three at reset, typically <= 8) */
#define HW_MA_OBS_DIM(A) (4 * (A) + 14)

/* bit layout of the packed per-car word (HwSaState.car_b[i].w, HwMaState car word 8) */
#define HW_BITS_TICK_MASK   0xFFFFu  /* single-agent only: simulator ticks since reset */
#define HW_BITS_LANE_SHIFT  16       /* lane id 1..3 */
#define HW_BIT_LEFT         (1u << 18)
#define HW_BIT_RIGHT        (1u << 19)
#define HW_BIT_DO_ACC       (1u << 20)
#define HW_BIT_DO_MAINT     (1u << 21)

/* simulation clock, derived exactly as the reference does (multi_lane_sim.py:358,371,888,910-913):
 * dt = 1/ticks, ticks_per_step = int(ticks/4), timeout_tick = first tick whose accumulated
 * run_time >= 60 s.  (36 Hz: 1/36, 9, 2160; real_time 60 Hz: 1/60, 15, 3600.) */
typedef struct HwSimParams {
    double dt;
    int32_t ticks_per_step;
    int32_t timeout_tick;
} HwSimParams;

/* Single-agent state, structure of arrays, 96 bytes per environment, every array 16-byte aligned.
 *   car_a[i] = {px, py, raw_x, vx}                      float4
 *   car_b[i] = {acc, steer, cruise_vel, bits}           float4 (w reinterpreted as uint32)
 *   obst[k*n + i] = {px, raw_x, vx, init_vx}            float4, k = lane-1 (the lane's newest obstacle)
 *   stat[i] = {n_obs_collisions, spawn_ctr, episode_return_milli (int32), episode_len}  uint4
 */
typedef struct HwSaState {
    float *car_a;
    float *car_b;
    float *obst;
    uint32_t *stat;
} HwSaState;

/* Outputs of one step.  `term` rows are written only for environments whose episode ended in this
 * step: {tick, n_obs_collisions, episode_return_milli, episode_len} (what Monitor reports).
 * `stats` (nullable) accumulates over calls: [0] episodes ended, [1] sum of episode lengths,
 * [2] sum of episode returns in milli-reward (two's complement), [3] sum of obstacle collisions,
 * [4] episodes ended by time-out, [5..7] reserved. */
typedef struct HwSaOut {
    float *obs;      /* [n][18] ([n][14] with HW_FLAG_COMPACT_OBS), post-reset observation where done */
    float *rew;      /* [n] */
    uint8_t *done;   /* [n] */
    int32_t *term;   /* [n][4], nullable */
    unsigned long long *stats; /* [8], nullable */
} HwSaOut;

int hw_abi_version(void);
const char *hw_error_string(int code);

/* reset(): environments whose mask byte is non-zero (mask NULL = all) go back to the start grid
 * (highway_env.py:55-65); obs (nullable) receives their reset observation. */
int hw_sa_reset(const HwSaState *state, const uint8_t *mask, float *obs, int64_t n, void *stream);

/* step(): `actions` is int32[n] (discrete) or float[n][2] (HW_FLAG_CONTINUOUS).  Spawn uniforms of
 * environment i come from Philox4x32-10 keyed by (seed, env_id0 + i, spawn_ctr). */
int hw_sa_step(const HwSaState *state, const void *actions, const HwSaOut *out, const HwSimParams *params,
               uint64_t seed, uint64_t env_id0, int64_t n, int flags, void *stream);

/* get_state(): observation of the current state, no side effects. */
int hw_sa_observe(const HwSaState *state, float *obs, int64_t n, void *stream);

/* `steps` consecutive step() calls in ONE launch with i.i.d. uniform random actions drawn in-kernel
 * (Philox stream keyed by (seed ^ action tag, env, action_ctr0 + t)); state stays in registers
 * between steps, only the last step's outputs are stored.  Used by the random-action benchmark
 * and by warm-up ("roll the batch forward"). */
int hw_sa_rollout_random(const HwSaState *state, const HwSaOut *out, const HwSimParams *params,
                         uint64_t seed, uint64_t env_id0, uint64_t action_ctr0, int steps,
                         int64_t n, int flags, void *stream);

/* Host-buffer step: HOST actions -> device -> step -> HOST obs/rew/done, all on `stream`, returns
 * after the copies have landed.  When `aux_stream` is a second (caller-owned, long-lived) stream, batches of
 * 262,144 environments and more run as two half-batch launches, the second one on `aux_stream`, so that its upload
 * and kernel overlap the first half's download; with aux_stream NULL the call is one upload, one launch, one download.
 * d_actions and d_out are caller-owned DEVICE staging buffers of the same shapes (h_obs and d_out->obs are [n][14] with
 * HW_FLAG_COMPACT_OBS); h_* should be page-locked
 * (and allocated on the GPU's NUMA node) for full copy bandwidth. */
int hw_sa_step_host(const HwSaState *state, const void *h_actions, void *d_actions, const HwSaOut *d_out,
                    float *h_obs, float *h_rew, uint8_t *h_done, const HwSimParams *params,
                    uint64_t seed, uint64_t env_id0, int64_t n, int flags, void *stream, void *aux_stream);

/* ------------------------------------------------------------------------------------------------
 * Multi-agent environments: A cars (1..5) and up to K live obstacles (3..16) per environment.
 *
 * State, structure of arrays over n environments (i = environment):
 *   cars[(a*9 + w)*n + i]  32-bit word w of car a: 0 px, 1 py, 2 raw_x, 3 vx, 4 vy, 5 acc, 6 steer,
 *                          7 cruise_vel (floats), 8 bits (uint32: lane << 16 | HW_BIT_* mode flags)
 *   obst[(k*n + i)*4 + j]  float4 {px, raw_x, vx, init_vx} of obstacle slot k; slots 0..nobs-1 are live and
 *                          kept in insertion order (the order pygame's sprite Group iterates); 16-byte aligned
 *   head[w*n + i]          uint32 words of the per-environment header:
 *        0  tick (bits 0-15) | leader car << 16 (3 bits) | nobs << 20 (5 bits)
 *        1  slot of lane_max_obs[lane] for lane 1,2,3: 5 bits each, 31 = that obstacle has been retired
 *        2  lane id of slot k in bits 2k..2k+1
 *        3  bit k: slot k still has pygame's default rect (spawned, not yet updated)
 *           bit 16: a spawn was dropped because all K slots were live; bit 17: a lane had no obstacle to observe
 *           bits 24..28: car a is done since an earlier step (is_done persists until reset, highway_core.py:239,:420;
 *           only ever non-zero with HW_FLAG_NO_AUTORESET)
 *        4  n_obs_collisions  5  n_agent_collisions  6  spawn_ctr  7  episode_len
 *        8,9  episode return (sum of every car's reward), IEEE double, low word first
 */
typedef struct HwMaState {
    float *cars;
    float *obst;
    uint32_t *head;
    int32_t num_agents;
    int32_t num_slots;
} HwMaState;

/* Outputs of one multi-agent step.  `term` rows are written only for environments in which some car
 * finished (the whole environment resets then, dummy_vec_ma_env.py:74-77):
 * {tick, n_obs_collisions, n_agent_collisions, episode_return, episode_len} as doubles.
 * `stats` (nullable) accumulates [0] episodes, [1] sum of lengths, [2] sum of returns in milli-reward,
 * [3] obstacle collisions, [4] time-outs, [5] car-car collisions, [6] dropped spawns. */
typedef struct HwMaOut {
    float *obs;      /* [n][A][4A+14], post-reset observation where any car is done */
    float *rew;      /* [n][A] */
    uint8_t *done;   /* [n][A] */
    double *term;    /* [n][5], nullable */
    unsigned long long *stats; /* [8], nullable */
} HwMaOut;

int hw_ma_reset(const HwMaState *state, const uint8_t *mask, float *obs, int64_t n, void *stream);

/* `actions` is int32[n][A] (the argmax of each car's action vector, multiagent_envs/highway_env.py:77-78)
 * or float[n][A][2] with HW_FLAG_CONTINUOUS (highway_ma_c_env.py:63-64). */
int hw_ma_step(const HwMaState *state, const void *actions, const HwMaOut *out, const HwSimParams *params,
               uint64_t seed, uint64_t env_id0, int64_t n, int flags, void *stream);

int hw_ma_observe(const HwMaState *state, float *obs, int64_t n, void *stream);

int hw_ma_step_host(const HwMaState *state, const void *h_actions, void *d_actions, const HwMaOut *d_out,
                    float *h_obs, float *h_rew, uint8_t *h_done, const HwSimParams *params,
                    uint64_t seed, uint64_t env_id0, int64_t n, int flags, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Rollout collection (consumer side; baselines/ppo2/runner.py:33, baselines/common/distributions.py CategoricalPd).
 * The tail of model.step for a categorical policy in ONE launch: per row, log-softmax over the first `n_actions` entries of
 * `head` (float32, or bfloat16 when head_is_bf16), one sample by inverse CDF on a Philox uniform keyed by
 * (seed, env_id0 + row, counter + *counter_offset), its negative log-probability, and the value (entry n_actions of the
 * row).  `counter_offset` (nullable) is a DEVICE counter the caller advances between rollouts, so that a CUDA graph holding
 * a whole rollout draws fresh numbers on every replay.  `stride` = elements per row of `head` (>= n_actions + 1).  Outputs
 * may point into [T, N] rollout buffers. */
/* float32 [n][cols] -> bfloat16 [n][dst_cols] (dst_cols even, >= cols), zero padded on the right: the policy's first GEMM wants its
 * K dimension aligned, the observation has 18 columns. */
int hw_cast_pad_bf16(const float *src, int cols, void *dst, int dst_cols, int64_t n, void *stream);

/* The same for the two-dimensional diagonal Gaussian policy of the continuous env (DiagGaussianPd): head row = {mean0, mean1, value},
 * actions[i] = mean + exp(logstd) * N(0, 1) (what the learner stores), env_actions[i] = that clamped to [-1, 1] (what the env acts on). */
int hw_policy_sample_gaussian2(const void *head, int head_is_bf16, int stride, const float *logstd /* device, 2 floats */, uint64_t seed,
                               uint64_t env_id0, uint64_t counter, const unsigned long long *counter_offset,
                               float *actions, float *env_actions, float *neglogp, float *values, int64_t n, void *stream);

/* GAE(gamma, lam) of ppo2's runner (baselines/ppo2/runner.py:53-65) over step-major [T][n] buffers in one launch:
 * dones is uint8 [T + 1][n] (row t: the episode ended before step t; row T: after the last step), last_values [n] the value
 * of the state after the last step.  Writes returns = advantages + values, and the advantages when `advantages` is non-NULL.
 * Precisions follow the runner's NumPy expressions: float32 gamma * V, float64 recursion, float32 advantages and returns. */
int hw_gae(const float *rewards, const float *values, const uint8_t *dones, const float *last_values, double gamma, double lam,
           float *returns, float *advantages, int T, int64_t n, void *stream);

/* n-step discounted returns of a2c's runner (baselines/a2c/runner.py:55-66, a2c/utils.py:147-153 discount_with_dones) over
 * step-major [T][n] float32 rewards in one launch: dones is uint8 [T + 1][n] with row t + 1 = the episode ended ON step t
 * (row 0, "ended before the first step", is not read), last_values [n] bootstraps the envs whose last step did not end an
 * episode.  The recursion runs in float64 like the reference's Python floats; out [T][n] float32 (may alias rewards). */
int hw_nstep_returns(const float *rewards, const uint8_t *dones, const float *last_values, double gamma, float *out, int T,
                     int64_t n, void *stream);

#define HW_POLICY_MAX_ACTIONS 8
int hw_policy_sample_discrete(const void *head, int head_is_bf16, int stride, int n_actions, uint64_t seed,
                              uint64_t env_id0, uint64_t counter, const unsigned long long *counter_offset,
                              int32_t *actions, float *neglogp, float *values, int64_t n, void *stream);

/* ---- transition store of the MADDPG / DDPG consumers ----
 * Replaces, per rollout step, the host loops of maddpg/trainer/ma_ddpg.py:228-262 (episode bookkeeping), :233-237 with
 * ma_ddpg_learner.py:390-396 (store_transition -> Memory.append per env and agent, ma_memory.py:20-31, :74-83) and
 * baselines/common/mpi_running_mean_std.py:41-48 (the float64 sum / sum of squares / count vector that is all-reduced). */
typedef struct HwRing {            /* per-agent ring buffers, stored together: row r of agent a at [r][a] */
    float *obs0;                   /* [limit][A][obs_dim] */
    float *obs1;                   /* [limit][A][obs_dim] */
    float *actions;                /* [limit][A][act_dim] */
    float *rewards;                /* [limit][A]   (already multiplied by reward_scale) */
    float *terminals1;             /* [limit][A]   0 / 1 */
    int64_t *state;                /* device words {start, length, appended_total}: RingBuffer.start / .length */
} HwRing;

typedef struct HwTransitionBlock { /* the (nenvs, A, ...) blocks one env step produced; strides in ELEMENTS between envs */
    const float *obs0;  int64_t stride_obs0;     /* [n][A][obs_dim]  observation the actions were computed from */
    const float *obs1;  int64_t stride_obs1;     /* [n][A][obs_dim]  observation the step returned (post-reset where done) */
    const float *actions; int64_t stride_actions;/* [n][A][act_dim] */
    const float *rewards; int64_t stride_rewards;/* [n][A] */
    const void *done;   int64_t stride_done;     /* [n][A] uint8 (env output) or float (done_is_float: a packed, gathered block) */
    int32_t done_is_float;
} HwTransitionBlock;

typedef struct HwEpisodeTracker {  /* nullable as a whole, or ep_reward == NULL: no bookkeeping */
    float *ep_reward;              /* [n][A] running sum of the current episode's rewards (float32, ma_ddpg.py:174) */
    int64_t *ep_step;              /* [n] */
    float *fin_reward;             /* [n][A] episode_reward of the envs whose episode ended on this step */
    int64_t *fin_step;             /* [n]    their episode_step; 0 = the env's episode goes on */
    double *epoch_reward;          /* [A], nullable: += episode_reward of every finished episode (epoch_episode_rewards) */
    uint64_t *epoch_counts;        /* [2], nullable: {episodes, sum of episode steps} */
} HwEpisodeTracker;

/* Append the n transitions of one env step to the ring in env order - what n x A Memory.append calls leave behind (the oldest
 * entries are overwritten once `limit` is reached; with n > limit only the last `limit` survive) - advance start / length on the
 * device, and run the episode bookkeeping.  Two launches (store, then a one-thread advance) on `stream`. */
int hw_ring_store(const HwRing *ring, const HwTransitionBlock *blk, const HwEpisodeTracker *ep, int64_t n, int num_agents,
                  int obs_dim, int act_dim, int64_t limit, float reward_scale, void *stream);

/* addvec of RunningMeanStd.update for a float32 [n][cols] block: out[0..cols) = column sums, out[cols..2cols) = column sums of
 * squares (float64, deterministic summation order), out[2*cols] = count.  workspace: hw_obs_moments_workspace(n, cols) bytes. */
int64_t hw_obs_moments_workspace(int64_t n, int cols);
int hw_obs_moments(const float *x, int64_t n, int cols, double count, double *workspace, double *out, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* HIGHWAY_B200_H */
