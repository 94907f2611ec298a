// Transition store of the MADDPG / DDPG consumers for sm_100a: what the reference does per rollout step on the host
//   for agent in trainers: agent.store_transition(obs_n, actions_n, rew_n, new_obs_n, done_n)   (maddpg/trainer/ma_ddpg.py:233-237)
//     -> for b in range(nenvs): memory.append(obs0[b][i], action[b][i], reward[b][i] * reward_scale, obs1[b][i], terminal1[b][i])
//        (ma_ddpg_learner.py:390-396, ma_memory.py:20-31 RingBuffer.append, :74-83 Memory.append)
//   episode_reward += rew_n; episode_step += 1; per-env bookkeeping when any(done_n[d])          (ma_ddpg.py:228-262)
//   obs_rms.update(np.array([obs0[b]]))  -> float64 sum / sum of squares / count                 (mpi_running_mean_std.py:41-48)
// as ONE launch over the (nenvs, A, ...) blocks the multi-agent VecEnv kernel left in HBM (plus a one-thread launch that advances
// the ring's start / length words, which live on the device so that a captured CUDA graph of a rollout keeps appending).
// HBM-bound copies: every thread moves 16 bytes of observation per access where the row length allows it.
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/highway_b200.h"

namespace hw {

struct RingPtrs {
    float *obs0, *obs1, *actions, *rewards, *terminals1;  // [limit][A][D], [limit][A][D], [limit][A][act], [limit][A], [limit][A]
    long long *state;                                     // {start, length, appended_total}
};

struct BlockPtrs {
    const float *obs0, *obs1, *actions, *rewards;  // rows of n envs; strides in floats between consecutive envs
    const void *done;                               // uint8 [n][A] (the env's output) or float (a packed, gathered block)
    long long s_obs0, s_obs1, s_act, s_rew, s_done;
    int done_is_float;
};

struct EpisodePtrs {  // all optional (NULL): the rollout loop's per-env metric tracking
    float *ep_reward;        // [n][A] running sum of this episode's rewards (float32 like the reference's array)
    long long *ep_step;      // [n]
    float *fin_reward;       // [n][A] episode_reward of the envs whose episode ended on this step (else untouched)
    long long *fin_step;     // [n]    episode_step of those envs, 0 for the others
    double *epoch_reward;    // [A]    += episode_reward[d] of every finished episode
    unsigned long long *epoch_counts;  // {episodes, sum of episode steps}
};

__device__ __forceinline__ bool env_done(const BlockPtrs &B, const long long env, const int A)
{
    bool any = false;
    for (int a = 0; a < A; ++a) {
        const long long j = env * B.s_done + a;
        any |= B.done_is_float ? (static_cast<const float *>(B.done)[j] != 0.f) : (static_cast<const uint8_t *>(B.done)[j] != 0);
    }
    return any;
}

template <int VEC>
__global__ void __launch_bounds__(256)
ring_store_kernel(const RingPtrs R, const BlockPtrs B, const EpisodePtrs E, const long long n, const int A, const int D, const int act_dim,
                  const long long limit, const float reward_scale, const long long skip)
{
    // the block's n transitions go to ring rows (start + length + b) % limit, b = 0..n-1, exactly where n successive appends put them;
    // when n > limit only the last `limit` survive (skip = n - limit rows are dropped up front)
    const long long start = R.state[0], length = R.state[1];
    const long long row0 = (start + length) % limit;
    const long long AD = (long long)A * D, ADv = AD / VEC;
    const long long total = (n - skip) * ADv;
    for (long long idx = (long long)blockIdx.x * 256 + threadIdx.x; idx < total; idx += (long long)gridDim.x * 256) {
        const long long b = idx / ADv + skip, c = (idx - (b - skip) * ADv) * VEC;   // env in the block, first column
        const long long row = (row0 + b) % limit;
        if (VEC == 4) {
            *reinterpret_cast<float4 *>(R.obs0 + row * AD + c) = *reinterpret_cast<const float4 *>(B.obs0 + b * B.s_obs0 + c);
            *reinterpret_cast<float4 *>(R.obs1 + row * AD + c) = *reinterpret_cast<const float4 *>(B.obs1 + b * B.s_obs1 + c);
        } else if (VEC == 2) {
            *reinterpret_cast<float2 *>(R.obs0 + row * AD + c) = *reinterpret_cast<const float2 *>(B.obs0 + b * B.s_obs0 + c);
            *reinterpret_cast<float2 *>(R.obs1 + row * AD + c) = *reinterpret_cast<const float2 *>(B.obs1 + b * B.s_obs1 + c);
        } else {
            R.obs0[row * AD + c] = B.obs0[b * B.s_obs0 + c];
            R.obs1[row * AD + c] = B.obs1[b * B.s_obs1 + c];
        }
    }
    // the small per-(env, agent) words: action vector, scaled reward, terminal flag, and the episode bookkeeping (all envs of the
    // block, also the ones a too-small ring drops)
    const long long pairs = n * A;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < pairs; i += (long long)gridDim.x * 256) {
        const long long b = i / A;
        const int a = (int)(i - b * A);
        const float r = B.rewards[b * B.s_rew + a];
        const long long j = b * B.s_done + a;
        const float term = B.done_is_float ? static_cast<const float *>(B.done)[j] : (float)static_cast<const uint8_t *>(B.done)[j];
        if (b >= skip) {
            const long long row = (row0 + b) % limit;
            for (int k = 0; k < act_dim; ++k) R.actions[(row * A + a) * act_dim + k] = B.actions[b * B.s_act + (long long)a * act_dim + k];
            R.rewards[row * A + a] = r * reward_scale;      // store_transition: reward *= self.reward_scale
            R.terminals1[row * A + a] = term;
        }
        if (E.ep_reward) {
            const float acc = E.ep_reward[i] + r;                      // episode_reward += rew_n (float32)
            const bool ended = env_done(B, b, A);
            if (ended) {
                E.fin_reward[i] = acc;
                E.ep_reward[i] = 0.f;
                if (E.epoch_reward) atomicAdd(&E.epoch_reward[a], (double)acc);
            } else {
                E.ep_reward[i] = acc;
            }
            if (a == 0) {
                const long long st = E.ep_step[b] + 1;
                E.fin_step[b] = ended ? st : 0;
                E.ep_step[b] = ended ? 0 : st;
                if (ended && E.epoch_counts) { atomicAdd(&E.epoch_counts[0], 1ull); atomicAdd(&E.epoch_counts[1], (unsigned long long)st); }
            }
        }
    }
}

__global__ void ring_advance_kernel(long long *state, const long long n, const long long limit)
{
    if (threadIdx.x || blockIdx.x) return;
    const long long start = state[0], length = state[1];
    const long long over = length + n > limit ? length + n - limit : 0;   // entries pushed out of a full ring
    state[0] = (start + over) % limit;
    state[1] = length + n > limit ? limit : length + n;
    state[2] += n;
}

// float64 column sums and sums of squares of a float32 [n][C] block, in two deterministic stages: stage 1 sums slabs of
// kMomentRows rows per block (thread c walks its column down the slab: coalesced across the threads of a block), stage 2 adds the
// slab results in slab order and appends the row count - the `addvec` RunningMeanStd.update builds before its Allreduce.
constexpr int kMomentRows = 64;

__global__ void __launch_bounds__(256)
moments_slab_kernel(const float *__restrict__ x, const long long n, const int C, double *__restrict__ partial /* [slabs][2][C] */)
{
    const long long slab = blockIdx.y;
    const int c = blockIdx.x * 256 + threadIdx.x;
    if (c >= C) return;
    const long long r0 = slab * kMomentRows, r1 = r0 + kMomentRows < n ? r0 + kMomentRows : n;
    // four interleaved accumulators (rows r0 + 4j + i) so that four loads are in flight per thread; combined in a fixed order
    // (eight in flight measured slower: 37 vs 21.5 us for a 45 MB block)
    double s[4] = {0.0, 0.0, 0.0, 0.0}, q[4] = {0.0, 0.0, 0.0, 0.0};
    long long r = r0;
    for (; r + 4 <= r1; r += 4) {
        float v[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = x[(r + i) * C + c];
#pragma unroll
        for (int i = 0; i < 4; ++i) { const double d = (double)v[i]; s[i] += d; q[i] += d * d; }
    }
    for (int i = 0; r < r1; ++r, ++i) { const double d = (double)x[r * C + c]; s[i] += d; q[i] += d * d; }
    partial[(slab * 2 + 0) * C + c] = (s[0] + s[1]) + (s[2] + s[3]);
    partial[(slab * 2 + 1) * C + c] = (q[0] + q[1]) + (q[2] + q[3]);
}

// stage 2: out[c] = sum over the slabs of partial[slab][c], c over the 2C {sum, sum of squares} columns.  A block owns 32 consecutive
// columns (coalesced 256-byte rows); its 32 slab groups (threadIdx.y) each add every 32nd slab in order, and the 32 group sums are
// then added in group order - a fixed summation tree, so the result does not depend on scheduling.
__global__ void __launch_bounds__(1024)
moments_finish_kernel(const double *__restrict__ partial, const long long slabs, const int C, const double count, double *__restrict__ out /* [2C + 1] */)
{
    __shared__ double part[32][33];
    const int c = blockIdx.x * 32 + threadIdx.x;
    double s = 0.0;
    if (c < 2 * C) {
        // eight loads in flight, added in slab order (the same sum as a plain loop; issued one by one the loop was a chain of 32
        // dependent L2 round trips on 11 blocks: 10 us)
        long long k = threadIdx.y;
        for (; k + 7 * 32 < slabs; k += 8 * 32) {
            double v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = partial[(k + 32 * i) * 2 * C + c];
#pragma unroll
            for (int i = 0; i < 8; ++i) s += v[i];
        }
        for (; k < slabs; k += 32) s += partial[k * 2 * C + c];
    }
    part[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0 && c < 2 * C) {
        double t = 0.0;
#pragma unroll
        for (int g = 0; g < 32; ++g) t += part[g][threadIdx.x];
        out[c] = t;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0 && threadIdx.y == 0) out[2 * C] = count;
}

}  // namespace hw

extern "C" int hw_ring_store(const HwRing *ring, const HwTransitionBlock *blk, const HwEpisodeTracker *ep, int64_t n, int num_agents,
                             int obs_dim, int act_dim, int64_t limit, float reward_scale, void *stream)
{
    if (!ring || !blk || n <= 0 || num_agents < 1 || obs_dim < 1 || act_dim < 1 || limit < 1) return HW_E_BADARG;
    if (!ring->obs0 || !ring->obs1 || !ring->actions || !ring->rewards || !ring->terminals1 || !ring->state) return HW_E_BADARG;
    if (!blk->obs0 || !blk->obs1 || !blk->actions || !blk->rewards || !blk->done) return HW_E_BADARG;
    hw::RingPtrs R = {ring->obs0, ring->obs1, ring->actions, ring->rewards, ring->terminals1, reinterpret_cast<long long *>(ring->state)};
    hw::BlockPtrs B = {blk->obs0, blk->obs1, blk->actions, blk->rewards, blk->done, blk->stride_obs0, blk->stride_obs1, blk->stride_actions,
                       blk->stride_rewards, blk->stride_done, blk->done_is_float};
    hw::EpisodePtrs E = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    if (ep && ep->ep_reward) {
        if (!ep->ep_step || !ep->fin_reward || !ep->fin_step) return HW_E_BADARG;
        E = {ep->ep_reward, reinterpret_cast<long long *>(ep->ep_step), ep->fin_reward, reinterpret_cast<long long *>(ep->fin_step),
             ep->epoch_reward, reinterpret_cast<unsigned long long *>(ep->epoch_counts)};
    }
    const long long AD = (long long)num_agents * obs_dim;
    const long long skip = n > limit ? n - limit : 0;
    auto aligned = [&](int v) {
        const uintptr_t m = (uintptr_t)(v * 4 - 1);
        return AD % v == 0 && blk->stride_obs0 % v == 0 && blk->stride_obs1 % v == 0 &&
               !(((uintptr_t)ring->obs0 | (uintptr_t)ring->obs1 | (uintptr_t)blk->obs0 | (uintptr_t)blk->obs1) & m);
    };
    const int vec = aligned(4) ? 4 : aligned(2) ? 2 : 1;
    const long long work = (n - skip) * (AD / vec);
    long long blocks = (work + 255) / 256;
    const long long cap = 148LL * 8 * 4;  // a few waves of 8 resident blocks per SM, grid-stride beyond
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    cudaStream_t s = (cudaStream_t)stream;
    if (vec == 4)      hw::ring_store_kernel<4><<<(unsigned)blocks, 256, 0, s>>>(R, B, E, n, num_agents, obs_dim, act_dim, limit, reward_scale, skip);
    else if (vec == 2) hw::ring_store_kernel<2><<<(unsigned)blocks, 256, 0, s>>>(R, B, E, n, num_agents, obs_dim, act_dim, limit, reward_scale, skip);
    else               hw::ring_store_kernel<1><<<(unsigned)blocks, 256, 0, s>>>(R, B, E, n, num_agents, obs_dim, act_dim, limit, reward_scale, skip);
    hw::ring_advance_kernel<<<1, 32, 0, s>>>(R.state, n, limit);
    return (int)cudaGetLastError();
}

extern "C" int64_t hw_obs_moments_workspace(int64_t n, int cols)
{
    if (n <= 0 || cols < 1) return 0;
    return ((n + hw::kMomentRows - 1) / hw::kMomentRows) * 2 * (int64_t)cols * (int64_t)sizeof(double);
}

extern "C" int hw_obs_moments(const float *x, int64_t n, int cols, double count, double *workspace, double *out, void *stream)
{
    if (!x || !workspace || !out || n <= 0 || cols < 1) return HW_E_BADARG;
    const long long slabs = (n + hw::kMomentRows - 1) / hw::kMomentRows;
    if (slabs > 65535) return HW_E_BADARG;  // gridDim.y
    cudaStream_t s = (cudaStream_t)stream;
    dim3 grid((unsigned)((cols + 255) / 256), (unsigned)slabs);
    hw::moments_slab_kernel<<<grid, 256, 0, s>>>(x, n, cols, workspace);
    hw::moments_finish_kernel<<<(unsigned)((2 * cols + 31) / 32), dim3(32, 32), 0, s>>>(workspace, slabs, cols, count, out);
    return (int)cudaGetLastError();
}
