// Consumer-side kernel of the on-device rollout collectors: the tail of ppo2's / a2c's
// model.step for a categorical policy, baselines/common/distributions.py CategoricalPd (sample, neglogp) applied to the
// network's head, fused into one launch and writing straight into the rollout's [T, N] buffers.
//
// Eagerly in torch this is a dozen small launches per step (log_softmax, exp, running sums, compares, gather, casts,
// copies) - a quarter of a rollout step at 10^5 environments once the env step itself takes 16 us.
#include <cuda_bf16.h>
#include "hw_common.cuh"
#include "../../include/highway_b200.h"

namespace hw {

constexpr uint32_t kPolicyTag = 0x504F4C21u;  // "POL!": action-sampling stream

template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<unsigned short>(unsigned short v) { return __uint_as_float((uint32_t)v << 16); }  // bf16

// head[i * stride + k], k < n_actions: logits; head[i * stride + n_actions]: value
template <typename T>
__global__ void __launch_bounds__(256)
policy_sample_discrete_kernel(const T *__restrict__ head, const int stride, const int n_actions, const uint64_t seed,
                              const uint64_t env_id0, const uint64_t counter0, const unsigned long long *__restrict__ counter_offset,
                              int32_t *__restrict__ actions,
                              float *__restrict__ neglogp, float *__restrict__ values, const int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const uint64_t counter = counter0 + (counter_offset ? (uint64_t)*counter_offset : 0ull);  // device-side part: advances under CUDA-graph replay
    const T *row = head + i * stride;
    float x[HW_POLICY_MAX_ACTIONS];
    float mx = -3.402823466e38f;
#pragma unroll
    for (int k = 0; k < HW_POLICY_MAX_ACTIONS; ++k) {
        x[k] = (k < n_actions) ? to_f32<T>(row[k]) : -3.402823466e38f;
        mx = fmaxf(mx, x[k]);
    }
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < HW_POLICY_MAX_ACTIONS; ++k) sum += (k < n_actions) ? expf(x[k] - mx) : 0.f;
    const float lse = mx + logf(sum);  // log_softmax(x)_k = x_k - lse
    // one uniform in [0, 1) per row and step: Philox keyed by the seed, counter = (step, global env id)
    const uint64_t id = env_id0 + (uint64_t)i;
    const uint4 w = philox4x32_10(make_uint4((uint32_t)counter, (uint32_t)id, (uint32_t)(id >> 32), kPolicyTag ^ (uint32_t)(counter >> 32)),
                                  make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    const float u = (float)(w.x >> 8) * (1.0f / 16777216.0f);
    // inverse CDF: the number of running sums of the probabilities that stay below u (the last action takes the rest)
    int a = 0;
    float c = 0.f, lp = x[0] - lse;
#pragma unroll
    for (int k = 0; k < HW_POLICY_MAX_ACTIONS - 1; ++k) {
        if (k < n_actions - 1) {
            c += expf(x[k] - lse);
            if (c < u) { a = k + 1; lp = x[k + 1] - lse; }
        }
    }
    actions[i] = a;
    neglogp[i] = -lp;
    values[i] = to_f32<T>(row[n_actions]);
}

// float32 [n][cols] -> bfloat16 [n][dst_cols], zero padded on the right (round to nearest even): the first GEMM of the policy wants
// its K dimension aligned to 8 elements, the observation has 18 columns
__global__ void __launch_bounds__(256)
cast_pad_bf16_kernel(const float *__restrict__ src, const int cols, unsigned short *__restrict__ dst, const int dst_cols, const int64_t total)
{
    const int64_t j = (int64_t)blockIdx.x * 256 + threadIdx.x;  // one thread per pair of output elements
    if (2 * j >= total) return;
    const int64_t row = (2 * j) / dst_cols;
    const int col = (int)((2 * j) - row * dst_cols);  // dst_cols is even, so a pair never straddles rows
    const float a = (col < cols) ? src[row * cols + col] : 0.f, b = (col + 1 < cols) ? src[row * cols + col + 1] : 0.f;
    const __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    reinterpret_cast<__nv_bfloat162 *>(dst)[j] = v;
}

// The same for a diagonal Gaussian policy (DiagGaussianPd): a = mean + exp(logstd) * N(0, 1) with the normals from Box-Muller
// on two Philox uniforms per pair, neglogp = 0.5 * sum(((a - mean) / std)^2) + 0.5 * log(2 pi) * d + sum(logstd); the
// unclipped sample is stored for the learner, its clamp to [-1, 1] is what the env acts on.
// head[i * stride + k], k < 2: means; head[i * stride + 2]: value
template <typename T>
__global__ void __launch_bounds__(256)
policy_sample_gaussian2_kernel(const T *__restrict__ head, const int stride, const float *__restrict__ logstd,
                               const uint64_t seed, const uint64_t env_id0, const uint64_t counter0,
                               const unsigned long long *__restrict__ counter_offset, float2 *__restrict__ actions,
                               float2 *__restrict__ env_actions, float *__restrict__ neglogp, float *__restrict__ values, const int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const uint64_t counter = counter0 + (counter_offset ? (uint64_t)*counter_offset : 0ull);
    const T *row = head + i * stride;
    const float logstd0 = logstd[0], logstd1 = logstd[1];  // a trained parameter: read on the device, also under graph replay
    const float m0 = to_f32<T>(row[0]), m1 = to_f32<T>(row[1]);
    const uint64_t id = env_id0 + (uint64_t)i;
    const uint4 w = philox4x32_10(make_uint4((uint32_t)counter, (uint32_t)id, (uint32_t)(id >> 32), kPolicyTag ^ (uint32_t)(counter >> 32)),
                                  make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    const float u1 = ((float)(w.x >> 8) + 1.0f) * (1.0f / 16777216.0f);  // (0, 1]
    const float u2 = (float)(w.y >> 8) * (1.0f / 16777216.0f);
    const float r = sqrtf(-2.0f * logf(u1));
    float sn, cs;
    sincospif(2.0f * u2, &sn, &cs);
    const float z0 = r * cs, z1 = r * sn;
    const float a0 = m0 + expf(logstd0) * z0, a1 = m1 + expf(logstd1) * z1;
    actions[i] = make_float2(a0, a1);
    env_actions[i] = make_float2(fminf(fmaxf(a0, -1.f), 1.f), fminf(fmaxf(a1, -1.f), 1.f));
    neglogp[i] = 0.5f * (z0 * z0 + z1 * z1) + 1.8378770664093453f + logstd0 + logstd1;
    values[i] = to_f32<T>(row[2]);
}

// Generalised advantage estimation exactly as ppo2's runner computes it (baselines/ppo2/runner.py:53-65), one thread per
// environment walking its T steps backwards; all arrays are [T][n] (step major), so a warp reads 32 neighbouring floats per step.
//   dones[t]  (t < T): the episode ended on the transition BEFORE step t;  dones[T]: after the last step
//   delta = r[t] + gamma * V[t+1] * (1 - dones[t+1]) - V[t];  A[t] = delta + gamma * lam * (1 - dones[t+1]) * A[t+1]
__global__ void __launch_bounds__(256)
gae_kernel(const float *__restrict__ rewards, const float *__restrict__ values, const uint8_t *__restrict__ dones,
           const float *__restrict__ last_values, const double gamma, const double lam, float *__restrict__ returns,
           float *__restrict__ advantages, const int T, const int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    // The precisions of the runner's NumPy expressions, operation by operation (ppo2/runner.py:57-65): rewards and values are
    // float32 arrays, `1.0 - dones` is float64, so `gamma * nextvalues` is a float32 product (the Python scalar adopts the
    // array's type) and everything after it float64; lastgaelam stays float64 across steps, mb_advs[t] is its float32
    // rounding, and mb_returns = mb_advs + mb_values is a float32 sum.
    const float g32 = (float)gamma;
    const double gl = gamma * lam;
    float nextvalue = last_values[i];
    double lastgaelam = 0.0;
    double nextnonterminal = 1.0 - (double)dones[(int64_t)T * n + i];
    for (int t = T - 1; t >= 0; --t) {
        const float v = values[(int64_t)t * n + i];
        const double delta = __dsub_rn(__dadd_rn((double)rewards[(int64_t)t * n + i], __dmul_rn((double)__fmul_rn(g32, nextvalue), nextnonterminal)), (double)v);
        lastgaelam = __dadd_rn(delta, __dmul_rn(__dmul_rn(gl, nextnonterminal), lastgaelam));
        const float adv = (float)lastgaelam;
        if (advantages) advantages[(int64_t)t * n + i] = adv;
        returns[(int64_t)t * n + i] = __fadd_rn(adv, v);
        nextvalue = v;
        nextnonterminal = 1.0 - (double)dones[(int64_t)t * n + i];
    }
}

// n-step discounted returns of a2c's runner (baselines/a2c/runner.py:55-66 with a2c/utils.py:147-153 discount_with_dones):
// per env, r = reward + gamma * r * (1. - done) walked backwards over Python floats (float64) - the float32 rewards, and
// the float32 bootstrap value appended as one more "reward" when the last step did not end the episode - then stored
// back into the float32 reward row.  dones: uint8 [T + 1][n] with row t + 1 = the episode ended ON step t (row 0 unused).
__global__ void __launch_bounds__(256)
nstep_returns_kernel(const float *__restrict__ rewards, const uint8_t *__restrict__ dones, const float *__restrict__ last_values,
                     const double gamma, float *__restrict__ out, const int T, const int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    // dones[-1] == 0: discount_with_dones(rewards + [value], dones + [0]) starts from r = value + gamma * 0 * 1 = value
    double r = dones[(int64_t)T * n + i] ? 0.0 : __dadd_rn((double)last_values[i], __dmul_rn(__dmul_rn(gamma, 0.0), 1.0));
    for (int t = T - 1; t >= 0; --t) {
        const double nd = 1.0 - (double)dones[(int64_t)(t + 1) * n + i];
        r = __dadd_rn((double)rewards[(int64_t)t * n + i], __dmul_rn(__dmul_rn(gamma, r), nd));
        out[(int64_t)t * n + i] = (float)r;
    }
}

}  // namespace hw

extern "C" int hw_nstep_returns(const float *rewards, const uint8_t *dones, const float *last_values, double gamma, float *out, int T,
                                int64_t n, void *stream)
{
    if (!rewards || !dones || !last_values || !out || T < 1 || n <= 0) return HW_E_BADARG;
    hw::nstep_returns_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(rewards, dones, last_values, gamma, out, T, n);
    return (int)cudaGetLastError();
}

extern "C" int hw_gae(const float *rewards, const float *values, const uint8_t *dones, const float *last_values, double gamma,
                      double lam, float *returns, float *advantages, int T, int64_t n, void *stream)
{
    if (!rewards || !values || !dones || !last_values || !returns || T < 1 || n <= 0) return HW_E_BADARG;
    hw::gae_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(rewards, values, dones, last_values, gamma, lam, returns,
                                                                                  advantages, T, n);
    return (int)cudaGetLastError();
}

extern "C" int hw_policy_sample_gaussian2(const void *head, int head_is_bf16, int stride, const float *logstd, uint64_t seed,
                                          uint64_t env_id0, uint64_t counter, const unsigned long long *counter_offset,
                                          float *actions, float *env_actions, float *neglogp, float *values, int64_t n, void *stream)
{
    if (!head || !logstd || !actions || !env_actions || !neglogp || !values || n <= 0 || stride < 3) return HW_E_BADARG;
    const unsigned grid = (unsigned)((n + 255) / 256);
    cudaStream_t s = (cudaStream_t)stream;
    float2 *a = reinterpret_cast<float2 *>(actions), *e = reinterpret_cast<float2 *>(env_actions);
    if (head_is_bf16)
        hw::policy_sample_gaussian2_kernel<unsigned short><<<grid, 256, 0, s>>>(static_cast<const unsigned short *>(head), stride, logstd,
                                                                                 seed, env_id0, counter, counter_offset, a, e, neglogp, values, n);
    else
        hw::policy_sample_gaussian2_kernel<float><<<grid, 256, 0, s>>>(static_cast<const float *>(head), stride, logstd, seed, env_id0,
                                                                        counter, counter_offset, a, e, neglogp, values, n);
    return (int)cudaGetLastError();
}

extern "C" int hw_cast_pad_bf16(const float *src, int cols, void *dst, int dst_cols, int64_t n, void *stream)
{
    if (!src || !dst || n <= 0 || cols < 1 || dst_cols < cols || (dst_cols & 1)) return HW_E_BADARG;
    const int64_t total = n * dst_cols;
    const unsigned grid = (unsigned)((total / 2 + 255) / 256);
    hw::cast_pad_bf16_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, cols, static_cast<unsigned short *>(dst), dst_cols, total);
    return (int)cudaGetLastError();
}

extern "C" int hw_policy_sample_discrete(const void *head, int head_is_bf16, int stride, int n_actions, uint64_t seed,
                                         uint64_t env_id0, uint64_t counter, const unsigned long long *counter_offset,
                                         int32_t *actions, float *neglogp, float *values, int64_t n, void *stream)
{
    if (!head || !actions || !neglogp || !values || n <= 0) return HW_E_BADARG;
    if (n_actions < 1 || n_actions > HW_POLICY_MAX_ACTIONS || stride < n_actions + 1) return HW_E_BADARG;
    const unsigned grid = (unsigned)((n + 255) / 256);
    cudaStream_t s = (cudaStream_t)stream;
    if (head_is_bf16)
        hw::policy_sample_discrete_kernel<unsigned short><<<grid, 256, 0, s>>>(static_cast<const unsigned short *>(head), stride, n_actions,
                                                                                seed, env_id0, counter, counter_offset, actions, neglogp, values, n);
    else
        hw::policy_sample_discrete_kernel<float><<<grid, 256, 0, s>>>(static_cast<const float *>(head), stride, n_actions, seed, env_id0,
                                                                       counter, counter_offset, actions, neglogp, values, n);
    return (int)cudaGetLastError();
}
