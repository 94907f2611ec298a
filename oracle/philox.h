/* Philox4x32-10 counter-based generator (Salmon, Moraes, Dror, Shaw, "Parallel random numbers:
 * as easy as 1, 2, 3", SC'11), restated from the paper.  TEST INFRASTRUCTURE: the oracle draws the
 * obstacle-respawn uniforms from the same (seed, env_id, spawn_ctr) counter as the CUDA kernel so
 * that parity is defined on (state, action, uniforms) triples (SURVEY.md section 7 "Randomness").
 *
 * The reference draws `random.uniform(70,100)` then `random.uniform(5,7)` from CPython's MT19937
 * (gym_highway/envs/multi_lane_sim.py:975-977); only the bit stream differs here, the mapping
 * word-pair -> 53-bit double -> a + (b-a)*u is CPython's (random_random in _randommodule.c).
 */
#ifndef HW_ORACLE_PHILOX_H
#define HW_ORACLE_PHILOX_H
#include <stdint.h>

static inline void hw_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* CPython random(): (a>>5, b>>6) -> (a*2^26 + b) / 2^53 */
static inline double hw_u53(uint32_t a, uint32_t b)
{
    return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) * (1.0 / 9007199254740992.0);
}

/* the two uniforms of spawn number `spawn_ctr` of environment `env_id` */
static inline void hw_spawn_uniforms(uint64_t seed, uint64_t env_id, uint32_t spawn_ctr, double *u_pos, double *u_vel)
{
    uint32_t ctr[4] = { spawn_ctr, (uint32_t)env_id, (uint32_t)(env_id >> 32), 0x48574159u /* "HWAY" */ };
    uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) };
    uint32_t w[4];
    hw_philox4x32_10(ctr, key, w);
    *u_pos = hw_u53(w[0], w[1]);
    *u_vel = hw_u53(w[2], w[3]);
}
#endif
