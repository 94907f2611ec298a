/* TEST INFRASTRUCTURE (oracle/): which tangent the CPU restatement evaluates.
 *
 * Default: libm's tan(), the function CPython's math.tan calls (multi_lane_sim.py:205, agent.py:131) - the reference's value.
 * hw_oracle_set_tan(1) switches to the polynomial the CUDA kernels use (tan_small, gym_highway_b200/csrc/hw_common.cuh: within 1 ulp
 * of libm, DESIGN.md deviation 3), so that a soak can tell a kernel bug from that declared deviation: a mismatch against the libm
 * oracle that disappears against this one is the 1-ulp tangent reaching a stored fp32 value, nothing else. */
#ifndef HW_ORACLE_TAN_SMALL_H
#define HW_ORACLE_TAN_SMALL_H
#include <math.h>

extern int hw_oracle_tan_poly; /* defined in sa_oracle.c */

static inline double hw_tan_small(double x)
{
    static const double c[12] = {
        0x1.20f68de6f5b19p-15, 0x1.790a294731d86p-16, 0x1.b8cea70623061p-14, 0x1.f054fa543c93bp-13,
        0x1.3598d277c0b71p-11, 0x1.7d9f32bbba0bcp-10, 0x1.d6d3fe7809e9dp-9,  0x1.226e34c1e860cp-7,
        0x1.664f48853966bp-6,  0x1.ba1ba1ba167cdp-5,  0x1.1111111111133p-3,  0x1.5555555555555p-2};
    const double z = x * x;
    double p = c[0];
    for (int i = 1; i < 12; ++i) p = fma(p, z, c[i]);
    return fma(x * z, p, x);
}

static inline double hw_tan(double x) { return hw_oracle_tan_poly ? hw_tan_small(x) : tan(x); }
#endif
