// check: q = fma(fma(-q0, c, x), rc, q0), q0 = x*rc, rc = RN(1/c)  ==  x / c  in binary64
//   c = 20    : x = car velocity in [0, 20]  (reward = vx/20 - 1): 2e9 random doubles + all floats' neighbourhoods
//   c = 1000  : x = every integer in [-2 000 000, 2 000 000]  (round(r, 3) = k / 1000)
// Markstein's theorem guarantees the result (rc correctly rounded, q0 faithful); this is the empirical cross-check.
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
static uint64_t s = 0x9E3779B97F4A7C15ull;
static inline uint64_t rnd(void){ s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; }
static inline double divc(double x, double c, double rc){ double q0 = x * rc; double r = fma(-q0, c, x); return fma(r, rc, q0); }
int main(void){
  long bad = 0; const double c = 20.0, rc = 1.0 / 20.0;
  for (long i = 0; i < 2000000000L; ++i) {
    double x = (double)(rnd() >> 11) * (1.0 / 9007199254740992.0) * 20.0;
    if (i & 1) { uint64_t u; memcpy(&u, &x, 8); u = (u & ~0xFFFFFFFull) + (rnd() & 7) - 3; memcpy(&x, &u, 8); }  // near float32-grid points
    double q = divc(x, c, rc), w = x / c;
    if (memcmp(&q, &w, 8)) bad++;
  }
  printf("c=20: 2e9 doubles in [0,20], mismatches %ld\n", bad);
  bad = 0; const double c2 = 1000.0, rc2 = 1.0 / 1000.0;
  for (long k = -2000000; k <= 2000000; ++k) { double x = (double)k; double q = divc(x, c2, rc2), w = x / c2; if (k != 0 && memcmp(&q, &w, 8)) bad++; }
  printf("c=1000: integers in [-2e6,2e6], mismatches %ld\n", bad);
  return 0; }
