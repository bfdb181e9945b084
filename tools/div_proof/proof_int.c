// for every odd integer divisor m in [3,255] (even divisors are m * 2^k: same significands) and every float x with 2^-80 <= |x| < 2^11:
// q = x*r; q = fma(fma(-q, m, x), r, q) with r = RN(1/m) must equal the IEEE quotient x / m.  Divisors that are m * 2^k are checked too
// (all w in 1..256) on a thinner grid to make sure the power-of-two scaling argument holds.
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
static inline float f_of(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline uint32_t u_of(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static unsigned long long check(float c, uint32_t e0, uint32_t e1, uint32_t stride)
{
    volatile float one = 1.0f;
    const float r = one / c;
    unsigned long long bad = 0;
#pragma omp parallel for reduction(+ : bad) schedule(static)
    for (long long i = (long long)e0 << 23; i < ((long long)e1 << 23); i += stride)
        for (int sgn = 0; sgn < 2; sgn++)
        {
            float x = f_of((uint32_t)i | ((uint32_t)sgn << 31));
            float want = x / c, q = x * r;
            float got = fmaf(fmaf(-q, c, x), r, q);
            bad += u_of(want) != u_of(got);
        }
    return bad;
}
int main()
{
    const uint32_t e0 = 127 - 80, e1 = 127 + 11; // biased exponents [2^-80, 2^11)
    unsigned long long total = 0;
    for (int m = 3; m <= 255; m += 2)
    {
        unsigned long long b = check((float)m, e0, e1, 1);
        total += b;
        if (b) printf("m=%d: %llu mismatches\n", m, b);
        if ((m & 31) == 31) { printf("... up to m=%d, mismatches so far %llu\n", m, total); fflush(stdout); }
    }
    for (int w = 1; w <= 256; w++)
        total += check((float)w, e0, e1, 7);
    printf("TOTAL mismatches: %llu\n", total);
    return 0;
}
