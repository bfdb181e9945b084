// exhaustive check: for a constant divisor c with r = RN(1/c), does  q = x*r; q = fma(fma(-q, c, x), r, q)  equal the IEEE quotient x / c
// for every float x?  (all 2^32 bit patterns; NaN inputs skipped)
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
static inline float f_of(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline uint32_t u_of(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
int main(int argc, char **argv)
{
    float c = (float)atof(argv[1]);
    uint32_t lo = argc > 2 ? (uint32_t)strtoul(argv[2], 0, 0) : 0u, hi = argc > 3 ? (uint32_t)strtoul(argv[3], 0, 0) : 0xffffffffu;
    volatile float one = 1.0f;
    const float r = one / c;
    unsigned long long bad = 0, badNormal = 0;
    uint32_t firstBad = 0;
#pragma omp parallel for reduction(+ : bad, badNormal) schedule(static)
    for (long long i = lo; i <= (long long)hi; i++)
    {
        float x = f_of((uint32_t)i);
        if (x != x) continue;
        float want = x / c;
        float q = x * r;
        float rem = fmaf(-q, c, x);
        float got = fmaf(rem, r, q);
        if (u_of(want) != u_of(got) && !(want != want && got != got))
        {
            bad++;
            float ax = fabsf(x);
            if (ax >= 1e-30f && ax <= 1e30f) { badNormal++; }
#pragma omp critical
            if (!firstBad) firstBad = (uint32_t)i;
        }
    }
    printf("c=%g r=%.9g  mismatches: %llu (of which |x| in [1e-30,1e30]: %llu) first bad bits 0x%08x x=%g\n", c, r, bad, badNormal, firstBad, f_of(firstBad));
    return 0;
}
