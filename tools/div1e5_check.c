// Exhaustive / randomised check that div1e5 (scab_common.cuh) == x / 1e5.  gcc -O2 -fopenmp -ffp-contract=off tools/div1e5_check.c -lm
#include <math.h>
#include <stdio.h>
#include <stdint.h>
#include <omp.h>
// exact x / 1e5 for integer-valued doubles x via one multiply and two FMAs?
static inline double div1e5(double x) {
    const double y = 1e-5;            // RN(1/1e5)
    double q = x * y;
    double r = fma(-q, 1e5, x);
    return fma(r, y, q);
}
int main() {
    long long bad = 0;
    const long long LIM = 1LL << 33;
    #pragma omp parallel for reduction(+:bad) schedule(static)
    for (long long i = -LIM; i <= LIM; i++) {
        double x = (double)i;
        if (div1e5(x) != x / 1e5) bad++;
    }
    printf("integers |x| <= 2^33: mismatches %lld\n", bad);
    // non-integers: random doubles in wide range
    long long bad2 = 0; uint64_t s = 88172645463325252ULL;
    for (long long k = 0; k < 400000000LL; k++) {
        s ^= s << 13; s ^= s >> 7; s ^= s << 17;
        double x = ldexp((double)(s >> 11), -53 + (int)((s & 63)) - 20);
        if (div1e5(x) != x / 1e5) bad2++;
    }
    printf("random doubles: mismatches %lld\n", bad2);
    return 0;
}
