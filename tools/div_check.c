/* div_check.c -- verifies on the CPU that the reciprocal-based division used by the CUDA kernels
 *     q0 = a*r; e0 = fma(-q0,b,a); q1 = fma(e0,r,q0); e1 = fma(-q1,b,a); q2 = fma(e1,r,q1)     (r = RN(1/b))
 * returns exactly the IEEE-754 quotient a/b (what Java computes in UniformMesh.XtoL, UM:158-159) for the
 * divisors a mesh uses.  Random dividends plus dividends constructed to land next to rounding midpoints of the
 * quotient (the only place a faithful-but-not-correct result could appear).
 * Build: gcc -O2 -ffp-contract=off -mfma -fopenmp tools/div_check.c -lm -o /tmp/div_check ; run: /tmp/div_check [n_per_b]
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static inline uint64_t sm64(uint64_t *s) { uint64_t z = (*s += 0x9E3779B97F4A7C15ull); z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); }
static inline double fastdiv(double a, double b, double r)
{
    double q0 = a * r, e0 = fma(-q0, b, a), q1 = fma(e0, r, q0), e1 = fma(-q1, b, a);
    return fma(e1, r, q1);
}
static inline double bits2d(uint64_t u) { double d; memcpy(&d, &u, 8); return d; }
static inline uint64_t d2bits(double d) { uint64_t u; memcpy(&u, &d, 8); return u; }

int main(int argc, char **argv)
{
    long n = argc > 1 ? atol(argv[1]) : 20000000L;
    double bs[64];
    int nb = 0;
    const double fixed[] = {1e-3, 5e-3, 2e-3, 0.5e-3, 1.0, 0.1, 0.25, 3.0, 1e-6, 7.3e-5, 1.0 / 3.0, 1e-2, 1.5e-3, 2.5e-4};
    for (unsigned k = 0; k < sizeof fixed / sizeof *fixed; k++) bs[nb++] = fixed[k];
    uint64_t s0 = 12345;
    while (nb < 40) bs[nb++] = ldexp(1.0 + (double)(sm64(&s0) >> 11) * 0x1p-53, (int)(sm64(&s0) % 40) - 30);
    long bad = 0, total = 0;
    for (int ib = 0; ib < nb; ib++) {
        const double b = bs[ib], r = 1.0 / b;
        long badb = 0;
#pragma omp parallel for reduction(+ : badb)
        for (long k = 0; k < n; k++) {
            uint64_t s = (uint64_t)k * 0x2545F4914F6CDD1Dull + ib;
            double a;
            const uint64_t w = sm64(&s);
            switch (w & 3) {
            case 0: a = (double)(sm64(&s) >> 11) * 0x1p-53 * 2.0 - 0.5; break;              /* x - x0 in a unit box */
            case 1: a = ldexp(1.0 + (double)(sm64(&s) >> 11) * 0x1p-53, (int)(sm64(&s) % 60) - 40); break; /* wide range */
            default: { /* next to a rounding midpoint of the quotient: a ~ (m + 1/2) ulp * b */
                const int e = (int)(sm64(&s) % 24) - 4;
                const uint64_t m = (1ull << 52) | (sm64(&s) >> 12);
                const double mid = ldexp((double)m, e - 52) + ldexp(1.0, e - 53); /* exact in long arithmetic? m+0.5 ulp: 54 bits */
                /* (m + 1/2)*2^(e-52) is not a double; approximate the product in long double and step around it */
                long double t = ((long double)m + 0.5L) * ldexpl(1.0L, e - 52) * (long double)b;
                a = (double)t;
                const int step = (int)(sm64(&s) % 5) - 2;
                a = bits2d(d2bits(a) + (uint64_t)(int64_t)step);
                (void)mid;
            }
            }
            if (w & 4) a = -a;
            const double q = a / b, f = (a == 0) ? a * r : fastdiv(a, b, r);
            if (d2bits(q) != d2bits(f)) badb++;
        }
        if (badb) printf("b=%.17g: %ld mismatches\n", b, badb);
        bad += badb;
        total += n;
    }
    printf("div_check: %ld divisions, %ld mismatches\n", total, bad);
    return bad != 0;
}
