/* check_div.c -- TEST INFRASTRUCTURE (see lmc_oracle.c header).
 *
 * Mode B's interpolation weight alpha = (t - t_before) / (t_after - t_before) (CS:1503) is an IEEE division of two
 * integers a < b.  The kernel forms it as q0 = a * RN(1/b); q = fma(fma(-b, q0, a), RN(1/b), q0) -- one Markstein
 * correction -- for b < 2^50.  This program compares that sequence with the hardware division over structured
 * (every a for typical IMU periods) and random (b up to 2^50, a near 0 / near b / uniform) operands and prints the
 * number of differing results (expected: 0).   usage: check_div [n_random]
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

static uint64_t s = 88172645463325252ULL;
static inline uint64_t rnd(void) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; }

static inline int differs(int64_t a, int64_t b)
{
    const double da = (double)a, db = (double)b, r = 1.0 / db;
    const double q0 = da * r;
    const double q1 = fma(fma(-db, q0, da), r, q0);
    return q1 != da / db;
}

int main(int argc, char** argv)
{
    const long n_random = argc > 1 ? atol(argv[1]) : 20000000L;
    static const int64_t bs[] = { 5000000, 10000000, 4999999, 5000001, 3333333, 1000000, 2500000, 20000000, 1, 2, 3, 7,
                                  1000, 999983, 4194304, 4194303, 1125899906842623LL, 1125899906842597LL };
    long bad = 0, n = 0;
    for (unsigned k = 0; k < sizeof bs / sizeof bs[0]; ++k) {
        const int64_t b = bs[k], lim = b < 2000000 ? b : 2000000;
        for (int64_t i = 0; i < lim; ++i, ++n) bad += differs(b <= 2000000 ? i : (int64_t)(rnd() % (uint64_t)b), b);
    }
    for (long i = 0; i < n_random; ++i, ++n) {
        const int sh = 1 + (int)(rnd() % 50);
        const int64_t b = 1 + (int64_t)(rnd() & ((1ULL << sh) - 1));
        int64_t a = (int64_t)(rnd() % (uint64_t)b);
        if (i & 1) a = b - 1 - (a % (b < 64 ? b : 64));
        bad += differs(a < 0 ? 0 : a, b);
    }
    printf("%ld %ld\n", n, bad);
    return bad != 0;
}
