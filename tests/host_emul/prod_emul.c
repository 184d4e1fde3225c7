/* CPU check of the final pass's float32 evaluation of RN32(RN64(x * g)) (epilogue.cu: prod2 / prod1): x float32, g a
 * non-negative float64 gain split on the host into gh = RZ32(g), gl = RN32(g - gh) >= 0.  The device forms
 *     p = RN(x gh), q = fma(x, -gh, p), e = fma(x, gl, -q), r = RN(p + e), rho = (p - r) + e, c = fma(rho, kappa, r)
 * and trusts r unless c != r (the sum sits within 2^-16 of a rounding boundary: the frame is redone in float64).
 * This program replays that with libm's correctly rounded fmaf over random and adversarial operands and fails if an
 * unflagged r differs from the float64 evaluation in a single bit (the sign of zero included).
 *   gcc -O2 -ffp-contract=off -o prod_emul prod_emul.c -lm && ./prod_emul [millions of random cases]  */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static uint64_t s[2] = {0x9E3779B97F4A7C15ull, 0xD1B54A32D192ED03ull};
static uint64_t rnd(void) {                      /* xorshift128+ */
    uint64_t a = s[0], b = s[1];
    s[0] = b;
    a ^= a << 23;
    s[1] = a ^ b ^ (a >> 17) ^ (b >> 26);
    return s[1] + b;
}
static uint32_t bits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static float from_bits(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static const float KAPPA = 1.0f + 1.52587890625e-05f;          /* 1 + 2^-16 */

static long flagged = 0, total = 0, wrong = 0;

static void one(float x, double g) {
    volatile float gh = (float)g;                /* the host's split: gh = g rounded TOWARD ZERO (gains are >= 0), so that */
    if ((double)gh > g) gh = nextafterf(gh, 0.0f);   /* gl >= 0 and the sign of a zero product comes out as numpy's */
    volatile float gl = (float)(g - (double)gh);
    volatile float p = x * gh;
    volatile float q = fmaf(x, -gh, p);
    volatile float e = fmaf(x, gl, -q);
    volatile float r = p + e;
    volatile float d = fmaf(r, -1.0f, p);
    volatile float rho = d + e;
    volatile float c = fmaf(rho, KAPPA, r);
    const float want = (float)((double)x * g);
    ++total;
    if (!(c == r)) { ++flagged; return; }
    if (bits(r) != bits(want)) {
        if (++wrong <= 10) fprintf(stderr, "MISMATCH x=%a g=%a got=%a want=%a\n", x, g, r, want);
    }
}

static float rand_x(void) {                      /* zero, or magnitude in [2^-60, 2^40), either sign */
    const uint64_t u = rnd();
    if ((u & 1023) == 0) return (u & 1024) ? -0.0f : 0.0f;
    const int ex = 127 - 60 + (int)((u >> 11) % 100);
    return from_bits(((uint32_t)(u >> 63) << 31) | ((uint32_t)ex << 23) | (uint32_t)((u >> 20) & 0x7fffff));
}
static double rand_g(void) {                     /* zero, or magnitude in [2^-16, 2^16]; mostly inside [0, 1] as the pan gains are */
    const uint64_t u = rnd();
    if ((u & 255) == 0) return 0.0;
    const uint64_t m = rnd() & 0xFFFFFFFFFFFFFull;
    int ex;
    if (u & 256) ex = 1023 - 16 + (int)((u >> 12) % 32); else ex = 1023 - 1 - (int)((u >> 12) % 4);
    uint64_t b = ((uint64_t)ex << 52) | m;
    if ((u & 3584) == 0) b &= ~0x1FFFFFFFull;    /* a gain that is a float32 value (gl = 0) */
    double g; memcpy(&g, &b, 8);
    return g;
}

int main(int argc, char** argv) {
    const long millions = argc > 1 ? atol(argv[1]) : 50;
    for (long i = 0; i < millions * 1000000L; ++i) one(rand_x(), rand_g());
    printf("random: cases %ld  flagged %ld (%.3g)  wrong %ld\n", total, flagged, (double)flagged / (double)total, wrong);
    /* adversarial: products that land on, or next to, float32 rounding boundaries.  x = 2^k * (small odd), g = float32
     * value + half an ulp +- a tiny amount, so that x g sits at a midpoint +- 2^-j ulp for every j up to the float64 grain */
    for (long i = 0; i < 4000000L; ++i) {
        const uint64_t u = rnd();
        const float ghf = from_bits((uint32_t)(126 - (u % 3)) << 23 | (uint32_t)((u >> 8) & 0x7fffff));
        const double ulp = ldexp(1.0, (int)(bits(ghf) >> 23) - 127 - 23);
        const int j = 1 + (int)((u >> 32) % 30);
        const double sgn = (u >> 40) & 1 ? 1.0 : -1.0;
        const double g = (double)ghf + 0.5 * ulp + sgn * ldexp(ulp, -j) * (double)((u >> 41) & 3);
        const float xs[] = {1.0f, 0.5f, 3.0f, 5.0f, 0.75f, -1.0f, -7.0f, 1.5f, 1.0f + ldexpf(1.0f, -23), from_bits(0x3f7fffffu)};
        for (unsigned k = 0; k < sizeof xs / sizeof *xs; ++k) one(xs[k], g);
    }
    /* double-rounding traps: x g just above / below a float32 midpoint by less than half a float64 ulp of the product */
    for (long i = 0; i < 4000000L; ++i) {
        const uint64_t u = rnd();
        const float x = from_bits(0x3f800000u | (uint32_t)(u & 0x7fffff));          /* [1, 2) */
        const float t = from_bits(0x3f000000u | (uint32_t)((u >> 23) & 0x7fffff));   /* target product mantissa in [0.5, 1) */
        const double mid = (double)t + ldexp(1.0, -25);                              /* midpoint above t */
        double g = mid / (double)x;                                                  /* x g ~ mid within 2^-53 */
        uint64_t gb; memcpy(&gb, &g, 8);
        gb += (int64_t)((u >> 46) % 7) - 3;                                         /* +- a few float64 ulps */
        memcpy(&g, &gb, 8);
        one(x, g);
        one(-x, g);
    }
    printf("all: cases %ld  flagged %ld (%.3g)  wrong %ld\n", total, flagged, (double)flagged / (double)total, wrong);
    return wrong ? 1 : 0;
}
