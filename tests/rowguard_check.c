/* Property behind the "row guard" of numerics.cuh / kernels_gradient.cu: if every non-zero FISTA
 * value of a 3x3 neighbourhood has magnitude >= 2^-35, every non-zero numerator the gradient
 * kernel divides (TV: a1*gx, a1*gy, a1*-(gx+gy); TGV: (s+gxx)+gyy, s+gxx, gyy+s, -s) has magnitude
 * >= 2^-60 — so one test per loaded value replaces one per numerator.  Random neighbourhoods whose
 * values crowd the threshold (and cancel heavily) are pushed through the kernel's own float
 * expressions (compute.c:79-81, :136-145, :98-103, :165-182). */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static uint64_t rng = 0x9E3779B97F4A7C15ull;
static uint64_t next(void) { rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17; return rng; }
static float from_bits(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

static uint32_t base_bits;                        /* per neighbourhood: a value just above 2^-35 */
static float value(void) {                         /* 0, or the base value a few ulps off (differences = a few quanta of 2^-58), or unrelated */
        const uint64_t r = next();
        if ((r & 15) == 0) return 0.f;
        if ((r & 15) == 1) {                       /* an unrelated value up to 2^-30 */
                const float v = from_bits(((uint32_t)(127 - 35 + (int)((r >> 4) % 6)) << 23) | ((uint32_t)(r >> 16) & 0x7fffff));
                return (r & 0x200000000ull) ? -v : v;
        }
        return from_bits(base_bits + (uint32_t)((r >> 8) % 7));
}

int main(int argc, char **argv) {
        const long n = argc > 1 ? atol(argv[1]) : 5000000;
        const float lim = 8.673617379884035e-19f;                       /* 2^-60 */
        const float a1s[3] = {1.f, (float)(1. / sqrtf(2.f)), (float)(1. / sqrtf(3.f))};   /* compute.c:90 for nc = 1, 2, 3 */
        long bad = 0, nonzero = 0;
        float smallest = INFINITY;
        for (long it = 0; it < n; it++) {
                base_bits = ((uint32_t)(127 - 35) << 23) | ((uint32_t)next() & 0x7ffff8);
                float y[4][4];                                          /* rows s-1..s+1 (+1 spare), columns x-1..x+1 (+1 spare) */
                for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) y[r][c] = value();
                /* first differences at (s, x), (s, x-1), (s-1, x): row index 1 = s, column index 1 = x */
                const float gx = y[1][2] - y[1][1], gy = y[2][1] - y[1][1];
                const float gx_l = y[1][1] - y[1][0], gy_l = y[2][0] - y[1][0];
                const float gx_u = y[0][2] - y[0][1], gy_u = y[1][1] - y[0][1];
                const float gxx = gx - gx_l, gyx = gy - gy_l, gxy = gx - gx_u, gyy = gy - gy_u;
                const float s = (gxy + gyx) * 0.5f;
                const float a1 = a1s[it % 3];
                const float num[7] = {a1 * -(gx + gy), a1 * gx, a1 * gy, (s + gxx) + gyy, s + gxx, gyy + s, -s};
                for (int k = 0; k < 7; k++) {
                        if (num[k] == 0.f) continue;
                        nonzero++;
                        if (fabsf(num[k]) < smallest) smallest = fabsf(num[k]);
                        if (fabsf(num[k]) < lim) {
                                if (bad < 10) printf("VIOLATION numerator %d = %a\n", k, num[k]);
                                bad++;
                        }
                }
        }
        /* ---- the two power-of-two rewrites of the TGV stage (kernels_gradient.cu) -------------------
         *   2*RN((u/2)^2)                      == RN(u * (u/2))
         *   a2 * (-(2gxx + 2s + 2gyy) / n)     == (-2 a2) * (((s + gxx) + gyy) / n)
         * on values in the range the fast path admits (differences >= 2^-58 in magnitude or 0, norms in [2^-40, 2^40]). */
        long rewrites = 0;
        for (long it = 0; it < n; it++) {
                const int e = (int)(next() % 70) - 50;                  /* magnitudes 2^-50 .. 2^19 */
                const float gxx = ldexpf((float)((int)(next() % 4001) - 2000) + (float)(next() & 0xffff) / 65536.f, e);
                const float gyy = ldexpf((float)((int)(next() % 4001) - 2000) + (float)(next() & 0xffff) / 65536.f, e - (int)(next() % 3));
                const float u = ldexpf((float)((int)(next() % 4001) - 2000) + (float)(next() & 0xffff) / 65536.f, e + (int)(next() % 3) - 1);
                const float sgm = u * 0.5f;
                const float a2 = 0.05f + (float)(next() % 1000) / 500.f;
                const float nn = sqrtf(gxx * gxx + 2.f * (sgm * sgm) + gyy * gyy);
                if (!(nn >= 9.094947017729282e-13f && nn <= 1.099511627776e12f)) continue;
                const float lhs1 = 2.f * (sgm * sgm), rhs1 = u * sgm;
                const float lhs2 = a2 * (-((2.f * gxx + 2.f * sgm) + 2.f * gyy) / nn);
                const float rhs2 = (-2.f * a2) * (((sgm + gxx) + gyy) / nn);
                rewrites++;
                if (!(lhs1 == rhs1) || !(lhs2 == rhs2)) {
                        if (bad < 10) printf("REWRITE MISMATCH gxx=%a s=%a gyy=%a n=%a: %a vs %a, %a vs %a\n", gxx, sgm, gyy, nn, lhs1, rhs1, lhs2, rhs2);
                        bad++;
                }
        }
        printf("rowguard_check: %ld rewrite pairs compared\n", rewrites);
        printf("rowguard_check: %ld non-zero numerators, %ld below 2^-60; smallest seen 2^%.2f\n", nonzero, bad, log2f(smallest));
        return bad ? 1 : 0;
}
