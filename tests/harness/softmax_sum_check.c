/* TEST INFRASTRUCTURE: runs csi-nn2_b200/csrc/softmax_sum.h (the code the CUDA softmax kernel uses
 * for its denominator) on the CPU against the literal float += double loop of
 * source/reference/softmax.c:53-55.  Returns the number of mismatching cases. */
#include <math.h>
#include <stdlib.h>

#include "../../csi-nn2_b200/csrc/softmax_sum.h"

static uint64_t rng_state;
static double rnd(void)
{
    rng_state = rng_state * 6364136223846793005ull + 1442695040888963407ull;
    return (double)(rng_state >> 11) / 9007199254740992.0;
}

/* the integer form, driven the way the kernel drives it: per binade, every term's increment
 * (all threads in the kernel), then the one-thread chain until the binade changes */
static float sum_integer_form(const double *e, int n, uint32_t *d)
{
    float acc = 0.f;
    int j = 0;
    while (j < n) {
        uint32_t bits;
        memcpy(&bits, &acc, sizeof bits);
        const int ex = (int)((bits >> 23) & 0xFF);
        if (ex == 0 || ex >= 0xFE) {
            j = b200_softmax_chain(0, e, j, n, &acc);
            continue;
        }
        for (int i = j; i < n; i++) d[i] = b200_softmax_term(e[i], ex - 127);
        j = b200_softmax_chain(d, e, j, n, &acc);
    }
    return acc;
}

/* the scan form of the CUDA kernel (csrc/softmax.cu), sequentially: per binade, all remaining terms'
 * increments, the first term at which the running integer sum leaves the binade (or that must be
 * literal), everything before it consumed at once, that one term by the literal step */
static float sum_scan_form(const double *e, int n_all, uint32_t *d)
{
    float acc = 0.f;
    int j0 = 0;
    while (j0 < n_all) {
        const int n = j0 + 64 < n_all ? j0 + 64 : n_all; /* the kernel scans windows of 64 terms */
        uint32_t bits;
        memcpy(&bits, &acc, sizeof bits);
        const uint32_t ex = (bits >> 23) & 0xFF;
        if (ex == 0 || ex >= 0xFE) {
            acc = b200_softmax_sum_literal(e + j0, 1, acc);
            j0++;
            continue;
        }
        const uint32_t a = (bits & 0x7FFFFFu) | 0x800000u;
        const uint64_t room = (1ull << 24) - a;
        uint64_t run = 0;
        int p = n;
        for (int j = j0; j < n; j++) {
            d[j] = b200_softmax_term(e[j], (int)ex - 127);
            const uint64_t nxt = run + ((d[j] & B200_SOFTMAX_LITERAL) ? (1ull << 40) : d[j]);
            if (nxt >= room) {
                p = j;
                break;
            }
            run = nxt;
        }
        const uint32_t nb = (ex << 23) | ((a + (uint32_t)run) & 0x7FFFFFu);
        memcpy(&acc, &nb, sizeof acc);
        if (p >= n) {
            j0 = n;
            continue;
        }
        acc = b200_softmax_sum_literal(e + p, 1, acc);
        j0 = p + 1;
    }
    return acc;
}

/* cases: how many sequences; n: terms per sequence; mode selects the distribution of the terms */
int softmax_sum_check(int cases, int n, int mode, uint64_t seed)
{
    double *e = (double *)malloc(sizeof(double) * (size_t)n);
    uint32_t *d = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)n);
    int bad = 0;
    rng_state = seed;
    for (int c = 0; c < cases; c++) {
        for (int j = 0; j < n; j++) {
            double v;
            switch (mode) {
                case 0: v = exp(-20.0 * rnd()); break;                 /* softmax-like: exp of non-positive logits */
                case 1: v = rnd() < 0.02 ? 1.0 : exp(-40.0 * rnd()); break; /* a few maxima, long tail */
                case 2: v = ldexp(rnd(), -(int)(60.0 * rnd())); break; /* wide dynamic range, many binade crossings */
                case 3: v = ldexp((double)(1 + (int)(rnd() * 7.0)), -24 - (int)(rnd() * 3.0)); break; /* exact ties */
                default: v = rnd() < 0.5 ? 0.0 : ldexp(rnd(), -140);   /* zeros and subnormal-range sums */
            }
            e[j] = v;
        }
        const float want = b200_softmax_sum_literal(e, n, 0.f);
        const float got = b200_softmax_sum(e, n);
        const float got_i = sum_integer_form(e, n, d);
        const float got_s = sum_scan_form(e, n, d);
        if (memcmp(&want, &got, sizeof want) != 0 || memcmp(&want, &got_i, sizeof want) != 0 ||
            memcmp(&want, &got_s, sizeof want) != 0)
            bad++;
    }
    free(e);
    free(d);
    return bad;
}
