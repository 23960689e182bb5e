/* softmax_sum.h -- the softmax denominator of source/reference/softmax.c:53-55,
 *     float acc = 0;  for j: acc += exp(x[j] - max);          (float += double)
 * i.e. acc <- (float)((double)acc + e[j]) term by term: the ORDER and the two roundings per term
 * (to double, then to float) fix the result bits, so the chain cannot be split over threads.  What
 * can be shortened is the chain itself: a float <-> double conversion pair per term costs far more
 * latency than the add.  While acc stays inside one binade [2^k, 2^(k+1)) its float rounding is
 * "round to a multiple of u = 2^(k-23)", which in double arithmetic is (s + C) - C with
 * C = 1.5 * 2^(k+29) (ulp_double(C) = u, C/u even so ties go to the same neighbour as the float
 * rounding).  So a block of terms is accumulated in double as  s = a + e;  a = (s + C) - C  -- three
 * dependent adds, no conversion -- and accepted if the block ends inside the binade (the terms are
 * non-negative, so the sum is monotone and no step left the binade either); otherwise the block is
 * redone with the literal conversion sequence.  Plain C so that the same code is checked on the CPU
 * against the literal loop (tests/harness/softmax_sum_check.c).
 */
#ifndef B200_SOFTMAX_SUM_H_
#define B200_SOFTMAX_SUM_H_

#include <stdint.h>
#include <string.h>

#ifdef __CUDACC__
#define B200_HD __device__ __forceinline__
#define B200_DADD(a, b) __dadd_rn((a), (b))
#else
#define B200_HD static inline
#define B200_DADD(a, b) ((double)((volatile double)(a) + (volatile double)(b)))
#endif

#define B200_SOFTMAX_BLOCK 8

B200_HD double b200_bits_to_double(uint64_t u)
{
    double d;
    memcpy(&d, &u, sizeof d);
    return d;
}

/* the literal sequence */
B200_HD float b200_softmax_sum_literal(const double *e, int n, float acc)
{
    for (int j = 0; j < n; j++) acc = (float)B200_DADD((double)acc, e[j]);
    return acc;
}

/* same result, short dependency chain; every e[j] >= 0 and finite */
B200_HD float b200_softmax_sum(const double *e, int n)
{
    float acc = 0.f;
    int j = 0;
    while (j < n) {
        const int nb = n - j < B200_SOFTMAX_BLOCK ? n - j : B200_SOFTMAX_BLOCK;
        uint32_t bits;
        memcpy(&bits, &acc, sizeof bits);
        const int ex = (int)((bits >> 23) & 0xFF);
        if (ex == 0 || ex >= 0xFE - 30) { /* zero / subnormal / huge: literal steps */
            acc = b200_softmax_sum_literal(e + j, nb, acc);
            j += nb;
            continue;
        }
        const int k = ex - 127;
        const double c = b200_bits_to_double(((uint64_t)(k + 29 + 1023) << 52) | (1ull << 51)); /* 1.5 * 2^(k+29) */
        const double hi = b200_bits_to_double((uint64_t)(k + 1 + 1023) << 52);                 /* 2^(k+1) */
        double a = (double)acc;
#ifdef __CUDACC__
#pragma unroll
#endif
        for (int i = 0; i < B200_SOFTMAX_BLOCK; i++) {
            if (i < nb) {
                const double s = B200_DADD(a, e[j + i]);
                a = B200_DADD(B200_DADD(s, c), -c);
            }
        }
        if (a < hi)
            acc = (float)a; /* exact: a is a multiple of 2^(k-23) below 2^(k+1) */
        else
            acc = b200_softmax_sum_literal(e + j, nb, acc);
        j += nb;
    }
    return acc;
}

/* ---- integer form ---------------------------------------------------------------------------
 * Measured on B200 the double adds above have the same ~16-cycle latency as the conversions, so the
 * chain did not get shorter.  Inside one binade the whole step is integer arithmetic: with
 * u = 2^(k-23), acc = A * u (A in [2^23, 2^24)) and one double ulp of the sum = u * 2^-29,
 *     RN_double(acc + e) = (A * 2^29 + r) * u * 2^-29,   r = RNE(e * 2^(52-k))       (exact scaling)
 *     RN_float(...)      = (A + (r >> 29) + up) * u,     up = low29(r) > 2^28, or == 2^28 and the sum odd
 * r, and with it the increment d = (r >> 29) + up, depends on the term and the binade only: all
 * threads compute d[j] for the current binade, one thread adds them up (a 4-cycle integer add per
 * term), and a term that ties, or a block of terms that leaves the binade, takes the literal float
 * <- double step.  b200_softmax_term / b200_softmax_chain are the two halves; the CPU check drives
 * them the way the kernel does. */
#define B200_SOFTMAX_LITERAL 0x80000000u

#ifdef __CUDACC__
#define B200_D2LL_RN(x) __double2ll_rn(x)
#define B200_DMUL(a, b) __dmul_rn((a), (b))
#else
#include <math.h>
#define B200_D2LL_RN(x) llrint(x)
#define B200_DMUL(a, b) ((double)((volatile double)(a) * (volatile double)(b)))
#endif

/* increment of term e while the float sum is in binade k (unbiased exponent), or LITERAL */
B200_HD uint32_t b200_softmax_term(double e, int k)
{
    const double hi = b200_bits_to_double((uint64_t)(k + 1 + 1023) << 52);      /* 2^(k+1) */
    const double scale = b200_bits_to_double((uint64_t)(52 - k + 1023) << 52);  /* 2^(52-k) */
    if (!(e < hi)) return B200_SOFTMAX_LITERAL; /* the term alone leaves the binade (also NaN) */
    const long long r = B200_D2LL_RN(B200_DMUL(e, scale));                       /* < 2^53 */
    const uint32_t low = (uint32_t)(r & ((1ll << 29) - 1));
    if (low == (1u << 28)) return B200_SOFTMAX_LITERAL;                          /* float rounding tie */
    return (uint32_t)(r >> 29) + (low > (1u << 28) ? 1u : 0u);
}

/* one thread: consume terms j0.. while the sum stays in the binade the d[] were computed for
 * (d == NULL: literal steps only); returns the index of the first unconsumed term */
B200_HD int b200_softmax_chain(const uint32_t *d, const double *e, int j0, int n, float *acc_io)
{
    float acc = *acc_io;
    int j = j0;
    uint32_t bits;
    memcpy(&bits, &acc, sizeof bits);
    const uint32_t ex = (bits >> 23) & 0xFF;
    if (d == 0 || ex == 0 || ex >= 0xFE) { /* zero / subnormal / overflowing sum: literal steps */
        const int nb = n - j < B200_SOFTMAX_BLOCK ? n - j : B200_SOFTMAX_BLOCK;
        *acc_io = b200_softmax_sum_literal(e + j, nb, acc);
        return j + nb;
    }
    uint32_t a = (bits & 0x7FFFFFu) | 0x800000u; /* A */
    while (j < n) {
        const int nb = n - j < B200_SOFTMAX_BLOCK ? n - j : B200_SOFTMAX_BLOCK;
        uint32_t a2 = a, flags = 0;
#ifdef __CUDACC__
#pragma unroll
#endif
        for (int i = 0; i < B200_SOFTMAX_BLOCK; i++) {
            if (i < nb) {
                const uint32_t v = d[j + i];
                flags |= v;
                a2 += v & 0x7FFFFFFFu;
            }
        }
        if ((flags & B200_SOFTMAX_LITERAL) || a2 >= (1u << 24)) {
            /* a tie, an oversized term or the end of the binade inside this block */
            const uint32_t abits = (ex << 23) | (a & 0x7FFFFFu);
            memcpy(&acc, &abits, sizeof acc);
            acc = b200_softmax_sum_literal(e + j, nb, acc);
            j += nb;
            memcpy(&bits, &acc, sizeof bits);
            if (((bits >> 23) & 0xFF) != ex) { /* new binade: the caller recomputes d[] */
                *acc_io = acc;
                return j;
            }
            a = (bits & 0x7FFFFFu) | 0x800000u;
            continue;
        }
        a = a2;
        j += nb;
    }
    const uint32_t abits = (ex << 23) | (a & 0x7FFFFFu);
    memcpy(&acc, &abits, sizeof acc);
    *acc_io = acc;
    return j;
}

#endif
