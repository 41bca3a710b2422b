/*
 * oracle/div_check.c -- exhaustive CPU checks of the two division sequences the CUDA kernels use in place of IEEE
 * divides.  THIS IS TEST INFRASTRUCTURE (tests/test_division_sequences.py), not product code.
 *
 *  (1) x / c for an integer count c in [2, 16] (TIES disjoint mean, ties.cu: div_by_count_fast; PCB's final / n,
 *      pcb.cu: pcb_div_count):  inv = RN(1 / c);  q0 = RN(x inv);  q = RN(q0 + (x - q0 c) inv)  -- the residual is
 *      exact in an FMA.  Claim: bit-identical to the IEEE quotient for every x in [2^-100, 2^100].  A float quotient
 *      depends on the mantissa only (apart from the exponent range, which the kernels guard), so the check runs over all
 *      2^23 mantissas at the exponents the caller names.
 *  (2) x / y with a prepared reciprocal r = RN(1 / y) (pcb.cu: pcb_div_by):  q0 = RN(x r), then twice
 *      q <- RN(q + (x - q y) r).  Claim: after the second correction the quotient is the correctly rounded one
 *      (Markstein: a faithful quotient corrected once with a correctly rounded reciprocal is correctly rounded; q0 can
 *      be 1.5 ulp off, so the first correction only makes it faithful).  Checked over all 2^23 mantissas of x for each
 *      divisor y the caller passes; the count of mismatches after ONE correction is returned too.
 *
 * fmaf() must be a true fused multiply-add: the loops are compiled for the FMA instruction where the CPU has it
 * (runtime dispatch) and fall back to libm's fmaf (correct, slower) elsewhere.  Built with -ffp-contract=off so that
 * nothing else is fused.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

typedef uint64_t u64;

static inline float from_bits(uint32_t b) { float f; memcpy(&f, &b, 4); return f; }
static inline uint32_t to_bits(float f) { uint32_t b; memcpy(&b, &f, 4); return b; }

#define COUNT_BODY(FMA)                                                                                   \
    u64 bad = 0;                                                                                          \
    const float fc = (float)c, inv = 1.0f / fc;                                                           \
    _Pragma("omp parallel for reduction(+ : bad) schedule(static)")                                       \
    for (int64_t m = 0; m < (1 << 23); ++m) {                                                             \
        for (int s = 0; s < 2; ++s) {                                                                     \
            const float x = from_bits(((uint32_t)s << 31) | ((uint32_t)(exp2 + 127) << 23) | (uint32_t)m); \
            const float q0 = x * inv;                                                                     \
            const float r = FMA(-q0, fc, x);                                                              \
            const float q = FMA(r, inv, q0);                                                              \
            bad += to_bits(q) != to_bits(x / fc);                                                         \
        }                                                                                                 \
    }                                                                                                     \
    return bad;

#define CHAIN_BODY(FMA)                                                                                   \
    u64 bad1 = 0, bad2 = 0;                                                                               \
    const float r = 1.0f / y;                                                                             \
    _Pragma("omp parallel for reduction(+ : bad1, bad2) schedule(static)")                                \
    for (int64_t m = 0; m < (1 << 23); ++m) {                                                             \
        const float x = from_bits(((uint32_t)(exp2 + 127) << 23) | (uint32_t)m);                          \
        const float want = x / y;                                                                         \
        float q = x * r;                                                                                  \
        q = FMA(FMA(-q, y, x), r, q);                                                                     \
        bad1 += to_bits(q) != to_bits(want);                                                              \
        q = FMA(FMA(-q, y, x), r, q);                                                                     \
        bad2 += to_bits(q) != to_bits(want);                                                              \
    }                                                                                                     \
    *after_one = bad1;                                                                                    \
    return bad2;

#if defined(__x86_64__)
__attribute__((target("fma"))) static u64 count_hw(int c, int exp2) { COUNT_BODY(__builtin_fmaf) }
__attribute__((target("fma"))) static u64 chain_hw(float y, int exp2, u64* after_one) { CHAIN_BODY(__builtin_fmaf) }
static int have_fma(void) { return __builtin_cpu_supports("fma"); }
#else
static u64 count_hw(int c, int exp2) { COUNT_BODY(fmaf) }
static u64 chain_hw(float y, int exp2, u64* after_one) { CHAIN_BODY(fmaf) }
static int have_fma(void) { return 0; }
#endif
static u64 count_sw(int c, int exp2) { COUNT_BODY(fmaf) }
static u64 chain_sw(float y, int exp2, u64* after_one) { CHAIN_BODY(fmaf) }

/* mismatches of sequence (1) against x / c over all mantissas and both signs of x = +-(1.m) * 2^exp2 */
u64 orc_check_div_count(int c, int exp2) { return have_fma() ? count_hw(c, exp2) : count_sw(c, exp2); }

/* mismatches of sequence (2) against x / y over all mantissas of x = (1.m) * 2^exp2; *after_one = after one correction */
u64 orc_check_div_chain(float y, int exp2, u64* after_one) {
    return have_fma() ? chain_hw(y, exp2, after_one) : chain_sw(y, exp2, after_one);
}
int orc_div_check_uses_fma_instruction(void) { return have_fma(); }
