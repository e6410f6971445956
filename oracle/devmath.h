/* TEST INFRASTRUCTURE ONLY (oracle).
 *
 * "Device arithmetic" mode of the oracle: C restatement of the two elementary functions the CUDA kernels evaluate
 * with their own table-driven algorithms instead of libm,
 *     dm_exp   smc-nuts_b200/csrc/common.cuh  fast_exp / fast_exp_tail / fast_exp_scale
 *     dm_log   smc-nuts_b200/csrc/common.cuh  fast_log
 * Every operation of those algorithms is an IEEE-754 fma / mul / add or an integer operation on the bit pattern, so a
 * restatement with C99 fma() reproduces them bit for bit on a CPU.  With orc_set_devmath(1) the oracle's model and
 * NUTS functions call these in place of exp / log1p; the oracle then predicts the PARITY device build
 * (libsmcnuts_b200_parity.so: -fmad=false, the oracle's statement order) exactly, including every tree of the
 * chaotic PRMwCD trajectories, which glibc-vs-device last-bit differences of exp otherwise decorrelate.
 * The default mode (glibc) is the one pinned against the reference goldens; the two modes differ only by these two
 * functions, each within 1 ulp / 2.2e-16 of the true value (tests/test_oracle_golden.py).
 * Tables: oracle/devmath_tables.h, generated independently with mpmath by oracle/gen_devmath_tables.py.
 */
#ifndef SMC_ORACLE_DEVMATH_H
#define SMC_ORACLE_DEVMATH_H
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "devmath_tables.h"

static inline int32_t dm_hi(double v) { uint64_t b; memcpy(&b, &v, 8); return (int32_t)(uint32_t)(b >> 32); }
static inline int32_t dm_lo(double v) { uint64_t b; memcpy(&b, &v, 8); return (int32_t)(uint32_t)b; }
static inline double dm_make(int32_t hi, int32_t lo) {
    uint64_t b = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo;
    double v;
    memcpy(&v, &b, 8);
    return v;
}

/* exp(x): x = (32 k + j) ln2/32 + r, exp(x) = 2^k T[j] (1 + q(r)), q a degree-6 polynomial in Estrin form */
static inline double dm_exp(double x) {
    const double magic = 6755399441055744.0;          /* 1.5 * 2^52 */
    const double l_hi = 0x1.62e42fef80000p-6, l_lo = 0x1.1cf79abc9e3b4p-41;
    double t = fma(x, DM_EXP_INVL, magic);
    const int32_t kp = dm_lo(t);
    const double tj = DM_EXP_T[kp & 31];
    t -= magic;
    double r = fma(t, -l_hi, x);
    r = fma(t, -l_lo, r);
    const double s = r * r;
    const double b0 = fma(1.0 / 6, r, 0.5);
    double b1 = fma(1.0 / 120, r, 1.0 / 24);
    b1 = fma(1.0 / 720, s, b1);
    const double q = fma(s, fma(s, b1, b0), r);
    const double p = fma(tj, q, tj);
    const int32_t k = kp >> 5;                         /* arithmetic shift, as on the device */
    if ((dm_hi(x) & 0x7fffffff) >= 0x40862000) {       /* |x| >= 708, inf, nan: two exact scaling steps */
        const int32_t k1 = k >> 1, k2 = k - k1;
        double res = p * dm_make((k1 + 1023) << 20, 0) * dm_make((k2 + 1023) << 20, 0);
        if (x > 709.782712893384) res = INFINITY;
        if (x < -745.2) res = 0.0;
        return (x != x) ? x : res;
    }
    return dm_make(dm_hi(p) + (int32_t)((uint32_t)k << 20), dm_lo(p));
}

/* log(u) for positive normal u: u = 2^e m, j = top 6 mantissa bits, r = m / c_j - 1, degree-7 polynomial */
static inline double dm_log(double u) {
    const int32_t hi = dm_hi(u);
    if (hi < 0x00100000 || hi >= 0x7ff00000) return log(u);
    const int j = (hi >> 14) & 63;
    const double m = dm_make((hi & 0x000fffff) | 0x3ff00000, dm_lo(u));
    const double e = (double)((hi >> 20) - 1023);
    const double r = fma(m, DM_LOG_INVC[j], -1.0);
    double p = fma(1.0 / 7, r, -1.0 / 6);
    p = fma(p, r, 0.2); p = fma(p, r, -0.25); p = fma(p, r, 1.0 / 3); p = fma(p, r, -0.5);
    const double l1 = fma(r * r, p, r);
    return fma(e, 0x1.62e42fefa2000p-1, DM_LOG_C[j]) + fma(e, 0x1.9ef35793c7673p-41, l1);
}
#endif
