/*
 * oracle/gomath.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * CPU restatement of the Go standard-library `math` functions that the elPrep
 * 5.1.3 BQSR path calls (reference call sites: filters/bqsr.go:562,566,599,
 * 600,611,625,662,705,785,995; filters/unpedantic.go:28-30;
 * filters/mark-optical-duplicates.go:536,578).
 *
 * The Go toolchain and its sources are NOT available in this image, so these
 * are restatements of the published algorithms Go's pure-Go implementations
 * follow (FreeBSD msun e_log.c, e_exp.c, e_lgamma_r.c; Go's own pow.go,
 * log10.go, floor.go).  PARITY UNPINNED at the last-ulp level: on amd64 the
 * real Go runtime uses an assembly Exp (optionally FMA), which can differ from
 * the pure algorithm in the last bit.  Every constant below was cross-checked
 * decimal-vs-hex (see tests/test_oracle_gomath.py).
 *
 * Build with -ffp-contract=off (amd64 Go never fuses multiply-add).
 */
#include <math.h>
#include <stdint.h>
#include <string.h>
#include "gomath.h"

static inline uint64_t f2u(double x) { uint64_t u; memcpy(&u, &x, 8); return u; }
static inline double u2f(uint64_t u) { double x; memcpy(&x, &u, 8); return x; }

/* math.Frexp (frexp.go): x = frac * 2**exp, 0.5 <= |frac| < 1 */
double gm_frexp(double f, int *e) {
    *e = 0;
    if (f == 0 || isinf(f) || isnan(f)) return f;
    /* normalize subnormals */
    if (fabs(f) < 2.2250738585072014e-308) { f *= (double)(1ULL << 52); *e = -52; }
    uint64_t x = f2u(f);
    *e += (int)((x >> 52) & 0x7ff) - 1022;
    x &= ~((uint64_t)0x7ff << 52);
    x |= (uint64_t)1022 << 52;
    return u2f(x);
}

/* math.Ldexp (ldexp.go) */
double gm_ldexp(double frac, int e) {
    if (frac == 0 || isinf(frac) || isnan(frac)) return frac;
    int fe;
    frac = gm_frexp(frac, &fe);
    /* gm_frexp returns frac in [0.5,1) => biased exponent field 1022 */
    e += fe;
    uint64_t x = f2u(frac);
    e += (int)((x >> 52) & 0x7ff) - 1023;
    if (e < -1075) return copysign(0.0, frac);
    if (e > 1023) return frac < 0 ? -INFINITY : INFINITY;
    double m = 1;
    if (e < -1022) { e += 53; m = 1.0 / (double)(1ULL << 53); }
    x &= ~((uint64_t)0x7ff << 52);
    x |= (uint64_t)(e + 1023) << 52;
    return m * u2f(x);
}

/* math.Modf for non-negative input (modf.go) */
static double gm_modf_pos(double f, double *frac) {
    if (f < 1) { *frac = f; return 0; }
    uint64_t x = f2u(f);
    unsigned e = (unsigned)((x >> 52) & 0x7ff) - 1023;
    if (e < 52) x &= ~(((uint64_t)1 << (52 - e)) - 1);
    double ip = u2f(x);
    *frac = f - ip;
    return ip;
}

/* math.Log (log.go; FreeBSD e_log.c) */
double gm_log(double x) {
    static const double Ln2Hi = 6.93147180369123816490e-01, Ln2Lo = 1.90821492927058770002e-10,
        L1 = 6.666666666666735130e-01, L2 = 3.999999999940941908e-01, L3 = 2.857142874366239149e-01,
        L4 = 2.222219843214978396e-01, L5 = 1.818357216161805012e-01, L6 = 1.531383769920937332e-01,
        L7 = 1.479819860511658591e-01;
    if (isnan(x) || (isinf(x) && x > 0)) return x;
    if (x < 0) return NAN;
    if (x == 0) return -INFINITY;
    int ki;
    double f1 = gm_frexp(x, &ki);
    if (f1 < 1.41421356237309504880168872420969808 / 2) { f1 *= 2; ki--; }
    double f = f1 - 1;
    double k = (double)ki;
    double s = f / (2 + f);
    double s2 = s * s;
    double s4 = s2 * s2;
    double t1 = s2 * (L1 + s4 * (L3 + s4 * (L5 + s4 * L7)));
    double t2 = s4 * (L2 + s4 * (L4 + s4 * L6));
    double R = t1 + t2;
    double hfsq = 0.5 * f * f;
    return k * Ln2Hi - ((hfsq - (s * (hfsq + R) + k * Ln2Lo)) - f);
}

/* math.Log2 / math.Log10 (log10.go) */
double gm_log2(double x) {
    int e;
    double frac = gm_frexp(x, &e);
    if (frac == 0.5) return (double)(e - 1);
    return gm_log(frac) * 1.4426950408889634 /* 1/Ln2 */ + (double)e;
}
double gm_log10(double x) { return gm_log2(x) * 0.3010299956639812 /* Ln2/Ln10 */; }

/* math.Exp (exp.go pure-Go version; FreeBSD e_exp.c) */
double gm_exp(double x) {
    static const double Ln2Hi = 6.93147180369123816490e-01, Ln2Lo = 1.90821492927058770002e-10,
        Log2e = 1.44269504088896338700e+00, Overflow = 7.09782712893383973096e+02,
        Underflow = -7.45133219101941108420e+02, NearZero = 1.0 / (1 << 28);
    static const double P1 = 1.66666666666666657415e-01, P2 = -2.77777777770155933842e-03,
        P3 = 6.61375632143793436117e-05, P4 = -1.65339022054652515390e-06, P5 = 4.13813679705723846039e-08;
    if (isnan(x) || (isinf(x) && x > 0)) return x;
    if (isinf(x)) return 0;
    if (x > Overflow) return INFINITY;
    if (x < Underflow) return 0;
    if (-NearZero < x && x < NearZero) return 1 + x;
    int k = 0;
    if (x < 0) k = (int)(Log2e * x - 0.5);
    else if (x > 0) k = (int)(Log2e * x + 0.5);
    double hi = x - (double)k * Ln2Hi;
    double lo = (double)k * Ln2Lo;
    double r = hi - lo;
    double t = r * r;
    double c = r - t * (P1 + t * (P2 + t * (P3 + t * (P4 + t * P5))));
    double y = 1 - ((lo - (r * c) / (2 - c)) - hi);
    return gm_ldexp(y, k);
}

/* math.Pow (pow.go) */
double gm_pow(double x, double y) {
    if (y == 0 || x == 1) return 1;
    if (y == 1) return x;
    if (isnan(x) || isnan(y)) return NAN;
    if (x == 0) {
        if (y < 0) return INFINITY; /* sign handling for odd integers not needed on this path */
        return 0;
    }
    if (isinf(y)) {
        if (x == -1) return 1;
        if ((fabs(x) < 1) == (y > 0)) return 0;
        return INFINITY;
    }
    if (isinf(x)) {
        if (x < 0) return NAN; /* not reached on this path */
        return y < 0 ? 0 : INFINITY;
    }
    if (y == 0.5) return sqrt(x);
    if (y == -0.5) return 1 / sqrt(x);
    double yf;
    double yi = gm_modf_pos(fabs(y), &yf);
    if (yf != 0 && x < 0) return NAN;
    if (yi >= 9.223372036854775808e18) {
        if (x == -1) return 1;
        if ((fabs(x) < 1) == (y > 0)) return 0;
        return INFINITY;
    }
    double a1 = 1.0;
    int ae = 0;
    if (yf != 0) {
        if (yf > 0.5) { yf--; yi++; }
        a1 = gm_exp(yf * gm_log(x));
    }
    int xe;
    double x1 = gm_frexp(x, &xe);
    for (int64_t i = (int64_t)yi; i != 0; i >>= 1) {
        if (xe < -(1 << 12) || (1 << 12) < xe) { ae += xe; break; }
        if (i & 1) { a1 *= x1; ae += xe; }
        x1 *= x1;
        xe <<= 1;
        if (x1 < .5) { x1 += x1; xe--; }
    }
    if (y < 0) { a1 = 1 / a1; ae = -ae; }
    return gm_ldexp(a1, ae);
}

/* math.Lgamma for x > 0 (lgamma.go; FreeBSD e_lgamma_r.c). Sign is always +1 there. */
static const double lgA[] = {7.72156649015328655494e-02, 3.22467033424113591611e-01, 6.73523010531292681824e-02,
    2.05808084325167332806e-02, 7.38555086081402883957e-03, 2.89051383673415629091e-03, 1.19270763183362067845e-03,
    5.10069792153511336608e-04, 2.20862790713908385557e-04, 1.08011567247583939954e-04, 2.52144565451257326939e-05,
    4.48640949618915160150e-05};
static const double lgR[] = {1.0, 1.39200533467621045958e+00, 7.21935547567138069525e-01, 1.71933865632803078993e-01,
    1.86459191715652901344e-02, 7.77942496381893596434e-04, 7.32668430744625636189e-06};
static const double lgS[] = {-7.72156649015328655494e-02, 2.14982415960608852501e-01, 3.25778796408930981787e-01,
    1.46350472652464452805e-01, 2.66422703033638609560e-02, 1.84028451407337715652e-03, 3.19475326584100867617e-05};
static const double lgT[] = {4.83836122723810047042e-01, -1.47587722994593911752e-01, 6.46249402391333854778e-02,
    -3.27885410759859649565e-02, 1.79706750811820387126e-02, -1.03142241298341437450e-02, 6.10053870246291332635e-03,
    -3.68452016781138256760e-03, 2.25964780900612472250e-03, -1.40346469989232843813e-03, 8.81081882437654011382e-04,
    -5.38595305356740546715e-04, 3.15632070903625950361e-04, -3.12754168375120860518e-04, 3.35529192635519073543e-04};
static const double lgU[] = {-7.72156649015328655494e-02, 6.32827064025093366517e-01, 1.45492250137234768737e+00,
    9.77717527963372745603e-01, 2.28963728064692451092e-01, 1.33810918536787660377e-02};
static const double lgV[] = {1.0, 2.45597793713041134822e+00, 2.12848976379893395361e+00, 7.69285150456672783825e-01,
    1.04222645593369134254e-01, 3.21709242282423911810e-03};
static const double lgW[] = {4.18938533204672725052e-01, 8.33333333333329678849e-02, -2.77777777728775536470e-03,
    7.93650558643019558500e-04, -5.95187557450339963135e-04, 8.36339918996282139126e-04, -1.63092934096575273989e-03};

double gm_lgamma(double x) {
    static const double Ymin = 1.461632144968362245, Two58 = 288230376151711744.0, Tiny = 1.0 / 1180591620717411303424.0,
        Tc = 1.46163214496836224576e+00, Tf = -1.21486290535849611461e-01, Tt = -3.63867699703950536541e-18;
    if (isnan(x) || isinf(x)) return x;
    if (x == 0) return INFINITY;
    if (x < 0) return NAN; /* negative branch (sinPi reflection) is unreachable on the elPrep path */
    if (x < Tiny) return -gm_log(x);
    double lg;
    if (x == 1 || x == 2) return 0;
    if (x < 2) {
        double y; int i;
        if (x <= 0.9) {
            lg = -gm_log(x);
            if (x >= (Ymin - 1 + 0.27)) { y = 1 - x; i = 0; }
            else if (x >= (Ymin - 1 - 0.23)) { y = x - (Tc - 1); i = 1; }
            else { y = x; i = 2; }
        } else {
            lg = 0;
            if (x >= (Ymin + 0.27)) { y = 2 - x; i = 0; }
            else if (x >= (Ymin - 0.23)) { y = x - Tc; i = 1; }
            else { y = x - 1; i = 2; }
        }
        if (i == 0) {
            double z = y * y;
            double p1 = lgA[0] + z * (lgA[2] + z * (lgA[4] + z * (lgA[6] + z * (lgA[8] + z * lgA[10]))));
            double p2 = z * (lgA[1] + z * (+lgA[3] + z * (lgA[5] + z * (lgA[7] + z * (lgA[9] + z * lgA[11])))));
            double p = y * p1 + p2;
            lg += (p - 0.5 * y);
        } else if (i == 1) {
            double z = y * y, w = z * y;
            double p1 = lgT[0] + w * (lgT[3] + w * (lgT[6] + w * (lgT[9] + w * lgT[12])));
            double p2 = lgT[1] + w * (lgT[4] + w * (lgT[7] + w * (lgT[10] + w * lgT[13])));
            double p3 = lgT[2] + w * (lgT[5] + w * (lgT[8] + w * (lgT[11] + w * lgT[14])));
            double p = z * p1 - (Tt - w * (p2 + y * p3));
            lg += (Tf + p);
        } else {
            double p1 = y * (lgU[0] + y * (lgU[1] + y * (lgU[2] + y * (lgU[3] + y * (lgU[4] + y * lgU[5])))));
            double p2 = 1 + y * (lgV[1] + y * (lgV[2] + y * (lgV[3] + y * (lgV[4] + y * lgV[5]))));
            lg += (-0.5 * y + p1 / p2);
        }
        return lg;
    }
    if (x < 8) {
        int i = (int)x;
        double y = x - (double)i;
        double p = y * (lgS[0] + y * (lgS[1] + y * (lgS[2] + y * (lgS[3] + y * (lgS[4] + y * (lgS[5] + y * lgS[6]))))));
        double q = 1 + y * (lgR[1] + y * (lgR[2] + y * (lgR[3] + y * (lgR[4] + y * (lgR[5] + y * lgR[6])))));
        lg = 0.5 * y + p / q;
        double z = 1.0;
        switch (i) {
        case 7: z *= (y + 6); /* fallthrough */
        case 6: z *= (y + 5); /* fallthrough */
        case 5: z *= (y + 4); /* fallthrough */
        case 4: z *= (y + 3); /* fallthrough */
        case 3: z *= (y + 2); lg += gm_log(z);
        }
        return lg;
    }
    if (x < Two58) {
        double t = gm_log(x);
        double z = 1 / x;
        double y = z * z;
        double w = lgW[0] + z * (lgW[1] + y * (lgW[2] + y * (lgW[3] + y * (lgW[4] + y * (lgW[5] + y * lgW[6])))));
        return (x - 0.5) * (t - 1) + w;
    }
    return x * (gm_log(x) - 1);
}

/* math.Round: half away from zero (floor.go) */
double gm_round(double x) { return round(x); }
