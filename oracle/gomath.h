/* oracle/gomath.h -- TEST INFRASTRUCTURE ONLY. See gomath.c. */
#ifndef ORACLE_GOMATH_H
#define ORACLE_GOMATH_H
#ifdef __cplusplus
extern "C" {
#endif
double gm_frexp(double f, int *e);
double gm_ldexp(double frac, int e);
double gm_log(double x);
double gm_log2(double x);
double gm_log10(double x);
double gm_exp(double x);
double gm_pow(double x, double y);
double gm_lgamma(double x);
double gm_round(double x);
#ifdef __cplusplus
}
#endif
#endif
