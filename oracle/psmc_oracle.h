/* oracle/psmc_oracle.h -- TEST INFRASTRUCTURE ONLY (see psmc_oracle.c header). */
#ifndef PSMC_ORACLE_H
#define PSMC_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_TINY 1e-25   /* khmm.h:28 HMM_TINY */
#define ORC_INF  1e300   /* khmm.h:29 HMM_INF  */
#define ORC_T_INF 1000.0 /* psmc.h:14 PSMC_T_INF */

typedef double (*orc_func_t)(int n, double *x, void *data);

int    orc_pattern(const char *pattern, int *n_free, int *par_map);
int    orc_update_hmm(int n, const int *par_map, const double *params, int n_params, double alpha0, int diverg,
                      const double *inp_ti,
                      double *a, double *e, double *a0, double *sigma, double *t, double *C_pi, double *C_sigma);
void   orc_avg_t(int n, const int *par_map, const double *params, int n_params, int diverg,
                 const double *t, const double *sigma, double C_pi, double C_sigma, double *avg_t);
int    orc_fwdbwd(int N, const double *a, const double *e, const double *a0, int L, const signed char *seq,
                  double *f, double *b, double *s);
double orc_lk(int L, const double *s);
int    orc_estep(int N, const double *a, const double *e, const double *a0,
                 int n_seqs, const int *L, const signed char *seqs,
                 double *LL, double *A, double *E, double *A0);
int    orc_decode(int N, const double *a, const double *e, const double *a0, int L, const signed char *seq,
                  int *best_k, double *best_p, double *post, double *p_recomb);
double orc_Q0(int N, const double *A, const double *E);
double orc_Q(int N, const double *a, const double *e, const double *A, const double *E, double Q0);
double orc_hj(orc_func_t func, int n, double *x, void *data, double r, double eps, int max_calls);
void   orc_struct_stats(int N, const double *A, double *RL, double *CL, double *RU, double *CU, double *AD);
int    orc_factors(int N, const double *a, double *U, double *V, double *W, double *Z, double *D);

#ifdef __cplusplus
}
#endif
#endif
