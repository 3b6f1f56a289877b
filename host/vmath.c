/* vmath.c -- the only translation unit built with -O3 -ffast-math: plain loops over exp/log that gcc turns into glibc
 * libmvec calls (<= 4 ulp).  Every function is compiled three times (AVX-512: _ZGVeN8v_*, 8 doubles per call; AVX2+FMA:
 * _ZGVdN4v_*; baseline) and the loader picks the best the host CPU supports (ifunc).  Used ONLY inside the
 * Hooke-Jeeves trial evaluations (psmch_model_update_fast / psmch_Q_fast); the model that is printed and sent to
 * the GPU is always recomputed with the scalar, bit-reproducible path. */
#include <math.h>

#if defined(__x86_64__)
#define VCLONES __attribute__((target_clones("avx512f", "avx2,fma", "default")))
#else
#define VCLONES /* other hosts (aarch64: Grace): one clone, the compiler's own vector math */
#endif

VCLONES void psmch_vexp(int n, const double *x, double *y)
{
	int i;
	for (i = 0; i < n; ++i) y[i] = exp(x[i]);
}

VCLONES void psmch_vlog(int n, const double *x, double *y)
{
	int i;
	for (i = 0; i < n; ++i) y[i] = log(x[i]);
}

VCLONES double psmch_vdot(int n, const double *a, const double *b)
{
	double s = 0.0;
	int i;
	for (i = 0; i < n; ++i) s += a[i] * b[i];
	return s;
}

/* dependency-free per-interval arithmetic of the model update (core.c:100-122 regrouped), vectorised by gcc;
 * sum_t[k] = t_k - t_0 is passed in.  Outputs sigma, the five factor arrays and the argument of the avg_t logarithm. */
VCLONES void psmch_vmodel_phase1(int N, const double *alp, const double *lam, const double *tau, const double *bet,
                         const double *qax, const double *sumt, double C_pi, double rho, double C_sigma,
                         double *sigma, double *U, double *V, double *W, double *Z, double *D, double *logarg)
{
	int k;
	const double icr = 1.0 / (C_pi * rho);
	for (k = 0; k < N; ++k) {
		const double ak1 = alp[k] - alp[k + 1], lak = lam[k];
		const double cpik = ak1 * (sumt[k] + lak) - alp[k + 1] * tau[k];
		const double pik = cpik / C_pi;
		const double sg = (ak1 * icr + pik * 0.5) / C_sigma;
		const double tmp = pik / (C_sigma * sg);
		const double qkk = (ak1 * ak1 * (bet[k] - lak / alp[k]) + 2 * lak * ak1 - 2 * alp[k + 1] * tau[k]) / cpik;
		sigma[k] = sg;
		logarg[k] = 1.0 - tmp;
		U[k] = tmp * (ak1 / cpik);
		V[k] = qax[k];
		W[k] = tmp * (qax[k] / cpik);
		Z[k] = ak1;
		D[k] = tmp * qkk + (1.0 - tmp);
	}
}

/* q_aux (core.c:93-94) and 1/alpha, vectorised */
VCLONES void psmch_vmodel_qaux(int n, const double *alp, const double *lam, const double *tau, const double *bet, double *qax)
{
	int l;
	for (l = 0; l < n; ++l) qax[l] = (alp[l] - alp[l + 1]) * (bet[l] - lam[l] / alp[l]) + tau[l];
}

VCLONES void psmch_vinv(int n, const double *x, double *y)
{
	int i;
	for (i = 0; i < n; ++i) y[i] = 1.0 / x[i];
}
