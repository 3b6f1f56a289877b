/* model.c -- population parameters -> PSMC HMM, in FACTORED form.
 *
 * The reference fills a dense (n+1)x(n+1) matrix on every objective evaluation (core.c:112-122,
 * O(N^2)).  The matrix is diagonal + rank-1 strictly lower + rank-1 strictly upper, so only the
 * O(N) factors are computed here:
 *     a[k][l] = U_k V_l (l<k)     U_k = tmp_k * (ak1_k / cpik_k)      V_l = q_aux[l]
 *     a[k][l] = W_k Z_l (l>k)     W_k = tmp_k * (q_aux[k] / cpik_k)   Z_l = alpha_l - alpha_{l+1}
 *     a[k][k] = D_k = tmp_k * q_kk + (1 - tmp_k),   tmp_k = pi_k / (C_sigma sigma_k)
 * with the same intermediate quantities and evaluation order as core.c:61-133 (so sigma, t, C_pi,
 * C_sigma, e are bit-identical and the factors agree with the dense entries to 1 ulp).
 * params = [theta, rho, max_t, lambda_free..., (dt)]  (core.c:26,40-48). */
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include "psmc_host.h"

int psmch_model_alloc(psmch_model_t *m, const psmch_space_t *sp)
{
	int N = sp->n + 1;
	double *blk;
	memset(m, 0, sizeof(*m));
	m->N = N;
	blk = (double*)calloc((size_t)sp->n_params + (N + 1) + 13 * (size_t)N + 8 + 4 * ((size_t)N + 2) + 21 * (size_t)N + 8, sizeof(double));
	if (blk == 0) return -1;
	m->params = blk; blk += sp->n_params;
	m->t = blk; blk += N + 1;
	m->sigma = blk; blk += N;
	m->e = blk; blk += 2 * N;
	m->U = blk; blk += N; m->V = blk; blk += N; m->W = blk; blk += N; m->Z = blk; blk += N; m->D = blk; blk += N;
	m->lam = blk; blk += N; m->alp = blk; blk += N + 1; m->bet = blk; blk += N; m->qax = blk; blk += N; m->tau = blk; blk += N + 1;
	m->vw = blk; /* 4*(N+2) + 21*N doubles of scratch for the vectorised trial evaluations */
	m->t_max_t = NAN;
	return 0;
}

void psmch_model_free(psmch_model_t *m)
{
	free(m->params);
	memset(m, 0, sizeof(*m));
}

void psmch_model_update(const psmch_space_t *sp, const double *params, psmch_model_t *m)
{
	const int n = sp->n, N = n + 1;
	int k, l;
	double theta, rho, max_t, dt = 0.0, sum_t, C_pi, C_sigma;
	double *lam = m->lam, *alp = m->alp, *bet = m->bet, *qax = m->qax, *tau = m->tau, *t = m->t;
	if (params != m->params) memcpy(m->params, params, sizeof(double) * sp->n_params);
	m->fast_valid = 0; m->t_max_t = NAN; /* t[] and e[] below come from scalar libm: the fast path's caches do not apply */
	theta = params[0]; rho = params[1]; max_t = params[2];
	for (k = 0; k < N; ++k) lam[k] = params[sp->par_map[k] + PSMCH_N_PARAMS];
	if (sp->inp_ti == 0) {                               /* core.c:9-14 */
		double beta = log(1.0 + max_t / sp->alpha0) / n;
		for (k = 0; k < n; ++k) t[k] = sp->alpha0 * (exp(beta * k) - 1);
		t[n] = max_t; t[n + 1] = PSMCH_T_INF;
	} else {                                             /* core.c:15-18 */
		memcpy(t, sp->inp_ti, sizeof(double) * (n + 1));
		t[n + 1] = PSMCH_T_INF;
	}
	if (sp->diverg) { dt = params[sp->n_params - 1]; if (dt < 0) dt = 0; }
	for (k = 0; k < N; ++k) tau[k] = t[k + 1] - t[k];
	alp[0] = 1.0;
	for (k = 1; k < N; ++k) alp[k] = alp[k - 1] * exp(-tau[k - 1] / lam[k - 1]);
	alp[N] = 0.0;
	bet[0] = 0.0;
	for (k = 1; k < N; ++k) bet[k] = bet[k - 1] + lam[k - 1] * (1.0 / alp[k] - 1.0 / alp[k - 1]);
	for (l = 0; l < n; ++l) qax[l] = (alp[l] - alp[l + 1]) * (bet[l] - lam[l] / alp[l]) + tau[l];
	qax[n] = 0.0;
	for (l = 0, C_pi = 0.0; l < N; ++l) C_pi += lam[l] * (alp[l] - alp[l + 1]);
	C_sigma = 1.0 / (C_pi * rho) + 0.5;
	m->C_pi = C_pi; m->C_sigma = C_sigma;
	for (k = 0, sum_t = 0.0; k < N; ++k) {
		const double ak1 = alp[k] - alp[k + 1], lak = lam[k];
		const double cpik = ak1 * (sum_t + lak) - alp[k + 1] * tau[k];
		const double pik = cpik / C_pi;
		double avg_t, tmp, qkk;
		m->sigma[k] = (ak1 / (C_pi * rho) + pik / 2.0) / C_sigma;
		avg_t = -log(1.0 - pik / (C_sigma * m->sigma[k])) / rho;
		if (isnan(avg_t) || avg_t < sum_t || avg_t > sum_t + tau[k]) /* core.c:109-110 */
			avg_t = sum_t + (lak - tau[k] * alp[k + 1] / (alp[k] - alp[k + 1]));
		qkk = (ak1 * ak1 * (bet[k] - lak / alp[k]) + 2 * lak * ak1 - 2 * alp[k + 1] * tau[k]) / cpik;
		tmp = pik / (C_sigma * m->sigma[k]);
		m->U[k] = tmp * (ak1 / cpik);
		m->V[k] = qax[k];
		m->W[k] = tmp * (qax[k] / cpik);
		m->Z[k] = ak1;
		m->D[k] = tmp * qkk + (1.0 - tmp);
		m->e[k] = exp(-theta * (avg_t + dt));
		m->e[N + k] = 1.0 - m->e[k];
		sum_t += tau[k];
	}
}

/* Same model, evaluated in array phases so that every exp/log goes through the vectorised loops of vmath.c
 * (results within a few ulp of psmch_model_update; used for the ~4000 trial points of one M-step only).
 * m->vw must provide 4*N+8 doubles of scratch. */
void psmch_vexp(int n, const double *x, double *y);
void psmch_vlog(int n, const double *x, double *y);
void psmch_vinv(int n, const double *x, double *y);
void psmch_vmodel_qaux(int n, const double *alp, const double *lam, const double *tau, const double *bet, double *qax);
void psmch_vmodel_phase1(int N, const double *alp, const double *lam, const double *tau, const double *bet, const double *qax,
                         const double *sumt, double C_pi, double rho, double C_sigma, double *sigma, double *U, double *V,
                         double *W, double *Z, double *D, double *logarg);

void psmch_model_update_fast(const psmch_space_t *sp, const double *params, psmch_model_t *m)
{
	const int n = sp->n, N = n + 1;
	int k, l;
	double theta, rho, max_t, dt = 0.0, sum_t, C_pi, C_sigma;
	double *lam = m->lam, *alp = m->alp, *bet = m->bet, *qax = m->qax, *tau = m->tau, *t = m->t;
	double *w0 = m->vw, *w1 = m->vw + (N + 2), *w2 = m->vw + 2 * (N + 2), *w3 = m->vw + 3 * (N + 2);
	if (params != m->params) memcpy(m->params, params, sizeof(double) * sp->n_params);
	theta = params[0]; rho = params[1]; max_t = params[2];
	for (k = 0; k < N; ++k) lam[k] = params[sp->par_map[k] + PSMCH_N_PARAMS];
	if (sp->inp_ti == 0) {
		if (!(max_t == m->t_max_t)) { /* most trial points move one lambda or theta/rho: the boundaries stay */
			const double beta = log(1.0 + max_t / sp->alpha0) / n;
			for (k = 0; k < n; ++k) w0[k] = beta * k;
			psmch_vexp(n, w0, w1);
			for (k = 0; k < n; ++k) t[k] = sp->alpha0 * (w1[k] - 1);
			t[n] = max_t; t[n + 1] = PSMCH_T_INF;
			m->t_max_t = max_t;
		}
	} else {
		memcpy(t, sp->inp_ti, sizeof(double) * (n + 1));
		t[n + 1] = PSMCH_T_INF;
	}
	if (sp->diverg) { dt = params[sp->n_params - 1]; if (dt < 0) dt = 0; }
	for (k = 0; k < N; ++k) tau[k] = t[k + 1] - t[k];
	for (k = 1; k < N; ++k) w0[k] = -tau[k - 1] / lam[k - 1];
	psmch_vexp(N - 1, w0 + 1, w1 + 1);
	alp[0] = 1.0;
	for (k = 1; k < N; ++k) alp[k] = alp[k - 1] * w1[k];
	alp[N] = 0.0;
	psmch_vinv(N, alp, w0); /* 1/alpha_k, k < N */
	bet[0] = 0.0;
	for (k = 1; k < N; ++k) bet[k] = bet[k - 1] + lam[k - 1] * (w0[k] - w0[k - 1]);
	psmch_vmodel_qaux(n, alp, lam, tau, bet, qax);
	qax[n] = 0.0;
	for (l = 0, C_pi = 0.0; l < N; ++l) C_pi += lam[l] * (alp[l] - alp[l + 1]);
	C_sigma = 1.0 / (C_pi * rho) + 0.5;
	m->C_pi = C_pi; m->C_sigma = C_sigma;
	/* phase 1: everything up to the argument of the logarithm (vectorised; w2 = interval starts) */
	for (k = 0, sum_t = 0.0; k < N; ++k) { w2[k] = sum_t; sum_t += tau[k]; }
	psmch_vmodel_phase1(N, alp, lam, tau, bet, qax, w2, C_pi, rho, C_sigma, m->sigma, m->U, m->V, m->W, m->Z, m->D, w0);
	psmch_vlog(N, w0, w1);
	/* phase 2: avg_t with the reference's fallback (core.c:109-110), then the emissions */
	for (k = 0; k < N; ++k) {
		double avg_t = -w1[k] / rho;
		if (isnan(avg_t) || avg_t < w2[k] || avg_t > w2[k] + tau[k])
			avg_t = w2[k] + (lam[k] - tau[k] * alp[k + 1] / (alp[k] - alp[k + 1]));
		w3[k] = -theta * (avg_t + dt);
	}
	psmch_vexp(N, w3, m->e);
	for (k = 0; k < N; ++k) m->e[N + k] = 1.0 - m->e[k];
	m->fast_valid = 1; /* w3 = log e[0][.] for psmch_Q_fast */
}

/* dense view (tests, diagnostics) */
void psmch_model_dense(const psmch_model_t *m, double *a)
{
	int N = m->N, k, l;
	for (k = 0; k < N; ++k)
		for (l = 0; l < N; ++l)
			a[(size_t)k * N + l] = l < k ? m->U[k] * m->V[l] : (l > k ? m->W[k] * m->Z[l] : m->D[k]);
}

/* average coalescent time per interval (used by the decoder's TC/DC lines); behaviour of core.c:135-162 */
void psmch_avg_t(const psmch_space_t *sp, const psmch_model_t *m, double *avg_t)
{
	const int N = sp->n + 1;
	const double rho = m->params[1];
	double dt = 0.0, sum_t = 0.0;
	int k;
	/* lam/alp/tau still describe m->params (psmch_model_update keeps them) */
	if (sp->diverg) { dt = m->params[sp->n_params - 1]; if (dt < 0) dt = 0; }
	for (k = 0; k < N; ++k) {
		const double ak1 = m->alp[k] - m->alp[k + 1], lak = m->lam[k];
		const double pik = (ak1 * (sum_t + lak) - m->alp[k + 1] * m->tau[k]) / m->C_pi;
		double v = -log(1.0 - pik / (m->C_sigma * m->sigma[k])) / rho;
		if (isnan(v) || v < sum_t || v > sum_t + m->tau[k])
			v = sum_t + (lak - m->tau[k] * m->alp[k + 1] / (m->alp[k] - m->alp[k + 1]));
		avg_t[k] = v + dt;
		sum_t += m->tau[k];
	}
}

void psmch_model_view(const psmch_model_t *m, psmc_b200_model *v)
{
	v->n_states = m->N;
	v->a0 = m->sigma; v->e = m->e;
	v->U = m->U; v->V = m->V; v->W = m->W; v->Z = m->Z; v->D = m->D;
}
