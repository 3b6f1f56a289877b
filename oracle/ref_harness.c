/* oracle/ref_harness.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Flat-array (ctypes-friendly) entry points around the UNMODIFIED reference objects
 * (khmm.o core.o cli.o em.o aux.o kmin.o compiled from /root/reference by oracle/Makefile).
 * Nothing here restates the algorithm: every number comes out of the reference's own functions
 *   hmm_forward / hmm_backward / hmm_lk / hmm_expect / hmm_add_expect   (khmm.c:145-359)
 *   hmm_post_decode / hmm_post_state                                   (khmm.c:264-293)
 *   psmc_update_hmm / psmc_new_data / psmc_avg_t                       (core.c:21-162)
 *   psmc_parse_pattern                                                 (cli.c:66-99)
 *   hmm_Q0 / hmm_Q                                                     (khmm.c:326-382)
 *   kmin_hj                                                            (kmin.c:68-107)
 * The harness only packs/unpacks the reference's pointer-of-pointer structs.
 * Built into oracle/_ref/libpsmcref.so; used by tests/ and by bench.py's reference arm.
 */
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include "psmc.h"
#include "khmm.h"
#include "kmin.h"

/* ---- model: params -> (a, e, a0, sigma, t, C_pi, C_sigma) via psmc_update_hmm ---------------- */

static psmc_par_t *mk_par(const char *pattern, double alpha, int diverg)
{
	psmc_par_t *pp = (psmc_par_t*)calloc(1, sizeof(psmc_par_t));
	pp->pattern = strdup(pattern);
	pp->par_map = psmc_parse_pattern(pattern, &pp->n_free, &pp->n);
	pp->alpha = alpha;
	if (diverg) pp->flag |= PSMC_F_DIVERG;
	return pp;
}
static void rm_par(psmc_par_t *pp)
{
	free(pp->pattern); free(pp->par_map); free(pp);
}

/* returns n (number of states is n+1); n_free written through pointer */
int ref_pattern(const char *pattern, int *n_free, int *par_map /* may be NULL; >= n+1 ints */)
{
	int n, nf, *pm, k;
	pm = psmc_parse_pattern(pattern, &nf, &n);
	if (n_free) *n_free = nf;
	if (par_map) for (k = 0; k <= n; ++k) par_map[k] = pm[k];
	free(pm);
	return n;
}

/* params layout: [theta, rho, max_t, lambda_free..., (dt)]  (core.c:26,40-48) */
int ref_update_hmm(const char *pattern, const double *params, double alpha, int diverg,
                   double *a /* N*N row-major */, double *e /* 2*N */, double *a0 /* N */,
                   double *sigma /* N */, double *t /* N+1 */, double *C_pi, double *C_sigma)
{
	psmc_par_t *pp = mk_par(pattern, alpha, diverg);
	psmc_data_t *pd = (psmc_data_t*)calloc(1, sizeof(psmc_data_t));
	int n = pp->n, N = n + 1, k, l;
	pd->n_params = pp->n_free + PSMC_N_PARAMS + (diverg ? 1 : 0);
	pd->hp = hmm_new_par(2, N);
	pd->sigma = (FLOAT*)calloc(N, sizeof(FLOAT));
	pd->post_sigma = (FLOAT*)calloc(N, sizeof(FLOAT));
	pd->t = (FLOAT*)malloc(sizeof(FLOAT) * (n + 2));
	pd->params = (FLOAT*)calloc(pd->n_params, sizeof(FLOAT));
	memcpy(pd->params, params, sizeof(FLOAT) * pd->n_params);
	psmc_update_hmm(pp, pd);
	for (k = 0; k < N; ++k) {
		for (l = 0; l < N; ++l) a[k * N + l] = pd->hp->a[k][l];
		e[k] = pd->hp->e[0][k]; e[N + k] = pd->hp->e[1][k];
		a0[k] = pd->hp->a0[k];
		if (sigma) sigma[k] = pd->sigma[k];
	}
	if (t) for (k = 0; k <= n; ++k) t[k] = pd->t[k];
	if (C_pi) *C_pi = pd->C_pi;
	if (C_sigma) *C_sigma = pd->C_sigma;
	psmc_delete_data(pd);
	rm_par(pp);
	return N;
}

/* ---- E-step: exactly the loop of psmc_em (em.c:33-55) on a caller-supplied dense model -------- */

static hmm_par_t *mk_hp(int N, const double *a, const double *e, const double *a0)
{
	hmm_par_t *hp = hmm_new_par(2, N);
	int k, l;
	for (k = 0; k < N; ++k) {
		for (l = 0; l < N; ++l) hp->a[k][l] = a[k * N + l];
		hp->e[0][k] = e[k]; hp->e[1][k] = e[N + k];
		hp->a0[k] = a0[k];
	}
	return hp;
}

/* seqs: concatenation of all sequences, values 0/1/2, lengths in L[].
 * Outputs: LL (sum over sequences), A[N*N], E[2*N] (rows b<m only, khmm.c:355), A0[N], Q0 (hmm_Q0 on the sum). */
int ref_estep(int N, const double *a, const double *e, const double *a0,
              int n_seqs, const int *L, const char *seqs,
              double *LL, double *A, double *E, double *A0, double *Q0)
{
	hmm_par_t *hp = mk_hp(N, a, e, a0);
	hmm_exp_t *he_sum = hmm_new_exp(hp);
	const char *p = seqs;
	double ll = 0.0;
	int i, k, l;
	hmm_pre_backward(hp);
	for (i = 0; i < n_seqs; ++i) {
		hmm_exp_t *he;
		hmm_data_t *hd;
		char *seq = (char*)calloc(L[i] + 1, 1);
		memcpy(seq, p, L[i]); p += L[i];
		hd = hmm_new_data(L[i], seq, hp);
		hmm_forward(hp, hd);
		hmm_backward(hp, hd);
		ll += hmm_lk(hd);
		he = hmm_expect(hp, hd);
		hmm_add_expect(he, he_sum);
		hmm_delete_exp(he);
		hmm_delete_data(hd);
		free(seq);
	}
	if (Q0) *Q0 = hmm_Q0(hp, he_sum);
	*LL = ll;
	for (k = 0; k < N; ++k) {
		for (l = 0; l < N; ++l) A[k * N + l] = he_sum->A[k][l];
		E[k] = he_sum->E[0][k]; E[N + k] = he_sum->E[1][k];
		if (A0) A0[k] = he_sum->A0[k];
	}
	hmm_delete_exp(he_sum);
	hmm_delete_par(hp);
	return 0;
}

/* ---- per-bin forward/backward of ONE sequence (small cases only): f,b are L*N, s is L (0-based) ---- */
int ref_fwdbwd(int N, const double *a, const double *e, const double *a0, int L, const char *seq,
               double *f, double *b, double *s)
{
	hmm_par_t *hp = mk_hp(N, a, e, a0);
	hmm_data_t *hd;
	char *sq = (char*)calloc(L + 1, 1);
	int u, k;
	memcpy(sq, seq, L);
	hmm_pre_backward(hp);
	hd = hmm_new_data(L, sq, hp);
	hmm_forward(hp, hd);
	hmm_backward(hp, hd);
	for (u = 1; u <= L; ++u) {
		for (k = 0; k < N; ++k) {
			f[(size_t)(u - 1) * N + k] = hd->f[u][k];
			b[(size_t)(u - 1) * N + k] = hd->b[u][k];
		}
		s[u - 1] = hd->s[u];
	}
	hmm_delete_data(hd); free(sq);
	hmm_delete_par(hp);
	return 0;
}

/* ---- decode of ONE sequence: hmm_post_decode path (aux.c:157-200) --------------------------------
 * best_k[L], best_p[L] (posterior of best state), post[L*N] (may be NULL), p_recomb[L] (may be NULL; aux.c:188-193) */
int ref_decode(int N, const double *a, const double *e, const double *a0, int L, const char *seq,
               int *best_k, double *best_p, double *post, double *p_recomb)
{
	hmm_par_t *hp = mk_hp(N, a, e, a0);
	hmm_data_t *hd;
	char *sq = (char*)calloc(L + 1, 1);
	double *prob = (double*)malloc(sizeof(double) * N);
	int u, l;
	memcpy(sq, seq, L);
	hmm_pre_backward(hp);
	hd = hmm_new_data(L, sq, hp);
	hmm_forward(hp, hd);
	hmm_backward(hp, hd);
	hmm_post_decode(hp, hd);
	for (u = 1; u <= L; ++u) {
		int x = hd->p[u];
		best_k[u - 1] = x;
		best_p[u - 1] = hd->f[u][x] * hd->b[u][x] * hd->s[u];
		if (post) {
			hmm_post_state(hp, hd, u, prob);
			for (l = 0; l < N; ++l) post[(size_t)(u - 1) * N + l] = prob[l];
		}
		if (p_recomb) {
			double p = 0.0;
			if (u < L) {
				FLOAT *fu = hd->f[u], *bu1 = hd->b[u + 1], *eu1 = hp->e[(int)hd->seq[u + 1]];
				for (l = 0; l < N; ++l) p += fu[l] * hp->a[l][l] * bu1[l] * eu1[l];
				p = 1.0 - p;
			}
			p_recomb[u - 1] = p;
		}
	}
	free(prob);
	hmm_delete_data(hd); free(sq);
	hmm_delete_par(hp);
	return 0;
}

/* ---- Q-function of the reference on caller-supplied dense stats ------------------------------- */
double ref_Q(int N, const double *a, const double *e, const double *a0,
             const double *A, const double *E, double *Q0_out)
{
	hmm_par_t *hp = mk_hp(N, a, e, a0);
	hmm_exp_t *he = hmm_new_exp(hp);
	double q;
	int k, l;
	for (k = 0; k < N; ++k) {
		for (l = 0; l < N; ++l) he->A[k][l] = A[k * N + l];
		he->E[0][k] = E[k]; he->E[1][k] = E[N + k];
	}
	hmm_Q0(hp, he);
	if (Q0_out) *Q0_out = he->Q0;
	q = hmm_Q(hp, he);
	hmm_delete_exp(he);
	hmm_delete_par(hp);
	return q;
}

/* ---- Hooke-Jeeves of the reference on a caller-supplied callback ------------------------------ */
double ref_kmin_hj(kmin_f func, int n, double *x, void *data, double r, double eps, int max_calls)
{
	return kmin_hj(func, n, x, data, r, eps, max_calls);
}
