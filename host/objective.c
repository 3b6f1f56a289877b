/* objective.c -- expected-count container and the EM objective on the factored model.
 *
 * hmm_Q (khmm.c:363-382) is  sum_b sum_k E[b][k] log e[b][k] + sum_kl A[k][l] log a[k][l] - Q0  with
 * -HMM_INF as soon as a used probability is <= 0.  With a[k][l] = U_k V_l / W_k Z_l / D_k the double
 * sum collapses to five O(N) sums over the marginals RL, CL, RU, CU, AD:
 *   sum_k ( RL_k log U_k + RU_k log W_k + AD_k log D_k ) + sum_l ( CL_l log V_l + CU_l log Z_l ). */
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include "psmc_host.h"

int psmch_counts_alloc(psmch_counts_t *c, int N, int dense)
{
	double *blk;
	memset(c, 0, sizeof(*c));
	c->N = N;
	blk = (double*)calloc(7 * (size_t)N + (dense ? (size_t)N * N : 0), sizeof(double));
	if (blk == 0) return -1;
	c->E = blk; blk += 2 * N;
	c->RL = blk; blk += N; c->CL = blk; blk += N; c->RU = blk; blk += N; c->CU = blk; blk += N; c->AD = blk; blk += N;
	if (dense) c->A = blk;
	return 0;
}

void psmch_counts_free(psmch_counts_t *c)
{
	free(c->E);
	memset(c, 0, sizeof(*c));
}

void psmch_counts_view(psmch_counts_t *c, psmc_b200_stats *v)
{
	v->LL = 0.0; v->E = c->E;
	v->RL = c->RL; v->CL = c->CL; v->RU = c->RU; v->CU = c->CU; v->AD = c->AD;
}

void psmch_counts_from_dense(psmch_counts_t *c)
{
	int N = c->N, k, l;
	for (k = 0; k < N; ++k) c->RL[k] = c->CL[k] = c->RU[k] = c->CU[k] = 0.0;
	for (k = 0; k < N; ++k)
		for (l = 0; l < N; ++l) {
			double v = c->A[(size_t)k * N + l];
			if (l < k) { c->RL[k] += v; c->CL[l] += v; }
			else if (l > k) { c->RU[k] += v; c->CU[l] += v; }
			else c->AD[k] = v;
		}
}

/* Offset of the objective (khmm.c:326-342).  It does not depend on the trial parameters, so it
 * shifts the printed QD values only.  The emission part is exact.  The transition part
 * sum_kl A log(A/rowsum) needs the dense counts; without them the row-wise lower bound built from the
 * marginals is used (each row treated as {lower block, diagonal, upper block}). */
double psmch_Q0(psmch_counts_t *c)
{
	int N = c->N, k, l, b;
	double sum = 0.0;
	for (k = 0; k < N; ++k) {
		double tot = 0.0;
		for (b = 0; b < 2; ++b) tot += c->E[(size_t)b * N + k];
		for (b = 0; b < 2; ++b) sum += c->E[(size_t)b * N + k] * log(c->E[(size_t)b * N + k] / tot);
	}
	if (c->A) {
		for (k = 0; k < N; ++k) {
			const double *Ak = c->A + (size_t)k * N;
			double tot = 0.0;
			for (l = 0; l < N; ++l) tot += Ak[l];
			for (l = 0; l < N; ++l) sum += Ak[l] * log(Ak[l] / tot);
		}
	} else {
		for (k = 0; k < N; ++k) {
			const double tot = c->RL[k] + c->AD[k] + c->RU[k];
			if (c->RL[k] > 0) sum += c->RL[k] * log(c->RL[k] / tot);
			if (c->AD[k] > 0) sum += c->AD[k] * log(c->AD[k] / tot);
			if (c->RU[k] > 0) sum += c->RU[k] * log(c->RU[k] / tot);
		}
	}
	return (c->Q0 = sum);
}

double psmch_Q(const psmch_model_t *m, const psmch_counts_t *c)
{
	const int N = m->N;
	double sum = 0.0;
	int k;
	for (k = 0; k < 2 * N; ++k) {                      /* khmm.c:366-372 */
		if (m->e[k] <= 0.0) return -PSMCH_INF;
		sum += c->E[k] * log(m->e[k]);
	}
	/* positivity of every a[k][l] (khmm.c:376): all U_k V_l (l<k), W_k Z_l (l>k), D_k must be > 0 */
	for (k = 0; k < N; ++k) {
		if (m->D[k] <= 0.0) return -PSMCH_INF;
		if (k > 0 && (m->U[k] * m->V[0] <= 0.0 || m->U[N - 1] * m->V[k - 1] <= 0.0)) return -PSMCH_INF;
		if (k < N - 1 && (m->W[k] * m->Z[N - 1] <= 0.0 || m->W[0] * m->Z[k + 1] <= 0.0)) return -PSMCH_INF;
	}
	for (k = 0; k < N; ++k) {
		double s = c->AD[k] * log(m->D[k]);
		if (k > 0) s += c->RL[k] * log(fabs(m->U[k])) + c->CU[k] * log(fabs(m->Z[k]));
		if (k < N - 1) s += c->RU[k] * log(fabs(m->W[k])) + c->CL[k] * log(fabs(m->V[k]));
		sum += s;
	}
	return sum - c->Q0;
}

/* psmch_Q with all 7N logarithms in one vectorised sweep (same guards, same value to rounding). */
void psmch_vlog(int n, const double *x, double *y);
double psmch_vdot(int n, const double *a, const double *b);

double psmch_Q_fast(const psmch_model_t *m, const psmch_counts_t *c)
{
	const int N = m->N;
	double *x = m->vw + 4 * (N + 2), *w = x + 7 * N, *lg = w + 7 * N; /* vw: 4*(N+2) used by the model, then 3 x 7N here */
	int k, n = 0;
	for (k = 0; k < 2 * N; ++k)
		if (m->e[k] <= 0.0) return -PSMCH_INF;
	for (k = 0; k < N; ++k) {
		if (m->D[k] <= 0.0) return -PSMCH_INF;
		if (k > 0 && (m->U[k] * m->V[0] <= 0.0 || m->U[N - 1] * m->V[k - 1] <= 0.0)) return -PSMCH_INF;
		if (k < N - 1 && (m->W[k] * m->Z[N - 1] <= 0.0 || m->W[0] * m->Z[k + 1] <= 0.0)) return -PSMCH_INF;
	}
	double q_e0 = 0.0;
	if (m->fast_valid) { /* log e[0][k] = -theta (avg_t_k + dt) is what the model update exponentiated: no logarithm needed */
		q_e0 = psmch_vdot(N, c->E, m->vw + 3 * (N + 2));
		for (k = N; k < 2 * N; ++k) { x[n] = m->e[k]; w[n] = c->E[k]; ++n; }
	} else {
		for (k = 0; k < 2 * N; ++k) { x[n] = m->e[k]; w[n] = c->E[k]; ++n; }
	}
	for (k = 0; k < N; ++k) { x[n] = m->D[k]; w[n] = c->AD[k]; ++n; }
	for (k = 1; k < N; ++k) {
		x[n] = fabs(m->U[k]); w[n] = c->RL[k]; ++n;
		x[n] = fabs(m->Z[k]); w[n] = c->CU[k]; ++n;
	}
	for (k = 0; k < N - 1; ++k) {
		x[n] = fabs(m->W[k]); w[n] = c->RU[k]; ++n;
		x[n] = fabs(m->V[k]); w[n] = c->CL[k]; ++n;
	}
	psmch_vlog(n, x, lg);
	return q_e0 + psmch_vdot(n, w, lg) - c->Q0;
}
