/* hj.c -- Hooke-Jeeves direct search, same search path as kmin.c:48-107 (the EM driver relies on the
 * exact sequence of trial points: the reference keeps the LAST EVALUATED point, not the optimum,
 * em.c:61-67, and that quirk is reproduced by the caller). */
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include "psmc_host.h"

typedef struct {
	psmch_func_t f;
	void *data;
	int n, calls;
	psmch_spec_t *spec; /* helper thread for the -step point of a probe, or NULL */
	double *last;       /* the last evaluated point in the order of the sequential search, or NULL */
} hj_t;

static double eval(hj_t *h, double *x)
{
	++h->calls;
	if (h->last) memcpy(h->last, x, sizeof(double) * h->n);
	return h->f(h->n, x, h->data);
}

/* probe each coordinate: +step, else -step, else stay (kmin.c:48-66).  With a helper, the -step point is evaluated
 * concurrently and counted only if the sequential search would have evaluated it. */
static double probe(hj_t *h, double *x, double fbest, double *step)
{
	int k;
	for (k = 0; k < h->n; ++k) {
		double v;
		if (h->spec) {
			const double xk = x[k];
			x[k] = (xk + step[k]) + ((0.0 - step[k]) + (0.0 - step[k])); /* the -step point, formed exactly as below */
			psmch_spec_submit(h->spec, x);
			x[k] = xk;
		}
		x[k] += step[k];
		v = eval(h, x);
		if (v < fbest) {
			fbest = v;
			if (h->spec) psmch_spec_wait(h->spec); /* discard */
			continue;
		}
		step[k] = 0.0 - step[k];
		x[k] += step[k] + step[k];
		if (h->spec) {
			++h->calls;
			if (h->last) memcpy(h->last, x, sizeof(double) * h->n);
			v = psmch_spec_wait(h->spec);
		} else {
			v = eval(h, x);
		}
		if (v < fbest) fbest = v;
		else x[k] -= step[k];
	}
	return fbest;
}

double psmch_hooke_jeeves(psmch_func_t f, int n, double *x, void *data, double r, double eps, int max_calls)
{
	return psmch_hooke_jeeves_spec(f, 0, n, x, data, r, eps, max_calls, 0, 0);
}

double psmch_hooke_jeeves_spec(psmch_func_t f, psmch_spec_t *spec, int n, double *x, void *data, double r, double eps, int max_calls,
                               double *last, int *n_calls)
{
	hj_t h;
	double *y = (double*)calloc(n, sizeof(double)), *step = (double*)calloc(n, sizeof(double));
	double fx, fy, radius = r;
	int k, done = 0;
	h.f = f; h.data = data; h.n = n; h.calls = 0; h.spec = spec; h.last = last;
	for (k = 0; k < n; ++k) {
		step[k] = fabs(x[k]) * r;
		if (step[k] == 0) step[k] = r;
	}
	fy = fx = eval(&h, x);
	while (!done) {
		memcpy(y, x, sizeof(double) * n);
		fy = probe(&h, y, fx, step);
		while (fy < fx) { /* pattern move (kmin.c:83-98) */
			for (k = 0; k < n; ++k) {
				const double prev = x[k];
				step[k] = y[k] > x[k] ? fabs(step[k]) : 0.0 - fabs(step[k]);
				x[k] = y[k];
				y[k] = y[k] + y[k] - prev;
			}
			fx = fy;
			if (h.calls >= max_calls) break;
			fy = eval(&h, y);
			fy = probe(&h, y, fy, step);
			if (fy >= fx) break;
			for (k = 0; k < n; ++k)
				if (fabs(y[k] - x[k]) > .5 * fabs(step[k])) break;
			if (k == n) break;
		}
		if (radius >= eps && h.calls < max_calls) {
			radius *= r;
			for (k = 0; k < n; ++k) step[k] *= r;
		} else done = 1;
	}
	free(y); free(step);
	if (n_calls) *n_calls = h.calls;
	return fy;
}
