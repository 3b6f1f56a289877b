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
	psmch_spec_t **spec; /* helper threads (spec.c) or NULL */
	int n_spec;          /* 0, 1, 3, 5 or 7 helpers are used (look-ahead over 1..4 coordinates) */
	double *last;        /* the last evaluated point in the order of the sequential search, or NULL */
} hj_t;

static double eval(hj_t *h, double *x)
{
	++h->calls;
	if (h->last) memcpy(h->last, x, sizeof(double) * h->n);
	return h->f(h->n, x, h->data);
}

/* the sequential probe of one coordinate given the two objective values (kmin.c:48-66): +step, else -step, else stay.
 * Returns 1 if the point moved.  Counts the calls and records the last evaluated point exactly as the sequential search
 * would have (vb is "evaluated" only if +step failed). */
static int settle(hj_t *h, double *x, double *fbest, double *step, int k, double va, double vb)
{
	x[k] += step[k];
	++h->calls;
	if (h->last) memcpy(h->last, x, sizeof(double) * h->n);
	if (va < *fbest) { *fbest = va; return 1; }
	step[k] = 0.0 - step[k];
	x[k] += step[k] + step[k];
	++h->calls;
	if (h->last) memcpy(h->last, x, sizeof(double) * h->n);
	if (vb < *fbest) { *fbest = vb; return 1; }
	x[k] -= step[k];
	return 0;
}

/* x with coordinate k at its -step trial value, formed exactly as the sequential code forms it */
static void submit_minus(psmch_spec_t *sp, double *x, const double *step, int k)
{
	const double xk = x[k];
	x[k] = (xk + step[k]) + ((0.0 - step[k]) + (0.0 - step[k]));
	psmch_spec_submit(sp, x);
	x[k] = xk;
}
static void submit_plus(psmch_spec_t *sp, double *x, const double *step, int k)
{
	const double xk = x[k];
	x[k] = xk + step[k];
	psmch_spec_submit(sp, x);
	x[k] = xk;
}

/* probe each coordinate: +step, else -step, else stay (kmin.c:48-66).  Both trial points of a coordinate are known in
 * advance, and so are those of the NEXT coordinate if this one fails in both directions (the usual case near convergence):
 * with one helper the -step point is evaluated concurrently, with three helpers also the two points of the next
 * coordinate.  Speculative values are used only where the sequential search would have evaluated the same point. */
static double probe(hj_t *h, double *x, double fbest, double *step)
{
	int k = 0;
	while (k < h->n) {
		if (h->n_spec == 0) {
			double va, vb;
			x[k] += step[k];
			va = eval(h, x);
			if (va < fbest) { fbest = va; ++k; continue; }
			step[k] = 0.0 - step[k];
			x[k] += step[k] + step[k];
			vb = eval(h, x);
			if (vb < fbest) fbest = vb;
			else x[k] -= step[k];
			++k;
			continue;
		}
		{	/* look ahead over `depth` coordinates: helper 0 takes -step of coordinate k, helpers 2j-1 / 2j the two points of
			 * coordinate k+j under the hypothesis that the coordinates before it fail twice (they then come back through the
			 * sequential code's arithmetic, which is not always the identity, so the same arithmetic is applied here) */
			enum { MAXD = 4 };
			int depth = (h->n_spec + 1) / 2, j, moved = 0;
			double va[MAXD], vb[MAXD], keep[MAXD] = {0.0, 0.0, 0.0, 0.0};
			if (depth > MAXD) depth = MAXD;
			if (depth > h->n - k) depth = h->n - k;
			for (j = 0; j < depth; ++j) {
				keep[j] = x[k + j];
				if (j > 0) submit_plus(h->spec[2 * j - 1], x, step, k + j);
				submit_minus(h->spec[2 * j], x, step, k + j);
				if (j + 1 < depth) { /* x[k+j] as it will be after failing twice */
					const double ns = 0.0 - step[k + j];
					x[k + j] = ((keep[j] + step[k + j]) + (ns + ns)) - ns;
				}
			}
			for (j = 0; j + 1 < depth; ++j) x[k + j] = keep[j];
			{ /* the caller's share: +step of coordinate k (not counted here: settle() does the book-keeping) */
				x[k] = keep[0] + step[k];
				va[0] = h->f(h->n, x, h->data);
				x[k] = keep[0];
			}
			for (j = 0; j < depth; ++j) {
				if (j > 0) va[j] = psmch_spec_wait(h->spec[2 * j - 1]);
				vb[j] = psmch_spec_wait(h->spec[2 * j]);
			}
			for (j = 0; j < depth && !moved; ++j) moved = settle(h, x, &fbest, step, k + j, va[j], vb[j]);
			k += j; /* the coordinates settled; the points evaluated for later ones are stale once the point has moved */
		}
	}
	return fbest;
}

double psmch_hooke_jeeves(psmch_func_t f, int n, double *x, void *data, double r, double eps, int max_calls)
{
	return psmch_hooke_jeeves_spec(f, 0, 0, n, x, data, r, eps, max_calls, 0, 0);
}

double psmch_hooke_jeeves_spec(psmch_func_t f, psmch_spec_t **spec, int n_spec, int n, double *x, void *data, double r, double eps,
                               int max_calls, double *last, int *n_calls)
{
	hj_t h;
	double *y = (double*)calloc(n, sizeof(double)), *step = (double*)calloc(n, sizeof(double));
	double fx, fy, radius = r;
	int k, done = 0;
	h.f = f; h.data = data; h.n = n; h.calls = 0; h.spec = spec; h.last = last;
	h.n_spec = (spec == 0 || n_spec <= 0) ? 0 : (n_spec >= 7 ? 7 : (n_spec >= 5 ? 5 : (n_spec >= 3 ? 3 : 1)));
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
