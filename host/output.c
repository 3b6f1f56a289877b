/* output.c -- per-round .psmc text and the -i parameter reader.
 * psmch_print_round reproduces aux.c:49-82 field by field (LK QD RI TR MT [DT] MM RS*(n+1) PA //);
 * psmch_read_param reproduces aux.c:84-113 (pattern, n_free+3 parameters, optional time intervals when
 * max_t < 0, optional divergence time). */
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include "psmc_host.h"

void psmch_print_round(const psmch_opts_t *o, const psmch_em_t *em, const psmch_seqs_t *sq, FILE *fp)
{
	const psmch_space_t *sp = &em->sp;
	const psmch_model_t *m = &em->model;
	const int N = sp->n + 1;
	const double n_recomb = sq->sum_L / m->C_sigma;
	double ri = 0.0;
	int k;
	fprintf(fp, "LK\t%lf\n", em->lk);
	fprintf(fp, "QD\t%lf -> %lf\n", em->Q0, em->Q1);
	for (k = 0; k < N; ++k) ri += m->sigma[k] * log(m->sigma[k] / em->post_sigma[k]);
	fprintf(fp, "RI\t%.10lf\n", ri);
	fprintf(fp, "TR\t%lf\t%lf\n", m->params[0], m->params[1]);
	fprintf(fp, "MT\t%lf\n", m->params[2]);
	if (sp->diverg) fprintf(fp, "DT\t%lf\n", m->params[sp->n_params - 1]);
	fprintf(fp, "MM\tC_pi: %lf, n_recomb: %lf\n", m->C_pi, n_recomb);
	for (k = 0; k < N; ++k)
		fprintf(fp, "RS\t%d\t%lf\t%lf\t%lf\t%lf\t%lf\n", k, m->t[k], m->params[sp->par_map[k] + PSMCH_N_PARAMS],
		        n_recomb * m->sigma[k], m->sigma[k], em->post_sigma[k]);
	fprintf(fp, "PA\t%s", sp->pattern);
	for (k = 0; k < sp->n_params; ++k) fprintf(fp, " %.9lf", m->params[k]);
	if (sp->inp_ti)
		for (k = 0; k < N; ++k) fprintf(fp, " %.9lf", sp->inp_ti[k]);
	fprintf(fp, "\n//\n");
	fflush(fp);
	(void)o;
}

int psmch_read_param(psmch_opts_t *o, psmch_space_t *sp)
{
	FILE *fp;
	char str[256];
	int k, n, n_free, *pm = 0;
	double v;
	if (o->pre_fn == 0) return -1;
	fp = fopen(o->pre_fn, "r");
	if (fp == 0) { fprintf(stderr, "psmc: cannot open parameter file '%s'\n", o->pre_fn); return -1; }
	if (fscanf(fp, "%255s", str) != 1) { fclose(fp); return -1; }
	n = psmch_parse_pattern(str, &n_free, &pm);
	if (n < 0) { fclose(fp); fprintf(stderr, "psmc: bad pattern '%s' in '%s'\n", str, o->pre_fn); return -1; }
	free(pm);
	free(o->pattern);
	o->pattern = strdup(str);
	o->inp_pa = (double*)calloc(n_free + PSMCH_N_PARAMS + 1, sizeof(double));
	for (k = 0; k < n_free + PSMCH_N_PARAMS; ++k)
		if (fscanf(fp, "%lf", &o->inp_pa[k]) != 1) o->inp_pa[k] = 0.0;
	if (o->inp_pa[2] < 0) { /* then the time intervals follow (aux.c:101-106) */
		o->inp_ti = (double*)calloc(n + 1, sizeof(double));
		for (k = 0; k <= n; ++k)
			if (fscanf(fp, "%lf", &o->inp_ti[k]) != 1) o->inp_ti[k] = 0.0;
	}
	if (fscanf(fp, "%lf", &v) > 0) { /* divergence time (aux.c:107-108) */
		o->inp_pa[n_free + PSMCH_N_PARAMS] = v;
		o->dt0 = v;
		o->flag |= PSMCH_F_DIVERG;
	}
	o->max_t = o->inp_pa[2];
	o->tr_ratio = o->inp_pa[0] / o->inp_pa[1];
	fclose(fp);
	(void)sp;
	return 0;
}
