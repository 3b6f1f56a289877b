/* decode.c -- text of the decoding modes (aux.c:129-232): TC lines, then per sequence either
 * PR (-s, per-bin scale factors), DC (-d, runs of the posterior-argmax state with their maximum
 * posterior) or DF (-D, recombination probability and the full posterior row per bin).
 * Posteriors come from psmc_b200_decode (forward/backward + argmax on the GPU). */
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include "psmc_host.h"

int psmch_decode(const psmch_opts_t *o, psmch_em_t *em, const psmch_seqs_t *sq, FILE *fp)
{
	const psmch_space_t *sp = &em->sp;
	const int N = sp->n + 1;
	const double theta = em->model.params[0];
	double *avg = (double*)malloc(sizeof(double) * N);
	int local[16] = {0}, primed[16] = {0};
	int i, k, l, rc = 0;
	psmc_b200_model mv;
	psmch_model_view(&em->model, &mv);
	psmch_avg_t(sp, &em->model, avg);
	for (k = 0; k < N; ++k) {
		if (avg[k] < em->model.t[k] || avg[k] > em->model.t[k + 1])
			fprintf(stderr, "ERROR: (%f <= %f <= %f) does not stand. Contact me if you see this.\n", em->model.t[k], avg[k], em->model.t[k + 1]);
		fprintf(fp, "TC\t%d\t%lf\t%lf\t%lf\n", k, em->model.t[k] * theta, avg[k] * theta, em->model.t[k + 1] * theta);
	}
	for (i = 0; i < sq->n_seqs && rc == 0; ++i) {
		const psmch_seq_t *s = sq->seqs + i;
		const int g = em->seq_owner[i], L = s->L, full = (o->flag & PSMCH_F_FULLDEC) && !(o->flag & PSMCH_F_PROB);
		int32_t *bk;
		double *bp, *post = 0, *prec = 0, *sc = 0;
		if (L == 0) { ++local[g]; continue; } /* nothing to decode; psmc_b200_decode counts the records as given to create */
		bk = (int32_t*)malloc(sizeof(int32_t) * L);
		bp = (double*)malloc(sizeof(double) * L);
		if (full) { post = (double*)malloc(sizeof(double) * (size_t)L * N); prec = (double*)malloc(sizeof(double) * L); }
		if (o->flag & PSMCH_F_PROB) sc = (double*)malloc(sizeof(double) * L);
		rc = psmc_b200_decode(em->ctx[g], primed[g] ? 0 : &mv, local[g], bk, bp, post, prec, sc);
		primed[g] = 1; ++local[g];
		if (rc != 0) {
			fprintf(stderr, "psmc: GPU decode failed: %s\n", psmc_b200_last_error());
		} else if (o->flag & PSMCH_F_PROB) { /* aux.c:159-164 */
			fprintf(fp, "PR\t%s\t%d", s->name, L);
			for (k = 0; k < L; ++k) fprintf(fp, "\t%.3f", sc[k]);
			fprintf(fp, "\n");
			fflush(fp);
		} else if (!full) { /* aux.c:165-182, 1-based positions */
			int start = 1, prev = bk[0];
			double p = bp[0];
			for (k = 2; k <= L; ++k) {
				if (prev != bk[k - 1]) {
					fprintf(fp, "DC\t%s\t%d\t%d\t%d\t%lf\t%.3lf\n", s->name, start, k - 1, prev, avg[prev] * theta, p);
					prev = bk[k - 1]; start = k; p = 0.0;
				}
				if (p < bp[k - 1]) p = bp[k - 1];
			}
			fprintf(fp, "DC\t%s\t%d\t%d\t%d\t%.3lf\t%.2lf\n", s->name, start, k - 1, prev, avg[prev] * theta, p);
			fflush(fp);
		} else { /* aux.c:183-200 */
			for (k = 1; k <= L; ++k) {
				const double *row = post + (size_t)(k - 1) * N;
				fprintf(fp, "DF\t%d\t%lf", k, prec[k - 1]);
				for (l = 0; l < N; ++l) fprintf(fp, "\t%.4f", row[l]);
				fprintf(fp, "\n");
			}
		}
		free(bk); free(bp); free(post); free(prec); free(sc);
	}
	free(avg);
	return rc;
}
