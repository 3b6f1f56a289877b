/* resamp.c -- bootstrap resampling of whole records with replacement until the total length is as
 * close as possible to the original (behaviour of aux.c:8-47, psmc_resamp), driven by a caller-supplied
 * uniform generator so that runs are reproducible under --seed (the reference seeds drand48 with
 * time^pid, main.c:11). */
#include <stdlib.h>
#include <string.h>
#include "psmc_host.h"

void psmch_resample(psmch_seqs_t *sq, double (*rnd)(void))
{
	int64_t L_ori = 0, L = 0;
	int i, n_new = 0, cap = 0;
	psmch_seq_t *ns = 0;
	for (i = 0; i < sq->n_seqs; ++i) L_ori += sq->seqs[i].L;
	for (;;) {
		const psmch_seq_t *s = sq->seqs + (int)(sq->n_seqs * rnd());
		/* the reference does this arithmetic in int (aux.c:16-17) */
		const int short_by = (int)(L_ori - L), over_by = (int)(L + s->L - L_ori);
		if (over_by <= 0 || (over_by > 0 && short_by > 0 && over_by < short_by)) {
			psmch_seq_t *d;
			if (n_new == cap) { cap = cap ? cap * 2 : 256; ns = (psmch_seq_t*)realloc(ns, sizeof(psmch_seq_t) * cap); }
			d = ns + n_new++;
			*d = *s;
			d->name = strdup(s->name);
			d->seq = (signed char*)malloc(s->L > 0 ? s->L : 1);
			memcpy(d->seq, s->seq, s->L);
			L += s->L;
		}
		if (short_by >= 0 && over_by >= 0) break;
	}
	for (i = 0; i < sq->n_seqs; ++i) { free(sq->seqs[i].name); free(sq->seqs[i].seq); }
	free(sq->seqs);
	sq->seqs = ns; sq->n_seqs = n_new;
	sq->sum_n = sq->sum_L = 0;
	for (i = 0; i < n_new; ++i) { sq->sum_n += ns[i].n_e; sq->sum_L += ns[i].L_e; }
}
