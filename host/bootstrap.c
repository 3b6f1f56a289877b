/* bootstrap.c -- the bootstrap pipeline in one process.
 *
 * The reference recipe (README:57-62) is: `splitfa in.psmcfa > split.psmcfa`, then 100 separate
 * `psmc -b split.psmcfa` processes, each of which parses the text again, draws records with
 * drand48 seeded by time^pid (main.c:11, aux.c:8-47) and copies every drawn record.  Here:
 *   - the splitfa rule (utils/splitfa.c:20-31) is applied in memory (--split),
 *   - the segments are uploaded ONCE per GPU and stay resident,
 *   - a replicate is a multiplicity vector over the resident segments (psmc_resamp draws whole records),
 *     handed to psmc_b200_set_multiplicity: drawn-twice segments are computed once and weighted, undrawn ones
 *     are skipped,
 *   - replicates are independent EM runs: they are dealt to (GPU, slot) worker threads, no collective;
 *     two slots per GPU keep the GPU busy while the other slot runs its host M-step,
 *   - replicate r draws from srand48(seed + r), so `--replicates R --seed S` reproduces the R runs
 *     `psmc -b --seed S`, `... --seed S+1`, ...; the outputs are concatenated in replicate order
 *     (what `cat round-*.psmc` gives, README:61). */
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <time.h>
#include <unistd.h>
#include <pthread.h>
#include "psmc_host.h"

/* ---- splitfa rule: pieces of `trunk` bins; a tail shorter than 1.5 trunks stays whole (splitfa.c:23-26) ---- */
int psmch_split(psmch_seqs_t *sq, int trunk)
{
	psmch_seq_t *ns = 0;
	int n_new = 0, cap = 0, r;
	if (trunk <= 0) return -1;
	for (r = 0; r < sq->n_seqs; ++r) {
		const psmch_seq_t *s = sq->seqs + r;
		int64_t i;
		int k = 0;
		for (i = 0; i < s->L; i += trunk) {
			const int whole_tail = (s->L - i < (int64_t)trunk * 3 / 2);
			const int64_t end = whole_tail ? s->L : (i + trunk < s->L ? i + trunk : s->L);
			psmch_seq_t *d;
			int64_t u;
			if (n_new == cap) { cap = cap ? cap * 2 : 256; ns = (psmch_seq_t*)realloc(ns, sizeof(psmch_seq_t) * cap); }
			d = ns + n_new++;
			d->L = (int32_t)(end - i);
			d->seq = (signed char*)malloc(d->L > 0 ? d->L : 1);
			memcpy(d->seq, s->seq + i, d->L);
			d->name = (char*)malloc(strlen(s->name) + 16);
			sprintf(d->name, "%s_%d", s->name, ++k);
			d->L_e = d->n_e = 0;
			for (u = 0; u < d->L; ++u) {
				if (d->seq[u] < 2) ++d->L_e;
				if (d->seq[u] == 1) ++d->n_e;
			}
			if (whole_tail) break;
		}
	}
	for (r = 0; r < sq->n_seqs; ++r) { free(sq->seqs[r].name); free(sq->seqs[r].seq); }
	free(sq->seqs);
	sq->seqs = ns; sq->n_seqs = n_new; /* sum_L and sum_n are unchanged by construction */
	return 0;
}

/* ---- one replicate as multiplicities: the draw loop of aux.c:14-32 without the copies ---- */
void psmch_draw(const psmch_seqs_t *sq, double (*rnd)(void), int32_t *mult, psmch_seqs_t *view)
{
	int64_t L_ori = 0, L = 0;
	int i;
	memset(view, 0, sizeof(*view));
	for (i = 0; i < sq->n_seqs; ++i) { L_ori += sq->seqs[i].L; mult[i] = 0; }
	for (;;) {
		const int j = (int)(sq->n_seqs * rnd());
		const psmch_seq_t *s = sq->seqs + j;
		const int short_by = (int)(L_ori - L), over_by = (int)(L + s->L - L_ori); /* int arithmetic as aux.c:16-17 */
		if (over_by <= 0 || (over_by > 0 && short_by > 0 && over_by < short_by)) {
			++mult[j];
			++view->n_seqs;
			view->sum_L += s->L_e;
			view->sum_n += s->n_e;
			L += s->L;
		}
		if (short_by >= 0 && over_by >= 0) break;
	}
}

/* ---- per-thread drand48 stream (erand48 on a thread-local state == drand48 after srand48) ---- */
static __thread unsigned short tl_x[3];
static void tl_seed(long seed)
{
	tl_x[0] = 0x330E; tl_x[1] = (unsigned short)(seed & 0xffff); tl_x[2] = (unsigned short)((seed >> 16) & 0xffff);
}
static double tl_rnd(void) { return erand48(tl_x); }

typedef struct {
	const psmch_opts_t *o;
	const psmch_seqs_t *sq;
	int gpu_slot;          /* index into devices[] */
	int n_rep;
	long seed0;
	int *next;             /* shared replicate counter */
	pthread_mutex_t *mu;
	char **out_buf;        /* per replicate */
	size_t *out_len;
	double *rep_ms;        /* per replicate wall time */
	int rc;
} worker_t;

static double now_ms(void)
{
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

static void *worker(void *arg)
{
	worker_t *w = (worker_t*)arg;
	const psmch_seqs_t *sq = w->sq;
	psmc_b200_ctx *ctx = 0;
	int32_t *L = (int32_t*)malloc(sizeof(int32_t) * (sq->n_seqs > 0 ? sq->n_seqs : 1));
	int32_t *mult = (int32_t*)malloc(sizeof(int32_t) * (sq->n_seqs > 0 ? sq->n_seqs : 1));
	const signed char **ptr = (const signed char**)malloc(sizeof(void*) * (sq->n_seqs > 0 ? sq->n_seqs : 1));
	psmch_space_t hdr;
	int i, n_states;
	w->rc = 0;
	if (psmch_space_init(&hdr, w->o->pattern ? w->o->pattern : "4+5*3+4", 0, w->o->alpha0) != 0) { w->rc = -1; goto done; }
	n_states = hdr.n + 1;
	for (i = 0; i < sq->n_seqs; ++i) { L[i] = sq->seqs[i].L; ptr[i] = sq->seqs[i].seq; }
	const double t_create = now_ms();
	if (psmc_b200_create(&ctx, sq->n_seqs, L, ptr, n_states, w->o->devices[w->gpu_slot], w->o->chunk_len, 0) != 0) {
		fprintf(stderr, "psmc: GPU E-step unavailable on device %d: %s\n", w->o->devices[w->gpu_slot], psmc_b200_last_error());
		w->rc = -1;
		goto done;
	}
	if (w->o->verbose) fprintf(stderr, "[psmc-b200] device %d: context with %d records created in %.1f ms\n", w->o->devices[w->gpu_slot], sq->n_seqs, now_ms() - t_create);
	for (;;) {
		psmch_opts_t o = *w->o;
		double e_ms = 0.0, m_ms = 0.0, t_plan;
		psmc_b200_info inf;
		psmch_seqs_t view;
		psmch_em_t em;
		int r, it;
		double t0;
		pthread_mutex_lock(w->mu);
		r = (*w->next)++;
		pthread_mutex_unlock(w->mu);
		if (r >= w->n_rep) break;
		t0 = now_ms();
		tl_seed(w->seed0 + r);
		psmch_draw(sq, tl_rnd, mult, &view);
		t_plan = now_ms();
		if (psmc_b200_set_multiplicity(ctx, mult) != 0) {
			fprintf(stderr, "psmc: replicate %d: %s\n", r, psmc_b200_last_error());
			w->rc = -1;
			break;
		}
		t_plan = now_ms() - t_plan;
		o.fpout = open_memstream(&w->out_buf[r], &w->out_len[r]);
		psmch_print_header(&o, 0, 0, 0);
		psmch_print_header(&o, &hdr, 0, 1);
		psmch_print_header(&o, 0, &view, 2);
		if (psmch_em_init_shared(&em, &o, &view, ctx, tl_rnd) != 0) { fclose(o.fpout); w->rc = -1; break; }
		fprintf(o.fpout, "RD\t0\n");
		psmch_print_round(&o, &em, &view, o.fpout);
		for (it = 0; it < o.n_iters; ++it) {
			if (psmch_em_iterate(&em, o.fpout) != 0) { w->rc = -1; break; }
			e_ms += em.t_estep_ms; m_ms += em.t_mstep_ms;
			fprintf(o.fpout, "RD\t%d\n", it + 1);
			psmch_print_round(&o, &em, &view, o.fpout);
		}
		psmch_em_free(&em);
		fclose(o.fpout);
		w->rep_ms[r] = now_ms() - t0;
		if (w->rc != 0) break;
		if (w->o->verbose) {
			psmc_b200_get_info(ctx, &inf);
			fprintf(stderr, "[psmc-b200] replicate %d on device %d: %d records drawn (%lld of %lld bins active), plan %.1f ms, E-steps %.1f ms, M-steps %.1f ms, total %.1f ms; chunks %d x %d, fallbacks %d\n",
			        r, w->o->devices[w->gpu_slot], view.n_seqs, (long long)inf.active_bins, (long long)inf.total_bins, t_plan, e_ms, m_ms, w->rep_ms[r], inf.n_chunks, inf.chunk_len, inf.fallbacks);
		}
	}
done:
	if (ctx) psmc_b200_destroy(ctx);
	psmch_space_free(&hdr);
	free(L); free(mult); free(ptr);
	return 0;
}

/* R replicates over o->n_gpus devices with `slots` concurrent replicates per device; text goes to o->fpout */
int psmch_bootstrap_run(const psmch_opts_t *o, const psmch_seqs_t *sq, int n_rep, int slots)
{
	const int n_workers = o->n_gpus * (slots < 1 ? 1 : slots);
	const long seed0 = o->seed >= 0 ? o->seed : (long)(time(0) ^ getpid());
	pthread_t *th = (pthread_t*)calloc(n_workers, sizeof(pthread_t));
	worker_t *w = (worker_t*)calloc(n_workers, sizeof(worker_t));
	char **buf = (char**)calloc(n_rep > 0 ? n_rep : 1, sizeof(char*));
	size_t *len = (size_t*)calloc(n_rep > 0 ? n_rep : 1, sizeof(size_t));
	double *ms = (double*)calloc(n_rep > 0 ? n_rep : 1, sizeof(double));
	pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER;
	int next = 0, i, rc = 0;
	const double t0 = now_ms();
	for (i = 0; i < n_workers; ++i) {
		w[i].o = o; w[i].sq = sq; w[i].gpu_slot = i % o->n_gpus; w[i].n_rep = n_rep; w[i].seed0 = seed0;
		w[i].next = &next; w[i].mu = &mu; w[i].out_buf = buf; w[i].out_len = len; w[i].rep_ms = ms;
		pthread_create(&th[i], 0, worker, &w[i]);
	}
	for (i = 0; i < n_workers; ++i) { pthread_join(th[i], 0); if (w[i].rc != 0) rc = -1; }
	for (i = 0; i < n_rep; ++i) {
		if (buf[i]) { if (rc == 0) fwrite(buf[i], 1, len[i], o->fpout); free(buf[i]); }
	}
	fflush(o->fpout);
	if (o->verbose || getenv("PSMC_B200_TIMING"))
		fprintf(stderr, "[psmc-b200] bootstrap: %d replicates x %d iterations on %d GPU(s) x %d slot(s): %.3f s\n", n_rep, o->n_iters, o->n_gpus, slots, (now_ms() - t0) * 1e-3);
	free(th); free(w); free(buf); free(len); free(ms);
	return rc;
}
