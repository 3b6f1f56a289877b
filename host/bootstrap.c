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

/* ---- batched replicates: B independent EM runs share every launch (psmc_b200_set_batch) -------------------------
 * The drawn records of B replicates are the work items of ONE chunk plan, so chunks are ~10x longer than for a single
 * replicate and the warm-up overlaps / repairs cost a few percent instead of more than the stored work.  The B host
 * M-steps of an iteration run on a pool of threads.  Replicate r still draws from srand48(seed + r) and prints the
 * text it would print alone. */
typedef struct { psmch_em_t *em; FILE **fp; int j0, j1, stride; } mstep_job_t;
static void *mstep_thread(void *arg)
{
	mstep_job_t *jb = (mstep_job_t*)arg;
	int j;
	for (j = jb->j0; j < jb->j1; j += jb->stride) psmch_em_mstep(&jb->em[j], jb->fp[j]);
	return 0;
}

typedef struct {
	const psmch_opts_t *o;
	const psmch_seqs_t *sq;
	int gpu_slot, n_rep, batch, n_threads;
	long seed0;
	int *next;
	pthread_mutex_t *mu;
	char **out_buf;
	size_t *out_len;
	int rc;
	double e_ms, m_ms;
} bworker_t;

static void *batch_worker(void *arg)
{
	bworker_t *w = (bworker_t*)arg;
	const psmch_seqs_t *sq = w->sq;
	const int ns = sq->n_seqs > 0 ? sq->n_seqs : 1, B = w->batch;
	psmc_b200_ctx *ctx = 0;
	int32_t *L = (int32_t*)malloc(sizeof(int32_t) * ns), *mult = (int32_t*)malloc(sizeof(int32_t) * (size_t)ns * B);
	const signed char **ptr = (const signed char**)malloc(sizeof(void*) * ns);
	psmch_em_t *em = (psmch_em_t*)calloc(B, sizeof(psmch_em_t));
	psmch_seqs_t *view = (psmch_seqs_t*)calloc(B, sizeof(psmch_seqs_t));
	psmch_opts_t *ob = (psmch_opts_t*)calloc(B, sizeof(psmch_opts_t));
	FILE **fp = (FILE**)calloc(B, sizeof(FILE*));
	psmc_b200_model *mv = (psmc_b200_model*)calloc(B, sizeof(psmc_b200_model));
	psmc_b200_stats *sv = (psmc_b200_stats*)calloc(B, sizeof(psmc_b200_stats));
	pthread_t *th = (pthread_t*)calloc(w->n_threads, sizeof(pthread_t));
	mstep_job_t *jobs = (mstep_job_t*)calloc(w->n_threads, sizeof(mstep_job_t));
	psmch_space_t hdr;
	int i, j;
	w->rc = 0;
	if (psmch_space_init(&hdr, w->o->pattern ? w->o->pattern : "4+5*3+4", 0, w->o->alpha0) != 0) { w->rc = -1; goto done; }
	for (i = 0; i < sq->n_seqs; ++i) { L[i] = sq->seqs[i].L; ptr[i] = sq->seqs[i].seq; }
	if (psmc_b200_create(&ctx, sq->n_seqs, L, ptr, hdr.n + 1, w->o->devices[w->gpu_slot], w->o->chunk_len, 0) != 0) {
		fprintf(stderr, "psmc: GPU E-step unavailable on device %d: %s\n", w->o->devices[w->gpu_slot], psmc_b200_last_error());
		w->rc = -1;
		goto done;
	}
	for (;;) {
		int r0, b, it, nt;
		double t0 = now_ms(), t_plan;
		psmc_b200_info inf;
		pthread_mutex_lock(w->mu);
		r0 = *w->next;
		b = w->n_rep - r0 < B ? w->n_rep - r0 : B;
		if (b > 0) *w->next += b;
		pthread_mutex_unlock(w->mu);
		if (b <= 0) break;
		for (j = 0; j < b; ++j) { /* replicate r0 + j: draw, header, initial model -- all from its own srand48(seed + r) stream */
			tl_seed(w->seed0 + r0 + j);
			psmch_draw(sq, tl_rnd, mult + (size_t)j * ns, &view[j]);
			ob[j] = *w->o;
			fp[j] = ob[j].fpout = open_memstream(&w->out_buf[r0 + j], &w->out_len[r0 + j]);
			psmch_print_header(&ob[j], 0, 0, 0);
			psmch_print_header(&ob[j], &hdr, 0, 1);
			psmch_print_header(&ob[j], 0, &view[j], 2);
			if (psmch_em_init_shared(&em[j], &ob[j], &view[j], ctx, tl_rnd) != 0) { w->rc = -1; break; }
			fprintf(fp[j], "RD\t0\n");
			psmch_print_round(&ob[j], &em[j], &view[j], fp[j]);
		}
		if (w->rc != 0) break;
		t_plan = now_ms();
		if (psmc_b200_set_batch(ctx, b, mult) != 0) {
			fprintf(stderr, "psmc: replicates %d..%d: %s\n", r0, r0 + b - 1, psmc_b200_last_error());
			w->rc = -1;
			break;
		}
		t_plan = now_ms() - t_plan;
		nt = b < w->n_threads ? b : w->n_threads;
		for (it = 0; it < w->o->n_iters && w->rc == 0; ++it) {
			double t1 = now_ms(), t2;
			for (j = 0; j < b; ++j) { psmch_model_view(&em[j].model, &mv[j]); psmch_counts_view(&em[j].counts, &sv[j]); }
			if (psmc_b200_estep_batch(ctx, b, mv, sv) != 0) {
				fprintf(stderr, "psmc: E-step failed: %s\n", psmc_b200_last_error());
				w->rc = -1;
				break;
			}
			for (j = 0; j < b; ++j) em[j].counts.LL = sv[j].LL;
			t2 = now_ms();
			w->e_ms += t2 - t1;
			for (j = 0; j < nt; ++j) {
				jobs[j].em = em; jobs[j].fp = fp; jobs[j].j0 = j; jobs[j].j1 = b; jobs[j].stride = nt;
				if (j > 0) pthread_create(&th[j], 0, mstep_thread, &jobs[j]);
			}
			mstep_thread(&jobs[0]);
			for (j = 1; j < nt; ++j) pthread_join(th[j], 0);
			w->m_ms += now_ms() - t2;
			for (j = 0; j < b; ++j) {
				fprintf(fp[j], "RD\t%d\n", it + 1);
				psmch_print_round(&ob[j], &em[j], &view[j], fp[j]);
			}
		}
		if (w->o->verbose) {
			psmc_b200_get_info(ctx, &inf);
			fprintf(stderr, "[psmc-b200] replicates %d..%d on device %d: %lld bins drawn, plan %.1f ms, chunks %d x %d (backward %d x %d), fallbacks %d, %.1f ms; last E-step on the device: %.1f ms "
			        "(forward %.1f [kernel %.1f], backward %.1f [kernel %.1f]), failed boundaries %d/%d, repair rounds %d\n",
			        r0, r0 + b - 1, w->o->devices[w->gpu_slot], (long long)inf.active_bins, t_plan, inf.n_chunks, inf.chunk_len, inf.n_chunks_bwd, inf.chunk_len_bwd, inf.fallbacks, now_ms() - t0,
			        inf.ms[5], inf.ms[2], inf.ms[6], inf.ms[3], inf.ms[7], inf.failed_fwd, inf.failed_bwd, inf.repair_rounds);
		}
		for (j = 0; j < b; ++j) { psmch_em_free(&em[j]); fclose(fp[j]); fp[j] = 0; }
		if (w->rc != 0) break;
	}
done:
	for (j = 0; j < B; ++j) if (fp[j]) fclose(fp[j]);
	if (ctx) psmc_b200_destroy(ctx);
	psmch_space_free(&hdr);
	free(L); free(mult); free(ptr); free(em); free(view); free(ob); free(fp); free(mv); free(sv); free(th); free(jobs);
	return 0;
}

/* how many replicates fit side by side on one device: the forward spill takes 8 * (NP + 1) bytes per drawn bin, the
 * repair operators 2 * NP * NP * 8 bytes per sub-chunk of 1536 bins; a replicate draws ~63 % distinct records */
static int auto_batch(const psmch_opts_t *o, const psmch_seqs_t *sq, int n_states, int workers_per_gpu)
{
	int64_t fr = 0, tot = 0, bins = 0;
	const int NP = n_states <= 32 ? 32 : (n_states <= 64 ? 64 : 128);
	double per_bin, per_rep;
	int i, b;
	for (i = 0; i < sq->n_seqs; ++i) bins += sq->seqs[i].L;
	if (psmc_b200_mem_info(o->devices[0], &fr, &tot) != 0 || bins <= 0) return 1;
	per_bin = 8.0 * (NP + 1) + 2.0 * NP * NP * 8.0 / 1536.0 * 1.125 + 16.0;
	per_rep = 0.72 * 1.125 * (double)bins * per_bin; /* (the library keeps 1/8 head room on the spill) */
	b = (int)(0.80 * (double)fr / workers_per_gpu / per_rep);
	return b < 1 ? 1 : (b > 64 ? 64 : b);
}

/* R replicates over o->n_gpus devices with `slots` concurrent replicates per device; text goes to o->fpout */
int psmch_bootstrap_run(const psmch_opts_t *o, const psmch_seqs_t *sq, int n_rep, int slots)
{
	const int n_workers_ = o->n_gpus * (slots < 1 ? 1 : slots), bsl_ = o->n_gpus * (o->batch_slots > 0 ? o->batch_slots : 1);
	const int n_workers = n_workers_ > bsl_ ? n_workers_ : bsl_;
	const long seed0 = o->seed >= 0 ? o->seed : (long)(time(0) ^ getpid());
	pthread_t *th = (pthread_t*)calloc(n_workers, sizeof(pthread_t));
	worker_t *w = (worker_t*)calloc(n_workers, sizeof(worker_t));
	char **buf = (char**)calloc(n_rep > 0 ? n_rep : 1, sizeof(char*));
	size_t *len = (size_t*)calloc(n_rep > 0 ? n_rep : 1, sizeof(size_t));
	double *ms = (double*)calloc(n_rep > 0 ? n_rep : 1, sizeof(double));
	pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER;
	int next = 0, i, rc = 0;
	const double t0 = now_ms();
	if (o->batch != 1 && !o->exact_qd && !(getenv("PSMC_B200_EXACT_QD") && atoi(getenv("PSMC_B200_EXACT_QD")) != 0)) {
		/* batched: one worker per (GPU, slot), every worker runs batches of B replicates in lock step */
		const int bslots = o->batch_slots > 0 ? o->batch_slots : 1, nw = o->n_gpus * bslots;
		bworker_t *bw = (bworker_t*)calloc(nw, sizeof(bworker_t));
		psmch_space_t hdr;
		long cores = sysconf(_SC_NPROCESSORS_ONLN);
		int B = o->batch, per;
		double e_ms = 0.0, m_ms = 0.0;
		if (psmch_space_init(&hdr, o->pattern ? o->pattern : "4+5*3+4", 0, o->alpha0) != 0) { free(bw); free(th); free(w); free(buf); free(len); free(ms); return -1; }
		if (B <= 0) B = auto_batch(o, sq, hdr.n + 1, bslots);
		per = (n_rep + nw - 1) / nw; /* a worker's share: split it into equal batches (13 replicates = 7 + 6, not 11 + 2) */
		if (B < 1) B = 1;
		{
			const int nb = (per + B - 1) / B;
			B = (per + nb - 1) / nb;
		}
		psmch_space_free(&hdr);
		if (cores < 1) cores = 1;
		for (i = 0; i < nw; ++i) {
			bw[i].o = o; bw[i].sq = sq; bw[i].gpu_slot = i % o->n_gpus; bw[i].n_rep = n_rep; bw[i].batch = B; bw[i].seed0 = seed0;
			bw[i].n_threads = (int)(cores / nw > 0 ? cores / nw : 1);
			bw[i].next = &next; bw[i].mu = &mu; bw[i].out_buf = buf; bw[i].out_len = len;
			pthread_create(&th[i], 0, batch_worker, &bw[i]);
		}
		for (i = 0; i < nw; ++i) { pthread_join(th[i], 0); if (bw[i].rc != 0) rc = -1; e_ms += bw[i].e_ms; m_ms += bw[i].m_ms; }
		for (i = 0; i < n_rep; ++i)
			if (buf[i]) { if (rc == 0) fwrite(buf[i], 1, len[i], o->fpout); free(buf[i]); }
		fflush(o->fpout);
		if (o->verbose || getenv("PSMC_B200_TIMING"))
			fprintf(stderr, "[psmc-b200] bootstrap: %d replicates x %d iterations on %d GPU(s), batches of %d x %d worker(s) per GPU: %.3f s (E-steps %.3f s, M-steps %.3f s summed over workers)\n",
			        n_rep, o->n_iters, o->n_gpus, B, bslots, (now_ms() - t0) * 1e-3, e_ms * 1e-3, m_ms * 1e-3);
		free(bw); free(th); free(w); free(buf); free(len); free(ms);
		return rc;
	}
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
