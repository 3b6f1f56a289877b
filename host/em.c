/* em.c -- one EM iteration around the GPU E-step.  Control flow of psmc_em (em.c:27-78):
 *   E-step over every sequence  (em.c:33-55)   -> psmc_b200_estep*  (one context per GPU, contigs sharded)
 *   Q0 offset, LK, Q before      (em.c:59-64)
 *   Hooke-Jeeves on -Q           (em.c:15-25,65) -> O(N) model update + O(N) objective per trial point
 *   IT line, posterior sigma     (em.c:66-74)
 * Reproduced quirks: the objective sees |x| (em.c:22); after the search the model is left at the LAST
 * EVALUATED trial point, not at the optimum (em.c:61-67 frees the optimum without copying it back). */
#define _GNU_SOURCE
#include <sched.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <time.h>
#include <unistd.h>
#include "psmc_host.h"

static double now_ms(void)
{
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

typedef struct { psmch_em_t *em; psmch_model_t *m; } aux_t;

/* a pure function of the point (the model instance is scratch): the caller and the speculative helper use one each */
static double objective(int n, double *x, void *data)
{
	aux_t *a = (aux_t*)data;
	psmch_em_t *em = a->em;
	psmch_model_t *m = a->m;
	int i;
	for (i = 0; i < n; ++i) m->params[i] = fabs(x[i]);
	if (em->exact_mstep) {
		psmch_model_update(&em->sp, m->params, m);
		return -psmch_Q(m, &em->counts);
	}
	psmch_model_update_fast(&em->sp, m->params, m); /* exp/log through libmvec: a few ulp from the scalar path */
	return -psmch_Q_fast(m, &em->counts);
}

/* parameter space, initial parameters (core.c:32-49) and the first model; sq gives sum_n / sum_L only */
static int init_model(psmch_em_t *em, const psmch_opts_t *o, const psmch_seqs_t *sq, double (*rnd)(void))
{
	int k;
	const char *pattern = o->pattern ? o->pattern : "4+5*3+4";
	memset(em, 0, sizeof(*em));
	if (psmch_space_init(&em->sp, pattern, (o->flag & PSMCH_F_DIVERG) ? 1 : 0, o->alpha0) < 0) {
		fprintf(stderr, "psmc: bad pattern '%s'\n", pattern);
		return -1;
	}
	if (o->inp_ti) {
		em->sp.inp_ti = (double*)malloc(sizeof(double) * (em->sp.n + 1));
		memcpy(em->sp.inp_ti, o->inp_ti, sizeof(double) * (em->sp.n + 1));
	}
	em->exact_qd = o->exact_qd || (getenv("PSMC_B200_EXACT_QD") && atoi(getenv("PSMC_B200_EXACT_QD")) != 0);
	if (psmch_model_alloc(&em->model, &em->sp) < 0 || psmch_counts_alloc(&em->counts, em->sp.n + 1, em->exact_qd) < 0) return -1;
	em->post_sigma = (double*)calloc(em->sp.n + 1, sizeof(double));
	/* initial parameters (core.c:32-49) */
	if (o->inp_pa) {
		memcpy(em->model.params, o->inp_pa, sizeof(double) * em->sp.n_params);
	} else {
		const double theta = -log(1.0 - (double)sq->sum_n / sq->sum_L);
		em->model.params[0] = theta;
		em->model.params[1] = theta / o->tr_ratio;
		em->model.params[2] = o->max_t;
		for (k = PSMCH_N_PARAMS; k < em->sp.n_free + PSMCH_N_PARAMS; ++k) {
			em->model.params[k] = 1.0 + (rnd() * 2.0 - 1.0) * o->ran_init;
			if (em->model.params[k] < 0.1) em->model.params[k] = 0.1;
		}
		if (em->sp.diverg) em->model.params[em->sp.n_params - 1] = o->dt0;
	}
	psmch_model_update(&em->sp, em->model.params, &em->model);
	em->exact_mstep = getenv("PSMC_B200_EXACT_MSTEP") != 0; /* scalar libm in every trial evaluation */
	{	/* helper threads of the M-step (spec.c): 3 if every process on this node can have 4 cores, else 1; PSMC_B200_MSTEP_SPEC=0/1/3 overrides */
		const char *env = getenv("PSMC_B200_MSTEP_SPEC"), *lws = getenv("LOCAL_WORLD_SIZE");
		long cores = sysconf(_SC_NPROCESSORS_ONLN);
		{	/* the cores this process may actually run on (cgroup / taskset / srun --cpus-per-task) */
			cpu_set_t set;
			if (sched_getaffinity(0, sizeof(set), &set) == 0 && CPU_COUNT(&set) > 0 && CPU_COUNT(&set) < cores) cores = CPU_COUNT(&set);
		}
		const int procs = (lws && atoi(lws) > 0) ? atoi(lws) : 1;
		const long per = cores / procs;
		em->spec_mstep = per >= 4 ? 3 : (per >= 2 ? 1 : 0); /* (5 or 7 helpers measured within noise of 3 on a 16-core host) */
		if (env) { const int v = atoi(env); em->spec_mstep = v >= 7 ? 7 : (v >= 5 ? 5 : (v >= 3 ? 3 : (v >= 1 ? 1 : 0))); }
	}
	em->n_seqs = sq->n_seqs;
	return 0;
}

/* EM on a context that already holds the sequences (bootstrap replicates: the multiplicities are set by the
 * caller; sq carries the replicate's n_seqs / sum_L / sum_n).  The context is borrowed, not destroyed. */
int psmch_em_init_shared(psmch_em_t *em, const psmch_opts_t *o, const psmch_seqs_t *sq, psmc_b200_ctx *ctx, double (*rnd)(void))
{
	if (init_model(em, o, sq, rnd) != 0) return -1;
	em->spec_mstep = 0; /* bootstrap workers already occupy the host cores */
	em->n_gpus = 1;
	em->ctx[0] = ctx;
	em->borrowed = 1;
	if (em->exact_qd && psmc_b200_set_dense(ctx, 1) != 0) { /* (idempotent: the rows are allocated once per context) */
		fprintf(stderr, "psmc: --exact-qd: %s\n", psmc_b200_last_error());
		return -1;
	}
	return 0;
}

int psmch_em_init(psmch_em_t *em, const psmch_opts_t *o, const psmch_seqs_t *sq, double (*rnd)(void))
{
	int g, i, rc;
	if (init_model(em, o, sq, rnd) != 0) return -1;
	for (i = 0, em->n_seqs = 0; i < sq->n_seqs; ++i) /* HMM_TINY terms: one per NON-EMPTY record, as the single-GPU path counts them */
		if (sq->seqs[i].L > 0) ++em->n_seqs;
	/* shard whole sequences over the GPUs: longest-processing-time first */
	em->n_gpus = o->n_gpus;
	em->seq_owner = (int*)calloc(sq->n_seqs > 0 ? sq->n_seqs : 1, sizeof(int));
	{
		int *order = (int*)malloc(sizeof(int) * (sq->n_seqs > 0 ? sq->n_seqs : 1)), j;
		int64_t load[16] = {0};
		for (i = 0; i < sq->n_seqs; ++i) order[i] = i;
		for (i = 1; i < sq->n_seqs; ++i) { /* insertion sort by length, descending, stable */
			int v = order[i];
			for (j = i - 1; j >= 0 && sq->seqs[order[j]].L < sq->seqs[v].L; --j) order[j + 1] = order[j];
			order[j + 1] = v;
		}
		for (i = 0; i < sq->n_seqs; ++i) {
			int best = 0;
			for (g = 1; g < em->n_gpus; ++g)
				if (load[g] < load[best]) best = g;
			em->seq_owner[order[i]] = best;
			load[best] += sq->seqs[order[i]].L;
		}
		free(order);
	}
	for (g = 0; g < em->n_gpus; ++g) {
		int32_t *L = (int32_t*)malloc(sizeof(int32_t) * (sq->n_seqs > 0 ? sq->n_seqs : 1)), ns = 0;
		const signed char **ptr = (const signed char**)malloc(sizeof(void*) * (sq->n_seqs > 0 ? sq->n_seqs : 1));
		for (i = 0; i < sq->n_seqs; ++i)
			if (em->seq_owner[i] == g) { L[ns] = sq->seqs[i].L; ptr[ns] = sq->seqs[i].seq; ++ns; }
		rc = psmc_b200_create(&em->ctx[g], ns, L, ptr, em->sp.n + 1, o->devices[g], o->chunk_len, 0);
		free(L); free(ptr);
		if (rc != 0) {
			fprintf(stderr, "psmc: GPU E-step unavailable on device %d: %s\n", o->devices[g], psmc_b200_last_error());
			return -1;
		}
		if (em->exact_qd && psmc_b200_set_dense(em->ctx[g], 1) != 0) {
			fprintf(stderr, "psmc: --exact-qd: %s\n", psmc_b200_last_error());
			return -1;
		}
	}
	return 0;
}

void psmch_em_free(psmch_em_t *em)
{
	int g;
	for (g = 0; g < em->n_gpus && !em->borrowed; ++g) psmc_b200_destroy(em->ctx[g]);
	for (g = 0; g < em->n_spec; ++g) { psmch_spec_stop(em->spec[g]); psmch_model_free(&em->model_spec[g]); free(em->spec_aux[g]); }
	psmch_counts_free(&em->counts);
	psmch_model_free(&em->model);
	psmch_space_free(&em->sp);
	free(em->post_sigma); free(em->seq_owner);
	memset(em, 0, sizeof(*em));
}

/* E-step on all GPUs: enqueue everywhere, then sum the raw statistic vectors in device order
 * (a fixed order keeps 1-GPU and N-GPU runs reproducible; the Python multi-process driver does the
 * same sum with one NCCL all-reduce per iteration). */
static int estep_all(psmch_em_t *em)
{
	const int N = em->sp.n + 1, len = 7 * N + 1;
	psmc_b200_model mv;
	psmc_b200_stats sv;
	int g, i, rc;
	psmch_model_view(&em->model, &mv);
	psmch_counts_view(&em->counts, &sv);
	if (em->n_gpus == 1) {
		rc = psmc_b200_estep(em->ctx[0], &mv, &sv);
		if (rc) return rc;
	} else {
		double *tot = (double*)calloc(len, sizeof(double)), *raw = (double*)malloc(sizeof(double) * len);
		for (g = 0; g < em->n_gpus; ++g)
			if ((rc = psmc_b200_estep_launch(em->ctx[g], &mv)) != 0) { free(tot); free(raw); return rc; }
		for (g = 0; g < em->n_gpus; ++g) {
			if ((rc = psmc_b200_estep_fetch_raw(em->ctx[g], raw)) != 0) { free(tot); free(raw); return rc; }
			for (i = 0; i < len; ++i) tot[i] += raw[i];
		}
		rc = psmc_b200_unpack_stats(N, tot, em->n_seqs, &sv);
		free(tot); free(raw);
		if (rc) return rc;
	}
	em->counts.LL = sv.LL;
	if (em->exact_qd && em->counts.A) { /* hmm_expect's dense A summed over all sequences (khmm.c:346-352), for hmm_Q0 only */
		double *tmp = (double*)malloc(sizeof(double) * (size_t)N * N);
		memset(em->counts.A, 0, sizeof(double) * (size_t)N * N);
		for (g = 0; g < em->n_gpus; ++g) {
			if ((rc = psmc_b200_dense_counts(em->ctx[g], tmp)) != 0) { free(tmp); return rc; }
			for (i = 0; i < N * N; ++i) em->counts.A[i] += tmp[i];
		}
		free(tmp);
	}
	return 0;
}

int psmch_em_estep(psmch_em_t *em)
{
	double t0 = now_ms();
	int rc = estep_all(em);
	if (rc != 0) fprintf(stderr, "psmc: E-step failed: %s\n", psmc_b200_last_error());
	em->t_estep_ms = now_ms() - t0;
	return rc;
}

/* install an already reduced raw statistics vector (multi-process drivers: NCCL all-reduce outside) */
int psmch_em_set_raw(psmch_em_t *em, const double *raw, int64_t n_seqs_total)
{
	psmc_b200_stats sv;
	int rc;
	psmch_counts_view(&em->counts, &sv);
	rc = psmc_b200_unpack_stats(em->sp.n + 1, raw, n_seqs_total, &sv);
	if (rc == 0) em->counts.LL = sv.LL;
	return rc;
}

/* maximisation on the counts currently in em->counts (em.c:56-74) */
int psmch_em_mstep(psmch_em_t *em, FILE *fpout)
{
	const int N = em->sp.n + 1, np = em->sp.n_params;
	double *x, *last, sum = 0.0, t1 = now_ms();
	aux_t aux;
	int k, n_calls = 0;
	psmch_Q0(&em->counts);
	em->lk = em->counts.LL;
	x = (double*)malloc(sizeof(double) * np);
	memcpy(x, em->model.params, sizeof(double) * np);
	aux.em = em; aux.m = &em->model;
	em->Q0 = psmch_Q(&em->model, &em->counts);
	while (em->n_spec < em->spec_mstep) { /* first M-step: start the helpers, each with its own model instance */
		const int j = em->n_spec;
		aux_t *ha = (aux_t*)calloc(1, sizeof(aux_t));
		if (ha == 0 || psmch_model_alloc(&em->model_spec[j], &em->sp) != 0) { free(ha); em->spec_mstep = em->n_spec; break; }
		ha->em = em; ha->m = &em->model_spec[j];
		em->spec_aux[j] = ha;
		em->spec[j] = psmch_spec_start(objective, np, ha);
		if (em->spec[j] == 0) { psmch_model_free(&em->model_spec[j]); free(ha); em->spec_mstep = em->n_spec; break; }
		++em->n_spec;
	}
	last = (double*)malloc(sizeof(double) * np);
	memcpy(last, x, sizeof(double) * np);
	for (k = 0; k < em->n_spec; ++k) psmch_spec_begin(em->spec[k]);
	em->Q1 = -psmch_hooke_jeeves_spec(objective, em->spec, em->n_spec, np, x, &aux, PSMCH_HJ_RADIUS,
	                                  PSMCH_HJ_EPS, PSMCH_HJ_MAXCALL, last, &n_calls);
	for (k = 0; k < em->n_spec; ++k) psmch_spec_end(em->spec[k]);
	em->hj_calls = n_calls;
	if (fpout) fprintf(fpout, "IT\t%d\n", n_calls);
	for (k = 0; k < np; ++k) em->model.params[k] = fabs(last[k]);
	free(last);
	free(x);
	/* the model stays at the LAST EVALUATED trial point (em.c:61-67), recomputed with the scalar bit-reproducible path */
	psmch_model_update(&em->sp, em->model.params, &em->model);
	for (k = 0; k < N; ++k) sum += em->counts.E[k] + em->counts.E[N + k];
	for (k = 0; k < N; ++k) em->post_sigma[k] = (em->counts.E[k] + em->counts.E[N + k]) / sum;
	em->t_mstep_ms = now_ms() - t1;
	return 0;
}

int psmch_em_iterate(psmch_em_t *em, FILE *fpout)
{
	int rc = psmch_em_estep(em);
	if (rc != 0) return rc;
	return psmch_em_mstep(em, fpout);
}
