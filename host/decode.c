/* decode.c -- text of the decoding modes (aux.c:129-232): TC lines, then per sequence either
 * PR (-s, per-bin scale factors), DC (-d, runs of the posterior-argmax state with their maximum
 * posterior) or DF (-D, recombination probability and the full posterior row per bin).
 *
 * -d / -D at genome scale (BASELINE configs[4]): psmc_b200_decode_run decodes EVERY sequence of a context in one go
 * on the fast path; for -d the runs are compacted on the device, so a 3 Gbp genome returns ~0.5 M runs instead of
 * 12 bytes per bin; for -D the rows come back as float and the ~14 GB of text are written by a pool of threads with a
 * hand-rolled %.4f (the reference spends its -D time in fprintf, aux.c:183-200).  -s keeps the one-sequence,
 * double-precision psmc_b200_decode. */
#define _GNU_SOURCE
#include <sched.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <time.h>
#include <unistd.h>
#include <pthread.h>
#include "psmc_host.h"

static double now_ms(void)
{
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

static int n_threads(void)
{
	long c = sysconf(_SC_NPROCESSORS_ONLN);
	cpu_set_t set;
	if (sched_getaffinity(0, sizeof(set), &set) == 0 && CPU_COUNT(&set) > 0 && CPU_COUNT(&set) < c) c = CPU_COUNT(&set);
	return c < 1 ? 1 : (c > 32 ? 32 : (int)c);
}

/* ---- one thread per GPU: the contexts decode their sequences concurrently ---- */
typedef struct {
	psmc_b200_ctx *ctx;
	const psmc_b200_model *mv;
	uint32_t what;
	int rc;
	char err[256];
	int64_t nr;
	int32_t *rs, *rst, *rl;
	uint8_t *rk;
	double *rm;
	psmc_b200_info inf;
} gpu_job_t;

static void *gpu_thread(void *arg)
{
	gpu_job_t *j = (gpu_job_t*)arg;
	j->rc = psmc_b200_decode_run(j->ctx, j->mv, j->what);
	if (j->rc == 0 && (j->what & PSMC_B200_DEC_RUNS)) {
		j->rc = psmc_b200_decode_get_runs(j->ctx, 0, 0, 0, 0, 0, 0, &j->nr);
		if (j->rc == 0) {
			j->rs = (int32_t*)malloc(sizeof(int32_t) * (j->nr + 1)); j->rst = (int32_t*)malloc(sizeof(int32_t) * (j->nr + 1));
			j->rl = (int32_t*)malloc(sizeof(int32_t) * (j->nr + 1)); j->rk = (uint8_t*)malloc(j->nr + 1); j->rm = (double*)malloc(sizeof(double) * (j->nr + 1));
			j->rc = psmc_b200_decode_get_runs(j->ctx, j->nr, j->rs, j->rst, j->rl, j->rk, j->rm, &j->nr);
		}
	}
	if (j->rc != 0) snprintf(j->err, sizeof(j->err), "%s", psmc_b200_last_error()); /* (the message is per thread) */
	else psmc_b200_get_info(j->ctx, &j->inf);
	return 0;
}

/* runs `what` on every context that holds sequences; returns 0 or the first error (message on stderr) */
static int decode_all_gpus(psmch_em_t *em, const int *cnt, const psmc_b200_model *mv, uint32_t what, gpu_job_t *job, double *gpu_ms, double *dec_ms)
{
	pthread_t th[16];
	int g, rc = 0;
	const double t0 = now_ms();
	for (g = 0; g < em->n_gpus; ++g) {
		memset(&job[g], 0, sizeof(job[g]));
		job[g].ctx = em->ctx[g]; job[g].mv = mv; job[g].what = what;
		if (cnt[g] > 0) {
			if (em->n_gpus > 1) pthread_create(&th[g], 0, gpu_thread, &job[g]);
			else gpu_thread(&job[g]); /* (one GPU: stay on the thread that owns the CUDA state) */
		}
	}
	for (g = 0; g < em->n_gpus; ++g) {
		if (cnt[g] == 0) continue;
		if (em->n_gpus > 1) pthread_join(th[g], 0);
		if (job[g].rc != 0 && rc == 0) { rc = job[g].rc; fprintf(stderr, "psmc: GPU decode failed on device slot %d: %s\n", g, job[g].err); }
		if (job[g].inf.decode_ms[0] + job[g].inf.decode_ms[1] > *gpu_ms) *gpu_ms = job[g].inf.decode_ms[0] + job[g].inf.decode_ms[1]; /* concurrent: the slowest GPU */
	}
	*dec_ms += now_ms() - t0;
	return rc;
}

/* ---- DC: one buffer per sequence, formatted in parallel, written in order ---- */
typedef struct {
	const psmch_seqs_t *sq;
	const double *avg;     /* avg_t per state */
	double theta;
	/* runs of all sequences of all contexts, already in global sequence order */
	const int64_t *first;  /* per sequence: index of its first run, first[n_seqs] = total */
	const int32_t *start, *len;
	const uint8_t *state;
	const double *maxp;
	char **buf;
	size_t *blen;
	int *next;
	pthread_mutex_t *mu;
} dc_job_t;

static void *dc_thread(void *arg)
{
	dc_job_t *j = (dc_job_t*)arg;
	for (;;) {
		int i;
		int64_t r;
		size_t cap, n = 0;
		char *b;
		const char *name;
		pthread_mutex_lock(j->mu);
		i = (*j->next)++;
		pthread_mutex_unlock(j->mu);
		if (i >= j->sq->n_seqs) break;
		if (j->sq->seqs[i].L == 0) continue;
		name = j->sq->seqs[i].name;
		cap = (size_t)(j->first[i + 1] - j->first[i]) * (strlen(name) + 80) + 64;
		b = (char*)malloc(cap);
		for (r = j->first[i]; r < j->first[i + 1]; ++r) { /* aux.c:165-182: 1-based positions; the last run of a sequence prints %.3lf / %.2lf */
			const int k = j->state[r], beg = j->start[r] + 1, end = j->start[r] + j->len[r];
			if (r + 1 < j->first[i + 1]) n += sprintf(b + n, "DC\t%s\t%d\t%d\t%d\t%lf\t%.3lf\n", name, beg, end, k, j->avg[k] * j->theta, j->maxp[r]);
			else n += sprintf(b + n, "DC\t%s\t%d\t%d\t%d\t%.3lf\t%.2lf\n", name, beg, end, k, j->avg[k] * j->theta, j->maxp[r]);
		}
		j->buf[i] = b; j->blen[i] = n;
	}
	return 0;
}

/* ---- DF: slices of one sequence formatted in parallel ---- */
typedef struct {
	int N, k0, k1;          /* bins [k0, k1), 0-based */
	const float *post;      /* rows of the whole sequence */
	const double *prec;
	char *buf;
	size_t n;
} df_job_t;

static inline char *put_fixed4(char *p, float v) /* "%.4f" of a value in [0, 10) */
{
	unsigned x = (unsigned)((double)v * 10000.0 + 0.5);
	const unsigned ip = x / 10000u;
	x -= ip * 10000u;
	*p++ = (char)('0' + ip); *p++ = '.';
	*p++ = (char)('0' + x / 1000u); x %= 1000u;
	*p++ = (char)('0' + x / 100u); x %= 100u;
	*p++ = (char)('0' + x / 10u);
	*p++ = (char)('0' + x % 10u);
	return p;
}

static void *df_thread(void *arg)
{
	df_job_t *j = (df_job_t*)arg;
	char *p = j->buf;
	int k, l;
	for (k = j->k0; k < j->k1; ++k) { /* aux.c:195-198 */
		const float *row = j->post + (size_t)k * j->N;
		p += sprintf(p, "DF\t%d\t%lf", k + 1, j->prec[k]);
		for (l = 0; l < j->N; ++l) {
			*p++ = '\t';
			if (row[l] >= 0.0f && row[l] < 9.9999f) p = put_fixed4(p, row[l]);
			else p += sprintf(p, "%.4f", row[l]);
		}
		*p++ = '\n';
	}
	j->n = (size_t)(p - j->buf);
	return 0;
}

int psmch_decode(const psmch_opts_t *o, psmch_em_t *em, const psmch_seqs_t *sq, FILE *fp)
{
	const psmch_space_t *sp = &em->sp;
	const int N = sp->n + 1, nt = n_threads();
	const double theta = em->model.params[0];
	double *avg = (double*)malloc(sizeof(double) * N);
	int *local = (int*)calloc(sq->n_seqs > 0 ? sq->n_seqs : 1, sizeof(int)); /* index of sequence i inside its context (records as given) */
	int cnt[16] = {0}, primed[16] = {0};
	int i, k, g, rc = 0;
	const double t_all = now_ms();
	double gpu_ms = 0.0, dec_ms = 0.0, fmt_ms = 0.0;
	int64_t bins = 0, total_runs = 0;
	psmc_b200_model mv;
	psmch_model_view(&em->model, &mv);
	psmch_avg_t(sp, &em->model, avg);
	for (k = 0; k < N; ++k) {
		if (avg[k] < em->model.t[k] || avg[k] > em->model.t[k + 1])
			fprintf(stderr, "ERROR: (%f <= %f <= %f) does not stand. Contact me if you see this.\n", em->model.t[k], avg[k], em->model.t[k + 1]);
		fprintf(fp, "TC\t%d\t%lf\t%lf\t%lf\n", k, em->model.t[k] * theta, avg[k] * theta, em->model.t[k + 1] * theta);
	}
	for (i = 0; i < sq->n_seqs; ++i) { local[i] = cnt[em->seq_owner[i]]++; bins += sq->seqs[i].L; }
	if (o->flag & PSMCH_F_PROB) { /* -s: per-bin scale factors (aux.c:159-164), one sequence at a time */
		for (i = 0; i < sq->n_seqs && rc == 0; ++i) {
			const psmch_seq_t *s = sq->seqs + i;
			const int L = s->L;
			int32_t *bk;
			double *bp, *sc;
			if (L == 0) continue;
			g = em->seq_owner[i];
			bk = (int32_t*)malloc(sizeof(int32_t) * L); bp = (double*)malloc(sizeof(double) * L); sc = (double*)malloc(sizeof(double) * L);
			rc = psmc_b200_decode(em->ctx[g], primed[g] ? 0 : &mv, local[i], bk, bp, 0, 0, sc);
			primed[g] = 1;
			if (rc != 0) fprintf(stderr, "psmc: GPU decode failed: %s\n", psmc_b200_last_error());
			else {
				fprintf(fp, "PR\t%s\t%d", s->name, L);
				for (k = 0; k < L; ++k) fprintf(fp, "\t%.3f", sc[k]);
				fprintf(fp, "\n");
				fflush(fp);
			}
			free(bk); free(bp); free(sc);
		}
	} else if (!(o->flag & PSMCH_F_FULLDEC)) { /* -d */
		int64_t *first = (int64_t*)calloc(sq->n_seqs + 2, sizeof(int64_t)), tot = 0, nr[16] = {0}, pos[16] = {0};
		int32_t *rs[16] = {0}, *rst[16] = {0}, *rl[16] = {0}, *st = 0, *ln = 0;
		uint8_t *rk[16] = {0}, *ks = 0;
		double *rm[16] = {0}, *mp = 0, t0;
		char **buf = (char**)calloc(sq->n_seqs > 0 ? sq->n_seqs : 1, sizeof(char*));
		size_t *blen = (size_t*)calloc(sq->n_seqs > 0 ? sq->n_seqs : 1, sizeof(size_t));
		pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER;
		gpu_job_t job[16];
		rc = decode_all_gpus(em, cnt, &mv, PSMC_B200_DEC_RUNS, job, &gpu_ms, &dec_ms); /* every GPU decodes its sequences; only the runs come back */
		for (g = 0; g < em->n_gpus; ++g) {
			rs[g] = job[g].rs; rst[g] = job[g].rst; rl[g] = job[g].rl; rk[g] = job[g].rk; rm[g] = job[g].rm; nr[g] = job[g].nr;
			tot += nr[g];
		}
		if (rc == 0) {
			/* runs of all contexts in global sequence order (a context returns its sequences in its own order) */
			st = (int32_t*)malloc(sizeof(int32_t) * (tot + 1)); ln = (int32_t*)malloc(sizeof(int32_t) * (tot + 1));
			ks = (uint8_t*)malloc(tot + 1); mp = (double*)malloc(sizeof(double) * (tot + 1));
			tot = 0;
			for (i = 0; i < sq->n_seqs; ++i) {
				g = em->seq_owner[i];
				first[i] = tot;
				while (pos[g] < nr[g] && rs[g][pos[g]] == local[i]) {
					st[tot] = rst[g][pos[g]]; ln[tot] = rl[g][pos[g]]; ks[tot] = rk[g][pos[g]]; mp[tot] = rm[g][pos[g]];
					++tot; ++pos[g];
				}
			}
			first[sq->n_seqs] = tot;
			total_runs = tot;
			t0 = now_ms();
			{
				dc_job_t job;
				pthread_t th[32];
				int next = 0, n = nt < sq->n_seqs ? nt : (sq->n_seqs > 0 ? sq->n_seqs : 1);
				job.sq = sq; job.avg = avg; job.theta = theta; job.first = first; job.start = st; job.len = ln; job.state = ks; job.maxp = mp;
				job.buf = buf; job.blen = blen; job.next = &next; job.mu = &mu;
				for (k = 1; k < n; ++k) pthread_create(&th[k], 0, dc_thread, &job);
				dc_thread(&job);
				for (k = 1; k < n; ++k) pthread_join(th[k], 0);
			}
			for (i = 0; i < sq->n_seqs; ++i)
				if (buf[i]) { fwrite(buf[i], 1, blen[i], fp); free(buf[i]); }
			fflush(fp);
			fmt_ms = now_ms() - t0;
		}
		for (g = 0; g < 16; ++g) { free(rs[g]); free(rst[g]); free(rl[g]); free(rk[g]); free(rm[g]); }
		free(first); free(st); free(ln); free(ks); free(mp); free(buf); free(blen);
	} else { /* -D */
		gpu_job_t job[16];
		rc = decode_all_gpus(em, cnt, &mv, PSMC_B200_DEC_POST, job, &gpu_ms, &dec_ms);
		for (i = 0; i < sq->n_seqs && rc == 0; ++i) {
			const psmch_seq_t *s = sq->seqs + i;
			const int L = s->L, line = 32 + 7 * N, blk = 16384;
			float *post;
			double *prec, t0;
			int b0;
			if (L == 0) continue;
			post = (float*)malloc(sizeof(float) * (size_t)L * N); prec = (double*)malloc(sizeof(double) * L);
			t0 = now_ms();
			rc = psmc_b200_decode_get_bins(em->ctx[em->seq_owner[i]], local[i], 0, 0, post, prec);
			dec_ms += now_ms() - t0;
			if (rc != 0) { fprintf(stderr, "psmc: GPU decode failed: %s\n", psmc_b200_last_error()); free(post); free(prec); break; }
			t0 = now_ms();
			for (b0 = 0; b0 < L; b0 += blk * nt) { /* nt slices of blk bins at a time: bounded memory, output in order */
				df_job_t job[32];
				pthread_t th[32];
				int n = 0;
				for (k = 0; k < nt && b0 + k * blk < L; ++k, ++n) {
					job[k].N = N; job[k].k0 = b0 + k * blk; job[k].k1 = job[k].k0 + blk < L ? job[k].k0 + blk : L;
					job[k].post = post; job[k].prec = prec; job[k].buf = (char*)malloc((size_t)line * (job[k].k1 - job[k].k0) + 64);
					if (k > 0) pthread_create(&th[k], 0, df_thread, &job[k]);
				}
				df_thread(&job[0]);
				for (k = 1; k < n; ++k) pthread_join(th[k], 0);
				for (k = 0; k < n; ++k) { fwrite(job[k].buf, 1, job[k].n, fp); free(job[k].buf); }
			}
			fmt_ms += now_ms() - t0;
			free(post); free(prec);
		}
	}
	if (o->verbose || getenv("PSMC_B200_TIMING"))
		fprintf(stderr, "[psmc-b200] decode: %.3f s for %lld bins on %d GPU(s) (device: E-step + decode kernels %.1f ms; decode calls %.1f ms wall; %lld runs; formatting + writing %.1f ms on %d threads)\n",
		        (now_ms() - t_all) * 1e-3, (long long)bins, em->n_gpus, gpu_ms, dec_ms, (long long)total_runs, fmt_ms, nt);
	free(avg); free(local);
	return rc;
}
