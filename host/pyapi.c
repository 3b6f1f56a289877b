/* pyapi.c -- flat, ctypes-friendly entry points of libpsmc_host.so (used by psmc_b200/host.py, the
 * tests and bench.py).  No numerics of its own: thin wrappers over model.c / objective.c / hj.c / em.c. */
#include <stdlib.h>
#include <string.h>
#include "psmc_host.h"

int psmch_py_pattern(const char *pattern, int *n_free, int *par_map /* may be NULL */)
{
	int *pm = 0, nf = 0, n = psmch_parse_pattern(pattern, &nf, &pm), k;
	if (n < 0) return n;
	if (n_free) *n_free = nf;
	if (par_map) for (k = 0; k <= n; ++k) par_map[k] = pm[k];
	free(pm);
	return n;
}

/* out: t[n+2], sigma[N], e[2N], U V W Z D [N each], cc[2] = {C_pi, C_sigma} */
int psmch_py_model(const char *pattern, const double *params, double alpha0, int diverg, const double *inp_ti,
                   double *t, double *sigma, double *e, double *U, double *V, double *W, double *Z, double *D, double *cc)
{
	psmch_space_t sp;
	psmch_model_t m;
	int N;
	if (psmch_space_init(&sp, pattern, diverg, alpha0) < 0) return -1;
	N = sp.n + 1;
	if (inp_ti) { sp.inp_ti = (double*)malloc(sizeof(double) * N); memcpy(sp.inp_ti, inp_ti, sizeof(double) * N); }
	psmch_model_alloc(&m, &sp);
	psmch_model_update(&sp, params, &m);
	memcpy(t, m.t, sizeof(double) * (N + 1));
	memcpy(sigma, m.sigma, sizeof(double) * N);
	memcpy(e, m.e, sizeof(double) * 2 * N);
	memcpy(U, m.U, sizeof(double) * N); memcpy(V, m.V, sizeof(double) * N); memcpy(W, m.W, sizeof(double) * N);
	memcpy(Z, m.Z, sizeof(double) * N); memcpy(D, m.D, sizeof(double) * N);
	cc[0] = m.C_pi; cc[1] = m.C_sigma;
	psmch_model_free(&m);
	psmch_space_free(&sp);
	return N;
}

int psmch_py_avg_t(const char *pattern, const double *params, double alpha0, int diverg, double *avg_t)
{
	psmch_space_t sp;
	psmch_model_t m;
	if (psmch_space_init(&sp, pattern, diverg, alpha0) < 0) return -1;
	psmch_model_alloc(&m, &sp);
	psmch_model_update(&sp, params, &m);
	psmch_avg_t(&sp, &m, avg_t);
	psmch_model_free(&m);
	psmch_space_free(&sp);
	return 0;
}

double psmch_py_hj(psmch_func_t f, int n, double *x, void *data, double r, double eps, int max_calls)
{
	return psmch_hooke_jeeves(f, n, x, data, r, eps, max_calls);
}

/* ---- an EM session behind an opaque handle ------------------------------------------------------- */
typedef struct {
	psmch_opts_t o;
	psmch_seqs_t sq;
	psmch_em_t em;
} session_t;

static double rnd48(void) { return drand48(); }

/* sequences are borrowed for the duration of the call only (the contexts keep their own packed copy) */
void *psmch_py_em_create(const char *pattern, int n_seqs, const int32_t *L, const signed char *cat,
                         double max_t, double tr_ratio, double alpha0, const double *init_params,
                         int n_gpus, const int *devices, int chunk_len)
{
	session_t *s = (session_t*)calloc(1, sizeof(session_t));
	const signed char *p = cat;
	int i, j;
	s->o.pattern = strdup(pattern);
	s->o.max_t = max_t; s->o.tr_ratio = tr_ratio; s->o.alpha0 = alpha0; s->o.dt0 = -1.0;
	s->o.n_gpus = n_gpus < 1 ? 1 : n_gpus;
	for (i = 0; i < 16; ++i) s->o.devices[i] = (devices && i < n_gpus) ? devices[i] : i;
	s->o.chunk_len = chunk_len;
	s->sq.n_seqs = n_seqs;
	s->sq.seqs = (psmch_seq_t*)calloc(n_seqs > 0 ? n_seqs : 1, sizeof(psmch_seq_t));
	for (i = 0; i < n_seqs; ++i) {
		psmch_seq_t *q = s->sq.seqs + i;
		q->L = L[i]; q->seq = (signed char*)p; q->name = 0;
		for (j = 0; j < L[i]; ++j)
			if (p[j] == 0 || p[j] == 1) { ++q->L_e; if (p[j] == 1) ++q->n_e; }
		s->sq.sum_L += q->L_e; s->sq.sum_n += q->n_e;
		p += L[i];
	}
	if (init_params) {
		int nf = 0, n = psmch_parse_pattern(pattern, &nf, 0);
		if (n < 0) { free(s->sq.seqs); free(s->o.pattern); free(s); return 0; }
		s->o.inp_pa = (double*)malloc(sizeof(double) * (nf + PSMCH_N_PARAMS + 1));
		memcpy(s->o.inp_pa, init_params, sizeof(double) * (nf + PSMCH_N_PARAMS));
	}
	if (psmch_em_init(&s->em, &s->o, &s->sq, rnd48) != 0) {
		free(s->sq.seqs); free(s->o.pattern); free(s->o.inp_pa); free(s);
		return 0;
	}
	for (i = 0; i < n_seqs; ++i) s->sq.seqs[i].seq = 0; /* borrowed */
	return s;
}

void psmch_py_em_destroy(void *h)
{
	session_t *s = (session_t*)h;
	if (!s) return;
	psmch_em_free(&s->em);
	free(s->sq.seqs); free(s->o.pattern); free(s->o.inp_pa);
	free(s);
}

int psmch_py_em_iterate(void *h) { return psmch_em_iterate(&((session_t*)h)->em, 0); }
int psmch_py_em_estep(void *h) { return psmch_em_estep(&((session_t*)h)->em); }
int psmch_py_em_mstep(void *h) { return psmch_em_mstep(&((session_t*)h)->em, 0); }
/* install parameters found elsewhere (multi-process drivers: the M-step runs on one rank, the others receive its result) */
int psmch_py_em_set_params(void *h, const double *params)
{
	session_t *s = (session_t*)h;
	memcpy(s->em.model.params, params, sizeof(double) * s->em.sp.n_params);
	psmch_model_update(&s->em.sp, s->em.model.params, &s->em.model);
	return 0;
}
int psmch_py_em_set_raw(void *h, const double *raw, long long n_seqs_total) { return psmch_em_set_raw(&((session_t*)h)->em, raw, n_seqs_total); }
void *psmch_py_em_ctx(void *h, int g) { session_t *s = (session_t*)h; return (g >= 0 && g < s->em.n_gpus) ? s->em.ctx[g] : 0; }
int psmch_py_em_launch(void *h)
{
	session_t *s = (session_t*)h;
	psmc_b200_model mv;
	int g, rc;
	psmch_model_view(&s->em.model, &mv);
	for (g = 0; g < s->em.n_gpus; ++g)
		if ((rc = psmc_b200_estep_launch(s->em.ctx[g], &mv)) != 0) return rc;
	return 0;
}
int psmch_py_em_dims(void *h, int *n, int *n_free, int *n_params)
{
	session_t *s = (session_t*)h;
	*n = s->em.sp.n; *n_free = s->em.sp.n_free; *n_params = s->em.sp.n_params;
	return 0;
}
/* scalars: lk, Q0, Q1, hj_calls, t_estep_ms, t_mstep_ms, C_pi, C_sigma, sum_L, sum_n */
void psmch_py_em_scalars(void *h, double *out)
{
	session_t *s = (session_t*)h;
	out[0] = s->em.lk; out[1] = s->em.Q0; out[2] = s->em.Q1; out[3] = s->em.hj_calls;
	out[4] = s->em.t_estep_ms; out[5] = s->em.t_mstep_ms; out[6] = s->em.model.C_pi; out[7] = s->em.model.C_sigma;
	out[8] = (double)s->sq.sum_L; out[9] = (double)s->sq.sum_n;
}
/* vectors: params[n_params], t[n+2], sigma[N], post_sigma[N] */
void psmch_py_em_vectors(void *h, double *params, double *t, double *sigma, double *post_sigma)
{
	session_t *s = (session_t*)h;
	const int N = s->em.sp.n + 1;
	memcpy(params, s->em.model.params, sizeof(double) * s->em.sp.n_params);
	memcpy(t, s->em.model.t, sizeof(double) * (N + 1));
	memcpy(sigma, s->em.model.sigma, sizeof(double) * N);
	memcpy(post_sigma, s->em.post_sigma, sizeof(double) * N);
}
/* counts of the last E-step: E[2N], RL CL RU CU AD [N each] */
void psmch_py_em_counts(void *h, double *E, double *RL, double *CL, double *RU, double *CU, double *AD)
{
	session_t *s = (session_t*)h;
	const int N = s->em.sp.n + 1;
	memcpy(E, s->em.counts.E, sizeof(double) * 2 * N);
	memcpy(RL, s->em.counts.RL, sizeof(double) * N); memcpy(CL, s->em.counts.CL, sizeof(double) * N);
	memcpy(RU, s->em.counts.RU, sizeof(double) * N); memcpy(CU, s->em.counts.CU, sizeof(double) * N);
	memcpy(AD, s->em.counts.AD, sizeof(double) * N);
}

/* ---- M-step on caller-supplied counts (no GPU involved): used by the CPU tests --------------------
 * params: in = start point, out = last evaluated point (the reference's quirk); res = {Q0(before), Q1, calls, Q0_offset}
 * A may be NULL (then the structured marginals are taken as given). */
typedef struct { psmch_space_t *sp; psmch_model_t *m; psmch_counts_t *c; int cnt, fast; } maux_t;
static double mobjective(int n, double *x, void *data)
{
	maux_t *a = (maux_t*)data;
	int i;
	++a->cnt;
	for (i = 0; i < n; ++i) a->m->params[i] = x[i] < 0 ? -x[i] : x[i];
	if (a->fast) {
		psmch_model_update_fast(a->sp, a->m->params, a->m);
		return -psmch_Q_fast(a->m, a->c);
	}
	psmch_model_update(a->sp, a->m->params, a->m);
	return -psmch_Q(a->m, a->c);
}
int psmch_py_mstep(const char *pattern, double alpha0, double *params, const double *E, const double *A,
                   const double *RL, const double *CL, const double *RU, const double *CU, const double *AD, double *res)
{
	psmch_space_t sp;
	psmch_model_t m;
	psmch_counts_t c;
	maux_t a;
	double *x;
	int N;
	if (psmch_space_init(&sp, pattern, 0, alpha0) < 0) return -1;
	N = sp.n + 1;
	psmch_model_alloc(&m, &sp);
	psmch_counts_alloc(&c, N, A != 0);
	memcpy(c.E, E, sizeof(double) * 2 * N);
	if (A) { memcpy(c.A, A, sizeof(double) * N * N); psmch_counts_from_dense(&c); }
	else {
		memcpy(c.RL, RL, sizeof(double) * N); memcpy(c.CL, CL, sizeof(double) * N); memcpy(c.RU, RU, sizeof(double) * N);
		memcpy(c.CU, CU, sizeof(double) * N); memcpy(c.AD, AD, sizeof(double) * N);
	}
	psmch_model_update(&sp, params, &m);
	res[3] = psmch_Q0(&c);
	res[0] = psmch_Q(&m, &c);
	x = (double*)malloc(sizeof(double) * sp.n_params);
	memcpy(x, params, sizeof(double) * sp.n_params);
	a.sp = &sp; a.m = &m; a.c = &c; a.cnt = 0; a.fast = getenv("PSMC_B200_EXACT_MSTEP") == 0;
	{	/* PSMC_B200_MSTEP_SPEC=1|3: speculative trial points on helper threads (spec.c), as the EM driver does */
		const int want = getenv("PSMC_B200_MSTEP_SPEC") ? atoi(getenv("PSMC_B200_MSTEP_SPEC")) : 0;
		psmch_model_t m2[7];
		maux_t a2[7];
		psmch_spec_t *spec[7] = {0, 0, 0, 0, 0, 0, 0};
		double *last = (double*)malloc(sizeof(double) * sp.n_params);
		int n_calls = 0, i, ns = 0;
		for (i = 0; i < (want >= 7 ? 7 : (want >= 5 ? 5 : (want >= 3 ? 3 : (want >= 1 ? 1 : 0)))); ++i) {
			if (psmch_model_alloc(&m2[i], &sp) != 0) break;
			a2[i] = a; a2[i].m = &m2[i];
			spec[i] = psmch_spec_start(mobjective, sp.n_params, &a2[i]);
			if (spec[i] == 0) { psmch_model_free(&m2[i]); break; }
			psmch_spec_begin(spec[i]);
			++ns;
		}
		memcpy(last, x, sizeof(double) * sp.n_params);
		res[1] = -psmch_hooke_jeeves_spec(mobjective, spec, ns, sp.n_params, x, &a, PSMCH_HJ_RADIUS, PSMCH_HJ_EPS,
		                                  PSMCH_HJ_MAXCALL, last, &n_calls);
		for (i = 0; i < ns; ++i) { psmch_spec_end(spec[i]); psmch_spec_stop(spec[i]); psmch_model_free(&m2[i]); }
		res[2] = n_calls;
		for (i = 0; i < sp.n_params; ++i) m.params[i] = last[i] < 0 ? -last[i] : last[i]; /* the last evaluated point (em.c:61-67) */
		free(last);
	}
	memcpy(params, m.params, sizeof(double) * sp.n_params);
	free(x);
	psmch_model_update(&sp, m.params, &m);
	psmch_counts_free(&c); psmch_model_free(&m); psmch_space_free(&sp);
	return 0;
}

/* ---- .psmcfa reader through the product parser ------------------------------------------------- */
void *psmch_py_read(const char *fn)
{
	psmch_seqs_t *sq = (psmch_seqs_t*)calloc(1, sizeof(psmch_seqs_t));
	if (psmch_read_psmcfa(fn, sq) != 0) { free(sq); return 0; }
	return sq;
}
int psmch_py_read_n(void *h) { return ((psmch_seqs_t*)h)->n_seqs; }
long long psmch_py_read_sum(void *h, int which) { return which ? ((psmch_seqs_t*)h)->sum_n : ((psmch_seqs_t*)h)->sum_L; }
int psmch_py_read_len(void *h, int i) { return ((psmch_seqs_t*)h)->seqs[i].L; }
const char *psmch_py_read_name(void *h, int i) { return ((psmch_seqs_t*)h)->seqs[i].name; }
const signed char *psmch_py_read_seq(void *h, int i) { return ((psmch_seqs_t*)h)->seqs[i].seq; }
void psmch_py_read_free(void *h) { psmch_free_seqs((psmch_seqs_t*)h); free(h); }
static double (*g_rnd_cb)(void);
void psmch_py_resample(void *h, double (*rnd)(void)) { g_rnd_cb = rnd; psmch_resample((psmch_seqs_t*)h, rnd); }

/* re-upload every sequence from host memory to the context that owns it (end-to-end benchmark leg) */
int psmch_py_em_upload(void *h, int n_seqs, const int32_t *L, const signed char *cat)
{
	session_t *s = (session_t*)h;
	const signed char **all = (const signed char**)malloc(sizeof(void*) * (n_seqs > 0 ? n_seqs : 1));
	const signed char **ptr = (const signed char**)malloc(sizeof(void*) * (n_seqs > 0 ? n_seqs : 1));
	int32_t *len = (int32_t*)malloc(sizeof(int32_t) * (n_seqs > 0 ? n_seqs : 1));
	const signed char *p = cat;
	int i, g, rc = 0;
	if (n_seqs != s->sq.n_seqs) { free(all); free(ptr); free(len); return -1; }
	for (i = 0; i < n_seqs; ++i) { all[i] = p; p += L[i]; }
	for (g = 0; g < s->em.n_gpus && rc == 0; ++g) {
		int ns = 0;
		for (i = 0; i < n_seqs; ++i)
			if (s->em.seq_owner[i] == g) { len[ns] = L[i]; ptr[ns] = all[i]; ++ns; }
		rc = psmc_b200_upload(s->em.ctx[g], ns, len, ptr);
	}
	free(all); free(ptr); free(len);
	return rc;
}

/* ---- bootstrap helpers (tests): the splitfa rule and the replicate draw --------------------------- */
static void seqs_from_lengths(psmch_seqs_t *sq, int n, const int32_t *L)
{
	int i;
	memset(sq, 0, sizeof(*sq));
	sq->n_seqs = n;
	sq->seqs = (psmch_seq_t*)calloc(n > 0 ? n : 1, sizeof(psmch_seq_t));
	for (i = 0; i < n; ++i) {
		char nm[32];
		int32_t u;
		psmch_seq_t *s = sq->seqs + i;
		s->L = L[i];
		s->seq = (signed char*)malloc(L[i] > 0 ? L[i] : 1);
		for (u = 0; u < L[i]; ++u) s->seq[u] = (signed char)((u * 7 + i) % 23 == 0 ? 1 : ((u + i) % 41 == 0 ? 2 : 0));
		for (u = 0; u < L[i]; ++u) { if (s->seq[u] < 2) ++s->L_e; if (s->seq[u] == 1) ++s->n_e; }
		sprintf(nm, "r%d", i);
		s->name = strdup(nm);
		sq->sum_L += s->L_e; sq->sum_n += s->n_e;
	}
}

/* piece lengths of the splitfa rule; rec[i] = source record, idx[i] = 1-based piece number; returns the count */
int psmch_py_split(int n, const int32_t *L, int trunk, int32_t *out_L, int32_t *rec, int32_t *idx, int cap)
{
	psmch_seqs_t sq;
	int i, m;
	seqs_from_lengths(&sq, n, L);
	if (psmch_split(&sq, trunk) != 0) { psmch_free_seqs(&sq); return -1; }
	m = sq.n_seqs;
	for (i = 0; i < m && i < cap; ++i) {
		out_L[i] = sq.seqs[i].L;
		sscanf(sq.seqs[i].name, "r%d_%d", &rec[i], &idx[i]);
	}
	psmch_free_seqs(&sq);
	return m;
}

static unsigned short py_x[3];
static double py_rnd(void) { return erand48(py_x); }

/* multiplicities of one replicate drawn with srand48(seed) semantics; view = {n_seqs, sum_L, sum_n};
 * also runs psmch_resample (the copying path of `psmc -b`) from the same seed and returns in mult_copy the
 * multiplicities it produced (matched by record name), so that a test can require them to agree */
int psmch_py_draw(int n, const int32_t *L, long seed, int32_t *mult, int64_t *view, int32_t *mult_copy)
{
	psmch_seqs_t sq, v;
	int i;
	seqs_from_lengths(&sq, n, L);
	py_x[0] = 0x330E; py_x[1] = (unsigned short)(seed & 0xffff); py_x[2] = (unsigned short)((seed >> 16) & 0xffff);
	psmch_draw(&sq, py_rnd, mult, &v);
	view[0] = v.n_seqs; view[1] = v.sum_L; view[2] = v.sum_n;
	srand48(seed);
	psmch_resample(&sq, rnd48);
	for (i = 0; i < n; ++i) mult_copy[i] = 0;
	for (i = 0; i < sq.n_seqs; ++i) ++mult_copy[atoi(sq.seqs[i].name + 1)];
	view[3] = sq.n_seqs; view[4] = sq.sum_L; view[5] = sq.sum_n;
	psmch_free_seqs(&sq);
	return 0;
}
