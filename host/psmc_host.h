/* psmc_host.h -- host side of the B200-native `psmc` driver (plain C, like the reference).
 *
 * Mirrors the reference's host interface for the path around the E-step (same option letters, same
 * .psmcfa grammar, same .psmc text), re-implemented from its behaviour:
 *   options / header text        cli.c:142-227        -> cli.c (this directory)
 *   pattern parser               cli.c:66-99          -> pattern.c
 *   .psmcfa reader               cli.c:103-138, kseq.h -> psmcfa.c
 *   params -> HMM                core.c:6-133         -> model.c   (O(N) factored form, not a dense matrix)
 *   Hooke-Jeeves                 kmin.c:48-107        -> hj.c
 *   EM iteration                 em.c:15-78           -> em.c      (E-step = psmc_b200_estep on the GPU)
 *   round printer, -i reader     aux.c:49-113         -> output.c
 *   bootstrap resampling         aux.c:8-47           -> resamp.c  (seedable); bootstrap.c (R replicates, one process)
 *   decode printer               aux.c:129-232        -> decode.c  (posteriors from psmc_b200_decode)
 */
#ifndef PSMC_HOST_H
#define PSMC_HOST_H

#include <stdint.h>
#include <stdio.h>
#include "psmc_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

#define PSMCH_VERSION "0.6.5-r74-dirty" /* the version string consumers see (cli.c:13); ours is in MM B200 line */
#define PSMCH_MAX_INTERVALS 4096 /* cap on the sum of the pattern lengths (the GPU path itself stops at 128 states) */
#define PSMCH_N_PARAMS 3                /* theta, rho, max_t (psmc.h:12) */
#define PSMCH_T_INF 1000.0              /* psmc.h:14 */
#define PSMCH_TINY 1e-25                /* khmm.h:28 */
#define PSMCH_INF 1e300                 /* khmm.h:29 */

#define PSMCH_F_DECODE 0x1
#define PSMCH_F_FULLDEC 0x2
#define PSMCH_F_SIMU 0x4
#define PSMCH_F_DIVERG 0x8
#define PSMCH_F_PROB 0x10

/* ---- sequences ------------------------------------------------------------------------------ */
typedef struct {
	int32_t L, L_e, n_e; /* bins, non-missing bins, het bins (cli.c:118-128) */
	signed char *seq;    /* values 0/1/2 */
	char *name;
} psmch_seq_t;

typedef struct {
	int n_seqs;
	psmch_seq_t *seqs;
	int64_t sum_L; /* sum of L_e (cli.c:133-137) */
	int64_t sum_n; /* sum of n_e */
} psmch_seqs_t;

int  psmch_read_psmcfa(const char *fn, psmch_seqs_t *out); /* "-" = stdin; gz or plain */
void psmch_free_seqs(psmch_seqs_t *s);
void psmch_resample(psmch_seqs_t *s, double (*rnd)(void)); /* aux.c:8-47; rnd() in [0,1) */
int  psmch_split(psmch_seqs_t *s, int trunk);               /* utils/splitfa.c:20-31 in memory */
void psmch_draw(const psmch_seqs_t *s, double (*rnd)(void), int32_t *mult, psmch_seqs_t *view); /* aux.c:14-32 as multiplicities; view: n_seqs, sum_L, sum_n of the replicate */

/* ---- parameter space and model ----------------------------------------------------------------- */
typedef struct {
	int n;        /* psmc's n; number of states N = n + 1 */
	int n_free;   /* free lambdas */
	int *par_map; /* n+1 entries */
	char *pattern;
	int n_params; /* n_free + 3 (+1 with the divergence parameter) */
	int diverg;
	double alpha0;  /* -l, default 0.1 (cli.c:151) */
	double *inp_ti; /* time intervals from -i (n+1 values) or NULL */
} psmch_space_t;

int  psmch_parse_pattern(const char *pattern, int *n_free, int **par_map); /* returns n or -1 */
int  psmch_space_init(psmch_space_t *sp, const char *pattern, int diverg, double alpha0);
void psmch_space_free(psmch_space_t *sp);

typedef struct {
	int N;
	double *params;                    /* n_params: the point the model currently describes */
	double *t;                         /* n+2 interval boundaries */
	double *sigma;                     /* N (also a0) */
	double *e;                         /* 2N */
	double *U, *V, *W, *Z, *D;         /* N each */
	double C_pi, C_sigma;
	double *lam, *alp, *bet, *qax, *tau; /* scratch */
	double *vw;                          /* scratch of the vectorised trial evaluations */
	int fast_valid;                      /* vw holds log e[0][k] = -theta (avg_t_k + dt) of the current point (set by psmch_model_update_fast) */
	double t_max_t;                      /* max_t the boundaries t[] were last computed for by the fast path (NaN = none) */
} psmch_model_t;

int  psmch_model_alloc(psmch_model_t *m, const psmch_space_t *sp);
void psmch_model_free(psmch_model_t *m);
void psmch_model_update(const psmch_space_t *sp, const double *params, psmch_model_t *m); /* core.c:61-133 */
void psmch_model_update_fast(const psmch_space_t *sp, const double *params, psmch_model_t *m); /* same, exp/log via libmvec: trial points of the M-step only */
void psmch_model_dense(const psmch_model_t *m, double *a /* N*N */);
void psmch_avg_t(const psmch_space_t *sp, const psmch_model_t *m, double *avg_t);           /* core.c:135-162 */
void psmch_model_view(const psmch_model_t *m, psmc_b200_model *v);

/* ---- expected counts and the EM objective ---------------------------------------------------- */
typedef struct {
	int N;
	double LL;
	double *E;                      /* 2N */
	double *RL, *CL, *RU, *CU, *AD; /* N each */
	double *A;                      /* N*N dense counts or NULL */
	double Q0;
} psmch_counts_t;

int    psmch_counts_alloc(psmch_counts_t *c, int N, int dense);
void   psmch_counts_free(psmch_counts_t *c);
void   psmch_counts_view(psmch_counts_t *c, psmc_b200_stats *v);
void   psmch_counts_from_dense(psmch_counts_t *c); /* marginals of c->A */
double psmch_Q0(psmch_counts_t *c);                 /* khmm.c:326-342; transition part needs c->A, else a marginal surrogate */
double psmch_Q(const psmch_model_t *m, const psmch_counts_t *c); /* khmm.c:363-382 in O(N) on the factored model */
double psmch_Q_fast(const psmch_model_t *m, const psmch_counts_t *c); /* same value to a few ulp; logs via libmvec (uses m->vw) */

/* ---- Hooke-Jeeves ---------------------------------------------------------------------------- */
typedef double (*psmch_func_t)(int n, double *x, void *data);
double psmch_hooke_jeeves(psmch_func_t f, int n, double *x, void *data, double r, double eps, int max_calls);
/* helper thread that evaluates the -step point of every probe concurrently (spec.c); same search path and call count */
typedef struct psmch_spec psmch_spec_t;
psmch_spec_t *psmch_spec_start(psmch_func_t f, int n, void *data);
void   psmch_spec_begin(psmch_spec_t *s);
void   psmch_spec_end(psmch_spec_t *s);
void   psmch_spec_submit(psmch_spec_t *s, const double *x);
double psmch_spec_wait(psmch_spec_t *s);
void   psmch_spec_stop(psmch_spec_t *s);
/* last (n doubles or NULL): the last evaluated point in the order of the sequential search; n_calls: evaluations counted */
double psmch_hooke_jeeves_spec(psmch_func_t f, psmch_spec_t **spec, int n_spec /* 0, 1, 3, 5 or 7 helpers */, int n, double *x, void *data,
                               double r, double eps, int max_calls, double *last, int *n_calls);
#define PSMCH_HJ_RADIUS 0.5
#define PSMCH_HJ_EPS 1e-7
#define PSMCH_HJ_MAXCALL 50000

/* ---- options, EM state, output ---------------------------------------------------------------- */
typedef struct {
	int flag, n_iters, cap_k, is_bootstrap;
	char *pattern, *pre_fn, *in_fn, *cnt_fn;
	FILE *fpout;
	double max_t, tr_ratio, alpha0, ran_init, dt0;
	double *inp_pa; /* parameters read by -i */
	double *inp_ti;
	/* B200 additions (long options / environment only; never change the .psmc text) */
	int n_gpus;          /* --gpus N   [1] */
	int devices[16];
	long seed;           /* --seed S   [-1 = time^pid as the reference, main.c:11] */
	int chunk_len;       /* --chunk L  [0 = auto] */
	int verbose;         /* --verbose  timing lines on stderr */
	int n_replicates;    /* --replicates R  R bootstrap replicates in this process (implies -b) [0] */
	int split_len;       /* --split[=T]  apply the splitfa rule with trunk size T bins first [off; 500000] */
	int slots;           /* --slots K  concurrent replicates per GPU [2] */
	int batch;           /* --batch B  replicates sharing one launch sequence [0 = as many as fit in device memory; 1 = one at a time (the older scheme)] */
	int batch_slots;     /* --batch-slots K  batched workers per GPU [1] */
	int exact_qd;        /* --exact-qd  dense transition counts for hmm_Q0 (khmm.c:336-340): the QD line as the original prints it */
} psmch_opts_t;

typedef struct {
	psmch_space_t sp;
	psmch_model_t model;
	psmch_counts_t counts;
	double *post_sigma;
	double lk, Q0, Q1;
	int n_gpus;
	psmc_b200_ctx *ctx[16];
	int *seq_owner; /* per sequence: which context holds it */
	int64_t n_seqs;
	int hj_calls;
	int exact_qd;    /* dense counts fetched after every E-step (counts.A) */
	int borrowed;    /* ctx[0] belongs to the caller (bootstrap replicates share one context per GPU slot) */
	int exact_mstep; /* PSMC_B200_EXACT_MSTEP: trial evaluations with scalar libm instead of libmvec */
	int spec_mstep;  /* speculative evaluator threads in the M-step: 0, 1, 3, 5 or 7 (PSMC_B200_MSTEP_SPEC; 0 for bootstrap workers) */
	int n_spec;      /* helpers running */
	psmch_model_t model_spec[7]; /* the helpers' own model instances */
	psmch_spec_t *spec[7];
	void *spec_aux[7];
	double t_estep_ms, t_mstep_ms; /* wall time of the last iteration */
} psmch_em_t;

int  psmch_parse_cli(int argc, char *argv[], psmch_opts_t *o);
void psmch_print_header(const psmch_opts_t *o, const psmch_space_t *sp, const psmch_seqs_t *sq, int stage);
int  psmch_read_param(psmch_opts_t *o, psmch_space_t *sp); /* aux.c:84-113 */
int  psmch_em_init(psmch_em_t *em, const psmch_opts_t *o, const psmch_seqs_t *sq, double (*rnd)(void));
int  psmch_em_init_shared(psmch_em_t *em, const psmch_opts_t *o, const psmch_seqs_t *sq, psmc_b200_ctx *ctx, double (*rnd)(void));
int  psmch_em_iterate(psmch_em_t *em, FILE *fpout); /* one psmc_em (em.c:27-78); prints the IT line */
int  psmch_em_estep(psmch_em_t *em);                 /* em.c:33-55 on the GPU(s) */
int  psmch_em_set_raw(psmch_em_t *em, const double *raw, int64_t n_seqs_total);
int  psmch_em_mstep(psmch_em_t *em, FILE *fpout);    /* em.c:56-74 */
void psmch_em_free(psmch_em_t *em);
void psmch_print_round(const psmch_opts_t *o, const psmch_em_t *em, const psmch_seqs_t *sq, FILE *fp); /* aux.c:49-82 */
int  psmch_decode(const psmch_opts_t *o, psmch_em_t *em, const psmch_seqs_t *sq, FILE *fp);          /* aux.c:129-232 */
int  psmch_bootstrap_run(const psmch_opts_t *o, const psmch_seqs_t *sq, int n_rep, int slots);     /* README:57-62 in one process */

#ifdef __cplusplus
}
#endif
#endif
