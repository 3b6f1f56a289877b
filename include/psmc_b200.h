/* psmc_b200.h -- C ABI of the B200-native PSMC E-step (libpsmc_b200.so).
 *
 * The reference (lh3/psmc) has no plugin/FFI layer; its seam for this path is the khmm.h C API as
 * used by exactly two call sites: psmc_em (em.c:33-55) and psmc_decode (aux.c:150-221).  The entry
 * points below are what a maintainer binds in place of those loops.
 * Plain pointers and sizes only; no torch / CUDA types in the signatures.
 *
 * Conventions
 *   - N = number of HMM states (= psmc_par_t::n + 1, core.c:27).  Supported on the GPU: 1 <= N <= 128.
 *   - observation symbols: 0 = hom, 1 = het, 2 (or anything else) = missing (cli.c:15-32, khmm.c:21).
 *   - all floating point is FP64 (khmm.h:25-27 FLOAT == double).
 *   - every function returns 0 on success or a negative PSMC_B200_E* code; psmc_b200_last_error()
 *     gives the message (the reference asserts/aborts instead: khmm.c:216,250,301).
 *   - there is NO CPU fallback: if no CUDA device is usable, create() fails with PSMC_B200_ENODEV.
 */
#ifndef PSMC_B200_H
#define PSMC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PSMC_B200_VERSION 100  /* 0.1.0 */

#define PSMC_B200_EINVAL   (-1) /* bad argument */
#define PSMC_B200_ENODEV   (-2) /* no usable CUDA device / CUDA runtime error at init */
#define PSMC_B200_ECUDA    (-3) /* CUDA error while running */
#define PSMC_B200_ESTRUCT  (-4) /* dense transition matrix is not diagonal + rank-1 lower + rank-1 upper (e.g. after -C, aux.c:115-127) */
#define PSMC_B200_ENUMERIC (-5) /* non-finite value in the model or in the results */

/* create() flags */
#define PSMC_B200_F_DEFAULT 0

typedef struct psmc_b200_ctx psmc_b200_ctx;

/* Factored PSMC model (derived from core.c:100-125):
 *   a[k][l] = U[k]*V[l] (l<k),  W[k]*Z[l] (l>k),  D[k] (l==k);   e[0][k], e[1][k];  a0[k].
 * It replaces hmm_par_t (khmm.h:31-37) on this path. */
typedef struct {
	int32_t n_states;
	const double *a0;  /* N      hmm_par_t::a0 (khmm.h:36) = sigma_k (core.c:123) */
	const double *e;   /* 2*N    hmm_par_t::e rows 0 and 1 (khmm.h:34); the missing row is implicit 1.0 (khmm.c:21) */
	const double *U, *V, *W, *Z, *D; /* N each */
} psmc_b200_model;

/* Expected counts of one E-step, summed over all sequences.  Replaces hmm_exp_t (khmm.h:49-53) as
 * filled by hmm_expect + hmm_add_expect (khmm.c:297-359) including the per-sequence HMM_TINY
 * initialisation (khmm.c:305-308):
 *   E[b*N+k]      = he_sum->E[b][k], b = 0,1
 *   RL[k] = sum_{l<k} A[k][l]   CL[l] = sum_{k>l} A[k][l]
 *   RU[k] = sum_{l>k} A[k][l]   CU[l] = sum_{k<l} A[k][l]   AD[k] = A[k][k]
 * which are sufficient for hmm_Q (khmm.c:363-382) under the factored model. */
typedef struct {
	double LL;         /* sum over sequences of hmm_lk (khmm.c:245-260, em.c:48) */
	double *E;         /* 2*N, caller-owned */
	double *RL, *CL, *RU, *CU, *AD; /* N each, caller-owned */
} psmc_b200_stats;

/* Layout/timing report of a context (for benchmarks and tests). */
typedef struct {
	int32_t device, n_states, n_states_padded, n_seqs, n_chunks, chunk_len;
	int64_t total_bins;
	int64_t bytes_obs, bytes_forward, bytes_transfer, bytes_total;
	/* device time of the kernels of the LAST run, milliseconds (CUDA events on the context's stream):
	 * [0] transfer-matrix kernel  [1] boundary-chain kernel  [2] forward kernel
	 * [3] backward+counts kernel  [4] reduction kernel        [5] whole E-step (first launch .. last)
	 * fast path: [2] / [3] include the repair rounds of their direction; [6] / [7] the forward / backward chunk kernel alone */
	float ms[8];
	int32_t launches; /* kernels launched by the last run */
	/* fast path (warm-up overlaps + boundary certificate): overlap in bins (0 = disabled), how many E-steps of this
	 * context had to be redone with the exact transfer-matrix path, and the largest boundary mismatches
	 * (Hilbert projective metric) seen by the last certificate */
	int32_t warm_len, fallbacks;
	double fwd_mismatch, bwd_mismatch;
	/* last run: boundary failures seen by the repair rounds (summed over rounds) and chunks they recomputed */
	int32_t failed_fwd, repaired_fwd, failed_bwd, repaired_bwd;
	/* psmc_b200_set_multiplicity: bins of the sequences with multiplicity > 0, and the sum of multiplicities */
	int64_t active_bins, n_seqs_effective;
	/* psmc_b200_set_batch: models per E-step (1 otherwise); the backward pass's own chunk plan */
	int32_t n_models, n_chunks_bwd, chunk_len_bwd;
	/* repair rounds the next E-step will enqueue (adapts to how deep the repairs cascade), and how many E-steps were
	 * redone on the fast path with more rounds after a failed certificate (before any exact fallback) */
	int32_t repair_rounds, warm_redos;
	/* last psmc_b200_decode_run: device ms of its E-step, of the decode kernel + run compaction, and wall ms of the call */
	float decode_ms[3];
	/* planned overlaps (mixing probe): 1 when the current chunk plans come from a probe (chunk_len is then the mean
	 * chunk length), how many plans were built from probes, the mean overlap per boundary in bins (the fixed overlap when
	 * not planned), and the boundaries the probe declared out of reach of any overlap (repaired with operators) */
	int32_t planned, probe_plans;
	float avg_overlap_fwd, avg_overlap_bwd;
	int32_t slow_fwd, slow_bwd;
} psmc_b200_info;

int  psmc_b200_version(void);
const char *psmc_b200_last_error(void);
int  psmc_b200_device_count(void);

/* Upload the sequences once (2-bit packed) and plan the chunked execution.
 * Replaces, for the whole run, the per-iteration per-sequence hmm_new_data copies
 * (em.c:42-44, khmm.c:37-45) and the (L+1)-row calloc of f and b (khmm.c:158-159,224).
 *   seqs[i] points to L[i] symbols (values 0/1/2, psmc_seq_t::seq, psmc.h:22-26).
 *   device: CUDA ordinal.  chunk_len: bins per chunk, 0 = choose automatically.
 * Sequences are immutable afterwards (as in the reference after psmc_parse_cli/psmc_resamp). */
int  psmc_b200_create(psmc_b200_ctx **out, int32_t n_seqs, const int32_t *L, const signed char *const *seqs,
                      int32_t n_states, int32_t device, int32_t chunk_len, uint32_t flags);
/* Same, with all sequences concatenated in one buffer (sum of L bytes). */
int  psmc_b200_create_cat(psmc_b200_ctx **out, int32_t n_seqs, const int32_t *L, const signed char *seqs_cat,
                          int32_t n_states, int32_t device, int32_t chunk_len, uint32_t flags);
void psmc_b200_destroy(psmc_b200_ctx *ctx);
/* Re-upload the observation tracks (same number of sequences and the same lengths as at create) from
 * host memory into the existing device buffers: pack to 2 bits/bin and one host->device copy.  This is
 * the per-call input transfer of a psmc_em-style call that owns only host buffers (em.c:42-44). */
int  psmc_b200_upload(psmc_b200_ctx *ctx, int32_t n_seqs, const int32_t *L, const signed char *const *seqs);
int  psmc_b200_upload_cat(psmc_b200_ctx *ctx, int32_t n_seqs, const int32_t *L, const signed char *seqs_cat);

/* Bootstrap replicates without re-uploading anything.  psmc_resamp (aux.c:8-47) draws WHOLE records with
 * replacement, so a replicate is a multiset of the resident sequences: mult[i] (i indexes the n_seqs records given
 * to create) says how often record i occurs.  The following E-steps return sum_i mult[i] * (LL_i, counts_i) - exactly
 * what em.c:33-55 computes on the resampled copy (khmm.c:346-359 adds the per-record counts), including one HMM_TINY
 * term per drawn record - and records with mult[i] == 0 are skipped: the chunk plans are rebuilt over the drawn
 * records only (host-side planning + a few hundred KB of tables; observations and work buffers stay in place).
 * mult == NULL restores multiplicity 1 for every record.  psmc_b200_decode refuses records with multiplicity 0. */
int  psmc_b200_set_multiplicity(psmc_b200_ctx *ctx, const int32_t *mult);

/* Batch mode: n_rep independent EM runs (bootstrap replicates, aux.c:8-47 / README:57-62) share ONE launch sequence.
 * Model r runs over the multiset mult[r*n_seqs + i] of the resident records (NULL = every record once in every model).
 * The drawn records of all models become the work items of the chunk plans: with a few hundred records per launch
 * the chunks are long, so the warm-up overlaps, the certificate and the repairs that a single genome's 22 contigs
 * need shrink to a few percent of the work.  The forward spill is reallocated to hold the
 * drawn records of all models (psmc_b200_mem_info tells what fits); on failure the context stays usable in single mode.
 * psmc_b200_estep_batch* take/return n_rep models/statistics; model r's statistics include the HMM_TINY terms of ITS
 * drawn records.  psmc_b200_set_multiplicity (or set_batch with n_rep = 1) returns to single mode.  No decode and no
 * dense counts in batch mode.  A model's result does not depend on what else is in the batch when chunk_len is fixed
 * (same bits); with the automatic plan it agrees to the certificate's tolerance. */
int  psmc_b200_set_batch(psmc_b200_ctx *ctx, int32_t n_rep, const int32_t *mult);
int  psmc_b200_estep_batch(psmc_b200_ctx *ctx, int32_t n_rep, const psmc_b200_model *models, psmc_b200_stats *outs);
int  psmc_b200_estep_batch_launch(psmc_b200_ctx *ctx, int32_t n_rep, const psmc_b200_model *models);
int  psmc_b200_estep_batch_finish(psmc_b200_ctx *ctx, int32_t n_rep, psmc_b200_stats *outs);
/* free / total device memory in bytes (cudaMemGetInfo), to size a batch: a model needs 8*(NP+1) bytes per drawn bin */
int  psmc_b200_mem_info(int32_t device, int64_t *free_bytes, int64_t *total_bytes);

/* One E-step: forward, backward, log-likelihood and expected counts over every sequence.
 * Replaces em.c:33-55 (hmm_pre_backward + the loop over hmm_forward / hmm_backward / hmm_lk /
 * hmm_expect / hmm_add_expect).  Host buffers in, host buffers out, synchronous. */
int  psmc_b200_estep(psmc_b200_ctx *ctx, const psmc_b200_model *model, psmc_b200_stats *out);

/* Same with the reference's dense hmm_par_t contents: a is N*N row-major (hmm_par_t::a), e is 2*N.
 * The factors are extracted and verified (relative tolerance tol, e.g. 1e-9); a matrix without the
 * PSMC structure (e.g. capped by psmc_cap_matrix, aux.c:115-127) is refused with PSMC_B200_ESTRUCT. */
int  psmc_b200_estep_dense(psmc_b200_ctx *ctx, int32_t n_states, const double *a0, const double *a,
                           const double *e, double tol, psmc_b200_stats *out);
int  psmc_b200_factorize(int32_t n_states, const double *a, double tol,
                         double *U, double *V, double *W, double *Z, double *D);

/* Asynchronous halves for multi-GPU drivers (one context per GPU, sequences sharded by the caller):
 * launch() enqueues the whole E-step on the context's stream and leaves the raw statistics vector
 *   [ LL | E0(N) E1(N) | RL(N) CL(N) RU(N) CU(N) AD(N) ]   (7*N+1 doubles, WITHOUT the TINY terms)
 * in device memory; device_stats() returns that device pointer (for an NCCL all-reduce);
 * finish() copies it back, adds n_seqs_total*HMM_TINY terms and unpacks it.
 * wait() blocks until the stream is idle. */
int  psmc_b200_estep_launch(psmc_b200_ctx *ctx, const psmc_b200_model *model);
void *psmc_b200_device_stats(psmc_b200_ctx *ctx);
int  psmc_b200_stats_len(const psmc_b200_ctx *ctx);
void *psmc_b200_stream(psmc_b200_ctx *ctx);
int  psmc_b200_wait(psmc_b200_ctx *ctx);
int  psmc_b200_estep_finish(psmc_b200_ctx *ctx, int64_t n_seqs_total, psmc_b200_stats *out);
/* wait for launch() and copy the raw 7*N+1 statistics vector to the host (for a host-side sum over GPUs) */
int  psmc_b200_estep_fetch_raw(psmc_b200_ctx *ctx, double *raw);
/* unpack a raw statistics vector that already lives on the host (e.g. after an all-reduce) */
int  psmc_b200_unpack_stats(int32_t n_states, const double *raw, int64_t n_seqs_total, psmc_b200_stats *out);

/* Posterior decoding of one sequence.  Replaces aux.c:157-158 (hmm_forward/hmm_backward) plus
 * hmm_post_decode (khmm.c:264-282; aux.c:167-182) and, when post/p_recomb are non-NULL,
 * hmm_post_state (khmm.c:286-293) and the recombination probability of aux.c:188-193.
 *   seq_id indexes the n_seqs records as given to create (empty records included; decoding one is EINVAL),
 *   best_k[L]: argmax_k f*b*s (first maximum wins), best_p[L]: its posterior,
 *   model may be NULL to reuse the forward pass of the previous estep/decode call on this context.
 *   post[L*N] (optional), p_recomb[L] (optional, 0 at the last bin), s_out[L] (optional; hmm_data_t::s, aux.c:159-164). */
int  psmc_b200_decode(psmc_b200_ctx *ctx, const psmc_b200_model *model, int32_t seq_id,
                      int32_t *best_k, double *best_p, double *post, double *p_recomb, double *s_out);

/* Decoding of EVERY sequence of the context in one go, on the fast path (what `psmc -d` / `-D` need at genome scale;
 * replaces aux.c:157-200 for all sequences).  decode_run = one complete E-step on the model (exact forward spill,
 * certified boundary directions) + one decode kernel over all chunks; results stay on the device until fetched:
 *   PSMC_B200_DEC_RUNS  runs of the posterior-argmax state with their maximum posterior -- the DC lines of aux.c:165-182 --
 *                       compacted on the device, so only the runs cross PCIe (~13 bytes per run instead of 12 per bin);
 *                       get_runs returns them for all sequences in (sequence, position) order: seq_id indexes the records
 *                       as given to create, start is the 0-based first bin, state = argmax_k f*b*s (first maximum wins),
 *                       max_p = the largest posterior of that state inside the run.  Call with cap = 0 to get *n_runs only.
 *   PSMC_B200_DEC_BINS  per bin: argmax state (uint8) and its posterior (float)            -> get_bins(best_k, best_p)
 *   PSMC_B200_DEC_POST  per bin: the full posterior row (float, hmm_post_state khmm.c:286-293) and the recombination
 *                       probability (double, aux.c:188-193; 0 at the last bin)             -> get_bins(post, p_recomb)
 * Outputs are reduced precision by design (uint8 / float): they feed %.3lf / %.4f text.  psmc_b200_decode above keeps the
 * double-precision, one-sequence interface.  Not available in batch mode. */
#define PSMC_B200_DEC_RUNS 1u
#define PSMC_B200_DEC_BINS 2u
#define PSMC_B200_DEC_POST 4u
int  psmc_b200_decode_run(psmc_b200_ctx *ctx, const psmc_b200_model *model, uint32_t what);
int  psmc_b200_decode_get_runs(psmc_b200_ctx *ctx, int64_t cap, int32_t *seq_id, int32_t *start, int32_t *len, uint8_t *state,
                               double *max_p, int64_t *n_runs);
int  psmc_b200_decode_get_bins(psmc_b200_ctx *ctx, int32_t seq_id, uint8_t *best_k, float *best_p, float *post, double *p_recomb);

/* Fast-path control: warm_len = bins of warm-up overlap per chunk (0 = always use the exact transfer-matrix
 * path; < 0 = keep), eps = certificate tolerance in Hilbert's projective metric (<= 0 = keep; default 1e-12).
 * Boundaries the overlap does not reach are repaired locally; if the final certificate still fails, that
 * E-step is silently redone with the transfer-matrix path (counted in psmc_b200_info::fallbacks). */
int  psmc_b200_set_warm(psmc_b200_ctx *ctx, int32_t warm_len, double eps);

/* Dense transition counts (option): hmm_expect's A[N][N] (khmm.c:305-316 summed over sequences as hmm_add_expect does,
 * khmm.c:346-352).  Nothing on the EM path needs them -- the M-step works on the five O(N) marginals -- except the
 * constant offset hmm_Q0 (khmm.c:336-340) of the printed QD line.  psmc_b200_set_dense(ctx, 1) makes every following
 * E-step also spill the backward rows (8N more bytes per bin: allocated here); psmc_b200_dense_counts then forms the
 * N x N matrix (row-major, caller-owned) of the LAST E-step.  At most 64 states. */
int  psmc_b200_set_dense(psmc_b200_ctx *ctx, int32_t on);
int  psmc_b200_dense_counts(psmc_b200_ctx *ctx, double *A);

int  psmc_b200_get_info(const psmc_b200_ctx *ctx, psmc_b200_info *info);

#ifdef __cplusplus
}
#endif
#endif
