// psmc_estep.cu -- B200 (sm_100a) E-step of the PSMC HMM behind the C ABI of include/psmc_b200.h.
//
// What this replaces (reference lh3/psmc): the per-iteration loop em.c:33-55 over
// hmm_forward (khmm.c:145-190), hmm_backward (khmm.c:210-241), hmm_lk (khmm.c:245-260),
// hmm_expect (khmm.c:297-324) and hmm_add_expect (khmm.c:346-359), and the forward/backward/posterior
// part of psmc_decode (aux.c:150-201).  Nothing here is translated from khmm.c: the reference is a
// dense O(N^2)-per-bin single-thread loop that materialises f and b; this is an O(N)-per-bin,
// chunk-parallel, exact formulation for the GPU.
//
// Algorithm (DESIGN.md has the derivation)
//   The PSMC transition matrix is diagonal + rank-1 strictly-lower + rank-1 strictly-upper
//   (core.c:100-123):  a[k][l] = U_k V_l (l<k), W_k Z_l (l>k), D_k (l=k).  One forward step is
//       f'[l] = e_x[l] * ( D_l f[l] + Z_l * sum_{k<l} W_k f[k] + V_l * sum_{k>l} U_k f[k] )
//   i.e. two exclusive scans; one backward step is the transposed pattern.  The states of one
//   sequence position live in the registers of a lane group (G lanes x SPL states per lane) and the
//   scans are warp shuffles.
//   The chain over bins is serial, so every sequence is cut into chunks and made exact again:
//     K1 transfer : per chunk, the N x N transfer operator T_c = prod_u diag(e_{x_u}) A^T, one column
//                   per lane group, per-column power-of-two scaling (exact).
//     K2 chain    : per sequence, v_{c+1} = normalise(T_c v_c) left to right (exact forward vector at
//                   every chunk start) and beta_c = T_{c+1}^T beta_{c+1} right to left (direction of
//                   the backward vector at every chunk end).
//     K3 forward  : one warp per chunk; scaled forward from the exact start; writes f_u (N doubles)
//                   and the scale s_u per bin to HBM; accumulates log-likelihood.
//     K4 backward : one warp per chunk; scaled backward from the exact end, re-reading f_u; accumulates
//                   the emission counts E[x][k] and the five O(N) marginals of the transition counts
//                   (RL, CL, RU, CU, AD) in registers; per-warp partials.
//     K5 reduce   : fixed-order tree over the partials -> one statistics vector in device memory
//                   (deterministic; this is the buffer an NCCL all-reduce would sum across GPUs).
//   Scaling follows the reference exactly (f normalised per bin, b_u divided by s_u, khmm.c:183-184,
//   226, 233), so gamma_u = f_u b_u s_u and xi_u = f_u a e b_{u+1} carry no extra factors.
//   Fast path (default): instead of K1/K2 for every chunk, every chunk warms up over an overlap with its
//   neighbour, every boundary is certified afterwards, and only the boundaries that fail are repaired with
//   transfer operators (launch_warm below).
//
// Files: psmc_kernels.cuh = all device code (both kernel generations); this file = the C ABI of
// include/psmc_b200.h, the chunk plans and the launch sequences.  tests/emu compiles both for the host.
//
// All arithmetic is FP64.  No tensor cores: there is no dense contraction on this path.

#ifndef PSMC_SIMT_EMU
#include <cuda_runtime.h>
#endif
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdarg.h>
#include <math.h>
#include <limits.h>
#if defined(__linux__)
#include <sched.h>
#endif
#include <vector>
#include <algorithm>
#include <thread>
#include <chrono>

#include "psmc_b200.h"

#define FULLMASK 0xffffffffu

// Every kernel launch goes through LAUNCH so that tests/emu can compile this very file for the host (a SIMT emulation
// used by the CPU test-suite only: it runs the kernels' source with one host thread per CUDA thread; never shipped).
#define PSMC_UNPAREN(...) __VA_ARGS__
#ifndef PSMC_SIMT_EMU
#define LAUNCH(kern, grid, block, stream, ...)                        \
	do {                                                              \
		auto kfn_ = PSMC_UNPAREN kern;                                \
		kfn_<<<(grid), (block), 0, (stream)>>>(__VA_ARGS__);          \
	} while (0)
#define LAUNCH_SMEM(kern, grid, block, smem, stream, ...)              \
	do {                                                              \
		auto kfn_ = PSMC_UNPAREN kern;                                \
		kfn_<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);     \
	} while (0)
#define PIN_REG(x) asm volatile("" : "+d"(x))
#define PIN_PTR(x) asm volatile("" : "+l"(x))
#define PIN_INT(x) asm volatile("" : "+r"(x))
#else
#define LAUNCH_SMEM(kern, grid, block, smem, stream, ...)             \
	do {                                                              \
		auto kfn_ = PSMC_UNPAREN kern;                                \
		simt_emu::launch((grid), (block), [&]() { kfn_(__VA_ARGS__); }, (smem)); \
	} while (0)
#define LAUNCH(kern, grid, block, stream, ...)                        \
	do {                                                              \
		auto kfn_ = PSMC_UNPAREN kern;                                \
		simt_emu::launch((grid), (block), [&]() { kfn_(__VA_ARGS__); }); \
	} while (0)
#define PIN_REG(x) ((void)(x))
#define PIN_PTR(x) ((void)(x))
#define PIN_INT(x) ((void)(x))
#endif
#define HMM_TINY_ 1e-25 /* khmm.h:28 */

// ------------------------------------------------------------------------------------------------
// error plumbing
// ------------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
static int set_err(int code, const char *fmt, ...)
{
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(g_err, sizeof(g_err), fmt, ap);
	va_end(ap);
	return code;
}
#define CUDA_TRY(call, code)                                                                     \
	do {                                                                                         \
		cudaError_t e_ = (call);                                                                 \
		if (e_ != cudaSuccess)                                                                   \
			return set_err(code, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
	} while (0)

#include "psmc_kernels.cuh"

// ================================================================================================
// host side
// ================================================================================================
struct psmc_b200_ctx {
	int device = 0;
	int N = 0, NP = 0, SPL = 0;
	int n_seqs = 0, n_chunks = 0, chunk_len = 0, n_k1 = 0;
	int64_t total_bins = 0;
	std::vector<int32_t> L;        // per kept sequence
	std::vector<int32_t> seq_c0, seq_nc;
	std::vector<int64_t> seq_gb0;
	std::vector<Chunk> chunks;
	cudaStream_t stream = nullptr, stream2 = nullptr;
	cudaEvent_t ev[8] = {}, ev_fork = nullptr, ev_join = nullptr;
	// device buffers
	uint32_t *d_obs = nullptr;
	Chunk *d_chunks = nullptr;
	int32_t *d_k1 = nullptr, *d_seq_c0 = nullptr, *d_seq_nc = nullptr, *d_Tex = nullptr;
	double *d_model = nullptr, *d_fhat = nullptr, *d_sc = nullptr, *d_T = nullptr, *d_vstart = nullptr, *d_bend = nullptr;
	double *d_part = nullptr, *d_llpart = nullptr, *d_stats = nullptr;
	double *d_fwarm = nullptr, *d_bwarm = nullptr, *d_bexact = nullptr;
	double *d_bsave[2] = {nullptr, nullptr}; // backward warm-start directions, double-buffered over E-steps
	int bsave_cur = 0;         // index of the buffer the NEXT E-step reads
	bool have_prev = false;    // fhat / d_bsave[bsave_cur] hold a previous E-step of the same data
	int warm_hot = 0;          // overlap when warm-started from the previous E-step (PSMC_B200_WARM_HOT; 0 = always cold: early EM iterations move the model too much for a short overlap to reach the certificate) // warm-up mode: boundary vectors for the certificate
	// sub-chunk tables of the repair rounds
	int n_sub = 0, sub_len = 1536;
	Chunk *d_sub = nullptr;
	int32_t *d_sub_parent = nullptr, *d_chunk_sub0 = nullptr, *d_Texsub = nullptr;
	double *d_Tsub = nullptr, *d_vsub = nullptr, *d_bsub = nullptr, *d_llsub = nullptr, *d_partsub = nullptr;
	int32_t *d_flag = nullptr; // n_chunks + 2 boundary flags of the current repair round (entries -1 and n_chunks are always 0)
	// The backward pass has its OWN chunk plan in the fast path: its kernel needs about twice the registers of the
	// forward kernel, so one resident wave holds half as many chunks; tying both passes to one plan would make the
	// forward chunks twice as long as necessary.  (Transfer mode and decode use the forward plan for both directions.)
	int n_chunks_b = 0, chunk_len_b = 0, n_sub_b = 0;
	Chunk *d_chunks_b = nullptr, *d_sub_b = nullptr;
	int32_t *d_sub_parent_b = nullptr, *d_chunk_sub0_b = nullptr, *d_flag_b = nullptr;
	unsigned long long *d_cert = nullptr, *h_cert = nullptr;            // [failed boundaries, max fwd mismatch bits, max bwd mismatch bits]
	int warm_len = 0;          // bins of forward warm-up overlap (0 = always use the transfer-matrix path)
	bool warm_bwd_fixed = false; // PSMC_B200_WARM_BWD given: no adaptive cap
	int warm_len_bwd = 0;      // bins of backward warm-up overlap (runs concurrently with the forward kernel, so it can be longer)
	double cert_eps = 1e-12;
	bool mode_warm = false, certified = true;
	int fallbacks = 0, repair_rounds = 3;
	// Repairs cascade: a round fixes the runs of boundaries that failed at its start, which can move the boundary BEHIND each
	// run (its left vector changed).  With chunks much shorter than the overlap (small inputs, many GPUs) a slow tract
	// needs many rounds.  rounds_cur adapts: two more than the deepest round of the last E-step that still saw a failure; if
	// the certificate fails, the E-step is first redone on the fast path with rounds_max rounds, then with exact operators.
	int rounds_cur = 3, rounds_max = 48, warm_redos = 0; // (rounds_max <= 64: one device counter per round and direction)
	bool rounds_fixed = false; // PSMC_B200_REPAIR_ROUNDS given
	bool redoing = false;
	int slots_fwd = 0, slots_bwd = 0; // resident chunks per SM of the chosen forward / backward kernels
	int g_bww = 8;              // lanes per chunk in the backward warm-up kernel (PSMC_B200_G_BWW)
	bool g2_fwd_auto = true;    // small shards (multi-GPU runs, single short contigs): fewer than g2_fwd_auto_max chunks per SM leave schedulers
	int g2_fwd_auto_max = 8;    // idle with 8-lane groups; 16-lane groups then give every chunk twice the lanes (measured on B200: -9 % forward at 3.6 M bins, +15 % at 7 M)
	// (never with a fixed chunk length, multiplicities or a batch: there a replicate's result does not depend on how it was scheduled, bit for bit)
	bool wide_fwd(int NP_) const { return gen == 2 && (g2_fwd == 16 || NP_ > 64 || (g2_fwd_auto && chunk_len_req <= 0 && !batch && !weighted && n_chunks <= g2_fwd_auto_max * sm_count)); }
	int g2_fwd = 8, g2_bww = 8; // generation 2: lanes per chunk of the forward / backward warm-up kernels at NP <= 64 (PSMC_B200_G2_FWD / _BWW: 8 or 16)
	bool dense = false;         // psmc_b200_set_dense: the backward pass also stores the rows g_u for psmc_b200_dense_counts
	bool dense_valid = false;   // ghat holds the rows of the last E-step
	double *d_ghat = nullptr, *d_cpart = nullptr, *d_cdense = nullptr;
	int cap_cpart = 0;
	bool staged_bwd = true;     // backward pass reads the forward spill through bulk-asynchronous copies into shared memory (PSMC_B200_TMA=0: register prefetch ring)
	int gen = 2;                // kernel generation (PSMC_B200_GEN=1: the Kogge-Stone kernels)
	int g_fwd = 16, g_bwd = 32; // lanes per chunk in the forward / backward kernels (PSMC_B200_G_FWD / PSMC_B200_G_BWD: 8, 16 or 32)
	long long rep_fwd_fail = 0, rep_fwd_chunks = 0, rep_bwd_fail = 0, rep_bwd_chunks = 0; // of the last run
	double mis_f = 0.0, mis_b = 0.0;
	// decode scratch (allocated on demand)
	int32_t *d_bestk = nullptr;
	double *d_bestp = nullptr, *d_post = nullptr, *d_prec = nullptr;
	int64_t dec_cap = 0, dec_post_cap = 0;
	// whole-context decoding on the fast path (psmc_b200_decode_run): per-bin outputs and run lists, indexed by forward-spill row
	uint8_t *d2_bestk = nullptr, *d2_run_state = nullptr, *d2_o_state = nullptr;
	float *d2_bestp = nullptr, *d2_post = nullptr;
	double *d2_prec = nullptr, *d2_run_maxp = nullptr, *d2_o_maxp = nullptr;
	int32_t *d2_run_start = nullptr, *d2_run_count = nullptr, *d2_o_seq = nullptr, *d2_o_start = nullptr;
	int64_t *d2_off = nullptr;
	int64_t d2_rows = 0, d2_rows_post = 0, d2_rows_runs = 0, d2_cap_o = 0;
	int d2_cap_chunks = 0;
	uint32_t dec_what = 0;       // what the last psmc_b200_decode_run produced
	float dec_ms[3] = {};        // E-step, decode kernel (+ run compaction), whole call (wall)
	std::vector<int32_t> r_seq, r_start, r_len; // merged runs of the last decode_run, in (sequence, position) order
	std::vector<uint8_t> r_state;
	std::vector<double> r_maxp;
	// pinned host staging
	double *h_model = nullptr, *h_stats = nullptr;
	uint32_t *h_obs = nullptr; // packed observations (pinned), words_obs 32-bit words
	int64_t words_obs = 0;
	std::vector<int64_t> seq_ow0;
	int64_t bytes_obs = 0, bytes_forward = 0, bytes_transfer = 0, bytes_total = 0;
	float ms[8] = {};
	int launches = 0;
	bool launched = false;
	bool fwd_valid = false; // fhat/sc/bend hold a complete forward pass + boundary chains
	// multiplicities (bootstrap replicates): the chunk plans cover the sequences with mult > 0 only
	int n_seqs_given = 0;            // as passed to create (including empty records)
	std::vector<int32_t> kept_of;    // given index -> kept index or -1
	std::vector<int32_t> mult;       // per kept sequence
	bool weighted = false;           // some multiplicity differs from 1
	int64_t n_seq_eff = 0;           // sum of multiplicities (the HMM_TINY terms, khmm.c:305-308)
	int64_t active_bins = 0;
	int chunk_len_req = 0;           // chunk length the caller / environment asked for (0 = one resident wave)
	int sm_count = 0;
	double *d_cw = nullptr, *d_cw_b = nullptr; // per-chunk multiplicity, forward / backward plan
	// Batch mode (psmc_b200_set_batch): n_rep models side by side in ONE launch sequence -- bootstrap replicates, each a
	// multiset of the resident records under its own model.  A "virtual sequence" is a (model, record) pair; the chunk
	// plans cover the virtual sequences, chunks carry their model index, k_reduce sums per model.
	// Mixing probe and planned overlaps (k_probe; DESIGN.md "Planned overlaps"): every now and then an E-step also measures
	// the local contraction rate of the chain; the chunk plans are then rebuilt with, per boundary, the overlap it needs
	// (plus a margin), chunk lengths balanced so that overlap + length is the same for every chunk, and the boundaries no
	// admissible overlap reaches known in advance (their operators are "predicted" from the first E-step of a plan on).
	bool probe_on = true;        // PSMC_B200_PROBE=0: fixed overlaps, uniform chunks (the round-1 plans)
	bool probe_batch = false;    // PSMC_B200_PROBE_BATCH=1: also in batch mode (measured on B200: no gain there -- chunks are ~100 k bins, the
	                             // overlaps a few percent of the work, and a probe costs a third of a batch E-step)
	int estep_no = 0;            // E-steps launched on the current sequences
	bool probe_now = false;      // the E-step in flight also runs k_probe
	bool plan_dirty = false;     // a probe result is waiting: re-plan before the next E-step
	bool have_probe = false, planned = false;
	float *d_probeK = nullptr, *h_probeK = nullptr;
	int64_t cap_probeK = 0, n_probeK = 0;
	std::vector<std::vector<double>> probe_cum; // per kept sequence: cum[j] = log contraction accumulated over the bins < 64 j (non-increasing); empty = not probed yet
	std::vector<int32_t> vseq_seq;              // virtual sequence -> kept sequence (batch mode: a record appears once per model that drew it)
	double probe_th = 35.0;      // e-folds an overlap must cover: ln(1 / 1e-12) = 27.6 + margin (start distance, probe accuracy); swept on B200
	int probe_hmax = 0, probe_hslow = 0, probe_lead = 0, probe_plans = 0; // 0 = derived from the fixed overlap: 5/3, 1/6 and 1/12 of it (20480 / 2048 / 1024 bins)
	int hmax() const { return probe_hmax > 0 ? probe_hmax : warm_len + 2 * warm_len / 3; }
	int hslow() const { return probe_hslow > 0 ? probe_hslow : std::max(64, warm_len / 6); }
	int lead() const { return probe_lead > 0 ? probe_lead : std::max(32, warm_len / 12); }
	int32_t *d_warm_f = nullptr, *d_warm_b = nullptr; // per chunk: overlap in front of (forward plan) / behind (backward plan) it
	double avg_warm_f = 0.0, avg_warm_b = 0.0;
	int slow_f = 0, slow_b = 0;
	uint64_t obs_hash = 0;
	int n_rep = 1;
	bool batch = false;
	std::vector<int32_t> bmult;          // batch mode: n_rep x n_seqs multiplicities (kept-record index)
	std::vector<int64_t> rep_seq_eff;    // per model: sum of multiplicities (HMM_TINY terms)
	std::vector<int32_t> rep_c0, rep_c0_b; // per model: first chunk in the forward / backward plan (n_rep + 1 entries)
	int32_t *d_rep_c0 = nullptr, *d_rep_c0_b = nullptr;
	int n_vseq = 0, cap_vseq = 0, cap_rep = 0;
	int64_t cap_bins = 0;                // rows allocated in fhat / sc (/ ghat)
	int cap_models = 0;                  // models d_model / h_model / d_stats / h_stats can hold
	// Slow-mixing boundaries are a property of the data (het-poor, low-TMRCA tracts): the chunks that needed a repair in
	// one E-step almost always need it in the next.  Their transfer operators depend on model + observations only, so they
	// are computed AHEAD of time on the side stream, hidden behind the forward kernel (d_pred*: round-1 failures of the
	// previous E-step, double-buffered; the backward plan has its own operator buffer so both can be in flight).
	int32_t *d_pred[2] = {nullptr, nullptr}, *d_pred_b[2] = {nullptr, nullptr};
	int pred_cur = 0;
	double *d_Tsub_b = nullptr;
	int32_t *d_Texsub_b = nullptr;
	cudaEvent_t ev_k1f = nullptr, ev_k1b = nullptr;
	bool predict = true; // PSMC_B200_PREDICT=0 disables
	int cap_chunks = 0, cap_chunks_b = 0, cap_sub = 0, cap_sub_b = 0, cap_k1 = 0; // allocated capacities of the plan buffers
	int64_t bytes_plan = 0;
	int replans = 0;
};

template <int NP>
static void chunk_slots(const psmc_b200_ctx *c, int *slots_fwd, int *slots_bwd);

extern "C" int psmc_b200_version(void) { return PSMC_B200_VERSION; }
extern "C" const char *psmc_b200_last_error(void) { return g_err; }
extern "C" int psmc_b200_device_count(void)
{
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
	return n;
}

static int pad_states(int N)
{
	if (N <= 32) return 32;
	if (N <= 64) return 64;
	return 128;
}

static void free_plan(psmc_b200_ctx *c)
{
	void **p[] = {(void **)&c->d_chunks, (void **)&c->d_k1, (void **)&c->d_Tex, (void **)&c->d_T, (void **)&c->d_vstart, (void **)&c->d_bend,
	              (void **)&c->d_part, (void **)&c->d_llpart, (void **)&c->d_fwarm, (void **)&c->d_bwarm, (void **)&c->d_bexact,
	              (void **)&c->d_bsave[0], (void **)&c->d_bsave[1], (void **)&c->d_chunks_b, (void **)&c->d_sub_b, (void **)&c->d_sub_parent_b,
	              (void **)&c->d_chunk_sub0_b, (void **)&c->d_flag_b, (void **)&c->d_flag, (void **)&c->d_sub, (void **)&c->d_sub_parent,
	              (void **)&c->d_chunk_sub0, (void **)&c->d_Tsub, (void **)&c->d_Texsub, (void **)&c->d_vsub, (void **)&c->d_bsub,
	              (void **)&c->d_llsub, (void **)&c->d_partsub, (void **)&c->d_cw, (void **)&c->d_cw_b,
	              (void **)&c->d_pred[0], (void **)&c->d_pred[1], (void **)&c->d_pred_b[0], (void **)&c->d_pred_b[1],
	              (void **)&c->d_Tsub_b, (void **)&c->d_Texsub_b, (void **)&c->d_seq_c0, (void **)&c->d_seq_nc, (void **)&c->d_rep_c0, (void **)&c->d_rep_c0_b,
	              (void **)&c->d_warm_f, (void **)&c->d_warm_b};
	for (auto q : p) { cudaFree(*q); *q = nullptr; }
	c->bytes_total -= c->bytes_plan;
	c->bytes_plan = 0;
	c->cap_chunks = c->cap_chunks_b = c->cap_sub = c->cap_sub_b = c->cap_k1 = c->cap_vseq = c->cap_rep = 0;
}

static void free_ctx(psmc_b200_ctx *c)
{
	if (!c) return;
	cudaSetDevice(c->device);
	free_plan(c);
	cudaFree(c->d_obs); cudaFree(c->d_probeK);
	if (c->h_probeK) cudaFreeHost(c->h_probeK);
	cudaFree(c->d_model); cudaFree(c->d_fhat); cudaFree(c->d_sc); cudaFree(c->d_cert);
	if (c->h_cert) cudaFreeHost(c->h_cert);
	cudaFree(c->d_stats);
	cudaFree(c->d_ghat); cudaFree(c->d_cpart); cudaFree(c->d_cdense);
	cudaFree(c->d_bestk); cudaFree(c->d_bestp); cudaFree(c->d_post); cudaFree(c->d_prec);
	cudaFree(c->d2_bestk); cudaFree(c->d2_run_state); cudaFree(c->d2_o_state); cudaFree(c->d2_bestp); cudaFree(c->d2_post);
	cudaFree(c->d2_prec); cudaFree(c->d2_run_maxp); cudaFree(c->d2_o_maxp); cudaFree(c->d2_run_start); cudaFree(c->d2_run_count);
	cudaFree(c->d2_o_seq); cudaFree(c->d2_o_start); cudaFree(c->d2_off);
	if (c->h_model) cudaFreeHost(c->h_model);
	if (c->h_stats) cudaFreeHost(c->h_stats);
	if (c->h_obs) cudaFreeHost(c->h_obs);
	for (int i = 0; i < 8; ++i)
		if (c->ev[i]) cudaEventDestroy(c->ev[i]);
	if (c->ev_fork) cudaEventDestroy(c->ev_fork);
	if (c->ev_join) cudaEventDestroy(c->ev_join);
	if (c->ev_k1f) cudaEventDestroy(c->ev_k1f);
	if (c->ev_k1b) cudaEventDestroy(c->ev_k1b);
	if (c->stream2) cudaStreamDestroy(c->stream2);
	if (c->stream) cudaStreamDestroy(c->stream);
	delete c;
}

extern "C" void psmc_b200_destroy(psmc_b200_ctx *ctx) { free_ctx(ctx); }

// 2 bits per bin, 16 bins per word, little end first; every sequence starts on a 128-byte boundary.
// Symbols: 0 hom, 1 het, everything else missing (cli.c:15-32 maps to {0,1,2}).
static inline uint32_t squeeze8(uint64_t x) // eight bytes holding 0..2 -> sixteen bits, byte k -> bits 2k..2k+1
{
	x = (x | (x >> 6)) & 0x000F000F000F000Full;
	x = (x | (x >> 12)) & 0x000000FF000000FFull;
	return (uint32_t)((x | (x >> 24)) & 0xFFFFull);
}
static void pack_range(const signed char *s, int64_t u0, int64_t u1, uint32_t *dst)
{
	int64_t w = u0 >> 4;
	const int64_t w_full = u1 >> 4; // words [w, w_full) are complete
	for (; w < w_full; ++w) {
		uint8_t t[16];
		const uint8_t *src = (const uint8_t *)s + (w << 4);
		for (int i = 0; i < 16; ++i) t[i] = src[i] > 1u ? 2u : src[i]; // vectorised by the host compiler (pminub)
		uint64_t lo, hi;
		memcpy(&lo, t, 8);
		memcpy(&hi, t + 8, 8);
		dst[w] = squeeze8(lo) | (squeeze8(hi) << 16);
	}
	if (w < (u1 + 15) >> 4) { // ragged last word, padding = missing
		const int64_t b0 = w << 4;
		uint32_t v = 0xaaaaaaaau;
		for (int64_t u = b0; u < u1; ++u) {
			const uint32_t x = (uint8_t)s[u];
			const int sh = (int)(u - b0) * 2;
			v = (v & ~(3u << sh)) | ((x > 1u ? 2u : x) << sh);
		}
		dst[w] = v;
	}
}
static uint64_t hash_obs(const psmc_b200_ctx *c)
{
	uint64_t h = 1469598103934665603ull; // word-wise multiply-xor over the packed tracks (~1 ms for a 3 Gbp genome)
	const uint64_t *p = (const uint64_t *)c->h_obs;
	const int64_t n = c->words_obs / 2;
	uint64_t a = h, b = h ^ 0x9e3779b97f4a7c15ull, d = h + 7, e = h + 13;
	int64_t i = 0;
	for (; i + 4 <= n; i += 4) {
		a = (a ^ p[i]) * 1099511628211ull; b = (b ^ p[i + 1]) * 1099511628211ull;
		d = (d ^ p[i + 2]) * 1099511628211ull; e = (e ^ p[i + 3]) * 1099511628211ull;
	}
	for (; i < n; ++i) a = (a ^ p[i]) * 1099511628211ull;
	return a ^ (b * 3) ^ (d * 5) ^ (e * 7);
}

static void pack_all(psmc_b200_ctx *c, const signed char *const *sp)
{
	// only the alignment padding behind every sequence needs the 'missing' fill; the rest is overwritten below
	for (int i = 0; i < c->n_seqs; ++i) {
		const int64_t w0 = c->seq_ow0[i] + ((int64_t)c->L[i] + 15) / 16;
		const int64_t w1 = (i + 1 < c->n_seqs) ? c->seq_ow0[i + 1] : c->words_obs;
		for (int64_t w = w0; w < w1; ++w) c->h_obs[w] = 0xaaaaaaaau;
	}
	if (c->n_seqs == 0)
		for (int64_t w = 0; w < c->words_obs; ++w) c->h_obs[w] = 0xaaaaaaaau;
	struct Job { const signed char *s; int64_t u0, u1; uint32_t *dst; };
	std::vector<Job> jobs;
	const int64_t step = 1 << 20; // multiple of 16
	for (int i = 0; i < c->n_seqs; ++i)
		for (int64_t u = 0; u < c->L[i]; u += step)
			jobs.push_back({sp[i], u, std::min<int64_t>(u + step, c->L[i]), c->h_obs + c->seq_ow0[i]});
	// threads: the cores this process may use (affinity mask), shared with the other local ranks of a multi-process job
	unsigned cores = std::max<unsigned>(1, std::thread::hardware_concurrency());
#if !defined(PSMC_SIMT_EMU) && defined(__linux__)
	{
		cpu_set_t set;
		if (sched_getaffinity(0, sizeof(set), &set) == 0 && CPU_COUNT(&set) > 0) cores = std::min<unsigned>(cores, (unsigned)CPU_COUNT(&set));
	}
#endif
	{
		const char *lws = getenv("LOCAL_WORLD_SIZE");
		if (lws && atoi(lws) > 1) cores = std::max<unsigned>(1, cores / (unsigned)atoi(lws));
	}
	unsigned nt = std::min<unsigned>(16, cores);
	if (jobs.size() < 4) nt = 1;
	if (nt == 1) {
		for (auto &j : jobs) pack_range(j.s, j.u0, j.u1, j.dst);
	} else {
		std::vector<std::thread> th;
		for (unsigned t = 0; t < nt; ++t)
			th.emplace_back([&jobs, t, nt]() {
				for (size_t k = t; k < jobs.size(); k += nt) pack_range(jobs[k].s, jobs[k].u0, jobs[k].u1, jobs[k].dst);
			});
		for (auto &x : th) x.join();
	}
}

// (Re)build both chunk plans over the sequences with multiplicity > 0 and upload them.  The packed observations,
// the forward spill (indexed by bin) and everything else that does not depend on the plan stay where they are.
struct VSeq { int32_t seq, rep, mult; int64_t gb0; };

// ---- planned overlaps: what the mixing probe says about a boundary ------------------------------------------------
// cum[j] = log contraction accumulated over the bins < 64 j of one sequence (cum[0] = 0, non-increasing).
static inline double probe_at(const std::vector<double> &cum, int64_t u) // accumulated over the bins < u (linear inside a cell)
{
	const int64_t j = u >> 6, nj = (int64_t)cum.size() - 1;
	if (j >= nj) return cum[nj];
	return cum[j] + (cum[j + 1] - cum[j]) * (double)(u & 63) / 64.0;
}
// bins of overlap a forward boundary in front of bin p needs: the shortest window [p - H, p) over which th e-folds are
// lost (p itself if the whole prefix does not lose that much: the warm-up then starts at the sequence start, exactly)
static int need_fwd(const std::vector<double> &cum, int64_t p, double th)
{
	const double target = probe_at(cum, p) + th;
	if (cum[0] < target) return (int)p;
	int64_t lo = 0, hi = p >> 6; // largest j <= p/64 with cum[j] >= target
	while (lo < hi) {
		const int64_t mid = (lo + hi + 1) >> 1;
		if (cum[mid] >= target) lo = mid; else hi = mid - 1;
	}
	return (int)(p - (lo << 6));
}
// bins of overlap a backward boundary behind bin q needs: the shortest window (q, q + H] over which th e-folds are lost
static int need_bwd(const std::vector<double> &cum, int64_t q, int64_t L, double th)
{
	const double target = probe_at(cum, q + 1) - th;
	const int64_t nj = (int64_t)cum.size() - 1;
	if (cum[nj] > target) return (int)(L - 1 - q);
	int64_t lo = (q + 1 + 63) >> 6, hi = nj; // smallest j >= (q+1)/64 with cum[j] <= target
	if (lo > hi) return (int)(L - 1 - q);
	while (lo < hi) {
		const int64_t mid = (lo + hi) >> 1;
		if (cum[mid] <= target) hi = mid; else lo = mid + 1;
	}
	return (int)std::min<int64_t>(L - 1 - q, (lo << 6) - (q + 1));
}
struct Piece { int32_t u0, len, warm, slow; };
// one sequence cut into chunks with overlap + length == T (as far as a minimum length and the ends allow); dir 0: the
// overlap of a chunk lies in front of it (forward plan), dir 1: behind it (backward plan).  Returns the number of chunks.
static int cut_sequence(const std::vector<double> &cum, int L, int T, int dir, double th, int hmax, int hslow, std::vector<Piece> *out, double sw = 2.0)
{
	// sw: a stored bin costs the forward kernel about twice a warm-up bin (normalisation, stores), so the budget of a chunk is
	// overlap + sw * length = T
	const int MINLEN = 512;
	int n = 0;
	if (dir == 0) {
		int pos = 0;
		while (pos < L) {
			int H = 0, slow = 0;
			if (pos > 0) {
				const int need = need_fwd(cum, pos, th);
				if (need > hmax) { H = std::min(pos, hslow); slow = 1; }
				else H = std::min(pos, need + 128);
			}
			int len = std::max(MINLEN, (int)((T - H) / sw));
			if (slow) { // this chunk will be recomputed through transfer operators (64 columns): end it where the slow tract ends
				for (int p = pos + MINLEN; p < pos + len && p < L; p += 256)
					if (need_fwd(cum, p, th) <= hmax) { len = p - pos; break; }
			}
			if (L - pos < len + len / 2) len = L - pos;
			if (out) out->push_back({pos, len, H, slow});
			pos += len;
			++n;
		}
	} else {
		// backward plan: the counting kernel only walks the chunk itself (the overlap is the side stream's warm-up kernel), so
		// its chunks stay equally long (T = chunk length here); only the overlap behind every chunk follows the probe
		const int nc = (L + T - 1) / T;
		for (int k = 0; k < nc; ++k) {
			const int64_t a = (int64_t)L * k / nc, b = (int64_t)L * (k + 1) / nc;
			int H = 0, slow = 0;
			if (b < L) {
				const int need = need_bwd(cum, b - 1, L, th);
				if (need > hmax) { H = std::min((int)(L - b), hslow); slow = 1; }
				else H = std::min((int)(L - b), need + 128);
			}
			if (out) out->push_back({(int32_t)a, (int32_t)(b - a), H, slow});
			++n;
		}
	}
	return n;
}

static int replan(psmc_b200_ctx *c)
{
	const int NP = c->NP;
	// virtual sequences: single mode = every resident record once (records with multiplicity 0 keep their rows in the
	// forward spill and get no chunks); batch mode = the drawn records of every model, rows packed densely
	std::vector<VSeq> vs;
	c->active_bins = 0; c->n_seq_eff = 0; c->weighted = false;
	c->rep_seq_eff.assign((size_t)c->n_rep, 0);
	int n_active = 0;
	int64_t rows = 0;
	for (int r = 0; r < c->n_rep; ++r)
		for (int i = 0; i < c->n_seqs; ++i) {
			const int32_t m = c->batch ? c->bmult[(size_t)r * c->n_seqs + i] : c->mult[i];
			if (m > 0) { c->active_bins += c->L[i]; ++n_active; }
			if (m != 1) c->weighted = true;
			c->rep_seq_eff[r] += m;
			if (c->batch && m == 0) continue;
			vs.push_back({i, r, m, rows});
			rows += c->L[i];
		}
	c->n_seq_eff = c->rep_seq_eff[0];
	c->n_vseq = (int)vs.size();
	c->vseq_seq.resize(vs.size());
	for (size_t v = 0; v < vs.size(); ++v) c->vseq_seq[v] = vs[v].seq;
	c->seq_c0.assign((size_t)std::max(c->n_vseq, 1), 0); c->seq_nc.assign((size_t)std::max(c->n_vseq, 1), 0); c->seq_gb0.assign((size_t)std::max(c->n_vseq, 1), 0);
	// The chunk kernels are latency-bound and every block lives as long as the kernel, so a plan must fit in ONE
	// resident wave of its kernel (one block too many doubles the kernel time; more chunks only add warm-up work).
	int chunk_len = c->chunk_len_req, chunk_len_b = c->chunk_len_req;
	if (chunk_len <= 0) {
		auto len_for = [&](int per_sm) {
			int64_t target = (int64_t)c->sm_count * per_sm - n_active; // every sequence rounds its chunk count up
			if (target < 1) target = 1;
			int64_t cl = (c->active_bins + target - 1) / target;
			if (cl < 512) cl = 512;
			return (int)std::min<int64_t>(cl, 1 << 24);
		};
		chunk_len = len_for(c->slots_fwd);
		chunk_len_b = len_for(c->slots_bwd);
	}
	c->chunk_len = chunk_len;
	c->chunk_len_b = chunk_len_b;
	std::vector<int32_t> k1;
	std::vector<double> cw, cw_b;
	// planned overlaps: one target T = overlap + length per plan, the smallest that fits the plan into one resident wave
	// (a record the probe has not seen yet -- batch mode: first drawn by this batch -- is cut uniformly with the fixed overlap)
	const bool planned = c->have_probe && c->probe_on && (!c->batch || c->probe_batch) && c->chunk_len_req <= 0 && c->warm_len > 0 && (int)c->probe_cum.size() == c->n_seqs;
	auto has_cum = [&](int seq) { return !c->probe_cum[seq].empty(); };
	auto cut_any = [&](int seq, int T, int dir, int hm, std::vector<Piece> *out) {
		const int Li = c->L[seq];
		if (has_cum(seq)) return cut_sequence(c->probe_cum[seq], Li, T, dir, c->probe_th, hm, c->hslow(), out);
		const int warm = dir == 0 ? c->warm_len : c->warm_len_bwd;
		const int clen = std::max(512, dir == 0 ? (T - warm) / 2 : T);
		const int nc0 = (Li + clen - 1) / clen;
		if (out)
			for (int k = 0; k < nc0; ++k) {
				const int64_t a = (int64_t)Li * k / nc0, b = (int64_t)Li * (k + 1) / nc0;
				out->push_back({(int32_t)a, (int32_t)(b - a), warm, 0});
			}
		return nc0;
	};
	int T_plan[2] = {0, 0}, hmax_plan[2] = {0, 0};
	if (planned) {
		// forward plan: the kernel runs overlap + length steps per chunk (a stored step costing two warm-up steps) and lasts as
		// long as its most expensive chunk, so no overlap may exceed T - 2 * 512 (boundaries that need more are "slow": short overlap, operators); T never drops below the fixed overlap's
		// reach (small shards: more slow boundaries would cost more than shorter chunks save).  backward plan: T is the chunk
		// length; its overlaps run in the side stream's warm-up kernel, which should not outlast the forward kernel.
		for (int dir = 0; dir < 2; ++dir) {
			const int target = std::max(1, c->sm_count * (dir == 0 ? c->slots_fwd : c->slots_bwd));
			auto cap = [&](int T) { return dir == 0 ? std::min(c->hmax(), T - 1024) : std::min(c->hmax(), std::max(hmax_plan[0], c->warm_len)); };
			auto count = [&](int T) {
				int64_t n = 0;
				for (int v = 0; v < c->n_vseq; ++v)
					if (vs[v].mult > 0) n += cut_any(vs[v].seq, T, dir, cap(T), nullptr);
				return n;
			};
			int lo = dir == 0 ? c->warm_len + 1024 : 512, hi = 1 << 25;
			while (lo < hi) { // count is non-increasing in T
				const int mid = lo + (hi - lo) / 2;
				if (count(mid) <= target) hi = mid; else lo = mid + 1;
			}
			T_plan[dir] = lo;
			hmax_plan[dir] = cap(lo);
		}
	}
	std::vector<int32_t> warm_f, warm_b, slow_f, slow_b;
	auto build_plan = [&](int clen, bool is_main, std::vector<Chunk> &chunks, std::vector<Chunk> &subs,
	                      std::vector<int32_t> &sub_parent, std::vector<int32_t> &chunk_sub0, std::vector<double> &w, std::vector<int32_t> &rep_c0,
	                      std::vector<int32_t> &warm_of, std::vector<int32_t> &slow_of) {
		rep_c0.assign((size_t)c->n_rep + 1, 0);
		int next_rep = 0;
		std::vector<Piece> pieces;
		for (int v = 0; v < c->n_vseq; ++v) {
			const int i = vs[v].seq, Li = c->L[i];
			pieces.clear();
			if (vs[v].mult > 0) {
				if (planned) {
					cut_any(i, T_plan[is_main ? 0 : 1], is_main ? 0 : 1, hmax_plan[is_main ? 0 : 1], &pieces);
				} else {
					const int nc0 = (Li + clen - 1) / clen;
					for (int k = 0; k < nc0; ++k) {
						const int64_t a = (int64_t)Li * k / nc0, b = (int64_t)Li * (k + 1) / nc0;
						pieces.push_back({(int32_t)a, (int32_t)(b - a), is_main ? c->warm_len : c->warm_len_bwd, 0});
					}
				}
			}
			const int nc = (int)pieces.size();
			const int64_t gb = vs[v].gb0;
			while (next_rep <= vs[v].rep) rep_c0[next_rep++] = (int32_t)chunks.size();
			if (is_main) {
				c->seq_c0[v] = (int)chunks.size();
				c->seq_nc[v] = nc;
				c->seq_gb0[v] = gb;
			}
			for (int k = 0; k < nc; ++k) {
				Chunk ch;
				ch.seq = v;
				ch.flags = (k == 0 ? CH_FIRST : 0) | (k == nc - 1 ? CH_LAST : 0);
				ch.u0 = pieces[k].u0;
				ch.len = pieces[k].len;
				ch.gb0 = gb + pieces[k].u0;
				ch.ow0 = c->seq_ow0[i];
				ch.Lseq = Li;
				ch.rep = vs[v].rep;
				if (is_main && nc > 1) k1.push_back((int)chunks.size());
				chunks.push_back(ch);
				w.push_back((double)vs[v].mult);
				warm_of.push_back(pieces[k].warm);
				slow_of.push_back(pieces[k].slow);
			}
		}
		while (next_rep <= c->n_rep) rep_c0[next_rep++] = (int32_t)chunks.size();
		// sub-chunks (repair granularity): every chunk split into equal pieces of about sub_len bins
		chunk_sub0.assign(chunks.size() + 1, 0);
		for (size_t ci = 0; ci < chunks.size(); ++ci) {
			const Chunk &pc = chunks[ci];
			const int ns = std::max(1, (pc.len + c->sub_len - 1) / c->sub_len);
			chunk_sub0[ci] = (int32_t)subs.size();
			for (int k = 0; k < ns; ++k) {
				Chunk sc_ = pc;
				const int64_t a = (int64_t)pc.len * k / ns, b = (int64_t)pc.len * (k + 1) / ns;
				sc_.u0 = pc.u0 + (int)a;
				sc_.len = (int)(b - a);
				sc_.gb0 = pc.gb0 + a;
				sc_.flags = ((pc.flags & CH_FIRST) && k == 0 ? CH_FIRST : 0) | ((pc.flags & CH_LAST) && k == ns - 1 ? CH_LAST : 0);
				subs.push_back(sc_);
				sub_parent.push_back((int32_t)ci);
			}
		}
		chunk_sub0[chunks.size()] = (int32_t)subs.size();
	};
	std::vector<Chunk> subs, chunks_b, subs_b;
	std::vector<int32_t> sub_parent, chunk_sub0, sub_parent_b, chunk_sub0_b;
	c->chunks.clear();
	build_plan(chunk_len, true, c->chunks, subs, sub_parent, chunk_sub0, cw, c->rep_c0, warm_f, slow_f);
	build_plan(chunk_len_b, false, chunks_b, subs_b, sub_parent_b, chunk_sub0_b, cw_b, c->rep_c0_b, warm_b, slow_b);
	c->planned = planned;
	if (planned) { // (what the info block reports for a planned context: the typical chunk, the mean overlaps)
		int64_t sl = 0, sw = 0, sl_b = 0, sw_b = 0;
		c->slow_f = c->slow_b = 0;
		for (size_t i = 0; i < c->chunks.size(); ++i) { sl += c->chunks[i].len; sw += warm_f[i]; c->slow_f += slow_f[i]; }
		for (size_t i = 0; i < chunks_b.size(); ++i) { sl_b += chunks_b[i].len; sw_b += warm_b[i]; c->slow_b += slow_b[i]; }
		c->chunk_len = c->chunks.empty() ? 0 : (int)(sl / (int64_t)c->chunks.size());
		c->chunk_len_b = chunks_b.empty() ? 0 : (int)(sl_b / (int64_t)chunks_b.size());
		c->avg_warm_f = c->chunks.empty() ? 0.0 : (double)sw / (double)c->chunks.size();
		c->avg_warm_b = chunks_b.empty() ? 0.0 : (double)sw_b / (double)chunks_b.size();
		++c->probe_plans;
	}
	c->n_chunks = (int)c->chunks.size();
	c->n_sub = (int)subs.size();
	c->n_chunks_b = (int)chunks_b.size();
	c->n_sub_b = (int)subs_b.size();
	c->n_k1 = (int)k1.size();
	CUDA_TRY(cudaStreamSynchronize(c->stream), PSMC_B200_ECUDA);
	CUDA_TRY(cudaStreamSynchronize(c->stream2), PSMC_B200_ECUDA);
	// forward spill: rows of every virtual sequence (grow-only; single mode needs total_bins rows, a batch the drawn records of all its models)
	if (rows > c->cap_bins) {
		cudaFree(c->d_fhat); cudaFree(c->d_sc); cudaFree(c->d_ghat);
		c->d_fhat = nullptr; c->d_sc = nullptr; c->d_ghat = nullptr;
		c->bytes_total -= c->bytes_forward;
		c->bytes_forward = 0; c->cap_bins = 0;
		// (batch mode: 1/8 head room, so that the next batch of the same size -- whose draws differ -- reuses the allocation)
		const size_t need = (size_t)std::max<int64_t>(c->batch ? rows + rows / 8 : rows, 1);
		// (+8 rows: the backward pass stages tiles of 8 rows aligned in the row index, the last one may reach past the spill)
		cudaError_t e1 = cudaMalloc((void **)&c->d_fhat, (need + 8) * NP * sizeof(double));
		cudaError_t e2 = e1 == cudaSuccess ? cudaMalloc((void **)&c->d_sc, (need + 8) * sizeof(double)) : e1;
		if (e2 == cudaSuccess) { // (the padding rows are read, never used: give them defined contents)
			cudaMemsetAsync(c->d_fhat + need * NP, 0, 8 * NP * sizeof(double), c->stream);
			cudaMemsetAsync(c->d_sc + need, 0, 8 * sizeof(double), c->stream);
		}
		cudaError_t e3 = (e2 == cudaSuccess && c->dense) ? cudaMalloc((void **)&c->d_ghat, need * NP * sizeof(double)) : e2;
		if (e3 != cudaSuccess) {
			cudaFree(c->d_fhat); cudaFree(c->d_sc); c->d_fhat = nullptr; c->d_sc = nullptr;
			cudaGetLastError();
			return set_err(PSMC_B200_ECUDA, "cudaMalloc of the forward spill (%lld bins x %d states) failed: %s", (long long)rows, NP, cudaGetErrorString(e3));
		}
		if (c->dense) cudaMemsetAsync(c->d_ghat, 0, need * NP * sizeof(double), c->stream);
		c->cap_bins = (int64_t)need;
		c->bytes_forward = (int64_t)need * NP * 8 * (c->dense ? 2 : 1) + (int64_t)need * 8;
		c->bytes_total += c->bytes_forward;
		c->dense_valid = false;
	}
	if (!c->d_chunks || c->n_chunks > c->cap_chunks || c->n_chunks_b > c->cap_chunks_b || c->n_sub > c->cap_sub || c->n_sub_b > c->cap_sub_b || c->n_k1 > c->cap_k1 ||
	    c->n_vseq > c->cap_vseq || c->n_rep > c->cap_rep) {
		// grow-only, with head room: a replicate never needs much more than the all-sequences plan
		free_plan(c);
		auto room = [](int n) { return n + n / 8 + 16; };
		const int cc = room(c->n_chunks), cb = room(c->n_chunks_b), cs = room(c->n_sub), csb = room(c->n_sub_b), ck = room(c->n_k1);
		const int cv = room(c->n_vseq), cr = room(c->n_rep);
		bool ok = true;
		auto alloc = [&](void **ptr, size_t bytes) {
			if (!ok) return;
			if (bytes == 0) bytes = 256;
			cudaError_t e_ = cudaMalloc(ptr, bytes);
			if (e_ != cudaSuccess) { set_err(PSMC_B200_ECUDA, "cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e_)); *ptr = nullptr; ok = false; return; }
			c->bytes_plan += (int64_t)bytes;
		};
		const size_t cm = (size_t)std::max(cc, cb);
		alloc((void **)&c->d_chunks, sizeof(Chunk) * (size_t)cc);
		alloc((void **)&c->d_k1, sizeof(int32_t) * (size_t)ck);
		alloc((void **)&c->d_T, sizeof(double) * (size_t)cc * NP * NP);
		alloc((void **)&c->d_Tex, sizeof(int32_t) * (size_t)cc * NP);
		alloc((void **)&c->d_vstart, sizeof(double) * (size_t)cc * NP);
		alloc((void **)&c->d_bend, sizeof(double) * (size_t)cc * NP);
		alloc((void **)&c->d_part, sizeof(double) * cm * S_COUNT * NP);
		alloc((void **)&c->d_llpart, sizeof(double) * (size_t)cc);
		alloc((void **)&c->d_fwarm, sizeof(double) * (size_t)cc * NP);
		alloc((void **)&c->d_bwarm, sizeof(double) * cm * NP);
		alloc((void **)&c->d_bexact, sizeof(double) * cm * NP);
		alloc((void **)&c->d_bsave[0], sizeof(double) * (size_t)cb * NP);
		alloc((void **)&c->d_bsave[1], sizeof(double) * (size_t)cb * NP);
		alloc((void **)&c->d_chunks_b, sizeof(Chunk) * (size_t)cb);
		alloc((void **)&c->d_sub_b, sizeof(Chunk) * (size_t)csb);
		alloc((void **)&c->d_sub_parent_b, sizeof(int32_t) * (size_t)csb);
		alloc((void **)&c->d_chunk_sub0_b, sizeof(int32_t) * (size_t)(cb + 1));
		alloc((void **)&c->d_flag_b, sizeof(int32_t) * (size_t)(cb + 2));
		alloc((void **)&c->d_flag, sizeof(int32_t) * (size_t)(cc + 2));
		alloc((void **)&c->d_sub, sizeof(Chunk) * (size_t)cs);
		alloc((void **)&c->d_sub_parent, sizeof(int32_t) * (size_t)cs);
		alloc((void **)&c->d_chunk_sub0, sizeof(int32_t) * (size_t)(cc + 1));
		alloc((void **)&c->d_Tsub, sizeof(double) * (size_t)cs * NP * NP);
		alloc((void **)&c->d_Texsub, sizeof(int32_t) * (size_t)cs * NP);
		alloc((void **)&c->d_Tsub_b, sizeof(double) * (size_t)csb * NP * NP);
		alloc((void **)&c->d_Texsub_b, sizeof(int32_t) * (size_t)csb * NP);
		for (int q = 0; q < 2; ++q) {
			alloc((void **)&c->d_pred[q], sizeof(int32_t) * (size_t)(cc + 2));
			alloc((void **)&c->d_pred_b[q], sizeof(int32_t) * (size_t)(cb + 2));
		}
		alloc((void **)&c->d_vsub, sizeof(double) * (size_t)cs * NP);
		alloc((void **)&c->d_bsub, sizeof(double) * (size_t)csb * NP);
		alloc((void **)&c->d_llsub, sizeof(double) * (size_t)cs);
		alloc((void **)&c->d_partsub, sizeof(double) * (size_t)csb * S_COUNT * NP);
		alloc((void **)&c->d_cw, sizeof(double) * (size_t)cc);
		alloc((void **)&c->d_cw_b, sizeof(double) * (size_t)cb);
		alloc((void **)&c->d_seq_c0, sizeof(int32_t) * (size_t)cv);
		alloc((void **)&c->d_seq_nc, sizeof(int32_t) * (size_t)cv);
		alloc((void **)&c->d_rep_c0, sizeof(int32_t) * (size_t)(cr + 1));
		alloc((void **)&c->d_rep_c0_b, sizeof(int32_t) * (size_t)(cr + 1));
		alloc((void **)&c->d_warm_f, sizeof(int32_t) * (size_t)cc);
		alloc((void **)&c->d_warm_b, sizeof(int32_t) * (size_t)cb);
		c->bytes_total += c->bytes_plan;
		if (!ok) { free_plan(c); return PSMC_B200_ECUDA; }
		c->cap_chunks = cc; c->cap_chunks_b = cb; c->cap_sub = cs; c->cap_sub_b = csb; c->cap_k1 = ck; c->cap_vseq = cv; c->cap_rep = cr;
	}
	c->bytes_transfer = (int64_t)c->n_chunks * NP * NP * 8;
	cudaStream_t st = c->stream;
#define UP(dst, vec)                                                                                                  \
	do {                                                                                                              \
		if (!(vec).empty())                                                                                           \
			CUDA_TRY(cudaMemcpyAsync(dst, (vec).data(), sizeof((vec)[0]) * (vec).size(), cudaMemcpyHostToDevice, st), PSMC_B200_ECUDA); \
	} while (0)
	UP(c->d_chunks, c->chunks); UP(c->d_sub, subs); UP(c->d_sub_parent, sub_parent); UP(c->d_chunk_sub0, chunk_sub0);
	UP(c->d_chunks_b, chunks_b); UP(c->d_sub_b, subs_b); UP(c->d_sub_parent_b, sub_parent_b); UP(c->d_chunk_sub0_b, chunk_sub0_b);
	UP(c->d_k1, k1); UP(c->d_cw, cw); UP(c->d_cw_b, cw_b);
	UP(c->d_seq_c0, c->seq_c0); UP(c->d_seq_nc, c->seq_nc); UP(c->d_rep_c0, c->rep_c0); UP(c->d_rep_c0_b, c->rep_c0_b);
	UP(c->d_warm_f, warm_f); UP(c->d_warm_b, warm_b);
#undef UP
	CUDA_TRY(cudaMemsetAsync(c->d_flag_b, 0, sizeof(int32_t) * (size_t)(c->n_chunks_b + 2), st), PSMC_B200_ECUDA);
	CUDA_TRY(cudaMemsetAsync(c->d_flag, 0, sizeof(int32_t) * (size_t)(c->n_chunks + 2), st), PSMC_B200_ECUDA);
	for (int q = 0; q < 2; ++q) { // no prediction for a new plan ...
		CUDA_TRY(cudaMemsetAsync(c->d_pred[q], 0, sizeof(int32_t) * (size_t)(c->n_chunks + 2), st), PSMC_B200_ECUDA);
		CUDA_TRY(cudaMemsetAsync(c->d_pred_b[q], 0, sizeof(int32_t) * (size_t)(c->n_chunks_b + 2), st), PSMC_B200_ECUDA);
	}
	if (planned) { // ... except the boundaries the probe already knows no overlap reaches: their operators are computed ahead of time
		if (!slow_f.empty()) CUDA_TRY(cudaMemcpyAsync(c->d_pred[c->pred_cur] + 1, slow_f.data(), sizeof(int32_t) * slow_f.size(), cudaMemcpyHostToDevice, st), PSMC_B200_ECUDA);
		if (!slow_b.empty()) CUDA_TRY(cudaMemcpyAsync(c->d_pred_b[c->pred_cur] + 1, slow_b.data(), sizeof(int32_t) * slow_b.size(), cudaMemcpyHostToDevice, st), PSMC_B200_ECUDA);
	}
	CUDA_TRY(cudaMemsetAsync(c->d_vstart, 0, sizeof(double) * (size_t)std::max(c->n_chunks, 1) * NP, st), PSMC_B200_ECUDA);
	CUDA_TRY(cudaMemsetAsync(c->d_bend, 0, sizeof(double) * (size_t)std::max(c->n_chunks, 1) * NP, st), PSMC_B200_ECUDA);
	CUDA_TRY(cudaStreamSynchronize(st), PSMC_B200_ECUDA); // the host vectors above go out of scope
	c->have_prev = false;
	c->fwd_valid = false;
	c->launched = false;
	++c->replans;
	return 0;
}

// model blocks and statistics vectors for n models (device + pinned staging), grow-only
static int ensure_models(psmc_b200_ctx *c, int n)
{
	if (n <= c->cap_models) return 0;
	const size_t mb = sizeof(double) * M_COUNT * c->NP * (size_t)n, sb = sizeof(double) * (size_t)(S_COUNT * c->N + 1) * (size_t)n;
	if (c->stream) CUDA_TRY(cudaStreamSynchronize(c->stream), PSMC_B200_ECUDA);
	cudaFree(c->d_model); cudaFree(c->d_stats);
	if (c->h_model) cudaFreeHost(c->h_model);
	if (c->h_stats) cudaFreeHost(c->h_stats);
	c->d_model = c->d_stats = c->h_model = c->h_stats = nullptr;
	c->cap_models = 0;
	CUDA_TRY(cudaMalloc((void **)&c->d_model, mb), PSMC_B200_ECUDA);
	CUDA_TRY(cudaMalloc((void **)&c->d_stats, sb), PSMC_B200_ECUDA);
	CUDA_TRY(cudaMallocHost((void **)&c->h_model, mb), PSMC_B200_ECUDA);
	CUDA_TRY(cudaMallocHost((void **)&c->h_stats, sb), PSMC_B200_ECUDA);
	c->cap_models = n;
	return 0;
}

extern "C" int psmc_b200_create(psmc_b200_ctx **out, int32_t n_seqs, const int32_t *L, const signed char *const *seqs,
                                int32_t n_states, int32_t device, int32_t chunk_len, uint32_t flags)
{
	(void)flags;
	if (!out) return set_err(PSMC_B200_EINVAL, "out is NULL");
	*out = nullptr;
	if (n_seqs < 0 || (n_seqs > 0 && (!L || !seqs))) return set_err(PSMC_B200_EINVAL, "bad sequence arguments");
	if (n_states < 1 || n_states > 128) return set_err(PSMC_B200_EINVAL, "n_states=%d unsupported on the GPU path (1..128)", n_states);
	int ndev = 0;
	if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0)
		return set_err(PSMC_B200_ENODEV, "no CUDA device available (there is no CPU fallback)");
	if (device < 0 || device >= ndev) return set_err(PSMC_B200_ENODEV, "device %d out of range (have %d)", device, ndev);
	CUDA_TRY(cudaSetDevice(device), PSMC_B200_ENODEV);

	psmc_b200_ctx *c = new psmc_b200_ctx();
	c->device = device;
	c->N = n_states;
	c->NP = pad_states(n_states);
	c->SPL = c->NP / 32;
	// keep non-empty sequences only (the reference reads uninitialised memory for L == 0; nothing to count there)
	std::vector<const signed char *> sp;
	c->n_seqs_given = n_seqs;
	c->kept_of.assign((size_t)std::max(n_seqs, 1), -1);
	for (int i = 0; i < n_seqs; ++i) {
		if (L[i] < 0) { free_ctx(c); return set_err(PSMC_B200_EINVAL, "negative sequence length"); }
		if (L[i] == 0) continue;
		c->kept_of[i] = (int32_t)c->L.size();
		c->L.push_back(L[i]);
		sp.push_back(seqs[i]);
		c->total_bins += L[i];
	}
	c->n_seqs = (int)c->L.size();
	c->mult.assign((size_t)c->n_seqs, 1);
	cudaDeviceProp prop;
	CUDA_TRY(cudaGetDeviceProperties(&prop, device), PSMC_B200_ENODEV);
	c->sm_count = prop.multiProcessorCount;
	{
		const char *env = getenv("PSMC_B200_GEN");
		if (env && atoi(env) == 1) c->gen = 1;
		env = getenv("PSMC_B200_TMA");
		if (env) c->staged_bwd = atoi(env) != 0;
		env = getenv("PSMC_B200_PROBE");
		if (env) c->probe_on = atoi(env) != 0;
		env = getenv("PSMC_B200_PROBE_BATCH");
		if (env) c->probe_batch = atoi(env) != 0;
		env = getenv("PSMC_B200_PROBE_TH");
		if (env && atof(env) > 0) c->probe_th = atof(env);
		env = getenv("PSMC_B200_PROBE_HMAX");
		if (env && atoi(env) > 0) c->probe_hmax = atoi(env);
		env = getenv("PSMC_B200_G2_FWD");
		if (env && atoi(env) == 16) c->g2_fwd = 16;
		if (env && atoi(env) == 8) c->g2_fwd_auto = false;
		env = getenv("PSMC_B200_G2_FWD_AUTO"); // chunks per SM up to which the forward kernel switches to 16-lane groups (0 = never)
		if (env && atoi(env) >= 0) c->g2_fwd_auto_max = atoi(env);
		env = getenv("PSMC_B200_G2_BWW");
		if (env && atoi(env) == 16) c->g2_bww = 16;
		env = getenv("PSMC_B200_G_FWD");
		if (env && (atoi(env) == 8 || atoi(env) == 16 || atoi(env) == 32)) c->g_fwd = atoi(env);
		env = getenv("PSMC_B200_G_BWD");
		if (env && (atoi(env) == 8 || atoi(env) == 16 || atoi(env) == 32)) c->g_bwd = atoi(env);
		c->g_bww = 8; // measured on B200: the narrow warm-up kernel leaves the most issue slots to the concurrent forward kernel
		env = getenv("PSMC_B200_G_BWW");
		if (env && (atoi(env) == 8 || atoi(env) == 16 || atoi(env) == 32)) c->g_bww = atoi(env);
	}
	if (chunk_len <= 0) {
		const char *env = getenv("PSMC_B200_CHUNK");
		if (env && atoi(env) > 0) chunk_len = atoi(env);
	}
	c->chunk_len_req = chunk_len > 0 ? chunk_len : 0;
	{ // resident chunk slots per SM of the forward / backward kernels (one wave each)
		int sf = 4, sb = 4;
		switch (c->NP) {
		case 32: chunk_slots<32>(c, &sf, &sb); break;
		case 64: chunk_slots<64>(c, &sf, &sb); break;
		default: chunk_slots<128>(c, &sf, &sb); break;
		}
		// measured on B200: beyond 16 forward chunks per SM the extra warm-up overlaps cost more than the shorter chunks save
		if (sf > 16) sf = 16;
		const char *env = getenv("PSMC_B200_CHUNKS_PER_SM");
		if (env && atoi(env) > 0) sf = sb = atoi(env);
		env = getenv("PSMC_B200_CHUNKS_PER_SM_FWD");
		if (env && atoi(env) > 0) sf = atoi(env);
		c->slots_fwd = sf; c->slots_bwd = sb;
	}
	{ // warm-up overlap: PSMC_B200_WARM=0 disables the fast path (always transfer matrices)
		const char *env = getenv("PSMC_B200_WARM");
		c->warm_len = env ? atoi(env) : 12288;
		if (c->warm_len < 0) c->warm_len = 0;
		env = getenv("PSMC_B200_WARM_BWD");
		c->warm_len_bwd = (env && atoi(env) > 0) ? atoi(env) : c->warm_len + c->warm_len / 3; // (it shares the SMs with the forward kernel: not free)
		c->warm_bwd_fixed = (env && atoi(env) > 0);
		env = getenv("PSMC_B200_WARM_HOT");
		if (env && atoi(env) >= 0) c->warm_hot = atoi(env);
		env = getenv("PSMC_B200_REPAIR_ROUNDS");
		if (env && atoi(env) >= 0) { c->repair_rounds = atoi(env); c->rounds_fixed = true; }
		c->rounds_cur = c->repair_rounds;
		env = getenv("PSMC_B200_CERT_EPS");
		if (env && atof(env) > 0) c->cert_eps = atof(env);
		env = getenv("PSMC_B200_SUB_LEN");
		if (env && atoi(env) >= 64) c->sub_len = atoi(env);
	}
	// packed observations: every sequence starts on a 128-byte boundary (512 bins)
	std::vector<int64_t> &ow0 = c->seq_ow0;
	ow0.resize(c->n_seqs);
	int64_t words = 0;
	for (int i = 0; i < c->n_seqs; ++i) {
		ow0[i] = words;
		int64_t w = ((int64_t)c->L[i] + 15) / 16;
		words += (w + 31) / 32 * 32;
	}
	c->words_obs = std::max<int64_t>(words, 32);
#define ALLOC(ptr, bytes)                                                                         \
	do {                                                                                          \
		size_t b_ = (size_t)(bytes);                                                              \
		if (b_ == 0) b_ = 256;                                                                    \
		cudaError_t e_ = cudaMalloc((void **)&(ptr), b_);                                         \
		if (e_ != cudaSuccess) {                                                                  \
			int rc_ = set_err(PSMC_B200_ECUDA, "cudaMalloc(%zu bytes) failed: %s", b_, cudaGetErrorString(e_)); \
			free_ctx(c);                                                                          \
			return rc_;                                                                           \
		}                                                                                         \
		c->bytes_total += (int64_t)b_;                                                            \
	} while (0)
	c->bytes_obs = c->words_obs * 4;
	ALLOC(c->d_obs, c->bytes_obs);
	ALLOC(c->d_cert, sizeof(unsigned long long) * 144); // 16 counters + one "flagged in this round" counter per repair round and direction
#undef ALLOC
#define CTRY(call)                                                                                \
	do {                                                                                          \
		cudaError_t e_ = (call);                                                                  \
		if (e_ != cudaSuccess) {                                                                  \
			int rc_ = set_err(PSMC_B200_ECUDA, "%s failed: %s", #call, cudaGetErrorString(e_));   \
			free_ctx(c);                                                                          \
			return rc_;                                                                           \
		}                                                                                         \
	} while (0)
	CTRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
	CTRY(cudaStreamCreateWithFlags(&c->stream2, cudaStreamNonBlocking));
	CTRY(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
	CTRY(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
	CTRY(cudaEventCreateWithFlags(&c->ev_k1f, cudaEventDisableTiming));
	CTRY(cudaEventCreateWithFlags(&c->ev_k1b, cudaEventDisableTiming));
	{
		const char *env = getenv("PSMC_B200_PREDICT");
		if (env && atoi(env) == 0) c->predict = false;
	}
	for (int i = 0; i < 8; ++i) CTRY(cudaEventCreate(&c->ev[i]));
	CTRY(cudaMallocHost((void **)&c->h_cert, sizeof(unsigned long long) * 16));
	CTRY(cudaMallocHost((void **)&c->h_obs, (size_t)c->bytes_obs));
	pack_all(c, sp.data());
	c->obs_hash = hash_obs(c);
	CTRY(cudaMemcpyAsync(c->d_obs, c->h_obs, (size_t)c->bytes_obs, cudaMemcpyHostToDevice, c->stream));
	CTRY(cudaStreamSynchronize(c->stream));
#undef CTRY
	{
		int rc = ensure_models(c, 1);
		if (rc == 0) rc = replan(c);
		if (rc != 0) { free_ctx(c); return rc; }
	}
	*out = c;
	return 0;
}

// Multiplicities of the resident sequences for the following E-steps (bootstrap replicates, aux.c:8-47).
extern "C" int psmc_b200_set_multiplicity(psmc_b200_ctx *c, const int32_t *mult)
{
	if (!c) return set_err(PSMC_B200_EINVAL, "ctx is NULL");
	std::vector<int32_t> m((size_t)c->n_seqs, 1);
	if (mult)
		for (int i = 0; i < c->n_seqs_given; ++i) {
			if (mult[i] < 0) return set_err(PSMC_B200_EINVAL, "negative multiplicity");
			if (c->kept_of[i] >= 0) m[c->kept_of[i]] = mult[i];
		}
	if (m == c->mult && !c->batch) return 0;
	CUDA_TRY(cudaSetDevice(c->device), PSMC_B200_ECUDA);
	c->mult.swap(m);
	c->batch = false;
	c->n_rep = 1;
	c->bmult.clear();
	return replan(c);
}

// Batch mode: n_rep models side by side, model r over the multiset mult[r * n_seqs + i] of the resident records.
extern "C" int psmc_b200_set_batch(psmc_b200_ctx *c, int32_t n_rep, const int32_t *mult)
{
	if (!c) return set_err(PSMC_B200_EINVAL, "ctx is NULL");
	if (n_rep < 1 || n_rep > 4096) return set_err(PSMC_B200_EINVAL, "n_rep=%d out of range (1..4096)", n_rep);
	if (c->dense) return set_err(PSMC_B200_EINVAL, "batch mode and dense counts (psmc_b200_set_dense) exclude each other");
	std::vector<int32_t> m((size_t)n_rep * std::max(c->n_seqs, 1), 1);
	if (mult)
		for (int r = 0; r < n_rep; ++r)
			for (int i = 0; i < c->n_seqs_given; ++i) {
				const int32_t v = mult[(size_t)r * c->n_seqs_given + i];
				if (v < 0) return set_err(PSMC_B200_EINVAL, "negative multiplicity");
				if (c->kept_of[i] >= 0) m[(size_t)r * c->n_seqs + c->kept_of[i]] = v;
			}
	CUDA_TRY(cudaSetDevice(c->device), PSMC_B200_ECUDA);
	int rc = ensure_models(c, n_rep);
	if (rc) return rc;
	c->bmult.swap(m);
	c->batch = true;
	c->n_rep = n_rep;
	c->estep_no = 0; // (the probe schedule restarts with every batch; records probed before keep their data)
	c->plan_dirty = false;
	rc = replan(c);
	if (rc) { // (typically: the forward spill of so many models does not fit) -- back to single mode, still usable
		c->batch = false; c->n_rep = 1; c->bmult.clear();
		char keep[512];
		snprintf(keep, sizeof(keep), "%s", g_err);
		if (replan(c) == 0) set_err(rc, "%s", keep);
	}
	return rc;
}

extern "C" int psmc_b200_mem_info(int32_t device, int64_t *free_bytes, int64_t *total_bytes)
{
	size_t f = 0, t = 0;
	CUDA_TRY(cudaSetDevice(device), PSMC_B200_ENODEV);
	CUDA_TRY(cudaMemGetInfo(&f, &t), PSMC_B200_ECUDA);
	if (free_bytes) *free_bytes = (int64_t)f;
	if (total_bytes) *total_bytes = (int64_t)t;
	return 0;
}

extern "C" int psmc_b200_create_cat(psmc_b200_ctx **out, int32_t n_seqs, const int32_t *L, const signed char *seqs_cat,
                                    int32_t n_states, int32_t device, int32_t chunk_len, uint32_t flags)
{
	if (n_seqs < 0 || (n_seqs > 0 && (!L || !seqs_cat))) return set_err(PSMC_B200_EINVAL, "bad sequence arguments");
	std::vector<const signed char *> p((size_t)std::max(n_seqs, 1));
	const signed char *q = seqs_cat;
	for (int i = 0; i < n_seqs; ++i) {
		p[i] = q;
		if (L[i] > 0) q += L[i];
	}
	return psmc_b200_create(out, n_seqs, L, p.data(), n_states, device, chunk_len, flags);
}

extern "C" int psmc_b200_upload(psmc_b200_ctx *c, int32_t n_seqs, const int32_t *L, const signed char *const *seqs)
{
	if (!c || (n_seqs > 0 && (!L || !seqs))) return set_err(PSMC_B200_EINVAL, "NULL argument");
	std::vector<const signed char *> sp;
	int k = 0;
	for (int i = 0; i < n_seqs; ++i) {
		if (L[i] == 0) continue;
		if (k >= c->n_seqs || L[i] != c->L[k]) return set_err(PSMC_B200_EINVAL, "upload: sequence lengths differ from the ones the context was created with");
		sp.push_back(seqs[i]);
		++k;
	}
	if (k != c->n_seqs) return set_err(PSMC_B200_EINVAL, "upload: %d sequences given, context holds %d", k, c->n_seqs);
	CUDA_TRY(cudaSetDevice(c->device), PSMC_B200_ECUDA);
	CUDA_TRY(cudaStreamSynchronize(c->stream), PSMC_B200_ECUDA);
	pack_all(c, sp.data());
	CUDA_TRY(cudaMemcpyAsync(c->d_obs, c->h_obs, (size_t)c->bytes_obs, cudaMemcpyHostToDevice, c->stream), PSMC_B200_ECUDA);
	{ // the mixing probe describes the OBSERVATIONS: it survives a re-upload of the same data, not new data
		const uint64_t h = hash_obs(c);
		if (h != c->obs_hash) {
			c->obs_hash = h;
			c->estep_no = 0;
			if (c->have_probe) { c->have_probe = false; c->probe_cum.clear(); c->plan_dirty = c->planned; }
		}
	}
	c->fwd_valid = false; // (have_prev stays: stale vectors are still legal warm starts, the certificate decides)
	return 0;
}

extern "C" int psmc_b200_upload_cat(psmc_b200_ctx *c, int32_t n_seqs, const int32_t *L, const signed char *seqs_cat)
{
	if (n_seqs < 0 || (n_seqs > 0 && (!L || !seqs_cat))) return set_err(PSMC_B200_EINVAL, "bad sequence arguments");
	std::vector<const signed char *> p((size_t)std::max(n_seqs, 1));
	const signed char *q = seqs_cat;
	for (int i = 0; i < n_seqs; ++i) {
		p[i] = q;
		if (L[i] > 0) q += L[i];
	}
	return psmc_b200_upload(c, n_seqs, L, p.data());
}

static int check_model(const psmc_b200_ctx *c, const psmc_b200_model *m)
{
	if (!m || !m->a0 || !m->e || !m->U || !m->V || !m->W || !m->Z || !m->D) return set_err(PSMC_B200_EINVAL, "model has NULL arrays");
	if (m->n_states != c->N) return set_err(PSMC_B200_EINVAL, "model has %d states, context has %d", m->n_states, c->N);
	const double *arr[7] = {m->a0, m->U, m->V, m->W, m->Z, m->D, m->e};
	for (int a = 0; a < 7; ++a)
		for (int k = 0; k < (a == 6 ? 2 : 1) * c->N; ++k)
			if (!isfinite(arr[a][k])) return set_err(PSMC_B200_ENUMERIC, "non-finite value in the model");
	return 0;
}

static void stage_model(psmc_b200_ctx *c, const psmc_b200_model *m, int r = 0)
{
	const int N = c->N, NP = c->NP;
	double *h = c->h_model + (size_t)r * M_COUNT * NP;
	memset(h, 0, sizeof(double) * M_COUNT * NP);
	for (int k = 0; k < N; ++k) {
		h[M_A0 * NP + k] = m->a0[k];
		h[M_E0 * NP + k] = m->e[k];
		h[M_E1 * NP + k] = m->e[N + k];
		h[M_U * NP + k] = (k > 0) ? m->U[k] : 0.0;       // U_0 multiplies an empty sum
		h[M_V * NP + k] = (k < N - 1) ? m->V[k] : 0.0;   // V_{N-1} never used
		h[M_W * NP + k] = (k < N - 1) ? m->W[k] : 0.0;   // W_{N-1} multiplies an empty sum
		h[M_Z * NP + k] = (k > 0) ? m->Z[k] : 0.0;       // Z_0 never used
		h[M_D * NP + k] = m->D[k];
	}
}

// ---- lane-group width dispatch: G lanes per chunk, 32/G chunks per warp, SPL = NP/G states per lane ----
// forward kernels are instantiated for SPL <= 8, backward kernels (7*SPL accumulators per lane) for SPL <= 4.
static inline int blocks_for(int n_chunks, int G) { const int per_block = 4 * (32 / G); return (n_chunks + per_block - 1) / per_block; }

// Kernel generation 2 (DualScan, see above) exists for G = 8 and 16.  Which (G, generation) a context uses:
//   forward / forward repair / backward warm-up: gen 2 with G = 8 (NP <= 64) or 16 (NP = 128)
//   backward / backward repair (7 accumulators per state): gen 2 with G = 8 (NP = 32) or 16 (NP = 64); NP = 128 keeps gen 1, G = 32
// PSMC_B200_GEN=1 selects generation 1 everywhere (then PSMC_B200_G_FWD / _BWD / _BWW choose the group widths as before).
template <int NP>
struct Gen2 {
	static constexpr int G_BWD = (NP == 32) ? 8 : 16;
	static constexpr bool BWD_OK = NP <= 64;
};

template <int NP>
static void run_forward(psmc_b200_ctx *c, int warm, int use_prev)
{
	cudaStream_t st = c->stream;
#define FWD(G_, V_) LAUNCH((k_forward<NP / G_, G_, V_>), blocks_for(c->n_chunks, G_), 128, st, c->d_chunks, c->n_chunks, c->d_obs, c->d_model, c->d_vstart, warm, use_prev, c->d_fhat, c->d_sc, c->d_llpart, c->d_fwarm, (c->planned && warm > 0 && !use_prev) ? c->d_warm_f : nullptr)
	if (c->wide_fwd(NP)) FWD(16, 2);
	else if (c->gen == 2) FWD(8, 2);
	else if (c->g_fwd == 8 && NP / 8 <= 8) FWD(8, 1);
	else if (c->g_fwd <= 16 && NP / 16 <= 8) FWD(16, 1);
	else FWD(32, 1);
#undef FWD
}
template <int NP>
static void run_forward_repair(psmc_b200_ctx *c, const unsigned long long *rcnt)
{
	cudaStream_t st = c->stream;
#define FWR(G_, V_) LAUNCH((k_forward_repair<NP / G_, G_, V_>), blocks_for(c->n_sub, G_), 128, st, c->d_sub, c->n_sub, c->d_sub_parent, c->d_chunk_sub0, c->d_obs, c->d_model, c->d_flag + 1, c->d_vsub, c->d_fhat, c->d_sc, c->d_llsub, c->d_fwarm, c->d_cert + 4, rcnt)
	if (c->wide_fwd(NP)) FWR(16, 2);
	else if (c->gen == 2) FWR(8, 2);
	else if (c->g_fwd == 8 && NP / 8 <= 8) FWR(8, 1);
	else if (c->g_fwd <= 16 && NP / 16 <= 8) FWR(16, 1);
	else FWR(32, 1);
#undef FWR
}
template <int NP>
static void run_backward(psmc_b200_ctx *c, const Chunk *chunks, int n, const double *bdir, int publish, double *bsave_next)
{
	cudaStream_t st = c->stream;
#define BWD(G_, V_) LAUNCH((k_backward<NP / G_, G_, V_>), blocks_for(n, G_), 128, st, chunks, n, c->d_obs, c->d_model, bdir, publish, c->d_fhat, c->d_sc, c->d_part, c->d_bexact, bsave_next, c->warm_hot, (c->dense && V_ == 2) ? c->d_ghat : nullptr)
	// (forcing 4 resident blocks per SM with __launch_bounds__(128, 4) was measured: the spills cost more than the occupancy gives)
	if constexpr (Gen2<NP>::BWD_OK) {
		if (c->gen == 2 && c->staged_bwd) {
			constexpr int GB = Gen2<NP>::G_BWD;
			typedef BackwardStaged<NP / GB, GB> BS;
			LAUNCH_SMEM((k_backward_staged<NP / GB, GB>), blocks_for(n, GB), 128, BS::SMEM_BYTES, st, chunks, n, c->d_obs, c->d_model, bdir, publish, c->d_fhat, c->d_sc,
			            c->d_part, c->d_bexact, bsave_next, c->warm_hot, c->dense ? c->d_ghat : nullptr, c->d_cert + 15);
			return;
		}
		if (c->gen == 2) {
			BWD(Gen2<NP>::G_BWD, 2);
			return;
		}
	}
	if (c->g_bwd == 8 && NP / 8 <= 4) BWD(8, 1);
	else if (c->g_bwd <= 16 && NP / 16 <= 4) BWD(16, 1);
	else BWD(32, 1);
#undef BWD
}
template <int NP>
static void run_backward_warm(psmc_b200_ctx *c, cudaStream_t st, int warm, int use_prev)
{
#define BWW(G_, V_) LAUNCH((k_backward_warm<NP / G_, G_, V_>), blocks_for(c->n_chunks_b, G_), 128, st, c->d_chunks_b, c->n_chunks_b, c->d_obs, c->d_model, warm, c->d_bwarm, use_prev ? c->d_bsave[c->bsave_cur] : nullptr, (c->planned && !use_prev) ? c->d_warm_b : nullptr)
	if (c->gen == 2 && (c->g2_bww == 16 || NP > 64)) BWW(16, 2);
	else if (c->gen == 2) BWW(8, 2);
	else if (c->g_bww == 8 && NP / 8 <= 8) BWW(8, 1);
	else if (c->g_bww <= 16 && NP / 16 <= 8) BWW(16, 1);
	else BWW(32, 1);
#undef BWW
}
template <int NP>
static void run_backward_repair(psmc_b200_ctx *c, const unsigned long long *rcnt)
{
	cudaStream_t st = c->stream;
#define BWR(G_, V_) LAUNCH((k_backward_repair<NP / G_, G_, V_>), blocks_for(c->n_sub_b, G_), 128, st, c->d_sub_b, c->n_sub_b, c->d_sub_parent_b, c->d_chunk_sub0_b, c->d_chunks_b, c->d_obs, c->d_model, c->d_flag_b + 1, c->d_bsub, c->d_fhat, c->d_sc, c->d_partsub, c->d_bwarm, c->d_bexact, c->d_cert + 4, (c->dense && V_ == 2) ? c->d_ghat : nullptr, rcnt)
	if constexpr (Gen2<NP>::BWD_OK) {
		if (c->gen == 2) {
			BWR(Gen2<NP>::G_BWD, 2);
			return;
		}
	}
	if (c->g_bwd == 8 && NP / 8 <= 4) BWR(8, 1);
	else if (c->g_bwd <= 16 && NP / 16 <= 4) BWR(16, 1);
	else BWR(32, 1);
#undef BWR
}

// resident chunk slots per SM of the selected forward / backward kernels (blocks of 128 threads)
template <int NP>
static void chunk_slots(const psmc_b200_ctx *c, int *slots_fwd, int *slots_bwd)
{
	int bf = 1, bb = 1, gf = 32, gb = 32;
#define OCC(K_, G_, V_, out_) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&out_, K_<NP / G_, G_, V_>, 128, 0)
	if (c->gen == 2 && (c->g2_fwd == 16 || NP > 64)) { gf = 16; OCC(k_forward, 16, 2, bf); }
	else if (c->gen == 2) { gf = 8; OCC(k_forward, 8, 2, bf); }
	else if (c->g_fwd == 8 && NP / 8 <= 8) { gf = 8; OCC(k_forward, 8, 1, bf); }
	else if (c->g_fwd <= 16 && NP / 16 <= 8) { gf = 16; OCC(k_forward, 16, 1, bf); }
	else OCC(k_forward, 32, 1, bf);
	bool done = false;
	if constexpr (Gen2<NP>::BWD_OK) {
		if (c->gen == 2 && c->staged_bwd) {
			constexpr int GB = Gen2<NP>::G_BWD;
			typedef BackwardStaged<NP / GB, GB> BS;
			gb = GB;
			cudaFuncSetAttribute(k_backward_staged<NP / GB, GB>, cudaFuncAttributeMaxDynamicSharedMemorySize, BS::SMEM_BYTES);
			cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bb, k_backward_staged<NP / GB, GB>, 128, BS::SMEM_BYTES);
			done = true;
		} else if (c->gen == 2) { gb = Gen2<NP>::G_BWD; OCC(k_backward, Gen2<NP>::G_BWD, 2, bb); done = true; }
	}
	if (done) {}
	else if (c->g_bwd == 8 && NP / 8 <= 4) { gb = 8; OCC(k_backward, 8, 1, bb); }
	else if (c->g_bwd <= 16 && NP / 16 <= 4) { gb = 16; OCC(k_backward, 16, 1, bb); }
	else OCC(k_backward, 32, 1, bb);
#undef OCC
	*slots_fwd = bf * 4 * (32 / gf);
	*slots_bwd = bb * 4 * (32 / gb);
}

template <int SPL>
static int launch_core(psmc_b200_ctx *c, bool with_counts)
{
	constexpr int NP = 32 * SPL;
	cudaStream_t st = c->stream;
	c->launches = 0;
	cudaEventRecord(c->ev[0], st);
	if (c->n_k1 > 0) {
		constexpr int G1 = (NP > 64) ? 16 : 8, SPL1 = NP / G1, COLS = 128 / G1;
		dim3 grid((unsigned)c->n_k1, NP / COLS);
		LAUNCH((k_transfer<SPL1, G1, COLS>), grid, COLS * G1, st, c->d_chunks, c->d_k1, c->d_obs, c->d_model, c->d_T, c->d_Tex, c->N, nullptr, 0, nullptr, c->n_k1, nullptr, 0, nullptr);
		++c->launches;
	}
	cudaEventRecord(c->ev[1], st);
	if (c->n_k1 > 0) {
		LAUNCH((k_chain<NP>), 2 * c->n_vseq, NP, st, c->d_chunks, c->d_seq_c0, c->d_seq_nc, c->d_T, c->d_Tex, c->d_model, c->d_vstart, c->d_bend, c->n_vseq);
		++c->launches;
	}
	cudaEventRecord(c->ev[2], st);
	if (c->n_chunks > 0) {
		run_forward<NP>(c, 0, 0);
		++c->launches;
	}
	cudaEventRecord(c->ev[3], st);
	if (with_counts) {
		if (c->n_chunks > 0) {
			run_backward<NP>(c, c->d_chunks, c->n_chunks, c->d_bend, 0, nullptr);
			++c->launches;
		}
		cudaEventRecord(c->ev[4], st);
		LAUNCH((k_reduce), dim3(1 + S_COUNT * c->N, c->n_rep), 256, st, c->d_part, c->d_llpart, c->d_rep_c0, c->d_rep_c0, c->N, NP, c->d_stats, c->weighted ? c->d_cw : nullptr, c->weighted ? c->d_cw : nullptr);
		++c->launches;
		cudaEventRecord(c->ev[5], st);
	}
	cudaError_t e = cudaGetLastError();
	if (e != cudaSuccess) return set_err(PSMC_B200_ECUDA, "kernel launch failed: %s", cudaGetErrorString(e));
	c->fwd_valid = true;
	if (with_counts) c->have_prev = false; // the transfer-mode backward pass saves no warm-start directions
	return 0;
}

// fast path: warm-up overlaps instead of transfer matrices; boundaries the overlap did not reach are
// repaired locally (k_forward_repair / k_backward_repair, a fixed number of rounds), then certified.
template <int SPL>
static int launch_warm(psmc_b200_ctx *c)
{
	constexpr int NP = 32 * SPL;
	cudaStream_t st = c->stream;
	cudaMemsetAsync(c->d_cert, 0, sizeof(unsigned long long) * 144, st);
	const int rounds = c->rounds_cur;
	// reach of a cascade front in chunks (see k_mark_fwd): overlap / chunk length, at least one chunk
	const int spread_f = std::min(16, std::max(1, (c->warm_len + c->chunk_len - 1) / std::max(c->chunk_len, 1)));
	const int spread_b = std::min(16, std::max(1, (c->warm_len_bwd + c->chunk_len_b - 1) / std::max(c->chunk_len_b, 1)));
	cudaEventRecord(c->ev[0], st);
	cudaEventRecord(c->ev[1], st);
	cudaEventRecord(c->ev[2], st);
	const int wpb = 4, nblk = (c->n_chunks + wpb - 1) / wpb;
	const int hot = (c->have_prev && c->warm_hot > 0) ? 1 : 0;
	const int wl = hot ? c->warm_hot : c->warm_len;
	// The backward warm-up needs only observations + model: it runs on the side stream, concurrently with the forward pass;
	// on small shards (multi-GPU) a long one would become the critical path, so it is capped at the forward kernel's length.
	// The side stream also computes, ahead of time, the operators of the chunks that failed in the previous E-step
	// (after the warm-up: forward operators, then backward operators).
	cudaEventRecord(c->ev_fork, st);
	cudaStreamWaitEvent(c->stream2, c->ev_fork, 0);
	const int wl_b = (c->warm_bwd_fixed || c->warm_len_bwd <= c->warm_len + c->chunk_len) ? c->warm_len_bwd : std::max(c->warm_len, c->warm_len + c->chunk_len);
	const int cap_b = wl_b;
	constexpr int G1 = (NP > 64) ? 16 : 8, SPL1 = NP / G1, COLS = 128 / G1;
	const dim3 gridT((unsigned)std::min(c->n_sub, 8 * c->sm_count), NP / COLS); // block rows stride over the sub-chunks
	const int nblk_b = (c->n_chunks_b + wpb - 1) / wpb;
	const dim3 gridTb((unsigned)std::min(c->n_sub_b, 8 * c->sm_count), NP / COLS);
	const int32_t *pred = c->predict ? c->d_pred[c->pred_cur] + 1 : nullptr, *pred_b = c->predict ? c->d_pred_b[c->pred_cur] + 1 : nullptr;
	int32_t *pred_next = c->d_pred[c->pred_cur ^ 1] + 1, *pred_next_b = c->d_pred_b[c->pred_cur ^ 1] + 1;
	auto side_k1f = [&]() {
		LAUNCH((k_transfer<SPL1, G1, COLS>), gridT, COLS * G1, c->stream2, c->d_sub, c->d_sub_parent, c->d_obs, c->d_model, c->d_Tsub, c->d_Texsub, c->N, pred, 3, nullptr, c->n_sub, c->d_chunk_sub0, 0, nullptr);
		cudaEventRecord(c->ev_k1f, c->stream2);
	};
	auto side_warm = [&]() {
		run_backward_warm<NP>(c, c->stream2, hot ? c->warm_hot : cap_b, hot);
		cudaEventRecord(c->ev_join, c->stream2);
	};
	side_warm();
	run_forward<NP>(c, wl, hot);
	cudaEventRecord(c->ev[6], st); // forward kernel done (the repair rounds follow)
	if (c->predict) side_k1f();
	if (c->predict) {
		LAUNCH((k_transfer<SPL1, G1, COLS>), gridTb, COLS * G1, c->stream2, c->d_sub_b, c->d_sub_parent_b, c->d_obs, c->d_model, c->d_Tsub_b, c->d_Texsub_b, c->N, pred_b, 3, nullptr, c->n_sub_b, c->d_chunk_sub0_b, 1, nullptr);
		cudaEventRecord(c->ev_k1b, c->stream2);
		cudaStreamWaitEvent(st, c->ev_k1f, 0);
	}
	for (int r = 0; r < rounds; ++r) {
		unsigned long long *rc_f = c->d_cert + 16 + r;
		LAUNCH((k_mark_fwd<SPL>), nblk, wpb * 32, st, c->d_chunks, c->n_chunks, c->N, c->cert_eps, c->d_fhat, c->d_fwarm, c->d_flag + 1, c->d_cert + 4, r == 0 ? pred_next : nullptr, c->d_cert + 8, r == 0 ? 0 : spread_f, r, rc_f);
		LAUNCH((k_transfer<SPL1, G1, COLS>), gridT, COLS * G1, st, c->d_sub, c->d_sub_parent, c->d_obs, c->d_model, c->d_Tsub, c->d_Texsub, c->N, c->d_flag + 1, 3, r == 0 ? pred : nullptr, c->n_sub, c->d_chunk_sub0, 0, rc_f);
		LAUNCH((k_chain_subs<NP>), c->n_chunks, NP, st, c->d_sub, c->n_sub, c->d_sub_parent, c->d_chunk_sub0, c->d_flag + 1, 0, c->d_Tsub, c->d_Texsub, c->d_fhat, c->d_bexact, c->d_vsub, c->d_bsub, rc_f);
		run_forward_repair<NP>(c, rc_f);
		LAUNCH((k_fold), c->n_chunks, 128, st, c->d_chunk_sub0, c->d_flag + 1, 0, NP, c->d_llsub, c->d_llpart, c->d_partsub, c->d_part, rc_f);
	}
	cudaEventRecord(c->ev[3], st);
	cudaStreamWaitEvent(st, c->ev_join, 0);
	run_backward<NP>(c, c->d_chunks_b, c->n_chunks_b, c->d_bwarm, 1, c->d_bsave[c->bsave_cur ^ 1]);
	cudaEventRecord(c->ev[7], st); // backward kernel done
	c->bsave_cur ^= 1;
	if (c->predict) cudaStreamWaitEvent(st, c->ev_k1b, 0);
	for (int r = 0; r < rounds; ++r) {
		unsigned long long *rc_b = c->d_cert + 16 + 64 + r;
		LAUNCH((k_mark_bwd<SPL>), nblk_b, wpb * 32, st, c->d_chunks_b, c->n_chunks_b, c->N, c->cert_eps, c->d_bwarm, c->d_bexact, c->d_flag_b + 1, c->d_cert + 4, r == 0 ? pred_next_b : nullptr, c->d_cert + 9, r == 0 ? 0 : spread_b, r, rc_b);
		LAUNCH((k_transfer<SPL1, G1, COLS>), gridTb, COLS * G1, st, c->d_sub_b, c->d_sub_parent_b, c->d_obs, c->d_model, c->d_Tsub_b, c->d_Texsub_b, c->N, c->d_flag_b + 1, 3, r == 0 ? pred_b : nullptr, c->n_sub_b, c->d_chunk_sub0_b, 1, rc_b);
		LAUNCH((k_chain_subs<NP>), c->n_chunks_b, NP, st, c->d_sub_b, c->n_sub_b, c->d_sub_parent_b, c->d_chunk_sub0_b, c->d_flag_b + 1, 1, c->d_Tsub_b, c->d_Texsub_b, c->d_fhat, c->d_bexact, c->d_vsub, c->d_bsub, rc_b);
		run_backward_repair<NP>(c, rc_b);
		LAUNCH((k_fold), c->n_chunks_b, 128, st, c->d_chunk_sub0_b, c->d_flag_b + 1, 1, NP, c->d_llsub, c->d_llpart, c->d_partsub, c->d_part, rc_b);
	}
	cudaEventRecord(c->ev[4], st);
	LAUNCH((k_reduce), dim3(1 + S_COUNT * c->N, c->n_rep), 256, st, c->d_part, c->d_llpart, c->d_rep_c0, c->d_rep_c0_b, c->N, NP, c->d_stats, c->weighted ? c->d_cw : nullptr, c->weighted ? c->d_cw_b : nullptr);
	LAUNCH((k_certify<SPL>), nblk, wpb * 32, st, c->d_chunks, c->n_chunks, c->N, c->d_fhat, c->d_fwarm, c->d_bwarm, c->d_bexact, c->cert_eps, 0, c->d_cert);
	LAUNCH((k_certify<SPL>), nblk_b, wpb * 32, st, c->d_chunks_b, c->n_chunks_b, c->N, c->d_fhat, c->d_fwarm, c->d_bwarm, c->d_bexact, c->cert_eps, 1, c->d_cert);
	cudaEventRecord(c->ev[5], st);
	if (c->probe_now && c->d_probeK) { // mixing probe next to the (now exact) forward pass; the result is read in sync_and_certify
		constexpr int GP = (NP > 64) ? 16 : 8;
		cudaMemsetAsync(c->d_probeK, 0, sizeof(float) * (size_t)c->n_probeK, st);
		LAUNCH((k_probe<NP / GP, GP>), blocks_for(c->n_chunks, GP), 128, st, c->d_chunks, c->n_chunks, c->d_obs, c->d_model, c->d_fhat, c->d_probeK, c->N, c->lead());
		cudaMemcpyAsync(c->h_probeK, c->d_probeK, sizeof(float) * (size_t)c->n_probeK, cudaMemcpyDeviceToHost, st);
	}
	c->launches = 6 + 10 * rounds + (c->predict ? 2 : 0) + ((c->probe_now && c->d_probeK) ? 1 : 0);
	c->pred_cur ^= 1;
	cudaError_t e = cudaGetLastError();
	if (e != cudaSuccess) return set_err(PSMC_B200_ECUDA, "kernel launch failed: %s", cudaGetErrorString(e));
	c->fwd_valid = false; // bend[] is not filled in this mode; decode runs its own forward pass
	c->mode_warm = true;
	c->certified = false;
	c->have_prev = true;
	return 0;
}

// the probe's pieces (k_probe) -> per sequence, the log contraction accumulated along a 64-bin grid
static void parse_probe(psmc_b200_ctx *c)
{
	if ((int)c->probe_cum.size() != c->n_seqs) c->probe_cum.assign((size_t)c->n_seqs, std::vector<double>());
	std::vector<char> touched((size_t)std::max(c->n_seqs, 1), 0);
	const float *K = c->h_probeK;
	int last_v = -1;
	for (int ci = 0; ci < c->n_chunks; ++ci) {
		const Chunk &ch = c->chunks[ci];
		const int seq = c->vseq_seq[ch.seq];
		std::vector<double> &cell = c->probe_cum[seq];
		if (ch.seq != last_v) { // a new virtual sequence: its record's cells start from zero (batch mode: the last model that drew a record wins)
			cell.assign((size_t)((c->L[seq] + 63) / 64) + 1, 0.0);
			touched[seq] = 1;
			last_v = ch.seq;
		}
		int64_t k = (ch.gb0 >> 6) + ci;
		int u = ch.u0;
		const int uend = ch.u0 + ch.len;
		while (u < uend) { // piece [u, ue]: up to the next global row with (row & 63) == 63, or the chunk's last bin
			const int64_t row = ch.gb0 + (u - ch.u0);
			const int ue = (int)std::min<int64_t>(uend - 1, u + (63 - (row & 63)));
			double v = (k < c->n_probeK) ? (double)K[k] : 0.0;
			if (!(v <= 0.0) || !std::isfinite(v)) v = 0.0;
			const double per_bin = v / (double)(ue - u + 1);
			for (int x = u; x <= ue;) { // spread over the record's own 64-bin cells
				const int xe = std::min(ue, x | 63);
				cell[(size_t)(x >> 6) + 1] += per_bin * (double)(xe - x + 1);
				x = xe + 1;
			}
			u = ue + 1;
			++k;
		}
	}
	for (int i = 0; i < c->n_seqs; ++i) { // cells -> prefix sums: cum[j] = accumulated over the bins < 64 j
		if (!touched[i]) continue;
		std::vector<double> &cum = c->probe_cum[i];
		for (size_t j = 1; j < cum.size(); ++j) cum[j] += cum[j - 1];
	}
	c->have_probe = true;
	c->plan_dirty = true;
}

static int launch_dispatch(psmc_b200_ctx *c, bool with_counts)
{
	c->mode_warm = false;
	c->certified = true;
	if (with_counts && c->warm_len > 0 && c->n_k1 > 0) {
		switch (c->SPL) {
		case 1: return launch_warm<1>(c);
		case 2: return launch_warm<2>(c);
		case 4: return launch_warm<4>(c);
		}
	}
	switch (c->SPL) {
	case 1: return launch_core<1>(c, with_counts);
	case 2: return launch_core<2>(c, with_counts);
	case 4: return launch_core<4>(c, with_counts);
	}
	return set_err(PSMC_B200_EINVAL, "internal: bad SPL");
}

static int launch_models(psmc_b200_ctx *c, int n_rep, const psmc_b200_model *models)
{
	if (!c) return set_err(PSMC_B200_EINVAL, "ctx is NULL");
	if (n_rep != c->n_rep) return set_err(PSMC_B200_EINVAL, "%d model(s) given, the context is set up for %d (psmc_b200_set_batch)", n_rep, c->n_rep);
	if (!models) return set_err(PSMC_B200_EINVAL, "models is NULL");
	for (int r = 0; r < n_rep; ++r) {
		int rc = check_model(c, models + r);
		if (rc) return rc;
	}
	CUDA_TRY(cudaSetDevice(c->device), PSMC_B200_ECUDA);
	CUDA_TRY(cudaStreamSynchronize(c->stream), PSMC_B200_ECUDA); // pinned staging buffer is reused
	if (c->plan_dirty) { // a mixing probe came back with the previous E-step: new plans with the overlaps it asks for
		c->plan_dirty = false;
		int rc = replan(c);
		if (rc) return rc;
	}
	// probe on the E-steps 0, 2, 8, 32, 128, ... of a context (the model moves fast early in EM, slowly later; a probe costs
	// about as much as the forward kernel, a stale plan a few more repaired boundaries)
	bool due = c->estep_no == 0;
	for (int e = 2; e <= c->estep_no && !due; e *= 4) due = (e == c->estep_no);
	if (due && c->estep_no == 0 && c->have_probe && (int)c->probe_cum.size() == c->n_seqs) { // (batch mode: records probed by an earlier batch keep their plan data)
		bool all = true;
		for (int v = 0; v < c->n_vseq && all; ++v) all = !c->probe_cum[c->vseq_seq[v]].empty();
		if (all) due = false;
	}
	c->probe_now = c->probe_on && (!c->batch || c->probe_batch) && c->warm_len > 0 && c->n_k1 > 0 && c->chunk_len_req <= 0 && due;
	if (c->probe_now) {
		const int64_t need = (c->cap_bins >> 6) + c->n_chunks + 2;
		if (need > c->cap_probeK) {
			cudaFree(c->d_probeK);
			if (c->h_probeK) cudaFreeHost(c->h_probeK);
			c->d_probeK = c->h_probeK = nullptr;
			c->cap_probeK = 0;
			if (cudaMalloc((void **)&c->d_probeK, sizeof(float) * (size_t)need) == cudaSuccess && cudaMallocHost((void **)&c->h_probeK, sizeof(float) * (size_t)need) == cudaSuccess) c->cap_probeK = need;
			else { cudaGetLastError(); c->probe_now = false; }
		}
		c->n_probeK = need;
	}
	++c->estep_no;
	for (int r = 0; r < n_rep; ++r) stage_model(c, models + r, r);
	CUDA_TRY(cudaMemcpyAsync(c->d_model, c->h_model, sizeof(double) * M_COUNT * c->NP * (size_t)n_rep, cudaMemcpyHostToDevice, c->stream), PSMC_B200_ECUDA);
	int rc = launch_dispatch(c, true);
	if (rc) return rc;
	c->launched = true;
	c->dense_valid = c->dense;
	return 0;
}

extern "C" int psmc_b200_estep_launch(psmc_b200_ctx *c, const psmc_b200_model *model)
{
	if (c && c->batch) return set_err(PSMC_B200_EINVAL, "the context is in batch mode: use psmc_b200_estep_batch* (or psmc_b200_set_multiplicity to leave it)");
	return launch_models(c, 1, model);
}

extern "C" int psmc_b200_estep_batch_launch(psmc_b200_ctx *c, int32_t n_rep, const psmc_b200_model *models)
{
	return launch_models(c, n_rep, models);
}

extern "C" void *psmc_b200_device_stats(psmc_b200_ctx *c) { return c ? (void *)c->d_stats : nullptr; }
extern "C" int psmc_b200_stats_len(const psmc_b200_ctx *c) { return c ? S_COUNT * c->N + 1 : 0; }
extern "C" void *psmc_b200_stream(psmc_b200_ctx *c) { return c ? (void *)c->stream : nullptr; }

static void collect_times(psmc_b200_ctx *c, bool with_counts)
{
	for (int i = 0; i < 8; ++i) c->ms[i] = 0.f;
	cudaEventElapsedTime(&c->ms[0], c->ev[0], c->ev[1]);
	cudaEventElapsedTime(&c->ms[1], c->ev[1], c->ev[2]);
	cudaEventElapsedTime(&c->ms[2], c->ev[2], c->ev[3]);
	if (with_counts) {
		cudaEventElapsedTime(&c->ms[3], c->ev[3], c->ev[4]);
		cudaEventElapsedTime(&c->ms[4], c->ev[4], c->ev[5]);
		cudaEventElapsedTime(&c->ms[5], c->ev[0], c->ev[5]);
		if (c->mode_warm) { // fast path: the two chunk kernels on their own (ms[2] / ms[3] include the repair rounds)
			cudaEventElapsedTime(&c->ms[6], c->ev[2], c->ev[6]);
			cudaEventElapsedTime(&c->ms[7], c->ev[3], c->ev[7]);
		}
	} else {
		cudaEventElapsedTime(&c->ms[5], c->ev[0], c->ev[3]);
	}
}

// Wait for the stream; in warm-up mode read the certificate and, if any boundary failed it, redo the
// E-step with the exact transfer-matrix path (same stream, same output buffers).
static int sync_and_certify(psmc_b200_ctx *c)
{
	const bool need = c->mode_warm && !c->certified;
	if (need) CUDA_TRY(cudaMemcpyAsync(c->h_cert, c->d_cert, sizeof(unsigned long long) * 16, cudaMemcpyDeviceToHost, c->stream), PSMC_B200_ECUDA);
	CUDA_TRY(cudaStreamSynchronize(c->stream), PSMC_B200_ECUDA);
	if (need && c->probe_now && c->h_probeK) parse_probe(c);
	c->probe_now = false;
	if (need) {
		c->certified = true;
		long long bf = (long long)c->h_cert[1], bb = (long long)c->h_cert[2];
		memcpy(&c->mis_f, &bf, sizeof(double));
		memcpy(&c->mis_b, &bb, sizeof(double));
		c->rep_fwd_fail = (long long)c->h_cert[4]; c->rep_fwd_chunks = (long long)c->h_cert[5];
		c->rep_bwd_fail = (long long)c->h_cert[6]; c->rep_bwd_chunks = (long long)c->h_cert[7];
		if (c->h_cert[15] > 0) return set_err(PSMC_B200_ECUDA, "internal: a staged tile of the forward spill never arrived (%llu waits timed out)", (unsigned long long)c->h_cert[15]);
		const bool cert_failed = c->h_cert[0] > 0;
		if (!c->rounds_fixed) { // the next E-step enqueues two rounds more than the deepest round that still saw a failure
			const int deepest = (int)std::max(c->h_cert[8], c->h_cert[9]);
			c->rounds_cur = std::min(c->rounds_max, std::max(c->repair_rounds, deepest + 2));
		}
		if (cert_failed && !c->redoing && !c->rounds_fixed && c->repair_rounds > 0) {
			// the cascade was deeper than the rounds enqueued: once more on the fast path with as many rounds as allowed
			++c->warm_redos;
			c->redoing = true;
			c->rounds_cur = c->rounds_max;
			int rc = launch_dispatch(c, true);
			if (rc == 0) rc = sync_and_certify(c);
			c->redoing = false;
			return rc;
		}
		if (cert_failed) {
			++c->fallbacks;
			const int keep = c->warm_len;
			c->warm_len = 0;
			int rc = launch_dispatch(c, true);
			c->warm_len = keep;
			if (rc) return rc;
			CUDA_TRY(cudaStreamSynchronize(c->stream), PSMC_B200_ECUDA);
		}
	}
	return 0;
}

extern "C" int psmc_b200_wait(psmc_b200_ctx *c)
{
	if (!c) return set_err(PSMC_B200_EINVAL, "ctx is NULL");
	CUDA_TRY(cudaSetDevice(c->device), PSMC_B200_ECUDA);
	int rc = sync_and_certify(c);
	if (rc) return rc;
	if (c->launched) collect_times(c, true);
	return 0;
}

extern "C" int psmc_b200_unpack_stats(int32_t N, const double *raw, int64_t n_seqs_total, psmc_b200_stats *out)
{
	if (!raw || !out || !out->E || !out->RL || !out->CL || !out->RU || !out->CU || !out->AD)
		return set_err(PSMC_B200_EINVAL, "NULL output arrays");
	for (int i = 0; i < S_COUNT * N + 1; ++i)
		if (!isfinite(raw[i])) return set_err(PSMC_B200_ENUMERIC, "non-finite value in the E-step statistics (entry %d)", i);
	// every sequence contributes HMM_TINY to every A[k][l] and E[b][l] (khmm.c:305-308, summed by khmm.c:346-359)
	const double tiny = (double)n_seqs_total * HMM_TINY_;
	out->LL = raw[0];
	const double *p = raw + 1;
	for (int k = 0; k < N; ++k) {
		out->E[k] = p[S_E0 * N + k] + tiny;
		out->E[N + k] = p[S_E1 * N + k] + tiny;
		out->RL[k] = p[S_RL * N + k] + tiny * k;
		out->CL[k] = p[S_CL * N + k] + tiny * (N - 1 - k);
		out->RU[k] = p[S_RU * N + k] + tiny * (N - 1 - k);
		out->CU[k] = p[S_CU * N + k] + tiny * k;
		out->AD[k] = p[S_AD * N + k] + tiny;
	}
	return 0;
}

static int fetch_stats(psmc_b200_ctx *c)
{
	if (!c->launched) return set_err(PSMC_B200_EINVAL, "estep finish/fetch without a launch");
	CUDA_TRY(cudaSetDevice(c->device), PSMC_B200_ECUDA);
	const size_t n = (size_t)(S_COUNT * c->N + 1) * (size_t)c->n_rep;
	int rc = sync_and_certify(c);
	if (rc) return rc;
	CUDA_TRY(cudaMemcpyAsync(c->h_stats, c->d_stats, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream), PSMC_B200_ECUDA);
	CUDA_TRY(cudaStreamSynchronize(c->stream), PSMC_B200_ECUDA);
	collect_times(c, true);
	c->launched = false;
	return 0;
}

extern "C" int psmc_b200_estep_finish(psmc_b200_ctx *c, int64_t n_seqs_total, psmc_b200_stats *out)
{
	if (!c) return set_err(PSMC_B200_EINVAL, "ctx is NULL");
	if (c->batch) return set_err(PSMC_B200_EINVAL, "the context is in batch mode: use psmc_b200_estep_batch_finish");
	int rc = fetch_stats(c);
	if (rc) return rc;
	if (n_seqs_total < 0) n_seqs_total = c->n_seq_eff;
	return psmc_b200_unpack_stats(c->N, c->h_stats, n_seqs_total, out);
}

extern "C" int psmc_b200_estep_batch_finish(psmc_b200_ctx *c, int32_t n_rep, psmc_b200_stats *outs)
{
	if (!c || !outs) return set_err(PSMC_B200_EINVAL, "NULL argument");
	if (n_rep != c->n_rep) return set_err(PSMC_B200_EINVAL, "%d result(s) asked, the context is set up for %d model(s)", n_rep, c->n_rep);
	int rc = fetch_stats(c);
	if (rc) return rc;
	const int len = S_COUNT * c->N + 1;
	for (int r = 0; r < n_rep; ++r) { // every model adds the HMM_TINY terms of ITS drawn records (khmm.c:305-308)
		rc = psmc_b200_unpack_stats(c->N, c->h_stats + (size_t)r * len, c->rep_seq_eff[r], outs + r);
		if (rc) return rc;
	}
	return 0;
}

extern "C" int psmc_b200_estep_batch(psmc_b200_ctx *c, int32_t n_rep, const psmc_b200_model *models, psmc_b200_stats *outs)
{
	int rc = psmc_b200_estep_batch_launch(c, n_rep, models);
	if (rc) return rc;
	return psmc_b200_estep_batch_finish(c, n_rep, outs);
}

extern "C" int psmc_b200_estep_fetch_raw(psmc_b200_ctx *c, double *raw)
{
	if (!c || !raw) return set_err(PSMC_B200_EINVAL, "NULL argument");
	if (c->batch) return set_err(PSMC_B200_EINVAL, "the context is in batch mode: use psmc_b200_estep_batch_finish");
	int rc = fetch_stats(c);
	if (rc) return rc;
	memcpy(raw, c->h_stats, sizeof(double) * (size_t)(S_COUNT * c->N + 1));
	return 0;
}

extern "C" int psmc_b200_estep(psmc_b200_ctx *c, const psmc_b200_model *model, psmc_b200_stats *out)
{
	int rc = psmc_b200_estep_launch(c, model);
	if (rc) return rc;
	return psmc_b200_estep_finish(c, c->n_seq_eff, out);
}

extern "C" int psmc_b200_factorize(int32_t N, const double *a, double tol, double *U, double *V, double *W, double *Z, double *D)
{
	if (N < 1 || !a || !U || !V || !W || !Z || !D) return set_err(PSMC_B200_EINVAL, "bad arguments");
	// V = last row, Z = first row, U = column 0 / a[N-1][0], W = last column / a[0][N-1]  
	for (int k = 0; k < N; ++k) {
		D[k] = a[(size_t)k * N + k];
		V[k] = a[(size_t)(N - 1) * N + k];
		Z[k] = a[k];
		U[k] = (N > 1) ? a[(size_t)k * N] / a[(size_t)(N - 1) * N] : 0.0;
		W[k] = (N > 1) ? a[(size_t)k * N + N - 1] / a[N - 1] : 0.0;
	}
	U[0] = 0.0; W[N - 1] = 0.0; V[N - 1] = 0.0; Z[0] = 0.0;
	for (int k = 0; k < N; ++k)
		for (int l = 0; l < N; ++l) {
			if (l == k) continue;
			const double r = (l < k) ? U[k] * V[l] : W[k] * Z[l];
			const double v = a[(size_t)k * N + l];
			if (!(fabs(r - v) <= tol * fabs(v) + 1e-300))
				return set_err(PSMC_B200_ESTRUCT, "transition matrix is not diagonal + rank-1 lower + rank-1 upper at (%d,%d): %g vs %g", k, l, v, r);
		}
	return 0;
}

extern "C" int psmc_b200_estep_dense(psmc_b200_ctx *c, int32_t N, const double *a0, const double *a, const double *e,
                                     double tol, psmc_b200_stats *out)
{
	if (!c) return set_err(PSMC_B200_EINVAL, "ctx is NULL");
	if (N != c->N) return set_err(PSMC_B200_EINVAL, "model has %d states, context has %d", N, c->N);
	std::vector<double> f((size_t)5 * N);
	int rc = psmc_b200_factorize(N, a, tol, &f[0], &f[N], &f[2 * N], &f[3 * N], &f[4 * N]);
	if (rc) return rc;
	psmc_b200_model m;
	m.n_states = N; m.a0 = a0; m.e = e;
	m.U = &f[0]; m.V = &f[N]; m.W = &f[2 * N]; m.Z = &f[3 * N]; m.D = &f[4 * N];
	return psmc_b200_estep(c, &m, out);
}

extern "C" int psmc_b200_decode(psmc_b200_ctx *c, const psmc_b200_model *model, int32_t seq_id, int32_t *best_k,
                                double *best_p, double *post, double *p_recomb, double *s_out)
{
	if (!c) return set_err(PSMC_B200_EINVAL, "ctx is NULL");
	if (c->batch) return set_err(PSMC_B200_EINVAL, "decode needs single-model mode (psmc_b200_set_multiplicity leaves batch mode)");
	// seq_id indexes the records AS GIVEN to create (empty records included), like psmc_b200_set_multiplicity
	if (seq_id < 0 || seq_id >= c->n_seqs_given) return set_err(PSMC_B200_EINVAL, "seq_id out of range");
	if (c->kept_of[seq_id] < 0) return set_err(PSMC_B200_EINVAL, "record %d is empty: nothing to decode", seq_id);
	if (!best_k || !best_p) return set_err(PSMC_B200_EINVAL, "best_k/best_p are NULL");
	const int given_id = seq_id;
	seq_id = c->kept_of[seq_id];
	if (c->mult[seq_id] <= 0) return set_err(PSMC_B200_EINVAL, "sequence %d has multiplicity 0 (psmc_b200_set_multiplicity)", given_id);
	int rc = 0;
	if (model) {
		rc = check_model(c, model);
		if (rc) return rc;
	} else if (!c->fwd_valid) {
		return set_err(PSMC_B200_EINVAL, "decode with model == NULL needs a previous estep/decode on this context");
	}
	CUDA_TRY(cudaSetDevice(c->device), PSMC_B200_ECUDA);
	CUDA_TRY(cudaStreamSynchronize(c->stream), PSMC_B200_ECUDA);
	const int Ls = c->L[seq_id];
	if (c->dec_cap < Ls) {
		cudaFree(c->d_bestk); cudaFree(c->d_bestp); cudaFree(c->d_prec);
		c->d_bestk = nullptr; c->d_bestp = nullptr; c->d_prec = nullptr;
		CUDA_TRY(cudaMalloc((void **)&c->d_bestk, sizeof(int32_t) * (size_t)Ls), PSMC_B200_ECUDA);
		CUDA_TRY(cudaMalloc((void **)&c->d_bestp, sizeof(double) * (size_t)Ls), PSMC_B200_ECUDA);
		CUDA_TRY(cudaMalloc((void **)&c->d_prec, sizeof(double) * (size_t)Ls), PSMC_B200_ECUDA);
		c->dec_cap = Ls;
	}
	if (post && c->dec_post_cap < (int64_t)Ls * c->N) {
		cudaFree(c->d_post); c->d_post = nullptr;
		CUDA_TRY(cudaMalloc((void **)&c->d_post, sizeof(double) * (size_t)Ls * c->N), PSMC_B200_ECUDA);
		c->dec_post_cap = (int64_t)Ls * c->N;
	}
	c->launches = 0;
	if (model) { // forward for every sequence (the chunk plan is global); later calls may pass model == NULL to reuse it
		stage_model(c, model);
		CUDA_TRY(cudaMemcpyAsync(c->d_model, c->h_model, sizeof(double) * M_COUNT * c->NP, cudaMemcpyHostToDevice, c->stream), PSMC_B200_ECUDA);
		rc = launch_dispatch(c, false);
		if (rc) return rc;
	}
	const int c0 = c->seq_c0[seq_id], nc = c->seq_nc[seq_id];
	const int wpb = 4, nblk = (nc + wpb - 1) / wpb;
	// per-sequence outputs are indexed by the sequence-local bin
	const double *fh = c->d_fhat;
	const double *scp = c->d_sc;
	switch (c->SPL) {
	case 1: LAUNCH((k_decode<1>), nblk, wpb * 32, c->stream, c->d_chunks, c0, nc, c->d_obs, c->d_model, c->d_bend, fh, scp, c->N, c->d_bestk, c->d_bestp, post ? c->d_post : nullptr, p_recomb ? c->d_prec : nullptr); break;
	case 2: LAUNCH((k_decode<2>), nblk, wpb * 32, c->stream, c->d_chunks, c0, nc, c->d_obs, c->d_model, c->d_bend, fh, scp, c->N, c->d_bestk, c->d_bestp, post ? c->d_post : nullptr, p_recomb ? c->d_prec : nullptr); break;
	case 4: LAUNCH((k_decode<4>), nblk, wpb * 32, c->stream, c->d_chunks, c0, nc, c->d_obs, c->d_model, c->d_bend, fh, scp, c->N, c->d_bestk, c->d_bestp, post ? c->d_post : nullptr, p_recomb ? c->d_prec : nullptr); break;
	}
	++c->launches;
	CUDA_TRY(cudaGetLastError(), PSMC_B200_ECUDA);
	CUDA_TRY(cudaMemcpyAsync(best_k, c->d_bestk, sizeof(int32_t) * (size_t)Ls, cudaMemcpyDeviceToHost, c->stream), PSMC_B200_ECUDA);
	CUDA_TRY(cudaMemcpyAsync(best_p, c->d_bestp, sizeof(double) * (size_t)Ls, cudaMemcpyDeviceToHost, c->stream), PSMC_B200_ECUDA);
	if (post) CUDA_TRY(cudaMemcpyAsync(post, c->d_post, sizeof(double) * (size_t)Ls * c->N, cudaMemcpyDeviceToHost, c->stream), PSMC_B200_ECUDA);
	if (p_recomb) CUDA_TRY(cudaMemcpyAsync(p_recomb, c->d_prec, sizeof(double) * (size_t)Ls, cudaMemcpyDeviceToHost, c->stream), PSMC_B200_ECUDA);
	if (s_out) CUDA_TRY(cudaMemcpyAsync(s_out, c->d_sc + c->seq_gb0[seq_id], sizeof(double) * (size_t)Ls, cudaMemcpyDeviceToHost, c->stream), PSMC_B200_ECUDA);
	CUDA_TRY(cudaStreamSynchronize(c->stream), PSMC_B200_ECUDA);
	if (model) collect_times(c, false);
	return 0;
}

// ---- decoding of every sequence of the context on the fast path ------------------------------------------------------
template <int NP>
static void launch_decode2(psmc_b200_ctx *c, const Chunk *chunks, int n, const double *bdir, uint32_t what)
{
	constexpr int G = (NP == 32) ? 8 : 16, SPL = NP / G;
	const bool bins = (what & (PSMC_B200_DEC_BINS | PSMC_B200_DEC_POST)) != 0, post = (what & PSMC_B200_DEC_POST) != 0, runs = (what & PSMC_B200_DEC_RUNS) != 0;
	LAUNCH((k_decode2<SPL, G>), blocks_for(n, G), 128, c->stream, chunks, 0, n, c->d_obs, c->d_model, bdir, c->d_fhat, c->d_sc, c->N, (int64_t)0,
	       bins ? c->d2_bestk : nullptr, bins ? c->d2_bestp : nullptr, post ? c->d2_post : nullptr, post ? c->d2_prec : nullptr,
	       runs ? c->d2_run_start : nullptr, runs ? c->d2_run_state : nullptr, runs ? c->d2_run_maxp : nullptr, runs ? c->d2_run_count : nullptr);
}

extern "C" int psmc_b200_decode_run(psmc_b200_ctx *c, const psmc_b200_model *model, uint32_t what)
{
	if (!c) return set_err(PSMC_B200_EINVAL, "ctx is NULL");
	if (c->batch) return set_err(PSMC_B200_EINVAL, "decode needs single-model mode (psmc_b200_set_multiplicity leaves batch mode)");
	if (!(what & (PSMC_B200_DEC_RUNS | PSMC_B200_DEC_BINS | PSMC_B200_DEC_POST))) return set_err(PSMC_B200_EINVAL, "nothing asked for");
	const auto t_wall = std::chrono::steady_clock::now();
	// 1. a complete E-step on this model: exact forward spill + certified boundary directions (counts are a by-product)
	int rc = launch_models(c, 1, model);
	if (rc) return rc;
	CUDA_TRY(cudaSetDevice(c->device), PSMC_B200_ECUDA);
	rc = sync_and_certify(c);
	if (rc) return rc;
	collect_times(c, true);
	c->launched = false;
	c->dec_ms[0] = c->ms[5];
	c->dec_what = 0;
	c->r_seq.clear(); c->r_start.clear(); c->r_len.clear(); c->r_state.clear(); c->r_maxp.clear();
	// which plan holds the certified directions: the fast path's backward plan (bwarm), else the boundary chains of the main plan (bend)
	const Chunk *chunks = c->mode_warm ? c->d_chunks_b : c->d_chunks;
	const int n = c->mode_warm ? c->n_chunks_b : c->n_chunks;
	const double *bdir = c->mode_warm ? c->d_bwarm : c->d_bend;
	const int64_t rows = std::max<int64_t>(c->cap_bins, 1);
#define DALLOC(ptr, type, count)                                                                                      \
	do {                                                                                                              \
		cudaFree(ptr); ptr = nullptr;                                                                                 \
		CUDA_TRY(cudaMalloc((void **)&(ptr), sizeof(type) * (size_t)(count)), PSMC_B200_ECUDA);                       \
	} while (0)
	if ((what & (PSMC_B200_DEC_BINS | PSMC_B200_DEC_POST)) && c->d2_rows < rows) {
		DALLOC(c->d2_bestk, uint8_t, rows); DALLOC(c->d2_bestp, float, rows);
		c->d2_rows = rows;
	}
	if ((what & PSMC_B200_DEC_POST) && c->d2_rows_post < rows) {
		DALLOC(c->d2_post, float, (size_t)rows * c->N); DALLOC(c->d2_prec, double, rows);
		c->d2_rows_post = rows;
	}
	if ((what & PSMC_B200_DEC_RUNS) && c->d2_rows_runs < rows) {
		DALLOC(c->d2_run_start, int32_t, rows); DALLOC(c->d2_run_state, uint8_t, rows); DALLOC(c->d2_run_maxp, double, rows);
		c->d2_rows_runs = rows;
	}
	const int nmax = std::max(std::max(c->n_chunks, c->n_chunks_b), 1);
	if ((what & PSMC_B200_DEC_RUNS) && c->d2_cap_chunks < nmax) {
		DALLOC(c->d2_run_count, int32_t, nmax); DALLOC(c->d2_off, int64_t, nmax);
		c->d2_cap_chunks = nmax;
	}
	cudaEventRecord(c->ev[0], c->stream);
	if (n > 0) {
		switch (c->NP) {
		case 32: launch_decode2<32>(c, chunks, n, bdir, what); break;
		case 64: launch_decode2<64>(c, chunks, n, bdir, what); break;
		default: launch_decode2<128>(c, chunks, n, bdir, what); break;
		}
	}
	CUDA_TRY(cudaGetLastError(), PSMC_B200_ECUDA);
	if ((what & PSMC_B200_DEC_RUNS) && n > 0) {
		// 2. compact the runs: per-chunk counts -> offsets (host: a few thousand numbers) -> dense list -> host, merged across chunk boundaries
		std::vector<int32_t> cnt((size_t)n);
		std::vector<int64_t> off((size_t)n);
		CUDA_TRY(cudaMemcpyAsync(cnt.data(), c->d2_run_count, sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToHost, c->stream), PSMC_B200_ECUDA);
		CUDA_TRY(cudaStreamSynchronize(c->stream), PSMC_B200_ECUDA);
		int64_t tot = 0;
		for (int i = 0; i < n; ++i) { off[i] = tot; tot += cnt[i]; }
		if (c->d2_cap_o < tot) {
			DALLOC(c->d2_o_seq, int32_t, tot); DALLOC(c->d2_o_start, int32_t, tot); DALLOC(c->d2_o_state, uint8_t, tot); DALLOC(c->d2_o_maxp, double, tot);
			c->d2_cap_o = tot;
		}
		CUDA_TRY(cudaMemcpyAsync(c->d2_off, off.data(), sizeof(int64_t) * (size_t)n, cudaMemcpyHostToDevice, c->stream), PSMC_B200_ECUDA);
		LAUNCH((k_runs_gather), std::min(n, 8 * c->sm_count), 128, c->stream, chunks, n, (int64_t)0, c->d2_run_count, c->d2_off, c->d2_run_start, c->d2_run_state,
		       c->d2_run_maxp, c->d2_o_seq, c->d2_o_start, c->d2_o_state, c->d2_o_maxp);
		std::vector<int32_t> q((size_t)tot), st((size_t)tot);
		std::vector<uint8_t> ks((size_t)tot);
		std::vector<double> mp((size_t)tot);
		CUDA_TRY(cudaMemcpyAsync(q.data(), c->d2_o_seq, sizeof(int32_t) * (size_t)tot, cudaMemcpyDeviceToHost, c->stream), PSMC_B200_ECUDA);
		CUDA_TRY(cudaMemcpyAsync(st.data(), c->d2_o_start, sizeof(int32_t) * (size_t)tot, cudaMemcpyDeviceToHost, c->stream), PSMC_B200_ECUDA);
		CUDA_TRY(cudaMemcpyAsync(ks.data(), c->d2_o_state, sizeof(uint8_t) * (size_t)tot, cudaMemcpyDeviceToHost, c->stream), PSMC_B200_ECUDA);
		CUDA_TRY(cudaMemcpyAsync(mp.data(), c->d2_o_maxp, sizeof(double) * (size_t)tot, cudaMemcpyDeviceToHost, c->stream), PSMC_B200_ECUDA);
		cudaEventRecord(c->ev[1], c->stream);
		CUDA_TRY(cudaStreamSynchronize(c->stream), PSMC_B200_ECUDA);
		std::vector<int32_t> given_of((size_t)std::max(c->n_seqs, 1), -1);
		for (int i = 0; i < c->n_seqs_given; ++i)
			if (c->kept_of[i] >= 0) given_of[c->kept_of[i]] = i;
		for (int64_t i = 0; i < tot; ++i) { // a run that continues across a chunk boundary was emitted in two pieces
			if (!c->r_seq.empty() && c->r_seq.back() == given_of[q[i]] && c->r_state.back() == ks[i]) {
				if (mp[i] > c->r_maxp.back()) c->r_maxp.back() = mp[i];
				continue;
			}
			c->r_seq.push_back(given_of[q[i]]); c->r_start.push_back(st[i]); c->r_state.push_back(ks[i]); c->r_maxp.push_back(mp[i]);
		}
		c->r_len.resize(c->r_seq.size());
		for (size_t i = 0; i < c->r_seq.size(); ++i) {
			const bool more = i + 1 < c->r_seq.size() && c->r_seq[i + 1] == c->r_seq[i];
			c->r_len[i] = (more ? c->r_start[i + 1] : c->L[c->kept_of[c->r_seq[i]]]) - c->r_start[i];
		}
	} else {
		cudaEventRecord(c->ev[1], c->stream);
		CUDA_TRY(cudaStreamSynchronize(c->stream), PSMC_B200_ECUDA);
	}
#undef DALLOC
	cudaEventElapsedTime(&c->dec_ms[1], c->ev[0], c->ev[1]);
	c->dec_ms[2] = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t_wall).count();
	c->dec_what = what;
	return 0;
}

extern "C" int psmc_b200_decode_get_runs(psmc_b200_ctx *c, int64_t cap, int32_t *seq_id, int32_t *start, int32_t *len, uint8_t *state,
                                         double *max_p, int64_t *n_runs)
{
	if (!c || !n_runs) return set_err(PSMC_B200_EINVAL, "NULL argument");
	if (!(c->dec_what & PSMC_B200_DEC_RUNS)) return set_err(PSMC_B200_EINVAL, "no runs: call psmc_b200_decode_run(ctx, model, PSMC_B200_DEC_RUNS) first");
	const int64_t n = (int64_t)c->r_seq.size();
	*n_runs = n;
	if (cap < n) return cap == 0 ? 0 : set_err(PSMC_B200_EINVAL, "capacity %lld < %lld runs", (long long)cap, (long long)n);
	if (n > 0 && (!seq_id || !start || !len || !state || !max_p)) return set_err(PSMC_B200_EINVAL, "NULL output arrays");
	for (int64_t i = 0; i < n; ++i) { seq_id[i] = c->r_seq[i]; start[i] = c->r_start[i]; len[i] = c->r_len[i]; state[i] = c->r_state[i]; max_p[i] = c->r_maxp[i]; }
	return 0;
}

extern "C" int psmc_b200_decode_get_bins(psmc_b200_ctx *c, int32_t seq_id, uint8_t *best_k, float *best_p, float *post, double *p_recomb)
{
	if (!c) return set_err(PSMC_B200_EINVAL, "ctx is NULL");
	if (seq_id < 0 || seq_id >= c->n_seqs_given) return set_err(PSMC_B200_EINVAL, "seq_id out of range");
	if (c->kept_of[seq_id] < 0) return set_err(PSMC_B200_EINVAL, "record %d is empty: nothing to decode", seq_id);
	const bool want_post = post || p_recomb;
	if (!(c->dec_what & (PSMC_B200_DEC_BINS | PSMC_B200_DEC_POST)) || (want_post && !(c->dec_what & PSMC_B200_DEC_POST)))
		return set_err(PSMC_B200_EINVAL, "not computed: call psmc_b200_decode_run with PSMC_B200_DEC_BINS / PSMC_B200_DEC_POST first");
	const int k = c->kept_of[seq_id];
	if (c->mult[k] <= 0) return set_err(PSMC_B200_EINVAL, "sequence %d has multiplicity 0 (psmc_b200_set_multiplicity)", seq_id);
	const size_t r0 = (size_t)c->seq_gb0[k], Ls = (size_t)c->L[k];
	CUDA_TRY(cudaSetDevice(c->device), PSMC_B200_ECUDA);
	if (best_k) CUDA_TRY(cudaMemcpyAsync(best_k, c->d2_bestk + r0, Ls, cudaMemcpyDeviceToHost, c->stream), PSMC_B200_ECUDA);
	if (best_p) CUDA_TRY(cudaMemcpyAsync(best_p, c->d2_bestp + r0, sizeof(float) * Ls, cudaMemcpyDeviceToHost, c->stream), PSMC_B200_ECUDA);
	if (post) CUDA_TRY(cudaMemcpyAsync(post, c->d2_post + r0 * c->N, sizeof(float) * Ls * c->N, cudaMemcpyDeviceToHost, c->stream), PSMC_B200_ECUDA);
	if (p_recomb) CUDA_TRY(cudaMemcpyAsync(p_recomb, c->d2_prec + r0, sizeof(double) * Ls, cudaMemcpyDeviceToHost, c->stream), PSMC_B200_ECUDA);
	CUDA_TRY(cudaStreamSynchronize(c->stream), PSMC_B200_ECUDA);
	return 0;
}

// ---- dense transition counts (option) ----
extern "C" int psmc_b200_set_dense(psmc_b200_ctx *c, int32_t on)
{
	if (!c) return set_err(PSMC_B200_EINVAL, "ctx is NULL");
	if (on && (c->NP > 64 || c->gen != 2))
		return set_err(PSMC_B200_EINVAL, "dense counts need the generation-2 backward kernels (at most 64 states)");
	CUDA_TRY(cudaSetDevice(c->device), PSMC_B200_ECUDA);
	if (on && c->batch) return set_err(PSMC_B200_EINVAL, "batch mode and dense counts exclude each other");
	if (on && !c->d_ghat) {
		const size_t bytes = (size_t)std::max<int64_t>(c->cap_bins, 1) * c->NP * sizeof(double);
		CUDA_TRY(cudaMalloc((void **)&c->d_ghat, bytes), PSMC_B200_ECUDA);
		CUDA_TRY(cudaMemsetAsync(c->d_ghat, 0, bytes, c->stream), PSMC_B200_ECUDA);
		if (!c->d_cdense) CUDA_TRY(cudaMalloc((void **)&c->d_cdense, sizeof(double) * (size_t)c->NP * c->NP), PSMC_B200_ECUDA);
		CUDA_TRY(cudaStreamSynchronize(c->stream), PSMC_B200_ECUDA);
		c->bytes_total += (int64_t)bytes;
		c->bytes_forward += (int64_t)bytes;
	}
	c->dense = on != 0;
	c->dense_valid = false;
	return 0;
}

template <int NP>
static void launch_dense(psmc_b200_ctx *c)
{
	LAUNCH((k_dense_chunk<NP>), c->n_chunks_b, 256, c->stream, c->d_chunks_b, c->d_fhat, c->d_ghat, c->d_cpart);
	LAUNCH((k_dense_reduce), NP * NP, 256, c->stream, c->d_cpart, c->n_chunks_b, NP * NP, c->weighted ? c->d_cw_b : nullptr, c->d_cdense);
}

extern "C" int psmc_b200_dense_counts(psmc_b200_ctx *c, double *A)
{
	if (!c || !A) return set_err(PSMC_B200_EINVAL, "NULL argument");
	if (!c->dense || !c->dense_valid) return set_err(PSMC_B200_EINVAL, "no dense rows: call psmc_b200_set_dense(ctx, 1) before the E-step");
	CUDA_TRY(cudaSetDevice(c->device), PSMC_B200_ECUDA);
	int rc = sync_and_certify(c);
	if (rc) return rc;
	const int NP = c->NP, N = c->N;
	if (c->n_chunks_b > c->cap_cpart) {
		cudaFree(c->d_cpart);
		c->d_cpart = nullptr;
		c->cap_cpart = 0;
		CUDA_TRY(cudaMalloc((void **)&c->d_cpart, sizeof(double) * (size_t)std::max(c->n_chunks_b, 1) * NP * NP), PSMC_B200_ECUDA);
		c->cap_cpart = c->n_chunks_b;
	}
	std::vector<double> C((size_t)NP * NP, 0.0);
	if (c->n_chunks_b > 0) {
		if (NP == 32) launch_dense<32>(c);
		else launch_dense<64>(c);
		cudaError_t e = cudaGetLastError();
		if (e != cudaSuccess) return set_err(PSMC_B200_ECUDA, "kernel launch failed: %s", cudaGetErrorString(e));
		CUDA_TRY(cudaMemcpyAsync(C.data(), c->d_cdense, sizeof(double) * (size_t)NP * NP, cudaMemcpyDeviceToHost, c->stream), PSMC_B200_ECUDA);
		CUDA_TRY(cudaStreamSynchronize(c->stream), PSMC_B200_ECUDA);
	}
	// A[k][l] = sum over records of ( HMM_TINY + sum_u f_{u-1}[k] a[k][l] e[l] b_u[l] )   (khmm.c:305-316, 346-352)
	const double *hm = c->h_model, tiny = HMM_TINY_ * (double)c->n_seq_eff;
	for (int k = 0; k < N; ++k)
		for (int l = 0; l < N; ++l) {
			const double a = l < k ? hm[M_U * NP + k] * hm[M_V * NP + l] : (l > k ? hm[M_W * NP + k] * hm[M_Z * NP + l] : hm[M_D * NP + k]);
			A[(size_t)k * N + l] = tiny + a * C[(size_t)k * NP + l];
		}
	return 0;
}

extern "C" int psmc_b200_set_warm(psmc_b200_ctx *c, int32_t warm_len, double eps)
{
	if (!c) return set_err(PSMC_B200_EINVAL, "ctx is NULL");
	if (warm_len >= 0) {
		c->warm_len = warm_len;
		c->warm_len_bwd = warm_len + warm_len / 3;
		if (c->have_probe || c->planned) c->plan_dirty = true; // the planned overlaps are capped relative to the fixed one: re-plan
	}
	if (eps > 0) c->cert_eps = eps;
	return 0;
}

extern "C" int psmc_b200_get_info(const psmc_b200_ctx *c, psmc_b200_info *info)
{
	if (!c || !info) return set_err(PSMC_B200_EINVAL, "NULL argument");
	memset(info, 0, sizeof(*info));
	info->device = c->device;
	info->n_states = c->N;
	info->n_states_padded = c->NP;
	info->n_seqs = c->n_seqs;
	info->n_chunks = c->n_chunks;
	info->chunk_len = c->chunk_len;
	info->total_bins = c->total_bins;
	info->bytes_obs = c->bytes_obs;
	info->bytes_forward = c->bytes_forward;
	info->bytes_transfer = c->bytes_transfer;
	info->bytes_total = c->bytes_total;
	for (int i = 0; i < 8; ++i) info->ms[i] = c->ms[i];
	info->launches = c->launches;
	info->warm_len = c->warm_len;
	info->fallbacks = c->fallbacks;
	info->fwd_mismatch = c->mis_f;
	info->bwd_mismatch = c->mis_b;
	info->repaired_fwd = (int32_t)c->rep_fwd_chunks;
	info->repaired_bwd = (int32_t)c->rep_bwd_chunks;
	info->failed_fwd = (int32_t)c->rep_fwd_fail;
	info->failed_bwd = (int32_t)c->rep_bwd_fail;
	info->active_bins = c->active_bins;
	info->n_seqs_effective = c->n_seq_eff;
	info->n_models = c->n_rep;
	info->repair_rounds = c->rounds_cur;
	info->planned = c->planned ? 1 : 0;
	info->probe_plans = c->probe_plans;
	info->avg_overlap_fwd = (float)(c->planned ? c->avg_warm_f : (double)c->warm_len);
	info->avg_overlap_bwd = (float)(c->planned ? c->avg_warm_b : (double)c->warm_len_bwd);
	info->slow_fwd = c->planned ? c->slow_f : 0;
	info->slow_bwd = c->planned ? c->slow_b : 0;
	for (int i = 0; i < 3; ++i) info->decode_ms[i] = c->dec_ms[i];
	info->warm_redos = c->warm_redos;
	info->n_chunks_bwd = c->n_chunks_b;
	info->chunk_len_bwd = c->chunk_len_b;
	return 0;
}
