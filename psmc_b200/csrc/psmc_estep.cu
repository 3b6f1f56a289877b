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
#include <vector>
#include <algorithm>
#include <thread>

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
#define PIN_REG(x) asm volatile("" : "+d"(x))
#else
#define LAUNCH(kern, grid, block, stream, ...)                        \
	do {                                                              \
		auto kfn_ = PSMC_UNPAREN kern;                                \
		simt_emu::launch((grid), (block), [&]() { kfn_(__VA_ARGS__); }); \
	} while (0)
#define PIN_REG(x) ((void)(x))
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

// ------------------------------------------------------------------------------------------------
// device-side layout
// ------------------------------------------------------------------------------------------------
struct Chunk {
	int32_t seq;     // sequence id
	int32_t flags;   // bit0: first chunk of its sequence, bit1: last chunk of its sequence
	int32_t u0;      // first bin of the chunk, sequence-local
	int32_t len;     // bins in the chunk (>= 1)
	int64_t gb0;     // global bin index of u0 (rows of fhat / entries of sc)
	int64_t ow0;     // first packed-observation word of the SEQUENCE (16 bins per 32-bit word)
	int32_t Lseq;    // length of the sequence
	int32_t pad_;
};
#define CH_FIRST 1
#define CH_LAST 2

// model arrays on the device, each NP doubles, contiguous: a0 e0 e1 U V W Z D
enum { M_A0 = 0, M_E0, M_E1, M_U, M_V, M_W, M_Z, M_D, M_COUNT };
// statistics rows: E0 E1 RL CL RU CU AD
enum { S_E0 = 0, S_E1, S_RL, S_CL, S_RU, S_CU, S_AD, S_COUNT };

// ------------------------------------------------------------------------------------------------
// lane-group primitives: a vector of NP = G*SPL states is held by G consecutive lanes, SPL states
// per lane (state = gl*SPL + i).  G is a power of two <= 32.
// ------------------------------------------------------------------------------------------------
// Kogge-Stone scans over the lanes of a group.  A step is "x += (source lane inside my group) ? shuffled x : 0";
// written as an FMA with a per-lane 0/1 mask it costs 2 SHFL + 1 DFMA (the obvious if/select form compiles to
// 2 SHFL + DADD + 2 FSEL + moves and made the chunk kernels issue-bound: 231 instructions per bin, ncu).
template <int G>
struct ScanMasks {
	static constexpr int STEPS = (G == 32 ? 5 : G == 16 ? 4 : G == 8 ? 3 : G == 4 ? 2 : G == 2 ? 1 : 0);
	double up[STEPS + 1], dn[STEPS + 1]; // [k]: distance 2^k; [STEPS] is unused padding for G == 1
	__device__ __forceinline__ void init(int gl)
	{
#pragma unroll
		for (int k = 0; k < STEPS; ++k) {
			up[k] = (gl >= (1 << k)) ? 1.0 : 0.0;
			dn[k] = (gl + (1 << k) < G) ? 1.0 : 0.0;
		}
		up[STEPS] = dn[STEPS] = 0.0;
	}
	// keep the masks in registers: without this ptxas re-derives every mask from a predicate in every iteration
	// (ISETP + FSEL + MOV per scan stage); only for kernels with registers to spare
	__device__ __forceinline__ void pin()
	{
#pragma unroll
		for (int k = 0; k < STEPS; ++k) {
			PIN_REG(up[k]);
			PIN_REG(dn[k]);
		}
	}
};
template <int G>
__device__ __forceinline__ double gscan_up(double t, const ScanMasks<G> &m) // exclusive prefix over the lanes of a group
{
	if (G == 1) return 0.0;
	double x = __shfl_up_sync(FULLMASK, t, 1, G) * m.up[0];
#pragma unroll
	for (int k = 0; k < ScanMasks<G>::STEPS; ++k) x = fma(__shfl_up_sync(FULLMASK, x, 1 << k, G), m.up[k], x);
	return x;
}
template <int G>
__device__ __forceinline__ double gscan_down(double t, const ScanMasks<G> &m) // exclusive suffix over the lanes of a group
{
	if (G == 1) return 0.0;
	double x = __shfl_down_sync(FULLMASK, t, 1, G) * m.dn[0];
#pragma unroll
	for (int k = 0; k < ScanMasks<G>::STEPS; ++k) x = fma(__shfl_down_sync(FULLMASK, x, 1 << k, G), m.dn[k], x);
	return x;
}
template <int G>
__device__ __forceinline__ double gsum(double t) // all-reduce over the lanes of a group (same bits in every lane)
{
#pragma unroll
	for (int d = G >> 1; d > 0; d >>= 1) t += __shfl_xor_sync(FULLMASK, t, d, G);
	return t;
}
// reciprocal of a positive normal double: hardware seed (20 bits) + two Newton steps (2^-80 before rounding, ~1 ulp;
// the IEEE division drags a slow path and a range check into the loop)
__device__ __forceinline__ double fast_rcp(double s)
{
	double r;
#ifndef PSMC_SIMT_EMU
	asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(s));
#else
	r = simt_emu::rcp_seed(s);
#endif
	double e = fma(-s, r, 1.0);
	r = fma(r, e, r);
	e = fma(-s, r, 1.0);
	return fma(r, e, r);
}
template <int G>
__device__ __forceinline__ int gmax_i(int t)
{
#pragma unroll
	for (int d = G >> 1; d > 0; d >>= 1) t = max(t, __shfl_xor_sync(FULLMASK, t, d, G));
	return t;
}

// out[i] = D[i] x[i] + pc[i] * sum_{j<i} pm[j] x[j] + sc[i] * sum_{j>i} sm[j] x[j]   (indices over the whole group)
//   forward  (A^T f): pm = W, pc = Z, sm = U, sc = V
//   backward (A g)  : pm = V, pc = U, sm = Z, sc = W
template <int SPL, int G>
__device__ __forceinline__ void semisep(const double (&x)[SPL], const double (&pm)[SPL], const double (&pc)[SPL],
                                        const double (&sm)[SPL], const double (&sc)[SPL], const double (&D)[SPL],
                                        const ScanMasks<G> &mk, double (&out)[SPL])
{
	double tp = 0.0, ts = 0.0;
#pragma unroll
	for (int i = 0; i < SPL; ++i) {
		tp = fma(x[i], pm[i], tp);
		ts = fma(x[i], sm[i], ts);
	}
	double p = gscan_up<G>(tp, mk), s = gscan_down<G>(ts, mk);
#pragma unroll
	for (int i = 0; i < SPL; ++i) {
		out[i] = fma(pc[i], p, D[i] * x[i]);
		p = fma(x[i], pm[i], p);
	}
#pragma unroll
	for (int i = SPL - 1; i >= 0; --i) {
		out[i] = fma(sc[i], s, out[i]);
		s = fma(x[i], sm[i], s);
	}
}

// exclusive prefix P[i] = sum_{j<i} pm[j] x[j] and exclusive suffix S[i] = sum_{j>i} sm[j] x[j]
template <int SPL, int G>
__device__ __forceinline__ void prefsuf(const double (&x)[SPL], const double (&pm)[SPL], const double (&sm)[SPL],
                                        const ScanMasks<G> &mk, double (&P)[SPL], double (&S)[SPL])
{
	double tp = 0.0, ts = 0.0;
#pragma unroll
	for (int i = 0; i < SPL; ++i) {
		tp = fma(x[i], pm[i], tp);
		ts = fma(x[i], sm[i], ts);
	}
	double p = gscan_up<G>(tp, mk), s = gscan_down<G>(ts, mk);
#pragma unroll
	for (int i = 0; i < SPL; ++i) {
		P[i] = p;
		p = fma(x[i], pm[i], p);
	}
#pragma unroll
	for (int i = SPL - 1; i >= 0; --i) {
		S[i] = s;
		s = fma(x[i], sm[i], s);
	}
}

template <int SPL>
__device__ __forceinline__ void load_vec(const double *__restrict__ p, double (&v)[SPL])
{
	if (SPL == 1) {
		v[0] = __ldg(p);
	} else {
#pragma unroll
		for (int i = 0; i < SPL; i += 2) {
			double2 t = __ldg(reinterpret_cast<const double2 *>(p + i));
			v[i] = t.x;
			v[i + 1] = t.y;
		}
	}
}
template <int SPL>
__device__ __forceinline__ void store_vec(double *__restrict__ p, const double (&v)[SPL])
{
	if (SPL == 1) {
		p[0] = v[0];
	} else {
#pragma unroll
		for (int i = 0; i < SPL; i += 2) *reinterpret_cast<double2 *>(p + i) = make_double2(v[i], v[i + 1]);
	}
}

__device__ __forceinline__ int obs_at(const uint32_t *__restrict__ obs, int64_t ow0, int u)
{
	return (__ldg(obs + ow0 + (u >> 4)) >> ((u & 15) * 2)) & 3;
}

// Loop bounds of a warp-per-chunk kernel come from a per-thread load; broadcasting them from lane 0 lets ptxas
// prove they are warp-uniform, otherwise every shuffle in the loop is wrapped in WARPSYNC/ENDCOLLECTIVE (3x the
// instruction count, measured with ncu on B200).
__device__ __forceinline__ Chunk uniform_chunk(Chunk ch)
{
	ch.seq = __shfl_sync(FULLMASK, ch.seq, 0);
	ch.flags = __shfl_sync(FULLMASK, ch.flags, 0);
	ch.u0 = __shfl_sync(FULLMASK, ch.u0, 0);
	ch.len = __shfl_sync(FULLMASK, ch.len, 0);
	ch.Lseq = __shfl_sync(FULLMASK, ch.Lseq, 0);
	return ch;
}

// exact power-of-two rescale helpers: k = floor(log2(x)) for a normal positive x
__device__ __forceinline__ int exponent_of(double x) { return ((__double2hiint(x) >> 20) & 0x7ff) - 1023; }
__device__ __forceinline__ double pow2i(int k) { return __hiloint2double((1023 + k) << 20, 0); } // |k| < 1022

// ================================================================================================
// Second-generation lane-group primitives (the chunk kernels are bound by the LATENCY of the per-bin dependency
// chain at one or two warps per scheduler, not by issue slots or bandwidth; measured with ncu on B200):
//   * DualScan<G>: exclusive prefix of one value AND exclusive suffix of another over the G lanes of a group in ONE
//     (G = 8) or TWO (G = 16) shuffle rounds instead of 1 + log2 G Kogge-Stone rounds.  With xor-partners, the lane
//     pair (l, l ^ m) needs exactly one value from each other: the lower lane's prefix term goes up, the upper lane's
//     suffix term goes down, so one 64-bit shuffle serves both scans.
//   * local prefix/suffix sums as trees, and everything that does not depend on the shuffled values moved in front
//     of them: after the shuffles a state costs two FMAs.
// ================================================================================================
template <int G>
struct DualScan;

template <>
struct DualScan<8> {
	bool h0, h1, h2;
	double u0, u1, u2, d0, d1, d2; // u_b = 1 iff bit b of my lane is set: the partners that differ first in bit b are BELOW me
	__device__ __forceinline__ void init(int gl)
	{
		h0 = (gl & 1) != 0; h1 = (gl & 2) != 0; h2 = (gl & 4) != 0;
		u0 = h0 ? 1.0 : 0.0; u1 = h1 ? 1.0 : 0.0; u2 = h2 ? 1.0 : 0.0;
		d0 = 1.0 - u0; d1 = 1.0 - u1; d2 = 1.0 - u2;
		PIN_REG(u0); PIN_REG(u1); PIN_REG(u2); PIN_REG(d0); PIN_REG(d1); PIN_REG(d2);
	}
	// P = sum_{j<gl} tp_j,  S = sum_{j>gl} ts_j
	__device__ __forceinline__ void run(double tp, double ts, double &P, double &S) const
	{
		// a partner above me needs my prefix term and sends its suffix term; a partner below me the other way round
		const double s0 = h0 ? ts : tp, s1 = h1 ? ts : tp, s2 = h2 ? ts : tp;
		const double r1 = __shfl_xor_sync(FULLMASK, s0, 1, 8);
		const double r2 = __shfl_xor_sync(FULLMASK, s1, 2, 8), r3 = __shfl_xor_sync(FULLMASK, s1, 3, 8);
		const double r4 = __shfl_xor_sync(FULLMASK, s2, 4, 8), r5 = __shfl_xor_sync(FULLMASK, s2, 5, 8);
		const double r6 = __shfl_xor_sync(FULLMASK, s2, 6, 8), r7 = __shfl_xor_sync(FULLMASK, s2, 7, 8);
		const double q1 = r2 + r3, q2 = (r4 + r5) + (r6 + r7);
		P = fma(u2, q2, fma(u1, q1, u0 * r1));
		S = fma(d2, q2, fma(d1, q1, d0 * r1));
	}
};

template <>
struct DualScan<16> {
	bool h2, h3;
	double u0, u1, u2, u3, d0, d1, d2, d3;
	__device__ __forceinline__ void init(int gl)
	{
		h2 = (gl & 4) != 0; h3 = (gl & 8) != 0;
		u0 = (gl & 1) ? 1.0 : 0.0; u1 = (gl & 2) ? 1.0 : 0.0; u2 = h2 ? 1.0 : 0.0; u3 = h3 ? 1.0 : 0.0;
		d0 = 1.0 - u0; d1 = 1.0 - u1; d2 = 1.0 - u2; d3 = 1.0 - u3;
		PIN_REG(u0); PIN_REG(u1); PIN_REG(u2); PIN_REG(u3); PIN_REG(d0); PIN_REG(d1); PIN_REG(d2); PIN_REG(d3);
	}
	__device__ __forceinline__ void run(double tp, double ts, double &P, double &S) const
	{
		// round 1, inside quads of lanes: both terms travel, because the quad totals are needed as well
		const double p1 = __shfl_xor_sync(FULLMASK, tp, 1, 16), s1 = __shfl_xor_sync(FULLMASK, ts, 1, 16);
		const double p2 = __shfl_xor_sync(FULLMASK, tp, 2, 16), s2 = __shfl_xor_sync(FULLMASK, ts, 2, 16);
		const double p3 = __shfl_xor_sync(FULLMASK, tp, 3, 16), s3 = __shfl_xor_sync(FULLMASK, ts, 3, 16);
		const double qp = p2 + p3, qs = s2 + s3;
		const double Pq = fma(u1, qp, u0 * p1), Sq = fma(d1, qs, d0 * s1);
		const double Tp = (tp + p1) + qp, Ts = (ts + s1) + qs;
		// round 2, between quads: one value per partner quad, as in DualScan<8>
		const double t2 = h2 ? Ts : Tp, t3 = h3 ? Ts : Tp;
		const double r4 = __shfl_xor_sync(FULLMASK, t2, 4, 16);
		const double r8 = __shfl_xor_sync(FULLMASK, t3, 8, 16), r12 = __shfl_xor_sync(FULLMASK, t3, 12, 16);
		const double q3 = r8 + r12;
		P = fma(u3, q3, fma(u2, r4, Pq));
		S = fma(d3, q3, fma(d2, r4, Sq));
	}
};

// exclusive prefix sums of a[0..SPL) inside a lane (lp[i] = a[0] + .. + a[i-1]) and the total, as a tree: the total
// is log2 SPL additions deep
template <int SPL>
__device__ __forceinline__ void local_prefix(const double (&a)[SPL], double (&lp)[SPL], double &tot)
{
	if (SPL == 1) {
		lp[0] = 0.0;
		tot = a[0];
	} else if (SPL == 2) {
		lp[0] = 0.0;
		lp[1] = a[0];
		tot = a[0] + a[1];
	} else if (SPL == 4) {
		const double p01 = a[0] + a[1], p23 = a[2] + a[3];
		lp[0] = 0.0; lp[1] = a[0]; lp[2] = p01; lp[3] = p01 + a[2];
		tot = p01 + p23;
	} else if (SPL == 8) {
		const double p01 = a[0] + a[1], p23 = a[2] + a[3], p45 = a[4] + a[5], p67 = a[6] + a[7];
		const double q03 = p01 + p23, q47 = p45 + p67, q05 = q03 + p45;
		lp[0] = 0.0; lp[1] = a[0]; lp[2] = p01; lp[3] = p01 + a[2];
		lp[4] = q03; lp[5] = q03 + a[4]; lp[6] = q05; lp[7] = q05 + a[6];
		tot = q03 + q47;
	} else {
		double t = 0.0;
#pragma unroll
		for (int i = 0; i < SPL; ++i) {
			lp[i] = t;
			t += a[i];
		}
		tot = t;
	}
}
template <int SPL>
__device__ __forceinline__ void local_suffix(const double (&c)[SPL], double (&ls)[SPL], double &tot)
{
	double r[SPL], lr[SPL];
#pragma unroll
	for (int i = 0; i < SPL; ++i) r[i] = c[SPL - 1 - i];
	local_prefix<SPL>(r, lr, tot);
#pragma unroll
	for (int i = 0; i < SPL; ++i) ls[i] = lr[SPL - 1 - i];
}
template <int SPL>
__device__ __forceinline__ double local_sum(const double (&a)[SPL])
{
	if (SPL == 4) return (a[0] + a[1]) + (a[2] + a[3]);
	if (SPL == 8) return ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]));
	double t = 0.0;
#pragma unroll
	for (int i = 0; i < SPL; ++i) t += a[i];
	return t;
}

// out[i] = D[i] x[i] + pc[i] * sum_{j<i} pm[j] x[j] + sc[i] * sum_{j>i} sm[j] x[j]  (same contract as semisep)
template <int SPL, int G>
__device__ __forceinline__ void semisep2(const double (&x)[SPL], const double (&pm)[SPL], const double (&pc)[SPL],
                                         const double (&sm)[SPL], const double (&sc)[SPL], const double (&D)[SPL],
                                         const DualScan<G> &ds, double (&out)[SPL])
{
	double a[SPL], c[SPL], lp[SPL], ls[SPL], tp, ts, P, S;
#pragma unroll
	for (int i = 0; i < SPL; ++i) {
		a[i] = x[i] * pm[i];
		c[i] = x[i] * sm[i];
	}
	local_prefix<SPL>(a, lp, tp);
	local_suffix<SPL>(c, ls, ts);
	ds.run(tp, ts, P, S);
#pragma unroll
	for (int i = 0; i < SPL; ++i) {
		double base = D[i] * x[i]; // independent of the shuffles (lp[0] and ls[SPL-1] are zero)
		if (i > 0) base = fma(pc[i], lp[i], base);
		if (i < SPL - 1) base = fma(sc[i], ls[i], base);
		out[i] = fma(sc[i], S, fma(pc[i], P, base));
	}
}
// full exclusive prefix P[i] = sum_{j<i} pm[j] x[j] and suffix S[i] = sum_{j>i} sm[j] x[j] (same contract as prefsuf)
template <int SPL, int G>
__device__ __forceinline__ void prefsuf2(const double (&x)[SPL], const double (&pm)[SPL], const double (&sm)[SPL],
                                         const DualScan<G> &ds, double (&P)[SPL], double (&S)[SPL])
{
	double a[SPL], c[SPL], lp[SPL], ls[SPL], tp, ts, p, s;
#pragma unroll
	for (int i = 0; i < SPL; ++i) {
		a[i] = x[i] * pm[i];
		c[i] = x[i] * sm[i];
	}
	local_prefix<SPL>(a, lp, tp);
	local_suffix<SPL>(c, ls, ts);
	ds.run(tp, ts, p, s);
#pragma unroll
	for (int i = 0; i < SPL; ++i) {
		P[i] = p + lp[i];
		S[i] = s + ls[i];
	}
}

// ------------------------------------------------------------------------------------------------
// K1: transfer operators.  grid = (n_k1_chunks, NP / COLS), block = COLS * G threads; the lane group
// of column j pushes the unit vector e_j through every bin of the chunk.
// T[c] is stored column-major (column j = 64 consecutive doubles), mantissas only; Tex[c][j] holds
// the power-of-two exponent of column j.
// ------------------------------------------------------------------------------------------------
// Is the operator of sub-chunk s used by the boundary chains (k_chain_subs) under the flag picture `flags`?  Its parent
// chunk must be flagged; and the sub-chunk at the far end of a chunk (the last one going forward, dir 0; the first one
// going backward, dir 1) only carries the vector into the NEXT chunk of the run, so it is needed only if that chunk is
// flagged as well (one operator in eight otherwise computed for nothing).  flags has guard entries at -1 and n.
__device__ __forceinline__ bool op_needed(const int32_t *__restrict__ flags, const int32_t *__restrict__ parent,
                                          const int32_t *__restrict__ chunk_sub0, int s, int dir)
{
	const int p = parent[s];
	if (!flags[p]) return false;
	if (dir == 0) return !(chunk_sub0[p + 1] - 1 == s) || flags[p + 1] != 0;
	return !(chunk_sub0[p] == s) || flags[p - 1] != 0;
}

template <int SPL, int G, int COLS>
__global__ void __launch_bounds__(COLS *G) k_transfer(const Chunk *__restrict__ chunks, const int32_t *__restrict__ k1_list,
                                                      const uint32_t *__restrict__ obs, const double *__restrict__ model,
                                                      double *__restrict__ T, int32_t *__restrict__ Tex, int N,
                                                      const int32_t *__restrict__ flag, int sel, const int32_t *__restrict__ skip,
                                                      int n_items, const int32_t *__restrict__ chunk_sub0, int dir)
{
	constexpr int NP = SPL * G;
	// sel 0: chunk list (transfer mode), one block row per listed chunk.  sel 3: repair rounds of the warm-up mode: `chunks`
	// is the SUB-chunk table; the block rows stride over it and work on the sub-chunks whose parent chunk is flagged (flag has
	// guard entries at -1 and n) -- a few hundred of ~19 000, so a block row per sub-chunk would spend 0.1 ms per launch on
	// scheduling empty blocks (measured).  k1_list = parent chunk of every sub-chunk in this mode; `skip` marks parents whose
	// operators are already there (computed ahead of time from the previous E-step's failures, see launch_warm).
	for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
	const int c = sel == 0 ? k1_list[item] : item;
	if (sel == 3 && (!op_needed(flag, k1_list, chunk_sub0, c, dir) || (skip && op_needed(skip, k1_list, chunk_sub0, c, dir)))) continue;
	const Chunk ch = chunks[c];
	const int gl = threadIdx.x % G;
	const int col = blockIdx.y * COLS + threadIdx.x / G;
	double cU[SPL], cV[SPL], cW[SPL], cZ[SPL], cD[SPL], e0[SPL], f[SPL];
	const int s0 = gl * SPL;
#pragma unroll
	for (int i = 0; i < SPL; ++i) {
		cU[i] = model[M_U * NP + s0 + i];
		cV[i] = model[M_V * NP + s0 + i];
		cW[i] = model[M_W * NP + s0 + i];
		cZ[i] = model[M_Z * NP + s0 + i];
		cD[i] = model[M_D * NP + s0 + i];
		e0[i] = model[M_E0 * NP + s0 + i];
		f[i] = (s0 + i == col && col < N) ? 1.0 : 0.0;
	}
	int ex = 0;
	const int uend = ch.u0 + ch.len;
	uint32_t word = 0;
	DualScan<G> ds;
	ds.init(gl);
	for (int u = ch.u0; u < uend; ++u) {
		if (u == ch.u0 || (u & 15) == 0) word = __ldg(obs + ch.ow0 + (u >> 4));
		const int x = (word >> ((u & 15) * 2)) & 3;
		double out[SPL];
		if (u == 0) { // first bin of the sequence: emission only (khmm.c:171-174 has no transition there)
#pragma unroll
			for (int i = 0; i < SPL; ++i) out[i] = f[i];
		} else {
			semisep2<SPL, G>(f, cW, cZ, cU, cV, cD, ds, out);
		}
		if (x == 0) {
#pragma unroll
			for (int i = 0; i < SPL; ++i) f[i] = out[i] * e0[i];
		} else if (x == 1) {
#pragma unroll
			for (int i = 0; i < SPL; ++i) f[i] = out[i] * (1.0 - e0[i]);
		} else {
#pragma unroll
			for (int i = 0; i < SPL; ++i) f[i] = out[i];
		}
		if (((u - ch.u0) & 15) == 15 || u == uend - 1) { // exact power-of-two rescale of the column
			double t = 0.0;
#pragma unroll
			for (int i = 0; i < SPL; ++i) t += f[i];
			t = gsum<G>(t);
			if (t > 1e-290 && t < 1e290) {
				const int k = exponent_of(t);
				const double sc = pow2i(-k);
#pragma unroll
				for (int i = 0; i < SPL; ++i) f[i] *= sc;
				ex += k;
			}
		}
	}
	double *Tc = T + (size_t)c * NP * NP + (size_t)col * NP + s0;
	store_vec<SPL>(Tc, f);
	if (gl == 0) Tex[(size_t)c * NP + col] = ex;
	}
}

// ------------------------------------------------------------------------------------------------
// K2: boundary chains.  One block per (sequence, direction); blockDim = NP.
//   dir 0: vstart[c+1] = normalise( T_c * vstart[c] ), starting from a0 at the first chunk.
//   dir 1: bend[c] = T_{c+1}^T * bend[c+1], starting from ones at the last chunk (direction only).
// ------------------------------------------------------------------------------------------------
template <int NP>
__device__ __forceinline__ double block_sum(double v, double *red)
{
	// NP threads, NP in {32,64,128}
	v = gsum<32>(v);
	if (NP > 32) {
		__syncthreads();
		if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
		__syncthreads();
		double t = 0.0;
#pragma unroll
		for (int w = 0; w < NP / 32; ++w) t += red[w];
		v = t;
	}
	return v;
}
template <int NP>
__device__ __forceinline__ int block_max_i(int v, int *red)
{
	v = gmax_i<32>(v);
	if (NP > 32) {
		__syncthreads();
		if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
		__syncthreads();
		int t = INT_MIN;
#pragma unroll
		for (int w = 0; w < NP / 32; ++w) t = max(t, red[w]);
		v = t;
	}
	return v;
}

// v <- normalise(T_c v): thread i holds v[i]; returns the new v[i]  (column-major mantissas + per-column exponents)
template <int NP>
__device__ __forceinline__ double chain_fwd_step(const double *__restrict__ Tc, const int32_t *__restrict__ Texc, double v,
                                                 double *vec, double *red, int *redi)
{
	const int i = threadIdx.x;
	const int ex = Texc[i];
	int e = (v > 0.0) ? ex + ilogb(v) : INT_MIN;
	const int emax = block_max_i<NP>(e, redi);
	__syncthreads();
	vec[i] = (v > 0.0) ? scalbn(v, ex - emax) : 0.0;
	__syncthreads();
	double acc = 0.0;
#pragma unroll 8
	for (int j = 0; j < NP; ++j) acc = fma(__ldg(Tc + (size_t)j * NP + i), vec[j], acc);
	const double tot = block_sum<NP>(acc, red);
	return acc / tot;
}
// b <- T_c^T b (direction only, rescaled so that the largest entry is ~1)
template <int NP>
__device__ __forceinline__ double chain_bwd_step(const double *__restrict__ Tc, const int32_t *__restrict__ Texc, double b,
                                                 double *vec, int *redi)
{
	const int i = threadIdx.x;
	const double *col = Tc + (size_t)i * NP; // column i of T_c
	__syncthreads();
	vec[i] = b;
	__syncthreads();
	double d = 0.0;
#pragma unroll 8
	for (int j = 0; j < NP; ++j) d = fma(__ldg(col + j), vec[j], d);
	const int ex = Texc[i];
	int e = (d > 0.0) ? ex + ilogb(d) : INT_MIN;
	const int emax = block_max_i<NP>(e, redi);
	return (d > 0.0) ? scalbn(d, ex - emax) : 0.0;
}

template <int NP>
__global__ void __launch_bounds__(NP) k_chain(const int32_t *__restrict__ seq_c0, const int32_t *__restrict__ seq_nc,
                                              const double *__restrict__ T, const int32_t *__restrict__ Tex,
                                              const double *__restrict__ model, double *__restrict__ vstart,
                                              double *__restrict__ bend, int n_seqs)
{
	__shared__ double vec[NP];
	__shared__ double red[4];
	__shared__ int redi[4];
	const int seq = blockIdx.x % n_seqs, dir = blockIdx.x / n_seqs;
	const int c0 = seq_c0[seq], nc = seq_nc[seq];
	const int i = threadIdx.x;
	if (nc <= 1) return;
	if (dir == 0) {
		double v = model[M_A0 * NP + i];
		for (int c = c0; c < c0 + nc - 1; ++c) {
			v = chain_fwd_step<NP>(T + (size_t)c * NP * NP, Tex + (size_t)c * NP, v, vec, red, redi);
			vstart[(size_t)(c + 1) * NP + i] = v;
		}
	} else {
		double b = 1.0;
		for (int c = c0 + nc - 2; c >= c0; --c) {
			b = chain_bwd_step<NP>(T + (size_t)(c + 1) * NP * NP, Tex + (size_t)(c + 1) * NP, b, vec, redi);
			bend[(size_t)c * NP + i] = b;
		}
	}
}

// Repair rounds of the warm-up mode.  Flagged chunks are repaired at SUB-chunk granularity (every chunk is pre-split
// into pieces of ~1.5k bins) so that the latency-bound pieces of a repair (transfer operators, recompute) are short.
// One block per chunk; only the HEAD of a run of failed boundaries works and walks the sub-chunks of its run.
//   dir 0 (forward): head = flag[c] && !flag[c-1]; the exact vector in front of the run is the last stored vector of
//          chunk c-1; vsub[s] (start vector of every sub-chunk of the run) follows through the sub-chunk operators.
//   dir 1 (backward): head = flag[c] && !flag[c+1]; the exact direction at the end of chunk c is bexact[c];
//          bsub[s] = direction of b at the last bin of sub-chunk s.
template <int NP>
__global__ void __launch_bounds__(NP) k_chain_subs(const Chunk *__restrict__ subs, int n_sub, const int32_t *__restrict__ parent,
                                                   const int32_t *__restrict__ chunk_sub0, const int32_t *__restrict__ flag, int dir,
                                                   const double *__restrict__ T, const int32_t *__restrict__ Tex,
                                                   const double *__restrict__ fhat, const double *__restrict__ bexact,
                                                   double *__restrict__ vsub, double *__restrict__ bsub)
{
	__shared__ double vec[NP];
	__shared__ double red[4];
	__shared__ int redi[4];
	const int c = blockIdx.x, i = threadIdx.x;
	if (dir == 0) {
		if (!flag[c] || flag[c - 1]) return;
		int s = chunk_sub0[c];
		double v = fhat[((size_t)subs[s].gb0 - 1) * NP + i];
		vsub[(size_t)s * NP + i] = v;
		while (s + 1 < n_sub && flag[parent[s + 1]]) {
			v = chain_fwd_step<NP>(T + (size_t)s * NP * NP, Tex + (size_t)s * NP, v, vec, red, redi);
			++s;
			vsub[(size_t)s * NP + i] = v;
		}
	} else {
		if (!flag[c] || flag[c + 1]) return;
		int s = chunk_sub0[c + 1] - 1;
		double b = bexact[(size_t)c * NP + i];
		bsub[(size_t)s * NP + i] = b;
		while (s - 1 >= 0 && flag[parent[s - 1]]) {
			b = chain_bwd_step<NP>(T + (size_t)s * NP * NP, Tex + (size_t)s * NP, b, vec, redi);
			--s;
			bsub[(size_t)s * NP + i] = b;
		}
	}
}

// after a repair round: fold the per-sub-chunk results of every flagged chunk back into the per-chunk arrays
// (dir 0: log-likelihood partials; dir 1: expected-count partials).  One block per chunk.
__global__ void __launch_bounds__(128) k_fold(const int32_t *__restrict__ chunk_sub0, const int32_t *__restrict__ flag, int dir, int NP,
                                              const double *__restrict__ llsub, double *__restrict__ llpart,
                                              const double *__restrict__ partsub, double *__restrict__ part)
{
	const int c = blockIdx.x;
	if (!flag[c]) return;
	const int s0 = chunk_sub0[c], s1 = chunk_sub0[c + 1];
	if (dir == 0) {
		if (threadIdx.x == 0) {
			double t = 0.0;
			for (int s = s0; s < s1; ++s) t += llsub[s];
			llpart[c] = t;
		}
	} else {
		const int n = S_COUNT * NP;
		for (int j = threadIdx.x; j < n; j += blockDim.x) {
			double t = 0.0;
			for (int s = s0; s < s1; ++s) t += partsub[(size_t)s * n + j];
			part[(size_t)c * n + j] = t;
		}
	}
}

// ------------------------------------------------------------------------------------------------
// Per-lane model constants of a warp that holds one state vector (G = 32 lanes x SPL states).
// ------------------------------------------------------------------------------------------------
// emission of symbol x in a state with hom-emission e0: x = 0 -> e0, x = 1 -> 1 - e0 (bit-identical to the host's
// e[1][k] = 1.0 - e[0][k], core.c:125), x = 2 (missing) -> 1 (khmm.c:21); branch-free: em = c1 * e0 + c0
__device__ __forceinline__ void emis_coef(int x, double &c0, double &c1)
{
	c0 = (x == 0) ? 0.0 : 1.0;
	c1 = (x == 0) ? 1.0 : ((x == 1) ? -1.0 : 0.0);
}

template <int SPL>
struct LaneModel {
	double U[SPL], V[SPL], W[SPL], Z[SPL], D[SPL], e0[SPL];
	__device__ __forceinline__ void load(const double *__restrict__ model, int s0, int NP)
	{
#pragma unroll
		for (int i = 0; i < SPL; ++i) {
			U[i] = model[M_U * NP + s0 + i];
			V[i] = model[M_V * NP + s0 + i];
			W[i] = model[M_W * NP + s0 + i];
			Z[i] = model[M_Z * NP + s0 + i];
			D[i] = model[M_D * NP + s0 + i];
			e0[i] = model[M_E0 * NP + s0 + i];
		}
	}
};

// Hilbert projective mismatch max_i(x_i/y_i) / min_i(x_i/y_i) - 1 of two vectors held like state vectors
// (states >= N ignored); 1e300 if their supports differ.  Same value in every lane.
template <int SPL>
__device__ __forceinline__ double warp_mismatch(const double (&x)[SPL], const double (&y)[SPL], int s0, int N)
{
	double mx = 0.0, mn = 1e300;
	bool bad = false;
#pragma unroll
	for (int i = 0; i < SPL; ++i) {
		if (s0 + i < N) {
			if ((x[i] > 0.0) != (y[i] > 0.0)) bad = true;
			else if (x[i] > 0.0) {
				const double r = x[i] / y[i];
				mx = fmax(mx, r);
				mn = fmin(mn, r);
			}
		}
	}
#pragma unroll
	for (int d = 16; d > 0; d >>= 1) {
		mx = fmax(mx, __shfl_xor_sync(FULLMASK, mx, d));
		mn = fmin(mn, __shfl_xor_sync(FULLMASK, mn, d));
	}
	bad = __any_sync(FULLMASK, bad);
	return (!bad && mn < 1e300 && mn > 0.0) ? mx / mn - 1.0 : 1e300;
}

// ------------------------------------------------------------------------------------------------
// Work assignment of the chunk kernels: a state vector is held by a GROUP of G consecutive lanes (SPL = NP/G
// states per lane), so a warp runs 32/G chunks side by side in lock step.  Narrow groups make the scans
// work-efficient (log2 G shuffle rounds shared by 32/G chunks) and give the FP64 pipe independent work.
// ------------------------------------------------------------------------------------------------
template <int G>
struct GroupId {
	int c, gl;
	bool valid;
	__device__ __forceinline__ GroupId(int n_chunks)
	{
		const int lane = threadIdx.x & 31, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
		gl = lane % G;
		c = warp * (32 / G) + lane / G;
		valid = c < n_chunks;
		if (!valid) c = n_chunks - 1; // idle groups shadow a real chunk so that every address they form stays legal
	}
};

// trip count of a lock-step loop: the largest count of the warp's groups, identical in every lane (and provably so
// for ptxas, which otherwise wraps every shuffle of the loop in WARPSYNC/ENDCOLLECTIVE)
__device__ __forceinline__ int warp_trips(int n)
{
#pragma unroll
	for (int d = 16; d > 0; d >>= 1) n = max(n, __shfl_xor_sync(FULLMASK, n, d));
	return n;
}

// ------------------------------------------------------------------------------------------------
// Forward over bins [ubeg, u0+len) of chunk ch by one lane group, starting from f (the normalised forward
// vector of bin ubeg-1, or a0 when ubeg == 0).  Bins >= u0 are stored (f_u, s_u) and enter the
// log-likelihood; bins < u0 are warm-up.  If fwarm_c != nullptr the vector reached at bin u0-1 is saved
// there.  On return f is the vector of the chunk's last bin.  (khmm.c:171-185 with O(N) transitions.)
// ------------------------------------------------------------------------------------------------
template <int SPL, int G>
__device__ __forceinline__ double forward_chunk(const Chunk &ch, bool valid, int ubeg, const LaneModel<SPL> &M, double (&f)[SPL],
                                                int gl, const uint32_t *__restrict__ obs, double *__restrict__ fhat,
                                                double *__restrict__ sc, double *__restrict__ fwarm_c)
{
	// The recursion carries g_u = (M_u g_{u-1}) * rho_u with the ONE-STEP-DELAYED scale rho_u = 1 / sum(g_{u-1}).
	// Then sum(g_u) = s_u exactly (the reference's scale factor, khmm.c:183-184) and f_u = g_u / s_u, while the
	// reduction and the division that produce rho_{u+1} overlap the next transition instead of sitting on the
	// dependency chain (the chain per bin is scan + combine only).  On entry f is normalised (rho = 1).
	constexpr int NP = SPL * G;
	const int s0 = gl * SPL;
	double ll = 0.0, prod = 1.0;
	const int uend = ch.u0 + ch.len;
	const int trips = warp_trips(valid ? uend - ubeg : 0);
	uint32_t word = 0;
	double g[SPL], rho = 1.0;
#pragma unroll
	for (int i = 0; i < SPL; ++i) g[i] = f[i];
	ScanMasks<G> mk;
	mk.init(gl);
	mk.pin();
	for (int t = 0; t < trips; ++t) {
		const int u = ubeg + t;
		const bool act = valid && u < uend;
		const int uo = act ? u : uend - 1; // idle groups keep reading a legal word
		if (t == 0 || (uo & 15) == 0) word = __ldg(obs + ch.ow0 + (uo >> 4));
		const int x = (word >> ((uo & 15) * 2)) & 3;
		double out[SPL];
		semisep<SPL, G>(g, M.W, M.Z, M.U, M.V, M.D, mk, out);
		double tsum = 0.0, c0, c1;
		emis_coef(x, c0, c1);
		c0 *= rho;
		c1 *= rho;
#pragma unroll
		for (int i = 0; i < SPL; ++i) {
			out[i] = ((u == 0) ? g[i] : out[i]) * fma(c1, M.e0[i], c0); // first bin of a sequence: no transition (khmm.c:171-174)
			tsum += out[i];
		}
		const double s = gsum<G>(tsum); // = s_u
		const double inv = fast_rcp(s);
		if (act) {
#pragma unroll
			for (int i = 0; i < SPL; ++i) g[i] = out[i];
			rho = inv;
			if (u >= ch.u0) {
				double fn[SPL];
#pragma unroll
				for (int i = 0; i < SPL; ++i) fn[i] = out[i] * inv;
				store_vec<SPL>(fhat + ((size_t)ch.gb0 + (u - ch.u0)) * NP + s0, fn);
				if (gl == 0) sc[ch.gb0 + (u - ch.u0)] = s;
				prod *= s; // running product with reset, as hmm_lk (khmm.c:251-258)
				if (prod < 1e-100 || prod > 1e100) {
					ll += log(prod);
					prod = 1.0;
				}
			} else if (u == ch.u0 - 1 && fwarm_c) {
				double fn[SPL];
#pragma unroll
				for (int i = 0; i < SPL; ++i) fn[i] = out[i] * inv;
				store_vec<SPL>(fwarm_c + s0, fn);
			}
		}
	}
#pragma unroll
	for (int i = 0; i < SPL; ++i) f[i] = g[i] * rho; // normalised vector of the last bin
	return ll + log(prod);
}

// ------------------------------------------------------------------------------------------------
// Second-generation forward over a chunk (same contract as forward_chunk).  Differences, all about latency:
//   * the recursion carries an UNNORMALISED vector g_u = q_u * diag(e_{x_u}) A^T g_{u-1}; q_u is 1 or an exact power
//     of two that is switched on when the sum has dropped below 2^-200 -- nothing on the per-bin dependency chain
//     depends on a reduction any more.  With S_u = sum(g_u): s_u = S_u / (q_u S_{u-1}) (the reference's scale factor,
//     khmm.c:183-184), f_u = g_u / S_u, and sum_u log s_u telescopes to log S_last - log S_before - log2(prod q) ln 2.
//   * during the warm-up overlap nothing else is computed; in the store phase the reduction, the reciprocal and the
//     stores of bin u-1 are issued together with the scan of bin u (software pipelining by one bin), so that their
//     shuffle latency hides behind the scan's.
//   * the lane groups of a warp are aligned at the END of their chunks, so that all of them are in the same phase.
// ------------------------------------------------------------------------------------------------
#ifndef PSMC_BOOST_BITS
#define PSMC_BOOST_BITS 200 /* the emulation tests also build with a small value so that boosts happen every few bins */
#endif
#define PSMC_BOOST_LOW pow2i(-PSMC_BOOST_BITS)
#define PSMC_BOOST_UP pow2i(PSMC_BOOST_BITS)
__device__ __forceinline__ int warp_min_i(int n)
{
#pragma unroll
	for (int d = 16; d > 0; d >>= 1) n = min(n, __shfl_xor_sync(FULLMASK, n, d));
	return n;
}
template <int SPL, int G>
struct ForwardRun {
	static constexpr int NP = SPL * G;
	const Chunk &ch;
	const LaneModel<SPL> &M;
	const bool valid;
	const int gl, s0, u0, uend;
	const uint32_t *__restrict__ obs;
	double *__restrict__ fhat, *__restrict__ sc, *__restrict__ fwarm_c;
	DualScan<G> ds;
	double g[SPL], ps, inv_prev, qc, rq_cur, Sstart;
	// `valid`, "has this group started" and "is there a warm-up vector to publish" are folded into bin indices the
	// loop compares u with (one ISETP each; a bool that lives across the loop gets re-derived from threadIdx in every
	// iteration once registers are tight -- measured):
	int u_store; // bins u-1 >= u_store are stored                (u0, or never for an idle shadow group)
	int u_first; // Sstart = S of the bin before bin u_first      (u0, or never)
	int u_warm;  // the vector of bin u_warm - 1 goes to fwarm_c  (u0 after a warm-up, or never)
	int u_boost; // boosts of bins >= u_boost enter the log-likelihood
	int kq, ubase, wlast, tpend, mystart;
	uint32_t wa, wb, wna, wnb; // packed observation words of the current block (word ia and the one after it) / of the next block (ina)
	int ia, ina;
	double *prow, *psc; // where the row / scale factor of the bin finished by the next store-phase step go (advance one bin per step)

	__device__ __forceinline__ ForwardRun(const Chunk &ch_, bool valid_, const LaneModel<SPL> &M_, int gl_,
	                                      const uint32_t *__restrict__ obs_, double *__restrict__ fhat_, double *__restrict__ sc_,
	                                      double *__restrict__ fwarm_c_)
	    : ch(ch_), M(M_), valid(valid_), gl(gl_), s0(gl_ * SPL), u0(ch_.u0), uend(ch_.u0 + ch_.len), obs(obs_),
	      fhat(fhat_), sc(sc_), fwarm_c(fwarm_c_)
	{
	}

	// A group whose chunk needs fewer steps than the longest one of its warp starts late: until then it computes on
	// whatever its registers hold (no per-step select keeps it idle) and picks up its real start vector here.
	__device__ __forceinline__ void start_due(int t, const double (&f)[SPL], double inv_before)
	{
		if (mystart == t) {
#pragma unroll
			for (int i = 0; i < SPL; ++i) g[i] = f[i];
			ps = local_sum<SPL>(g);
			inv_prev = inv_before;
			rq_cur = 1.0;
			qc = 1.0;
		}
		tpend = warp_min_i(mystart > t ? mystart : INT_MAX);
	}

	// A block = the steps [t, tstop) between two multiples of 16 (or up to a late start / the end of the phase).  Makes the
	// words prefetched for a block starting at t current and fetches those of the block starting at tstop (indices clamped
	// to the sequence, so idle groups read legal words too; the loads have a whole block to complete).
	__device__ __forceinline__ int begin_block(int t, int tend)
	{
		const int tstop = min(min(tend, tpend), (t & ~15) + 16);
		wa = wna; wb = wnb; ia = ina;
		ina = (ubase + tstop) >> 4;
		wna = __ldg(obs + ch.ow0 + min(max(ina, 0), wlast));
		wnb = __ldg(obs + ch.ow0 + min(max(ina + 1, 0), wlast));
		return tstop;
	}

	// bin u = ubase + t is formed from the current vector (bin u-1); BOOK: finish bin u-1 (sum, reciprocal, stores) alongside
	template <bool BOOK>
	__device__ __forceinline__ void step(int t)
	{
		const int u = ubase + t;
		// the (at most two) packed words of this block of <= 16 bins were fetched during the previous block (begin_block):
		// no load and no address arithmetic in here
		const int x = ((((u >> 4) == ia) ? wa : wb) >> ((u & 15) * 2)) & 3;
		double c0, c1;
		emis_coef(x, c0, c1);
		c0 *= qc;
		c1 *= qc;
		double out[SPL];
		semisep2<SPL, G>(g, M.W, M.Z, M.U, M.V, M.D, ds, out);
		if (BOOK) { // (written after the scan so that the scheduler issues the scan's shuffles first and this reduction in their shadow)
			const double S1 = gsum<G>(ps); // = S_{u-1}
			const double inv1 = fast_rcp(S1);
			const bool st_f = u - 1 >= u_store, st_w = u == u_warm;
			if (st_f || st_w) {
				double fn[SPL];
#pragma unroll
				for (int i = 0; i < SPL; ++i) fn[i] = g[i] * inv1;
				store_vec<SPL>(st_f ? prow : fwarm_c + s0, fn);
			}
			if (st_f && gl == 0) *psc = S1 * inv_prev * rq_cur;
			prow += NP;
			psc += 1;
			if (u == u_first) Sstart = S1;
			const bool boosted = qc != 1.0;
			rq_cur = boosted ? PSMC_BOOST_LOW : 1.0;
			if (boosted && u >= u_boost) kq += PSMC_BOOST_BITS;
			inv_prev = inv1;
			qc = (S1 < PSMC_BOOST_LOW) ? PSMC_BOOST_UP : 1.0; // acts on the NEXT bin: off the dependency chain
		} else {
			qc = 1.0; // (the warm-up phase decides once per block of 16 bins, see run)
		}
#pragma unroll
		for (int i = 0; i < SPL; ++i) g[i] = out[i] * fma(c1, M.e0[i], c0);
		ps = local_sum<SPL>(g);
	}

	__device__ __forceinline__ double run(int ubeg_in, double (&f)[SPL])
	{
		int ubeg = ubeg_in;
		const bool warmed = ubeg_in < u0;
		wlast = (ch.Lseq - 1) >> 4;
		ds.init(gl);
		Sstart = 1.0;
		kq = 0;
		double inv_before = 1.0; // 1 / S of the bin before the start vector's bin (only the first bin of a sequence needs it)
		const double S_init = gsum<G>(local_sum<SPL>(f)); // (every lane of the warp takes part)
		bool have_start = false;
		if (ubeg_in == 0) {
			// first bin of a sequence: emission only, no transition (khmm.c:171-174); f is a0 here.  Done in front of the loop
			// so that the loop body has no special case: the start vector becomes bin 0 (unnormalised) and bin 1 is formed first.
			const int x = __ldg(obs + ch.ow0) & 3;
			double c0, c1;
			emis_coef(x, c0, c1);
#pragma unroll
			for (int i = 0; i < SPL; ++i) f[i] *= fma(c1, M.e0[i], c0);
			inv_before = fast_rcp(S_init);
			if (u0 == 0) {
				Sstart = S_init;
				have_start = true;
			}
			ubeg = 1;
		}
		// from here on: f = vector of bin ubeg-1, the group forms bins ubeg .. uend-1 and is aligned with the other groups at the END
		const int never = INT_MAX / 2;
		u_store = valid ? u0 : never;
		u_first = (valid && !have_start) ? u0 : never;
		u_warm = (valid && warmed && fwarm_c != nullptr) ? u0 : never;
		u_boost = valid ? max(u0, ubeg) : never;
		const int mytrips = valid ? uend - ubeg : 0;
		const int trips = warp_trips(mytrips);
		const int tB = warp_min_i(valid ? trips - (uend - max(u0, ubeg)) : trips); // first step in which some group forms a bin it stores
		ubase = uend - trips;
		mystart = valid ? trips - mytrips : INT_MAX;
		ina = ubase >> 4; // words of the first block (begin_block makes them current)
		wna = __ldg(obs + ch.ow0 + min(max(ina, 0), wlast));
		wnb = __ldg(obs + ch.ow0 + min(max(ina + 1, 0), wlast));
#pragma unroll
		for (int i = 0; i < SPL; ++i) g[i] = f[i];
		ps = local_sum<SPL>(g);
		inv_prev = inv_before;
		rq_cur = 1.0;
		qc = 1.0;
		tpend = warp_min_i(mystart);
		int t = 0;
		while (t < tB) { // warm-up phase, in blocks of at most 16 bins: late starts and the boost decision sit between blocks
			if (t == tpend) start_due(t, f, inv_before);
			const int tstop = begin_block(t, tB);
			qc = (gsum<G>(ps) < PSMC_BOOST_LOW) ? PSMC_BOOST_UP : 1.0;
			for (; t < tstop; ++t) step<false>(t);
		}
		qc = 1.0;
		{ // integer arithmetic: for idle groups the address lies outside the buffers until their first stored bin (never dereferenced)
			const long long row0 = (long long)ch.gb0 + ((long long)ubase + t - 1 - u0);
			prow = (double *)((char *)fhat + (row0 * NP + s0) * (long long)sizeof(double));
			psc = (double *)((char *)sc + row0 * (long long)sizeof(double));
		}
		while (t < trips) { // store phase
			if (t == tpend) start_due(t, f, inv_before);
			const int tstop = begin_block(t, trips);
			for (; t < tstop; ++t) step<true>(t);
		}
		if (mytrips == 0) { // nothing formed in the loop (a one-bin chunk at the start of a sequence): the start vector is the last bin
#pragma unroll
			for (int i = 0; i < SPL; ++i) g[i] = f[i];
			ps = local_sum<SPL>(g);
			inv_prev = inv_before;
			rq_cur = 1.0;
		}
		// finish the last bin
		const double S1 = gsum<G>(ps), inv1 = fast_rcp(S1);
#pragma unroll
		for (int i = 0; i < SPL; ++i) f[i] = g[i] * inv1;
		if (valid) {
			store_vec<SPL>(fhat + ((size_t)ch.gb0 + (ch.len - 1)) * NP + s0, f);
			if (gl == 0) sc[ch.gb0 + (ch.len - 1)] = S1 * inv_prev * rq_cur;
		}
		return (log(S1) - log(Sstart)) - (double)kq * 0.69314718055994530942;
	}
};

template <int SPL, int G>
__device__ __forceinline__ double forward_chunk2(const Chunk &ch, bool valid, int ubeg, const LaneModel<SPL> &M, double (&f)[SPL],
                                                 int gl, const uint32_t *__restrict__ obs, double *__restrict__ fhat,
                                                 double *__restrict__ sc, double *__restrict__ fwarm_c)
{
	ForwardRun<SPL, G> r(ch, valid, M, gl, obs, fhat, sc, fwarm_c);
	return r.run(ubeg, f);
}

// ------------------------------------------------------------------------------------------------
// FP32 pre-warm-up.  An overlap only has to deliver a start vector that is exact to 1e-12 AFTER its last few thousand
// bins; what happens in its early part is forgotten geometrically.  So the early part runs in FP32 (one SHFL per
// value instead of two, FFMA at full rate and 4 cycles of latency instead of DFMA at half rate and 9) in its own short
// kernel, and the FP64 overlap of the forward / backward warm-up kernel starts from its result instead of from the
// stationary vector.  FP32 leaves a relative error of ~1e-6 in the vector; the FP64 part contracts it like any other
// start error (the certificate decides, as always).  8-lane groups only (NP <= 64).
// ------------------------------------------------------------------------------------------------
template <typename T>
struct DualScan8T { // DualScan<8> in the scalar type T
	bool h0, h1, h2;
	T u0, u1, u2, d0, d1, d2;
	__device__ __forceinline__ void init(int gl)
	{
		h0 = (gl & 1) != 0; h1 = (gl & 2) != 0; h2 = (gl & 4) != 0;
		u0 = h0 ? T(1) : T(0); u1 = h1 ? T(1) : T(0); u2 = h2 ? T(1) : T(0);
		d0 = T(1) - u0; d1 = T(1) - u1; d2 = T(1) - u2;
	}
	__device__ __forceinline__ void run(T tp, T ts, T &P, T &S) const
	{
		const T s0 = h0 ? ts : tp, s1 = h1 ? ts : tp, s2 = h2 ? ts : tp;
		const T r1 = __shfl_xor_sync(FULLMASK, s0, 1, 8);
		const T r2 = __shfl_xor_sync(FULLMASK, s1, 2, 8), r3 = __shfl_xor_sync(FULLMASK, s1, 3, 8);
		const T r4 = __shfl_xor_sync(FULLMASK, s2, 4, 8), r5 = __shfl_xor_sync(FULLMASK, s2, 5, 8);
		const T r6 = __shfl_xor_sync(FULLMASK, s2, 6, 8), r7 = __shfl_xor_sync(FULLMASK, s2, 7, 8);
		const T q1 = r2 + r3, q2 = (r4 + r5) + (r6 + r7);
		P = fma(u2, q2, fma(u1, q1, u0 * r1));
		S = fma(d2, q2, fma(d1, q1, d0 * r1));
	}
};
template <typename T, int SPL>
__device__ __forceinline__ void local_prefix_t(const T (&a)[SPL], T (&lp)[SPL], T &tot)
{
	T t = T(0);
	if (SPL == 8) {
		const T p01 = a[0] + a[1], p23 = a[2] + a[3], p45 = a[4] + a[5], p67 = a[6] + a[7];
		const T q03 = p01 + p23, q47 = p45 + p67, q05 = q03 + p45;
		lp[0] = T(0); lp[1] = a[0]; lp[2] = p01; lp[3] = p01 + a[2];
		lp[4] = q03; lp[5] = q03 + a[4]; lp[6] = q05; lp[7] = q05 + a[6];
		tot = q03 + q47;
		return;
	}
#pragma unroll
	for (int i = 0; i < SPL; ++i) {
		lp[i] = t;
		t += a[i];
	}
	tot = t;
}
template <typename T, int SPL>
__device__ __forceinline__ T local_sum_t(const T (&a)[SPL])
{
	T t = T(0);
#pragma unroll
	for (int i = 0; i < SPL; ++i) t += a[i];
	return t;
}
template <typename T>
__device__ __forceinline__ T gsum8_t(T t)
{
#pragma unroll
	for (int d = 4; d > 0; d >>= 1) t += __shfl_xor_sync(FULLMASK, t, d, 8);
	return t;
}
// out[i] = D[i] x[i] + pc[i] * sum_{j<i} pm[j] x[j] + sc[i] * sum_{j>i} sm[j] x[j]  (semisep2 in the scalar type T, 8 lanes)
template <typename T, int SPL>
__device__ __forceinline__ void semisep2_t(const T (&x)[SPL], const T (&pm)[SPL], const T (&pc)[SPL], const T (&sm)[SPL],
                                           const T (&sc)[SPL], const T (&D)[SPL], const DualScan8T<T> &ds, T (&out)[SPL])
{
	T a[SPL], c[SPL], r[SPL], lp[SPL], lr[SPL], tp, ts, P, S;
#pragma unroll
	for (int i = 0; i < SPL; ++i) {
		a[i] = x[i] * pm[i];
		c[i] = x[i] * sm[i];
		r[SPL - 1 - i] = c[i];
	}
	local_prefix_t<T, SPL>(a, lp, tp);
	local_prefix_t<T, SPL>(r, lr, ts); // suffix sums = prefix sums of the reversed array
	ds.run(tp, ts, P, S);
#pragma unroll
	for (int i = 0; i < SPL; ++i) {
		T base = D[i] * x[i];
		if (i > 0) base = fma(pc[i], lp[i], base);
		if (i < SPL - 1) base = fma(sc[i], lr[SPL - 1 - i], base);
		out[i] = fma(sc[i], S, fma(pc[i], P, base));
	}
}

// DIR 0: for every forward chunk whose FP64 overlap [u0 - warm64, u0) does not reach the start of its sequence, the forward
//        vector of bin u0 - warm64 - 1 from warm32 bins further left (sum-normalised, as doubles) -> start[c].
// DIR 1: for every backward chunk whose FP64 overlap (ulast, ulast + warm64] does not reach the end of its sequence, the
//        direction of b at bin ulast + warm64 from warm32 bins further right -> start[c].
// Same loop structure as the FP64 warm-up (groups aligned at the end, late starts between 16-bin blocks, boosts).
template <int SPL, int DIR>
__global__ void __launch_bounds__(128) k_prewarm(const Chunk *__restrict__ chunks, int n_chunks, const uint32_t *__restrict__ obs,
                                                 const double *__restrict__ model, int warm64, int warm32, double *__restrict__ start)
{
	typedef float T;
	constexpr int G = 8, NP = SPL * G;
	const GroupId<G> id(n_chunks);
	if (!__any_sync(FULLMASK, id.valid)) return;
	const int c = id.c, gl = id.gl, s0 = gl * SPL;
	const Chunk ch = chunks[c];
	T cU[SPL], cV[SPL], cW[SPL], cZ[SPL], cD[SPL], e0[SPL], v[SPL];
#pragma unroll
	for (int i = 0; i < SPL; ++i) {
		cU[i] = (T)model[M_U * NP + s0 + i];
		cV[i] = (T)model[M_V * NP + s0 + i];
		cW[i] = (T)model[M_W * NP + s0 + i];
		cZ[i] = (T)model[M_Z * NP + s0 + i];
		cD[i] = (T)model[M_D * NP + s0 + i];
		e0[i] = (T)model[M_E0 * NP + s0 + i];
		v[i] = DIR == 0 ? (T)model[M_A0 * NP + s0 + i] : T(1);
	}
	const int ulast = ch.u0 + ch.len - 1;
	// bins processed: DIR 0: ua .. ub ascending (ub = u0 - warm64 - 1);  DIR 1: ua .. ub descending (ub = ulast + warm64 + 1)
	bool need;
	int ua, ub;
	if (DIR == 0) {
		need = id.valid && !(ch.flags & CH_FIRST) && ch.u0 - warm64 > 0;
		ub = ch.u0 - warm64 - 1;
		ua = max(0, ub - warm32 + 1);
	} else {
		need = id.valid && !(ch.flags & CH_LAST) && ulast + warm64 < ch.Lseq - 1;
		ub = ulast + warm64 + 1;
		ua = min(ch.Lseq - 1, ub + warm32 - 1);
	}
	const int mytrips = need ? (DIR == 0 ? ub - ua + 1 : ua - ub + 1) : 0;
	const int trips = warp_trips(mytrips);
	const int mystart = mytrips > 0 ? trips - mytrips : INT_MAX;
	int tpend = warp_min_i(mystart);
	const int wlast = (ch.Lseq - 1) >> 4;
	const int ufirst = DIR == 0 ? ub - trips + 1 : ub + trips - 1; // bin of step 0 (may lie outside the sequence: clamped words)
	const int uprev = DIR == 0 ? ufirst - 1 : ufirst + 1;          // as if this bin had just been processed
	uint32_t word = __ldg(obs + ch.ow0 + min(max(uprev >> 4, 0), wlast));
	uint32_t wnext = __ldg(obs + ch.ow0 + min(max((uprev >> 4) + (DIR == 0 ? 1 : -1), 0), wlast));
	DualScan8T<T> ds;
	ds.init(gl);
	T g[SPL], q = T(1);
#pragma unroll
	for (int i = 0; i < SPL; ++i) g[i] = v[i];
	int t = 0;
	while (t < trips) {
		if (t == tpend) { // warp-uniform, rare
			if (mystart == t) {
#pragma unroll
				for (int i = 0; i < SPL; ++i) g[i] = v[i];
			}
			tpend = warp_min_i(mystart > t ? mystart : INT_MAX);
		}
		const int tstop = min(min(trips, tpend), (t & ~15) + 16);
		q = (gsum8_t<T>(local_sum_t<T, SPL>(g)) < T(9.094947e-13)) ? T(1.0995116e12) : T(1); // 2^-40 / 2^40
		for (; t < tstop; ++t) {
			const int u = DIR == 0 ? ufirst + t : ufirst - t;
			if ((u & 15) == (DIR == 0 ? 0 : 15)) {
				word = wnext;
				wnext = __ldg(obs + ch.ow0 + min(max((u >> 4) + (DIR == 0 ? 1 : -1), 0), wlast));
			}
			const int x = (word >> ((u & 15) * 2)) & 3;
			const T c0 = (x == 0 ? T(0) : T(1)) * q, c1 = (x == 0 ? T(1) : (x == 1 ? T(-1) : T(0))) * q;
			q = T(1);
			T out[SPL];
			if (DIR == 0) {
				semisep2_t<T, SPL>(g, cW, cZ, cU, cV, cD, ds, out);
#pragma unroll
				for (int i = 0; i < SPL; ++i) g[i] = out[i] * fma(c1, e0[i], c0);
			} else {
				T h[SPL];
#pragma unroll
				for (int i = 0; i < SPL; ++i) h[i] = fma(c1, e0[i], c0) * g[i];
				semisep2_t<T, SPL>(h, cV, cU, cZ, cW, cD, ds, g);
			}
		}
	}
	const T tot = gsum8_t<T>(local_sum_t<T, SPL>(g));
	if (need) {
		const double inv = 1.0 / (double)tot;
		double o[SPL];
#pragma unroll
		for (int i = 0; i < SPL; ++i) o[i] = fmax((double)g[i] * inv, 1e-300); // (a flushed component restarts positive)
		store_vec<SPL>(start + (size_t)c * NP + s0, o);
	}
}

// ------------------------------------------------------------------------------------------------
// K3: forward.  One lane group per chunk.
//   warm == 0 : the exact start vector comes from the boundary chain (vstart, transfer mode).
//   warm  > 0 : the group starts `warm` bins to the LEFT of its chunk from the stationary vector, runs
//               the same recursion without storing (the HMM forgets its start geometrically) and saves
//               the vector it reached at the bin before its chunk (fwarm[c]) for the certificate.
// ------------------------------------------------------------------------------------------------
template <int SPL, int G, int VER>
__global__ void __launch_bounds__(128) k_forward(const Chunk *__restrict__ chunks, int n_chunks,
                                                 const uint32_t *__restrict__ obs, const double *__restrict__ model,
                                                 const double *__restrict__ vstart, int warm, int use_prev, double *__restrict__ fhat,
                                                 double *__restrict__ sc, double *__restrict__ llpart, double *__restrict__ fwarm,
                                                 const int32_t *__restrict__ order, const int32_t *__restrict__ warm_of,
                                                 const double *__restrict__ pre_start)
{
	constexpr int NP = SPL * G;
	const GroupId<G> id(n_chunks);
	if (!__any_sync(FULLMASK, id.valid)) return;
	// pre_start: start vectors of the overlaps from the FP32 pre-warm-up (k_prewarm), nullptr = stationary start.
	// order: chunks sorted by their number of steps, so that the groups of a warp finish together (adaptive overlaps);
	// warm_of: this chunk's own overlap (nullptr: `warm` for every chunk)
	const int c = order ? order[id.c] : id.c, gl = id.gl, s0 = gl * SPL;
	const Chunk ch = chunks[c];
	LaneModel<SPL> M;
	M.load(model, s0, NP);
	double f[SPL];
	int ubeg = ch.u0;
	if ((ch.flags & CH_FIRST) || warm > 0) {
		if (!(ch.flags & CH_FIRST)) ubeg = max(0, ch.u0 - (warm_of ? min(warm_of[c], warm) : warm));
		if (use_prev && ubeg > 0) {
			// warm start: the vector the PREVIOUS E-step stored for bin ubeg-1 (the parameters moved only a little since;
			// any positive vector is a legal start -- the certificate decides -- so a stale or concurrently rewritten row is harmless)
			const double *row = fhat + ((size_t)ch.gb0 - (size_t)(ch.u0 - ubeg) - 1) * NP + s0;
#pragma unroll
			for (int i = 0; i < SPL; ++i) f[i] = fmax(row[i], 1e-300);
		} else if (pre_start && ubeg > 0) {
			load_vec<SPL>(pre_start + (size_t)c * NP + s0, f);
		} else {
#pragma unroll
			for (int i = 0; i < SPL; ++i) f[i] = model[M_A0 * NP + s0 + i];
		}
	} else {
		load_vec<SPL>(vstart + (size_t)c * NP + s0, f);
	}
	double ll;
	if constexpr (VER == 2) ll = forward_chunk2<SPL, G>(ch, id.valid, ubeg, M, f, gl, obs, fhat, sc, fwarm + (size_t)c * NP);
	else ll = forward_chunk<SPL, G>(ch, id.valid, ubeg, M, f, gl, obs, fhat, sc, fwarm + (size_t)c * NP);
	if (gl == 0 && id.valid) llpart[c] = ll;
}

// ------------------------------------------------------------------------------------------------
// Adaptive overlaps.  How many bins of warm-up a boundary needs is a property of the data around it (het-poor,
// low-TMRCA tracts mix slowly) and varies by an order of magnitude between boundaries; the mismatch the certificate
// measures anyway tells, per boundary, whether the overlap of this E-step was ample (mismatch at the rounding floor),
// tight, or too short.  Additive-decrease / multiplicative-increase on that signal: shrink by 1/8 while the mismatch
// stays at the floor, grow by 1/2 as soon as it leaves the safe band (the boundary still passes at 1e-12, or is
// repaired).  All on the device, inside the mark kernels of the first repair round; the next E-step reads the new lengths.
// ------------------------------------------------------------------------------------------------
// w: the overlap used in this E-step; tight: the longest overlap seen so far whose mismatch was NOT at the floor (memory,
// so that a boundary approaches its need from above once instead of probing it again and again).
__device__ __forceinline__ int adapt_overlap(int w, double mismatch, int w_max, int32_t *tight)
{
	const int w_min = 1536;
	int t = *tight;
	if (!(mismatch <= 2e-14)) { // off the floor (or failed): remember, and keep a safe distance from now on
		t = max(t, w);
		*tight = t;
	}
	const int floor_w = max(w_min, (t + (t >> 1)) & ~15); // 1.5 x the tightest length seen
	if (mismatch <= 2e-14) w = max(floor_w, (w - (w >> 4)) & ~15); // 1/16 per E-step: a step multiplies the mismatch by < 10
	else w = max(floor_w, w);
	return max(min(w, w_max), 16);
}

// order[] = chunk indices sorted (stably) by their number of steps, longest first: every thread ranks one chunk
// (O(n^2 / threads); n is a few thousand).  steps = (overlap unless the chunk starts / ends its sequence) + chunk length.
__global__ void __launch_bounds__(256) k_order(const Chunk *__restrict__ chunks, int n_chunks, const int32_t *__restrict__ warm_of,
                                               int warm_cap, int edge_flag, int32_t *__restrict__ order)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_chunks) return;
	const int ki = ((chunks[i].flags & edge_flag) ? 0 : min(warm_of[i], warm_cap)) + chunks[i].len;
	int rank = 0;
	for (int j = 0; j < n_chunks; ++j) {
		const int kj = ((chunks[j].flags & edge_flag) ? 0 : min(warm_of[j], warm_cap)) + chunks[j].len;
		rank += (kj > ki || (kj == ki && j < i)) ? 1 : 0;
	}
	order[rank] = i;
}

// ------------------------------------------------------------------------------------------------
// K3r: forward repair round (warm-up mode).  Boundary c (between chunks c-1 and c) "fails" when
// fwarm[c] differs from the last stored vector of chunk c-1 by more than eps (Hilbert metric).
// A warp acts only as the HEAD of a run of failed boundaries (its own fails, its left neighbour's
// passes, so the left neighbour's stored vectors are final): it recomputes chunk c from the exact
// vector, then keeps going through the following chunks whose boundaries also failed at kernel start
// (nobody else touches those).  The boundary after the run is re-examined by the next round.
// stat[0] += failed boundaries seen at kernel start, stat[1] += chunks recomputed.
// ------------------------------------------------------------------------------------------------
template <int SPL>
__device__ __forceinline__ double fwd_boundary_mismatch(const Chunk &ch, int c, const double *__restrict__ fhat,
                                                        const double *__restrict__ fwarm, int s0, int N)
{
	constexpr int NP = SPL * 32;
	double x[SPL], y[SPL];
	load_vec<SPL>(fwarm + (size_t)c * NP + s0, x);
	load_vec<SPL>(fhat + ((size_t)ch.gb0 - 1) * NP + s0, y);
	return warp_mismatch<SPL>(x, y, s0, N);
}

// flag_f[c] = 1 iff the boundary in front of chunk c currently fails (evaluated once per round, so that the
// repair kernel takes its decisions on a frozen picture); stat[0] += failures
template <int SPL>
__global__ void __launch_bounds__(128) k_mark_fwd(const Chunk *__restrict__ chunks, int n_chunks, int N, double eps,
                                                  const double *__restrict__ fhat, const double *__restrict__ fwarm,
                                                  int32_t *__restrict__ flag_f, unsigned long long *__restrict__ stat,
                                                  int32_t *__restrict__ pred_next, int32_t *__restrict__ warm_of, int warm_max, int32_t *__restrict__ tight)
{
	const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	if (c >= n_chunks) return;
	const int gl = threadIdx.x & 31;
	const Chunk ch = chunks[c];
	int fl = 0;
	double m = 0.0;
	if (!(ch.flags & CH_FIRST)) {
		m = fwd_boundary_mismatch<SPL>(ch, c, fhat, fwarm, gl * SPL, N);
		fl = m > eps ? 1 : 0;
	}
	if (gl == 0) {
		flag_f[c] = fl;
		if (pred_next) pred_next[c] = fl;
		if (fl) atomicAdd(&stat[0], 1ull);
		if (warm_of && pred_next && !(ch.flags & CH_FIRST)) warm_of[c] = adapt_overlap(warm_of[c], m, warm_max, tight + c); // first round only
	}
}

template <int SPL, int G, int VER>
__global__ void __launch_bounds__(128) k_forward_repair(const Chunk *__restrict__ subs, int n_sub, const int32_t *__restrict__ parent,
                                                        const int32_t *__restrict__ chunk_sub0,
                                                        const uint32_t *__restrict__ obs, const double *__restrict__ model,
                                                        const int32_t *__restrict__ flag_f, const double *__restrict__ vsub,
                                                        double *__restrict__ fhat, double *__restrict__ sc,
                                                        double *__restrict__ llsub, double *__restrict__ fwarm,
                                                        unsigned long long *__restrict__ stat)
{
	constexpr int NP = SPL * G;
	const GroupId<G> id(n_sub);
	const int s = id.c, gl = id.gl, s0 = gl * SPL;
	const int pc = parent[s];
	const bool valid = id.valid && flag_f[pc] != 0;
	if (!__any_sync(FULLMASK, valid)) return;
	const Chunk ch = subs[s];
	LaneModel<SPL> M;
	M.load(model, s0, NP);
	double f[SPL];
	load_vec<SPL>(vsub + (size_t)s * NP + s0, f);                                              // exact vector of the bin before the sub-chunk (k_chain_subs)
	if (valid && chunk_sub0[pc] == s) store_vec<SPL>(fwarm + (size_t)pc * NP + s0, f);        // the chunk boundary agrees by construction from now on
	double ll;
	if constexpr (VER == 2) ll = forward_chunk2<SPL, G>(ch, valid, ch.u0, M, f, gl, obs, fhat, sc, nullptr);
	else ll = forward_chunk<SPL, G>(ch, valid, ch.u0, M, f, gl, obs, fhat, sc, nullptr);
	if (gl == 0 && valid) {
		llsub[s] = ll;
		atomicAdd(&stat[1], 1ull);
	}
}

// ------------------------------------------------------------------------------------------------
// Backward over chunk ch by one lane group from b = b_{ulast} (reference scaling), accumulating the expected
// counts into part_c; on return b is the vector of bin u0-1 (if u0 > 0).  (khmm.c:226-235, 310-318.)
// ------------------------------------------------------------------------------------------------
template <int SPL, int G>
__device__ __forceinline__ void backward_chunk(const Chunk &ch, bool valid, const LaneModel<SPL> &M, double (&b)[SPL], int gl,
                                               const uint32_t *__restrict__ obs, const double *__restrict__ fhat,
                                               const double *__restrict__ sc, double *__restrict__ part_c,
                                               double *__restrict__ bsave_c = nullptr, int usave = -1)
{
	constexpr int NP = SPL * G, PF = 4;
	const int s0 = gl * SPL;
	double aE0[SPL], aE1[SPL], aRL[SPL], aCL[SPL], aRU[SPL], aCU[SPL], aAD[SPL];
#pragma unroll
	for (int i = 0; i < SPL; ++i) aE0[i] = aE1[i] = aRL[i] = aCL[i] = aRU[i] = aCU[i] = aAD[i] = 0.0;
	const int ulast = ch.u0 + ch.len - 1;
	const double *frow = fhat + ((size_t)ch.gb0 + (ch.len - 1)) * NP + s0; // row of bin ulast
	const double *srow = sc + ch.gb0 + (ch.len - 1);
	double fu[SPL], su;
	load_vec<SPL>(frow, fu);
	su = __ldg(srow);
	// software prefetch ring: nf[j], ns[j] hold row (u-1-j) while bin u is processed
	double nf[PF][SPL], ns[PF];
#pragma unroll
	for (int j = 0; j < PF; ++j) {
		const int uu = ulast - 1 - j;
		if (uu >= 0) {
			load_vec<SPL>(frow - (size_t)(1 + j) * NP, nf[j]);
			ns[j] = __ldg(srow - (1 + j));
		} else {
#pragma unroll
			for (int i = 0; i < SPL; ++i) nf[j][i] = 0.0;
			ns[j] = 1.0;
		}
	}
	const int trips = warp_trips(valid ? ch.len : 0);
	uint32_t word = 0;
	ScanMasks<G> mk;
	mk.init(gl);
	for (int t = 0; t < trips; ++t) {
		const int u = ulast - t;
		const bool act = valid && u >= ch.u0;
		const int uo = act ? u : ch.u0;
		if (t == 0 || (uo & 15) == 15) word = __ldg(obs + ch.ow0 + (uo >> 4));
		const int x = (word >> ((uo & 15) * 2)) & 3;
		if (act && u == usave && bsave_c) store_vec<SPL>(bsave_c + s0, b); // warm start of the left neighbour's next overlap
		// emission counts: bins 0..L-2 only (khmm.c:310, 317)
		if (act && u != ch.Lseq - 1) {
			const double w0 = (x == 0) ? su : 0.0, w1 = (x == 1) ? su : 0.0;
#pragma unroll
			for (int i = 0; i < SPL; ++i) {
				const double fb = fu[i] * b[i];
				aE0[i] = fma(fb, w0, aE0[i]);
				aE1[i] = fma(fb, w1, aE1[i]);
			}
		}
		const bool trans = act && u > 0; // no transition into the first bin of a sequence
		// rotate the prefetch ring
		double fm[SPL], sm = ns[0];
#pragma unroll
		for (int i = 0; i < SPL; ++i) fm[i] = nf[0][i];
		if (act) {
#pragma unroll
			for (int j = 0; j + 1 < PF; ++j) {
#pragma unroll
				for (int i = 0; i < SPL; ++i) nf[j][i] = nf[j + 1][i];
				ns[j] = ns[j + 1];
			}
			const int uu = u - 1 - PF;
			if (uu >= 0) {
				const size_t back = (size_t)(ulast - uu);
				load_vec<SPL>(frow - back * NP, nf[PF - 1]);
				ns[PF - 1] = __ldg(srow - back);
			}
		}
		// transition u-1 -> u (khmm.c:313-318 for the counts, khmm.c:230-234 for b_{u-1})
		double g[SPL], Pg[SPL], Sg[SPL], Pf[SPL], Sf[SPL], c0, c1;
		emis_coef(x, c0, c1);
#pragma unroll
		for (int i = 0; i < SPL; ++i) g[i] = fma(c1, M.e0[i], c0) * b[i];
		prefsuf<SPL, G>(g, M.V, M.Z, mk, Pg, Sg);  // Pg = sum_{l<k} V_l g_l, Sg = sum_{l>k} Z_l g_l
		prefsuf<SPL, G>(fm, M.W, M.U, mk, Pf, Sf); // Pf = sum_{k<l} W_k f_k, Sf = sum_{k>l} U_k f_k
		if (trans) {
			const double inv = fast_rcp(sm);
#pragma unroll
			for (int i = 0; i < SPL; ++i) {
				aRL[i] = fma(fm[i], Pg[i], aRL[i]);
				aRU[i] = fma(fm[i], Sg[i], aRU[i]);
				aAD[i] = fma(fm[i], g[i], aAD[i]);
				aCL[i] = fma(g[i], Sf[i], aCL[i]);
				aCU[i] = fma(g[i], Pf[i], aCU[i]);
				const double bb = fma(M.U[i], Pg[i], fma(M.W[i], Sg[i], M.D[i] * g[i]));
				b[i] = bb * inv;
				fu[i] = fm[i];
			}
			su = sm;
		}
	}
	if (valid) {
		double *po = part_c + s0;
#pragma unroll
		for (int i = 0; i < SPL; ++i) {
			po[S_E0 * NP + i] = aE0[i];
			po[S_E1 * NP + i] = aE1[i];
			po[S_RL * NP + i] = aRL[i] * M.U[i];
			po[S_CL * NP + i] = aCL[i] * M.V[i];
			po[S_RU * NP + i] = aRU[i] * M.W[i];
			po[S_CU * NP + i] = aCU[i] * M.Z[i];
			po[S_AD * NP + i] = aAD[i] * M.D[i];
		}
	}
}

// ------------------------------------------------------------------------------------------------
// Second-generation backward over a chunk (same contract as backward_chunk, except for which chunk owns which
// emission count: gamma_{u-1} = f_{u-1} (A diag(e_{x_u}) b_u) is accumulated with the TRANSITION u-1 -> u, which the
// chunk of bin u owns -- this covers bins 0..L-2 exactly once, as khmm.c:310-318 does, and needs neither f_u nor s_u).
// The f rows come from a ring of PF prefetched rows with static slots (the loop is unrolled PF times), the four scans
// of a bin are two DualScans, and the scans of f_{u-1} do not depend on b, so they overlap the chain through b.
// ------------------------------------------------------------------------------------------------
template <int SPL, int G>
struct BackwardRun {
#ifndef PSMC_BWD_PF
#define PSMC_BWD_PF 4
#endif
	static constexpr int NP = SPL * G, PF = PSMC_BWD_PF;
	const Chunk &ch;
	const LaneModel<SPL> &M;
	const bool valid;
	const int gl, s0, ulast;
	const uint32_t *__restrict__ obs;
	const double *__restrict__ frow, *__restrict__ srow; // row of bin ulast
	double *__restrict__ bsave_c;
	const int usave;
	double *__restrict__ grow; // dense-count option: g_u = e_{x_u} b_u of every transition goes to ghat (this is the row of bin ulast), or nullptr
	DualScan<G> ds;
	double aE0[SPL], aE1[SPL], aRL[SPL], aCL[SPL], aRU[SPL], aCU[SPL], aAD[SPL];
	double nf[PF][SPL], ns[PF];
	uint32_t word, wprev;
	int xu;

	__device__ __forceinline__ BackwardRun(const Chunk &ch_, bool valid_, const LaneModel<SPL> &M_, int gl_, const uint32_t *__restrict__ obs_,
	                                       const double *__restrict__ fhat, const double *__restrict__ sc, double *__restrict__ bsave_c_, int usave_,
	                                       double *__restrict__ ghat)
	    : ch(ch_), M(M_), valid(valid_), gl(gl_), s0(gl_ * SPL), ulast(ch_.u0 + ch_.len - 1), obs(obs_),
	      frow(fhat + ((size_t)ch_.gb0 + (ch_.len - 1)) * NP + gl_ * SPL), srow(sc + ch_.gb0 + (ch_.len - 1)), bsave_c(bsave_c_), usave(usave_),
	      grow(ghat ? ghat + ((size_t)ch_.gb0 + (ch_.len - 1)) * NP + gl_ * SPL : nullptr)
	{
	}

	template <int J>
	__device__ __forceinline__ void step(int t, double (&b)[SPL])
	{
		const int u = ulast - t;
		const bool act = valid && u >= ch.u0;
		const bool trans = act && u > 0; // no transition into the first bin of a sequence
		if (act && u == usave && bsave_c) store_vec<SPL>(bsave_c + s0, b); // warm start of the left neighbour's next overlap
		// row u-1 from the ring, then refill the slot with row u-1-PF
		double fm[SPL];
#pragma unroll
		for (int i = 0; i < SPL; ++i) fm[i] = nf[J][i];
		const double sm = ns[J];
		if (act && u - 1 - PF >= 0) {
			const size_t back = (size_t)(t + 1 + PF);
			load_vec<SPL>(frow - back * NP, nf[J]);
			ns[J] = __ldg(srow - back);
		}
		// symbol of bin u-1 (the emission counts of this transition belong to it)
		const int v = u - 1;
		int xm = 2;
		if (trans) {
			if (t > 0 && (v & 15) == 15) {
				word = wprev;
				wprev = __ldg(obs + ch.ow0 + max((v >> 4) - 1, 0));
			}
			xm = (word >> ((v & 15) * 2)) & 3;
		}
		double g[SPL], Pg[SPL], Sg[SPL], Pf[SPL], Sf[SPL], c0, c1;
		emis_coef(xu, c0, c1);
#pragma unroll
		for (int i = 0; i < SPL; ++i) g[i] = fma(c1, M.e0[i], c0) * b[i];
		prefsuf2<SPL, G>(g, M.V, M.Z, ds, Pg, Sg);  // Pg = sum_{l<k} V_l g_l, Sg = sum_{l>k} Z_l g_l
		prefsuf2<SPL, G>(fm, M.W, M.U, ds, Pf, Sf); // Pf = sum_{k<l} W_k f_k, Sf = sum_{k>l} U_k f_k
		if (trans) {
			if (grow) store_vec<SPL>(grow - (size_t)t * NP, g);
			const double inv = fast_rcp(sm);
			const double w0 = (xm == 0) ? 1.0 : 0.0, w1 = (xm == 1) ? 1.0 : 0.0;
#pragma unroll
			for (int i = 0; i < SPL; ++i) {
				aRL[i] = fma(fm[i], Pg[i], aRL[i]);
				aRU[i] = fma(fm[i], Sg[i], aRU[i]);
				aAD[i] = fma(fm[i], g[i], aAD[i]);
				aCL[i] = fma(g[i], Sf[i], aCL[i]);
				aCU[i] = fma(g[i], Pf[i], aCU[i]);
				const double bb = fma(M.U[i], Pg[i], fma(M.W[i], Sg[i], M.D[i] * g[i])); // = b_{u-1} s_{u-1} (khmm.c:230-234)
				const double gam = fm[i] * bb;                                         // posterior of bin u-1 (khmm.c:317)
				aE0[i] = fma(gam, w0, aE0[i]);
				aE1[i] = fma(gam, w1, aE1[i]);
				b[i] = bb * inv;
			}
			xu = xm;
		}
	}

	__device__ __forceinline__ void run(double (&b)[SPL], double *__restrict__ part_c)
	{
#pragma unroll
		for (int i = 0; i < SPL; ++i) aE0[i] = aE1[i] = aRL[i] = aCL[i] = aRU[i] = aCU[i] = aAD[i] = 0.0;
#pragma unroll
		for (int j = 0; j < PF; ++j) {
			if (ulast - 1 - j >= 0) {
				load_vec<SPL>(frow - (size_t)(1 + j) * NP, nf[j]);
				ns[j] = __ldg(srow - (1 + j));
			} else {
#pragma unroll
				for (int i = 0; i < SPL; ++i) nf[j][i] = 0.0;
				ns[j] = 1.0;
			}
		}
		ds.init(gl);
		xu = (__ldg(obs + ch.ow0 + (ulast >> 4)) >> ((ulast & 15) * 2)) & 3;
		const int v0 = max(ulast - 1, 0);
		word = __ldg(obs + ch.ow0 + (v0 >> 4));
		wprev = __ldg(obs + ch.ow0 + max((v0 >> 4) - 1, 0));
		const int trips = (warp_trips(valid ? ch.len : 0) + PF - 1) / PF * PF;
		for (int t = 0; t < trips; t += PF) {
			step<0>(t, b);
			if (PF > 1) step<(PF > 1 ? 1 : 0)>(t + 1, b);
			if (PF > 2) step<(PF > 2 ? 2 : 0)>(t + 2, b);
			if (PF > 3) step<(PF > 3 ? 3 : 0)>(t + 3, b);
		}
		if (valid) {
			double *po = part_c + s0;
#pragma unroll
			for (int i = 0; i < SPL; ++i) {
				po[S_E0 * NP + i] = aE0[i];
				po[S_E1 * NP + i] = aE1[i];
				po[S_RL * NP + i] = aRL[i] * M.U[i];
				po[S_CL * NP + i] = aCL[i] * M.V[i];
				po[S_RU * NP + i] = aRU[i] * M.W[i];
				po[S_CU * NP + i] = aCU[i] * M.Z[i];
				po[S_AD * NP + i] = aAD[i] * M.D[i];
			}
		}
	}
};

template <int SPL, int G>
__device__ __forceinline__ void backward_chunk2(const Chunk &ch, bool valid, const LaneModel<SPL> &M, double (&b)[SPL], int gl,
                                                const uint32_t *__restrict__ obs, const double *__restrict__ fhat,
                                                const double *__restrict__ sc, double *__restrict__ part_c,
                                                double *__restrict__ bsave_c = nullptr, int usave = -1, double *__restrict__ ghat = nullptr)
{
	BackwardRun<SPL, G> r(ch, valid, M, gl, obs, fhat, sc, bsave_c, usave, ghat);
	r.run(b, part_c);
}

// b_{ulast} in the reference's scaling from a direction beta: sum_k f[k] b[k] s = 1 (khmm.c:237 sanity identity)
template <int SPL, int G>
__device__ __forceinline__ void scale_boundary(const Chunk &ch, const double (&beta)[SPL], double (&b)[SPL], int gl,
                                               const double *__restrict__ fhat, const double *__restrict__ sc)
{
	constexpr int NP = SPL * G;
	double fu[SPL], dot = 0.0;
	load_vec<SPL>(fhat + ((size_t)ch.gb0 + (ch.len - 1)) * NP + gl * SPL, fu);
	const double su = __ldg(sc + ch.gb0 + (ch.len - 1));
#pragma unroll
	for (int i = 0; i < SPL; ++i) dot = fma(fu[i], beta[i], dot);
	dot = gsum<G>(dot);
	const double v = 1.0 / (su * dot);
#pragma unroll
	for (int i = 0; i < SPL; ++i) b[i] = beta[i] * v;
}

// sum-normalised copy of b to dst (direction of the backward vector at a chunk boundary)
template <int SPL, int G>
__device__ __forceinline__ void publish_direction(const double (&b)[SPL], double *__restrict__ dst, int gl, bool doit)
{
	double t = 0.0, nb[SPL];
#pragma unroll
	for (int i = 0; i < SPL; ++i) t += b[i];
	t = 1.0 / gsum<G>(t);
#pragma unroll
	for (int i = 0; i < SPL; ++i) nb[i] = b[i] * t;
	if (doit) store_vec<SPL>(dst + gl * SPL, nb);
}

// ------------------------------------------------------------------------------------------------
// K4w: the backward warm-up on its own (needs only the observations and the model, so it runs on a second stream
// concurrently with the forward pass): direction of b at the last bin of every chunk that does not end its sequence,
// from `warm` bins to the right (or from the previous E-step's saved direction), sum-normalised into bwarm[c].
template <int SPL, int G, int VER>
__global__ void __launch_bounds__(128) k_backward_warm(const Chunk *__restrict__ chunks, int n_chunks,
                                                       const uint32_t *__restrict__ obs, const double *__restrict__ model,
                                                       int warm, double *__restrict__ bwarm, const double *__restrict__ bsave_prev,
                                                       const int32_t *__restrict__ order, const int32_t *__restrict__ warm_of,
                                                       const double *__restrict__ pre_start)
{
	constexpr int NP = SPL * G;
	const GroupId<G> id(n_chunks);
	if (!__any_sync(FULLMASK, id.valid)) return;
	const int c = order ? order[id.c] : id.c, gl = id.gl, s0 = gl * SPL; // (order / warm_of: adaptive overlaps, see k_forward)
	const Chunk ch = chunks[c];
	LaneModel<SPL> M;
	M.load(model, s0, NP);
	const int ulast = ch.u0 + ch.len - 1;
	const bool is_last = (ch.flags & CH_LAST) != 0;
	double beta[SPL];
#pragma unroll
	for (int i = 0; i < SPL; ++i) beta[i] = 1.0;
	if (warm_of) warm = min(warm_of[c], warm);
	int z0 = is_last ? ulast : min(ch.Lseq - 1, ulast + warm);
	if (pre_start && !is_last && ulast + warm < ch.Lseq - 1) load_vec<SPL>(pre_start + (size_t)c * NP + s0, beta); // (k_prewarm, same condition)
	if (bsave_prev && !is_last) {
		// warm start: the direction the right neighbour saved during the PREVIOUS E-step at the bin
		// min(ulast + warm, last bin of the right neighbour) -- see usave in k_backward
		const Chunk nx = chunks[c + 1];
		z0 = min(ulast + warm, nx.u0 + nx.len - 1);
		const double *row = bsave_prev + (size_t)c * NP + s0;
#pragma unroll
		for (int i = 0; i < SPL; ++i) beta[i] = fmax(row[i], 1e-300);
	}
	const int trips = warp_trips(id.valid ? z0 - ulast : 0);
	if constexpr (VER == 2) {
		// Unnormalised recursion with a power-of-two boost when the vector has shrunk (only the direction matters); the sum
		// that decides it is taken every 16th bin and acts on the NEXT bin, off the dependency chain.  The groups of a warp
		// are aligned at the END (bin ulast + 1); a group with a shorter overlap starts late and until then computes on
		// whatever its registers hold -- no per-step select.
		DualScan<G> ds;
		ds.init(gl);
		const int mytrips = id.valid ? z0 - ulast : 0;
		const int mystart = mytrips > 0 ? trips - mytrips : INT_MAX;
		int tpend = warp_min_i(mystart);
		const int wlast = (ch.Lseq - 1) >> 4, ufirst = ulast + trips;
		// packed words: the (at most two) words of a block of <= 16 bins are fetched during the previous block, so the inner
		// loop has no load and no address arithmetic; indices are clamped to the sequence (idle groups read legal words too)
		int ina = ufirst >> 4, ia;
		uint32_t wa, wb, wna = __ldg(obs + ch.ow0 + min(max(ina, 0), wlast)), wnb = __ldg(obs + ch.ow0 + min(max(ina - 1, 0), wlast));
		double q = 1.0, bc[SPL];
#pragma unroll
		for (int i = 0; i < SPL; ++i) bc[i] = beta[i];
		int t = 0;
		while (t < trips) { // blocks of at most 16 bins: late starts and the boost decision sit between blocks
			if (t == tpend) { // warp-uniform, rare
				if (mystart == t) {
#pragma unroll
					for (int i = 0; i < SPL; ++i) bc[i] = beta[i];
				}
				tpend = warp_min_i(mystart > t ? mystart : INT_MAX);
			}
			const int tstop = min(min(trips, tpend), (t & ~15) + 16);
			wa = wna; wb = wnb; ia = ina;
			ina = (ufirst - tstop) >> 4; // the next block starts at step tstop
			wna = __ldg(obs + ch.ow0 + min(max(ina, 0), wlast));
			wnb = __ldg(obs + ch.ow0 + min(max(ina - 1, 0), wlast));
			q = (gsum<G>(local_sum<SPL>(bc)) < PSMC_BOOST_LOW) ? PSMC_BOOST_UP : 1.0;
			for (; t < tstop; ++t) {
				const int u = ufirst - t; // bin whose emission enters; the step yields the direction of bin u-1
				const int x = ((((u >> 4) == ia) ? wa : wb) >> ((u & 15) * 2)) & 3;
				double g[SPL], c0, c1;
				emis_coef(x, c0, c1);
				c0 *= q;
				c1 *= q;
				q = 1.0;
#pragma unroll
				for (int i = 0; i < SPL; ++i) g[i] = fma(c1, M.e0[i], c0) * bc[i];
				semisep2<SPL, G>(g, M.V, M.U, M.Z, M.W, M.D, ds, bc);
			}
		}
		if (mytrips > 0) {
#pragma unroll
			for (int i = 0; i < SPL; ++i) beta[i] = bc[i];
		}
		publish_direction<SPL, G>(beta, bwarm + (size_t)c * NP, gl, id.valid && !is_last);
		return;
	}
	uint32_t word = 0;
	ScanMasks<G> mk;
	mk.init(gl);
	mk.pin();
	for (int t = 0; t < trips; ++t) {
		const int u = z0 - t;
		const bool act = id.valid && u > ulast;
		const int uo = act ? u : ulast;
		if (t == 0 || (uo & 15) == 15) word = __ldg(obs + ch.ow0 + (uo >> 4));
		const int x = (word >> ((uo & 15) * 2)) & 3;
		double g[SPL], out[SPL], c0, c1;
		emis_coef(x, c0, c1);
#pragma unroll
		for (int i = 0; i < SPL; ++i) g[i] = fma(c1, M.e0[i], c0) * beta[i];
		semisep<SPL, G>(g, M.V, M.U, M.Z, M.W, M.D, mk, out);
		double scl = 1.0;
		if ((t & 7) == 7) { // exact power-of-two rescale, same factor in every lane of the group
			double tt = 0.0;
#pragma unroll
			for (int i = 0; i < SPL; ++i) tt += out[i];
			tt = gsum<G>(tt);
			if (tt > 1e-290 && tt < 1e290) scl = pow2i(-exponent_of(tt));
		}
		if (act) {
#pragma unroll
			for (int i = 0; i < SPL; ++i) beta[i] = out[i] * scl;
		}
	}
	publish_direction<SPL, G>(beta, bwarm + (size_t)c * NP, gl, id.valid && !is_last);
}

// ------------------------------------------------------------------------------------------------
// K4: backward + expected counts.  One lane group per chunk.  Per-chunk partials: part[c][S_COUNT][NP].
// The direction of b at the chunk's last bin is read from `bdir`: the boundary chain's bend (transfer mode) or the
// warm-up result bwarm (fast path, K4w).  With publish != 0 the direction computed for the last bin of chunk c-1
// goes to bexact[c-1] for the certificate; usave/bsave_next feed the optional warm start of the next E-step.
// ------------------------------------------------------------------------------------------------
template <int SPL, int G, int VER>
__global__ void __launch_bounds__(128) k_backward(const Chunk *__restrict__ chunks, int n_chunks,
                                                  const uint32_t *__restrict__ obs, const double *__restrict__ model,
                                                  const double *__restrict__ bdir, int publish, const double *__restrict__ fhat,
                                                  const double *__restrict__ sc, double *__restrict__ part,
                                                  double *__restrict__ bexact, double *__restrict__ bsave_next, int warm_next,
                                                  double *__restrict__ ghat)
{
	constexpr int NP = SPL * G;
	const GroupId<G> id(n_chunks);
	if (!__any_sync(FULLMASK, id.valid)) return;
	const int c = id.c, gl = id.gl, s0 = gl * SPL;
	const Chunk ch = chunks[c];
	LaneModel<SPL> M;
	M.load(model, s0, NP);
	const int ulast = ch.u0 + ch.len - 1;
	const bool is_last = (ch.flags & CH_LAST) != 0;
	double beta[SPL], b[SPL];
	load_vec<SPL>(bdir + (size_t)c * NP + s0, beta);
	// (the groups of a warp differ in is_last: everything containing a shuffle runs unconditionally, then selects)
	scale_boundary<SPL, G>(ch, beta, b, gl, fhat, sc);
	if (is_last) { // khmm.c:226: b_L[k] = 1/s_L
		const double v = 1.0 / __ldg(sc + ch.gb0 + (ch.len - 1));
#pragma unroll
		for (int i = 0; i < SPL; ++i) b[i] = v;
	}
	// the bin whose b the left neighbour will start its next overlap from (inside this chunk)
	const int usave = (ch.flags & CH_FIRST) ? -1 : min(ch.u0 - 1 + warm_next, ulast);
	if constexpr (VER == 2)
		backward_chunk2<SPL, G>(ch, id.valid, M, b, gl, obs, fhat, sc, part + (size_t)c * S_COUNT * NP,
		                        bsave_next ? bsave_next + (size_t)(c > 0 ? c - 1 : 0) * NP : nullptr, usave, ghat);
	else
		backward_chunk<SPL, G>(ch, id.valid, M, b, gl, obs, fhat, sc, part + (size_t)c * S_COUNT * NP,
		                       bsave_next ? bsave_next + (size_t)(c > 0 ? c - 1 : 0) * NP : nullptr, usave);
	// b now belongs to the last bin of chunk c-1: publish its direction for the certificate
	publish_direction<SPL, G>(b, bexact + (size_t)(c > 0 ? c - 1 : 0) * NP, gl, id.valid && publish && !(ch.flags & CH_FIRST));
}

// ------------------------------------------------------------------------------------------------
// K4r: backward repair round (warm-up mode), mirror image of K3r.  Boundary c (at the END of chunk c)
// fails when bwarm[c] differs from bexact[c].  The head of a run (its own boundary fails, the boundary at
// the end of chunk c+1 passes or chunk c+1 ends its sequence, so bexact[c] is final) recomputes chunk c
// from bexact[c], republishes bexact[c-1] and keeps going left while the next boundary fails.
// ------------------------------------------------------------------------------------------------
template <int SPL>
__device__ __forceinline__ double bwd_boundary_mismatch(int c, const double *__restrict__ bwarm, const double *__restrict__ bexact, int s0, int N)
{
	constexpr int NP = SPL * 32;
	double x[SPL], y[SPL];
	load_vec<SPL>(bwarm + (size_t)c * NP + s0, x);
	load_vec<SPL>(bexact + (size_t)c * NP + s0, y);
	return warp_mismatch<SPL>(x, y, s0, N);
}

template <int SPL>
__global__ void __launch_bounds__(128) k_mark_bwd(const Chunk *__restrict__ chunks, int n_chunks, int N, double eps,
                                                  const double *__restrict__ bwarm, const double *__restrict__ bexact,
                                                  int32_t *__restrict__ flag_b, unsigned long long *__restrict__ stat,
                                                  int32_t *__restrict__ pred_next, int32_t *__restrict__ warm_of, int warm_max, int32_t *__restrict__ tight)
{
	const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	if (c >= n_chunks) return;
	const int gl = threadIdx.x & 31;
	int fl = 0;
	double m = 0.0;
	const bool inner = !(chunks[c].flags & CH_LAST);
	if (inner) {
		m = bwd_boundary_mismatch<SPL>(c, bwarm, bexact, gl * SPL, N);
		fl = m > eps ? 1 : 0;
	}
	if (gl == 0) {
		flag_b[c] = fl;
		if (pred_next) pred_next[c] = fl;
		if (fl) atomicAdd(&stat[2], 1ull);
		if (warm_of && pred_next && inner) warm_of[c] = adapt_overlap(warm_of[c], m, warm_max, tight + c); // first round only
	}
}

template <int SPL, int G, int VER>
__global__ void __launch_bounds__(128) k_backward_repair(const Chunk *__restrict__ subs, int n_sub, const int32_t *__restrict__ parent,
                                                         const int32_t *__restrict__ chunk_sub0, const Chunk *__restrict__ chunks,
                                                         const uint32_t *__restrict__ obs, const double *__restrict__ model,
                                                         const int32_t *__restrict__ flag_b, const double *__restrict__ bsub,
                                                         const double *__restrict__ fhat, const double *__restrict__ sc,
                                                         double *__restrict__ partsub, double *__restrict__ bwarm,
                                                         double *__restrict__ bexact, unsigned long long *__restrict__ stat,
                                                         double *__restrict__ ghat)
{
	constexpr int NP = SPL * G;
	const GroupId<G> id(n_sub);
	const int s = id.c, gl = id.gl, s0 = gl * SPL;
	const int pc = parent[s];
	const bool valid = id.valid && flag_b[pc] != 0;
	if (!__any_sync(FULLMASK, valid)) return;
	const Chunk ch = subs[s];
	LaneModel<SPL> M;
	M.load(model, s0, NP);
	double beta[SPL], b[SPL];
	load_vec<SPL>(bsub + (size_t)s * NP + s0, beta);                                                           // exact direction at the sub-chunk's last bin (k_chain_subs)
	publish_direction<SPL, G>(beta, bwarm + (size_t)pc * NP, gl, valid && chunk_sub0[pc + 1] - 1 == s);       // the chunk boundary agrees by construction from now on
	scale_boundary<SPL, G>(ch, beta, b, gl, fhat, sc);
	if constexpr (VER == 2) backward_chunk2<SPL, G>(ch, valid, M, b, gl, obs, fhat, sc, partsub + (size_t)s * S_COUNT * NP, nullptr, -1, ghat);
	else backward_chunk<SPL, G>(ch, valid, M, b, gl, obs, fhat, sc, partsub + (size_t)s * S_COUNT * NP);
	if (gl == 0 && valid) atomicAdd(&stat[3], 1ull);
	// the first sub-chunk of a chunk ends at the boundary to chunk pc-1: publish the direction computed here
	publish_direction<SPL, G>(b, bexact + (size_t)(pc > 0 ? pc - 1 : 0) * NP, gl,
	                          valid && chunk_sub0[pc] == s && !(chunks[pc].flags & CH_FIRST));
}

// ------------------------------------------------------------------------------------------------
// K4c: final certificate of the warm-up mode.  One warp per internal chunk boundary c|c+1.
//   forward : fwarm[c+1] vs the last stored vector of chunk c;   backward: bwarm[c] vs bexact[c].
// If every boundary agrees to eps (Hilbert projective metric), the boundary vectors are a fixed point of
// the exact recursion and, by induction from the exactly known sequence start (forward) and sequence end
// (backward), every chunk was computed from the exact vector (to eps).
// cert[0] = boundaries that fail, cert[1] / cert[2] = largest forward / backward mismatch (double bits).
// ------------------------------------------------------------------------------------------------
template <int SPL>
__global__ void __launch_bounds__(128) k_certify(const Chunk *__restrict__ chunks, int n_chunks, int N,
                                                 const double *__restrict__ fhat, const double *__restrict__ fwarm,
                                                 const double *__restrict__ bwarm, const double *__restrict__ bexact,
                                                 double eps, int dir, unsigned long long *__restrict__ cert)
{
	// dir 0: forward boundaries of the forward plan; dir 1: backward boundaries of the backward plan
	const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	if (c >= n_chunks) return;
	const int gl = threadIdx.x & 31, s0 = gl * SPL;
	const Chunk ch = chunks[c];
	if (ch.flags & CH_LAST) return;
	const double m = dir == 0 ? fwd_boundary_mismatch<SPL>(chunks[c + 1], c + 1, fhat, fwarm, s0, N)
	                          : bwd_boundary_mismatch<SPL>(c, bwarm, bexact, s0, N);
	if (gl == 0) {
		if (!(m <= eps)) atomicAdd(&cert[0], 1ull);
		atomicMax(&cert[1 + dir], (unsigned long long)__double_as_longlong(m));
	}
}

// ------------------------------------------------------------------------------------------------
// Dense transition counts (option): A[k][l] = a[k][l] * C[k][l] with C = sum_u f_{u-1}[k] g_u[l] over the transitions
// (khmm.c:313-316), needed only for the constant offset hmm_Q0 of the printed QD line (khmm.c:336-340).  The backward
// kernel stores the rows g_u next to the forward spill; C is then a tall-skinny product F^T G: one block per backward
// chunk accumulates its NP x NP partial in registers (16 x 16 threads, a (NP/16)^2 tile each), rows staged through
// shared memory in slabs of 32 bins; a fixed-order reduction over the chunks follows (deterministic, weighted by the
// multiplicity of the chunk's record).  The pairing (row of bin u-1, row of bin u) never crosses a record: u >= max(u0, 1).
// ------------------------------------------------------------------------------------------------
template <int NP>
__global__ void __launch_bounds__(256) k_dense_chunk(const Chunk *__restrict__ chunks, const double *__restrict__ fhat,
                                                     const double *__restrict__ ghat, double *__restrict__ cpart)
{
	constexpr int TL = NP / 16, SLAB = 32;
	__shared__ __align__(16) double sf[SLAB][NP], sg[SLAB][NP];
	const Chunk ch = chunks[blockIdx.x];
	const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
	double acc[TL][TL];
#pragma unroll
	for (int i = 0; i < TL; ++i)
#pragma unroll
		for (int j = 0; j < TL; ++j) acc[i][j] = 0.0;
	const int ulo = max(ch.u0, 1), uhi = ch.u0 + ch.len; // transitions into bins [ulo, uhi)
	for (int ub = ulo; ub < uhi; ub += SLAB) {
		const int nb = min(SLAB, uhi - ub);
		for (int idx = threadIdx.x; idx < nb * NP; idx += 256) {
			const int r = idx / NP, k = idx % NP;
			const size_t row = (size_t)ch.gb0 + (size_t)(ub + r - ch.u0);
			sf[r][k] = fhat[(row - 1) * NP + k];
			sg[r][k] = ghat[row * NP + k];
		}
		__syncthreads();
		for (int r = 0; r < nb; ++r) {
			double fv[TL], gv[TL];
#pragma unroll
			for (int i = 0; i < TL; i += 2) { // 128-bit shared loads (TL is 2 or 4, the tiles are 16-byte aligned)
				const double2 a = *reinterpret_cast<const double2 *>(&sf[r][ty * TL + i]);
				const double2 b = *reinterpret_cast<const double2 *>(&sg[r][tx * TL + i]);
				fv[i] = a.x; fv[i + 1] = a.y;
				gv[i] = b.x; gv[i + 1] = b.y;
			}
#pragma unroll
			for (int i = 0; i < TL; ++i)
#pragma unroll
				for (int j = 0; j < TL; ++j) acc[i][j] = fma(fv[i], gv[j], acc[i][j]);
		}
		__syncthreads();
	}
	double *out = cpart + (size_t)blockIdx.x * NP * NP;
#pragma unroll
	for (int i = 0; i < TL; ++i)
#pragma unroll
		for (int j = 0; j < TL; ++j) out[(size_t)(ty * TL + i) * NP + tx * TL + j] = acc[i][j];
}

// C[e] = sum_c w[c] * cpart[c][e] in a fixed order; one block per entry e of the NP x NP matrix
__global__ void __launch_bounds__(256) k_dense_reduce(const double *__restrict__ cpart, int n_chunks, int n_entries,
                                                      const double *__restrict__ w, double *__restrict__ out)
{
	__shared__ double sh[256];
	const int e = blockIdx.x;
	double acc = 0.0;
	for (int c = threadIdx.x; c < n_chunks; c += 256) {
		const double v = cpart[(size_t)c * n_entries + e];
		acc += w ? w[c] * v : v;
	}
	sh[threadIdx.x] = acc;
	__syncthreads();
	for (int d = 128; d > 0; d >>= 1) {
		if (threadIdx.x < d) sh[threadIdx.x] += sh[threadIdx.x + d];
		__syncthreads();
	}
	if (threadIdx.x == 0) out[e] = sh[0];
}

// ------------------------------------------------------------------------------------------------
// K5: deterministic reduction of the per-chunk partials.
// out layout (7*N+1 doubles): [ LL | E0(N) E1(N) | RL(N) CL(N) RU(N) CU(N) AD(N) ]
// grid = 1 + S_COUNT*N blocks, block = 256 threads; block 0 reduces LL.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_reduce(const double *__restrict__ part, const double *__restrict__ llpart,
                                                int n_chunks, int n_part, int N, int NP, double *__restrict__ out,
                                                const double *__restrict__ w_ll, const double *__restrict__ w_part)
{
	// w_ll / w_part: multiplicity of the sequence every chunk belongs to (bootstrap replicates, aux.c:8-47); NULL = 1
	__shared__ double sh[256];
	const int o = blockIdx.x;
	double acc = 0.0;
	if (o == 0) {
		for (int c = threadIdx.x; c < n_chunks; c += 256) acc += w_ll ? w_ll[c] * llpart[c] : llpart[c];
	} else {
		const int row = (o - 1) / N, k = (o - 1) % N;
		const double *p = part + (size_t)row * NP + k;
		for (int c = threadIdx.x; c < n_part; c += 256) {
			const double v = p[(size_t)c * S_COUNT * NP];
			acc += w_part ? w_part[c] * v : v;
		}
	}
	sh[threadIdx.x] = acc;
	__syncthreads();
	for (int d = 128; d > 0; d >>= 1) {
		if (threadIdx.x < d) sh[threadIdx.x] += sh[threadIdx.x + d];
		__syncthreads();
	}
	if (threadIdx.x == 0) out[o] = sh[0];
}

// ------------------------------------------------------------------------------------------------
// K6: decode backward.  One warp per chunk of ONE sequence: posterior argmax / max, optional full
// posterior and recombination probability (aux.c:167-200, khmm.c:264-293).
// ------------------------------------------------------------------------------------------------
template <int SPL>
__global__ void __launch_bounds__(128) k_decode(const Chunk *__restrict__ chunks, int c_first, int n_chunks_seq,
                                                const uint32_t *__restrict__ obs, const double *__restrict__ model,
                                                const double *__restrict__ bend, const double *__restrict__ fhat,
                                                const double *__restrict__ sc, int N, int32_t *__restrict__ best_k,
                                                double *__restrict__ best_p, double *__restrict__ post,
                                                double *__restrict__ p_recomb)
{
	constexpr int G = 32, NP = SPL * G;
	const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	if (w >= n_chunks_seq) return;
	const int c = c_first + w;
	const int gl = threadIdx.x & 31;
	const Chunk ch = uniform_chunk(chunks[c]);
	const int s0 = gl * SPL;
	double cU[SPL], cV[SPL], cW[SPL], cZ[SPL], cD[SPL], e0[SPL], e1[SPL];
#pragma unroll
	for (int i = 0; i < SPL; ++i) {
		cU[i] = model[M_U * NP + s0 + i];
		cV[i] = model[M_V * NP + s0 + i];
		cW[i] = model[M_W * NP + s0 + i];
		cZ[i] = model[M_Z * NP + s0 + i];
		cD[i] = model[M_D * NP + s0 + i];
		e0[i] = model[M_E0 * NP + s0 + i];
		e1[i] = model[M_E1 * NP + s0 + i];
	}
	const int ulast = ch.u0 + ch.len - 1;
	const double *frow = fhat + ((size_t)ch.gb0 + (ch.len - 1)) * NP + s0;
	const double *srow = sc + ch.gb0 + (ch.len - 1);
	double fu[SPL], b[SPL], su;
	load_vec<SPL>(frow, fu);
	su = __ldg(srow);
	if (ch.flags & CH_LAST) {
		const double v = 1.0 / su;
#pragma unroll
		for (int i = 0; i < SPL; ++i) b[i] = v;
	} else {
		double beta[SPL], dot = 0.0;
		load_vec<SPL>(bend + (size_t)c * NP + s0, beta);
#pragma unroll
		for (int i = 0; i < SPL; ++i) dot = fma(fu[i], beta[i], dot);
		dot = gsum<G>(dot);
		const double v = 1.0 / (su * dot);
#pragma unroll
		for (int i = 0; i < SPL; ++i) b[i] = beta[i] * v;
	}
	uint32_t word = 0;
	ScanMasks<G> mk;
	mk.init(gl);
	for (int u = ulast; u >= ch.u0; --u) {
		if (u == ulast || (u & 15) == 15) word = __ldg(obs + ch.ow0 + (u >> 4));
		const int x = (word >> ((u & 15) * 2)) & 3;
		// posterior of bin u: gamma[k] = f*b*s (khmm.c:274); first maximum wins
		double gmaxv = -1.0;
		int garg = 0x7fffffff;
#pragma unroll
		for (int i = 0; i < SPL; ++i) {
			const double gm = fu[i] * b[i] * su;
			if (s0 + i < N) {
				if (post) post[(size_t)u * N + s0 + i] = gm;
				if (gm > gmaxv) {
					gmaxv = gm;
					garg = s0 + i;
				}
			}
		}
#pragma unroll
		for (int d = 16; d > 0; d >>= 1) {
			const double ov = __shfl_xor_sync(FULLMASK, gmaxv, d);
			const int oa = __shfl_xor_sync(FULLMASK, garg, d);
			if (ov > gmaxv || (ov == gmaxv && oa < garg)) {
				gmaxv = ov;
				garg = oa;
			}
		}
		if (gl == 0) {
			best_k[u] = garg;
			best_p[u] = gmaxv;
		}
		if (u == ch.Lseq - 1 && p_recomb && gl == 0) p_recomb[u] = 0.0;
		if (u == 0) break;
		double fm[SPL], g[SPL], out[SPL];
		load_vec<SPL>(frow - (size_t)(ulast - (u - 1)) * NP, fm);
		const double sm = __ldg(srow - (ulast - (u - 1)));
#pragma unroll
		for (int i = 0; i < SPL; ++i) {
			const double em = (x == 0) ? e0[i] : ((x == 1) ? e1[i] : 1.0);
			g[i] = em * b[i];
		}
		if (p_recomb) { // aux.c:188-193 for bin u-1: 1 - sum_l f_{u-1}[l] a[l][l] b_u[l] e_u[l]
			double t = 0.0;
#pragma unroll
			for (int i = 0; i < SPL; ++i) t = fma(fm[i] * cD[i], g[i], t);
			t = gsum<G>(t);
			if (gl == 0) p_recomb[u - 1] = 1.0 - t;
		}
		semisep<SPL, G>(g, cV, cU, cZ, cW, cD, mk, out);
		const double inv = 1.0 / sm;
#pragma unroll
		for (int i = 0; i < SPL; ++i) {
			b[i] = out[i] * inv;
			fu[i] = fm[i];
		}
		su = sm;
	}
}

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
	int slots_fwd = 0, slots_bwd = 0; // resident chunks per SM of the chosen forward / backward kernels
	int g_bww = 8;              // lanes per chunk in the backward warm-up kernel (PSMC_B200_G_BWW)
	int g2_fwd = 8, g2_bww = 8; // generation 2: lanes per chunk of the forward / backward warm-up kernels at NP <= 64 (PSMC_B200_G2_FWD / _BWW: 8 or 16)
	int side_order = 0;         // PSMC_B200_SIDE_ORDER, see launch_warm
	bool dense = false;         // psmc_b200_set_dense: the backward pass also stores the rows g_u for psmc_b200_dense_counts
	bool dense_valid = false;   // ghat holds the rows of the last E-step
	double *d_ghat = nullptr, *d_cpart = nullptr, *d_cdense = nullptr;
	int cap_cpart = 0;
	int warm32 = 0, warm32_b = 0; // bins of FP32 pre-warm-up in front of the FP64 overlaps (PSMC_B200_WARM32 / _WARM32_BWD; 0 = none)
	double *d_pre_f = nullptr, *d_pre_b = nullptr; // its results: start vectors of the forward / backward overlaps
	bool adapt = false;         // adaptive per-boundary overlaps (PSMC_B200_ADAPT=1; measured on B200: no gain -- the kernels' duration is set by
	                            // the slowest boundaries either way and the extra failures while adapting cost more than the shorter warm-ups save)
	int adapt_max_chunks = 16384; // (k_order ranks in O(n^2))
	int warm_max_b = 0;         // ceiling of an adaptive backward overlap
	int32_t *d_warm_f = nullptr, *d_warm_b = nullptr, *d_order_f = nullptr, *d_order_b = nullptr, *d_tight_f = nullptr, *d_tight_b = nullptr;
	int gen = 2;                // kernel generation (PSMC_B200_GEN=1: the Kogge-Stone kernels)
	int g_fwd = 16, g_bwd = 32; // lanes per chunk in the forward / backward kernels (PSMC_B200_G_FWD / PSMC_B200_G_BWD: 8, 16 or 32)
	long long rep_fwd_fail = 0, rep_fwd_chunks = 0, rep_bwd_fail = 0, rep_bwd_chunks = 0; // of the last run
	double mis_f = 0.0, mis_b = 0.0;
	// decode scratch (allocated on demand)
	int32_t *d_bestk = nullptr;
	double *d_bestp = nullptr, *d_post = nullptr, *d_prec = nullptr;
	int64_t dec_cap = 0, dec_post_cap = 0;
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
	              (void **)&c->d_Tsub_b, (void **)&c->d_Texsub_b,
	              (void **)&c->d_warm_f, (void **)&c->d_warm_b, (void **)&c->d_order_f, (void **)&c->d_order_b, (void **)&c->d_tight_f, (void **)&c->d_tight_b, (void **)&c->d_pre_f, (void **)&c->d_pre_b};
	for (auto q : p) { cudaFree(*q); *q = nullptr; }
	c->bytes_total -= c->bytes_plan;
	c->bytes_plan = 0;
	c->cap_chunks = c->cap_chunks_b = c->cap_sub = c->cap_sub_b = c->cap_k1 = 0;
}

static void free_ctx(psmc_b200_ctx *c)
{
	if (!c) return;
	cudaSetDevice(c->device);
	free_plan(c);
	cudaFree(c->d_obs); cudaFree(c->d_seq_c0); cudaFree(c->d_seq_nc);
	cudaFree(c->d_model); cudaFree(c->d_fhat); cudaFree(c->d_sc); cudaFree(c->d_cert);
	if (c->h_cert) cudaFreeHost(c->h_cert);
	cudaFree(c->d_stats);
	cudaFree(c->d_ghat); cudaFree(c->d_cpart); cudaFree(c->d_cdense);
	cudaFree(c->d_bestk); cudaFree(c->d_bestp); cudaFree(c->d_post); cudaFree(c->d_prec);
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
	unsigned nt = std::min<unsigned>(16, std::max<unsigned>(1, std::thread::hardware_concurrency()));
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
// adaptive overlaps start from the configured lengths, in plan order
static int reset_overlaps(psmc_b200_ctx *c)
{
	if (!c->d_warm_f) return 0;
	c->warm_max_b = std::max(c->warm_len_bwd, 2 * c->warm_len);
	std::vector<int32_t> wf((size_t)std::max(c->n_chunks, 1), c->warm_len), wb((size_t)std::max(c->n_chunks_b, 1), c->warm_len_bwd);
	std::vector<int32_t> of((size_t)std::max(c->n_chunks, 1)), ob((size_t)std::max(c->n_chunks_b, 1));
	for (size_t i = 0; i < of.size(); ++i) of[i] = (int32_t)i;
	for (size_t i = 0; i < ob.size(); ++i) ob[i] = (int32_t)i;
	CUDA_TRY(cudaStreamSynchronize(c->stream), PSMC_B200_ECUDA);
	CUDA_TRY(cudaMemcpyAsync(c->d_warm_f, wf.data(), sizeof(int32_t) * (size_t)c->n_chunks, cudaMemcpyHostToDevice, c->stream), PSMC_B200_ECUDA);
	CUDA_TRY(cudaMemcpyAsync(c->d_warm_b, wb.data(), sizeof(int32_t) * (size_t)c->n_chunks_b, cudaMemcpyHostToDevice, c->stream), PSMC_B200_ECUDA);
	CUDA_TRY(cudaMemcpyAsync(c->d_order_f, of.data(), sizeof(int32_t) * (size_t)c->n_chunks, cudaMemcpyHostToDevice, c->stream), PSMC_B200_ECUDA);
	CUDA_TRY(cudaMemcpyAsync(c->d_order_b, ob.data(), sizeof(int32_t) * (size_t)c->n_chunks_b, cudaMemcpyHostToDevice, c->stream), PSMC_B200_ECUDA);
	CUDA_TRY(cudaMemsetAsync(c->d_tight_f, 0, sizeof(int32_t) * (size_t)std::max(c->n_chunks, 1), c->stream), PSMC_B200_ECUDA);
	CUDA_TRY(cudaMemsetAsync(c->d_tight_b, 0, sizeof(int32_t) * (size_t)std::max(c->n_chunks_b, 1), c->stream), PSMC_B200_ECUDA);
	CUDA_TRY(cudaStreamSynchronize(c->stream), PSMC_B200_ECUDA);
	return 0;
}

static int replan(psmc_b200_ctx *c)
{
	const int NP = c->NP;
	c->active_bins = 0; c->n_seq_eff = 0; c->weighted = false;
	int n_active = 0;
	for (int i = 0; i < c->n_seqs; ++i) {
		if (c->mult[i] > 0) { c->active_bins += c->L[i]; ++n_active; }
		if (c->mult[i] != 1) c->weighted = true;
		c->n_seq_eff += c->mult[i];
	}
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
	auto build_plan = [&](int clen, bool is_main, std::vector<Chunk> &chunks, std::vector<Chunk> &subs,
	                      std::vector<int32_t> &sub_parent, std::vector<int32_t> &chunk_sub0, std::vector<double> &w) {
		int64_t gb = 0;
		for (int i = 0; i < c->n_seqs; ++i) {
			const int Li = c->L[i];
			const int nc = c->mult[i] > 0 ? (Li + clen - 1) / clen : 0;
			if (is_main) {
				c->seq_c0[i] = (int)chunks.size();
				c->seq_nc[i] = nc;
				c->seq_gb0[i] = gb;
			}
			for (int k = 0; k < nc; ++k) {
				Chunk ch;
				const int64_t a = (int64_t)Li * k / nc, b = (int64_t)Li * (k + 1) / nc;
				ch.seq = i;
				ch.flags = (k == 0 ? CH_FIRST : 0) | (k == nc - 1 ? CH_LAST : 0);
				ch.u0 = (int)a;
				ch.len = (int)(b - a);
				ch.gb0 = gb + a;
				ch.ow0 = c->seq_ow0[i];
				ch.Lseq = Li;
				ch.pad_ = 0;
				if (is_main && nc > 1) k1.push_back((int)chunks.size());
				chunks.push_back(ch);
				w.push_back((double)c->mult[i]);
			}
			gb += Li;
		}
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
	build_plan(chunk_len, true, c->chunks, subs, sub_parent, chunk_sub0, cw);
	build_plan(chunk_len_b, false, chunks_b, subs_b, sub_parent_b, chunk_sub0_b, cw_b);
	c->n_chunks = (int)c->chunks.size();
	c->n_sub = (int)subs.size();
	c->n_chunks_b = (int)chunks_b.size();
	c->n_sub_b = (int)subs_b.size();
	c->n_k1 = (int)k1.size();
	CUDA_TRY(cudaStreamSynchronize(c->stream), PSMC_B200_ECUDA);
	CUDA_TRY(cudaStreamSynchronize(c->stream2), PSMC_B200_ECUDA);
	if (!c->d_chunks || c->n_chunks > c->cap_chunks || c->n_chunks_b > c->cap_chunks_b || c->n_sub > c->cap_sub || c->n_sub_b > c->cap_sub_b || c->n_k1 > c->cap_k1) {
		// grow-only, with head room: a replicate never needs much more than the all-sequences plan
		free_plan(c);
		auto room = [](int n) { return n + n / 8 + 16; };
		const int cc = room(c->n_chunks), cb = room(c->n_chunks_b), cs = room(c->n_sub), csb = room(c->n_sub_b), ck = room(c->n_k1);
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
		alloc((void **)&c->d_pre_f, sizeof(double) * (size_t)cc * NP);
		alloc((void **)&c->d_pre_b, sizeof(double) * (size_t)cb * NP);
		alloc((void **)&c->d_warm_f, sizeof(int32_t) * (size_t)cc);
		alloc((void **)&c->d_warm_b, sizeof(int32_t) * (size_t)cb);
		alloc((void **)&c->d_order_f, sizeof(int32_t) * (size_t)cc);
		alloc((void **)&c->d_order_b, sizeof(int32_t) * (size_t)cb);
		alloc((void **)&c->d_tight_f, sizeof(int32_t) * (size_t)cc);
		alloc((void **)&c->d_tight_b, sizeof(int32_t) * (size_t)cb);
		alloc((void **)&c->d_cw, sizeof(double) * (size_t)cc);
		alloc((void **)&c->d_cw_b, sizeof(double) * (size_t)cb);
		c->bytes_total += c->bytes_plan;
		if (!ok) { free_plan(c); return PSMC_B200_ECUDA; }
		c->cap_chunks = cc; c->cap_chunks_b = cb; c->cap_sub = cs; c->cap_sub_b = csb; c->cap_k1 = ck;
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
	UP(c->d_seq_c0, c->seq_c0); UP(c->d_seq_nc, c->seq_nc);
#undef UP
	CUDA_TRY(cudaMemsetAsync(c->d_flag_b, 0, sizeof(int32_t) * (size_t)(c->n_chunks_b + 2), st), PSMC_B200_ECUDA);
	CUDA_TRY(cudaMemsetAsync(c->d_flag, 0, sizeof(int32_t) * (size_t)(c->n_chunks + 2), st), PSMC_B200_ECUDA);
	for (int q = 0; q < 2; ++q) { // no prediction for a new plan
		CUDA_TRY(cudaMemsetAsync(c->d_pred[q], 0, sizeof(int32_t) * (size_t)(c->n_chunks + 2), st), PSMC_B200_ECUDA);
		CUDA_TRY(cudaMemsetAsync(c->d_pred_b[q], 0, sizeof(int32_t) * (size_t)(c->n_chunks_b + 2), st), PSMC_B200_ECUDA);
	}
	CUDA_TRY(cudaMemsetAsync(c->d_vstart, 0, sizeof(double) * (size_t)std::max(c->n_chunks, 1) * NP, st), PSMC_B200_ECUDA);
	CUDA_TRY(cudaMemsetAsync(c->d_bend, 0, sizeof(double) * (size_t)std::max(c->n_chunks, 1) * NP, st), PSMC_B200_ECUDA);
	CUDA_TRY(cudaStreamSynchronize(st), PSMC_B200_ECUDA); // the host vectors above go out of scope
	{
		int rc = reset_overlaps(c);
		if (rc) return rc;
	}
	c->have_prev = false;
	c->fwd_valid = false;
	c->launched = false;
	++c->replans;
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
		env = getenv("PSMC_B200_WARM32");
		if (env && atoi(env) >= 0) c->warm32 = c->warm32_b = atoi(env);
		env = getenv("PSMC_B200_WARM32_BWD");
		if (env && atoi(env) >= 0) c->warm32_b = atoi(env);
		env = getenv("PSMC_B200_ADAPT");
		if (env) c->adapt = atoi(env) != 0;
		env = getenv("PSMC_B200_SIDE_ORDER");
		if (env) c->side_order = atoi(env);
		env = getenv("PSMC_B200_G2_FWD");
		if (env && atoi(env) == 16) c->g2_fwd = 16;
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
		if (env && atoi(env) >= 0) c->repair_rounds = atoi(env);
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
	c->seq_c0.resize(c->n_seqs); c->seq_nc.resize(c->n_seqs); c->seq_gb0.resize(c->n_seqs);
	const int NP = c->NP;
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
	c->bytes_forward = c->total_bins * NP * 8 + c->total_bins * 8;
	ALLOC(c->d_obs, c->bytes_obs);
	ALLOC(c->d_seq_c0, sizeof(int32_t) * (size_t)c->n_seqs);
	ALLOC(c->d_seq_nc, sizeof(int32_t) * (size_t)c->n_seqs);
	ALLOC(c->d_model, sizeof(double) * M_COUNT * NP);
	ALLOC(c->d_fhat, (size_t)c->total_bins * NP * 8);
	ALLOC(c->d_sc, (size_t)c->total_bins * 8);
	ALLOC(c->d_stats, sizeof(double) * (size_t)(S_COUNT * c->N + 1));
	ALLOC(c->d_cert, sizeof(unsigned long long) * 8);
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
	CTRY(cudaMallocHost((void **)&c->h_model, sizeof(double) * M_COUNT * NP));
	CTRY(cudaMallocHost((void **)&c->h_stats, sizeof(double) * (size_t)(S_COUNT * c->N + 1)));
	CTRY(cudaMallocHost((void **)&c->h_cert, sizeof(unsigned long long) * 8));
	CTRY(cudaMallocHost((void **)&c->h_obs, (size_t)c->bytes_obs));
	pack_all(c, sp.data());
	CTRY(cudaMemcpyAsync(c->d_obs, c->h_obs, (size_t)c->bytes_obs, cudaMemcpyHostToDevice, c->stream));
	CTRY(cudaStreamSynchronize(c->stream));
#undef CTRY
	{
		int rc = replan(c);
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
	if (m == c->mult) return 0;
	CUDA_TRY(cudaSetDevice(c->device), PSMC_B200_ECUDA);
	c->mult.swap(m);
	return replan(c);
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

static void stage_model(psmc_b200_ctx *c, const psmc_b200_model *m)
{
	const int N = c->N, NP = c->NP;
	double *h = c->h_model;
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

// FP32 pre-warm-up (k_prewarm): generation 2, 8-lane groups (NP <= 64), PSMC_B200_WARM32 > 0
template <int NP>
static bool prewarm_on(const psmc_b200_ctx *c) { return NP <= 64 && c->gen == 2 && c->g2_fwd == 8 && c->g2_bww == 8 && c->warm32 > 0 && c->d_pre_f != nullptr; }

template <int NP>
static void run_forward(psmc_b200_ctx *c, int warm, int use_prev)
{
	cudaStream_t st = c->stream;
	const bool ad = c->adapt && warm > 0 && !use_prev && c->n_chunks <= c->adapt_max_chunks;
	const bool pre = prewarm_on<NP>(c) && warm > 0 && !use_prev && !ad;
	if (pre) LAUNCH((k_prewarm<(NP <= 64 ? NP / 8 : 8), 0>), blocks_for(c->n_chunks, 8), 128, st, c->d_chunks, c->n_chunks, c->d_obs, c->d_model, warm, c->warm32, c->d_pre_f);
#define FWD(G_, V_) LAUNCH((k_forward<NP / G_, G_, V_>), blocks_for(c->n_chunks, G_), 128, st, c->d_chunks, c->n_chunks, c->d_obs, c->d_model, c->d_vstart, warm, use_prev, c->d_fhat, c->d_sc, c->d_llpart, c->d_fwarm, ad ? c->d_order_f : nullptr, ad ? c->d_warm_f : nullptr, pre ? c->d_pre_f : nullptr)
	if (c->gen == 2 && (c->g2_fwd == 16 || NP > 64)) FWD(16, 2);
	else if (c->gen == 2) FWD(8, 2);
	else if (c->g_fwd == 8 && NP / 8 <= 8) FWD(8, 1);
	else if (c->g_fwd <= 16 && NP / 16 <= 8) FWD(16, 1);
	else FWD(32, 1);
#undef FWD
}
template <int NP>
static void run_forward_repair(psmc_b200_ctx *c)
{
	cudaStream_t st = c->stream;
#define FWR(G_, V_) LAUNCH((k_forward_repair<NP / G_, G_, V_>), blocks_for(c->n_sub, G_), 128, st, c->d_sub, c->n_sub, c->d_sub_parent, c->d_chunk_sub0, c->d_obs, c->d_model, c->d_flag + 1, c->d_vsub, c->d_fhat, c->d_sc, c->d_llsub, c->d_fwarm, c->d_cert + 4)
	if (c->gen == 2 && (c->g2_fwd == 16 || NP > 64)) FWR(16, 2);
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
	const bool ad = c->adapt && !use_prev && c->n_chunks_b <= c->adapt_max_chunks;
	const bool pre = prewarm_on<NP>(c) && !use_prev && !ad;
	if (pre) LAUNCH((k_prewarm<(NP <= 64 ? NP / 8 : 8), 1>), blocks_for(c->n_chunks_b, 8), 128, st, c->d_chunks_b, c->n_chunks_b, c->d_obs, c->d_model, warm, c->warm32_b, c->d_pre_b);
#define BWW(G_, V_) LAUNCH((k_backward_warm<NP / G_, G_, V_>), blocks_for(c->n_chunks_b, G_), 128, st, c->d_chunks_b, c->n_chunks_b, c->d_obs, c->d_model, warm, c->d_bwarm, use_prev ? c->d_bsave[c->bsave_cur] : nullptr, ad ? c->d_order_b : nullptr, ad ? c->d_warm_b : nullptr, pre ? c->d_pre_b : nullptr)
	if (c->gen == 2 && (c->g2_bww == 16 || NP > 64)) BWW(16, 2);
	else if (c->gen == 2) BWW(8, 2);
	else if (c->g_bww == 8 && NP / 8 <= 8) BWW(8, 1);
	else if (c->g_bww <= 16 && NP / 16 <= 8) BWW(16, 1);
	else BWW(32, 1);
#undef BWW
}
template <int NP>
static void run_backward_repair(psmc_b200_ctx *c)
{
	cudaStream_t st = c->stream;
#define BWR(G_, V_) LAUNCH((k_backward_repair<NP / G_, G_, V_>), blocks_for(c->n_sub_b, G_), 128, st, c->d_sub_b, c->n_sub_b, c->d_sub_parent_b, c->d_chunk_sub0_b, c->d_chunks_b, c->d_obs, c->d_model, c->d_flag_b + 1, c->d_bsub, c->d_fhat, c->d_sc, c->d_partsub, c->d_bwarm, c->d_bexact, c->d_cert + 4, (c->dense && V_ == 2) ? c->d_ghat : nullptr)
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
		if (c->gen == 2) { gb = Gen2<NP>::G_BWD; OCC(k_backward, Gen2<NP>::G_BWD, 2, bb); done = true; }
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
		LAUNCH((k_transfer<SPL1, G1, COLS>), grid, COLS * G1, st, c->d_chunks, c->d_k1, c->d_obs, c->d_model, c->d_T, c->d_Tex, c->N, nullptr, 0, nullptr, c->n_k1, nullptr, 0);
		++c->launches;
	}
	cudaEventRecord(c->ev[1], st);
	if (c->n_k1 > 0) {
		LAUNCH((k_chain<NP>), 2 * c->n_seqs, NP, st, c->d_seq_c0, c->d_seq_nc, c->d_T, c->d_Tex, c->d_model, c->d_vstart, c->d_bend, c->n_seqs);
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
		LAUNCH((k_reduce), 1 + S_COUNT * c->N, 256, st, c->d_part, c->d_llpart, c->n_chunks, c->n_chunks, c->N, NP, c->d_stats, c->weighted ? c->d_cw : nullptr, c->weighted ? c->d_cw : nullptr);
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
	cudaMemsetAsync(c->d_cert, 0, sizeof(unsigned long long) * 8, st);
	cudaEventRecord(c->ev[0], st);
	cudaEventRecord(c->ev[1], st);
	cudaEventRecord(c->ev[2], st);
	const int wpb = 4, nblk = (c->n_chunks + wpb - 1) / wpb;
	const int hot = (c->have_prev && c->warm_hot > 0) ? 1 : 0;
	const int wl = hot ? c->warm_hot : c->warm_len;
	const bool adf = c->adapt && !hot && c->n_chunks <= c->adapt_max_chunks, adb = c->adapt && !hot && c->n_chunks_b <= c->adapt_max_chunks;
	// The backward warm-up needs only observations + model: it runs on the side stream, concurrently with the forward pass;
	// on small shards (multi-GPU) a long one would become the critical path, so it is capped at the forward kernel's length.
	// The side stream also computes, ahead of time, the operators of the chunks that failed in the previous E-step.
	// Order on the side stream (PSMC_B200_SIDE_ORDER): 0 = warm-up, forward operators, backward operators;
	//                                                  1 = forward operators first (the forward repair rounds wait for them only).
	cudaEventRecord(c->ev_fork, st);
	cudaStreamWaitEvent(c->stream2, c->ev_fork, 0);
	const int wl_b = (c->warm_bwd_fixed || c->warm_len_bwd <= c->warm_len + c->chunk_len) ? c->warm_len_bwd : std::max(c->warm_len, c->warm_len + c->chunk_len);
	// adaptive overlaps may grow beyond the default, up to warm_max_b (but never beyond what hides behind the forward kernel)
	const int cap_b = adb ? std::min(c->warm_max_b, std::max(wl_b, c->warm_len + c->chunk_len)) : wl_b;
	constexpr int G1 = (NP > 64) ? 16 : 8, SPL1 = NP / G1, COLS = 128 / G1;
	const dim3 gridT((unsigned)std::min(c->n_sub, 8 * c->sm_count), NP / COLS); // block rows stride over the sub-chunks
	const int nblk_b = (c->n_chunks_b + wpb - 1) / wpb;
	const dim3 gridTb((unsigned)std::min(c->n_sub_b, 8 * c->sm_count), NP / COLS);
	const int32_t *pred = c->predict ? c->d_pred[c->pred_cur] + 1 : nullptr, *pred_b = c->predict ? c->d_pred_b[c->pred_cur] + 1 : nullptr;
	int32_t *pred_next = c->d_pred[c->pred_cur ^ 1] + 1, *pred_next_b = c->d_pred_b[c->pred_cur ^ 1] + 1;
	auto side_k1f = [&]() {
		LAUNCH((k_transfer<SPL1, G1, COLS>), gridT, COLS * G1, c->stream2, c->d_sub, c->d_sub_parent, c->d_obs, c->d_model, c->d_Tsub, c->d_Texsub, c->N, pred, 3, nullptr, c->n_sub, c->d_chunk_sub0, 0);
		cudaEventRecord(c->ev_k1f, c->stream2);
	};
	auto side_warm = [&]() {
		run_backward_warm<NP>(c, c->stream2, hot ? c->warm_hot : cap_b, hot);
		cudaEventRecord(c->ev_join, c->stream2);
	};
	if (c->side_order == 1) {
		run_forward<NP>(c, wl, hot);
		cudaEventRecord(c->ev[6], st); // forward kernel done (the repair rounds follow)
		if (c->predict) side_k1f();
		side_warm();
	} else {
		side_warm();
		run_forward<NP>(c, wl, hot);
		cudaEventRecord(c->ev[6], st);
		if (c->predict) side_k1f();
	}
	if (c->predict) {
		LAUNCH((k_transfer<SPL1, G1, COLS>), gridTb, COLS * G1, c->stream2, c->d_sub_b, c->d_sub_parent_b, c->d_obs, c->d_model, c->d_Tsub_b, c->d_Texsub_b, c->N, pred_b, 3, nullptr, c->n_sub_b, c->d_chunk_sub0_b, 1);
		cudaEventRecord(c->ev_k1b, c->stream2);
		cudaStreamWaitEvent(st, c->ev_k1f, 0);
	}
	for (int r = 0; r < c->repair_rounds; ++r) {
		LAUNCH((k_mark_fwd<SPL>), nblk, wpb * 32, st, c->d_chunks, c->n_chunks, c->N, c->cert_eps, c->d_fhat, c->d_fwarm, c->d_flag + 1, c->d_cert + 4, r == 0 ? pred_next : nullptr, adf ? c->d_warm_f : nullptr, c->warm_len, c->d_tight_f);
		LAUNCH((k_transfer<SPL1, G1, COLS>), gridT, COLS * G1, st, c->d_sub, c->d_sub_parent, c->d_obs, c->d_model, c->d_Tsub, c->d_Texsub, c->N, c->d_flag + 1, 3, r == 0 ? pred : nullptr, c->n_sub, c->d_chunk_sub0, 0);
		LAUNCH((k_chain_subs<NP>), c->n_chunks, NP, st, c->d_sub, c->n_sub, c->d_sub_parent, c->d_chunk_sub0, c->d_flag + 1, 0, c->d_Tsub, c->d_Texsub, c->d_fhat, c->d_bexact, c->d_vsub, c->d_bsub);
		run_forward_repair<NP>(c);
		LAUNCH((k_fold), c->n_chunks, 128, st, c->d_chunk_sub0, c->d_flag + 1, 0, NP, c->d_llsub, c->d_llpart, c->d_partsub, c->d_part);
	}
	cudaEventRecord(c->ev[3], st);
	cudaStreamWaitEvent(st, c->ev_join, 0);
	run_backward<NP>(c, c->d_chunks_b, c->n_chunks_b, c->d_bwarm, 1, c->d_bsave[c->bsave_cur ^ 1]);
	cudaEventRecord(c->ev[7], st); // backward kernel done
	c->bsave_cur ^= 1;
	if (c->predict) cudaStreamWaitEvent(st, c->ev_k1b, 0);
	for (int r = 0; r < c->repair_rounds; ++r) {
		LAUNCH((k_mark_bwd<SPL>), nblk_b, wpb * 32, st, c->d_chunks_b, c->n_chunks_b, c->N, c->cert_eps, c->d_bwarm, c->d_bexact, c->d_flag_b + 1, c->d_cert + 4, r == 0 ? pred_next_b : nullptr, adb ? c->d_warm_b : nullptr, cap_b, c->d_tight_b);
		LAUNCH((k_transfer<SPL1, G1, COLS>), gridTb, COLS * G1, st, c->d_sub_b, c->d_sub_parent_b, c->d_obs, c->d_model, c->d_Tsub_b, c->d_Texsub_b, c->N, c->d_flag_b + 1, 3, r == 0 ? pred_b : nullptr, c->n_sub_b, c->d_chunk_sub0_b, 1);
		LAUNCH((k_chain_subs<NP>), c->n_chunks_b, NP, st, c->d_sub_b, c->n_sub_b, c->d_sub_parent_b, c->d_chunk_sub0_b, c->d_flag_b + 1, 1, c->d_Tsub_b, c->d_Texsub_b, c->d_fhat, c->d_bexact, c->d_vsub, c->d_bsub);
		run_backward_repair<NP>(c);
		LAUNCH((k_fold), c->n_chunks_b, 128, st, c->d_chunk_sub0_b, c->d_flag_b + 1, 1, NP, c->d_llsub, c->d_llpart, c->d_partsub, c->d_part);
	}
	cudaEventRecord(c->ev[4], st);
	LAUNCH((k_reduce), 1 + S_COUNT * c->N, 256, st, c->d_part, c->d_llpart, c->n_chunks, c->n_chunks_b, c->N, NP, c->d_stats, c->weighted ? c->d_cw : nullptr, c->weighted ? c->d_cw_b : nullptr);
	LAUNCH((k_certify<SPL>), nblk, wpb * 32, st, c->d_chunks, c->n_chunks, c->N, c->d_fhat, c->d_fwarm, c->d_bwarm, c->d_bexact, c->cert_eps, 0, c->d_cert);
	LAUNCH((k_certify<SPL>), nblk_b, wpb * 32, st, c->d_chunks_b, c->n_chunks_b, c->N, c->d_fhat, c->d_fwarm, c->d_bwarm, c->d_bexact, c->cert_eps, 1, c->d_cert);
	// adaptive overlaps: sort the chunks by their new number of steps for the next E-step (two tiny launches)
	if (adf) LAUNCH((k_order), (c->n_chunks + 255) / 256, 256, st, c->d_chunks, c->n_chunks, c->d_warm_f, c->warm_len, CH_FIRST, c->d_order_f);
	if (adb) LAUNCH((k_order), (c->n_chunks_b + 255) / 256, 256, st, c->d_chunks_b, c->n_chunks_b, c->d_warm_b, cap_b, CH_LAST, c->d_order_b);
	cudaEventRecord(c->ev[5], st);
	c->launches = 6 + 10 * c->repair_rounds + (c->predict ? 2 : 0) + (adf ? 1 : 0) + (adb ? 1 : 0) + ((prewarm_on<NP>(c) && !hot && !adf) ? 1 : 0) + ((prewarm_on<NP>(c) && !hot && !adb) ? 1 : 0);
	c->pred_cur ^= 1;
	cudaError_t e = cudaGetLastError();
	if (e != cudaSuccess) return set_err(PSMC_B200_ECUDA, "kernel launch failed: %s", cudaGetErrorString(e));
	c->fwd_valid = false; // bend[] is not filled in this mode; decode runs its own forward pass
	c->mode_warm = true;
	c->certified = false;
	c->have_prev = true;
	return 0;
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

extern "C" int psmc_b200_estep_launch(psmc_b200_ctx *c, const psmc_b200_model *model)
{
	if (!c) return set_err(PSMC_B200_EINVAL, "ctx is NULL");
	int rc = check_model(c, model);
	if (rc) return rc;
	CUDA_TRY(cudaSetDevice(c->device), PSMC_B200_ECUDA);
	CUDA_TRY(cudaStreamSynchronize(c->stream), PSMC_B200_ECUDA); // pinned staging buffer is reused
	stage_model(c, model);
	CUDA_TRY(cudaMemcpyAsync(c->d_model, c->h_model, sizeof(double) * M_COUNT * c->NP, cudaMemcpyHostToDevice, c->stream), PSMC_B200_ECUDA);
	rc = launch_dispatch(c, true);
	if (rc) return rc;
	c->launched = true;
	c->dense_valid = c->dense;
	return 0;
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
	if (need) CUDA_TRY(cudaMemcpyAsync(c->h_cert, c->d_cert, sizeof(unsigned long long) * 8, cudaMemcpyDeviceToHost, c->stream), PSMC_B200_ECUDA);
	CUDA_TRY(cudaStreamSynchronize(c->stream), PSMC_B200_ECUDA);
	if (need) {
		c->certified = true;
		long long bf = (long long)c->h_cert[1], bb = (long long)c->h_cert[2];
		memcpy(&c->mis_f, &bf, sizeof(double));
		memcpy(&c->mis_b, &bb, sizeof(double));
		c->rep_fwd_fail = (long long)c->h_cert[4]; c->rep_fwd_chunks = (long long)c->h_cert[5];
		c->rep_bwd_fail = (long long)c->h_cert[6]; c->rep_bwd_chunks = (long long)c->h_cert[7];
		if (c->h_cert[0] > 0) {
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

extern "C" int psmc_b200_estep_finish(psmc_b200_ctx *c, int64_t n_seqs_total, psmc_b200_stats *out)
{
	if (!c) return set_err(PSMC_B200_EINVAL, "ctx is NULL");
	if (!c->launched) return set_err(PSMC_B200_EINVAL, "estep_finish without estep_launch");
	CUDA_TRY(cudaSetDevice(c->device), PSMC_B200_ECUDA);
	const int n = S_COUNT * c->N + 1;
	int rc = sync_and_certify(c);
	if (rc) return rc;
	CUDA_TRY(cudaMemcpyAsync(c->h_stats, c->d_stats, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream), PSMC_B200_ECUDA);
	CUDA_TRY(cudaStreamSynchronize(c->stream), PSMC_B200_ECUDA);
	collect_times(c, true);
	c->launched = false;
	if (n_seqs_total < 0) n_seqs_total = c->n_seq_eff;
	return psmc_b200_unpack_stats(c->N, c->h_stats, n_seqs_total, out);
}

extern "C" int psmc_b200_estep_fetch_raw(psmc_b200_ctx *c, double *raw)
{
	if (!c || !raw) return set_err(PSMC_B200_EINVAL, "NULL argument");
	if (!c->launched) return set_err(PSMC_B200_EINVAL, "estep_fetch_raw without estep_launch");
	CUDA_TRY(cudaSetDevice(c->device), PSMC_B200_ECUDA);
	const int n = S_COUNT * c->N + 1;
	int rc = sync_and_certify(c);
	if (rc) return rc;
	CUDA_TRY(cudaMemcpyAsync(c->h_stats, c->d_stats, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream), PSMC_B200_ECUDA);
	CUDA_TRY(cudaStreamSynchronize(c->stream), PSMC_B200_ECUDA);
	collect_times(c, true);
	c->launched = false;
	memcpy(raw, c->h_stats, sizeof(double) * n);
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
	if (seq_id < 0 || seq_id >= c->n_seqs) return set_err(PSMC_B200_EINVAL, "seq_id out of range");
	if (!best_k || !best_p) return set_err(PSMC_B200_EINVAL, "best_k/best_p are NULL");
	if (c->mult[seq_id] <= 0) return set_err(PSMC_B200_EINVAL, "sequence %d has multiplicity 0 (psmc_b200_set_multiplicity)", seq_id);
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

// ---- dense transition counts (option) ----
extern "C" int psmc_b200_set_dense(psmc_b200_ctx *c, int32_t on)
{
	if (!c) return set_err(PSMC_B200_EINVAL, "ctx is NULL");
	if (on && (c->NP > 64 || c->gen != 2))
		return set_err(PSMC_B200_EINVAL, "dense counts need the generation-2 backward kernels (at most 64 states)");
	CUDA_TRY(cudaSetDevice(c->device), PSMC_B200_ECUDA);
	if (on && !c->d_ghat) {
		const size_t bytes = (size_t)std::max<int64_t>(c->total_bins, 1) * c->NP * sizeof(double);
		CUDA_TRY(cudaMalloc((void **)&c->d_ghat, bytes), PSMC_B200_ECUDA);
		CUDA_TRY(cudaMemsetAsync(c->d_ghat, 0, bytes, c->stream), PSMC_B200_ECUDA);
		CUDA_TRY(cudaMalloc((void **)&c->d_cdense, sizeof(double) * (size_t)c->NP * c->NP), PSMC_B200_ECUDA);
		CUDA_TRY(cudaStreamSynchronize(c->stream), PSMC_B200_ECUDA);
		c->bytes_total += (int64_t)bytes;
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
		int rc = reset_overlaps(c);
		if (rc) return rc;
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
	return 0;
}
