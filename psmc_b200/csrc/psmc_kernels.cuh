// psmc_kernels.cuh -- device code of the B200 (sm_100a) E-step: lane-group primitives, the chunk kernels of both
// generations, transfer operators, boundary chains, certificate, reductions, decode, dense counts.
// Included by psmc_estep.cu only (which holds the C ABI and the host orchestration); see the header comment there and
// DESIGN.md for the algorithm.  Reference being replaced: khmm.c:145-324 (hmm_forward / hmm_backward / hmm_lk / hmm_expect)
// and aux.c:150-201 (posterior decoding) of lh3/psmc -- nothing here is translated from it.
#pragma once
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
	int32_t rep;     // which model of the batch this chunk runs under (0 unless psmc_b200_set_batch: bootstrap replicates side by side)
};
#define CH_FIRST 1
#define CH_LAST 2

// model arrays on the device, each NP doubles, contiguous: a0 e0 e1 U V W Z D; one such block per model of the batch
enum { M_A0 = 0, M_E0, M_E1, M_U, M_V, M_W, M_Z, M_D, M_COUNT };
#define MODEL_OF(model, ch, NP_) ((model) + (size_t)(ch).rep * (M_COUNT * (NP_)))
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
                                                      int n_items, const int32_t *__restrict__ chunk_sub0, int dir,
                                                      const unsigned long long *__restrict__ round_cnt)
{
	constexpr int NP = SPL * G;
	if (round_cnt && *round_cnt == 0) return; // nothing flagged in this repair round
	// sel 0: chunk list (transfer mode), one block row per listed chunk.  sel 3: repair rounds of the warm-up mode: `chunks`
	// is the SUB-chunk table; the block rows stride over it and work on the sub-chunks whose parent chunk is flagged (flag has
	// guard entries at -1 and n) -- a few hundred of ~19 000, so a block row per sub-chunk would spend 0.1 ms per launch on
	// scheduling empty blocks (measured).  k1_list = parent chunk of every sub-chunk in this mode; `skip` marks parents whose
	// operators are already there (computed ahead of time from the previous E-step's failures, see launch_warm).
	for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
	const int c = sel == 0 ? k1_list[item] : item;
	if (sel == 3 && (!op_needed(flag, k1_list, chunk_sub0, c, dir) || (skip && op_needed(skip, k1_list, chunk_sub0, c, dir)))) continue;
	const Chunk ch = chunks[c];
	const double *__restrict__ mdl = MODEL_OF(model, ch, NP);
	const int gl = threadIdx.x % G;
	const int col = blockIdx.y * COLS + threadIdx.x / G;
	double cU[SPL], cV[SPL], cW[SPL], cZ[SPL], cD[SPL], e0[SPL], f[SPL];
	const int s0 = gl * SPL;
#pragma unroll
	for (int i = 0; i < SPL; ++i) {
		cU[i] = mdl[M_U * NP + s0 + i];
		cV[i] = mdl[M_V * NP + s0 + i];
		cW[i] = mdl[M_W * NP + s0 + i];
		cZ[i] = mdl[M_Z * NP + s0 + i];
		cD[i] = mdl[M_D * NP + s0 + i];
		e0[i] = mdl[M_E0 * NP + s0 + i];
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
	// four independent accumulators: the chain kernels are pure latency (one block walks a run of operators), and a single
	// accumulator serialises NP dependent FMAs behind NP loads
	double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll 4
	for (int j = 0; j < NP; j += 4) {
		a0 = fma(__ldg(Tc + (size_t)j * NP + i), vec[j], a0);
		a1 = fma(__ldg(Tc + (size_t)(j + 1) * NP + i), vec[j + 1], a1);
		a2 = fma(__ldg(Tc + (size_t)(j + 2) * NP + i), vec[j + 2], a2);
		a3 = fma(__ldg(Tc + (size_t)(j + 3) * NP + i), vec[j + 3], a3);
	}
	const double acc = (a0 + a1) + (a2 + a3);
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
	double d0 = 0.0, d1 = 0.0, d2 = 0.0, d3 = 0.0;
#pragma unroll 4
	for (int j = 0; j < NP; j += 4) {
		d0 = fma(__ldg(col + j), vec[j], d0);
		d1 = fma(__ldg(col + j + 1), vec[j + 1], d1);
		d2 = fma(__ldg(col + j + 2), vec[j + 2], d2);
		d3 = fma(__ldg(col + j + 3), vec[j + 3], d3);
	}
	const double d = (d0 + d1) + (d2 + d3);
	const int ex = Texc[i];
	int e = (d > 0.0) ? ex + ilogb(d) : INT_MIN;
	const int emax = block_max_i<NP>(e, redi);
	return (d > 0.0) ? scalbn(d, ex - emax) : 0.0;
}

template <int NP>
__global__ void __launch_bounds__(NP) k_chain(const Chunk *__restrict__ chunks, const int32_t *__restrict__ seq_c0, const int32_t *__restrict__ seq_nc,
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
		double v = MODEL_OF(model, chunks[c0], NP)[M_A0 * NP + i];
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
                                                   double *__restrict__ vsub, double *__restrict__ bsub, const unsigned long long *__restrict__ round_cnt)
{
	__shared__ double vec[NP];
	if (*round_cnt == 0) return;
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
                                              const double *__restrict__ partsub, double *__restrict__ part, const unsigned long long *__restrict__ round_cnt)
{
	if (*round_cnt == 0) return;
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
	double *pwarm;      // fwarm_c + s0, and "am I the group's first lane": made opaque (PIN_*) before the store loop, otherwise ptxas
	int lane0;          // re-derives both from SR_TID.X in every iteration (two S2R per bin, 5 % of the kernel's stall samples: ncu r02)

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
				store_vec<SPL>(st_f ? prow : pwarm, fn);
			}
			if (st_f && lane0) *psc = S1 * inv_prev * rq_cur;
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
			pwarm = fwarm_c ? fwarm_c + s0 : prow;
			lane0 = gl == 0 ? 1 : 0;
			PIN_PTR(pwarm);
			PIN_INT(lane0);
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
                                                 const int32_t *__restrict__ warm_of)
{
	constexpr int NP = SPL * G;
	const GroupId<G> id(n_chunks);
	if (!__any_sync(FULLMASK, id.valid)) return;
	const int c = id.c, gl = id.gl, s0 = gl * SPL;
	const Chunk ch = chunks[c];
	LaneModel<SPL> M;
	M.load(MODEL_OF(model, ch, NP), s0, NP);
	double f[SPL];
	int ubeg = ch.u0;
	if ((ch.flags & CH_FIRST) || warm > 0) {
		if (!(ch.flags & CH_FIRST)) ubeg = max(0, ch.u0 - (warm_of ? warm_of[c] : warm)); // (warm_of: this boundary's own overlap, from the mixing probe)
		if (use_prev && ubeg > 0) {
			// warm start: the vector the PREVIOUS E-step stored for bin ubeg-1 (the parameters moved only a little since;
			// any positive vector is a legal start -- the certificate decides -- so a stale or concurrently rewritten row is harmless)
			const double *row = fhat + ((size_t)ch.gb0 - (size_t)(ch.u0 - ubeg) - 1) * NP + s0;
#pragma unroll
			for (int i = 0; i < SPL; ++i) f[i] = fmax(row[i], 1e-300);
		} else {
#pragma unroll
			for (int i = 0; i < SPL; ++i) f[i] = MODEL_OF(model, ch, NP)[M_A0 * NP + s0 + i];
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
// Mixing probe.  How many bins of overlap a boundary needs is a property of the data in front of it: the chain forgets
// its start at the local rate of its second Lyapunov exponent, ~1/120 per bin on average but 50x slower inside long
// low-TMRCA tracts.  The probe measures that rate everywhere, once in a while: next to the exact forward pass (already in
// fhat) every chunk runs a second, slightly perturbed chain h through the same bins and records, per piece of <= 64 bins,
// the logarithm of the factor by which the projective (Hilbert) distance between h and the exact vector shrank.  The
// host sums the pieces along every sequence; the overlap a boundary at bin p needs is then the shortest window in front
// of p over which the sum reaches log(eps) minus a safety margin (psmc_estep.cu, plan_from_probe) -- typically 3-5 k bins
// instead of the 12 k a fixed overlap must allow for, and known in advance for the boundaries no overlap can reach.
// Pieces end at the global rows r with (r & 63) == 63 and at the chunk's last row; piece k of chunk c goes to
// K[(gb0 >> 6) + c + k] (disjoint ranges for different chunks: no atomics, deterministic).  The perturbed chain starts
// `lead` bins in front of the chunk so that it has turned into the slowest-decaying direction when measuring starts.
// ------------------------------------------------------------------------------------------------
template <int SPL, int G>
__global__ void __launch_bounds__(128) k_probe(const Chunk *__restrict__ chunks, int n_chunks, const uint32_t *__restrict__ obs,
                                               const double *__restrict__ model, const double *__restrict__ fhat, float *__restrict__ K,
                                               int N, int lead)
{
	constexpr int NP = SPL * G;
	const GroupId<G> id(n_chunks);
	if (!__any_sync(FULLMASK, id.valid)) return;
	const int c = id.c, gl = id.gl, s0 = gl * SPL;
	const Chunk ch = chunks[c];
	LaneModel<SPL> M;
	M.load(MODEL_OF(model, ch, NP), s0, NP);
	const int uend = ch.u0 + ch.len, ub = max(0, ch.u0 - lead);
	const double kappa = 1e-3;
	double h[SPL], pat[SPL];
#pragma unroll
	for (int i = 0; i < SPL; ++i) pat[i] = 1.0 - 2.0 * (double)(s0 + i) / (double)(NP - 1); // a smooth relative perturbation across the states
	if (ub == 0) {
#pragma unroll
		for (int i = 0; i < SPL; ++i) h[i] = MODEL_OF(model, ch, NP)[M_A0 * NP + s0 + i];
	} else {
		load_vec<SPL>(fhat + ((size_t)ch.gb0 - (size_t)(ch.u0 - ub) - 1) * NP + s0, h);
	}
#pragma unroll
	for (int i = 0; i < SPL; ++i) h[i] *= fma(kappa, pat[i], 1.0);
	DualScan<G> ds;
	ds.init(gl);
	const int trips = warp_trips(id.valid ? uend - ub : 0);
	const int64_t kbase = (ch.gb0 >> 6) + c;
	double d_prev = -1.0; // distance at the previous knot (< 0: not measured yet)
	int k = 0;
	uint32_t word = 0;
	for (int t = 0; t < trips; ++t) {
		const int u = ub + t;
		const bool act = id.valid && u < uend;
		const int uo = act ? u : uend - 1;
		if (t == 0 || (uo & 15) == 0) word = __ldg(obs + ch.ow0 + (uo >> 4));
		const int x = (word >> ((uo & 15) * 2)) & 3;
		double out[SPL], c0, c1;
		semisep2<SPL, G>(h, M.W, M.Z, M.U, M.V, M.D, ds, out);
		emis_coef(x, c0, c1);
		if (act) {
#pragma unroll
			for (int i = 0; i < SPL; ++i) h[i] = ((u == 0) ? h[i] : out[i]) * fma(c1, M.e0[i], c0); // first bin of a sequence: no transition
		}
		// knots: the bin in front of the chunk (reference distance), every 64th global row, the chunk's last row
		const int64_t row = ch.gb0 + (u - ch.u0);
		const bool knot = act && (u == ch.u0 - 1 || (row & 63) == 63 || u == uend - 1); // (in front of the chunk: renormalisation only)
		if (__any_sync(FULLMASK, knot)) {
			double f[SPL], mx = 0.0, mn = 1e300, tot = 0.0;
			if (knot) load_vec<SPL>(fhat + (size_t)row * NP + s0, f);
			else {
#pragma unroll
				for (int i = 0; i < SPL; ++i) f[i] = 1.0;
			}
#pragma unroll
			for (int i = 0; i < SPL; ++i) {
				tot += h[i];
				if (s0 + i < N && f[i] > 0.0 && h[i] > 0.0) {
					const double r = h[i] * fast_rcp(f[i]);
					mx = fmax(mx, r);
					mn = fmin(mn, r);
				}
			}
#pragma unroll
			for (int d = G >> 1; d > 0; d >>= 1) {
				mx = fmax(mx, __shfl_xor_sync(FULLMASK, mx, d, G));
				mn = fmin(mn, __shfl_xor_sync(FULLMASK, mn, d, G));
			}
			tot = gsum<G>(tot);
			if (knot) {
				double dist = (mn < 1e300 && mn > 0.0) ? mx / mn - 1.0 : 1.0;
				if (!(dist > 1e-15)) dist = 1e-15;
				if (u >= ch.u0) {
					// (a chunk that starts its sequence, or has no lead, measures its first piece from the perturbation it was given)
					const double ref = d_prev > 0.0 ? d_prev : 2.0 * kappa;
					K[kbase + k] = (float)fmin(log(dist / ref), 0.0);
					++k;
				}
				d_prev = dist;
				// keep the perturbation small but far from the rounding floor: rescale the relative deviation to kappa
				const double sc_ = (dist < 1e-6) ? kappa / dist : 1.0, inv = 1.0 / tot;
#pragma unroll
				for (int i = 0; i < SPL; ++i) {
					if (s0 + i < N && f[i] > 0.0 && h[i] > 0.0 && sc_ != 1.0) h[i] = f[i] * fma(h[i] * fast_rcp(f[i] * mn) - 1.0, sc_, 1.0);
					else h[i] *= inv; // (sum-normalised: the chain is carried unnormalised between knots)
				}
				if (sc_ != 1.0) d_prev = kappa; // (the relative deviations now span [0, dist * sc_])
			}
		}
	}
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
                                                  int32_t *__restrict__ pred_next, unsigned long long *__restrict__ last_round, int spread, int round,
                                                  unsigned long long *__restrict__ round_cnt)
{
	// round_cnt: chunks flagged by THIS round -- the other kernels of the round return at once when it stays zero.
	// spread > 0 (rounds after the first): a boundary that fails NOW fails because the repair of the previous round changed
	// the vector to its left -- a cascade front.  The chunks behind it would fail one per round; chunk c is therefore also
	// flagged when one of the `spread` boundaries to its left fails (spread = overlap / chunk length: the front's reach).
	const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	if (c >= n_chunks) return;
	const int gl = threadIdx.x & 31;
	int fl = 0, own = 0;
	for (int j = 0; j <= spread && c - j >= 0; ++j) {
		const Chunk ch = chunks[c - j];
		if (ch.flags & CH_FIRST) break; // no boundary in front of the first chunk of a sequence, and nothing reaches across it
		const int f = fwd_boundary_mismatch<SPL>(ch, c - j, fhat, fwarm, gl * SPL, N) > eps ? 1 : 0;
		if (j == 0) own = f;
		fl |= f;
		if (fl) break;
	}
	if (gl == 0) {
		flag_f[c] = fl;
		if (fl) atomicAdd(round_cnt, 1ull);
		if (pred_next) pred_next[c] = fl;
		if (own) atomicAdd(&stat[0], 1ull);
		if (own) atomicMax(last_round, (unsigned long long)(round + 1)); // the deepest round that still saw a failure: the host sizes the next E-step's rounds by it
	}
}

template <int SPL, int G, int VER>
__global__ void __launch_bounds__(128) k_forward_repair(const Chunk *__restrict__ subs, int n_sub, const int32_t *__restrict__ parent,
                                                        const int32_t *__restrict__ chunk_sub0,
                                                        const uint32_t *__restrict__ obs, const double *__restrict__ model,
                                                        const int32_t *__restrict__ flag_f, const double *__restrict__ vsub,
                                                        double *__restrict__ fhat, double *__restrict__ sc,
                                                        double *__restrict__ llsub, double *__restrict__ fwarm,
                                                        unsigned long long *__restrict__ stat, const unsigned long long *__restrict__ round_cnt)
{
	constexpr int NP = SPL * G;
	if (*round_cnt == 0) return;
	const GroupId<G> id(n_sub);
	const int s = id.c, gl = id.gl, s0 = gl * SPL;
	const int pc = parent[s];
	const bool valid = id.valid && flag_f[pc] != 0;
	if (!__any_sync(FULLMASK, valid)) return;
	const Chunk ch = subs[s];
	LaneModel<SPL> M;
	M.load(MODEL_OF(model, ch, NP), s0, NP);
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

// ------------------------------------------------------------------------------------------------
// Bulk-asynchronous staging of the forward spill for the backward pass (TMA unit, 1-D bulk copies + mbarrier).
// The register prefetch ring of BackwardRun costs 40 registers and its loads share the warp's six scoreboards with the
// packed-observation loads: ncu shows ~9 % of the backward kernel's cycles in long-scoreboard stalls at the consumers
// (profiles/r02_*).  Here every chunk (lane group) owns a ring of STAGES tiles in shared memory; a tile is ROWS
// consecutive rows of fhat (ROWS x 512 B at 64 states) plus their scale factors, aligned to a multiple of ROWS in the
// global row index so that both bulk copies are 16-byte aligned and all groups of all warps cross tile boundaries in
// the same step.  The group's first lane issues cp.async.bulk (global -> shared, completion on the tile's mbarrier) two
// tiles ahead; the sixteen lanes wait on the mbarrier's phase parity and read their states with 128-bit LDS.  No
// scoreboard is involved, the loads are in flight 16+ bins ahead, and the ring lives in shared memory, not registers.
// ------------------------------------------------------------------------------------------------
#ifndef PSMC_SIMT_EMU
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init()
{
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// orders the generic-proxy reads of a stage (already consumed, warp-synchronised) before the async-proxy write that refills it
__device__ __forceinline__ void proxy_fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_load(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes),
	             "r"(smem_u32(bar))
	             : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
	uint32_t ok;
	asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
	return ok != 0;
}
#define PSMC_DYN_SMEM(name) extern __shared__ __align__(128) unsigned char name[]
#else
// host emulation of the same protocol (tests/emu): a barrier is {pending bytes, arrivals left, completed phases}
struct EmuBar { int32_t tx; int16_t arrivals, phases; };
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) { EmuBar *b = (EmuBar *)bar; b->tx = 0; b->arrivals = (int16_t)count; b->phases = 0; }
__device__ __forceinline__ void mbar_fence_init() {}
__device__ __forceinline__ void proxy_fence_async() {}
__device__ __forceinline__ void emu_bar_check(EmuBar *b, int count) { if (b->arrivals <= 0 && b->tx == 0) { b->arrivals = (int16_t)count; ++b->phases; } }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) { EmuBar *b = (EmuBar *)bar; b->tx += (int32_t)bytes; --b->arrivals; }
__device__ __forceinline__ void bulk_load(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
	EmuBar *b = (EmuBar *)bar;
	memcpy(dst, src, bytes);
	b->tx -= (int32_t)bytes;
	emu_bar_check(b, 1);
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
	EmuBar *b = (EmuBar *)bar;
	if (((uint32_t)b->phases & 1u) != parity) return true;
	simt_emu::yield(); // let the producer lane run
	return false;
}
#define PSMC_DYN_SMEM(name) unsigned char *name = simt_emu::dyn_smem()
#endif
// bounded wait: a protocol error must not hang the GPU (the flag makes the E-step fail loudly on the host)
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity, unsigned long long *err_flag)
{
	for (int i = 0; i < (1 << 22); ++i)
		if (mbar_try_wait(bar, parity)) return;
	if (err_flag) atomicAdd(err_flag, 1ull);
}

template <int SPL, int G>
struct BackwardStaged {
#ifndef PSMC_BWD_ROWS
#define PSMC_BWD_ROWS 8   /* rows per tile; the loop body is unrolled ROWS times (measured: 4 rows x 6 stages is 3 % slower although the 8-row body is 41 KB of code) */
#define PSMC_BWD_STAGES 3 /* tiles in flight per chunk: (STAGES - 1) * ROWS rows of look-ahead */
#endif
	static constexpr int NP = SPL * G, ROWS = PSMC_BWD_ROWS, STAGES = PSMC_BWD_STAGES;
	static constexpr int TILE_BYTES = ROWS * NP * 8 + ROWS * 8;      // rows, then their scale factors
	static constexpr int GROUP_BYTES = STAGES * TILE_BYTES;
	static constexpr int GROUPS_PER_BLOCK = 128 / G;
	static constexpr int SMEM_BYTES = GROUPS_PER_BLOCK * GROUP_BYTES + GROUPS_PER_BLOCK * STAGES * 8;
	const Chunk &ch;
	const LaneModel<SPL> &M;
	const bool valid;
	const int gl, s0, ulast;
	const uint32_t *__restrict__ obs;
	const double *__restrict__ fhat, *__restrict__ sc;
	double *__restrict__ bsave_c;
	const int usave;
	double *__restrict__ grow; // dense-count option (row of bin ulast), or nullptr
	unsigned char *tiles;      // this group's ring
	uint64_t *bars;            // this group's STAGES barriers
	unsigned long long *err_flag;
	DualScan<G> ds;
	double aE0[SPL], aE1[SPL], aRL[SPL], aCL[SPL], aRU[SPL], aCU[SPL], aAD[SPL];
	uint32_t word, wprev;
	int xu;
	int64_t rtop; // global row of the f row used by step 0 of the aligned walk (rtop % ROWS == ROWS - 1)
	int64_t rlo;  // lowest global row this chunk reads

	__device__ __forceinline__ BackwardStaged(const Chunk &ch_, bool valid_, const LaneModel<SPL> &M_, int gl_, const uint32_t *__restrict__ obs_,
	                                          const double *__restrict__ fhat_, const double *__restrict__ sc_, double *__restrict__ bsave_c_, int usave_,
	                                          double *__restrict__ ghat, unsigned char *smem, int group_in_block, unsigned long long *err)
	    : ch(ch_), M(M_), valid(valid_), gl(gl_), s0(gl_ * SPL), ulast(ch_.u0 + ch_.len - 1), obs(obs_), fhat(fhat_), sc(sc_), bsave_c(bsave_c_),
	      usave(usave_), grow(ghat ? ghat + ((size_t)ch_.gb0 + (ch_.len - 1)) * NP + gl_ * SPL : nullptr),
	      tiles(smem + (size_t)group_in_block * GROUP_BYTES), bars((uint64_t *)(smem + (size_t)GROUPS_PER_BLOCK * GROUP_BYTES) + group_in_block * STAGES),
	      err_flag(err)
	{
	}

	__device__ __forceinline__ void issue(int64_t tile, int stage) // called by the group's first lane
	{
		unsigned char *dst = tiles + (size_t)stage * TILE_BYTES;
		mbar_expect_tx(&bars[stage], (uint32_t)TILE_BYTES);
		bulk_load(dst, fhat + (size_t)tile * ROWS * NP, (uint32_t)(ROWS * NP * 8), &bars[stage]);
		bulk_load(dst + ROWS * NP * 8, sc + (size_t)tile * ROWS, (uint32_t)(ROWS * 8), &bars[stage]);
	}

	// one transition; fm / sm = the f row of bin u-1 and its scale factor (from the staged tile)
	__device__ __forceinline__ void step(int u, int t_of_chunk, double (&b)[SPL], const double (&fm)[SPL], double sm)
	{
		const bool act = valid && u >= ch.u0 && u <= ulast;
		const bool trans = act && u > 0; // no transition into the first bin of a sequence
		if (act && u == usave && bsave_c) store_vec<SPL>(bsave_c + s0, b);
		const int v = u - 1;
		int xm = 2;
		if (trans) {
			if (u != ulast && (v & 15) == 15) {
				word = wprev;
				wprev = __ldg(obs + ch.ow0 + max((v >> 4) - 1, 0));
			}
			xm = (word >> ((v & 15) * 2)) & 3;
		}
		double g[SPL], Pg[SPL], Sg[SPL], Pf[SPL], Sf[SPL], c0, c1;
		emis_coef(xu, c0, c1);
#pragma unroll
		for (int i = 0; i < SPL; ++i) g[i] = fma(c1, M.e0[i], c0) * b[i];
		prefsuf2<SPL, G>(g, M.V, M.Z, ds, Pg, Sg);
		prefsuf2<SPL, G>(fm, M.W, M.U, ds, Pf, Sf);
		if (trans) {
			if (grow) store_vec<SPL>(grow - (size_t)t_of_chunk * NP, g);
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

	// the same transition for a bin strictly inside its chunk (every lane group of the warp: u0 < u < ulast, u > 0, no save
	// point): no activity predicates, no branches around the accumulation
	__device__ __forceinline__ void step_interior(int u, int t_of_chunk, double (&b)[SPL], const double (&fm)[SPL], double sm)
	{
		const int v = u - 1;
		if ((v & 15) == 15) {
			word = wprev;
			wprev = __ldg(obs + ch.ow0 + max((v >> 4) - 1, 0));
		}
		const int xm = (word >> ((v & 15) * 2)) & 3;
		double g[SPL], Pg[SPL], Sg[SPL], Pf[SPL], Sf[SPL], c0, c1;
		emis_coef(xu, c0, c1);
#pragma unroll
		for (int i = 0; i < SPL; ++i) g[i] = fma(c1, M.e0[i], c0) * b[i];
		prefsuf2<SPL, G>(g, M.V, M.Z, ds, Pg, Sg);
		prefsuf2<SPL, G>(fm, M.W, M.U, ds, Pf, Sf);
		if (grow) store_vec<SPL>(grow - (size_t)t_of_chunk * NP, g);
		const double inv = fast_rcp(sm);
		const double w0 = (xm == 0) ? 1.0 : 0.0, w1 = (xm == 1) ? 1.0 : 0.0;
#pragma unroll
		for (int i = 0; i < SPL; ++i) {
			aRL[i] = fma(fm[i], Pg[i], aRL[i]);
			aRU[i] = fma(fm[i], Sg[i], aRU[i]);
			aAD[i] = fma(fm[i], g[i], aAD[i]);
			aCL[i] = fma(g[i], Sf[i], aCL[i]);
			aCU[i] = fma(g[i], Pf[i], aCU[i]);
			const double bb = fma(M.U[i], Pg[i], fma(M.W[i], Sg[i], M.D[i] * g[i]));
			const double gam = fm[i] * bb;
			aE0[i] = fma(gam, w0, aE0[i]);
			aE1[i] = fma(gam, w1, aE1[i]);
			b[i] = bb * inv;
		}
		xu = xm;
	}

	__device__ __forceinline__ void run(double (&b)[SPL], double *__restrict__ part_c)
	{
#pragma unroll
		for (int i = 0; i < SPL; ++i) aE0[i] = aE1[i] = aRL[i] = aCL[i] = aRU[i] = aCU[i] = aAD[i] = 0.0;
		ds.init(gl);
		xu = (__ldg(obs + ch.ow0 + (ulast >> 4)) >> ((ulast & 15) * 2)) & 3;
		const int v0 = max(ulast - 1, 0);
		word = __ldg(obs + ch.ow0 + (v0 >> 4));
		wprev = __ldg(obs + ch.ow0 + max((v0 >> 4) - 1, 0));
		// aligned walk over f rows: step t uses global row rtop - t (the row of bin u - 1, u = that row's bin + 1)
		const int64_t r_first = ch.gb0 + (ch.len - 1) - 1;               // f row of the chunk's first step (bin ulast)
		rtop = r_first >= 0 ? (r_first | (ROWS - 1)) : (ROWS - 1);
		rlo = ch.gb0 - (ch.u0 > 0 ? 1 : 0);                               // the left neighbour's last row is needed by bin u0
		if (rlo < 0) rlo = 0;
		const int my_tiles = (valid && r_first >= rlo) ? (int)(rtop / ROWS - rlo / ROWS + 1) : 0;
		const int n_tiles = warp_trips(my_tiles);
		const int64_t tile_top = rtop / ROWS;
		if (gl == 0) {
#pragma unroll
			for (int k = 0; k < STAGES; ++k)
				if (k < my_tiles) issue(tile_top - k, k);
		}
		__syncwarp();
		for (int k = 0; k < n_tiles; ++k) {
			const int stage = k % STAGES;
			const uint32_t parity = (uint32_t)((k / STAGES) & 1);
			if (k < my_tiles) mbar_wait(&bars[stage], parity, err_flag);
			const double *rows = (const double *)(tiles + (size_t)stage * TILE_BYTES);
			const double *scs = rows + ROWS * NP;
			{ // bulk of the chunk: all ROWS bins of this tile strictly inside the chunk, for every lane group of the warp
				const int u_hi = (int)(rtop - (int64_t)k * ROWS - ch.gb0) + ch.u0 + 1, u_lo = u_hi - (ROWS - 1);
				const bool inner = valid && k < my_tiles && u_hi < ulast && u_lo > max(ch.u0, 0) && !(bsave_c && usave >= u_lo && usave <= u_hi);
				if (!__any_sync(FULLMASK, !inner)) {
#pragma unroll
					for (int i = 0; i < ROWS; ++i) {
						double fm[SPL];
						const double *rp = rows + (size_t)(ROWS - 1 - i) * NP + s0;
#pragma unroll
						for (int q = 0; q < SPL; q += 2) {
							const double2 t2 = *reinterpret_cast<const double2 *>(rp + q);
							fm[q] = t2.x; fm[q + 1] = t2.y;
						}
						step_interior(u_hi - i, ulast - (u_hi - i), b, fm, scs[ROWS - 1 - i]);
					}
					__syncwarp();
					if (gl == 0 && k + STAGES < my_tiles) {
						proxy_fence_async();
						issue(tile_top - (k + STAGES), stage);
					}
					continue;
				}
			}
#pragma unroll
			for (int i = 0; i < ROWS; ++i) {
				const int64_t row = rtop - ((int64_t)k * ROWS + i);   // f row of this step
				const int u = (int)(row - ch.gb0) + ch.u0 + 1;        // the bin whose b is current
				double fm[SPL];
				const double *rp = rows + (size_t)(ROWS - 1 - i) * NP + s0;
				if (k < my_tiles) {
#pragma unroll
					for (int q = 0; q < SPL; q += 2) {
						const double2 t2 = *reinterpret_cast<const double2 *>(rp + q);
						fm[q] = t2.x; fm[q + 1] = t2.y;
					}
				} else {
#pragma unroll
					for (int q = 0; q < SPL; ++q) fm[q] = 0.0;
				}
				const double sm = (k < my_tiles) ? scs[ROWS - 1 - i] : 1.0;
				step(u, ulast - u, b, fm, sm);
			}
			__syncwarp(); // every lane of the group has consumed this stage: its first lane may refill it
			if (gl == 0 && k + STAGES < my_tiles) {
				proxy_fence_async();
				issue(tile_top - (k + STAGES), stage);
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
                                                       const int32_t *__restrict__ warm_of)
{
	constexpr int NP = SPL * G;
	const GroupId<G> id(n_chunks);
	if (!__any_sync(FULLMASK, id.valid)) return;
	const int c = id.c, gl = id.gl, s0 = gl * SPL;
	const Chunk ch = chunks[c];
	LaneModel<SPL> M;
	M.load(MODEL_OF(model, ch, NP), s0, NP);
	const int ulast = ch.u0 + ch.len - 1;
	const bool is_last = (ch.flags & CH_LAST) != 0;
	double beta[SPL];
#pragma unroll
	for (int i = 0; i < SPL; ++i) beta[i] = 1.0;
	if (warm_of) warm = warm_of[c]; // this boundary's own overlap (mixing probe)
	int z0 = is_last ? ulast : min(ch.Lseq - 1, ulast + warm);
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
	M.load(MODEL_OF(model, ch, NP), s0, NP);
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

// K4 with the forward spill staged through shared memory by bulk-asynchronous copies (BackwardStaged); same contract as
// k_backward<SPL, G, 2>.  Dynamic shared memory: BackwardStaged<SPL, G>::SMEM_BYTES.
template <int SPL, int G>
__global__ void __launch_bounds__(128) k_backward_staged(const Chunk *__restrict__ chunks, int n_chunks, const uint32_t *__restrict__ obs,
                                                         const double *__restrict__ model, const double *__restrict__ bdir, int publish,
                                                         const double *__restrict__ fhat, const double *__restrict__ sc, double *__restrict__ part,
                                                         double *__restrict__ bexact, double *__restrict__ bsave_next, int warm_next,
                                                         double *__restrict__ ghat, unsigned long long *__restrict__ err_flag)
{
	typedef BackwardStaged<SPL, G> BS;
	constexpr int NP = SPL * G;
	PSMC_DYN_SMEM(smem);
	{ // one barrier per (group, stage), armed by the group's first lane
		uint64_t *bars = (uint64_t *)(smem + (size_t)BS::GROUPS_PER_BLOCK * BS::GROUP_BYTES);
		if (threadIdx.x < BS::GROUPS_PER_BLOCK * BS::STAGES) mbar_init(&bars[threadIdx.x], 1);
		mbar_fence_init();
		__syncthreads();
	}
	const GroupId<G> id(n_chunks);
	if (!__any_sync(FULLMASK, id.valid)) return;
	const int c = id.c, gl = id.gl, s0 = gl * SPL;
	const Chunk ch = chunks[c];
	LaneModel<SPL> M;
	M.load(MODEL_OF(model, ch, NP), s0, NP);
	const int ulast = ch.u0 + ch.len - 1;
	const bool is_last = (ch.flags & CH_LAST) != 0;
	double beta[SPL], b[SPL];
	load_vec<SPL>(bdir + (size_t)c * NP + s0, beta);
	scale_boundary<SPL, G>(ch, beta, b, gl, fhat, sc);
	if (is_last) { // khmm.c:226: b_L[k] = 1/s_L
		const double v = 1.0 / __ldg(sc + ch.gb0 + (ch.len - 1));
#pragma unroll
		for (int i = 0; i < SPL; ++i) b[i] = v;
	}
	const int usave = (ch.flags & CH_FIRST) ? -1 : min(ch.u0 - 1 + warm_next, ulast);
	BS r(ch, id.valid, M, gl, obs, fhat, sc, bsave_next ? bsave_next + (size_t)(c > 0 ? c - 1 : 0) * NP : nullptr, usave, ghat, smem,
	     (int)(threadIdx.x / G), err_flag);
	r.run(b, part + (size_t)c * S_COUNT * NP);
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
                                                  int32_t *__restrict__ pred_next, unsigned long long *__restrict__ last_round, int spread, int round,
                                                  unsigned long long *__restrict__ round_cnt)
{
	// (spread: see k_mark_fwd; the backward cascade runs to the left, so chunk c looks at the boundaries to its right)
	const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	if (c >= n_chunks) return;
	const int gl = threadIdx.x & 31;
	int fl = 0, own = 0;
	for (int j = 0; j <= spread && c + j < n_chunks; ++j) {
		if (chunks[c + j].flags & CH_LAST) break; // no boundary behind the last chunk of a sequence
		const int f = bwd_boundary_mismatch<SPL>(c + j, bwarm, bexact, gl * SPL, N) > eps ? 1 : 0;
		if (j == 0) own = f;
		fl |= f;
		if (fl) break;
	}
	if (gl == 0) {
		flag_b[c] = fl;
		if (fl) atomicAdd(round_cnt, 1ull);
		if (pred_next) pred_next[c] = fl;
		if (own) atomicAdd(&stat[2], 1ull);
		if (own) atomicMax(last_round, (unsigned long long)(round + 1));
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
                                                         double *__restrict__ ghat, const unsigned long long *__restrict__ round_cnt)
{
	constexpr int NP = SPL * G;
	if (*round_cnt == 0) return;
	const GroupId<G> id(n_sub);
	const int s = id.c, gl = id.gl, s0 = gl * SPL;
	const int pc = parent[s];
	const bool valid = id.valid && flag_b[pc] != 0;
	if (!__any_sync(FULLMASK, valid)) return;
	const Chunk ch = subs[s];
	LaneModel<SPL> M;
	M.load(MODEL_OF(model, ch, NP), s0, NP);
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
                                                const int32_t *__restrict__ rep_f, const int32_t *__restrict__ rep_b, int N, int NP,
                                                double *__restrict__ out, const double *__restrict__ w_ll, const double *__restrict__ w_part)
{
	// w_ll / w_part: multiplicity of the sequence every chunk belongs to (bootstrap replicates, aux.c:8-47); NULL = 1.
	// blockIdx.y = model of the batch: its chunks are [rep_f[r], rep_f[r+1]) in the forward plan (log-likelihood partials)
	// and [rep_b[r], rep_b[r+1]) in the backward plan (count partials); its statistics vector is out + r * (7N+1).
	// The summation order depends only on the position of a chunk inside its own model's range, so a model gets the
	// same bits whether it runs alone or inside a batch (same chunk plan provided).
	__shared__ double sh[256];
	const int o = blockIdx.x, r = blockIdx.y;
	double acc = 0.0;
	if (o == 0) {
		const int c0 = rep_f[r], c1 = rep_f[r + 1];
		for (int c = c0 + threadIdx.x; c < c1; c += 256) acc += w_ll ? w_ll[c] * llpart[c] : llpart[c];
	} else {
		const int c0 = rep_b[r], c1 = rep_b[r + 1];
		const int row = (o - 1) / N, k = (o - 1) % N;
		const double *p = part + (size_t)row * NP + k;
		for (int c = c0 + threadIdx.x; c < c1; c += 256) {
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
	if (threadIdx.x == 0) out[(size_t)r * (S_COUNT * N + 1) + o] = sh[0];
}

// ------------------------------------------------------------------------------------------------
// K6b: decoding on the generation-2 lane groups, for EVERY chunk of the plan at once (aux.c:150-201, khmm.c:264-293).
// Runs after a complete, certified E-step on the same model: fhat / sc hold the exact forward pass and bdir[c] the exact
// direction of b at the last bin of chunk c (the fast path's certified bwarm, or the boundary chain's bend), so every
// chunk decodes independently.  Per transition u-1 -> u the posterior of bin u-1 is f_{u-1} (A diag(e_u) b_u) -- the same
// quantity the counting kernel accumulates (khmm.c:317) -- so a step is one DualScan plus an argmax over the group.
// Outputs, all optional: per bin the posterior-argmax state (uint8) and its posterior (float), the full posterior row
// (float) and the recombination probability (double, aux.c:188-193); and, written by the group's first lane, the RUNS of
// the argmax state with their maximum posterior (what `psmc -d` prints, aux.c:165-182): run j of chunk c goes to entry
// gb0 + j of the run arrays (at most one run per bin), in the order the kernel meets them, i.e. right to left.
// ------------------------------------------------------------------------------------------------
template <int SPL, int G>
struct DecodeRun {
	static constexpr int NP = SPL * G, PF = 4;
	const Chunk &ch;
	const LaneModel<SPL> &M;
	const bool valid;
	const int gl, s0, ulast, N;
	const uint32_t *__restrict__ obs;
	const double *__restrict__ frow, *__restrict__ srow; // row of bin ulast
	uint8_t *__restrict__ best_k;
	float *__restrict__ best_p, *__restrict__ post;
	double *__restrict__ p_recomb;
	int32_t *__restrict__ run_start;
	uint8_t *__restrict__ run_state;
	double *__restrict__ run_maxp;
	const int64_t obase; // output entry of bin u: gb0 - out_base + (u - u0)
	DualScan<G> ds;
	double nf[PF][SPL], ns[PF];
	uint32_t word, wprev;
	int xu, cur_state, n_runs, ulo;
	double cur_max;

	__device__ __forceinline__ DecodeRun(const Chunk &ch_, bool valid_, const LaneModel<SPL> &M_, int gl_, int N_, const uint32_t *__restrict__ obs_,
	                                     const double *__restrict__ fhat, const double *__restrict__ sc, int64_t out_base, uint8_t *bk, float *bp, float *po,
	                                     double *pr, int32_t *rs, uint8_t *rk, double *rm)
	    : ch(ch_), M(M_), valid(valid_), gl(gl_), s0(gl_ * SPL), ulast(ch_.u0 + ch_.len - 1), N(N_), obs(obs_),
	      frow(fhat + ((size_t)ch_.gb0 + (ch_.len - 1)) * NP + gl_ * SPL), srow(sc + ch_.gb0 + (ch_.len - 1)), best_k(bk), best_p(bp), post(po),
	      p_recomb(pr), run_start(rs), run_state(rk), run_maxp(rm), obase(ch_.gb0 - out_base - ch_.u0)
	{
	}

	// posterior row of bin v is in gam (this lane's states); act: the bin belongs to this group's chunk
	__device__ __forceinline__ void emit(int v, const double (&gam)[SPL], bool act)
	{
		double bv = -1.0;
		int ba = 0x7fffffff;
#pragma unroll
		for (int i = 0; i < SPL; ++i)
			if (s0 + i < N && gam[i] > bv) { // first maximum wins (khmm.c:270-275)
				bv = gam[i];
				ba = s0 + i;
			}
#pragma unroll
		for (int d = G >> 1; d > 0; d >>= 1) {
			const double ov = __shfl_xor_sync(FULLMASK, bv, d, G);
			const int oa = __shfl_xor_sync(FULLMASK, ba, d, G);
			if (ov > bv || (ov == bv && oa < ba)) {
				bv = ov;
				ba = oa;
			}
		}
		if (!act) return;
		const int64_t o = obase + v;
		if (post) {
#pragma unroll
			for (int i = 0; i < SPL; ++i)
				if (s0 + i < N) post[(size_t)o * N + s0 + i] = (float)gam[i];
		}
		if (gl == 0) {
			if (best_k) {
				best_k[o] = (uint8_t)ba;
				best_p[o] = (float)bv;
			}
			if (run_start) {
				if (v == ulast) {
					cur_state = ba;
					cur_max = bv;
				} else if (ba != cur_state) { // the run that started at bin v + 1 is complete
					const int64_t e = obase + ch.u0 + n_runs; // = gb0 - out_base + n_runs
					run_start[e] = v + 1;
					run_state[e] = (uint8_t)cur_state;
					run_maxp[e] = cur_max;
					++n_runs;
					cur_state = ba;
					cur_max = bv;
				} else if (bv > cur_max) {
					cur_max = bv;
				}
			}
		}
	}

	template <int J>
	__device__ __forceinline__ void step(int t, double (&b)[SPL])
	{
		const int u = ulast - t;            // transition u-1 -> u: yields the posterior of bin u-1
		const bool act = valid && u > ch.u0; // (bin u-1 belongs to this chunk)
		const bool actp = valid && u > ulo;  // recombination probability of bin u-1: also across the boundary to the left neighbour
		double fm[SPL];
#pragma unroll
		for (int i = 0; i < SPL; ++i) fm[i] = nf[J][i];
		const double sm = ns[J];
		if (valid && u - 1 - PF >= ulo) { // (no other rows of the left neighbour are needed)
			const size_t back = (size_t)(t + 1 + PF);
			load_vec<SPL>(frow - back * NP, nf[J]);
			ns[J] = __ldg(srow - back);
		}
		const int v = u - 1;
		int xm = 2;
		if (act) {
			if (t > 0 && (v & 15) == 15) {
				word = wprev;
				wprev = __ldg(obs + ch.ow0 + max((v >> 4) - 1, 0));
			}
			xm = (word >> ((v & 15) * 2)) & 3;
		}
		double g[SPL], Pg[SPL], Sg[SPL], gam[SPL], c0, c1;
		emis_coef(xu, c0, c1);
#pragma unroll
		for (int i = 0; i < SPL; ++i) g[i] = fma(c1, M.e0[i], c0) * b[i];
		prefsuf2<SPL, G>(g, M.V, M.Z, ds, Pg, Sg);
		const double inv = fast_rcp(act ? sm : 1.0);
		double tr = 0.0;
#pragma unroll
		for (int i = 0; i < SPL; ++i) {
			const double bb = fma(M.U[i], Pg[i], fma(M.W[i], Sg[i], M.D[i] * g[i])); // = b_{u-1} s_{u-1} (khmm.c:230-234)
			gam[i] = fm[i] * bb;                                                   // posterior of bin u-1
			tr = fma(fm[i] * M.D[i], g[i], tr);                                    // no-recombination mass (aux.c:188-191)
			if (act) b[i] = bb * inv;
		}
		if (p_recomb) {
			tr = gsum<G>(tr);
			if (actp && gl == 0) p_recomb[obase + v] = 1.0 - tr;
		}
		emit(v, gam, act);
		if (act) xu = xm;
	}

	__device__ __forceinline__ void run(double (&b)[SPL])
	{
		// lowest bin whose row is read: the chunk's first bin, or (recombination probability asked, not the first chunk of
		// its sequence) the left neighbour's last bin
		ulo = (p_recomb && !(ch.flags & CH_FIRST)) ? ch.u0 - 1 : ch.u0;
#pragma unroll
		for (int j = 0; j < PF; ++j) {
			if (valid && ulast - 1 - j >= ulo) {
				load_vec<SPL>(frow - (size_t)(1 + j) * NP, nf[j]);
				ns[j] = __ldg(srow - (1 + j));
			} else {
#pragma unroll
				for (int i = 0; i < SPL; ++i) nf[j][i] = 0.0;
				ns[j] = 1.0;
			}
		}
		ds.init(gl);
		n_runs = 0;
		cur_state = 0;
		cur_max = 0.0;
		xu = (__ldg(obs + ch.ow0 + (ulast >> 4)) >> ((ulast & 15) * 2)) & 3;
		const int v0 = max(ulast - 1, 0);
		word = __ldg(obs + ch.ow0 + (v0 >> 4));
		wprev = __ldg(obs + ch.ow0 + max((v0 >> 4) - 1, 0));
		{ // the chunk's last bin: posterior = f b s (khmm.c:274); nothing follows it inside this chunk
			double fu[SPL], gam[SPL];
			load_vec<SPL>(frow, fu);
			const double su = __ldg(srow);
#pragma unroll
			for (int i = 0; i < SPL; ++i) gam[i] = fu[i] * b[i] * su;
			if (p_recomb && valid && gl == 0 && (ch.flags & CH_LAST)) p_recomb[obase + ulast] = 0.0; // aux.c:194
			emit(ulast, gam, valid);
		}
		const int trips = (warp_trips(valid ? ulast - ulo : 0) + PF - 1) / PF * PF;
		for (int t = 0; t < trips; t += PF) {
			step<0>(t, b);
			step<1>(t + 1, b);
			step<2>(t + 2, b);
			step<3>(t + 3, b);
		}
		if (valid && gl == 0 && run_start) { // the run that reaches the chunk's first bin
			const int64_t e = obase + ch.u0 + n_runs;
			run_start[e] = ch.u0;
			run_state[e] = (uint8_t)cur_state;
			run_maxp[e] = cur_max;
			++n_runs;
		}
	}
};

template <int SPL, int G>
__global__ void __launch_bounds__(128) k_decode2(const Chunk *__restrict__ chunks, int c_first, int n_chunks, const uint32_t *__restrict__ obs,
                                                 const double *__restrict__ model, const double *__restrict__ bdir,
                                                 const double *__restrict__ fhat, const double *__restrict__ sc, int N, int64_t out_base,
                                                 uint8_t *__restrict__ best_k, float *__restrict__ best_p, float *__restrict__ post,
                                                 double *__restrict__ p_recomb, int32_t *__restrict__ run_start, uint8_t *__restrict__ run_state,
                                                 double *__restrict__ run_maxp, int32_t *__restrict__ run_count)
{
	constexpr int NP = SPL * G;
	const GroupId<G> id(n_chunks);
	if (!__any_sync(FULLMASK, id.valid)) return;
	const int c = c_first + id.c, gl = id.gl, s0 = gl * SPL;
	const Chunk ch = chunks[c];
	LaneModel<SPL> M;
	M.load(MODEL_OF(model, ch, NP), s0, NP);
	double beta[SPL], b[SPL];
	load_vec<SPL>(bdir + (size_t)c * NP + s0, beta);
	scale_boundary<SPL, G>(ch, beta, b, gl, fhat, sc); // b of the chunk's last bin in the reference's scaling
	if (ch.flags & CH_LAST) { // khmm.c:226: b_L[k] = 1/s_L
		const double v = 1.0 / __ldg(sc + ch.gb0 + (ch.len - 1));
#pragma unroll
		for (int i = 0; i < SPL; ++i) b[i] = v;
	}
	DecodeRun<SPL, G> r(ch, id.valid, M, gl, N, obs, fhat, sc, out_base, best_k, best_p, post, p_recomb, run_start, run_state, run_maxp);
	r.run(b);
	if (id.valid && gl == 0 && run_count) run_count[c] = r.n_runs;
}

// dense list of runs: chunk c's runs (stored right to left at gb0 - out_base + j) go, left to right, to off[c] ...
__global__ void __launch_bounds__(128) k_runs_gather(const Chunk *__restrict__ chunks, int n_chunks, int64_t out_base, const int32_t *__restrict__ run_count,
                                                     const int64_t *__restrict__ off, const int32_t *__restrict__ run_start,
                                                     const uint8_t *__restrict__ run_state, const double *__restrict__ run_maxp,
                                                     int32_t *__restrict__ o_seq, int32_t *__restrict__ o_start, uint8_t *__restrict__ o_state,
                                                     double *__restrict__ o_maxp)
{
	for (int c = blockIdx.x; c < n_chunks; c += gridDim.x) {
		const Chunk ch = chunks[c];
		const int n = run_count[c];
		const int64_t src = ch.gb0 - out_base, dst = off[c];
		for (int j = threadIdx.x; j < n; j += blockDim.x) {
			const int64_t e = src + (n - 1 - j);
			o_seq[dst + j] = ch.seq;
			o_start[dst + j] = run_start[e];
			o_state[dst + j] = run_state[e];
			o_maxp[dst + j] = run_maxp[e];
		}
	}
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

