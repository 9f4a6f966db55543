// Vertically implicit column solve, fast path (FP64, sm_100a): vertical
// order 1, terrain-following metric through the column constants of
// tb200_fast.cuh.
//
// Same system and same elimination as k_column_implicit /
// k_column_implicit_window (tb200_column.cuh) - reference
// VerticalDynamicsFEM::StepImplicit with PrepareColumn, BuildF,
// BuildJacobianF_LOR_RhoTheta_Pi, BuildJacobianF_Diffusion
// (VerticalDynamicsFEM.cpp:1230-1638, 1839-3187) and LAPACK dgbsv
// (dgbtf2 + dgbtrs with partial pivoting, LinearAlgebra.cpp:156-202) - with
// the generality that made those kernels latency-bound stripped away:
//
//  * one thread per unique column, the threads of a block march over the levels
//    in lockstep, so every operator coefficient is a broadcast read of the
//    per-level window table in shared memory (no dependent global loads);
//  * the 3-D metric (21 doubles per level in the general kernel) is rebuilt
//    from 10 column constants held in registers;
//  * state inputs slide through a register window and are prefetched one level
//    ahead; the rows of level k+1 are generated in registers (no staging)
//    while the 5 x 9 register block eliminates the rows of level k, dgbtf2 step
//    for step: same pivot search, interchange, reciprocal scaling, rank-1
//    update; rows that are structurally zero in the pivot column enter late;
//  * finished rows of U and the forward-substituted right-hand side stream to
//    a per-warp scratch [entry][lane] (coalesced); the solve is bound by that
//    traffic, so entries of a row that are zero in every lane of the warp are
//    not stored (the pivot order this matrix takes leaves 2-4 of the 8
//    off-diagonal entries of a row of U zero): a warp vote builds the row's mask,
//    kept in shared memory.  The back substitution reads the rows once, in
//    dtbsv order, one level prefetched ahead, and scatters x = x0 - delta to
//    the column and its duplicates.
#ifndef TB200_COLUMN_FAST_CUH
#define TB200_COLUMN_FAST_CUH

#include <type_traits>

#include "tb200_platform.h"
#include "tb200_device.h"
#include "tb200_fast.cuh"
#include "tb200_column.cuh"

#ifndef TBC_THREADS
#define TBC_THREADS 128
#endif

struct ColumnFastArgs {
	const int * col_node;   // [ncols] local node address (e*NN+n) solved
	const int * col_dups;   // [ncols][3] duplicates receiving a copy, -1 unused
	int ncols;
	int col0;
	double * ws;            // scratch [block][10 n][TBC_THREADS]
	double dt;
	double upwind_coeff;
	int * info;
	const double * colc;
	const double * lev;
	double * inc;           // optional: receives x_new - x_old (may alias the input)
};

// rows of one level in LAPACK band form relative to their own diagonal:
// entry b of row r is A(r, r - 4 + b)
struct TbcRows {
	double p[9], w[9], r[9];
	double fp, fw, fr;
};

__device__ __forceinline__ double tb_signw(double xd, double cx2) {
	// BuildJacobianF_Diffusion :2876-2884
	if (xd > 0.0) return 1.0 * cx2;
	if (xd < 0.0) return -1.0 * cx2;
	return 0.0;
}

// One batch of TBC_THREADS columns: batch `vblock` of the launch, scratch slot
// `pblock` (the block's own: a block that walks several batches reuses it, so
// that the rows of U it wrote for the previous batch - read back and dead by
// then - are overwritten while they still sit in L2 instead of being written
// back to DRAM).
__device__ __forceinline__ void tb_column_fast_batch(
	const DevLayout & lay, const DevPhys & ph, const ColumnFastArgs & ca,
	const double * in, double * out, const double * slev, unsigned * smask,
	int vblock, int pblock
) {
	const int L = lay.nlev;
	const int NN = lay.nn;
	const int n = 3 * (L + 1);
	const unsigned FULL = 0xffffffffu;

	int tcol = vblock * blockDim.x + threadIdx.x;
	const bool live = (tcol < ca.ncols);
	if (!live) tcol = ca.ncols - 1;      // keep the warp converged; no stores

	const int node = ca.col_node[ca.col0 + tcol];
	const long long e = node / NN;
	const int nd = node % NN;
	const size_t ebase = (size_t)e * lay.nrows * NN;
	const double * inU = in + ebase + (size_t)lay.rowoff[0] * NN + nd;
	const double * inV = in + ebase + (size_t)lay.rowoff[1] * NN + nd;
	const double * inP = in + ebase + (size_t)lay.rowoff[2] * NN + nd;
	const double * inW = in + ebase + (size_t)lay.rowoff[3] * NN + nd;
	const double * inR = in + ebase + (size_t)lay.rowoff[4] * NN + nd;

	// column constants
	const double * cc = ca.colc + (size_t)e * TBF_NC * NN + nd;
	const double cA0 = cc[TBF_A0 * NN], cA1 = cc[TBF_A1 * NN], cB1 = cc[TBF_B1 * NN];
	const double cJ = cc[TBF_JAC * NN];
	const double invj = 1.0 / cJ;
	const double cA2 = cc[TBF_A2 * NN], cB2 = cc[TBF_B2 * NN];
	const double cX0 = cc[TBF_X0 * NN], cX2 = cc[TBF_X2 * NN];
	const double gdxr = ph.g * cc[TBF_DXR * NN];

	const double dInvDeltaT = 1.0 / ca.dt;
	const double upw = ca.upwind_coeff;

	// scratch of this block: step j -> [10 j + c][thread], c = 0..8 row j of U,
	// c = 9 right-hand side
	// scratch of this warp: entries [slot][lane]; a step stores the pivot, the
	// right-hand side and the off-diagonal entries of its mask
	const int S = 32;
	double * sc = ca.ws
		+ ((size_t)pblock * (TBC_THREADS / 32) + (threadIdx.x >> 5)) * (size_t)(10 * n) * S
		+ (threadIdx.x & 31);
	double * scp = sc;      // next free slot

	// ---- sliding input window ------------------------------------------------------
	// levels k-1, k, k+1 (m, 0, p) and the prefetched level k+2 (q)
	double Um = 0.0, U0 = 0.0, Up = 0.0, Uq = 0.0;
	double Vm = 0.0, V0 = 0.0, Vp = 0.0, Vq = 0.0;
	double Pm = 0.0, P0 = 0.0, Pp = 0.0, Pq = 0.0;
	double Rm = 0.0, R0 = 0.0, Rp = 0.0, Rq = 0.0;
	double Wm = 0.0, W0 = 0.0, Wp = 0.0, Wq = 0.0;   // interfaces k-1, k, k+1, k+2
	// derived, cached across levels
	double exn_m = 0.0, exn_0 = 0.0;    // Exner pressure on levels k-1, k
	double ken_m = 0.0, ken_0 = 0.0;    // kinetic energy
	double xdn_m = 0.0, xdn_0 = 0.0;    // xi-dot on levels
	double xde_0 = 0.0, xde_p = 0.0;    // xi-dot on interfaces k, k+1
	double seP_0 = 0.0, seP_p = 0.0, seR_0 = 0.0, seR_p = 0.0;  // rho-theta, rho on interfaces k, k+1
	double cxe2_0 = 0.0, cxe2_p = 0.0;  // ContraMetricXi[2] on interfaces k, k+1

	// Rows of level k from the window (all names relative to k).  lv = window
	// table row of level k.
	auto assemble = [&](int k, TbcRows & o) {
		const double * lv = slev + (size_t)k * TBF_LW;
#pragma unroll
		for (int b = 0; b < 9; b++) { o.p[b] = 0.0; o.w[b] = 0.0; o.r[b] = 0.0; }
		double fP = 0.0, fW = 0.0, fR = 0.0;
		const double i0 = lv[TBF_CILO + 0], i1 = lv[TBF_CILO + 1];     // interface k   <- levels k-1, k
		const double h1 = lv[TBF_CIHI + 1], h2 = lv[TBF_CIHI + 2];     // interface k+1 <- levels k, k+1
		if (k < L) {
			// BuildF: conservative fluxes (:2219-2254, 2301-2333)
			const double de0 = lv[TBF_DEN + 0], de1 = lv[TBF_DEN + 1];
			const double mfe0 = (k >= 1) ? cJ * seR_0 * xde_0 : 0.0;
			const double pfe0 = (k >= 1) ? cJ * seP_0 * xde_0 : 0.0;
			const double mfe1 = (k + 1 < L) ? cJ * seR_p * xde_p : 0.0;
			const double pfe1 = (k + 1 < L) ? cJ * seP_p * xde_p : 0.0;
			double dmfn = 0.0, dpfn = 0.0;
			dmfn += de0 * mfe0; dpfn += de0 * pfe0;
			dmfn += de1 * mfe1; dpfn += de1 * pfe1;
			fR = dmfn * invj;
			fP += dpfn * invj;
			// upwind penalty of rho-theta and rho (:2687-2713)
			const double pl0 = lv[TBF_CPL + 0], pl1 = lv[TBF_CPL + 1], pl2 = lv[TBF_CPL + 2];
			const double pr0 = lv[TBF_CPR + 0], pr1 = lv[TBF_CPR + 1], pr2 = lv[TBF_CPR + 2];
			{
				double auxP = 0.0, auxR = 0.0;
				if (k <= L - 2) {
					double a = 0.0, b = 0.0;
					a += pl0 * Pm; a += pl1 * P0; a += pl2 * Pp;
					b += pl0 * Rm; b += pl1 * R0; b += pl2 * Rp;
					auxP += a * fabs(xde_p);
					auxR += b * fabs(xde_p);
				}
				if (k >= 1) {
					double a = 0.0, b = 0.0;
					a += pr0 * Pm; a += pr1 * P0; a += pr2 * Pp;
					b += pr0 * Rm; b += pr1 * R0; b += pr2 * Rp;
					auxP += a * fabs(xde_0);
					auxR += b * fabs(xde_0);
				}
				fP -= auxP;
				fR -= auxR;
			}
			// Jacobian of the flux terms (:3059-3091).  Row P: b = 1 P(k-1), 4 P(k),
			// 5 W(k), 7 P(k+1), 8 W(k+1); row R: b = 1 R(k-1), 3 W(k), 4 R(k),
			// 6 W(k+1), 7 R(k+1)
			{
				const double dm = de0;
				if (k != 0) {
					const double c = dm * cJ * invj * cxe2_0;
					o.p[5] += c * seP_0;
					o.r[3] += c * seR_0;
				}
				const double v0 = dm * cJ * invj * i0 * xde_0;
				const double v1 = dm * cJ * invj * i1 * xde_0;
				if (k >= 1) { o.r[1] += v0; o.p[1] += v0; }
				o.r[4] += v1; o.p[4] += v1;
			}
			{
				const double dm = de1;
				if (k + 1 != L) {
					const double c = dm * cJ * invj * cxe2_p;
					o.p[8] += c * seP_p;
					o.r[6] += c * seR_p;
				}
				const double v0 = dm * cJ * invj * h1 * xde_p;
				const double v1 = dm * cJ * invj * h2 * xde_p;
				o.r[4] += v0; o.p[4] += v0;
				if (k + 1 < L) { o.r[7] += v1; o.p[7] += v1; }
			}
		}
		if (k >= 1 && k < L) {
			// BuildF: vertical velocity on interfaces (:2533-2589)
			const double dn0 = lv[TBF_DNE + 0], dn1 = lv[TBF_DNE + 1];   // levels k-1, k
			double dPe = 0.0, dkee = 0.0;
			if (dn0 != 0.0) { dPe += dn0 * exn_m; dkee += dn0 * ken_m; }
			if (dn1 != 0.0) { dPe += dn1 * exn_0; dkee += dn1 * ken_0; }
			const double se = lv[TBF_SE];
			double seU = 0.0, seV = 0.0;
			seU += i0 * Um; seU += i1 * U0;
			seV += i0 * Vm; seV += i1 * V0;
			double f = dPe * seP_0 / seR_0;
			f += gdxr;
			const double ca2 = se * cA2, cb2 = se * cB2;
			const double dConUa = cA0 * seU + cA1 * seV + ca2 * W0;
			const double dConUb = cA1 * seU + cB1 * seV + cb2 * W0;
			double dUa = 0.0, dUb = 0.0;
			dUa += dn0 * Um; dUa += dn1 * U0;
			dUb += dn0 * Vm; dUb += dn1 * V0;
			const double dCurlTerm = -dConUa * dUa - dConUb * dUb;
			f += (dkee + dCurlTerm);
			fW = f;
			// Jacobian rows of w (:3094-3140).  Row W: b = 0 P(k-1), 1 W(k-1),
			// 2 R(k-1), 3 P(k), 4 W(k), 5 R(k), 7 W(k+1)
			const double dRHSWCoeffA = seP_0 * ph.R / (seR_0 * ph.cv);
			if (dn0 != 0.0) o.w[0] += dRHSWCoeffA * dn0 * exn_m / Pm;
			if (dn1 != 0.0) o.w[3] += dRHSWCoeffA * dn1 * exn_0 / P0;
			const double dRHSWCoeffB = 1.0 / (seR_0 * seR_0) * dPe;
			{
				const double c = dRHSWCoeffB * i0;
				o.w[0] += c * seR_0;
				o.w[2] += -c * seP_0;
			}
			{
				const double c = dRHSWCoeffB * i1;
				o.w[3] += c * seR_0;
				o.w[5] += -c * seP_0;
			}
			// Clark-form vertical advection of w
			if (dn0 != 0.0) {
				o.w[1] += lv[TBF_IEN1 + 0] * dn0 * xdn_m;
				o.w[4] += lv[TBF_IEN1 + 1] * dn0 * xdn_m;
			}
			if (dn1 != 0.0) {
				o.w[4] += lv[TBF_CW + 0] * dn1 * xdn_0;
				o.w[7] += lv[TBF_CW + 1] * dn1 * xdn_0;
			}
		}
		{
			// upwinding of w on interfaces: F (:2676-2686), Jacobian (:2870-2897)
			double d2 = 0.0;
			if (k > 0 && k < L) {
				d2 += lv[TBF_DDE + 0] * Wm; d2 += lv[TBF_DDE + 1] * W0; d2 += lv[TBF_DDE + 2] * Wp;
			}
			const double axd = fabs(xde_0);
			fW -= upw * axd * d2;
			o.w[4] -= upw * tb_signw(xde_0, cxe2_0) * d2;
			if (k >= 1) o.w[1] -= upw * axd * lv[TBF_DDE + 0];
			o.w[4] -= upw * axd * lv[TBF_DDE + 1];
			if (k < L) o.w[7] -= upw * axd * lv[TBF_DDE + 2];
		}
		if (k < L) {
			// upwind penalty of rho-theta and rho: Jacobian (:2899-2968); the
			// "right" terms (interface k) come before the "left" ones (k+1)
			const double pl0 = lv[TBF_CPL + 0], pl1 = lv[TBF_CPL + 1], pl2 = lv[TBF_CPL + 2];
			const double pr0 = lv[TBF_CPR + 0], pr1 = lv[TBF_CPR + 1], pr2 = lv[TBF_CPR + 2];
			if (k >= 1) {
				const double wgt = fabs(xde_0);
				const double sw = tb_signw(xde_0, cxe2_0);
				o.p[5] -= sw * pr0 * Pm; o.p[5] -= sw * pr1 * P0; o.p[5] -= sw * pr2 * Pp;
				o.p[1] -= wgt * pr0; o.p[4] -= wgt * pr1; o.p[7] -= wgt * pr2;
			}
			if (k <= L - 2) {
				const double wgt = fabs(xde_p);
				const double sw = tb_signw(xde_p, cxe2_p);
				o.p[8] -= sw * pl0 * Pm; o.p[8] -= sw * pl1 * P0; o.p[8] -= sw * pl2 * Pp;
				o.p[1] -= wgt * pl0; o.p[4] -= wgt * pl1; o.p[7] -= wgt * pl2;
			}
			if (k >= 1) {
				const double wgt = fabs(xde_0);
				const double sw = tb_signw(xde_0, cxe2_0);
				o.r[3] -= sw * pr0 * Rm; o.r[3] -= sw * pr1 * R0; o.r[3] -= sw * pr2 * Rp;
				o.r[1] -= wgt * pr0; o.r[4] -= wgt * pr1; o.r[7] -= wgt * pr2;
			}
			if (k <= L - 2) {
				const double wgt = fabs(xde_p);
				const double sw = tb_signw(xde_p, cxe2_p);
				o.r[6] -= sw * pl0 * Rm; o.r[6] -= sw * pl1 * R0; o.r[6] -= sw * pl2 * Rp;
				o.r[1] -= wgt * pl0; o.r[4] -= wgt * pl1; o.r[7] -= wgt * pl2;
			}
		}
		if (k == 0 || k == L) fW = 0.0;       // :2747-2758
		o.p[4] += dInvDeltaT;
		o.w[4] += dInvDeltaT;
		o.r[4] += dInvDeltaT;
		o.fp = fP; o.fw = fW; o.fr = fR;
	};

	// Advance the window so that it is centred on level k (called with
	// k = 0, 1, ..., L in order) and refresh the cached column quantities
	// (PrepareColumn, :1839-2179).
	auto advance = [&](int k) {
		const double * lv = slev + (size_t)k * TBF_LW;
		if (k == 0) {
			U0 = inU[0]; V0 = inV[0]; P0 = inP[0]; R0 = inR[0];
			W0 = inW[0]; Wp = inW[NN];
			if (L > 1) {
				Up = inU[NN]; Vp = inV[NN]; Pp = inP[NN]; Rp = inR[NN];
				Wq = inW[(size_t)2 * NN];
			}
			if (L > 2) {
				const size_t o2 = (size_t)2 * NN;
				Uq = inU[o2]; Vq = inV[o2]; Pq = inP[o2]; Rq = inR[o2];
			}
		} else {
			Um = U0; U0 = Up; Up = Uq;
			Vm = V0; V0 = Vp; Vp = Vq;
			Pm = P0; P0 = Pp; Pp = Pq;
			Rm = R0; R0 = Rp; Rp = Rq;
			Wm = W0; W0 = Wp; Wp = Wq;
			if (k + 2 < L) {
				const size_t o2 = (size_t)(k + 2) * NN;
				Uq = inU[o2]; Vq = inV[o2]; Pq = inP[o2]; Rq = inR[o2];
			}
			if (k + 2 <= L) Wq = inW[(size_t)(k + 2) * NN];
		}
		// interface k <- interface k+1 of the previous level
		xde_0 = xde_p; seP_0 = seP_p; seR_0 = seR_p; cxe2_0 = cxe2_p;
		if (k == 0) {
			const double se = lv[TBF_SE];
			cxe2_0 = cX0 + (se * se) * cX2;
			xde_0 = 0.0;
		}
		// interface k+1: interpolated rho-theta, rho, xi-dot (:1960-2086)
		if (k + 1 <= L) {
			const double se1 = lv[TBF_SE1];
			cxe2_p = cX0 + (se1 * se1) * cX2;
			if (k + 1 < L) {
				const double h1 = lv[TBF_CIHI + 1], h2 = lv[TBF_CIHI + 2];
				double sU = 0.0, sV = 0.0, sP = 0.0, sR = 0.0;
				sU += h1 * U0; sU += h2 * Up;
				sV += h1 * V0; sV += h2 * Vp;
				sP += h1 * P0; sP += h2 * Pp;
				sR += h1 * R0; sR += h2 * Rp;
				seP_p = sP; seR_p = sR;
				xde_p = (se1 * cA2) * sU + (se1 * cB2) * sV + cxe2_p * Wp;
			} else {
				seP_p = 0.0; seR_p = 0.0; xde_p = 0.0;
			}
		}
		// level k: w on the level, Exner pressure, xi-dot, kinetic energy
		exn_m = exn_0; ken_m = ken_0; xdn_m = xdn_0;
		if (k < L) {
			const double sn = lv[TBF_SN];
			double wn = 0.0;
			wn += lv[TBF_CW + 0] * W0; wn += lv[TBF_CW + 1] * Wp;
			exn_0 = tb_exner(ph, P0);
			const double cx0 = sn * cA2, cx1 = sn * cB2;
			const double cx2 = cX0 + (sn * sn) * cX2;
			const double dConUa = cA0 * U0 + cA1 * V0 + cx0 * wn;
			const double dConUb = cA1 * U0 + cB1 * V0 + cx1 * wn;
			const double dConUx = cx0 * U0 + cx1 * V0 + cx2 * wn;
			xdn_0 = dConUx;
			ken_0 = 0.5 * (dConUa * U0 + dConUb * V0 + dConUx * wn);
		}
	};

	// ---- dgbtf2 + forward substitution on the register block -----------------------
	// B[r][c] = A(j + r, j + c), bb[r] = b(j + r) with j = 3k fixed over the three
	// steps of level k: rows 0-2 are the (partly eliminated) rows of level k,
	// rows 3-5 those of level k+1; step s works on rows s..s+4 and columns
	// s..s+8 (row 6 of the third step is P(k+2), zero in that pivot column, and
	// is left out).  The names are static inside a level; the block moves by
	// three rows and columns once per level.
	double B[6][11];
	double bb[6];
#pragma unroll
	for (int r = 0; r < 6; r++) {
#pragma unroll
		for (int c = 0; c < 11; c++) B[r][c] = 0.0;
		bb[r] = 0.0;
	}
	int info = 0;
	int jstep = 0;

	// elimination step s of the level
	auto eliminate = [&](auto sc) {
		constexpr int s = decltype(sc)::value;
		constexpr int rlast = (s + 4 < 5) ? (s + 4) : 5;
		int jp = 0;
		double amax = fabs(B[s][s]);
#pragma unroll
		for (int r = s + 1; r <= rlast; r++) {
			const double v = fabs(B[r][s]);
			if (v > amax) { amax = v; jp = r - s; }
		}
		// interchange rows j and j + jp (warp-uniform in practice)
		if (jp != 0) {
#define TBC_SWAP(R) { _Pragma("unroll") for (int c = s; c < s + 9; c++) { const double t0 = B[s][c]; B[s][c] = B[R][c]; B[R][c] = t0; } \
			const double t1 = bb[s]; bb[s] = bb[R]; bb[R] = t1; }
			if (jp == 1) TBC_SWAP(s + 1)
			else if (jp == 2) TBC_SWAP(s + 2)
			else if (jp == 3) TBC_SWAP(s + 3)
			else TBC_SWAP((s + 4 <= 5) ? (s + 4) : 5)
#undef TBC_SWAP
		}
		if (B[s][s] != 0.0) {
			const double rcp = 1.0 / B[s][s];
			const double bj = bb[s];
#pragma unroll
			for (int r = s + 1; r <= rlast; r++) {
				const double mult = B[r][s] * rcp;
#pragma unroll
				for (int c = s + 1; c < s + 9; c++) {
					B[r][c] -= mult * B[s][c];
				}
				bb[r] -= mult * bj;
			}
		} else if (info == 0) {
			info = jstep + 1;
		}
		{
			unsigned m = 0;
#pragma unroll
			for (int c = 1; c < 9; c++) {
				if (__any_sync(FULL, B[s][s + c] != 0.0)) m |= (1u << c);
			}
			scp[0] = B[s][s];
			scp[S] = bb[s];
			int slot = 2;
#pragma unroll
			for (int c = 1; c < 9; c++) {
				if (m & (1u << c)) {
					scp[slot * S] = B[s][s + c];
					slot++;
				}
			}
			scp += slot * S;
			if ((threadIdx.x & 31) == 0) smask[jstep] = m;
		}
		jstep++;
	};

	{
		// level 0: rows 0, 1, 2 at j = 0, block column = r - 4 + b
		TbcRows o;
		advance(0);
		assemble(0, o);
#pragma unroll
		for (int b = 4; b < 9; b++) B[0][b - 4] = o.p[b];
#pragma unroll
		for (int b = 3; b < 9; b++) B[1][b - 3] = o.w[b];
#pragma unroll
		for (int b = 2; b < 9; b++) B[2][b - 2] = o.r[b];
		bb[0] = o.fp; bb[1] = o.fw; bb[2] = o.fr;
	}
	for (int k = 0; k <= L; k++) {
		if (k + 1 <= L) {
			TbcRows o;
			advance(k + 1);
			assemble(k + 1, o);
			// P(k+1) -> row 3 (block column b - 1), W(k+1) -> row 4 (column b),
			// R(k+1) -> row 5 (column b + 1); the rest of the rows is zero
#pragma unroll
			for (int b = 1; b < 9; b++) B[3][b - 1] = o.p[b];
#pragma unroll
			for (int b = 0; b < 9; b++) B[4][b] = o.w[b];
#pragma unroll
			for (int b = 0; b < 9; b++) B[5][b + 1] = o.r[b];
			bb[3] = o.fp; bb[4] = o.fw; bb[5] = o.fr;
		} else {
#pragma unroll
			for (int r = 3; r < 6; r++) {
#pragma unroll
				for (int c = 0; c < 11; c++) B[r][c] = 0.0;
				bb[r] = 0.0;
			}
		}
		B[3][8] = 0.0; B[3][9] = 0.0; B[3][10] = 0.0;
		B[4][9] = 0.0; B[4][10] = 0.0;
		B[5][0] = 0.0; B[5][10] = 0.0;
		eliminate(std::integral_constant<int, 0>());          // j = 3k
		eliminate(std::integral_constant<int, 1>());          // j = 3k + 1
		eliminate(std::integral_constant<int, 2>());          // j = 3k + 2
		// rows and columns 3.. become 0..
#pragma unroll
		for (int r = 0; r < 3; r++) {
#pragma unroll
			for (int c = 0; c < 8; c++) B[r][c] = B[r + 3][c + 3];
			B[r][8] = 0.0; B[r][9] = 0.0; B[r][10] = 0.0;
			bb[r] = bb[r + 3];
		}
	}
	__syncwarp();          // the row masks of this warp are complete
	if (info != 0) {
		if (live) atomicMax(ca.info, ca.col0 + tcol + 1);
		return;
	}
	if (!live) return;

	// ---- back substitution in dtbsv's order; x = x0 - delta scattered as each
	//      unknown appears ----------------------------------------------------------
	const int * dups = ca.col_dups + (size_t)(ca.col0 + tcol) * 3;
	const int d0 = dups[0], d1 = dups[1], d2 = dups[2];
	const size_t ob0 = (d0 >= 0) ? ((size_t)(d0 / NN) * lay.nrows * NN + (d0 % NN)) : 0;
	const size_t ob1 = (d1 >= 0) ? ((size_t)(d1 / NN) * lay.nrows * NN + (d1 % NN)) : 0;
	const size_t ob2 = (d2 >= 0) ? ((size_t)(d2 / NN) * lay.nrows * NN + (d2 % NN)) : 0;
	const int rowP = lay.rowoff[2], rowW = lay.rowoff[3], rowR = lay.rowoff[4];
	double xr[8];      // x(j+1) .. x(j+8)
#pragma unroll
	for (int c = 0; c < 8; c++) xr[c] = 0.0;
	bool nan_seen = false;
	// level k: steps 3k+2, 3k+1, 3k; the scratch rows and the old state of the
	// next level to be processed (k-1) are loaded while level k is computed
	double ucur[3][10], x0cur[3];
	// rows of level k (steps 3k+2, 3k+1, 3k) from the compacted scratch; scp
	// walks backwards
	auto load_level = [&](int k, double (&u)[3][10]) {
#pragma unroll
		for (int c3 = 2; c3 >= 0; c3--) {
			const unsigned m = smask[3 * k + c3];
			scp -= (2 + __popc(m)) * S;
			u[c3][0] = scp[0];
			u[c3][9] = scp[S];
			int slot = 2;
#pragma unroll
			for (int c = 1; c < 9; c++) {
				if (m & (1u << c)) {
					u[c3][c] = scp[slot * S];
					slot++;
				} else {
					u[c3][c] = 0.0;
				}
			}
		}
	};
	load_level(L, ucur);
	x0cur[0] = 0.0; x0cur[2] = 0.0;
	x0cur[1] = in[ebase + (size_t)(rowW + L) * NN + nd];
	for (int k = L; k >= 0; k--) {
		double unext[3][10], x0next[3];
		if (k > 0) {
			load_level(k - 1, unext);
			x0next[0] = in[ebase + (size_t)(rowP + k - 1) * NN + nd];
			x0next[1] = in[ebase + (size_t)(rowW + k - 1) * NN + nd];
			x0next[2] = in[ebase + (size_t)(rowR + k - 1) * NN + nd];
		} else {
#pragma unroll
			for (int c3 = 0; c3 < 3; c3++) {
#pragma unroll
				for (int c = 0; c < 10; c++) unext[c3][c] = 0.0;
				x0next[c3] = 0.0;
			}
		}
#pragma unroll
		for (int c3 = 2; c3 >= 0; c3--) {
			double acc = ucur[c3][9];
#pragma unroll
			for (int c = 8; c >= 1; c--) {
				acc -= xr[c - 1] * ucur[c3][c];
			}
			double xj = acc;
			if (xj != 0.0) xj = xj / ucur[c3][0];
#pragma unroll
			for (int c = 7; c >= 1; c--) xr[c] = xr[c - 1];
			xr[0] = xj;
			if (k == 0 && c3 == 0 && !(xj == xj)) nan_seen = true;
			if (c3 == 1 || k < L) {
				const int row = ((c3 == 0) ? rowP : ((c3 == 1) ? rowW : rowR)) + k;
				const size_t ro = (size_t)row * NN;
				const double xnew = x0cur[c3] - xj;
				out[ebase + ro + nd] = xnew;
				if (d0 >= 0) out[ob0 + ro] = xnew;
				if (d1 >= 0) out[ob1 + ro] = xnew;
				if (d2 >= 0) out[ob2 + ro] = xnew;
				if (ca.inc != 0) {
					// Grid::LinearCombineData({+1, -1}) of the new and the old state
					const double dx = xnew - x0cur[c3];
					ca.inc[ebase + ro + nd] = dx;
					if (d0 >= 0) ca.inc[ob0 + ro] = dx;
					if (d1 >= 0) ca.inc[ob1 + ro] = dx;
					if (d2 >= 0) ca.inc[ob2 + ro] = dx;
				}
			}
		}
#pragma unroll
		for (int c3 = 0; c3 < 3; c3++) {
#pragma unroll
			for (int c = 0; c < 10; c++) ucur[c3][c] = unext[c3][c];
			x0cur[c3] = x0next[c3];
		}
	}
	if (nan_seen) atomicMax(ca.info, ca.col0 + tcol + 1);
}


__global__ void __launch_bounds__(TBC_THREADS)
k_column_fast(
	DevLayout lay, DevPhys ph, ColumnFastArgs ca,
	const double * in, double * out,  // may alias
	int nbatches
) {
	TB_DYN_SMEM(double, slev);          // [L+1][TBF_LW], then row masks [warps][n]
	const int L = lay.nlev;
	const int n = 3 * (L + 1);
	unsigned * smask = reinterpret_cast<unsigned *>(slev + (size_t)(L + 1) * TBF_LW)
		+ (size_t)(threadIdx.x >> 5) * n;
	for (int q = threadIdx.x; q < (L + 1) * TBF_LW; q += blockDim.x) {
		slev[q] = ca.lev[q];
	}
	__syncthreads();
	// warps are independent from here on (no block barrier in a batch)
	for (int vb = blockIdx.x; vb < nbatches; vb += gridDim.x) {
		tb_column_fast_batch(lay, ph, ca, in, out, slev, smask, vb, blockIdx.x);
		__syncwarp();
	}
}

#endif
