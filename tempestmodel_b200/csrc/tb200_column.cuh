// Vertically implicit column solve (FP64): one thread per unique column.
//
// Restates VerticalDynamicsFEM::StepImplicit and its helpers for the
// configuration the reference compiles (src/atm/Defines.h: USE_DIRECTSOLVE,
// USE_JACOBIAN_DIAGONAL, FORMULATION_RHOTHETA_PI; VerticalDynamicsFEM.cpp:36-46
// upwinding on rho-theta, w, rho; Clark-form implicit vertical advection of w)
// under Lorenz staggering:
//   SetupReferenceColumn   VerticalDynamicsFEM.cpp:1643-1835
//   PrepareColumn          :1839-2179
//   BuildF                 :2183-2780
//   BuildJacobianF_LOR_RhoTheta_Pi + _Diffusion   :2977-3187, :2784-2973
//   LAPACK dgbsv (dgbtf2 + dgbtrs, partial pivoting)  LinearAlgebra.cpp:156-202
//   x = x0 - J^{-1} F and scatter to duplicates   :1483-1633
//
// All per-column work arrays live in a global workspace laid out
// [entry][column] so that the threads of a warp (adjacent columns) touch
// consecutive addresses.
#ifndef TB200_COLUMN_CUH
#define TB200_COLUMN_CUH

#include "tb200_platform.h"
#include "tb200_device.h"
#include "tb200_kernels.cuh"

struct ColumnArgs {
	const int * col_node;   // [ncols] local node address (e*NN+n) solved
	const int * col_dups;   // [ncols][3] duplicates receiving a copy, -1 unused
	int ncols;              // columns in this launch
	int col0;               // first column of this launch
	double * ws;            // workspace [entries][ws_stride]
	int ws_stride;
	double dt;
	int offd;               // m_nJacobianFOffD (kl = ku)
	int fe_nodes;           // nodes per vertical finite element
	double upwind_coeff;    // m_dUpwindCoeff
	int * info;             // device flag: first failing column + 1
	int assemble_only;      // debugging: 1 stop before the solve, 2 after it;
	                        // 3: fully explicit vertical step (update -= dt F)
	// uniform diffusion (BuildF :2594-2636; k_column_implicit only): reference state in
	// the state layout and the coefficients divided by ztop^2; 0 = off
	const double * ref = 0;
	double uni_s = 0.0;
	double uni_v = 0.0;
	// --vmassfluxlevels (BuildF :2229-2243, 2301-2315; k_column_implicit only)
	int mass_flux_levels = 0;
};

// number of workspace entries per column
__host__ __device__ inline int tb_column_ws_entries(int L, int offd) {
	const int n = 3 * (L + 1);
	return 24 * (L + 1) + 2 * n + n * (3 * offd + 1);
}

// Banded LU with partial pivoting and solve, nrhs = 1: LAPACK dgbsv =
// dgbtf2 + dgbtrs('N').  ab(r, j) is band row r (0-based, 0..ldab-1) of
// column j: the reference's row-major [n][ldab] array handed to Fortran as
// AB(ldab, n) (LinearAlgebra.cpp:156-202).  Strided accessors.
template <typename AB, typename BV>
__device__ inline int tb_dgbsv(int n, int kl, int ku, AB ab, BV b) {
	const int kv = ku + kl;
	int info = 0;
	// zero the fill-in part of the first superdiagonal columns
	for (int j = ku + 1; j < ((kv < n) ? kv : n); j++) {
		for (int i = kv - j; i < kl; i++) {
			ab(i, j) = 0.0;
		}
	}
	int ju = 0;
	for (int j = 0; j < n; j++) {
		if (j + kv < n) {
			for (int i = 0; i < kl; i++) {
				ab(i, j + kv) = 0.0;
			}
		}
		const int km = (kl < n - 1 - j) ? kl : (n - 1 - j);
		// idamax
		int jp = 0;
		double amax = fabs(ab(kv, j));
		for (int i = 1; i <= km; i++) {
			const double v = fabs(ab(kv + i, j));
			if (v > amax) {
				amax = v;
				jp = i;
			}
		}
		const int piv = jp + j;
		if (ab(kv + jp, j) != 0.0) {
			int cand = j + ku + jp;
			if (cand > n - 1) cand = n - 1;
			if (cand > ju) ju = cand;
			if (jp != 0) {
				for (int c = 0; c <= ju - j; c++) {
					const double tmp = ab(kv + jp - c, j + c);
					ab(kv + jp - c, j + c) = ab(kv - c, j + c);
					ab(kv - c, j + c) = tmp;
				}
			}
			if (km > 0) {
				const double r = 1.0 / ab(kv, j);
				for (int i = 1; i <= km; i++) {
					ab(kv + i, j) *= r;
				}
				for (int c = 1; c <= ju - j; c++) {
					const double y = ab(kv - c, j + c);
					if (y != 0.0) {
						for (int i = 1; i <= km; i++) {
							ab(kv + i - c, j + c) -= ab(kv + i, j) * y;
						}
					}
				}
			}
		} else if (info == 0) {
			info = j + 1;
		}
		// forward substitution of dgbtrs fused in (same arithmetic)
		if (j < n - 1) {
			if (piv != j) {
				const double tmp = b(piv);
				b(piv) = b(j);
				b(j) = tmp;
			}
			const double bj = b(j);
			for (int i = 1; i <= km; i++) {
				b(j + i) -= ab(kv + i, j) * bj;
			}
		}
	}
	if (info != 0) return info;
	// dtbsv upper, no transpose, non-unit, bandwidth kv
	for (int j = n - 1; j >= 0; j--) {
		if (b(j) != 0.0) {
			b(j) = b(j) / ab(kv, j);
			const double temp = b(j);
			const int lo = (j - kv > 0) ? (j - kv) : 0;
			for (int i = j - 1; i >= lo; i--) {
				b(i) -= temp * ab(kv - (j - i), j);
			}
		}
	}
	return 0;
}

struct WsAcc {
	double * p;
	int stride;
	__device__ __forceinline__ double & operator()(int i) const {
		return p[(size_t)i * stride];
	}
};

struct WsBand {
	double * p;
	int stride;
	int ldab;
	__device__ __forceinline__ double & operator()(int r, int j) const {
		return p[(size_t)(j * ldab + r) * stride];
	}
};

// Standalone batched band solve (pinning against LAPACK)
__global__ void k_band_solve(
	int ncols, int n, int kl, int ku, double * ab, double * b, int * info
) {
	const int tcol = blockIdx.x * blockDim.x + threadIdx.x;
	if (tcol >= ncols) return;
	const int ldab = 2 * kl + ku + 1;
	WsBand A = {ab + tcol, ncols, ldab};
	WsAcc B = {b + tcol, ncols};
	const int r = tb_dgbsv(n, kl, ku, A, B);
	if (r != 0) atomicMax(info, r);
}

// apply a column operator to a workspace column
__device__ __forceinline__ double tb_ws_apply(const DevOp & op, const WsAcc & in, int k) {
	double o = 0.0;
	const int b = op.begin[k];
	const int e = op.end[k];
	const double * c = op.coeff + (size_t)k * op.width;
	for (int l = b; l < e; l++) {
		o += c[l - b] * in(l);
	}
	return o;
}

__device__ __forceinline__ double tb_op_coeff(const DevOp & op, int k, int l) {
	return op.coeff[(size_t)k * op.width + (l - op.begin[k])];
}

__global__ void k_column_implicit(
	DevLayout lay, DevGeom g, DevOps ops, DevPhys ph, ColumnArgs ca,
	const double * in, double * out   // may alias: StepImplicit(i, i, ...)
) {
	const int tcol = blockIdx.x * blockDim.x + threadIdx.x;
	if (tcol >= ca.ncols) return;

	const int UIx = 0, VIx = 1, PIx = 2, WIx = 3, RIx = 4;
	const int FP = 0, FW = 1, FR = 2;
	const int NN = lay.nn;
	const int L = lay.nlev;
	const int n = 3 * (L + 1);
	const int offd = ca.offd;
	const int ldab = 3 * offd + 1;

	// (col_node == 0: every element-local node is its own column - the fully
	// explicit vertical step visits all of them, VerticalDynamicsFEM.cpp:723-724)
	const int node = (ca.col_node != 0) ? ca.col_node[ca.col0 + tcol] : (ca.col0 + tcol);
	const long long e = node / NN;
	const int nd = node % NN;
	const size_t ebase = (size_t)e * lay.nrows * NN;
	const size_t g3 = (size_t)e * L * NN + nd;
	const size_t g3e = (size_t)e * (L + 1) * NN + nd;

	// workspace carve-up
	double * w0 = ca.ws + tcol;
	const int S = ca.ws_stride;
	int cur = 0;
#define TB_WS(name, len) WsAcc name = {w0 + (size_t)cur * S, S}; cur += (len)
	TB_WS(snU, L + 1); TB_WS(snV, L + 1); TB_WS(snP, L + 1); TB_WS(snW, L + 1); TB_WS(snR, L + 1);
	TB_WS(seU, L + 1); TB_WS(seV, L + 1); TB_WS(seW, L + 1); TB_WS(seR, L + 1); TB_WS(seP, L + 1);
	TB_WS(exn, L + 1); TB_WS(dPe, L + 1); TB_WS(xdn, L + 1); TB_WS(xde, L + 1);
	TB_WS(mfe, L + 1); TB_WS(dmfn, L + 1); TB_WS(pfe, L + 1); TB_WS(dpfn, L + 1);
	TB_WS(ken, L + 1); TB_WS(dkee, L + 1); TB_WS(dUa, L + 1); TB_WS(dUb, L + 1);
	TB_WS(ddW, L + 1); TB_WS(aux, L + 1);
	TB_WS(x0, n); TB_WS(F, n);
#undef TB_WS
	WsBand DG = {w0 + (size_t)cur * S, S, ldab};

	const DevOp & opInterpN2E = ops.op[0];
	const DevOp & opInterpE2N = ops.op[1];
	const DevOp & opDiffN2E = ops.op[3];
	const DevOp & opDiffE2N = ops.op[4];
	const DevOp & opDDE2E = ops.op[7];
	const DevOp & opPenL = ops.op[8];
	const DevOp & opPenR = ops.op[9];

	const double * inU = in + ebase + (size_t)lay.rowoff[UIx] * NN + nd;
	const double * inV = in + ebase + (size_t)lay.rowoff[VIx] * NN + nd;
	const double * inP = in + ebase + (size_t)lay.rowoff[PIx] * NN + nd;
	const double * inW = in + ebase + (size_t)lay.rowoff[WIx] * NN + nd;
	const double * inR = in + ebase + (size_t)lay.rowoff[RIx] * NN + nd;

	// ---- SetupReferenceColumn (:1643-1835) --------------------------------
	for (int k = 0; k < L; k++) {
		snU(k) = inU[(size_t)k * NN];
		snV(k) = inV[(size_t)k * NN];
	}
	for (int k = 0; k <= L; k++) {
		seU(k) = tb_ws_apply(opInterpN2E, snU, k);
		seV(k) = tb_ws_apply(opInterpN2E, snV, k);
		dUa(k) = tb_ws_apply(opDiffN2E, snU, k);
		dUb(k) = tb_ws_apply(opDiffN2E, snV, k);
	}
	for (int q = 0; q < n; q++) {
		x0(q) = 0.0;
	}
	for (int k = 0; k < L; k++) {
		x0(3 * k + FP) = inP[(size_t)k * NN];
		x0(3 * k + FR) = inR[(size_t)k * NN];
	}
	for (int k = 0; k <= L; k++) {
		x0(3 * k + FW) = inW[(size_t)k * NN];
	}

	// ---- PrepareColumn (:1839-2179) ----------------------------------------
	for (int k = 0; k < L; k++) {
		snP(k) = x0(3 * k + FP);
		seW(k) = x0(3 * k + FW);
		snR(k) = x0(3 * k + FR);
	}
	seW(L) = x0(3 * L + FW);
	for (int k = 0; k < L; k++) {
		snW(k) = tb_ws_apply(opInterpE2N, seW, k);
	}
	for (int k = 0; k <= L; k++) {
		seR(k) = tb_ws_apply(opInterpN2E, snR, k);
		seP(k) = tb_ws_apply(opInterpN2E, snP, k);
	}
	for (int k = 0; k < L; k++) {
		exn(k) = ph.cp * exp(ph.exner_c1 * log(ph.exner_c2 * snP(k)));
	}
	for (int k = 0; k <= L; k++) {
		dPe(k) = tb_ws_apply(opDiffN2E, exn, k);
	}
	for (int k = 0; k < L; k++) {
		const size_t o = g3 + (size_t)k * NN;
		xdn(k) = g.cx[0][o] * snU(k) + g.cx[1][o] * snV(k) + g.cx[2][o] * snW(k);
	}
	for (int k = 1; k < L; k++) {
		const size_t o = g3e + (size_t)k * NN;
		xde(k) = g.cxe[0][o] * seU(k) + g.cxe[1][o] * seV(k) + g.cxe[2][o] * seW(k);
	}
	xde(0) = 0.0;
	xde(L) = 0.0;
	// second derivative of w for interface upwinding (:2091-2102)
	for (int k = 0; k <= L; k++) {
		ddW(k) = tb_ws_apply(opDDE2E, seW, k);
	}

	// ---- BuildF (:2183-2780) -------------------------------------------------
	for (int q = 0; q < n; q++) {
		F(q) = 0.0;
	}
	mfe(0) = 0.0; mfe(L) = 0.0; pfe(0) = 0.0; pfe(L) = 0.0;
	if (!ca.mass_flux_levels) {
		for (int k = 1; k < L; k++) {
			const double je = g.jace[g3e + (size_t)k * NN];
			mfe(k) = je * seR(k) * xde(k);
			pfe(k) = je * seP(k) * xde(k);
		}
	} else {
		// fluxes on levels, differentiated with zero boundary fluxes
		for (int k = 0; k < L; k++) {
			const double jn = g.jac[g3 + (size_t)k * NN];
			mfe(k) = jn * snR(k) * xdn(k);
			pfe(k) = jn * snP(k) * xdn(k);
		}
	}
	{
		const DevOp & opFlux = ca.mass_flux_levels ? ops.op[10] : opDiffE2N;
		for (int k = 0; k < L; k++) {
			const double invj = 1.0 / g.jac[g3 + (size_t)k * NN];
			dmfn(k) = tb_ws_apply(opFlux, mfe, k);
			dpfn(k) = tb_ws_apply(opFlux, pfe, k);
			F(3 * k + FR) = dmfn(k) * invj;
			F(3 * k + FP) += dpfn(k) * invj;
		}
	}
	// kinetic energy on levels (:2433-2467)
	for (int k = 0; k < L; k++) {
		const size_t o = g3 + (size_t)k * NN;
		const double dCovUa = snU(k), dCovUb = snV(k), dCovUx = snW(k);
		const double dConUa = g.ca[0][o] * dCovUa + g.ca[1][o] * dCovUb + g.ca[2][o] * dCovUx;
		const double dConUb = g.cb[0][o] * dCovUa + g.cb[1][o] * dCovUb + g.cb[2][o] * dCovUx;
		const double dConUx = g.cx[0][o] * dCovUa + g.cx[1][o] * dCovUb + g.cx[2][o] * dCovUx;
		ken(k) = 0.5 * (dConUa * dCovUa + dConUb * dCovUb + dConUx * dCovUx);
	}
	for (int k = 0; k <= L; k++) {
		dkee(k) = tb_ws_apply(opDiffN2E, ken, k);
	}
	// vertical velocity on interfaces (:2533-2589)
	for (int k = 1; k < L; k++) {
		const size_t o = g3e + (size_t)k * NN;
		const double dPressureGradientForce = dPe(k) * seP(k) / seR(k);
		double f = dPressureGradientForce;
		f += ph.g * g.dre[2][o];
		const double dCovUa = seU(k), dCovUb = seV(k), dCovUx = seW(k);
		const double dConUa = g.cae[0][o] * dCovUa + g.cae[1][o] * dCovUb + g.cae[2][o] * dCovUx;
		const double dConUb = g.cbe[0][o] * dCovUa + g.cbe[1][o] * dCovUb + g.cbe[2][o] * dCovUx;
		const double dCurlTerm = -dConUa * dUa(k) - dConUb * dUb(k);
		f += (dkee(k) + dCurlTerm);
		F(3 * k + FW) = f;
	}
	// uniform diffusion of rho theta and w minus their reference (:2101-2160,
	// 2594-2636; the Jacobian does not carry it)
	if (ca.ref != 0) {
		const DevOp & opDDN2N = ops.op[6];
		const double * rP = ca.ref + ebase + (size_t)lay.rowoff[PIx] * NN + nd;
		const double * rW = ca.ref + ebase + (size_t)lay.rowoff[WIx] * NN + nd;
		for (int k = 0; k < L; k++) {
			aux(k) = rP[(size_t)k * NN];
		}
		for (int k = 0; k < L; k++) {
			const double dd = tb_ws_apply(opDDN2N, snP, k);
			const double dUniform = dd - tb_ws_apply(opDDN2N, aux, k);
			F(3 * k + FP) -= ca.uni_s * dUniform;
		}
		for (int k = 0; k <= L; k++) {
			aux(k) = rW[(size_t)k * NN];
		}
		for (int k = 1; k < L; k++) {
			const double dUniform = ddW(k) - tb_ws_apply(opDDE2E, aux, k);
			F(3 * k + FW) -= ca.uni_v * dUniform;
		}
	}
	// vertical upwinding (:2640-2713)
	const int vo = ca.fe_nodes;
	const int nfe = L / vo;
	for (int c = 2; c < 5; c++) {
		if (c == WIx) {
			ddW(0) = 0.0;
			ddW(L) = 0.0;
			for (int k = 0; k <= L; k++) {
				F(3 * k + FW) -= ca.upwind_coeff * fabs(xde(k)) * ddW(k);
			}
		} else {
			const WsAcc & sn = (c == PIx) ? snP : snR;
			const int fc = (c == PIx) ? FP : FR;
			for (int k = 0; k < L; k++) {
				aux(k) = 0.0;
			}
			for (int a = 0; a < nfe - 1; a++) {
				const double wgt = fabs(xde((a + 1) * vo));
				for (int ii = 0; ii < vo; ii++) {
					const int k = a * vo + ii;
					aux(k) += tb_ws_apply(opPenL, sn, k) * wgt;
				}
			}
			for (int a = 1; a < nfe; a++) {
				const double wgt = fabs(xde(a * vo));
				for (int ii = 0; ii < vo; ii++) {
					const int k = a * vo + ii;
					aux(k) += tb_ws_apply(opPenR, sn, k) * wgt;
				}
			}
			for (int k = 0; k < L; k++) {
				F(3 * k + fc) -= aux(k);
			}
		}
	}
	F(3 * 0 + FW) = 0.0;
	F(3 * L + FW) = 0.0;

	// ---- BuildJacobianF_LOR_RhoTheta_Pi (:2977-3187) -------------------------
	for (int j = 0; j < n; j++) {
		for (int r = 0; r < ldab; r++) {
			DG(r, j) = 0.0;
		}
	}
	// MatFIx(c0,k0,c1,k1): column 3*k0+c0, row 3*k1+c1 (VerticalDynamicsFEM.h:110-119)
#define TB_MAT(c0, k0, c1, k1) \
	DG(2 * offd + (3 * (k1) + (c1)) - (3 * (k0) + (c0)), 3 * (k0) + (c0))

	const double dInvDeltaT = 1.0 / ca.dt;

	for (int k = 0; k < L; k++) {
		const double invj = 1.0 / g.jac[g3 + (size_t)k * NN];
		for (int m = opDiffE2N.begin[k]; m < opDiffE2N.end[k]; m++) {
			const double je = g.jace[g3e + (size_t)m * NN];
			const double dm = tb_op_coeff(opDiffE2N, k, m);
			if ((m != 0) && (m != L)) {
				const double dMassFluxCoeff =
					dm * je * invj * g.cxe[2][g3e + (size_t)m * NN];
				TB_MAT(FW, m, FP, k) += dMassFluxCoeff * seP(m);
				TB_MAT(FW, m, FR, k) += dMassFluxCoeff * seR(m);
			}
			for (int q = opInterpN2E.begin[m]; q < opInterpN2E.end[m]; q++) {
				const double dCoeffVerticalFlux =
					dm * je * invj * tb_op_coeff(opInterpN2E, m, q) * xde(m);
				TB_MAT(FR, q, FR, k) += dCoeffVerticalFlux;
				TB_MAT(FP, q, FP, k) += dCoeffVerticalFlux;
			}
		}
	}
	for (int k = 1; k < L; k++) {
		const double dRHSWCoeffA = seP(k) * ph.R / (seR(k) * ph.cv);
		for (int m = opDiffN2E.begin[k]; m < opDiffN2E.end[k]; m++) {
			TB_MAT(FP, m, FW, k) +=
				dRHSWCoeffA * tb_op_coeff(opDiffN2E, k, m) * exn(m) / snP(m);
		}
		const double dRHSWCoeffB = 1.0 / (seR(k) * seR(k)) * dPe(k);
		for (int q = opInterpN2E.begin[k]; q < opInterpN2E.end[k]; q++) {
			const double dRHSWCoeffC = dRHSWCoeffB * tb_op_coeff(opInterpN2E, k, q);
			TB_MAT(FP, q, FW, k) += dRHSWCoeffC * seR(k);
			TB_MAT(FR, q, FW, k) += -dRHSWCoeffC * seP(k);
		}
	}
	// dW_k/dW_m (Clark form)
	for (int k = 1; k < L; k++) {
		for (int l = opDiffN2E.begin[k]; l < opDiffN2E.end[k]; l++) {
			for (int m = opInterpE2N.begin[l]; m < opInterpE2N.end[l]; m++) {
				TB_MAT(FW, m, FW, k) +=
					tb_op_coeff(opInterpE2N, l, m) * tb_op_coeff(opDiffN2E, k, l) * xdn(l);
			}
		}
	}
	// ---- BuildJacobianF_Diffusion (:2784-2973) -------------------------------
	for (int c = 2; c < 5; c++) {
		if (c == WIx) {
			for (int k = 0; k <= L; k++) {
				double dSignWeight;
				const double cx2 = g.cxe[2][g3e + (size_t)k * NN];
				if (xde(k) > 0.0) {
					dSignWeight = 1.0 * cx2;
				} else if (xde(k) < 0.0) {
					dSignWeight = -1.0 * cx2;
				} else {
					dSignWeight = 0.0;
				}
				TB_MAT(FW, k, FW, k) -= ca.upwind_coeff * dSignWeight * ddW(k);
			}
			for (int k = 0; k <= L; k++) {
				for (int q = opDDE2E.begin[k]; q < opDDE2E.end[k]; q++) {
					TB_MAT(FW, q, FW, k) -=
						ca.upwind_coeff * fabs(xde(k)) * tb_op_coeff(opDDE2E, k, q);
				}
			}
		} else {
			const WsAcc & sn = (c == PIx) ? snP : snR;
			const int fc = (c == PIx) ? FP : FR;
			for (int a = 1; a < nfe; a++) {
				const int ke = a * vo;
				const double xd = xde(ke);
				const double dWeight = fabs(xd);
				const double cx2 = g.cxe[2][g3e + (size_t)ke * NN];
				double dSignWeight;
				if (xd > 0.0) {
					dSignWeight = 1.0 * cx2;
				} else if (xd < 0.0) {
					dSignWeight = -1.0 * cx2;
				} else {
					dSignWeight = 0.0;
				}
				const int kLeftBegin = (a - 1) * vo;
				const int kLeftEnd = a * vo;
				const int kRightBegin = a * vo;
				const int kRightEnd = (a + 1) * vo;
				for (int k = kLeftBegin; k < kLeftEnd; k++) {
					for (int q = opPenL.begin[k]; q < opPenL.end[k]; q++) {
						TB_MAT(FW, kLeftEnd, fc, k) -=
							dSignWeight * tb_op_coeff(opPenL, k, q) * sn(q);
					}
				}
				for (int k = kRightBegin; k < kRightEnd; k++) {
					for (int q = opPenR.begin[k]; q < opPenR.end[k]; q++) {
						TB_MAT(FW, kRightBegin, fc, k) -=
							dSignWeight * tb_op_coeff(opPenR, k, q) * sn(q);
					}
				}
				for (int k = kLeftBegin; k < kLeftEnd; k++) {
					for (int q = opPenL.begin[k]; q < opPenL.end[k]; q++) {
						TB_MAT(fc, q, fc, k) -= dWeight * tb_op_coeff(opPenL, k, q);
					}
				}
				for (int k = kRightBegin; k < kRightEnd; k++) {
					for (int q = opPenR.begin[k]; q < opPenR.end[k]; q++) {
						TB_MAT(fc, q, fc, k) -= dWeight * tb_op_coeff(opPenR, k, q);
					}
				}
			}
		}
	}
	// identity components (:3172-3176)
	for (int k = 0; k <= L; k++) {
		TB_MAT(FP, k, FP, k) += dInvDeltaT;
		TB_MAT(FW, k, FW, k) += dInvDeltaT;
		TB_MAT(FR, k, FR, k) += dInvDeltaT;
	}
#undef TB_MAT

	if (ca.assemble_only == 1) return;

	if (ca.assemble_only == 3) {
		// VerticalDynamicsFEM::StepExplicit with m_fFullyExplicit (:748-793):
		// Evaluate(column state) = BuildF, update -= dt * F on rho theta, w, rho
		double * oP = out + ebase + nd + (size_t)lay.rowoff[PIx] * NN;
		double * oW = out + ebase + nd + (size_t)lay.rowoff[WIx] * NN;
		double * oR = out + ebase + nd + (size_t)lay.rowoff[RIx] * NN;
		for (int k = 0; k < L; k++) {
			oP[(size_t)k * NN] -= ca.dt * F(3 * k + FP);
			oR[(size_t)k * NN] -= ca.dt * F(3 * k + FR);
		}
		for (int k = 0; k <= L; k++) {
			oW[(size_t)k * NN] -= ca.dt * F(3 * k + FW);
		}
		return;
	}

	// ---- direct solve and update (:1457-1536) -------------------------------
	const int r = tb_dgbsv(n, offd, offd, DG, F);
	if (r != 0 || !(F(0) == F(0))) {
		atomicMax(ca.info, ca.col0 + tcol + 1);
	}

	if (ca.assemble_only == 2) return;

	const int * dups = ca.col_dups + (size_t)(ca.col0 + tcol) * 3;
	for (int q = -1; q < 3; q++) {
		int tgt = node;
		if (q >= 0) {
			tgt = dups[q];
			if (tgt < 0) continue;
		}
		const long long te = tgt / NN;
		const int tn = tgt % NN;
		const size_t tb = (size_t)te * lay.nrows * NN + tn;
		double * oP = out + tb + (size_t)lay.rowoff[PIx] * NN;
		double * oW = out + tb + (size_t)lay.rowoff[WIx] * NN;
		double * oR = out + tb + (size_t)lay.rowoff[RIx] * NN;
		for (int k = 0; k < L; k++) {
			oP[(size_t)k * NN] = x0(3 * k + FP) - F(3 * k + FP);
			oR[(size_t)k * NN] = x0(3 * k + FR) - F(3 * k + FR);
		}
		for (int k = 0; k <= L; k++) {
			oW[(size_t)k * NN] = x0(3 * k + FW) - F(3 * k + FW);
		}
	}
}


///////////////////////////////////////////////////////////////////////////////
// One warp per column, all work arrays in shared memory.
//
// Same arithmetic as k_column_implicit (which stays as the reference
// implementation of the solve on the device and as the fallback for columns
// too tall for shared memory); here lane k assembles the rows of level k, and
// the band LU runs warp-wide: pivot search by shuffle reduction, the rank-1
// update of the (kl x kl+ku) trailing block one entry per lane.  Every
// Jacobian entry is accumulated by a single lane in the reference's statement
// order, so the matrix is identical to the one the thread-per-column kernel
// (and the reference) build.

__host__ __device__ inline int tb_column_warp_smem_doubles(int L, int offd) {
	const int n = 3 * (L + 1);
	const int lp = L + 2;
	return 21 * lp + n + n * (3 * offd + 1) + 8;
}

struct SmAcc {
	double * p;
	__device__ __forceinline__ double & operator()(int i) const { return p[i]; }
};

struct SmBand {
	double * p;
	int ldab;
	__device__ __forceinline__ double & operator()(int r, int j) const {
		return p[j * ldab + r];
	}
};

__device__ __forceinline__ double tb_sm_apply(const DevOp & op, const SmAcc & in, int k) {
	double o = 0.0;
	const int b = op.begin[k];
	const int e = op.end[k];
	const double * c = op.coeff + (size_t)k * op.width;
	for (int l = b; l < e; l++) {
		o += c[l - b] * in(l);
	}
	return o;
}

__global__ void k_column_implicit_warp(
	DevLayout lay, DevGeom g, DevOps ops, DevPhys ph, ColumnArgs ca,
	const double * in, double * out, int smem_per_warp
) {
	TB_DYN_SMEM(double, smem_all);
	const int warp = threadIdx.x >> 5;
	const int lane = threadIdx.x & 31;
	const int wpb = blockDim.x >> 5;
	const int tcol = blockIdx.x * wpb + warp;
	if (tcol >= ca.ncols) return;      // whole warp leaves together
	const unsigned FULL = 0xffffffffu;

	const int UIx = 0, VIx = 1, PIx = 2, WIx = 3, RIx = 4;
	const int FP = 0, FW = 1, FR = 2;
	const int NN = lay.nn;
	const int L = lay.nlev;
	const int n = 3 * (L + 1);
	const int offd = ca.offd;
	const int ldab = 3 * offd + 1;
	const int lp = L + 2;

	const int node = ca.col_node[ca.col0 + tcol];
	const long long e = node / NN;
	const int nd = node % NN;
	const size_t ebase = (size_t)e * lay.nrows * NN;
	const size_t g3 = (size_t)e * L * NN + nd;
	const size_t g3e = (size_t)e * (L + 1) * NN + nd;

	double * w0 = smem_all + (size_t)warp * smem_per_warp;
	int cur = 0;
#define TB_SM(name) SmAcc name = {w0 + cur}; cur += lp
	TB_SM(snU); TB_SM(snV); TB_SM(snP); TB_SM(snW); TB_SM(snR);
	TB_SM(seU); TB_SM(seV); TB_SM(seW); TB_SM(seR); TB_SM(seP);
	TB_SM(exn); TB_SM(dPe); TB_SM(xdn); TB_SM(xde); TB_SM(mfe); TB_SM(pfe);
	TB_SM(ken); TB_SM(dkee); TB_SM(dUa); TB_SM(dUb); TB_SM(ddW);
#undef TB_SM
	SmAcc F = {w0 + cur}; cur += n;
	SmBand DG = {w0 + cur, ldab};

	const DevOp & opInterpN2E = ops.op[0];
	const DevOp & opInterpE2N = ops.op[1];
	const DevOp & opDiffN2E = ops.op[3];
	const DevOp & opDiffE2N = ops.op[4];
	const DevOp & opDDE2E = ops.op[7];
	const DevOp & opPenL = ops.op[8];
	const DevOp & opPenR = ops.op[9];

	const double * inU = in + ebase + (size_t)lay.rowoff[UIx] * NN + nd;
	const double * inV = in + ebase + (size_t)lay.rowoff[VIx] * NN + nd;
	const double * inP = in + ebase + (size_t)lay.rowoff[PIx] * NN + nd;
	const double * inW = in + ebase + (size_t)lay.rowoff[WIx] * NN + nd;
	const double * inR = in + ebase + (size_t)lay.rowoff[RIx] * NN + nd;

	// ---- SetupReferenceColumn / PrepareColumn ---------------------------------
	for (int k = lane; k <= L; k += 32) {
		if (k < L) {
			snU(k) = inU[(size_t)k * NN];
			snV(k) = inV[(size_t)k * NN];
			snP(k) = inP[(size_t)k * NN];
			snR(k) = inR[(size_t)k * NN];
		}
		seW(k) = inW[(size_t)k * NN];
	}
	__syncwarp();
	for (int k = lane; k <= L; k += 32) {
		seU(k) = tb_sm_apply(opInterpN2E, snU, k);
		seV(k) = tb_sm_apply(opInterpN2E, snV, k);
		dUa(k) = tb_sm_apply(opDiffN2E, snU, k);
		dUb(k) = tb_sm_apply(opDiffN2E, snV, k);
		seR(k) = tb_sm_apply(opInterpN2E, snR, k);
		seP(k) = tb_sm_apply(opInterpN2E, snP, k);
		double d2 = tb_sm_apply(opDDE2E, seW, k);
		if (k == 0 || k == L) d2 = 0.0;      // BuildF :2676-2679
		ddW(k) = d2;
		if (k < L) {
			snW(k) = tb_sm_apply(opInterpE2N, seW, k);
			exn(k) = ph.cp * exp(ph.exner_c1 * log(ph.exner_c2 * snP(k)));
		}
	}
	__syncwarp();
	for (int k = lane; k <= L; k += 32) {
		dPe(k) = tb_sm_apply(opDiffN2E, exn, k);
		if (k < L) {
			const size_t o = g3 + (size_t)k * NN;
			const double dCovUa = snU(k), dCovUb = snV(k), dCovUx = snW(k);
			const double cx0 = g.cx[0][o], cx1 = g.cx[1][o], cx2 = g.cx[2][o];
			xdn(k) = cx0 * dCovUa + cx1 * dCovUb + cx2 * dCovUx;
			const double dConUa = g.ca[0][o] * dCovUa + g.ca[1][o] * dCovUb + g.ca[2][o] * dCovUx;
			const double dConUb = g.cb[0][o] * dCovUa + g.cb[1][o] * dCovUb + g.cb[2][o] * dCovUx;
			const double dConUx = cx0 * dCovUa + cx1 * dCovUb + cx2 * dCovUx;
			ken(k) = 0.5 * (dConUa * dCovUa + dConUb * dCovUb + dConUx * dCovUx);
		}
		double xd = 0.0;
		if (k >= 1 && k < L) {
			const size_t o = g3e + (size_t)k * NN;
			xd = g.cxe[0][o] * seU(k) + g.cxe[1][o] * seV(k) + g.cxe[2][o] * seW(k);
		}
		xde(k) = xd;
		double mf = 0.0, pf = 0.0;
		if (k >= 1 && k < L) {
			const double je = g.jace[g3e + (size_t)k * NN];
			mf = je * seR(k) * xd;
			pf = je * seP(k) * xd;
		}
		mfe(k) = mf;
		pfe(k) = pf;
	}
	__syncwarp();
	for (int k = lane; k <= L; k += 32) {
		dkee(k) = tb_sm_apply(opDiffN2E, ken, k);
	}
	for (int q = lane; q < n * ldab; q += 32) {
		DG.p[q] = 0.0;
	}
	__syncwarp();

	const int vo = ca.fe_nodes;
	const int nfe = L / vo;
	const double dInvDeltaT = 1.0 / ca.dt;
#define TB_MAT(c0, k0, c1, k1) \
	DG(2 * offd + (3 * (k1) + (c1)) - (3 * (k0) + (c0)), 3 * (k0) + (c0))

	// ---- BuildF and the Jacobian rows of level k --------------------------------
	for (int k = lane; k <= L; k += 32) {
		double fP = 0.0, fW = 0.0, fR = 0.0;
		if (k < L) {
			const double invj = 1.0 / g.jac[g3 + (size_t)k * NN];
			const double dmfn = tb_sm_apply(opDiffE2N, mfe, k);
			const double dpfn = tb_sm_apply(opDiffE2N, pfe, k);
			fR = dmfn * invj;
			fP += dpfn * invj;
			// upwind penalty of rho-theta and rho (BuildF :2687-2713)
			const int a = k / vo;
			for (int c = 0; c < 2; c++) {
				const SmAcc & sn = (c == 0) ? snP : snR;
				double aux = 0.0;
				if (a <= nfe - 2) {
					aux += tb_sm_apply(opPenL, sn, k) * fabs(xde((a + 1) * vo));
				}
				if (a >= 1) {
					aux += tb_sm_apply(opPenR, sn, k) * fabs(xde(a * vo));
				}
				if (c == 0) fP -= aux; else fR -= aux;
			}
			// Jacobian: conservative flux terms (:3059-3091)
			for (int m = opDiffE2N.begin[k]; m < opDiffE2N.end[k]; m++) {
				const double je = g.jace[g3e + (size_t)m * NN];
				const double dm = tb_op_coeff(opDiffE2N, k, m);
				if ((m != 0) && (m != L)) {
					const double dMassFluxCoeff =
						dm * je * invj * g.cxe[2][g3e + (size_t)m * NN];
					TB_MAT(FW, m, FP, k) += dMassFluxCoeff * seP(m);
					TB_MAT(FW, m, FR, k) += dMassFluxCoeff * seR(m);
				}
				for (int q = opInterpN2E.begin[m]; q < opInterpN2E.end[m]; q++) {
					const double dCoeffVerticalFlux =
						dm * je * invj * tb_op_coeff(opInterpN2E, m, q) * xde(m);
					TB_MAT(FR, q, FR, k) += dCoeffVerticalFlux;
					TB_MAT(FP, q, FP, k) += dCoeffVerticalFlux;
				}
			}
		}
		if (k >= 1 && k < L) {
			const size_t o = g3e + (size_t)k * NN;
			const double dPressureGradientForce = dPe(k) * seP(k) / seR(k);
			double f = dPressureGradientForce;
			f += ph.g * g.dre[2][o];
			const double dCovUa = seU(k), dCovUb = seV(k), dCovUx = seW(k);
			const double dConUa = g.cae[0][o] * dCovUa + g.cae[1][o] * dCovUb + g.cae[2][o] * dCovUx;
			const double dConUb = g.cbe[0][o] * dCovUa + g.cbe[1][o] * dCovUb + g.cbe[2][o] * dCovUx;
			const double dCurlTerm = -dConUa * dUa(k) - dConUb * dUb(k);
			f += (dkee(k) + dCurlTerm);
			fW = f;
			// Jacobian rows of w (:3094-3140)
			const double dRHSWCoeffA = seP(k) * ph.R / (seR(k) * ph.cv);
			for (int m = opDiffN2E.begin[k]; m < opDiffN2E.end[k]; m++) {
				TB_MAT(FP, m, FW, k) +=
					dRHSWCoeffA * tb_op_coeff(opDiffN2E, k, m) * exn(m) / snP(m);
			}
			const double dRHSWCoeffB = 1.0 / (seR(k) * seR(k)) * dPe(k);
			for (int q = opInterpN2E.begin[k]; q < opInterpN2E.end[k]; q++) {
				const double dRHSWCoeffC = dRHSWCoeffB * tb_op_coeff(opInterpN2E, k, q);
				TB_MAT(FP, q, FW, k) += dRHSWCoeffC * seR(k);
				TB_MAT(FR, q, FW, k) += -dRHSWCoeffC * seP(k);
			}
			for (int l = opDiffN2E.begin[k]; l < opDiffN2E.end[k]; l++) {
				for (int m = opInterpE2N.begin[l]; m < opInterpE2N.end[l]; m++) {
					TB_MAT(FW, m, FW, k) +=
						tb_op_coeff(opInterpE2N, l, m) * tb_op_coeff(opDiffN2E, k, l) * xdn(l);
				}
			}
		}
		// upwinding of w on interfaces: F (:2676-2686), Jacobian (:2870-2897)
		{
			fW -= ca.upwind_coeff * fabs(xde(k)) * ddW(k);
			double dSignWeight;
			const double cx2 = g.cxe[2][g3e + (size_t)k * NN];
			if (xde(k) > 0.0) {
				dSignWeight = 1.0 * cx2;
			} else if (xde(k) < 0.0) {
				dSignWeight = -1.0 * cx2;
			} else {
				dSignWeight = 0.0;
			}
			TB_MAT(FW, k, FW, k) -= ca.upwind_coeff * dSignWeight * ddW(k);
			for (int q = opDDE2E.begin[k]; q < opDDE2E.end[k]; q++) {
				TB_MAT(FW, q, FW, k) -=
					ca.upwind_coeff * fabs(xde(k)) * tb_op_coeff(opDDE2E, k, q);
			}
		}
		// upwind penalty of rho-theta and rho: Jacobian (:2899-2968); a row in
		// finite element a receives its "right" terms (iteration a of the
		// reference loop) before its "left" terms (iteration a + 1)
		if (k < L) {
			const int a = k / vo;
			for (int c = 0; c < 2; c++) {
				const SmAcc & sn = (c == 0) ? snP : snR;
				const int fc = (c == 0) ? FP : FR;
				for (int side = 0; side < 2; side++) {
					const bool right = (side == 0);
					if (right && a < 1) continue;
					if (!right && a > nfe - 2) continue;
					const int ke = right ? (a * vo) : ((a + 1) * vo);
					const DevOp & op = right ? opPenR : opPenL;
					const double xd = xde(ke);
					const double dWeight = fabs(xd);
					const double cx2 = g.cxe[2][g3e + (size_t)ke * NN];
					double dSignWeight;
					if (xd > 0.0) {
						dSignWeight = 1.0 * cx2;
					} else if (xd < 0.0) {
						dSignWeight = -1.0 * cx2;
					} else {
						dSignWeight = 0.0;
					}
					for (int q = op.begin[k]; q < op.end[k]; q++) {
						TB_MAT(FW, ke, fc, k) -= dSignWeight * tb_op_coeff(op, k, q) * sn(q);
					}
					for (int q = op.begin[k]; q < op.end[k]; q++) {
						TB_MAT(fc, q, fc, k) -= dWeight * tb_op_coeff(op, k, q);
					}
				}
			}
		}
		if (k == 0 || k == L) fW = 0.0;      // :2747-2758
		TB_MAT(FP, k, FP, k) += dInvDeltaT;
		TB_MAT(FW, k, FW, k) += dInvDeltaT;
		TB_MAT(FR, k, FR, k) += dInvDeltaT;
		F(3 * k + FP) = fP;
		F(3 * k + FW) = fW;
		F(3 * k + FR) = fR;
	}
#undef TB_MAT
	__syncwarp();

	// ---- dgbsv: dgbtf2 with the forward substitution fused in, warp-wide ------
	const int kl = offd, ku = offd, kv = 2 * offd;
	int info = 0;
	int ju = 0;
	// dgbtf2 clears the fill-in rows before it uses them.  That matters: above
	// vertical order 1 the assembly leaves entries one place outside the band
	// (coefficients of 1e-13 that are zero analytically), which the band
	// storage folds into the fill-in rows of the neighbouring column - LAPACK,
	// and so the reference, drops them there.
	for (int q = lane; q < (kv - ku - 1) * kl; q += 32) {
		const int j = ku + 1 + q / kl;
		const int i = q % kl;
		if (j < n && i >= kv - j) DG(i, j) = 0.0;
	}
	__syncwarp();
	for (int j = 0; j < n; j++) {
		if (j + kv < n) {
			for (int i = lane; i < kl; i += 32) {
				DG(i, j + kv) = 0.0;
			}
			__syncwarp();
		}
		const int km = (kl < n - 1 - j) ? kl : (n - 1 - j);
		// idamax over rows j..j+km (first maximum)
		double v = -1.0;
		int jp = 0;
		for (int i = lane; i <= km; i += 32) {
			const double a = fabs(DG(kv + i, j));
			if (a > v) { v = a; jp = i; }
		}
		for (int off = 16; off > 0; off >>= 1) {
			const double ov = __shfl_down_sync(FULL, v, off);
			const int oi = __shfl_down_sync(FULL, jp, off);
			if (ov > v || (ov == v && oi < jp)) { v = ov; jp = oi; }
		}
		jp = __shfl_sync(FULL, jp, 0);
		const int piv = jp + j;
		const double pv = DG(kv + jp, j);
		if (pv != 0.0) {
			int cand = j + ku + jp;
			if (cand > n - 1) cand = n - 1;
			if (cand > ju) ju = cand;
			if (jp != 0) {
				for (int c = lane; c <= ju - j; c += 32) {
					const double tmp = DG(kv + jp - c, j + c);
					DG(kv + jp - c, j + c) = DG(kv - c, j + c);
					DG(kv - c, j + c) = tmp;
				}
			}
			__syncwarp();
			if (km > 0) {
				const double r = 1.0 / DG(kv, j);
				__syncwarp();
				for (int i = 1 + lane; i <= km; i += 32) {
					DG(kv + i, j) *= r;
				}
				__syncwarp();
				const int nc = ju - j;
				for (int q = lane; q < km * nc; q += 32) {
					const int i = 1 + q / nc;
					const int c = 1 + q % nc;
					const double y = DG(kv - c, j + c);
					if (y != 0.0) {
						DG(kv + i - c, j + c) -= DG(kv + i, j) * y;
					}
				}
			}
		} else if (info == 0) {
			info = j + 1;
		}
		if (j < n - 1) {
			if (lane == 0 && piv != j) {
				const double tmp = F(piv);
				F(piv) = F(j);
				F(j) = tmp;
			}
			__syncwarp();
			const double bj = F(j);
			for (int i = 1 + lane; i <= km; i += 32) {
				F(j + i) -= DG(kv + i, j) * bj;
			}
		}
		__syncwarp();
	}
	if (info == 0) {
		// dtbsv: upper, no transpose, non-unit diagonal, bandwidth kv
		for (int j = n - 1; j >= 0; j--) {
			const double bj = F(j);
			if (bj != 0.0) {
				const double temp = bj / DG(kv, j);
				__syncwarp();
				if (lane == 0) F(j) = temp;
				const int lo = (j - kv > 0) ? (j - kv) : 0;
				for (int i = j - 1 - lane; i >= lo; i -= 32) {
					F(i) -= temp * DG(kv - (j - i), j);
				}
			}
			__syncwarp();
		}
	}
	if (info != 0 || !(F(0) == F(0))) {
		if (lane == 0) atomicMax(ca.info, ca.col0 + tcol + 1);
	}
	if (ca.assemble_only) return;

	// ---- x = x0 - delta, scattered to the column and its duplicates ------------
	const int * dups = ca.col_dups + (size_t)(ca.col0 + tcol) * 3;
	for (int q = -1; q < 3; q++) {
		int tgt = node;
		if (q >= 0) {
			tgt = dups[q];
			if (tgt < 0) continue;
		}
		const long long te = tgt / NN;
		const int tn = tgt % NN;
		const size_t tb = (size_t)te * lay.nrows * NN + tn;
		double * oP = out + tb + (size_t)lay.rowoff[PIx] * NN;
		double * oW = out + tb + (size_t)lay.rowoff[WIx] * NN;
		double * oR = out + tb + (size_t)lay.rowoff[RIx] * NN;
		for (int k = lane; k <= L; k += 32) {
			if (k < L) {
				oP[(size_t)k * NN] = snP(k) - F(3 * k + FP);
				oR[(size_t)k * NN] = snR(k) - F(3 * k + FR);
			}
			oW[(size_t)k * NN] = seW(k) - F(3 * k + FW);
		}
	}
}


///////////////////////////////////////////////////////////////////////////////
// One thread per column, register-resident elimination block
// (vertical order 1: kl = ku = 4).
//
// The banded Jacobian is never materialised.  Rows are generated level by
// level straight from the state and the metric terms (inputs in a sliding
// register window, each loaded once; three rows of a level are staged in
// shared memory) and enter a 5 x 9 block of registers that holds rows
// j..j+4, columns j..j+8 of the partially eliminated matrix - all LAPACK's
// dgbtf2 touches at step j.  After the step the finished row j of U and the
// forward-substituted right-hand side stream to a global scratch laid out
// [entry][column]; the block shifts by one row and column and the next row
// enters.  The elimination is dgbtf2 step for step (same pivot search, row
// interchange, reciprocal scaling, rank-1 update), the back substitution
// subtracts in dtbsv's order, and every Jacobian entry is accumulated in the
// reference's statement order, so the arithmetic matches k_column_implicit.
// Per column the global traffic is the inputs (~6 kB) plus 10 n doubles
// written and read once (~15 kB at L = 30).

#define TBW_KL 4
#define TBW_KV 8
#define TBW_THREADS 128
#define TBW_STG 30     // staged doubles per thread: 3 rows x 9 band entries + 3 F

__host__ __device__ inline size_t tb_column_window_smem_bytes() {
	return (size_t)TBW_STG * TBW_THREADS * sizeof(double);
}
// scratch doubles per column: rows of U (9 n) + right-hand side (n)
__host__ __device__ inline int tb_column_window_scratch(int L) {
	return 10 * 3 * (L + 1);
}

__global__ void __launch_bounds__(TBW_THREADS)
k_column_implicit_window(
	DevLayout lay, DevGeom g, DevOps ops, DevPhys ph, ColumnArgs ca,
	const double * in, double * out   // may alias
) {
	TB_DYN_SMEM(double, sm);
	const int t = threadIdx.x;
	const int tcol = blockIdx.x * TBW_THREADS + t;
	if (tcol >= ca.ncols) return;

	const int UIx = 0, VIx = 1, PIx = 2, WIx = 3, RIx = 4;
	const int FP = 0, FW = 1, FR = 2;
	const int NN = lay.nn;
	const int L = lay.nlev;
	const int n = 3 * (L + 1);
	const int kl = TBW_KL;

	const int node = ca.col_node[ca.col0 + tcol];
	const long long e = node / NN;
	const int nd = node % NN;
	const size_t ebase = (size_t)e * lay.nrows * NN;
	const size_t g3 = (size_t)e * L * NN + nd;
	const size_t g3e = (size_t)e * (L + 1) * NN + nd;

	double * stg = sm + t;
	// staged rows of the level being generated: A(i, c), i in 3k..3k+2,
	// band entry c - i + 4 in 0..8
#define WIN(i, c) stg[(((i) - 3 * kcur) * 9 + ((c) - (i) + 4)) * TBW_THREADS]
#define FRING(i) stg[(27 + (i) - 3 * kcur) * TBW_THREADS]
	// scratch: U row j -> entries [9 j, 9 j + 9), right-hand side -> 9 n + j
	double * sc = ca.ws + tcol;
	const size_t S = (size_t)ca.ws_stride;

	const DevOp & opInterpN2E = ops.op[0];
	const DevOp & opInterpE2N = ops.op[1];
	const DevOp & opDiffN2E = ops.op[3];
	const DevOp & opDiffE2N = ops.op[4];
	const DevOp & opDDE2E = ops.op[7];
	const DevOp & opPenL = ops.op[8];
	const DevOp & opPenR = ops.op[9];

	const double * inU = in + ebase + (size_t)lay.rowoff[UIx] * NN + nd;
	const double * inV = in + ebase + (size_t)lay.rowoff[VIx] * NN + nd;
	const double * inP = in + ebase + (size_t)lay.rowoff[PIx] * NN + nd;
	const double * inW = in + ebase + (size_t)lay.rowoff[WIx] * NN + nd;
	const double * inR = in + ebase + (size_t)lay.rowoff[RIx] * NN + nd;

	// ---- column quantities (PrepareColumn, VerticalDynamicsFEM.cpp:1839-2179) --
	// Inputs are held in a register window that slides with the level being
	// assembled (levels kcur-2..kcur+1, interfaces kcur-1..kcur+1): each state
	// value is loaded from global memory once.  Anything outside the window
	// (never the case for order-1 operators) falls back to a global load.
	int kcur = 0;
	double Uw0 = 0.0, Uw1 = 0.0, Uw2 = 0.0, Uw3 = 0.0;
	double Vw0 = 0.0, Vw1 = 0.0, Vw2 = 0.0, Vw3 = 0.0;
	double Pw0 = 0.0, Pw1 = 0.0, Pw2 = 0.0, Pw3 = 0.0;
	double Rw0 = 0.0, Rw1 = 0.0, Rw2 = 0.0, Rw3 = 0.0;
	double Ww0 = 0.0, Ww1 = 0.0, Ww2 = 0.0;
	auto getU = [&](int l) {
		const int d = l - (kcur - 2);
		return (d == 0) ? Uw0 : (d == 1) ? Uw1 : (d == 2) ? Uw2 : (d == 3) ? Uw3 : inU[(size_t)l * NN];
	};
	auto getV = [&](int l) {
		const int d = l - (kcur - 2);
		return (d == 0) ? Vw0 : (d == 1) ? Vw1 : (d == 2) ? Vw2 : (d == 3) ? Vw3 : inV[(size_t)l * NN];
	};
	auto getP = [&](int l) {
		const int d = l - (kcur - 2);
		return (d == 0) ? Pw0 : (d == 1) ? Pw1 : (d == 2) ? Pw2 : (d == 3) ? Pw3 : inP[(size_t)l * NN];
	};
	auto getR = [&](int l) {
		const int d = l - (kcur - 2);
		return (d == 0) ? Rw0 : (d == 1) ? Rw1 : (d == 2) ? Rw2 : (d == 3) ? Rw3 : inR[(size_t)l * NN];
	};
	auto getW = [&](int m) {
		const int d = m - (kcur - 1);
		return (d == 0) ? Ww0 : (d == 1) ? Ww1 : (d == 2) ? Ww2 : inW[(size_t)m * NN];
	};
	auto slide = [&](int k) {
		// window for level k; called with k = 0, 1, 2, ... in order
		kcur = k;
		if (k == 0) {
			Uw2 = inU[0]; Vw2 = inV[0]; Pw2 = inP[0]; Rw2 = inR[0];
			if (L > 1) {
				Uw3 = inU[NN]; Vw3 = inV[NN]; Pw3 = inP[NN]; Rw3 = inR[NN];
			}
			Ww1 = inW[0];
			Ww2 = inW[NN];
		} else {
			Uw0 = Uw1; Uw1 = Uw2; Uw2 = Uw3;
			Vw0 = Vw1; Vw1 = Vw2; Vw2 = Vw3;
			Pw0 = Pw1; Pw1 = Pw2; Pw2 = Pw3;
			Rw0 = Rw1; Rw1 = Rw2; Rw2 = Rw3;
			if (k + 1 < L) {
				const size_t o = (size_t)(k + 1) * NN;
				Uw3 = inU[o]; Vw3 = inV[o]; Pw3 = inP[o]; Rw3 = inR[o];
			}
			Ww0 = Ww1; Ww1 = Ww2;
			if (k + 1 <= L) Ww2 = inW[(size_t)(k + 1) * NN];
		}
	};
#define TB_APPLY(op, m, get) ([&]() { \
		double o_ = 0.0; \
		const int b_ = (op).begin[m], e_ = (op).end[m]; \
		const double * c_ = (op).coeff + (size_t)(m) * (op).width; \
		for (int l_ = b_; l_ < e_; l_++) o_ += c_[l_ - b_] * get(l_); \
		return o_; }())
	auto seU = [&](int m) { return TB_APPLY(opInterpN2E, m, getU); };
	auto seV = [&](int m) { return TB_APPLY(opInterpN2E, m, getV); };
	auto seP = [&](int m) { return TB_APPLY(opInterpN2E, m, getP); };
	auto seR = [&](int m) { return TB_APPLY(opInterpN2E, m, getR); };
	auto snW = [&](int l) { return TB_APPLY(opInterpE2N, l, getW); };
	auto exn = [&](int l) {
		return ph.cp * exp(ph.exner_c1 * log(ph.exner_c2 * getP(l)));
	};
	auto xdn = [&](int l) {
		const size_t o = g3 + (size_t)l * NN;
		return g.cx[0][o] * getU(l) + g.cx[1][o] * getV(l) + g.cx[2][o] * snW(l);
	};
	auto ken = [&](int l) {
		const size_t o = g3 + (size_t)l * NN;
		const double dCovUa = getU(l), dCovUb = getV(l);
		const double dCovUx = snW(l);
		const double cx0 = g.cx[0][o], cx1 = g.cx[1][o], cx2 = g.cx[2][o];
		const double dConUa = g.ca[0][o] * dCovUa + g.ca[1][o] * dCovUb + g.ca[2][o] * dCovUx;
		const double dConUb = g.cb[0][o] * dCovUa + g.cb[1][o] * dCovUb + g.cb[2][o] * dCovUx;
		const double dConUx = cx0 * dCovUa + cx1 * dCovUb + cx2 * dCovUx;
		return 0.5 * (dConUa * dCovUa + dConUb * dCovUb + dConUx * dCovUx);
	};
	auto ddW = [&](int m) {
		if (m <= 0 || m >= L) return 0.0;
		return TB_APPLY(opDDE2E, m, getW);
	};
	// interface metrics of the interface above the current level are reused
	// as those of the current interface one level later
	double cxe0_k = 0.0, cxe1_k = 0.0, cxe2_k = 0.0, jace_k = 0.0;
	double cxe0_p = 0.0, cxe1_p = 0.0, cxe2_p = 0.0, jace_p = 0.0;
	auto load_edge_metrics = [&](int k) {
		if (k == 0) {
			cxe0_k = g.cxe[0][g3e]; cxe1_k = g.cxe[1][g3e]; cxe2_k = g.cxe[2][g3e];
			jace_k = g.jace[g3e];
		} else {
			cxe0_k = cxe0_p; cxe1_k = cxe1_p; cxe2_k = cxe2_p; jace_k = jace_p;
		}
		if (k + 1 <= L) {
			const size_t o = g3e + (size_t)(k + 1) * NN;
			cxe0_p = g.cxe[0][o]; cxe1_p = g.cxe[1][o]; cxe2_p = g.cxe[2][o];
			jace_p = g.jace[o];
		}
	};
	auto cxe2_at = [&](int m) {
		return (m == kcur) ? cxe2_k : (m == kcur + 1) ? cxe2_p : g.cxe[2][g3e + (size_t)m * NN];
	};
	auto jace_at = [&](int m) {
		return (m == kcur) ? jace_k : (m == kcur + 1) ? jace_p : g.jace[g3e + (size_t)m * NN];
	};
	auto xde = [&](int m) {
		if (m <= 0 || m >= L) return 0.0;
		double c0, c1, c2;
		if (m == kcur) { c0 = cxe0_k; c1 = cxe1_k; c2 = cxe2_k; }
		else if (m == kcur + 1) { c0 = cxe0_p; c1 = cxe1_p; c2 = cxe2_p; }
		else {
			const size_t o = g3e + (size_t)m * NN;
			c0 = g.cxe[0][o]; c1 = g.cxe[1][o]; c2 = g.cxe[2][o];
		}
		return c0 * seU(m) + c1 * seV(m) + c2 * getW(m);
	};

	const int vo = ca.fe_nodes;
	const int nfe = L / vo;
	const double dInvDeltaT = 1.0 / ca.dt;


	// exn / ken of the two levels around the current interface are reused
	int cache_l = -1000;
	double exn_a = 0.0, exn_b = 0.0, ken_a = 0.0, ken_b = 0.0;   // levels cache_l-1, cache_l
	double xdn_a = 0.0, xdn_b = 0.0;
	double xde_k = 0.0, xde_kp1 = 0.0;

	// ---- rows of level k into the window (BuildF + Jacobian) -------------------
	auto assemble_level = [&](int k) {
		// slide the input window, then the caches: interface quantities at k
		// and k+1, level quantities at k-1 and k
		slide(k);
		load_edge_metrics(k);
#pragma unroll
		for (int q = 0; q < 27; q++) {
			stg[q * TBW_THREADS] = 0.0;
		}
		xde_k = (k == 0) ? 0.0 : xde_kp1;
		xde_kp1 = xde(k + 1);
		if (k < L) {
			if (cache_l == k - 1 && k >= 1) {
				exn_a = exn_b; ken_a = ken_b; xdn_a = xdn_b;
			} else if (k >= 1) {
				exn_a = exn(k - 1); ken_a = ken(k - 1); xdn_a = xdn(k - 1);
			}
			exn_b = exn(k); ken_b = ken(k); xdn_b = xdn(k);
			cache_l = k;
		}
		auto xdn_at = [&](int l) {
			if (l == cache_l - 1 && l >= 0) return xdn_a;
			if (l == cache_l) return xdn_b;
			return xdn(l);
		};
		auto exn_at = [&](int l) {
			if (l == cache_l - 1 && l >= 0) return exn_a;
			if (l == cache_l) return exn_b;
			return exn(l);
		};
		auto ken_at = [&](int l) {
			if (l == cache_l - 1 && l >= 0) return ken_a;
			if (l == cache_l) return ken_b;
			return ken(l);
		};
		auto xde_at = [&](int m) {
			if (m == k) return xde_k;
			if (m == k + 1) return xde_kp1;
			return xde(m);
		};
		auto mfe = [&](int m) {
			if (m <= 0 || m >= L) return 0.0;
			return jace_at(m) * seR(m) * xde_at(m);
		};
		auto pfe = [&](int m) {
			if (m <= 0 || m >= L) return 0.0;
			return jace_at(m) * seP(m) * xde_at(m);
		};

		double fP = 0.0, fW = 0.0, fR = 0.0;
		const int rP = 3 * k + FP, rW = 3 * k + FW, rR = 3 * k + FR;
		if (k < L) {
			const double invj = 1.0 / g.jac[g3 + (size_t)k * NN];
			double dmfn = 0.0, dpfn = 0.0;
			for (int m = opDiffE2N.begin[k]; m < opDiffE2N.end[k]; m++) {
				const double c = tb_op_coeff(opDiffE2N, k, m);
				dmfn += c * mfe(m);
				dpfn += c * pfe(m);
			}
			fR = dmfn * invj;
			fP += dpfn * invj;
			const int a = k / vo;
			for (int cc = 0; cc < 2; cc++) {
				double aux = 0.0;
				if (a <= nfe - 2) {
					const double pl = (cc == 0) ? TB_APPLY(opPenL, k, getP) : TB_APPLY(opPenL, k, getR);
					aux += pl * fabs(xde_at((a + 1) * vo));
				}
				if (a >= 1) {
					const double pr = (cc == 0) ? TB_APPLY(opPenR, k, getP) : TB_APPLY(opPenR, k, getR);
					aux += pr * fabs(xde_at(a * vo));
				}
				if (cc == 0) fP -= aux; else fR -= aux;
			}
			for (int m = opDiffE2N.begin[k]; m < opDiffE2N.end[k]; m++) {
				const double je = jace_at(m);
				const double dm = tb_op_coeff(opDiffE2N, k, m);
				if ((m != 0) && (m != L)) {
					const double dMassFluxCoeff =
						dm * je * invj * cxe2_at(m);
					WIN(rP, 3 * m + FW) += dMassFluxCoeff * seP(m);
					WIN(rR, 3 * m + FW) += dMassFluxCoeff * seR(m);
				}
				for (int q = opInterpN2E.begin[m]; q < opInterpN2E.end[m]; q++) {
					const double dCoeffVerticalFlux =
						dm * je * invj * tb_op_coeff(opInterpN2E, m, q) * xde_at(m);
					WIN(rR, 3 * q + FR) += dCoeffVerticalFlux;
					WIN(rP, 3 * q + FP) += dCoeffVerticalFlux;
				}
			}
		}
		if (k >= 1 && k < L) {
			const size_t o = g3e + (size_t)k * NN;
			double dPe = 0.0, dkee = 0.0;
			for (int l = opDiffN2E.begin[k]; l < opDiffN2E.end[k]; l++) {
				const double c = tb_op_coeff(opDiffN2E, k, l);
				if (c != 0.0) {
					dPe += c * exn_at(l);
					dkee += c * ken_at(l);
				}
			}
			const double sePk = seP(k), seRk = seR(k);
			const double dPressureGradientForce = dPe * sePk / seRk;
			double f = dPressureGradientForce;
			f += ph.g * g.dre[2][o];
			const double dCovUa = seU(k), dCovUb = seV(k), dCovUx = getW(k);
			const double dConUa = g.cae[0][o] * dCovUa + g.cae[1][o] * dCovUb + g.cae[2][o] * dCovUx;
			const double dConUb = g.cbe[0][o] * dCovUa + g.cbe[1][o] * dCovUb + g.cbe[2][o] * dCovUx;
			const double dUa = TB_APPLY(opDiffN2E, k, getU);
			const double dUb = TB_APPLY(opDiffN2E, k, getV);
			const double dCurlTerm = -dConUa * dUa - dConUb * dUb;
			f += (dkee + dCurlTerm);
			fW = f;
			const double dRHSWCoeffA = sePk * ph.R / (seRk * ph.cv);
			for (int m = opDiffN2E.begin[k]; m < opDiffN2E.end[k]; m++) {
				const double c = tb_op_coeff(opDiffN2E, k, m);
				if (c != 0.0) {
					WIN(rW, 3 * m + FP) += dRHSWCoeffA * c * exn_at(m) / getP(m);
				}
			}
			const double dRHSWCoeffB = 1.0 / (seRk * seRk) * dPe;
			for (int q = opInterpN2E.begin[k]; q < opInterpN2E.end[k]; q++) {
				const double dRHSWCoeffC = dRHSWCoeffB * tb_op_coeff(opInterpN2E, k, q);
				WIN(rW, 3 * q + FP) += dRHSWCoeffC * seRk;
				WIN(rW, 3 * q + FR) += -dRHSWCoeffC * sePk;
			}
			for (int l = opDiffN2E.begin[k]; l < opDiffN2E.end[k]; l++) {
				const double cl = tb_op_coeff(opDiffN2E, k, l);
				if (cl == 0.0) continue;
				const double xn = xdn_at(l);
				for (int m = opInterpE2N.begin[l]; m < opInterpE2N.end[l]; m++) {
					WIN(rW, 3 * m + FW) += tb_op_coeff(opInterpE2N, l, m) * cl * xn;
				}
			}
		}
		{
			const double d2 = ddW(k);
			fW -= ca.upwind_coeff * fabs(xde_k) * d2;
			double dSignWeight;
			const double cx2 = cxe2_k;
			if (xde_k > 0.0) {
				dSignWeight = 1.0 * cx2;
			} else if (xde_k < 0.0) {
				dSignWeight = -1.0 * cx2;
			} else {
				dSignWeight = 0.0;
			}
			WIN(rW, rW) -= ca.upwind_coeff * dSignWeight * d2;
			for (int q = opDDE2E.begin[k]; q < opDDE2E.end[k]; q++) {
				WIN(rW, 3 * q + FW) -=
					ca.upwind_coeff * fabs(xde_k) * tb_op_coeff(opDDE2E, k, q);
			}
		}
		if (k < L) {
			const int a = k / vo;
			for (int cc = 0; cc < 2; cc++) {
				const int fc = (cc == 0) ? FP : FR;
				const int rr = 3 * k + fc;
				for (int side = 0; side < 2; side++) {
					const bool right = (side == 0);
					if (right && a < 1) continue;
					if (!right && a > nfe - 2) continue;
					const int ke = right ? (a * vo) : ((a + 1) * vo);
					const DevOp & op = right ? opPenR : opPenL;
					const double xd = xde_at(ke);
					const double dWeight = fabs(xd);
					const double cx2 = cxe2_at(ke);
					double dSignWeight;
					if (xd > 0.0) {
						dSignWeight = 1.0 * cx2;
					} else if (xd < 0.0) {
						dSignWeight = -1.0 * cx2;
					} else {
						dSignWeight = 0.0;
					}
					for (int q = op.begin[k]; q < op.end[k]; q++) {
						WIN(rr, 3 * ke + FW) -=
							dSignWeight * tb_op_coeff(op, k, q) * ((cc == 0) ? getP(q) : getR(q));
					}
					for (int q = op.begin[k]; q < op.end[k]; q++) {
						WIN(rr, 3 * q + fc) -= dWeight * tb_op_coeff(op, k, q);
					}
				}
			}
		}
		if (k == 0 || k == L) fW = 0.0;
		WIN(rP, rP) += dInvDeltaT;
		WIN(rW, rW) += dInvDeltaT;
		WIN(rR, rR) += dInvDeltaT;
		FRING(rP) = fP;
		FRING(rW) = fW;
		FRING(rR) = fR;
	};

	// ---- dgbtf2 + forward substitution on the register block -------------------
	// B[r][c] = A(j + r, j + c), bb[r] = b(j + r)
	double B[5][9];
	double bb[5];
#pragma unroll
	for (int r = 0; r < 5; r++) {
#pragma unroll
		for (int c = 0; c < 9; c++) B[r][c] = 0.0;
		bb[r] = 0.0;
	}
	// rows 0..4 (levels 0 and 1)
	assemble_level(0);
#pragma unroll
	for (int r = 0; r < 3; r++) {
#pragma unroll
		for (int c = 0; c < 9; c++) {
			// row r holds columns r-4..r+4: column c is band entry c - r + 4
			if (c - r + 4 >= 0 && c - r + 4 <= 8) B[r][c] = stg[(r * 9 + (c - r + 4)) * TBW_THREADS];
		}
		bb[r] = stg[(27 + r) * TBW_THREADS];
	}
	if (L >= 1) {
		assemble_level(1);
#pragma unroll
		for (int r = 3; r < 5; r++) {
#pragma unroll
			for (int c = 0; c < 9; c++) {
				if (c - r + 4 >= 0 && c - r + 4 <= 8) {
					B[r][c] = stg[((r - 3) * 9 + (c - r + 4)) * TBW_THREADS];
				}
			}
			bb[r] = stg[(27 + r - 3) * TBW_THREADS];
		}
	}
	int info = 0;
	for (int j = 0; j < n; j++) {
		const int km = (kl < n - 1 - j) ? kl : (n - 1 - j);
		// idamax over rows j..j+km (first maximum)
		int jp = 0;
		double amax = fabs(B[0][0]);
#pragma unroll
		for (int r = 1; r < 5; r++) {
			const double v = fabs(B[r][0]);
			if (r <= km && v > amax) { amax = v; jp = r; }
		}
		// interchange rows j and j + jp (entries beyond ju are zero in both)
		if (jp != 0) {
#pragma unroll
			for (int c = 0; c < 9; c++) {
				const double t0 = B[0][c];
				const double s0 = (jp == 1) ? B[1][c] : (jp == 2) ? B[2][c] : (jp == 3) ? B[3][c] : B[4][c];
				B[0][c] = s0;
				if (jp == 1) B[1][c] = t0;
				if (jp == 2) B[2][c] = t0;
				if (jp == 3) B[3][c] = t0;
				if (jp == 4) B[4][c] = t0;
			}
			if (j < n - 1) {
				const double t0 = bb[0];
				const double s0 = (jp == 1) ? bb[1] : (jp == 2) ? bb[2] : (jp == 3) ? bb[3] : bb[4];
				bb[0] = s0;
				if (jp == 1) bb[1] = t0;
				if (jp == 2) bb[2] = t0;
				if (jp == 3) bb[3] = t0;
				if (jp == 4) bb[4] = t0;
			}
		}
		if (B[0][0] != 0.0) {
			if (km > 0) {
				const double rcp = 1.0 / B[0][0];
				const double bj = bb[0];
#pragma unroll
				for (int r = 1; r < 5; r++) {
					if (r <= km) {
						const double mult = B[r][0] * rcp;
#pragma unroll
						for (int c = 1; c < 9; c++) {
							const double y = B[0][c];
							if (y != 0.0) B[r][c] -= mult * y;
						}
						bb[r] -= mult * bj;
					}
				}
			}
		} else if (info == 0) {
			info = j + 1;
		}
		// row j of U and entry j of the right-hand side are final
#pragma unroll
		for (int c = 0; c < 9; c++) {
			sc[(size_t)(9 * j + c) * S] = B[0][c];
		}
		sc[(size_t)(9 * n + j) * S] = bb[0];
		// shift the block; row j + 5 enters
#pragma unroll
		for (int r = 0; r < 4; r++) {
#pragma unroll
			for (int c = 0; c < 8; c++) B[r][c] = B[r + 1][c + 1];
			B[r][8] = 0.0;
			bb[r] = bb[r + 1];
		}
		const int inew = j + 5;
		if (inew < n) {
			if (inew % 3 == 0) assemble_level(inew / 3);
			const int rl = inew % 3;
#pragma unroll
			for (int c = 0; c < 9; c++) B[4][c] = stg[(rl * 9 + c) * TBW_THREADS];
			bb[4] = stg[(27 + rl) * TBW_THREADS];
		} else {
#pragma unroll
			for (int c = 0; c < 9; c++) B[4][c] = 0.0;
			bb[4] = 0.0;
		}
	}
	if (info != 0) {
		atomicMax(ca.info, ca.col0 + tcol + 1);
		return;
	}
	if (ca.assemble_only) return;

	// ---- back substitution in dtbsv's order; x = x0 - delta is scattered as
	//      each unknown appears ------------------------------------------------
	const int * dups = ca.col_dups + (size_t)(ca.col0 + tcol) * 3;
	const int d0 = dups[0], d1 = dups[1], d2 = dups[2];
	double xr[8];      // x(j+1) .. x(j+8)
#pragma unroll
	for (int c = 0; c < 8; c++) xr[c] = 0.0;
	bool nan_seen = false;
	for (int j = n - 1; j >= 0; j--) {
		double acc = sc[(size_t)(9 * n + j) * S];
#pragma unroll
		for (int c = 8; c >= 1; c--) {
			const double xc = xr[c - 1];
			if (xc != 0.0) acc -= xc * sc[(size_t)(9 * j + c) * S];
		}
		double xj = acc;
		if (xj != 0.0) xj = xj / sc[(size_t)(9 * j) * S];
#pragma unroll
		for (int c = 7; c >= 1; c--) xr[c] = xr[c - 1];
		xr[0] = xj;
		if (j == 0 && !(xj == xj)) nan_seen = true;
		const int k = j / 3;
		const int c3 = j - 3 * k;
		if (c3 == FW || k < L) {
			const int rowoff = (c3 == FP) ? lay.rowoff[PIx] : ((c3 == FW) ? lay.rowoff[WIx] : lay.rowoff[RIx]);
			const double x0 = in[ebase + (size_t)(rowoff + k) * NN + nd];
			const double xnew = x0 - xj;
			out[ebase + (size_t)(rowoff + k) * NN + nd] = xnew;
			if (d0 >= 0) out[((size_t)(d0 / NN) * lay.nrows + rowoff + k) * NN + (d0 % NN)] = xnew;
			if (d1 >= 0) out[((size_t)(d1 / NN) * lay.nrows + rowoff + k) * NN + (d1 % NN)] = xnew;
			if (d2 >= 0) out[((size_t)(d2 / NN) * lay.nrows + rowoff + k) * NN + (d2 % NN)] = xnew;
		}
	}
	if (nan_seen) atomicMax(ca.info, ca.col0 + tcol + 1);
#undef WIN
#undef FRING
}

#endif
