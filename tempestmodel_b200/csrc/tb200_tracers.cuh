// Tracer kernels of the vertical dynamics and the positive-definite filters.
//
//   VerticalDynamicsFEM::UpdateColumnTracers     VerticalDynamicsFEM.cpp:3783-4282
//   VerticalDynamicsFEM::FilterNegativeTracers   VerticalDynamicsFEM.cpp:4286-4347
//   HorizontalDynamicsFEM::FilterNegativeTracers HorizontalDynamicsFEM.cpp:213-317
//
// General kernels (any vertical order, stored 3-D metric): one thread per
// column / per (element, tracer, level).  The column update restates the
// reference for its compiled configuration (implicit vertical, upwinding on
// rho and tracers, no uniform diffusion): one banded matrix per column
// (kl = ku = 2 vo - 1), LAPACK dgbtrf + dgbtrs per tracer (the factorisation is
// repeated per tracer here: same arithmetic).
#ifndef TB200_TRACERS_CUH
#define TB200_TRACERS_CUH

#include "tb200_platform.h"
#include "tb200_device.h"
#include "tb200_column.cuh"
#include "tb200_kernels.cuh"

struct TracerColumnArgs {
	const int * col_node;
	const int * col_dups;
	int ncols;
	int col0;
	double * ws;              // workspace [entries][ws_stride]
	int ws_stride;
	double dt;
	int fe_nodes;             // nodes per vertical finite element
	int kl;                   // 2 vo - 1
	const double * w_old;     // [e][L+1][NN]: w before the implicit solve
	int * info;
	// column-constant metric (tb200_fast.cuh) instead of the stored 3-D arrays:
	// colc [e][15][NN], lev [L+1][TBF_LW]; 0 = read the arrays of DevGeom
	const double * colc;
	const double * lev;
	// --explicitvertical (VerticalDynamicsFEM.cpp:802-810): every element-local
	// node is a column (col_node = 0), xi-dot from the initial w, no matrix
	// beyond the identity / dt, no jump terms, no duplicates to copy to
	int fully_explicit;
};

// entries of the column constants / level rows this kernel reads (tb200_fast.cuh)
#define TBT_C_JAC 3
#define TBT_C_A2 6
#define TBT_C_B2 7
#define TBT_C_X0 8
#define TBT_C_X2 9
#define TBT_C_NC 15
#define TBT_L_SE 18
#define TBT_L_LW 32

__host__ __device__ inline int tb_tracer_ws_entries(int L, int kl) {
	return 8 * (L + 1) + (3 * kl + 1) * L;
}

__global__ void k_column_tracers(
	DevLayout lay, DevGeom g, DevOps ops, TracerColumnArgs ta,
	const double * st_in,     // state before the solve (u, v)
	const double * st_out,    // state after the solve (w)
	const double * tr_in,     // instance holding the initial tracers
	double * tr_out           // instance whose tracers are updated
) {
	const int tcol = blockIdx.x * blockDim.x + threadIdx.x;
	if (tcol >= ta.ncols) return;
	const int UIx = 0, VIx = 1, WIx = 3;
	const int NN = lay.nn;
	const int L = lay.nlev;
	const int kl = ta.kl;
	const int ldab = 3 * kl + 1;

	const bool expl = (ta.fully_explicit != 0);
	const int node = expl ? (ta.col0 + tcol) : ta.col_node[ta.col0 + tcol];
	const long long e = node / NN;
	const int nd = node % NN;
	const size_t ebase = (size_t)e * lay.nrows * NN;
	const size_t g3 = (size_t)e * L * NN + nd;
	const size_t g3e = (size_t)e * (L + 1) * NN + nd;

	double * w0 = ta.ws + tcol;
	const int S = ta.ws_stride;
	int cur = 0;
#define TB_WS(name, len) WsAcc name = {w0 + (size_t)cur * S, S}; cur += (len)
	TB_WS(snU, L + 1); TB_WS(snV, L + 1);
	TB_WS(xdi, L + 1);        // xi-dot on interfaces from the state before the solve
	TB_WS(xdn, L + 1);        // ... from the updated w
	TB_WS(qn, L + 1); TB_WS(qe, L + 1); TB_WS(mf, L + 1); TB_WS(F, L + 1);
#undef TB_WS
	WsBand AB = {w0 + (size_t)cur * S, S, ldab};

	const DevOp & opInterpN2E = ops.op[0];
	const DevOp & opDiffE2N = ops.op[4];
	const DevOp & opPenL = ops.op[8];
	const DevOp & opPenR = ops.op[9];

	const double * inU = st_in + ebase + (size_t)lay.rowoff[UIx] * NN + nd;
	const double * inV = st_in + ebase + (size_t)lay.rowoff[VIx] * NN + nd;
	const double * wNew = st_out + ebase + (size_t)lay.rowoff[WIx] * NN + nd;
	// (explicit: the initial w throughout, :4048-4061)
	const double * wOld = expl ? (st_in + ebase + (size_t)lay.rowoff[WIx] * NN + nd) : (ta.w_old + g3e);

	// metric of the column: Jacobian on level k / interface m, xi-row of the
	// contravariant metric on interface m
	const bool fastm = (ta.colc != 0);
	const double * ccol = fastm ? (ta.colc + (size_t)e * TBT_C_NC * NN + nd) : 0;
	const double cJ = fastm ? ccol[TBT_C_JAC * NN] : 0.0;
	const double cA2 = fastm ? ccol[TBT_C_A2 * NN] : 0.0;
	const double cB2 = fastm ? ccol[TBT_C_B2 * NN] : 0.0;
	const double cX0 = fastm ? ccol[TBT_C_X0 * NN] : 0.0;
	const double cX2 = fastm ? ccol[TBT_C_X2 * NN] : 0.0;
	auto jac_at = [&](int k) { return fastm ? cJ : g.jac[g3 + (size_t)k * NN]; };
	auto jace_at = [&](int m) { return fastm ? cJ : g.jace[g3e + (size_t)m * NN]; };
	auto cxe_at = [&](int m, double & c0, double & c1, double & c2) {
		if (fastm) {
			const double se = ta.lev[(size_t)m * TBT_L_LW + TBT_L_SE];
			c0 = se * cA2; c1 = se * cB2; c2 = cX0 + (se * se) * cX2;
		} else {
			const size_t o = g3e + (size_t)m * NN;
			c0 = g.cxe[0][o]; c1 = g.cxe[1][o]; c2 = g.cxe[2][o];
		}
	};

	// SetupReferenceColumn: u, v on interfaces (:1643-1835)
	for (int k = 0; k < L; k++) {
		snU(k) = inU[(size_t)k * NN];
		snV(k) = inV[(size_t)k * NN];
	}
	// xi-dot on interfaces before (:3933-3950) and after (:4078-4093) the solve
	for (int k = 0; k <= L; k++) {
		double a = 0.0, b = 0.0;
		if (k >= 1 && k < L) {
			const double ue = tb_ws_apply(opInterpN2E, snU, k);
			const double ve = tb_ws_apply(opInterpN2E, snV, k);
			double c0, c1, c2;
			cxe_at(k, c0, c1, c2);
			a = c0 * ue + c1 * ve + c2 * wOld[(size_t)k * NN];
			b = c0 * ue + c1 * ve + c2 * wNew[(size_t)k * NN];
		}
		xdi(k) = a;
		xdn(k) = b;
	}

	const int vo = ta.fe_nodes;
	const int nfe = L / vo;
	const int * dups = expl ? 0 : (ta.col_dups + (size_t)(ta.col0 + tcol) * 3);

	for (int c = 0; c < lay.ntr; c++) {
		// ---- matrix (:3953-4016) ---------------------------------------------------
		for (int j = 0; j < L; j++) {
			for (int r = 0; r < ldab; r++) AB(r, j) = 0.0;
		}
		// TracerMatFIx(n, k): column n, row k -> band row 2 kl + k - n
#define TB_TMAT(n, k) AB(2 * kl + (k) - (n), (n))
		// (only with implicit advection, :3908)
		for (int k = 0; k < (expl ? 0 : L); k++) {
			const double jn = jac_at(k);
			for (int m = opDiffE2N.begin[k]; m < opDiffE2N.end[k]; m++) {
				const double je = jace_at(m);
				for (int n = opInterpN2E.begin[m]; n < opInterpN2E.end[m]; n++) {
					TB_TMAT(n, k) +=
						tb_op_coeff(opDiffE2N, k, m)
						* je
						/ jn
						* tb_op_coeff(opInterpN2E, m, n)
						* xdi(m);
				}
			}
		}
		for (int a = 1; a < (expl ? 0 : nfe); a++) {
			const int kLeftBegin = (a - 1) * vo, kLeftEnd = a * vo;
			const int kRightBegin = a * vo, kRightEnd = (a + 1) * vo;
			const double dWeight = fabs(xdi(kLeftEnd));
			for (int k = kLeftBegin; k < kLeftEnd; k++) {
				for (int n = opPenL.begin[k]; n < opPenL.end[k]; n++) {
					TB_TMAT(n, k) -= dWeight * tb_op_coeff(opPenL, k, n);
				}
			}
			for (int k = kRightBegin; k < kRightEnd; k++) {
				for (int n = opPenR.begin[k]; n < opPenR.end[k]; n++) {
					TB_TMAT(n, k) -= dWeight * tb_op_coeff(opPenR, k, n);
				}
			}
		}
		for (int k = 0; k < L; k++) {
			TB_TMAT(k, k) += 1.0 / ta.dt;
		}
#undef TB_TMAT

		// ---- right-hand side (:4096-4230) ------------------------------------------
		const double * tin = tr_in + ebase + (size_t)(lay.troff + c * L) * NN + nd;
		for (int k = 0; k < L; k++) {
			qn(k) = tin[(size_t)k * NN];
		}
		for (int k = 0; k <= L; k++) {
			qe(k) = tb_ws_apply(opInterpN2E, qn, k);
		}
		for (int k = 0; k <= L; k++) {
			mf(k) = jace_at(k) * qe(k) * xdn(k);
		}
		mf(0) = 0.0;
		mf(L) = 0.0;
		for (int k = 0; k < L; k++) {
			F(k) = tb_ws_apply(opDiffE2N, mf, k) / jac_at(k);
		}
		// upwind penalty with the weights of the state before the solve
		// (LinearColumnDiscPenaltyFEM::Apply, LinearColumnOperatorFEM.cpp:1863-1887)
		for (int k = 0; k < L; k++) {
			const int a = k / vo;
			double aux = 0.0;
			if (a <= nfe - 2) {
				aux += tb_ws_apply(opPenL, qn, k) * fabs(xdi((a + 1) * vo));
			}
			if (a >= 1) {
				aux += tb_ws_apply(opPenR, qn, k) * fabs(xdi(a * vo));
			}
			F(k) -= aux;
		}
		// jump terms from the change of w (:4178-4228; implicit advection only)
		for (int a = 1; a < (expl ? 0 : nfe); a++) {
			const int kLeftBegin = (a - 1) * vo, kLeftEnd = a * vo;
			const int kRightBegin = a * vo, kRightEnd = (a + 1) * vo;
			const double xd = xdi(kLeftEnd);
			double cx0_, cx1_, cx2;
			cxe_at(kLeftEnd, cx0_, cx1_, cx2);
			double dSignWeight;
			if (xd > 0.0) {
				dSignWeight = 1.0 * cx2;
			} else if (xd < 0.0) {
				dSignWeight = -1.0 * cx2;
			} else {
				dSignWeight = 0.0;
			}
			const double dLeftJumpConUx = dSignWeight
				* (wNew[(size_t)kLeftEnd * NN] - wOld[(size_t)kLeftEnd * NN]);
			for (int k = kLeftBegin; k < kLeftEnd; k++) {
				for (int n = opPenL.begin[k]; n < opPenL.end[k]; n++) {
					F(k) -= tb_op_coeff(opPenL, k, n) * qn(n) * dLeftJumpConUx;
				}
			}
			const double dRightJumpConUx = dSignWeight
				* (wNew[(size_t)kRightBegin * NN] - wOld[(size_t)kRightBegin * NN]);
			for (int k = kRightBegin; k < kRightEnd; k++) {
				for (int n = opPenR.begin[k]; n < opPenR.end[k]; n++) {
					F(k) -= tb_op_coeff(opPenR, k, n) * qn(n) * dRightJumpConUx;
				}
			}
		}

		// ---- DGBTRF + DGBTRS (:4026-4043, 4233-4262) -----------------------------
		const int r = tb_dgbsv(L, kl, kl, AB, F);
		if (r != 0 || !(F(0) == F(0))) {
			atomicMax(ta.info, ta.col0 + tcol + 1);
		}

		// ---- update and copy to the duplicates (:4265-4281, 1544-1633) -------------
		double * own = tr_out + ebase + (size_t)(lay.troff + c * L) * NN + nd;
		for (int k = 0; k < L; k++) {
			const double v = own[(size_t)k * NN] - F(k);
			own[(size_t)k * NN] = v;
			for (int q = 0; q < (expl ? 0 : 3); q++) {
				const int tgt = dups[q];
				if (tgt < 0) continue;
				tr_out[(size_t)(tgt / NN) * lay.nrows * NN + (tgt % NN)
					+ (size_t)(lay.troff + c * L + k) * NN] = v;
			}
		}
	}
}

// HorizontalDynamicsFEM::FilterNegativeTracers: per (element, tracer, level),
// mass-preserving clip of negatives within the element (:262-309)
__global__ void k_filter_tracers_element(
	DevLayout lay, const double * area_node, double * data
) {
	const int NN = lay.nn;
	const int L = lay.nlev;
	const long long nitems = lay.nelem * lay.ntr * L;
	const long long item = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (item >= nitems) return;
	const long long e = item / (lay.ntr * L);
	const int r = (int)(item % (lay.ntr * L));
	const int k = r % L;
	double * p = data + ((size_t)e * lay.nrows + lay.troff + r) * NN;
	const double * a = area_node + ((size_t)e * L + k) * NN;
	double dTotalMass = 0.0;
	double dNonNegativeMass = 0.0;
	for (int n = 0; n < NN; n++) {
		const double v = p[n];
		const double dPointwiseMass = v * a[n];
		dTotalMass += dPointwiseMass;
		if (v >= 0.0) {
			dNonNegativeMass += dPointwiseMass;
		}
	}
	const double dR = dTotalMass / dNonNegativeMass;
	for (int n = 0; n < NN; n++) {
		const double v = p[n];
		p[n] = (v > 0.0) ? v * dR : 0.0;
	}
}

// VerticalDynamicsFEM::FilterNegativeTracers: per (node, tracer), within the column.
// Optionally fused with what the time schemes put around it:
//  - combine: Grid::LinearCombineData(ca) of the tracer rows first (the Strang
//    carry-over, TimestepSchemeStrang.cpp:470-482), in k_lincomb's operation order;
//  - inc: Grid::LinearCombineData({+1, -1}) afterwards, inc = filtered - inc (the
//    tail of the Strang step, :658-672; inc holds the tracers from before the
//    column update).
__global__ void k_filter_tracers_column(
	DevLayout lay, const double * area_node, double * data,
	CombineArgs ca, int combine, double * inc
) {
	const int NN = lay.nn;
	const int L = lay.nlev;
	const long long nitems = lay.nelem * NN * lay.ntr;
	const long long item = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (item >= nitems) return;
	// consecutive threads -> consecutive nodes of an element
	const int n = (int)(item % NN);
	const long long ec = item / NN;
	const int c = (int)(ec % lay.ntr);
	const long long e = ec / lay.ntr;
	const size_t off0 = ((size_t)e * lay.nrows + lay.troff + c * L) * NN + n;
	double * p = data + off0;
	const double * a = area_node + (size_t)e * L * NN + n;
	double dTotalMass = 0.0;
	double dNonNegativeMass = 0.0;
	for (int k = 0; k < L; k++) {
		double v;
		if (combine) {
			const size_t off = off0 + (size_t)k * NN;
			v = 0.0;
			if (ca.scale_dst) {
				v = data[off] * ca.cdst;
			}
			for (int m = 0; m < ca.nsrc; m++) {
				v += ca.src[m][off] * ca.coeff[m];
			}
			p[(size_t)k * NN] = v;
		} else {
			v = p[(size_t)k * NN];
		}
		const double dPointwiseMass = v * a[(size_t)k * NN];
		dTotalMass += dPointwiseMass;
		if (v >= 0.0) {
			dNonNegativeMass += dPointwiseMass;
		}
	}
	const double dR = dTotalMass / dNonNegativeMass;
	for (int k = 0; k < L; k++) {
		const double v = p[(size_t)k * NN];
		const double vf = (v > 0.0) ? v * dR : 0.0;
		p[(size_t)k * NN] = vf;
		if (inc != 0) {
			double * q = inc + off0 + (size_t)k * NN;
			double w = q[0] * -1.0;
			w += vf * 1.0;
			q[0] = w;
		}
	}
}

#endif
