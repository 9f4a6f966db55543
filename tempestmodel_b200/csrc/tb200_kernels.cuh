// Element kernels of the horizontal dynamics (FP64, sm_100a).
//
// One thread per (level, node) of a spectral element; the np-term contractions
// with the 1-D derivative / stiffness matrices run over element tiles staged
// in shared memory.  These are HBM-bandwidth-bound stencils (about 2 flop/B),
// not GEMMs: the order-4 contractions are far too small for tensor cores.
//
// Every formula cites the reference statement it restates; the summation order
// inside each np-sum follows the reference so that results agree to rounding.
#ifndef TB200_KERNELS_CUH
#define TB200_KERNELS_CUH

#include "tb200_platform.h"
#include "tb200_device.h"

///////////////////////////////////////////////////////////////////////////////
// Column operator application: LinearColumnOperator::Apply
// (reference src/atm/LinearColumnOperator.h:82-100).

__device__ __forceinline__ double tb_col_apply(
	const DevOp & op, const double * col, int stride, int k
) {
	double o = 0.0;
	const int b = op.begin[k];
	const int e = op.end[k];
	const double * c = op.coeff + (size_t)k * op.width;
	for (int l = b; l < e; l++) {
		o += c[l - b] * col[(size_t)l * stride];
	}
	return o;
}

///////////////////////////////////////////////////////////////////////////////
// Host layout <-> device layout.
//
// Host: DataArray4D [c][iA][iB][k] with halo (reference DataArray4D.h:507-530).
// One block per element: the source is read with k fastest (coalesced runs of
// np*nlev doubles), transposed through shared memory, written node-fastest.

template <bool TO_DEVICE>
__global__ void k_transpose_state(
	DevLayout lay,
	double * dev,                 // device instance (element-major)
	double * host,                // staged copy of one host array
	long long elem0, int nea, int neb, int halo,
	int c_first, int c_count,     // host components handled
	int host_nlev,                // k extent of the host array
	int dev_row0_of_first,        // unused when rows are looked up per comp
	const int * rowmap            // [c_count] first device row of host comp, -1 = skip
) {
	TB_DYN_SMEM(double, tile);    // [host_nlev][nn + 1]
	const int np = lay.np;
	const int nn = lay.nn;
	const int e_local = blockIdx.x;
	const int a = e_local / neb;
	const int b = e_local % neb;
	const int wb = neb * np + 2 * halo;
	const int wa = nea * np + 2 * halo;
	const long long e = elem0 + e_local;
	const int nt = blockDim.x;

	for (int cc = 0; cc < c_count; cc++) {
		const int row0 = rowmap[cc];
		if (row0 < 0) continue;
		const int c = c_first + cc;
		const size_t hbase = (size_t)c * wa * wb * host_nlev;
		if (TO_DEVICE) {
			// read host: for node (i,j): host_nlev contiguous values
			for (int idx = threadIdx.x; idx < nn * host_nlev; idx += nt) {
				const int n = idx / host_nlev;
				const int k = idx % host_nlev;
				const int i = n / np, j = n % np;
				const size_t h = hbase
					+ ((size_t)(a * np + i + halo) * wb + (b * np + j + halo)) * host_nlev + k;
				tile[k * (nn + 1) + n] = host[h];
			}
			__syncthreads();
			for (int idx = threadIdx.x; idx < nn * host_nlev; idx += nt) {
				const int k = idx / nn;
				const int n = idx % nn;
				dev[((size_t)e * lay.nrows + row0 + k) * nn + n] = tile[k * (nn + 1) + n];
			}
			__syncthreads();
		} else {
			for (int idx = threadIdx.x; idx < nn * host_nlev; idx += nt) {
				const int k = idx / nn;
				const int n = idx % nn;
				tile[k * (nn + 1) + n] = dev[((size_t)e * lay.nrows + row0 + k) * nn + n];
			}
			__syncthreads();
			for (int idx = threadIdx.x; idx < nn * host_nlev; idx += nt) {
				const int n = idx / host_nlev;
				const int k = idx % host_nlev;
				const int i = n / np, j = n % np;
				const size_t h = hbase
					+ ((size_t)(a * np + i + halo) * wb + (b * np + j + halo)) * host_nlev + k;
				host[h] = tile[k * (nn + 1) + n];
			}
			__syncthreads();
		}
	}
}

// Geometry upload: host [iA][iB][k][m] (m = vector component, nm of them) ->
// device array m: [e][k][n].  dst[m] may be null.
struct GeomDst { double * p[3]; };

__global__ void k_transpose_geom(
	int np, int nn, long long elem0, int nea, int neb, int halo,
	const double * host, int nlev, int nm, GeomDst dst
) {
	const int e_local = blockIdx.x;
	const int a = e_local / neb;
	const int b = e_local % neb;
	const int wb = neb * np + 2 * halo;
	const long long e = elem0 + e_local;
	for (int idx = threadIdx.x; idx < nn * nlev * nm; idx += blockDim.x) {
		const int m = idx / (nn * nlev);
		const int r = idx % (nn * nlev);
		const int k = r / nn;
		const int n = r % nn;
		const int i = n / np, j = n % np;
		const size_t h =
			(((size_t)(a * np + i + halo) * wb + (b * np + j + halo)) * nlev + k) * nm + m;
		if (dst.p[m] != nullptr) {
			dst.p[m][((size_t)e * nlev + k) * nn + n] = host[h];
		}
	}
}

// Derived slots the reference keeps on the host: W on levels and U,V on
// interfaces (HorizontalDynamicsFEM.cpp:817-831), written into a staged host
// array in the reference layout.
__global__ void k_fill_derived(
	DevLayout lay, DevOps ops, const double * dev,
	double * host, long long elem0, int nea, int neb, int halo,
	int src_comp, int dst_comp, int op_id, int host_nlev
) {
	const int np = lay.np, nn = lay.nn;
	const int e_local = blockIdx.x;
	const int a = e_local / neb;
	const int b = e_local % neb;
	const int wb = neb * np + 2 * halo;
	const int wa = nea * np + 2 * halo;
	const long long e = elem0 + e_local;
	const double * col0 = dev + ((size_t)e * lay.nrows + lay.rowoff[src_comp]) * nn;
	for (int idx = threadIdx.x; idx < nn * host_nlev; idx += blockDim.x) {
		const int n = idx / host_nlev;
		const int k = idx % host_nlev;
		const int i = n / np, j = n % np;
		const double v = tb_col_apply(ops.op[op_id], col0 + n, nn, k);
		const size_t h = (size_t)dst_comp * wa * wb * host_nlev
			+ ((size_t)(a * np + i + halo) * wb + (b * np + j + halo)) * host_nlev + k;
		host[h] = v;
	}
}

///////////////////////////////////////////////////////////////////////////////
// Grid::LinearCombineData / CopyData / ZeroData on rows [row0,row1) of every
// element (reference GridPatch.cpp:1402-1553, DataArray4D.h:297-440).
// dest = c[dest]*dest (or 0), then += c[m]*inst[m] in ascending m.

#define TB_MAXINST 12

struct CombineArgs {
	const double * src[TB_MAXINST];
	double coeff[TB_MAXINST];
	int nsrc;
	double cdst;      // coefficient of the destination itself
	int scale_dst;    // 0: dest starts at 0; 1: dest *= cdst
};

__global__ void k_lincomb(
	DevLayout lay, CombineArgs ca, double * dst, int row0, int row1
) {
	const int nn = lay.nn;
	const long long per_e = (long long)(row1 - row0) * nn;
	const long long total = lay.nelem * per_e;
	for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	     idx < total; idx += (long long)gridDim.x * blockDim.x
	) {
		const long long e = idx / per_e;
		const long long r = idx % per_e;
		const size_t off = ((size_t)e * lay.nrows + row0) * nn + r;
		double v = 0.0;
		if (ca.scale_dst) {
			v = dst[off] * ca.cdst;
		}
		for (int m = 0; m < ca.nsrc; m++) {
			v += ca.src[m][off] * ca.coeff[m];
		}
		dst[off] = v;
	}
}

///////////////////////////////////////////////////////////////////////////////
// HorizontalDynamicsFEM::StepShallowWater
// (reference HorizontalDynamicsFEM.cpp:321-647).
// A block holds ITEMS (element, level) pairs; thread = (item, node).

template <int NP, int ITEMS>
__global__ void __launch_bounds__(NP * NP * ITEMS)
k_sw_explicit(
	DevLayout lay, DevGeom g, DevTables t,
	const double * __restrict__ in, double * __restrict__ out,
	double dt, double grav
) {
	const int NN = NP * NP;
	__shared__ double sUa[ITEMS][NN];
	__shared__ double sUb[ITEMS][NN];
	__shared__ double sKE[ITEMS][NN];
	__shared__ double sFa[ITEMS][NN];
	__shared__ double sFb[ITEMS][NN];

	const int L = lay.nlev;
	const int it = threadIdx.x / NN;
	const int n = threadIdx.x % NN;
	const int i = n / NP, j = n % NP;
	const long long nitems = lay.nelem * L;
	long long item = (long long)blockIdx.x * ITEMS + it;
	const bool active = (item < nitems);
	if (!active) item = nitems - 1;
	const long long e = item / L;
	const int k = (int)(item % L);

	const size_t ebase = (size_t)e * lay.nrows * NN;
	const size_t g2 = (size_t)e * NN + n;
	const double dInvDA = g.inv_da[e];
	const double dInvDB = g.inv_db[e];

	const double dCovUa = in[ebase + (size_t)(lay.rowoff[0] + k) * NN + n];
	const double dCovUb = in[ebase + (size_t)(lay.rowoff[1] + k) * NN + n];
	const double dH = in[ebase + (size_t)(lay.rowoff[2] + k) * NN + n];
	const double dJ2D = g.j2d[g2];

	// Contravariant velocities (:432-438)
	const double dConUa = g.a0[g2] * dCovUa + g.a1[g2] * dCovUb;
	const double dConUb = g.b0[g2] * dCovUa + g.b1[g2] * dCovUb;
	// Specific kinetic energy plus pointwise pressure (:441-446)
	double dKE = 0.5 * (dConUa * dCovUa + dConUb * dCovUb);
	dKE += grav * dH;
	// Base fluxes and height flux (:459-479)
	const double dAlphaBaseFlux = dJ2D * dConUa;
	const double dBetaBaseFlux = dJ2D * dConUb;
	const double dZs = g.zs[g2];

	sUa[it][n] = dCovUa;
	sUb[it][n] = dCovUb;
	sKE[it][n] = dKE;
	sFa[it][n] = dAlphaBaseFlux * (dH - dZs);
	sFb[it][n] = dBetaBaseFlux * (dH - dZs);
	__syncthreads();

	// np-sums (:520-569)
	double dDaMassFluxA = 0.0, dCovDaUb = 0.0, dDaKE = 0.0;
	double dDbMassFluxB = 0.0, dCovDbUa = 0.0, dDbKE = 0.0;
#pragma unroll
	for (int s = 0; s < NP; s++) {
		dDaMassFluxA -= sFa[it][s * NP + j] * t.st[i * NP + s];
		dCovDaUb += sUb[it][s * NP + j] * t.dx[s * NP + i];
		dDaKE += sKE[it][s * NP + j] * t.dx[s * NP + i];
	}
#pragma unroll
	for (int s = 0; s < NP; s++) {
		dDbMassFluxB -= sFb[it][i * NP + s] * t.st[j * NP + s];
		dCovDbUa += sUa[it][i * NP + s] * t.dx[s * NP + j];
		dDbKE += sKE[it][i * NP + s] * t.dx[s * NP + j];
	}
	dDaMassFluxA *= dInvDA;
	dCovDaUb *= dInvDA;
	dDaKE *= dInvDA;
	dDbMassFluxB *= dInvDB;
	dCovDbUa *= dInvDB;
	dDbKE *= dInvDB;

	// Momentum update (:572-608)
	double dLocalUpdateUa = 0.0;
	double dLocalUpdateUb = 0.0;
	const double dZetaXi = (dCovDaUb - dCovDbUa);
	const double dCovUCrossZetaA = dConUb * dZetaXi;
	const double dCovUCrossZetaB = -dConUa * dZetaXi;
	const double dF = g.f[g2];
	dLocalUpdateUa += dF * dJ2D * dConUb;
	dLocalUpdateUb -= dF * dJ2D * dConUa;
	dLocalUpdateUa += -dDaKE + dCovUCrossZetaA;
	dLocalUpdateUb += -dDbKE + dCovUCrossZetaB;

	const double dInvJacobian2D = 1.0 / dJ2D;
	if (active) {
		const size_t oU = ebase + (size_t)(lay.rowoff[0] + k) * NN + n;
		const size_t oV = ebase + (size_t)(lay.rowoff[1] + k) * NN + n;
		const size_t oH = ebase + (size_t)(lay.rowoff[2] + k) * NN + n;
		out[oU] += dt * dLocalUpdateUa;
		out[oV] += dt * dLocalUpdateUb;
		// Height update (:611-616)
		out[oH] -= dt * dInvJacobian2D * (dDaMassFluxA + dDbMassFluxB);
	}

	// Tracers (:619-640)
	for (int c = 0; c < lay.ntr; c++) {
		__syncthreads();
		const size_t oT = ebase + (size_t)(lay.troff + c * L + k) * NN + n;
		const double q = in[oT];
		sFa[it][n] = dAlphaBaseFlux * q;
		sFb[it][n] = dBetaBaseFlux * q;
		__syncthreads();
		double dDaTracerFluxA = 0.0, dDbTracerFluxB = 0.0;
#pragma unroll
		for (int s = 0; s < NP; s++) {
			dDaTracerFluxA -= sFa[it][s * NP + j] * t.st[i * NP + s];
			dDbTracerFluxB -= sFb[it][i * NP + s] * t.st[j * NP + s];
		}
		dDaTracerFluxA *= dInvDA;
		dDbTracerFluxB *= dInvDB;
		if (active) {
			out[oT] -= dt * dInvJacobian2D * (dDaTracerFluxA + dDbTracerFluxB);
		}
	}
}

///////////////////////////////////////////////////////////////////////////////
// HorizontalDynamicsFEM::StepNonhydrostaticPrimitive
// (reference HorizontalDynamicsFEM.cpp:701-1783, FORMULATION_RHOTHETA_PI,
// Lorenz staggering) fused - when DO_V - with VerticalDynamicsFEM::StepExplicit
// (reference VerticalDynamicsFEM.cpp:616-1159, implicit-vertical branch:
// xi-dot on interfaces :816-828 and upwind penalty on U,V :998-1023).
//
// One block per element.  The element's U, V (levels) and W (interfaces)
// columns are staged in shared memory once; levels are processed KB at a time.

struct NHArgs {
	double dt;
	int xz;             // GridGLL::GetIsCartesianXZ
	int fe_nodes;       // nodes per vertical finite element (vertorder; 1 for FV)
};

// Base of the stage result: the current content of the update instance, or a
// linear combination of instances formed on the fly (Grid::CopyData /
// LinearCombineData fused into the stage; same operation order as k_lincomb).
struct StageBase {
	const double * src[TB_MAXINST];
	double coeff[TB_MAXINST];
	int nsrc;
	double cdst;
	int scale_dst;
	int use_out;        // 1: base = out as it is
};

__device__ __forceinline__ double tb_stage_base(
	const StageBase & sb, const double * out, size_t off
) {
	if (sb.use_out) return out[off];
	double v = 0.0;
	if (sb.scale_dst) v = out[off] * sb.cdst;
	for (int m = 0; m < sb.nsrc; m++) {
		v += sb.src[m][off] * sb.coeff[m];
	}
	return v;
}

// ---- terrain-following metric on the fly --------------------------------------
// GridPatchCSGLL::EvaluateGeometricTerms (GridPatchCSGLL.cpp:344-553)

__device__ __forceinline__ ColMetric tb_col_metric(const DevGeom & g, size_t g2) {
	ColMetric c;
	c.a0 = g.a0[g2]; c.a1 = g.a1[g2]; c.b0 = g.b0[g2]; c.b1 = g.b1[g2];
	c.j2d = g.j2d[g2];
	c.onepx2 = 1.0; c.onepy2 = 1.0; c.xy = 0.0; c.msod = 0.0;
	c.inv_dxr = 0.0; c.inv_dxr2 = 0.0; c.dxr = 0.0; c.dazs = 0.0; c.dbzs = 0.0;
	if (g.analytic) {
		const double X = g.tx[g2], Y = g.ty[g2];
		const double xx = __dmul_rn(X, X), yy = __dmul_rn(Y, Y);
		const double d2 = __dadd_rn(__dadd_rn(1.0, xx), yy);
		c.onepx2 = __dadd_rn(1.0, xx);
		c.onepy2 = __dadd_rn(1.0, yy);
		c.xy = __dmul_rn(X, Y);
		const double aa = __dmul_rn(g.radius, g.radius);
		const double scale = __ddiv_rn(__ddiv_rn(__ddiv_rn(d2, c.onepx2), c.onepy2), aa);
		c.dxr = __dadd_rn(g.ztop, -g.zs[g2]);
		c.msod = __ddiv_rn(-scale, c.dxr);
		c.inv_dxr = __ddiv_rn(1.0, c.dxr);
		c.inv_dxr2 = __ddiv_rn(1.0, __dmul_rn(c.dxr, c.dxr));
		c.dazs = g.tda[g2];
		c.dbzs = g.tdb[g2];
	}
	return c;
}

__device__ __forceinline__ LevMetric tb_lev_metric(const ColMetric & c, double eta) {
	LevMetric m;
	const double w = __dadd_rn(1.0, -eta);
	m.dar = __dmul_rn(w, c.dazs);
	m.dbr = __dmul_rn(w, c.dbzs);
	m.jac = __dmul_rn(c.dxr, c.j2d);
	m.a2 = __dmul_rn(c.msod, __dadd_rn(__dmul_rn(c.onepy2, m.dar), __dmul_rn(c.xy, m.dbr)));
	m.b2 = __dmul_rn(c.msod, __dadd_rn(__dmul_rn(c.xy, m.dar), __dmul_rn(c.onepx2, m.dbr)));
	m.x2 = __dadd_rn(c.inv_dxr2,
		-__dmul_rn(c.inv_dxr, __dadd_rn(__dmul_rn(m.a2, m.dar), __dmul_rn(m.b2, m.dbr))));
	return m;
}

// shared-memory doubles needed by k_nh_explicit
__host__ __device__ inline size_t tb_nh_smem_doubles(int L, int nn, int kb) {
	return (size_t)nn * (3 * L + (L + 1) + 3 * L) + (size_t)kb * 7 * nn;
}

template <int NP, bool DO_H, bool DO_V>
__global__ void k_nh_explicit(
	DevLayout lay, DevGeom g, DevTables t, DevOps ops, DevPhys ph, NHArgs args,
	StageBase sb, const double * __restrict__ in, double * out, int KB
) {
	const int NN = NP * NP;
	const int UIx = 0, VIx = 1, PIx = 2, WIx = 3, RIx = 4;
	const int L = lay.nlev;
	const double dt = args.dt;

	TB_DYN_SMEM(double, sm);
	double * sU = sm;                    // [L][NN]   covariant u_alpha (initial)
	double * sV = sU + L * NN;           // [L][NN]
	double * sW = sV + L * NN;           // [L+1][NN] covariant w on interfaces
	double * sZX = sW + (L + 1) * NN;    // [L][NN]   (u x zeta)_xi
	double * sUn = sZX + L * NN;         // [L][NN]   U after the horizontal update
	double * sVn = sUn + L * NN;         // [L][NN]
	double * tile = sVn + L * NN;        // [KB][7][NN]

	const long long e = blockIdx.x;
	const int kk = threadIdx.x / NN;
	const int n = threadIdx.x % NN;
	const int i = n / NP, j = n % NP;
	const int nt = blockDim.x;

	const size_t ebase = (size_t)e * lay.nrows * NN;
	const size_t offU = ebase + (size_t)lay.rowoff[UIx] * NN;
	const size_t offV = ebase + (size_t)lay.rowoff[VIx] * NN;
	const size_t offP = ebase + (size_t)lay.rowoff[PIx] * NN;
	const size_t offW = ebase + (size_t)lay.rowoff[WIx] * NN;
	const size_t offR = ebase + (size_t)lay.rowoff[RIx] * NN;
	const double * inU = in + offU;
	const double * inV = in + offV;
	const double * inP = in + offP;
	const double * inW = in + offW;
	const double * inR = in + offR;

	for (int idx = threadIdx.x; idx < L * NN; idx += nt) {
		sU[idx] = inU[idx];
		sV[idx] = inV[idx];
	}
	for (int idx = threadIdx.x; idx < (L + 1) * NN; idx += nt) {
		sW[idx] = inW[idx];
	}
	__syncthreads();

	const double dInvDA = g.inv_da[e];
	const double dInvDB = g.inv_db[e];
	const size_t g2 = (size_t)e * NN + n;
	const size_t g3 = (size_t)e * L * NN;
	const size_t g3e = (size_t)e * (L + 1) * NN;
	const ColMetric cm = tb_col_metric(g, g2);

	// xi-row of the contravariant metric on interface m of this thread's column
	auto cxe_at = [&](int m, double & c0, double & c1, double & c2) {
		if (g.analytic) {
			const LevMetric lm = tb_lev_metric(cm, g.reta_e[m]);
			c0 = lm.a2; c1 = lm.b2; c2 = lm.x2;
		} else {
			const size_t om = g3e + (size_t)m * NN + n;
			c0 = g.cxe[0][om]; c1 = g.cxe[1][om]; c2 = g.cxe[2][om];
		}
	};

	double * tWn = tile + (size_t)kk * 7 * NN;  // covariant w on levels
	double * tKE = tWn + NN;
	double * tEX = tKE + NN;
	double * tFaR = tEX + NN;
	double * tFbR = tFaR + NN;
	double * tFaP = tFbR + NN;
	double * tFbP = tFaP + NN;

	for (int k0 = 0; k0 < L; k0 += KB) {
		const int k = k0 + kk;
		const bool active = (k < L);
		const int kc = active ? k : (L - 1);
		const size_t o = (size_t)kc * NN + n;   // offset inside a component

		double dCovUa = 0.0, dCovUb = 0.0, dCovUx = 0.0;
		double dConUa = 0.0, dConUb = 0.0, dConUx = 0.0;
		double dJac = 1.0, dRho = 1.0, dRhoTheta = 1.0;
		double dAlphaBaseFlux = 0.0, dBetaBaseFlux = 0.0;
		double dDerivR0 = 0.0, dDerivR1 = 0.0;

		if (DO_H) {
			dCovUa = sU[o];
			dCovUb = sV[o];
			// InterpolateREdgeToNode(W) (:817-819, GridPatchGLL.cpp:109-143)
			dCovUx = tb_col_apply(ops.op[1], sW + n, NN, kc);

			double m0, m1, m2, m3, m4, m5;
			if (g.analytic) {
				const LevMetric lm = tb_lev_metric(cm, g.reta_n[kc]);
				dJac = lm.jac;
				m0 = cm.a0; m1 = cm.a1; m2 = lm.a2;
				m3 = cm.b1; m4 = lm.b2; m5 = lm.x2;
				dDerivR0 = lm.dar; dDerivR1 = lm.dbr;
			} else {
				dJac = g.jac[g3 + o];
				m0 = g.ca[0][g3 + o];
				m1 = g.ca[1][g3 + o];
				m2 = g.ca[2][g3 + o];
				m3 = g.cb[1][g3 + o];
				m4 = g.cb[2][g3 + o];
				m5 = g.cx[2][g3 + o];
				dDerivR0 = g.dr[0][g3 + o];
				dDerivR1 = g.dr[1][g3 + o];
			}

			// Contravariant velocities (:916-929)
			dConUa = m0 * dCovUa + m1 * dCovUb + m2 * dCovUx;
			dConUb = m1 * dCovUa + m3 * dCovUb + m4 * dCovUx;
			dConUx = m2 * dCovUa + m4 * dCovUb + m5 * dCovUx;

			dRho = inR[o];
			dRhoTheta = inP[o];

			tWn[n] = dCovUx;
			// Specific kinetic energy (:932-935)
			tKE[n] = 0.5 * (dConUa * dCovUa + dConUb * dCovUb + dConUx * dCovUx);
			// Exner pressure (:949-951, PhysicalConstants.h:397-399)
			tEX[n] = ph.cp * exp(ph.exner_c1 * log(ph.exner_c2 * dRhoTheta));
			// Fluxes (:1050-1077)
			dAlphaBaseFlux = dJac * dConUa;
			dBetaBaseFlux = dJac * dConUb;
			tFaR[n] = dAlphaBaseFlux * dRho;
			tFbR[n] = dBetaBaseFlux * dRho;
			tFaP[n] = dAlphaBaseFlux * dRhoTheta;
			tFbP[n] = dBetaBaseFlux * dRhoTheta;
		}
		__syncthreads();

		double uNew = 0.0, vNew = 0.0;
		double dInvJacobian = 1.0;

		if (DO_H) {
			// U cross relative vorticity (:966-1039)
			const double dCovDxUa = tb_col_apply(ops.op[2], sU + n, NN, kc);
			const double dCovDxUb = tb_col_apply(ops.op[2], sV + n, NN, kc);
			double dCovDaUb = 0.0, dCovDaUx = 0.0, dCovDbUa = 0.0, dCovDbUx = 0.0;
#pragma unroll
			for (int s = 0; s < NP; s++) {
				dCovDaUb += sV[(size_t)kc * NN + s * NP + j] * t.dx[s * NP + i];
				dCovDaUx += tWn[s * NP + j] * t.dx[s * NP + i];
				dCovDbUa += sU[(size_t)kc * NN + i * NP + s] * t.dx[s * NP + j];
				dCovDbUx += tWn[i * NP + s] * t.dx[s * NP + j];
			}
			dCovDaUb *= dInvDA;
			dCovDaUx *= dInvDA;
			dCovDbUa *= dInvDB;
			dCovDbUx *= dInvDB;

			const double dJZetaA = (dCovDbUx - dCovDxUb);
			const double dJZetaB = (dCovDxUa - dCovDaUx);
			const double dJZetaX = (dCovDaUb - dCovDbUa);

			const double dUCrossZetaA = dConUb * dJZetaX - dConUx * dJZetaB;
			const double dUCrossZetaB = dConUx * dJZetaA - dConUa * dJZetaX;
			const double dUCrossZetaX = -dConUa * dCovDaUx - dConUb * dCovDbUx;

			// Pointwise update (:1183-1421)
			dInvJacobian = 1.0 / dJac;
			double dDaKE = 0.0, dDbKE = 0.0, dDaP = 0.0, dDbP = 0.0;
			double dDaRhoFluxA = 0.0, dDaPressureFluxA = 0.0;
			double dDbRhoFluxB = 0.0, dDbPressureFluxB = 0.0;
#pragma unroll
			for (int s = 0; s < NP; s++) {
				dDaRhoFluxA -= tFaR[s * NP + j] * t.st[i * NP + s];
				dDaPressureFluxA -= tFaP[s * NP + j] * t.st[i * NP + s];
				dDaP += tEX[s * NP + j] * t.dx[s * NP + i];
				dDaKE += tKE[s * NP + j] * t.dx[s * NP + i];
			}
#pragma unroll
			for (int s = 0; s < NP; s++) {
				dDbRhoFluxB -= tFbR[i * NP + s] * t.st[j * NP + s];
				dDbPressureFluxB -= tFbP[i * NP + s] * t.st[j * NP + s];
				dDbP += tEX[i * NP + s] * t.dx[s * NP + j];
				dDbKE += tKE[i * NP + s] * t.dx[s * NP + j];
			}
			dDaRhoFluxA *= dInvDA;
			dDbRhoFluxB *= dInvDB;
			dDaPressureFluxA *= dInvDA;
			dDbPressureFluxB *= dInvDB;
			dDaP *= dInvDA;
			dDbP *= dInvDB;
			dDaKE *= dInvDA;
			dDbKE *= dInvDB;

			double dLocalUpdateUa = 0.0;
			double dLocalUpdateUb = 0.0;
			dLocalUpdateUa += dUCrossZetaA;
			dLocalUpdateUb += dUCrossZetaB;

			// Coriolis (:1330-1338)
			const double dF = g.f[g2];
			const double dJ2D = cm.j2d;
			dLocalUpdateUa += dF * dJ2D * dConUb;
			dLocalUpdateUb -= dF * dJ2D * dConUa;

			// Pressure gradient force, RHOTHETA_PI (:1348-1353)
			const double dPressureGradientForceUa = dDaP * dRhoTheta / dRho;
			const double dPressureGradientForceUb = dDbP * dRhoTheta / dRho;

			// Gravity (:1363-1364)
			const double dDaPhi = ph.g * dDerivR0;
			const double dDbPhi = ph.g * dDerivR1;

			const double dDaUpdate = dPressureGradientForceUa + dDaKE + dDaPhi;
			const double dDbUpdate = dPressureGradientForceUb + dDbKE + dDbPhi;
			dLocalUpdateUa -= dDaUpdate;
			dLocalUpdateUb -= dDbUpdate;

			if (active) {
				uNew = tb_stage_base(sb, out, offU + o) + dt * dLocalUpdateUa;
				vNew = tb_stage_base(sb, out, offV + o);
				if (!args.xz) {
					vNew += dt * dLocalUpdateUb;
				}
				sUn[o] = uNew;
				sVn[o] = vNew;
				sZX[o] = dUCrossZetaX;
				// Density and rho-theta (:1399-1421)
				out[offR + o] = tb_stage_base(sb, out, offR + o)
					- dt * dInvJacobian * (dDaRhoFluxA + dDbRhoFluxB);
				out[offP + o] = tb_stage_base(sb, out, offP + o)
					- dt * dInvJacobian * (dDaPressureFluxA + dDbPressureFluxB);
			}
		} else if (active) {
			uNew = tb_stage_base(sb, out, offU + o);
			vNew = tb_stage_base(sb, out, offV + o);
		}

		if (DO_V && active) {
			// Upwind penalty on U and V (VerticalDynamicsFEM.cpp:998-1023,
			// LinearColumnOperatorFEM.cpp:1863-1887)
			const int vo = args.fe_nodes;
			const int nfe = L / vo;
			const int a = k / vo;
			if (a <= nfe - 2) {
				const int m = (a + 1) * vo;
				const size_t om = (size_t)m * NN + n;
				const double ue = tb_col_apply(ops.op[0], sU + n, NN, m);
				const double ve = tb_col_apply(ops.op[0], sV + n, NN, m);
				double c0, c1, c2;
				cxe_at(m, c0, c1, c2);
				const double xd = c0 * ue + c1 * ve + c2 * sW[om];
				const double w = dt * fabs(xd);
				uNew += tb_col_apply(ops.op[8], sU + n, NN, k) * w;
				vNew += tb_col_apply(ops.op[8], sV + n, NN, k) * w;
			}
			if (a >= 1) {
				const int m = a * vo;
				const size_t om = (size_t)m * NN + n;
				const double ue = tb_col_apply(ops.op[0], sU + n, NN, m);
				const double ve = tb_col_apply(ops.op[0], sV + n, NN, m);
				double c0, c1, c2;
				cxe_at(m, c0, c1, c2);
				const double xd = c0 * ue + c1 * ve + c2 * sW[om];
				const double w = dt * fabs(xd);
				uNew += tb_col_apply(ops.op[9], sU + n, NN, k) * w;
				vNew += tb_col_apply(ops.op[9], sV + n, NN, k) * w;
			}
		}
		if (active) {
			out[offU + o] = uNew;
			out[offV + o] = vNew;
		}

		// Tracers (:1531-1553)
		if (DO_H) {
			for (int c = 0; c < lay.ntr; c++) {
				__syncthreads();
				const size_t oT = ebase + (size_t)(lay.troff + c * L + kc) * NN + n;
				const double q = in[oT];
				tFaR[n] = dAlphaBaseFlux * q;
				tFbR[n] = dBetaBaseFlux * q;
				__syncthreads();
				double dDaTracerFluxA = 0.0, dDbTracerFluxB = 0.0;
#pragma unroll
				for (int s = 0; s < NP; s++) {
					dDaTracerFluxA -= tFaR[s * NP + j] * t.st[i * NP + s];
					dDbTracerFluxB -= tFbR[i * NP + s] * t.st[j * NP + s];
				}
				dDaTracerFluxA *= dInvDA;
				dDbTracerFluxB *= dInvDB;
				if (active) {
					out[oT] = tb_stage_base(sb, out, oT)
						- dt * dInvJacobian * (dDaTracerFluxA + dDbTracerFluxB);
				}
			}
		} else if (!sb.use_out) {
			for (int c = 0; c < lay.ntr; c++) {
				const size_t oT = ebase + (size_t)(lay.troff + c * L + kc) * NN + n;
				if (active) out[oT] = tb_stage_base(sb, out, oT);
			}
		}
		__syncthreads();
	}

	// Vertical velocity on interfaces (:1612-1660)
	if (DO_H) {
		for (int idx = threadIdx.x; idx < (L + 1) * NN; idx += nt) {
			const int k = idx / NN;
			const int nc = idx % NN;
			if (k == 0) {
				const double dU0 = tb_col_apply(ops.op[0], sUn + nc, NN, 0);
				const double dV0 = tb_col_apply(ops.op[0], sVn + nc, NN, 0);
				double c0, c1, c2;
				if (g.analytic) {
					const ColMetric cmc = tb_col_metric(g, (size_t)e * NN + nc);
					const LevMetric lm = tb_lev_metric(cmc, g.reta_e[0]);
					c0 = lm.a2; c1 = lm.b2; c2 = lm.x2;
				} else {
					const size_t oe = g3e + nc;
					c0 = g.cxe[0][oe]; c1 = g.cxe[1][oe]; c2 = g.cxe[2][oe];
				}
				out[offW + nc] = -(c0 * dU0 + c1 * dV0) / c2;
			} else if (k < L) {
				const double dUCrossZetaX = tb_col_apply(ops.op[0], sZX + nc, NN, k);
				out[offW + idx] = tb_stage_base(sb, out, offW + idx) + dt * dUCrossZetaX;
			} else if (!sb.use_out) {
				out[offW + idx] = tb_stage_base(sb, out, offW + idx);
			}
		}
	} else if (!sb.use_out) {
		for (int idx = threadIdx.x; idx < (L + 1) * NN; idx += nt) {
			out[offW + idx] = tb_stage_base(sb, out, offW + idx);
		}
		for (int idx = threadIdx.x; idx < L * NN; idx += nt) {
			out[offP + idx] = tb_stage_base(sb, out, offP + idx);
			out[offR + idx] = tb_stage_base(sb, out, offR + idx);
		}
	}
}

///////////////////////////////////////////////////////////////////////////////
// Uniform diffusion of u and v in the column (VerticalDynamicsFEM::StepExplicit,
// VerticalDynamicsFEM.cpp:1058-1106): second derivative on levels of the velocity and
// of the reference velocity, update += dt coeff (DD u - DD u_ref), coeff = nu / ztop^2.
// One thread per element-local node.  stale != 0: the "state" column is that one
// column [2][L] for every node (what the reference's work array holds outside
// --explicitvertical, see uniform_diffusion_vertical_uv in tb200_api.cu).
__global__ void k_stale_column_uv(
	DevLayout lay, const double * __restrict__ in, long long node, double * __restrict__ stale
) {
	const int NN = lay.nn;
	const int L = lay.nlev;
	const size_t ebase = (size_t)(node / NN) * lay.nrows * NN + (size_t)(node % NN);
	for (int q = threadIdx.x; q < 2 * L; q += blockDim.x) {
		const int c = q / L, k = q % L;
		stale[q] = in[ebase + (size_t)(lay.rowoff[c] + k) * NN];
	}
}

__global__ void k_uniform_diffusion_uv(
	DevLayout lay, DevOps ops, const double * __restrict__ in,
	const double * __restrict__ ref, double * __restrict__ out,
	double dt, double coeff, const double * __restrict__ stale
) {
	const int NN = lay.nn;
	const int L = lay.nlev;
	const long long node = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (node >= lay.nelem * (long long)NN) return;
	const long long e = node / NN;
	const int n = (int)(node % NN);
	const size_t ebase = (size_t)e * lay.nrows * NN + n;
	const DevOp & opDDN2N = ops.op[6];
	for (int c = 0; c < 2; c++) {
		const size_t o = ebase + (size_t)lay.rowoff[c] * NN;
		for (int k = 0; k < L; k++) {
			const double dDiffDiffState = (stale != 0)
				? tb_col_apply(opDDN2N, stale + (size_t)c * L, 1, k)
				: tb_col_apply(opDDN2N, in + o, NN, k);
			const double dDiffDiffRef = tb_col_apply(opDDN2N, ref + o, NN, k);
			out[o + (size_t)k * NN] += dt * coeff * (dDiffDiffState - dDiffDiffRef);
		}
	}
}

///////////////////////////////////////////////////////////////////////////////
// HorizontalDynamicsFEM::ApplyScalarHyperdiffusion
// (reference HorizontalDynamicsFEM.cpp:1867-2203) on rows [row0,row1) of every
// element, each row being one (component, level) slab; jac_sel picks the
// level or interface Jacobian for the row.

struct HyperRows {
	int nranges;
	int row0[TB_MAXC + 1];
	int row1[TB_MAXC + 1];
	int onedge[TB_MAXC + 1];
};

template <int NP, int ITEMS>
__global__ void __launch_bounds__(NP * NP * ITEMS)
k_hyper_scalar(
	DevLayout lay, DevGeom g, DevTables t, HyperRows hr, int nrows_sel,
	const double * __restrict__ in, double * __restrict__ out,
	double dt, double nu, int scale_nu,
	const double * __restrict__ ref   // fRemoveRefState (:2060-2070): subtracted from the field; 0 = no
) {
	const int NN = NP * NP;
	__shared__ double sPsi[ITEMS][NN];
	__shared__ double sGa[ITEMS][NN];
	__shared__ double sGb[ITEMS][NN];

	const int L = lay.nlev;
	const int it = threadIdx.x / NN;
	const int n = threadIdx.x % NN;
	const int i = n / NP, j = n % NP;
	const long long nitems = lay.nelem * nrows_sel;
	long long item = (long long)blockIdx.x * ITEMS + it;
	const bool active = (item < nitems);
	if (!active) item = nitems - 1;
	const long long e = item / nrows_sel;
	int rs = (int)(item % nrows_sel);
	// locate the row
	int row = 0, klev = 0, onedge = 0;
	for (int q = 0; q < hr.nranges; q++) {
		const int len = hr.row1[q] - hr.row0[q];
		if (rs < len) {
			row = hr.row0[q] + rs;
			onedge = hr.onedge[q];
			klev = rs % (onedge ? (L + 1) : L);
			break;
		}
		rs -= len;
	}

	const size_t off = ((size_t)e * lay.nrows + row) * NN + n;
	const size_t g2 = (size_t)e * NN + n;
	const double dInvDA = g.inv_da[e];
	const double dInvDB = g.inv_db[e];
	const double dJac = onedge
		? g.jace[((size_t)e * (L + 1) + klev) * NN + n]
		: g.jac[((size_t)e * L + klev) * NN + n];

	double dBufferState = in[off];
	if (ref != 0) {
		dBufferState -= ref[off];
	}
	sPsi[it][n] = dBufferState;
	__syncthreads();

	// Pointwise gradient (:2073-2107)
	double dDaPsi = 0.0, dDbPsi = 0.0;
#pragma unroll
	for (int s = 0; s < NP; s++) {
		dDaPsi += sPsi[it][s * NP + j] * t.dx[s * NP + i];
		dDbPsi += sPsi[it][i * NP + s] * t.dx[s * NP + j];
	}
	dDaPsi *= dInvDA;
	dDbPsi *= dInvDB;
	sGa[it][n] = dJac * (g.a0[g2] * dDaPsi + g.a1[g2] * dDbPsi);
	sGb[it][n] = dJac * (g.b0[g2] * dDaPsi + g.b1[g2] * dDbPsi);
	__syncthreads();

	// Integral term (:2126-2165)
	const double dInvJacobian = 1.0 / dJac;
	double dUpdateA = 0.0, dUpdateB = 0.0;
#pragma unroll
	for (int s = 0; s < NP; s++) {
		dUpdateA += sGa[it][s * NP + j] * t.st[i * NP + s];
		dUpdateB += sGb[it][i * NP + s] * t.st[j * NP + s];
	}
	dUpdateA *= dInvDA;
	dUpdateB *= dInvDB;

	double dLocalNu = nu;
	if (scale_nu) {
		dLocalNu *= g.nu_scale[e];
	}
	if (active) {
		out[off] -= dt * dInvJacobian * dLocalNu * (dUpdateA + dUpdateB);
	}
}

///////////////////////////////////////////////////////////////////////////////
// HorizontalDynamicsFEM::ApplyVectorHyperdiffusion with
// GridPatchCSGLL::ComputeCurlAndDiv inlined (reference
// HorizontalDynamicsFEM.cpp:2207-2414, GridPatchCSGLL.cpp:1132-1305).

template <int NP, int ITEMS>
__global__ void __launch_bounds__(NP * NP * ITEMS)
k_hyper_vector(
	DevLayout lay, DevGeom g, DevTables t,
	const double * __restrict__ in, double * __restrict__ out,
	double dt, double nu_div, double nu_vort, int scale_nu, int xz
) {
	const int NN = NP * NP;
	__shared__ double sUa[ITEMS][NN];
	__shared__ double sUb[ITEMS][NN];
	__shared__ double sJ[ITEMS][NN];     // 2-D Jacobian
	__shared__ double sCa[ITEMS][NN];    // contravariant u^alpha
	__shared__ double sCb[ITEMS][NN];    // contravariant u^beta
	__shared__ double sDiv[ITEMS][NN];
	__shared__ double sCurl[ITEMS][NN];

	const int L = lay.nlev;
	const int it = threadIdx.x / NN;
	const int n = threadIdx.x % NN;
	const int i = n / NP, j = n % NP;
	const long long nitems = lay.nelem * L;
	long long item = (long long)blockIdx.x * ITEMS + it;
	const bool active = (item < nitems);
	if (!active) item = nitems - 1;
	const long long e = item / L;
	const int k = (int)(item % L);

	const size_t ebase = (size_t)e * lay.nrows * NN;
	const size_t oU = ebase + (size_t)(lay.rowoff[0] + k) * NN + n;
	const size_t oV = ebase + (size_t)(lay.rowoff[1] + k) * NN + n;
	const size_t g2 = (size_t)e * NN + n;
	const double dInvDA = g.inv_da[e];
	const double dInvDB = g.inv_db[e];

	const double dUa = in[oU];
	const double dUb = in[oV];
	const double dJ2D = g.j2d[g2];
	const double a0 = g.a0[g2], a1 = g.a1[g2], b0 = g.b0[g2], b1 = g.b1[g2];

	// Contravariant velocities (GridPatchCSGLL.cpp:1207-1218)
	sUa[it][n] = dUa;
	sUb[it][n] = dUb;
	sJ[it][n] = dJ2D;
	sCa[it][n] = +a0 * dUa + a1 * dUb;
	sCb[it][n] = +b0 * dUa + b1 * dUb;
	__syncthreads();

	// Curl and divergence (GridPatchCSGLL.cpp:1262-1299); the products
	// J2D(s) * conU(s) * Dx follow the reference's left-to-right order.
	double dDaUb = 0.0, dDbUa = 0.0, dDaJUa = 0.0, dDbJUb = 0.0;
#pragma unroll
	for (int s = 0; s < NP; s++) {
		dDaUb += sUb[it][s * NP + j] * t.dx[s * NP + i];
		dDbUa += sUa[it][i * NP + s] * t.dx[s * NP + j];
		dDaJUa += sJ[it][s * NP + j] * sCa[it][s * NP + j] * t.dx[s * NP + i];
		dDbJUb += sJ[it][i * NP + s] * sCb[it][i * NP + s] * t.dx[s * NP + j];
	}
	dDaUb *= dInvDA;
	dDbUa *= dInvDB;
	dDaJUa *= dInvDA;
	dDbJUb *= dInvDB;
	const double dInvJacobian2D = 1.0 / dJ2D;
	sDiv[it][n] = (dDaJUa + dDbJUb) * dInvJacobian2D;
	sCurl[it][n] = (dDaUb - dDbUa) * dInvJacobian2D;
	__syncthreads();

	// Hyperviscosity sums (:2366-2407)
	double dDaDiv = 0.0, dDbDiv = 0.0, dDaCurl = 0.0, dDbCurl = 0.0;
#pragma unroll
	for (int s = 0; s < NP; s++) {
		dDaDiv -= t.st[i * NP + s] * sDiv[it][s * NP + j];
		dDbDiv -= t.st[j * NP + s] * sDiv[it][i * NP + s];
		dDaCurl -= t.st[i * NP + s] * sCurl[it][s * NP + j];
		dDbCurl -= t.st[j * NP + s] * sCurl[it][i * NP + s];
	}
	dDaDiv *= dInvDA;
	dDbDiv *= dInvDB;
	dDaCurl *= dInvDA;
	dDbCurl *= dInvDB;

	double dLocalNuDiv = nu_div;
	double dLocalNuVort = nu_vort;
	if (scale_nu) {
		dLocalNuDiv = dLocalNuDiv * g.nu_scale[e];
		dLocalNuVort = dLocalNuVort * g.nu_scale[e];
	}
	const double dUpdateUa =
		+dLocalNuDiv * dDaDiv
		- dLocalNuVort * dJ2D * (b0 * dDaCurl + b1 * dDbCurl);
	const double dUpdateUb =
		+dLocalNuDiv * dDbDiv
		+ dLocalNuVort * dJ2D * (a0 * dDaCurl + a1 * dDbCurl);
	if (active) {
		out[oU] -= dt * dUpdateUa;
		if (!xz) {
			out[oV] -= dt * dUpdateUb;
		}
	}
}

///////////////////////////////////////////////////////////////////////////////
// HorizontalDynamicsFEM::ApplyRayleighFriction
// (reference HorizontalDynamicsFEM.cpp:2418-2536): ten implicit relaxation
// sub-cycles towards the reference state wherever the strength is non-zero;
// u, v, rho-theta, w (u, rho-theta, w on x-z slices) - not rho.

__global__ void k_rayleigh(
	DevLayout lay, const double * ray_node, const double * ray_redge,
	const double * ref, double * data, double dt, int xz
) {
	const int NN = lay.nn;
	const int L = lay.nlev;
	const int nRayleighCycles = 10;
	const double dRayleighFactor = 1.0 / nRayleighCycles;
	const long long total = lay.nelem * (long long)lay.nrows_state * NN;
	for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	     idx < total; idx += (long long)gridDim.x * blockDim.x
	) {
		const int n = (int)(idx % NN);
		const long long er = idx / NN;
		const int row = (int)(er % lay.nrows_state);
		const long long e = er / lay.nrows_state;
		int c = 0;
		while (c + 1 < lay.ncomp && row >= lay.rowoff[c + 1]) c++;
		// nEffectiveC (:2444-2464)
		if (c == 4) continue;
		if (xz && c == 1) continue;
		const int k = row - lay.rowoff[c];
		const double dNu = lay.onedge[c]
			? ray_redge[((size_t)e * (L + 1) + k) * NN + n]
			: ray_node[((size_t)e * L + k) * NN + n];
		if (dNu == 0.0) continue;
		const size_t off = ((size_t)e * lay.nrows + row) * NN + n;
		const double r = ref[off];
		double x = data[off];
		for (int si = 0; si < nRayleighCycles; si++) {
			const double dNuNode = 1.0 / (1.0 + dRayleighFactor * dt * dNu);
			x = dNuNode * x + (1.0 - dNuNode) * r;
		}
		data[off] = x;
	}
}

#endif
