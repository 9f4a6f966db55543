// Conservation diagnostics (kernels); host side in tb200_api.cu.
#ifndef TB200_DIAG_CUH
#define TB200_DIAG_CUH

#include "tb200_platform.h"
#include "tb200_device.h"
#include "tb200_kernels.cuh"

///////////////////////////////////////////////////////////////////////////////
// Conservation diagnostics on the device: Grid::ComputeTotalEnergy,
// ComputeTotalPotentialEnstrophy, ComputeTotalVerticalMomentum
// (reference Grid.cpp:529-623, GridPatch.cpp:925-1288).
//
// The reference's routines read W on levels and rho on interfaces - slots the
// Lorenz-staggered state does not carry and the device does not store; they are
// formed here with the reference's interpolation operators
// (Grid::InterpolateREdgeToNode / InterpolateNodeToREdge, Grid.cpp:843-863), i.e.
// the value the reference returns once those slots are refreshed.

struct DiagArgs {
	double g, gamma, pscale;      // PhysicalConstants: G, Gamma, PressureScaling
	int shallow;
	int what;                     // 0 energy, 1 potential enstrophy, 2 vertical momentum
	const double * vort;          // shallow water: DSS'd relative vorticity [e][nrows][NN] row 0..
};

__device__ __forceinline__ double tb_op_row(
	const DevOp & op, const double * col, size_t stride, int k
) {
	// LinearColumnOperator::Apply for one output row (LinearColumnOperator.h:62-236)
	double v = 0.0;
	const int b = op.begin[k], e = op.end[k];
	for (int l = b; l < e; l++) {
		v += op.coeff[(size_t)k * op.width + (l - b)] * col[(size_t)l * stride];
	}
	return v;
}

__global__ void k_diagnostic(
	DevLayout lay, DevGeom g, DevOps ops, DiagArgs da, const double * data,
	const double * area_node, const double * area_redge, double * sums
) {
	__shared__ double red[128];
	const int NN = lay.nn;
	const int L = lay.nlev;
	const long long ncol = lay.nelem * NN;
	double acc = 0.0;
	for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	     idx < ncol; idx += (long long)gridDim.x * blockDim.x
	) {
		const long long e = idx / NN;
		const int n = (int)(idx % NN);
		const size_t ebase = (size_t)e * lay.nrows * NN + n;
		const size_t g2 = (size_t)e * NN + n;
		const double * an = area_node + (size_t)e * L * NN + n;
		if (da.shallow) {
			const double u = data[ebase + (size_t)lay.rowoff[0] * NN];
			const double v = data[ebase + (size_t)lay.rowoff[1] * NN];
			const double h = data[ebase + (size_t)lay.rowoff[2] * NN];
			const double zs = g.zs[g2];
			if (da.what == 0) {
				// GridPatch.cpp:964-995
				double dUdotU = +g.b1[g2] * u * u - 2.0 * g.a1[g2] * u * v + g.a0[g2] * v * v;
				dUdotU *= g.j2d[g2] * g.j2d[g2];
				const double dKineticEnergy = 0.5 * dUdotU * (h - zs);
				const double dPotentialEnergy = 0.5 * da.g * (h * h - zs * zs);
				acc += an[0] * (dKineticEnergy + dPotentialEnergy);
			} else if (da.what == 1) {
				// GridPatch.cpp:1180-1199 (planetary vorticity 2 Omega sin(lat) = CoriolisF)
				const double dAbsoluteVorticity = da.vort[ebase] + g.f[g2];
				acc += an[0] * 0.5 * dAbsoluteVorticity * dAbsoluteVorticity / (h - zs);
			}
			continue;
		}
		const double * U = data + ebase + (size_t)lay.rowoff[0] * NN;
		const double * V = data + ebase + (size_t)lay.rowoff[1] * NN;
		const double * P = data + ebase + (size_t)lay.rowoff[2] * NN;
		const double * W = data + ebase + (size_t)lay.rowoff[3] * NN;
		const double * R = data + ebase + (size_t)lay.rowoff[4] * NN;
		if (da.what == 1) {
			// nonhydrostatic "potential enstrophy" of the reference: zonal momentum
			// (GridPatch.cpp:1203-1218)
			for (int k = 0; k < L; k++) {
				acc += an[(size_t)k * NN] * R[(size_t)k * NN] * U[(size_t)k * NN];
			}
			continue;
		}
		const ColMetric cm = tb_col_metric(g, g2);
		const size_t g3 = (size_t)e * L * NN + n;
		const size_t g3e = (size_t)e * (L + 1) * NN + n;
		for (int k = 0; k < L; k++) {
			const double dCovUa = U[(size_t)k * NN];
			const double dCovUb = V[(size_t)k * NN];
			const double dRho = R[(size_t)k * NN];
			// W on levels: InterpolateREdgeToNode (GridPatchGLL.cpp:113-143)
			const double dCovUx = tb_op_row(ops.op[TB200_OP_INTERP_E2N], W, NN, k);
			if (da.what == 2) {
				acc += an[(size_t)k * NN] * dRho * dCovUx;      // GridPatch.cpp:1273-1282
				continue;
			}
			double a0, a1, a2, b0, b1, b2, x0, x1;
			if (g.analytic) {
				const LevMetric lm = tb_lev_metric(cm, g.reta_n[k]);
				a0 = cm.a0; a1 = cm.a1; a2 = lm.a2;
				b0 = cm.b0; b1 = cm.b1; b2 = lm.b2;
				x0 = lm.a2; x1 = lm.b2;
			} else {
				const size_t o = g3 + (size_t)k * NN;
				a0 = g.ca[0][o]; a1 = g.ca[1][o]; a2 = g.ca[2][o];
				b0 = g.cb[0][o]; b1 = g.cb[1][o]; b2 = g.cb[2][o];
				x0 = g.cx[0][o]; x1 = g.cx[1][o];
			}
			// GridPatch.cpp:1063-1112
			const double dConUa = a0 * dCovUa + a1 * dCovUb + a2 * dCovUx;
			const double dConUb = b0 * dCovUa + b1 * dCovUb + b2 * dCovUx;
			double dUdotU = dConUa * dCovUa + dConUb * dCovUb;
			dUdotU += x0 * dCovUa * dCovUx;
			dUdotU += x1 * dCovUb * dCovUx;
			const double dKineticEnergy = 0.5 * dRho * dUdotU;
			const double dPressure = da.pscale * exp(log(P[(size_t)k * NN]) * da.gamma);
			const double dInternalEnergy = dPressure / (da.gamma - 1.0);
			const double zs = g.zs[g2];
			const double dZ = zs + g.reta_n[k] * (g.ztop - zs);     // GridPatchCSGLL.cpp:643-645
			const double dPotentialEnergy = da.g * dRho * dZ;
			acc += an[(size_t)k * NN] * (dKineticEnergy + dInternalEnergy + dPotentialEnergy);
		}
		if (da.what == 0) {
			// vertical kinetic energy on interfaces (GridPatch.cpp:1117-1134), rho there
			// by InterpolateNodeToREdge
			const double * ae = area_redge + (size_t)e * (L + 1) * NN + n;
			for (int k = 0; k <= L; k++) {
				const double dCovUx = W[(size_t)k * NN];
				const double dRhoE = tb_op_row(ops.op[TB200_OP_INTERP_N2E], R, NN, k);
				double x2;
				if (g.analytic) {
					x2 = tb_lev_metric(cm, g.reta_e[k]).x2;
				} else {
					x2 = g.cxe[2][g3e + (size_t)k * NN];
				}
				acc += ae[(size_t)k * NN] * (0.5 * dRhoE * x2 * dCovUx * dCovUx);
			}
		}
	}
	red[threadIdx.x] = acc;
	__syncthreads();
	for (int s = blockDim.x / 2; s > 0; s >>= 1) {
		if ((int)threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
		__syncthreads();
	}
	if (threadIdx.x == 0) atomicAdd(&sums[0], red[0]);
}

// relative vorticity of a shallow-water state, element by element
// (GridPatchCSGLL::ComputeVorticityDivergence, GridPatchCSGLL.cpp:1309-1460),
// into row 0.. of `out` (an instance used as scratch)
template <int NP>
__global__ void __launch_bounds__(NP * NP * 8) k_sw_vorticity(
	DevLayout lay, DevGeom g, DevTables t, const double * in, double * out
) {
	const int NN = NP * NP;
	__shared__ double sUa[8][NN];
	__shared__ double sUb[8][NN];
	const int it = threadIdx.x / NN;
	const int n = threadIdx.x % NN;
	const int i = n / NP, j = n % NP;
	const int L = lay.nlev;
	const long long nitems = lay.nelem * L;
	long long item = (long long)blockIdx.x * 8 + it;
	const bool active = (item < nitems);
	if (!active) item = nitems - 1;
	const long long e = item / L;
	const int k = (int)(item % L);
	const size_t ebase = (size_t)e * lay.nrows * NN;
	sUa[it][n] = in[ebase + (size_t)(lay.rowoff[0] + k) * NN + n];
	sUb[it][n] = in[ebase + (size_t)(lay.rowoff[1] + k) * NN + n];
	__syncthreads();
	double dDaUb = 0.0, dDbUa = 0.0;
#pragma unroll
	for (int s = 0; s < NP; s++) {
		dDaUb += sUb[it][s * NP + j] * t.dx[s * NP + i];
		dDbUa += sUa[it][i * NP + s] * t.dx[s * NP + j];
	}
	dDaUb *= g.inv_da[e];
	dDbUa *= g.inv_db[e];
	if (active) {
		out[ebase + (size_t)(lay.rowoff[0] + k) * NN + n] =
			(dDaUb - dDbUa) / g.j2d[(size_t)e * NN + n];
	}
}


#endif
