// Output-side interpolation: state and tracers of an instance
// evaluated at arbitrary points of the sphere and on arbitrary REta levels without
// bringing the instance back to the host.
//
//   GridPatchCSGLL::InterpolateData       src/atm/GridPatchCSGLL.cpp:1365-1780
//   Grid::ReduceInterpolate               src/atm/Grid.cpp:866-990
//   CubedSphereTrans::CoVecTransRLLFromABP  src/atm/CubedSphereTrans.cpp:640-729
//
// The caller locates each point (patch, element, Lagrangian coefficients of the
// element's GLL nodes: PolynomialInterp::LagrangianPolynomialCoeffs) and hands
// in the vertical operators (LinearColumnInterpFEM, dense with their [begin, end)
// windows); the device does the reduction: horizontal interpolation of every
// row of the element, the column operator, and the conversion to primitive
// variables (w / DerivR, covariant wind -> zonal / meridional).
#ifndef TB200_OUTPUT_CUH
#define TB200_OUTPUT_CUH

#include "tb200_platform.h"
#include "tb200_device.h"

#define TB_INTERP_MAXLEV 256

struct InterpArgs {
	int npts, nout;
	const int * elem;          // [npts] device element of the point, -1: not on this rank
	const double * ca;         // [npts][np] coefficients along alpha
	const double * cb;         // [npts][np] coefficients along beta
	const double * vop;        // [nout][nin] dense column operator
	const int * vbegin;        // [nout]
	const int * vend;          // [nout]
	int nin;                   // levels (L) or interfaces (L + 1) of the rows
	int row0;                  // first row of the component inside an element
	long long estride;         // rows per element of the array the rows are read from
	const double * derivr;     // w -> primitive: DerivR[2] at the rows' location
	                           // [e][nin][NN], or 0
	const double * zs;         // ... or ztop - zs (terrain-following, uniform levels)
	double ztop;
	int divide_derivr;
	double * out;              // [nout][npts] of this component
};

// one thread per point: horizontal interpolation of the nin rows of one
// component, then the column operator (:1640-1722)
__global__ void k_interpolate_component(DevLayout lay, InterpArgs a, const double * data) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= a.npts) return;
	const int e = a.elem[i];
	if (e < 0) return;                 // left at zero: another rank's point (MPI_Reduce sums)
	const int np = lay.np;
	const int NN = lay.nn;
	double col[TB_INTERP_MAXLEV];
	const double * ca = a.ca + (size_t)i * np;
	const double * cb = a.cb + (size_t)i * np;
	const double * src = data + ((size_t)e * a.estride + a.row0) * NN;
	for (int k = 0; k < a.nin; k++) {
		double v = 0.0;
		double dr = 1.0;
		if (a.divide_derivr) {
			// DerivR[2] at the element's first node (iA, iB), as the reference reads it
			dr = (a.derivr != 0) ? a.derivr[((size_t)e * a.nin + k) * NN]
			                     : (a.ztop - a.zs[(size_t)e * NN]);
		}
		for (int m = 0; m < np; m++) {
			for (int n = 0; n < np; n++) {
				if (a.divide_derivr) {
					v += ca[m] * cb[n] * src[(size_t)k * NN + m * np + n] / dr;
				} else {
					v += ca[m] * cb[n] * src[(size_t)k * NN + m * np + n];
				}
			}
		}
		col[k] = v;
	}
	// LinearColumnOperator::Apply (LinearColumnOperator.h:163-171)
	for (int ko = 0; ko < a.nout; ko++) {
		double o = 0.0;
		for (int l = a.vbegin[ko]; l < a.vend[ko]; l++) {
			o += a.vop[(size_t)ko * a.nin + l] * col[l];
		}
		a.out[(size_t)ko * a.npts + i] = o;
	}
}

// conversion of the interpolated covariant wind to zonal / meridional components
// (:1735-1776), in place on out_u, out_v [nout][npts]
__global__ void k_interpolate_wind(
	int npts, int nout, const int * elem, const int * panel, const double * alpha,
	const double * beta, double radius, double * out_u, double * out_v
) {
	const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (idx >= (long long)npts * nout) return;
	const int i = (int)(idx % npts);
	if (elem[i] < 0) return;
	const double dX = tan(alpha[i]);
	const double dY = tan(beta[i]);
	const int nP = panel[i];
	const double dUalpha = out_u[idx] / radius;
	const double dUbeta = out_v[idx] / radius;
	double dUlon, dUlat;
	const double dDelta2 = 1.0 + dX * dX + dY * dY;
	if ((nP > 3) && (fabs(dX) < 1.0e-13) && (fabs(dY) < 1.0e-13)) {
		dUlon = (nP == 4) ? dUalpha : (- dUalpha);
		dUlat = dUbeta;
	} else if (nP < 4) {
		dUlon =
			  dDelta2 / (1.0 + dX * dX) * dUalpha
			+ dDelta2 * dX * dY / (1.0 + dX * dX) / (1.0 + dY * dY) * dUbeta;
		dUlat =
			  dDelta2 / sqrt(1.0 + dX * dX) / (1.0 + dY * dY) * dUbeta;
		const double lat = atan(dY / sqrt(1.0 + dX * dX));
		dUlon *= cos(lat);
	} else {
		const double dRadius2 = (dX * dX + dY * dY);
		const double dRadius = sqrt(dRadius2);
		const double sgn = (nP == 4) ? 1.0 : -1.0;
		dUlon = sgn * (
			- dDelta2 * dY / (1.0 + dX * dX) / dRadius2 * dUalpha
			+ dDelta2 * dX / (1.0 + dY * dY) / dRadius2 * dUbeta);
		dUlat = sgn * (
			- dDelta2 * dX / (1.0 + dX * dX) / dRadius * dUalpha
			- dDelta2 * dY / (1.0 + dY * dY) / dRadius * dUbeta);
		const double lat = 0.5 * 3.14159265358979323846 - atan(sqrt(dX * dX + dY * dY));
		dUlon *= cos(lat);
	}
	out_u[idx] = dUlon;
	out_v[idx] = dUlat;
}

// ---- derived output fields on levels, one array shaped like a small instance ----------
// field[((e * 3 L) + row) * NN + n]: rows 0 .. L-1 vorticity, L .. 2L-1 divergence,
// 2L .. 3L-1 temperature (so that the averaging kernels of the DSS apply to it)

// GridPatch::ComputeTemperature (GridPatch.cpp:641-700, FORMULATION_RHOTHETA_PI):
// T = p / (rho R), p = PressureFromRhoTheta(rho theta)
__global__ void k_output_temperature(
	DevLayout lay, double pressure_scaling, double gamma, double R,
	const double * data, double * field
) {
	const int NN = lay.nn;
	const int L = lay.nlev;
	const long long total = lay.nelem * (long long)L * NN;
	for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	     idx < total; idx += (long long)gridDim.x * blockDim.x
	) {
		const long long e = idx / ((long long)L * NN);
		const int r = (int)(idx % ((long long)L * NN));
		const size_t ebase = (size_t)e * lay.nrows * NN;
		const double rt = data[ebase + (size_t)lay.rowoff[2] * NN + r];
		const double rho = data[ebase + (size_t)lay.rowoff[4] * NN + r];
		const double dPressure = pressure_scaling * exp(log(rt) * gamma);
		field[((size_t)e * 3 * L + 2 * L) * NN + r] = dPressure / (rho * R);
	}
}

// GridPatchCSGLL::ComputeCurlAndDiv (GridPatchCSGLL.cpp:1132-1305) per (element,
// level): one thread per node, the element's tile in shared memory.  ITEMS
// (element, level) pairs per block.
template <int NP, int ITEMS>
__global__ void __launch_bounds__(NP * NP * ITEMS)
k_output_curl_div(
	DevLayout lay, DevGeom g, DevTables t, const double * data, double * field
) {
	const int NN = NP * NP;
	__shared__ double sUa[ITEMS][NN];
	__shared__ double sUb[ITEMS][NN];
	__shared__ double sJa[ITEMS][NN];    // Jacobian2D * contravariant u^alpha
	__shared__ double sJb[ITEMS][NN];
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
	const double dUa = data[ebase + (size_t)(lay.rowoff[0] + k) * NN + n];
	const double dUb = data[ebase + (size_t)(lay.rowoff[1] + k) * NN + n];
	const double conUa = + g.a0[g2] * dUa + g.a1[g2] * dUb;
	const double conUb = + g.b0[g2] * dUa + g.b1[g2] * dUb;
	sUa[it][n] = dUa;
	sUb[it][n] = dUb;
	sJa[it][n] = g.j2d[g2] * conUa;
	sJb[it][n] = g.j2d[g2] * conUb;
	__syncthreads();
	double dDaUb = 0.0, dDbUa = 0.0, dDaJUa = 0.0, dDbJUb = 0.0;
#pragma unroll
	for (int s = 0; s < NP; s++) {
		dDaUb += sUb[it][s * NP + j] * t.dx[s * NP + i];
		dDbUa += sUa[it][i * NP + s] * t.dx[s * NP + j];
		dDaJUa += sJa[it][s * NP + j] * t.dx[s * NP + i];
		dDbJUb += sJb[it][i * NP + s] * t.dx[s * NP + j];
	}
	dDaUb *= g.inv_da[e];
	dDbUa *= g.inv_db[e];
	dDaJUa *= g.inv_da[e];
	dDbJUb *= g.inv_db[e];
	const double dInvJacobian2D = 1.0 / g.j2d[g2];
	if (active) {
		const size_t o = ((size_t)e * 3 * L + k) * NN + n;
		field[o + (size_t)L * NN] = (dDaJUa + dDbJUb) * dInvJacobian2D;
		field[o] = (dDaUb - dDbUa) * dInvJacobian2D;
	}
}

#endif
