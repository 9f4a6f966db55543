// Column physics as device workflow steps.
//
//   HeldSuarezPhysics::Perform   src/atm/HeldSuarezPhysics.cpp:62-301
//
// Held-Suarez forcing (Rayleigh friction of the horizontal wind in the boundary
// layer, Newtonian relaxation of temperature towards a zonally symmetric
// equilibrium) acts pointwise on instance 0 once per step; in the reference it
// is a WorkflowProcess on the host arrays, which would put a full state
// round trip over the bus between any two device steps (Model.cpp:477-481).
// Here it is one streaming pass over u, v, rho-theta (read rho) on the device.
//
// The reference as compiled (FORMULATION_RHOTHETA_PI, Lorenz staggering: every
// prognostic variable but w on levels) takes the surface pressure from
// PressureFromRhoTheta(dataREdge[R][0] * dataREdge[T][0]) (:112-115): the slots
// of rho and rho-theta on the lowest *interface* of instance 0.  The dynamics
// never writes those slots (they keep what EvaluateTestCase put there), and
// their product is not a pressure argument in this formulation; the values are
// what they are.  The device gets that product once per column
// (tb200_upload_held_suarez) and reproduces the arithmetic as written.
#ifndef TB200_PHYSICS_CUH
#define TB200_PHYSICS_CUH

#include "tb200_platform.h"
#include "tb200_device.h"

struct HeldSuarezArgs {
	const double * latitude;        // [e][NN]
	const double * surface_product; // [e][NN]: dataREdge[R][0] * dataREdge[T][0]
	double dt;
	double pressure_scaling;        // PhysicalConstants m_dPressureScaling
	double gamma, kappa, R, p0;
};

// parameters of HeldSuarezPhysics.cpp:25-47
#define TB_HS_BOUNDARY_SIGMA 0.7
#define TB_HS_K_FRICTION (1.0 / 86400.0)
#define TB_HS_KA ((1.0 / 40.0) / 86400.0)
#define TB_HS_KS ((1.0 / 4.0) / 86400.0)
#define TB_HS_DELTA_TY 60.0
#define TB_HS_DELTA_THETA_Z 10.0
#define TB_HS_MINIMUM_T 200.0
#define TB_HS_MAXIMUM_T 315.0

// PhysicalConstants::PressureFromRhoTheta (PhysicalConstants.h:382-384)
__device__ __forceinline__ double tb_pressure_from_rhotheta(const HeldSuarezArgs & a, double rhotheta) {
	return a.pressure_scaling * exp(log(rhotheta) * a.gamma);
}

// one thread per (element, level, node)
__global__ void k_held_suarez(DevLayout lay, HeldSuarezArgs a, double * data) {
	const int NN = lay.nn;
	const int L = lay.nlev;
	const long long total = lay.nelem * (long long)L * NN;
	for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	     idx < total; idx += (long long)gridDim.x * blockDim.x
	) {
		const long long e = idx / ((long long)L * NN);
		const int r = (int)(idx % ((long long)L * NN));
		const int k = r / NN;
		const int n = r % NN;
		const size_t ebase = (size_t)e * lay.nrows * NN;
		double * pU = data + ebase + (size_t)(lay.rowoff[0] + k) * NN + n;
		double * pV = data + ebase + (size_t)(lay.rowoff[1] + k) * NN + n;
		double * pT = data + ebase + (size_t)(lay.rowoff[2] + k) * NN + n;
		const double dRho = data[ebase + (size_t)(lay.rowoff[4] + k) * NN + n];
		const double dRhoTheta = pT[0];
		const size_t c2 = (size_t)e * NN + n;

		// :112-115
		const double dSurfacePressure = tb_pressure_from_rhotheta(a, a.surface_product[c2]);

		// velocity: :121-140 (pressure from rho * theta-slot)
		{
			const double dPressure = tb_pressure_from_rhotheta(a, dRho * dRhoTheta);
			const double dSigma = dPressure / dSurfacePressure;
			double dBoundaryScale =
				(dSigma - TB_HS_BOUNDARY_SIGMA) / (1.0 - TB_HS_BOUNDARY_SIGMA);
			if (dBoundaryScale < 0.0) {
				dBoundaryScale = 0.0;
			}
			pU[0] = pU[0] / (1.0 + TB_HS_K_FRICTION * dBoundaryScale * a.dt);
			pV[0] = pV[0] / (1.0 + TB_HS_K_FRICTION * dBoundaryScale * a.dt);
		}

		// rho-theta on levels: :144-211
		{
			const double dPressure = tb_pressure_from_rhotheta(a, dRhoTheta);
			const double dSigma = dPressure / dSurfacePressure;
			double dBoundaryScale =
				(dSigma - TB_HS_BOUNDARY_SIGMA) / (1.0 - TB_HS_BOUNDARY_SIGMA);
			if (dBoundaryScale < 0.0) {
				dBoundaryScale = 0.0;
			}
			const double dT = dPressure / (dRho * a.R);
			const double dLat = a.latitude[c2];
			const double dSinLat = sin(dLat);
			const double dCosLat = cos(dLat);
			const double dCos4Lat = dCosLat * dCosLat * dCosLat * dCosLat;
			const double dKT = TB_HS_KA + (TB_HS_KS - TB_HS_KA) * dBoundaryScale * dCos4Lat;
			double dTeq =
				TB_HS_MAXIMUM_T
				- TB_HS_DELTA_TY * dSinLat * dSinLat
				- TB_HS_DELTA_THETA_Z * log(dPressure / a.p0) * dCosLat * dCosLat;
			dTeq *= pow(dPressure / a.p0, a.kappa);
			if (dTeq < TB_HS_MINIMUM_T) {
				dTeq = TB_HS_MINIMUM_T;
			}
			const double dDH = - dKT / a.gamma * (1.0 + (a.gamma - 1.0) * dTeq / dT);
			const double dH = - dKT / a.gamma * (1.0 - dTeq / dT);
			pT[0] = dRhoTheta * (1.0 + a.dt / (1.0 - a.dt * dDH) * dH);
		}
	}
}

///////////////////////////////////////////////////////////////////////////////
// Kessler warm-rain microphysics as a device workflow step.
//
//   KesslerPhysics::Perform   test/dcmip2016/KesslerPhysics.cpp:84-285
//   KESSLER                   test/dcmip2016/interface/kessler.f90:62-185
//
// PARITY UNPINNED: the reference's kernel is Fortran and the image has no
// Fortran compiler, so it cannot be run here; the device kernel is held to a C
// restatement of the same source (oracle/kessler_port.c, which states how the
// mixed-precision declarations of the Fortran are read).  One thread per
// element-local column (the reference visits every node of a patch, duplicates
// included); the column arrays live in a coalesced scratch [entry][thread].
// Tracers 0, 1, 2 are rho qv, rho qc, rho qr.

struct KesslerArgs {
	const double * zs;           // topography [e][NN]
	const double * reta_n;       // REta of the levels
	double ztop;
	double dt;
	double pressure_scaling, gamma, R;
	double * precip;             // accumulated precipitation [e][NN] (UserData2D[0])
	double * ws;                 // scratch: 9 * L doubles per column of the launch
	long long col0, ncols;       // columns (element-local nodes) of this launch
};

// assignment to a default-real (single precision) Fortran variable
__device__ __forceinline__ double tb_f32(double x) { return (double)(float)x; }

__global__ void k_kessler(DevLayout lay, KesslerArgs a, double * data) {
	const int NN = lay.nn;
	const int nz = lay.nlev;
	const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= a.ncols) return;
	const long long col = a.col0 + t;
	const long long e = col / NN;
	const int n = (int)(col % NN);
	const size_t ebase = (size_t)e * lay.nrows * NN + n;
	double * pT = data + ebase + (size_t)lay.rowoff[2] * NN;
	double * pR = data + ebase + (size_t)lay.rowoff[4] * NN;
	double * pQv = data + ebase + (size_t)(lay.troff + 0 * nz) * NN;
	double * pQc = data + ebase + (size_t)(lay.troff + 1 * nz) * NN;
	double * pQr = data + ebase + (size_t)(lay.troff + 2 * nz) * NN;
	const double zs = a.zs[(size_t)e * NN + n];

	// scratch [entry][thread of the launch]
	const size_t S = (size_t)a.ncols;
	double * w0 = a.ws + t;
#define TB_KW(q, k) w0[((size_t)(q) * nz + (k)) * S]
	enum { THETA = 0, QV, QC, QR, RHOD, PK, VELQR, SED, PC };
#define TB_KZ(k) (zs + a.reta_n[k] * (a.ztop - zs))

	const double f2x = 17.27;
	const double f5 = 237.3 * f2x * 2500000.0 / 1003.0;
	const double xk = .2875;
	const double psl = 1000.0;
	const double rhoqr = 1000.0;
	const double e1364 = (double)0.1364f, e875 = (double).875f;
	const double e2046 = (double).2046f, e525 = (double).525f, c001 = (double).001f;

	// KesslerPhysics.cpp:143-221
	for (int k = 0; k < nz; k++) {
		const double dRho = pR[(size_t)k * NN];
		const double rqv = pQv[(size_t)k * NN], rqc = pQc[(size_t)k * NN], rqr = pQr[(size_t)k * NN];
		const double dRhoD = dRho - rqv - rqc - rqr;
		const double thetav = pT[(size_t)k * NN] / dRho;
		const double dPressure = a.pressure_scaling * exp(log(dRho * thetav) * a.gamma);
		const double dTv = dPressure / (dRho * a.R);
		double qv = rqv / dRho; if (qv < 0.0) qv = 0.0;
		double qc = rqc / dRho; if (qc < 0.0) qc = 0.0;
		double qr = rqr / dRho; if (qr < 0.0) qr = 0.0;
		TB_KW(THETA, k) = thetav / (1.0 + 0.61 * qv);
		TB_KW(QV, k) = qv;
		TB_KW(QC, k) = qc;
		TB_KW(QR, k) = qr;
		TB_KW(RHOD, k) = dRhoD;
		TB_KW(PK, k) = dTv / thetav;
	}

	// kessler.f90:101-185
	const double rho1 = TB_KW(RHOD, 0);
	for (int k = 0; k < nz; k++) {
		const double rho = TB_KW(RHOD, k), pk = TB_KW(PK, k);
		const double r = tb_f32(0.001 * rho);
		const double rhalf = tb_f32(sqrt(rho1 / rho));
		TB_KW(PC, k) = tb_f32(3.8 / (pow(pk, (double)1.f / xk) * psl));
		TB_KW(VELQR, k) = tb_f32(36.34 * pow(TB_KW(QR, k) * r, e1364) * rhalf);
	}
	double dt_max = a.dt;
	for (int k = 0; k < nz - 1; k++) {
		const double v = TB_KW(VELQR, k);
		if (v != 0.0) {
			dt_max = fmin(dt_max, 0.8 * (TB_KZ(k + 1) - TB_KZ(k)) / v);
		}
	}
	const int rainsplit = (int)ceil(a.dt / dt_max);
	const double dt0 = a.dt / (double)rainsplit;

	double precl = 0.0;
	for (int nt = 1; nt <= rainsplit; nt++) {
		precl = precl + rho1 * TB_KW(QR, 0) * TB_KW(VELQR, 0) / rhoqr;

		for (int k = 0; k < nz - 1; k++) {
			const double r0 = tb_f32(0.001 * TB_KW(RHOD, k)), r1 = tb_f32(0.001 * TB_KW(RHOD, k + 1));
			TB_KW(SED, k) = tb_f32(dt0 * (r1 * TB_KW(QR, k + 1) * TB_KW(VELQR, k + 1)
				- r0 * TB_KW(QR, k) * TB_KW(VELQR, k)) / (r0 * (TB_KZ(k + 1) - TB_KZ(k))));
		}
		TB_KW(SED, nz - 1) = tb_f32(-dt0 * TB_KW(QR, nz - 1) * TB_KW(VELQR, nz - 1)
			/ ((double).5f * (TB_KZ(nz - 1) - TB_KZ(nz - 2))));

		for (int k = 0; k < nz; k++) {
			double qc = TB_KW(QC, k), qr = TB_KW(QR, k), qv = TB_KW(QV, k), theta = TB_KW(THETA, k);
			const double pk = TB_KW(PK, k), pc = TB_KW(PC, k);
			const double r = tb_f32(0.001 * TB_KW(RHOD, k));
			const double qrprod = qc - (qc - dt0 * fmax(c001 * (qc - .001), 0.0))
				/ (1.0 + dt0 * 2.2 * pow(qr, e875));
			qc = fmax(qc - qrprod, 0.0);
			qr = fmax(qr + qrprod + TB_KW(SED, k), 0.0);

			const double pt = pk * theta;
			const double qvs = pc * exp(f2x * (pt - 273.0) / (pt - 36.0));
			const double prod = (qv - qvs) / (1.0 + qvs * f5 / ((pt - 36.0) * (pt - 36.0)));

			const double rq = r * qr;
			const double ern = fmin(fmin(
				dt0 * (((1.6 + 124.9 * pow(rq, e2046)) * pow(rq, e525))
					/ (2550000.0 * pc / (3.8 * qvs) + 540000.0))
					* (fmax(qvs - qv, 0.0) / (r * qvs)),
				fmax(-prod - qc, 0.0)), qr);

			theta = theta + 2500000.0 / (1003.0 * pk) * (fmax(prod, -qc) - ern);
			qv = fmax(qv - fmax(prod, -qc) + ern, 0.0);
			qc = qc + fmax(prod, -qc);
			qr = qr - ern;
			TB_KW(THETA, k) = theta; TB_KW(QV, k) = qv; TB_KW(QC, k) = qc; TB_KW(QR, k) = qr;
		}

		if (nt != rainsplit) {
			for (int k = 0; k < nz; k++) {
				const double rho = TB_KW(RHOD, k);
				const double r = tb_f32(0.001 * rho);
				const double rhalf = tb_f32(sqrt(rho1 / rho));
				TB_KW(VELQR, k) = tb_f32(36.34 * pow(TB_KW(QR, k) * r, e1364) * rhalf);
			}
		}
	}
	precl = precl / (double)rainsplit;

	// KesslerPhysics.cpp:235-279
	a.precip[(size_t)e * NN + n] += precl * a.dt;
	for (int k = 0; k < nz; k++) {
		const double qv = TB_KW(QV, k), qc = TB_KW(QC, k), qr = TB_KW(QR, k);
		const double rho = TB_KW(RHOD, k) / (1.0 - qv - qc - qr);
		pR[(size_t)k * NN] = rho;
		pQv[(size_t)k * NN] = qv * rho;
		pQc[(size_t)k * NN] = qc * rho;
		pQr[(size_t)k * NN] = qr * rho;
		pT[(size_t)k * NN] = rho * TB_KW(THETA, k) * (1.0 + 0.61 * qv);
	}
#undef TB_KW
#undef TB_KZ
}

#endif
