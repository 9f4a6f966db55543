/*
 * Kessler warm-rain microphysics: CPU restatement.  TEST INFRASTRUCTURE ONLY
 * (only tests/ may load this; the product never does).
 *
 * PARITY UNPINNED.  The reference's implementation is Fortran
 * (/root/reference/test/dcmip2016/interface/kessler.f90, called per column by
 * KesslerPhysics::Perform, test/dcmip2016/KesslerPhysics.cpp:84-285) and the
 * image has no Fortran compiler, so the reference itself cannot be run here
 * and there are no golden vectors for this path in the reference's tests.
 * What follows restates the published algorithm line by line:
 *
 *   kessler_column()   kessler.f90:62-185  (Klemp, Skamarock and Park 2015)
 *   kessler_physics()  KesslerPhysics.cpp:143-279 (FORMULATION_RHOTHETA_PI,
 *                      rho theta on levels)
 *
 * Reading of the Fortran that a compiler would fix and this file states:
 *  - "REAL, DIMENSION(nz) :: r, rhalf, velqr, sed, pc" are default (single
 *    precision) reals: assignments to them round to float;
 *  - real literals without a kind suffix (.001, 0.1364, .875, .2046, .525, 1.,
 *    .5, 0.) are single-precision constants promoted to double where they meet
 *    double operands;
 *  - amax1 / amin1 / dim applied to double arguments are evaluated in double
 *    (ifort; gfortran needs -fallow-argument-mismatch for this source).
 */
#include <math.h>

static double dmax(double a, double b) { return (a > b) ? a : b; }
static double dmin(double a, double b) { return (a < b) ? a : b; }

/* kessler.f90:62-185; arrays of length nz, surface first; work: 5 * nz floats */
void kessler_column(
	double * theta, double * qv, double * qc, double * qr,
	const double * rho, const double * pk, double dt, const double * z,
	int nz, double * precl, float * work
) {
	float * r = work, * rhalf = work + nz, * velqr = work + 2 * nz;
	float * sed = work + 3 * nz, * pc = work + 4 * nz;
	const double f2x = 17.27;
	const double f5 = 237.3 * f2x * 2500000.0 / 1003.0;
	const double xk = .2875;
	const double psl = 1000.0;
	const double rhoqr = 1000.0;
	const double e1364 = (double)0.1364f, e875 = (double).875f;
	const double e2046 = (double).2046f, e525 = (double).525f, c001 = (double).001f;
	int k, nt, rainsplit;
	double dt_max, dt0;

	for (k = 0; k < nz; k++) {
		r[k] = (float)(0.001 * rho[k]);
		rhalf[k] = (float)sqrt(rho[0] / rho[k]);
		pc[k] = (float)(3.8 / (pow(pk[k], (double)1.f / xk) * psl));
		velqr[k] = (float)(36.34 * pow(qr[k] * (double)r[k], e1364) * (double)rhalf[k]);
	}

	dt_max = dt;
	for (k = 0; k < nz - 1; k++) {
		if ((double)velqr[k] != 0.0) {
			dt_max = dmin(dt_max, 0.8 * (z[k + 1] - z[k]) / (double)velqr[k]);
		}
	}
	rainsplit = (int)ceil(dt / dt_max);
	dt0 = dt / (double)rainsplit;

	*precl = 0.0;
	for (nt = 1; nt <= rainsplit; nt++) {
		*precl = *precl + rho[0] * qr[0] * (double)velqr[0] / rhoqr;

		for (k = 0; k < nz - 1; k++) {
			sed[k] = (float)(dt0 * ((double)r[k + 1] * qr[k + 1] * (double)velqr[k + 1]
				- (double)r[k] * qr[k] * (double)velqr[k]) / ((double)r[k] * (z[k + 1] - z[k])));
		}
		sed[nz - 1] = (float)(-dt0 * qr[nz - 1] * (double)velqr[nz - 1]
			/ ((double).5f * (z[nz - 1] - z[nz - 2])));

		for (k = 0; k < nz; k++) {
			double qrprod, qvs, prod, ern, rq, pt;
			qrprod = qc[k] - (qc[k] - dt0 * dmax(c001 * (qc[k] - .001), 0.0))
				/ (1.0 + dt0 * 2.2 * pow(qr[k], e875));
			qc[k] = dmax(qc[k] - qrprod, 0.0);
			qr[k] = dmax(qr[k] + qrprod + (double)sed[k], 0.0);

			pt = pk[k] * theta[k];
			qvs = (double)pc[k] * exp(f2x * (pt - 273.0) / (pt - 36.0));
			prod = (qv[k] - qvs) / (1.0 + qvs * f5 / ((pt - 36.0) * (pt - 36.0)));

			rq = (double)r[k] * qr[k];
			ern = dmin(dmin(
				dt0 * (((1.6 + 124.9 * pow(rq, e2046)) * pow(rq, e525))
					/ (2550000.0 * (double)pc[k] / (3.8 * qvs) + 540000.0))
					* (dmax(qvs - qv[k], 0.0) / ((double)r[k] * qvs)),
				dmax(-prod - qc[k], 0.0)), qr[k]);

			theta[k] = theta[k] + 2500000.0 / (1003.0 * pk[k]) * (dmax(prod, -qc[k]) - ern);
			qv[k] = dmax(qv[k] - dmax(prod, -qc[k]) + ern, 0.0);
			qc[k] = qc[k] + dmax(prod, -qc[k]);
			qr[k] = qr[k] - ern;
		}

		if (nt != rainsplit) {
			for (k = 0; k < nz; k++) {
				velqr[k] = (float)(36.34 * pow(qr[k] * (double)r[k], e1364) * (double)rhalf[k]);
			}
		}
	}
	*precl = *precl / (double)rainsplit;
}

/*
 * KesslerPhysics::Perform for one column (KesslerPhysics.cpp:143-279):
 * rhotheta, rho, the three tracer densities rho qv, rho qc, rho qr on nz levels
 * (updated in place), heights z of the levels; *precip accumulates precl * dt.
 * pressure_scaling, gamma, R: PhysicalConstants::PressureFromRhoTheta
 * (PhysicalConstants.h:382-384).  work: 7 * nz doubles + 5 * nz floats.
 */
void kessler_physics_column(
	double * rhotheta, double * rho, double * rqv, double * rqc, double * rqr,
	const double * z, int nz, double dt, double pressure_scaling, double gamma, double R,
	double * precip, double * work
) {
	double * theta = work, * qv = work + nz, * qc = work + 2 * nz, * qr = work + 3 * nz;
	double * rhod = work + 4 * nz, * pk = work + 5 * nz, * thetav = work + 6 * nz;
	float * fwork = (float *)(work + 7 * nz);
	double precl = 0.0;
	int k;
	for (k = 0; k < nz; k++) {
		const double dRho = rho[k];
		const double dRhoD = dRho - rqv[k] - rqc[k] - rqr[k];
		double dPressure, dTv;
		thetav[k] = rhotheta[k] / rho[k];
		dPressure = pressure_scaling * exp(log(dRho * thetav[k]) * gamma);
		dTv = dPressure / (dRho * R);
		qv[k] = rqv[k] / rho[k];
		if (qv[k] < 0.0) qv[k] = 0.0;
		qc[k] = rqc[k] / rho[k];
		if (qc[k] < 0.0) qc[k] = 0.0;
		qr[k] = rqr[k] / rho[k];
		if (qr[k] < 0.0) qr[k] = 0.0;
		theta[k] = thetav[k] / (1.0 + 0.61 * qv[k]);
		rhod[k] = dRhoD;
		pk[k] = dTv / thetav[k];
	}
	kessler_column(theta, qv, qc, qr, rhod, pk, dt, z, nz, &precl, fwork);
	*precip += precl * dt;
	for (k = 0; k < nz; k++) {
		rho[k] = rhod[k] / (1.0 - qv[k] - qc[k] - qr[k]);
		rqv[k] = qv[k] * rho[k];
		rqc[k] = qc[k] * rho[k];
		rqr[k] = qr[k] * rho[k];
	}
	for (k = 0; k < nz; k++) {
		rhotheta[k] = rho[k] * theta[k] * (1.0 + 0.61 * qv[k]);
	}
}

/* batch driver for the tests: ncol columns, arrays [ncol][nz] */
void kessler_physics_batch(
	double * rhotheta, double * rho, double * rqv, double * rqc, double * rqr,
	const double * z, int ncol, int nz, double dt, double pressure_scaling, double gamma,
	double R, double * precip, double * work
) {
	int c;
	for (c = 0; c < ncol; c++) {
		kessler_physics_column(rhotheta + (long)c * nz, rho + (long)c * nz, rqv + (long)c * nz,
			rqc + (long)c * nz, rqr + (long)c * nz, z + (long)c * nz, nz, dt,
			pressure_scaling, gamma, R, precip + c, work);
	}
}
