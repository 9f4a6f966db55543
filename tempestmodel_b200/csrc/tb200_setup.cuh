// Device-side set-up of the cubed-sphere runs: the 2-D metric and
// the initial state are evaluated where they are used instead of being built
// on the host (the reference spends ~15 minutes of one core on them at ne = 120)
// and copied over the bus.
//
//   CubedSphereTrans::RLLFromXYP                    src/atm/CubedSphereTrans.cpp:200-266
//   GridPatchCSGLL::EvaluateGeometricTerms, 2-D     src/atm/GridPatchCSGLL.cpp:295-343
//   BaroclinicWaveJWTest::EvaluateTopography        test/nonhydro_sphere/BaroclinicWaveJWTest.cpp:170-204
//   ...::CalculateGeopotentialTemperature           :210-289
//   ...::EtaFromRLL                                 :297-345
//   ...::EvaluateReferenceState / PointwiseState    :351-413
//   GridPatchCSGLL::EvaluateTestCase                src/atm/GridPatchCSGLL.cpp:578-920
//   CubedSphereTrans::CoVecTransABPFromRLL          src/atm/CubedSphereTrans.cpp:549-636
//   EquationSet::ConvertComponents                  src/atm/EquationSet.cpp:153-155
//
// Inputs already on the device: X = tan(alpha), Y = tan(beta) per node
// (tb200_set_terrain_metric), the topography (tb200_upload_geometry) and the
// vertical coordinate (tb200_set_vertical_coordinate).
#ifndef TB200_SETUP_CUH
#define TB200_SETUP_CUH

#include "tb200_platform.h"
#include "tb200_device.h"

#define TB_PI 3.14159265358979323846

// longitude / latitude of a point (X, Y) on a panel
__device__ __forceinline__ void tb_rll_from_xyp(double X, double Y, int panel, double & lon, double & lat) {
	if (panel < 4) {
		lon = atan(X) + 0.5 * TB_PI * panel;
		lat = atan(Y / sqrt(1.0 + X * X));
	} else if (panel == 4) {
		if (fabs(X) > 2.220446049250313e-16) {
			lon = atan2(X, -Y);
		} else if (Y <= 0.0) {
			lon = 0.0;
		} else {
			lon = TB_PI;
		}
		lat = 0.5 * TB_PI - atan(sqrt(X * X + Y * Y));
	} else {
		if (fabs(X) > 2.220446049250313e-16) {
			lon = atan2(X, Y);
		} else if (Y > 0.0) {
			lon = 0.0;
		} else {
			lon = TB_PI;
		}
		lat = -0.5 * TB_PI + atan(sqrt(X * X + Y * Y));
	}
	if (lon < 0.0) {
		lon += 2.0 * TB_PI;
	}
}

struct CsGeomOut {
	double * j2d;
	double * a0; double * a1;   // ContraMetric2DA
	double * b0; double * b1;   // ContraMetric2DB
	double * coriolis;
	double * lon; double * lat;
};

// one thread per node of the elements [elem0, elem0 + nelem) of one patch
__global__ void k_cs_geometry_2d(
	int nn, long long elem0, long long nelem, int panel, double radius, double omega,
	const double * tx, const double * ty, CsGeomOut o
) {
	const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (idx >= nelem * nn) return;
	const size_t q = (size_t)elem0 * nn + idx;
	const double X = tx[q], Y = ty[q];
	const double a = radius;
	const double d2 = 1.0 + X * X + Y * Y;
	const double d = sqrt(d2);
	double j2d = (1.0 + X * X) * (1.0 + Y * Y) / (d * d * d);
	j2d = j2d * (a * a);
	const double scale = d2 / (1.0 + X * X) / (1.0 + Y * Y) / (a * a);
	double lon, lat;
	tb_rll_from_xyp(X, Y, panel, lon, lat);
	o.j2d[q] = j2d;
	o.a0[q] = scale * (1.0 + Y * Y);
	o.a1[q] = scale * X * Y;
	o.b0[q] = scale * X * Y;
	o.b1[q] = scale * (1.0 + X * X);
	o.coriolis[q] = 2.0 * omega * sin(lat);
	o.lon[q] = lon;
	o.lat[q] = lat;
}

// ---- Jablonowski-Williamson baroclinic wave --------------------------------------

struct JWParams {
	double eta0, tropopause_eta, t0, delta_t, lapse_rate, u0, up;
	double pert_lon, pert_lat, pert_r;
	int perturbation;            // 1: exponential bump in the zonal wind
	double g, R, p0, omega, radius, pressure_scaling, gamma;
	double ztop;
};

__device__ __forceinline__ void tb_jw_geopotential_temperature(
	const JWParams & P, double dEta, double dLat, double & dGeopotential, double & dTemperature
) {
	const double dAuxEta = 0.5 * TB_PI * (dEta - P.eta0);
	double dAvgTemperature = P.t0 * pow(dEta, P.R * P.lapse_rate / P.g);
	if (dEta < P.tropopause_eta) {
		dAvgTemperature += P.delta_t * pow(P.tropopause_eta - dEta, 5.0);
	}
	const double dSinLat = sin(dLat);
	const double dSinLat2 = dSinLat * dSinLat;
	const double dSinLat3 = dSinLat * dSinLat2;
	const double dSinLat4 = dSinLat * dSinLat3;
	const double dSinLat5 = dSinLat * dSinLat4;
	const double dSinLat6 = dSinLat * dSinLat5;
	const double dCosLat = cos(dLat);
	const double dCosLat2 = dCosLat * dCosLat;
	const double dCosLat3 = dCosLat * dCosLat2;
	const double dRefProfile1 =
		P.u0 * pow(cos(dAuxEta), 1.5)
			* (-2.0 * dSinLat6 * (dCosLat2 + 1.0 / 3.0) + 10.0 / 63.0);
	const double dRefProfile2 =
		P.radius * P.omega
			* (8.0 / 5.0 * dCosLat3 * (dSinLat2 + 2.0 / 3.0) - 0.25 * TB_PI);
	dTemperature = 2.0 * dRefProfile1 + dRefProfile2;
	dTemperature =
		dAvgTemperature
		+ 0.75 * dEta * TB_PI * P.u0 / P.R
			* sin(dAuxEta) * sqrt(cos(dAuxEta)) * dTemperature;
	double dAvgGeopotential =
		P.t0 * P.g / P.lapse_rate
			* (1.0 - pow(dEta, P.R * P.lapse_rate / P.g));
	if (dEta < P.tropopause_eta) {
		const double dEta2 = dEta * dEta;
		const double dEta3 = dEta * dEta2;
		const double dEta4 = dEta * dEta3;
		const double dEta5 = dEta * dEta4;
		const double dTropoEta = P.tropopause_eta;
		const double dTropoEta2 = dTropoEta * dTropoEta;
		const double dTropoEta3 = dTropoEta * dTropoEta2;
		const double dTropoEta4 = dTropoEta * dTropoEta3;
		const double dTropoEta5 = dTropoEta * dTropoEta4;
		dAvgGeopotential -= P.R * P.delta_t * (
			(log(dEta / P.tropopause_eta) + 137.0 / 60.0) * dTropoEta5
			- 5.0 * dTropoEta4 * dEta
			+ 5.0 * dTropoEta3 * dEta2
			- (10.0 / 3.0) * dTropoEta2 * dEta3
			+ 5.0 / 4.0 * dTropoEta * dEta4
			- 1.0 / 5.0 * dEta5);
	}
	dGeopotential = dAvgGeopotential
		+ P.u0 * pow(cos(dAuxEta), 1.5) * (dRefProfile1 + dRefProfile2);
}

// surface height of the test (EvaluateTopography), one thread per node
__global__ void k_jw_topography(
	int nn, long long elem0, long long nelem, JWParams P, const double * lat, double * zs
) {
	const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (idx >= nelem * nn) return;
	const size_t q = (size_t)elem0 * nn + idx;
	const double dLat = lat[q];
	const double dAuxEta = 0.5 * TB_PI * (1.0 - P.eta0);
	const double dSinLat = sin(dLat);
	const double dSinLat2 = dSinLat * dSinLat;
	const double dSinLat6 = dSinLat2 * dSinLat * dSinLat * dSinLat * dSinLat;
	const double dCosLat = cos(dLat);
	const double dCosLat2 = dCosLat * dCosLat;
	const double dCosLat3 = dCosLat * dCosLat2;
	const double dRefProfile1 =
		P.u0 * pow(cos(dAuxEta), 1.5)
			* (-2.0 * dSinLat6 * (dCosLat2 + 1.0 / 3.0) + 10.0 / 63.0);
	const double dRefProfile2 =
		P.radius * P.omega
			* (8.0 / 5.0 * dCosLat3 * (dSinLat2 + 2.0 / 3.0) - 0.25 * TB_PI);
	const double dSurfGeopotential = P.u0 * pow(cos(dAuxEta), 1.5) * (dRefProfile1 + dRefProfile2);
	zs[q] = dSurfGeopotential / P.g;
}

// Initial state of the elements [elem0, elem0 + nelem) of one patch, written in
// the device layout: covariant u_alpha, u_beta, rho theta, rho on levels, w = 0
// on interfaces.  One thread per (element, level, node); info[0] receives a
// non-zero value when the Newton iteration for eta does not converge (the
// reference throws "Maximum number of iterations exceeded.").
__global__ void k_jw_state(
	DevLayout lay, long long elem0, long long nelem, int panel, JWParams P,
	const double * tx, const double * ty, const double * lon_, const double * lat_,
	const double * zs_, const double * reta_n, double * data, int * fail
) {
	const int NN = lay.nn;
	const int L = lay.nlev;
	const long long total = nelem * (long long)(L + 1) * NN;
	for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	     idx < total; idx += (long long)gridDim.x * blockDim.x
	) {
		const long long el = idx / ((long long)(L + 1) * NN);
		const int r = (int)(idx % ((long long)(L + 1) * NN));
		const int k = r / NN;
		const int n = r % NN;
		const long long e = elem0 + el;
		const size_t ebase = (size_t)e * lay.nrows * NN;
		// w on interfaces: dState[3] = 0
		data[ebase + (size_t)(lay.rowoff[3] + k) * NN + n] = 0.0;
		if (k >= L) continue;
		const size_t q = (size_t)e * NN + n;
		const double X = tx[q], Y = ty[q];
		const double dLon = lon_[q], dLat = lat_[q];
		const double zs = zs_[q];
		const double dZ = zs + reta_n[k] * (P.ztop - zs);

		// EtaFromRLL
		double dEta = 1.0e-7;
		double dGeopotential = 0.0, dTemperature = 0.0;
		bool converged = false;
		for (int it = 0; it < 25; it++) {
			tb_jw_geopotential_temperature(P, dEta, dLat, dGeopotential, dTemperature);
			const double dF = - P.g * dZ + dGeopotential;
			const double dDiffF = - P.R / dEta * dTemperature;
			const double dNewEta = dEta - dF / dDiffF;
			const bool done = (fabs(dEta - dNewEta) < 1.0e-13);
			dEta = dNewEta;
			if (done) { converged = true; break; }
		}
		if (!converged) atomicMax(fail, 1);

		// EvaluateReferenceState + perturbation
		double dUlon =
			P.u0 * pow(cos(0.5 * TB_PI * (dEta - P.eta0)), 1.5)
				* sin(2.0 * dLat) * sin(2.0 * dLat);
		const double dPressure = P.p0 * dEta;
		const double dRho = dPressure / (P.R * dTemperature);
		// PhysicalConstants::RhoThetaFromPressure (PhysicalConstants.h:389-391)
		const double dRhoTheta = exp(log(dPressure / P.pressure_scaling) / P.gamma);
		if (P.perturbation) {
			double dGreatCircleR =
				acos(sin(P.pert_lat) * sin(dLat)
					+ cos(P.pert_lat) * cos(dLat) * cos(dLon - P.pert_lon));
			dGreatCircleR /= P.pert_r;
			if (dGreatCircleR < 1.0) {
				dUlon += P.up * exp( - dGreatCircleR * dGreatCircleR);
			}
		}
		const double dTheta = dRhoTheta / dRho;

		// zonal / meridional wind (times the radius) -> covariant components
		const double ulon = dUlon * P.radius, ulat = 0.0 * P.radius;
		const double d2 = 1.0 + X * X + Y * Y;
		double ua, ub;
		if (panel < 4) {
			const double lat = atan(Y / sqrt(1.0 + X * X));
			const double ul = ulon / cos(lat);
			ua = (1.0 + X * X) / d2 * ul - X * Y * sqrt(1.0 + X * X) / d2 * ulat;
			ub = sqrt(1.0 + X * X) * (1.0 + Y * Y) / d2 * ulat;
		} else {
			const double rad = sqrt(X * X + Y * Y);
			const bool pole = (fabs(X) < 1.0e-13) && (fabs(Y) < 1.0e-13);
			const double rs = pole ? 1.0 : rad;
			const double sgn = (panel == 4) ? 1.0 : -1.0;
			const double lat = sgn * (0.5 * TB_PI - atan(rad));
			const double cl = pole ? 1.0 : cos(lat);
			const double ul = ulon / cl;
			ua = sgn * (-Y * (1.0 + X * X) / d2 * ul - X * (1.0 + X * X) / (d2 * rs) * ulat);
			ub = sgn * (+X * (1.0 + Y * Y) / d2 * ul - Y * (1.0 + Y * Y) / (d2 * rs) * ulat);
			if (pole) {
				ua = sgn * ulon;
				ub = ulat;
			}
		}
		data[ebase + (size_t)(lay.rowoff[0] + k) * NN + n] = ua;
		data[ebase + (size_t)(lay.rowoff[1] + k) * NN + n] = ub;
		// EquationSet::ConvertComponents: theta -> rho theta
		data[ebase + (size_t)(lay.rowoff[2] + k) * NN + n] = dTheta * dRho;
		data[ebase + (size_t)(lay.rowoff[4] + k) * NN + n] = dRho;
	}
}

#endif
