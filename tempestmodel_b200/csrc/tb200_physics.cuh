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

#endif
