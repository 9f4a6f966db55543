///////////////////////////////////////////////////////////////////////////////
///
///	\file    ref_dump.cpp
///
///	Dump-hook driver for the parity oracle.  TEST INFRASTRUCTURE ONLY.
///
///	Links the UNMODIFIED reference objects (oracle/_ref/libtempestref.a) and
///	calls the reference plugin methods one at a time (all public:
///	HorizontalDynamics.h:57-167, VerticalDynamics.h:53-126, Grid.h:228,548-566,
///	TimestepScheme.h:55-117), writing the raw arrays after every operation to
///	one binary file that tests/ read with tests/refdump.py.
///
///	usage: ref_dump --case (sw2|jw|bubble) --out FILE --npatch N
///	                --script "op;op;..."  [reference command-line flags]
///
///	script ops (instances are the reference's state-instance indices):
///	  dump:TAG[,I,J..]      write every (or the listed) state/tracer instance
///	  hexp:IN,OUT,DT        HorizontalDynamics::StepExplicit
///	  vexp:IN,OUT,DT        VerticalDynamics::StepExplicit
///	  vimp:IN,OUT,DT        VerticalDynamics::StepImplicit
///	  dss:INST              Grid::PostProcessSubstage(INST, State) (+Tracers)
///	  hasc:IN,OUT,WORK,DT   HorizontalDynamics::StepAfterSubCycle
///	  copy:SRC,DST          Grid::CopyData (State and Tracers)
///	  lincomb:DST,c0,c1,..  Grid::LinearCombineData (State and Tracers)
///	  step:N                N calls of TimestepScheme::Step (first = first call)
///	  hs:SECONDS            HeldSuarezPhysics::Perform with that forcing interval
///	                        (a WorkflowProcess: acts on instance 0)
///	  interp:TAG,NLON,NLAT,NZ,PRIM  Grid::ReduceInterpolate of instance 0 (state at
///	                        both locations, tracers) to an NLON x NLAT latitude-longitude
///	                        grid and NZ uniform REta levels, as OutputManagerReference
///	                        does (OutputManagerReference.cpp:180-222, 585-612);
///	                        PRIM = fConvertToPrimitive
///	  checksum:TAG          Grid::Checksum of instance 0 -> record
///	  addw:INST,AMP         test data: add a smooth non-zero W on interfaces
///	  perturb:INST,EPS      test data: relative pseudo-random noise of size EPS
///	  energy:TAG,INST       Grid::ComputeTotalEnergy / ComputeTotalPotentialEnstrophy /
///	                        ComputeTotalVerticalMomentum of INST -> record, after
///	                        refreshing the slots those routines read at the location
///	                        the state does not live on (W on levels, rho on
///	                        interfaces) with the reference's own interpolation
///	(--nogeometry 1 leaves the 3-D metric arrays out of the file: large grids)
///
///	The test-case classes live in the reference's driver sources next to a
///	main(); they are included (not copied) with main renamed.
///
///////////////////////////////////////////////////////////////////////////////

#define main SWTest2_reference_main
#include "shallowwater_sphere/SWTest2.cpp"
#undef main
#define main BaroclinicWaveJW_reference_main
#include "nonhydro_sphere/BaroclinicWaveJWTest.cpp"
#undef main
#define main ThermalBubble_reference_main
#include "nonhydro_xz/ThermalBubbleCartesianTest.cpp"
#undef main

#include "GridCSGLL.h"
#include "GridCartesianGLL.h"
#include "GridPatchGLL.h"
#include "HorizontalDynamicsFEM.h"
#include "VerticalDynamicsFEM.h"
#include "HeldSuarezPhysics.h"
#include "LinearColumnOperatorFEM.h"

#include <cstdio>
#include <cstdint>
#include <sstream>
#include <vector>
#include <string>

///////////////////////////////////////////////////////////////////////////////

///	<summary>
///		Test data: the reference's Jablonowski-Williamson case carrying
///		analytic tracer densities (a smooth field and two narrow cosine bells
///		with zero background, which the transport drives negative so that the
///		positive-definite filters act).
///	</summary>
class JWTracerTest : public BaroclinicWaveJWTest {
public:
	JWTracerTest(
		double dAlpha, double dZtop, PerturbationType ePert, int nTracers,
		double dRayleigh, double dDiffS = 0.0, double dDiffV = 0.0
	) :
		BaroclinicWaveJWTest(dAlpha, dZtop, ePert),
		m_nTracers(nTracers),
		m_dRayleigh(dRayleigh),
		m_dSpongeTop(dZtop),
		m_dDiffS(dDiffS),
		m_dDiffV(dDiffV)
	{ }

	///	<summary>
	///		Test data: uniform diffusion coefficients (TestCase.h:78-84; the
	///		Cartesian cases of test/nonhydro_xz set them to 75 or 300 m^2/s).
	///	</summary>
	virtual void GetUniformDiffusionCoeffs(
		double & dScalarUniformDiffusionCoeff,
		double & dVectorUniformDiffusionCoeff
	) const {
		dScalarUniformDiffusionCoeff = m_dDiffS;
		dVectorUniformDiffusionCoeff = m_dDiffV;
	}

	///	<summary>
	///		Test data: a sponge layer in the upper 40 % of the domain whose
	///		strength also varies horizontally.
	///	</summary>
	virtual bool HasRayleighFriction() const {
		return (m_dRayleigh > 0.0);
	}

	virtual double EvaluateRayleighStrength(
		double dZ, double dLon, double dLat
	) const {
		const double dZ0 = 0.6 * m_dSpongeTop;
		if (dZ <= dZ0) {
			return 0.0;
		}
		const double dS = sin(0.5 * M_PI * (dZ - dZ0) / (m_dSpongeTop - dZ0));
		return m_dRayleigh * dS * dS * (1.0 + 0.3 * cos(dLon) * cos(dLat));
	}

	virtual void EvaluatePointwiseState(
		const PhysicalConstants & phys,
		const Time & time,
		double dZ, double dLon, double dLat,
		double * dState, double * dTracer
	) const {
		BaroclinicWaveJWTest::EvaluatePointwiseState(
			phys, time, dZ, dLon, dLat, dState, dTracer);
		const double dRho = dState[4];
		for (int c = 0; c < m_nTracers; c++) {
			double dQ;
			if (c == 0) {
				dQ = 1.0e-3 * (2.0 + sin(dLon) * cos(dLat))
					* exp(- dZ / 8000.0);
			} else {
				// cosine bell centred at (lon0, lat0, z0)
				const double dLon0 = 0.5 + 1.7 * c;
				const double dLat0 = 0.6 - 0.5 * c;
				const double dZ0 = 4000.0 + 3000.0 * c;
				const double dR = acos(
					sin(dLat0) * sin(dLat)
					+ cos(dLat0) * cos(dLat) * cos(dLon - dLon0));
				const double dRz = fabs(dZ - dZ0) / 6000.0;
				const double dD = sqrt(dR * dR / (0.9 * 0.9) + dRz * dRz);
				dQ = (dD < 1.0) ? 0.5e-2 * (1.0 + cos(M_PI * dD)) : 0.0;
			}
			dTracer[c] = dRho * dQ;
		}
	}

private:
	int m_nTracers;
	double m_dRayleigh;
	double m_dSpongeTop;
	double m_dDiffS;
	double m_dDiffV;
};

///	<summary>
///		Test data: Williamson 2 carrying analytic tracer densities (a smooth
///		field and a cosine bell), for the tracer transport of
///		HorizontalDynamicsFEM::StepShallowWater (:449-453, 612-640).
///	</summary>
class SWTracerTest : public ShallowWaterTestCase2 {
public:
	SWTracerTest(double dH0, double dU0, double dAlpha, int nTracers) :
		ShallowWaterTestCase2(dH0, dU0, dAlpha),
		m_nTracers(nTracers)
	{ }

	virtual void EvaluatePointwiseState(
		const PhysicalConstants & phys,
		const Time & time,
		double dZ, double dLon, double dLat,
		double * dState, double * dTracer
	) const {
		ShallowWaterTestCase2::EvaluatePointwiseState(
			phys, time, dZ, dLon, dLat, dState, dTracer);
		for (int c = 0; c < m_nTracers; c++) {
			if (c == 0) {
				dTracer[c] = dState[2] * 1.0e-3 * (2.0 + sin(dLon) * cos(dLat));
			} else {
				const double dLon0 = 0.5 + 1.7 * c;
				const double dLat0 = 0.6 - 0.5 * c;
				const double dR = acos(
					sin(dLat0) * sin(dLat)
					+ cos(dLat0) * cos(dLat) * cos(dLon - dLon0));
				dTracer[c] = (dR < 0.9)
					? dState[2] * 0.5e-2 * (1.0 + cos(M_PI * dR / 0.9)) : 0.0;
			}
		}
	}

private:
	int m_nTracers;
};

///	<summary>
///		Test data: the reference's thermal bubble with uniform diffusion
///		switched on (its own coefficients are zero,
///		ThermalBubbleCartesianTest.cpp:144-150).
///	</summary>
class BubbleDiffusionTest : public ThermalBubbleCartesianTest {
public:
	BubbleDiffusionTest(double dDiffS, double dDiffV) :
		ThermalBubbleCartesianTest(300.0, 0.5, 250.0, 500.0, 350.0, 3.14159265),
		m_dDiffS(dDiffS),
		m_dDiffV(dDiffV)
	{ }

	virtual void GetUniformDiffusionCoeffs(
		double & dScalarUniformDiffusionCoeff,
		double & dVectorUniformDiffusionCoeff
	) const {
		dScalarUniformDiffusionCoeff = m_dDiffS;
		dVectorUniformDiffusionCoeff = m_dDiffV;
	}

private:
	double m_dDiffS;
	double m_dDiffV;
};

///////////////////////////////////////////////////////////////////////////////

static FILE * g_fp = NULL;

// --vmassfluxlevels (recorded with the geometry)
static int g_nMassFluxLevels = 0;

static void WriteRecord(
	const std::string & strName,
	int iType, // 0 = double, 1 = int32
	const std::vector<uint64_t> & vecDims,
	const void * pData
) {
	uint32_t nName = strName.size();
	uint32_t nType = iType;
	uint32_t nDim = vecDims.size();
	fwrite(&nName, 4, 1, g_fp);
	fwrite(strName.c_str(), 1, nName, g_fp);
	fwrite(&nType, 4, 1, g_fp);
	fwrite(&nDim, 4, 1, g_fp);
	uint64_t nTotal = 1;
	for (size_t d = 0; d < vecDims.size(); d++) {
		fwrite(&(vecDims[d]), 8, 1, g_fp);
		nTotal *= vecDims[d];
	}
	fwrite(pData, (iType == 0) ? 8 : 4, nTotal, g_fp);
}

static void WriteScalarD(const std::string & strName, double d) {
	std::vector<uint64_t> dims(1, 1);
	WriteRecord(strName, 0, dims, &d);
}

static void WriteScalarI(const std::string & strName, int i) {
	std::vector<uint64_t> dims(1, 1);
	WriteRecord(strName, 1, dims, &i);
}

static void Write1D(const std::string & strName, const DataArray1D<double> & a) {
	std::vector<uint64_t> dims(1, a.GetRows());
	WriteRecord(strName, 0, dims, &(a[0]));
}

static void Write1I(const std::string & strName, const DataArray1D<int> & a) {
	std::vector<uint64_t> dims(1, a.GetRows());
	WriteRecord(strName, 1, dims, &(a[0]));
}

static void Write2D(const std::string & strName, const DataArray2D<double> & a) {
	std::vector<uint64_t> dims(2);
	dims[0] = a.GetRows(); dims[1] = a.GetColumns();
	if (dims[0] * dims[1] == 0) return;
	WriteRecord(strName, 0, dims, &(a[0][0]));
}

static void Write3D(const std::string & strName, const DataArray3D<double> & a) {
	std::vector<uint64_t> dims(3);
	dims[0] = a.GetSize(0); dims[1] = a.GetSize(1); dims[2] = a.GetSize(2);
	if (dims[0] * dims[1] * dims[2] == 0) return;
	WriteRecord(strName, 0, dims, &(a[0][0][0]));
}

static void Write4D(const std::string & strName, const DataArray4D<double> & a) {
	std::vector<uint64_t> dims(4);
	dims[0] = a.GetSize(0); dims[1] = a.GetSize(1);
	dims[2] = a.GetSize(2); dims[3] = a.GetSize(3);
	if (dims[0] * dims[1] * dims[2] * dims[3] == 0) return;
	WriteRecord(strName, 0, dims, &(a[0][0][0][0]));
}

static void WriteOp(const std::string & strName, const LinearColumnOperator & op) {
	if (op.GetCoeffs().GetRows() == 0) return;
	Write2D(strName + ".coeff", op.GetCoeffs());
	Write1I(strName + ".begin", op.GetIxBegin());
	Write1I(strName + ".end", op.GetIxEnd());
}

static std::string P(int n, const char * sz) {
	char buf[64];
	snprintf(buf, 64, "patch%d.", n);
	return std::string(buf) + sz;
}

///////////////////////////////////////////////////////////////////////////////

static void DumpGeometry(Model & model, bool fArrays3D) {
	GridGLL * pGrid = dynamic_cast<GridGLL*>(model.GetGrid());
	const PhysicalConstants & phys = model.GetPhysicalConstants();
	const EquationSet & eqn = model.GetEquationSet();

	WriteScalarI("grid.npatch", pGrid->GetActivePatchCount());
	WriteScalarI("grid.np", pGrid->GetHorizontalOrder());
	WriteScalarI("grid.vertorder", pGrid->GetVerticalOrder());
	WriteScalarI("grid.nlev", pGrid->GetRElements());
	WriteScalarI("grid.ncomp", eqn.GetComponents());
	WriteScalarI("grid.ntracers", eqn.GetTracers());
	WriteScalarI("grid.eqntype", (int)eqn.GetType());
	WriteScalarI("grid.ninstances", model.GetComponentDataInstances());
	WriteScalarI("grid.ntracerinstances", model.GetTracerDataInstances());
	WriteScalarI("grid.xz", pGrid->GetIsCartesianXZ() ? 1 : 0);
	WriteScalarI("grid.iscartesian",
		(dynamic_cast<GridCartesianGLL*>(pGrid) != NULL) ? 1 : 0);
	WriteScalarD("grid.ztop", pGrid->GetZtop());
	WriteScalarD("grid.reflength", pGrid->GetReferenceLength());
	WriteScalarI("grid.vdisc_fv",
		(pGrid->GetVerticalDiscretization() == Grid::VerticalDiscretization_FiniteVolume) ? 1 : 0);
	WriteScalarD("grid.diffs",
		pGrid->HasUniformDiffusion() ? pGrid->GetScalarUniformDiffusionCoeff() : 0.0);
	WriteScalarD("grid.diffv",
		pGrid->HasUniformDiffusion() ? pGrid->GetVectorUniformDiffusionCoeff() : 0.0);
	{
		DataArray1D<int> loc(eqn.GetComponents());
		for (int c = 0; c < eqn.GetComponents(); c++) {
			loc[c] = (pGrid->GetVarLocation(c) == DataLocation_REdge) ? 1 : 0;
		}
		Write1I("grid.varloc", loc);
	}
	Write1D("grid.retalevels", pGrid->GetREtaLevels());
	Write1D("grid.retainterfaces", pGrid->GetREtaInterfaces());

	WriteScalarD("phys.g", phys.GetG());
	WriteScalarD("phys.R", phys.GetR());
	WriteScalarD("phys.cp", phys.GetCp());
	WriteScalarD("phys.cv", phys.GetCv());
	WriteScalarD("phys.p0", phys.GetP0());
	WriteScalarD("phys.omega", phys.GetOmega());
	WriteScalarD("phys.radius", phys.GetEarthRadius());

	Write2D("table.dxbasis1d", pGrid->GetDxBasis1D());
	Write2D("table.stiffness1d", pGrid->GetStiffness1D());
	Write1D("table.gllweights1d", pGrid->GetGLLWeights1D());

	WriteOp("op.interp_n2e", pGrid->GetOpInterpNodeToREdge());
	WriteOp("op.interp_e2n", pGrid->GetOpInterpREdgeToNode());
	WriteOp("op.diff_n2n", pGrid->GetOpDiffNodeToNode());
	WriteOp("op.diff_n2e", pGrid->GetOpDiffNodeToREdge());
	WriteOp("op.diff_e2n", pGrid->GetOpDiffREdgeToNode());
	WriteOp("op.diff_e2e", pGrid->GetOpDiffREdgeToREdge());
	WriteOp("op.diffdiff_n2n", pGrid->GetOpDiffDiffNodeToNode());
	WriteOp("op.diffdiff_e2e", pGrid->GetOpDiffDiffREdgeToREdge());
	WriteOp("op.penalty_left", pGrid->GetOpPenaltyNodeToNode().GetLeftOp());
	WriteOp("op.penalty_right", pGrid->GetOpPenaltyNodeToNode().GetRightOp());
	{
		// m_opDiffNodeToNodeZeroBoundaries (GridGLL.h:413, used by BuildF under
		// --vmassfluxlevels) has no accessor: recover its coefficients by applying
		// GridGLL::DifferentiateNodeToNode(., ., true) to the unit vectors
		const int nL = pGrid->GetRElements();
		DataArray2D<double> dC(nL, nL);
		DataArray1D<int> iBegin(nL);
		DataArray1D<int> iEnd(nL);
		DataArray1D<double> dIn(nL);
		DataArray1D<double> dOut(nL);
		for (int l = 0; l < nL; l++) {
			dIn.Zero();
			dOut.Zero();
			dIn[l] = 1.0;
			pGrid->DifferentiateNodeToNode(&(dIn[0]), &(dOut[0]), true);
			for (int k = 0; k < nL; k++) {
				dC[k][l] = dOut[k];
			}
		}
		for (int k = 0; k < nL; k++) {
			iBegin[k] = 0;
			iEnd[k] = 0;
			bool fAny = false;
			for (int l = 0; l < nL; l++) {
				if (dC[k][l] != 0.0) {
					if (!fAny) iBegin[k] = l;
					iEnd[k] = l + 1;
					fAny = true;
				}
			}
		}
		Write2D("op.diff_n2n_zb.coeff", dC);
		Write1I("op.diff_n2n_zb.begin", iBegin);
		Write1I("op.diff_n2n_zb.end", iEnd);
		WriteScalarI("grid.massfluxlevels", g_nMassFluxLevels);
	}

	for (int n = 0; n < pGrid->GetActivePatchCount(); n++) {
		GridPatchGLL * pPatch =
			dynamic_cast<GridPatchGLL*>(pGrid->GetActivePatch(n));
		const PatchBox & box = pPatch->GetPatchBox();

		WriteScalarI(P(n, "index"), pPatch->GetPatchIndex());
		WriteScalarI(P(n, "panel"), box.GetPanel());
		WriteScalarI(P(n, "halo"), box.GetHaloElements());
		WriteScalarI(P(n, "nelem_a"), pPatch->GetElementCountA());
		WriteScalarI(P(n, "nelem_b"), pPatch->GetElementCountB());
		WriteScalarI(P(n, "a_global_begin"), box.GetAGlobalInteriorBegin());
		WriteScalarI(P(n, "b_global_begin"), box.GetBGlobalInteriorBegin());
		WriteScalarD(P(n, "delta_a"), pPatch->GetElementDeltaA());
		WriteScalarD(P(n, "delta_b"), pPatch->GetElementDeltaB());
		const bool fCartesian = (dynamic_cast<GridCartesianGLL*>(pGrid) != NULL);
		if (!fCartesian) {
			DataArray1D<int> nb(8);
			for (int d = 0; d < 8; d++) {
				nb[d] = pPatch->GetNeighborPanel((Direction)d);
			}
			Write1I(P(n, "neighbor_panels"), nb);
		}
		Write1D(P(n, "anode"), pPatch->GetANodes());
		Write1D(P(n, "bnode"), pPatch->GetBNodes());
		if (!fCartesian) {
			// m_dXNode / m_dYNode (GridPatchCSGLL.cpp:205-213) are protected:
			// same expression, same libm
			DataArray1D<double> dX(pPatch->GetANodes().GetRows());
			DataArray1D<double> dY(pPatch->GetBNodes().GetRows());
			for (int i = 0; i < dX.GetRows(); i++) dX[i] = tan(pPatch->GetANode(i));
			for (int j = 0; j < dY.GetRows(); j++) dY[j] = tan(pPatch->GetBNode(j));
			Write1D(P(n, "xnode"), dX);
			Write1D(P(n, "ynode"), dY);
		}
		Write3D(P(n, "topographyderiv"), pPatch->GetTopographyDeriv());
		Write2D(P(n, "lon"), pPatch->GetLongitude());
		Write2D(P(n, "lat"), pPatch->GetLatitude());
		Write2D(P(n, "jacobian2d"), pPatch->GetJacobian2D());
		Write3D(P(n, "contrametric2da"), pPatch->GetContraMetric2DA());
		Write3D(P(n, "contrametric2db"), pPatch->GetContraMetric2DB());
		Write2D(P(n, "coriolis"), pPatch->GetCoriolisF());
		Write2D(P(n, "topography"), pPatch->GetTopography());
		if (!fArrays3D) continue;
		Write3D(P(n, "jacobian"), pPatch->GetJacobian());
		Write3D(P(n, "jacobianredge"), pPatch->GetJacobianREdge());
		Write4D(P(n, "contrametrica"), pPatch->GetContraMetricA());
		Write4D(P(n, "contrametricb"), pPatch->GetContraMetricB());
		Write4D(P(n, "contrametricxi"), pPatch->GetContraMetricXi());
		Write4D(P(n, "contrametricaredge"), pPatch->GetContraMetricAREdge());
		Write4D(P(n, "contrametricbredge"), pPatch->GetContraMetricBREdge());
		Write4D(P(n, "contrametricxiredge"), pPatch->GetContraMetricXiREdge());
		Write4D(P(n, "derivrnode"), pPatch->GetDerivRNode());
		Write4D(P(n, "derivrredge"), pPatch->GetDerivRREdge());
		Write3D(P(n, "elementareanode"), pPatch->GetElementAreaNode());
		Write3D(P(n, "elementarearedge"), pPatch->GetElementAreaREdge());
		Write3D(P(n, "zlevels"), pPatch->GetZLevels());
		Write3D(P(n, "zinterfaces"), pPatch->GetZInterfaces());
		Write3D(P(n, "rayleighnode"), pPatch->GetRayleighStrength(DataLocation_Node));
		Write3D(P(n, "rayleighredge"), pPatch->GetRayleighStrength(DataLocation_REdge));
		Write4D(P(n, "refstatenode"), pPatch->GetReferenceState(DataLocation_Node));
		Write4D(P(n, "refstateredge"), pPatch->GetReferenceState(DataLocation_REdge));
	}
}

///////////////////////////////////////////////////////////////////////////////

static void DumpState(
	Model & model, const std::string & strTag, int iOnly = -1
) {
	Grid * pGrid = model.GetGrid();
	const EquationSet & eqn = model.GetEquationSet();
	for (int n = 0; n < pGrid->GetActivePatchCount(); n++) {
		GridPatch * pPatch = pGrid->GetActivePatch(n);
		for (int m = 0; m < model.GetComponentDataInstances(); m++) {
			if ((iOnly >= 0) && (m != iOnly)) continue;
			char buf[128];
			snprintf(buf, 128, "%s.patch%d.inst%d.", strTag.c_str(), n, m);
			Write4D(std::string(buf) + "node",
				pPatch->GetDataState(m, DataLocation_Node));
			Write4D(std::string(buf) + "redge",
				pPatch->GetDataState(m, DataLocation_REdge));
		}
		if (eqn.GetTracers() != 0) {
			for (int m = 0; m < model.GetTracerDataInstances(); m++) {
				if ((iOnly >= 0) && (m != iOnly)) continue;
				char buf[128];
				snprintf(buf, 128, "%s.patch%d.inst%d.", strTag.c_str(), n, m);
				Write4D(std::string(buf) + "tracers",
					pPatch->GetDataTracers(m));
			}
		}
	}
}

///////////////////////////////////////////////////////////////////////////////

static std::vector<std::string> Split(const std::string & s, char c) {
	std::vector<std::string> out;
	std::stringstream ss(s);
	std::string item;
	while (std::getline(ss, item, c)) {
		if (item.size() != 0) out.push_back(item);
	}
	return out;
}

static void RunScript(Model & model, const std::string & strScript) {
	Grid * pGrid = model.GetGrid();
	const EquationSet & eqn = model.GetEquationSet();
	HorizontalDynamics * pH = model.GetHorizontalDynamics();
	VerticalDynamics * pV = model.GetVerticalDynamics();
	TimestepScheme * pT = model.GetTimestepScheme();

	Time time = model.GetStartTime();
	bool fFirst = true;

	std::vector<std::string> vecOps = Split(strScript, ';');
	for (size_t o = 0; o < vecOps.size(); o++) {
		std::vector<std::string> kv = Split(vecOps[o], ':');
		const std::string & op = kv[0];
		std::vector<std::string> a;
		if (kv.size() > 1) a = Split(kv[1], ',');

		if (op == "dump") {
			if (a.size() > 1) {
				for (size_t i = 1; i < a.size(); i++) {
					DumpState(model, a[0], atoi(a[i].c_str()));
				}
			} else {
				DumpState(model, a[0]);
			}
		} else if (op == "addw") {
			// Test-data helper: add a smooth, everywhere non-zero vertical
			// velocity to interior interfaces of one instance so that xi-dot
			// is nowhere rounding noise (the reference's Jacobian carries
			// sign(xi-dot), VerticalDynamicsFEM.cpp:2876-2884, which makes
			// columns of exactly zero wind - JW equator and poles - flip
			// with the last bit of their input).
			const int iInst = atoi(a[0].c_str());
			const double dAmp = atof(a[1].c_str());
			const int nL = pGrid->GetRElements();
			for (int n = 0; n < pGrid->GetActivePatchCount(); n++) {
				GridPatch * pPatch = pGrid->GetActivePatch(n);
				DataArray4D<double> & dataREdge =
					pPatch->GetDataState(iInst, DataLocation_REdge);
				const DataArray2D<double> & dLon = pPatch->GetLongitude();
				const DataArray2D<double> & dLat = pPatch->GetLatitude();
				for (int i = 0; i < dataREdge.GetSize(1); i++) {
				for (int j = 0; j < dataREdge.GetSize(2); j++) {
				for (int k = 1; k < nL; k++) {
					dataREdge(3,i,j,k) += dAmp
						* (1.0 + 0.5 * sin(dLon(i,j)) * cos(dLat(i,j)))
						* sin(M_PI * static_cast<double>(k) / static_cast<double>(nL));
				}
				}
				}
			}
		} else if (op == "perturb") {
			// Test-data helper: multiply every state value of one instance by
			// (1 + EPS * r), r a fixed pseudo-random sequence in [-1,1]; used to
			// measure the reference's own sensitivity to last-bit changes.
			const int iInst = atoi(a[0].c_str());
			const double dEps = atof(a[1].c_str());
			unsigned long long u = 88172645463325252ull;
			for (int n = 0; n < pGrid->GetActivePatchCount(); n++) {
				GridPatch * pPatch = pGrid->GetActivePatch(n);
				for (int l = 0; l < 2; l++) {
					DataArray4D<double> & data = pPatch->GetDataState(
						iInst, (l == 0) ? DataLocation_Node : DataLocation_REdge);
					double * p = &(data(0,0,0,0));
					const size_t nTotal = data.GetTotalSize();
					for (size_t q = 0; q < nTotal; q++) {
						u ^= u << 13; u ^= u >> 7; u ^= u << 17;
						const double r = 2.0 * (double)(u >> 11) / 9007199254740992.0 - 1.0;
						p[q] *= (1.0 + dEps * r);
					}
				}
			}
		} else if (op == "hexp") {
			pH->StepExplicit(atoi(a[0].c_str()), atoi(a[1].c_str()), time, atof(a[2].c_str()));
		} else if (op == "vexp") {
			pV->StepExplicit(atoi(a[0].c_str()), atoi(a[1].c_str()), time, atof(a[2].c_str()));
		} else if (op == "vimp") {
			pV->StepImplicit(atoi(a[0].c_str()), atoi(a[1].c_str()), time, atof(a[2].c_str()));
		} else if (op == "dss") {
			pGrid->PostProcessSubstage(atoi(a[0].c_str()), DataType_State);
			pGrid->PostProcessSubstage(atoi(a[0].c_str()), DataType_Tracers);
		} else if (op == "hasc") {
			pH->StepAfterSubCycle(
				atoi(a[0].c_str()), atoi(a[1].c_str()), atoi(a[2].c_str()),
				time, atof(a[3].c_str()));
		} else if (op == "copy") {
			pGrid->CopyData(atoi(a[0].c_str()), atoi(a[1].c_str()), DataType_State);
			pGrid->CopyData(atoi(a[0].c_str()), atoi(a[1].c_str()), DataType_Tracers);
		} else if (op == "lincomb") {
			DataArray1D<double> dCoeff(model.GetComponentDataInstances());
			for (size_t i = 1; i < a.size(); i++) {
				dCoeff[i-1] = atof(a[i].c_str());
			}
			pGrid->LinearCombineData(dCoeff, atoi(a[0].c_str()), DataType_State);
			pGrid->LinearCombineData(dCoeff, atoi(a[0].c_str()), DataType_Tracers);
		} else if (op == "step") {
			int nSteps = atoi(a[0].c_str());
			double dDeltaT = model.GetDeltaT().GetSeconds();
			for (int s = 0; s < nSteps; s++) {
				pT->Step(fFirst, false, time, dDeltaT);
				fFirst = false;
				time += model.GetDeltaT();
			}
		} else if (op == "energy") {
			// GridPatch::ComputeTotalEnergy (GridPatch.cpp:925-1138) reads W on
			// levels and rho on interfaces, which the Lorenz-staggered state
			// does not carry: refresh them first (Grid.cpp:843-863)
			const int iInst = atoi(a[1].c_str());
			if (eqn.GetType() == EquationSet::PrimitiveNonhydrostaticEquations) {
				if (pGrid->GetVarLocation(3) == DataLocation_REdge) {
					pGrid->InterpolateREdgeToNode(3, iInst);
				}
				if (pGrid->GetVarLocation(4) == DataLocation_Node) {
					pGrid->InterpolateNodeToREdge(4, iInst);
				}
			}
			DataArray1D<double> dDiag(3);
			dDiag[0] = pGrid->ComputeTotalEnergy(iInst);
			dDiag[1] = pGrid->ComputeTotalPotentialEnstrophy(iInst);
			dDiag[2] = (eqn.GetType() == EquationSet::PrimitiveNonhydrostaticEquations) ?
				pGrid->ComputeTotalVerticalMomentum(iInst) : 0.0;
			Write1D(a[0] + ".energy", dDiag);
		} else if (op == "hs") {
			const double dSeconds = atof(a[0].c_str());
			const int iSec = static_cast<int>(dSeconds);
			const int iMicro = static_cast<int>((dSeconds - iSec) * 1.0e6 + 0.5);
			HeldSuarezPhysics hs(model, Time(0, 0, 0, iSec, iMicro, Time::CalendarNoLeap, Time::TypeDelta));
			hs.Perform(time);
		} else if (op == "interp") {
			const int nLon = atoi(a[1].c_str());
			const int nLat = atoi(a[2].c_str());
			const int nZ = atoi(a[3].c_str());
			const bool fPrim = (atoi(a[4].c_str()) != 0);
			const int nPts = nLon * nLat;
			DataArray1D<double> dLonDeg(nPts), dLatDeg(nPts);
			int ix = 0;
			for (int j = 0; j < nLat; j++) {
			for (int i = 0; i < nLon; i++) {
				dLonDeg[ix] = 360.0 * (static_cast<double>(i) + 0.5) / static_cast<double>(nLon);
				dLatDeg[ix] = -90.0 + 180.0 * (static_cast<double>(j) + 0.5) / static_cast<double>(nLat);
				ix++;
			}
			}
			DataArray1D<double> dAlpha(nPts), dBeta(nPts);
			DataArray1D<int> iPatch(nPts);
			pGrid->ConvertReferenceToPatchCoord(dLonDeg, dLatDeg, dAlpha, dBeta, iPatch);
			DataArray1D<double> dREta(nZ);
			for (int k = 0; k < nZ; k++) {
				dREta[k] = (static_cast<double>(k) + 0.5) / static_cast<double>(nZ);
			}
			const int nComp = eqn.GetComponents();
			DataArray3D<double> dNode(nComp, nZ, nPts);
			DataArray3D<double> dREdge(nComp, nZ, nPts);
			pGrid->ReduceInterpolate(
				DataType_State, dREta, dAlpha, dBeta, iPatch, dNode,
				DataLocation_Node, true, fPrim);
			pGrid->ReduceInterpolate(
				DataType_State, dREta, dAlpha, dBeta, iPatch, dREdge,
				DataLocation_REdge, true, fPrim);
			Write1D(a[0] + ".alpha", dAlpha);
			Write1D(a[0] + ".beta", dBeta);
			Write1I(a[0] + ".ipatch", iPatch);
			Write1D(a[0] + ".reta", dREta);
			Write3D(a[0] + ".node", dNode);
			Write3D(a[0] + ".redge", dREdge);
			if (eqn.GetTracers() != 0) {
				DataArray3D<double> dTr(eqn.GetTracers(), nZ, nPts);
				pGrid->ReduceInterpolate(
					DataType_Tracers, dREta, dAlpha, dBeta, iPatch, dTr,
					DataLocation_None, true, fPrim);
				Write3D(a[0] + ".tracers", dTr);
			}
			// derived fields of instance 0 (OutputManagerReference.cpp:640-700)
			{
				pGrid->ComputeVorticityDivergence(0);
				DataArray3D<double> dV(1, nZ, nPts), dD(1, nZ, nPts);
				pGrid->ReduceInterpolate(
					DataType_Vorticity, dREta, dAlpha, dBeta, iPatch, dV);
				pGrid->ReduceInterpolate(
					DataType_Divergence, dREta, dAlpha, dBeta, iPatch, dD);
				Write3D(a[0] + ".vorticity", dV);
				Write3D(a[0] + ".divergence", dD);
				if (eqn.GetType() == EquationSet::PrimitiveNonhydrostaticEquations) {
					pGrid->ComputeTemperature(0);
					DataArray3D<double> dT(1, nZ, nPts);
					pGrid->ReduceInterpolate(
						DataType_Temperature, dREta, dAlpha, dBeta, iPatch, dT);
					Write3D(a[0] + ".temperature", dT);
				}
			}
			// the vertical operators GridPatchCSGLL::InterpolateData builds (:1470-1487)
			GridGLL * pGridGLL = dynamic_cast<GridGLL*>(pGrid);
			LinearColumnInterpFEM opN, opE;
			opN.Initialize(
				LinearColumnInterpFEM::InterpSource_Levels, pGridGLL->GetVerticalOrder(),
				pGrid->GetREtaLevels(), pGrid->GetREtaInterfaces(), dREta);
			opE.Initialize(
				LinearColumnInterpFEM::InterpSource_Interfaces, pGridGLL->GetVerticalOrder(),
				pGrid->GetREtaLevels(), pGrid->GetREtaInterfaces(), dREta);
			WriteOp(a[0] + ".vop_node", opN);
			WriteOp(a[0] + ".vop_redge", opE);
		} else if (op == "checksum") {
			DataArray1D<double> dSums;
			pGrid->Checksum(DataType_State, dSums, 0, ChecksumType_Sum);
			Write1D(a[0] + ".checksum", dSums);
		} else {
			_EXCEPTION1("ref_dump: unknown op \"%s\"", op.c_str());
		}
	}
}

///////////////////////////////////////////////////////////////////////////////

int main(int argc, char ** argv) {

	TempestInitialize(&argc, &argv);

try {
	std::string strCase;
	std::string strOut;
	std::string strScript;
	int nPatch;
	double dZtop;
	std::string strPert;
	double dU0, dH0, dAlpha;
	int nTracers;
	double dRayleigh;
	int nNoGeometry;
	double dDiffS;
	double dDiffV;
	int nXZ;

	BeginTempestCommandLine("RefDump");
		SetDefaultResolution(4);
		SetDefaultResolutionY(1);
		SetDefaultLevels(1);
		SetDefaultOutputDeltaT("200s");
		SetDefaultDeltaT("200s");
		SetDefaultEndTime("0s");
		SetDefaultHorizontalOrder(4);
		SetDefaultVerticalOrder(1);

		CommandLineString(strCase, "case", "sw2");
		CommandLineString(strOut, "out", "ref_dump.bin");
		CommandLineString(strScript, "script", "dump:ic");
		CommandLineInt(nPatch, "npatch", 6);
		CommandLineDouble(dZtop, "ztop", 30000.0);
		CommandLineString(strPert, "pert", "Exp");
		CommandLineDouble(dU0, "u0", 38.61068277);
		CommandLineDouble(dH0, "h0", 2998.104995);
		CommandLineDouble(dAlpha, "alpha", 0.0);
		CommandLineInt(nTracers, "ntracers", 0);
		CommandLineDouble(dRayleigh, "rayleigh", 0.0);
		CommandLineInt(nNoGeometry, "nogeometry", 0);
		CommandLineInt(nXZ, "xz", 1);
		CommandLineDouble(dDiffS, "diffs", 0.0);
		CommandLineDouble(dDiffV, "diffv", 0.0);

		ParseCommandLine(argc, argv);
	EndTempestCommandLine(argv)

	g_nMassFluxLevels = _tempestvars.fForceMassFluxOnLevels ? 1 : 0;

	g_fp = fopen(strOut.c_str(), "wb");
	if (g_fp == NULL) {
		_EXCEPTION1("Cannot open \"%s\"", strOut.c_str());
	}
	fwrite("TB2DUMP1", 1, 8, g_fp);

	Model * pModel;
	TestCase * pTest;

	if (strCase == "sw2") {
		if (nTracers > 0) {
			EquationSet eqn(EquationSet::ShallowWaterEquations);
			for (int c = 0; c < nTracers; c++) {
				char szName[16];
				snprintf(szName, 16, "HQ%d", c);
				eqn.InsertTracer(szName, szName);
			}
			UserDataMeta metaUserData;
			pModel = new Model(eqn, metaUserData);
			pTest = new SWTracerTest(dH0, dU0, dAlpha, nTracers);
		} else {
			pModel = new Model(EquationSet::ShallowWaterEquations);
			pTest = new ShallowWaterTestCase2(dH0, dU0, dAlpha);
		}
	} else if (strCase == "jw") {
		STLStringHelper::ToLower(strPert);
		const BaroclinicWaveJWTest::PerturbationType ePert =
			(strPert == "exp") ?
				BaroclinicWaveJWTest::PerturbationType_Exp :
				BaroclinicWaveJWTest::PerturbationType_None;
		if ((nTracers > 0) || (dRayleigh > 0.0) || (dDiffS != 0.0) || (dDiffV != 0.0)) {
			EquationSet eqn(EquationSet::PrimitiveNonhydrostaticEquations);
			for (int c = 0; c < nTracers; c++) {
				char szName[16];
				snprintf(szName, 16, "RhoQ%d", c);
				eqn.InsertTracer(szName, szName);
			}
			UserDataMeta metaUserData;
			pModel = new Model(eqn, metaUserData);
			pTest = new JWTracerTest(
				dAlpha, dZtop, ePert, nTracers, dRayleigh, dDiffS, dDiffV);
		} else {
			pModel = new Model(EquationSet::PrimitiveNonhydrostaticEquations);
			pTest = new BaroclinicWaveJWTest(dAlpha, dZtop, ePert);
		}
	} else if (strCase == "bubble") {
		pModel = new Model(EquationSet::PrimitiveNonhydrostaticEquations);
		pTest = new BubbleDiffusionTest(dDiffS, dDiffV);
	} else {
		_EXCEPTIONT("--case must be sw2, jw or bubble");
	}
	Model & model = (*pModel);

	if (strCase == "bubble") {
		ThermalBubbleCartesianTest * pBubble =
			dynamic_cast<ThermalBubbleCartesianTest*>(pTest);
		// --xz 0: a three-dimensional periodic box (as
		// ThermalBubbleCartesian3DTest sets its grid up) instead of the x-z slice
		TempestSetupCartesianModel(
			model, pBubble->m_dGDim, 0.0, pBubble->m_iLatBC, (nXZ != 0));
		const double XL = std::abs(pBubble->m_dGDim[1] - pBubble->m_dGDim[0]);
		model.GetGrid()->SetReferenceLength((XL < 110000.0) ? XL : 110000.0);

	} else {
		// Same sequence as _TempestSetupCubedSphereModel
		// (TempestInitialize.h:476-586) with an explicit patch count.
		model.SetDeltaT(_tempestvars.timeDeltaT);
		model.SetEndTime(_tempestvars.timeEndTime);
		_TempestSetupMethodOfLines(model, _tempestvars);

		// --vdisc (TempestInitialize.h:486-497)
		STLStringHelper::ToLower(_tempestvars.strVerticalDiscretization);
		const bool fFiniteVolume = (_tempestvars.strVerticalDiscretization == "fv");

		GridCSGLL * pGrid = new GridCSGLL(model);
		pGrid->DefineParameters();
		pGrid->SetParameters(
			_tempestvars.nLevels,
			nPatch,
			_tempestvars.nResolutionX,
			4,
			_tempestvars.nHorizontalOrder,
			_tempestvars.nVerticalOrder,
			fFiniteVolume
				? Grid::VerticalDiscretization_FiniteVolume
				: Grid::VerticalDiscretization_FiniteElement,
			Grid::VerticalStaggering_Lorenz);
		pGrid->InitializeDataLocal();
		model.SetGrid(pGrid, nPatch);
		_TempestSetupOutputManagers(model, _tempestvars);
	}

	model.SetTestCase(pTest);

	// With end time == start time Model::Go runs exactly its head
	// (Model.cpp:347-366: EvaluateGeometricTerms, ApplyBoundaryConditions,
	// Initialize of scheme / horizontal / vertical dynamics) and returns
	// before the first output and the step loop.
	model.Go();

	WriteScalarD("run.dt", model.GetDeltaT().GetSeconds());
	WriteScalarI("run.hypervisorder", _tempestvars.nHyperviscosityOrder);
	WriteScalarI("run.nohypervis", _tempestvars.fNoHyperviscosity ? 1 : 0);
	WriteScalarD("run.nu_scalar", _tempestvars.dNuScalar);
	WriteScalarD("run.nu_div", _tempestvars.dNuDiv);
	WriteScalarD("run.nu_vort", _tempestvars.dNuVort);

	DumpGeometry(model, nNoGeometry == 0);
	RunScript(model, strScript);

	fclose(g_fp);
	delete pModel;

} catch(Exception & e) {
	std::cout << e.ToString() << std::endl;
	return 1;
}
	TempestDeinitialize();
	return 0;
}
