///////////////////////////////////////////////////////////////////////////////
///
///	\file    b200_driver.cpp
///
///	The reference's own driver flow (TempestInitialize.h setup macros, Model::Go,
///	checksum output manager, error norms) with the B200 plugins selected by
///	--b200 (none | plugins | scheme):
///	  none     every plugin is the reference's (CPU)
///	  plugins  HorizontalDynamicsB200 + VerticalDynamicsB200 under the
///	           reference's TimestepScheme and host DSS
///	  scheme   TimestepSchemeB200: the whole step on the device
///	The test-case classes are the reference's (included, main renamed).
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

///	<summary>
///		The reference's thermal bubble with uniform diffusion switched on
///		(--diffs / --diffv; its own coefficients are zero,
///		ThermalBubbleCartesianTest.cpp:144-150 - the other Cartesian cases of
///		test/nonhydro_xz run with 75 or 300 m^2/s).
///	</summary>
class ThermalBubbleDiffusionTest : public ThermalBubbleCartesianTest {
public:
	ThermalBubbleDiffusionTest(double dDiffS, double dDiffV) :
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

#include "TempestB200.h"
#include "HeldSuarezPhysics.h"
#include "GridCSGLL.h"
#include "GridCartesianGLL.h"
#include "VerticalDynamicsStub.h"

int main(int argc, char ** argv) {

	TempestInitialize(&argc, &argv);

try {
	std::string strCase;
	std::string strMode;
	int nPatch;
	double dZtop;
	std::string strPert;
	int nEager;
	int nHeldSuarez;
	double dDiffS;
	double dDiffV;

	BeginTempestCommandLine("B200Driver");
		SetDefaultResolution(8);
		SetDefaultResolutionY(1);
		SetDefaultLevels(10);
		SetDefaultOutputDeltaT("200s");
		SetDefaultDeltaT("200s");
		SetDefaultEndTime("600s");
		SetDefaultHorizontalOrder(4);
		SetDefaultVerticalOrder(1);

		CommandLineString(strCase, "case", "jw");
		CommandLineString(strMode, "b200", "scheme");
		CommandLineInt(nPatch, "npatch", 6);
		CommandLineDouble(dZtop, "ztop", 10000.0);
		CommandLineString(strPert, "pert", "Exp");
		CommandLineInt(nEager, "b200eager", 0);
		CommandLineInt(nHeldSuarez, "heldsuarez", 0);
		CommandLineDouble(dDiffS, "diffs", 0.0);
		CommandLineDouble(dDiffV, "diffv", 0.0);

		ParseCommandLine(argc, argv);
	EndTempestCommandLine(argv)

	const bool fSW = (strCase == "sw2");
	Model model(fSW ? EquationSet::ShallowWaterEquations
	                : EquationSet::PrimitiveNonhydrostaticEquations);

	model.SetDeltaT(_tempestvars.timeDeltaT);
	model.SetEndTime(_tempestvars.timeEndTime);

	STLStringHelper::ToLower(_tempestvars.strTimestepScheme);
	const int iScheme = tb200_scheme_from_name(_tempestvars.strTimestepScheme.c_str());
	if (iScheme < 0) {
		_EXCEPTION1("--timescheme \"%s\" is not implemented by libtempest_b200",
			_tempestvars.strTimestepScheme.c_str());
	}

	TimestepSchemeB200 * pSchemeB200 = NULL;
	if (strMode == "none") {
		_TempestSetupMethodOfLines(model, _tempestvars);

	} else {
		if (strMode == "scheme") {
			TimestepSchemeB200 * pScheme = new TimestepSchemeB200(model, iScheme);
			pSchemeB200 = pScheme;
			// the output managers of _TempestSetupOutputManagers all fire every
			// --outputtime (TempestInitialize.h:413-472): instance 0 only comes
			// back to the host for them (--b200eager 1: every step)
			if (nEager == 0) {
				pScheme->HostReadsEvery(
					model.GetStartTime(), _tempestvars.timeOutputDeltaT);
			}
			model.SetTimestepScheme(pScheme);
		} else if (iScheme == TB200_SCHEME_ARS343) {
			model.SetTimestepScheme(new TimestepSchemeARS343(model));
		} else if (iScheme == TB200_SCHEME_STRANG_KGU35) {
			model.SetTimestepScheme(new TimestepSchemeStrang(model));
		} else {
			_EXCEPTIONT("--mode plugins drives the reference's strang or ars343 scheme");
		}
		// same arguments as the "v1" branches of _TempestSetupMethodOfLines
		// (TempestInitialize.h:296-366)
		if (_tempestvars.fNoHyperviscosity) {
			_tempestvars.nHyperviscosityOrder = 0;
			_tempestvars.dNuScalar = 0.0;
			_tempestvars.dNuDiv = 0.0;
			_tempestvars.dNuVort = 0.0;
		}
		model.SetHorizontalDynamics(
			new HorizontalDynamicsB200(
				model,
				_tempestvars.nHorizontalOrder,
				_tempestvars.nHyperviscosityOrder,
				_tempestvars.dNuScalar,
				_tempestvars.dNuDiv,
				_tempestvars.dNuVort,
				_tempestvars.dInstepNuDiv));
		if (_tempestvars.nLevels == 1) {
			model.SetVerticalDynamics(new VerticalDynamicsStub(model));
		} else {
			model.SetVerticalDynamics(
				new VerticalDynamicsB200(
					model,
					_tempestvars.nHorizontalOrder,
					_tempestvars.nVerticalOrder,
					_tempestvars.nVerticalHyperdiffOrder,
					_tempestvars.fExplicitVertical,
					!_tempestvars.fNoReferenceState,
					_tempestvars.fForceMassFluxOnLevels));
		}
	}

	// --vdisc (TempestInitialize.h:486-497)
	STLStringHelper::ToLower(_tempestvars.strVerticalDiscretization);
	const Grid::VerticalDiscretization eVerticalDiscretization =
		(_tempestvars.strVerticalDiscretization == "fv")
			? Grid::VerticalDiscretization_FiniteVolume
			: Grid::VerticalDiscretization_FiniteElement;

	ThermalBubbleCartesianTest * pBubble = NULL;
	if (strCase == "bubble") {
		// _TempestSetupCartesianModel (TempestInitialize.h:590-706), x-z slice
		pBubble = new ThermalBubbleDiffusionTest(dDiffS, dDiffV);
		GridCartesianGLL * pGrid = new GridCartesianGLL(model);
		pGrid->DefineParameters();
		pGrid->SetParameters(
			_tempestvars.nLevels,
			1,
			_tempestvars.nResolutionX,
			_tempestvars.nResolutionY,
			4,
			_tempestvars.nHorizontalOrder,
			_tempestvars.nVerticalOrder,
			pBubble->m_dGDim,
			0.0,
			pBubble->m_iLatBC,
			true,
			eVerticalDiscretization,
			Grid::VerticalStaggering_Lorenz);
		pGrid->InitializeDataLocal();
		model.SetGrid(pGrid);
		const double XL = std::abs(pBubble->m_dGDim[1] - pBubble->m_dGDim[0]);
		pGrid->SetReferenceLength((XL < 110000.0) ? XL : 110000.0);
	} else {
		// _TempestSetupCubedSphereModel (TempestInitialize.h:476-586)
		GridCSGLL * pGrid = new GridCSGLL(model);
		pGrid->DefineParameters();
		pGrid->SetParameters(
			_tempestvars.nLevels,
			nPatch,
			_tempestvars.nResolutionX,
			4,
			_tempestvars.nHorizontalOrder,
			_tempestvars.nVerticalOrder,
			eVerticalDiscretization,
			Grid::VerticalStaggering_Lorenz);
		pGrid->InitializeDataLocal();
		model.SetGrid(pGrid, nPatch);
	}
	_TempestSetupOutputManagers(model, _tempestvars);

	if (pBubble != NULL) {
		model.SetTestCase(pBubble);
	} else if (fSW) {
		model.SetTestCase(new ShallowWaterTestCase2(2998.104995, 38.61068277, 0.0));
	} else {
		STLStringHelper::ToLower(strPert);
		model.SetTestCase(
			new BaroclinicWaveJWTest(
				0.0, dZtop,
				(strPert == "exp") ?
					BaroclinicWaveJWTest::PerturbationType_Exp :
					BaroclinicWaveJWTest::PerturbationType_None));
	}

	// --heldsuarez 1: Held-Suarez forcing every step (HeldSuarezTest.cpp:373-377),
	// the reference's host process under --b200 none, the device step otherwise
	if ((nHeldSuarez != 0) && (!fSW) && (pBubble == NULL)) {
		if (strMode == "none") {
			model.AttachWorkflowProcess(
				new HeldSuarezPhysics(model, model.GetDeltaT()));
		} else {
			model.AttachWorkflowProcess(
				new HeldSuarezPhysicsB200(model, model.GetDeltaT(), pSchemeB200));
		}
	}

	AnnounceBanner("SIMULATION");
	model.Go();

	AnnounceBanner("RESULTS");
	model.ComputeErrorNorms();
	AnnounceBanner();

} catch(Exception & e) {
	std::cout << e.ToString() << std::endl;
	return 1;
}
	TempestDeinitialize();
	return 0;
}
