///////////////////////////////////////////////////////////////////////////////
///
///	\file    TempestB200.cpp
///
///	Reference-side binding of libtempest_b200 (see TempestB200.h).
///
///////////////////////////////////////////////////////////////////////////////

#include "TempestB200.h"

#include "GridCSGLL.h"
#include "GridCartesianGLL.h"
#include "CubedSphereTrans.h"
#include "EquationSet.h"
#include "PhysicalConstants.h"
#include "Exception.h"

#include <map>
#include <cmath>
#include <cstring>
#include <cstdlib>

///////////////////////////////////////////////////////////////////////////////

static std::map<Model *, B200Bridge *> g_mapBridges;

B200Bridge & B200Bridge::Get(Model & model) {
	std::map<Model *, B200Bridge *>::iterator it = g_mapBridges.find(&model);
	if (it == g_mapBridges.end()) {
		B200Bridge * p = new B200Bridge(model);
		g_mapBridges[&model] = p;
		return *p;
	}
	return *(it->second);
}

B200Bridge::B200Bridge(Model & model) :
	m_model(model),
	m_pCtx(NULL),
	m_fInitialized(false),
	m_nHypervisOrder(0),
	m_dNuScalar(0.0),
	m_dNuDiv(0.0),
	m_dNuVort(0.0),
	m_fFullyExplicit(false),
	m_fUseReferenceState(true),
	m_fMassFluxOnLevels(false)
{ }

void B200Bridge::SetUseReferenceState(bool fUseReferenceState) {
	m_fUseReferenceState = fUseReferenceState;
}

void B200Bridge::SetMassFluxOnLevels(bool fMassFluxOnLevels) {
	m_fMassFluxOnLevels = fMassFluxOnLevels;
}

void B200Bridge::SetFullyExplicit(bool fFullyExplicit) {
	m_fFullyExplicit = fFullyExplicit;
}

void B200Bridge::SetHyperviscosity(
	int nOrder, double dNuScalar, double dNuDiv, double dNuVort
) {
	m_nHypervisOrder = nOrder;
	m_dNuScalar = dNuScalar;
	m_dNuDiv = dNuDiv;
	m_dNuVort = dNuVort;
}

void B200Bridge::Check(int iResult) {
	if (iResult != 0) {
		_EXCEPTION1("tempest_b200: %s", tb200_last_error(m_pCtx));
	}
}

///////////////////////////////////////////////////////////////////////////////
// Global node ids: the integer point of the cube surface a GLL node sits on
// (panel orientation of CubedSphereTrans::XYZFromXYP, CubedSphereTrans.cpp:25-83).

static void LatticePoint(int p, long s, long t, long n, long & x, long & y, long & z) {
	switch (p) {
		case 0: x = n; y = s; z = t; break;
		case 1: x = -s; y = n; z = t; break;
		case 2: x = -n; y = -s; z = t; break;
		case 3: x = s; y = -n; z = t; break;
		case 4: x = -t; y = s; z = n; break;
		default: x = t; y = s; z = -n; break;
	}
}

static long UniqueIndex(long g, int np) {
	return (g / np) * (np - 1) + (g % np);
}

///////////////////////////////////////////////////////////////////////////////

void B200Bridge::Initialize() {
	if (m_fInitialized) {
		return;
	}
	GridGLL * pGrid = dynamic_cast<GridGLL *>(m_model.GetGrid());
	if (pGrid == NULL) {
		_EXCEPTIONT("tempest_b200 requires a GridGLL");
	}
	GridCSGLL * pGridCS = dynamic_cast<GridCSGLL *>(pGrid);
	GridCartesianGLL * pGridCart = dynamic_cast<GridCartesianGLL *>(pGrid);
	if ((pGridCS == NULL) && (pGridCart == NULL)) {
		_EXCEPTIONT("tempest_b200: GridCSGLL or GridCartesianGLL expected");
	}
	if (pGridCart != NULL) {
		// GridPatchCartesianGLL::ApplyBoundaryConditions (no-flux / no-slip
		// walls) is not implemented on the device
		for (int d = 0; d < 4; d++) {
			if (pGrid->GetBoundaryCondition((Direction)d) !=
			    Grid::BoundaryCondition_Periodic
			) {
				_EXCEPTIONT("tempest_b200: only periodic Cartesian boundaries "
					"are implemented");
			}
		}
	}
	const EquationSet & eqn = m_model.GetEquationSet();
	const PhysicalConstants & phys = m_model.GetPhysicalConstants();

	tb200_config cfg;
	memset(&cfg, 0, sizeof(cfg));
	cfg.np = pGrid->GetHorizontalOrder();
	cfg.nlev = pGrid->GetRElements();
	cfg.vertical_order = pGrid->GetVerticalOrder();
	cfg.ncomp = eqn.GetComponents();
	cfg.ntracers = eqn.GetTracers();
	cfg.ninstances = m_model.GetComponentDataInstances();
	cfg.eqn_type =
		(eqn.GetType() == EquationSet::ShallowWaterEquations)
			? TB200_EQN_SHALLOW_WATER : TB200_EQN_PRIMITIVE_NONHYDRO;
	cfg.cartesian_xz = pGrid->GetIsCartesianXZ() ? 1 : 0;
	for (int c = 0; c < cfg.ncomp; c++) {
		cfg.comp_on_redge[c] =
			(pGrid->GetVarLocation(c) == DataLocation_REdge) ? 1 : 0;
	}
	cfg.device = -1;
	cfg.g = phys.GetG();
	cfg.R = phys.GetR();
	cfg.cp = phys.GetCp();
	cfg.cv = phys.GetCv();
	cfg.p0 = phys.GetP0();
	cfg.omega = phys.GetOmega();
	cfg.earth_radius = phys.GetEarthRadius();
	cfg.ztop = pGrid->GetZtop();
	cfg.ref_length = pGrid->GetReferenceLength();
	cfg.hypervis_order = m_nHypervisOrder;
	cfg.nu_scalar = m_dNuScalar;
	cfg.nu_div = m_dNuDiv;
	cfg.nu_vort = m_dNuVort;
	cfg.fully_explicit = m_fFullyExplicit ? 1 : 0;
	cfg.off_centering = 0.0;

	int iResult = tb200_create(&cfg, &m_pCtx);
	Check(iResult);

	const int np = cfg.np;
	const int ne = pGrid->GetABaseResolution();
	const long nLattice = (long)(np - 1) * ne;

	// Patches (all on this rank: single-rank MPI build)
	for (int n = 0; n < pGrid->GetActivePatchCount(); n++) {
		GridPatchGLL * pPatch =
			dynamic_cast<GridPatchGLL *>(pGrid->GetActivePatch(n));
		const PatchBox & box = pPatch->GetPatchBox();
		Check(tb200_add_patch(
			m_pCtx, pPatch->GetPatchIndex(), box.GetPanel(),
			pPatch->GetElementCountA(), pPatch->GetElementCountB(),
			box.GetHaloElements(),
			pPatch->GetElementDeltaA(), pPatch->GetElementDeltaB(), 0));
	}
	Check(tb200_commit_layout(m_pCtx));

	// Tables and vertical column operators, from the host objects
	Check(tb200_set_tables(
		m_pCtx,
		&(pGrid->GetDxBasis1D()[0][0]),
		&(pGrid->GetStiffness1D()[0][0]),
		&(pGrid->GetGLLWeights1D()[0])));

	if (cfg.nlev > 1) {
		// (the operators the reference exposes; TB200_OP_DIFF_N2N_ZB follows below)
		const LinearColumnOperator * apOps[TB200_OP_DIFF_N2N_ZB] = {
			&(pGrid->GetOpInterpNodeToREdge()),
			&(pGrid->GetOpInterpREdgeToNode()),
			&(pGrid->GetOpDiffNodeToNode()),
			&(pGrid->GetOpDiffNodeToREdge()),
			&(pGrid->GetOpDiffREdgeToNode()),
			&(pGrid->GetOpDiffREdgeToREdge()),
			&(pGrid->GetOpDiffDiffNodeToNode()),
			&(pGrid->GetOpDiffDiffREdgeToREdge()),
			&(pGrid->GetOpPenaltyNodeToNode().GetLeftOp()),
			&(pGrid->GetOpPenaltyNodeToNode().GetRightOp())};
		const int nAccessibleOps = (int)(sizeof(apOps) / sizeof(apOps[0]));
		for (int o = 0; o < nAccessibleOps; o++) {
			const DataArray2D<double> & dCoeff = apOps[o]->GetCoeffs();
			if (dCoeff.GetRows() == 0) {
				continue;
			}
			Check(tb200_set_column_op(
				m_pCtx, o, dCoeff.GetRows(), dCoeff.GetColumns(),
				&(dCoeff[0][0]),
				&(apOps[o]->GetIxBegin()[0]),
				&(apOps[o]->GetIxEnd()[0])));
		}
	}

	// --vmassfluxlevels: BuildF differentiates level fluxes with
	// m_opDiffNodeToNodeZeroBoundaries (GridGLL.h:413), which has no accessor;
	// GridGLL::DifferentiateNodeToNode(., ., true) applied to the unit vectors
	// returns its coefficients exactly
	if (m_fMassFluxOnLevels && (cfg.eqn_type == TB200_EQN_PRIMITIVE_NONHYDRO) &&
	    (pGrid->GetRElements() > 1)
	) {
		const int nL = pGrid->GetRElements();
		std::vector<double> vecCoeff((size_t)nL * nL, 0.0);
		std::vector<int> vecBegin(nL, 0), vecEnd(nL, 0);
		DataArray1D<double> dIn(nL);
		DataArray1D<double> dOut(nL);
		for (int l = 0; l < nL; l++) {
			dIn.Zero();
			dOut.Zero();
			dIn[l] = 1.0;
			pGrid->DifferentiateNodeToNode(&(dIn[0]), &(dOut[0]), true);
			for (int k = 0; k < nL; k++) {
				vecCoeff[(size_t)k * nL + l] = dOut[k];
			}
		}
		for (int k = 0; k < nL; k++) {
			bool fAny = false;
			for (int l = 0; l < nL; l++) {
				if (vecCoeff[(size_t)k * nL + l] != 0.0) {
					if (!fAny) vecBegin[k] = l;
					vecEnd[k] = l + 1;
					fAny = true;
				}
			}
		}
		Check(tb200_set_column_op(
			m_pCtx, TB200_OP_DIFF_N2N_ZB, nL, nL,
			&(vecCoeff[0]), &(vecBegin[0]), &(vecEnd[0])));
		Check(tb200_set_mass_flux_on_levels(m_pCtx, 1));
	}

	// Geometry, node ids, seam transforms per patch
	for (int n = 0; n < pGrid->GetActivePatchCount(); n++) {
		GridPatchGLL * pPatch =
			dynamic_cast<GridPatchGLL *>(pGrid->GetActivePatch(n));
		const PatchBox & box = pPatch->GetPatchBox();
		const int ixPatch = pPatch->GetPatchIndex();
		const int nPanel = box.GetPanel();
		const int nHalo = box.GetHaloElements();

		tb200_geometry geo;
		memset(&geo, 0, sizeof(geo));
		geo.jacobian2d = &(pPatch->GetJacobian2D()[0][0]);
		geo.contrametric2da = &(pPatch->GetContraMetric2DA()[0][0][0]);
		geo.contrametric2db = &(pPatch->GetContraMetric2DB()[0][0][0]);
		geo.coriolis = &(pPatch->GetCoriolisF()[0][0]);
		geo.topography = &(pPatch->GetTopography()[0][0]);
		geo.jacobian = &(pPatch->GetJacobian()[0][0][0]);
		geo.jacobian_redge = &(pPatch->GetJacobianREdge()[0][0][0]);
		if (cfg.eqn_type == TB200_EQN_PRIMITIVE_NONHYDRO) {
			geo.contrametrica = &(pPatch->GetContraMetricA()[0][0][0][0]);
			geo.contrametricb = &(pPatch->GetContraMetricB()[0][0][0][0]);
			geo.contrametricxi = &(pPatch->GetContraMetricXi()[0][0][0][0]);
			geo.contrametrica_redge = &(pPatch->GetContraMetricAREdge()[0][0][0][0]);
			geo.contrametricb_redge = &(pPatch->GetContraMetricBREdge()[0][0][0][0]);
			geo.contrametricxi_redge = &(pPatch->GetContraMetricXiREdge()[0][0][0][0]);
			geo.derivr_node = &(pPatch->GetDerivRNode()[0][0][0][0]);
			geo.derivr_redge = &(pPatch->GetDerivRREdge()[0][0][0][0]);
		}
		Check(tb200_upload_geometry(m_pCtx, ixPatch, &geo));

		// Rayleigh friction (Grid::HasRayleighFriction, GridPatch.h:1053-1078)
		if (pGrid->HasRayleighFriction()) {
			Check(tb200_upload_rayleigh(
				m_pCtx, ixPatch,
				&(pPatch->GetRayleighStrength(DataLocation_Node)[0][0][0]),
				&(pPatch->GetRayleighStrength(DataLocation_REdge)[0][0][0]),
				&(pPatch->GetReferenceState(DataLocation_Node)[0][0][0][0]),
				&(pPatch->GetReferenceState(DataLocation_REdge)[0][0][0][0])));
		}

		// Uniform diffusion acts on the state minus the reference state
		// (Grid::HasUniformDiffusion, Grid.h:872-874)
		if (pGrid->HasUniformDiffusion()) {
			Check(tb200_upload_reference_state(
				m_pCtx, ixPatch,
				&(pPatch->GetReferenceState(DataLocation_Node)[0][0][0][0]),
				&(pPatch->GetReferenceState(DataLocation_REdge)[0][0][0][0])));
		}

		// On-the-fly terrain-following metric: m_dXNode / m_dYNode
		// (GridPatchCSGLL.cpp:205-213) and the topography derivatives
		if (cfg.eqn_type == TB200_EQN_PRIMITIVE_NONHYDRO) {
			std::vector<double> vecX(box.GetATotalWidth()), vecY(box.GetBTotalWidth());
			for (int i = 0; i < box.GetATotalWidth(); i++) {
				vecX[i] = tan(pPatch->GetANode(i));
			}
			for (int j = 0; j < box.GetBTotalWidth(); j++) {
				vecY[j] = tan(pPatch->GetBNode(j));
			}
			Check(tb200_set_terrain_metric(
				m_pCtx, ixPatch, &(vecX[0]), &(vecY[0]),
				&(pPatch->GetTopographyDeriv()[0][0][0])));
		}

		const int nWA = box.GetAInteriorWidth();
		const int nWB = box.GetBInteriorWidth();
		const long gA0 = box.GetAGlobalInteriorBegin();
		const long gB0 = box.GetBGlobalInteriorBegin();

		std::vector<int64_t> vecIds((size_t)nWA * nWB);
		std::vector<int> vecIa, vecIb, vecSrc;
		std::vector<double> vecM;
		const long m = 2 * nLattice + 1;
		if (pGridCart != NULL) {
			// periodic global node index: duplicates across element edges,
			// patch edges and the periodic boundaries share one id
			// (GridCartesianGLL.cpp:380-432)
			const long nUA = (long)(np - 1) * pGrid->GetABaseResolution();
			const long nUB = (long)(np - 1) * pGrid->GetBBaseResolution();
			for (int i = 0; i < nWA; i++) {
			for (int j = 0; j < nWB; j++) {
				const long uA = UniqueIndex(gA0 + i, np) % nUA;
				const long uB = UniqueIndex(gB0 + j, np) % nUB;
				vecIds[(size_t)i * nWB + j] = uA * nUB + uB;
			}
			}
		}
		for (int i = 0; (pGridCart == NULL) && (i < nWA); i++) {
		for (int j = 0; j < nWB; j++) {
			const long s = 2 * UniqueIndex(gA0 + i, np) - nLattice;
			const long t = 2 * UniqueIndex(gB0 + j, np) - nLattice;
			long x, y, z;
			LatticePoint(nPanel, s, t, nLattice, x, y, z);
			vecIds[(size_t)i * nWB + j] =
				((x + nLattice) * m + (y + nLattice)) * m + (z + nLattice);

			// Nodes on a panel edge: covector re-basing from every other
			// panel containing the point, with the reference's own formulas
			// (CubedSphereTrans::CoVecPanelTrans, CubedSphereTrans.h:1751)
			if ((labs(s) != nLattice) && (labs(t) != nLattice)) {
				continue;
			}
			const bool fOn[6] = {
				x == nLattice, y == nLattice, x == -nLattice,
				y == -nLattice, z == nLattice, z == -nLattice};
			const double dX = tan(pPatch->GetANode(i + nHalo));
			const double dY = tan(pPatch->GetBNode(j + nHalo));
			for (int q = 0; q < 6; q++) {
				if ((!fOn[q]) || (q == nPanel)) {
					continue;
				}
				double dM[4];
				for (int u = 0; u < 2; u++) {
					double dA = (u == 0) ? 1.0 : 0.0;
					double dB = (u == 0) ? 0.0 : 1.0;
					CubedSphereTrans::CoVecPanelTrans(q, nPanel, dA, dB, dX, dY);
					dM[0 + u] = dA;
					dM[2 + u] = dB;
				}
				vecIa.push_back(i);
				vecIb.push_back(j);
				vecSrc.push_back(q);
				for (int u = 0; u < 4; u++) {
					vecM.push_back(dM[u]);
				}
			}
		}
		}
		Check(tb200_set_node_ids(m_pCtx, ixPatch, &(vecIds[0])));
		Check(tb200_set_seam_transforms(
			m_pCtx, ixPatch, (int)vecIa.size(),
			vecIa.empty() ? NULL : &(vecIa[0]),
			vecIb.empty() ? NULL : &(vecIb[0]),
			vecSrc.empty() ? NULL : &(vecSrc[0]),
			vecM.empty() ? NULL : &(vecM[0])));
	}
	if (cfg.eqn_type == TB200_EQN_PRIMITIVE_NONHYDRO) {
		Check(tb200_set_vertical_coordinate(
			m_pCtx, &(pGrid->GetREtaLevels()[0]), &(pGrid->GetREtaInterfaces()[0])));
	}
	Check(tb200_build_connectivity(m_pCtx));
	if (pGrid->GetVerticalDiscretization() ==
	    Grid::VerticalDiscretization_FiniteVolume
	) {
		Check(tb200_set_vertical_discretization(m_pCtx, 1));
	}
	if (pGrid->HasUniformDiffusion()) {
		// (--norefstate leaves the column's reference arrays zero while the
		// horizontal part still removes the reference state,
		// VerticalDynamicsFEM.cpp:1756: not restated)
		if (!m_fUseReferenceState) {
			_EXCEPTIONT("B200 plugins: uniform diffusion with --norefstate "
				"is not supported");
		}
		Check(tb200_set_uniform_diffusion(
			m_pCtx,
			pGrid->GetScalarUniformDiffusionCoeff(),
			pGrid->GetVectorUniformDiffusionCoeff()));
	}

	// Pin the state containers (one contiguous block each, DataContainer.cpp:77-147):
	// bus copies then run asynchronously at full rate
	for (int n = 0; n < pGrid->GetActivePatchCount(); n++) {
		GridPatch * pPatch = pGrid->GetActivePatch(n);
		DataContainer * apDC[2] = {
			&(pPatch->GetDataContainerActiveState()),
			&(pPatch->GetDataContainerBufferState())};
		for (int q = 0; q < 2; q++) {
			if (apDC[q]->GetTotalByteSize() == 0) {
				continue;
			}
			if (tb200_host_register(
					m_pCtx, apDC[q]->GetPointer(), apDC[q]->GetTotalByteSize()) == 0
			) {
				m_vecPinned.push_back(apDC[q]->GetPointer());
			}
		}
	}

	// FunctionTimer groups of the reference around the device calls
	const char * szTiming = getenv("TB200_TIMING");
	if ((szTiming != NULL) && (szTiming[0] == '1')) {
		Check(tb200_set_timing_hooks(
			m_pCtx, &B200Bridge::TimerBegin, &B200Bridge::TimerEnd, this));
	}
	m_fInitialized = true;
}

///////////////////////////////////////////////////////////////////////////////

void B200Bridge::TimerBegin(void * pUser, const char * szGroup) {
	B200Bridge * pBridge = static_cast<B200Bridge *>(pUser);
	pBridge->m_vecTimers.push_back(new FunctionTimer(szGroup));
}

void B200Bridge::TimerEnd(void * pUser, const char * szGroup) {
	B200Bridge * pBridge = static_cast<B200Bridge *>(pUser);
	if (!pBridge->m_vecTimers.empty()) {
		delete pBridge->m_vecTimers.back();
		pBridge->m_vecTimers.pop_back();
	}
}

///////////////////////////////////////////////////////////////////////////////

void B200Bridge::Upload(int iInstance) {
	Grid * pGrid = m_model.GetGrid();
	const bool fTracers = (m_model.GetEquationSet().GetTracers() != 0);
	for (int n = 0; n < pGrid->GetActivePatchCount(); n++) {
		GridPatch * pPatch = pGrid->GetActivePatch(n);
		Check(tb200_upload_state_async(
			m_pCtx, pPatch->GetPatchIndex(), iInstance,
			&(pPatch->GetDataState(iInstance, DataLocation_Node)[0][0][0][0]),
			&(pPatch->GetDataState(iInstance, DataLocation_REdge)[0][0][0][0]),
			fTracers ? &(pPatch->GetDataTracers(iInstance)[0][0][0][0]) : NULL));
	}
	Check(tb200_transfer_sync(m_pCtx));
}

void B200Bridge::Download(int iInstance) {
	Grid * pGrid = m_model.GetGrid();
	const bool fTracers = (m_model.GetEquationSet().GetTracers() != 0);
	for (int n = 0; n < pGrid->GetActivePatchCount(); n++) {
		GridPatch * pPatch = pGrid->GetActivePatch(n);
		Check(tb200_download_state_async(
			m_pCtx, pPatch->GetPatchIndex(), iInstance,
			&(pPatch->GetDataState(iInstance, DataLocation_Node)[0][0][0][0]),
			&(pPatch->GetDataState(iInstance, DataLocation_REdge)[0][0][0][0]),
			fTracers ? &(pPatch->GetDataTracers(iInstance)[0][0][0][0]) : NULL,
			1));
	}
	Check(tb200_transfer_sync(m_pCtx));
}

///////////////////////////////////////////////////////////////////////////////
// HorizontalDynamicsB200

HorizontalDynamicsB200::HorizontalDynamicsB200(
	Model & model,
	int nHorizontalOrder,
	int nHyperviscosityOrder,
	double dNuScalar,
	double dNuDiv,
	double dNuVort,
	double dInstepNuDiv
) :
	HorizontalDynamics(model)
{
	B200Bridge::Get(model).SetHyperviscosity(
		nHyperviscosityOrder, dNuScalar, dNuDiv, dNuVort);
}

void HorizontalDynamicsB200::Initialize() {
	B200Bridge::Get(m_model).Initialize();
}

void HorizontalDynamicsB200::StepExplicit(
	int iDataInitial, int iDataUpdate, const Time & time, double dDeltaT
) {
	// Used under one of the reference's own TimestepSchemes the state lives
	// on the host between plugin calls: move both instances across.
	B200Bridge & b = B200Bridge::Get(m_model);
	b.Upload(iDataInitial);
	b.Upload(iDataUpdate);
	b.Check(tb200_h_step_explicit(b.Ctx(), iDataInitial, iDataUpdate, dDeltaT));
	// the reference leaves W on levels and U,V on interfaces in the input
	// instance (HorizontalDynamicsFEM.cpp:817-831); download refreshes them
	b.Download(iDataInitial);
	b.Download(iDataUpdate);
}

void HorizontalDynamicsB200::StepAfterSubCycle(
	int iDataInitial, int iDataUpdate, int iDataWorking,
	const Time & time, double dDeltaT
) {
	B200Bridge & b = B200Bridge::Get(m_model);
	b.Upload(iDataInitial);
	b.Check(tb200_h_step_after_subcycle(
		b.Ctx(), iDataInitial, iDataUpdate, iDataWorking, dDeltaT));
	b.Download(iDataUpdate);
	// the working instance is scratch: the reference writes it only in the
	// order-4 branch (HorizontalDynamicsFEM.cpp:2687-2713) and nothing reads it
	// afterwards; leave the host copy alone
}

///////////////////////////////////////////////////////////////////////////////
// VerticalDynamicsB200

VerticalDynamicsB200::VerticalDynamicsB200(
	Model & model,
	int nHorizontalOrder,
	int nVerticalOrder,
	int nHypervisOrder,
	bool fFullyExplicit,
	bool fUseReferenceState,
	bool fForceMassFluxOnLevels
) :
	VerticalDynamics(model)
{
	// --explicitvertical (VerticalDynamicsFEM.cpp:748-793, 1240-1242)
	B200Bridge::Get(model).SetFullyExplicit(fFullyExplicit);
	B200Bridge::Get(model).SetUseReferenceState(fUseReferenceState);
	B200Bridge::Get(model).SetMassFluxOnLevels(fForceMassFluxOnLevels);
}

void VerticalDynamicsB200::Initialize() {
	B200Bridge::Get(m_model).Initialize();
}

void VerticalDynamicsB200::StepExplicit(
	int iDataInitial, int iDataUpdate, const Time & time, double dDeltaT
) {
	B200Bridge & b = B200Bridge::Get(m_model);
	b.Upload(iDataInitial);
	b.Upload(iDataUpdate);
	b.Check(tb200_v_step_explicit(b.Ctx(), iDataInitial, iDataUpdate, dDeltaT));
	b.Download(iDataUpdate);
}

void VerticalDynamicsB200::StepImplicitTermsExplicitly(
	int iDataInitial, int iDataUpdate, const Time & time, double dDeltaT
) {
	B200Bridge & b = B200Bridge::Get(m_model);
	b.Upload(iDataInitial);
	b.Upload(iDataUpdate);
	b.Check(tb200_v_step_implicit_terms_explicitly(
		b.Ctx(), iDataInitial, iDataUpdate, dDeltaT));
	b.Download(iDataUpdate);
}

void VerticalDynamicsB200::StepImplicit(
	int iDataInitial, int iDataUpdate, const Time & time, double dDeltaT
) {
	B200Bridge & b = B200Bridge::Get(m_model);
	b.Upload(iDataInitial);
	if (iDataUpdate != iDataInitial) {
		b.Upload(iDataUpdate);
	}
	b.Check(tb200_v_step_implicit(b.Ctx(), iDataInitial, iDataUpdate, dDeltaT));
	b.Check(tb200_check_errors(b.Ctx()));
	b.Download(iDataUpdate);
}

void VerticalDynamicsB200::FilterNegativeTracers(int iDataUpdate) {
	if (m_model.GetEquationSet().GetTracers() == 0) {
		return;
	}
	B200Bridge & b = B200Bridge::Get(m_model);
	b.Upload(iDataUpdate);
	b.Check(tb200_v_filter_negative_tracers(b.Ctx(), iDataUpdate));
	b.Download(iDataUpdate);
}

///////////////////////////////////////////////////////////////////////////////
// TimestepSchemeB200

TimestepSchemeB200::TimestepSchemeB200(Model & model, int iScheme) :
	TimestepScheme(model),
	m_iScheme(iScheme),
	m_fLazy(false),
	m_fDeviceCurrent(false),
	m_fHostCopy(false)
{
	if (tb200_scheme_instances(iScheme) < 0) {
		_EXCEPTIONT("tempest_b200: time scheme not implemented");
	}
}

int TimestepSchemeB200::GetComponentDataInstances() const {
	return tb200_scheme_instances(m_iScheme);
}

int TimestepSchemeB200::GetTracerDataInstances() const {
	return tb200_scheme_instances(m_iScheme);
}

void TimestepSchemeB200::Initialize() {
}

void TimestepSchemeB200::Step(
	bool fFirstStep, bool fLastStep, const Time & time, double dDeltaT
) {
	B200Bridge & b = B200Bridge::Get(m_model);
	b.Initialize();
	// Between steps the host may read and change instance 0 (workflow
	// processes, Model.cpp:477-481; outputs, :484-509): it is the only
	// instance that crosses the bus.  The carry-over instance of the Strang
	// scheme stays on the device.
	if (!m_fLazy || !m_fDeviceCurrent) {
		b.Upload(0);
	}
	b.Check(tb200_step(b.Ctx(), m_iScheme, fFirstStep ? 1 : 0, fLastStep ? 1 : 0, dDeltaT));
	m_fDeviceCurrent = true;

	// Who touches instance 0 on the host before the next step?  Model::Go asks
	// its workflow processes and output managers with the time at the end of
	// this step (Model.cpp:470-509).
	Time timeNext = time;
	timeNext += m_model.GetDeltaT();
	if (timeNext >= m_model.GetEndTime()) {
		timeNext = m_model.GetEndTime();
	}
	bool fHostReads = (!m_fLazy) || fLastStep;
	for (size_t q = 0; q < m_vecNextRead.size(); q++) {
		if ((!m_vecReadFrequency[q].IsZero()) && (!(timeNext < m_vecNextRead[q]))) {
			fHostReads = true;
			m_vecNextRead[q] += m_vecReadFrequency[q];
		}
	}
	for (size_t q = 0; q < m_vecProcesses.size(); q++) {
		if (m_vecProcesses[q]->IsReady(timeNext)) {
			// its Perform() changes the host copy: upload before the next step
			fHostReads = true;
			m_fDeviceCurrent = false;
		}
	}
	if (fHostReads) {
		b.Check(tb200_check_errors(b.Ctx()));
		b.Download(0);
	}
	m_fHostCopy = fHostReads;
}

void TimestepSchemeB200::HostReadsEvery(
	const Time & timeStart, const Time & timeFrequency
) {
	m_fLazy = true;
	Time timeNext = timeStart;
	timeNext += timeFrequency;
	m_vecNextRead.push_back(timeNext);
	m_vecReadFrequency.push_back(timeFrequency);
}

void TimestepSchemeB200::HostProcess(WorkflowProcess * pProcess) {
	m_fLazy = true;
	m_vecProcesses.push_back(pProcess);
}

///////////////////////////////////////////////////////////////////////////////

///////////////////////////////////////////////////////////////////////////////
// HeldSuarezPhysicsB200

HeldSuarezPhysicsB200::HeldSuarezPhysicsB200(
	Model & model, const Time & timeFrequency, TimestepSchemeB200 * pScheme
) :
	WorkflowProcess(model, timeFrequency),
	m_pScheme(pScheme),
	m_fUploaded(false)
{ }

void HeldSuarezPhysicsB200::Perform(const Time & time) {
	B200Bridge & b = B200Bridge::Get(m_model);
	b.Initialize();
	Grid * pGrid = m_model.GetGrid();
	if (!m_fUploaded) {
		// latitude and the slots the reference takes its surface pressure from
		// (HeldSuarezPhysics.cpp:95-115): rho and the theta slot on the lowest
		// interface of instance 0, which the dynamics never writes
		for (int n = 0; n < pGrid->GetActivePatchCount(); n++) {
			GridPatch * pPatch = pGrid->GetActivePatch(n);
			const DataArray2D<double> & dataLatitude = pPatch->GetLatitude();
			const DataArray4D<double> & dataREdge =
				pPatch->GetDataState(0, DataLocation_REdge);
			DataArray2D<double> dProduct(dataLatitude.GetRows(), dataLatitude.GetColumns());
			for (int i = 0; i < dProduct.GetRows(); i++) {
			for (int j = 0; j < dProduct.GetColumns(); j++) {
				dProduct[i][j] = dataREdge[4][i][j][0] * dataREdge[2][i][j][0];
			}
			}
			b.Check(tb200_upload_held_suarez(
				b.Ctx(), pPatch->GetPatchIndex(), &(dataLatitude[0][0]), &(dProduct[0][0])));
		}
		m_fUploaded = true;
	}
	const double dDeltaT = m_timeFrequency.GetSeconds();
	if (m_pScheme == NULL) {
		// host time scheme: instance 0 lives on the host between plugin calls
		b.Upload(0);
		b.Check(tb200_held_suarez(b.Ctx(), dDeltaT));
		b.Download(0);
	} else {
		b.Check(tb200_held_suarez(b.Ctx(), dDeltaT));
		if (m_pScheme->HostCopyIsCurrent()) {
			b.Download(0);
		}
	}
	WorkflowProcess::Perform(time);
}
