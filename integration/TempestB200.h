///////////////////////////////////////////////////////////////////////////////
///
///	\file    TempestB200.h
///
///	Reference-side binding of libtempest_b200: plugin classes a maintainer
///	adds to the reference tree.  They derive from the reference's own plugin
///	interfaces (HorizontalDynamics.h:34-174, VerticalDynamics.h:30-134,
///	TimestepScheme.h:32-124) and are the only translation unit that sees both
///	the reference headers and the C ABI (include/tempest_b200.h).
///
///	  HorizontalDynamicsB200  replaces HorizontalDynamicsFEM   (--hmethod B200)
///	  VerticalDynamicsB200    replaces VerticalDynamicsFEM     (--vmethod B200)
///	  TimestepSchemeB200      replaces TimestepSchemeStrang / ARS343 and keeps
///	                          every state instance on the device for the whole
///	                          step (--timescheme b200/strang, b200/ars343)
///
///	Errors of the library surface as the reference's Exception
///	(src/base/Exception.h:25-49) carrying tb200_last_error().
///
///////////////////////////////////////////////////////////////////////////////

#ifndef _TEMPESTB200_H_
#define _TEMPESTB200_H_

#include "Model.h"
#include "GridGLL.h"
#include "GridPatchGLL.h"
#include "HorizontalDynamics.h"
#include "VerticalDynamics.h"
#include "TimestepScheme.h"

#include "FunctionTimer.h"
#include "WorkflowProcess.h"

#include "tempest_b200.h"

#include <vector>

///////////////////////////////////////////////////////////////////////////////

///	<summary>
///		The device context shared by the three plugins of one Model.
///	</summary>
class B200Bridge {

public:
	///	<summary>
	///		Bridge of a model (created on first use).
	///	</summary>
	static B200Bridge & Get(Model & model);

	///	<summary>
	///		Set the parameters the HorizontalDynamicsFEM constructor takes.
	///	</summary>
	void SetHyperviscosity(int nOrder, double dNuScalar, double dNuDiv, double dNuVort);

	///	<summary>
	///		The fFullyExplicit argument of the VerticalDynamicsFEM constructor
	///		(--explicitvertical).
	///	</summary>
	void SetFullyExplicit(bool fFullyExplicit);

	///	<summary>
	///		The fUseReferenceState argument of the VerticalDynamicsFEM
	///		constructor (false under --norefstate).
	///	</summary>
	void SetUseReferenceState(bool fUseReferenceState);

	///	<summary>
	///		The fForceMassFluxOnLevels argument of the VerticalDynamicsFEM
	///		constructor (--vmassfluxlevels).
	///	</summary>
	void SetMassFluxOnLevels(bool fMassFluxOnLevels);

	///	<summary>
	///		Create the context, describe the grid, upload geometry and tables.
	///		Called from the plugins' Initialize(), i.e. after
	///		Grid::EvaluateGeometricTerms (Model.cpp:347-355).
	///	</summary>
	void Initialize();

	///	<summary>
	///		Host instance -> device, device -> host.  All patches are enqueued
	///		(bus copies and layout conversions overlap), then the host waits once.
	///	</summary>
	void Upload(int iInstance);
	void Download(int iInstance);

	///	<summary>
	///		Open / close one of the reference's FunctionTimer groups
	///		(tb200_set_timing_hooks); installed when TB200_TIMING=1.
	///	</summary>
	static void TimerBegin(void * pUser, const char * szGroup);
	static void TimerEnd(void * pUser, const char * szGroup);

	///	<summary>
	///		Throw the reference's Exception if a C-ABI call failed.
	///	</summary>
	void Check(int iResult);

	tb200_ctx * Ctx() { return m_pCtx; }

private:
	B200Bridge(Model & model);

	Model & m_model;
	tb200_ctx * m_pCtx;
	bool m_fInitialized;
	std::vector<void *> m_vecPinned;
	std::vector<FunctionTimer *> m_vecTimers;
	int m_nHypervisOrder;
	double m_dNuScalar, m_dNuDiv, m_dNuVort;
	bool m_fFullyExplicit;
	bool m_fUseReferenceState;
	bool m_fMassFluxOnLevels;
};

///////////////////////////////////////////////////////////////////////////////

class HorizontalDynamicsB200 : public HorizontalDynamics {
public:
	///	<summary>
	///		Same arguments as HorizontalDynamicsFEM (HorizontalDynamicsFEM.h).
	///	</summary>
	HorizontalDynamicsB200(
		Model & model,
		int nHorizontalOrder,
		int nHyperviscosityOrder,
		double dNuScalar,
		double dNuDiv,
		double dNuVort,
		double dInstepNuDiv = 0.0);

	virtual void Initialize();
	virtual void StepExplicit(int iDataInitial, int iDataUpdate, const Time & time, double dDeltaT);
	virtual void StepAfterSubCycle(int iDataInitial, int iDataUpdate, int iDataWorking, const Time & time, double dDeltaT);
};

class VerticalDynamicsB200 : public VerticalDynamics {
public:
	///	<summary>
	///		Same arguments as VerticalDynamicsFEM (VerticalDynamicsFEM.h).
	///	</summary>
	VerticalDynamicsB200(
		Model & model,
		int nHorizontalOrder,
		int nVerticalOrder,
		int nHypervisOrder = 0,
		bool fFullyExplicit = false,
		bool fUseReferenceState = true,
		bool fForceMassFluxOnLevels = false);

	virtual void Initialize();
	virtual void StepExplicit(int iDataInitial, int iDataUpdate, const Time & time, double dDeltaT);
	virtual void StepImplicit(int iDataInitial, int iDataUpdate, const Time & time, double dDeltaT);
	virtual void StepImplicitTermsExplicitly(
		int iDataInitial, int iDataUpdate, const Time & time, double dDeltaT);
	///	<summary>
	///		Column-wise positive-definite filter (VerticalDynamics.h:123).
	///	</summary>
	virtual void FilterNegativeTracers(int iDataUpdate);
};

class TimestepSchemeB200 : public TimestepScheme {
public:
	TimestepSchemeB200(Model & model, int iScheme);

	virtual int GetComponentDataInstances() const;
	virtual int GetTracerDataInstances() const;
	virtual void Initialize();
	///	<summary>
	///		The whole step runs on the device.  Between steps the host only
	///		touches state instance 0, and only when an output manager fires
	///		(Model.cpp:484-509) or a workflow process is ready (:477-481): the
	///		instance is downloaded at the end of a step only then (and on the
	///		last step), and uploaded at the start of a step only when the host
	///		may have changed it (first step, after a workflow process).
	///	</summary>
	virtual void Step(bool fFirstStep, bool fLastStep, const Time & time, double dDeltaT);

	///	<summary>
	///		Tell the scheme who reads instance 0 on the host between steps.
	///		An output manager created with output frequency `timeFrequency`
	///		fires at start + k * frequency (OutputManager::IsOutputNeeded,
	///		OutputManager.cpp:83-98; a zero frequency only writes the initial
	///		and final state).  Without any registration the scheme is
	///		conservative: instance 0 crosses the bus both ways every step.
	///	</summary>
	void HostReadsEvery(const Time & timeStart, const Time & timeFrequency);

	///	<summary>
	///		A workflow process that reads and changes instance 0 on the host
	///		(e.g. column physics): asked through its own IsReady(); instance 0
	///		is uploaded again before the step that follows its Perform().
	///	</summary>
	void HostProcess(WorkflowProcess * pProcess);

	///	<summary>
	///		Did the last Step() leave a copy of instance 0 on the host (an
	///		output manager is about to read it)?  A workflow process that
	///		changes instance 0 on the device refreshes that copy.
	///	</summary>
	bool HostCopyIsCurrent() const { return m_fHostCopy; }

private:
	int m_iScheme;
	bool m_fLazy;
	bool m_fDeviceCurrent;
	bool m_fHostCopy;
	std::vector<Time> m_vecNextRead;
	std::vector<Time> m_vecReadFrequency;
	std::vector<WorkflowProcess *> m_vecProcesses;
};

///////////////////////////////////////////////////////////////////////////////

///	<summary>
///		HeldSuarezPhysics (src/atm/HeldSuarezPhysics.h:27-45) as a device
///		workflow step: same constructor, same Perform(), the forcing is applied
///		to instance 0 where it lives.  With TimestepSchemeB200 (pScheme given)
///		that is the device between steps - no bus traffic unless an output
///		manager is about to read the host copy; under a host time scheme the
///		instance is uploaded before and downloaded after.
///	</summary>
class HeldSuarezPhysicsB200 : public WorkflowProcess {
public:
	HeldSuarezPhysicsB200(
		Model & model, const Time & timeFrequency, TimestepSchemeB200 * pScheme = NULL);

	virtual void Perform(const Time & time);

private:
	TimestepSchemeB200 * m_pScheme;
	bool m_fUploaded;
};

///////////////////////////////////////////////////////////////////////////////

#endif
