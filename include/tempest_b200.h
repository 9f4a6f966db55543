/*
 * tempest_b200.h - C ABI of the B200-native Tempest dynamical-core hot path.
 *
 * The library (libtempest_b200.so, hand-written FP64 CUDA for sm_100a) is the
 * drop-in for the reference's per-timestep path between Model.cpp:420
 * (m_pTimestepScheme->Step) and its return.  Every entry point below cites the
 * reference interface it replaces (paths relative to the reference root).
 *
 * Conventions
 *  - plain pointers and sizes only; no C++ or torch types cross this boundary;
 *  - every function returns 0 on success, non-zero on failure;
 *    tb200_last_error() returns the message (the C++ shells turn it into the
 *    reference's Exception, src/base/Exception.h:25-49);
 *  - one caller thread per context (the reference is single threaded per rank);
 *  - HOST arrays use the reference layout: state [c][iA][iB][k], k fastest,
 *    with the one-node halo (src/base/DataArray4D.h:507-530,
 *    src/atm/GridPatch.cpp:341-357); the device layout is private.
 *  - "instance" is the reference's state-instance index
 *    (TimestepScheme::GetComponentDataInstances, TimestepScheme.h:55-63).
 */
#ifndef TEMPEST_B200_H
#define TEMPEST_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct tb200_ctx tb200_ctx;

#define TB200_MAX_COMPONENTS 8

/* EquationSet::Type (src/atm/EquationSet.h) */
#define TB200_EQN_SHALLOW_WATER 1
#define TB200_EQN_PRIMITIVE_NONHYDRO 2

/* DataType selector (src/atm/DataType.h): bit mask */
#define TB200_DATA_STATE 1
#define TB200_DATA_TRACERS 2
/* derived output fields (tb200_compute_output_fields, tb200_interpolate) */
#define TB200_DATA_TEMPERATURE 4
#define TB200_DATA_VORTICITY 8
#define TB200_DATA_DIVERGENCE 16

/* Vertical column operators of GridGLL (src/atm/GridGLL.h:357-448) */
enum tb200_column_op {
	TB200_OP_INTERP_N2E = 0,   /* m_opInterpNodeToREdge  */
	TB200_OP_INTERP_E2N = 1,   /* m_opInterpREdgeToNode  */
	TB200_OP_DIFF_N2N = 2,     /* m_opDiffNodeToNode     */
	TB200_OP_DIFF_N2E = 3,     /* m_opDiffNodeToREdge    */
	TB200_OP_DIFF_E2N = 4,     /* m_opDiffREdgeToNode    */
	TB200_OP_DIFF_E2E = 5,     /* m_opDiffREdgeToREdge   */
	TB200_OP_DIFFDIFF_N2N = 6, /* m_opDiffDiffNodeToNode */
	TB200_OP_DIFFDIFF_E2E = 7, /* m_opDiffDiffREdgeToREdge */
	TB200_OP_PENALTY_LEFT = 8, /* m_opPenaltyNodeToNode.GetLeftOp()  */
	TB200_OP_PENALTY_RIGHT = 9,/* m_opPenaltyNodeToNode.GetRightOp() */
	TB200_OP_DIFF_N2N_ZB = 10, /* m_opDiffNodeToNodeZeroBoundaries (no accessor in the
	                            * reference: GridGLL::DifferentiateNodeToNode(in, out,
	                            * true) applied to unit vectors); --vmassfluxlevels only */
	TB200_OP_COUNT = 11
};

/* Time schemes (src/atm/TimestepScheme*.cpp) */
enum tb200_scheme {
	TB200_SCHEME_STRANG_KGU35 = 0, /* TimestepSchemeStrang (default)   */
	TB200_SCHEME_ARS343 = 1,       /* TimestepSchemeARS343             */
	TB200_SCHEME_ARS232 = 2,       /* TimestepSchemeARS232             */
	TB200_SCHEME_ARS222 = 3,       /* TimestepSchemeARS222             */
	TB200_SCHEME_ARS443 = 4,       /* TimestepSchemeARS443             */
	TB200_SCHEME_STRANG_RK4 = 5,
	TB200_SCHEME_STRANG_SSP3 = 6,
	TB200_SCHEME_STRANG_FE = 7,
	TB200_SCHEME_STRANG_SSPRK53 = 8,
	TB200_SCHEME_ERK_KGU35 = 9,    /* TimestepSchemeERK (default "erk") */
	TB200_SCHEME_ERK_FE = 10,
	TB200_SCHEME_ERK_RK4 = 11,
	TB200_SCHEME_ERK_SSP3 = 12,
	TB200_SCHEME_ERK_SSPRK53 = 13,
	TB200_SCHEME_GARK2 = 14,       /* TimestepSchemeGARK2   ("gark2")    */
	TB200_SCHEME_SSP3332 = 15,     /* TimestepSchemeSSP3332 ("ssp3_332") */
	TB200_SCHEME_ARK232 = 16       /* TimestepSchemeARK232  ("ark232")   */
};

/*
 * Static configuration: what Model, EquationSet, GridGLL, PhysicalConstants
 * and the HorizontalDynamicsFEM / VerticalDynamicsFEM constructors hold
 * (src/atm/TempestInitialize.h:112-144,185-409; PhysicalConstants.h).
 */
typedef struct {
	int np;              /* GridGLL::GetHorizontalOrder (nodes per element edge):
	                      * 3, 4, 5 or 6; the column-constant kernels are np = 4 */
	int nlev;            /* Grid::GetRElements                                   */
	int vertical_order;  /* GridGLL::GetVerticalOrder                            */
	int ncomp;           /* EquationSet::GetComponents (3 SW, 5 nonhydro)        */
	int ntracers;        /* EquationSet::GetTracers                              */
	int ninstances;      /* Model::GetComponentDataInstances                     */
	int eqn_type;        /* TB200_EQN_*                                          */
	int cartesian_xz;    /* GridGLL::GetIsCartesianXZ                            */
	int comp_on_redge[TB200_MAX_COMPONENTS]; /* Grid::GetVarLocation == REdge   */
	int device;          /* CUDA device ordinal (-1: current device)             */
	/* PhysicalConstants */
	double g, R, cp, cv, p0, omega, earth_radius;
	double ztop;             /* Grid::GetZtop             */
	double ref_length;       /* Grid::GetReferenceLength  */
	/* HorizontalDynamicsFEM ctor (HorizontalDynamicsFEM.cpp:44-73) */
	int hypervis_order;      /* 0 (--nohypervis), 2 or 4 */
	double nu_scalar, nu_div, nu_vort;
	/* VerticalDynamicsFEM ctor (VerticalDynamicsFEM.cpp:53-78) */
	int fully_explicit;      /* --explicitvertical */
	/* TimestepSchemeStrang off-centering (TimestepSchemeStrang.h:53) */
	double off_centering;
} tb200_config;

/* Host pointers to one patch's geometric arrays in the reference layout
 * (allocated in GridPatch::InitializeDataLocal, GridPatch.cpp:102-285).
 * W_A x W_B include the halo.  Pointers that a configuration does not need
 * (all 3-D arrays for shallow water) may be NULL. */
typedef struct {
	const double * jacobian2d;         /* [W_A][W_B]         GetJacobian2D         */
	const double * contrametric2da;    /* [W_A][W_B][2]      GetContraMetric2DA    */
	const double * contrametric2db;    /* [W_A][W_B][2]      GetContraMetric2DB    */
	const double * coriolis;           /* [W_A][W_B]         GetCoriolisF          */
	const double * topography;         /* [W_A][W_B]         GetTopography         */
	const double * jacobian;           /* [W_A][W_B][L]      GetJacobian           */
	const double * jacobian_redge;     /* [W_A][W_B][L+1]    GetJacobianREdge      */
	const double * contrametrica;      /* [W_A][W_B][L][3]   GetContraMetricA      */
	const double * contrametricb;      /* [W_A][W_B][L][3]   GetContraMetricB      */
	const double * contrametricxi;     /* [W_A][W_B][L][3]   GetContraMetricXi     */
	const double * contrametrica_redge;/* [W_A][W_B][L+1][3] GetContraMetricAREdge */
	const double * contrametricb_redge;/* [W_A][W_B][L+1][3] GetContraMetricBREdge */
	const double * contrametricxi_redge;/*[W_A][W_B][L+1][3] GetContraMetricXiREdge*/
	const double * derivr_node;        /* [W_A][W_B][L][3]   GetDerivRNode         */
	const double * derivr_redge;       /* [W_A][W_B][L+1][3] GetDerivRREdge        */
} tb200_geometry;

/* ---- lifetime ----------------------------------------------------------- */

/* Replaces the constructors of HorizontalDynamicsFEM / VerticalDynamicsFEM /
 * TimestepScheme* (TempestInitialize.h:185-409). */
int tb200_create(const tb200_config * cfg, tb200_ctx ** out);
int tb200_destroy(tb200_ctx * ctx);
const char * tb200_last_error(const tb200_ctx * ctx);
const char * tb200_version(void);

/* Launch all work on this CUDA stream (a cudaStream_t passed as void*). */
int tb200_set_stream(tb200_ctx * ctx, void * cuda_stream);
/* Block until all queued work is complete (FunctionTimer groups). */
int tb200_sync(tb200_ctx * ctx);
/* tb200_sync + report deferred device-side failures of the column solve
 * ("Solution failed" / "Inversion failure", VerticalDynamicsFEM.cpp:1461-1481). */
int tb200_check_errors(tb200_ctx * ctx);

/* ---- grid description (Grid::NewPatch / GridPatchGLL ctor) --------------- */

/* Register one patch of the global grid (GridPatch.h:331 index, PatchBox.h:74
 * panel, GridPatchGLL.h:94 element spacing) and the rank that owns it
 * (Grid::DistributePatches, Grid.cpp:1038-1062).  EVERY patch of the grid is
 * registered on every rank; only patches with owner_rank == this rank hold
 * data.  Patches must be added before tb200_commit_layout.  halo is
 * PatchBox::GetHaloElements (1) and only describes the HOST arrays. */
int tb200_add_patch(tb200_ctx * ctx, int patch_index, int panel,
                    int nelem_a, int nelem_b, int halo,
                    double delta_a, double delta_b, int owner_rank);
/* Allocate device storage for every state instance and the geometry. */
int tb200_commit_layout(tb200_ctx * ctx);

/* GridGLL::GetDxBasis1D / GetStiffness1D / GetGLLWeights1D
 * (GridGLL.cpp:101-180); [np][np], [np][np], [np]. */
int tb200_set_tables(tb200_ctx * ctx, const double * dx_basis,
                     const double * stiffness, const double * gll_weights);
/* LinearColumnOperator coefficient table + per-row band
 * (LinearColumnOperator.h:62-236): coeff[nout][nin], begin[nout], end[nout]. */
int tb200_set_column_op(tb200_ctx * ctx, int op, int nout, int nin,
                        const double * coeff, const int * begin, const int * end);
/* Upload one patch's metric terms (GridPatchCSGLL::EvaluateGeometricTerms,
 * GridPatchCSGLL.cpp:295-574 / GridPatchCartesianGLL.cpp:197-460 outputs). */
int tb200_upload_geometry(tb200_ctx * ctx, int patch_index, const tb200_geometry * g);

/* Optional: let the kernels evaluate the terrain-following cubed-sphere metric
 * on the fly instead of reading the stored 3-D arrays
 * (GridPatchCSGLL::EvaluateGeometricTerms, GridPatchCSGLL.cpp:344-553; same
 * expressions, same bits).  xnode[W_A], ynode[W_B] = tan of GetANode/GetBNode
 * (GridPatchCSGLL.cpp:205-213), topography_deriv = GetTopographyDeriv()
 * [W_A][W_B][2]; reta_* = Grid::GetREtaLevels / GetREtaInterfaces. */
int tb200_set_terrain_metric(tb200_ctx * ctx, int patch_index, const double * xnode,
                             const double * ynode, const double * topography_deriv);
int tb200_set_vertical_coordinate(tb200_ctx * ctx, const double * reta_levels,
                                  const double * reta_interfaces);

/* Fast path of the nonhydrostatic kernels (vertical order 1, terrain-following
 * metric with level-independent layer depth): the 3-D metric arrays are
 * replaced by 13 constants per column, checked against the uploaded arrays
 * (relative deviation <= 1e-13) before the path is enabled.  Returns 1 when the
 * fast kernels are in use, 0 when the general kernels run (reason: see
 * tb200_fast_path_reason), -1 on error.  TB200_STAGE_KERNEL=generic disables it. */
int tb200_fast_path(tb200_ctx * ctx);
const char * tb200_fast_path_reason(tb200_ctx * ctx);
double tb200_fast_path_metric_error(tb200_ctx * ctx);

/*
 * Connectivity.  The reference finds coincident nodes through its exchange
 * buffer topology (Grid.cpp:1066-1573, Connectivity.cpp:47-744); here the host
 * names them: node_ids[W_Aint][W_Bint] (interior nodes only, no halo) gives
 * every element-local node a global id, equal ids are one physical node.
 * Direct stiffness summation averages over equal ids
 * (GridCSGLL::ApplyDSS, GridCSGLL.cpp:435-781).
 */
int tb200_set_node_ids(tb200_ctx * ctx, int patch_index, const int64_t * node_ids);
/* (ids are supplied for every patch of the grid, local or not, so that each
 * rank can derive the same send/receive slot order without communication) */
/* Nodes on a panel seam: covector re-basing of (u_alpha,u_beta) from a source
 * panel to this node's panel (CubedSphereTrans::CoVecPanelTrans,
 * CubedSphereTrans.h:1751-2275; GridPatchCSGLL::TransformHaloVelocities,
 * GridPatchCSGLL.cpp:1783-1924).  For n seam nodes: ia[n], ib[n] (interior
 * indices), src_panel[n], m[n][4] row-major 2x2 so that
 * (ua,ub)_this = M (ua,ub)_src. */
int tb200_set_seam_transforms(tb200_ctx * ctx, int patch_index, int n,
                              const int * ia, const int * ib,
                              const int * src_panel, const double * m);
/* Build the device-side averaging groups from the ids above. */
int tb200_build_connectivity(tb200_ctx * ctx);

/* ---- state movement (DataContainer / DataArray4D seam) -------- */

/* Host (reference layout, halo included) -> device instance.  node/redge are
 * GridPatch::GetDataState(inst, DataLocation_Node / _REdge), tracers is
 * GetDataTracers(inst); NULL skips.  Only the components located at the array
 * (Grid.cpp:281-287) are transferred. */
int tb200_upload_state(tb200_ctx * ctx, int patch_index, int inst,
                       const double * node, const double * redge,
                       const double * tracers);
/* Device instance -> host; also fills the derived slots the reference keeps
 * (W on nodes, U,V on interfaces: HorizontalDynamicsFEM.cpp:817-831) when
 * fill_derived != 0.  Halo entries are left untouched. */
int tb200_download_state(tb200_ctx * ctx, int patch_index, int inst,
                         double * node, double * redge, double * tracers,
                         int fill_derived);

/* The same without waiting for the bus: the copies and layout conversions of
 * consecutive calls overlap (two staging buffers, a copy stream).  The host
 * arrays of an upload may be changed, and those of a download read, after
 * tb200_transfer_sync.  Host memory should be pinned (tb200_host_register)
 * for the copies to run asynchronously. */
int tb200_upload_state_async(tb200_ctx * ctx, int patch_index, int inst,
                             const double * node, const double * redge,
                             const double * tracers);
int tb200_download_state_async(tb200_ctx * ctx, int patch_index, int inst,
                               double * node, double * redge, double * tracers,
                               int fill_derived);
int tb200_transfer_sync(tb200_ctx * ctx);
/* Pin / unpin the host memory state arrays live in: the reference's
 * DataContainer blocks (src/base/DataContainer.cpp:77-147; one contiguous
 * allocation per container, GridPatch.cpp:287-480). */
int tb200_host_register(tb200_ctx * ctx, void * ptr, size_t bytes);
int tb200_host_unregister(tb200_ctx * ctx, void * ptr);

/* ---- Grid::CopyData / LinearCombineData / ZeroData (Grid.cpp:1585-1632,
 *      GridPatch.cpp:1402-1553) ------------------------------------------- */
int tb200_copy(tb200_ctx * ctx, int src, int dst, int data_mask);
int tb200_lincomb(tb200_ctx * ctx, const double * coeff, int ncoeff, int dst,
                  int data_mask);
int tb200_zero(tb200_ctx * ctx, int inst, int data_mask);

/* ---- dynamics plugins ---------------------------------------------------- */

/* HorizontalDynamicsFEM::StepExplicit (HorizontalDynamicsFEM.cpp:1787-1863):
 * StepShallowWater (:321-647) or StepNonhydrostaticPrimitive (:701-1783). */
int tb200_h_step_explicit(tb200_ctx * ctx, int in, int out, double dt);
/* VerticalDynamicsFEM::StepExplicit (VerticalDynamicsFEM.cpp:616-1159). */
int tb200_v_step_explicit(tb200_ctx * ctx, int in, int out, double dt);
/* Both of the above in one pass over the state (same results). */
int tb200_hv_step_explicit(tb200_ctx * ctx, int in, int out, double dt);
/* Grid::LinearCombineData(coeff, out) (GridPatch.cpp:1433-1520) followed by
 * both explicit plugins, the combination formed inside the stage kernel:
 * HorizontalDynamics::StepExplicitCombine (HorizontalDynamics.h:97-106). */
int tb200_hv_step_explicit_combine(tb200_ctx * ctx, const double * coeff, int ncoeff,
                                   int in, int out, double dt);
/* ... followed by PostProcessSubstage(out) (DSS of state and tracers): the explicit
 * substage of every time scheme.  On several ranks the halo exchange overlaps the
 * elements that do not feed it when TB200_OVERLAP=1. */
int tb200_hv_step_explicit_combine_dss(tb200_ctx * ctx, const double * coeff, int ncoeff,
                                       int in, int out, double dt);
/* VerticalDynamicsFEM::StepImplicit (VerticalDynamicsFEM.cpp:1230-1638). */
int tb200_v_step_implicit(tb200_ctx * ctx, int in, int out, double dt);
/* Grid::CopyData(src -> dst) + VerticalDynamics::StepImplicit(dst, dst, ...) as the
 * time schemes issue them back to back (TimestepSchemeStrang.cpp:644-650,
 * TimestepSchemeARS343.cpp:169-172): only the rows the solve does not
 * overwrite are copied. */
int tb200_copy_v_step_implicit(tb200_ctx * ctx, int src, int dst, double dt);
/* ... followed by Grid::LinearCombineData({+1 dst, -1 src} -> src): dst = solve(src),
 * src = dst - src (the tail of TimestepSchemeStrang::Step, :644-672). */
int tb200_copy_v_step_implicit_diff(tb200_ctx * ctx, int src, int dst, double dt);
/* The same tail when the state to solve already sits in `inst`
 * (StepAfterSubCycle written straight into it): StepImplicit(inst, inst) in
 * place, `inc` = new - old state (zero in the u, v rows).  Returns 2 without
 * doing anything when the fast column kernel does not apply
 * (tb200_v_step_implicit_inc_available == 0). */
int tb200_v_step_implicit_inc_available(tb200_ctx * ctx);
int tb200_v_step_implicit_inc(tb200_ctx * ctx, int inst, int inc, double dt);
/* GridGLL::PostProcessSubstage -> ApplyDSS (GridGLL.cpp:571-583,
 * GridCSGLL.cpp:435-781, GridCartesianGLL.cpp:508-654). */
int tb200_dss(tb200_ctx * ctx, int inst, int data_mask);
/* HorizontalDynamicsFEM::StepAfterSubCycle (HorizontalDynamicsFEM.cpp:2637-2726):
 * scalar + vector hyperdiffusion, DSS, Rayleigh friction. */
int tb200_h_step_after_subcycle(tb200_ctx * ctx, int in, int out, int work, double dt);
/* VerticalDynamics::FilterNegativeTracers + HorizontalDynamicsFEM one
 * (VerticalDynamicsFEM.cpp:4286-4347, HorizontalDynamicsFEM.cpp:213-317). */
int tb200_filter_negative_tracers(tb200_ctx * ctx, int inst);
/* VerticalDynamicsFEM::FilterNegativeTracers (VerticalDynamicsFEM.cpp:4286-4347):
 * the column-wise filter (the one above is HorizontalDynamicsFEM's element-wise
 * filter, HorizontalDynamicsFEM.cpp:213-317). */
int tb200_v_filter_negative_tracers(tb200_ctx * ctx, int inst);
/* ---- output-side interpolation -----------------------------------------
 * Grid::ReduceInterpolate (src/atm/Grid.cpp:866-990) -> GridPatchCSGLL::InterpolateData
 * (src/atm/GridPatchCSGLL.cpp:1365-1780) of state instance `inst` on the device: the
 * state (data_type TB200_DATA_STATE; only_location -1 every component, 0 those on
 * levels, 1 those on interfaces = eOnlyVariablesAt) or the tracers
 * (TB200_DATA_TRACERS) at npts points and nout REta values.  Per point: its patch,
 * the element it lies in (indices inside the patch, :1605-1627), the np coefficients
 * of PolynomialInterp::LagrangianPolynomialCoeffs along alpha and beta on that
 * element's nodes (:1631-1641), alpha and beta themselves (wind conversion).  Per
 * location the LinearColumnInterpFEM operator as a dense [nout][levels or
 * interfaces] matrix with its [begin, end) windows (null = identity, nout = that
 * count).  convert_to_primitive: w divided by DerivR[2], the covariant wind
 * converted to zonal / meridional components (CubedSphereTrans::CoVecTransRLLFromABP,
 * :1655-1690, 1735-1776).  The reference state is included (fIncludeReferenceState).
 * out: host array [components or tracers][nout][npts]; points of patches that are
 * not local stay zero (the reference sums the ranks' arrays). */
/* Grid::ComputeVorticityDivergence (GridPatchCSGLL::ComputeCurlAndDiv,
 * src/atm/GridPatchCSGLL.cpp:1132-1361) and Grid::ComputeTemperature
 * (src/atm/GridPatch.cpp:641-700) of state instance `inst`: relative vorticity and
 * divergence of the horizontal wind element by element, temperature (nonhydrostatic
 * equations) on levels, kept on the device for tb200_interpolate with data_type
 * TB200_DATA_VORTICITY / _DIVERGENCE / _TEMPERATURE (one component, on levels). */
int tb200_compute_output_fields(tb200_ctx * ctx, int inst);
int tb200_interpolate(tb200_ctx * ctx, int inst, int data_type, int only_location,
                      int npts, const int * patch_index, const int * elem_a, const int * elem_b,
                      const double * ca, const double * cb,
                      const double * alpha, const double * beta, int nout,
                      const double * vop_node, const int * vbegin_node, const int * vend_node,
                      const double * vop_redge, const int * vbegin_redge, const int * vend_redge,
                      int convert_to_primitive, double * out);

/* ---- device-side set-up of cubed-sphere runs ------------------------
 * The 2-D metric and the initial state evaluated on the device instead of being
 * built on the host and copied over the bus.  Needs tb200_set_terrain_metric
 * (X = tan alpha, Y = tan beta per node), tb200_set_vertical_coordinate and, for
 * the state, the topography (tb200_evaluate_jw_topography or tb200_upload_geometry). */

/* What GridPatchCSGLL::EvaluateGeometricTerms (src/atm/GridPatchCSGLL.cpp:295-343)
 * leaves in GetJacobian2D(), GetContraMetric2DA/B(), GetCoriolisF(), GetLongitude()
 * and GetLatitude() (CubedSphereTrans::RLLFromXYP, CubedSphereTrans.cpp:200-266);
 * the latitude also feeds tb200_held_suarez. */
int tb200_evaluate_geometry_cs(tb200_ctx * ctx, int patch_index, double radius, double omega);

/* Debugging aid (cf. tb200_debug_column_assembly): a per-column array in the device's
 * element-major order [element][np * np]; which = 0 Jacobian2D, 1, 2 ContraMetric2DA,
 * 3, 4 ContraMetric2DB, 5 CoriolisF, 6 topography, 7 longitude, 8 latitude,
 * 9 accumulated precipitation of tb200_kessler. */
int tb200_debug_column_field(tb200_ctx * ctx, int which, double * out);

/* Parameters of BaroclinicWaveJWTest (test/nonhydro_sphere/BaroclinicWaveJWTest.cpp:41-134)
 * and the constants of the sphere it takes from PhysicalConstants. */
typedef struct {
	double eta0, tropopause_eta, t0, delta_t, lapse_rate, u0, up;
	double pert_lon, pert_lat, pert_r;
	int perturbation;        /* 1: PerturbationType_Exp, 0: none */
	double omega, radius;    /* PhysicalConstants::GetOmega(), GetEarthRadius() */
} tb200_jw_test;

/* BaroclinicWaveJWTest::EvaluateTopography (:170-204) on the nodes of a patch. */
int tb200_evaluate_jw_topography(tb200_ctx * ctx, int patch_index, const tb200_jw_test * test);
/* GridPatchCSGLL::EvaluateTestCase (GridPatchCSGLL.cpp:578-920) with
 * BaroclinicWaveJWTest::EvaluatePointwiseState (:297-413): covariant u_alpha, u_beta,
 * rho theta, rho on levels and w = 0 on interfaces of state instance `inst`.  Fails with
 * the reference's "Maximum number of iterations exceeded." when the Newton
 * iteration for eta does not converge. */
int tb200_evaluate_jw_state(tb200_ctx * ctx, int patch_index, int inst, const tb200_jw_test * test);

/* HeldSuarezPhysics::Perform (src/atm/HeldSuarezPhysics.cpp:62-301) as a device
 * workflow step on instance 0: boundary-layer friction of u, v and Newtonian
 * relaxation of rho-theta, one streaming pass.  Inputs per column, [iA][iB] in the
 * reference layout with halo: GridPatch::GetLatitude() and the product
 * dataREdge[R][..][0] * dataREdge[T][..][0] of instance 0 the reference takes its
 * "surface pressure" from (:112-115; the dynamics never changes those slots). */
int tb200_upload_held_suarez(tb200_ctx * ctx, int patch_index,
                             const double * latitude, const double * surface_product);
int tb200_held_suarez(tb200_ctx * ctx, double dt);
/* KesslerPhysics::Perform (test/dcmip2016/KesslerPhysics.cpp:84-285, calling KESSLER,
 * test/dcmip2016/interface/kessler.f90:62-185) on instance 0: warm-rain microphysics per
 * column on rho theta, rho and the first three tracers (rho qv, rho qc, rho qr);
 * precipitation accumulates per column (UserData2D[0]; tb200_debug_column_field(9)).
 * PARITY UNPINNED - the reference kernel is Fortran and cannot be built in this image;
 * the device kernel is held to the C restatement oracle/kessler_port.c. */
int tb200_kessler(tb200_ctx * ctx, double dt);
/* Grid::LinearCombineData(coeff -> dst) (GridPatch.cpp:1433-1520) of state and
 * tracers followed by VerticalDynamics::FilterNegativeTracers(dst), as
 * TimestepSchemeStrang::Step issues them at the start of a step (:470-482);
 * the tracer combination is formed inside the filter kernel. */
int tb200_lincomb_v_filter(tb200_ctx * ctx, const double * coeff, int ncoeff, int dst);

/* TimestepScheme::Step (TimestepSchemeStrang.cpp:450-674,
 * TimestepSchemeARS343.cpp:146-235, ...): one full time step on the device. */
/* Scheme id of a --timescheme string (TempestInitialize.h:192-291, lower case:
 * strang[/kgu35|fe|rk4|rk3|ssprk53], erk[/...], ars222, ars232, ars343, ars443);
 * -1 if the scheme is not implemented. */
/* VerticalDynamicsFEM::StepImplicitTermsExplicitly (VerticalDynamicsFEM.cpp:439-612;
 * TimestepSchemeARK232): out -= dt * BuildF(in) on rho theta, w, rho of every node. */
int tb200_v_step_implicit_terms_explicitly(tb200_ctx * ctx, int in, int out, double dt);
int tb200_scheme_from_name(const char * name);
int tb200_scheme_instances(int scheme);
int tb200_step(tb200_ctx * ctx, int scheme, int first_step, int last_step, double dt);

/* ---- diagnostics (Grid::Checksum, GridPatch.cpp:744-835) ----------------- */
/* Area-weighted sum of every component of an instance over the local patches;
 * sums[ncomp].  element_area_* in the reference layout are supplied once. */
/* Rayleigh friction: GridPatch::GetRayleighStrength(Node / REdge) [W_A][W_B][L(+1)]
 * and GridPatch::GetReferenceState(Node / REdge) in the state layout.  Once
 * uploaded, tb200_h_step_after_subcycle applies
 * HorizontalDynamicsFEM::ApplyRayleighFriction (:2418-2536) after the
 * hyperdiffusion (APPLY_RAYLEIGH_WITH_HYPERVIS, Defines.h:74). */
/* --vdisc FV (Grid::VerticalDiscretization_FiniteVolume; the default is the finite-
 * element discretisation): the column operators passed to tb200_set_column_op are
 * then the reference's finite-volume ones; on this side every level becomes its own
 * element for the penalty terms and the Jacobian band is the narrower one of
 * VerticalDynamicsFEM.cpp:174-185.  General kernels.  Call before the first step. */
int tb200_set_vertical_discretization(tb200_ctx * ctx, int finite_volume);
/* --vmassfluxlevels (fForceMassFluxOnLevels of the VerticalDynamicsFEM constructor):
 * BuildF forms the mass and rho-theta fluxes on levels and differentiates them with
 * TB200_OP_DIFF_N2N_ZB (VerticalDynamicsFEM.cpp:2229-2243, 2301-2315); the Jacobian
 * stays that of the interface fluxes, as in the reference.  General kernels. */
int tb200_set_mass_flux_on_levels(tb200_ctx * ctx, int on);
/* GridPatch::GetReferenceState(Node / REdge) of a local patch alone (zero until
 * uploaded, as in a test case without TestCase::HasReferenceState). */
int tb200_upload_reference_state(tb200_ctx * ctx, int patch_index,
                                 const double * ref_node, const double * ref_redge);
/* Uniform diffusion (Grid::HasUniformDiffusion, Grid.cpp:399-415, coefficients from
 * TestCase::GetUniformDiffusionCoeffs): second-order diffusion of the state minus the
 * reference state - horizontally at the end of HorizontalDynamicsFEM::StepExplicit
 * (:1817-1858: u, v; rho theta; w), in the column in VerticalDynamicsFEM::StepExplicit
 * (u, v, :1058-1106) and BuildF (rho theta, w, :2594-2636; not in the Jacobian).
 * General kernels; with tracers the step fails as the reference does (:3914-3917).
 * The column terms remove the reference state, as with the default fUseReferenceState
 * of the VerticalDynamicsFEM constructor (--norefstate is not restated).
 * Both zero = off (the default). */
int tb200_set_uniform_diffusion(tb200_ctx * ctx, double scalar_coeff, double vector_coeff);
int tb200_upload_rayleigh(tb200_ctx * ctx, int patch_index,
                          const double * strength_node, const double * strength_redge,
                          const double * ref_node, const double * ref_redge);
int tb200_upload_element_area(tb200_ctx * ctx, int patch_index,
                              const double * area_node, const double * area_redge);
int tb200_checksum(tb200_ctx * ctx, int inst, double * sums);
/* Conservation diagnostics of an instance over the local patches
 * (Grid::ComputeTotalEnergy, Grid.cpp:529-556 / GridPatch.cpp:925-1138;
 * Grid::ComputeTotalPotentialEnstrophy, Grid.cpp:560-590 / GridPatch.cpp:1142-1230;
 * Grid::ComputeTotalVerticalMomentum, Grid.cpp:594-623 / GridPatch.cpp:1234-1288).
 * W on levels and rho on interfaces, which the reference's routines read and
 * the Lorenz-staggered state does not carry, are formed with
 * Grid::InterpolateREdgeToNode / InterpolateNodeToREdge (Grid.cpp:843-863).
 * Shallow water: the potential enstrophy needs the DSS'd relative vorticity
 * (GridGLL::ComputeVorticityDivergence, GridGLL.cpp:587-602), built in the
 * scratch instance `work`; other equation sets ignore `work`. */
int tb200_total_energy(tb200_ctx * ctx, int inst, double * energy);
int tb200_total_potential_enstrophy(tb200_ctx * ctx, int inst, int work, double * enstrophy);
int tb200_total_vertical_momentum(tb200_ctx * ctx, int inst, double * momentum);

/* ---- multi-GPU halo traffic (Grid::Exchange, Grid.cpp:627-685) ----------- */
/* Patches are partitioned over ranks (one process per GPU).  Per exchange the
 * library packs the local nodes that other ranks share into a device send
 * buffer ordered [destination rank][slot][row], calls the callback - which
 * must move send_counts[r] doubles to rank r and receive recv_counts[r]
 * doubles from rank r into recvbuf (same ordering), enqueued on the stream of
 * tb200_set_stream, e.g. one NCCL all-to-all - and then averages.
 * Must be called before tb200_add_patch. */
typedef int (*tb200_exchange_fn)(void * user, double * sendbuf, double * recvbuf,
                                 const int64_t * send_counts,
                                 const int64_t * recv_counts, int nranks);
int tb200_set_exchange(tb200_ctx * ctx, int rank, int nranks,
                       tb200_exchange_fn fn, void * user);
/* Peer-memory exchange over NVLink / NVSwitch (replaces the callback once
 * attached; stands in for the MPI_Isend / MPI_Irecv / MPI_Waitall of
 * Grid::Exchange, Grid.cpp:627-685, Connectivity.cpp:941-1120): the pack kernel stores every shared node straight into the
 * receive buffer of the rank that averages it and raises a flag there; the
 * consumer's stream waits on its flags.  After tb200_build_connectivity every
 * rank calls tb200_peer_export (allocates its receive area; handle = 64-byte
 * CUDA IPC handle, recv_offsets[r] = first slot of source rank r,
 * recv_total = slots in all), the host all-gathers the three, and every rank
 * calls tb200_peer_attach with handles[nranks][64], my_offset_at[r] = where
 * rank r expects this rank's slots (= rank r's recv_offsets[this rank]) and
 * recv_totals[r].  All ranks of one node, one GPU each. */
int tb200_peer_export(tb200_ctx * ctx, void * handle, int64_t * recv_offsets,
                      int64_t * recv_total);
int tb200_peer_attach(tb200_ctx * ctx, const void * handles,
                      const int64_t * my_offset_at, const int64_t * recv_totals);
/* Back to the callback exchange (e.g. another rank could not attach). */
int tb200_peer_detach(tb200_ctx * ctx);
/* Nodes sent to / received from each rank per exchange (after
 * tb200_build_connectivity); arrays of nranks entries. */
int tb200_exchange_counts(tb200_ctx * ctx, int64_t * send_nodes, int64_t * recv_nodes);

/* ---- timing hooks (FunctionTimer groups) ----------------------- */
/* The reference times its plugins with FunctionTimer groups
 * ("HorizontalStepNonhydrostaticPrimitive", HorizontalDynamicsFEM.cpp:708;
 * "StepAfterSubCycle", :2645; "VerticalStepExplicit", VerticalDynamicsFEM.cpp:630;
 * "VerticalStepImplicit", :1237; "Communicate", Grid.cpp:636) and prints their
 * averages at the end of a run (Model.cpp:640-688).  With hooks set, every
 * entry point that stands for one of those groups calls begin(user, group)
 * when it starts and - after waiting for its device work - end(user, group),
 * also inside tb200_step; the C++ shells open and close the reference's own
 * FunctionTimer in them.  The fused explicit stage (horizontal + vertical
 * explicit in one kernel) reports under the horizontal group.  NULL hooks
 * (default): no synchronisation, no calls. */
typedef void (*tb200_timing_fn)(void * user, const char * group);
int tb200_set_timing_hooks(tb200_ctx * ctx, tb200_timing_fn begin, tb200_timing_fn end,
                           void * user);

/* ---- introspection for tests and the bench -------------------------------- */
/* Number of kernels this context has launched since creation. */
int64_t tb200_launch_count(const tb200_ctx * ctx);
/* Averaging groups (shared nodes) that the stage and hyperdiffusion kernels
 * average themselves when a DSS follows them (all members neighbours inside one
 * patch); 0 when the fused DSS is unavailable.  TB200_DSS_FUSED=0 turns it off. */
int64_t tb200_fused_group_count(const tb200_ctx * ctx);
/* Total columns (element-local nodes incl. duplicates) held by this context. */
int64_t tb200_column_count(const tb200_ctx * ctx);
/* Direct banded solve used by the implicit step, exposed for pinning against
 * LAPACK dgbsv (src/base/LinearAlgebra.cpp:156-202): ncols independent systems,
 * ab[ncols][n][ldab] in the reference's row-major band storage, b[ncols][n]
 * overwritten with the solution.  Runs on the device. */
int tb200_test_band_solve(tb200_ctx * ctx, int ncols, int n, int kl, int ku,
                          const double * ab, double * b);

/* Debugging aid (cf. USE_JACOBIAN_DEBUG / BootstrapJacobian,
 * VerticalDynamicsFEM.cpp:1163-1226): assemble F and the banded Jacobian of
 * the implicit solve for instance `in` without solving and return the work
 * arrays of unique column `col` (24 arrays of nlev+1, x0, F, band matrix). */
int tb200_debug_column_assembly(tb200_ctx * ctx, int in, double dt, int col,
                                double * ws_out, int nentries);

#ifdef __cplusplus
}
#endif

#endif
