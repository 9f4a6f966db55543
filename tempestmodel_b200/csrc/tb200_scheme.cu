// TimestepScheme::Step on the device, and Grid::Checksum.
//
// Stage sequencing of the reference's time schemes: every Grid::CopyData /
// LinearCombineData / StepExplicit / StepImplicit / PostProcessSubstage call
// of the reference becomes the matching C-ABI call on device-resident state
// instances, in the same order with the same coefficients.
//   TimestepSchemeStrang::Step   reference TimestepSchemeStrang.cpp:450-674
//   TimestepSchemeARS343::Step   reference TimestepSchemeARS343.cpp:146-235
//   TimestepSchemeARS222::Step   reference TimestepSchemeARS222.cpp:50-118
//   TimestepSchemeARS232::Step   reference TimestepSchemeARS232.cpp:51-150
//   TimestepSchemeARS443::Step   reference TimestepSchemeARS443.cpp:52-191
//   TimestepSchemeERK::Step      reference TimestepSchemeERK.cpp:143-304

#include <cstring>
#include <cmath>
#include <vector>

#include "tb200_ctx.h"

#define TB_FAIL(ctx, msg) \
	do { (ctx)->err = (msg); return 1; } while (0)

#define TB_CHECK(ctx, call) \
	do { \
		cudaError_t e__ = (call); \
		if (e__ != cudaSuccess) { \
			(ctx)->err = std::string(#call) + ": " + cudaGetErrorString(e__); \
			return 1; \
		} \
	} while (0)

#define TRY(call) do { if ((call) != 0) return 1; } while (0)

static const int ALL = TB200_DATA_STATE | TB200_DATA_TRACERS;

// --timescheme strings of TempestInitialize.h:192-291 (lower case)
extern "C" int tb200_scheme_from_name(const char * name) {
	static const struct { const char * n; int id; } tab[] = {
		{"strang", TB200_SCHEME_STRANG_KGU35}, {"strang/kgu35", TB200_SCHEME_STRANG_KGU35},
		{"strang/fe", TB200_SCHEME_STRANG_FE}, {"strang/rk4", TB200_SCHEME_STRANG_RK4},
		{"strang/rk3", TB200_SCHEME_STRANG_SSP3}, {"strang/ssprk53", TB200_SCHEME_STRANG_SSPRK53},
		{"erk", TB200_SCHEME_ERK_KGU35}, {"erk/kgu35", TB200_SCHEME_ERK_KGU35},
		{"erk/fe", TB200_SCHEME_ERK_FE}, {"erk/rk4", TB200_SCHEME_ERK_RK4},
		{"erk/rk3", TB200_SCHEME_ERK_SSP3}, {"erk/ssprk53", TB200_SCHEME_ERK_SSPRK53},
		{"ars222", TB200_SCHEME_ARS222}, {"ars232", TB200_SCHEME_ARS232},
		{"ars343", TB200_SCHEME_ARS343}, {"ars443", TB200_SCHEME_ARS443},
		{"gark2", TB200_SCHEME_GARK2}, {"ssp3_332", TB200_SCHEME_SSP3332},
		{"ark232", TB200_SCHEME_ARK232}};
	if (name == 0) return -1;
	for (size_t q = 0; q < sizeof(tab) / sizeof(tab[0]); q++) {
		if (strcmp(name, tab[q].n) == 0) return tab[q].id;
	}
	return -1;
}

extern "C" int tb200_scheme_instances(int scheme) {
	switch (scheme) {
		case TB200_SCHEME_STRANG_KGU35:
		case TB200_SCHEME_STRANG_RK4:
		case TB200_SCHEME_STRANG_SSP3:
		case TB200_SCHEME_STRANG_FE:
		case TB200_SCHEME_STRANG_SSPRK53:
			return 5;   // TimestepSchemeStrang.h:61-70
		case TB200_SCHEME_ERK_KGU35:
		case TB200_SCHEME_ERK_FE:
		case TB200_SCHEME_ERK_RK4:
		case TB200_SCHEME_ERK_SSP3:
		case TB200_SCHEME_ERK_SSPRK53:
			return 5;   // TimestepSchemeERK.h:61-70
		case TB200_SCHEME_ARS343:
			return 7;   // TimestepSchemeARS343.h:49-58
		case TB200_SCHEME_ARS222:
			return 4;   // TimestepSchemeARS222.h:48-57
		case TB200_SCHEME_ARS232:
			return 7;   // TimestepSchemeARS232.h:48-57
		case TB200_SCHEME_ARS443:
			return 10;  // TimestepSchemeARS443.h:48-57
		case TB200_SCHEME_GARK2:
			return 5;   // TimestepSchemeGARK2.h:48-57
		case TB200_SCHEME_SSP3332:
			return 9;   // TimestepSchemeSSP3332.h:48-57
		case TB200_SCHEME_ARK232:
			return 8;   // TimestepSchemeARK232.h:51-60
		default:
			return -1;
	}
}

// One explicit substage: H.StepExplicit, V.StepExplicit, DSS of state and
// tracers (e.g. TimestepSchemeStrang.cpp:548-553).
static int substage(tb200_ctx * ctx, int in, int out, double dt) {
	TRY(tb200_hv_step_explicit(ctx, in, out, dt));
	TRY(tb200_dss(ctx, out, ALL));
	return 0;
}

static int lincomb(tb200_ctx * ctx, const std::vector<double> & c, int dst) {
	return tb200_lincomb(ctx, c.data(), (int)c.size(), dst, ALL);
}

// CopyData(src -> out) / LinearCombineData(c -> out) + explicit substage, the
// combination formed inside the stage kernel
static int substage_from(tb200_ctx * ctx, std::vector<double> c, int in, int out, double dt) {
	if ((int)c.size() <= out) c.resize(out + 1, 0.0);
	TRY(tb200_hv_step_explicit_combine_dss(ctx, c.data(), (int)c.size(), in, out, dt));
	return 0;
}

static std::vector<double> copy_of(int src) {
	std::vector<double> c(src + 1, 0.0);
	c[src] = 1.0;
	return c;
}

///////////////////////////////////////////////////////////////////////////////

static int step_strang(tb200_ctx * ctx, int scheme, int first, int last, double dt) {
	const double offc = ctx->cfg.off_centering;
	const double half = 0.5 * dt;

	// :470-482
	if (first) {
		TRY(tb200_v_step_implicit(ctx, 0, 0, half));
	} else {
		const std::vector<double> carry = {1.0, 1.0};
		// LinearCombineData + pVerticalDynamics->FilterNegativeTracers(0),
		// TimestepSchemeStrang.cpp:470-480
		TRY(tb200_lincomb_v_filter(ctx, carry.data(), (int)carry.size(), 0));
	}

	if (scheme == TB200_SCHEME_STRANG_FE) {
		// :485-492
		TRY(tb200_copy(ctx, 0, 4, ALL));
		TRY(substage(ctx, 0, 4, dt));

	} else if (scheme == TB200_SCHEME_STRANG_RK4) {
		// :495-522
		TRY(tb200_copy(ctx, 0, 1, ALL));
		TRY(substage(ctx, 0, 1, half));
		TRY(tb200_copy(ctx, 0, 2, ALL));
		TRY(substage(ctx, 1, 2, half));
		TRY(tb200_copy(ctx, 0, 3, ALL));
		TRY(substage(ctx, 2, 3, dt));
		const std::vector<double> rk4 = {-1.0 / 3.0, 1.0 / 3.0, 2.0 / 3.0, 1.0 / 3.0, 0.0};
		TRY(lincomb(ctx, rk4, 4));
		TRY(substage(ctx, 3, 4, dt / 6.0));

	} else if (scheme == TB200_SCHEME_STRANG_SSP3) {
		// :525-546
		TRY(tb200_copy(ctx, 0, 1, ALL));
		TRY(substage(ctx, 0, 1, dt));
		const std::vector<double> a = {3.0 / 4.0, 1.0 / 4.0, 0.0};
		TRY(lincomb(ctx, a, 2));
		TRY(substage(ctx, 1, 2, 0.25 * dt));
		const std::vector<double> b = {1.0 / 3.0, 0.0, 2.0 / 3.0, 0.0, 0.0};
		TRY(lincomb(ctx, b, 4));
		TRY(substage(ctx, 2, 4, (2.0 / 3.0) * dt));

	} else if (scheme == TB200_SCHEME_STRANG_SSPRK53) {
		// :586-627
		const double s1 = 0.377268915331368;
		TRY(substage_from(ctx, copy_of(0), 0, 1, s1 * dt));
		TRY(substage_from(ctx, copy_of(1), 1, 2, s1 * dt));
		const std::vector<double> a = {0.355909775063327, 0.0, 0.644090224936674, 0.0};
		TRY(substage_from(ctx, a, 2, 3, 0.242995220537396 * dt));
		const std::vector<double> b = {0.367933791638137, 0.0, 0.0, 0.632066208361863};
		TRY(substage_from(ctx, b, 3, 0, 0.238458932846290 * dt));
		const std::vector<double> c = {0.762406163401431, 0.0, 0.237593836598569, 0.0, 0.0};
		TRY(substage_from(ctx, c, 0, 4, 0.287632146308408 * dt));

	} else {
		// Kinnmark-Gray-Ullrich (3,5), :548-585
		TRY(substage_from(ctx, copy_of(0), 0, 1, dt / 5.0));
		TRY(substage_from(ctx, copy_of(0), 1, 2, dt / 5.0));
		TRY(substage_from(ctx, copy_of(0), 2, 3, dt / 3.0));
		TRY(substage_from(ctx, copy_of(0), 3, 2, 2.0 * dt / 3.0));
		const std::vector<double> kgu = {-1.0 / 4.0, 5.0 / 4.0, 0.0, 0.0, 0.0};
		TRY(substage_from(ctx, kgu, 2, 4, 3.0 * dt / 4.0));
	}

	const double dOffCenterDeltaT = 0.5 * (1.0 + offc) * dt;
	if (offc == 0.0 && !last && tb200_v_step_implicit_inc_available(ctx)) {
		// hyperdiffusion written straight into instance 0 (no CopyData(1 -> 0)),
		// solve in place, instance 1 receives the {+1, -1} combination
		TRY(tb200_h_step_after_subcycle(ctx, 4, 0, 2, dt));
		TRY(tb200_v_step_implicit_inc(ctx, 0, 1, dOffCenterDeltaT));
		return 0;
	}

	// hyperdiffusion, :638-641 (the CopyData(4 -> 1) in front of it repeats the
	// one StepAfterSubCycle starts with, HorizontalDynamicsFEM.cpp:2661)
	TRY(tb200_h_step_after_subcycle(ctx, 4, 1, 2, dt));

	// vertical step, :644-657: CopyData(1 -> 0), StepImplicit(0, 0)
	if (offc == 0.0 && !last) {
		// the off-centring combination is 1 * instance 0: the solve and the final
		// {+1, -1} combination run as one call
		TRY(tb200_copy_v_step_implicit_diff(ctx, 1, 0, dOffCenterDeltaT));
		return 0;
	}
	TRY(tb200_copy_v_step_implicit(ctx, 1, 0, dOffCenterDeltaT));
	const std::vector<double> oc = {(2.0 - offc) / 2.0, offc / 2.0};
	TRY(lincomb(ctx, oc, 0));
	if (!last) {
		const std::vector<double> fin = {+1.0, -1.0};
		TRY(lincomb(ctx, fin, 1));
	}
	return 0;
}

///////////////////////////////////////////////////////////////////////////////
// ARS(3,4,3) of Ascher, Ruuth & Spiteri (1997): tableau and the stage
// combinations that express each new stage base through stored instances
// (reference TimestepSchemeARS343.cpp:30-141).

struct ARS343Coefficients {
	double diag_exp[4];
	double diag_imp[4];
	std::vector<double> u2, u3, u4;
};

static ARS343Coefficients ars343_coefficients() {
	ARS343Coefficients k;
	const double gam = 0.4358665215084590;
	const double b1 = -1.5 * gam * gam + 4.0 * gam - 0.25;
	const double b2 = 1.5 * gam * gam - 5.0 * gam + 1.25;
	const double a42 = 0.5529291480359398;
	const double a43 = 0.5529291480359398;
	const double a31 =
		(1.0 - 4.5 * gam + 1.5 * gam * gam) * a42
		+ (2.75 - 10.5 * gam + 3.75 * gam * gam) * a43
		- 3.5 + 13 * gam - 4.5 * gam * gam;
	const double a32 =
		(-1.0 + 4.5 * gam - 1.5 * gam * gam) * a42
		+ (-2.75 + 10.5 * gam - 3.75 * gam * gam) * a43
		+ 4.0 - 12.5 * gam + 4.5 * gam * gam;
	const double a41 = 1.0 - a42 - a43;
	// rows = stages 1..4, columns = earlier stages
	const double I[4][4] = {
		{gam, 0., 0., 0.},
		{0.5 * (1.0 - gam), gam, 0., 0.},
		{b1, b2, gam, 0.},
		{b1, b2, gam, 0.}};
	const double E[4][4] = {
		{gam, 0., 0., 0.},
		{a31, a32, 0., 0.},
		{a41, a42, a43, 0.},
		{0., b1, b2, gam}};
	for (int s = 0; s < 4; s++) {
		k.diag_exp[s] = E[s][s];
		k.diag_imp[s] = I[s][s];
	}
	// raw combination of stage row r (1..3): instance 0 carries u^n, the pair
	// (2j+1, 2j+2) carries the explicit / implicit result of stage j
	std::vector<double> raw[4];
	for (int r = 1; r < 4; r++) {
		raw[r].assign(7, 0.0);
		raw[r][0] = 1.0 - E[r][0] / E[0][0];
		for (int j = 0; j < r; j++) {
			raw[r][2 * j + 1] = E[r][j] / E[j][j] - I[r][j] / I[j][j];
			raw[r][2 * j + 2] = I[r][j] / I[j][j];
		}
	}
	// the explicit increment of a stage j >= 1 is measured from that stage's
	// own base, itself a combination of earlier instances: fold it back
	k.u2 = raw[1];
	k.u3 = raw[2];
	k.u4 = raw[3];
	const double c37 = -E[2][1] / E[1][1];
	for (int q = 0; q < 3; q++) k.u3[q] += c37 * k.u2[q];
	const double c47 = -E[3][1] / E[1][1];
	const double c48 = -E[3][2] / E[2][2];
	for (int q = 0; q < 3; q++) k.u4[q] += c47 * k.u2[q] + c48 * k.u3[q];
	for (int q = 3; q < 5; q++) k.u4[q] += c48 * k.u3[q];
	return k;
}

static int step_ars343(tb200_ctx * ctx, int first, int last, double dt) {
	(void)first;
	(void)last;
	static const ARS343Coefficients k = ars343_coefficients();
	// :161-234
	TRY(substage_from(ctx, copy_of(0), 0, 1, k.diag_exp[0] * dt));
	TRY(tb200_copy_v_step_implicit(ctx, 1, 2, k.diag_imp[0] * dt));

	TRY(substage_from(ctx, k.u2, 2, 3, k.diag_exp[1] * dt));
	TRY(tb200_copy_v_step_implicit(ctx, 3, 4, k.diag_imp[1] * dt));

	TRY(substage_from(ctx, k.u3, 4, 5, k.diag_exp[2] * dt));
	TRY(tb200_copy_v_step_implicit(ctx, 5, 6, k.diag_imp[2] * dt));

	TRY(substage_from(ctx, k.u4, 6, 1, k.diag_exp[3] * dt));

	// (CopyData(1 -> 0) is the first thing StepAfterSubCycle does)
	TRY(tb200_h_step_after_subcycle(ctx, 1, 0, 2, dt));
	return 0;
}

///////////////////////////////////////////////////////////////////////////////
// TimestepSchemeERK: explicit Runge-Kutta on the horizontal dynamics only
// (reference TimestepSchemeERK.cpp:143-304; VerticalDynamics is never called).

static int substage_h(tb200_ctx * ctx, const std::vector<double> & c, int in, int out, double dt) {
	std::vector<double> cc(c);
	if ((int)cc.size() <= out) cc.resize(out + 1, 0.0);
	TRY(tb200_lincomb(ctx, cc.data(), (int)cc.size(), out, ALL));
	TRY(tb200_h_step_explicit(ctx, in, out, dt));
	TRY(tb200_dss(ctx, out, ALL));
	return 0;
}

static int step_erk(tb200_ctx * ctx, int scheme, double dt) {
	const double half = 0.5 * dt;
	if (scheme == TB200_SCHEME_ERK_FE) {
		TRY(substage_h(ctx, copy_of(0), 0, 4, dt));
	} else if (scheme == TB200_SCHEME_ERK_RK4) {
		TRY(substage_h(ctx, copy_of(0), 0, 1, half));
		TRY(substage_h(ctx, copy_of(0), 1, 2, half));
		TRY(substage_h(ctx, copy_of(0), 2, 3, dt));
		const std::vector<double> rk4 = {-1.0 / 3.0, 1.0 / 3.0, 2.0 / 3.0, 1.0 / 3.0, 0.0};
		TRY(substage_h(ctx, rk4, 3, 4, dt / 6.0));
	} else if (scheme == TB200_SCHEME_ERK_SSP3) {
		TRY(substage_h(ctx, copy_of(0), 0, 1, dt));
		const std::vector<double> a = {3.0 / 4.0, 1.0 / 4.0, 0.0};
		TRY(substage_h(ctx, a, 1, 2, 0.25 * dt));
		const std::vector<double> b = {1.0 / 3.0, 0.0, 2.0 / 3.0, 0.0, 0.0};
		TRY(substage_h(ctx, b, 2, 4, (2.0 / 3.0) * dt));
	} else if (scheme == TB200_SCHEME_ERK_KGU35) {
		TRY(substage_h(ctx, copy_of(0), 0, 1, dt / 5.0));
		TRY(substage_h(ctx, copy_of(0), 1, 2, dt / 5.0));
		TRY(substage_h(ctx, copy_of(0), 2, 3, dt / 3.0));
		TRY(substage_h(ctx, copy_of(0), 3, 2, 2.0 * dt / 3.0));
		const std::vector<double> kgu = {-1.0 / 4.0, 5.0 / 4.0, 0.0, 0.0, 0.0};
		TRY(substage_h(ctx, kgu, 2, 4, 3.0 * dt / 4.0));
	} else {
		const double s1 = 0.377268915331368;
		TRY(substage_h(ctx, copy_of(0), 0, 1, s1 * dt));
		TRY(substage_h(ctx, copy_of(1), 1, 2, s1 * dt));
		const std::vector<double> a = {0.355909775063327, 0.0, 0.644090224936674, 0.0};
		TRY(substage_h(ctx, a, 2, 3, 0.242995220537396 * dt));
		const std::vector<double> b = {0.367933791638137, 0.0, 0.0, 0.632066208361863};
		TRY(substage_h(ctx, b, 3, 0, 0.238458932846290 * dt));
		const std::vector<double> c = {0.762406163401431, 0.0, 0.237593836598569, 0.0, 0.0};
		TRY(substage_h(ctx, c, 0, 4, 0.287632146308408 * dt));
	}
	// :300-303 (the CopyData(4 -> 0) repeats the one StepAfterSubCycle starts with)
	TRY(tb200_h_step_after_subcycle(ctx, 4, 0, 2, dt));
	return 0;
}

///////////////////////////////////////////////////////////////////////////////
// ARS(2,2,2), ARS(2,3,2), ARS(4,4,3) of Ascher, Ruuth & Spiteri (1997).  The
// stage combinations are the reference's m_du?fCombo vectors
// (TimestepSchemeARS222.cpp:67-72, ARS232.cpp:68-86, ARS443.cpp:70-108).

// raw[r][.]: combination that forms the base of explicit stage r + 1 from
// u^n (instance 0) and the explicit / implicit results of the earlier stages
static void ars_combo(
	const double * E, const double * I, int ns, int r, std::vector<double> & c, int len
) {
	c.assign(len, 0.0);
	c[0] = 1.0 - E[r * ns + 0] / E[0];
	c[1] = E[r * ns + 0] / E[0] - I[r * ns + 0] / I[0];
	c[2] = I[r * ns + 0] / I[0];
	for (int j = 1; j < r; j++) {
		c[2 * j + 1] = E[r * ns + j] / E[j * ns + j] - I[r * ns + j] / I[j * ns + j];
		c[2 * j + 2] = I[r * ns + j] / I[j * ns + j];
	}
}

// CopyData(State only) + StepImplicit + DSS, as ARS222 / ARS443 issue them
// (the reference copies the State twice and never the Tracers, ARS222.cpp:88-89)
static int implicit_stage_dss(tb200_ctx * ctx, int src, int dst, double dt) {
	if (ctx->lay.ntr == 0) {
		TRY(tb200_copy_v_step_implicit(ctx, src, dst, dt));
	} else {
		TRY(tb200_copy(ctx, src, dst, TB200_DATA_STATE));
		TRY(tb200_v_step_implicit(ctx, dst, dst, dt));
	}
	TRY(tb200_dss(ctx, dst, ALL));
	return 0;
}

static int step_ars222(tb200_ctx * ctx, double dt) {
	const double gam = 1.0 - 0.5 * std::sqrt(2.0);
	const double del = 1.0 - 1.0 / (2.0 * gam);
	const double I[4] = {gam, 0.0, 1.0 - gam, gam};
	const double E[4] = {gam, 0.0, del, 1.0 - del};
	std::vector<double> u2;
	ars_combo(E, I, 2, 1, u2, 4);
	// stage 1 (:75-93)
	TRY(substage_from(ctx, copy_of(0), 0, 1, E[0] * dt));
	TRY(implicit_stage_dss(ctx, 1, 2, I[0] * dt));
	// stage 2 (:96-110)
	TRY(substage_from(ctx, u2, 2, 3, E[3] * dt));
	TRY(tb200_v_step_implicit(ctx, 3, 3, I[3] * dt));
	TRY(tb200_dss(ctx, 3, ALL));
	// hyperdiffusion (:113-117)
	TRY(tb200_copy(ctx, 3, 2, ALL));
	TRY(tb200_h_step_after_subcycle(ctx, 2, 1, 3, dt));
	TRY(tb200_copy(ctx, 1, 0, ALL));
	return 0;
}

static int step_ars232(tb200_ctx * ctx, double dt) {
	const double gam = 1.0 - 1.0 / std::sqrt(2.0);
	const double del = -(2.0 * std::sqrt(2.0)) / 3.0;
	const double I[9] = {gam, 0., 0., 1.0 - gam, gam, 0., 1.0 - gam, gam, 0.};
	const double E[9] = {gam, 0., 0., del, 1.0 - del, 0., 0., 1.0 - gam, gam};
	std::vector<double> u2, u3;
	ars_combo(E, I, 3, 1, u2, 6);
	ars_combo(E, I, 3, 2, u3, 7);
	u3[5] = -E[2 * 3 + 1] / E[1 * 3 + 1];     // :85
	// stage 1 (:89-105)
	TRY(substage_from(ctx, copy_of(0), 0, 1, E[0] * dt));
	TRY(tb200_copy_v_step_implicit(ctx, 1, 2, I[0] * dt));
	// stage 2 (:108-126): the combination is kept in instance 5
	TRY(tb200_lincomb(ctx, u2.data(), (int)u2.size(), 5, ALL));
	TRY(substage_from(ctx, copy_of(5), 2, 3, E[4] * dt));
	TRY(tb200_copy_v_step_implicit(ctx, 3, 4, I[4] * dt));
	// stage 3 (:129-137)
	TRY(substage_from(ctx, u3, 4, 6, E[8] * dt));
	// hyperdiffusion (:146-150)
	TRY(tb200_copy(ctx, 6, 2, ALL));
	TRY(tb200_h_step_after_subcycle(ctx, 2, 1, 6, dt));
	TRY(tb200_copy(ctx, 1, 0, ALL));
	return 0;
}

static int step_ars443(tb200_ctx * ctx, double dt) {
	const double I[16] = {
		1. / 2., 0., 0., 0.,
		1. / 6., 1. / 2., 0., 0.,
		-1. / 2., 1. / 2., 1. / 2., 0.,
		3. / 2., -3. / 2., 1. / 2., 1. / 2.};
	const double E[16] = {
		1. / 2., 0., 0., 0.,
		11. / 18., 1. / 18., 0., 0.,
		5. / 6., -5. / 6., 1. / 2., 0.,
		1. / 4., 7. / 4., 3. / 4., -7. / 4.};
	std::vector<double> u2, u3, u4;
	ars_combo(E, I, 4, 1, u2, 8);
	ars_combo(E, I, 4, 2, u3, 9);
	ars_combo(E, I, 4, 3, u4, 10);
	u3[7] = -E[2 * 4 + 1] / E[1 * 4 + 1];     // :88
	u4[7] = -E[3 * 4 + 1] / E[1 * 4 + 1];     // :104
	u4[8] = -E[3 * 4 + 2] / E[2 * 4 + 2];     // :105
	// stage 1 (:111-127)
	TRY(substage_from(ctx, copy_of(0), 0, 1, E[0] * dt));
	TRY(implicit_stage_dss(ctx, 1, 2, I[0] * dt));
	// stage 2 (:130-148): combination kept in instance 7
	TRY(tb200_lincomb(ctx, u2.data(), (int)u2.size(), 7, ALL));
	TRY(substage_from(ctx, copy_of(7), 2, 3, E[5] * dt));
	TRY(implicit_stage_dss(ctx, 3, 4, I[5] * dt));
	// stage 3 (:151-169): combination kept in instance 8
	TRY(tb200_lincomb(ctx, u3.data(), (int)u3.size(), 8, ALL));
	TRY(substage_from(ctx, copy_of(8), 4, 5, E[10] * dt));
	TRY(implicit_stage_dss(ctx, 5, 6, I[10] * dt));
	// stage 4 (:172-185)
	TRY(substage_from(ctx, u4, 6, 9, E[15] * dt));
	TRY(tb200_v_step_implicit(ctx, 9, 9, I[15] * dt));
	TRY(tb200_dss(ctx, 9, ALL));
	// hyperdiffusion (:187-191)
	TRY(tb200_copy(ctx, 9, 2, ALL));
	TRY(tb200_h_step_after_subcycle(ctx, 2, 1, 9, dt));
	TRY(tb200_copy(ctx, 1, 0, ALL));
	return 0;
}

///////////////////////////////////////////////////////////////////////////////
// TimestepSchemeGARK2 (second-order IMEX GARK, Sandu & Guenther 2013, example 7;
// reference TimestepSchemeGARK2.cpp:25-141).  The reference copies the State twice
// and never the Tracers before its first implicit stage (:107-108); so does this.

static int step_gark2(tb200_ctx * ctx, double dt) {
	const double gam = 1.0 - 0.5 * std::sqrt(2.0);
	const double alpha = 0.5;
	const double Imp[2][2] = {{gam, 0.0}, {1.0 - gam, gam}};
	const double Exp[2][2] = {{0.0, 0.0}, {1.0, 0.0}};
	const double EI[2][2] = {{0.0, 0.0}, {1.0, 0.0}};
	const double IE[2][2] = {{gam, 0.0}, {alpha, 1.0 - alpha}};
	std::vector<double> u2f(4, 0.0), u3f(5, 0.0);
	u2f[0] = 1.0 - Exp[1][0] / IE[0][0];
	u2f[1] = Exp[1][0] / IE[0][0] - EI[1][0] / Imp[0][0];
	u2f[2] = EI[1][0] / Imp[0][0];
	u3f[0] = 1.0 - IE[1][0] / IE[0][0];
	u3f[1] = IE[1][0] / IE[0][0] - Imp[1][0] / Imp[0][0];
	u3f[2] = Imp[1][0] / Imp[0][0];
	// stage 1 (:96-112)
	TRY(substage_from(ctx, copy_of(0), 0, 1, IE[0][0] * dt));
	TRY(tb200_copy(ctx, 1, 2, TB200_DATA_STATE));
	TRY(tb200_v_step_implicit(ctx, 2, 2, Imp[0][0] * dt));
	TRY(tb200_dss(ctx, 2, ALL));
	// (:115-116)
	TRY(lincomb(ctx, u2f, 3));
	// stage 2 (:120-134)
	TRY(substage_from(ctx, u3f, 3, 4, IE[1][1] * dt));
	TRY(tb200_v_step_implicit(ctx, 4, 4, Imp[1][1] * dt));
	TRY(tb200_dss(ctx, 4, ALL));
	// hyperdiffusion (:137-141)
	TRY(tb200_copy(ctx, 4, 2, ALL));
	TRY(tb200_h_step_after_subcycle(ctx, 2, 1, 4, dt));
	TRY(tb200_copy(ctx, 1, 0, ALL));
	return 0;
}

///////////////////////////////////////////////////////////////////////////////
// TimestepSchemeSSP3332 (Pareschi & Russo SSP3(3,3,2); reference
// TimestepSchemeSSP3332.cpp:25-181).  Every implicit stage copies the State only.

static int step_ssp3332(tb200_ctx * ctx, double dt) {
	const double gam = 1.0 - 1.0 / sqrt(2.0);
	const double Imp[4][4] = {
		{gam, 0., 0., 0.},
		{(1.0 - 2.0 * gam), gam, 0., 0.},
		{0.5 - gam, 0.0, gam, 0.},
		{1.0 / 6.0, 1.0 / 6.0, 2.0 / 3.0, 0.}};
	const double Exp[4][4] = {
		{0.0, 0., 0., 0.},
		{1.0, 0., 0., 0.},
		{0.25, 0.25, 0., 0.},
		{1.0 / 6.0, 1.0 / 6.0, 2.0 / 3.0, 0.}};
	std::vector<double> u2f(8, 0.0), u3f(9, 0.0), u4f(9, 0.0);
	// (:62-105, the same expressions in the same order)
	u2f[0] = 1.0 - Imp[1][0] / Imp[0][0];
	u2f[2] = Imp[1][0] / Imp[0][0];
	u3f[0] = 1.0 + Imp[1][0] / Imp[0][0] *
	               Exp[2][0] / Exp[1][0] -
	               Exp[2][0] / Exp[1][0] -
	               Imp[2][0] / Imp[0][0];
	u3f[2] = Imp[2][0] / Imp[0][0] -
	         Imp[1][0] / Imp[0][0] *
	         Exp[2][0] / Exp[1][0];
	u3f[3] = Exp[2][0] / Exp[1][0];
	u4f[0] = 1.0 - Exp[3][0] / Imp[0][0];
	u4f[2] = Imp[3][0] / Imp[0][0];
	u4f[3] = Exp[3][0] / Exp[1][0] -
	         Imp[3][1] / Imp[1][1];
	u4f[4] = Imp[3][1] / Imp[1][1];
	u4f[5] = Exp[3][1] / Exp[2][1] -
	         Imp[3][2] / Imp[2][2];
	u4f[6] = Imp[3][2] / Imp[2][2];
	u4f[7] = -Exp[3][0] / Exp[1][0];
	u4f[8] = -Exp[3][1] / Exp[2][1];
	// stage 1: implicit only (:109-115)
	TRY(tb200_copy(ctx, 0, 2, TB200_DATA_STATE));
	TRY(tb200_v_step_implicit(ctx, 2, 2, Imp[0][0] * dt));
	TRY(tb200_dss(ctx, 2, ALL));
	// stage 2 (:118-137): the combination is kept in instance 7
	TRY(lincomb(ctx, u2f, 7));
	TRY(substage_from(ctx, copy_of(7), 2, 3, Exp[1][0] * dt));
	TRY(tb200_copy(ctx, 3, 4, TB200_DATA_STATE));
	TRY(tb200_v_step_implicit(ctx, 4, 4, Imp[1][1] * dt));
	TRY(tb200_dss(ctx, 4, ALL));
	// stage 3 (:140-159): combination kept in instance 8
	TRY(lincomb(ctx, u3f, 8));
	TRY(substage_from(ctx, copy_of(8), 4, 5, Exp[2][1] * dt));
	TRY(tb200_copy(ctx, 5, 6, TB200_DATA_STATE));
	TRY(tb200_v_step_implicit(ctx, 6, 6, Imp[2][2] * dt));
	TRY(tb200_dss(ctx, 6, ALL));
	// final stage (:162-170)
	TRY(substage_from(ctx, u4f, 6, 1, Exp[3][2] * dt));
	// hyperdiffusion (:173-177)
	TRY(tb200_copy(ctx, 1, 2, ALL));
	TRY(tb200_h_step_after_subcycle(ctx, 2, 1, 3, dt));
	TRY(tb200_copy(ctx, 1, 0, ALL));
	return 0;
}

///////////////////////////////////////////////////////////////////////////////
// TimestepSchemeARK232 (reference TimestepSchemeARK232.cpp:25-226): the first
// explicit stage in two sub-cycles that overwrite instance 0 on the way (:184-189),
// the implicit terms of the first stage evaluated explicitly
// (VerticalDynamicsFEM::StepImplicitTermsExplicitly, VerticalDynamicsFEM.cpp:439-612).

static int step_ark232(tb200_ctx * ctx, double dt) {
	const double gam = 1.0 - 1.0 / std::sqrt(2.0);
	const double del = 1.0 / (2.0 * std::sqrt(2.0));
	const double alpha = 1.0 / 6.0 *
	                     (3.0 + 2.0 * std::sqrt(2.0));
	const double Imp[3][3] = {
		{gam, gam, 0.},
		{del, del, gam},
		{del, del, gam}};
	const double Exp[3][3] = {
		{2.0 * gam, 0., 0.},
		{1.0 - alpha, alpha, 0.},
		{del, del, gam}};
	std::vector<double> u2f(7, 0.0), u3f(8, 0.0);
	u2f[0] = 1.0 - Exp[1][0] / Exp[0][0];
	u2f[1] = Exp[1][0] / Exp[0][0] -
	         Imp[1][0] / Imp[0][0];
	u2f[2] = Imp[1][0] / Imp[0][0] -
	         Imp[1][1] / Imp[0][1];
	u2f[3] = Imp[1][1] / Imp[0][1];
	u3f[0] = 1.0 - Exp[2][0] / Exp[0][0];
	u3f[1] = Exp[2][0] / Exp[0][0] -
	         Imp[2][0] / Imp[0][0];
	u3f[2] = Imp[2][0] / Imp[0][0] -
	         Imp[2][1] / Imp[0][1];
	u3f[3] = Imp[2][1] / Imp[0][1];
	u3f[4] = Exp[2][1] / Exp[1][1] -
	         Imp[2][2] / Imp[1][2];
	u3f[5] = Imp[2][2] / Imp[1][2];
	u3f[6] = -Exp[2][1] / Exp[1][1];
	// SubcycleStageExplicit(ExpCf[0][0], iNS = 2, 0 -> 1) (:170-192)
	{
		const int iNS = 2;
		for (int n = 0; n < iNS; n++) {
			TRY(substage_from(ctx, copy_of(0), 0, 1, Exp[0][0] * dt / iNS));
			if (n < iNS - 1) {
				TRY(tb200_copy(ctx, 1, 0, ALL));
			}
		}
	}
	// SubcycleStageImplicitExplicitly(ImpCf[0][0], iNS = 1, 1 -> 2) (:196-226)
	TRY(tb200_copy(ctx, 1, 2, ALL));
	TRY(tb200_v_step_implicit_terms_explicitly(ctx, 1, 2, Imp[0][0] * dt / 1));
	TRY(tb200_dss(ctx, 2, ALL));
	// (:121-127)
	TRY(tb200_copy(ctx, 2, 3, TB200_DATA_STATE));
	TRY(tb200_v_step_implicit(ctx, 3, 3, Imp[0][1] * dt));
	TRY(tb200_dss(ctx, 3, ALL));
	// stage 2 (:130-149): combination kept in instance 6
	TRY(lincomb(ctx, u2f, 6));
	TRY(substage_from(ctx, copy_of(6), 3, 4, Exp[1][1] * dt));
	TRY(tb200_copy(ctx, 4, 5, TB200_DATA_STATE));
	TRY(tb200_v_step_implicit(ctx, 5, 5, Imp[1][2] * dt));
	TRY(tb200_dss(ctx, 5, ALL));
	// stage 3 (:152-160)
	TRY(substage_from(ctx, u3f, 5, 7, Exp[2][2] * dt));
	// hyperdiffusion (:163-167)
	TRY(tb200_copy(ctx, 7, 2, ALL));
	TRY(tb200_h_step_after_subcycle(ctx, 7, 1, 3, dt));
	TRY(tb200_copy(ctx, 1, 0, ALL));
	return 0;
}

int tb_mirror_errors(tb200_ctx * ctx);    // tb200_api.cu
int tb_poll_errors(tb200_ctx * ctx);

static int step_dispatch(tb200_ctx * ctx, int scheme, int first, int last, double dt) {
	switch (scheme) {
		case TB200_SCHEME_ARS343: return step_ars343(ctx, first, last, dt);
		case TB200_SCHEME_ARS222: return step_ars222(ctx, dt);
		case TB200_SCHEME_ARS232: return step_ars232(ctx, dt);
		case TB200_SCHEME_ARS443: return step_ars443(ctx, dt);
		case TB200_SCHEME_GARK2: return step_gark2(ctx, dt);
		case TB200_SCHEME_SSP3332: return step_ssp3332(ctx, dt);
		case TB200_SCHEME_ARK232: return step_ark232(ctx, dt);
		case TB200_SCHEME_ERK_KGU35:
		case TB200_SCHEME_ERK_FE:
		case TB200_SCHEME_ERK_RK4:
		case TB200_SCHEME_ERK_SSP3:
		case TB200_SCHEME_ERK_SSPRK53:
			return step_erk(ctx, scheme, dt);
		default:
			return step_strang(ctx, scheme, first, last, dt);
	}
}

extern "C" int tb200_step(tb200_ctx * ctx, int scheme, int first, int last, double dt) {
	const int need = tb200_scheme_instances(scheme);
	if (need < 0) TB_FAIL(ctx, "time scheme not implemented");
	if ((int)ctx->inst.size() < need) TB_FAIL(ctx, "not enough state instances for this scheme");
	// a failure recorded by an earlier step (column solve, peer time-out): stop
	TRY(tb_poll_errors(ctx));
	TRY(step_dispatch(ctx, scheme, first, last, dt));
	return tb_mirror_errors(ctx);
}

///////////////////////////////////////////////////////////////////////////////
// Grid::Checksum, ChecksumType_Sum (reference GridPatch.cpp:744-835,
// Grid.cpp:460-524): area-weighted sum of every component.

__global__ void k_checksum(
	DevLayout lay, const double * data, const double * area_node,
	const double * area_redge, double * sums
) {
	__shared__ double red[256];
	const int c = blockIdx.y;
	const int nn = lay.nn;
	const int nl = lay.rowlev[c];
	const double * area = lay.onedge[c] ? area_redge : area_node;
	const long long per_e = (long long)nl * nn;
	const long long total = lay.nelem * per_e;
	double acc = 0.0;
	for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	     idx < total; idx += (long long)gridDim.x * blockDim.x
	) {
		const long long e = idx / per_e;
		const long long r = idx % per_e;
		acc += data[((size_t)e * lay.nrows + lay.rowoff[c]) * nn + r] * area[(size_t)e * per_e + r];
	}
	red[threadIdx.x] = acc;
	__syncthreads();
	for (int s = blockDim.x / 2; s > 0; s >>= 1) {
		if ((int)threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
		__syncthreads();
	}
	if (threadIdx.x == 0) atomicAdd(&sums[c], red[0]);
}

extern "C" int tb200_checksum(tb200_ctx * ctx, int inst, double * sums) {
	if (ctx->d_area_node == 0) TB_FAIL(ctx, "element areas not uploaded");
	if (inst < 0 || inst >= (int)ctx->inst.size()) TB_FAIL(ctx, "invalid state instance");
	TB_CHECK(ctx, cudaMemsetAsync(ctx->d_sums, 0, 64 * sizeof(double), ctx->stream));
	auto kfn = k_checksum;
	TB_LAUNCH(kfn, dim3(148, ctx->lay.ncomp), dim3(256), 0, ctx->stream,
		ctx->lay, (const double *)ctx->inst[inst], (const double *)ctx->d_area_node,
		(const double *)ctx->d_area_redge, ctx->d_sums);
	ctx->launches++;
	ctx->writes++;
	TB_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
	TB_CHECK(ctx, cudaMemcpy(sums, ctx->d_sums, ctx->lay.ncomp * sizeof(double),
		cudaMemcpyDeviceToHost));
	return 0;
}
