// Device-side parameter blocks shared by the host code and the kernels.
//
// DEVICE DATA LAYOUT (private to the library)
// -------------------------------------------
// All local patches are flattened into one list of spectral elements.  Every
// array is element-major with the np*np nodes of an element fastest:
//
//     value(e, row, n) = base[(e * nrows + row) * NN + n],   n = i*np + j
//
// "row" enumerates (component, level): component c owns rows
// [rowoff[c], rowoff[c]+rowlev[c]) with rowlev = L on model levels and L+1 on
// interfaces; tracer q owns rows [troff + q*L, troff + (q+1)*L).  One row of one
// element is NN*8 = 128 bytes at np = 4, so an element's whole state is one
// contiguous, 128-byte aligned block that a thread block streams with fully
// coalesced loads, and a column (fixed e,n) walks rows with stride NN.
// Only the valid slots of the reference's state are stored (U,V,rho-theta,rho
// on levels, W on interfaces under Lorenz staggering, Grid.cpp:281-287); the
// reference's derived copies (W on levels, U,V on interfaces) are recomputed
// where they are needed.  There is no halo: duplicates of a node in other
// elements/patches are reached through the averaging groups (tb200_dss.cuh).
#ifndef TB200_DEVICE_H
#define TB200_DEVICE_H

#define TB_MAXC 8
#define TB_MAXNP 8
#define TB_NOPS 11

struct DevLayout {
	int np, nn;
	int nlev;
	int ncomp, ntr;
	int nrows_state;      // rows per element holding state components
	int nrows;            // rows per element (state + tracers)
	int rowoff[TB_MAXC];
	int rowlev[TB_MAXC];
	int onedge[TB_MAXC];
	int troff;
	long long nelem;
};

struct DevOp {
	const double * coeff; // [nout][width], entry (k, l - begin[k])
	const int * begin;    // [nout]
	const int * end;      // [nout]
	int width, nout, nin;
};

struct DevOps {
	DevOp op[TB_NOPS];
};

struct DevGeom {
	// per element [e]
	const double * inv_da;
	const double * inv_db;
	const double * nu_scale;    // (deltaA / reference length)^3.2
	// 2-D metric [e][NN]
	const double * j2d;
	const double * a0; const double * a1;   // ContraMetric2DA
	const double * b0; const double * b1;   // ContraMetric2DB
	const double * f;                       // Coriolis
	const double * zs;                      // topography
	// 3-D metric on levels [e][L][NN]
	const double * jac;
	const double * ca[3]; const double * cb[3]; const double * cx[3];
	const double * dr[3];
	// 3-D metric on interfaces [e][L+1][NN]
	const double * jace;
	const double * cae[3]; const double * cbe[3]; const double * cxe[3];
	const double * dre[3];
	// Terrain-following cubed-sphere metric evaluated on the fly (optional):
	// gnomonic X, Y and the topography derivatives per column [e][NN], the
	// eta levels / interfaces, model top and radius
	// (GridPatchCSGLL::EvaluateGeometricTerms, GridPatchCSGLL.cpp:344-553)
	int analytic;
	const double * tx; const double * ty;
	const double * tda; const double * tdb;
	const double * reta_n; const double * reta_e;
	double ztop, radius;
};

// Column constants and per-level values of the terrain-following metric,
// evaluated with the reference's expression order and without FMA contraction
// so that they equal the stored arrays bit for bit.
struct ColMetric {
	double a0, a1, b0, b1, j2d;
	double onepx2, onepy2, xy;
	double msod;        // (-scale) / dxr
	double inv_dxr, inv_dxr2, dxr;
	double dazs, dbzs;
};

struct LevMetric {
	double jac;
	double a2, b2, x2;  // ContraMetricA[2], ContraMetricB[2], ContraMetricXi[2]
	double dar, dbr;    // DerivR[0], DerivR[1]
};

struct DevTables {
	double dx[TB_MAXNP * TB_MAXNP];  // dx[s*np+i] = dDxBasis1D(s,i)
	double st[TB_MAXNP * TB_MAXNP];  // st[i*np+s] = dStiffness1D(i,s)
};

struct DevPhys {
	double g, R, cp, cv, p0;
	double exner_c1;   // R / (cp - R)
	double exner_c2;   // R / p0
	int exner25;       // exponent R / cv is 2/5 (dry air of PhysicalConstants.h): tb_exner
};

#endif
