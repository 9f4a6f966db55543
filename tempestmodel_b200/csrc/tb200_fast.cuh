// Fast path of the explicit stage (FP64, sm_100a): vertical order 1,
// terrain-following metric with a level-independent layer depth.
//
// Same mathematics as k_nh_explicit (tb200_kernels.cuh), reorganised around the
// B200's two scarce resources for this stencil - FP64 issue slots and HBM
// bytes:
//
//  * one block per element, one thread per (level k, element row i); the
//    thread owns the four nodes (i, j = 0..3) of its row, i.e. 32 contiguous
//    bytes of every 128-byte state row -> 16-byte vector loads / stores, whole
//    element streamed as one contiguous 19 kB block;
//  * beta-derivatives (sums over j) are register-only; alpha-derivatives (sums
//    over i) read the (k) tile of a field from shared memory as 16-byte
//    vectors - five tiles + the V column, one __syncwarp, no block barrier;
//  * the vertical operators (interpolation / differentiation / penalty, taken
//    from the host tables, never hard-coded) are pre-windowed on the host to
//    dense 3-coefficient rows around k and loaded once per thread;
//  * the 3-D metric arrays of the reference (13 doubles per node) are replaced
//    by 15 constants per column (TBF_*), exact for every terrain-following
//    metric of the form  dR/dalpha = s(eta) dzs/dalpha, dR/dxi = ztop - zs
//    (GridPatchCSGLL.cpp:344-553; GridPatchCartesianGLL.cpp:262-330 with flat
//    terrain).  The host verifies them against the uploaded reference arrays
//    before enabling this path (tb200_api.cu: fast_prepare).
//  * Grid::CopyData / LinearCombineData of the stage base is formed on the fly.
//
// Summation order inside every np-sum and column operator follows the
// reference; products are contracted to FMA by nvcc.  Results agree with the
// reference to rounding (tests/test_parity.py, 1e-12 of the tendency).
#ifndef TB200_FAST_CUH
#define TB200_FAST_CUH

#include "tb200_platform.h"
#include "tb200_device.h"
#include "tb200_kernels.cuh"
#include "tb200_tma.cuh"

// ---- per-column constants: colc[(e * TBF_NC + q) * 16 + n] ---------------------
#define TBF_A0 0       // ContraMetric2DA[0]
#define TBF_A1 1       // ContraMetric2DA[1] (= ContraMetric2DB[0])
#define TBF_B1 2       // ContraMetric2DB[1]
#define TBF_JAC 3      // Jacobian = dxr * Jacobian2D (levels and interfaces)
#define TBF_INVJAC 4
#define TBF_FJ 5       // CoriolisF * Jacobian2D
#define TBF_A2 6       // ContraMetricA[2] = s * A2, A2 = -(a0 dazs + a1 dbzs) / dxr
#define TBF_B2 7       // ContraMetricB[2] = s * B2
#define TBF_X0 8       // ContraMetricXi[2] = X0 + s^2 X2, X0 = 1 / dxr^2
#define TBF_X2 9       //                     X2 = -(A2 dazs + B2 dbzs) / dxr
#define TBF_GDA 10     // g * dzs/dalpha   (g * DerivR[0] = s * GDA)
#define TBF_GDB 11
#define TBF_DXR 12     // DerivR[2]
#define TBF_J2D 13     // Jacobian2D
#define TBF_B0 14      // ContraMetric2DB[0]
#define TBF_NC 15

// ---- per-level operator windows: lev[k * TBF_LW + q], k = 0..L ------------------
#define TBF_CW 0       // [2] InterpREdgeToNode row k on W[k], W[k+1]
#define TBF_CD 2       // [3] DiffNodeToNode row k on levels k-1, k, k+1
#define TBF_CILO 5     // [3] InterpNodeToREdge row k   on levels k-1, k, k+1
#define TBF_CIHI 8     // [3] InterpNodeToREdge row k+1 on levels k-1, k, k+1
#define TBF_CPL 11     // [3] penalty (left)  row k on levels k-1, k, k+1
#define TBF_CPR 14     // [3] penalty (right) row k
#define TBF_SN 17      // s on level k
#define TBF_SE 18      // s on interface k
#define TBF_SE1 19     // s on interface k+1
#define TBF_CB0 20     // [3] InterpNodeToREdge row 0 on levels 0, 1, 2 (row k = 0 only)
// column solve (tb200_column_fast.cuh)
#define TBF_DNE 23     // [2] DiffNodeToREdge row k on levels k-1, k
#define TBF_DEN 25     // [2] DiffREdgeToNode row k on interfaces k, k+1
#define TBF_DDE 27     // [3] DiffDiffREdgeToREdge row k on interfaces k-1, k, k+1
#define TBF_IEN1 30    // [2] InterpREdgeToNode row k-1 on W[k-1], W[k]
#define TBF_LW 32
#define TBF_LWK 24     // entries of a row the explicit stage reads (0 .. TBF_CB0 + 2)
#define TBF_LWS 26     // row stride of those in shared memory (bank spread: 8 rows, 8 bank pairs)

#define TBF_RS 18      // shared-memory row stride (doubles): 16 nodes + 2 pad

// Exner pressure cp (R rho-theta / p0)^(R/cv) (PhysicalConstants.h:397-399).  The
// reference evaluates cp * exp((R/cv) * log(x)): two libm calls, ~120 FP64-pipe
// instructions on the device.  With the reference's constants R/cv = 287/717.5 =
// 2/5 exactly, so x^(2/5) is computed as x r^3 with r = x^(-1/5) from one
// division-free Newton step on a single-precision seed, followed by one
// correction of y on y^5 = x^2 whose residual is formed with an FMA: ~25 FP64
// instructions, <= 1 ulp (exp(c log x) itself is off by up to 2.4 ulp).  Any
// other exponent takes the libm form.
__device__ __forceinline__ double tb_pow25(double x) {
#ifdef TB200_EMU
	double r = (double)powf((float)x, -0.2f);
#else
	double r = (double)__powf((float)x, -0.2f);
#endif
	double r2 = r * r, r4 = r2 * r2, r5 = r4 * r;
	r = r * fma(-x, r5, 6.0) * 0.2;
	r2 = r * r;
	const double y = x * r2 * r;
	r4 = r2 * r2;
	const double r8 = r4 * r4;
	const double s = r8 * r2;                     // ~ 1 / x^2
	const double y2 = y * y, y4 = y2 * y2, y5 = y4 * y;
	const double d = fma(x, x, -y5);
	return fma(y, d * s * 0.2, y);
}

__device__ __forceinline__ double tb_exner(const DevPhys & ph, double rhotheta) {
	const double x = ph.exner_c2 * rhotheta;
	if (ph.exner25 && x > 1.0e-6 && x < 1.0e6) {
		return ph.cp * tb_pow25(x);
	}
	return ph.cp * exp(ph.exner_c1 * log(x));
}

// Elements a launch works on: all of them (list == 0, n = nelem) or an index
// list (multi-rank: the elements that feed the halo exchange first, the rest on
// a second stream while the exchange is in flight).
struct ElemList {
	const int * list;
	int n;
};

__device__ __forceinline__ long long tb_elem(const ElemList & el, long long w) {
	return (el.list != 0) ? (long long)el.list[w] : w;
}

struct FastArgs {
	const double * colc;
	const double * lev;
	const double * inv_da;
	const double * inv_db;
	double dt;
	int xz;
};

// 4 consecutive doubles <-> registers (two 16-byte accesses)
__device__ __forceinline__ void tb_ld4(const double * p, double (&v)[4]) {
#ifdef TB200_EMU
	v[0] = p[0]; v[1] = p[1]; v[2] = p[2]; v[3] = p[3];
#else
	const double2 a = *reinterpret_cast<const double2 *>(p);
	const double2 b = *reinterpret_cast<const double2 *>(p + 2);
	v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
#endif
}

__device__ __forceinline__ void tb_st4(double * p, const double (&v)[4]) {
#ifdef TB200_EMU
	p[0] = v[0]; p[1] = v[1]; p[2] = v[2]; p[3] = v[3];
#else
	*reinterpret_cast<double2 *>(p) = make_double2(v[0], v[1]);
	*reinterpret_cast<double2 *>(p + 2) = make_double2(v[2], v[3]);
#endif
}

__device__ __forceinline__ void tb_ld2(const double * p, double (&v)[2]) {
#ifdef TB200_EMU
	v[0] = p[0]; v[1] = p[1];
#else
	const double2 a = *reinterpret_cast<const double2 *>(p);
	v[0] = a.x; v[1] = a.y;
#endif
}

__device__ __forceinline__ void tb_st2(double * p, const double (&v)[2]) {
#ifdef TB200_EMU
	p[0] = v[0]; p[1] = v[1];
#else
	*reinterpret_cast<double2 *>(p) = make_double2(v[0], v[1]);
#endif
}

// Stage base of 2 consecutive values (same operation order as k_lincomb)
__device__ __forceinline__ void tb_stage_base2(
	const StageBase & sb, const double * out, size_t off, double (&v)[2]
) {
	if (sb.use_out) {
		tb_ld2(out + off, v);
		return;
	}
	if (sb.scale_dst) {
		tb_ld2(out + off, v);
		v[0] = v[0] * sb.cdst;
		v[1] = v[1] * sb.cdst;
	} else {
		v[0] = 0.0;
		v[1] = 0.0;
	}
	for (int m = 0; m < sb.nsrc; m++) {
		double s[2];
		tb_ld2(sb.src[m] + off, s);
		const double c = sb.coeff[m];
		v[0] += s[0] * c;
		v[1] += s[1] * c;
	}
}

// Stage base of 4 consecutive values (same operation order as k_lincomb)
__device__ __forceinline__ void tb_stage_base4(
	const StageBase & sb, const double * out, size_t off, double (&v)[4]
) {
	if (sb.use_out) {
		tb_ld4(out + off, v);
		return;
	}
	if (sb.scale_dst) {
		tb_ld4(out + off, v);
#pragma unroll
		for (int j = 0; j < 4; j++) v[j] = v[j] * sb.cdst;
	} else {
#pragma unroll
		for (int j = 0; j < 4; j++) v[j] = 0.0;
	}
	for (int m = 0; m < sb.nsrc; m++) {
		double s[4];
		tb_ld4(sb.src[m] + off, s);
		const double c = sb.coeff[m];
#pragma unroll
		for (int j = 0; j < 4; j++) v[j] += s[j] * c;
	}
}

__host__ __device__ inline size_t tb_fast_stage_smem_doubles(int L, bool do_h) {
	// sU, sV [L], sW [L+1]; tiles Wn, KE, EX, FaR, FaP, ZX, AU, AV [L]; Un, Vn [3]
	const size_t rows = do_h ? (size_t)(3 * L + 1 + 8 * L + 6) : (size_t)(3 * L + 1);
	return rows * TBF_RS;
}

// alpha-derivative of a field whose level-k tile sits in shared memory, for
// the node pair j = 2 jh, 2 jh + 1:   out[q] = sum_s tile[s][2 jh + q] * c[s]
__device__ __forceinline__ void tb_cross_sum2(
	const double * tile, const double (&c)[4], double (&o)[2]
) {
	double r0[2], r1[2], r2[2], r3[2];
	tb_ld2(tile, r0);
	tb_ld2(tile + 4, r1);
	tb_ld2(tile + 8, r2);
	tb_ld2(tile + 12, r3);
#pragma unroll
	for (int q = 0; q < 2; q++) {
		double a = 0.0;
		a += r0[q] * c[0];
		a += r1[q] * c[1];
		a += r2[q] * c[2];
		a += r3[q] * c[3];
		o[q] = a;
	}
}

#define TBF_THREADS 128
#ifndef TBF_MINBLOCKS
#define TBF_MINBLOCKS 3
#endif
#define TBF_KB (TBF_THREADS / 4)     // levels per pass

template <bool DO_H, bool DO_V>
__global__ void __launch_bounds__(TBF_THREADS, TBF_MINBLOCKS)
k_nh_stage_fast(
	DevLayout lay, DevTables t, DevPhys ph, FastArgs fa, StageBase sb,
	const double * __restrict__ in, double * out
) {
	const int NP = 4, NN = 16, RS = TBF_RS;
	const int L = lay.nlev;
	const double dt = fa.dt;

	TB_DYN_SMEM(double, sm);
	double * sU = sm;                       // [L][RS]   covariant u_alpha
	double * sV = sU + (size_t)L * RS;      // [L][RS]
	double * sW = sV + (size_t)L * RS;      // [L+1][RS] covariant w on interfaces
	double * tWn = sW + (size_t)(L + 1) * RS;
	double * tKE = tWn + (size_t)L * RS;
	double * tEX = tKE + (size_t)L * RS;
	double * tFaR = tEX + (size_t)L * RS;
	double * tFaP = tFaR + (size_t)L * RS;
	double * tZX = tFaP + (size_t)L * RS;
	double * tAU = tZX + (size_t)L * RS;    // vertical-explicit increments of U, V
	double * tAV = tAU + (size_t)L * RS;
	double * sUn = tAV + (size_t)L * RS;    // [3][RS] U after the horizontal update, levels 0..2
	double * sVn = sUn + 3 * RS;

	const long long e = blockIdx.x;
	const int kq = threadIdx.x >> 2;
	const int i = threadIdx.x & 3;

	const size_t ebase = (size_t)e * lay.nrows * NN;
	const size_t offU = ebase + (size_t)lay.rowoff[0] * NN;
	const size_t offV = ebase + (size_t)lay.rowoff[1] * NN;
	const size_t offP = ebase + (size_t)lay.rowoff[2] * NN;
	const size_t offW = ebase + (size_t)lay.rowoff[3] * NN;
	const size_t offR = ebase + (size_t)lay.rowoff[4] * NN;

	// ---- stage the U, V (levels) and W (interfaces) columns of the element ---------
	for (int k = kq; k <= L; k += TBF_KB) {
		double x[4];
		if (k < L) {
			tb_ld4(in + offU + (size_t)k * NN + i * NP, x);
			tb_st4(sU + k * RS + i * NP, x);
			tb_ld4(in + offV + (size_t)k * NN + i * NP, x);
			tb_st4(sV + k * RS + i * NP, x);
		}
		tb_ld4(in + offW + (size_t)k * NN + i * NP, x);
		tb_st4(sW + k * RS + i * NP, x);
	}
	double p[4], r[4];
	if (DO_H) {
		const int k = (kq < L) ? kq : (L - 1);
		tb_ld4(in + offP + (size_t)k * NN + i * NP, p);
		tb_ld4(in + offR + (size_t)k * NN + i * NP, r);
	}

	// constants of my four columns
	const double * cc = fa.colc + (size_t)e * TBF_NC * NN + i * NP;
	const double dInvDA = fa.inv_da[e];
	const double dInvDB = fa.inv_db[e];

	__syncthreads();

	for (int k0 = 0; k0 < L; k0 += TBF_KB) {
		const int k = k0 + kq;
		const bool active = (k < L);
		const int kc = active ? k : (L - 1);
		const size_t o4 = (size_t)kc * NN + i * NP;    // my 4 nodes inside a component
		const double * lv = fa.lev + (size_t)kc * TBF_LW;
		const double sn = lv[TBF_SN];
		double cA2[4], cB2[4], cX0[4], cX2[4];
		tb_ld4(cc + TBF_A2 * NN, cA2);
		tb_ld4(cc + TBF_B2 * NN, cB2);
		tb_ld4(cc + TBF_X0 * NN, cX0);
		tb_ld4(cc + TBF_X2 * NN, cX2);

		const int km = (kc > 0) ? kc - 1 : 0;
		const int kp = (kc < L - 1) ? kc + 1 : L - 1;
		double u[4], v[4], w0[4], um[4], up[4], vm[4], vp[4], wp[4];
		tb_ld4(sU + kc * RS + i * NP, u);
		tb_ld4(sV + kc * RS + i * NP, v);
		tb_ld4(sW + kc * RS + i * NP, w0);
		tb_ld4(sU + km * RS + i * NP, um);
		tb_ld4(sU + kp * RS + i * NP, up);
		tb_ld4(sV + km * RS + i * NP, vm);
		tb_ld4(sV + kp * RS + i * NP, vp);
		tb_ld4(sW + (kc + 1) * RS + i * NP, wp);
		if (DO_H && k0 > 0) {
			tb_ld4(in + offP + o4, p);
			tb_ld4(in + offR + o4, r);
		}

		// ---- vertical explicit part: upwind penalty on U and V ----------------------
		// (VerticalDynamicsFEM.cpp:816-828, 998-1023; LinearColumnOperatorFEM.cpp:1863-1887)
		double addU[4], addV[4];
#pragma unroll
		for (int j = 0; j < 4; j++) { addU[j] = 0.0; addV[j] = 0.0; }
		if (DO_V) {
			const double se0 = lv[TBF_SE], se1 = lv[TBF_SE1];
			const bool hi = (kc <= L - 2);      // interface k+1 is interior
			const bool lo = (kc >= 1);          // interface k is interior
			const double h0 = lv[TBF_CIHI + 0], h1 = lv[TBF_CIHI + 1], h2 = lv[TBF_CIHI + 2];
			const double l0 = lv[TBF_CILO + 0], l1 = lv[TBF_CILO + 1], l2 = lv[TBF_CILO + 2];
			const double pl0 = lv[TBF_CPL + 0], pl1 = lv[TBF_CPL + 1], pl2 = lv[TBF_CPL + 2];
			const double pr0 = lv[TBF_CPR + 0], pr1 = lv[TBF_CPR + 1], pr2 = lv[TBF_CPR + 2];
#pragma unroll
			for (int j = 0; j < 4; j++) {
				if (hi) {
					double ue = 0.0, ve = 0.0;
					ue += h0 * um[j]; ue += h1 * u[j]; ue += h2 * up[j];
					ve += h0 * vm[j]; ve += h1 * v[j]; ve += h2 * vp[j];
					const double c0 = se1 * cA2[j], c1 = se1 * cB2[j];
					const double c2 = cX0[j] + (se1 * se1) * cX2[j];
					const double xd = c0 * ue + c1 * ve + c2 * wp[j];
					const double wgt = dt * fabs(xd);
					double pu = 0.0, pv = 0.0;
					pu += pl0 * um[j]; pu += pl1 * u[j]; pu += pl2 * up[j];
					pv += pl0 * vm[j]; pv += pl1 * v[j]; pv += pl2 * vp[j];
					addU[j] += pu * wgt;
					addV[j] += pv * wgt;
				}
				if (lo) {
					double ue = 0.0, ve = 0.0;
					ue += l0 * um[j]; ue += l1 * u[j]; ue += l2 * up[j];
					ve += l0 * vm[j]; ve += l1 * v[j]; ve += l2 * vp[j];
					const double c0 = se0 * cA2[j], c1 = se0 * cB2[j];
					const double c2 = cX0[j] + (se0 * se0) * cX2[j];
					const double xd = c0 * ue + c1 * ve + c2 * w0[j];
					const double wgt = dt * fabs(xd);
					double pu = 0.0, pv = 0.0;
					pu += pr0 * um[j]; pu += pr1 * u[j]; pu += pr2 * up[j];
					pv += pr0 * vm[j]; pv += pr1 * v[j]; pv += pr2 * vp[j];
					addU[j] += pu * wgt;
					addV[j] += pv * wgt;
				}
			}
		}

		if (!DO_H) {
			// VerticalDynamicsFEM::StepExplicit alone
			if (active) {
				double bu[4], bv[4];
				tb_stage_base4(sb, out, offU + o4, bu);
				tb_stage_base4(sb, out, offV + o4, bv);
#pragma unroll
				for (int j = 0; j < 4; j++) { bu[j] += addU[j]; bv[j] += addV[j]; }
				tb_st4(out + offU + o4, bu);
				tb_st4(out + offV + o4, bv);
				if (!sb.use_out) {
					double b[4];
					tb_stage_base4(sb, out, offP + o4, b); tb_st4(out + offP + o4, b);
					tb_stage_base4(sb, out, offR + o4, b); tb_st4(out + offR + o4, b);
					tb_stage_base4(sb, out, offW + o4, b); tb_st4(out + offW + o4, b);
					if (k == L - 1) {
						const size_t oL = offW + (size_t)L * NN + i * NP;
						tb_stage_base4(sb, out, oL, b); tb_st4(out + oL, b);
					}
				}
			}
			continue;
		}

		// ---- horizontal part (HorizontalDynamicsFEM.cpp:876-1660) -------------------
		double cA0[4], cA1[4], cB1[4], cJ[4];
		tb_ld4(cc + TBF_A0 * NN, cA0);
		tb_ld4(cc + TBF_A1 * NN, cA1);
		tb_ld4(cc + TBF_B1 * NN, cB1);
		tb_ld4(cc + TBF_JAC * NN, cJ);

		double wn[4], conUa[4], conUb[4], conUx[4], ke[4], ex[4], fbR[4], fbP[4];
		double dxUa[4], dxUb[4], theta[4];
		{
			double faR[4], faP[4];
			const double sn2 = sn * sn;
			const double cw0 = lv[TBF_CW + 0], cw1 = lv[TBF_CW + 1];
			const double d0 = lv[TBF_CD + 0], d1 = lv[TBF_CD + 1], d2 = lv[TBF_CD + 2];
#pragma unroll
			for (int j = 0; j < 4; j++) {
				// InterpolateREdgeToNode(W) (:817-819)
				double x = 0.0;
				x += cw0 * w0[j];
				x += cw1 * wp[j];
				wn[j] = x;
				const double m2 = sn * cA2[j], m4 = sn * cB2[j];
				const double m5 = cX0[j] + sn2 * cX2[j];
				// Contravariant velocities (:916-929)
				conUa[j] = cA0[j] * u[j] + cA1[j] * v[j] + m2 * x;
				conUb[j] = cA1[j] * u[j] + cB1[j] * v[j] + m4 * x;
				conUx[j] = m2 * u[j] + m4 * v[j] + m5 * x;
				// Specific kinetic energy (:932-935)
				ke[j] = 0.5 * (conUa[j] * u[j] + conUb[j] * v[j] + conUx[j] * x);
				// Exner pressure (:949-951, PhysicalConstants.h:397-399)
				ex[j] = tb_exner(ph, p[j]);
				// Fluxes (:1050-1077)
				const double fa_ = cJ[j] * conUa[j];
				const double fb_ = cJ[j] * conUb[j];
				faR[j] = fa_ * r[j];
				fbR[j] = fb_ * r[j];
				faP[j] = fa_ * p[j];
				fbP[j] = fb_ * p[j];
				theta[j] = p[j] / r[j];
				// DifferentiateNodeToNode of u_alpha, u_beta (:974-1001)
				double d = 0.0;
				d += d0 * um[j]; d += d1 * u[j]; d += d2 * up[j];
				dxUa[j] = d;
				d = 0.0;
				d += d0 * vm[j]; d += d1 * v[j]; d += d2 * vp[j];
				dxUb[j] = d;
			}
			if (active) {
				tb_st4(tWn + k * RS + i * NP, wn);
				tb_st4(tKE + k * RS + i * NP, ke);
				tb_st4(tEX + k * RS + i * NP, ex);
				tb_st4(tFaR + k * RS + i * NP, faR);
				tb_st4(tFaP + k * RS + i * NP, faP);
			}
		}
		__syncwarp();

		if (DO_V && active) {
			tb_st4(tAU + k * RS + i * NP, addU);
			tb_st4(tAV + k * RS + i * NP, addV);
		}

		// alpha-derivatives: sums over s of tile[s][j] (rows held by the other
		// three threads of the level); beta-derivatives: sums over my own row.
		// Two nodes at a time to bound the live registers.
		double dxI[4], stI[4];
#pragma unroll
		for (int s = 0; s < 4; s++) {
			dxI[s] = t.dx[s * NP + i];
			stI[s] = t.st[i * NP + s];
		}
#pragma unroll
		for (int jh = 0; jh < 2; jh++) {
			double dCovDaUb[2], dCovDaUx[2], dDaP[2], dDaKE[2], dDaRhoFluxA[2], dDaPressureFluxA[2];
			tb_cross_sum2(sV + kc * RS + 2 * jh, dxI, dCovDaUb);
			tb_cross_sum2(tWn + kc * RS + 2 * jh, dxI, dCovDaUx);
			tb_cross_sum2(tEX + kc * RS + 2 * jh, dxI, dDaP);
			tb_cross_sum2(tKE + kc * RS + 2 * jh, dxI, dDaKE);
			tb_cross_sum2(tFaR + kc * RS + 2 * jh, stI, dDaRhoFluxA);
			tb_cross_sum2(tFaP + kc * RS + 2 * jh, stI, dDaPressureFluxA);

			const size_t o2 = o4 + 2 * jh;
			double cIJ[2], cFJ[2], cGA[2], cGB[2];
			tb_ld2(cc + TBF_INVJAC * NN + 2 * jh, cIJ);
			tb_ld2(cc + TBF_FJ * NN + 2 * jh, cFJ);
			tb_ld2(cc + TBF_GDA * NN + 2 * jh, cGA);
			tb_ld2(cc + TBF_GDB * NN + 2 * jh, cGB);
			double bU[2], bV[2], bP[2], bR[2];
			tb_stage_base2(sb, out, offU + o2, bU);
			tb_stage_base2(sb, out, offV + o2, bV);
			tb_stage_base2(sb, out, offP + o2, bP);
			tb_stage_base2(sb, out, offR + o2, bR);

			double zx[2];
#pragma unroll
			for (int q = 0; q < 2; q++) {
				const int j = 2 * jh + q;
				double dCovDbUa = 0.0, dCovDbUx = 0.0, dDbP = 0.0, dDbKE = 0.0;
				double dDbRhoFluxB = 0.0, dDbPressureFluxB = 0.0;
#pragma unroll
				for (int s = 0; s < 4; s++) {
					dCovDbUa += u[s] * t.dx[s * NP + j];
					dCovDbUx += wn[s] * t.dx[s * NP + j];
					dDbRhoFluxB -= fbR[s] * t.st[j * NP + s];
					dDbPressureFluxB -= fbP[s] * t.st[j * NP + s];
					dDbP += ex[s] * t.dx[s * NP + j];
					dDbKE += ke[s] * t.dx[s * NP + j];
				}
				const double aDaUb = dCovDaUb[q] * dInvDA;
				const double aDaUx = dCovDaUx[q] * dInvDA;
				const double aDbUa = dCovDbUa * dInvDB;
				const double aDbUx = dCovDbUx * dInvDB;

				// U cross relative vorticity (:966-1039)
				const double dJZetaA = (aDbUx - dxUb[j]);
				const double dJZetaB = (dxUa[j] - aDaUx);
				const double dJZetaX = (aDaUb - aDbUa);
				const double dUCrossZetaA = conUb[j] * dJZetaX - conUx[j] * dJZetaB;
				const double dUCrossZetaB = conUx[j] * dJZetaA - conUa[j] * dJZetaX;
				zx[q] = -conUa[j] * aDaUx - conUb[j] * aDbUx;

				// flux divergences; the alpha sums were accumulated with a plus
				// sign (:1222-1301 subtract term by term)
				const double aDaRho = -(dDaRhoFluxA[q] * dInvDA);
				const double aDbRho = dDbRhoFluxB * dInvDB;
				const double aDaPre = -(dDaPressureFluxA[q] * dInvDA);
				const double aDbPre = dDbPressureFluxB * dInvDB;
				const double aDaP = dDaP[q] * dInvDA;
				const double aDbP = dDbP * dInvDB;
				const double aDaKE = dDaKE[q] * dInvDA;
				const double aDbKE = dDbKE * dInvDB;

				double dLocalUpdateUa = 0.0, dLocalUpdateUb = 0.0;
				dLocalUpdateUa += dUCrossZetaA;
				dLocalUpdateUb += dUCrossZetaB;
				// Coriolis (:1330-1338)
				dLocalUpdateUa += cFJ[q] * conUb[j];
				dLocalUpdateUb -= cFJ[q] * conUa[j];
				// Pressure gradient force, RHOTHETA_PI (:1348-1353): theta * grad Pi
				const double dPGFa = aDaP * theta[j];
				const double dPGFb = aDbP * theta[j];
				// Gravity (:1363-1364): g * DerivR
				const double dDaPhi = sn * cGA[q];
				const double dDbPhi = sn * cGB[q];
				dLocalUpdateUa -= dPGFa + aDaKE + dDaPhi;
				dLocalUpdateUb -= dPGFb + aDbKE + dDbPhi;

				bU[q] = bU[q] + dt * dLocalUpdateUa;
				if (!fa.xz) bV[q] += dt * dLocalUpdateUb;
				// Density and rho-theta (:1399-1421)
				bR[q] = bR[q] - dt * cIJ[q] * (aDaRho + aDbRho);
				bP[q] = bP[q] - dt * cIJ[q] * (aDaPre + aDbPre);
			}
			if (active) {
				tb_st2(out + offR + o2, bR);
				tb_st2(out + offP + o2, bP);
				tb_st2(tZX + k * RS + i * NP + 2 * jh, zx);
				if (k < 3) {
					tb_st2(sUn + k * RS + i * NP + 2 * jh, bU);
					tb_st2(sVn + k * RS + i * NP + 2 * jh, bV);
				}
				if (DO_V) {
					double au[2], av[2];
					tb_ld2(tAU + k * RS + i * NP + 2 * jh, au);
					tb_ld2(tAV + k * RS + i * NP + 2 * jh, av);
					bU[0] += au[0]; bU[1] += au[1];
					bV[0] += av[0]; bV[1] += av[1];
				}
				tb_st2(out + offU + o2, bU);
				tb_st2(out + offV + o2, bV);
			}
		}
	}
	if (!DO_H) return;
	__syncthreads();

	// ---- vertical velocity on interfaces (:1612-1660) ------------------------------
	for (int k = kq; k <= L; k += TBF_KB) {
		const size_t o4 = (size_t)k * NN + i * NP;
		const double * lv = fa.lev + (size_t)k * TBF_LW;
		double bW[4];
		if (k == 0) {
			// bottom boundary: no flow through the surface, with the updated u
			// extrapolated to the surface
			const double se0 = lv[TBF_SE];
			double a[4], b[4], c[4], a2[4], b2[4], c2[4];
			double cA2[4], cB2[4], cX0[4], cX2[4];
			tb_ld4(cc + TBF_A2 * NN, cA2);
			tb_ld4(cc + TBF_B2 * NN, cB2);
			tb_ld4(cc + TBF_X0 * NN, cX0);
			tb_ld4(cc + TBF_X2 * NN, cX2);
			const int l2 = (L > 2) ? 2 : (L - 1);
			const int l1 = (L > 1) ? 1 : 0;
			tb_ld4(sUn + i * NP, a); tb_ld4(sUn + l1 * RS + i * NP, b); tb_ld4(sUn + l2 * RS + i * NP, c);
			tb_ld4(sVn + i * NP, a2); tb_ld4(sVn + l1 * RS + i * NP, b2); tb_ld4(sVn + l2 * RS + i * NP, c2);
#pragma unroll
			for (int j = 0; j < 4; j++) {
				double dU0 = 0.0, dV0 = 0.0;
				dU0 += lv[TBF_CB0 + 0] * a[j]; dU0 += lv[TBF_CB0 + 1] * b[j]; dU0 += lv[TBF_CB0 + 2] * c[j];
				dV0 += lv[TBF_CB0 + 0] * a2[j]; dV0 += lv[TBF_CB0 + 1] * b2[j]; dV0 += lv[TBF_CB0 + 2] * c2[j];
				const double c0 = se0 * cA2[j], c1 = se0 * cB2[j];
				const double cx2 = cX0[j] + (se0 * se0) * cX2[j];
				bW[j] = -(c0 * dU0 + c1 * dV0) / cx2;
			}
			tb_st4(out + offW + o4, bW);
		} else if (k < L) {
			double zm[4], z0[4];
			tb_ld4(tZX + (k - 1) * RS + i * NP, zm);
			tb_ld4(tZX + k * RS + i * NP, z0);
			tb_stage_base4(sb, out, offW + o4, bW);
#pragma unroll
			for (int j = 0; j < 4; j++) {
				double x = 0.0;
				x += lv[TBF_CILO + 0] * zm[j];
				x += lv[TBF_CILO + 1] * z0[j];
				bW[j] += dt * x;
			}
			tb_st4(out + offW + o4, bW);
		} else if (!sb.use_out) {
			tb_stage_base4(sb, out, offW + o4, bW);
			tb_st4(out + offW + o4, bW);
		}
	}
}

///////////////////////////////////////////////////////////////////////////////
// Pipelined variant of the fused stage: persistent blocks, the element's input
// state and stage base prefetched one element ahead with cp.async
// (global -> shared, no register staging), so that HBM latency is off the
// critical path and every SM keeps ~4 elements (2 blocks x 2 buffers) of loads
// in flight.  Rows are 128 bytes (16 nodes); the 16-byte chunk c of row r is
// stored at chunk c ^ (r & 1): the two levels that share a quarter-warp then
// hit different bank groups for both access patterns of the kernel (own 32
// bytes of a row; the whole row, broadcast within the level's four threads).
// Stage base = one source instance with coefficient 1 (Grid::CopyData), which
// is every stage of KGU35 but the last, every ARS stage's first half, and the
// plugin calls on a pre-filled update instance; other combinations take
// k_nh_stage_fast.

#ifdef TB200_EMU
__device__ __forceinline__ void tb_cp16(double * dst, const double * src) {
	dst[0] = src[0]; dst[1] = src[1];
}
__device__ __forceinline__ void tb_cp_commit() {}
template <int N> __device__ __forceinline__ void tb_cp_wait() {}
#else
__device__ __forceinline__ void tb_cp16(double * dst, const double * src) {
	const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
	asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" :: "r"(d), "l"(src) : "memory");
}
__device__ __forceinline__ void tb_cp_commit() {
	asm volatile("cp.async.commit_group;\n" ::: "memory");
}
template <int N> __device__ __forceinline__ void tb_cp_wait() {
	asm volatile("cp.async.wait_group %0;\n" :: "n"(N) : "memory");
}
#endif

// read-only global data (column constants): non-coherent loads
__device__ __forceinline__ void tb_ld2g(const double * p, double (&v)[2]) {
#ifdef TB200_EMU
	v[0] = p[0]; v[1] = p[1];
#else
	const double2 a = __ldg(reinterpret_cast<const double2 *>(p));
	v[0] = a.x; v[1] = a.y;
#endif
}
__device__ __forceinline__ void tb_ld4g(const double * p, double (&v)[4]) {
	double a[2], b[2];
	tb_ld2g(p, a);
	tb_ld2g(p + 2, b);
	v[0] = a[0]; v[1] = a[1]; v[2] = b[0]; v[3] = b[1];
}

// my four nodes (i, 0..3) of a swizzled row
__device__ __forceinline__ void tb_ld4s(const double * row, int par, int i, double (&v)[4]) {
	double a[2], b[2];
	tb_ld2(row + (((2 * i) ^ par) << 1), a);
	tb_ld2(row + (((2 * i + 1) ^ par) << 1), b);
	v[0] = a[0]; v[1] = a[1]; v[2] = b[0]; v[3] = b[1];
}

__device__ __forceinline__ void tb_st4s(double * row, int par, int i, const double (&v)[4]) {
	const double a[2] = {v[0], v[1]};
	const double b[2] = {v[2], v[3]};
	tb_st2(row + (((2 * i) ^ par) << 1), a);
	tb_st2(row + (((2 * i + 1) ^ par) << 1), b);
}

// out[q] = sum_s row[s][2 jh + q] * c[s] on a swizzled row
__device__ __forceinline__ void tb_cross_sum2s(
	const double * row, int par, int jh, const double (&c)[4], double (&o)[2]
) {
	double r0[2], r1[2], r2[2], r3[2];
	tb_ld2(row + (((0 + jh) ^ par) << 1), r0);
	tb_ld2(row + (((2 + jh) ^ par) << 1), r1);
	tb_ld2(row + (((4 + jh) ^ par) << 1), r2);
	tb_ld2(row + (((6 + jh) ^ par) << 1), r3);
#pragma unroll
	for (int q = 0; q < 2; q++) {
		double a = 0.0;
		a += r0[q] * c[0];
		a += r1[q] * c[1];
		a += r2[q] * c[2];
		a += r3[q] * c[3];
		o[q] = a;
	}
}

#define TBP_MAXSRC 2

// tensor maps of the instances a pipelined kernel streams: input, up to two
// stage-base sources (hyperdiffusion: field, base)
struct PipeMaps {
	TbMap in;
	TbMap b0;
	TbMap b1;
};

// 1024-byte aligned start of the dynamic shared memory
__device__ __forceinline__ double * tb_smem_aligned(double * raw) {
#ifdef TB200_EMU
	return raw;
#else
	const unsigned a = (unsigned)__cvta_generic_to_shared(raw);
	return raw + (((1024u - (a & 1023u)) & 1023u) >> 3);
#endif
}

// Stage base of the pipelined kernel: up to TBP_MAXSRC instances (the update
// instance itself counts as one when its coefficient is non-zero), combined in
// Grid::LinearCombineData's order.
struct PipeBase {
	const double * src[TBP_MAXSRC];
	double coeff[TBP_MAXSRC];
	int nsrc;          // 0: the base is the input instance itself (first stage)
};

__host__ __device__ inline size_t tb_pipe_smem_doubles(int nrows, int L, int nsrc, bool fuse = false) {
	// in[2], base[nsrc], tiles Wn, KE, EX, FaR, FaP, ZX, Un/Vn[3],
	// column constants [2], operator windows [L+1]
	// (one source: its buffer is doubled and fetched one element ahead)
	// fused DSS: + beta carry [nrows][2] and corner pair averages [nrows]
	// + 1 KiB so that the element buffers start on a 1024-byte boundary (128-byte
	// swizzle of the bulk tensor copies), + 4 transaction barriers
	return tb_tma_buffer_doubles(nrows) * (2 + (nsrc == 1 ? 2 : nsrc)) + (size_t)(6 * L + 6) * 16
		+ 2 * TBF_NC * 16 + (size_t)(L + 1) * TBF_LWS + (fuse ? (size_t)nrows * 3 : 0) + 128 + 16
		+ 2 * 4 * 16;
}

// my node pair of the stage base at swizzled element offset off
template <int NSRC>
__device__ __forceinline__ void tb_pipe_base2(
	const PipeBase & pb, const double * inb, const double * b0, const double * b1,
	int off, double (&v)[2]
) {
	if (NSRC == 0) {
		tb_ld2(inb + off, v);
		return;
	}
	// first term: value * coefficient (= 0 + value * coefficient of
	// LinearCombineData, and the scaled destination when it leads the list)
	double s[2];
	tb_ld2(b0 + off, s);
	v[0] = s[0] * pb.coeff[0];
	v[1] = s[1] * pb.coeff[0];
	if (NSRC == 2) {
		tb_ld2(b1 + off, s);
		v[0] += s[0] * pb.coeff[1];
		v[1] += s[1] * pb.coeff[1];
	}
}


///////////////////////////////////////////////////////////////////////////////
// Direct stiffness summation fused into the pipelined kernels.
//
// GridCSGLL::ApplyDSS (GridCSGLL.cpp:435-781) averages the duplicates of every
// node shared between elements after each explicit stage and each Laplacian
// application.  As a separate pass it reads and writes the whole state again
// (every 32-byte sector of a row holds an edge node): 7 sweeps per Strang step.
// Here the averaging groups whose members are neighbours inside one patch - all
// but the patch edges, panel seams and rank boundaries - are averaged by the
// kernel that produces the values, while they are in registers or L2:
//
//  * blocks walk *strips* of beta-consecutive elements (element index e, e+1,
//    ...).  Thread (k, i) owns the nodes (i, 0..3) of a row; the beta-edge it
//    shares with the previous element of the strip - node (i, 0) here, (i, 3)
//    there, i = 1, 2 - is averaged in registers against a value carried in
//    shared memory (the store of (i, 3) is deferred by one element);
//  * the alpha-edge row i = 0 and the corners are finished TBF_LAG elements later
//    (tb_fuse_alpha): own raw row and the row i = 3 of the element one
//    alpha-row down - processed at the same pace by another block, found in L2 -
//    are averaged pairwise in alpha, corners then in beta against the pair
//    average kept from the previous element: the association order of the
//    reference (alpha pass, then beta pass) and of the averaging-group kernels
//    (tb200_dss.cuh), hence the same bits;
//  * an element's "done" stamp (launch epoch) is published after its raw values
//    are stored; consumers acquire it before they read the neighbour row.  A
//    block only ever waits for elements of strips that were taken up before its
//    own, all blocks are resident: no deadlock.  The lag keeps the wait off the
//    critical path: finishing an element right after it is produced makes every
//    alpha-row trail the one below by the publish-to-observe latency (2-3 us),
//    74 rows deep in every round of strips (measured: +1.5 ms per launch).
//
// Groups that are not fused (strip ends, patch edges, seams, other ranks) stay
// raw and go through the averaging-group kernels afterwards (a few per cent).

struct FuseArgs {
	const int * strip_first;   // [nstrips] first element of the strip
	const int * strip_len;     // [nstrips] number of beta-consecutive elements
	const int * strip_neb;     // [nstrips] element stride to the alpha-neighbour row in the patch, 0: none
	int nstrips;
	unsigned * done;           // [nelem] epoch of the launch that last produced the element
	unsigned epoch;
};

#define TBF_LAG 3              // elements between producing a row and finishing its alpha edge

struct FuseElem {
	long long e;               // element, -1: none
	int neb;                   // stride to the alpha-neighbour row (fa)
	int fa;                    // alpha-neighbour row in the same patch
	int fb;                    // the strip has a previous element
	int defer;                 // the strip has a next element
};

__device__ __forceinline__ void tb_flag_publish(unsigned * p, unsigned v) {
#ifdef TB200_EMU
	*p = v;
#else
	// release at gpu scope, cumulative over the stores of the other threads of the
	// block (ordered before it by the barrier): MEMBAR.ALL.GPU + STG.STRONG.  No
	// __threadfence() here: it would also invalidate the L1 (CCTL.IVALL), which
	// holds the operator windows of the fused kernels.
#if defined(TBF_PUB_SC)
	__threadfence();
	asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
#elif defined(TBF_PUB_ALLFENCE)
	asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
#else
	asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
#endif
#endif
}

__device__ __forceinline__ void tb_flag_wait(const unsigned * p, unsigned v) {
#ifdef TB200_EMU
	// the emulation runs fused launches on one block, strips in order: the
	// producer has finished
	if (*p != v) { fprintf(stderr, "tb200 emu: fused DSS read an element that was not produced\n"); abort(); }
#else
	// Relaxed polling: ld.acquire.gpu invalidates the whole L1 on every iteration
	// (LDG.STRONG + CCTL.IVALL).  The data read afterwards is loaded with ld.cg
	// (L2, never L1) by the same thread behind the loop's exit branch, so no stale
	// line can be observed and no acquire fence is needed.
	unsigned x;
	do {
		asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(x) : "l"(p) : "memory");
	} while (x != v);
#endif
}

// 4 consecutive doubles from L2 (written by another block a moment ago)
__device__ __forceinline__ void tb_ld4cg(const double * p, double (&v)[4]) {
#ifdef TB200_EMU
	v[0] = p[0]; v[1] = p[1]; v[2] = p[2]; v[3] = p[3];
#else
	const double2 a = __ldcg(reinterpret_cast<const double2 *>(p));
	const double2 b = __ldcg(reinterpret_cast<const double2 *>(p + 2));
	v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
#endif
}

// Walk of one block over its strips: strip s = blockIdx.x, + gridDim.x, ...;
// the strip's description is read once, at its first element.
struct FuseWalk {
	int s, t;                  // strip, position inside it
	int first, len, neb;
};

__device__ __forceinline__ void tb_walk_load(const FuseArgs & fz, FuseWalk & wk) {
	wk.first = __ldg(fz.strip_first + wk.s);
	wk.len = __ldg(fz.strip_len + wk.s);
	wk.neb = __ldg(fz.strip_neb + wk.s);
}

__device__ __forceinline__ FuseElem tb_walk_elem(const FuseWalk & wk) {
	FuseElem fe;
	fe.e = (long long)wk.first + wk.t;
	fe.neb = wk.neb;
	fe.fa = (wk.neb > 0) ? 1 : 0;
	fe.fb = (wk.t > 0) ? 1 : 0;
	fe.defer = (wk.t + 1 < wk.len) ? 1 : 0;
	return fe;
}

// advance to the next element of the block; false when there is none
__device__ __forceinline__ bool tb_walk_next(const FuseArgs & fz, FuseWalk & wk, int stride) {
	if (wk.t + 1 < wk.len) {
		wk.t++;
		return true;
	}
	wk.s += stride;
	wk.t = 0;
	if (wk.s >= fz.nstrips) return false;
	tb_walk_load(fz, wk);
	return true;
}

// Store the node pair (i, 2 jh), (i, 2 jh + 1) of one row of the element at oute.
template <bool FUSE>
__device__ __forceinline__ void tb_out2(
	const FuseElem & fe, double * carry, double * oute, size_t esz,
	int row, int i, int jh, const double (&v)[2]
) {
	double * p = oute + (size_t)row * 16 + 4 * i + 2 * jh;
	if (FUSE && (i == 1 || i == 2)) {
		double w[2] = {v[0], v[1]};
		if (jh == 0) {
			if (fe.fb) {
				// members in group order: (a, b-1)(i, 3), then (a, b)(i, 0)
				w[0] = 0.5 * (carry[row * 2 + (i - 1)] + w[0]);
				*(p - esz + 3) = w[0];
			}
			tb_st2(p, w);
		} else if (fe.defer) {
			carry[row * 2 + (i - 1)] = w[1];
			p[0] = w[0];
		} else {
			tb_st2(p, w);
		}
		return;
	}
	tb_st2(p, v);
}

// ... the four nodes (i, 0..3)
template <bool FUSE>
__device__ __forceinline__ void tb_out4(
	const FuseElem & fe, double * carry, double * oute, size_t esz,
	int row, int i, const double (&v)[4]
) {
	const double a[2] = {v[0], v[1]};
	const double b[2] = {v[2], v[3]};
	tb_out2<FUSE>(fe, carry, oute, esz, row, i, 0, a);
	tb_out2<FUSE>(fe, carry, oute, esz, row, i, 1, b);
}

// Alpha-edge row and corners of element pe, TBF_LAG elements after it was
// produced, in two parts so that the L2 round trip of the loads overlaps the
// prefetch of the next element: tb_alpha_load (rows tid, tid + 128 of the own
// raw row i = 0 and of the row i = 3 one alpha-row down), tb_alpha_finish.
#define TBF_AROWS 2            // rows per thread: nrows <= 2 * TBF_THREADS

struct AlphaRegs {
	double x[TBF_AROWS][4];
	double nb[TBF_AROWS][4];
};

__device__ __forceinline__ unsigned tb_flag_peek(const unsigned * p) {
#ifdef TB200_EMU
	return *p;
#else
	unsigned x;
	asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(x) : "l"(p) : "memory");
	return x;
#endif
}

// `seen`: the neighbour's stamp as peeked one iteration ago (tb_alpha_peek): in
// the steady state it already carries this launch's epoch and nothing is polled.
// The stamp of the element one alpha-row down covers the corner's fourth member
// (e - neb - 1) too: every alpha-row of a patch is cut into the same strips, so
// that element precedes e - neb in the same strip of the same block.
__device__ __forceinline__ void tb_alpha_load(
	const FuseArgs & fz, const FuseElem & pe, const double * out, size_t esz, int nrows,
	int tid, unsigned seen, AlphaRegs & ar
) {
	if (pe.e < 0 || !pe.fa) return;
	// (skipping this poll when a peek of the stamp one iteration earlier had
	// already shown the epoch was measured to return stale rows on the B200,
	// with every publisher-side fence tried: the poll stays)
	(void)seen;
	tb_flag_wait(fz.done + (pe.e - pe.neb), fz.epoch);
#pragma unroll
	for (int q = 0; q < TBF_AROWS; q++) {
		const int r = tid + q * TBF_THREADS;
		if (r < nrows) {
			tb_ld4cg(out + (size_t)pe.e * esz + (size_t)r * 16, ar.x[q]);
			tb_ld4cg(out + (size_t)(pe.e - pe.neb) * esz + (size_t)r * 16 + 12, ar.nb[q]);
		}
	}
}

__device__ __forceinline__ unsigned tb_alpha_peek(const FuseArgs & fz, const FuseElem & pe) {
	if (pe.e < 0 || !pe.fa) return 0u;
	return tb_flag_peek(fz.done + (pe.e - pe.neb));
}

__device__ __forceinline__ void tb_alpha_finish(
	const FuseElem & pe, double * out, size_t esz, int nrows, double * aprev, int tid,
	const AlphaRegs & ar
) {
	if (pe.e < 0 || !pe.fa) return;
#pragma unroll
	for (int q = 0; q < TBF_AROWS; q++) {
		const int r = tid + q * TBF_THREADS;
		if (r >= nrows) continue;
		double * own = out + (size_t)pe.e * esz + (size_t)r * 16;               // (0, 0..3)
		double * nbp = out + (size_t)(pe.e - pe.neb) * esz + (size_t)r * 16 + 12;   // (3, 0..3) one row down
		double A[4];
#pragma unroll
		for (int j = 0; j < 4; j++) A[j] = 0.5 * (ar.nb[q][j] + ar.x[q][j]);
		if (pe.fb) {
			// corner: pair average of the previous element, then this one
			const double C = 0.5 * (aprev[r] + A[0]);
			const double o[2] = {C, A[1]};
			tb_st2(own, o);
			tb_st2(nbp, o);
			*(own - esz + 3) = C;       // (a, b-1)(0, 3)
			*(nbp - esz + 3) = C;       // (a-1, b-1)(3, 3)
		} else {
			own[1] = A[1];
			nbp[1] = A[1];
		}
		own[2] = A[2];
		nbp[2] = A[2];
		if (pe.defer) aprev[r] = A[3];
	}
}

__device__ __forceinline__ void tb_fuse_alpha(
	const FuseArgs & fz, const FuseElem & pe, double * out, size_t esz, int nrows,
	double * aprev, int tid
) {
	AlphaRegs ar;
	tb_alpha_load(fz, pe, out, esz, nrows, tid, 0u, ar);
	tb_alpha_finish(pe, out, esz, nrows, aprev, tid, ar);
}

#ifndef TBP_MINBLOCKS
#define TBP_MINBLOCKS 2
#endif

template <bool DO_V, int NSRC, bool FUSE>
__global__ void __launch_bounds__(TBF_THREADS, TBP_MINBLOCKS)
k_nh_stage_pipe(
	DevLayout lay, DevTables t, DevPhys ph, FastArgs fa,
	const double * __restrict__ in, PipeBase pb, double * out, ElemList el, FuseArgs fz,
	const __grid_constant__ PipeMaps maps
) {
	const int NP = 4, NN = 16;
	const int L = lay.nlev;
	const double dt = fa.dt;
	const int nrows = lay.nrows_state;                   // rows the kernel handles (tracer rows follow)
	const long long nstride = lay.nrows;                 // rows per element in global memory
	const int rU = lay.rowoff[0], rV = lay.rowoff[1], rP = lay.rowoff[2];
	const int rW = lay.rowoff[3], rR = lay.rowoff[4];

	TB_DYN_SMEM(double, sm_raw);
	double * sm = tb_smem_aligned(sm_raw);
	const size_t esz = (size_t)nstride * NN;             // element stride in global memory
	const size_t ebuf = tb_tma_buffer_doubles(nrows);    // element buffer in shared memory
	double * inb0 = sm;
	// one source: [2][ebuf], fetched one element ahead like the input;
	// two sources: [2][ebuf], one buffer each, fetched at the start of the element
	double * bsb0 = sm + 2 * ebuf;
	const bool ahead = (NSRC == 1);
	double * tWn = sm + (2 + (NSRC == 1 ? 2 : NSRC)) * ebuf;
	double * tKE = tWn + (size_t)L * NN;
	double * tEX = tKE + (size_t)L * NN;
	double * tFaR = tEX + (size_t)L * NN;
	double * tFaP = tFaR + (size_t)L * NN;
	double * tZX = tFaP + (size_t)L * NN;
	double * sUn = tZX + (size_t)L * NN;     // [3][16] U after the horizontal update
	double * sVn = sUn + 3 * NN;
	double * scc0 = sVn + 3 * NN;            // [2][TBF_NC][16] column constants
	double * slev = scc0 + 2 * TBF_NC * NN;  // [L+1][TBF_LWS] operator windows
	double * carry = slev + (size_t)(L + 1) * TBF_LWS;   // FUSE: [nrows][2] beta carry of rows i = 1, 2
	double * aprev = carry + (size_t)nrows * 2;   // FUSE: [nrows] alpha pair average of node (0, 3)
	// transaction barriers: input slot 0 / 1 (+ its base when fetched ahead, column
	// constants; slot 0 also the operator windows), the two-source base
	tb_mbar_t * bars = reinterpret_cast<tb_mbar_t *>(
		(FUSE ? aprev + (size_t)nrows : slev + (size_t)(L + 1) * TBF_LWS));
	// Decoupled warps (single level pass, no fused DSS): no block barrier inside
	// the element loop.  A warp only waits for the data it reads: the element
	// (bars[0 / 1], by bulk copy), the (u x zeta)_xi row of the level below its
	// first one (zdone, from the warp below; kept in zb, double-buffered by
	// element parity so that a warp running ahead does not overwrite it), and the
	// issuing thread for every warp to have left a buffer before it is refilled
	// (empty[0 / 1], four arrivals).  Warps drift up to one element apart instead
	// of meeting twice per element.
	tb_mbar_t * empty = bars + 3;              // [2]
	tb_mbar_t * zdone = bars + 5;              // [3 warp boundaries][2 element parities]
	double * zb = reinterpret_cast<double *>(bars + 16);   // [2][4][16]
	// Measured on the B200 (ne = 120, L = 30) and left off: 1.32 / 1.40 / 1.44 ms
	// against 1.21 / 1.25 / 1.28 ms with the two block barriers per element.
#if defined(TBF_DECOUPLED) && !defined(TB200_EMU)
	const bool decoupled = !FUSE && (L <= TBF_KB);
#else
	const bool decoupled = false;
#endif

	const int tid = threadIdx.x;
	const int kq = tid >> 2;
	const int i = tid & 3;
	const int warp = tid >> 5;

	double dxI[4], stI[4];
#pragma unroll
	for (int s = 0; s < 4; s++) {
		dxI[s] = t.dx[s * NP + i];
		stI[s] = t.st[i * NP + s];
	}

	// element walk: positions blockIdx.x, + gridDim.x, ... of the element list,
	// or (FUSE) the strips blockIdx.x, + gridDim.x, ..., each element by element
	long long w = blockIdx.x;          // position in the element list / strip
	if (w >= (FUSE ? (long long)fz.nstrips : (long long)el.n)) return;
	FuseWalk wk;
	wk.s = (int)w; wk.t = 0; wk.first = 0; wk.len = 0; wk.neb = 0;
	AlphaRegs areg;
	FuseElem cur, prev;
	prev.e = -1; prev.neb = 0; prev.fa = 0; prev.fb = 0; prev.defer = 0;
	// the alpha edges trail the walk by TBF_LAG elements: a second walker over the
	// same strips; lag_el = its element once `behind` has reached TBF_LAG, lag_nx
	// the one after it (its stamp is peeked one iteration ahead)
	FuseWalk wa = wk;
	FuseElem lag_el = prev, lag_nx = prev;
	int behind = 0;
	unsigned seen = 0u;
	if (FUSE) {
		tb_walk_load(fz, wk);
		cur = tb_walk_elem(wk);
		wa = wk;
		lag_nx = cur;
	} else {
		cur = prev;
		cur.e = tb_elem(el, w);
	}
	long long e = cur.e;
	(void)in;
	const unsigned ebytes = tb_tma_element_bytes(maps.in);
	const unsigned cbytes = TBF_NC * NN * sizeof(double);
	if (tid == 0) {
		tb_mbar_init(&bars[0], 1);
		tb_mbar_init(&bars[1], 1);
		tb_mbar_init(&bars[2], 1);
		tb_mbar_init(&empty[0], TBF_THREADS / 32);
		tb_mbar_init(&empty[1], TBF_THREADS / 32);
		for (int q = 0; q < 6; q++) tb_mbar_init(&zdone[q], 1);
		tb_mbar_fence_init();
	}
	__syncthreads();
	if (tid == 0) {
		// first element, its column constants and the operator windows
		tb_mbar_expect(&bars[0], ebytes * (ahead ? 2u : 1u) + cbytes
			+ (unsigned)((L + 1) * TBF_LWK * sizeof(double)));
		tb_tma_element(inb0, maps.in, e * nstride, nrows, &bars[0]);
		if (ahead) tb_tma_element(bsb0, maps.b0, e * nstride, nrows, &bars[0]);
		tb_bulk_1d(scc0, fa.colc + (size_t)e * TBF_NC * NN, cbytes, &bars[0]);
		for (int r = 0; r <= L; r++) {
			tb_bulk_1d(slev + (size_t)r * TBF_LWS, fa.lev + (size_t)r * TBF_LW,
				TBF_LWK * sizeof(double), &bars[0]);
		}
	}

	bool valid = true;
	for (int it = 0; valid; it++) {
		e = cur.e;
		// the element after this one
		FuseElem nxt = cur;
		bool has_next;
		if (FUSE) {
			has_next = tb_walk_next(fz, wk, (int)gridDim.x);
			if (has_next) nxt = tb_walk_elem(wk);
		} else {
			w += gridDim.x;
			has_next = (w < el.n);
			if (has_next) nxt.e = tb_elem(el, w);
		}
		const int buf = it & 1;
		const double * inb = inb0 + (size_t)buf * ebuf;
		// the data of this element (issued one iteration ago) has landed, and
		// every thread is done with the previous element
		tb_mbar_wait(&bars[buf], (unsigned)(it >> 1) & 1u);
		if (!decoupled) __syncthreads();
		if (FUSE) {
			// the previous element's raw values are stored: publish it; finish the
			// alpha edge and corners of the element produced TBF_LAG iterations ago
			// (the row one down was published long since)
			if (tid == 0 && prev.e >= 0) tb_flag_publish(fz.done + prev.e, fz.epoch);
			if (behind == TBF_LAG) {
				// lag_nx becomes the element whose alpha edge is finished now
				lag_el = lag_nx;
				tb_alpha_load(fz, lag_el, out, esz, nrows, tid, seen, areg);
				if (tb_walk_next(fz, wa, (int)gridDim.x)) lag_nx = tb_walk_elem(wa);
				seen = tb_alpha_peek(fz, lag_nx);
			} else {
				lag_el.e = -1;
			}
		}
		// stage base.  Two sources: fetched now for this element (consumed at the
		// end of its level loop, behind bars[2]).  One source: fetched below, one
		// element ahead.  One thread issues the bulk copies.
		const double * bp0 = ahead ? (bsb0 + (size_t)buf * ebuf) : bsb0;
		const double * bp1 = bsb0 + ebuf;
		// decoupled, one source or none: the prefetch is issued after this warp's
		// level loop (below), when every warp has long left the other buffer
#ifdef TBF_ISSUE_LATE
		const bool issue_late = decoupled && (NSRC != 2);
#else
		const bool issue_late = false;
#endif
		if (tid == 0 && !issue_late) {
			// every warp has left the previous element (its buffers are refilled now)
			if (decoupled && it > 0) tb_mbar_wait(&empty[buf ^ 1], (unsigned)((it - 1) >> 1) & 1u);
			if (NSRC == 2) {
				tb_mbar_expect(&bars[2], 2u * ebytes);
				tb_tma_element(bsb0, maps.b0, e * nstride, nrows, &bars[2]);
				tb_tma_element(bsb0 + ebuf, maps.b1, e * nstride, nrows, &bars[2]);
			}
			// prefetch the next element of this block into the other buffer
			if (has_next) {
				const long long en = nxt.e;
				tb_mbar_expect(&bars[buf ^ 1], ebytes * (ahead ? 2u : 1u) + cbytes);
				tb_tma_element(inb0 + (size_t)(buf ^ 1) * ebuf, maps.in, en * nstride, nrows, &bars[buf ^ 1]);
				if (ahead) {
					tb_tma_element(bsb0 + (size_t)(buf ^ 1) * ebuf, maps.b0, en * nstride, nrows, &bars[buf ^ 1]);
				}
				tb_bulk_1d(scc0 + (size_t)(buf ^ 1) * TBF_NC * NN,
					fa.colc + (size_t)en * TBF_NC * NN, cbytes, &bars[buf ^ 1]);
			}
		}
		if (FUSE) tb_alpha_finish(lag_el, out, esz, nrows, aprev, tid, areg);

		const size_t ebase = (size_t)e * esz;
		const double * cc = scc0 + (size_t)buf * TBF_NC * NN + i * NP;
		const double dInvDA = __ldg(fa.inv_da + e);
		const double dInvDB = __ldg(fa.inv_db + e);

		for (int k0 = 0; k0 < L; k0 += TBF_KB) {
			const int k = k0 + kq;
			const bool active = (k < L);
			const int kc = active ? k : (L - 1);
			const int km = (kc > 0) ? kc - 1 : 0;
			const int kp = (kc < L - 1) ? kc + 1 : L - 1;
			const double * lv = slev + (size_t)kc * TBF_LWS;
#define LV(q) lv[(q)]
			const double sn = LV(TBF_SN);
			const int tp = kc & 1;                         // tile parity

			double cA2[4], cB2[4], cX0[4], cX2[4];
			tb_ld4(cc + TBF_A2 * NN, cA2);
			tb_ld4(cc + TBF_B2 * NN, cB2);
			tb_ld4(cc + TBF_X0 * NN, cX0);
			tb_ld4(cc + TBF_X2 * NN, cX2);

			double u[4], wn[4], conUa[4], conUb[4], conUx[4], ke[4], ex[4], fbR[4], fbP[4];
			double dxUa[4], dxUb[4], theta[4];
			{
				double v[4], w0[4], wp[4], p[4], r[4], um[4], up[4], vm[4], vp[4];
				tb_ld4s(inb + (size_t)(rU + kc) * NN, (rU + kc) & 7, i, u);
				tb_ld4s(inb + (size_t)(rV + kc) * NN, (rV + kc) & 7, i, v);
				tb_ld4s(inb + (size_t)(rW + kc) * NN, (rW + kc) & 7, i, w0);
				tb_ld4s(inb + (size_t)(rW + kc + 1) * NN, (rW + kc + 1) & 7, i, wp);
				tb_ld4s(inb + (size_t)(rP + kc) * NN, (rP + kc) & 7, i, p);
				tb_ld4s(inb + (size_t)(rR + kc) * NN, (rR + kc) & 7, i, r);
				tb_ld4s(inb + (size_t)(rU + km) * NN, (rU + km) & 7, i, um);
				tb_ld4s(inb + (size_t)(rU + kp) * NN, (rU + kp) & 7, i, up);
				tb_ld4s(inb + (size_t)(rV + km) * NN, (rV + km) & 7, i, vm);
				tb_ld4s(inb + (size_t)(rV + kp) * NN, (rV + kp) & 7, i, vp);
				double cA0[4], cA1[4], cB1[4], cJ[4];
				tb_ld4(cc + TBF_A0 * NN, cA0);
				tb_ld4(cc + TBF_A1 * NN, cA1);
				tb_ld4(cc + TBF_B1 * NN, cB1);
				tb_ld4(cc + TBF_JAC * NN, cJ);
				double faR[4], faP[4];
				const double sn2 = sn * sn;
				const double cw0 = LV(TBF_CW + 0), cw1 = LV(TBF_CW + 1);
				const double d0 = LV(TBF_CD + 0), d1 = LV(TBF_CD + 1), d2 = LV(TBF_CD + 2);
#pragma unroll
				for (int j = 0; j < 4; j++) {
					// InterpolateREdgeToNode(W) (:817-819)
					double x = 0.0;
					x += cw0 * w0[j];
					x += cw1 * wp[j];
					wn[j] = x;
					const double m2 = sn * cA2[j], m4 = sn * cB2[j];
					const double m5 = cX0[j] + sn2 * cX2[j];
					// Contravariant velocities (:916-929)
					conUa[j] = cA0[j] * u[j] + cA1[j] * v[j] + m2 * x;
					conUb[j] = cA1[j] * u[j] + cB1[j] * v[j] + m4 * x;
					conUx[j] = m2 * u[j] + m4 * v[j] + m5 * x;
					// Specific kinetic energy (:932-935)
					ke[j] = 0.5 * (conUa[j] * u[j] + conUb[j] * v[j] + conUx[j] * x);
					// Exner pressure (:949-951, PhysicalConstants.h:397-399)
					ex[j] = tb_exner(ph, p[j]);
					// Fluxes (:1050-1077)
					const double fa_ = cJ[j] * conUa[j];
					const double fb_ = cJ[j] * conUb[j];
					faR[j] = fa_ * r[j];
					fbR[j] = fb_ * r[j];
					faP[j] = fa_ * p[j];
					fbP[j] = fb_ * p[j];
					theta[j] = p[j] / r[j];
					// DifferentiateNodeToNode of u_alpha, u_beta (:974-1001)
					double d = 0.0;
					d += d0 * um[j]; d += d1 * u[j]; d += d2 * up[j];
					dxUa[j] = d;
					d = 0.0;
					d += d0 * vm[j]; d += d1 * v[j]; d += d2 * vp[j];
					dxUb[j] = d;
				}
				if (active) {
					tb_st4s(tWn + (size_t)k * NN, tp, i, wn);
					tb_st4s(tKE + (size_t)k * NN, tp, i, ke);
					tb_st4s(tEX + (size_t)k * NN, tp, i, ex);
					tb_st4s(tFaR + (size_t)k * NN, tp, i, faR);
					tb_st4s(tFaP + (size_t)k * NN, tp, i, faP);
				}
			}
			__syncwarp();

#pragma unroll
			for (int jh = 0; jh < 2; jh++) {
				double dCovDaUb[2], dCovDaUx[2], dDaP[2], dDaKE[2], dDaRhoFluxA[2], dDaPressureFluxA[2];
				tb_cross_sum2s(inb + (size_t)(rV + kc) * NN, (rV + kc) & 7, jh, dxI, dCovDaUb);
				tb_cross_sum2s(tWn + (size_t)kc * NN, tp, jh, dxI, dCovDaUx);
				tb_cross_sum2s(tEX + (size_t)kc * NN, tp, jh, dxI, dDaP);
				tb_cross_sum2s(tKE + (size_t)kc * NN, tp, jh, dxI, dDaKE);
				tb_cross_sum2s(tFaR + (size_t)kc * NN, tp, jh, stI, dDaRhoFluxA);
				tb_cross_sum2s(tFaP + (size_t)kc * NN, tp, jh, stI, dDaPressureFluxA);

				double cIJ[2], cFJ[2], cGA[2], cGB[2];
				tb_ld2(cc + TBF_INVJAC * NN + 2 * jh, cIJ);
				tb_ld2(cc + TBF_FJ * NN + 2 * jh, cFJ);
				tb_ld2(cc + TBF_GDA * NN + 2 * jh, cGA);
				tb_ld2(cc + TBF_GDB * NN + 2 * jh, cGB);
				// stage base: my node pair of the four level components
				const int ch = 2 * i + jh;
				double bU[2], bV[2], bP[2], bR[2];
				if (jh == 0 && NSRC == 2) tb_mbar_wait(&bars[2], (unsigned)it & 1u);   // the bases have landed
				tb_pipe_base2<NSRC>(pb, inb, bp0, bp1, (rU + kc) * NN + ((ch ^ ((rU + kc) & 7)) << 1), bU);
				tb_pipe_base2<NSRC>(pb, inb, bp0, bp1, (rV + kc) * NN + ((ch ^ ((rV + kc) & 7)) << 1), bV);
				tb_pipe_base2<NSRC>(pb, inb, bp0, bp1, (rP + kc) * NN + ((ch ^ ((rP + kc) & 7)) << 1), bP);
				tb_pipe_base2<NSRC>(pb, inb, bp0, bp1, (rR + kc) * NN + ((ch ^ ((rR + kc) & 7)) << 1), bR);

				double zx[2];
#pragma unroll
				for (int q = 0; q < 2; q++) {
					const int j = 2 * jh + q;
					double dCovDbUa = 0.0, dCovDbUx = 0.0, dDbP = 0.0, dDbKE = 0.0;
					double dDbRhoFluxB = 0.0, dDbPressureFluxB = 0.0;
#pragma unroll
					for (int s = 0; s < 4; s++) {
						dCovDbUa += u[s] * t.dx[s * NP + j];
						dCovDbUx += wn[s] * t.dx[s * NP + j];
						dDbRhoFluxB -= fbR[s] * t.st[j * NP + s];
						dDbPressureFluxB -= fbP[s] * t.st[j * NP + s];
						dDbP += ex[s] * t.dx[s * NP + j];
						dDbKE += ke[s] * t.dx[s * NP + j];
					}
					const double aDaUb = dCovDaUb[q] * dInvDA;
					const double aDaUx = dCovDaUx[q] * dInvDA;
					const double aDbUa = dCovDbUa * dInvDB;
					const double aDbUx = dCovDbUx * dInvDB;

					// U cross relative vorticity (:966-1039)
					const double dJZetaA = (aDbUx - dxUb[j]);
					const double dJZetaB = (dxUa[j] - aDaUx);
					const double dJZetaX = (aDaUb - aDbUa);
					const double dUCrossZetaA = conUb[j] * dJZetaX - conUx[j] * dJZetaB;
					const double dUCrossZetaB = conUx[j] * dJZetaA - conUa[j] * dJZetaX;
					zx[q] = -conUa[j] * aDaUx - conUb[j] * aDbUx;

					const double aDaRho = -(dDaRhoFluxA[q] * dInvDA);
					const double aDbRho = dDbRhoFluxB * dInvDB;
					const double aDaPre = -(dDaPressureFluxA[q] * dInvDA);
					const double aDbPre = dDbPressureFluxB * dInvDB;
					const double aDaP = dDaP[q] * dInvDA;
					const double aDbP = dDbP * dInvDB;
					const double aDaKE = dDaKE[q] * dInvDA;
					const double aDbKE = dDbKE * dInvDB;

					double dLocalUpdateUa = 0.0, dLocalUpdateUb = 0.0;
					dLocalUpdateUa += dUCrossZetaA;
					dLocalUpdateUb += dUCrossZetaB;
					// Coriolis (:1330-1338)
					dLocalUpdateUa += cFJ[q] * conUb[j];
					dLocalUpdateUb -= cFJ[q] * conUa[j];
					// Pressure gradient force (:1348-1353), gravity (:1363-1364)
					const double dPGFa = aDaP * theta[j];
					const double dPGFb = aDbP * theta[j];
					const double dDaPhi = sn * cGA[q];
					const double dDbPhi = sn * cGB[q];
					dLocalUpdateUa -= dPGFa + aDaKE + dDaPhi;
					dLocalUpdateUb -= dPGFb + aDbKE + dDbPhi;

					bU[q] = bU[q] + dt * dLocalUpdateUa;
					if (!fa.xz) bV[q] += dt * dLocalUpdateUb;
					// Density and rho-theta (:1399-1421)
					bR[q] = bR[q] - dt * cIJ[q] * (aDaRho + aDbRho);
					bP[q] = bP[q] - dt * cIJ[q] * (aDaPre + aDbPre);
				}
				if (active) {
					tb_out2<FUSE>(cur, carry, out + ebase, esz, rR + k, i, jh, bR);
					tb_out2<FUSE>(cur, carry, out + ebase, esz, rP + k, i, jh, bP);
					tb_st2(tZX + (size_t)k * NN + ((ch ^ tp) << 1), zx);
					if (decoupled && (k & 7) == 7) {
						tb_st2(zb + ((size_t)(it & 1) * 4 + warp) * NN + 4 * i + 2 * jh, zx);
					}
					if (k < 3) {
						tb_st2(sUn + k * NN + 4 * i + 2 * jh, bU);
						tb_st2(sVn + k * NN + 4 * i + 2 * jh, bV);
					}
				}
				if (DO_V) {
					// upwind penalty on U and V for this node pair
					// (VerticalDynamicsFEM.cpp:816-828, 998-1023)
					const double se0 = LV(TBF_SE), se1 = LV(TBF_SE1);
					// the windows of the skipped sides (top / bottom level) are zero
					double u0[2], v0[2], um[2], up[2], vm[2], vp[2], w0[2], wp[2];
					tb_ld2(inb + (size_t)(rU + kc) * NN + ((ch ^ ((rU + kc) & 7)) << 1), u0);
					tb_ld2(inb + (size_t)(rV + kc) * NN + ((ch ^ ((rV + kc) & 7)) << 1), v0);
					tb_ld2(inb + (size_t)(rU + km) * NN + ((ch ^ ((rU + km) & 7)) << 1), um);
					tb_ld2(inb + (size_t)(rU + kp) * NN + ((ch ^ ((rU + kp) & 7)) << 1), up);
					tb_ld2(inb + (size_t)(rV + km) * NN + ((ch ^ ((rV + km) & 7)) << 1), vm);
					tb_ld2(inb + (size_t)(rV + kp) * NN + ((ch ^ ((rV + kp) & 7)) << 1), vp);
					tb_ld2(inb + (size_t)(rW + kc) * NN + ((ch ^ ((rW + kc) & 7)) << 1), w0);
					tb_ld2(inb + (size_t)(rW + kc + 1) * NN + ((ch ^ ((rW + kc + 1) & 7)) << 1), wp);
#pragma unroll
					for (int q = 0; q < 2; q++) {
						const int j = 2 * jh + q;
						double au = 0.0, av = 0.0;
						{
							double ue = 0.0, ve = 0.0;
							ue += LV(TBF_CIHI + 0) * um[q]; ue += LV(TBF_CIHI + 1) * u0[q]; ue += LV(TBF_CIHI + 2) * up[q];
							ve += LV(TBF_CIHI + 0) * vm[q]; ve += LV(TBF_CIHI + 1) * v0[q]; ve += LV(TBF_CIHI + 2) * vp[q];
							const double c0 = se1 * cA2[j], c1 = se1 * cB2[j];
							const double c2 = cX0[j] + (se1 * se1) * cX2[j];
							const double xd = c0 * ue + c1 * ve + c2 * wp[q];
							const double wgt = dt * fabs(xd);
							double pu = 0.0, pv = 0.0;
							pu += LV(TBF_CPL + 0) * um[q]; pu += LV(TBF_CPL + 1) * u0[q]; pu += LV(TBF_CPL + 2) * up[q];
							pv += LV(TBF_CPL + 0) * vm[q]; pv += LV(TBF_CPL + 1) * v0[q]; pv += LV(TBF_CPL + 2) * vp[q];
							au += pu * wgt;
							av += pv * wgt;
						}
						{
							double ue = 0.0, ve = 0.0;
							ue += LV(TBF_CILO + 0) * um[q]; ue += LV(TBF_CILO + 1) * u0[q]; ue += LV(TBF_CILO + 2) * up[q];
							ve += LV(TBF_CILO + 0) * vm[q]; ve += LV(TBF_CILO + 1) * v0[q]; ve += LV(TBF_CILO + 2) * vp[q];
							const double c0 = se0 * cA2[j], c1 = se0 * cB2[j];
							const double c2 = cX0[j] + (se0 * se0) * cX2[j];
							const double xd = c0 * ue + c1 * ve + c2 * w0[q];
							const double wgt = dt * fabs(xd);
							double pu = 0.0, pv = 0.0;
							pu += LV(TBF_CPR + 0) * um[q]; pu += LV(TBF_CPR + 1) * u0[q]; pu += LV(TBF_CPR + 2) * up[q];
							pv += LV(TBF_CPR + 0) * vm[q]; pv += LV(TBF_CPR + 1) * v0[q]; pv += LV(TBF_CPR + 2) * vp[q];
							au += pu * wgt;
							av += pv * wgt;
						}
						bU[q] += au;
						bV[q] += av;
					}
				}
				if (active) {
					tb_out2<FUSE>(cur, carry, out + ebase, esz, rU + k, i, jh, bU);
					tb_out2<FUSE>(cur, carry, out + ebase, esz, rV + k, i, jh, bV);
				}
			}
		}
		if (decoupled) {
			// this warp's tiles are complete; tell the warp above, wait for the one below
			__syncwarp();
			if ((tid & 31) == 0 && warp < 3) tb_mbar_arrive(&zdone[warp * 2 + (it & 1)]);
			if (issue_late && tid == 0 && has_next) {
				if (it > 0) tb_mbar_wait(&empty[buf ^ 1], (unsigned)((it - 1) >> 1) & 1u);
				const long long en = nxt.e;
				tb_mbar_expect(&bars[buf ^ 1], ebytes * (ahead ? 2u : 1u) + cbytes);
				tb_tma_element(inb0 + (size_t)(buf ^ 1) * ebuf, maps.in, en * nstride, nrows, &bars[buf ^ 1]);
				if (ahead) {
					tb_tma_element(bsb0 + (size_t)(buf ^ 1) * ebuf, maps.b0, en * nstride, nrows, &bars[buf ^ 1]);
				}
				tb_bulk_1d(scc0 + (size_t)(buf ^ 1) * TBF_NC * NN,
					fa.colc + (size_t)en * TBF_NC * NN, cbytes, &bars[buf ^ 1]);
			}
			if (warp > 0) tb_mbar_wait(&zdone[(warp - 1) * 2 + (it & 1)], (unsigned)(it >> 1) & 1u);
		} else {
			__syncthreads();
		}

		// ---- vertical velocity on interfaces (:1612-1660) --------------------------
		for (int k = kq; k <= L; k += TBF_KB) {
			const double * lv = slev + (size_t)k * TBF_LWS;
			double bW[4];
			if (k == 0) {
				const double se0 = LV(TBF_SE);
				double a[4], b[4], c[4], a2[4], b2[4], c2[4];
				double cA2[4], cB2[4], cX0[4], cX2[4];
				tb_ld4(cc + TBF_A2 * NN, cA2);
				tb_ld4(cc + TBF_B2 * NN, cB2);
				tb_ld4(cc + TBF_X0 * NN, cX0);
				tb_ld4(cc + TBF_X2 * NN, cX2);
				const int l2 = (L > 2) ? 2 : (L - 1);
				const int l1 = (L > 1) ? 1 : 0;
				tb_ld4(sUn + i * NP, a); tb_ld4(sUn + l1 * NN + i * NP, b); tb_ld4(sUn + l2 * NN + i * NP, c);
				tb_ld4(sVn + i * NP, a2); tb_ld4(sVn + l1 * NN + i * NP, b2); tb_ld4(sVn + l2 * NN + i * NP, c2);
#pragma unroll
				for (int j = 0; j < 4; j++) {
					double dU0 = 0.0, dV0 = 0.0;
					dU0 += LV(TBF_CB0 + 0) * a[j]; dU0 += LV(TBF_CB0 + 1) * b[j]; dU0 += LV(TBF_CB0 + 2) * c[j];
					dV0 += LV(TBF_CB0 + 0) * a2[j]; dV0 += LV(TBF_CB0 + 1) * b2[j]; dV0 += LV(TBF_CB0 + 2) * c2[j];
					const double c0 = se0 * cA2[j], c1 = se0 * cB2[j];
					const double cx2 = cX0[j] + (se0 * se0) * cX2[j];
					bW[j] = -(c0 * dU0 + c1 * dV0) / cx2;
				}
			} else {
				{
					const int par = (rW + k) & 7;
					double lo[2], hi[2];
					tb_pipe_base2<NSRC>(pb, inb, bp0, bp1, (rW + k) * NN + (((2 * i) ^ par) << 1), lo);
					tb_pipe_base2<NSRC>(pb, inb, bp0, bp1, (rW + k) * NN + (((2 * i + 1) ^ par) << 1), hi);
					bW[0] = lo[0]; bW[1] = lo[1]; bW[2] = hi[0]; bW[3] = hi[1];
				}
				if (k < L) {
					double zm[4], z0[4];
					if (decoupled && (k & 7) == 0) {
						// the level below belongs to the warp below
						tb_ld4(zb + ((size_t)(it & 1) * 4 + (warp - 1)) * NN + 4 * i, zm);
					} else {
						tb_ld4s(tZX + (size_t)(k - 1) * NN, (k - 1) & 1, i, zm);
					}
					tb_ld4s(tZX + (size_t)k * NN, k & 1, i, z0);
#pragma unroll
					for (int j = 0; j < 4; j++) {
						double x = 0.0;
						x += LV(TBF_CILO + 0) * zm[j];
						x += LV(TBF_CILO + 1) * z0[j];
						bW[j] += dt * x;
					}
				}
			}
			tb_out4<FUSE>(cur, carry, out + ebase, esz, rW + k, i, bW);
		}
		if (decoupled) {
			// this warp has left the element (its buffers, tiles and boundary row)
			__syncwarp();
			if ((tid & 31) == 0) tb_mbar_arrive(&empty[buf]);
		}
#undef LV
		if (FUSE && behind < TBF_LAG) behind++;
		prev = cur;
		cur = nxt;
		valid = has_next;
	}
	if (FUSE) {
		// the last element of this block, then the alpha edges still pending
		// (oldest first: the corner pair average passes from one to the next)
		__syncthreads();
		if (tid == 0 && prev.e >= 0) tb_flag_publish(fz.done + prev.e, fz.epoch);
		for (int q = 0; q < behind; q++) {
			tb_fuse_alpha(fz, lag_nx, out, esz, nrows, aprev, tid);
			if (tb_walk_next(fz, wa, (int)gridDim.x)) lag_nx = tb_walk_elem(wa);
		}
	}
}

///////////////////////////////////////////////////////////////////////////////
// Hyperdiffusion, one pass over all prognostic fields:
//   out = base - dt nu L(fld)       (base = 0 when HAS_BASE is false)
// with L the scalar Laplacian on rho-theta, w, rho
// (HorizontalDynamicsFEM::ApplyScalarHyperdiffusion, :1867-2203) and the vector
// Laplacian on (u_alpha, u_beta) (ApplyVectorHyperdiffusion with
// GridPatchCSGLL::ComputeCurlAndDiv inlined, :2207-2414,
// GridPatchCSGLL.cpp:1132-1305).  The reference's order-4 sequence
//   work = 0; work -= L(in); DSS(work); out = in; out -= (-dt) nu_loc L(work); DSS(out)
// (:2687-2713) becomes two launches of this kernel around the DSS calls, with
// the ZeroData / CopyData passes folded in: 8 S bytes per node in all
//  instead of 13 S.  Same persistent, cp.async-pipelined skeleton
// and the same (level, element row) thread layout as k_nh_stage_pipe.

struct HyperFastArgs {
	const double * colc;
	const double * inv_da;
	const double * inv_db;
	const double * nu_scale;
	double dt;
	double nu_scalar, nu_div, nu_vort;
	int scale_nu;
	int xz;
};

__host__ __device__ inline size_t tb_hyper_smem_doubles(int nrows, int L, bool has_base, bool fuse = false) {
	// fld[2] (+ base[2]), tiles GaP, GaR, JUa, Div, Curl [L], GaW [L+1], column constants [2]
	// (+ fused DSS: beta carry [nrows][2], corner pair averages [nrows])
	return tb_tma_buffer_doubles(nrows) * (has_base ? 4 : 2) + (size_t)(6 * L + 1) * 16 + 2 * TBF_NC * 16
		+ (fuse ? (size_t)nrows * 3 : 0) + 128 + 4;
}

// beta-direction sum over my own row: o = sum_s x[s] * c[s*4 + j] (c = dx) or
// c[j*4 + s] (c = st)
#define TB_ROW_DX(x, j) ((((0.0 + (x)[0] * t.dx[0 * 4 + (j)]) + (x)[1] * t.dx[1 * 4 + (j)]) \
	+ (x)[2] * t.dx[2 * 4 + (j)]) + (x)[3] * t.dx[3 * 4 + (j)])
#define TB_ROW_ST(x, j) ((((0.0 + (x)[0] * t.st[(j) * 4 + 0]) + (x)[1] * t.st[(j) * 4 + 1]) \
	+ (x)[2] * t.st[(j) * 4 + 2]) + (x)[3] * t.st[(j) * 4 + 3])

// alpha-direction sum over a swizzled row for all four nodes of my row
__device__ __forceinline__ void tb_cross_sum4s(
	const double * row, int par, const double (&c)[4], double (&o)[4]
) {
	double lo[2], hi[2];
	tb_cross_sum2s(row, par, 0, c, lo);
	tb_cross_sum2s(row, par, 1, c, hi);
	o[0] = lo[0]; o[1] = lo[1]; o[2] = hi[0]; o[3] = hi[1];
}

template <bool HAS_BASE, bool FUSE>
__global__ void __launch_bounds__(TBF_THREADS, 2)
k_hyper_pipe(
	DevLayout lay, DevTables t, HyperFastArgs ha,
	const double * __restrict__ fld, const double * base, double * out, ElemList el, FuseArgs fz,
	const __grid_constant__ PipeMaps maps
) {
	const int NP = 4, NN = 16;
	const int L = lay.nlev;
	const int nrows = lay.nrows_state;                   // rows the kernel handles (tracer rows follow)
	const long long nstride = lay.nrows;                 // rows per element in global memory
	const int rU = lay.rowoff[0], rV = lay.rowoff[1], rP = lay.rowoff[2];
	const int rW = lay.rowoff[3], rR = lay.rowoff[4];

	TB_DYN_SMEM(double, sm_raw);
	double * sm = tb_smem_aligned(sm_raw);
	const size_t esz = (size_t)nstride * NN;             // element stride in global memory
	const size_t ebuf = tb_tma_buffer_doubles(nrows);    // element buffer in shared memory
	double * fb0 = sm;
	double * bb0 = sm + 2 * ebuf;
	double * tGP = sm + (HAS_BASE ? 4 : 2) * ebuf;
	double * tGR = tGP + (size_t)L * NN;
	double * tJU = tGR + (size_t)L * NN;
	double * tDV = tJU + (size_t)L * NN;
	double * tCL = tDV + (size_t)L * NN;
	double * tGW = tCL + (size_t)L * NN;         // [L+1]
	double * scc0 = tGW + (size_t)(L + 1) * NN;  // [2][TBF_NC][16]
	double * carry = scc0 + 2 * TBF_NC * NN;     // FUSE: [nrows][2]
	double * aprev = carry + (size_t)nrows * 2;  // FUSE: [nrows]
	tb_mbar_t * bars = reinterpret_cast<tb_mbar_t *>(FUSE ? aprev + (size_t)nrows : carry);

	const int tid = threadIdx.x;
	const int kq = tid >> 2;
	const int i = tid & 3;

	double dxI[4], stI[4];
#pragma unroll
	for (int s = 0; s < 4; s++) {
		dxI[s] = t.dx[s * NP + i];
		stI[s] = t.st[i * NP + s];
	}

	long long w = blockIdx.x;          // position in the element list / strip
	if (w >= (FUSE ? (long long)fz.nstrips : (long long)el.n)) return;
	FuseWalk wk;
	wk.s = (int)w; wk.t = 0; wk.first = 0; wk.len = 0; wk.neb = 0;
	AlphaRegs areg;
	FuseElem cur, prev;
	prev.e = -1; prev.neb = 0; prev.fa = 0; prev.fb = 0; prev.defer = 0;
	// the alpha edges trail the walk by TBF_LAG elements: a second walker over the
	// same strips; lag_el = its element once `behind` has reached TBF_LAG, lag_nx
	// the one after it (its stamp is peeked one iteration ahead)
	FuseWalk wa = wk;
	FuseElem lag_el = prev, lag_nx = prev;
	int behind = 0;
	unsigned seen = 0u;
	if (FUSE) {
		tb_walk_load(fz, wk);
		cur = tb_walk_elem(wk);
		wa = wk;
		lag_nx = cur;
	} else {
		cur = prev;
		cur.e = tb_elem(el, w);
	}
	long long e = cur.e;
	(void)fld; (void)base;
	const unsigned ebytes = tb_tma_element_bytes(maps.in);
	const unsigned cbytes = TBF_NC * NN * sizeof(double);
	if (tid == 0) {
		tb_mbar_init(&bars[0], 1);
		tb_mbar_init(&bars[1], 1);
		tb_mbar_fence_init();
	}
	__syncthreads();
	if (tid == 0) {
		tb_mbar_expect(&bars[0], ebytes * (HAS_BASE ? 2u : 1u) + cbytes);
		tb_tma_element(fb0, maps.in, e * nstride, nrows, &bars[0]);
		if (HAS_BASE) tb_tma_element(bb0, maps.b0, e * nstride, nrows, &bars[0]);
		tb_bulk_1d(scc0, ha.colc + (size_t)e * TBF_NC * NN, cbytes, &bars[0]);
	}

	bool valid = true;
	for (int it = 0; valid; it++) {
		e = cur.e;
		FuseElem nxt = cur;
		bool has_next;
		if (FUSE) {
			has_next = tb_walk_next(fz, wk, (int)gridDim.x);
			if (has_next) nxt = tb_walk_elem(wk);
		} else {
			w += gridDim.x;
			has_next = (w < el.n);
			if (has_next) nxt.e = tb_elem(el, w);
		}
		const int buf = it & 1;
		const double * fb = fb0 + (size_t)buf * ebuf;
		const double * bb = bb0 + (size_t)buf * ebuf;
		tb_mbar_wait(&bars[buf], (unsigned)(it >> 1) & 1u);
#if defined(TBF_PUB_ALLFENCE) && !defined(TB200_EMU)
		if (FUSE) asm volatile("fence.acq_rel.gpu;" ::: "memory");
#endif
		__syncthreads();
		if (FUSE) {
			// the previous element's raw values are stored: publish it; finish the
			// alpha edge and corners of the element produced TBF_LAG iterations ago
			// (the row one down was published long since)
			if (tid == 0 && prev.e >= 0) tb_flag_publish(fz.done + prev.e, fz.epoch);
			if (behind == TBF_LAG) {
				// lag_nx becomes the element whose alpha edge is finished now
				lag_el = lag_nx;
				tb_alpha_load(fz, lag_el, out, esz, nrows, tid, seen, areg);
				if (tb_walk_next(fz, wa, (int)gridDim.x)) lag_nx = tb_walk_elem(wa);
				seen = tb_alpha_peek(fz, lag_nx);
			} else {
				lag_el.e = -1;
			}
		}
		if (tid == 0 && has_next) {
			// the next element of this block, by bulk tensor copies
			const long long en = nxt.e;
			tb_mbar_expect(&bars[buf ^ 1], ebytes * (HAS_BASE ? 2u : 1u) + cbytes);
			tb_tma_element(fb0 + (size_t)(buf ^ 1) * ebuf, maps.in, en * nstride, nrows, &bars[buf ^ 1]);
			if (HAS_BASE) {
				tb_tma_element(bb0 + (size_t)(buf ^ 1) * ebuf, maps.b0, en * nstride, nrows, &bars[buf ^ 1]);
			}
			tb_bulk_1d(scc0 + (size_t)(buf ^ 1) * TBF_NC * NN,
				ha.colc + (size_t)en * TBF_NC * NN, cbytes, &bars[buf ^ 1]);
		}
		if (FUSE) tb_alpha_finish(lag_el, out, esz, nrows, aprev, tid, areg);

		const size_t ebase = (size_t)e * esz;
		const double * cc = scc0 + (size_t)buf * TBF_NC * NN + i * NP;
		const double dInvDA = __ldg(ha.inv_da + e);
		const double dInvDB = __ldg(ha.inv_db + e);
		const double nus = ha.scale_nu ? __ldg(ha.nu_scale + e) : 1.0;
		// HorizontalDynamicsFEM.cpp:1970-1975, 2226-2233
		const double dNuS = ha.scale_nu ? ha.nu_scalar * nus : ha.nu_scalar;
		const double dNuD = ha.scale_nu ? ha.nu_div * nus : ha.nu_div;
		const double dNuV = ha.scale_nu ? ha.nu_vort * nus : ha.nu_vort;

		double cA0[4], cA1[4], cB0[4], cB1[4], cJ[4], cIJ[4], cJ2[4];
		tb_ld4(cc + TBF_A0 * NN, cA0);
		tb_ld4(cc + TBF_A1 * NN, cA1);
		tb_ld4(cc + TBF_B0 * NN, cB0);
		tb_ld4(cc + TBF_B1 * NN, cB1);
		tb_ld4(cc + TBF_JAC * NN, cJ);
		tb_ld4(cc + TBF_INVJAC * NN, cIJ);
		tb_ld4(cc + TBF_J2D * NN, cJ2);

		for (int k0 = 0; k0 <= L; k0 += TBF_KB) {
			const int k = k0 + kq;
			const bool wact = (k <= L);             // interface row exists
			const bool lact = (k < L);              // level rows exist
			const int kw = wact ? k : L;
			const int kc = lact ? k : (L - 1);
			const int tp = kc & 1, tpw = kw & 1;

			// ---- first round: gradients -> fluxes ------------------------------------
			double gbP[4], gbR[4], gbW[4], jub[4], u[4], v[4];
			{
				double x[4], da[4], gaP[4], gaR[4], gaW[4], jua[4];
				// rho-theta
				tb_ld4s(fb + (size_t)(rP + kc) * NN, (rP + kc) & 7, i, x);
				tb_cross_sum4s(fb + (size_t)(rP + kc) * NN, (rP + kc) & 7, dxI, da);
#pragma unroll
				for (int j = 0; j < 4; j++) {
					const double dDa = da[j] * dInvDA;
					const double dDb = TB_ROW_DX(x, j) * dInvDB;
					gaP[j] = cJ[j] * (cA0[j] * dDa + cA1[j] * dDb);
					gbP[j] = cJ[j] * (cB0[j] * dDa + cB1[j] * dDb);
				}
				// rho
				tb_ld4s(fb + (size_t)(rR + kc) * NN, (rR + kc) & 7, i, x);
				tb_cross_sum4s(fb + (size_t)(rR + kc) * NN, (rR + kc) & 7, dxI, da);
#pragma unroll
				for (int j = 0; j < 4; j++) {
					const double dDa = da[j] * dInvDA;
					const double dDb = TB_ROW_DX(x, j) * dInvDB;
					gaR[j] = cJ[j] * (cA0[j] * dDa + cA1[j] * dDb);
					gbR[j] = cJ[j] * (cB0[j] * dDa + cB1[j] * dDb);
				}
				// w (interfaces; JacobianREdge = Jacobian for this metric)
				tb_ld4s(fb + (size_t)(rW + kw) * NN, (rW + kw) & 7, i, x);
				tb_cross_sum4s(fb + (size_t)(rW + kw) * NN, (rW + kw) & 7, dxI, da);
#pragma unroll
				for (int j = 0; j < 4; j++) {
					const double dDa = da[j] * dInvDA;
					const double dDb = TB_ROW_DX(x, j) * dInvDB;
					gaW[j] = cJ[j] * (cA0[j] * dDa + cA1[j] * dDb);
					gbW[j] = cJ[j] * (cB0[j] * dDa + cB1[j] * dDb);
				}
				// velocities: J2D * contravariant components (GridPatchCSGLL.cpp:1207-1218)
				tb_ld4s(fb + (size_t)(rU + kc) * NN, (rU + kc) & 7, i, u);
				tb_ld4s(fb + (size_t)(rV + kc) * NN, (rV + kc) & 7, i, v);
#pragma unroll
				for (int j = 0; j < 4; j++) {
					jua[j] = cJ2[j] * (+cA0[j] * u[j] + cA1[j] * v[j]);
					jub[j] = cJ2[j] * (+cB0[j] * u[j] + cB1[j] * v[j]);
				}
				if (lact) {
					tb_st4s(tGP + (size_t)k * NN, tp, i, gaP);
					tb_st4s(tGR + (size_t)k * NN, tp, i, gaR);
					tb_st4s(tJU + (size_t)k * NN, tp, i, jua);
				}
				if (wact) tb_st4s(tGW + (size_t)k * NN, tpw, i, gaW);
			}
			__syncwarp();

			// ---- scalar Laplacians (:2126-2165) ----------------------------------------
			{
				double ua[4], o[4];
				tb_cross_sum4s(tGP + (size_t)kc * NN, tp, stI, ua);
				if (HAS_BASE) tb_ld4s(bb + (size_t)(rP + kc) * NN, (rP + kc) & 7, i, o);
#pragma unroll
				for (int j = 0; j < 4; j++) {
					const double dUpdateA = ua[j] * dInvDA;
					const double dUpdateB = TB_ROW_ST(gbP, j) * dInvDB;
					const double b = HAS_BASE ? o[j] : 0.0;
					o[j] = b - ha.dt * cIJ[j] * dNuS * (dUpdateA + dUpdateB);
				}
				if (lact) tb_out4<FUSE>(cur, carry, out + ebase, esz, rP + k, i, o);
				tb_cross_sum4s(tGR + (size_t)kc * NN, tp, stI, ua);
				if (HAS_BASE) tb_ld4s(bb + (size_t)(rR + kc) * NN, (rR + kc) & 7, i, o);
#pragma unroll
				for (int j = 0; j < 4; j++) {
					const double dUpdateA = ua[j] * dInvDA;
					const double dUpdateB = TB_ROW_ST(gbR, j) * dInvDB;
					const double b = HAS_BASE ? o[j] : 0.0;
					o[j] = b - ha.dt * cIJ[j] * dNuS * (dUpdateA + dUpdateB);
				}
				if (lact) tb_out4<FUSE>(cur, carry, out + ebase, esz, rR + k, i, o);
				tb_cross_sum4s(tGW + (size_t)kw * NN, tpw, stI, ua);
				if (HAS_BASE) tb_ld4s(bb + (size_t)(rW + kw) * NN, (rW + kw) & 7, i, o);
#pragma unroll
				for (int j = 0; j < 4; j++) {
					const double dUpdateA = ua[j] * dInvDA;
					const double dUpdateB = TB_ROW_ST(gbW, j) * dInvDB;
					const double b = HAS_BASE ? o[j] : 0.0;
					o[j] = b - ha.dt * cIJ[j] * dNuS * (dUpdateA + dUpdateB);
				}
				if (wact) tb_out4<FUSE>(cur, carry, out + ebase, esz, rW + k, i, o);
			}

			// ---- curl and divergence (GridPatchCSGLL.cpp:1262-1299) --------------------
			double dv[4], cl[4];
			{
				double daUb[4], daJUa[4];
				tb_cross_sum4s(fb + (size_t)(rV + kc) * NN, (rV + kc) & 7, dxI, daUb);
				tb_cross_sum4s(tJU + (size_t)kc * NN, tp, dxI, daJUa);
#pragma unroll
				for (int j = 0; j < 4; j++) {
					const double dDaUb = daUb[j] * dInvDA;
					const double dDbUa = TB_ROW_DX(u, j) * dInvDB;
					const double dDaJUa = daJUa[j] * dInvDA;
					const double dDbJUb = TB_ROW_DX(jub, j) * dInvDB;
					const double dInvJacobian2D = 1.0 / cJ2[j];
					dv[j] = (dDaJUa + dDbJUb) * dInvJacobian2D;
					cl[j] = (dDaUb - dDbUa) * dInvJacobian2D;
				}
				if (lact) {
					tb_st4s(tDV + (size_t)k * NN, tp, i, dv);
					tb_st4s(tCL + (size_t)k * NN, tp, i, cl);
				}
			}
			__syncwarp();

			// ---- vector Laplacian (:2366-2407) -----------------------------------------
			{
				double daDiv[4], daCurl[4], oU[4], oV[4];
				tb_cross_sum4s(tDV + (size_t)kc * NN, tp, stI, daDiv);
				tb_cross_sum4s(tCL + (size_t)kc * NN, tp, stI, daCurl);
				if (HAS_BASE) {
					tb_ld4s(bb + (size_t)(rU + kc) * NN, (rU + kc) & 7, i, oU);
					tb_ld4s(bb + (size_t)(rV + kc) * NN, (rV + kc) & 7, i, oV);
				}
#pragma unroll
				for (int j = 0; j < 4; j++) {
					const double dDaDiv = -daDiv[j] * dInvDA;
					const double dDbDiv = -TB_ROW_ST(dv, j) * dInvDB;
					const double dDaCurl = -daCurl[j] * dInvDA;
					const double dDbCurl = -TB_ROW_ST(cl, j) * dInvDB;
					const double dUpdateUa =
						+dNuD * dDaDiv
						- dNuV * cJ2[j] * (cB0[j] * dDaCurl + cB1[j] * dDbCurl);
					const double dUpdateUb =
						+dNuD * dDbDiv
						+ dNuV * cJ2[j] * (cA0[j] * dDaCurl + cA1[j] * dDbCurl);
					const double bu = HAS_BASE ? oU[j] : 0.0;
					const double bv = HAS_BASE ? oV[j] : 0.0;
					oU[j] = bu - ha.dt * dUpdateUa;
					oV[j] = ha.xz ? bv : (bv - ha.dt * dUpdateUb);
				}
				if (lact) {
					tb_out4<FUSE>(cur, carry, out + ebase, esz, rU + k, i, oU);
					tb_out4<FUSE>(cur, carry, out + ebase, esz, rV + k, i, oV);
				}
			}
		}
		if (FUSE && behind < TBF_LAG) behind++;
		prev = cur;
		cur = nxt;
		valid = has_next;
	}
	if (FUSE) {
		// the last element of this block, then the alpha edges still pending
		// (oldest first: the corner pair average passes from one to the next)
		__syncthreads();
		if (tid == 0 && prev.e >= 0) tb_flag_publish(fz.done + prev.e, fz.epoch);
		for (int q = 0; q < behind; q++) {
			tb_fuse_alpha(fz, lag_nx, out, esz, nrows, aprev, tid);
			if (tb_walk_next(fz, wa, (int)gridDim.x)) lag_nx = tb_walk_elem(wa);
		}
	}
}

///////////////////////////////////////////////////////////////////////////////
// Column constants from the 2-D metric, the topography derivatives and the
// layer depth dxr = ztop - zs (GridPatchCSGLL.cpp:344-553).

__global__ void k_fast_colc(
	long long ncol, DevGeom g, double grav, double * colc
) {
	const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (idx >= ncol) return;
	const long long e = idx / 16;
	const int n = (int)(idx % 16);
	const double a0 = g.a0[idx], a1 = g.a1[idx], b1 = g.b1[idx];
	const double j2d = g.j2d[idx];
	const double dazs = (g.tda != 0) ? g.tda[idx] : 0.0;
	const double dbzs = (g.tdb != 0) ? g.tdb[idx] : 0.0;
	const double dxr = g.ztop - g.zs[idx];
	const double jac = dxr * j2d;
	const double A2 = -(a0 * dazs + a1 * dbzs) / dxr;
	const double B2 = -(a1 * dazs + b1 * dbzs) / dxr;
	double * c = colc + (size_t)e * TBF_NC * 16 + n;
	c[TBF_A0 * 16] = a0;
	c[TBF_A1 * 16] = a1;
	c[TBF_B1 * 16] = b1;
	c[TBF_JAC * 16] = jac;
	c[TBF_INVJAC * 16] = 1.0 / jac;
	c[TBF_FJ * 16] = g.f[idx] * j2d;
	c[TBF_A2 * 16] = A2;
	c[TBF_B2 * 16] = B2;
	c[TBF_X0 * 16] = 1.0 / (dxr * dxr);
	c[TBF_X2 * 16] = -(A2 * dazs + B2 * dbzs) / dxr;
	c[TBF_GDA * 16] = grav * dazs;
	c[TBF_GDB * 16] = grav * dbzs;
	c[TBF_DXR * 16] = dxr;
	c[TBF_J2D * 16] = j2d;
	c[TBF_B0 * 16] = g.b0[idx];
}

// Largest deviation of the metric rebuilt from the column constants from the
// uploaded reference arrays (relative to the natural scale of each entry):
// one value per thread in errs[], the host takes the maximum.
__device__ __forceinline__ void tb_fast_dev(double ref, double got, double scale, double & worst) {
	const double d = fabs(ref - got);
	const double s = fabs(scale);
	const double rel = (s > 0.0) ? d / s : d;
	if (!(rel == rel)) worst = 1.0e300;
	else if (rel > worst) worst = rel;
}

__global__ void k_fast_verify(
	DevLayout lay, DevGeom g, double grav, const double * colc, const double * lev,
	double * errs
) {
	const int L = lay.nlev;
	const long long ncol = lay.nelem * 16;
	double worst = 0.0;
	for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	     idx < ncol; idx += (long long)gridDim.x * blockDim.x
	) {
		const long long e = idx / 16;
		const int n = (int)(idx % 16);
		const double * c = colc + (size_t)e * TBF_NC * 16 + n;
		const double a0 = c[TBF_A0 * 16], a1 = c[TBF_A1 * 16], b1 = c[TBF_B1 * 16];
		const double mag = fabs(a0) + fabs(b1);      // scale of the horizontal metric
		const double dxr = c[TBF_DXR * 16];
		for (int k = 0; k <= L; k++) {
			const size_t oe = ((size_t)e * (L + 1) + k) * 16 + n;
			const double se = lev[(size_t)k * TBF_LW + TBF_SE];
			const double x2e = c[TBF_X0 * 16] + se * se * c[TBF_X2 * 16];
			const double mix = sqrt(mag * x2e);
			tb_fast_dev(g.jace[oe], c[TBF_JAC * 16], g.jace[oe], worst);
			tb_fast_dev(g.cae[0][oe], a0, mag, worst);
			tb_fast_dev(g.cae[1][oe], a1, mag, worst);
			tb_fast_dev(g.cbe[1][oe], b1, mag, worst);
			tb_fast_dev(g.cae[2][oe], se * c[TBF_A2 * 16], mix, worst);
			tb_fast_dev(g.cbe[2][oe], se * c[TBF_B2 * 16], mix, worst);
			tb_fast_dev(g.cxe[0][oe], se * c[TBF_A2 * 16], mix, worst);
			tb_fast_dev(g.cxe[1][oe], se * c[TBF_B2 * 16], mix, worst);
			tb_fast_dev(g.cxe[2][oe], x2e, x2e, worst);
			tb_fast_dev(g.dre[2][oe], dxr, dxr, worst);
			if (k < L) {
				const size_t on = ((size_t)e * L + k) * 16 + n;
				const double s = lev[(size_t)k * TBF_LW + TBF_SN];
				const double x2 = c[TBF_X0 * 16] + s * s * c[TBF_X2 * 16];
				const double mixn = sqrt(mag * x2);
				tb_fast_dev(g.jac[on], c[TBF_JAC * 16], g.jac[on], worst);
				tb_fast_dev(g.ca[0][on], a0, mag, worst);
				tb_fast_dev(g.ca[1][on], a1, mag, worst);
				tb_fast_dev(g.cb[0][on], a1, mag, worst);
				tb_fast_dev(g.cb[1][on], b1, mag, worst);
				tb_fast_dev(g.ca[2][on], s * c[TBF_A2 * 16], mixn, worst);
				tb_fast_dev(g.cb[2][on], s * c[TBF_B2 * 16], mixn, worst);
				tb_fast_dev(g.cx[0][on], s * c[TBF_A2 * 16], mixn, worst);
				tb_fast_dev(g.cx[1][on], s * c[TBF_B2 * 16], mixn, worst);
				tb_fast_dev(g.cx[2][on], x2, x2, worst);
				tb_fast_dev(g.dr[2][on], dxr, dxr, worst);
				// g * DerivR[0,1]
				tb_fast_dev(grav * g.dr[0][on], s * c[TBF_GDA * 16], grav * g.dr[0][on], worst);
				tb_fast_dev(grav * g.dr[1][on], s * c[TBF_GDB * 16], grav * g.dr[1][on], worst);
			}
		}
	}
	errs[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = worst;
}

#endif
