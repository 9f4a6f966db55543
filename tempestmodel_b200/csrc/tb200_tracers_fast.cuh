// Tracer transport on the column-constant path (vertical order 1,
// terrain-following metric: tb200_fast.cuh), np = 4.
//
//   HorizontalDynamicsFEM::StepNonhydrostaticPrimitive, tracer part
//       HorizontalDynamicsFEM.cpp:1050-1077 (fluxes), :1531-1553 (update)
//   HorizontalDynamicsFEM::FilterNegativeTracers      :213-317
//   HorizontalDynamicsFEM::ApplyScalarHyperdiffusion  :1867-2203 (tracer rows)
//
// The state rows of an element go through k_nh_stage_pipe / k_hyper_pipe; the
// tracer rows that follow them in the element block go through the kernels
// here.  They stream: every tracer value is read once and written once per
// pass, the mass fluxes J u^alpha, J u^beta are rebuilt from the element's u, v,
// w rows (3 L + 1 rows against ntr * L tracer rows per source) and the column
// constants, in the arithmetic of the stage kernel.
//
// Thread = (level k, element row i) owning the nodes (i, 0..3): 32 contiguous
// bytes of every 128-byte row.  Sums over j (beta) are register-only; sums over
// i (alpha) and the 16-node sums of the positivity filter take the other three
// rows of the level from the neighbouring lanes by shuffles - no shared memory,
// no block barrier.  Summation order follows the reference.
#ifndef TB200_TRACERS_FAST_CUH
#define TB200_TRACERS_FAST_CUH

#include "tb200_fast.cuh"

#define TBT_THREADS 128
#define TBT_KB (TBT_THREADS / 4)

// the four values of row s (0..3) of my level, from the lane that owns it
__device__ __forceinline__ void tb_row_from(const double (&mine)[4], int s, double (&o)[4]) {
#pragma unroll
	for (int j = 0; j < 4; j++) {
		o[j] = __shfl_sync(0xffffffffu, mine[j], s, 4);
	}
}

// alpha-direction sum for my four nodes: o[j] = sum_s x(s, j) * c[s]
__device__ __forceinline__ void tb_cross_sum4w(const double (&x)[4], const double (&c)[4], double (&o)[4]) {
#pragma unroll
	for (int j = 0; j < 4; j++) o[j] = 0.0;
#pragma unroll
	for (int s = 0; s < 4; s++) {
		double r[4];
		tb_row_from(x, s, r);
#pragma unroll
		for (int j = 0; j < 4; j++) o[j] += r[j] * c[s];
	}
}

// HorizontalDynamicsFEM::FilterNegativeTracers on one (element, tracer, level):
// v = my four values, a = the element areas of my four nodes.  The pointwise
// masses v * area are formed by the lane that owns the node and summed by every
// lane in node order.  "value >= 0" is tested on the mass: the areas are
// positive, so the two differ only where a negative value underflows to a mass
// of -0, which adds nothing to either sum.
__device__ __forceinline__ void tb_filter_level(double (&v)[4], const double (&a)[4]) {
	double pm[4];
#pragma unroll
	for (int j = 0; j < 4; j++) pm[j] = v[j] * a[j];
	double dTotalMass = 0.0;
	double dNonNegativeMass = 0.0;
#pragma unroll
	for (int s = 0; s < 4; s++) {
		double r[4];
		tb_row_from(pm, s, r);
#pragma unroll
		for (int j = 0; j < 4; j++) {
			const double dPointwiseMass = r[j];
			dTotalMass += dPointwiseMass;
			if (dPointwiseMass >= 0.0) {
				dNonNegativeMass += dPointwiseMass;
			}
		}
	}
	const double dR = dTotalMass / dNonNegativeMass;
#pragma unroll
	for (int j = 0; j < 4; j++) {
		v[j] = (v[j] > 0.0) ? v[j] * dR : 0.0;
	}
}

struct TracerFastArgs {
	const double * colc;      // column constants [e][TBF_NC][16]
	const double * lev;       // operator windows [L+1][TBF_LW]
	const double * inv_da;
	const double * inv_db;
	const double * area;      // element areas [e][L][16] (positivity filter), or 0
	double dt;
};

// Stage base (Grid::CopyData / LinearCombineData) + horizontal transport + the
// element-wise positivity filter of every tracer of an element.
__global__ void __launch_bounds__(TBT_THREADS, 4)
k_tracer_stage(
	DevLayout lay, DevTables t, TracerFastArgs ta, StageBase sb,
	const double * __restrict__ in, double * out, ElemList el
) {
	const int NN = 16;
	const int L = lay.nlev;
	const long long e = tb_elem(el, blockIdx.x);
	const int kq = threadIdx.x >> 2;
	const int i = threadIdx.x & 3;

	const size_t ebase = (size_t)e * lay.nrows * NN;
	const double * inU = in + ebase + (size_t)lay.rowoff[0] * NN;
	const double * inV = in + ebase + (size_t)lay.rowoff[1] * NN;
	const double * inW = in + ebase + (size_t)lay.rowoff[3] * NN;
	const double * cc = ta.colc + (size_t)e * TBF_NC * NN + i * 4;
	const double dInvDA = __ldg(ta.inv_da + e);
	const double dInvDB = __ldg(ta.inv_db + e);
	const double dt = ta.dt;

	double stI[4];
#pragma unroll
	for (int s = 0; s < 4; s++) stI[s] = t.st[i * 4 + s];

	double cA0[4], cA1[4], cB1[4], cJ[4], cIJ[4], cA2[4], cB2[4];
	tb_ld4g(cc + TBF_A0 * NN, cA0);
	tb_ld4g(cc + TBF_A1 * NN, cA1);
	tb_ld4g(cc + TBF_B1 * NN, cB1);
	tb_ld4g(cc + TBF_JAC * NN, cJ);
	tb_ld4g(cc + TBF_INVJAC * NN, cIJ);
	tb_ld4g(cc + TBF_A2 * NN, cA2);
	tb_ld4g(cc + TBF_B2 * NN, cB2);

	for (int k0 = 0; k0 < L; k0 += TBT_KB) {
		const int k = k0 + kq;
		const bool active = (k < L);
		const int kc = active ? k : (L - 1);
		const size_t o4 = (size_t)kc * NN + i * 4;
		const double * lv = ta.lev + (size_t)kc * TBF_LW;
		const double sn = __ldg(lv + TBF_SN);
		const double cw0 = __ldg(lv + TBF_CW + 0), cw1 = __ldg(lv + TBF_CW + 1);

		double u[4], v[4], w0[4], wp[4];
		tb_ld4(inU + o4, u);
		tb_ld4(inV + o4, v);
		tb_ld4(inW + o4, w0);
		tb_ld4(inW + o4 + NN, wp);

		// mass fluxes per unit tracer density (:916-929, 1050-1077)
		double fa[4], fb[4];
#pragma unroll
		for (int j = 0; j < 4; j++) {
			double x = 0.0;
			x += cw0 * w0[j];
			x += cw1 * wp[j];
			const double m2 = sn * cA2[j], m4 = sn * cB2[j];
			const double conUa = cA0[j] * u[j] + cA1[j] * v[j] + m2 * x;
			const double conUb = cA1[j] * u[j] + cB1[j] * v[j] + m4 * x;
			fa[j] = cJ[j] * conUa;
			fb[j] = cJ[j] * conUb;
		}

		double area[4] = {1.0, 1.0, 1.0, 1.0};
		if (ta.area != 0) {
			tb_ld4g(ta.area + ((size_t)e * L + kc) * NN + i * 4, area);
		}

		for (int c = 0; c < lay.ntr; c++) {
			const size_t oT = ebase + (size_t)(lay.troff + c * L) * NN + o4;
			double q[4], b[4];
			tb_ld4(in + oT, q);
			tb_stage_base4(sb, out, oT, b);
			double fA[4], fB[4], dDa[4];
#pragma unroll
			for (int j = 0; j < 4; j++) {
				fA[j] = fa[j] * q[j];
				fB[j] = fb[j] * q[j];
			}
			// :1536-1545: dDaTracerFluxA -= flux(s, j) * stiffness(i, s)
#pragma unroll
			for (int j = 0; j < 4; j++) dDa[j] = 0.0;
#pragma unroll
			for (int s = 0; s < 4; s++) {
				double r[4];
				tb_row_from(fA, s, r);
#pragma unroll
				for (int j = 0; j < 4; j++) dDa[j] -= r[j] * stI[s];
			}
#pragma unroll
			for (int j = 0; j < 4; j++) {
				double dDb = 0.0;
#pragma unroll
				for (int s = 0; s < 4; s++) dDb -= fB[s] * t.st[j * 4 + s];
				const double dDaTracerFluxA = dDa[j] * dInvDA;
				const double dDbTracerFluxB = dDb * dInvDB;
				b[j] = b[j] - dt * cIJ[j] * (dDaTracerFluxA + dDbTracerFluxB);
			}
			if (ta.area != 0) tb_filter_level(b, area);
			if (active) tb_st4(out + oT, b);
		}
	}
}

///////////////////////////////////////////////////////////////////////////////
// Pipelined variant of k_tracer_stage: persistent blocks, everything an element
// needs - its u, v, w rows, its tracer rows, the tracer rows of the stage-base
// sources, column constants and element areas - arrives by bulk asynchronous
// copies (cp.async.bulk, contiguous runs of 128-byte rows, completion on an
// mbarrier) one element ahead, so that a whole element per block is always in
// flight and no warp waits on a global load.  Same arithmetic as k_tracer_stage.

struct TracerBase {
	const double * src[2];    // instances the base is combined from (Grid::LinearCombineData order)
	double coeff[2];
	int nsrc;                 // 0: base = the input tracers (CopyData of the input instance)
	int first_is_dst;         // src[0] is the update instance itself (scaled by coeff[0])
	int exact;                // one source taken as it is (the update instance was pre-filled)
};

__host__ __device__ inline size_t tb_tracer_pipe_buffer_doubles(int L, int ntr, int nsrc) {
	// u, v [2 L], w [L + 1], tracers [ntr L], base [nsrc][ntr L], column constants, areas [L]
	return ((size_t)(3 * L + 1) + (size_t)ntr * L * (1 + nsrc) + TBF_NC + L) * 16;
}
__host__ __device__ inline size_t tb_tracer_pipe_smem_doubles(int L, int ntr, int nsrc) {
	return 2 * tb_tracer_pipe_buffer_doubles(L, ntr, nsrc) + 16 + 4;
}

__global__ void __launch_bounds__(TBT_THREADS, 2)
k_tracer_stage_pipe(
	DevLayout lay, DevTables t, TracerFastArgs ta, TracerBase tbse,
	const double * __restrict__ in, double * out, ElemList el
) {
	const int NN = 16;
	const int L = lay.nlev;
	const int ntr = lay.ntr;
	const int nsrc = tbse.nsrc;
	TB_DYN_SMEM(double, sm_raw);
	double * sm = tb_smem_aligned(sm_raw);
	const size_t bufd = tb_tracer_pipe_buffer_doubles(L, ntr, nsrc);
	tb_mbar_t * bars = reinterpret_cast<tb_mbar_t *>(sm + 2 * bufd);

	const int tid = threadIdx.x;
	const int kq = tid >> 2;
	const int i = tid & 3;
	const size_t esz = (size_t)lay.nrows * NN;
	const size_t trows = (size_t)ntr * L * NN;        // doubles of the tracer rows of an element
	const unsigned bytes = (unsigned)(bufd * sizeof(double));

	double stI[4];
#pragma unroll
	for (int s = 0; s < 4; s++) stI[s] = t.st[i * 4 + s];

	long long w = blockIdx.x;
	if (w >= el.n) return;
	if (tid == 0) {
		tb_mbar_init(&bars[0], 1);
		tb_mbar_init(&bars[1], 1);
		tb_mbar_fence_init();
	}
	__syncthreads();

	// buffer layout (doubles): u [L], v [L], w [L+1], tracers, base 0, base 1, colc, area
	const size_t oU = 0, oV = (size_t)L * NN, oW = (size_t)2 * L * NN;
	const size_t oT = oW + (size_t)(L + 1) * NN;
	const size_t oB = oT + trows;
	const size_t oC = oB + (size_t)nsrc * trows;
	const size_t oA = oC + (size_t)TBF_NC * NN;

	auto fetch = [&](long long e, int buf) {
		double * d = sm + (size_t)buf * bufd;
		const double * ge = in + (size_t)e * esz;
		tb_mbar_expect(&bars[buf], bytes);
		tb_bulk_1d(d + oU, ge + (size_t)lay.rowoff[0] * NN, (unsigned)(L * NN * 8), &bars[buf]);
		tb_bulk_1d(d + oV, ge + (size_t)lay.rowoff[1] * NN, (unsigned)(L * NN * 8), &bars[buf]);
		tb_bulk_1d(d + oW, ge + (size_t)lay.rowoff[3] * NN, (unsigned)((L + 1) * NN * 8), &bars[buf]);
		tb_bulk_1d(d + oT, ge + (size_t)lay.troff * NN, (unsigned)(trows * 8), &bars[buf]);
		for (int m = 0; m < nsrc; m++) {
			tb_bulk_1d(d + oB + (size_t)m * trows,
				tbse.src[m] + (size_t)e * esz + (size_t)lay.troff * NN, (unsigned)(trows * 8), &bars[buf]);
		}
		tb_bulk_1d(d + oC, ta.colc + (size_t)e * TBF_NC * NN, (unsigned)(TBF_NC * NN * 8), &bars[buf]);
		tb_bulk_1d(d + oA, ta.area + (size_t)e * L * NN, (unsigned)(L * NN * 8), &bars[buf]);
	};

	long long e = tb_elem(el, w);
	if (tid == 0) fetch(e, 0);

	for (int it = 0; ; it++) {
		w += gridDim.x;
		const bool has_next = (w < el.n);
		const long long en = has_next ? tb_elem(el, w) : 0;
		const int buf = it & 1;
		// this element has landed, and every thread is done with the other buffer
		tb_mbar_wait(&bars[buf], (unsigned)(it >> 1) & 1u);
		__syncthreads();
		if (tid == 0 && has_next) fetch(en, buf ^ 1);

		const double * d = sm + (size_t)buf * bufd;
		const double * cc = d + oC + i * 4;
		const double dInvDA = __ldg(ta.inv_da + e);
		const double dInvDB = __ldg(ta.inv_db + e);
		const double dt = ta.dt;
		double cA0[4], cA1[4], cB1[4], cJ[4], cIJ[4], cA2[4], cB2[4];
		tb_ld4(cc + TBF_A0 * NN, cA0);
		tb_ld4(cc + TBF_A1 * NN, cA1);
		tb_ld4(cc + TBF_B1 * NN, cB1);
		tb_ld4(cc + TBF_JAC * NN, cJ);
		tb_ld4(cc + TBF_INVJAC * NN, cIJ);
		tb_ld4(cc + TBF_A2 * NN, cA2);
		tb_ld4(cc + TBF_B2 * NN, cB2);
		double * oute = out + (size_t)e * esz + (size_t)lay.troff * NN;

		for (int k0 = 0; k0 < L; k0 += TBT_KB) {
			const int k = k0 + kq;
			const bool active = (k < L);
			const int kc = active ? k : (L - 1);
			const size_t o4 = (size_t)kc * NN + i * 4;
			const double * lv = ta.lev + (size_t)kc * TBF_LW;
			const double sn = __ldg(lv + TBF_SN);
			const double cw0 = __ldg(lv + TBF_CW + 0), cw1 = __ldg(lv + TBF_CW + 1);

			double u[4], v[4], w0[4], wp[4];
			tb_ld4(d + oU + o4, u);
			tb_ld4(d + oV + o4, v);
			tb_ld4(d + oW + o4, w0);
			tb_ld4(d + oW + o4 + NN, wp);

			// mass fluxes per unit tracer density (:916-929, 1050-1077)
			double fa[4], fb[4];
#pragma unroll
			for (int j = 0; j < 4; j++) {
				double x = 0.0;
				x += cw0 * w0[j];
				x += cw1 * wp[j];
				const double m2 = sn * cA2[j], m4 = sn * cB2[j];
				const double conUa = cA0[j] * u[j] + cA1[j] * v[j] + m2 * x;
				const double conUb = cA1[j] * u[j] + cB1[j] * v[j] + m4 * x;
				fa[j] = cJ[j] * conUa;
				fb[j] = cJ[j] * conUb;
			}
			double area[4];
			tb_ld4(d + oA + o4, area);

			for (int c = 0; c < ntr; c++) {
				const size_t oc = (size_t)c * L * NN + o4;
				double q[4], b[4];
				tb_ld4(d + oT + oc, q);
				// stage base in Grid::LinearCombineData's order (k_lincomb)
				if (nsrc == 0) {
#pragma unroll
					for (int j = 0; j < 4; j++) b[j] = q[j];
				} else if (tbse.exact) {
					tb_ld4(d + oB + oc, b);
				} else {
					int m0 = 0;
					if (tbse.first_is_dst) {
						double s0[4];
						tb_ld4(d + oB + oc, s0);
#pragma unroll
						for (int j = 0; j < 4; j++) b[j] = s0[j] * tbse.coeff[0];
						m0 = 1;
					} else {
#pragma unroll
						for (int j = 0; j < 4; j++) b[j] = 0.0;
					}
					for (int m = m0; m < nsrc; m++) {
						double s1[4];
						tb_ld4(d + oB + (size_t)m * trows + oc, s1);
						const double cf = tbse.coeff[m];
#pragma unroll
						for (int j = 0; j < 4; j++) b[j] += s1[j] * cf;
					}
				}
				double fA[4], fB[4], dDa[4];
#pragma unroll
				for (int j = 0; j < 4; j++) {
					fA[j] = fa[j] * q[j];
					fB[j] = fb[j] * q[j];
				}
				// :1536-1545: dDaTracerFluxA -= flux(s, j) * stiffness(i, s)
#pragma unroll
				for (int j = 0; j < 4; j++) dDa[j] = 0.0;
#pragma unroll
				for (int s = 0; s < 4; s++) {
					double r[4];
					tb_row_from(fA, s, r);
#pragma unroll
					for (int j = 0; j < 4; j++) dDa[j] -= r[j] * stI[s];
				}
#pragma unroll
				for (int j = 0; j < 4; j++) {
					double dDb = 0.0;
#pragma unroll
					for (int s = 0; s < 4; s++) dDb -= fB[s] * t.st[j * 4 + s];
					const double dDaTracerFluxA = dDa[j] * dInvDA;
					const double dDbTracerFluxB = dDb * dInvDB;
					b[j] = b[j] - dt * cIJ[j] * (dDaTracerFluxA + dDbTracerFluxB);
				}
				tb_filter_level(b, area);
				if (active) tb_st4(oute + oc, b);
			}
		}
		if (!has_next) break;
		e = en;
	}
}

// out = (HAS_BASE ? base : 0) - dt nu L(fld) on the tracer rows of an element,
// optionally followed by the element-wise positivity filter
// (HorizontalDynamicsFEM.cpp:2073-2165; the order-4 sequence of
// StepAfterSubCycle, :2687-2713, calls it twice around a DSS).
template <bool HAS_BASE>
__global__ void __launch_bounds__(TBT_THREADS, 4)
k_tracer_hyper(
	DevLayout lay, DevTables t, HyperFastArgs ha, const double * area_node,
	const double * __restrict__ fld, const double * base, double * out, ElemList el
) {
	const int NN = 16;
	const int L = lay.nlev;
	const int nr = lay.ntr * L;
	const long long e = tb_elem(el, blockIdx.x);
	const int rq = threadIdx.x >> 2;
	const int i = threadIdx.x & 3;

	const size_t tbase = ((size_t)e * lay.nrows + lay.troff) * NN;
	const double * cc = ha.colc + (size_t)e * TBF_NC * NN + i * 4;
	const double dInvDA = __ldg(ha.inv_da + e);
	const double dInvDB = __ldg(ha.inv_db + e);
	const double nus = ha.scale_nu ? __ldg(ha.nu_scale + e) : 1.0;
	const double dNuS = ha.scale_nu ? ha.nu_scalar * nus : ha.nu_scalar;

	double dxI[4], stI[4];
#pragma unroll
	for (int s = 0; s < 4; s++) {
		dxI[s] = t.dx[s * 4 + i];
		stI[s] = t.st[i * 4 + s];
	}
	double cA0[4], cA1[4], cB0[4], cB1[4], cJ[4], cIJ[4];
	tb_ld4g(cc + TBF_A0 * NN, cA0);
	tb_ld4g(cc + TBF_A1 * NN, cA1);
	tb_ld4g(cc + TBF_B0 * NN, cB0);
	tb_ld4g(cc + TBF_B1 * NN, cB1);
	tb_ld4g(cc + TBF_JAC * NN, cJ);
	tb_ld4g(cc + TBF_INVJAC * NN, cIJ);

	for (int r0 = 0; r0 < nr; r0 += TBT_KB) {
		const int r = r0 + rq;
		const bool active = (r < nr);
		const int rc = active ? r : (nr - 1);
		const size_t o4 = tbase + (size_t)rc * NN + i * 4;

		double x[4], da[4], ga[4], gb[4], ua[4], o[4];
		tb_ld4(fld + o4, x);
		if (HAS_BASE) tb_ld4(base + o4, o);
		// pointwise gradient (:2073-2107)
		tb_cross_sum4w(x, dxI, da);
#pragma unroll
		for (int j = 0; j < 4; j++) {
			const double dDa = da[j] * dInvDA;
			const double dDb = TB_ROW_DX(x, j) * dInvDB;
			ga[j] = cJ[j] * (cA0[j] * dDa + cA1[j] * dDb);
			gb[j] = cJ[j] * (cB0[j] * dDa + cB1[j] * dDb);
		}
		// integral term (:2126-2165)
		tb_cross_sum4w(ga, stI, ua);
#pragma unroll
		for (int j = 0; j < 4; j++) {
			const double dUpdateA = ua[j] * dInvDA;
			const double dUpdateB = TB_ROW_ST(gb, j) * dInvDB;
			const double b = HAS_BASE ? o[j] : 0.0;
			o[j] = b - ha.dt * cIJ[j] * dNuS * (dUpdateA + dUpdateB);
		}
		if (area_node != 0) {
			double area[4];
			tb_ld4g(area_node + ((size_t)e * L + (rc % L)) * NN + i * 4, area);
			tb_filter_level(o, area);
		}
		if (active) tb_st4(out + o4, o);
	}
}

///////////////////////////////////////////////////////////////////////////////
// VerticalDynamicsFEM::UpdateColumnTracers (VerticalDynamicsFEM.cpp:3783-4282)
// for vertical order 1 on the column-constant metric.
//
// With one level per vertical element the tracer matrix (:3953-4016) is
// tridiagonal and the same for every tracer of a column.  One thread per unique
// column marches down the levels once: it builds row k of the matrix and the
// right-hand sides of NT tracers in registers (xi-dot on the interfaces before /
// after the solve, upwind penalty, jump terms: :3933-3950, 4078-4230), runs the
// LAPACK dgbtf2 step (kl = ku = 1: pivot search over two rows, interchange,
// reciprocal scaling, rank-1 update) and the forward substitution of dgbtrs on
// all NT right-hand sides at once - NT independent dependency chains per
// thread.  Only the rows of U and the substituted right-hand sides are kept
// (block-private scratch in global memory, [entry][thread]: coalesced, written
// once and read once); the back substitution (dtbsv) marches up,
// subtracts the solution from the update instance and copies it to the
// duplicates of the column (:4265-4281, 1544-1633).

struct TracerColumnFastArgs {
	const int * col_node;
	const int * col_dups;
	int ncols;
	const double * colc;
	const double * lev;
	const double * w_old;     // [e][L+1][NN]: w before the implicit solve
	double dt;
	int c0;                   // first tracer of this pass
	int col0;                 // first column of this launch
	double * ws;              // scratch: (3 + NT) * L doubles per column of the launch
	double * keep;            // instance that receives the tracers from before the update, or 0
	int * info;
};

#define TBT_COL_THREADS 128

template <int NT>
__global__ void __launch_bounds__(TBT_COL_THREADS)
k_column_tracers_fast(
	DevLayout lay, TracerColumnFastArgs ta,
	const double * st_in,     // state before the solve (u, v)
	const double * st_out,    // state after the solve (w)
	const double * tr_in,     // instance holding the initial tracers
	double * tr_out           // instance whose tracers are updated
) {
	TB_DYN_SMEM(double, slev);          // [L+1][TBF_LW]
	const int L = lay.nlev;
	const int NN = lay.nn;
	const int T = TBT_COL_THREADS;
	for (int q = threadIdx.x; q < (L + 1) * TBF_LW; q += T) {
		slev[q] = ta.lev[q];
	}
	__syncthreads();
	// scratch of this block: U [3][L][T], Y [NT][L][T]
	double * sU = ta.ws + (size_t)blockIdx.x * (size_t)(3 + NT) * L * T + threadIdx.x;   // sU[(r * L + j) * T]
	double * sY = sU + (size_t)3 * L * T;                                            // sY[(c * L + j) * T]

	int tcol = blockIdx.x * T + threadIdx.x;
	const bool live = (tcol < ta.ncols);
	if (!live) tcol = ta.ncols - 1;
	tcol += ta.col0;
	const int node = ta.col_node[tcol];
	const long long e = node / NN;
	const int nd = node % NN;
	const size_t ebase = (size_t)e * lay.nrows * NN;
	const double * inU = st_in + ebase + (size_t)lay.rowoff[0] * NN + nd;
	const double * inV = st_in + ebase + (size_t)lay.rowoff[1] * NN + nd;
	const double * wNew = st_out + ebase + (size_t)lay.rowoff[3] * NN + nd;
	const double * wOld = ta.w_old + (size_t)e * (L + 1) * NN + nd;
	const int ntr = lay.ntr;
	// tracers of this pass (a short last pass repeats its last tracer, unstored)
	const double * tin[NT];
#pragma unroll
	for (int c = 0; c < NT; c++) {
		const int cc = (ta.c0 + c < ntr) ? (ta.c0 + c) : (ntr - 1);
		tin[c] = tr_in + ebase + (size_t)(lay.troff + cc * L) * NN + nd;
	}

	const double * ccol = ta.colc + (size_t)e * TBF_NC * NN + nd;
	const double cJ = ccol[TBF_JAC * NN];
	const double cA2 = ccol[TBF_A2 * NN], cB2 = ccol[TBF_B2 * NN];
	const double cX0 = ccol[TBF_X0 * NN], cX2 = ccol[TBF_X2 * NN];
	const double dInvDeltaT = 1.0 / ta.dt;
	const double cIJ = 1.0 / cJ;

	// interface m (1 <= m <= L-1) from levels m-1, m: xi-dot before (xi) and after
	// (xn) the solve, jump factor sign(xi-dot) g^{xi xi} (w_new - w_old)
	// (:3933-3950, 4078-4093, 4178-4228)
	double uP = inU[0], vP = inV[0];            // level m-1
	double uN = (L > 1) ? inU[NN] : 0.0;        // level m (prefetched)
	double vN = (L > 1) ? inV[NN] : 0.0;
	double woN = wOld[NN], wnN = wNew[NN];      // interface m (prefetched)

	double xi0 = 0.0, xn0 = 0.0, jp0 = 0.0;     // interface r
	double qm[NT], q0[NT], qp[NT];              // levels r-1, r, r+1
#pragma unroll
	for (int c = 0; c < NT; c++) {
		qm[c] = 0.0;
		q0[c] = tin[c][0];
		qp[c] = (L > 1) ? tin[c][NN] : 0.0;
	}
	double mf0[NT];                              // mass flux on interface r
#pragma unroll
	for (int c = 0; c < NT; c++) mf0[c] = 0.0;

	// row r-1 after its elimination steps: entries in columns r-1, r, r+1
	double pd = 0.0, ps1 = 0.0, ps2 = 0.0;
	double pb[NT];
#pragma unroll
	for (int c = 0; c < NT; c++) pb[c] = 0.0;
	int info = 0;

	for (int r = 0; r < L; r++) {
		const double * lv = slev + (size_t)r * TBF_LW;
		const double * lv1 = lv + TBF_LW;            // level / interface r+1
		// ---- interface r+1 ------------------------------------------------------------
		double xi1 = 0.0, xn1 = 0.0, jp1 = 0.0;
		double qq[NT];                               // level r+2
		const bool inner = (r + 1 < L);              // interface r+1 is interior
		{
			const double u1 = uN, v1 = vN, wo1 = woN, wn1 = wnN;
			// prefetch level r+2 / interface r+2
			const int k2 = (r + 2 < L) ? (r + 2) : (L - 1);
			uN = inU[(size_t)k2 * NN];
			vN = inV[(size_t)k2 * NN];
			const int m2 = (r + 2 < L) ? (r + 2) : L;
			woN = wOld[(size_t)m2 * NN];
			wnN = wNew[(size_t)m2 * NN];
#pragma unroll
			for (int c = 0; c < NT; c++) qq[c] = (r + 2 < L) ? tin[c][(size_t)(r + 2) * NN] : 0.0;
			if (inner) {
				const double i0 = lv1[TBF_CILO + 0], i1 = lv1[TBF_CILO + 1];
				double ue = 0.0, ve = 0.0;
				ue += i0 * uP; ue += i1 * u1;
				ve += i0 * vP; ve += i1 * v1;
				const double se = lv1[TBF_SE];
				const double c0 = se * cA2, c1 = se * cB2;
				const double c2 = cX0 + (se * se) * cX2;
				xi1 = c0 * ue + c1 * ve + c2 * wo1;
				xn1 = c0 * ue + c1 * ve + c2 * wn1;
				double dSignWeight;
				if (xi1 > 0.0) {
					dSignWeight = 1.0 * c2;
				} else if (xi1 < 0.0) {
					dSignWeight = -1.0 * c2;
				} else {
					dSignWeight = 0.0;
				}
				jp1 = dSignWeight * (wn1 - wo1);
			}
			uP = u1; vP = v1;
		}
		const double ax0 = fabs(xi0), ax1 = fabs(xi1);

		// ---- row r of the matrix (:3953-4016) ---------------------------------------------
		const double d0 = lv[TBF_DEN + 0], d1 = lv[TBF_DEN + 1];   // DiffREdgeToNode(r; r, r+1)
		const double pl0 = lv[TBF_CPL + 0], pl1 = lv[TBF_CPL + 1], pl2 = lv[TBF_CPL + 2];
		const double pr0 = lv[TBF_CPR + 0], pr1 = lv[TBF_CPR + 1], pr2 = lv[TBF_CPR + 2];
		const double i00 = lv[TBF_CILO + 0], i01 = lv[TBF_CILO + 1];   // InterpNodeToREdge(r; r-1, r)
		const double i10 = lv1[TBF_CILO + 0], i11 = lv1[TBF_CILO + 1]; // InterpNodeToREdge(r+1; r, r+1)
		double sub = 0.0, diag = 0.0, sup = 0.0;
		// (the reference forms coeff * J(interface) / J(level): the two Jacobians
		// are the same number for this metric)
		const double e0 = d0, e1 = d1;
		if (r >= 1) {
			sub += e0 * i00 * xi0;
			diag += e0 * i01 * xi0;
		}
		if (inner) {
			diag += e1 * i10 * xi1;
			sup += e1 * i11 * xi1;
		}
		if (r >= 1) {
			sub -= ax0 * pr0; diag -= ax0 * pr1; sup -= ax0 * pr2;
		}
		if (inner) {
			sub -= ax1 * pl0; diag -= ax1 * pl1; sup -= ax1 * pl2;
		}
		diag += dInvDeltaT;

		// ---- right-hand sides (:4096-4230) --------------------------------------------------
		double F[NT];
#pragma unroll
		for (int c = 0; c < NT; c++) {
			double mf1 = 0.0;
			if (inner) {
				double qe = 0.0;
				qe += i10 * q0[c]; qe += i11 * qp[c];
				mf1 = cJ * qe * xn1;
			}
			double f = 0.0;
			f += d0 * mf0[c]; f += d1 * mf1;
			f = f * cIJ;
			mf0[c] = mf1;
			double aux = 0.0;
			if (inner) {
				double a = 0.0;
				a += pl0 * qm[c]; a += pl1 * q0[c]; a += pl2 * qp[c];
				aux += a * ax1;
			}
			if (r >= 1) {
				double a = 0.0;
				a += pr0 * qm[c]; a += pr1 * q0[c]; a += pr2 * qp[c];
				aux += a * ax0;
			}
			f -= aux;
			if (r >= 1) {
				f -= pr0 * qm[c] * jp0; f -= pr1 * q0[c] * jp0; f -= pr2 * qp[c] * jp0;
			}
			if (inner) {
				f -= pl0 * qm[c] * jp1; f -= pl1 * q0[c] * jp1; f -= pl2 * qp[c] * jp1;
			}
			F[c] = f;
			qm[c] = q0[c]; q0[c] = qp[c]; qp[c] = qq[c];
		}
		xi0 = xi1; xn0 = xn1; jp0 = jp1;
		(void)xn0;

		// ---- dgbtf2 step j = r-1 on rows j (pd, ps1, ps2) and r (sub, diag, sup) ----------
		if (r == 0) {
			pd = diag; ps1 = sup; ps2 = 0.0;
#pragma unroll
			for (int c = 0; c < NT; c++) pb[c] = F[c];
			continue;
		}
		const int j = r - 1;
		double a0 = pd, a1 = ps1, a2 = ps2;       // pivot row
		double b0 = sub, b1 = diag, b2 = sup;     // the other row
		const bool swap = (fabs(sub) > fabs(pd));
		if (swap) {
			a0 = sub; a1 = diag; a2 = sup;
			b0 = pd; b1 = ps1; b2 = ps2;
		}
		double l = b0;
		if (a0 != 0.0) {
			const double rp = 1.0 / a0;
			l = b0 * rp;
			if (a1 != 0.0) b1 -= l * a1;
			if (a2 != 0.0) b2 -= l * a2;
		} else if (info == 0) {
			info = j + 1;
		}
		sU[(size_t)(0 * L + j) * T] = a0;
		sU[(size_t)(1 * L + j) * T] = a1;
		sU[(size_t)(2 * L + j) * T] = a2;
#pragma unroll
		for (int c = 0; c < NT; c++) {
			double bj = pb[c], bn = F[c];
			if (swap) { bj = F[c]; bn = pb[c]; }
			bn -= l * bj;
			sY[(size_t)(c * L + j) * T] = bj;
			pb[c] = bn;
		}
		pd = b1; ps1 = b2; ps2 = 0.0;
	}
	// last step: no rows below
	if (pd == 0.0 && info == 0) info = L;
	sU[(size_t)(0 * L + (L - 1)) * T] = pd;
	sU[(size_t)(1 * L + (L - 1)) * T] = 0.0;
	sU[(size_t)(2 * L + (L - 1)) * T] = 0.0;
#pragma unroll
	for (int c = 0; c < NT; c++) sY[(size_t)(c * L + (L - 1)) * T] = pb[c];

	// ---- dtbsv + update + duplicates ------------------------------------------------------
	const int * dups = ta.col_dups + (size_t)tcol * 3;
	const int d0 = dups[0], d1 = dups[1], d2 = dups[2];
	const size_t ob0 = (d0 >= 0) ? ((size_t)(d0 / NN) * lay.nrows * NN + (d0 % NN)) : 0;
	const size_t ob1 = (d1 >= 0) ? ((size_t)(d1 / NN) * lay.nrows * NN + (d1 % NN)) : 0;
	const size_t ob2 = (d2 >= 0) ? ((size_t)(d2 / NN) * lay.nrows * NN + (d2 % NN)) : 0;
	double x1[NT], x2[NT];
	bool bad = (info != 0);
#pragma unroll
	for (int c = 0; c < NT; c++) { x1[c] = 0.0; x2[c] = 0.0; }
	// rows of U, substituted right-hand sides and the values to be updated of level
	// j - 1 are loaded while level j is computed
	double nu0 = sU[(size_t)(0 * L + (L - 1)) * T];
	double nu1 = sU[(size_t)(1 * L + (L - 1)) * T];
	double nu2 = sU[(size_t)(2 * L + (L - 1)) * T];
	double ny[NT], nown[NT];
#pragma unroll
	for (int c = 0; c < NT; c++) {
		const int cc = (ta.c0 + c < ntr) ? (ta.c0 + c) : (ntr - 1);
		ny[c] = sY[(size_t)(c * L + (L - 1)) * T];
		nown[c] = tr_out[ebase + (size_t)(lay.troff + cc * L + (L - 1)) * NN + nd];
	}
	for (int j = L - 1; j >= 0; j--) {
		const double u0 = nu0, u1 = nu1, u2 = nu2;
		double y[NT], oldv[NT];
#pragma unroll
		for (int c = 0; c < NT; c++) { y[c] = ny[c]; oldv[c] = nown[c]; }
		if (j > 0) {
			nu0 = sU[(size_t)(0 * L + j - 1) * T];
			nu1 = sU[(size_t)(1 * L + j - 1) * T];
			nu2 = sU[(size_t)(2 * L + j - 1) * T];
#pragma unroll
			for (int c = 0; c < NT; c++) {
				const int cc = (ta.c0 + c < ntr) ? (ta.c0 + c) : (ntr - 1);
				ny[c] = sY[(size_t)(c * L + j - 1) * T];
				nown[c] = tr_out[ebase + (size_t)(lay.troff + cc * L + j - 1) * NN + nd];
			}
		}
		const double ru0 = 1.0 / u0;       // one reciprocal for the NT right-hand sides
#pragma unroll
		for (int c = 0; c < NT; c++) {
			double b = y[c];
			b -= x2[c] * u2;
			b -= x1[c] * u1;
			b = b * ru0;
			x2[c] = x1[c];
			x1[c] = b;
			if (!(b == b)) bad = true;
			if (live && ta.c0 + c < ntr) {
				const size_t row = (size_t)(lay.troff + (ta.c0 + c) * L + j) * NN;
				double * own = tr_out + ebase + row + nd;
				const double old = oldv[c];
				const double v = old - b;
				if (ta.keep != 0) {
					ta.keep[ebase + row + nd] = old;
					if (d0 >= 0) ta.keep[ob0 + row] = tr_out[ob0 + row];
					if (d1 >= 0) ta.keep[ob1 + row] = tr_out[ob1 + row];
					if (d2 >= 0) ta.keep[ob2 + row] = tr_out[ob2 + row];
				}
				own[0] = v;
				if (d0 >= 0) tr_out[ob0 + row] = v;
				if (d1 >= 0) tr_out[ob1 + row] = v;
				if (d2 >= 0) tr_out[ob2 + row] = v;
			}
		}
	}
	if (bad && live) atomicMax(ta.info, tcol + 1);
}

#endif
