// C ABI of libtempest_b200 (see include/tempest_b200.h).
//
// Host-side plumbing only: context, layout, uploads, connectivity, dispatch.
// All arithmetic on model data happens in the kernels of tb200_kernels.cuh,
// tb200_dss.cuh and tb200_column.cuh; there is no CPU fallback.

#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <cmath>
#include <algorithm>
#include <unordered_map>

#include "tb200_ctx.h"
#include "tb200_kernels.cuh"
#include "tb200_dss.cuh"
#include "tb200_column.cuh"
#include "tb200_fast.cuh"
#include "tb200_column_fast.cuh"
#ifndef TB200_EMU
#include <cuda.h>
#endif
#include "tb200_tracers.cuh"
#include "tb200_tracers_fast.cuh"
#include "tb200_diag.cuh"
#include "tb200_physics.cuh"
#include "tb200_setup.cuh"
#include "tb200_output.cuh"

static_assert(TBT_C_JAC == TBF_JAC && TBT_C_A2 == TBF_A2 && TBT_C_B2 == TBF_B2
	&& TBT_C_X0 == TBF_X0 && TBT_C_X2 == TBF_X2 && TBT_C_NC == TBF_NC
	&& TBT_L_SE == TBF_SE && TBT_L_LW == TBF_LW, "column-constant indices of tb200_tracers.cuh");

#define TB_CHECK(ctx, call) \
	do { \
		cudaError_t e__ = (call); \
		if (e__ != cudaSuccess) { \
			(ctx)->err = std::string(#call) + ": " + cudaGetErrorString(e__); \
			return 1; \
		} \
	} while (0)

#define TB_FAIL(ctx, msg) \
	do { (ctx)->err = (msg); return 1; } while (0)

#define TB_KERNEL_CHECK(ctx) \
	do { \
		(ctx)->launches++; \
		(ctx)->writes++; \
		cudaError_t e__ = cudaGetLastError(); \
		if (e__ != cudaSuccess) { \
			(ctx)->err = std::string("kernel launch: ") + cudaGetErrorString(e__); \
			return 1; \
		} \
	} while (0)

// error check of launches that were counted one by one
#define TB_LAUNCH_CHECK(ctx) \
	do { \
		cudaError_t e__ = cudaGetLastError(); \
		if (e__ != cudaSuccess) { \
			(ctx)->err = std::string("kernel launch: ") + cudaGetErrorString(e__); \
			return 1; \
		} \
	} while (0)

// FunctionTimer group of an entry point (tb200_set_timing_hooks): begin at
// construction, device work awaited and end at destruction.
struct TimingScope {
	tb200_ctx * ctx;
	const char * group;
	TimingScope(tb200_ctx * c, const char * g) : ctx(c), group(g) {
		if (ctx->timing_begin != 0) ctx->timing_begin(ctx->timing_user, group);
	}
	~TimingScope() {
		if (ctx->timing_end != 0) {
			cudaStreamSynchronize(ctx->stream);
			ctx->timing_end(ctx->timing_user, group);
		}
	}
};

extern "C" int tb200_set_timing_hooks(
	tb200_ctx * ctx, tb200_timing_fn begin, tb200_timing_fn end, void * user
) {
	ctx->timing_begin = begin;
	ctx->timing_end = end;
	ctx->timing_user = user;
	return 0;
}

static const int kItems = 8;   // (element, level) pairs per block in the slab kernels

// the general element kernels are templates on the horizontal order
#define TB_NP_SWITCH(np, ...) \
	switch (np) { \
		case 3: { constexpr int NPV = 3; __VA_ARGS__ } break; \
		case 4: { constexpr int NPV = 4; __VA_ARGS__ } break; \
		case 5: { constexpr int NPV = 5; __VA_ARGS__ } break; \
		case 6: { constexpr int NPV = 6; __VA_ARGS__ } break; \
		default: TB_FAIL(ctx, "horizontal order not instantiated"); \
	}

template <typename T>
static int dalloc(tb200_ctx * ctx, T ** p, size_t count) {
	void * q = 0;
	cudaError_t e = cudaMalloc(&q, std::max<size_t>(count, 1) * sizeof(T));
	if (e != cudaSuccess) {
		ctx->err = std::string("cudaMalloc: ") + cudaGetErrorString(e);
		return 1;
	}
	ctx->allocs.push_back(q);
	*p = (T *)q;
	return 0;
}

template <typename T>
static int dupload(tb200_ctx * ctx, T ** p, const std::vector<T> & v) {
	if (dalloc(ctx, p, v.size())) return 1;
	if (v.size() != 0) {
		TB_CHECK(ctx, cudaMemcpy(*p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
	}
	return 0;
}

// Grid of a persistent kernel: resident blocks per SM (registers and shared
// memory both counted by the runtime) times the SM count, so that every block
// of the launch is resident and walks the same number of elements.
// Tensor map of a state instance: [nelem * nrows][16] doubles, boxes of whole
// 128-byte rows, 128-byte swizzle (tb200_tma.cuh).  cuTensorMapEncodeTiled is a
// driver entry point; it is looked up through the runtime.
static int make_tensor_map(tb200_ctx * ctx, const double * base, TbMap * out) {
	const DevLayout & lay = ctx->lay;
	memset(out, 0, sizeof(TbMap));
	out->base = base;
	out->nbox = tb_tma_nbox(lay.nrows_state);
	out->boxrows = tb_tma_boxrows(lay.nrows_state);
#ifndef TB200_EMU
	typedef CUresult (*encode_fn)(
		CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
		const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
		CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
	static encode_fn encode = 0;
	if (encode == 0) {
		void * fn = 0;
		cudaDriverEntryPointQueryResult qres;
		TB_CHECK(ctx, cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
		if (fn == 0 || qres != cudaDriverEntryPointSuccess) {
			TB_FAIL(ctx, "cuTensorMapEncodeTiled is not available in this driver");
		}
		encode = (encode_fn)fn;
	}
	static_assert(sizeof(CUtensorMap) == 128, "CUtensorMap size");
	const cuuint64_t gdim[2] = {16, (cuuint64_t)lay.nelem * (cuuint64_t)lay.nrows};
	const cuuint64_t gstride[1] = {128};
	const cuuint32_t box[2] = {16, (cuuint32_t)out->boxrows};
	const cuuint32_t estride[2] = {1, 1};
	const CUresult rc = encode(
		reinterpret_cast<CUtensorMap *>(out->desc), CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2,
		const_cast<double *>(base), gdim, gstride, box, estride,
		CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
		CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
	if (rc != CUDA_SUCCESS) {
		char buf[96];
		snprintf(buf, 96, "cuTensorMapEncodeTiled failed (%d)", (int)rc);
		TB_FAIL(ctx, buf);
	}
#endif
	return 0;
}

// tensor map of the instance a device pointer belongs to
static const TbMap & tensor_map_of(const tb200_ctx * ctx, const double * p) {
	for (size_t m = 0; m < ctx->inst.size(); m++) {
		if (ctx->inst[m] == p) return ctx->tmaps[m];
	}
	return ctx->tmaps[0];
}

template <typename K>
static long long persistent_blocks(
	tb200_ctx * ctx, K kfn, int threads, size_t smem, long long nwork, int reserve_sms = 0
) {
	int per_sm = 1;
#ifndef TB200_EMU
	if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kfn, threads, smem) != cudaSuccess
		|| per_sm < 1) {
		per_sm = 1;
	}
#else
	(void)kfn; (void)threads; (void)smem;
	per_sm = 2;
#endif
	// reserve_sms: SMs left free for the pack kernel and the collective that run
	// next to this launch (persistent blocks would otherwise hold every SM)
	long long nb = (long long)std::max(1, ctx->sm_count - reserve_sms) * per_sm;
	const char * fb = getenv("TB200_PIPE_BLOCKS");   // tests: force the multi-element loop
	if (fb != 0 && atoi(fb) > 0) nb = atoi(fb);
	if (nb > nwork) nb = nwork;
	return nb;
}

// Blocks of a launch with the fused DSS: all resident (they wait for each other's
// elements), at most one per strip.  The emulation runs blocks one after another:
// one block, strips in order.
template <typename K>
static long long fused_blocks(tb200_ctx * ctx, K kfn, int threads, size_t smem) {
#ifdef TB200_EMU
	(void)kfn; (void)threads; (void)smem; (void)ctx;
	return 1;
#else
	int per_sm = 1;
	if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kfn, threads, smem) != cudaSuccess
		|| per_sm < 1) {
		per_sm = 1;
	}
	long long nb = (long long)ctx->sm_count * per_sm;
	if (nb > ctx->nstrips) nb = ctx->nstrips;
	return nb;
#endif
}

// Multi-rank overlap of the halo exchange with compute: an element kernel that
// is followed by a DSS runs first on the elements that own nodes of the send
// list (same stream as the pack + exchange), and on all other elements on a
// second stream while the exchange is in flight; dss_rows joins the two before
// it averages.  part: 0 = every element on ctx->stream; 1 = exchange-feeding
// elements on ctx->stream; 2 = the rest on ctx->stream2.
// TB200_OVERLAP=1 turns the exchange / compute overlap on (off by default until
// it is measured to pay on the target node count)
static bool overlap_wanted() {
	const char * ov = getenv("TB200_OVERLAP");
	return ov != 0 && strcmp(ov, "1") == 0;
}

static int overlap_reserve() {
	const char * r = getenv("TB200_OVERLAP_RESERVE");
	return (r != 0) ? atoi(r) : 8;
}

static bool split_enabled(const tb200_ctx * ctx) {
	return ctx->want_split && ctx->nranks > 1 && ctx->d_elist_bnd != 0 && ctx->n_int > 0;
}

static ElemList elem_list(const tb200_ctx * ctx, int part) {
	ElemList el;
	if (part == 1) { el.list = ctx->d_elist_bnd; el.n = ctx->n_bnd; }
	else if (part == 2) { el.list = ctx->d_elist_int; el.n = ctx->n_int; }
	else { el.list = 0; el.n = (int)ctx->lay.nelem; }
	return el;
}

static int split_fork(tb200_ctx * ctx) {
#ifndef TB200_EMU
	if (ctx->stream2 == 0) {
		TB_CHECK(ctx, cudaStreamCreateWithFlags(&ctx->stream2, cudaStreamNonBlocking));
		TB_CHECK(ctx, cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
		TB_CHECK(ctx, cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming));
	}
	TB_CHECK(ctx, cudaEventRecord(ctx->ev_fork, ctx->stream));
	TB_CHECK(ctx, cudaStreamWaitEvent(ctx->stream2, ctx->ev_fork, 0));
#endif
	return 0;
}

static int split_mark(tb200_ctx * ctx) {
#ifndef TB200_EMU
	TB_CHECK(ctx, cudaEventRecord(ctx->ev_join, ctx->stream2));
#endif
	ctx->split_pending = true;
	return 0;
}

static int split_join(tb200_ctx * ctx) {
	if (!ctx->split_pending) return 0;
#ifndef TB200_EMU
	TB_CHECK(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));
#endif
	ctx->split_pending = false;
	return 0;
}

// the next launch of a pipelined kernel may average the in-patch groups itself
static bool fuse_enabled(const tb200_ctx * ctx) {
	if (!ctx->fuse_ready || !ctx->fuse_want || split_enabled(ctx)) return false;
	if (ctx->lay.ntr > 0) return false;                           // tracer rows are averaged by the group kernels
	if (ctx->lay.nrows > TBF_AROWS * TBF_THREADS) return false;   // rows per thread of the alpha phase
	// Off unless TB200_DSS_FUSED=1: measured on the B200 (ne = 120, L = 30) the
	// fused kernels cut the DRAM traffic of stage + DSS from 8.5 to 5.7 GB but run
	// 2.7-3.7 ms against 2.0-2.3 ms for stage kernel + separate DSS pass - flag
	// polls, the alpha phase and its scattered 8-byte stores sit on the critical
	// path of a kernel that is latency-bound at 8 warps per SM
	// (profiles/r2_fused_dss_experiment.txt).  Kept, with its bit-for-bit tests,
	// as the starting point for a cluster / DSMEM variant.
	const char * e = getenv("TB200_DSS_FUSED");
	if (e == 0 || strcmp(e, "1") != 0) return false;
	if (getenv("TB200_PIPE_BLOCKS") != 0) return false;     // tests of the strided walk
	return true;
}

static FuseArgs fuse_args(tb200_ctx * ctx) {
	FuseArgs fz;
	memset(&fz, 0, sizeof(fz));
	fz.strip_first = ctx->d_strip_first;
	fz.strip_len = ctx->d_strip_len;
	fz.strip_neb = ctx->d_strip_neb;
	fz.nstrips = ctx->nstrips;
	fz.done = ctx->d_done;
	fz.epoch = ctx->fuse_epoch;
	return fz;
}

static PatchInfo * find_patch(tb200_ctx * ctx, int patch_index) {
	std::map<int, int>::iterator it = ctx->patch_pos.find(patch_index);
	if (it == ctx->patch_pos.end()) return 0;
	return &ctx->patches[it->second];
}

///////////////////////////////////////////////////////////////////////////////

extern "C" const char * tb200_version(void) {
#ifdef TB200_EMU
	return "tempest-b200 0.1 (host emulation build - tests only)";
#else
	return "tempest-b200 0.1 (CUDA sm_100a)";
#endif
}

extern "C" const char * tb200_last_error(const tb200_ctx * ctx) {
	return ctx ? ctx->err.c_str() : "null context";
}

extern "C" int tb200_create(const tb200_config * cfg, tb200_ctx ** out) {
	if (cfg == 0 || out == 0) return 1;
	tb200_ctx * ctx = new tb200_ctx();
	*out = ctx;
	ctx->cfg = *cfg;
	// (--order: the general kernels are instantiated for these orders; the
	// column-constant path is np = 4 only)
	if (cfg->np != 3 && cfg->np != 4 && cfg->np != 5 && cfg->np != 6) {
		TB_FAIL(ctx, "only np = 3, 4, 5, 6 kernels are instantiated in this build");
	}
	if (cfg->nlev < 1 || cfg->ncomp < 1 || cfg->ncomp > TB_MAXC) {
		TB_FAIL(ctx, "invalid nlev / ncomp");
	}
	if (cfg->ninstances < 1 || cfg->ninstances > TB_MAXINST) {
		TB_FAIL(ctx, "invalid ninstances");
	}
	if (cfg->eqn_type == TB200_EQN_SHALLOW_WATER) {
		if (cfg->ncomp != 3) TB_FAIL(ctx, "shallow water needs 3 components");
	} else if (cfg->eqn_type == TB200_EQN_PRIMITIVE_NONHYDRO) {
		if (cfg->ncomp != 5) TB_FAIL(ctx, "nonhydrostatic needs 5 components");
		// Lorenz staggering only (reference default --vstagger LOR)
		const int want[5] = {0, 0, 0, 1, 0};
		for (int c = 0; c < 5; c++) {
			if ((cfg->comp_on_redge[c] != 0) != (want[c] != 0)) {
				TB_FAIL(ctx, "only Lorenz staggering (W on interfaces) is supported");
			}
		}
	} else {
		TB_FAIL(ctx, "unsupported equation set");
	}
#ifndef TB200_EMU
	int ndev = 0;
	if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
		TB_FAIL(ctx, "no CUDA device: libtempest_b200 has no CPU fallback");
	}
#endif
	if (cfg->device >= 0) {
		TB_CHECK(ctx, cudaSetDevice(cfg->device));
	}
	ctx->sm_count = 148;
#ifndef TB200_EMU
	{
		int dev = 0, sms = 0;
		if (cudaGetDevice(&dev) == cudaSuccess
			&& cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess
			&& sms > 0) {
			ctx->sm_count = sms;
		}
	}
#endif

	DevLayout & lay = ctx->lay;
	memset(&lay, 0, sizeof(lay));
	lay.np = cfg->np;
	lay.nn = cfg->np * cfg->np;
	lay.nlev = cfg->nlev;
	lay.ncomp = cfg->ncomp;
	lay.ntr = cfg->ntracers;
	int row = 0;
	for (int c = 0; c < cfg->ncomp; c++) {
		lay.rowoff[c] = row;
		lay.onedge[c] = cfg->comp_on_redge[c] ? 1 : 0;
		lay.rowlev[c] = cfg->nlev + lay.onedge[c];
		row += lay.rowlev[c];
	}
	lay.nrows_state = row;
	lay.troff = row;
	lay.nrows = row + cfg->ntracers * cfg->nlev;

	ctx->phys.g = cfg->g;
	ctx->phys.R = cfg->R;
	ctx->phys.cp = cfg->cp;
	ctx->phys.cv = cfg->cv;
	ctx->phys.p0 = cfg->p0;
	ctx->phys.exner_c1 = cfg->R / (cfg->cp - cfg->R);
	ctx->phys.exner_c2 = cfg->R / cfg->p0;
	{
		// R / cv == 2/5 (to rounding): the fast kernels take tb_exner's Newton form;
		// TB200_EXNER=libm keeps exp(c log x)
		const char * ex = getenv("TB200_EXNER");
		ctx->phys.exner25 = (fabs(ctx->phys.exner_c1 - 0.4) <= 1.0e-15
			&& !(ex != 0 && strcmp(ex, "libm") == 0)) ? 1 : 0;
	}

	// m_nJacobianFOffD (VerticalDynamicsFEM.cpp:188-201, FE discretisation)
	switch (cfg->vertical_order) {
		case 1: ctx->offd = 4; break;
		case 2: ctx->offd = 9; break;
		case 3: ctx->offd = 15; break;
		case 4: ctx->offd = 22; break;
		case 5: ctx->offd = 30; break;
		default: TB_FAIL(ctx, "unsupported vertical order");
	}
	ctx->fe_nodes = cfg->vertical_order;
	ctx->finite_volume = 0;
	ctx->mass_flux_levels = 0;
	memset(&ctx->ops, 0, sizeof(ctx->ops));
	memset(&ctx->geom, 0, sizeof(ctx->geom));
	memset(&ctx->tables, 0, sizeof(ctx->tables));
	ctx->carry_full = (getenv("TB200_CARRY_FULL") != 0);
	TB_CHECK(ctx, cudaMallocHost((void **)&ctx->h_info, 4 * sizeof(int)));
	memset(ctx->h_info, 0, 4 * sizeof(int));
	return 0;
}

extern "C" int tb200_destroy(tb200_ctx * ctx) {
	if (ctx == 0) return 0;
	cudaDeviceSynchronize();
#ifndef TB200_EMU
	for (size_t r = 0; r < ctx->peer_base.size(); r++) {
		if (ctx->peer_base[r] != 0) cudaIpcCloseMemHandle(ctx->peer_base[r]);
	}
	if (ctx->peer_area != 0) cudaFree(ctx->peer_area);
#endif
	for (size_t i = 0; i < ctx->allocs.size(); i++) {
		cudaFree(ctx->allocs[i]);
	}
	if (ctx->h_info != 0) cudaFreeHost(ctx->h_info);
#ifndef TB200_EMU
	if (ctx->copy_stream != 0) {
		cudaStreamDestroy(ctx->copy_stream);
		for (int q = 0; q < 2; q++) {
			cudaEventDestroy(ctx->ev_stage_free[q]);
			cudaEventDestroy(ctx->ev_stage_full[q]);
		}
		cudaEventDestroy(ctx->ev_compute);
	}
#endif
	delete ctx;
	return 0;
}

extern "C" int tb200_set_stream(tb200_ctx * ctx, void * s) {
	ctx->stream = (cudaStream_t)s;
	return 0;
}

extern "C" int tb200_sync(tb200_ctx * ctx) {
	TB_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
	return 0;
}

extern "C" int64_t tb200_launch_count(const tb200_ctx * ctx) {
	return ctx->launches;
}

extern "C" int64_t tb200_fused_group_count(const tb200_ctx * ctx) {
	return ctx->fuse_ready ? (int64_t)(ctx->ngroups - ctx->nrem) : 0;
}

extern "C" int64_t tb200_column_count(const tb200_ctx * ctx) {
	return (int64_t)ctx->lay.nelem * ctx->lay.nn;
}

///////////////////////////////////////////////////////////////////////////////

extern "C" int tb200_set_exchange(
	tb200_ctx * ctx, int rank, int nranks, tb200_exchange_fn fn, void * user
) {
	if (ctx->patches.size() != 0) {
		TB_FAIL(ctx, "tb200_set_exchange must precede tb200_add_patch");
	}
	if (nranks < 1 || rank < 0 || rank >= nranks) TB_FAIL(ctx, "invalid rank");
	if (nranks > 1 && fn == 0) TB_FAIL(ctx, "exchange callback required");
	ctx->rank = rank;
	ctx->nranks = nranks;
	ctx->exch_fn = fn;
	ctx->exch_user = user;
	return 0;
}

extern "C" int tb200_add_patch(
	tb200_ctx * ctx, int patch_index, int panel, int nelem_a, int nelem_b,
	int halo, double delta_a, double delta_b, int owner_rank
) {
	if (ctx->committed) TB_FAIL(ctx, "layout already committed");
	if (ctx->patch_pos.count(patch_index)) TB_FAIL(ctx, "duplicate patch index");
	if (owner_rank < 0 || owner_rank >= ctx->nranks) TB_FAIL(ctx, "invalid owner rank");
	PatchInfo p;
	p.index = patch_index;
	p.panel = panel;
	p.nea = nelem_a;
	p.neb = nelem_b;
	p.halo = halo;
	p.owner = owner_rank;
	p.da = delta_a;
	p.db = delta_b;
	p.elem0 = -1;
	ctx->patch_pos[patch_index] = (int)ctx->patches.size();
	ctx->patches.push_back(p);
	return 0;
}

extern "C" int tb200_commit_layout(tb200_ctx * ctx) {
	if (ctx->committed) TB_FAIL(ctx, "layout already committed");
	DevLayout & lay = ctx->lay;
	const int np = lay.np, nn = lay.nn, L = lay.nlev;

	long long nelem = 0;
	size_t stage = 0;
	for (size_t p = 0; p < ctx->patches.size(); p++) {
		PatchInfo & pi = ctx->patches[p];
		if (pi.owner != ctx->rank) continue;
		pi.elem0 = nelem;
		nelem += (long long)pi.nea * pi.neb;
		const size_t wa = pi.nea * np + 2 * pi.halo;
		const size_t wb = pi.neb * np + 2 * pi.halo;
		const size_t ncmax = std::max(std::max(lay.ncomp, lay.ntr), 3);
		stage = std::max(stage, wa * wb * (size_t)(L + 1) * ncmax);
	}
	if (nelem == 0) TB_FAIL(ctx, "no local patches");
	if (nelem * nn >= (1ll << 30)) TB_FAIL(ctx, "too many local nodes for 32-bit node addresses");
	lay.nelem = nelem;

	const size_t inst_doubles = (size_t)nelem * lay.nrows * nn;
	ctx->inst.resize(ctx->cfg.ninstances);
	for (int m = 0; m < ctx->cfg.ninstances; m++) {
		if (dalloc(ctx, &ctx->inst[m], inst_doubles)) return 1;
		TB_CHECK(ctx, cudaMemset(ctx->inst[m], 0, inst_doubles * sizeof(double)));
	}
	ctx->tmaps.resize(ctx->cfg.ninstances);
	for (int m = 0; m < ctx->cfg.ninstances && lay.np == 4; m++) {
		if (make_tensor_map(ctx, ctx->inst[m], &ctx->tmaps[m])) return 1;
	}
	ctx->stage_doubles = stage;
	if (dalloc(ctx, &ctx->d_stage, stage)) return 1;
	if (dalloc(ctx, &ctx->d_rowmap, 64)) return 1;

	// per-element spacing and viscosity scaling
	// (HorizontalDynamicsFEM.cpp:1970-1975: nu * (deltaA / ref)^3.2)
	std::vector<double> ida(nelem), idb(nelem), nus(nelem);
	for (size_t p = 0; p < ctx->patches.size(); p++) {
		const PatchInfo & pi = ctx->patches[p];
		if (pi.elem0 < 0) continue;
		for (long long q = 0; q < (long long)pi.nea * pi.neb; q++) {
			ida[pi.elem0 + q] = 1.0 / pi.da;
			idb[pi.elem0 + q] = 1.0 / pi.db;
			nus[pi.elem0 + q] = (ctx->cfg.ref_length != 0.0)
				? pow(pi.da / ctx->cfg.ref_length, 3.2) : 1.0;
		}
	}
	if (dupload(ctx, &ctx->d_inv_da, ida)) return 1;
	if (dupload(ctx, &ctx->d_inv_db, idb)) return 1;
	if (dupload(ctx, &ctx->d_nu_scale, nus)) return 1;

	for (int q = 0; q < 7; q++) {
		if (dalloc(ctx, &ctx->g2d[q], (size_t)nelem * nn)) return 1;
		TB_CHECK(ctx, cudaMemset(ctx->g2d[q], 0, (size_t)nelem * nn * sizeof(double)));
	}
	const bool need3d = (ctx->cfg.eqn_type == TB200_EQN_PRIMITIVE_NONHYDRO);
	// The 3-D metric arrays (26 values per node against 5 of the state) are
	// allocated when the host uploads them (tb200_upload_geometry).  A host that
	// supplies only the 2-D metric, the topography derivatives and the vertical
	// coordinate runs on the column constants / the on-the-fly metric and keeps
	// that memory for the state: ne = 240, L = 60 holds 5 instances in 67 GB
	// instead of 136 GB.
	DevGeom & g = ctx->geom;
	g.inv_da = ctx->d_inv_da; g.inv_db = ctx->d_inv_db; g.nu_scale = ctx->d_nu_scale;
	g.j2d = ctx->g2d[0]; g.a0 = ctx->g2d[1]; g.a1 = ctx->g2d[2];
	g.b0 = ctx->g2d[3]; g.b1 = ctx->g2d[4]; g.f = ctx->g2d[5]; g.zs = ctx->g2d[6];
	if (dalloc(ctx, &ctx->d_sums, 64)) return 1;
	if (dalloc(ctx, &ctx->d_info, 4)) return 1;
	TB_CHECK(ctx, cudaMemset(ctx->d_info, 0, 4 * sizeof(int)));

	// unique columns of the implicit solve and the duplicates that receive
	// a copy (VerticalDynamicsFEM.cpp:1315-1334, 1544-1633)
	if (need3d) {
		std::vector<int> cn, cd;
		for (size_t p = 0; p < ctx->patches.size(); p++) {
			const PatchInfo & pi = ctx->patches[p];
			if (pi.elem0 < 0) continue;
			for (int a = 0; a < pi.nea; a++)
			for (int b = 0; b < pi.neb; b++) {
				const int iEnd = (a == pi.nea - 1) ? np : np - 1;
				const int jEnd = (b == pi.neb - 1) ? np : np - 1;
				for (int i = 0; i < iEnd; i++)
				for (int j = 0; j < jEnd; j++) {
					const long long e = pi.elem0 + (long long)a * pi.neb + b;
					cn.push_back((int)(e * nn + i * np + j));
					int d[3] = {-1, -1, -1};
					int nd = 0;
					const bool da_ = (i == 0 && a > 0);
					const bool db_ = (j == 0 && b > 0);
					if (da_) {
						const long long e2 = pi.elem0 + (long long)(a - 1) * pi.neb + b;
						d[nd++] = (int)(e2 * nn + (np - 1) * np + j);
					}
					if (db_) {
						const long long e2 = pi.elem0 + (long long)a * pi.neb + (b - 1);
						d[nd++] = (int)(e2 * nn + i * np + (np - 1));
					}
					if (da_ && db_) {
						const long long e2 = pi.elem0 + (long long)(a - 1) * pi.neb + (b - 1);
						d[nd++] = (int)(e2 * nn + (np - 1) * np + (np - 1));
					}
					cd.push_back(d[0]); cd.push_back(d[1]); cd.push_back(d[2]);
				}
			}
		}
		ctx->ncols = (int)cn.size();
		if (dupload(ctx, &ctx->d_col_node, cn)) return 1;
		if (dupload(ctx, &ctx->d_col_dups, cd)) return 1;
		ctx->ws_cols = std::min(ctx->ncols, 1 << 17);
		const size_t wsd = (size_t)tb_column_ws_entries(L, ctx->offd) * ctx->ws_cols;
		if (dalloc(ctx, &ctx->d_ws, wsd)) return 1;
	}
	ctx->committed = true;
	return 0;
}

extern "C" int tb200_set_tables(
	tb200_ctx * ctx, const double * dx, const double * st, const double * w
) {
	const int np = ctx->lay.np;
	for (int q = 0; q < np * np; q++) {
		ctx->tables.dx[q] = dx[q];
		ctx->tables.st[q] = st[q];
	}
	(void)w;
	return 0;
}

extern "C" int tb200_set_column_op(
	tb200_ctx * ctx, int op, int nout, int nin,
	const double * coeff, const int * begin, const int * end
) {
	if (op < 0 || op >= TB_NOPS) TB_FAIL(ctx, "invalid column operator id");
	HostOp & h = ctx->hops[op];
	h.nout = nout;
	h.nin = nin;
	h.begin.assign(nout, 0);
	h.end.assign(nout, 0);
	int width = 1;
	for (int k = 0; k < nout; k++) {
		int b = std::max(begin[k], 0);
		int e = std::min(end[k], nin);
		if (e < b) e = b;
		h.begin[k] = b;
		h.end[k] = e;
		width = std::max(width, e - b);
	}
	h.width = width;
	h.coeff.assign((size_t)nout * width, 0.0);
	for (int k = 0; k < nout; k++) {
		for (int l = h.begin[k]; l < h.end[k]; l++) {
			h.coeff[(size_t)k * width + (l - h.begin[k])] = coeff[(size_t)k * nin + l];
		}
	}
	if (dupload(ctx, &h.d_coeff, h.coeff)) return 1;
	if (dupload(ctx, &h.d_begin, h.begin)) return 1;
	if (dupload(ctx, &h.d_end, h.end)) return 1;
	DevOp & d = ctx->ops.op[op];
	d.coeff = h.d_coeff;
	d.begin = h.d_begin;
	d.end = h.d_end;
	d.width = width;
	d.nout = nout;
	d.nin = nin;
	return 0;
}

///////////////////////////////////////////////////////////////////////////////

// 3-D metric arrays: device storage on first use
static int ensure_geom3d(tb200_ctx * ctx, int q0, int nq, bool edge) {
	const DevLayout & lay = ctx->lay;
	const size_t count = (size_t)lay.nelem * (lay.nlev + (edge ? 1 : 0)) * lay.nn;
	double ** arr = edge ? ctx->g3e : ctx->g3n;
	for (int q = q0; q < q0 + nq; q++) {
		if (arr[q] != 0) continue;
		if (dalloc(ctx, &arr[q], count)) return 1;
		TB_CHECK(ctx, cudaMemset(arr[q], 0, count * sizeof(double)));
	}
	DevGeom & g = ctx->geom;
	g.jac = ctx->g3n[0];
	g.jace = ctx->g3e[0];
	for (int m = 0; m < 3; m++) {
		g.ca[m] = ctx->g3n[1 + m]; g.cb[m] = ctx->g3n[4 + m];
		g.cx[m] = ctx->g3n[7 + m]; g.dr[m] = ctx->g3n[10 + m];
		g.cae[m] = ctx->g3e[1 + m]; g.cbe[m] = ctx->g3e[4 + m];
		g.cxe[m] = ctx->g3e[7 + m]; g.dre[m] = ctx->g3e[10 + m];
	}
	return 0;
}

static int upload_geom_array(
	tb200_ctx * ctx, const PatchInfo & pi, const double * host,
	int nlev, int nm, double * d0, double * d1, double * d2
) {
	if (host == 0) return 0;
	const int np = ctx->lay.np;
	const size_t wa = pi.nea * np + 2 * pi.halo;
	const size_t wb = pi.neb * np + 2 * pi.halo;
	const size_t count = wa * wb * nlev * nm;
	if (count > ctx->stage_doubles) TB_FAIL(ctx, "staging buffer too small");
	TB_CHECK(ctx, cudaMemcpyAsync(ctx->d_stage, host, count * sizeof(double),
		cudaMemcpyHostToDevice, ctx->stream));
	GeomDst dst;
	dst.p[0] = d0; dst.p[1] = d1; dst.p[2] = d2;
	auto kfn = k_transpose_geom;
	TB_LAUNCH_FLAT(kfn, dim3(pi.nea * pi.neb), dim3(256), 0, ctx->stream,
		np, ctx->lay.nn, pi.elem0, pi.nea, pi.neb, pi.halo,
		(const double *)ctx->d_stage, nlev, nm, dst);
	TB_KERNEL_CHECK(ctx);
	TB_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
	return 0;
}

extern "C" int tb200_upload_geometry(
	tb200_ctx * ctx, int patch_index, const tb200_geometry * gh
) {
	if (!ctx->committed) TB_FAIL(ctx, "commit the layout first");
	PatchInfo * pi = find_patch(ctx, patch_index);
	if (pi == 0 || pi->elem0 < 0) TB_FAIL(ctx, "not a local patch");
	const int L = ctx->lay.nlev;
	double ** g2 = ctx->g2d;
	double ** gn = ctx->g3n;
	double ** ge = ctx->g3e;
	if (upload_geom_array(ctx, *pi, gh->jacobian2d, 1, 1, g2[0], 0, 0)) return 1;
	if (upload_geom_array(ctx, *pi, gh->contrametric2da, 1, 2, g2[1], g2[2], 0)) return 1;
	if (upload_geom_array(ctx, *pi, gh->contrametric2db, 1, 2, g2[3], g2[4], 0)) return 1;
	if (upload_geom_array(ctx, *pi, gh->coriolis, 1, 1, g2[5], 0, 0)) return 1;
	if (upload_geom_array(ctx, *pi, gh->topography, 1, 1, g2[6], 0, 0)) return 1;
	if (gh->jacobian != 0 && ensure_geom3d(ctx, 0, 1, false)) return 1;
	if (gh->jacobian_redge != 0 && ensure_geom3d(ctx, 0, 1, true)) return 1;
	if (upload_geom_array(ctx, *pi, gh->jacobian, L, 1, gn[0], 0, 0)) return 1;
	if (upload_geom_array(ctx, *pi, gh->jacobian_redge, L + 1, 1, ge[0], 0, 0)) return 1;
	if (ctx->cfg.eqn_type == TB200_EQN_PRIMITIVE_NONHYDRO) {
		if (gh->contrametrica != 0 && ensure_geom3d(ctx, 1, 3, false)) return 1;
		if (gh->contrametricb != 0 && ensure_geom3d(ctx, 4, 3, false)) return 1;
		if (gh->contrametricxi != 0 && ensure_geom3d(ctx, 7, 3, false)) return 1;
		if (gh->derivr_node != 0 && ensure_geom3d(ctx, 10, 3, false)) return 1;
		if (gh->contrametrica_redge != 0 && ensure_geom3d(ctx, 1, 3, true)) return 1;
		if (gh->contrametricb_redge != 0 && ensure_geom3d(ctx, 4, 3, true)) return 1;
		if (gh->contrametricxi_redge != 0 && ensure_geom3d(ctx, 7, 3, true)) return 1;
		if (gh->derivr_redge != 0 && ensure_geom3d(ctx, 10, 3, true)) return 1;
		if (upload_geom_array(ctx, *pi, gh->contrametrica, L, 3, gn[1], gn[2], gn[3])) return 1;
		if (upload_geom_array(ctx, *pi, gh->contrametricb, L, 3, gn[4], gn[5], gn[6])) return 1;
		if (upload_geom_array(ctx, *pi, gh->contrametricxi, L, 3, gn[7], gn[8], gn[9])) return 1;
		if (upload_geom_array(ctx, *pi, gh->derivr_node, L, 3, gn[10], gn[11], gn[12])) return 1;
		if (upload_geom_array(ctx, *pi, gh->contrametrica_redge, L + 1, 3, ge[1], ge[2], ge[3])) return 1;
		if (upload_geom_array(ctx, *pi, gh->contrametricb_redge, L + 1, 3, ge[4], ge[5], ge[6])) return 1;
		if (upload_geom_array(ctx, *pi, gh->contrametricxi_redge, L + 1, 3, ge[7], ge[8], ge[9])) return 1;
		if (upload_geom_array(ctx, *pi, gh->derivr_redge, L + 1, 3, ge[10], ge[11], ge[12])) return 1;
		if (gh->contrametrica != 0 && gh->contrametricxi_redge != 0 && gh->derivr_node != 0) {
			ctx->geometry3d_uploaded = true;
		}
	}
	ctx->fast_state = 0;
	return 0;
}

// Terrain-following cubed-sphere metric evaluated on the fly
// (GridPatchCSGLL::EvaluateGeometricTerms, GridPatchCSGLL.cpp:344-553): the
// gnomonic node coordinates m_dXNode / m_dYNode (= tan of GetANode / GetBNode,
// GridPatchCSGLL.cpp:205-213) and GridPatch::GetTopographyDeriv().
extern "C" int tb200_set_terrain_metric(
	tb200_ctx * ctx, int patch_index, const double * xnode, const double * ynode,
	const double * topography_deriv
) {
	if (!ctx->committed) TB_FAIL(ctx, "commit the layout first");
	PatchInfo * pi = find_patch(ctx, patch_index);
	if (pi == 0 || pi->elem0 < 0) TB_FAIL(ctx, "not a local patch");
	const int np = ctx->lay.np;
	const size_t nn = ctx->lay.nn;
	const size_t wa = pi->nea * np + 2 * pi->halo;
	const size_t wb = pi->neb * np + 2 * pi->halo;
	if (ctx->d_tx == 0) {
		if (dalloc(ctx, &ctx->d_tx, (size_t)ctx->lay.nelem * nn)) return 1;
		if (dalloc(ctx, &ctx->d_ty, (size_t)ctx->lay.nelem * nn)) return 1;
		if (dalloc(ctx, &ctx->d_tda, (size_t)ctx->lay.nelem * nn)) return 1;
		if (dalloc(ctx, &ctx->d_tdb, (size_t)ctx->lay.nelem * nn)) return 1;
	}
	std::vector<double> hx(wa * wb), hy(wa * wb);
	for (size_t i = 0; i < wa; i++) {
		for (size_t j = 0; j < wb; j++) {
			hx[i * wb + j] = xnode[i];
			hy[i * wb + j] = ynode[j];
		}
	}
	if (upload_geom_array(ctx, *pi, hx.data(), 1, 1, ctx->d_tx, 0, 0)) return 1;
	if (upload_geom_array(ctx, *pi, hy.data(), 1, 1, ctx->d_ty, 0, 0)) return 1;
	if (upload_geom_array(ctx, *pi, topography_deriv, 1, 2, ctx->d_tda, ctx->d_tdb, 0)) return 1;
	pi->has_terrain = true;
	return 0;
}

// Grid::GetREtaLevels / GetREtaInterfaces; enables the on-the-fly metric once
// every local patch has its terrain arrays (TB200_METRIC=stored disables it).
extern "C" int tb200_set_vertical_coordinate(
	tb200_ctx * ctx, const double * reta_levels, const double * reta_interfaces
) {
	if (!ctx->committed) TB_FAIL(ctx, "commit the layout first");
	const int L = ctx->lay.nlev;
	std::vector<double> a(reta_levels, reta_levels + L), b(reta_interfaces, reta_interfaces + L + 1);
	ctx->reta_n_h = a;
	ctx->reta_e_h = b;
	ctx->fast_state = 0;
	if (dupload(ctx, &ctx->d_reta_n, a)) return 1;
	if (dupload(ctx, &ctx->d_reta_e, b)) return 1;
	bool all = true;
	for (size_t p = 0; p < ctx->patches.size(); p++) {
		if (ctx->patches[p].elem0 >= 0 && !ctx->patches[p].has_terrain) all = false;
	}
	const char * force = getenv("TB200_METRIC");
	if (force != 0 && strcmp(force, "stored") == 0) all = false;
	DevGeom & g = ctx->geom;
	g.analytic = all ? 1 : 0;
	g.tx = ctx->d_tx; g.ty = ctx->d_ty; g.tda = ctx->d_tda; g.tdb = ctx->d_tdb;
	g.reta_n = ctx->d_reta_n; g.reta_e = ctx->d_reta_e;
	g.ztop = ctx->cfg.ztop;
	g.radius = ctx->cfg.earth_radius;
	return 0;
}

extern "C" int tb200_upload_element_area(
	tb200_ctx * ctx, int patch_index, const double * area_node, const double * area_redge
) {
	if (!ctx->committed) TB_FAIL(ctx, "commit the layout first");
	PatchInfo * pi = find_patch(ctx, patch_index);
	if (pi == 0 || pi->elem0 < 0) TB_FAIL(ctx, "not a local patch");
	const int L = ctx->lay.nlev;
	const size_t nn = ctx->lay.nn;
	if (ctx->d_area_node == 0) {
		if (dalloc(ctx, &ctx->d_area_node, (size_t)ctx->lay.nelem * L * nn)) return 1;
		if (dalloc(ctx, &ctx->d_area_redge, (size_t)ctx->lay.nelem * (L + 1) * nn)) return 1;
	}
	if (upload_geom_array(ctx, *pi, area_node, L, 1, ctx->d_area_node, 0, 0)) return 1;
	if (upload_geom_array(ctx, *pi, area_redge, L + 1, 1, ctx->d_area_redge, 0, 0)) return 1;
	return 0;
}

extern "C" int tb200_upload_state(
	tb200_ctx * ctx, int patch_index, int inst,
	const double * node, const double * redge, const double * tracers);

// Rayleigh friction (GridPatch::GetRayleighStrength, GetReferenceState): the
// strength on levels / interfaces [W_A][W_B][L(+1)] and the reference state in
// the layout of a state instance.  Optional; StepAfterSubCycle applies the
// friction once any patch has a non-zero strength (Grid::HasRayleighFriction).
// zero reference state until one is uploaded (a test case without one,
// TestCase::HasReferenceState, leaves the reference's arrays zero as well)
static int refstate_alloc(tb200_ctx * ctx) {
	if (ctx->d_refstate != 0) return 0;
	const DevLayout & lay = ctx->lay;
	const size_t n = (size_t)lay.nelem * lay.nrows * lay.nn;
	if (dalloc(ctx, &ctx->d_refstate, n)) return 1;
	TB_CHECK(ctx, cudaMemset(ctx->d_refstate, 0, n * sizeof(double)));
	return 0;
}

// GridPatch::GetReferenceState(Node / REdge) of a local patch, in the state layout
extern "C" int tb200_upload_reference_state(
	tb200_ctx * ctx, int patch_index, const double * ref_node, const double * ref_redge
) {
	if (!ctx->committed) TB_FAIL(ctx, "commit the layout first");
	PatchInfo * pi = find_patch(ctx, patch_index);
	if (pi == 0 || pi->elem0 < 0) TB_FAIL(ctx, "not a local patch");
	if (refstate_alloc(ctx)) return 1;
	// through the state path into a scratch "instance"
	ctx->inst.push_back(ctx->d_refstate);
	const int slot = (int)ctx->inst.size() - 1;
	const int rc = tb200_upload_state(ctx, patch_index, slot, ref_node, ref_redge, 0);
	ctx->inst.pop_back();
	return rc;
}

// --vdisc FV (Grid::VerticalDiscretization_FiniteVolume): the column operators the
// caller supplies are the finite-volume ones (even orders only,
// LinearColumnOperatorFEM.cpp:227); here only the counts change - every level is
// its own element for the penalty terms (VerticalDynamicsFEM.cpp:646-650, 2649-2654)
// and the declared Jacobian band is narrower (:174-185).  The workspace was sized
// for the finite-element band at tb200_create, which is the wider one.
extern "C" int tb200_set_vertical_discretization(tb200_ctx * ctx, int finite_volume) {
	int fe_offd = 0;
	switch (ctx->cfg.vertical_order) {
		case 1: fe_offd = 4; break;
		case 2: fe_offd = 9; break;
		case 3: fe_offd = 15; break;
		case 4: fe_offd = 22; break;
		case 5: fe_offd = 30; break;
		default: TB_FAIL(ctx, "unsupported vertical order");
	}
	if (!finite_volume) {
		ctx->finite_volume = 0;
		ctx->fe_nodes = ctx->cfg.vertical_order;
		ctx->offd = fe_offd;
	} else {
		int offd = 0;
		if (ctx->cfg.vertical_order <= 2) offd = 4;
		else if (ctx->cfg.vertical_order == 4) offd = 7;
		else if (ctx->cfg.vertical_order == 6) offd = 10;
		else TB_FAIL(ctx, "UNIMPLEMENTED: At this vertical order");
		if (offd > fe_offd) TB_FAIL(ctx, "column workspace too small for this band");
		ctx->finite_volume = 1;
		ctx->fe_nodes = 1;
		ctx->offd = offd;
	}
	ctx->fast_state = 0;
	return 0;
}

extern "C" int tb200_set_mass_flux_on_levels(tb200_ctx * ctx, int on) {
	ctx->mass_flux_levels = on ? 1 : 0;
	ctx->fast_state = 0;        // the column-constant path declines (fast_prepare)
	return 0;
}

// Grid::HasUniformDiffusion with TestCase::GetUniformDiffusionCoeffs (Grid.cpp:399-415)
extern "C" int tb200_set_uniform_diffusion(
	tb200_ctx * ctx, double scalar_coeff, double vector_coeff
) {
	ctx->uniform_s = scalar_coeff;
	ctx->uniform_v = vector_coeff;
	ctx->fast_state = 0;        // the column-constant path declines (fast_prepare)
	return 0;
}

static inline bool uniform_on(const tb200_ctx * ctx) {
	return ctx->uniform_s != 0.0 || ctx->uniform_v != 0.0;
}

extern "C" int tb200_upload_rayleigh(
	tb200_ctx * ctx, int patch_index,
	const double * strength_node, const double * strength_redge,
	const double * ref_node, const double * ref_redge
) {
	if (!ctx->committed) TB_FAIL(ctx, "commit the layout first");
	PatchInfo * pi = find_patch(ctx, patch_index);
	if (pi == 0 || pi->elem0 < 0) TB_FAIL(ctx, "not a local patch");
	const DevLayout & lay = ctx->lay;
	const int L = lay.nlev;
	const size_t nn = lay.nn;
	if (ctx->d_ray_node == 0) {
		if (dalloc(ctx, &ctx->d_ray_node, (size_t)lay.nelem * L * nn)) return 1;
		if (dalloc(ctx, &ctx->d_ray_redge, (size_t)lay.nelem * (L + 1) * nn)) return 1;
		TB_CHECK(ctx, cudaMemset(ctx->d_ray_node, 0, (size_t)lay.nelem * L * nn * sizeof(double)));
		TB_CHECK(ctx, cudaMemset(ctx->d_ray_redge, 0, (size_t)lay.nelem * (L + 1) * nn * sizeof(double)));
	}
	if (upload_geom_array(ctx, *pi, strength_node, L, 1, ctx->d_ray_node, 0, 0)) return 1;
	if (upload_geom_array(ctx, *pi, strength_redge, L + 1, 1, ctx->d_ray_redge, 0, 0)) return 1;
	if (tb200_upload_reference_state(ctx, patch_index, ref_node, ref_redge)) return 1;
	ctx->has_rayleigh = true;
	return 0;
}

///////////////////////////////////////////////////////////////////////////////
// Device-side set-up (tb200_setup.cuh)

// 2-D metric, Coriolis parameter, longitude and latitude of a cubed-sphere patch
// from X = tan(alpha), Y = tan(beta) (tb200_set_terrain_metric): what
// GridPatchCSGLL::EvaluateGeometricTerms leaves in GetJacobian2D(),
// GetContraMetric2DA/B(), GetCoriolisF(), GetLongitude(), GetLatitude().
extern "C" int tb200_evaluate_geometry_cs(
	tb200_ctx * ctx, int patch_index, double radius, double omega
) {
	if (!ctx->committed) TB_FAIL(ctx, "commit the layout first");
	PatchInfo * pi = find_patch(ctx, patch_index);
	if (pi == 0 || pi->elem0 < 0) TB_FAIL(ctx, "not a local patch");
	if (ctx->d_tx == 0) TB_FAIL(ctx, "node coordinates not set (tb200_set_terrain_metric)");
	const DevLayout & lay = ctx->lay;
	if (ctx->d_hs_lat == 0) {
		if (dalloc(ctx, &ctx->d_hs_lat, (size_t)lay.nelem * lay.nn)) return 1;
		if (dalloc(ctx, &ctx->d_hs_sp, (size_t)lay.nelem * lay.nn)) return 1;
	}
	if (ctx->d_lon == 0) {
		if (dalloc(ctx, &ctx->d_lon, (size_t)lay.nelem * lay.nn)) return 1;
	}
	CsGeomOut o;
	o.j2d = ctx->g2d[0]; o.a0 = ctx->g2d[1]; o.a1 = ctx->g2d[2];
	o.b0 = ctx->g2d[3]; o.b1 = ctx->g2d[4]; o.coriolis = ctx->g2d[5];
	o.lon = ctx->d_lon; o.lat = ctx->d_hs_lat;
	const long long nelem = (long long)pi->nea * pi->neb;
	const long long total = nelem * lay.nn;
	auto kfn = k_cs_geometry_2d;
	TB_LAUNCH_FLAT(kfn, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, ctx->stream,
		lay.nn, (long long)pi->elem0, nelem, pi->panel, radius, omega,
		(const double *)ctx->d_tx, (const double *)ctx->d_ty, o);
	TB_KERNEL_CHECK(ctx);
	pi->has_lonlat = true;
	return 0;
}

// Read-back of a per-column array in the device's element-major order
// [element][np * np] (tests of the device-side set-up): 0 Jacobian2D,
// 1, 2 ContraMetric2DA, 3, 4 ContraMetric2DB, 5 Coriolis, 6 topography,
// 7 longitude, 8 latitude, 9 accumulated precipitation (Kessler).
extern "C" int tb200_debug_column_field(tb200_ctx * ctx, int which, double * out) {
	const DevLayout & lay = ctx->lay;
	const double * src = 0;
	if (which >= 0 && which < 7) src = ctx->g2d[which];
	else if (which == 7) src = ctx->d_lon;
	else if (which == 8) src = ctx->d_hs_lat;
	else if (which == 9) src = ctx->d_precip;
	if (src == 0) TB_FAIL(ctx, "column field not available");
	TB_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
	TB_CHECK(ctx, cudaMemcpy(out, src, (size_t)lay.nelem * lay.nn * sizeof(double), cudaMemcpyDeviceToHost));
	return 0;
}

static JWParams jw_params(const tb200_ctx * ctx, const tb200_jw_test * t) {
	JWParams P;
	P.eta0 = t->eta0; P.tropopause_eta = t->tropopause_eta; P.t0 = t->t0;
	P.delta_t = t->delta_t; P.lapse_rate = t->lapse_rate; P.u0 = t->u0; P.up = t->up;
	P.pert_lon = t->pert_lon; P.pert_lat = t->pert_lat; P.pert_r = t->pert_r;
	P.perturbation = t->perturbation;
	P.g = ctx->cfg.g; P.R = ctx->cfg.R; P.p0 = ctx->cfg.p0;
	P.omega = t->omega; P.radius = t->radius;
	// PhysicalConstants.h:361-375
	P.gamma = ctx->cfg.cp / (ctx->cfg.cp - ctx->cfg.R);
	P.pressure_scaling = ctx->cfg.p0 * pow(ctx->cfg.R / ctx->cfg.p0, P.gamma);
	P.ztop = ctx->cfg.ztop;
	return P;
}

// BaroclinicWaveJWTest::EvaluateTopography on the nodes of a patch (into the
// topography array of the geometry); needs tb200_evaluate_geometry_cs.
extern "C" int tb200_evaluate_jw_topography(
	tb200_ctx * ctx, int patch_index, const tb200_jw_test * test
) {
	PatchInfo * pi = find_patch(ctx, patch_index);
	if (pi == 0 || pi->elem0 < 0) TB_FAIL(ctx, "not a local patch");
	if (!pi->has_lonlat) TB_FAIL(ctx, "longitude / latitude not evaluated (tb200_evaluate_geometry_cs)");
	const DevLayout & lay = ctx->lay;
	const long long nelem = (long long)pi->nea * pi->neb;
	const long long total = nelem * lay.nn;
	auto kfn = k_jw_topography;
	TB_LAUNCH_FLAT(kfn, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, ctx->stream,
		lay.nn, (long long)pi->elem0, nelem, jw_params(ctx, test),
		(const double *)ctx->d_hs_lat, ctx->g2d[6]);
	TB_KERNEL_CHECK(ctx);
	return 0;
}

// GridPatchCSGLL::EvaluateTestCase for BaroclinicWaveJWTest: the initial state of
// a patch written straight into instance `inst` (tracers are not touched).
extern "C" int tb200_evaluate_jw_state(
	tb200_ctx * ctx, int patch_index, int inst, const tb200_jw_test * test
) {
	PatchInfo * pi = find_patch(ctx, patch_index);
	if (pi == 0 || pi->elem0 < 0) TB_FAIL(ctx, "not a local patch");
	if (inst < 0 || inst >= (int)ctx->inst.size()) TB_FAIL(ctx, "invalid instance");
	if (ctx->cfg.eqn_type != TB200_EQN_PRIMITIVE_NONHYDRO) TB_FAIL(ctx, "the JW test needs the nonhydrostatic equation set");
	if (!pi->has_lonlat) TB_FAIL(ctx, "longitude / latitude not evaluated (tb200_evaluate_geometry_cs)");
	if (ctx->d_reta_n == 0) TB_FAIL(ctx, "vertical coordinate not set (tb200_set_vertical_coordinate)");
	const DevLayout & lay = ctx->lay;
	const long long nelem = (long long)pi->nea * pi->neb;
	const long long total = nelem * (long long)(lay.nlev + 1) * lay.nn;
	long long nb = (total + 127) / 128;
	if (nb > 148 * 32) nb = 148 * 32;
	TB_CHECK(ctx, cudaMemset(ctx->d_info + 2, 0, sizeof(int)));
	auto kfn = k_jw_state;
	TB_LAUNCH_FLAT(kfn, dim3((unsigned)nb), dim3(128), 0, ctx->stream,
		lay, (long long)pi->elem0, nelem, pi->panel, jw_params(ctx, test),
		(const double *)ctx->d_tx, (const double *)ctx->d_ty,
		(const double *)ctx->d_lon, (const double *)ctx->d_hs_lat,
		(const double *)ctx->g2d[6], (const double *)ctx->d_reta_n,
		ctx->inst[inst], ctx->d_info + 2);
	TB_KERNEL_CHECK(ctx);
	int failed = 0;
	TB_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
	TB_CHECK(ctx, cudaMemcpy(&failed, ctx->d_info + 2, sizeof(int), cudaMemcpyDeviceToHost));
	// BaroclinicWaveJWTest.cpp:337-339
	if (failed != 0) TB_FAIL(ctx, "Maximum number of iterations exceeded.");
	return 0;
}

///////////////////////////////////////////////////////////////////////////////
// Output-side interpolation (tb200_output.cuh)

namespace {
struct DevTemp {
	std::vector<void *> p;
	~DevTemp() { for (size_t q = 0; q < p.size(); q++) cudaFree(p[q]); }
	template <typename T> T * up(tb200_ctx * ctx, const T * host, size_t n, bool * ok) {
		void * d = 0;
		if (cudaMalloc(&d, std::max<size_t>(n, 1) * sizeof(T)) != cudaSuccess) { *ok = false; return 0; }
		p.push_back(d);
		if (host != 0 && n > 0
			&& cudaMemcpy(d, host, n * sizeof(T), cudaMemcpyHostToDevice) != cudaSuccess) { *ok = false; }
		(void)ctx;
		return (T *)d;
	}
};
}

// Derived output fields of an instance, kept on the device (tb200_output.cuh):
// GridGLL::ComputeVorticityDivergence (GridGLL.cpp:587-601) = ComputeCurlAndDiv on
// every patch + ApplyDSS of vorticity and divergence; Grid::ComputeTemperature.
static int dss_rows(tb200_ctx * ctx, int inst, int row0, int row1, bool is_state, bool remainder_only);

extern "C" int tb200_compute_output_fields(tb200_ctx * ctx, int inst) {
	if (!ctx->committed) TB_FAIL(ctx, "commit the layout first");
	if (inst < 0 || inst >= (int)ctx->inst.size()) TB_FAIL(ctx, "invalid instance");
	if (!ctx->connectivity_built) TB_FAIL(ctx, "connectivity not built");
	const DevLayout & lay = ctx->lay;
	if (lay.ncomp < 2) TB_FAIL(ctx, "Insufficient components for vorticity calculation");
	const int L = lay.nlev;
	if (3 * L > lay.nrows && ctx->nranks > 1) {
		TB_FAIL(ctx, "output fields: exchange buffers too small for this layout");
	}
	const size_t n = (size_t)lay.nelem * 3 * L * lay.nn;
	if (ctx->d_outfield == 0 && dalloc(ctx, &ctx->d_outfield, n)) return 1;
	if (ctx->cfg.eqn_type == TB200_EQN_PRIMITIVE_NONHYDRO) {
		for (int c = 2; c < 5; c++) {
			if (c != 3 && lay.onedge[c]) TB_FAIL(ctx, "output temperature: rho theta and rho on levels only");
		}
		const double gamma = ctx->cfg.cp / (ctx->cfg.cp - ctx->cfg.R);
		const double pscale = ctx->cfg.p0 * pow(ctx->cfg.R / ctx->cfg.p0, gamma);
		long long nb = ((long long)lay.nelem * L * lay.nn + 255) / 256;
		if (nb > 148 * 16) nb = 148 * 16;
		auto kfn = k_output_temperature;
		TB_LAUNCH_FLAT(kfn, dim3((unsigned)nb), dim3(256), 0, ctx->stream,
			lay, pscale, gamma, ctx->cfg.R, (const double *)ctx->inst[inst], ctx->d_outfield);
		TB_KERNEL_CHECK(ctx);
	}
	{
		const long long nitems = lay.nelem * L;
		TB_NP_SWITCH(lay.np,
			auto kfn = k_output_curl_div<NPV, kItems>;
			TB_LAUNCH(kfn, dim3((unsigned)((nitems + kItems - 1) / kItems)), dim3(NPV * NPV * kItems), 0,
				ctx->stream, lay, ctx->geom, ctx->tables, (const double *)ctx->inst[inst],
				ctx->d_outfield);)
		TB_KERNEL_CHECK(ctx);
	}
	// ApplyDSS(0, DataType_Vorticity / DataType_Divergence): the field array is laid
	// out like an instance of 3 L rows per element, so the averaging (and exchange)
	// kernels run on it with that row count; scalars, no seam re-basing
	const DevLayout saved = ctx->lay;
	ctx->lay.nrows = 3 * L;
	ctx->lay.nrows_state = 3 * L;
	ctx->inst.push_back(ctx->d_outfield);
	const int slot = (int)ctx->inst.size() - 1;
	const int rc = dss_rows(ctx, slot, 0, 2 * L, false, false);
	ctx->inst.pop_back();
	ctx->lay = saved;
	return rc;
}

// Grid::ReduceInterpolate (Grid.cpp:866-990) / GridPatchCSGLL::InterpolateData
// (GridPatchCSGLL.cpp:1365-1780) of instance `inst`: data_type TB200_DATA_STATE or
// TB200_DATA_TRACERS; only_location -1 all components, 0 those on levels, 1 those
// on interfaces (eOnlyVariablesAt); per point its patch, element (elem_a, elem_b:
// element indices inside the patch), the np Lagrangian coefficients along alpha
// and beta of that element's GLL nodes, and its alpha, beta (primitive wind);
// per location the column operator (dense [nout][nin] with [begin, end) windows;
// a null operator = identity, nout = nin).  out: [ncomp or ntracers][nout][npts],
// zero for points of patches that are not local (the reference sums over ranks)
// and for components that were not asked for.
extern "C" int tb200_interpolate(
	tb200_ctx * ctx, int inst, int data_type, int only_location,
	int npts, const int * patch_index, const int * elem_a, const int * elem_b,
	const double * ca, const double * cb, const double * alpha, const double * beta,
	int nout,
	const double * vop_node, const int * vbegin_node, const int * vend_node,
	const double * vop_redge, const int * vbegin_redge, const int * vend_redge,
	int convert_to_primitive, double * out
) {
	if (!ctx->committed) TB_FAIL(ctx, "commit the layout first");
	if (inst < 0 || inst >= (int)ctx->inst.size()) TB_FAIL(ctx, "invalid instance");
	const DevLayout & lay = ctx->lay;
	const bool tracers = (data_type == TB200_DATA_TRACERS);
	// derived fields (tb200_compute_output_fields): one component on levels
	int derived = -1;
	if (data_type == TB200_DATA_TEMPERATURE) derived = 0;
	if (data_type == TB200_DATA_VORTICITY) derived = 1;
	if (data_type == TB200_DATA_DIVERGENCE) derived = 2;
	if (derived >= 0 && ctx->d_outfield == 0) {
		TB_FAIL(ctx, "output fields not computed (tb200_compute_output_fields)");
	}
	if (!tracers && derived < 0 && data_type != TB200_DATA_STATE) TB_FAIL(ctx, "Invalid DataType");
	const int ncomp = (derived >= 0) ? 1 : (tracers ? lay.ntr : lay.ncomp);
	if (npts <= 0 || nout <= 0 || ncomp == 0) return 0;
	if (lay.nlev + 1 > TB_INTERP_MAXLEV) TB_FAIL(ctx, "too many levels for the interpolation kernel");
	const int np = lay.np;
	std::vector<int> elem(npts), panel(npts);
	for (int i = 0; i < npts; i++) {
		PatchInfo * pi = find_patch(ctx, patch_index[i]);
		elem[i] = -1;
		panel[i] = 0;
		if (pi == 0) TB_FAIL(ctx, "unknown patch");
		panel[i] = pi->panel;
		if (pi->elem0 < 0) continue;
		if (elem_a[i] < 0 || elem_a[i] >= pi->nea || elem_b[i] < 0 || elem_b[i] >= pi->neb) {
			TB_FAIL(ctx, "Point out of range");          // GridPatchCSGLL.cpp:1596-1602
		}
		elem[i] = (int)(pi->elem0 + (long long)elem_a[i] * pi->neb + elem_b[i]);
	}
	// identity operators when none is given
	std::vector<double> idn, ide;
	std::vector<int> bn, en, be, ee;
	const int L = lay.nlev;
	auto identity = [&](int n, std::vector<double> & m, std::vector<int> & b, std::vector<int> & e) {
		m.assign((size_t)n * n, 0.0); b.resize(n); e.resize(n);
		for (int q = 0; q < n; q++) { m[(size_t)q * n + q] = 1.0; b[q] = q; e[q] = q + 1; }
	};
	if (vop_node == 0) {
		if (nout != L && !tracers) { /* checked per component below */ }
		identity(L, idn, bn, en);
	}
	if (vop_redge == 0) identity(L + 1, ide, be, ee);
	bool ok = true;
	DevTemp tmp;
	int * d_elem = tmp.up(ctx, elem.data(), (size_t)npts, &ok);
	int * d_panel = tmp.up(ctx, panel.data(), (size_t)npts, &ok);
	double * d_ca = tmp.up(ctx, ca, (size_t)npts * np, &ok);
	double * d_cb = tmp.up(ctx, cb, (size_t)npts * np, &ok);
	double * d_alpha = tmp.up(ctx, alpha, (size_t)npts, &ok);
	double * d_beta = tmp.up(ctx, beta, (size_t)npts, &ok);
	const int nout_n = (vop_node != 0) ? nout : L;
	const int nout_e = (vop_redge != 0) ? nout : (L + 1);
	double * d_vn = tmp.up(ctx, vop_node != 0 ? vop_node : idn.data(), (size_t)nout_n * L, &ok);
	int * d_bn = tmp.up(ctx, vop_node != 0 ? vbegin_node : bn.data(), (size_t)nout_n, &ok);
	int * d_en = tmp.up(ctx, vop_node != 0 ? vend_node : en.data(), (size_t)nout_n, &ok);
	double * d_ve = tmp.up(ctx, vop_redge != 0 ? vop_redge : ide.data(), (size_t)nout_e * (L + 1), &ok);
	int * d_be = tmp.up(ctx, vop_redge != 0 ? vbegin_redge : be.data(), (size_t)nout_e, &ok);
	int * d_ee = tmp.up(ctx, vop_redge != 0 ? vend_redge : ee.data(), (size_t)nout_e, &ok);
	const size_t per_comp = (size_t)nout * npts;
	double * d_out = tmp.up(ctx, (const double *)0, per_comp * ncomp, &ok);
	if (!ok) TB_FAIL(ctx, "interpolation: device allocation failed");
	TB_CHECK(ctx, cudaMemset(d_out, 0, per_comp * ncomp * sizeof(double)));
	TB_CHECK(ctx, cudaStreamSynchronize(ctx->stream));

	const bool on_levels_only = tracers || (derived >= 0);
	for (int c = 0; c < ncomp; c++) {
		const int onedge = on_levels_only ? 0 : lay.onedge[c];
		if (!on_levels_only && only_location >= 0 && only_location != onedge) continue;
		if ((onedge ? nout_e : nout_n) != nout) {
			TB_FAIL(ctx, "InterpData dimension mismatch (1)");      // Grid.cpp:938-940
		}
		InterpArgs a;
		a.npts = npts; a.nout = nout;
		a.elem = d_elem; a.ca = d_ca; a.cb = d_cb;
		a.vop = onedge ? d_ve : d_vn;
		a.vbegin = onedge ? d_be : d_bn;
		a.vend = onedge ? d_ee : d_en;
		a.nin = onedge ? (L + 1) : L;
		// derived array: rows 0.. vorticity, L.. divergence, 2L.. temperature
		const int drow[3] = {2 * L, 0, L};
		a.row0 = (derived >= 0) ? drow[derived] : (tracers ? (lay.troff + c * L) : lay.rowoff[c]);
		a.estride = (derived >= 0) ? 3 * L : lay.nrows;
		// w -> primitive (:1655-1690): divided by DerivR[2] at the element's first node
		a.divide_derivr = (!on_levels_only && convert_to_primitive && c == 3
			&& ctx->cfg.eqn_type == TB200_EQN_PRIMITIVE_NONHYDRO) ? 1 : 0;
		a.derivr = onedge ? ctx->geom.dre[2] : ctx->geom.dr[2];
		a.zs = ctx->g2d[6];
		a.ztop = ctx->cfg.ztop;
		a.out = d_out + per_comp * c;
		auto kfn = k_interpolate_component;
		TB_LAUNCH_FLAT(kfn, dim3((npts + 63) / 64), dim3(64), 0, ctx->stream,
			lay, a, (derived >= 0) ? (const double *)ctx->d_outfield
			                       : (const double *)ctx->inst[inst]);
		TB_KERNEL_CHECK(ctx);
	}
	if (!on_levels_only && convert_to_primitive && lay.ncomp >= 2
		&& (only_location < 0 || only_location == lay.onedge[0])) {
		const long long total = (long long)npts * nout;
		auto kfn = k_interpolate_wind;
		TB_LAUNCH_FLAT(kfn, dim3((unsigned)((total + 127) / 128)), dim3(128), 0, ctx->stream,
			npts, nout, (const int *)d_elem, (const int *)d_panel,
			(const double *)d_alpha, (const double *)d_beta, ctx->cfg.earth_radius,
			d_out, d_out + per_comp);
		TB_KERNEL_CHECK(ctx);
	}
	TB_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
	TB_CHECK(ctx, cudaMemcpy(out, d_out, per_comp * ncomp * sizeof(double), cudaMemcpyDeviceToHost));
	return 0;
}

///////////////////////////////////////////////////////////////////////////////
// Column physics (tb200_physics.cuh)

// Per-column inputs of HeldSuarezPhysics::Perform (HeldSuarezPhysics.cpp:95-115):
// GridPatch::GetLatitude() and the product dataREdge[R][0] * dataREdge[T][0] of
// instance 0, both [iA][iB] in the reference's layout with halo.
extern "C" int tb200_upload_held_suarez(
	tb200_ctx * ctx, int patch_index, const double * latitude, const double * surface_product
) {
	if (!ctx->committed) TB_FAIL(ctx, "commit the layout first");
	PatchInfo * pi = find_patch(ctx, patch_index);
	if (pi == 0 || pi->elem0 < 0) TB_FAIL(ctx, "not a local patch");
	if (latitude == 0 || surface_product == 0) TB_FAIL(ctx, "Held-Suarez: latitude and surface product needed");
	const DevLayout & lay = ctx->lay;
	if (ctx->d_hs_lat == 0) {
		if (dalloc(ctx, &ctx->d_hs_lat, (size_t)lay.nelem * lay.nn)) return 1;
		if (dalloc(ctx, &ctx->d_hs_sp, (size_t)lay.nelem * lay.nn)) return 1;
	}
	if (upload_geom_array(ctx, *pi, latitude, 1, 1, ctx->d_hs_lat, 0, 0)) return 1;
	if (upload_geom_array(ctx, *pi, surface_product, 1, 1, ctx->d_hs_sp, 0, 0)) return 1;
	pi->has_held_suarez = true;
	return 0;
}

// HeldSuarezPhysics::Perform on instance 0 (the reference's workflow processes
// always act on instance 0, Model.cpp:477-481)
extern "C" int tb200_held_suarez(tb200_ctx * ctx, double dt) {
	const DevLayout & lay = ctx->lay;
	if (ctx->cfg.eqn_type != TB200_EQN_PRIMITIVE_NONHYDRO) {
		TB_FAIL(ctx, "Held-Suarez physics needs the nonhydrostatic equation set");
	}
	for (size_t p = 0; p < ctx->patches.size(); p++) {
		// (the latitude array alone may exist from tb200_evaluate_geometry_cs)
		if (ctx->patches[p].elem0 >= 0 && !ctx->patches[p].has_held_suarez) {
			TB_FAIL(ctx, "Held-Suarez inputs not uploaded (tb200_upload_held_suarez)");
		}
	}
	for (int c = 0; c < 5; c++) {
		if (c != 3 && lay.onedge[c]) TB_FAIL(ctx, "Held-Suarez physics: Lorenz staggering only");
	}
	HeldSuarezArgs a;
	a.latitude = ctx->d_hs_lat;
	a.surface_product = ctx->d_hs_sp;
	a.dt = dt;
	// PhysicalConstants.h:361-375
	a.kappa = ctx->cfg.R / ctx->cfg.cp;
	a.gamma = ctx->cfg.cp / (ctx->cfg.cp - ctx->cfg.R);
	a.pressure_scaling = ctx->cfg.p0 * pow(ctx->cfg.R / ctx->cfg.p0, a.gamma);
	a.R = ctx->cfg.R;
	a.p0 = ctx->cfg.p0;
	const long long total = lay.nelem * (long long)lay.nlev * lay.nn;
	long long nb = (total + 255) / 256;
	if (nb > 148 * 16) nb = 148 * 16;
	auto kfn = k_held_suarez;
	TB_LAUNCH_FLAT(kfn, dim3((unsigned)nb), dim3(256), 0, ctx->stream, lay, a, ctx->inst[0]);
	TB_KERNEL_CHECK(ctx);
	return 0;
}

// KesslerPhysics::Perform on instance 0 (state and the first three tracers:
// rho qv, rho qc, rho qr); accumulates the precipitation per column
// (GridPatch::GetUserData2D()[0], read back with tb200_debug_column_field(9)).
extern "C" int tb200_kessler(tb200_ctx * ctx, double dt) {
	const DevLayout & lay = ctx->lay;
	if (ctx->cfg.eqn_type != TB200_EQN_PRIMITIVE_NONHYDRO) {
		TB_FAIL(ctx, "Kessler physics needs the nonhydrostatic equation set");
	}
	if (lay.ntr < 3) TB_FAIL(ctx, "Kessler physics needs three tracers (rho qv, rho qc, rho qr)");
	if (lay.nlev < 2) TB_FAIL(ctx, "Kessler physics needs at least two levels");
	for (int c = 0; c < 5; c++) {
		if (c != 3 && lay.onedge[c]) TB_FAIL(ctx, "Kessler physics: Lorenz staggering only");
	}
	if (ctx->d_reta_n == 0) TB_FAIL(ctx, "vertical coordinate not set (tb200_set_vertical_coordinate)");
	if (ctx->d_precip == 0) {
		if (dalloc(ctx, &ctx->d_precip, (size_t)lay.nelem * lay.nn)) return 1;
		TB_CHECK(ctx, cudaMemset(ctx->d_precip, 0, (size_t)lay.nelem * lay.nn * sizeof(double)));
	}
	KesslerArgs a;
	a.zs = ctx->g2d[6];
	a.reta_n = ctx->d_reta_n;
	a.ztop = ctx->cfg.ztop;
	a.dt = dt;
	a.gamma = ctx->cfg.cp / (ctx->cfg.cp - ctx->cfg.R);
	a.pressure_scaling = ctx->cfg.p0 * pow(ctx->cfg.R / ctx->cfg.p0, a.gamma);
	a.R = ctx->cfg.R;
	a.precip = ctx->d_precip;
	a.ws = ctx->d_ws;
	const size_t ws_doubles = (size_t)tb_column_ws_entries(lay.nlev, ctx->offd) * ctx->ws_cols;
	const long long total = lay.nelem * (long long)lay.nn;
	long long chunk = (long long)(ws_doubles / ((size_t)9 * lay.nlev)) / 128 * 128;
	if (chunk < 128) TB_FAIL(ctx, "column workspace too small for the Kessler step");
	for (long long c0 = 0; c0 < total; c0 += chunk) {
		a.col0 = c0;
		a.ncols = std::min(chunk, total - c0);
		auto kfn = k_kessler;
		TB_LAUNCH_FLAT(kfn, dim3((unsigned)((a.ncols + 127) / 128)), dim3(128), 0, ctx->stream,
			lay, a, ctx->inst[0]);
		TB_KERNEL_CHECK(ctx);
	}
	return 0;
}

///////////////////////////////////////////////////////////////////////////////
// State movement

// Host <-> device state movement.
//
// A host array (reference layout [c][iA][iB][k], one-node halo) crosses the bus
// as pitched copies of the interior of its valid component slices into / out of
// a staging buffer of the same shape, and k_transpose_state converts between
// that and the element-major device layout.  Two staging buffers and a copy
// stream pipeline the two: while one array is on the bus the previous one is
// being transposed (upload) or the next one is (download); per-buffer events
// order producer and consumer, nothing synchronises with the host until the
// caller asks (tb200_transfer_sync; the plain tb200_upload_state /
// tb200_download_state do that before they return - an upload only waits for
// its bus copies, not for the transposes behind them).

static int xfer_setup(tb200_ctx * ctx) {
	if (ctx->copy_stream != 0) return 0;
#ifndef TB200_EMU
	TB_CHECK(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
	for (int q = 0; q < 2; q++) {
		TB_CHECK(ctx, cudaEventCreateWithFlags(&ctx->ev_stage_free[q], cudaEventDisableTiming));
		TB_CHECK(ctx, cudaEventCreateWithFlags(&ctx->ev_stage_full[q], cudaEventDisableTiming));
	}
	TB_CHECK(ctx, cudaEventCreateWithFlags(&ctx->ev_compute, cudaEventDisableTiming));
#else
	ctx->copy_stream = (cudaStream_t)1;
#endif
	ctx->stage_buf[0] = ctx->d_stage;
	if (dalloc(ctx, &ctx->stage_buf[1], ctx->stage_doubles)) return 1;
	// device row of every host component: node array, interface array, tracers
	const DevLayout & lay = ctx->lay;
	std::vector<int> maps(64, -1);
	for (int c = 0; c < lay.ncomp && c < 8; c++) {
		maps[c] = lay.onedge[c] ? -1 : lay.rowoff[c];
		maps[8 + c] = lay.onedge[c] ? lay.rowoff[c] : -1;
	}
	for (int q = 0; q < lay.ntr && q < 48; q++) maps[16 + q] = lay.troff + q * lay.nlev;
	TB_CHECK(ctx, cudaMemcpy(ctx->d_rowmap, maps.data(), 64 * sizeof(int), cudaMemcpyHostToDevice));
	ctx->rowmap_h = maps;
	return 0;
}

// pitched copy of the interior of one component slice between the host array
// and the staging buffer (halo entries are not touched on either side)
static int copy_interior(
	tb200_ctx * ctx, const PatchInfo & pi, double * host, double * stage, size_t comp,
	int host_nlev, bool to_device
) {
	const int np = ctx->lay.np;
	const size_t wa = pi.nea * np + 2 * pi.halo;
	const size_t wb = pi.neb * np + 2 * pi.halo;
	const size_t slice = wa * wb * host_nlev;
	const size_t pitch = wb * host_nlev * sizeof(double);
	const size_t width = (wb - 2 * pi.halo) * host_nlev * sizeof(double);
	const size_t off = comp * slice + ((size_t)pi.halo * wb + pi.halo) * host_nlev;
	if (to_device) {
		TB_CHECK(ctx, cudaMemcpy2DAsync(stage + off, pitch, host + off, pitch,
			width, wa - 2 * pi.halo, cudaMemcpyHostToDevice, ctx->copy_stream));
	} else {
		TB_CHECK(ctx, cudaMemcpy2DAsync(host + off, pitch, stage + off, pitch,
			width, wa - 2 * pi.halo, cudaMemcpyDeviceToHost, ctx->copy_stream));
	}
	return 0;
}

// One host array of one patch.  map0: first entry of the array's row map in
// d_rowmap (0 node, 8 interfaces, 16 tracers); derived: components whose
// derived copies the reference keeps at this location (download only).
static int move_state_array(
	tb200_ctx * ctx, const PatchInfo & pi, int inst, double * host,
	int ncomp_host, int host_nlev, int map0, bool to_device,
	const std::vector<int> & derived = std::vector<int>()
) {
	if (host == 0) return 0;
	if (xfer_setup(ctx)) return 1;
	const int np = ctx->lay.np, nn = ctx->lay.nn;
	const size_t wa = pi.nea * np + 2 * pi.halo;
	const size_t wb = pi.neb * np + 2 * pi.halo;
	const size_t slice = wa * wb * host_nlev;
	if (slice * ncomp_host > ctx->stage_doubles) TB_FAIL(ctx, "staging buffer too small");
	const int slot = (int)(ctx->xfer_jobs++ & 1);
	double * stage = ctx->stage_buf[slot];
	const int * rowmap = ctx->d_rowmap + map0;
	const size_t smem = (size_t)host_nlev * (nn + 1) * sizeof(double);
	if (to_device) {
#ifndef TB200_EMU
		// the transpose that last read this buffer is done
		TB_CHECK(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_stage_free[slot], 0));
#endif
		for (int c = 0; c < ncomp_host; c++) {
			if (ctx->rowmap_h[map0 + c] < 0) continue;
			if (copy_interior(ctx, pi, host, stage, c, host_nlev, true)) return 1;
		}
#ifndef TB200_EMU
		TB_CHECK(ctx, cudaEventRecord(ctx->ev_stage_full[slot], ctx->copy_stream));
		TB_CHECK(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_stage_full[slot], 0));
#endif
		auto kfn = k_transpose_state<true>;
		TB_LAUNCH(kfn, dim3(pi.nea * pi.neb), dim3(256), smem, ctx->stream,
			ctx->lay, ctx->inst[inst], stage, pi.elem0, pi.nea, pi.neb, pi.halo,
			0, ncomp_host, host_nlev, 0, rowmap);
		TB_KERNEL_CHECK(ctx);
#ifndef TB200_EMU
		TB_CHECK(ctx, cudaEventRecord(ctx->ev_stage_free[slot], ctx->stream));
#endif
	} else {
#ifndef TB200_EMU
		// the bus copy that last read this buffer is done
		TB_CHECK(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_stage_free[slot], 0));
#endif
		auto kfn = k_transpose_state<false>;
		TB_LAUNCH(kfn, dim3(pi.nea * pi.neb), dim3(256), smem, ctx->stream,
			ctx->lay, ctx->inst[inst], stage, pi.elem0, pi.nea, pi.neb, pi.halo,
			0, ncomp_host, host_nlev, 0, rowmap);
		TB_KERNEL_CHECK(ctx);
		for (size_t q = 0; q < derived.size(); q++) {
			// W on levels / U, V on interfaces (HorizontalDynamicsFEM.cpp:817-831)
			auto kfd = k_fill_derived;
			const int c = derived[q];
			TB_LAUNCH_FLAT(kfd, dim3(pi.nea * pi.neb), dim3(256), 0, ctx->stream,
				ctx->lay, ctx->ops, (const double *)ctx->inst[inst], stage,
				pi.elem0, pi.nea, pi.neb, pi.halo, c, c, (map0 == 0) ? 1 : 0, host_nlev);
			TB_KERNEL_CHECK(ctx);
		}
#ifndef TB200_EMU
		TB_CHECK(ctx, cudaEventRecord(ctx->ev_stage_full[slot], ctx->stream));
		TB_CHECK(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_stage_full[slot], 0));
#endif
		for (int c = 0; c < ncomp_host; c++) {
			bool take = ctx->rowmap_h[map0 + c] >= 0;
			for (size_t q = 0; q < derived.size(); q++) take = take || derived[q] == c;
			if (!take) continue;
			if (copy_interior(ctx, pi, host, stage, c, host_nlev, false)) return 1;
		}
#ifndef TB200_EMU
		TB_CHECK(ctx, cudaEventRecord(ctx->ev_stage_free[slot], ctx->copy_stream));
#endif
	}
	return 0;
}

extern "C" int tb200_transfer_sync(tb200_ctx * ctx) {
#ifndef TB200_EMU
	if (ctx->copy_stream != 0) TB_CHECK(ctx, cudaStreamSynchronize(ctx->copy_stream));
#endif
	return 0;
}

extern "C" int tb200_upload_state_async(
	tb200_ctx * ctx, int patch_index, int inst,
	const double * node, const double * redge, const double * tracers
) {
	if (!ctx->committed) TB_FAIL(ctx, "commit the layout first");
	PatchInfo * pi = find_patch(ctx, patch_index);
	if (pi == 0 || pi->elem0 < 0) TB_FAIL(ctx, "not a local patch");
	if (inst < 0 || inst >= (int)ctx->inst.size()) TB_FAIL(ctx, "invalid instance");
	const DevLayout & lay = ctx->lay;
	if (move_state_array(ctx, *pi, inst, (double *)node, lay.ncomp, lay.nlev, 0, true)) return 1;
	bool anyedge = false;
	for (int c = 0; c < lay.ncomp; c++) anyedge = anyedge || lay.onedge[c];
	if (anyedge) {
		if (move_state_array(ctx, *pi, inst, (double *)redge, lay.ncomp, lay.nlev + 1, 8, true)) return 1;
	}
	if (lay.ntr > 0 && tracers != 0) {
		if (lay.ntr > 48) TB_FAIL(ctx, "more than 48 tracers");
		if (move_state_array(ctx, *pi, inst, (double *)tracers, lay.ntr, lay.nlev, 16, true)) return 1;
	}
	return 0;
}

extern "C" int tb200_upload_state(
	tb200_ctx * ctx, int patch_index, int inst,
	const double * node, const double * redge, const double * tracers
) {
	if (tb200_upload_state_async(ctx, patch_index, inst, node, redge, tracers)) return 1;
	// the host arrays are free once their bus copies are done; the transposes
	// run on behind them in stream order
	return tb200_transfer_sync(ctx);
}

extern "C" int tb200_download_state_async(
	tb200_ctx * ctx, int patch_index, int inst,
	double * node, double * redge, double * tracers, int fill_derived
) {
	if (!ctx->committed) TB_FAIL(ctx, "commit the layout first");
	PatchInfo * pi = find_patch(ctx, patch_index);
	if (pi == 0 || pi->elem0 < 0) TB_FAIL(ctx, "not a local patch");
	if (inst < 0 || inst >= (int)ctx->inst.size()) TB_FAIL(ctx, "invalid instance");
	const DevLayout & lay = ctx->lay;
	const bool nh = (ctx->cfg.eqn_type == TB200_EQN_PRIMITIVE_NONHYDRO);
	std::vector<int> derived;
	if (node != 0) {
		derived.clear();
		if (fill_derived && nh) derived.push_back(3);
		if (move_state_array(ctx, *pi, inst, node, lay.ncomp, lay.nlev, 0, false, derived)) return 1;
	}
	bool anyedge = false;
	for (int c = 0; c < lay.ncomp; c++) anyedge = anyedge || lay.onedge[c];
	if (redge != 0 && anyedge) {
		derived.clear();
		if (fill_derived && nh) { derived.push_back(0); derived.push_back(1); }
		if (move_state_array(ctx, *pi, inst, redge, lay.ncomp, lay.nlev + 1, 8, false, derived)) return 1;
	}
	if (lay.ntr > 0 && tracers != 0) {
		if (lay.ntr > 48) TB_FAIL(ctx, "more than 48 tracers");
		derived.clear();
		if (move_state_array(ctx, *pi, inst, tracers, lay.ntr, lay.nlev, 16, false, derived)) return 1;
	}
	return 0;
}

extern "C" int tb200_download_state(
	tb200_ctx * ctx, int patch_index, int inst,
	double * node, double * redge, double * tracers, int fill_derived
) {
	if (tb200_download_state_async(ctx, patch_index, inst, node, redge, tracers, fill_derived)) return 1;
	return tb200_transfer_sync(ctx);
}

// Pin / unpin host memory the state arrays live in (the reference's
// DataContainer blocks, DataContainer.cpp:77-147) so that the bus copies run
// asynchronously at full rate.  The C++ shells only see the C ABI.
extern "C" int tb200_host_register(tb200_ctx * ctx, void * p, size_t bytes) {
	TB_CHECK(ctx, cudaHostRegister(p, bytes, 0));
	return 0;
}

extern "C" int tb200_host_unregister(tb200_ctx * ctx, void * p) {
	TB_CHECK(ctx, cudaHostUnregister(p));
	return 0;
}

///////////////////////////////////////////////////////////////////////////////
// Copy / LinearCombine / Zero

static void mask_rows(const tb200_ctx * ctx, int mask, int & row0, int & row1) {
	const DevLayout & lay = ctx->lay;
	row0 = (mask & TB200_DATA_STATE) ? 0 : lay.nrows_state;
	row1 = (mask & TB200_DATA_TRACERS) ? lay.nrows : lay.nrows_state;
}

static int launch_combine(tb200_ctx * ctx, const CombineArgs & ca, int dst, int row0, int row1) {
	if (row1 <= row0) return 0;
	const DevLayout & lay = ctx->lay;
	const long long total = lay.nelem * (long long)(row1 - row0) * lay.nn;
	const int block = 256;
	long long nb = (total + block - 1) / block;
	if (nb > 148 * 16) nb = 148 * 16;
	auto kfn = k_lincomb;
	TB_LAUNCH_FLAT(kfn, dim3((unsigned)nb), dim3(block), 0, ctx->stream,
		lay, ca, ctx->inst[dst], row0, row1);
	TB_KERNEL_CHECK(ctx);
	return 0;
}

extern "C" int tb200_copy(tb200_ctx * ctx, int src, int dst, int mask) {
	const int ni = (int)ctx->inst.size();
	if (src < 0 || src >= ni || dst < 0 || dst >= ni) TB_FAIL(ctx, "Invalid index in CopyData.");
	if (src == dst) return 0;
	int row0, row1;
	mask_rows(ctx, mask, row0, row1);
	if (row1 <= row0) return 0;
	if (row0 == 0 && row1 == ctx->lay.nrows) {
		const size_t bytes = (size_t)ctx->lay.nelem * ctx->lay.nrows * ctx->lay.nn * sizeof(double);
		TB_CHECK(ctx, cudaMemcpyAsync(ctx->inst[dst], ctx->inst[src], bytes,
			cudaMemcpyDeviceToDevice, ctx->stream));
		ctx->writes++;              // not a kernel launch, still a write
		return 0;
	}
	CombineArgs ca;
	memset(&ca, 0, sizeof(ca));
	ca.nsrc = 1;
	ca.src[0] = ctx->inst[src];
	ca.coeff[0] = 1.0;
	ca.scale_dst = 0;
	return launch_combine(ctx, ca, dst, row0, row1);
}

extern "C" int tb200_lincomb(
	tb200_ctx * ctx, const double * coeff, int ncoeff, int dst, int mask
) {
	const int ni = (int)ctx->inst.size();
	if (dst < 0 || dst >= ni) TB_FAIL(ctx, "Invalid ixDest index in LinearCombineData.");
	if (dst >= ncoeff) TB_FAIL(ctx, "Destination index out of coefficient bounds");
	if (ncoeff > ni) TB_FAIL(ctx, "Too many elements in coefficient vector.");
	CombineArgs ca;
	memset(&ca, 0, sizeof(ca));
	ca.cdst = coeff[dst];
	ca.scale_dst = (coeff[dst] == 0.0) ? 0 : 1;
	for (int m = 0; m < ncoeff; m++) {
		if (m == dst || coeff[m] == 0.0) continue;
		ca.src[ca.nsrc] = ctx->inst[m];
		ca.coeff[ca.nsrc] = coeff[m];
		ca.nsrc++;
	}
	// dest = 1.0 * dest and nothing added: bitwise no-op
	if (ca.nsrc == 0 && ca.scale_dst && ca.cdst == 1.0) return 0;
	int row0, row1;
	mask_rows(ctx, mask, row0, row1);
	// dest += c * (the increment the Strang tail just wrote): its u, v rows are
	// zero, so those rows of dest stay as they are (Strang carry-over,
	// TimestepSchemeStrang.cpp:470-482)
	if (ca.nsrc == 1 && ca.scale_dst && ca.cdst == 1.0 && ctx->uvzero_inst >= 0
		&& ca.src[0] == ctx->inst[ctx->uvzero_inst] && ctx->writes == ctx->uvzero_writes
		&& row0 == 0 && ctx->lay.rowoff[0] == 0
		&& ctx->lay.rowoff[1] == ctx->lay.rowlev[0]
		&& !ctx->carry_full                     // TB200_CARRY_FULL=1 (tests): combine every row
	) {
		row0 = ctx->lay.rowoff[1] + ctx->lay.rowlev[1];
	}
	return launch_combine(ctx, ca, dst, row0, row1);
}

extern "C" int tb200_zero(tb200_ctx * ctx, int inst, int mask) {
	const int ni = (int)ctx->inst.size();
	if (inst < 0 || inst >= ni) TB_FAIL(ctx, "Invalid ixData index in ZeroData.");
	CombineArgs ca;
	memset(&ca, 0, sizeof(ca));
	int row0, row1;
	mask_rows(ctx, mask, row0, row1);
	return launch_combine(ctx, ca, inst, row0, row1);
}


///////////////////////////////////////////////////////////////////////////////
// Fast path set-up (tb200_fast.cuh): operator windows, column constants and
// their verification against the uploaded reference metric.

// dense window of row `row` of operator h over inputs [first, first + nw):
// false if the row has a non-zero coefficient outside the window
static bool op_window(const HostOp & h, int row, int first, int nw, double * w) {
	for (int q = 0; q < nw; q++) w[q] = 0.0;
	if (row < 0 || row >= h.nout) return true;
	for (int l = h.begin[row]; l < h.end[row]; l++) {
		const double c = h.coeff[(size_t)row * h.width + (l - h.begin[row])];
		if (l >= first && l < first + nw) {
			w[l - first] = c;
		} else if (c != 0.0) {
			return false;
		}
	}
	return true;
}

static int fast_prepare(tb200_ctx * ctx) {
	if (ctx->fast_state != 0) return 0;
	ctx->fast_state = -1;
	const DevLayout & lay = ctx->lay;
	const int L = lay.nlev;
	const char * force = getenv("TB200_STAGE_KERNEL");
	if (force != 0 && strcmp(force, "generic") == 0) { ctx->fast_reason = "TB200_STAGE_KERNEL=generic"; return 0; }
	if (ctx->cfg.eqn_type != TB200_EQN_PRIMITIVE_NONHYDRO || lay.np != 4 || L < 2) {
		ctx->fast_reason = "not a nonhydrostatic np=4 configuration"; return 0;
	}
	if (ctx->cfg.vertical_order != 1) { ctx->fast_reason = "vertical order > 1"; return 0; }
	if (ctx->finite_volume) { ctx->fast_reason = "finite-volume vertical discretisation"; return 0; }
	if (ctx->mass_flux_levels) { ctx->fast_reason = "--vmassfluxlevels (general kernels)"; return 0; }
	if (uniform_on(ctx)) { ctx->fast_reason = "uniform diffusion (general kernels)"; return 0; }
	if ((int)ctx->reta_n_h.size() != L || (int)ctx->reta_e_h.size() != L + 1) {
		ctx->fast_reason = "vertical coordinate not set"; return 0;
	}
	for (size_t p = 0; p < ctx->patches.size(); p++) {
		if (ctx->patches[p].elem0 >= 0 && !ctx->patches[p].has_terrain) {
			ctx->fast_reason = "topography derivatives not set"; return 0;
		}
	}
	for (int q = 0; q < TB_NOPS; q++) {
		if (q == 5 || q == 6 || q == 10) continue;
		if (ctx->hops[q].nout == 0) { ctx->fast_reason = "column operators not set"; return 0; }
	}
	// operator windows
	std::vector<double> lev((size_t)(L + 1) * TBF_LW, 0.0);
	bool ok = true;
	for (int k = 0; k <= L; k++) {
		double * row = &lev[(size_t)k * TBF_LW];
		if (k < L) {
			ok = ok && op_window(ctx->hops[1], k, k, 2, row + TBF_CW);
			ok = ok && op_window(ctx->hops[2], k, k - 1, 3, row + TBF_CD);
			// penalty rows: the left operator acts on all but the top level, the
			// right one on all but the bottom level (VerticalDynamicsFEM.cpp:1009-1023)
			if (k <= L - 2) ok = ok && op_window(ctx->hops[8], k, k - 1, 3, row + TBF_CPL);
			if (k >= 1) ok = ok && op_window(ctx->hops[9], k, k - 1, 3, row + TBF_CPR);
			ok = ok && op_window(ctx->hops[4], k, k, 2, row + TBF_DEN);
			row[TBF_SN] = 1.0 - ctx->reta_n_h[k];
		}
		if (k >= 1 && k < L) {
			ok = ok && op_window(ctx->hops[0], k, k - 1, 3, row + TBF_CILO);
			if (row[TBF_CILO + 2] != 0.0) ok = false;   // interface k reads levels k-1, k
			ok = ok && op_window(ctx->hops[3], k, k - 1, 2, row + TBF_DNE);
		}
		if (k + 1 < L) {
			ok = ok && op_window(ctx->hops[0], k + 1, k - 1, 3, row + TBF_CIHI);
		}
		if (k >= 1) {
			ok = ok && op_window(ctx->hops[1], k - 1, k - 1, 2, row + TBF_IEN1);
		}
		ok = ok && op_window(ctx->hops[7], k, k - 1, 3, row + TBF_DDE);
		row[TBF_SE] = 1.0 - ctx->reta_e_h[k];
		row[TBF_SE1] = (k < L) ? (1.0 - ctx->reta_e_h[k + 1]) : 0.0;
	}
	ok = ok && op_window(ctx->hops[0], 0, 0, 3, &lev[TBF_CB0]);
	if (!ok) { ctx->fast_reason = "a column operator row is wider than the order-1 window"; return 0; }
	if (ctx->d_lev == 0) {
		if (dalloc(ctx, &ctx->d_lev, lev.size())) return 1;
		if (dalloc(ctx, &ctx->d_colc, (size_t)lay.nelem * TBF_NC * lay.nn)) return 1;
	}
	TB_CHECK(ctx, cudaMemcpy(ctx->d_lev, lev.data(), lev.size() * sizeof(double), cudaMemcpyHostToDevice));
	DevGeom g = ctx->geom;
	g.tda = ctx->d_tda; g.tdb = ctx->d_tdb; g.ztop = ctx->cfg.ztop;
	const long long ncol = lay.nelem * lay.nn;
	{
		auto kfn = k_fast_colc;
		TB_LAUNCH_FLAT(kfn, dim3((unsigned)((ncol + 255) / 256)), dim3(256), 0, ctx->stream,
			ncol, g, ctx->cfg.g, ctx->d_colc);
		TB_KERNEL_CHECK(ctx);
	}
	ctx->fast_metric_error = 0.0;
	if (ctx->geometry3d_uploaded) {
		const int nb = 148, nt = 128;
		double * d_errs = 0;
		if (dalloc(ctx, &d_errs, (size_t)nb * nt)) return 1;
		auto kfn = k_fast_verify;
		TB_LAUNCH_FLAT(kfn, dim3(nb), dim3(nt), 0, ctx->stream,
			lay, ctx->geom, ctx->cfg.g, (const double *)ctx->d_colc, (const double *)ctx->d_lev, d_errs);
		TB_KERNEL_CHECK(ctx);
		TB_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
		std::vector<double> errs((size_t)nb * nt);
		TB_CHECK(ctx, cudaMemcpy(errs.data(), d_errs, errs.size() * sizeof(double), cudaMemcpyDeviceToHost));
		double worst = 0.0;
		for (size_t q = 0; q < errs.size(); q++) worst = std::max(worst, errs[q]);
		ctx->fast_metric_error = worst;
		if (!(worst <= 1.0e-13)) {
			char buf[160];
			snprintf(buf, 160, "column constants deviate from the uploaded metric by %.3e", worst);
			ctx->fast_reason = buf;
			return 0;
		}
	}
	TB_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
	ctx->fast_state = 1;
	ctx->fast_reason = "";
	return 0;
}

// 1: fast path in use; 0: generic kernels (tb200_fast_path_reason tells why)
extern "C" int tb200_fast_path(tb200_ctx * ctx) {
	if (fast_prepare(ctx)) return -1;
	return (ctx->fast_state == 1) ? 1 : 0;
}

extern "C" const char * tb200_fast_path_reason(tb200_ctx * ctx) {
	return ctx->fast_reason.c_str();
}

extern "C" double tb200_fast_path_metric_error(tb200_ctx * ctx) {
	return ctx->fast_metric_error;
}

///////////////////////////////////////////////////////////////////////////////
// Dynamics

// The general kernels read the stored level / interface Jacobians (and, without
// the on-the-fly metric, every 3-D metric array): refuse to run them on a
// context whose host supplied the lean geometry only.
static int need_metric3d(tb200_ctx * ctx, const char * what) {
	const bool have_jac = (ctx->g3n[0] != 0 && ctx->g3e[0] != 0);
	const bool have_all = ctx->geometry3d_uploaded;
	if (have_jac && (have_all || ctx->geom.analytic
		|| ctx->cfg.eqn_type != TB200_EQN_PRIMITIVE_NONHYDRO)) return 0;
	ctx->err = std::string(what) + ": the 3-D metric arrays were not uploaded "
		"(lean geometry runs the column-constant kernels only)";
	return 1;
}

static int check_ops(tb200_ctx * ctx) {
	for (int q = 0; q < TB_NOPS; q++) {
		if (q == 5 || q == 6 || q == 10) continue;   // not used by the hot path
		if (ctx->ops.op[q].coeff == 0) TB_FAIL(ctx, "vertical column operators not set");
	}
	if (ctx->mass_flux_levels && ctx->ops.op[10].coeff == 0) {
		TB_FAIL(ctx, "--vmassfluxlevels needs the zero-boundaries DiffNodeToNode operator");
	}
	if ((ctx->uniform_s != 0.0 || ctx->uniform_v != 0.0) && ctx->ops.op[6].coeff == 0) {
		TB_FAIL(ctx, "uniform diffusion needs the DiffDiffNodeToNode operator");
	}
	return 0;
}

static StageBase stage_base_out() {
	StageBase sb;
	memset(&sb, 0, sizeof(sb));
	sb.use_out = 1;
	return sb;
}

// Tracers ride the column-constant path too (tb200_tracers_fast.cuh) unless
// TB200_TRACER_KERNEL=generic asks for the general kernels (tests).
static bool fast_with_tracers(const tb200_ctx * ctx) {
	if (ctx->lay.ntr == 0) return true;
	const char * force = getenv("TB200_TRACER_KERNEL");
	return !(force != 0 && strcmp(force, "generic") == 0);
}

// The explicit stage on the column-constant path?  (state rows: k_nh_stage_pipe /
// k_nh_stage_fast; tracer rows: k_tracer_stage)
static bool stage_fast_ok(tb200_ctx * ctx) {
	if (fast_prepare(ctx)) return false;
	return ctx->fast_state == 1 && fast_with_tracers(ctx);
}

static int nh_launch_state(
	tb200_ctx * ctx, int in, int out, double dt, bool do_h, bool do_v,
	const StageBase & sb
);

// Tracer rows of an explicit stage on the fast path: stage base, horizontal
// transport and the element-wise positivity filter in one pass (do_h); the
// vertical explicit step leaves tracers alone (VerticalDynamicsFEM.cpp:616-1159),
// only a stage base that is not already in place is formed.
static int tracer_stage_fast(
	tb200_ctx * ctx, int in, int out, double dt, bool do_h, const StageBase & sb
) {
	const DevLayout & lay = ctx->lay;
	if (!do_h) {
		if (sb.use_out) return 0;
		CombineArgs ca;
		memset(&ca, 0, sizeof(ca));
		ca.cdst = sb.cdst;
		ca.scale_dst = sb.scale_dst;
		ca.nsrc = sb.nsrc;
		for (int m = 0; m < sb.nsrc; m++) { ca.src[m] = sb.src[m]; ca.coeff[m] = sb.coeff[m]; }
		return launch_combine(ctx, ca, out, lay.nrows_state, lay.nrows);
	}
	if (ctx->d_area_node == 0) TB_FAIL(ctx, "element areas not uploaded (tracer filter)");
	TracerFastArgs ta;
	ta.colc = ctx->d_colc;
	ta.lev = ctx->d_lev;
	ta.inv_da = ctx->d_inv_da;
	ta.inv_db = ctx->d_inv_db;
	ta.area = ctx->d_area_node;
	ta.dt = dt;
	const ElemList el = elem_list(ctx, 0);
	if (el.n == 0) return 0;
	// stage base of at most two terms: pipelined kernel (bulk copies one element
	// ahead); TB200_TRACER_STAGE_KERNEL=plain keeps the block-per-element kernel
	TracerBase tbse;
	memset(&tbse, 0, sizeof(tbse));
	bool fits = true;
	if (sb.use_out) {
		tbse.src[0] = ctx->inst[out]; tbse.coeff[0] = 1.0; tbse.nsrc = 1; tbse.exact = 1;
	} else {
		if (sb.scale_dst) {
			tbse.src[0] = ctx->inst[out]; tbse.coeff[0] = sb.cdst; tbse.nsrc = 1; tbse.first_is_dst = 1;
		}
		for (int m = 0; m < sb.nsrc; m++) {
			if (tbse.nsrc == 2) { fits = false; break; }
			tbse.src[tbse.nsrc] = sb.src[m]; tbse.coeff[tbse.nsrc] = sb.coeff[m]; tbse.nsrc++;
		}
		if (tbse.nsrc == 0) fits = false;
		if (fits && tbse.nsrc == 1 && !tbse.first_is_dst && tbse.coeff[0] == 1.0
			&& tbse.src[0] == ctx->inst[in]) {
			tbse.nsrc = 0;        // base = the input tracers: read once
		}
	}
	const size_t smem = tb_tracer_pipe_smem_doubles(lay.nlev, lay.ntr, tbse.nsrc) * sizeof(double) + 1024;
	const char * force = getenv("TB200_TRACER_STAGE_KERNEL");
	if (fits && smem <= 227 * 1024 - 1024 && !(force != 0 && strcmp(force, "plain") == 0)) {
		auto kfn = k_tracer_stage_pipe;
#ifndef TB200_EMU
		TB_CHECK(ctx, cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
#endif
		const dim3 grid((unsigned)persistent_blocks(ctx, kfn, TBT_THREADS, smem, el.n));
		TB_LAUNCH(kfn, grid, dim3(TBT_THREADS), smem, ctx->stream,
			lay, ctx->tables, ta, tbse, (const double *)ctx->inst[in], ctx->inst[out], el);
	} else {
		auto kfn = k_tracer_stage;
		TB_LAUNCH(kfn, dim3((unsigned)el.n), dim3(TBT_THREADS), 0, ctx->stream,
			lay, ctx->tables, ta, sb, (const double *)ctx->inst[in], ctx->inst[out], el);
	}
	ctx->launches++;
	ctx->writes++;
	TB_LAUNCH_CHECK(ctx);
	return 0;
}

// One explicit stage (either plugin or both) of state and tracers.  *filtered
// tells the caller that HorizontalDynamicsFEM::FilterNegativeTracers has been
// applied to the tracers of `out` already.
static int nh_launch(
	tb200_ctx * ctx, int in, int out, double dt, bool do_h, bool do_v,
	const StageBase & sb, bool * filtered = 0
) {
	if (filtered != 0) *filtered = false;
	if (check_ops(ctx)) return 1;
	if (fast_prepare(ctx)) return 1;
	if (nh_launch_state(ctx, in, out, dt, do_h, do_v, sb)) return 1;
	if (ctx->lay.ntr > 0 && stage_fast_ok(ctx)) {
		if (tracer_stage_fast(ctx, in, out, dt, do_h, sb)) return 1;
		if (filtered != 0 && do_h) *filtered = true;
	}
	return 0;
}

static int nh_launch_state(
	tb200_ctx * ctx, int in, int out, double dt, bool do_h, bool do_v,
	const StageBase & sb
) {
	const DevLayout & lay = ctx->lay;
	if (stage_fast_ok(ctx)) {
		FastArgs fa;
		fa.colc = ctx->d_colc;
		fa.lev = ctx->d_lev;
		fa.inv_da = ctx->d_inv_da;
		fa.inv_db = ctx->d_inv_db;
		fa.dt = dt;
		fa.xz = ctx->cfg.cartesian_xz;
		// stage base from at most TBP_MAXSRC instances: pipelined kernel
		const char * nopipe = getenv("TB200_STAGE_KERNEL");
		PipeBase pb;
		memset(&pb, 0, sizeof(pb));
		bool fits = true;
		if (sb.use_out) {
			pb.src[0] = ctx->inst[out]; pb.coeff[0] = 1.0; pb.nsrc = 1;
		} else {
			if (sb.scale_dst) {
				pb.src[0] = ctx->inst[out]; pb.coeff[0] = sb.cdst; pb.nsrc = 1;
			}
			for (int m = 0; m < sb.nsrc; m++) {
				if (pb.nsrc == TBP_MAXSRC) { fits = false; break; }
				pb.src[pb.nsrc] = sb.src[m]; pb.coeff[pb.nsrc] = sb.coeff[m]; pb.nsrc++;
			}
			if (pb.nsrc == 0) fits = false;       // an all-zero base: general kernel
		}
		// base = the input instance, copied: read it once
		if (fits && pb.nsrc == 1 && pb.coeff[0] == 1.0 && pb.src[0] == ctx->inst[in]) {
			pb.nsrc = 0;
		}
		if (do_h && fits && !(nopipe != 0 && strcmp(nopipe, "fast") == 0)) {
			// DSS of `out` follows: the kernel averages the in-patch groups itself
			bool fuse = do_v && fuse_enabled(ctx);
			if (fuse && tb_pipe_smem_doubles(lay.nrows_state, lay.nlev, pb.nsrc, true) * sizeof(double)
					> 227 * 1024 - 1024) fuse = false;
			const size_t smem = tb_pipe_smem_doubles(lay.nrows_state, lay.nlev, pb.nsrc, fuse) * sizeof(double);
			PipeMaps maps;
			maps.in = tensor_map_of(ctx, ctx->inst[in]);
			maps.b0 = tensor_map_of(ctx, pb.nsrc > 0 ? pb.src[0] : ctx->inst[in]);
			maps.b1 = tensor_map_of(ctx, pb.nsrc > 1 ? pb.src[1] : ctx->inst[in]);
			if (smem <= 227 * 1024 - 1024) {
				const dim3 block(TBF_THREADS);
#ifndef TB200_EMU
#define TB_PIPE_ATTR(kfn) TB_CHECK(ctx, cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem))
#else
#define TB_PIPE_ATTR(kfn)
#endif
#define TB_PIPE_LAUNCH(V, N) { \
					if (fuse) { \
						auto kfn = k_nh_stage_pipe<V, N, true>; \
						TB_PIPE_ATTR(kfn); \
						ctx->fuse_epoch++; \
						const FuseArgs fz = fuse_args(ctx); \
						const dim3 grid((unsigned)fused_blocks(ctx, kfn, TBF_THREADS, smem)); \
						TB_LAUNCH(kfn, grid, block, smem, ctx->stream, \
							lay, ctx->tables, ctx->phys, fa, \
							(const double *)ctx->inst[in], pb, ctx->inst[out], elem_list(ctx, 0), fz, maps); \
						ctx->launches++; \
						ctx->writes++; \
						ctx->fuse_done = true; \
					} else { \
					auto kfn = k_nh_stage_pipe<V, N, false>; \
					TB_PIPE_ATTR(kfn); \
					const FuseArgs fz = fuse_args(ctx); \
					const bool split = split_enabled(ctx); \
					if (split && split_fork(ctx)) return 1; \
					for (int part = split ? 1 : 0; part <= (split ? 2 : 0); part++) { \
						const ElemList el = elem_list(ctx, part); \
						if (el.n == 0) continue; \
						const dim3 grid((unsigned)persistent_blocks(ctx, kfn, TBF_THREADS, smem, el.n, \
							(part == 2) ? overlap_reserve() : 0)); \
						TB_LAUNCH(kfn, grid, block, smem, (part == 2) ? ctx->stream2 : ctx->stream, \
							lay, ctx->tables, ctx->phys, fa, \
							(const double *)ctx->inst[in], pb, ctx->inst[out], el, fz, maps); \
						ctx->launches++; \
						ctx->writes++; \
					} \
					if (split && split_mark(ctx)) return 1; } }
				if (do_v) {
					if (pb.nsrc == 0) TB_PIPE_LAUNCH(true, 0)
					else if (pb.nsrc == 1) TB_PIPE_LAUNCH(true, 1)
					else TB_PIPE_LAUNCH(true, 2)
				} else {
					if (pb.nsrc == 0) TB_PIPE_LAUNCH(false, 0)
					else if (pb.nsrc == 1) TB_PIPE_LAUNCH(false, 1)
					else TB_PIPE_LAUNCH(false, 2)
				}
#undef TB_PIPE_LAUNCH
#undef TB_PIPE_ATTR
				TB_LAUNCH_CHECK(ctx);
				return 0;
			}
		}
		const size_t smem = tb_fast_stage_smem_doubles(lay.nlev, do_h) * sizeof(double);
		const dim3 grid((unsigned)lay.nelem), block(TBF_THREADS);
#ifndef TB200_EMU
#define TB_FAST_ATTR(kfn) TB_CHECK(ctx, cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem))
#else
#define TB_FAST_ATTR(kfn)
#endif
		if (do_h && do_v) {
			auto kfn = k_nh_stage_fast<true, true>;
			TB_FAST_ATTR(kfn);
			TB_LAUNCH(kfn, grid, block, smem, ctx->stream, lay, ctx->tables, ctx->phys, fa, sb,
				(const double *)ctx->inst[in], ctx->inst[out]);
		} else if (do_h) {
			auto kfn = k_nh_stage_fast<true, false>;
			TB_FAST_ATTR(kfn);
			TB_LAUNCH(kfn, grid, block, smem, ctx->stream, lay, ctx->tables, ctx->phys, fa, sb,
				(const double *)ctx->inst[in], ctx->inst[out]);
		} else {
			auto kfn = k_nh_stage_fast<false, true>;
			TB_FAST_ATTR(kfn);
			TB_LAUNCH(kfn, grid, block, smem, ctx->stream, lay, ctx->tables, ctx->phys, fa, sb,
				(const double *)ctx->inst[in], ctx->inst[out]);
		}
#undef TB_FAST_ATTR
		TB_KERNEL_CHECK(ctx);
		return 0;
	}
	if (need_metric3d(ctx, "explicit stage (general kernel)")) return 1;
	NHArgs a;
	a.dt = dt;
	a.xz = ctx->cfg.cartesian_xz;
	a.fe_nodes = ctx->fe_nodes;
	// levels per pass: at most 16 (256 threads), chunks of equal size
	const int maxkb = std::max(1, 256 / lay.nn);
	const int nchunk = (lay.nlev + maxkb - 1) / maxkb;
	const int KB = (lay.nlev + nchunk - 1) / nchunk;
	const size_t smem = tb_nh_smem_doubles(lay.nlev, lay.nn, KB) * sizeof(double);
	const dim3 grid((unsigned)lay.nelem), block(KB * lay.nn);
#ifndef TB200_EMU
#define TB_NH_ATTR(k) TB_CHECK(ctx, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem))
#else
#define TB_NH_ATTR(k) (void)0
#endif
	if (do_h && do_v) {
		TB_NP_SWITCH(lay.np,
			auto kfn = k_nh_explicit<NPV, true, true>;
			TB_NH_ATTR(kfn);
			TB_LAUNCH(kfn, grid, block, smem, ctx->stream, lay, ctx->geom, ctx->tables,
				ctx->ops, ctx->phys, a, sb, (const double *)ctx->inst[in], ctx->inst[out], KB);)
	} else if (do_h) {
		TB_NP_SWITCH(lay.np,
			auto kfn = k_nh_explicit<NPV, true, false>;
			TB_NH_ATTR(kfn);
			TB_LAUNCH(kfn, grid, block, smem, ctx->stream, lay, ctx->geom, ctx->tables,
				ctx->ops, ctx->phys, a, sb, (const double *)ctx->inst[in], ctx->inst[out], KB);)
	} else {
		TB_NP_SWITCH(lay.np,
			auto kfn = k_nh_explicit<NPV, false, true>;
			TB_NH_ATTR(kfn);
			TB_LAUNCH(kfn, grid, block, smem, ctx->stream, lay, ctx->geom, ctx->tables,
				ctx->ops, ctx->phys, a, sb, (const double *)ctx->inst[in], ctx->inst[out], KB);)
	}
	TB_KERNEL_CHECK(ctx);
	return 0;
}

static int check_inst2(tb200_ctx * ctx, int in, int out) {
	const int ni = (int)ctx->inst.size();
	if (in < 0 || in >= ni || out < 0 || out >= ni) TB_FAIL(ctx, "invalid state instance");
	return 0;
}

static int uniform_diffusion_horizontal(tb200_ctx * ctx, int in, int out, double dt);
static int uniform_diffusion_vertical_uv(tb200_ctx * ctx, int in, int out, double dt);
static int stale_column_update(tb200_ctx * ctx, int in);

extern "C" int tb200_h_step_explicit(tb200_ctx * ctx, int in, int out, double dt) {
	TimingScope ts(ctx, (ctx->cfg.eqn_type == TB200_EQN_SHALLOW_WATER)
		? "HorizontalStepShallowWater" : "HorizontalStepNonhydrostaticPrimitive");
	if (check_inst2(ctx, in, out)) return 1;
	// HorizontalDynamicsFEM.cpp:1793
	if (in == out) TB_FAIL(ctx, "HorizontalDynamics Step must have iDataInitial != iDataUpdate");
	const DevLayout & lay = ctx->lay;
	if (ctx->cfg.eqn_type == TB200_EQN_SHALLOW_WATER) {
		const long long nitems = lay.nelem * lay.nlev;
		TB_NP_SWITCH(lay.np,
			auto kfn = k_sw_explicit<NPV, kItems>;
			TB_LAUNCH(kfn, dim3((unsigned)((nitems + kItems - 1) / kItems)), dim3(NPV * NPV * kItems), 0,
				ctx->stream, lay, ctx->geom, ctx->tables,
				(const double *)ctx->inst[in], ctx->inst[out], dt, ctx->cfg.g);)
		TB_KERNEL_CHECK(ctx);
	} else {
		bool filtered = false;
		if (nh_launch(ctx, in, out, dt, true, false, stage_base_out(), &filtered)) return 1;
		if (filtered) return 0;
	}
	if (uniform_on(ctx) && uniform_diffusion_horizontal(ctx, in, out, dt)) return 1;
	return tb200_filter_negative_tracers(ctx, out);
}

// uniform diffusion terms of BuildF (VerticalDynamicsFEM.cpp:2594-2636)
static int column_uniform_args(tb200_ctx * ctx, ColumnArgs & ca) {
	ca.mass_flux_levels = ctx->mass_flux_levels;
	if (!uniform_on(ctx)) return 0;
	if (refstate_alloc(ctx)) return 1;
	const double ztop = ctx->cfg.ztop;
	ca.ref = ctx->d_refstate;
	ca.uni_s = ctx->uniform_s / (ztop * ztop);
	ca.uni_v = ctx->uniform_v / (ztop * ztop);
	return 0;
}

// Fully explicit vertical step: BuildF of every element-local column of `in`
// (k_column_implicit in its explicit mode), out -= dt * F.
static int explicit_vertical_columns(tb200_ctx * ctx, int in, int out, double dt) {
	const DevLayout & lay = ctx->lay;
	if (check_ops(ctx)) return 1;
	if (need_metric3d(ctx, "explicit vertical step")) return 1;
	ColumnArgs ca;
	ca.col_node = 0;
	ca.col_dups = 0;
	ca.ws = ctx->d_ws;
	ca.ws_stride = ctx->ws_cols;
	ca.dt = dt;
	ca.offd = ctx->offd;
	ca.fe_nodes = ctx->fe_nodes;
	ca.upwind_coeff = (1.0 / 2.0) * pow(1.0 / static_cast<double>(lay.nlev), 1.0);
	ca.info = ctx->d_info;
	ca.assemble_only = 3;
	if (column_uniform_args(ctx, ca)) return 1;
	const long long total = lay.nelem * (long long)lay.nn;
	for (long long c0 = 0; c0 < total; c0 += ctx->ws_cols) {
		ca.col0 = (int)c0;
		ca.ncols = (int)std::min<long long>(ctx->ws_cols, total - c0);
		auto kfn = k_column_implicit;
		TB_LAUNCH_FLAT(kfn, dim3((ca.ncols + 63) / 64), dim3(64), 0, ctx->stream,
			lay, ctx->geom, ctx->ops, ctx->phys, ca,
			(const double *)ctx->inst[in], ctx->inst[out]);
		TB_KERNEL_CHECK(ctx);
	}
	if (lay.ntr == 0) return 0;
	// UpdateColumnTracers in its explicit branches (:802-810, 4048-4171): the column
	// flux of every tracer with xi-dot of the initial state, out -= dt * F; the
	// column filter belongs to StepImplicit (:1637), which does nothing here
	TracerColumnArgs ta;
	memset(&ta, 0, sizeof(ta));
	ta.ws = ctx->d_ws;
	ta.ws_stride = ctx->ws_cols;
	ta.dt = dt;
	ta.fe_nodes = ctx->fe_nodes;
	ta.kl = 2 * ctx->cfg.vertical_order - 1;
	ta.info = ctx->d_info;
	ta.fully_explicit = 1;
	if (tb_tracer_ws_entries(lay.nlev, ta.kl) > tb_column_ws_entries(lay.nlev, ctx->offd)) {
		TB_FAIL(ctx, "column workspace too small for the tracer update");
	}
	for (long long c0 = 0; c0 < total; c0 += ctx->ws_cols) {
		ta.col0 = (int)c0;
		ta.ncols = (int)std::min<long long>(ctx->ws_cols, total - c0);
		auto kfn = k_column_tracers;
		TB_LAUNCH_FLAT(kfn, dim3((ta.ncols + 63) / 64), dim3(64), 0, ctx->stream,
			lay, ctx->geom, ctx->ops, ta,
			(const double *)ctx->inst[in], (const double *)ctx->inst[in],
			(const double *)ctx->inst[in], ctx->inst[out]);
		TB_KERNEL_CHECK(ctx);
	}
	return 0;
}

extern "C" int tb200_v_step_explicit(tb200_ctx * ctx, int in, int out, double dt) {
	TimingScope ts(ctx, "VerticalStepExplicit");
	if (check_inst2(ctx, in, out)) return 1;
	if (ctx->cfg.eqn_type != TB200_EQN_PRIMITIVE_NONHYDRO || ctx->lay.nlev == 1) {
		return 0;   // VerticalDynamicsStub (TempestInitialize.h:362-364)
	}
	if (in == out) TB_FAIL(ctx, "VerticalDynamics StepExplicit must have iDataInitial != iDataUpdate");
	if (ctx->cfg.fully_explicit) {
		// --explicitvertical (VerticalDynamicsFEM.cpp:748-793): rho theta, w and rho are
		// advanced with the column tendencies as well; general kernels
		if (explicit_vertical_columns(ctx, in, out, dt)) return 1;
	}
	if (nh_launch(ctx, in, out, dt, false, true, stage_base_out())) return 1;
	if (uniform_on(ctx)) return uniform_diffusion_vertical_uv(ctx, in, out, dt);
	return 0;
}

// VerticalDynamicsFEM::StepImplicitTermsExplicitly (VerticalDynamicsFEM.cpp:439-612):
// the column tendencies BuildF of every node of `in`, out -= dt * F (rho theta, w,
// rho); used by TimestepSchemeARK232.
extern "C" int tb200_v_step_implicit_terms_explicitly(tb200_ctx * ctx, int in, int out, double dt) {
	TimingScope ts(ctx, "VerticalStepImplicit");
	if (check_inst2(ctx, in, out)) return 1;
	if (ctx->cfg.eqn_type != TB200_EQN_PRIMITIVE_NONHYDRO || ctx->lay.nlev == 1) return 0;
	if (ctx->lay.ntr > 0) {
		TB_FAIL(ctx, "StepImplicitTermsExplicitly with tracers is not implemented");
	}
	return explicit_vertical_columns(ctx, in, out, dt);
}

extern "C" int tb200_hv_step_explicit(tb200_ctx * ctx, int in, int out, double dt) {
	if (check_inst2(ctx, in, out)) return 1;
	if (ctx->cfg.eqn_type != TB200_EQN_PRIMITIVE_NONHYDRO || ctx->lay.nlev == 1) {
		return tb200_h_step_explicit(ctx, in, out, dt);
	}
	if (in == out) TB_FAIL(ctx, "HorizontalDynamics Step must have iDataInitial != iDataUpdate");
	if ((ctx->lay.ntr > 0 && !stage_fast_ok(ctx)) || ctx->cfg.fully_explicit || uniform_on(ctx)) {
		// the tracer filter sits between the two plugins in the reference;
		// --explicitvertical adds the column tendencies in the vertical plugin
		if (tb200_h_step_explicit(ctx, in, out, dt)) return 1;
		return tb200_v_step_explicit(ctx, in, out, dt);
	}
	// (fast path with tracers: the vertical explicit step does not touch tracers,
	// so the filter commutes with it and is applied by the tracer kernel)
	TimingScope ts(ctx, "HorizontalStepNonhydrostaticPrimitive");
	return nh_launch(ctx, in, out, dt, true, true, stage_base_out());
}

// The explicit substage as the time schemes issue it: combination + both
// explicit plugins + PostProcessSubstage(out) (e.g. TimestepSchemeStrang.cpp:548-553).
// On several ranks the halo exchange of the DSS overlaps the elements that do
// not feed it.
extern "C" int tb200_hv_step_explicit_combine(
	tb200_ctx * ctx, const double * coeff, int ncoeff, int in, int out, double dt);
static int dss_instance(tb200_ctx * ctx, int inst, int mask, bool remainder_only);

extern "C" int tb200_hv_step_explicit_combine_dss(
	tb200_ctx * ctx, const double * coeff, int ncoeff, int in, int out, double dt
) {
	ctx->want_split = overlap_wanted();
	ctx->fuse_want = true;
	ctx->fuse_done = false;
	int rc = tb200_hv_step_explicit_combine(ctx, coeff, ncoeff, in, out, dt);
	ctx->want_split = false;
	ctx->fuse_want = false;
	// the stage kernel has averaged the in-patch groups: the rest (patch edges,
	// seams, strip ends, other ranks) goes through the group kernels
	if (rc == 0) rc = dss_instance(ctx, out, TB200_DATA_STATE | TB200_DATA_TRACERS, ctx->fuse_done);
	ctx->fuse_done = false;
	if (split_join(ctx)) return 1;
	return rc;
}

// Grid::LinearCombineData(coeff, out) (or CopyData when ncoeff == 0: copy of
// instance -ncoeff_src) followed by both explicit plugins, in one pass:
// HorizontalDynamics::StepExplicitCombine (HorizontalDynamics.h:97-106).
extern "C" int tb200_hv_step_explicit_combine(
	tb200_ctx * ctx, const double * coeff, int ncoeff, int in, int out, double dt
) {
	if (check_inst2(ctx, in, out)) return 1;
	const int ni = (int)ctx->inst.size();
	if (out >= ncoeff) TB_FAIL(ctx, "Destination index out of coefficient bounds");
	if (ncoeff > ni) TB_FAIL(ctx, "Too many elements in coefficient vector.");
	const bool fusable =
		(ctx->cfg.eqn_type == TB200_EQN_PRIMITIVE_NONHYDRO) && (ctx->lay.nlev > 1)
		&& (ctx->lay.ntr == 0 || stage_fast_ok(ctx)) && (in != out)
		&& !ctx->cfg.fully_explicit && !uniform_on(ctx);
	if (!fusable) {
		if (tb200_lincomb(ctx, coeff, ncoeff, out, TB200_DATA_STATE | TB200_DATA_TRACERS)) return 1;
		return tb200_hv_step_explicit(ctx, in, out, dt);
	}
	TimingScope ts(ctx, "HorizontalStepNonhydrostaticPrimitive");
	StageBase sb;
	memset(&sb, 0, sizeof(sb));
	sb.cdst = coeff[out];
	sb.scale_dst = (coeff[out] == 0.0) ? 0 : 1;
	for (int m = 0; m < ncoeff; m++) {
		if (m == out || coeff[m] == 0.0) continue;
		sb.src[sb.nsrc] = ctx->inst[m];
		sb.coeff[sb.nsrc] = coeff[m];
		sb.nsrc++;
	}
	if (fast_prepare(ctx)) return 1;
	const int nterms = sb.nsrc + (sb.scale_dst ? 1 : 0);
	if (ctx->fast_state == 1 && (nterms > TBP_MAXSRC || nterms == 0)
		&& getenv("TB200_STAGE_KERNEL") == 0) {
		// several sources: a streaming combine, then the pipelined stage kernel on
		// the pre-filled update instance (same operation order as the fused form)
		if (tb200_lincomb(ctx, coeff, ncoeff, out, TB200_DATA_STATE | TB200_DATA_TRACERS)) return 1;
		return nh_launch(ctx, in, out, dt, true, true, stage_base_out());
	}
	return nh_launch(ctx, in, out, dt, true, true, sb);
}

static int column_solve(tb200_ctx * ctx, int in, int out, double dt);
static int column_tracers(tb200_ctx * ctx, int in, int out, double dt);
static int filter_tracers(tb200_ctx * ctx, int inst, bool column, const CombineArgs * comb, double * inc);

extern "C" int tb200_v_step_implicit(tb200_ctx * ctx, int in, int out, double dt) {
	if (check_inst2(ctx, in, out)) return 1;
	if (ctx->cfg.eqn_type != TB200_EQN_PRIMITIVE_NONHYDRO || ctx->lay.nlev == 1) return 0;
	if (ctx->cfg.fully_explicit) return 0;   // VerticalDynamicsFEM.cpp:1240-1242
	if (check_ops(ctx)) return 1;
	const DevLayout & lay = ctx->lay;
	if (lay.ntr > 0) {
		// w before the solve is needed by the tracer update (the solve may run
		// in place): keep a copy [e][L+1][NN]
		const size_t wrow = (size_t)(lay.nlev + 1) * lay.nn;
		if (ctx->d_wold == 0) {
			if (dalloc(ctx, &ctx->d_wold, (size_t)lay.nelem * wrow)) return 1;
		}
		TB_CHECK(ctx, cudaMemcpy2DAsync(
			ctx->d_wold, wrow * sizeof(double),
			ctx->inst[in] + (size_t)lay.rowoff[3] * lay.nn, (size_t)lay.nrows * lay.nn * sizeof(double),
			wrow * sizeof(double), (size_t)lay.nelem, cudaMemcpyDeviceToDevice, ctx->stream));
		if (column_solve(ctx, in, out, dt)) return 1;
		return column_tracers(ctx, in, out, dt);
	}
	if (column_solve(ctx, in, out, dt)) return 1;
	return stale_column_update(ctx, in);
}

// UpdateColumnTracers for every unique column, then the column filter
// (VerticalDynamicsFEM.cpp:1528-1536, 1637)
static int column_tracers(tb200_ctx * ctx, int in, int out, double dt) {
	const DevLayout & lay = ctx->lay;
	if (fast_prepare(ctx)) return 1;
	const bool fastm = (ctx->fast_state == 1) && fast_with_tracers(ctx);
	if (!fastm && need_metric3d(ctx, "tracer column update")) return 1;
	const char * forcek = getenv("TB200_TRACER_COLUMN_KERNEL");   // "ws": general kernel, column-constant metric
	if (fastm && !(forcek != 0 && strcmp(forcek, "ws") == 0)) {
		// tridiagonal kernel, up to 6 tracers per pass
		TracerColumnFastArgs fa;
		fa.col_node = ctx->d_col_node;
		fa.col_dups = ctx->d_col_dups;
		fa.ncols = ctx->ncols;
		fa.colc = ctx->d_colc;
		fa.lev = ctx->d_lev;
		fa.w_old = ctx->d_wold;
		fa.dt = dt;
		fa.info = ctx->d_info;
		fa.keep = ctx->tracer_keep;
		const size_t ws_doubles = (size_t)tb_column_ws_entries(lay.nlev, ctx->offd) * ctx->ws_cols;
		const size_t smem = (size_t)(lay.nlev + 1) * TBF_LW * sizeof(double);
		for (int c0 = 0; c0 < lay.ntr; c0 += 6) {
			fa.c0 = c0;
			const int nt = std::min(6, lay.ntr - c0);
			// columns per launch: what the scratch holds, whole blocks
			long long chunk = (long long)(ws_doubles / ((size_t)(3 + nt) * lay.nlev))
				/ TBT_COL_THREADS * TBT_COL_THREADS;
			if (chunk < TBT_COL_THREADS) TB_FAIL(ctx, "column workspace too small for the tracer update");
			for (int col0 = 0; col0 < ctx->ncols; col0 += (int)chunk) {
				fa.col0 = col0;
				fa.ncols = (int)std::min<long long>(chunk, ctx->ncols - col0);
				fa.ws = ctx->d_ws;
				const dim3 grid((fa.ncols + TBT_COL_THREADS - 1) / TBT_COL_THREADS), block(TBT_COL_THREADS);
#ifndef TB200_EMU
#define TB_TRC_ATTR(kfn) TB_CHECK(ctx, cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem))
#else
#define TB_TRC_ATTR(kfn)
#endif
#define TB_TRC_LAUNCH(N) { \
					auto kfn = k_column_tracers_fast<N>; \
					TB_TRC_ATTR(kfn); \
					TB_LAUNCH(kfn, grid, block, smem, ctx->stream, lay, fa, \
						(const double *)ctx->inst[in], (const double *)ctx->inst[out], \
						(const double *)ctx->inst[in], ctx->inst[out]); }
				switch (nt) {
					case 1: TB_TRC_LAUNCH(1) break;
					case 2: TB_TRC_LAUNCH(2) break;
					case 3: TB_TRC_LAUNCH(3) break;
					case 4: TB_TRC_LAUNCH(4) break;
					case 5: TB_TRC_LAUNCH(5) break;
					default: TB_TRC_LAUNCH(6) break;
				}
#undef TB_TRC_LAUNCH
#undef TB_TRC_ATTR
				ctx->launches++;
				ctx->writes++;
				TB_LAUNCH_CHECK(ctx);
			}
		}
		return filter_tracers(ctx, out, true, 0, ctx->tracer_inc);
	}
	if (ctx->tracer_keep != 0) {
		// the tracers from before the update, for the increment formed after the filter
		CombineArgs keep;
		memset(&keep, 0, sizeof(keep));
		keep.nsrc = 1;
		keep.src[0] = ctx->inst[out];
		keep.coeff[0] = 1.0;
		int kdst = -1;
		for (int m = 0; m < (int)ctx->inst.size(); m++) if (ctx->inst[m] == ctx->tracer_keep) kdst = m;
		if (kdst < 0) TB_FAIL(ctx, "invalid tracer keep instance");
		if (launch_combine(ctx, keep, kdst, lay.nrows_state, lay.nrows)) return 1;
	}
	TracerColumnArgs ta;
	ta.colc = fastm ? ctx->d_colc : 0;
	ta.lev = fastm ? ctx->d_lev : 0;
	ta.col_node = ctx->d_col_node;
	ta.col_dups = ctx->d_col_dups;
	ta.ws = ctx->d_ws;
	ta.ws_stride = ctx->ws_cols;
	ta.dt = dt;
	ta.fe_nodes = ctx->fe_nodes;
	ta.kl = 2 * ctx->cfg.vertical_order - 1;
	ta.w_old = ctx->d_wold;
	ta.info = ctx->d_info;
	ta.fully_explicit = 0;
	if (tb_tracer_ws_entries(lay.nlev, ta.kl) > tb_column_ws_entries(lay.nlev, ctx->offd)) {
		TB_FAIL(ctx, "column workspace too small for the tracer update");
	}
	for (int c0 = 0; c0 < ctx->ncols; c0 += ctx->ws_cols) {
		ta.col0 = c0;
		ta.ncols = std::min(ctx->ws_cols, ctx->ncols - c0);
		const int block = 64;
		auto kfn = k_column_tracers;
		TB_LAUNCH_FLAT(kfn, dim3((ta.ncols + block - 1) / block), dim3(block), 0, ctx->stream,
			lay, ctx->geom, ctx->ops, ta,
			(const double *)ctx->inst[in], (const double *)ctx->inst[out],
			(const double *)ctx->inst[in], ctx->inst[out]);
		TB_KERNEL_CHECK(ctx);
	}
	return filter_tracers(ctx, out, true, 0, ctx->tracer_inc);
}

static int column_solve(tb200_ctx * ctx, int in, int out, double dt) {
	TimingScope ts(ctx, "VerticalStepImplicit");
	const DevLayout & lay = ctx->lay;
	ColumnArgs ca;
	ca.col_node = ctx->d_col_node;
	ca.col_dups = ctx->d_col_dups;
	ca.ws = ctx->d_ws;
	ca.ws_stride = ctx->ws_cols;
	ca.dt = dt;
	ca.offd = ctx->offd;
	ca.fe_nodes = ctx->fe_nodes;
	// m_dUpwindCoeff (VerticalDynamicsFEM.cpp:520-521)
	ca.upwind_coeff = (1.0 / 2.0) * pow(1.0 / static_cast<double>(lay.nlev), 1.0);
	ca.info = ctx->d_info;
	ca.assemble_only = 0;
	if (column_uniform_args(ctx, ca)) return 1;

	// vertical order 1 + terrain-following metric: specialised kernel
	// (tb200_column_fast.cuh); TB200_COLUMN_KERNEL = thread | warp | window selects
	// one of the general kernels instead
	if (fast_prepare(ctx)) return 1;
	if (ctx->fast_state == 1 && getenv("TB200_COLUMN_KERNEL") == 0) {
		ColumnFastArgs fa;
		fa.col_node = ctx->d_col_node;
		fa.col_dups = ctx->d_col_dups;
		fa.ws = ctx->d_ws;
		fa.dt = dt;
		fa.upwind_coeff = ca.upwind_coeff;
		fa.info = ctx->d_info;
		fa.colc = ctx->d_colc;
		fa.lev = ctx->d_lev;
		fa.inc = ctx->column_inc;
		const size_t smem = (size_t)(lay.nlev + 1) * TBF_LW * sizeof(double)
			+ (size_t)(TBC_THREADS / 32) * 3 * (lay.nlev + 1) * sizeof(unsigned);
		auto kfn = k_column_fast;
#ifndef TB200_EMU
		TB_CHECK(ctx, cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
#endif
		// scratch: 10 n doubles per column, columns padded to whole blocks
		const size_t ws_doubles = (size_t)tb_column_ws_entries(lay.nlev, ctx->offd) * ctx->ws_cols;
		const size_t per_block = (size_t)30 * (lay.nlev + 1) * TBC_THREADS;
		const long long nbatches_all = (ctx->ncols + TBC_THREADS - 1) / TBC_THREADS;
		// Persistent blocks (default): every resident block walks the batches
		// blockIdx.x, + gridDim.x, ... and reuses its own scratch slot, so that
		// the rows of U of a finished batch are overwritten in L2 instead of being
		// written back.  TB200_COLUMN_PERSISTENT=0: one block per batch, chunked
		// launches sized by the scratch (round-1 behaviour).
		const char * pers = getenv("TB200_COLUMN_PERSISTENT");
		long long grid_p = persistent_blocks(ctx, kfn, TBC_THREADS, smem, nbatches_all);
		if (!(pers != 0 && strcmp(pers, "0") == 0) && (size_t)grid_p * per_block <= ws_doubles) {
			fa.col0 = 0;
			fa.ncols = ctx->ncols;
			// (Measured and dropped: an access-policy window marking part of the scratch
			// as persisting in L2 - cudaLimitPersistingL2CacheSize at its maximum - made
			// this solve 2x slower, 5.7 ms, and every other kernel of the step with it:
			// the carve-out is taken from the L2 the streams of the step live in.)
			TB_LAUNCH(kfn, dim3((unsigned)grid_p), dim3(TBC_THREADS),
				smem, ctx->stream, lay, ctx->phys, fa,
				(const double *)ctx->inst[in], ctx->inst[out], (int)nbatches_all);
			TB_KERNEL_CHECK(ctx);
			return 0;
		}
		long long chunk = (long long)(ws_doubles / ((size_t)30 * (lay.nlev + 1))) / TBC_THREADS * TBC_THREADS;
		if (chunk < TBC_THREADS) TB_FAIL(ctx, "column workspace too small");
		if (chunk > ctx->ncols) chunk = ((ctx->ncols + TBC_THREADS - 1) / TBC_THREADS) * TBC_THREADS;
		for (int c0 = 0; c0 < ctx->ncols; c0 += (int)chunk) {
			fa.col0 = c0;
			fa.ncols = (int)std::min<long long>(chunk, ctx->ncols - c0);
			const int nb = (fa.ncols + TBC_THREADS - 1) / TBC_THREADS;
			TB_LAUNCH(kfn, dim3(nb), dim3(TBC_THREADS),
				smem, ctx->stream, lay, ctx->phys, fa,
				(const double *)ctx->inst[in], ctx->inst[out], nb);
			TB_KERNEL_CHECK(ctx);
		}
		return 0;
	}

	if (need_metric3d(ctx, "implicit column solve (general kernel)")) return 1;
	// one warp per column with all work arrays in shared memory, unless the
	// column is too tall for it (or TB200_COLUMN_KERNEL=thread asks for the
	// thread-per-column implementation)
	const size_t warp_bytes =
		(size_t)tb_column_warp_smem_doubles(lay.nlev, ctx->offd) * sizeof(double);
	const size_t sm_budget = 227 * 1024;
	int wpb = 0;
	{
		int best = 0;
		for (int w = 1; w <= 8; w++) {
			const size_t blk = warp_bytes * w + 1024;
			if (blk > sm_budget) break;
			const int per_sm = (int)(sm_budget / blk) * w;
			if (per_sm > best) { best = per_sm; wpb = w; }
		}
	}
	// default: thread per column with a sliding band window in shared memory
	// (vertical order 1); TB200_COLUMN_KERNEL = thread | warp | window overrides
	const char * force = getenv("TB200_COLUMN_KERNEL");
	bool use_window = (ctx->offd == TBW_KL) && (ctx->cfg.vertical_order == 1) && !ctx->finite_volume;
	if (force != 0 && strcmp(force, "thread") == 0) { wpb = 0; use_window = false; }
	// (the uniform diffusion terms of BuildF are in the thread-per-column kernel only)
	if (uniform_on(ctx) || ctx->mass_flux_levels) { wpb = 0; use_window = false; force = "thread"; }
	if (force != 0 && strcmp(force, "warp") == 0) use_window = false;
	if (force == 0 || strcmp(force, "warp") != 0) { if (use_window) wpb = 0; }
	if (use_window) {
		const size_t smem = tb_column_window_smem_bytes();
		auto kfn = k_column_implicit_window;
#ifndef TB200_EMU
		TB_CHECK(ctx, cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
#endif
		for (int c0 = 0; c0 < ctx->ncols; c0 += ctx->ws_cols) {
			ca.col0 = c0;
			ca.ncols = std::min(ctx->ws_cols, ctx->ncols - c0);
			TB_LAUNCH_FLAT(kfn, dim3((ca.ncols + TBW_THREADS - 1) / TBW_THREADS), dim3(TBW_THREADS),
				smem, ctx->stream, lay, ctx->geom, ctx->ops, ctx->phys, ca,
				(const double *)ctx->inst[in], ctx->inst[out]);
			TB_KERNEL_CHECK(ctx);
		}
		return 0;
	}

	if (wpb > 0) {
		ca.col0 = 0;
		ca.ncols = ctx->ncols;
		const size_t smem = warp_bytes * wpb;
		auto kfn = k_column_implicit_warp;
#ifndef TB200_EMU
		TB_CHECK(ctx, cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
#endif
		TB_LAUNCH(kfn, dim3((ca.ncols + wpb - 1) / wpb), dim3(32 * wpb), smem, ctx->stream,
			lay, ctx->geom, ctx->ops, ctx->phys, ca,
			(const double *)ctx->inst[in], ctx->inst[out], (int)(warp_bytes / sizeof(double)));
		TB_KERNEL_CHECK(ctx);
		return 0;
	}
	for (int c0 = 0; c0 < ctx->ncols; c0 += ctx->ws_cols) {
		ca.col0 = c0;
		ca.ncols = std::min(ctx->ws_cols, ctx->ncols - c0);
		const int block = 64;
		auto kfn = k_column_implicit;
		TB_LAUNCH_FLAT(kfn, dim3((ca.ncols + block - 1) / block), dim3(block), 0, ctx->stream,
			lay, ctx->geom, ctx->ops, ctx->phys, ca,
			(const double *)ctx->inst[in], ctx->inst[out]);
		TB_KERNEL_CHECK(ctx);
	}
	return 0;
}

// CopyData(src -> dst) followed by StepImplicit(dst, dst): the column solve
// reads src and writes rho-theta, w, rho of dst (every node is a solved column
// or the duplicate of one), so only the rows the solve leaves alone (u, v) are
// copied.
extern "C" int tb200_copy_v_step_implicit(tb200_ctx * ctx, int src, int dst, double dt) {
	if (check_inst2(ctx, src, dst)) return 1;
	const DevLayout & lay = ctx->lay;
	const bool solve = (ctx->cfg.eqn_type == TB200_EQN_PRIMITIVE_NONHYDRO) && lay.nlev > 1
		&& !ctx->cfg.fully_explicit;
	if (src == dst || !solve || !fast_with_tracers(ctx) || fast_prepare(ctx) || ctx->fast_state != 1
		|| getenv("TB200_COLUMN_KERNEL") != 0) {
		if (tb200_copy(ctx, src, dst, TB200_DATA_STATE | TB200_DATA_TRACERS)) return 1;
		return tb200_v_step_implicit(ctx, dst, dst, dt);
	}
	CombineArgs ca;
	memset(&ca, 0, sizeof(ca));
	ca.nsrc = 1;
	ca.src[0] = ctx->inst[src];
	ca.coeff[0] = 1.0;
	ca.scale_dst = 0;
	// rows of u and v (components 0 and 1 are adjacent)
	if (launch_combine(ctx, ca, dst, lay.rowoff[0], lay.rowoff[1] + lay.rowlev[1])) return 1;
	// the tracer update subtracts from the copy (VerticalDynamicsFEM.cpp:4265-4281)
	if (launch_combine(ctx, ca, dst, lay.nrows_state, lay.nrows)) return 1;
	return tb200_v_step_implicit(ctx, src, dst, dt);
}

// CopyData(src -> dst), StepImplicit(dst, dst), LinearCombineData({+1, -1} -> src):
// dst = solve(src) and src = dst - src, the tail of a Strang step without
// off-centring (TimestepSchemeStrang.cpp:644-672).  The increment of the solved
// rows is written by the column kernel itself; u and v are untouched by the
// solve, their increment is +0.
extern "C" int tb200_copy_v_step_implicit_diff(tb200_ctx * ctx, int src, int dst, double dt) {
	if (check_inst2(ctx, src, dst)) return 1;
	const DevLayout & lay = ctx->lay;
	const bool solve = (ctx->cfg.eqn_type == TB200_EQN_PRIMITIVE_NONHYDRO) && lay.nlev > 1
		&& !ctx->cfg.fully_explicit;
	if (src == dst || !solve || !fast_with_tracers(ctx) || fast_prepare(ctx) || ctx->fast_state != 1
		|| getenv("TB200_COLUMN_KERNEL") != 0) {
		if (tb200_copy_v_step_implicit(ctx, src, dst, dt)) return 1;
		const double fin[2] = {+1.0, -1.0};
		std::vector<double> c(std::max(src, dst) + 1, 0.0);
		c[dst] = fin[0];
		c[src] = fin[1];
		return tb200_lincomb(ctx, c.data(), (int)c.size(), src, TB200_DATA_STATE | TB200_DATA_TRACERS);
	}
	CombineArgs ca;
	memset(&ca, 0, sizeof(ca));
	ca.nsrc = 1;
	ca.src[0] = ctx->inst[src];
	ca.coeff[0] = 1.0;
	ca.scale_dst = 0;
	const int uv0 = lay.rowoff[0], uv1 = lay.rowoff[1] + lay.rowlev[1];
	if (launch_combine(ctx, ca, dst, uv0, uv1)) return 1;
	if (launch_combine(ctx, ca, dst, lay.nrows_state, lay.nrows)) return 1;
	ctx->column_inc = ctx->inst[src];
	// tracers: src still holds the values from before the update; the column
	// filter leaves dst - src there
	ctx->tracer_keep = 0;
	ctx->tracer_inc = (lay.ntr > 0) ? ctx->inst[src] : 0;
	const int rc = tb200_v_step_implicit(ctx, src, dst, dt);
	ctx->column_inc = 0;
	ctx->tracer_inc = 0;
	if (rc) return 1;
	// u, v rows of the increment
	CombineArgs zero;
	memset(&zero, 0, sizeof(zero));
	if (launch_combine(ctx, zero, src, uv0, uv1)) return 1;
	ctx->uvzero_inst = src;
	ctx->uvzero_writes = ctx->writes;
	return 0;
}

// Strang tail when the state to be solved already sits in the destination
// (hyperdiffusion written straight into it): StepImplicit(inst, inst) in place,
// and `inc` receives Grid::LinearCombineData({+1, -1}) of the new and the old
// state (zero in the u, v rows).  Returns 2, having done nothing, when the
// fast column kernel does not apply; the caller then takes the reference's
// sequence of calls.
extern "C" int tb200_v_step_implicit_inc_available(tb200_ctx * ctx) {
	const DevLayout & lay = ctx->lay;
	const bool solve = (ctx->cfg.eqn_type == TB200_EQN_PRIMITIVE_NONHYDRO) && lay.nlev > 1
		&& !ctx->cfg.fully_explicit;
	if (!solve || !fast_with_tracers(ctx) || fast_prepare(ctx) || ctx->fast_state != 1
		|| getenv("TB200_COLUMN_KERNEL") != 0 || getenv("TB200_CARRY_FULL") != 0) {
		return 0;
	}
	return 1;
}

extern "C" int tb200_v_step_implicit_inc(tb200_ctx * ctx, int inst, int inc, double dt) {
	if (check_inst2(ctx, inst, inc)) return 1;
	const DevLayout & lay = ctx->lay;
	if (inst == inc || !tb200_v_step_implicit_inc_available(ctx)) return 2;
	// tracers: the column kernel leaves the values from before the update in
	// `inc`, the column filter turns them into new - old
	ctx->tracer_keep = (lay.ntr > 0) ? ctx->inst[inc] : 0;
	ctx->tracer_inc = ctx->tracer_keep;
	ctx->column_inc = ctx->inst[inc];
	const int rc = tb200_v_step_implicit(ctx, inst, inst, dt);
	ctx->column_inc = 0;
	ctx->tracer_keep = 0;
	ctx->tracer_inc = 0;
	if (rc) return 1;
	// u, v rows of the increment
	const int uv0 = lay.rowoff[0], uv1 = lay.rowoff[1] + lay.rowlev[1];
	CombineArgs zero;
	memset(&zero, 0, sizeof(zero));
	if (launch_combine(ctx, zero, inc, uv0, uv1)) return 1;
	ctx->uvzero_inst = inc;
	ctx->uvzero_writes = ctx->writes;
	return 0;
}

// Debugging aid: assemble F and the banded Jacobian of the first launch chunk
// without solving and copy the workspace of one column back
// (cf. USE_JACOBIAN_DEBUG / BootstrapJacobian, VerticalDynamicsFEM.cpp:1163-1226).
extern "C" int tb200_debug_column_assembly(
	tb200_ctx * ctx, int in, double dt, int col, double * ws_out, int nentries
) {
	// nentries < 0: also run the band solve (F then holds the Newton update)
	const int mode = (nentries < 0) ? 2 : 1;
	if (nentries < 0) nentries = -nentries;
	if (check_ops(ctx)) return 1;
	const DevLayout & lay = ctx->lay;
	ColumnArgs ca;
	ca.col_node = ctx->d_col_node;
	ca.col_dups = ctx->d_col_dups;
	ca.ws = ctx->d_ws;
	ca.ws_stride = ctx->ws_cols;
	ca.dt = dt;
	ca.offd = ctx->offd;
	ca.fe_nodes = ctx->fe_nodes;
	ca.upwind_coeff = (1.0 / 2.0) * pow(1.0 / static_cast<double>(lay.nlev), 1.0);
	ca.info = ctx->d_info;
	ca.assemble_only = mode;
	if (column_uniform_args(ctx, ca)) return 1;
	ca.col0 = 0;
	ca.ncols = std::min(ctx->ws_cols, ctx->ncols);
	if (col < 0 || col >= ca.ncols) TB_FAIL(ctx, "column out of range");
	auto kfn = k_column_implicit;
	TB_LAUNCH_FLAT(kfn, dim3((ca.ncols + 63) / 64), dim3(64), 0, ctx->stream,
		lay, ctx->geom, ctx->ops, ctx->phys, ca,
		(const double *)ctx->inst[in], ctx->inst[in]);
	TB_KERNEL_CHECK(ctx);
	TB_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
	const int total = tb_column_ws_entries(lay.nlev, ctx->offd);
	std::vector<double> all((size_t)total * ctx->ws_cols);
	TB_CHECK(ctx, cudaMemcpy(all.data(), ctx->d_ws, all.size() * sizeof(double), cudaMemcpyDeviceToHost));
	for (int q = 0; q < nentries && q < total; q++) {
		ws_out[q] = all[(size_t)q * ctx->ws_cols + col];
	}
	return 0;
}

// Deferred error of the column solve ("Solution failed" / "Inversion failure",
// VerticalDynamicsFEM.cpp:1461-1481); checked at tb200_sync-like points.
static int check_column_info(tb200_ctx * ctx) {
	if (ctx->d_info == 0) return 0;
	int both[2] = {0, 0};
	TB_CHECK(ctx, cudaMemcpy(both, ctx->d_info, 2 * sizeof(int), cudaMemcpyDeviceToHost));
	if (both[1] != 0) {
		char buf[128];
		snprintf(buf, 128, "peer-memory exchange: rank %d did not signal within the time limit", both[1] - 1);
		TB_FAIL(ctx, buf);
	}
	const int info = both[0];
	if (info != 0) {
		char buf[128];
		snprintf(buf, 128, "Inversion failure in column %d", info - 1);
		int zero = 0;
		cudaMemcpy(ctx->d_info, &zero, sizeof(int), cudaMemcpyHostToDevice);
		TB_FAIL(ctx, buf);
	}
	return 0;
}

extern "C" int tb200_check_errors(tb200_ctx * ctx) {
	TB_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
	return check_column_info(ctx);
}

// Stream-ordered copy of the device-side failure record into pinned host
// memory (end of every tb200_step), and the test on it that needs no
// synchronisation (start of every tb200_step): a failed column solve or a peer
// that stopped signalling stops the run at most two steps later even when the
// caller never asks (the reference throws at once, VerticalDynamicsFEM.cpp:1461-1481).
int tb_mirror_errors(tb200_ctx * ctx) {
	if (ctx->d_info == 0 || ctx->h_info == 0) return 0;
	TB_CHECK(ctx, cudaMemcpyAsync(ctx->h_info, ctx->d_info, 2 * sizeof(int),
		cudaMemcpyDeviceToHost, ctx->stream));
	return 0;
}

int tb_poll_errors(tb200_ctx * ctx) {
	if (ctx->h_info == 0) return 0;
	const volatile int * h = ctx->h_info;
	if (h[0] == 0 && h[1] == 0) return 0;
	return tb200_check_errors(ctx);
}

// HorizontalDynamicsFEM::FilterNegativeTracers (element-wise) and
// VerticalDynamicsFEM::FilterNegativeTracers (column-wise); both need the
// element areas (tb200_upload_element_area).
static int filter_tracers(
	tb200_ctx * ctx, int inst, bool column, const CombineArgs * comb = 0, double * inc = 0
) {
	const DevLayout & lay = ctx->lay;
	if (lay.ntr == 0) return 0;
	if (inst < 0 || inst >= (int)ctx->inst.size()) TB_FAIL(ctx, "invalid state instance");
	if (ctx->d_area_node == 0) TB_FAIL(ctx, "element areas not uploaded (tracer filter)");
	const int block = 128;
	if (column) {
		const long long nitems = lay.nelem * lay.nn * lay.ntr;
		auto kfn = k_filter_tracers_column;
		CombineArgs none;
		memset(&none, 0, sizeof(none));
		TB_LAUNCH_FLAT(kfn, dim3((unsigned)((nitems + block - 1) / block)), dim3(block), 0,
			ctx->stream, lay, (const double *)ctx->d_area_node, ctx->inst[inst],
			(comb != 0) ? *comb : none, (comb != 0) ? 1 : 0, inc);
	} else {
		const long long nitems = lay.nelem * lay.ntr * lay.nlev;
		auto kfn = k_filter_tracers_element;
		TB_LAUNCH_FLAT(kfn, dim3((unsigned)((nitems + block - 1) / block)), dim3(block), 0,
			ctx->stream, lay, (const double *)ctx->d_area_node, ctx->inst[inst]);
	}
	TB_KERNEL_CHECK(ctx);
	return 0;
}

extern "C" int tb200_filter_negative_tracers(tb200_ctx * ctx, int inst) {
	return filter_tracers(ctx, inst, false);
}

extern "C" int tb200_v_filter_negative_tracers(tb200_ctx * ctx, int inst) {
	return filter_tracers(ctx, inst, true);
}

// Grid::LinearCombineData(coeff -> dst) of state and tracers followed by
// VerticalDynamics::FilterNegativeTracers(dst): the carry-over at the start of a
// Strang step (TimestepSchemeStrang.cpp:470-482).  The combination of the tracer
// rows is formed inside the filter kernel.
extern "C" int tb200_lincomb_v_filter(tb200_ctx * ctx, const double * coeff, int ncoeff, int dst) {
	const DevLayout & lay = ctx->lay;
	if (lay.ntr == 0 || ctx->d_area_node == 0) {
		if (tb200_lincomb(ctx, coeff, ncoeff, dst, TB200_DATA_STATE | TB200_DATA_TRACERS)) return 1;
		return tb200_v_filter_negative_tracers(ctx, dst);
	}
	if (tb200_lincomb(ctx, coeff, ncoeff, dst, TB200_DATA_STATE)) return 1;
	CombineArgs ca;
	memset(&ca, 0, sizeof(ca));
	ca.cdst = coeff[dst];
	ca.scale_dst = (coeff[dst] == 0.0) ? 0 : 1;
	for (int m = 0; m < ncoeff; m++) {
		if (m == dst || coeff[m] == 0.0) continue;
		ca.src[ca.nsrc] = ctx->inst[m];
		ca.coeff[ca.nsrc] = coeff[m];
		ca.nsrc++;
	}
	return filter_tracers(ctx, dst, true, &ca, 0);
}

///////////////////////////////////////////////////////////////////////////////
// DSS

static int ensure_buffers(tb200_ctx * ctx, size_t rows) {
	if (ctx->nranks == 1 || rows <= ctx->buf_rows) return 0;
	if (dalloc(ctx, &ctx->d_sendbuf, (size_t)ctx->nsend_total * rows)) return 1;
	if (dalloc(ctx, &ctx->d_recvbuf, (size_t)ctx->nrecv_total * rows)) return 1;
	ctx->buf_rows = rows;
	return 0;
}

// ---- peer-memory exchange -----------------------------------------------------
static const size_t kPeerFlagDoubles = 64;     // 512 bytes of flags in front of the buffers

static double * peer_buffer(void * base, int64_t recv_total, size_t rows, int parity) {
	return (double *)base + kPeerFlagDoubles + (size_t)parity * (size_t)recv_total * rows;
}

// how long a rank waits for a neighbour's flag before reporting it missing:
// TB200_PEER_TIMEOUT_S seconds, default 120 (ranks may legitimately drift apart,
// e.g. while one writes output)
static unsigned long long peer_timeout_ns() {
	static const unsigned long long ns = []() {
		const char * e = getenv("TB200_PEER_TIMEOUT_S");
		const double sec = (e != 0 && atof(e) > 0.0) ? atof(e) : 120.0;
		return (unsigned long long)(sec * 1e9);
	}();
	return ns;
}

static int peer_exchange(tb200_ctx * ctx, int inst, int row0, int nsel, const double ** recvbuf) {
	const DevLayout & lay = ctx->lay;
	if ((size_t)nsel > ctx->peer_rows) TB_FAIL(ctx, "peer exchange: more rows than the buffers hold");
	const unsigned long long seq = ++ctx->peer_seq;
	const int parity = (int)(seq & 1);
	PeerPtrs pp;
	memset(&pp, 0, sizeof(pp));
	unsigned wait_mask = 0;
	bool any_send = false;
	for (int r = 0; r < ctx->nranks; r++) {
		if (r == ctx->rank || ctx->peer_base[r] == 0) continue;
		pp.recv[r] = peer_buffer(ctx->peer_base[r], ctx->peer_recv_total[r], ctx->peer_rows, parity);
		if (ctx->send_count[r] > 0) {
			pp.flag[r] = (unsigned long long *)ctx->peer_base[r] + ctx->rank;
			any_send = true;
		}
		if (ctx->recv_count[r] > 0) wait_mask |= (1u << r);
	}
	if (ctx->nsend_total > 0) {
		const long long total = (long long)ctx->nsend_total * nsel;
		long long nb = (total + 255) / 256;
		if (nb > 148 * 8) nb = 148 * 8;
		auto kfn = k_dss_pack_peer;
		TB_LAUNCH_FLAT(kfn, dim3((unsigned)nb), dim3(256), 0, ctx->stream,
			lay, (const int *)ctx->d_send_nodes, (const int *)ctx->d_send_rank,
			(const int *)ctx->d_send_slot, ctx->nsend_total,
			(const double *)ctx->inst[inst], pp, row0, nsel,
			ctx->d_peer_ticket, ctx->nranks, ctx->rank, seq);
		TB_KERNEL_CHECK(ctx);
	}
	(void)any_send;
	(void)wait_mask;
	// no signal / wait launches: the last block of the pack kernel raises the
	// flags, the averaging kernel waits for the ranks each of its groups reads from
	*recvbuf = peer_buffer(ctx->peer_area, ctx->nrecv_total, ctx->peer_rows, parity);
	return 0;
}

extern "C" int tb200_peer_export(
	tb200_ctx * ctx, void * handle, int64_t * recv_offsets, int64_t * recv_total
) {
#ifdef TB200_EMU
	(void)handle; (void)recv_offsets; (void)recv_total;
	TB_FAIL(ctx, "peer-memory exchange needs the CUDA build");
#else
	if (!ctx->connectivity_built) TB_FAIL(ctx, "connectivity not built");
	if (ctx->nranks < 2) TB_FAIL(ctx, "peer-memory exchange needs more than one rank");
	if (ctx->nranks > TB200_MAX_PEERS) TB_FAIL(ctx, "peer-memory exchange: too many ranks");
	if (ctx->peer_area != 0) TB_FAIL(ctx, "peer-memory exchange already exported");
	const DevLayout & lay = ctx->lay;
	ctx->peer_rows = (size_t)lay.nrows;
	const size_t doubles = kPeerFlagDoubles + 2 * (size_t)ctx->nrecv_total * ctx->peer_rows;
	TB_CHECK(ctx, cudaMalloc(&ctx->peer_area, doubles * sizeof(double)));
	TB_CHECK(ctx, cudaMemset(ctx->peer_area, 0, doubles * sizeof(double)));
	cudaIpcMemHandle_t h;
	TB_CHECK(ctx, cudaIpcGetMemHandle(&h, ctx->peer_area));
	static_assert(sizeof(h) == 64, "IPC handle size");
	memcpy(handle, &h, sizeof(h));
	int64_t off = 0;
	for (int r = 0; r < ctx->nranks; r++) {
		recv_offsets[r] = off;
		off += ctx->recv_count[r];
	}
	*recv_total = ctx->nrecv_total;
	if (ctx->d_info == 0) {
		if (dalloc(ctx, &ctx->d_info, 4)) return 1;
		TB_CHECK(ctx, cudaMemset(ctx->d_info, 0, 4 * sizeof(int)));
	}
	if (ctx->d_peer_ticket == 0) {
		if (dalloc(ctx, &ctx->d_peer_ticket, 1)) return 1;
		TB_CHECK(ctx, cudaMemset(ctx->d_peer_ticket, 0, sizeof(unsigned)));
	}
	return 0;
#endif
}

extern "C" int tb200_peer_attach(
	tb200_ctx * ctx, const void * handles, const int64_t * my_offset_at,
	const int64_t * recv_totals
) {
#ifdef TB200_EMU
	(void)handles; (void)my_offset_at; (void)recv_totals;
	TB_FAIL(ctx, "peer-memory exchange needs the CUDA build");
#else
	if (ctx->peer_area == 0) TB_FAIL(ctx, "tb200_peer_export first");
	ctx->peer_base.assign(ctx->nranks, (void *)0);
	ctx->peer_recv_total.assign(recv_totals, recv_totals + ctx->nranks);
	for (int r = 0; r < ctx->nranks; r++) {
		if (r == ctx->rank || (ctx->send_count[r] == 0 && ctx->recv_count[r] == 0)) continue;
		cudaIpcMemHandle_t h;
		memcpy(&h, (const char *)handles + (size_t)r * 64, sizeof(h));
		void * p = 0;
		TB_CHECK(ctx, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
		ctx->peer_base[r] = p;
	}
	std::vector<int> slot(ctx->send_rank.size());
	for (size_t q = 0; q < slot.size(); q++) {
		slot[q] = (int)(my_offset_at[ctx->send_rank[q]] + ctx->send_j[q]);
	}
	if (dupload(ctx, &ctx->d_send_rank, ctx->send_rank)) return 1;
	if (dupload(ctx, &ctx->d_send_slot, slot)) return 1;
	ctx->peer_ready = true;
	return 0;
#endif
}

extern "C" int tb200_peer_detach(tb200_ctx * ctx) {
	ctx->peer_ready = false;      // back to the callback; mappings are released at destroy
	return 0;
}

static int dss_rows(
	tb200_ctx * ctx, int inst, int row0, int row1, bool is_state, bool remainder_only = false
) {
	if (row1 <= row0) return 0;
	const DevLayout & lay = ctx->lay;
	const int nsel = row1 - row0;
	const double * recvbuf = ctx->d_recvbuf;
	// Grid::Exchange ("Communicate", Grid.cpp:636): one per DSS, also on one rank
	// (the reference exchanges halos with itself)
	TimingScope ts(ctx, "Communicate");
	if (ctx->nranks > 1 && ctx->peer_ready) {
		if (peer_exchange(ctx, inst, row0, nsel, &recvbuf)) return 1;
	} else if (ctx->nranks > 1) {
		if (ensure_buffers(ctx, (size_t)lay.nrows)) return 1;
		if (ctx->nsend_total > 0) {
			const long long total = (long long)ctx->nsend_total * nsel;
			long long nb = (total + 255) / 256;
			if (nb > 148 * 8) nb = 148 * 8;
			auto kfn = k_dss_pack;
			TB_LAUNCH_FLAT(kfn, dim3((unsigned)nb), dim3(256), 0, ctx->stream,
				lay, (const int *)ctx->d_send_nodes, ctx->nsend_total,
				(const double *)ctx->inst[inst], ctx->d_sendbuf, row0, nsel);
			TB_KERNEL_CHECK(ctx);
		}
		std::vector<int64_t> sc(ctx->nranks), rc(ctx->nranks);
		for (int r = 0; r < ctx->nranks; r++) {
			sc[r] = ctx->send_count[r] * nsel;
			rc[r] = ctx->recv_count[r] * nsel;
		}
		if (ctx->exch_fn(ctx->exch_user, ctx->d_sendbuf, ctx->d_recvbuf,
				sc.data(), rc.data(), ctx->nranks) != 0) {
			TB_FAIL(ctx, "exchange callback failed");
		}
		recvbuf = ctx->d_recvbuf;
	}
	// elements that do not feed the exchange may still be in flight on stream2
	if (split_join(ctx)) return 1;
	if (ctx->ngroups == 0) return 0;
	DssArgs a;
	a.members = ctx->d_members;
	a.flags = ctx->d_flags;
	a.ngroups = ctx->ngroups;
	if (remainder_only) {
		// the groups the fused kernels left raw
		a.members = ctx->d_rem_members;
		a.flags = ctx->d_rem_flags;
		a.ngroups = ctx->nrem;
	}
	a.nlocal = (int)(lay.nelem * lay.nn);
	a.recv = recvbuf;
	a.row0 = row0;
	a.row1 = row1;
	a.uv_row0 = is_state ? lay.rowoff[0] : -1;
	a.uv_row1 = is_state ? (lay.rowoff[1] + lay.rowlev[1]) : -1;
	a.nsel = nsel;
	a.sel_row0 = row0;
	a.rows_fastest = 0;
	a.peer_flags = 0;
	a.peer_seq = 0;
	a.peer_timeout_ns = 0;
	a.info = ctx->d_info;
	if (ctx->nranks > 1 && ctx->peer_ready) {
		a.peer_flags = (const unsigned long long *)ctx->peer_area;
		a.peer_seq = ctx->peer_seq;
		a.peer_timeout_ns = peer_timeout_ns();
	}
	if (a.ngroups > 0) {
		const int block = 128;
		static const bool classes = []() {  // TB200_DSS_KERNEL=generic: row-at-a-time kernel for every group
			const char * e = getenv("TB200_DSS_KERNEL");
			return !(e != 0 && strcmp(e, "generic") == 0);
		}();
		// one batch of TBD_B rows per thread measured best (0.76 ms against
		// 0.79 ms with 38 rows per thread at ne=120, L=30); the row-at-a-time
		// kernel wants its prologue amortised over more rows
		int gy = classes ? (nsel + TBD_B - 1) / TBD_B : std::min(nsel, 4);
		{
			const char * g = getenv("TB200_DSS_GY");
			if (g != 0 && atoi(g) > 0) gy = std::min(nsel, atoi(g));
		}
		// block order: consecutive blocks walk the row chunks of the same groups,
		// i.e. an element's contiguous block of rows is streamed by blocks that run
		// together (ne = 120, L = 30: 0.757 -> 0.711 ms per pass, two batches of
		// TBD_B rows per block; TB200_DSS_ORDER=groups: groups fastest, one batch)
		// On several ranks the passes are short and interleaved with the exchange:
		// measured on 8 GPUs the old order wins there (2.60 against 2.64 ms per step,
		// twice each), so the new one is the default on one rank only
		// (TB200_DSS_ORDER=rows / groups forces either).
		static const int order_env = []() {
			const char * e = getenv("TB200_DSS_ORDER");
			if (e != 0 && strcmp(e, "groups") == 0) return 0;
			if (e != 0 && strcmp(e, "rows") == 0) return 1;
			return -1;
		}();
		const bool rows_fastest = (order_env >= 0) ? (order_env == 1) : (ctx->nranks == 1);
		const int gx = (a.ngroups + block - 1) / block;
		if (rows_fastest && classes && gx <= 65535 && getenv("TB200_DSS_GY") == 0) {
			gy = (nsel + 2 * TBD_B - 1) / (2 * TBD_B);
		}
		a.rows_fastest = (rows_fastest && gx <= 65535) ? 1 : 0;
		const dim3 grid = a.rows_fastest ? dim3(gy, gx) : dim3(gx, gy);
		if (classes) {
			auto kfn = k_dss_fast;
			TB_LAUNCH_FLAT(kfn, grid, dim3(block), 0,
				ctx->stream, lay, a, ctx->inst[inst]);
		} else {
			auto kfn = k_dss_scalar;
			TB_LAUNCH_FLAT(kfn, grid, dim3(block), 0,
				ctx->stream, lay, a, ctx->inst[inst]);
		}
		TB_KERNEL_CHECK(ctx);
	}
	if (is_state && ctx->nseam > 0) {
		// seam groups index the full group list
		a.members = ctx->d_members;
		a.flags = ctx->d_flags;
		a.ngroups = ctx->ngroups;
		SeamArgs sa;
		sa.group = ctx->d_seam_group;
		sa.mats = ctx->d_seam_mats;
		sa.nseam = ctx->nseam;
		sa.nlev_u = lay.rowlev[0];
		const int total = sa.nseam * sa.nlev_u;
		auto kfn = k_dss_seam_vector;
		TB_LAUNCH_FLAT(kfn, dim3((total + 127) / 128), dim3(128), 0, ctx->stream,
			lay, a, sa, ctx->inst[inst]);
		TB_KERNEL_CHECK(ctx);
	}
	return 0;
}

// DSS of rows that hold scalars (no covector re-basing at panel seams):
// the vorticity of the shallow-water enstrophy diagnostic
static int tb_dss_scalar_rows(tb200_ctx * ctx, int inst, int row0, int row1) {
	if (!ctx->connectivity_built) TB_FAIL(ctx, "connectivity not built");
	return dss_rows(ctx, inst, row0, row1, false);
}

static int dss_instance(tb200_ctx * ctx, int inst, int mask, bool remainder_only) {
	if (!ctx->connectivity_built) TB_FAIL(ctx, "connectivity not built");
	if (inst < 0 || inst >= (int)ctx->inst.size()) TB_FAIL(ctx, "invalid state instance");
	const DevLayout & lay = ctx->lay;
	if ((mask & TB200_DATA_STATE) && (mask & TB200_DATA_TRACERS) && lay.ntr > 0 && !remainder_only) {
		// state and tracers in one exchange and one averaging pass (the reference
		// exchanges them together as well, Grid::Exchange: Grid.cpp:627-685)
		return dss_rows(ctx, inst, 0, lay.nrows, true, false);
	}
	if (mask & TB200_DATA_STATE) {
		if (dss_rows(ctx, inst, 0, lay.nrows_state, true, remainder_only)) return 1;
	}
	if ((mask & TB200_DATA_TRACERS) && lay.ntr > 0) {
		if (dss_rows(ctx, inst, lay.nrows_state, lay.nrows, false, remainder_only)) return 1;
	}
	return 0;
}

extern "C" int tb200_dss(tb200_ctx * ctx, int inst, int mask) {
	return dss_instance(ctx, inst, mask, false);
}

///////////////////////////////////////////////////////////////////////////////
// Hyperdiffusion (HorizontalDynamicsFEM::StepAfterSubCycle, :2637-2726)

// component >= 0: that state component alone (iComponent, :1988-1996), ref != 0: the
// reference state is removed from the field first (fRemoveRefState)
static int hyper_scalar(
	tb200_ctx * ctx, int in, int out, double dt, double nu, bool scale,
	int component = -1, const double * ref = 0
) {
	const DevLayout & lay = ctx->lay;
	if (need_metric3d(ctx, "scalar hyperdiffusion (general kernel)")) return 1;
	HyperRows hr;
	memset(&hr, 0, sizeof(hr));
	int nsel = 0;
	// components 2.. of the state (:1983-1993), then every tracer
	for (int c = 2; c < lay.ncomp; c++) {
		if (component >= 0 && c != component) continue;
		hr.row0[hr.nranges] = lay.rowoff[c];
		hr.row1[hr.nranges] = lay.rowoff[c] + lay.rowlev[c];
		hr.onedge[hr.nranges] = lay.onedge[c];
		nsel += lay.rowlev[c];
		hr.nranges++;
	}
	if (lay.ntr > 0 && component < 0) {
		hr.row0[hr.nranges] = lay.troff;
		hr.row1[hr.nranges] = lay.nrows;
		hr.onedge[hr.nranges] = 0;
		nsel += lay.ntr * lay.nlev;
		hr.nranges++;
	}
	const long long nitems = lay.nelem * nsel;
	TB_NP_SWITCH(lay.np,
		auto kfn = k_hyper_scalar<NPV, kItems>;
		TB_LAUNCH(kfn, dim3((unsigned)((nitems + kItems - 1) / kItems)), dim3(NPV * NPV * kItems), 0,
			ctx->stream, lay, ctx->geom, ctx->tables, hr, nsel,
			(const double *)ctx->inst[in], ctx->inst[out], dt, nu, scale ? 1 : 0, ref);)
	TB_KERNEL_CHECK(ctx);
	return 0;
}

static int hyper_vector_from(
	tb200_ctx * ctx, const double * in, int out, double dt, double nud, double nuv, bool scale
) {
	const DevLayout & lay = ctx->lay;
	const long long nitems = lay.nelem * lay.nlev;
	TB_NP_SWITCH(lay.np,
		auto kfn = k_hyper_vector<NPV, kItems>;
		TB_LAUNCH(kfn, dim3((unsigned)((nitems + kItems - 1) / kItems)), dim3(NPV * NPV * kItems), 0,
			ctx->stream, lay, ctx->geom, ctx->tables,
			in, ctx->inst[out], dt, nud, nuv, scale ? 1 : 0,
			ctx->cfg.cartesian_xz);)
	TB_KERNEL_CHECK(ctx);
	return 0;
}

static int hyper_vector(
	tb200_ctx * ctx, int in, int out, double dt, double nud, double nuv, bool scale
) {
	return hyper_vector_from(ctx, (const double *)ctx->inst[in], out, dt, nud, nuv, scale);
}

// Uniform diffusion at the end of HorizontalDynamicsFEM::StepExplicit (:1817-1858):
// second-order diffusion of the velocity, of rho theta and of w, each minus the same
// operator on the reference state.
static int uniform_diffusion_horizontal(tb200_ctx * ctx, int in, int out, double dt) {
	if (ctx->lay.ntr > 0) {
		// (the reference's implicit column update of the tracers throws "Not
		// implemented" with uniform diffusion, VerticalDynamicsFEM.cpp:3914-3917)
		TB_FAIL(ctx, "uniform diffusion with tracers is not implemented");
	}
	if (refstate_alloc(ctx)) return 1;
	const double nuv = ctx->uniform_v;
	const double nus = ctx->uniform_s;
	if (hyper_vector(ctx, in, out, dt, -nuv, -nuv, false)) return 1;
	if (hyper_vector_from(ctx, (const double *)ctx->d_refstate, out, dt, nuv, nuv, false)) return 1;
	if (ctx->cfg.eqn_type == TB200_EQN_PRIMITIVE_NONHYDRO) {
		if (hyper_scalar(ctx, in, out, dt, nus, false, 2, ctx->d_refstate)) return 1;
		if (hyper_scalar(ctx, in, out, dt, nuv, false, 3, ctx->d_refstate)) return 1;
	}
	return 0;
}

// Uniform diffusion of u and v in the column (VerticalDynamicsFEM::StepExplicit,
// :1058-1106), after the vertical advection of the velocity.
//
// What the reference differentiates there is its work array m_dStateNode[UIx/VIx],
// which StepExplicit fills only under --explicitvertical (SetupReferenceColumn,
// :751-756; the copy at the head of the column loop is commented out, :724-745).
// Otherwise the array still holds the column SetupReferenceColumn saw last: the
// last column of the most recent StepImplicit - the last interior node of the last
// active patch, u and v of that call's initial instance (:1334-1345) - or zeros
// before the first StepImplicit.  Every node then gets dt nu / ztop^2 (DD of that one
// column - DD of its own reference column).  Restated as it is: the results are
// the reference's.
static int stale_column_alloc(tb200_ctx * ctx) {
	if (ctx->d_stale_uv != 0) return 0;
	const size_t n = (size_t)2 * ctx->lay.nlev;
	if (dalloc(ctx, &ctx->d_stale_uv, n)) return 1;
	TB_CHECK(ctx, cudaMemset(ctx->d_stale_uv, 0, n * sizeof(double)));
	return 0;
}

// after StepImplicit(in, .): remember u, v of its last column
static int stale_column_update(tb200_ctx * ctx, int in) {
	if (!uniform_on(ctx) || ctx->cfg.fully_explicit) return 0;
	if (stale_column_alloc(ctx)) return 1;
	const PatchInfo * last = 0;
	for (size_t p = 0; p < ctx->patches.size(); p++) {
		const PatchInfo & pi = ctx->patches[p];
		if (pi.elem0 >= 0 && (last == 0 || pi.index > last->index)) last = &pi;
	}
	if (last == 0) return 0;
	const DevLayout & lay = ctx->lay;
	// element (nea - 1, neb - 1) of the patch, node (np - 1, np - 1)
	const long long e = last->elem0 + (long long)last->nea * last->neb - 1;
	const long long node = e * lay.nn + (lay.nn - 1);
	auto kfn = k_stale_column_uv;
	TB_LAUNCH_FLAT(kfn, dim3(1), dim3(64), 0, ctx->stream,
		lay, (const double *)ctx->inst[in], node, ctx->d_stale_uv);
	TB_KERNEL_CHECK(ctx);
	return 0;
}

static int uniform_diffusion_vertical_uv(tb200_ctx * ctx, int in, int out, double dt) {
	const DevLayout & lay = ctx->lay;
	if (check_ops(ctx)) return 1;
	if (refstate_alloc(ctx)) return 1;
	if (stale_column_alloc(ctx)) return 1;
	const double ztop = ctx->cfg.ztop;
	const double coeff = ctx->uniform_v / (ztop * ztop);
	const long long total = lay.nelem * (long long)lay.nn;
	auto kfn = k_uniform_diffusion_uv;
	TB_LAUNCH_FLAT(kfn, dim3((unsigned)((total + 127) / 128)), dim3(128), 0, ctx->stream,
		lay, ctx->ops, (const double *)ctx->inst[in], (const double *)ctx->d_refstate,
		ctx->inst[out], dt, coeff,
		ctx->cfg.fully_explicit ? (const double *)0 : (const double *)ctx->d_stale_uv);
	TB_KERNEL_CHECK(ctx);
	return 0;
}

static bool hyper_fast_ok(tb200_ctx * ctx) {
	if (fast_prepare(ctx)) return false;
	const char * force = getenv("TB200_HYPER_KERNEL");
	if (force != 0 && strcmp(force, "generic") == 0) return false;
	return ctx->fast_state == 1 && fast_with_tracers(ctx);
}

// out = (base >= 0 ? inst[base] : 0) - dt nu L(inst[fld]) on the tracer rows
// (+ the positivity filter of HorizontalDynamicsFEM.cpp:2717 after the last pass)
static int hyper_tracers_fast(
	tb200_ctx * ctx, int fld, int base, int out, double dt, double nus, bool scale, bool filter
) {
	const DevLayout & lay = ctx->lay;
	if (lay.ntr == 0) return 0;
	if (filter && ctx->d_area_node == 0) TB_FAIL(ctx, "element areas not uploaded (tracer filter)");
	HyperFastArgs ha;
	memset(&ha, 0, sizeof(ha));
	ha.colc = ctx->d_colc;
	ha.inv_da = ctx->d_inv_da;
	ha.inv_db = ctx->d_inv_db;
	ha.nu_scale = ctx->d_nu_scale;
	ha.dt = dt;
	ha.nu_scalar = nus;
	ha.scale_nu = scale ? 1 : 0;
	const ElemList el = elem_list(ctx, 0);
	if (el.n == 0) return 0;
	const double * area = filter ? ctx->d_area_node : 0;
	if (base >= 0) {
		auto kfn = k_tracer_hyper<true>;
		TB_LAUNCH(kfn, dim3((unsigned)el.n), dim3(TBT_THREADS), 0, ctx->stream,
			lay, ctx->tables, ha, area,
			(const double *)ctx->inst[fld], (const double *)ctx->inst[base], ctx->inst[out], el);
	} else {
		auto kfn = k_tracer_hyper<false>;
		TB_LAUNCH(kfn, dim3((unsigned)el.n), dim3(TBT_THREADS), 0, ctx->stream,
			lay, ctx->tables, ha, area,
			(const double *)ctx->inst[fld], (const double *)0, ctx->inst[out], el);
	}
	ctx->launches++;
	ctx->writes++;
	TB_LAUNCH_CHECK(ctx);
	return 0;
}

// out = (base >= 0 ? inst[base] : 0) - dt nu L(inst[fld]), all prognostic fields
static int hyper_fast(
	tb200_ctx * ctx, int fld, int base, int out, double dt,
	double nus, double nud, double nuv, bool scale
) {
	const DevLayout & lay = ctx->lay;
	HyperFastArgs ha;
	ha.colc = ctx->d_colc;
	ha.inv_da = ctx->d_inv_da;
	ha.inv_db = ctx->d_inv_db;
	ha.nu_scale = ctx->d_nu_scale;
	ha.dt = dt;
	ha.nu_scalar = nus;
	ha.nu_div = nud;
	ha.nu_vort = nuv;
	ha.scale_nu = scale ? 1 : 0;
	ha.xz = ctx->cfg.cartesian_xz;
	const bool has_base = (base >= 0);
	bool fuse = fuse_enabled(ctx);
	if (fuse && tb_hyper_smem_doubles(lay.nrows_state, lay.nlev, has_base, true) * sizeof(double)
			> 227 * 1024 - 1024) fuse = false;
	const size_t smem = tb_hyper_smem_doubles(lay.nrows_state, lay.nlev, has_base, fuse) * sizeof(double);
	if (smem > 227 * 1024 - 1024) TB_FAIL(ctx, "column too tall for the fused hyperdiffusion kernel");
	const dim3 block(TBF_THREADS);
	const FuseArgs fz0 = fuse_args(ctx);
	PipeMaps maps;
	maps.in = tensor_map_of(ctx, ctx->inst[fld]);
	maps.b0 = tensor_map_of(ctx, ctx->inst[has_base ? base : fld]);
	maps.b1 = maps.b0;
	if (fuse) {
		// this launch also averages the in-patch groups of `out` (fused DSS)
		ctx->fuse_epoch++;
		const FuseArgs fz = fuse_args(ctx);
		if (has_base) {
			auto kfn = k_hyper_pipe<true, true>;
#ifndef TB200_EMU
			TB_CHECK(ctx, cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
#endif
			const dim3 grid((unsigned)fused_blocks(ctx, kfn, TBF_THREADS, smem));
			TB_LAUNCH(kfn, grid, block, smem, ctx->stream, lay, ctx->tables, ha,
				(const double *)ctx->inst[fld], (const double *)ctx->inst[base], ctx->inst[out],
				elem_list(ctx, 0), fz, maps);
		} else {
			auto kfn = k_hyper_pipe<false, true>;
#ifndef TB200_EMU
			TB_CHECK(ctx, cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
#endif
			const dim3 grid((unsigned)fused_blocks(ctx, kfn, TBF_THREADS, smem));
			TB_LAUNCH(kfn, grid, block, smem, ctx->stream, lay, ctx->tables, ha,
				(const double *)ctx->inst[fld], (const double *)0, ctx->inst[out],
				elem_list(ctx, 0), fz, maps);
		}
		ctx->launches++;
		ctx->writes++;
		ctx->fuse_done = true;
		TB_LAUNCH_CHECK(ctx);
		return 0;
	}
	if (has_base) {
		auto kfn = k_hyper_pipe<true, false>;
#ifndef TB200_EMU
		TB_CHECK(ctx, cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
#endif
		const bool split = split_enabled(ctx);
		if (split && split_fork(ctx)) return 1;
		for (int part = split ? 1 : 0; part <= (split ? 2 : 0); part++) {
			const ElemList el = elem_list(ctx, part);
			if (el.n == 0) continue;
			const dim3 grid((unsigned)persistent_blocks(ctx, kfn, TBF_THREADS, smem, el.n,
				(part == 2) ? overlap_reserve() : 0));
			TB_LAUNCH(kfn, grid, block, smem, (part == 2) ? ctx->stream2 : ctx->stream,
				lay, ctx->tables, ha,
				(const double *)ctx->inst[fld], (const double *)ctx->inst[base], ctx->inst[out], el, fz0, maps);
			ctx->launches++;
			ctx->writes++;
		}
		if (split && split_mark(ctx)) return 1;
	} else {
		auto kfn = k_hyper_pipe<false, false>;
#ifndef TB200_EMU
		TB_CHECK(ctx, cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
#endif
		const bool split = split_enabled(ctx);
		if (split && split_fork(ctx)) return 1;
		for (int part = split ? 1 : 0; part <= (split ? 2 : 0); part++) {
			const ElemList el = elem_list(ctx, part);
			if (el.n == 0) continue;
			const dim3 grid((unsigned)persistent_blocks(ctx, kfn, TBF_THREADS, smem, el.n,
				(part == 2) ? overlap_reserve() : 0));
			TB_LAUNCH(kfn, grid, block, smem, (part == 2) ? ctx->stream2 : ctx->stream,
				lay, ctx->tables, ha,
				(const double *)ctx->inst[fld], (const double *)0, ctx->inst[out], el, fz0, maps);
			ctx->launches++;
			ctx->writes++;
		}
		if (split && split_mark(ctx)) return 1;
	}
	TB_LAUNCH_CHECK(ctx);
	return 0;
}

static int h_step_after_subcycle_impl(
	tb200_ctx * ctx, int in, int out, int work, double dt
) {
	const int ni = (int)ctx->inst.size();
	if (in < 0 || in >= ni || out < 0 || out >= ni || work < 0 || work >= ni) {
		TB_FAIL(ctx, "invalid state instance");
	}
	if (in == work) TB_FAIL(ctx, "Invalid indices -- initial and working data must be distinct");
	if (out == work) TB_FAIL(ctx, "Invalid indices -- working and update data must be distinct");
	const tb200_config & c = ctx->cfg;
	const int all = TB200_DATA_STATE | TB200_DATA_TRACERS;
	const bool active = !((c.nu_scalar == 0.0) && (c.nu_div == 0.0) && (c.nu_vort == 0.0));
	if (!(active && c.hypervis_order == 4 && hyper_fast_ok(ctx))) {
		if (tb200_copy(ctx, in, out, all)) return 1;
	}
	if ((c.nu_scalar == 0.0) && (c.nu_div == 0.0) && (c.nu_vort == 0.0)) {
	} else if (c.hypervis_order == 0) {
	} else if (c.hypervis_order == 2) {
		if (hyper_scalar(ctx, in, out, dt, c.nu_scalar, false)) return 1;
		if (hyper_vector(ctx, in, out, -dt, c.nu_div, c.nu_vort, false)) return 1;
		if (tb200_filter_negative_tracers(ctx, out)) return 1;
		if (tb200_dss(ctx, out, all)) return 1;
	} else if (c.hypervis_order == 4 && hyper_fast_ok(ctx)) {
		// both Laplacian applications with ZeroData / CopyData folded in; each is
		// followed by a DSS whose exchange overlaps the elements that do not feed it
		const bool want = overlap_wanted();
		ctx->want_split = want;
		ctx->fuse_want = true;
		ctx->fuse_done = false;
		int rc = hyper_fast(ctx, in, -1, work, 1.0, 1.0, 1.0, 1.0, false);
		if (rc == 0) rc = hyper_tracers_fast(ctx, in, -1, work, 1.0, 1.0, false, false);
		ctx->want_split = false;
		if (rc == 0) rc = dss_instance(ctx, work, all, ctx->fuse_done);
		if (split_join(ctx)) return 1;
		if (rc) { ctx->fuse_want = false; return 1; }
		ctx->want_split = want;
		ctx->fuse_done = false;
		rc = hyper_fast(ctx, work, in, out, -dt, c.nu_scalar, c.nu_div, c.nu_vort, true);
		if (rc == 0) rc = hyper_tracers_fast(ctx, work, in, out, -dt, c.nu_scalar, true, true);
		ctx->want_split = false;
		ctx->fuse_want = false;
		if (rc == 0) rc = dss_instance(ctx, out, all, ctx->fuse_done);
		ctx->fuse_done = false;
		if (split_join(ctx)) return 1;
		return rc;
	} else if (c.hypervis_order == 4) {
		if (tb200_zero(ctx, work, all)) return 1;
		if (hyper_scalar(ctx, in, work, 1.0, 1.0, false)) return 1;
		if (hyper_vector(ctx, in, work, 1.0, 1.0, 1.0, false)) return 1;
		if (tb200_dss(ctx, work, all)) return 1;
		if (hyper_scalar(ctx, work, out, -dt, c.nu_scalar, true)) return 1;
		if (hyper_vector(ctx, work, out, -dt, c.nu_div, c.nu_vort, true)) return 1;
		if (tb200_filter_negative_tracers(ctx, out)) return 1;
		if (tb200_dss(ctx, out, all)) return 1;
	} else {
		TB_FAIL(ctx, "Invalid viscosity order");
	}
	return 0;
}

// Rayleigh damping after the hyperdiffusion (APPLY_RAYLEIGH_WITH_HYPERVIS,
// Defines.h:74; HorizontalDynamicsFEM.cpp:2720-2725)
extern "C" int tb200_h_step_after_subcycle(
	tb200_ctx * ctx, int in, int out, int work, double dt
) {
	TimingScope ts(ctx, "StepAfterSubCycle");
	if (h_step_after_subcycle_impl(ctx, in, out, work, dt)) return 1;
	if (!ctx->has_rayleigh) return 0;
	const DevLayout & lay = ctx->lay;
	const long long total = lay.nelem * (long long)lay.nrows_state * lay.nn;
	long long nb = (total + 255) / 256;
	if (nb > 148 * 16) nb = 148 * 16;
	auto kfn = k_rayleigh;
	TB_LAUNCH_FLAT(kfn, dim3((unsigned)nb), dim3(256), 0, ctx->stream,
		lay, (const double *)ctx->d_ray_node, (const double *)ctx->d_ray_redge,
		(const double *)ctx->d_refstate, ctx->inst[out], dt, ctx->cfg.cartesian_xz);
	TB_KERNEL_CHECK(ctx);
	return 0;
}

///////////////////////////////////////////////////////////////////////////////
// Connectivity

extern "C" int tb200_set_node_ids(tb200_ctx * ctx, int patch_index, const int64_t * ids) {
	PatchInfo * pi = find_patch(ctx, patch_index);
	if (pi == 0) TB_FAIL(ctx, "unknown patch");
	const size_t n = (size_t)pi->nea * ctx->lay.np * pi->neb * ctx->lay.np;
	pi->ids.assign(ids, ids + n);
	return 0;
}

extern "C" int tb200_set_seam_transforms(
	tb200_ctx * ctx, int patch_index, int n, const int * ia, const int * ib,
	const int * src_panel, const double * m
) {
	PatchInfo * pi = find_patch(ctx, patch_index);
	if (pi == 0) TB_FAIL(ctx, "unknown patch");
	pi->seams.clear();
	for (int q = 0; q < n; q++) {
		SeamEntry s;
		s.ia = ia[q];
		s.ib = ib[q];
		s.src_panel = src_panel[q];
		for (int r = 0; r < 4; r++) s.m[r] = m[4 * q + r];
		pi->seams.push_back(s);
	}
	return 0;
}

struct Member {
	int ppos;       // position of the patch in ctx->patches
	int ia, ib;
	long long addr; // local node address, -1 when remote
};

// Order of the members of an averaging group = association order of the average
// (pairs (0, 1) and (2, 3), then the pair of pair averages; tb200_dss.cuh).  It
// must not depend on how a panel is cut into patches - a run on 24 patches has
// to reproduce the bits of the run on 6 - and it has to pair the reference's way:
// GridCSGLL::ApplyDSS averages across alpha first, then across beta
// (GridCSGLL.cpp:435-781), so the two members of an alpha pair share the
// element-local column j.  Key: panel, then element-local j (the element ending
// at the node first: j = np-1), then element-local i likewise; patch index and
// position only break ties (slot order of the exchange lists).
static bool member_less(const tb200_ctx * ctx, const Member & x, const Member & y) {
	const PatchInfo & px = ctx->patches[x.ppos];
	const PatchInfo & py = ctx->patches[y.ppos];
	if (px.panel != py.panel) return px.panel < py.panel;
	const int np = ctx->lay.np;
	const int jx = (x.ib % np == np - 1) ? 0 : 1, jy = (y.ib % np == np - 1) ? 0 : 1;
	if (jx != jy) return jx < jy;
	const int ix = (x.ia % np == np - 1) ? 0 : 1, iy = (y.ia % np == np - 1) ? 0 : 1;
	if (ix != iy) return ix < iy;
	if (px.index != py.index) return px.index < py.index;
	if (x.ib != y.ib) return x.ib < y.ib;
	return x.ia < y.ia;
}

// Strips of the fused DSS (tb200_fast.cuh) and the averaging groups it leaves to
// the group kernels.  A strip is a run of beta-consecutive elements of one
// alpha-row of a patch; strips are handed to the persistent blocks round-robin,
// in (patch, alpha-row, chunk) order, so that the strip one alpha-row down - whose
// values the alpha-edge averaging reads - is walked at the same pace by another
// block.  Fused are the groups whose members are exactly
//   (e-1)(i,3), e(i,0)              i = 1, 2, e not first in its strip
//   (e-neb)(3,j), e(0,j)            j = 1, 2, e not in the first alpha-row
//   (e-neb-1)(3,3), (e-1)(0,3), (e-neb)(3,0), e(0,0)      both of the above
// with all members local and on one panel (flag bit 1).
static int build_fuse(tb200_ctx * ctx, const std::vector<int> & members, const std::vector<int> & flags) {
	const DevLayout & lay = ctx->lay;
	ctx->fuse_ready = false;
	if (lay.np != 4 || ctx->cfg.eqn_type != TB200_EQN_PRIMITIVE_NONHYDRO) return 0;
	const long long nelem = lay.nelem;
	const int nn = lay.nn;
	// strip length: about 8 strips per resident block, whole chunks of a row
	const long long nblocks = 2ll * ctx->sm_count;
	long long target = nelem / (8 * nblocks);
	target = std::max(4ll, std::min(64ll, target));
	{
		const char * e = getenv("TB200_STRIP");      // tests: any strip length
		if (e != 0 && atoi(e) > 0) target = atoi(e);
	}
	std::vector<int> first, len, nebs;
	std::vector<char> fa(nelem, 0), fb(nelem, 0);
	std::vector<int> enb(nelem, 0);
	for (size_t p = 0; p < ctx->patches.size(); p++) {
		const PatchInfo & pi = ctx->patches[p];
		if (pi.elem0 < 0) continue;
		const int nchunk = (int)((pi.neb + target - 1) / target);
		for (int a = 0; a < pi.nea; a++) {
			for (int c = 0; c < nchunk; c++) {
				const int b0 = (int)((long long)pi.neb * c / nchunk);
				const int b1 = (int)((long long)pi.neb * (c + 1) / nchunk);
				if (b1 <= b0) continue;
				const long long e0 = pi.elem0 + (long long)a * pi.neb + b0;
				first.push_back((int)e0);
				len.push_back(b1 - b0);
				nebs.push_back((a > 0) ? pi.neb : 0);
				for (int b = b0; b < b1; b++) {
					const long long e = pi.elem0 + (long long)a * pi.neb + b;
					fa[e] = (a > 0) ? 1 : 0;
					fb[e] = (b > b0) ? 1 : 0;
					enb[e] = pi.neb;
				}
			}
		}
	}
	// groups the kernels average themselves
	const int ngroups = (int)flags.size();
	std::vector<int> rem_members, rem_flags;
	long long nfused = 0;
	for (int gi = 0; gi < ngroups; gi++) {
		bool fused = false;
		if (flags[gi] & 2) {
			const int * m = &members[(size_t)gi * 4];
			const int cnt = (m[2] >= 0) ? 4 : 2;
			int top = 0;
			for (int q = 1; q < cnt; q++) if (m[q] / nn > m[top] / nn) top = q;
			const long long e = m[top] / nn;
			const int n = m[top] % nn, i = n / 4, j = n % 4;
			auto has = [&](long long ee, int ii, int jj) {
				const int addr = (int)(ee * nn + ii * 4 + jj);
				for (int q = 0; q < cnt; q++) if (m[q] == addr) return true;
				return false;
			};
			if (cnt == 2 && j == 0 && (i == 1 || i == 2)) {
				fused = fb[e] && has(e - 1, i, 3);
			} else if (cnt == 2 && i == 0 && (j == 1 || j == 2)) {
				fused = fa[e] && has(e - enb[e], 3, j);
			} else if (cnt == 4 && i == 0 && j == 0) {
				fused = fa[e] && fb[e] && has(e - 1, 0, 3) && has(e - enb[e], 3, 0)
					&& has(e - enb[e] - 1, 3, 3);
			}
		}
		if (fused) {
			nfused++;
		} else {
			for (int q = 0; q < 4; q++) rem_members.push_back(members[(size_t)gi * 4 + q]);
			rem_flags.push_back(flags[gi]);
		}
	}
	ctx->nstrips = (int)first.size();
	ctx->nrem = (int)rem_flags.size();
	if (ctx->nstrips == 0 || nfused == 0) return 0;
	if (dupload(ctx, &ctx->d_strip_first, first)) return 1;
	if (dupload(ctx, &ctx->d_strip_len, len)) return 1;
	if (dupload(ctx, &ctx->d_strip_neb, nebs)) return 1;
	if (dupload(ctx, &ctx->d_rem_members, rem_members)) return 1;
	if (dupload(ctx, &ctx->d_rem_flags, rem_flags)) return 1;
	if (ctx->d_done == 0) {
		if (dalloc(ctx, &ctx->d_done, (size_t)nelem)) return 1;
		TB_CHECK(ctx, cudaMemset(ctx->d_done, 0, (size_t)nelem * sizeof(unsigned)));
	}
	ctx->fuse_epoch = 0;
	ctx->fuse_ready = true;
	return 0;
}

extern "C" int tb200_build_connectivity(tb200_ctx * ctx) {
	if (!ctx->committed) TB_FAIL(ctx, "commit the layout first");
	const int np = ctx->lay.np, nn = ctx->lay.nn;

	// collect the boundary nodes of every element of every patch by id
	std::unordered_map<int64_t, std::vector<Member> > byid;
	for (size_t p = 0; p < ctx->patches.size(); p++) {
		const PatchInfo & pi = ctx->patches[p];
		const int wbi = pi.neb * np;
		if (pi.ids.size() != (size_t)pi.nea * np * wbi) {
			TB_FAIL(ctx, "node ids missing for a patch");
		}
		for (int a = 0; a < pi.nea; a++)
		for (int b = 0; b < pi.neb; b++)
		for (int i = 0; i < np; i++)
		for (int j = 0; j < np; j++) {
			if (i != 0 && i != np - 1 && j != 0 && j != np - 1) continue;
			Member m;
			m.ppos = (int)p;
			m.ia = a * np + i;
			m.ib = b * np + j;
			m.addr = (pi.elem0 >= 0)
				? (pi.elem0 + (long long)a * pi.neb + b) * nn + i * np + j : -1;
			byid[pi.ids[(size_t)m.ia * wbi + m.ib]].push_back(m);
		}
	}

	const int nlocal = (int)(ctx->lay.nelem * nn);
	// exchange lists: for the ordered pair (src rank -> dst rank) the nodes of
	// src that share an id with a node of dst, in (patch, ib, ia) order.
	// send_lists[d] = my nodes sent to d; recv_index[(s, patch, ia, ib)] = slot.
	std::vector<std::vector<Member> > send_lists(ctx->nranks), recv_lists(ctx->nranks);

	struct Group { std::vector<Member> mem; bool seam; };
	std::vector<Group> groups;

	for (std::unordered_map<int64_t, std::vector<Member> >::iterator it = byid.begin();
	     it != byid.end(); ++it
	) {
		std::vector<Member> & mem = it->second;
		if (mem.size() < 2) continue;
		if (mem.size() > 4) TB_FAIL(ctx, "node shared by more than four elements");
		bool anylocal = false, anyremote = false, seam = false;
		for (size_t q = 0; q < mem.size(); q++) {
			if (mem[q].addr >= 0) anylocal = true; else anyremote = true;
			if (ctx->patches[mem[q].ppos].panel != ctx->patches[mem[0].ppos].panel) seam = true;
		}
		if (!anylocal) continue;
		std::sort(mem.begin(), mem.end(),
			[ctx](const Member & x, const Member & y) { return member_less(ctx, x, y); });
		if (anyremote) {
			for (size_t q = 0; q < mem.size(); q++) {
				const int owner = ctx->patches[mem[q].ppos].owner;
				if (owner == ctx->rank) {
					// this node goes to every other rank present in the group
					std::vector<int> seen;
					for (size_t r = 0; r < mem.size(); r++) {
						const int o2 = ctx->patches[mem[r].ppos].owner;
						if (o2 != ctx->rank && std::find(seen.begin(), seen.end(), o2) == seen.end()) {
							seen.push_back(o2);
							send_lists[o2].push_back(mem[q]);
						}
					}
				} else {
					recv_lists[owner].push_back(mem[q]);
				}
			}
		}
		Group gr;
		gr.mem = mem;
		gr.seam = seam;
		groups.push_back(gr);
	}

	// deterministic slot order on both sides
	std::vector<int> send_nodes;
	ctx->send_rank.clear();
	ctx->send_j.clear();
	ctx->send_count.assign(ctx->nranks, 0);
	ctx->recv_count.assign(ctx->nranks, 0);
	std::map<std::pair<int, std::pair<int, int> >, int> recv_slot;  // (patch index,(ia,ib)) -> slot
	int slot = 0;
	for (int r = 0; r < ctx->nranks; r++) {
		std::vector<Member> & sl = send_lists[r];
		std::sort(sl.begin(), sl.end(),
			[ctx](const Member & x, const Member & y) { return member_less(ctx, x, y); });
		// a node can be listed once per group only, groups are disjoint: no duplicates
		for (size_t q = 0; q < sl.size(); q++) {
			send_nodes.push_back((int)sl[q].addr);
			ctx->send_rank.push_back(r);
			ctx->send_j.push_back((int)q);
		}
		ctx->send_count[r] = (int64_t)sl.size();
		std::vector<Member> & rl = recv_lists[r];
		std::sort(rl.begin(), rl.end(),
			[ctx](const Member & x, const Member & y) { return member_less(ctx, x, y); });
		for (size_t q = 0; q < rl.size(); q++) {
			recv_slot[std::make_pair(ctx->patches[rl[q].ppos].index,
				std::make_pair(rl[q].ia, rl[q].ib))] = slot++;
		}
		ctx->recv_count[r] = (int64_t)rl.size();
	}
	ctx->nsend_total = (int)send_nodes.size();
	if (ctx->nranks > 1) {
		// elements that own a node of the send list, and the rest
		std::vector<char> feeds((size_t)ctx->lay.nelem, 0);
		for (size_t q = 0; q < send_nodes.size(); q++) feeds[send_nodes[q] / nn] = 1;
		std::vector<int> lb, li;
		for (long long e = 0; e < ctx->lay.nelem; e++) {
			(feeds[e] ? lb : li).push_back((int)e);
		}
		ctx->n_bnd = (int)lb.size();
		ctx->n_int = (int)li.size();
		if (dupload(ctx, &ctx->d_elist_bnd, lb)) return 1;
		if (dupload(ctx, &ctx->d_elist_int, li)) return 1;
	}
	ctx->nrecv_total = slot;
	if (dupload(ctx, &ctx->d_send_nodes, send_nodes)) return 1;

	// order groups by their first local member for memory locality
	std::sort(groups.begin(), groups.end(), [](const Group & x, const Group & y) {
		long long ax = -1, ay = -1;
		for (size_t q = 0; q < x.mem.size(); q++) if (x.mem[q].addr >= 0) { ax = x.mem[q].addr; break; }
		for (size_t q = 0; q < y.mem.size(); q++) if (y.mem[q].addr >= 0) { ay = y.mem[q].addr; break; }
		return ax < ay;
	});

	std::vector<int> members(groups.size() * 4, -1), flags(groups.size(), 0);
	std::vector<int> seam_group;
	std::vector<double> seam_mats;
	for (size_t gi = 0; gi < groups.size(); gi++) {
		const Group & gr = groups[gi];
		for (size_t q = 0; q < gr.mem.size(); q++) {
			const Member & m = gr.mem[q];
			if (m.addr >= 0) {
				members[gi * 4 + q] = (int)m.addr;
			} else {
				members[gi * 4 + q] = nlocal + recv_slot[std::make_pair(
					ctx->patches[m.ppos].index, std::make_pair(m.ia, m.ib))];
			}
		}
		// ranks whose nodes this group reads from the receive buffer
		int rankmask = 0;
		for (size_t q = 0; q < gr.mem.size(); q++) {
			if (gr.mem[q].addr < 0) rankmask |= 1 << ctx->patches[gr.mem[q].ppos].owner;
		}
		if (!gr.seam) {
			bool local = (gr.mem.size() == 2 || gr.mem.size() == 4);
			for (size_t q = 0; q < gr.mem.size(); q++) local = local && gr.mem[q].addr >= 0;
			flags[gi] = (local ? 2 : 0) | (rankmask << 8);
			continue;
		}
		flags[gi] = 1 | (rankmask << 8);
		seam_group.push_back((int)gi);
		const size_t base = seam_mats.size();
		seam_mats.resize(base + 64, 0.0);
		for (size_t t = 0; t < gr.mem.size(); t++) {
			const PatchInfo & pt = ctx->patches[gr.mem[t].ppos];
			for (size_t s = 0; s < gr.mem.size(); s++) {
				const PatchInfo & ps = ctx->patches[gr.mem[s].ppos];
				double * M = &seam_mats[base + (t * 4 + s) * 4];
				if (ps.panel == pt.panel) {
					M[0] = 1.0; M[1] = 0.0; M[2] = 0.0; M[3] = 1.0;
					continue;
				}
				if (gr.mem[t].addr < 0) continue;   // remote targets are not written
				bool found = false;
				for (size_t q = 0; q < pt.seams.size(); q++) {
					const SeamEntry & se = pt.seams[q];
					if (se.ia == gr.mem[t].ia && se.ib == gr.mem[t].ib && se.src_panel == ps.panel) {
						for (int r = 0; r < 4; r++) M[r] = se.m[r];
						found = true;
						break;
					}
				}
				if (!found) TB_FAIL(ctx, "missing seam transform for a panel-boundary node");
			}
		}
	}
	ctx->ngroups = (int)groups.size();
	ctx->nseam = (int)seam_group.size();
	if (build_fuse(ctx, members, flags)) return 1;
	if (dupload(ctx, &ctx->d_members, members)) return 1;
	if (dupload(ctx, &ctx->d_flags, flags)) return 1;
	if (dupload(ctx, &ctx->d_seam_group, seam_group)) return 1;
	if (dupload(ctx, &ctx->d_seam_mats, seam_mats)) return 1;
	ctx->connectivity_built = true;
	return 0;
}

extern "C" int tb200_exchange_counts(
	tb200_ctx * ctx, int64_t * send_nodes, int64_t * recv_nodes
) {
	for (int r = 0; r < ctx->nranks; r++) {
		send_nodes[r] = ctx->send_count.size() ? ctx->send_count[r] : 0;
		recv_nodes[r] = ctx->recv_count.size() ? ctx->recv_count[r] : 0;
	}
	return 0;
}

///////////////////////////////////////////////////////////////////////////////
// Band solve test hook

extern "C" int tb200_test_band_solve(
	tb200_ctx * ctx, int ncols, int n, int kl, int ku, const double * ab, double * b
) {
	const int ldab = 2 * kl + ku + 1;
	if (ctx->d_info == 0) {
		if (dalloc(ctx, &ctx->d_info, 4)) return 1;
		TB_CHECK(ctx, cudaMemset(ctx->d_info, 0, 4 * sizeof(int)));
	}
	// host [col][n][ldab] -> device [(j*ldab + r)][col]
	std::vector<double> hab((size_t)ncols * n * ldab), hb((size_t)ncols * n);
	for (int c = 0; c < ncols; c++) {
		for (int q = 0; q < n * ldab; q++) hab[(size_t)q * ncols + c] = ab[(size_t)c * n * ldab + q];
		for (int q = 0; q < n; q++) hb[(size_t)q * ncols + c] = b[(size_t)c * n + q];
	}
	double * dab = 0;
	double * db = 0;
	if (dupload(ctx, &dab, hab)) return 1;
	if (dupload(ctx, &db, hb)) return 1;
	auto kfn = k_band_solve;
	TB_LAUNCH_FLAT(kfn, dim3((ncols + 63) / 64), dim3(64), 0, ctx->stream,
		ncols, n, kl, ku, dab, db, ctx->d_info);
	TB_KERNEL_CHECK(ctx);
	TB_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
	TB_CHECK(ctx, cudaMemcpy(hb.data(), db, hb.size() * sizeof(double), cudaMemcpyDeviceToHost));
	for (int c = 0; c < ncols; c++) {
		for (int q = 0; q < n; q++) b[(size_t)c * n + q] = hb[(size_t)q * ncols + c];
	}
	return check_column_info(ctx);
}

///////////////////////////////////////////////////////////////////////////////
// Conservation diagnostics (kernels: tb200_diag.cuh)


static int diagnostic(tb200_ctx * ctx, int inst, int what, const double * vort, double * value) {
	if (ctx->d_area_node == 0) TB_FAIL(ctx, "element areas not uploaded");
	if (inst < 0 || inst >= (int)ctx->inst.size()) TB_FAIL(ctx, "invalid state instance");
	const bool sw = (ctx->cfg.eqn_type == TB200_EQN_SHALLOW_WATER);
	if (!sw) {
		if (ctx->ops.op[TB200_OP_INTERP_E2N].coeff == 0 || ctx->ops.op[TB200_OP_INTERP_N2E].coeff == 0) {
			TB_FAIL(ctx, "vertical column operators not set");
		}
		if (!ctx->geom.analytic && !ctx->geometry3d_uploaded) {
			TB_FAIL(ctx, "diagnostics need the 3-D metric arrays or the terrain metric");
		}
		if (ctx->d_reta_n == 0) TB_FAIL(ctx, "vertical coordinate not set (tb200_set_vertical_coordinate)");
	}
	DiagArgs da;
	da.g = ctx->cfg.g;
	da.gamma = ctx->cfg.cp / (ctx->cfg.cp - ctx->cfg.R);              // PhysicalConstants.h:368
	da.pscale = ctx->cfg.p0 * pow(ctx->cfg.R / ctx->cfg.p0, da.gamma);  // PhysicalConstants.h:375
	da.shallow = sw ? 1 : 0;
	da.what = what;
	da.vort = vort;
	DevGeom g = ctx->geom;
	g.reta_n = ctx->d_reta_n; g.reta_e = ctx->d_reta_e; g.ztop = ctx->cfg.ztop;
	TB_CHECK(ctx, cudaMemsetAsync(ctx->d_sums, 0, 64 * sizeof(double), ctx->stream));
	auto kfn = k_diagnostic;
	TB_LAUNCH(kfn, dim3(2 * 148), dim3(128), 0, ctx->stream,
		ctx->lay, g, ctx->ops, da, (const double *)ctx->inst[inst],
		(const double *)ctx->d_area_node, (const double *)ctx->d_area_redge, ctx->d_sums);
	TB_KERNEL_CHECK(ctx);
	TB_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
	TB_CHECK(ctx, cudaMemcpy(value, ctx->d_sums, sizeof(double), cudaMemcpyDeviceToHost));
	return 0;
}

extern "C" int tb200_total_energy(tb200_ctx * ctx, int inst, double * energy) {
	return diagnostic(ctx, inst, 0, 0, energy);
}

extern "C" int tb200_total_vertical_momentum(tb200_ctx * ctx, int inst, double * momentum) {
	if (ctx->cfg.eqn_type == TB200_EQN_SHALLOW_WATER) {
		// GridPatch.cpp:1262-1265
		TB_FAIL(ctx, "ComputeTotalVerticalMomentum() Not implemented for ShallowWaterEquations");
	}
	return diagnostic(ctx, inst, 2, 0, momentum);
}

extern "C" int tb200_total_potential_enstrophy(
	tb200_ctx * ctx, int inst, int work, double * enstrophy
) {
	if (ctx->cfg.eqn_type != TB200_EQN_SHALLOW_WATER) {
		return diagnostic(ctx, inst, 1, 0, enstrophy);
	}
	// Grid::ComputeVorticityDivergence + DSS of the vorticity (GridGLL.cpp:587-602)
	const int ni = (int)ctx->inst.size();
	if (work < 0 || work >= ni || work == inst) {
		TB_FAIL(ctx, "potential enstrophy of a shallow-water state needs a scratch instance");
	}
	const DevLayout & lay = ctx->lay;
	const long long nitems = lay.nelem * lay.nlev;
	TB_NP_SWITCH(lay.np,
		auto kfn = k_sw_vorticity<NPV>;
		TB_LAUNCH(kfn, dim3((unsigned)((nitems + 7) / 8)), dim3(NPV * NPV * 8), 0, ctx->stream,
			lay, ctx->geom, ctx->tables, (const double *)ctx->inst[inst], ctx->inst[work]);)
	TB_KERNEL_CHECK(ctx);
	if (tb_dss_scalar_rows(ctx, work, lay.rowoff[0], lay.rowoff[0] + lay.nlev)) return 1;
	return diagnostic(ctx, inst, 1, ctx->inst[work], enstrophy);
}
