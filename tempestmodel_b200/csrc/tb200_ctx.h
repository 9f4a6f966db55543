// Host-side context of the library (private).
#ifndef TB200_CTX_H
#define TB200_CTX_H

#include <string>
#include <vector>
#include <map>
#include <cstdint>

#include "tb200_platform.h"
#include "tb200_device.h"
#include "tb200_tma.cuh"
#include "../../include/tempest_b200.h"

struct SeamEntry {
	int ia, ib, src_panel;
	double m[4];
};

struct PatchInfo {
	int index, panel, nea, neb, halo, owner;
	double da, db;
	long long elem0;           // first local element, -1 when owned elsewhere
	std::vector<int64_t> ids;  // [nea*np][neb*np] global node ids
	std::vector<SeamEntry> seams;
	bool has_terrain;
	bool has_lonlat;      // tb200_evaluate_geometry_cs has run for this patch
	bool has_held_suarez; // tb200_upload_held_suarez has run for this patch
	PatchInfo() : has_terrain(false), has_lonlat(false), has_held_suarez(false) {}
};

struct HostOp {
	int nout, nin, width;
	std::vector<double> coeff;   // [nout][width]
	std::vector<int> begin, end;
	double * d_coeff;
	int * d_begin;
	int * d_end;
	HostOp() : nout(0), nin(0), width(0), d_coeff(0), d_begin(0), d_end(0) {}
};

struct tb200_ctx {
	tb200_config cfg;
	std::string err;
	cudaStream_t stream;
	bool committed;
	bool connectivity_built;

	std::vector<PatchInfo> patches;
	std::map<int, int> patch_pos;     // patch index -> position in patches

	DevLayout lay;
	DevTables tables;
	DevOps ops;
	DevGeom geom;
	DevPhys phys;
	HostOp hops[TB_NOPS];

	std::vector<double *> inst;       // device state instances
	std::vector<TbMap> tmaps;         // their tensor maps (bulk tensor copies of the pipelined kernels)
	std::vector<void *> allocs;       // everything to cudaFree

	// staging for host <-> device layout conversion
	double * d_stage;
	size_t stage_doubles;
	int * d_rowmap;                   // [64] device row of host component: node 0.., interfaces 8.., tracers 16..
	std::vector<int> rowmap_h;
	// pipelined state transfers: copy stream, two staging buffers and their events
	cudaStream_t copy_stream;
	double * stage_buf[2];
	cudaEvent_t ev_stage_free[2], ev_stage_full[2], ev_compute;
	unsigned long long xfer_jobs;

	// mutable geometry arrays (device), same pointers as in geom
	double * d_inv_da; double * d_inv_db; double * d_nu_scale;
	double * g2d[7];
	double * g3n[13];
	double * g3e[13];
	double * d_area_node;             // [e][L][NN] element area (checksum)
	double * d_area_redge;            // [e][L+1][NN]
	double * d_sums;
	double * d_tx; double * d_ty; double * d_tda; double * d_tdb;
	double * d_reta_n; double * d_reta_e;

	// averaging groups
	int ngroups;
	int * d_members; int * d_flags;
	int nseam;
	int * d_seam_group; double * d_seam_mats;
	// multi-rank exchange
	int rank, nranks;
	tb200_exchange_fn exch_fn;
	void * exch_user;
	int nsend_total, nrecv_total;
	std::vector<int64_t> send_count, recv_count;   // nodes per peer
	int * d_send_nodes;
	double * d_sendbuf; double * d_recvbuf;
	// peer-memory exchange (tb200_peer_export / tb200_peer_attach)
	std::vector<int> send_rank, send_j;     // per send slot: destination, index in its block
	int * d_send_rank; int * d_send_slot;
	void * peer_area;                       // [64 flags][2][nrecv_total * peer_rows]
	size_t peer_rows;
	std::vector<void *> peer_base;          // IPC mappings of the peers' areas
	std::vector<int64_t> peer_recv_total;
	bool peer_ready;
	unsigned long long peer_seq;
	unsigned * d_peer_ticket;               // blocks of the pack kernel that have finished
	size_t buf_rows;                  // rows per slot the buffers are sized for

	// DSS fused into the pipelined kernels (tb200_fast.cuh: FuseArgs)
	bool fuse_ready;                  // strips and the list of unfused groups are built
	int nstrips;
	int * d_strip_first; int * d_strip_len; int * d_strip_neb;
	unsigned * d_done;                // [nelem] launch epoch per element
	unsigned fuse_epoch;
	int nrem;                         // averaging groups the fused kernels leave raw
	int * d_rem_members; int * d_rem_flags;
	bool fuse_want;                   // the caller follows the launch with a DSS of `out`
	bool fuse_done;                   // ... and the launch did the fused part of it

	// implicit column solve
	int ncols;
	int * d_col_node; int * d_col_dups;
	// Strang tail with tracers: instance that receives the tracers from before the
	// column update (0: tracer_inc holds them already) / the increment new - old
	double * tracer_keep; double * tracer_inc;
	// Held-Suarez forcing (tb200_physics.cuh): latitude and the surface product per column
	double * d_hs_lat; double * d_hs_sp;
	double * d_lon;       // longitude per column (device-side set-up)
	double * d_precip;    // accumulated precipitation per column (Kessler)
	double * d_outfield;  // vorticity, divergence, temperature on levels [e][3 L][NN]
	double * d_ws; int ws_cols; int * d_info;
	double * d_ray_node; double * d_ray_redge; double * d_refstate;   // Rayleigh friction
	bool has_rayleigh;
	// uniform diffusion (Grid::HasUniformDiffusion, Grid.cpp:399-415): scalar and
	// vector coefficients; acts on the state minus the reference state (d_refstate)
	double uniform_s; double uniform_v;
	double * d_stale_uv;   // [2][nlev]: see uniform_diffusion_vertical_uv (tb200_api.cu)
	double * column_inc;              // set while tb200_copy_v_step_implicit_diff runs
	double * d_wold;                  // w before the implicit solve (tracer update)
	int offd;
	int fe_nodes;          // levels per vertical element: the order (FE), 1 (FV)
	int finite_volume;     // Grid::VerticalDiscretization_FiniteVolume
	int mass_flux_levels;  // --vmassfluxlevels

	// exchange / compute overlap on several ranks
	cudaStream_t stream2;
	cudaEvent_t ev_fork, ev_join;
	int * d_elist_bnd; int n_bnd;     // elements owning nodes of the send list
	int * d_elist_int; int n_int;     // all other elements
	bool want_split, split_pending;

	// FunctionTimer hooks (tb200_set_timing_hooks); depth: only the outermost
	// entry point of a group reports
	tb200_timing_fn timing_begin, timing_end;
	void * timing_user;

	int64_t launches;
	// Every operation that writes device state bumps `writes`: kernel launches
	// (TB_KERNEL_CHECK and the counted persistent launches) and the
	// cudaMemcpyAsync of tb200_copy.  uvzero_inst is the instance whose u, v
	// rows are known to be zero (the increment written by the Strang tail) as
	// long as `writes` still equals uvzero_writes, i.e. nothing has written
	// anything since; -1 = none
	int64_t writes;
	int uvzero_inst;
	int64_t uvzero_writes;
	bool carry_full;                  // TB200_CARRY_FULL at tb200_create (tests)
	// pinned host mirror of d_info, refreshed by an asynchronous copy at the end
	// of every tb200_step: tb200_step refuses to start once it shows a failure
	int * h_info;
	int sm_count;

	// fast path (tb200_fast.cuh): 0 = not examined, 1 = enabled, -1 = unavailable
	int fast_state;
	std::string fast_reason;          // why the fast path is unavailable
	double fast_metric_error;         // deviation from the uploaded 3-D metric
	double * d_colc;                  // [e][TBF_NC][NN] column constants
	double * d_lev;                   // [L+1][TBF_LW] operator windows
	std::vector<double> reta_n_h, reta_e_h;
	bool geometry3d_uploaded;

	tb200_ctx() :
		stream(0), committed(false), connectivity_built(false),
		stream2(0), ev_fork(0), ev_join(0), d_elist_bnd(0), n_bnd(0), d_elist_int(0), n_int(0),
		want_split(false), split_pending(false),
		d_stage(0), stage_doubles(0), d_rowmap(0), copy_stream(0), xfer_jobs(0),
		d_inv_da(0), d_inv_db(0), d_nu_scale(0),
		d_area_node(0), d_area_redge(0), d_sums(0),
		d_tx(0), d_ty(0), d_tda(0), d_tdb(0), d_reta_n(0), d_reta_e(0),
		ngroups(0), d_members(0), d_flags(0), nseam(0), d_seam_group(0),
		d_seam_mats(0), rank(0), nranks(1), exch_fn(0), exch_user(0),
		nsend_total(0), nrecv_total(0), d_send_nodes(0), d_sendbuf(0),
		d_recvbuf(0), d_send_rank(0), d_send_slot(0), peer_area(0), peer_rows(0),
		peer_ready(false), peer_seq(0), d_peer_ticket(0), buf_rows(0), ncols(0), d_col_node(0), d_col_dups(0), tracer_keep(0), tracer_inc(0), d_hs_lat(0), d_hs_sp(0), d_lon(0), d_precip(0),
		d_ws(0), ws_cols(0), d_info(0), d_ray_node(0), d_ray_redge(0), d_refstate(0), has_rayleigh(false), uniform_s(0.0), uniform_v(0.0), d_stale_uv(0),
		column_inc(0), d_wold(0), offd(4), launches(0), writes(0), uvzero_inst(-1), uvzero_writes(0),
		timing_begin(0), timing_end(0), timing_user(0),
		carry_full(false), h_info(0),
		fuse_ready(false), nstrips(0), d_strip_first(0), d_strip_len(0), d_strip_neb(0),
		d_done(0), fuse_epoch(0), nrem(0), d_rem_members(0), d_rem_flags(0),
		fuse_want(false), fuse_done(false),
		fast_state(0), fast_metric_error(0.0), d_colc(0), d_lev(0),
		geometry3d_uploaded(false)
	{
		stage_buf[0] = stage_buf[1] = 0;
		for (int i = 0; i < 2; i++) { ev_stage_free[i] = 0; ev_stage_full[i] = 0; }
		ev_compute = 0;
		for (int i = 0; i < 7; i++) g2d[i] = 0;
		d_outfield = 0;
		for (int i = 0; i < 13; i++) { g3n[i] = 0; g3e[i] = 0; }
	}
};

#endif
