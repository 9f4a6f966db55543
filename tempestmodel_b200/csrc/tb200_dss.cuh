// Direct stiffness summation and halo traffic.
//
// The reference copies one-node strips into patch halos (Grid::Exchange,
// Grid.cpp:627-685; ExchangeBuffer::Pack/Unpack, Connectivity.cpp:47-744),
// re-bases halo velocities at panel seams
// (GridPatchCSGLL::TransformHaloVelocities, GridPatchCSGLL.cpp:1783-1924) and
// then averages duplicates pairwise in alpha, then beta, with a 3-way average
// at cube corners (GridCSGLL::ApplyDSS, GridCSGLL.cpp:435-781).  The net
// effect on every physical node is the mean of its duplicates expressed in
// the owner's basis.  The device has no halo: each physical node shared by
// 2..4 element-local nodes is one *averaging group*; a thread gathers the
// members, averages them in the reference's association order
// (alpha pairs first, then beta) and scatters the result to the local members.
// Members owned by other ranks are read from the receive buffer that the
// exchange filled from the peers' packed send buffers.
#ifndef TB200_DSS_CUH
#define TB200_DSS_CUH

#include "tb200_platform.h"
#include "tb200_device.h"

struct DssArgs {
	const int * members;      // [ngroups][4], -1 = unused; >= nlocal: remote slot
	const int * flags;        // [ngroups] bit0: group spans a panel seam; bit1: 2 or 4 members, all local, no seam;
	                          // bits 8..23: ranks that own remote members of the group
	// peer-memory exchange: the kernel itself waits for the source ranks' flags
	// (0: the exchange is complete when the kernel starts)
	const unsigned long long * peer_flags;   // [rank] sequence number of the last exchange that rank delivered
	unsigned long long peer_seq;
	unsigned long long peer_timeout_ns;
	int * info;               // [1]: rank + 1 of a peer that did not deliver in time
	int ngroups;
	int nlocal;               // number of local element nodes (nelem * NN)
	const double * recv;      // [slot][nsel]
	int row0, row1;           // rows averaged
	int uv_row0, uv_row1;     // rows of (U,V) skipped for seam groups (vector path)
	int nsel;                 // rows per remote slot in this exchange
	int sel_row0;             // first row carried by the exchange
	int rows_fastest;         // 1: blockIdx.x walks the row chunks, blockIdx.y the groups
};

// offset of local node m = e * nn + n inside row 0 of its element
__device__ __forceinline__ size_t tb_dss_base(const DevLayout & lay, int m) {
	if (lay.nn == 16) {
		return (size_t)(m >> 4) * lay.nrows * 16 + (m & 15);
	}
	const int e = m / lay.nn;
	const int n = m % lay.nn;
	return (size_t)e * lay.nrows * lay.nn + n;
}

__device__ __forceinline__ double tb_dss_load(
	const DevLayout & lay, const DssArgs & a, const double * data, int m, int r
) {
	if (m < a.nlocal) {
		return data[tb_dss_base(lay, m) + (size_t)r * lay.nn];
	}
	return a.recv[(size_t)(m - a.nlocal) * a.nsel + (r - a.sel_row0)];
}

// member reference resolved once per thread: local nodes -> pointer to row 0,
// remote nodes -> pointer into the receive buffer
struct DssRef {
	const double * p;
	double * w;        // writable alias for local members, null otherwise
	int stride;        // doubles per row step
};

__device__ __forceinline__ DssRef tb_dss_ref(
	const DevLayout & lay, const DssArgs & a, double * data, int m
) {
	DssRef r;
	if (m < 0) {
		r.p = 0; r.w = 0; r.stride = 0;
	} else if (m < a.nlocal) {
		r.w = data + tb_dss_base(lay, m);
		r.p = r.w;
		r.stride = lay.nn;
	} else {
		r.p = a.recv + (size_t)(m - a.nlocal) * a.nsel - a.sel_row0;
		r.w = 0;
		r.stride = 1;
	}
	return r;
}

// value of a member at row r: local memory, or the receive buffer - filled by
// another GPU's stores while this kernel may already be running, hence read
// from L2 (ld.cg), never through L1
__device__ __forceinline__ double tb_dss_get(const DssRef & m, int r) {
#ifndef TB200_EMU
	if (m.w == 0) return __ldcg(m.p + (size_t)r * m.stride);
#endif
	return m.p[(size_t)r * m.stride];
}

// Wait until every rank in `mask` has delivered exchange a.peer_seq.  On a
// time-out the failure is recorded and false returned: the caller poisons the
// group (NaN) instead of averaging stale halo data.
__device__ __forceinline__ bool tb_dss_wait_peers(const DssArgs & a, unsigned mask) {
#ifndef TB200_EMU
	if (*(volatile int *)(a.info + 1) != 0) return false;     // a peer is gone: do not wait again
	unsigned long long t0;
	asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
	while (mask != 0) {
		const int r = __ffs(mask) - 1;
		unsigned long long f;
		asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(f) : "l"(a.peer_flags + r) : "memory");
		if (f >= a.peer_seq) {
			mask &= mask - 1;
			continue;
		}
		unsigned long long t1;
		asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
		if (t1 - t0 > a.peer_timeout_ns) {
			atomicMax(a.info + 1, r + 1);
			return false;
		}
	}
#endif
	return true;
}

// block coordinates: which batch of groups, which chunk of rows.  With
// rows_fastest consecutive blocks work on the successive row chunks of the same
// groups, i.e. walk an element's contiguous block of rows together.
__device__ __forceinline__ int tb_dss_group_block(const DssArgs & a) {
	return a.rows_fastest ? blockIdx.y : blockIdx.x;
}
__device__ __forceinline__ int tb_dss_row_block(const DssArgs & a) {
	return a.rows_fastest ? blockIdx.x : blockIdx.y;
}
__device__ __forceinline__ int tb_dss_row_blocks(const DssArgs & a) {
	return a.rows_fastest ? gridDim.x : gridDim.y;
}

__device__ __forceinline__ void tb_dss_generic(
	const DevLayout & lay, const DssArgs & a, int gidx, double * data
) {
	const int m2 = a.members[4 * gidx + 2];
	const int m3 = a.members[4 * gidx + 3];
	bool delivered = true;
	{
		const unsigned mask = ((unsigned)a.flags[gidx] >> 8) & 0xffffu;
		if (mask != 0 && a.peer_flags != 0) delivered = tb_dss_wait_peers(a, mask);
	}
	const DssRef r0 = tb_dss_ref(lay, a, data, a.members[4 * gidx + 0]);
	const DssRef r1 = tb_dss_ref(lay, a, data, a.members[4 * gidx + 1]);
	const DssRef r2 = tb_dss_ref(lay, a, data, m2);
	const DssRef r3 = tb_dss_ref(lay, a, data, m3);
	const bool seam = (a.flags[gidx] & 1) != 0;
	// blockIdx.y owns a contiguous range of rows: the rows of an element are
	// adjacent in memory, so a block walks whole DRAM pages
	const int rows_per = (a.row1 - a.row0 + tb_dss_row_blocks(a) - 1) / tb_dss_row_blocks(a);
	const int rbeg = a.row0 + tb_dss_row_block(a) * rows_per;
	const int rend = (rbeg + rows_per < a.row1) ? (rbeg + rows_per) : a.row1;
#pragma unroll 4
	for (int r = rbeg; r < rend; r++) {
		if (seam && r >= a.uv_row0 && r < a.uv_row1) continue;
		const double v0 = tb_dss_get(r0, r);
		const double v1 = tb_dss_get(r1, r);
		double avg;
		if (m2 < 0) {
			avg = 0.5 * (v0 + v1);
		} else if (m3 < 0) {
			const double v2 = tb_dss_get(r2, r);
			avg = (1.0 / 3.0) * (v0 + v1 + v2);
		} else {
			const double v2 = tb_dss_get(r2, r);
			const double v3 = tb_dss_get(r3, r);
			avg = 0.5 * (0.5 * (v0 + v1) + 0.5 * (v2 + v3));
		}
		// a peer never delivered: poison the node rather than average stale data
		// (tb200_check_errors / the next tb200_step report the rank)
		if (!delivered) avg = nan("");
		if (r0.w != 0) r0.w[(size_t)r * r0.stride] = avg;
		if (r1.w != 0) r1.w[(size_t)r * r1.stride] = avg;
		if (r2.w != 0) r2.w[(size_t)r * r2.stride] = avg;
		if (r3.w != 0) r3.w[(size_t)r * r3.stride] = avg;
	}
}

__global__ void k_dss_scalar(DevLayout lay, DssArgs a, double * data) {
	const int gidx = tb_dss_group_block(a) * blockDim.x + threadIdx.x;
	if (gidx >= a.ngroups) return;
	tb_dss_generic(lay, a, gidx, data);
}

// Groups whose 2 or 4 members are all local and off the panel seams (flag
// bit 1; every group but the seam and rank-boundary ones): the loads of B rows
// are issued before the first store (the members alias as far as the compiler
// can tell, so the row-at-a-time loop above serialises on every store) and the
// member pointers just step by one row.  Groups stay ordered by the address of
// their first member, so every row of an element is touched within one wave
// of blocks and comes from DRAM once.
template <int B>
__device__ __forceinline__ void tb_dss_local(
	const DevLayout & lay, const DssArgs & a, int gidx, double * data
) {
	const int rows_per = (a.row1 - a.row0 + tb_dss_row_blocks(a) - 1) / tb_dss_row_blocks(a);
	const int rbeg = a.row0 + tb_dss_row_block(a) * rows_per;
	const int rend = (rbeg + rows_per < a.row1) ? (rbeg + rows_per) : a.row1;
	const int nn = lay.nn;
	const int4 m = ((const int4 *)a.members)[gidx];
	const bool four = m.z >= 0;
	double * p0 = data + tb_dss_base(lay, m.x) + (size_t)rbeg * nn;
	double * p1 = data + tb_dss_base(lay, m.y) + (size_t)rbeg * nn;
	double * p2 = four ? data + tb_dss_base(lay, m.z) + (size_t)rbeg * nn : p0;
	double * p3 = four ? data + tb_dss_base(lay, m.w) + (size_t)rbeg * nn : p1;
	int r = rbeg;
	for (; r + B <= rend; r += B) {
		double v0[B], v1[B], v2[B], v3[B];
#pragma unroll
		for (int b = 0; b < B; b++) {
			v0[b] = p0[b * nn];
			v1[b] = p1[b * nn];
			if (four) {
				v2[b] = p2[b * nn];
				v3[b] = p3[b * nn];
			}
		}
#pragma unroll
		for (int b = 0; b < B; b++) {
			double avg = 0.5 * (v0[b] + v1[b]);
			if (four) avg = 0.5 * (avg + 0.5 * (v2[b] + v3[b]));
			p0[b * nn] = avg;
			p1[b * nn] = avg;
			if (four) {
				p2[b * nn] = avg;
				p3[b * nn] = avg;
			}
		}
		p0 += B * nn; p1 += B * nn; p2 += B * nn; p3 += B * nn;
	}
	for (; r < rend; r++) {
		const double v0 = p0[0], v1 = p1[0];
		double avg = 0.5 * (v0 + v1);
		if (four) avg = 0.5 * (avg + 0.5 * (p2[0] + p3[0]));
		p0[0] = avg;
		p1[0] = avg;
		if (four) {
			p2[0] = avg;
			p3[0] = avg;
		}
		p0 += nn; p1 += nn; p2 += nn; p3 += nn;
	}
}

#ifndef TBD_B
#define TBD_B 4
#endif
#ifndef TBD_MINB
#define TBD_MINB 8
#endif
__global__ void __launch_bounds__(128, TBD_MINB) k_dss_fast(DevLayout lay, DssArgs a, double * data) {
	const int gidx = tb_dss_group_block(a) * blockDim.x + threadIdx.x;
	if (gidx >= a.ngroups) return;
	if (a.flags[gidx] & 2) {
		tb_dss_local<TBD_B>(lay, a, gidx, data);
	} else {
		tb_dss_generic(lay, a, gidx, data);
	}
}

// Seam groups: (u_alpha, u_beta) of every member is re-based into each local
// target's panel with the 2x2 matrix mats[sg][t][s] (identity when t and s are
// on the same panel), restating CubedSphereTrans::CoVecPanelTrans
// (CubedSphereTrans.h:1751-2275).
struct SeamArgs {
	const int * group;        // [nseam] index into the group list
	const double * mats;      // [nseam][4][4][4]
	int nseam;
	int nlev_u;               // levels of U (and V)
};

__global__ void k_dss_seam_vector(DevLayout lay, DssArgs a, SeamArgs sa, double * data) {
	const int idx = blockIdx.x * blockDim.x + threadIdx.x;
	if (idx >= sa.nseam * sa.nlev_u) return;
	const int sg = idx / sa.nlev_u;
	const int k = idx % sa.nlev_u;
	const int gidx = sa.group[sg];
	int ms[4];
	double u[4], v[4];
	int cnt = 0;
	for (int q = 0; q < 4; q++) {
		ms[q] = a.members[4 * gidx + q];
		if (ms[q] >= 0) {
			u[q] = tb_dss_load(lay, a, data, ms[q], a.uv_row0 + k);
			v[q] = tb_dss_load(lay, a, data, ms[q], a.uv_row0 + sa.nlev_u + k);
			cnt++;
		} else {
			u[q] = 0.0;
			v[q] = 0.0;
		}
	}
	const double * M = sa.mats + (size_t)sg * 64;
	for (int tq = 0; tq < 4; tq++) {
		const int m = ms[tq];
		if (m < 0 || m >= a.nlocal) continue;
		double tu[4], tv[4];
		for (int s = 0; s < 4; s++) {
			const double * mm = M + (tq * 4 + s) * 4;
			tu[s] = mm[0] * u[s] + mm[1] * v[s];
			tv[s] = mm[2] * u[s] + mm[3] * v[s];
		}
		double au, av;
		if (cnt == 2) {
			au = 0.5 * (tu[0] + tu[1]);
			av = 0.5 * (tv[0] + tv[1]);
		} else if (cnt == 3) {
			au = (1.0 / 3.0) * (tu[0] + tu[1] + tu[2]);
			av = (1.0 / 3.0) * (tv[0] + tv[1] + tv[2]);
		} else {
			au = 0.5 * (0.5 * (tu[0] + tu[1]) + 0.5 * (tu[2] + tu[3]));
			av = 0.5 * (0.5 * (tv[0] + tv[1]) + 0.5 * (tv[2] + tv[3]));
		}
		const size_t b = tb_dss_base(lay, m);
		data[b + (size_t)(a.uv_row0 + k) * lay.nn] = au;
		data[b + (size_t)(a.uv_row0 + sa.nlev_u + k) * lay.nn] = av;
	}
}

// Pack the local nodes other ranks need: send[slot][nsel] (Grid::Exchange
// pack step, GridPatch.cpp:1292-1340).
__global__ void k_dss_pack(
	DevLayout lay, const int * send_nodes, int nsend,
	const double * data, double * send, int row0, int nsel
) {
	const long long total = (long long)nsend * nsel;
	for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	     idx < total; idx += (long long)gridDim.x * blockDim.x
	) {
		const int slot = (int)(idx / nsel);
		const int r = (int)(idx % nsel);
		const int m = send_nodes[slot];
		send[idx] = data[tb_dss_base(lay, m) + (size_t)(row0 + r) * lay.nn];
	}
}

// ---- peer-memory exchange (NVLink / NVSwitch P2P) ----------------------------
// Instead of packing into a send buffer and handing it to a collective, the
// pack kernel stores every shared node straight into the receive buffer of the
// rank that averages it (an IPC mapping of the peer's memory); a flag per
// source rank, raised after the kernel, tells the consumer the slots of this
// exchange have landed.  Receive buffers alternate by exchange parity: a rank
// that signals exchange s has finished averaging exchange s - 2 (in-order
// stream), and nobody can be more than one exchange ahead of a neighbour it
// waits for.
#define TB200_MAX_PEERS 16

struct PeerPtrs {
	double * recv[TB200_MAX_PEERS];               // peer's receive buffer of this parity
	unsigned long long * flag[TB200_MAX_PEERS];   // peer's flag for this rank
};

// The last block to finish raises this rank's flag at every neighbour (ticket
// counter; every thread's stores are fenced at system scope before its block
// takes a ticket): no separate signal launch.
__global__ void k_dss_pack_peer(
	DevLayout lay, const int * send_nodes, const int * send_rank, const int * send_slot,
	int nsend, const double * data, PeerPtrs pp, int row0, int nsel,
	unsigned * ticket, int nranks, int me, unsigned long long seq
) {
	const long long total = (long long)nsend * nsel;
	for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	     idx < total; idx += (long long)gridDim.x * blockDim.x
	) {
		const int slot = (int)(idx / nsel);
		const int r = (int)(idx % nsel);
		const int m = send_nodes[slot];
		pp.recv[send_rank[slot]][(size_t)send_slot[slot] * nsel + r] =
			data[tb_dss_base(lay, m) + (size_t)(row0 + r) * lay.nn];
	}
#ifndef TB200_EMU
	__threadfence_system();
	__syncthreads();
	if (threadIdx.x == 0) {
		const unsigned t = atomicAdd(ticket, 1u);
		if (t == gridDim.x - 1) {
			*ticket = 0u;
			__threadfence_system();
			for (int r = 0; r < nranks; r++) {
				if (r != me && pp.flag[r] != 0) {
					asm volatile("st.release.sys.global.u64 [%0], %1;"
						:: "l"(pp.flag[r]), "l"(seq) : "memory");
				}
			}
		}
	}
#endif
}


#endif
