// Bulk asynchronous copies (TMA) and transaction barriers of sm_100a, as used by
// the pipelined element kernels (tb200_fast.cuh).
//
// An element's state is one contiguous block of nrows rows of 128 bytes (16
// nodes): the textbook 2-D tile.  One tensor map per state instance describes the
// instance as a [nelem * nrows][16] array of doubles with SWIZZLE_128B, so that
// one elected thread moves an element into shared memory with a single
// cp.async.bulk.tensor.2d (UTMALDG) and the hardware stores the 16-byte chunk c
// of row r at chunk c ^ (r & 7) of that row - conflict-free for both access
// patterns of the kernels (a thread's own 32 bytes of a row; a whole row,
// broadcast within the level's four threads).  Completion is signalled on an
// mbarrier by transaction bytes; consumers wait on its phase parity.
//
// The host emulation (TB200_EMU) performs the same copies synchronously with
// the same swizzle; barriers are no-ops there.
#ifndef TB200_TMA_CUH
#define TB200_TMA_CUH

#include "tb200_platform.h"

// Tensor map of one state instance + the plain pointer (emulation, 1-D copies).
struct alignas(64) TbMap {
	unsigned char desc[128];      // CUtensorMap (cuTensorMapEncodeTiled)
	const double * base;
	int boxrows;                  // rows per copy (<= 256)
	int nbox;                     // copies per element
	int pad_[12];
};

// offset (doubles) of node pair `chunk` (0..7) of row r in a swizzled element buffer
__device__ __forceinline__ int tb_swz(int r, int chunk) {
	return ((r << 3) | (chunk ^ (r & 7))) << 1;
}

#ifdef TB200_EMU

typedef unsigned long long tb_mbar_t;
__device__ __forceinline__ void tb_mbar_init(tb_mbar_t *, int) {}
__device__ __forceinline__ void tb_mbar_fence_init() {}
__device__ __forceinline__ void tb_mbar_expect(tb_mbar_t *, unsigned) {}
__device__ __forceinline__ void tb_mbar_wait(tb_mbar_t *, unsigned) {}
__device__ __forceinline__ void tb_mbar_arrive(tb_mbar_t *) {}

// element rows [row0, row0 + nrows) of the instance -> swizzled buffer
__device__ __forceinline__ void tb_tma_element(
	double * dst, const TbMap & m, long long row0, int nrows, tb_mbar_t *
) {
	const double * src = m.base + (size_t)row0 * 16;
	for (int r = 0; r < nrows; r++) {
		for (int c = 0; c < 8; c++) {
			dst[tb_swz(r, c)] = src[r * 16 + 2 * c];
			dst[tb_swz(r, c) + 1] = src[r * 16 + 2 * c + 1];
		}
	}
}

__device__ __forceinline__ void tb_bulk_1d(double * dst, const double * src, unsigned bytes, tb_mbar_t *) {
	for (unsigned q = 0; q < bytes / 8; q++) dst[q] = src[q];
}

#else

typedef unsigned long long tb_mbar_t;

__device__ __forceinline__ unsigned tb_smem_u32(const void * p) {
	return (unsigned)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void tb_mbar_init(tb_mbar_t * bar, int count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(tb_smem_u32(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void tb_mbar_fence_init() {
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

// one arrival + the bytes the copies issued next will deliver
__device__ __forceinline__ void tb_mbar_expect(tb_mbar_t * bar, unsigned bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
		:: "r"(tb_smem_u32(bar)), "r"(bytes) : "memory");
}

// plain arrival (release at CTA scope): "this warp is done with ..."
__device__ __forceinline__ void tb_mbar_arrive(tb_mbar_t * bar) {
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(tb_smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void tb_mbar_wait(tb_mbar_t * bar, unsigned parity) {
	const unsigned a = tb_smem_u32(bar);
	unsigned done;
	do {
		asm volatile(
			"{\n\t.reg .pred p;\n\t"
			"mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
			"selp.b32 %0, 1, 0, p;\n\t}"
			: "=r"(done) : "r"(a), "r"(parity) : "memory");
	} while (!done);
}

// element rows [row0, row0 + nrows) of the instance -> swizzled buffer (1024-byte
// aligned), m.nbox boxes of m.boxrows rows each; one thread calls this
__device__ __forceinline__ void tb_tma_element(
	double * dst, const TbMap & m, long long row0, int nrows, tb_mbar_t * bar
) {
	(void)nrows;
	const unsigned b = tb_smem_u32(bar);
	for (int q = 0; q < m.nbox; q++) {
		const unsigned d = tb_smem_u32(dst + (size_t)q * m.boxrows * 16);
		const int c1 = (int)(row0 + (long long)q * m.boxrows);
		asm volatile(
			"cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
			" [%0], [%1, {%3, %4}], [%2];"
			:: "r"(d), "l"(reinterpret_cast<unsigned long long>(m.desc)), "r"(b), "r"(0), "r"(c1)
			: "memory");
	}
}

// contiguous bytes (multiple of 16, 16-byte aligned on both sides)
__device__ __forceinline__ void tb_bulk_1d(double * dst, const double * src, unsigned bytes, tb_mbar_t * bar) {
	asm volatile(
		"cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
		:: "r"(tb_smem_u32(dst)), "l"(src), "r"(bytes), "r"(tb_smem_u32(bar)) : "memory");
}

#endif

// bytes one tb_tma_element delivers
__host__ __device__ inline unsigned tb_tma_element_bytes(const TbMap & m) {
	return (unsigned)m.nbox * (unsigned)m.boxrows * 128u;
}

// Boxes of an element: at most 256 rows each, a multiple of 8 rows (1 KiB, the
// period of the swizzle) so that every box starts on a 1024-byte boundary of the
// buffer; the last box may read a few rows of the next element (or zeros past
// the end of the instance), which nobody looks at.
__host__ __device__ inline int tb_tma_nbox(int nrows) { return (nrows + 255) / 256; }
__host__ __device__ inline int tb_tma_boxrows(int nrows) {
	const int nbox = tb_tma_nbox(nrows);
	return (((nrows + nbox - 1) / nbox) + 7) / 8 * 8;
}

// doubles of shared memory an element buffer needs (whole boxes)
__host__ __device__ inline size_t tb_tma_buffer_doubles(int nrows) {
	return (size_t)tb_tma_nbox(nrows) * tb_tma_boxrows(nrows) * 16;
}

#endif
