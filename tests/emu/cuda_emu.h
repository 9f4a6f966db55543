// Host emulation of the small CUDA subset the tempest-b200 kernels use.
//
// TEST INFRASTRUCTURE ONLY.  This header lets the *same* kernel sources under
// tempestmodel_b200/csrc be compiled with g++ into tests/emu/libtb200_emu.so,
// so that kernel logic (indexing, operator order, connectivity) is checked
// against the oracle in this GPU-less container before GPU minutes are spent.
// The product library (libtempest_b200.so, nvcc, sm_100a) never includes it
// and the Python package never loads the emulation library.
//
// Model: blocks run one after another; the threads of a block are ucontext
// fibers on one OS thread, scheduled round-robin; __syncthreads and the
// warp shuffles are fiber barriers.
#ifndef TB200_CUDA_EMU_H
#define TB200_CUDA_EMU_H

#include <ucontext.h>
#include <cstdlib>
#include <cstring>
#include <cstdio>
#include <cmath>
#include <cstdint>
#include <vector>
#include <functional>
#include <algorithm>

#define TB200_EMU 1

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __grid_constant__
#define __shared__ static

struct int4 { int x, y, z, w; };

struct dim3 {
	unsigned x, y, z;
	dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct uint3_emu { unsigned x, y, z; };

namespace tbemu {

struct State {
	uint3_emu threadIdx, blockIdx;
	dim3 blockDim, gridDim;
	unsigned char * dyn_smem;
};
inline State & S() { static State s; return s; }

enum { RUN = 0, AT_BAR = 1, AT_WBAR = 2, DONE = 3 };

struct Fiber {
	ucontext_t ctx;
	int state;
	unsigned tid;
};

struct Sched {
	ucontext_t main;
	std::vector<Fiber> fibers;
	std::vector<unsigned char *> stacks;
	const std::function<void()> * body;
	unsigned nthreads;
	unsigned cur;
	double wbuf[32][32];
	unsigned long long wbuf_u[32][32];
};
inline Sched & G() { static Sched g; return g; }

static const size_t kStack = 512 * 1024;

inline void SetThread(unsigned t) {
	State & s = S();
	s.threadIdx.x = t % s.blockDim.x;
	s.threadIdx.y = (t / s.blockDim.x) % s.blockDim.y;
	s.threadIdx.z = t / (s.blockDim.x * s.blockDim.y);
}

inline void Trampoline() {
	Sched & g = G();
	(*g.body)();
	g.fibers[g.cur].state = DONE;
	swapcontext(&g.fibers[g.cur].ctx, &g.main);
}

inline void Yield(int newstate) {
	Sched & g = G();
	unsigned me = g.cur;
	g.fibers[me].state = newstate;
	swapcontext(&g.fibers[me].ctx, &g.main);
	SetThread(me);
}

inline void RunBlock(const std::function<void()> & body, unsigned nthreads) {
	Sched & g = G();
	g.body = &body;
	g.nthreads = nthreads;
	if (g.fibers.size() < nthreads) g.fibers.resize(nthreads);
	while (g.stacks.size() < nthreads) {
		g.stacks.push_back((unsigned char *)malloc(kStack));
	}
	for (unsigned t = 0; t < nthreads; t++) {
		Fiber & f = g.fibers[t];
		getcontext(&f.ctx);
		f.ctx.uc_stack.ss_sp = g.stacks[t];
		f.ctx.uc_stack.ss_size = kStack;
		f.ctx.uc_link = &g.main;
		f.state = RUN;
		f.tid = t;
		makecontext(&f.ctx, (void (*)())Trampoline, 0);
	}
	for (;;) {
		unsigned ndone = 0, nbar = 0;
		for (unsigned t = 0; t < nthreads; t++) {
			if (g.fibers[t].state == RUN) {
				g.cur = t;
				SetThread(t);
				swapcontext(&g.main, &g.fibers[t].ctx);
			}
		}
		// release warp barriers
		for (unsigned w = 0; w * 32 < nthreads; w++) {
			unsigned lo = w * 32, hi = std::min(nthreads, lo + 32);
			unsigned nw = 0, nalive = 0;
			for (unsigned t = lo; t < hi; t++) {
				if (g.fibers[t].state == AT_WBAR) nw++;
				if (g.fibers[t].state != DONE) nalive++;
			}
			if (nw != 0 && nw == nalive) {
				for (unsigned t = lo; t < hi; t++) {
					if (g.fibers[t].state == AT_WBAR) g.fibers[t].state = RUN;
				}
			}
		}
		bool anyrun = false;
		for (unsigned t = 0; t < nthreads; t++) {
			if (g.fibers[t].state == DONE) ndone++;
			if (g.fibers[t].state == AT_BAR) nbar++;
			if (g.fibers[t].state == RUN) anyrun = true;
		}
		if (ndone == nthreads) break;
		if (anyrun) continue;
		if (nbar != 0 && nbar + ndone == nthreads) {
			for (unsigned t = 0; t < nthreads; t++) {
				if (g.fibers[t].state == AT_BAR) g.fibers[t].state = RUN;
			}
			continue;
		}
		fprintf(stderr, "cuda_emu: deadlock (divergent barrier)\n");
		abort();
	}
}

// Launch with fibers (kernels that synchronise or shuffle)
inline void Launch(dim3 grid, dim3 block, size_t smem, const std::function<void()> & body) {
	State & s = S();
	s.gridDim = grid;
	s.blockDim = block;
	std::vector<unsigned char> dyn(smem + 16);
	s.dyn_smem = dyn.data();
	unsigned nthreads = block.x * block.y * block.z;
	for (unsigned bz = 0; bz < grid.z; bz++)
	for (unsigned by = 0; by < grid.y; by++)
	for (unsigned bx = 0; bx < grid.x; bx++) {
		s.blockIdx.x = bx; s.blockIdx.y = by; s.blockIdx.z = bz;
		RunBlock(body, nthreads);
	}
}

// Launch without fibers (kernels with no block- or warp-level communication)
inline void LaunchFlat(dim3 grid, dim3 block, size_t smem, const std::function<void()> & body) {
	State & s = S();
	s.gridDim = grid;
	s.blockDim = block;
	std::vector<unsigned char> dyn(smem + 16);
	s.dyn_smem = dyn.data();
	unsigned nthreads = block.x * block.y * block.z;
	for (unsigned bz = 0; bz < grid.z; bz++)
	for (unsigned by = 0; by < grid.y; by++)
	for (unsigned bx = 0; bx < grid.x; bx++) {
		s.blockIdx.x = bx; s.blockIdx.y = by; s.blockIdx.z = bz;
		for (unsigned t = 0; t < nthreads; t++) {
			SetThread(t);
			body();
		}
	}
}

}  // namespace tbemu

#define threadIdx (tbemu::S().threadIdx)
#define blockIdx (tbemu::S().blockIdx)
#define blockDim (tbemu::S().blockDim)
#define gridDim (tbemu::S().gridDim)

inline void __syncthreads() { tbemu::Yield(tbemu::AT_BAR); }
inline void __syncwarp(unsigned = 0xffffffffu) { tbemu::Yield(tbemu::AT_WBAR); }
inline void __threadfence() {}
inline void __threadfence_block() {}
inline void __threadfence_system() {}

inline double __shfl_sync(unsigned, double v, int src, int width = 32) {
	tbemu::Sched & g = tbemu::G();
	unsigned t = g.cur, w = t / 32, lane = t % 32;
	g.wbuf[w][lane] = v;
	tbemu::Yield(tbemu::AT_WBAR);
	unsigned base = lane & ~(unsigned)(width - 1);
	double r = g.wbuf[w][base + ((unsigned)src % (unsigned)width)];
	tbemu::Yield(tbemu::AT_WBAR);
	return r;
}
inline double __shfl_xor_sync(unsigned m, double v, int lanemask, int width = 32) {
	unsigned lane = tbemu::G().cur % 32;
	return __shfl_sync(m, v, (int)((lane ^ (unsigned)lanemask) % (unsigned)width), width);
}
inline double __shfl_down_sync(unsigned m, double v, unsigned delta, int width = 32) {
	unsigned lane = tbemu::G().cur % 32;
	unsigned l = lane % width;
	return __shfl_sync(m, v, (int)((l + delta < (unsigned)width) ? l + delta : l), width);
}

inline int __shfl_sync(unsigned m, int v, int src, int width = 32) {
	return (int)__shfl_sync(m, (double)v, src, width);
}
inline int __shfl_down_sync(unsigned m, int v, unsigned delta, int width = 32) {
	return (int)__shfl_down_sync(m, (double)v, delta, width);
}

inline int __any_sync(unsigned, int pred) {
	tbemu::Sched & g = tbemu::G();
	unsigned t = g.cur, w = t / 32, lane = t % 32;
	// lanes beyond the block's thread count do not exist: their slots must not vote
	const unsigned nlanes = (g.nthreads - w * 32 < 32) ? (g.nthreads - w * 32) : 32;
	g.wbuf[w][lane] = pred ? 1.0 : 0.0;
	tbemu::Yield(tbemu::AT_WBAR);
	int r = 0;
	for (unsigned l = 0; l < nlanes; l++) r |= (g.wbuf[w][l] != 0.0);
	tbemu::Yield(tbemu::AT_WBAR);
	return r;
}
inline int __popc(unsigned v) { return __builtin_popcount(v); }

inline double __ldg(const double * p) { return *p; }
inline int __ldg(const int * p) { return *p; }
inline double atomicAdd(double * p, double v) { double o = *p; *p += v; return o; }
inline int atomicAdd(int * p, int v) { int o = *p; *p += v; return o; }
inline int atomicMax(int * p, int v) { int o = *p; if (v > o) *p = v; return o; }
inline int atomicOr(int * p, int v) { int o = *p; *p |= v; return o; }
inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
inline double __drcp_rn(double a) { return 1.0 / a; }
inline double __ddiv_rn(double a, double b) { return a / b; }

// ---- runtime API subset ----------------------------------------------------
typedef int cudaError_t;
typedef void * cudaStream_t;
typedef void * cudaEvent_t;
enum { cudaSuccess = 0, cudaErrorEmu = 1 };
enum cudaMemcpyKind {
	cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice,
	cudaMemcpyHostToHost, cudaMemcpyDefault
};
enum { cudaHostRegisterDefault = 0, cudaStreamNonBlocking = 1 };
inline const char * cudaGetErrorString(cudaError_t) { return "emulated"; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaGetDevice(int * d) { *d = 0; return cudaSuccess; }
inline cudaError_t cudaGetDeviceCount(int * n) { *n = 1; return cudaSuccess; }
// fresh device memory is NOT zero: poison it (NaN pattern) so that reads of
// never-written entries show up in the emulated tests
inline cudaError_t cudaMalloc(void ** p, size_t n) { *p = malloc(n ? n : 1); if (*p) memset(*p, 0xFF, n ? n : 1); return *p ? cudaSuccess : cudaErrorEmu; }
inline cudaError_t cudaFree(void * p) { free(p); return cudaSuccess; }
inline cudaError_t cudaMallocHost(void ** p, size_t n) { *p = malloc(n ? n : 1); return cudaSuccess; }
inline cudaError_t cudaFreeHost(void * p) { free(p); return cudaSuccess; }
inline cudaError_t cudaMemcpy(void * d, const void * s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void * d, const void * s, size_t n, cudaMemcpyKind, cudaStream_t = 0) { memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpy2DAsync(void * d, size_t dp, const void * s, size_t sp, size_t w, size_t h, cudaMemcpyKind, cudaStream_t = 0) {
	for (size_t r = 0; r < h; r++) memmove((char *)d + r * dp, (const char *)s + r * sp, w);
	return cudaSuccess;
}
inline cudaError_t cudaMemset(void * d, int v, size_t n) { memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void * d, int v, size_t n, cudaStream_t = 0) { memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaStreamCreate(cudaStream_t * s) { *s = 0; return cudaSuccess; }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t * s, unsigned) { *s = 0; return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
inline cudaError_t cudaHostRegister(void *, size_t, unsigned) { return cudaSuccess; }
inline cudaError_t cudaHostUnregister(void *) { return cudaSuccess; }
template <typename T>
inline cudaError_t cudaFuncSetAttribute(T, int, int) { return cudaSuccess; }
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };

#endif
