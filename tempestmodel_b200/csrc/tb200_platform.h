// Platform switch: real CUDA (product) or the host emulation used by the
// GPU-less kernel-logic tests (tests/emu/cuda_emu.h, never shipped).
#ifndef TB200_PLATFORM_H
#define TB200_PLATFORM_H

#ifdef TB200_EMU
#include "cuda_emu.h"
// Kernels that use __syncthreads / shuffles
#define TB_LAUNCH(kfn, grid, block, smem, stream, ...) \
	tbemu::Launch((grid), (block), (smem), [&]() { kfn(__VA_ARGS__); })
// Kernels without intra-block communication
#define TB_LAUNCH_FLAT(kfn, grid, block, smem, stream, ...) \
	tbemu::LaunchFlat((grid), (block), (smem), [&]() { kfn(__VA_ARGS__); })
#define TB_DYN_SMEM(T, name) T * name = reinterpret_cast<T *>(tbemu::S().dyn_smem)
#else
#include <cuda_runtime.h>
#define TB_LAUNCH(kfn, grid, block, smem, stream, ...) \
	kfn<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define TB_LAUNCH_FLAT(kfn, grid, block, smem, stream, ...) \
	kfn<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define TB_DYN_SMEM(T, name) \
	extern __shared__ __align__(16) unsigned char tb_dyn_smem_raw[]; \
	T * name = reinterpret_cast<T *>(tb_dyn_smem_raw)
#endif

#endif
