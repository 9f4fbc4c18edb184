// Launch / shared-memory macros.  The product is compiled by nvcc for sm_100a.  -DQPIC_EMU (g++) swaps in the
// development-only fiber emulator of tools/emu/ so kernel logic can be debugged in the GPU-less build container;
// that build produces a separate debug library which the package never loads (tools/emu/qpic_emu.h).
#pragma once

#ifdef QPIC_EMU
#include "qpic_emu.h"
#else
#include <cuda_runtime.h>
#define QPIC_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define QPIC_LAUNCH_FLAT(kernel, grid, block, smem, stream, ...) kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define QPIC_DYN_SMEM(type, name)                                        \
    extern __shared__ __align__(16) unsigned char qpic_dyn_smem_raw[];  \
    type * name = reinterpret_cast<type *>(qpic_dyn_smem_raw)
#endif
