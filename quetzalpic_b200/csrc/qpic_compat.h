// Launch / shared-memory macros.  The product is compiled by nvcc for sm_100a.  -DQPIC_EMU (g++) swaps in the
// development-only fiber emulator of tools/emu/ so kernel logic can be debugged in the GPU-less build container;
// that build produces a separate debug library which the package never loads (tools/emu/qpic_emu.h).
#pragma once

#ifdef QPIC_EMU
#include "qpic_emu.h"
#else
#include <cuda_runtime.h>
#include <cstdlib>
#include <utility>
// Kernels are launched with programmatic stream serialization (PDL): the next kernel of the stream is staged while this
// one drains; every kernel begins with griddepcontrol.wait (QPIC_KERNEL_BEGIN).  Measured on B200 (tools/timeline.py,
// profiles/r02_timeline_*.txt): the graph's kernel-to-kernel gaps are already ~1.3 us and the staged blocks of the next
// kernel slow the running one down (0.98 -> 1.05 ms/slice), so the attribute is OFF unless QPIC_PDL=1.
namespace qpic {
inline bool pdl_enabled() {
    static const bool on = [] {
        const char * e = getenv("QPIC_PDL");
        return e && e[0] == '1';
    }();
    return on;
}
template <class... Exp, class... Act>
inline cudaError_t launch_kernel(void (*kernel)(Exp...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Act &&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Act>(args)...);
}
// cooperative launch: the driver guarantees that all blocks of the grid are resident at the same time (what a grid-wide barrier
// inside the kernel needs); valid inside stream capture (the attribute becomes a kernel node attribute)
}  // namespace qpic
#define QPIC_LAUNCH(kernel, grid, block, smem, stream, ...) (void)qpic::launch_kernel(kernel, dim3(grid), dim3(block), (smem), (stream), __VA_ARGS__)
#define QPIC_LAUNCH_FLAT(kernel, grid, block, smem, stream, ...) (void)qpic::launch_kernel(kernel, dim3(grid), dim3(block), (smem), (stream), __VA_ARGS__)
#define QPIC_DYN_SMEM(type, name)                                        \
    extern __shared__ __align__(16) unsigned char qpic_dyn_smem_raw[];  \
    type * name = reinterpret_cast<type *>(qpic_dyn_smem_raw)
#endif
