// DEVELOPMENT / TEST TOOL ONLY -- see qpic_emu.h.  Fiber scheduler for the CPU kernel emulator.
#include "qpic_emu.h"

#undef threadIdx
#undef blockIdx
#undef blockDim
#undef gridDim

namespace qemu {

// thread_local: two contexts can be driven from two host threads (the slab-decomposed sweep, tools/emu/emu_dist_check.py)
thread_local Block * g_block = nullptr;
thread_local Fiber * g_fiber = nullptr;

static const size_t STACK_BYTES = 256 * 1024;

static void fiber_entry() {
    Block * b = g_block;
    Fiber * f = g_fiber;
    b->body();
    f->done = true;
    b->live--;
    b->warp_live[f->linear / 32]--;
    // a finished thread no longer takes part in barriers: release waiters if it was the last one missing
    if (b->live > 0 && b->bar_arrived >= b->live) {
        b->bar_arrived = 0;
        b->bar_gen++;
    }
    const int w = f->linear / 32;
    if (b->warp_live[w] > 0 && b->warp_arrived[w] >= b->warp_live[w]) {
        b->warp_arrived[w] = 0;
        b->warp_gen[w]++;
    }
    swapcontext(&f->ctx, &b->sched);
}

void yield_fiber() { swapcontext(&g_fiber->ctx, &g_block->sched); }

void sync_threads() {
    Block * b = g_block;
    if (b->flat) {
        std::fprintf(stderr, "qemu: __syncthreads inside a kernel launched as FLAT\n");
        std::abort();
    }
    const unsigned long long gen = b->bar_gen;
    b->bar_arrived++;
    if (b->bar_arrived >= b->live) {
        b->bar_arrived = 0;
        b->bar_gen++;
        return;
    }
    while (b->bar_gen == gen) yield_fiber();
}

void sync_warp() {
    Block * b = g_block;
    if (b->flat) return;
    const int w = g_fiber->linear / 32;
    const unsigned long long gen = b->warp_gen[w];
    b->warp_arrived[w]++;
    if (b->warp_arrived[w] >= b->warp_live[w]) {
        b->warp_arrived[w] = 0;
        b->warp_gen[w]++;
        return;
    }
    while (b->warp_gen[w] == gen) yield_fiber();
}

void launch(dim3 grid, dim3 block, size_t smem, bool flat, const std::function<void()> & body) {
    Block blk;
    blk.grid = grid;
    blk.block = block;
    blk.nthreads = (int)(block.x * block.y * block.z);
    blk.flat = flat;
    blk.body = body;
    blk.dyn_smem = (char *)std::malloc(smem ? smem : 16);
    const int nwarps = (blk.nthreads + 31) / 32;
    Block * saved_block = g_block;
    Fiber * saved_fiber = g_fiber;
    g_block = &blk;

    if (flat) {
        Fiber f;
        g_fiber = &f;
        for (unsigned bz = 0; bz < grid.z; bz++)
            for (unsigned by = 0; by < grid.y; by++)
                for (unsigned bx = 0; bx < grid.x; bx++) {
                    blk.bid = uint3{bx, by, bz};
                    int lin = 0;
                    for (unsigned tz = 0; tz < block.z; tz++)
                        for (unsigned ty = 0; ty < block.y; ty++)
                            for (unsigned tx = 0; tx < block.x; tx++) {
                                f.tid = uint3{tx, ty, tz};
                                f.linear = lin++;
                                body();
                            }
                }
    } else {
        blk.fibers.resize(blk.nthreads);
        for (auto & f : blk.fibers) f.stack = (char *)std::malloc(STACK_BYTES);
        for (unsigned bz = 0; bz < grid.z; bz++)
            for (unsigned by = 0; by < grid.y; by++)
                for (unsigned bx = 0; bx < grid.x; bx++) {
                    blk.bid = uint3{bx, by, bz};
                    blk.live = blk.nthreads;
                    blk.bar_arrived = 0;
                    blk.warp_arrived.assign(nwarps, 0);
                    blk.warp_gen.assign(nwarps, 0);
                    blk.warp_live.assign(nwarps, 0);
                    int lin = 0;
                    for (unsigned tz = 0; tz < block.z; tz++)
                        for (unsigned ty = 0; ty < block.y; ty++)
                            for (unsigned tx = 0; tx < block.x; tx++) {
                                Fiber & f = blk.fibers[lin];
                                f.tid = uint3{tx, ty, tz};
                                f.linear = lin;
                                f.done = false;
                                blk.warp_live[lin / 32]++;
                                getcontext(&f.ctx);
                                f.ctx.uc_stack.ss_sp = f.stack;
                                f.ctx.uc_stack.ss_size = STACK_BYTES;
                                f.ctx.uc_link = &blk.sched;
                                makecontext(&f.ctx, (void (*)())fiber_entry, 0);
                                lin++;
                            }
                    // round-robin until every fiber has finished
                    while (blk.live > 0) {
                        bool progressed = false;
                        for (int t = 0; t < blk.nthreads; t++) {
                            Fiber & f = blk.fibers[t];
                            if (f.done) continue;
                            g_fiber = &f;
                            const int live_before = blk.live;
                            const unsigned long long gen_before = blk.bar_gen;
                            swapcontext(&blk.sched, &f.ctx);
                            (void)live_before;
                            (void)gen_before;
                            progressed = true;
                        }
                        if (!progressed) break;
                    }
                }
        for (auto & f : blk.fibers) std::free(f.stack);
    }
    std::free(blk.dyn_smem);
    g_block = saved_block;
    g_fiber = saved_fiber;
}

}  // namespace qemu
