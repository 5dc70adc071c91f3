// TEST INFRASTRUCTURE -- runtime of the CPU emulation in cuda_emul.h: the fiber switch, the per-block scheduler and the
// block-parallel launcher.  See cuda_emul.h for the model.
#include "cuda_emul.h"

#include <sys/mman.h>

#include <mutex>

// void emul_switch(void **save_sp, void *load_sp): save the callee-saved registers and the stack pointer of the running
// context, continue the context whose stack pointer is load_sp (System V x86-64 ABI).
asm(R"(
.text
.globl emul_switch
.type emul_switch,@function
emul_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size emul_switch,.-emul_switch
)");

namespace emul {

thread_local Block *g_blk = nullptr;
static thread_local char t_anchor;
char *smem_anchor() { return &t_anchor; }

static std::atomic<int> g_order_mode{[] {
    const char *e = std::getenv("MVP_EMUL_ORDER");
    return !e ? 0 : !std::strcmp(e, "reverse") ? 1 : !std::strcmp(e, "random") ? 2 : 0;
}()};

static thread_local unsigned long long t_progress;   // bumped whenever a collective completes or a lane exits

static void complete_warp(WarpSync &s) {
    const unsigned my = s.gen;
    std::memcpy(s.res[my & 1], s.val, sizeof(s.val));
    s.res_mask[my & 1] = s.lanes_arrived_mask;
    s.arrived = 0;
    s.lanes_arrived_mask = 0;
    s.gen = my + 1;
}

void lane_exited(Block *b, int t) {
    WarpSync &s = b->warp[t >> 5];
    if (s.arrived > 0 && s.arrived == warp_active(b, t >> 5)) complete_warp(s);
    if (b->bar_arrived > 0 && b->bar_arrived == b->alive) {
        const unsigned my = b->bar_gen;
        b->bar_or[my & 1] = b->bar_acc;
        b->bar_acc = 0;
        b->bar_arrived = 0;
        b->bar_gen = my + 1;
    }
}

static void fiber_entry() {
    Block *b = g_blk;
    (*b->body)();
    Lane &l = b->lane[b->cur];
    l.done = true;
    emul_switch(&l.sp, b->sched_sp);
    std::abort();   // a finished fiber is never resumed
}

static void prepare_stack(Lane &l) {
    if (!l.stack) {
        void *m = mmap(nullptr, kStackBytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        if (m == MAP_FAILED) { std::perror("emul: mmap"); std::abort(); }
        l.stack = (char *)m;
    }
    uintptr_t top = ((uintptr_t)l.stack + kStackBytes) & ~(uintptr_t)15;
    void **sp = (void **)top - 8;
    for (int i = 0; i < 6; ++i) sp[i] = nullptr;   // r15 r14 r13 r12 rbx rbp
    sp[6] = (void *)fiber_entry;                   // popped by `ret`
    sp[7] = nullptr;                               // fake return address: fiber_entry starts with rsp = 8 mod 16
    l.sp = sp;
}

static void run_block(Block *b, unsigned bx, unsigned by, unsigned bz) {
    b->bid = make_uint3(bx, by, bz);
    const int n = b->nthreads;
    for (int t = 0; t < n; ++t) {
        Lane &l = b->lane[t];
        l.linear = t;
        l.tid = make_uint3(t % b->bdim.x, (t / b->bdim.x) % b->bdim.y, t / (b->bdim.x * b->bdim.y));
        l.done = false;
        prepare_stack(l);
    }
    for (int w = 0; w < (n + 31) / 32; ++w) {
        WarpSync &s = b->warp[w];
        s.arrived = 0; s.gen = 0; s.lanes_arrived_mask = 0; s.res_mask[0] = s.res_mask[1] = 0;
    }
    b->alive = n;
    b->bar_arrived = 0; b->bar_gen = 0; b->bar_acc = 0; b->bar_or[0] = b->bar_or[1] = 0;
    g_blk = b;
    unsigned long long last = ~0ull;
    int idle = 0;
    // Lane schedule of a pass.  Real warps give no ordering between lanes outside collectives, so the tests run the kernels
    // under several schedules (MVP_EMUL_ORDER = forward | reverse | random): code that only works when lane 0 runs first
    // (emul_set_lane_order() changes it at run time)
    // (e.g. a missing __syncwarp between a shared-memory write and another lane's read) fails under one of them.
    const int order_mode = g_order_mode.load();
    unsigned rng = 0x9e3779b9u ^ (bx * 73856093u) ^ (by * 19349663u) ^ (bz * 83492791u);
    std::vector<int> order(n);
    for (int t = 0; t < n; ++t) order[t] = order_mode == 1 ? n - 1 - t : t;
    while (b->alive > 0) {
        if (order_mode == 2) {
            for (int i = n - 1; i > 0; --i) {
                rng = rng * 1664525u + 1013904223u;
                std::swap(order[i], order[(rng >> 8) % (unsigned)(i + 1)]);
            }
        }
        for (int oi = 0; oi < n; ++oi) {
            const int t = order[oi];
            Lane &l = b->lane[t];
            if (l.done) continue;
            b->cur = t;
            emul_switch(&b->sched_sp, l.sp);
            if (l.done) {
                b->alive--;
                t_progress++;
                lane_exited(b, t);
            }
        }
        // every collective completion changes a generation counter; a full pass without any is a deadlock
        unsigned long long sig = t_progress;
        for (int w = 0; w < (n + 31) / 32; ++w) sig = sig * 1315423911ull + b->warp[w].gen;
        sig = sig * 1315423911ull + b->bar_gen;
        if (sig == last) {
            if (++idle > 4) {
                std::fprintf(stderr, "emul: deadlock in block (%u,%u,%u): %d threads alive, none can make progress\n", bx, by, bz, b->alive);
                std::abort();
            }
        } else {
            idle = 0;
        }
        last = sig;
    }
    g_blk = nullptr;
}

void launch(dim3 grid, dim3 block, size_t dyn_smem, const std::function<void()> &body) {
    const size_t nblocks = (size_t)grid.x * grid.y * grid.z;
    const int nthreads = (int)(block.x * block.y * block.z);
    if (nblocks == 0) return;
    if (nthreads <= 0 || nthreads > kMaxThreads) { std::fprintf(stderr, "emul: bad block size %d\n", nthreads); std::abort(); }
    unsigned hw = std::thread::hardware_concurrency();
    if (const char *e = std::getenv("MVP_EMUL_THREADS")) hw = (unsigned)std::atoi(e);
    const size_t nworkers = std::max<size_t>(1, std::min<size_t>(std::min<size_t>(hw ? hw : 1, 16), nblocks));
    std::atomic<size_t> next{0};
    static std::mutex pool_mutex;
    static std::vector<Block *> pool_free;   // blocks (with their lanes' stacks) are reused across launches
    auto worker = [&] {
        Block *b = nullptr;
        {
            std::lock_guard<std::mutex> g(pool_mutex);
            if (!pool_free.empty()) { b = pool_free.back(); pool_free.pop_back(); }
        }
        if (!b) b = new Block();
        b->nthreads = nthreads;
        b->bdim = block;
        b->gdim = grid;
        b->body = &body;
        if (dyn_smem > b->dyn_bytes) {
            std::free(b->dyn_smem);
            b->dyn_smem = (char *)std::aligned_alloc(128, (dyn_smem + 127) / 128 * 128);
            b->dyn_bytes = dyn_smem;
        }
        for (;;) {
            const size_t i = next.fetch_add(1);
            if (i >= nblocks) break;
            const unsigned bx = (unsigned)(i % grid.x), by = (unsigned)((i / grid.x) % grid.y), bz = (unsigned)(i / ((size_t)grid.x * grid.y));
            run_block(b, bx, by, bz);
        }
        std::lock_guard<std::mutex> g(pool_mutex);
        pool_free.push_back(b);
    };
    if (nworkers == 1) {
        worker();
        return;
    }
    std::vector<std::thread> pool;
    for (size_t i = 0; i < nworkers; ++i) pool.emplace_back(worker);
    for (auto &t : pool) t.join();
}

}  // namespace emul

// lane schedule for subsequent launches: 0 forward, 1 reverse, 2 random (seeded per block)
extern "C" void emul_set_lane_order(int mode) { emul::g_order_mode.store(mode); }
