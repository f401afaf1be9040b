// cuda_emu.cpp — runtime of the test-only warp emulator (see cuda_emu.h).
#include "cuda_emu.h"

namespace sse {
uint8_t smem[emu::MAX_SMEM] __attribute__((aligned(16)));
}

// x86-64 SysV context switch: callee-saved registers + stack pointer.
asm(R"(
.text
.globl emu_switch
.type emu_switch,@function
emu_switch:
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
.size emu_switch,.-emu_switch
)");

namespace emu {

Cta g_cta;
Fiber *g_cur = nullptr;
emu_dim3 g_blockIdx, g_blockDim, g_gridDim;
uint64_t g_clock = 0;

static const char *op_name(int op) {
    switch (op) {
        case OP_BALLOT: return "ballot";
        case OP_SHFL: return "shfl";
        case OP_SHFL_UP: return "shfl_up";
        case OP_SHFL_DOWN: return "shfl_down";
        case OP_SHFL_XOR: return "shfl_xor";
        case OP_SYNCWARP: return "syncwarp";
    }
    return "-";
}

static void report() {
    fprintf(stderr, "[cuda_emu] block %u, %zu threads:\n", g_blockIdx.x, g_cta.f.size());
    for (const Fiber &f : g_cta.f) {
        if (f.done) continue;
        fprintf(stderr, "  thread %3u (warp %d lane %2d): %s%s mask=%08x%s\n", f.tid.x, f.warp, f.lane,
                f.arrived ? "waiting in " : "running", f.arrived ? op_name(f.op) : "", f.cmask, f.at_bar ? " [at __syncthreads]" : "");
    }
}

[[noreturn]] void die(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    fprintf(stderr, "[cuda_emu] FATAL: ");
    vfprintf(stderr, fmt, ap);
    fprintf(stderr, "\n");
    va_end(ap);
    if (g_cur) fprintf(stderr, "[cuda_emu] in block %u thread %u\n", g_blockIdx.x, g_cur->tid.x);
    report();
    fflush(stderr);
    abort();
}

void yield() { emu_switch(&g_cur->sp, g_cta.sched_sp); }

static uint64_t g_polls = 0;
void poll_yield() {
    if (++g_polls > (1ull << 34)) die("livelock: 2^34 polls without the launch ending");
    ++g_cta.progress;  // other fibers may have published what this one waits for; deadlocks show up as the poll limit
    yield();
}

static inline int src_lane(int op, int lane, int arg, int width) {
    const int base = lane & ~(width - 1), rel = lane & (width - 1);
    switch (op) {
        case OP_SHFL: return base | (arg & (width - 1));
        case OP_SHFL_UP: return rel - arg >= 0 ? lane - arg : lane;
        case OP_SHFL_DOWN: return rel + arg < width ? lane + arg : lane;
        case OP_SHFL_XOR: return ((rel ^ arg) < width) ? (base | (rel ^ arg)) : lane;
    }
    return lane;
}

uint64_t collective(int op, uint32_t mask, uint64_t val, int arg, int width) {
    Fiber *me = g_cur;
    if (!((mask >> me->lane) & 1u)) die("%s: calling lane %d is not in its own mask %08x", op_name(op), me->lane, mask);
    if (width < 1 || width > 32 || (width & (width - 1))) die("%s: bad width %d", op_name(op), width);
    Fiber *w = &g_cta.f[(size_t)me->warp * 32];
    const int nl = (int)std::min<size_t>(32, g_cta.f.size() - (size_t)me->warp * 32);
    me->op = op;
    me->cmask = mask;
    me->val = val;
    me->arg = arg;
    me->width = width;
    me->arrived = true;
    me->released = false;
    bool all = true;
    for (int j = 0; j < 32; ++j) {
        if (!((mask >> j) & 1u)) continue;
        if (j >= nl) die("%s: mask %08x names lane %d of a %d-thread warp", op_name(op), mask, j, nl);
        if (w[j].done) die("%s: mask %08x names lane %d, which has exited", op_name(op), mask, j);
        if (!w[j].arrived) { all = false; continue; }
        if (w[j].cmask != mask || w[j].op != op)
            die("divergent collective: lane %d is in %s(mask %08x), lane %d in %s(mask %08x)", me->lane, op_name(op), mask, j,
                op_name(w[j].op), w[j].cmask);
    }
    if (all) {
        uint32_t ballot = 0;
        if (op == OP_BALLOT)
            for (int j = 0; j < 32; ++j)
                if (((mask >> j) & 1u) && w[j].val) ballot |= 1u << j;
        for (int j = 0; j < 32; ++j) {
            if (!((mask >> j) & 1u)) continue;
            uint64_t r = 0;
            if (op == OP_BALLOT) r = ballot;
            else if (op != OP_SYNCWARP) {
                const int s = src_lane(op, j, w[j].arg, w[j].width);
                if (!((mask >> s) & 1u)) die("%s: lane %d reads lane %d, which is not in mask %08x", op_name(op), j, s, mask);
                r = w[s].val;
            }
            w[j].result = r;
        }
        for (int j = 0; j < 32; ++j)
            if ((mask >> j) & 1u) { w[j].arrived = false; w[j].released = true; }
        ++g_cta.progress;
    }
    while (!me->released) yield();
    me->released = false;
    return me->result;
}

void cta_barrier() {
    Fiber *me = g_cur;
    const uint64_t gen = g_cta.bar_gen;
    me->bar_gen = gen;
    me->at_bar = true;
    ++g_cta.bar_count;
    if (g_cta.bar_count + g_cta.n_exited == (int)g_cta.f.size()) {
        g_cta.bar_count = 0;
        ++g_cta.bar_gen;
        ++g_cta.progress;
    }
    while (g_cta.bar_gen == gen) yield();
    me->at_bar = false;
}

static void fiber_main() {
    g_cta.body(g_cta.body_arg);
    Fiber *me = g_cur;
    me->done = true;
    ++g_cta.n_exited;
    ++g_cta.progress;
    if (g_cta.bar_count > 0 && g_cta.bar_count + g_cta.n_exited == (int)g_cta.f.size()) {  // exited threads count as arrived
        g_cta.bar_count = 0;
        ++g_cta.bar_gen;
    }
    yield();
    die("resumed a finished fiber");
}

static std::vector<char *> g_stacks;

void run_grid(unsigned grid, unsigned block, size_t smem_bytes, void (*body)(void *), void *arg) {
    if (block == 0 || block > 1024) die("bad block size %u", block);
    if (smem_bytes > (size_t)MAX_SMEM) die("dynamic shared memory %zu exceeds %d", smem_bytes, MAX_SMEM);
    while (g_stacks.size() < block) {
        void *p = nullptr;
        if (posix_memalign(&p, 64, STACK_BYTES)) die("out of memory for fiber stacks");
        g_stacks.push_back((char *)p);
    }
    g_gridDim.x = grid;
    g_polls = 0;
    g_blockDim.x = block;
    for (unsigned b = 0; b < grid; ++b) {
        g_blockIdx.x = b;
        memset(sse::smem, 0xCD, MAX_SMEM);  // shared memory is uninitialised on the device
        Cta &c = g_cta;
        c.f.assign(block, Fiber());
        c.progress = 0;
        c.bar_gen = 0;
        c.bar_count = 0;
        c.n_exited = 0;
        c.body = body;
        c.body_arg = arg;
        for (unsigned t = 0; t < block; ++t) {
            Fiber &f = c.f[t];
            f.tid.x = t;
            f.lane = (int)(t & 31u);
            f.warp = (int)(t >> 5);
            f.stack = g_stacks[t];
            // initial frame: six callee-saved registers, then the entry address; after `ret` the stack pointer
            // is 8 mod 16, as at any function entry
            uintptr_t top = ((uintptr_t)f.stack + STACK_BYTES) & ~(uintptr_t)15;
            void **sp = (void **)(top - 64);
            for (int i = 0; i < 6; ++i) sp[i] = nullptr;
            sp[6] = (void *)&fiber_main;
            sp[7] = nullptr;
            f.sp = sp;
        }
        unsigned remaining = block;
        while (remaining) {
            const uint64_t before = c.progress;
            for (unsigned t = 0; t < block; ++t) {
                Fiber &f = c.f[t];
                if (f.done) continue;
                if (f.arrived && !f.released) continue;            // still waiting for its collective
                if (f.at_bar && f.bar_gen == c.bar_gen) continue;  // still waiting at __syncthreads
                g_cur = &f;
                emu_switch(&c.sched_sp, f.sp);
                if (f.done) --remaining;
            }
            g_cur = nullptr;
            if (remaining && c.progress == before) die("deadlock: no thread of block %u can make progress", b);
        }
    }
    check_redzones("after a kernel launch");
}

// ---- device memory with red zones ----------------------------------------------------------------------
static std::map<void *, size_t> g_allocs;

static bool zone_ok(const uint8_t *p) {
    for (size_t i = 0; i < REDZONE; ++i)
        if (p[i] != 0xEE) return false;
    return true;
}

void check_redzones(const char *when) {
    for (auto &kv : g_allocs) {
        const uint8_t *u = (const uint8_t *)kv.first;
        if (!zone_ok(u - REDZONE)) die("out-of-bounds write BELOW the %zu-byte device allocation %p detected %s", kv.second, kv.first, when);
        if (!zone_ok(u + kv.second)) die("out-of-bounds write ABOVE the %zu-byte device allocation %p detected %s", kv.second, kv.first, when);
    }
}

}  // namespace emu

cudaError_t cudaMalloc(void **p, size_t bytes) {
    void *raw = nullptr;
    if (posix_memalign(&raw, 256, bytes + 2 * emu::REDZONE + 256)) return cudaErrorMemoryAllocation;
    uint8_t *u = (uint8_t *)raw;
    memset(u, 0xEE, emu::REDZONE);
    const size_t padded = (bytes + 255) & ~(size_t)255;  // keep the upper zone right behind the block
    (void)padded;
    memset(u + emu::REDZONE, 0xA5, bytes);               // device memory is uninitialised
    memset(u + emu::REDZONE + bytes, 0xEE, emu::REDZONE);
    *p = u + emu::REDZONE;
    emu::g_allocs[*p] = bytes;
    return cudaSuccess;
}

cudaError_t cudaFree(void *p) {
    if (!p) return cudaSuccess;
    auto it = emu::g_allocs.find(p);
    if (it == emu::g_allocs.end()) emu::die("cudaFree of unknown pointer %p", p);
    emu::check_redzones("at cudaFree");
    emu::g_allocs.erase(it);
    free((uint8_t *)p - emu::REDZONE);
    return cudaSuccess;
}
