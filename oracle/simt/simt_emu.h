/*
 * simt_emu.h -- a small CPU SIMT emulator.  TEST INFRASTRUCTURE (lives under oracle/): it lets the CUDA kernel
 * SOURCE of dynamic-obstacle-avoidance-mpc_b200/csrc/ be compiled with g++ and executed on the host, thread block by
 * thread block, so that
 *   (1) a "mirror" build of the solver kernels exists whose every floating-point operation, reduction order and
 *       shuffle pattern is the GPU's (the kernels are written with explicit fma() and compiled with -fmad=false on the
 *       device / -ffp-contract=off here): tests demand BIT-IDENTICAL outputs, not tolerances;
 *   (2) kernels can be debugged in a container that has no GPU.
 * Only tests/, __graft_entry__.smoke() and bench.py's checker legs may load what is built from this header.
 *
 * Model: one host thread executes one thread block at a time; every CUDA thread is a coroutine with its own stack.
 * A thread runs until it reaches a warp/block collective (__shfl*_sync, __reduce_*_sync, __syncwarp, __syncthreads,
 * named barriers, mbarrier waits), deposits its operand and yields to the next runnable thread; the last arrival flips
 * the barrier generation and everybody picks up its result.  Memory is the host's: shared memory is a per-block buffer,
 * "global" memory plain arrays.  There is no modelling of races between collectives (the GPU run is the judge of
 * those; a missing barrier typically shows up as a bit difference between the two).
 */
#pragma once
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <functional>
#include <thread>
#include <vector>

#if !defined(__x86_64__)
#error "simt_emu.h: the coroutine switch is written for x86-64 (SysV ABI)"
#endif

#define SIMT_EMU 1
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
#define __constant__ static const
#define __shared__ static thread_local          /* static shared variables: one block at a time per host thread */
#define __launch_bounds__(...)
#define __maxnreg__(...)

namespace simt {

struct uint3_ { unsigned x, y, z; };

enum { MAX_THREADS = 1024, MAX_WARPS = 32, N_NAMED = 16, STACK_BYTES = 256 * 1024 };

struct Barrier {
    int arrived = 0;
    unsigned gen = 0;
    uint64_t slot[2][32];            // warp collectives: one 64-bit operand per lane, double-buffered by generation parity
};

struct Cta;
struct Lane {
    void* sp = nullptr;
    char* stack = nullptr;
    bool done = true;
    const unsigned* wait_gen_ptr = nullptr;      // blocked until *wait_gen_ptr != wait_gen
    unsigned wait_gen = 0;
    const volatile uint64_t* wait_word = nullptr;   // or until *wait_word != wait_word_val (mbarrier phase)
    uint64_t wait_word_val = 0;
    uint3_ tid{0, 0, 0};
    int lin = 0;
};

struct Cta {
    Lane lanes[MAX_THREADS];
    int nthreads = 0, live = 0, cur = 0;
    uint3_ bid{0, 0, 0}, bdim{1, 1, 1}, gdim{1, 1, 1};
    char* smem = nullptr;
    size_t smem_bytes = 0, smem_used = 0;
    Barrier warp_bar[MAX_WARPS];
    Barrier sub_bar[MAX_WARPS][4];               // partial-mask collectives: keyed by the lowest lane of the mask / 8
    Barrier cta_bar, named[N_NAMED];
    void* sched_sp = nullptr;
    const std::function<void()>* body = nullptr;
};

extern thread_local Cta* g_cta;
extern thread_local Lane* g_lane;

extern "C" void simt_switch(void** save_sp, void* load_sp);

#ifdef SIMT_EMU_IMPL
thread_local Cta* g_cta = nullptr;
thread_local Lane* g_lane = nullptr;
__asm__(
    ".text\n.globl simt_switch\n.type simt_switch,@function\nsimt_switch:\n"
    "  pushq %rbp\n  pushq %rbx\n  pushq %r12\n  pushq %r13\n  pushq %r14\n  pushq %r15\n"
    "  movq %rsp, (%rdi)\n  movq %rsi, %rsp\n"
    "  popq %r15\n  popq %r14\n  popq %r13\n  popq %r12\n  popq %rbx\n  popq %rbp\n  ret\n"
    ".size simt_switch,.-simt_switch\n");
#endif

inline bool runnable(const Lane& l)
{
    if (l.done) return false;
    if (l.wait_gen_ptr && *l.wait_gen_ptr == l.wait_gen) return false;
    if (l.wait_word && *l.wait_word == l.wait_word_val) return false;
    return true;
}

// give up the processor: run the next runnable thread of the block (or return to the scheduler when none is left)
inline void yield_lane()
{
    Cta* c = g_cta;
    Lane* me = g_lane;
    const int n = c->nthreads;
    int i = me->lin;
    for (int step = 0; step < n; step++) {
        i = i + 1 == n ? 0 : i + 1;
        Lane& l = c->lanes[i];
        if (runnable(l)) {
            if (&l == me) return;
            g_lane = &l;
            simt_switch(&me->sp, l.sp);
            return;                                // resumed: g_lane was set by whoever switched to us
        }
    }
    if (c->live == 0) {                            // last thread finished
        g_lane = nullptr;
        simt_switch(&me->sp, c->sched_sp);
        return;
    }
    fprintf(stderr, "simt_emu: deadlock in block (%u,%u,%u): %d live threads, none runnable (thread %d yielded last)\n",
            c->bid.x, c->bid.y, c->bid.z, c->live, me->lin);
    abort();
}

inline void block_on_gen(const unsigned* gen_ptr, unsigned gen)
{
    Lane* me = g_lane;
    me->wait_gen_ptr = gen_ptr; me->wait_gen = gen;
    yield_lane();
    me->wait_gen_ptr = nullptr;
}

// arrive at a counting barrier; returns after `expected` threads have arrived
inline void barrier_arrive_wait(Barrier& b, int expected)
{
    const unsigned g = b.gen;
    if (++b.arrived >= expected) { b.arrived = 0; b.gen = g + 1; return; }
    block_on_gen(&b.gen, g);
}

inline Barrier& warp_barrier(unsigned mask, int& expected)
{
    Cta* c = g_cta;
    const int w = g_lane->lin >> 5;
    if (mask == 0xffffffffu) {
        // a partially filled last warp: only the existing threads take part
        const int in_warp = c->nthreads - (w << 5);
        expected = in_warp < 32 ? in_warp : 32;
        return c->warp_bar[w];
    }
    expected = __builtin_popcount(mask);
    return c->sub_bar[w][(__builtin_ctz(mask) >> 3) & 3];
}

template <class T>
inline uint64_t to_bits(T v) { uint64_t u = 0; memcpy(&u, &v, sizeof(T)); return u; }
template <class T>
inline T from_bits(uint64_t u) { T v; memcpy(&v, &u, sizeof(T)); return v; }

// every lane of `mask` deposits v; afterwards slot[0..31] of the returned buffer holds all operands
template <class T>
inline const uint64_t* exchange(unsigned mask, T v)
{
    static_assert(sizeof(T) <= 8, "64-bit operands at most");
    int expected;
    Barrier& b = warp_barrier(mask, expected);
    const unsigned g = b.gen;
    uint64_t* buf = b.slot[g & 1];
    buf[g_lane->lin & 31] = to_bits(v);
    if (++b.arrived >= expected) { b.arrived = 0; b.gen = g + 1; }
    else block_on_gen(&b.gen, g);
    return buf;
}

void run_block(Cta& c, const std::function<void()>& body, uint3_ bid, uint3_ gdim, uint3_ bdim, size_t smem_bytes);

#ifdef SIMT_EMU_IMPL
static void lane_trampoline()
{
    Cta* c = g_cta;
    (*c->body)();
    Lane* me = g_lane;
    me->done = true;
    c->live--;
    // wake-up conditions of threads waiting at __syncthreads are NOT adjusted for exited threads (kernels here never
    // exit a thread ahead of a block barrier)
    yield_lane();
    fprintf(stderr, "simt_emu: finished thread resumed\n");
    abort();
}

void run_block(Cta& c, const std::function<void()>& body, uint3_ bid, uint3_ gdim, uint3_ bdim, size_t smem_bytes)
{
    const int n = (int)(bdim.x * bdim.y * bdim.z);
    if (n > MAX_THREADS) { fprintf(stderr, "simt_emu: block too large\n"); abort(); }
    c.nthreads = n; c.live = n; c.bid = bid; c.gdim = gdim; c.bdim = bdim; c.body = &body;
    if (c.smem_bytes < smem_bytes + 1024) {
        free(c.smem);
        c.smem_bytes = smem_bytes + 1024;
        if (posix_memalign((void**)&c.smem, 128, c.smem_bytes)) abort();
    }
    memset(c.smem, 0xA5, c.smem_bytes);            // uninitialised shared memory is garbage on the GPU, too (and the
    c.smem_used = smem_bytes;                      // bytes past the launch's size are a canary, checked after the block)
    for (auto& b : c.warp_bar) { b.arrived = 0; }
    for (auto& w : c.sub_bar) for (auto& b : w) b.arrived = 0;
    c.cta_bar.arrived = 0;
    for (auto& b : c.named) b.arrived = 0;
    for (int i = 0; i < n; i++) {
        Lane& l = c.lanes[i];
        if (!l.stack) { if (posix_memalign((void**)&l.stack, 64, STACK_BYTES)) abort(); }
        l.done = false; l.wait_gen_ptr = nullptr; l.wait_word = nullptr; l.lin = i;
        l.tid.x = i % bdim.x; l.tid.y = (i / bdim.x) % bdim.y; l.tid.z = i / (bdim.x * bdim.y);
        uintptr_t top = ((uintptr_t)(l.stack + STACK_BYTES)) & ~(uintptr_t)15;
        void** sp = (void**)top;
        *--sp = nullptr;                           // fake return address of the trampoline (keeps rsp = 8 mod 16 at entry)
        *--sp = (void*)&lane_trampoline;
        for (int r = 0; r < 6; r++) *--sp = nullptr;
        l.sp = (void*)sp;
    }
    g_cta = &c;
    g_lane = &c.lanes[0];
    simt_switch(&c.sched_sp, c.lanes[0].sp);
    g_cta = nullptr;
    for (size_t i = c.smem_used; i < c.smem_bytes; i++) {
        if ((unsigned char)c.smem[i] != 0xA5) {
            fprintf(stderr, "simt_emu: block (%u,%u,%u) wrote shared memory at byte %zu, past the %zu bytes of the launch\n", bid.x, bid.y, bid.z, i, c.smem_used);
            abort();
        }
    }
}
#endif

// run a grid: `host_threads` OS threads pull blocks from a counter (blocks of one grid are independent in CUDA, too)
inline void run_grid(unsigned grid_x, unsigned block_x, size_t smem_bytes, int host_threads, const std::function<void()>& body)
{
    std::atomic<unsigned> next{0};
    auto worker = [&]() {
        Cta* c = new Cta();
        for (;;) {
            const unsigned b = next.fetch_add(1);
            if (b >= grid_x) break;
            run_block(*c, body, uint3_{b, 0, 0}, uint3_{grid_x, 1, 1}, uint3_{block_x, 1, 1}, smem_bytes);
        }
        for (int i = 0; i < MAX_THREADS; i++) free(c->lanes[i].stack);
        free(c->smem);
        delete c;
    };
    if (host_threads <= 1) { worker(); return; }
    std::vector<std::thread> th;
    for (int t = 0; t < host_threads; t++) th.emplace_back(worker);
    for (auto& t : th) t.join();
}

}   // namespace simt

// ---------------------------------------------------------------- the CUDA device vocabulary the kernels use
#define threadIdx (simt::g_lane->tid)
#define blockIdx (simt::g_cta->bid)
#define blockDim (simt::g_cta->bdim)
#define gridDim (simt::g_cta->gdim)
#define SIMT_DYNAMIC_SMEM(type, name) type* const name = reinterpret_cast<type*>(simt::g_cta->smem)

template <class T>
inline T __shfl_sync(unsigned mask, T v, int src, int width = 32)
{
    const int lane = simt::g_lane->lin & 31;
    const uint64_t* buf = simt::exchange(mask, v);
    const int base = lane & ~(width - 1);
    return simt::from_bits<T>(buf[base + (src & (width - 1))]);
}
template <class T>
inline T __shfl_xor_sync(unsigned mask, T v, int lanemask, int width = 32)
{
    const int lane = simt::g_lane->lin & 31;
    const uint64_t* buf = simt::exchange(mask, v);
    const int src = lane ^ lanemask;
    const int base = lane & ~(width - 1);
    return simt::from_bits<T>(buf[(src >= base && src < base + width) ? src : lane]);
}
template <class T>
inline T __shfl_down_sync(unsigned mask, T v, unsigned delta, int width = 32)
{
    const int lane = simt::g_lane->lin & 31;
    const uint64_t* buf = simt::exchange(mask, v);
    const int base = lane & ~(width - 1);
    const int src = lane + (int)delta;
    return simt::from_bits<T>(buf[src < base + width ? src : lane]);
}
template <class T>
inline T __shfl_up_sync(unsigned mask, T v, unsigned delta, int width = 32)
{
    const int lane = simt::g_lane->lin & 31;
    const uint64_t* buf = simt::exchange(mask, v);
    const int base = lane & ~(width - 1);
    const int src = lane - (int)delta;
    return simt::from_bits<T>(buf[src >= base ? src : lane]);
}
inline unsigned __reduce_max_sync(unsigned mask, unsigned v)
{
    const uint64_t* buf = simt::exchange(mask, v);
    unsigned r = 0;
    for (int i = 0; i < 32; i++) if ((mask >> i) & 1u) { const unsigned x = (unsigned)buf[i]; r = x > r ? x : r; }
    return r;
}
inline unsigned __reduce_min_sync(unsigned mask, unsigned v)
{
    const uint64_t* buf = simt::exchange(mask, v);
    unsigned r = 0xffffffffu;
    for (int i = 0; i < 32; i++) if ((mask >> i) & 1u) { const unsigned x = (unsigned)buf[i]; r = x < r ? x : r; }
    return r;
}
inline unsigned __ballot_sync(unsigned mask, int pred)
{
    const uint64_t* buf = simt::exchange(mask, (unsigned)(pred != 0));
    unsigned r = 0;
    for (int i = 0; i < 32; i++) if (((mask >> i) & 1u) && buf[i]) r |= 1u << i;
    return r;
}
inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
inline int __all_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) == mask; }
inline void __syncwarp(unsigned mask = 0xffffffffu) { (void)simt::exchange(mask, 0u); }
inline void __syncthreads() { simt::barrier_arrive_wait(simt::g_cta->cta_bar, simt::g_cta->nthreads); }
// bar.sync id, nthreads (the kernels reach it through b2_named_barrier())
inline void simt_named_barrier(int id, int nthreads) { simt::barrier_arrive_wait(simt::g_cta->named[id & 15], nthreads); }

inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }

inline int __double2hiint(double x) { uint64_t u; memcpy(&u, &x, 8); return (int)(u >> 32); }
inline int __double2loint(double x) { uint64_t u; memcpy(&u, &x, 8); return (int)(u & 0xffffffffu); }
inline double __hiloint2double(int hi, int lo)
{
    const uint64_t u = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo;
    double x; memcpy(&x, &u, 8); return x;
}
inline double __longlong_as_double(long long v) { double x; memcpy(&x, &v, 8); return x; }
inline long long __double_as_longlong(double x) { long long v; memcpy(&v, &x, 8); return v; }
inline double __dmul_rn(double a, double b) { return a * b; }
inline double __dadd_rn(double a, double b) { return a + b; }
inline double __fma_rn(double a, double b, double c) { return fma(a, b, c); }
inline double rsqrt(double x) { return 1.0 / sqrt(x); }     // (the kernels themselves never call rsqrt: not correctly rounded on the GPU)
