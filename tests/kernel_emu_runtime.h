// TEST INFRASTRUCTURE (see tests/_kernel_emu.py): a host stand-in for the CUDA execution model, good enough to run the
// library's CUDA-core kernels from their own source text on a machine without a GPU.
//
//   * one block at a time; every CUDA thread of the block is a fiber (ucontext) with its own stack, scheduled round-robin
//     on ONE OS thread, so "atomics" are plain read-modify-writes and the run is deterministic;
//   * __syncthreads / __syncwarp / __shfl_*_sync / __ballot_sync / __any_sync / __all_sync are rendezvous points: a fiber
//     deposits its value, yields until its block / warp has arrived, then reads the snapshot;
//   * polling loops on volatile shared memory (the CTC wavefront) make progress because the host versions of the polling
//     loads yield;
//   * `__shared__` variables become function-level statics (blocks run one after the other), dynamic shared memory is one
//     buffer per launch, poisoned with NaN bytes before every block;
//   * a block in which every live fiber waits and none can be released is reported as a deadlock (launch returns 1).
//
// NOT modelled: memory ordering, bank conflicts, alignment faults, occupancy, anything asynchronous (cp.async is executed
// eagerly by the host replacement the test provides), tensor cores / TMA / TMEM.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <ucontext.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <type_traits>
#include <vector>

namespace emu {

struct Fiber {
  ucontext_t ctx;
  char* stack = nullptr;
  bool done = false;
  uint3 tid;
  int warp, lane;
};
struct Warp {
  uint64_t buf[32], snap[32];
  int arrived = 0, nlanes = 0;
  unsigned long gen = 0;
  unsigned live_mask = 0;       // lanes that have not exited
  unsigned snap_mask = 0;       // live_mask when the last collective completed (a lane may exit before its peers read the snapshot)
};

static std::vector<Fiber> fibers;
static std::vector<Warp> warps;
static ucontext_t sched_ctx;
static std::function<void()> body;
static int cur = -1, live = 0, bar_arrived = 0;
static unsigned long bar_gen = 0;
static long long n_yields = 0, n_events = 0;       // events: releases, exits, polling yields (anything that can unblock someone)
static unsigned char* dyn_smem = nullptr;
static uint3 tid, bid;
static dim3 bdim, gdim;
static const size_t kStack = 256 * 1024;
static long long yield_budget = 50000000LL;      // per block; emu_set_yield_budget() changes it

static inline void yield_wait() {
  ++n_yields;
  swapcontext(&fibers[cur].ctx, &sched_ctx);
}
static inline void yield_poll() {                  // a polling loop: counts as potential progress
  ++n_events;
  yield_wait();
}
static void release_block_if_complete() {
  if (live > 0 && bar_arrived == live) {
    bar_arrived = 0;
    ++bar_gen;
    ++n_events;
  }
}
static void release_warp_if_complete(Warp& w) {
  if (w.nlanes > 0 && w.arrived == w.nlanes) {
    w.arrived = 0;
    std::memcpy(w.snap, w.buf, sizeof(w.snap));
    w.snap_mask = w.live_mask;
    ++w.gen;
    ++n_events;
  }
}
static void trampoline() {
  body();
  Fiber& f = fibers[cur];
  f.done = true;
  --live;
  ++n_events;
  Warp& w = warps[f.warp];
  --w.nlanes;
  w.live_mask &= ~(1u << f.lane);
  release_block_if_complete();                      // exited threads no longer count for barriers (as on the hardware)
  release_warp_if_complete(w);
  swapcontext(&f.ctx, &sched_ctx);
}

static inline void syncthreads() {
  ++bar_arrived;
  const unsigned long g = bar_gen;
  release_block_if_complete();
  while (bar_gen == g) yield_wait();
}
// deposit `bits`, wait for the warp, return the snapshot of all lanes' deposits
static inline const uint64_t* warp_exchange(uint64_t bits) {
  Fiber& f = fibers[cur];
  Warp& w = warps[f.warp];
  w.buf[f.lane] = bits;
  ++w.arrived;
  const unsigned long g = w.gen;
  release_warp_if_complete(w);
  while (w.gen == g) yield_wait();
  return w.snap;
}
template <typename T>
static inline uint64_t to_bits(T v) {
  static_assert(sizeof(T) <= 8, "shuffle operand wider than 64 bits");
  uint64_t b = 0;
  std::memcpy(&b, &v, sizeof(T));
  return b;
}
template <typename T>
static inline T from_bits(uint64_t b) {
  T v;
  std::memcpy(&v, &b, sizeof(T));
  return v;
}
template <typename T>
static inline T shfl_from(T v, int src) {
  const int lane = fibers[cur].lane, warp = fibers[cur].warp;
  const uint64_t* snap = warp_exchange(to_bits(v));
  if (src < 0 || src > 31 || !((warps[warp].snap_mask >> src) & 1u)) src = lane;     // inactive source: undefined on hardware
  return from_bits<T>(snap[src]);
}

// returns 0 on success, 1 on deadlock, 2 when the yield budget is exhausted (a polling loop that never ends); launch() adds
// 3: a block wrote past the end of its dynamic shared memory (the 256 guard bytes behind it changed)
static int run_block(unsigned nthreads) {
  fibers.resize(nthreads);
  warps.assign((nthreads + 31) / 32, Warp());
  live = (int)nthreads;
  bar_arrived = 0;
  for (unsigned i = 0; i < nthreads; ++i) {
    Fiber& f = fibers[i];
    if (!f.stack) f.stack = (char*)std::malloc(kStack);
    f.done = false;
    f.tid = make_uint3(i % bdim.x, (i / bdim.x) % bdim.y, i / (bdim.x * bdim.y));
    f.warp = i / 32;
    f.lane = i % 32;
    warps[f.warp].nlanes++;
    warps[f.warp].live_mask |= 1u << f.lane;
    getcontext(&f.ctx);
    f.ctx.uc_stack.ss_sp = f.stack;
    f.ctx.uc_stack.ss_size = kStack;
    f.ctx.uc_link = &sched_ctx;
    makecontext(&f.ctx, trampoline, 0);
  }
  const long long budget = n_yields + yield_budget;
  while (live > 0) {
    const long long ev0 = n_events;
    for (unsigned i = 0; i < nthreads; ++i) {
      if (fibers[i].done) continue;
      cur = (int)i;
      tid = fibers[i].tid;
      swapcontext(&sched_ctx, &fibers[i].ctx);
    }
    if (live > 0 && n_events == ev0) return 1;
    if (n_yields > budget) return 2;
  }
  return 0;
}

static int launch(int gx, int gy, int gz, int bx, int by, int bz, size_t smem_bytes, const std::function<void()>& fn) {
  gdim = dim3(gx, gy, gz);
  bdim = dim3(bx, by, bz);
  body = fn;
  const size_t n = smem_bytes + 256;
  unsigned char* raw = (unsigned char*)std::malloc(n + 128);
  dyn_smem = (unsigned char*)(((uintptr_t)raw + 127) & ~(uintptr_t)127);
  int rc = 0;
  for (int z = 0; z < gz && !rc; ++z)
    for (int y = 0; y < gy && !rc; ++y)
      for (int x = 0; x < gx && !rc; ++x) {
        std::memset(dyn_smem, 0xFF, n);              // NaN poison: a read of unwritten shared memory shows up in the result
        bid = make_uint3(x, y, z);
        rc = run_block((unsigned)(bx * by * bz));
        for (size_t i = smem_bytes; i < n && !rc; ++i)
          if (dyn_smem[i] != 0xFF) rc = 3;           // a write past the dynamic shared memory the launch asked for
      }
  std::free(raw);
  dyn_smem = nullptr;
  return rc;
}

}  // namespace emu

extern "C" void emu_set_yield_budget(long long n) { emu::yield_budget = n; }

// ------------------------------------------------------------------------------------------------ CUDA surface
#undef __shared__
#define __shared__ static
#define __launch_bounds__(...)
#define threadIdx emu::tid
#define blockIdx emu::bid
#define blockDim emu::bdim
#define gridDim emu::gdim
#define warpSize 32

static inline void __syncthreads() { emu::syncthreads(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { emu::warp_exchange(0); }
static inline void __threadfence() {}
static inline void __threadfence_block() {}
static inline void __threadfence_system() {}
static inline void __nanosleep(unsigned) { emu::yield_poll(); }
static inline void __trap() {
  std::fprintf(stderr, "emu: __trap()\n");
  std::abort();
}
static inline unsigned __activemask() { return emu::warps[emu::fibers[emu::cur].warp].live_mask; }

template <typename T>
static inline T __shfl_sync(unsigned, T v, int src, int width = 32) {
  const int lane = emu::fibers[emu::cur].lane;
  return emu::shfl_from(v, (lane / width) * width + (src % width));
}
template <typename T>
static inline T __shfl_up_sync(unsigned, T v, unsigned delta, int width = 32) {
  const int lane = emu::fibers[emu::cur].lane;
  return emu::shfl_from(v, (lane % width) < (int)delta ? lane : lane - (int)delta);
}
template <typename T>
static inline T __shfl_down_sync(unsigned, T v, unsigned delta, int width = 32) {
  const int lane = emu::fibers[emu::cur].lane;
  return emu::shfl_from(v, (lane % width) + (int)delta >= width ? lane : lane + (int)delta);
}
template <typename T>
static inline T __shfl_xor_sync(unsigned, T v, int lane_mask, int width = 32) {
  const int lane = emu::fibers[emu::cur].lane;
  const int src = lane ^ lane_mask;
  return emu::shfl_from(v, (src / width) != (lane / width) ? lane : src);
}
static inline unsigned __ballot_sync(unsigned mask, int pred) {
  const int warp = emu::fibers[emu::cur].warp;
  const uint64_t* snap = emu::warp_exchange(pred ? 1 : 0);
  unsigned r = 0;
  for (int i = 0; i < 32; ++i)
    if (((emu::warps[warp].snap_mask & mask) >> i) & 1u) r |= (snap[i] ? 1u : 0u) << i;
  return r;
}
static inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
static inline int __all_sync(unsigned mask, int pred) {
  const unsigned votes = __ballot_sync(mask, pred);
  const unsigned live = emu::warps[emu::fibers[emu::cur].warp].snap_mask & mask;
  return (votes & live) == live;
}

template <typename T>
static inline T __ldg(const T* p) { return *p; }

// single OS thread, cooperative scheduling: a read-modify-write between two yields is atomic
template <typename T, typename U>
static inline T atomicAdd(T* p, U v) { T old = *p; *p = (T)(old + (T)v); return old; }
template <typename T, typename U>
static inline T atomicSub(T* p, U v) { T old = *p; *p = (T)(old - (T)v); return old; }
template <typename T, typename U>
static inline T atomicMax(T* p, U v) { T old = *p; if ((T)v > old) *p = (T)v; return old; }
template <typename T, typename U>
static inline T atomicMin(T* p, U v) { T old = *p; if ((T)v < old) *p = (T)v; return old; }
template <typename T, typename U>
static inline T atomicExch(T* p, U v) { T old = *p; *p = (T)v; return old; }
template <typename T, typename U>
static inline T atomicOr(T* p, U v) { T old = *p; *p = (T)(old | (T)v); return old; }
template <typename T, typename U>
static inline T atomicAnd(T* p, U v) { T old = *p; *p = (T)(old & (T)v); return old; }
template <typename T, typename U, typename V>
static inline T atomicCAS(T* p, U cmp, V v) { T old = *p; if (old == (T)cmp) *p = (T)v; return old; }

// min / max with CUDA's mixed-type overloads
template <typename A, typename B>
static inline typename std::common_type<A, B>::type min(A a, B b) {
  typedef typename std::common_type<A, B>::type C;
  return (C)a < (C)b ? (C)a : (C)b;
}
template <typename A, typename B>
static inline typename std::common_type<A, B>::type max(A a, B b) {
  typedef typename std::common_type<A, B>::type C;
  return (C)a > (C)b ? (C)a : (C)b;
}

// integer / bit intrinsics
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
static inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline unsigned __brev(unsigned x) {
  unsigned r = 0;
  for (int i = 0; i < 32; ++i) r |= ((x >> i) & 1u) << (31 - i);
  return r;
}
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((uint64_t)a * b) >> 32); }
static inline int __mulhi(int a, int b) { return (int)(((int64_t)a * b) >> 32); }
static inline int __float_as_int(float f) { return emu::from_bits<int>(emu::to_bits(f)); }
static inline unsigned __float_as_uint(float f) { return emu::from_bits<unsigned>(emu::to_bits(f)); }
static inline float __int_as_float(int i) { return emu::from_bits<float>(emu::to_bits(i)); }
static inline float __uint_as_float(unsigned i) { return emu::from_bits<float>(emu::to_bits(i)); }
static inline unsigned __float2uint_rn(float f) { return f <= 0.f ? 0u : (unsigned)std::nearbyintf(f); }
static inline int __float2int_rn(float f) { return (int)std::nearbyintf(f); }
static inline int __float2int_rd(float f) { return (int)std::floor(f); }
static inline float __int2float_rn(int i) { return (float)i; }
static inline float __uint2float_rn(unsigned i) { return (float)i; }

// float intrinsics (the fast-math ones are exact here; tolerances in the tests are the GPU tests' own).  glibc declares some of
// these names itself (__expf, __logf, ...), hence macros onto emu_ functions.
static inline float emu_expf(float x) { return std::exp(x); }
static inline float emu_exp2f(float x) { return std::exp2(x); }
static inline float emu_logf(float x) { return std::log(x); }
static inline float emu_log2f(float x) { return std::log2(x); }
static inline float emu_sinf(float x) { return std::sin(x); }
static inline float emu_cosf(float x) { return std::cos(x); }
static inline void emu_sincosf(float x, float* s, float* c) { *s = std::sin(x); *c = std::cos(x); }
static inline void emu_sincospif(float x, float* s, float* c) {
  *s = (float)std::sin((double)x * 3.14159265358979323846);
  *c = (float)std::cos((double)x * 3.14159265358979323846);
}
static inline float emu_rsqrtf(float x) { return 1.f / std::sqrt(x); }
static inline float emu_saturatef(float x) { return x < 0.f ? 0.f : (x > 1.f ? 1.f : x); }
#define __expf(x) emu_expf(x)
#define __exp2f(x) emu_exp2f(x)
#define __logf(x) emu_logf(x)
#define __log2f(x) emu_log2f(x)
#define __sinf(x) emu_sinf(x)
#define __cosf(x) emu_cosf(x)
#define __sincosf(x, s, c) emu_sincosf(x, s, c)
#define sincospif(x, s, c) emu_sincospif(x, s, c)
#define rsqrtf(x) emu_rsqrtf(x)
#define __saturatef(x) emu_saturatef(x)
#define __fdividef(a, b) ((a) / (b))
#define __frcp_rn(x) (1.f / (x))
#define __fmaf_rn(a, b, c) std::fma((float)(a), (float)(b), (float)(c))
#define __fmul_rn(a, b) ((a) * (b))
#define __fadd_rn(a, b) ((a) + (b))
#define __fsub_rn(a, b) ((a) - (b))
#define __fdiv_rn(a, b) ((a) / (b))
#define __fsqrt_rn(x) std::sqrt((float)(x))

// ------------------------------------------------------------------------------------------------ library host surface
// what the kernels' translation units expect from common.cuh / runtime.cu (host helpers that sit beside the kernels)
#include "../include/w2l_sm100.h"
namespace w2l {
static inline void set_error(const char*, ...) {}
static inline int num_sms() { return 148; }
static inline int gemm_sms() { return 148; }
static inline int after_launch(const char*) { return 0; }
static inline int check_cuda(cudaError_t, const char*) { return 0; }
#define W2L_REQUIRE(cond, ...) \
  do {                         \
    if (!(cond)) return W2L_ERR_INVALID_ARGUMENT; \
  } while (0)
#define W2L_CUDA(expr) \
  do {                 \
    (void)0;           \
  } while (0)
}  // namespace w2l
