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

struct NamedBar {
  int arrived = 0;
  unsigned long gen = 0;
};
static NamedBar named[16];
static inline void named_barrier(int id, int count) {      // bar.sync id, count
  NamedBar& b = named[id & 15];
  const unsigned long g = b.gen;
  if (++b.arrived == count) {
    b.arrived = 0;
    ++b.gen;
    ++n_events;
  }
  while (b.gen == g) yield_wait();
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
  for (auto& nb : named) nb.arrived = 0;
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
  // W2L_EMU_SCHEDULE=reverse | random[:seed]: the order in which the fibers of a block are resumed in every round.  A kernel whose
  // result depends on it has a race (a missing barrier, or code that counts on a warp running in lockstep): racecheck, on the host.
  static int sched_mode = -1;
  static unsigned long long rng = 0x9E3779B97F4A7C15ull;
  if (sched_mode < 0) {
    const char* e = std::getenv("W2L_EMU_SCHEDULE");
    sched_mode = !e ? 0 : (std::strncmp(e, "reverse", 7) == 0 ? 1 : (std::strncmp(e, "random", 6) == 0 ? 2 : 0));
    if (sched_mode == 2 && e[6] == ':') rng ^= std::strtoull(e + 7, nullptr, 10) * 0xD1B54A32D192ED03ull;
  }
  std::vector<unsigned> order(nthreads);
  for (unsigned i = 0; i < nthreads; ++i) order[i] = sched_mode == 1 ? nthreads - 1 - i : i;
  while (live > 0) {
    const long long ev0 = n_events;
    if (sched_mode == 2)
      for (unsigned i = nthreads - 1; i > 0; --i) {            // Fisher-Yates with a xorshift generator
        rng ^= rng << 13;
        rng ^= rng >> 7;
        rng ^= rng << 17;
        std::swap(order[i], order[rng % (i + 1)]);
      }
    for (unsigned k = 0; k < nthreads; ++k) {
      const unsigned i = order[k];
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
  unsigned char* raw = (unsigned char*)std::malloc(n + 1024);
  dyn_smem = (unsigned char*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
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
#include <cuda.h>
#include <cstdarg>
#include "../include/w2l_sm100.h"

// CUDA runtime calls made by the C wrappers: memsets are performed, attribute settings succeed
#define cudaMemsetAsync(p, v, n, st) (std::memset((p), (v), (n)), cudaSuccess)
#define cudaFuncSetAttribute(...) cudaSuccess
#define cudaPeekAtLastError() cudaSuccess
#define cudaGetLastError() cudaSuccess

namespace w2l {
static char g_err[512] = "";
static long long g_launches = 0;
static inline void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
static int g_sm_budget = 0;                                  // w2l_set_sm_budget stand-in (emu_set_sm_budget)
static inline int num_sms() { return 148; }
static inline int gemm_sms() { return (g_sm_budget > 0 && g_sm_budget < 148) ? g_sm_budget : 148; }
static inline int after_launch(const char*) {
  ++g_launches;
  return 0;
}
static inline int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return W2L_OK;
  set_error("CUDA error %d at %s", (int)e, what);
  return W2L_ERR_CUDA;
}
#define W2L_REQUIRE(cond, ...)                 \
  do {                                         \
    if (!(cond)) {                             \
      w2l::set_error(__VA_ARGS__);             \
      return W2L_ERR_INVALID_ARGUMENT;         \
    }                                          \
  } while (0)
#define W2L_CUDA(expr)                                     \
  do {                                                     \
    int _rc = w2l::check_cuda((expr), #expr);              \
    if (_rc != W2L_OK) return _rc;                         \
  } while (0)

// ---- TMA: the tensor map is opaque to the kernels; here its 128 bytes hold what make_tensor_map was given
struct EmuTensorMap {
  const unsigned char* base;
  int32_t elem_bytes, rank, swizzle128;
  uint64_t dims[4];
  uint64_t strides[3];          // bytes, for dims 1..
  uint32_t box[4];
};
static_assert(sizeof(EmuTensorMap) <= sizeof(CUtensorMap), "emulated tensor map must fit the opaque CUtensorMap");
static inline int make_tensor_map(CUtensorMap* map, const void* base, int elem_bytes, int rank, const uint64_t* dims,
                                  const uint64_t* strides_bytes, const uint32_t* box, bool swizzle128) {
  // the documented constraints of cuTensorMapEncodeTiled that the library's call sites must meet
  W2L_REQUIRE(rank >= 1 && rank <= 4 && ((uintptr_t)base & 15) == 0, "tensor map: rank %d / base alignment", rank);
  for (int i = 0; i + 1 < rank; ++i) W2L_REQUIRE(strides_bytes[i] % 16 == 0, "tensor map: stride %d = %llu not a multiple of 16", i, (unsigned long long)strides_bytes[i]);
  for (int i = 0; i < rank; ++i) W2L_REQUIRE(box[i] >= 1 && box[i] <= 256 && dims[i] >= 1, "tensor map: box/dim %d", i);
  W2L_REQUIRE(!swizzle128 || box[0] * (uint32_t)elem_bytes <= 128, "tensor map: 128B swizzle needs an inner box of at most 128 bytes");
  EmuTensorMap m;
  std::memset(&m, 0, sizeof(m));
  m.base = (const unsigned char*)base;
  m.elem_bytes = elem_bytes;
  m.rank = rank;
  m.swizzle128 = swizzle128;
  for (int i = 0; i < 4; ++i) {
    m.dims[i] = i < rank ? dims[i] : 1;
    m.box[i] = i < rank ? box[i] : 1;
    if (i > 0) m.strides[i - 1] = i < rank ? strides_bytes[i - 1] : 0;
  }
  std::memset(map, 0, sizeof(*map));
  std::memcpy(map, &m, sizeof(m));
  return W2L_OK;
}
}  // namespace w2l

// ------------------------------------------------------------------------------------------------ sm_100 device surface
// Functional stand-ins for the inline-PTX wrappers of common.cuh (mbarrier, TMA, tcgen05 / TMEM).  Everything is synchronous:
// a TMA load lands before the call returns, an MMA has completed when it is issued, so tcgen05.commit is a plain arrive.
#define __grid_constant__
template <typename T>
static inline T __ldcg(const T* p) { return *p; }

namespace w2l {
// shared-memory "addresses" are offsets into the launch's dynamic shared memory (1024-byte aligned, like the hardware's window)
static inline uint32_t smem_u32(const void* p) { return (uint32_t)((const unsigned char*)p - emu::dyn_smem); }
static inline unsigned char* smem_ptr(uint32_t a) { return emu::dyn_smem + a; }
// 128-byte swizzle: address bits [4,7) ^= bits [7,10) of the ABSOLUTE shared-memory address (TMA writes and tcgen05 reads agree on it)
static inline uint32_t swz128(uint32_t a) { return a ^ (((a >> 7) & 7u) << 4); }
static inline uint32_t elect_one_sync() { return emu::fibers[emu::cur].lane == 0; }

struct EmuMbar {                 // mbarrier.shared::cta.b64
  uint32_t phase : 1, expected : 15, pending : 16;
  int32_t tx;
};
static_assert(sizeof(EmuMbar) == 8, "mbarrier is a 64-bit object");
static inline void mbar_check(EmuMbar* b) {
  if (b->pending == 0 && b->tx == 0) {
    b->phase ^= 1u;
    b->pending = b->expected;
    ++emu::n_events;
  }
}
static inline void mbar_init(uint64_t* bar, uint32_t count) {
  EmuMbar* b = (EmuMbar*)bar;
  b->phase = 0;
  b->expected = count;
  b->pending = count;
  b->tx = 0;
}
static inline void fence_barrier_init() {}
static inline void fence_proxy_async() {}
static inline void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {      // mbarrier.arrive.expect_tx
  EmuMbar* b = (EmuMbar*)bar;
  b->tx += (int32_t)bytes;
  b->pending -= 1;
  mbar_check(b);
}
static inline void mbar_arrive(uint64_t* bar) {
  EmuMbar* b = (EmuMbar*)bar;
  if (b->pending == 0) {
    std::fprintf(stderr, "emu: mbarrier over-arrival\n");
    std::abort();
  }
  b->pending -= 1;
  mbar_check(b);
}
static inline void mbar_complete_tx(uint64_t* bar, uint32_t bytes) {
  EmuMbar* b = (EmuMbar*)bar;
  b->tx -= (int32_t)bytes;
  mbar_check(b);
}
static inline bool mbar_try_wait(uint64_t* bar, uint32_t parity) { return ((EmuMbar*)bar)->phase != (parity & 1u); }
static inline void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) emu::yield_poll();
}

// cp.async.bulk.tensor.{3,4}d: box [b3][b2][b1][b0] lands densely (inner row = box0 elements), out-of-range elements as zero,
// the whole box is credited to the barrier
static inline void tma_load_nd(void* smem_dst, const CUtensorMap* mp, uint64_t* bar, const int (&c)[4]) {
  EmuTensorMap m;
  std::memcpy(&m, mp, sizeof(m));
  const uint32_t dst = smem_u32(smem_dst);
  const uint32_t row_bytes = m.box[0] * (uint32_t)m.elem_bytes;
  if (((long long)c[0] * m.elem_bytes) & 15) {        // seen on B200 (round 2, call 12): "illegal instruction" from such a load
    std::fprintf(stderr, "emu: TMA load with an innermost coordinate (%d x %d bytes) that is not 16-byte aligned\n", c[0], m.elem_bytes);
    std::abort();
  }
  uint32_t lin = 0;
  for (uint32_t i3 = 0; i3 < m.box[3]; ++i3)
    for (uint32_t i2 = 0; i2 < m.box[2]; ++i2)
      for (uint32_t i1 = 0; i1 < m.box[1]; ++i1) {
        const long long g3 = (long long)c[3] + i3, g2 = (long long)c[2] + i2, g1 = (long long)c[1] + i1;
        const bool row_in = g3 >= 0 && g3 < (long long)m.dims[3] && g2 >= 0 && g2 < (long long)m.dims[2] && g1 >= 0 && g1 < (long long)m.dims[1];
        const unsigned char* src = m.base + g1 * (long long)m.strides[0] + g2 * (long long)m.strides[1] + g3 * (long long)m.strides[2];
        for (uint32_t i0 = 0; i0 < m.box[0]; ++i0) {
          const long long g0 = (long long)c[0] + i0;
          const uint32_t a = dst + lin + i0 * (uint32_t)m.elem_bytes;
          unsigned char* d = smem_ptr(m.swizzle128 ? swz128(a) : a);
          if (row_in && g0 >= 0 && g0 < (long long)m.dims[0]) std::memcpy(d, src + g0 * m.elem_bytes, m.elem_bytes);
          else std::memset(d, 0, m.elem_bytes);
        }
        lin += row_bytes;
      }
  mbar_complete_tx(bar, lin);
}
static inline void tma_prefetch_desc(const CUtensorMap*) {}
static inline void tma_load_3d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  const int c[4] = {c0, c1, c2, 0};
  tma_load_nd(smem_dst, m, bar, c);
}
static inline void tma_load_4d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
  const int c[4] = {c0, c1, c2, c3};
  tma_load_nd(smem_dst, m, bar, c);
}

// TMEM: 128 lanes x 512 32-bit columns per SM; addresses are (lane << 16) | column
static uint32_t emu_tmem[128][512];
static inline void tmem_alloc(uint32_t* smem_result, uint32_t) { *smem_result = 0; }
static inline void tmem_relinquish() {}
static inline void tmem_dealloc(uint32_t, uint32_t) {}
static inline void tc_fence_before() {}
static inline void tc_fence_after() {}
static inline void tmem_ld_wait() {}
static inline void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {      // tcgen05.ld.32x32b.x32: thread i reads lane base+i
  const uint32_t lane = (taddr >> 16) + (uint32_t)emu::fibers[emu::cur].lane, col = taddr & 0xFFFFu;
  for (int i = 0; i < 32; ++i) r[i] = emu_tmem[lane][col + i];
}
static inline void umma_commit(uint64_t* bar) { mbar_arrive(bar); }

// one operand element of a 128B-swizzled shared-memory matrix descriptor (sm_100 "version 1", see common.cuh make_smem_desc):
//   K-major : row r, contraction index k (0..15) at start + (r / 8) * SBO + (r % 8) * 128 + k * 2
//   MN-major: mn index, contraction index k at start + (mn / 64) * LBO + (k / 8) * SBO + (k % 8) * 128 + (mn % 64) * 2
static inline float umma_operand(uint64_t desc, bool mn_major, int r, int k) {
  const uint32_t start = (uint32_t)(desc & 0x3FFF) << 4, lbo = (uint32_t)((desc >> 16) & 0x3FFF) << 4, sbo = (uint32_t)((desc >> 32) & 0x3FFF) << 4;
  const uint32_t a = mn_major ? start + (uint32_t)(r / 64) * lbo + (uint32_t)(k / 8) * sbo + (uint32_t)(k % 8) * 128u + (uint32_t)(r % 64) * 2u
                              : start + (uint32_t)(r / 8) * sbo + (uint32_t)(r % 8) * 128u + (uint32_t)k * 2u;
  uint16_t bits;
  std::memcpy(&bits, smem_ptr(swz128(a)), 2);
  return emu::from_bits<float>((uint64_t)bits << 16);
}
// tcgen05.mma.cta_group::1.kind::f16, bf16 x bf16 -> fp32: D[128 x N] (+)= A[128 x 16] * B[N x 16]^T
static inline void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  const int N = (int)((idesc >> 17) & 0x3F) << 3, M = (int)((idesc >> 24) & 0x1F) << 4;
  const bool a_mn = (idesc >> 15) & 1u, b_mn = (idesc >> 16) & 1u;
  if (M != 128 || N < 16 || N > 256 || ((a_desc >> 61) & 7) != 2 || ((b_desc >> 61) & 7) != 2) {
    std::fprintf(stderr, "emu: unsupported tcgen05.mma shape / swizzle (M=%d N=%d)\n", M, N);
    std::abort();
  }
  static float A[128][16], B[256][16];
  for (int m = 0; m < M; ++m)
    for (int k = 0; k < 16; ++k) A[m][k] = umma_operand(a_desc, a_mn, m, k);
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < 16; ++k) B[n][k] = umma_operand(b_desc, b_mn, n, k);
  const uint32_t lane0 = d_tmem >> 16, col0 = d_tmem & 0xFFFFu;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      float acc = accumulate ? emu::from_bits<float>(emu_tmem[lane0 + m][col0 + n]) : 0.f;
      for (int k = 0; k < 16; ++k) acc += A[m][k] * B[n][k];
      emu_tmem[lane0 + m][col0 + n] = (uint32_t)emu::to_bits(acc);
    }
}
// tcgen05.mma.cta_group::1.kind::tf32: fp32 operands read with a 10-bit mantissa (low 13 bits dropped), K = 8 per instruction;
// K-major only (the library transposes the operands of the fp32-faithful weight gradient instead of using MN-major tf32)
static inline float umma_operand_tf32(uint64_t desc, int r, int k) {
  const uint32_t start = (uint32_t)(desc & 0x3FFF) << 4, sbo = (uint32_t)((desc >> 32) & 0x3FFF) << 4;
  const uint32_t a = start + (uint32_t)(r / 8) * sbo + (uint32_t)(r % 8) * 128u + (uint32_t)k * 4u;
  uint32_t bits;
  std::memcpy(&bits, smem_ptr(swz128(a)), 4);
  return emu::from_bits<float>((uint64_t)(bits & 0xFFFFE000u));
}
static inline void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  const int N = (int)((idesc >> 17) & 0x3F) << 3, M = (int)((idesc >> 24) & 0x1F) << 4;
  const bool a_mn = (idesc >> 15) & 1u, b_mn = (idesc >> 16) & 1u;
  if (M != 128 || N < 16 || N > 256 || a_mn || b_mn || ((idesc >> 7) & 7) != 2 || ((idesc >> 10) & 7) != 2 || ((a_desc >> 61) & 7) != 2 ||
      ((b_desc >> 61) & 7) != 2) {
    std::fprintf(stderr, "emu: unsupported tcgen05.mma.kind::tf32 shape / majorness / format (M=%d N=%d)\n", M, N);
    std::abort();
  }
  static float A[128][8], B[256][8];
  for (int m = 0; m < M; ++m)
    for (int k = 0; k < 8; ++k) A[m][k] = umma_operand_tf32(a_desc, m, k);
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < 8; ++k) B[n][k] = umma_operand_tf32(b_desc, n, k);
  const uint32_t lane0 = d_tmem >> 16, col0 = d_tmem & 0xFFFFu;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      float acc = accumulate ? emu::from_bits<float>(emu_tmem[lane0 + m][col0 + n]) : 0.f;
      for (int k = 0; k < 8; ++k) acc += A[m][k] * B[n][k];
      emu_tmem[lane0 + m][col0 + n] = (uint32_t)emu::to_bits(acc);
    }
}
static inline void emu_red_add_v4(float* dst, float a, float b, float c, float d) {     // red.global.add.v4.f32
  dst[0] += a;
  dst[1] += b;
  dst[2] += c;
  dst[3] += d;
}
}  // namespace w2l
