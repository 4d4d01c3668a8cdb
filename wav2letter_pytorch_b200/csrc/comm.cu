// Gradient all-reduce (mean) over NVLink peer memory for the data-parallel training step.
// Replaces the bucketed NCCL all-reduce Lightning DDP runs for the reference (README.md:40, config.yaml:21) on the one
// exchange the hot path has: averaging the weight gradients of the replicas.
//
// All ranks hold the gradients in a SYMMETRIC arena (same offsets on every GPU; peers mapped through CUDA IPC / multicast by
// the host).  One small kernel per layer runs beside the remaining backward GEMMs (no shared memory, a handful of CTAs, so it
// co-resides with the persistent GEMM CTAs instead of taking SMs from them):
//   1. barrier-in : every rank has finished this layer's wgrad (flags in peer memory, release/acquire at system scope)
//   2. reduce     : rank r owns slice r of the tensor.
//        NVLS path: multimem.ld_reduce pulls the SUM of all replicas' values out of the NVSwitch, multimem.st broadcasts the
//                   mean back to every replica -- the arithmetic happens in the switch, the SMs only move 1/world of the tensor;
//        P2P  path: the owner loads its slice from every peer in rank order (fixed order => bit-identical results everywhere),
//                   and stores the mean into every peer's copy.
//   3. barrier-out: every rank's slice has landed everywhere before the optimizer may read.
// Flags only grow (one sequence number per call), so a rank that runs ahead can never be mistaken for one that is behind.
#include "common.cuh"

namespace w2l {

constexpr int kCommThreads = 512;
constexpr int kCommMaxWorld = 16;

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 multimem_ld_reduce_add(const float* mc) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(mc)
               : "memory");
  return v;
}
__device__ __forceinline__ void multimem_st(float* mc, float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// peer memory changes under our feet between steps: read it at system scope (never from a stale L1 line)
__device__ __forceinline__ float4 ld_sys_v4(const float* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}

struct CommPeers {
  float* data[kCommMaxWorld];        // peer r's arena (this rank's view of it)
  uint32_t* flags[kCommMaxWorld];    // peer r's flag block: [gridDim.x][world] uint32
};

// Flag slot [cta][src] of rank `dst` is written only by CTA `cta` of rank `src`.  Sequence numbers are compared as wrapped
// differences so that the 32-bit counter may roll over.
__device__ __forceinline__ uint64_t global_timer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// Ranks must reach the same call within `timeout_ns` of each other (default 10 minutes, W2L_COMM_TIMEOUT_S): rank-0-only work
// such as writing a checkpoint belongs between two host-side barriers, not between two training steps of the other ranks.
// A waiting thread backs off with __nanosleep so that the polling CTAs leave the SM's issue slots to the GEMM CTAs beside them.
__device__ __forceinline__ void cta_barrier(const CommPeers& peers, int rank, int world, uint32_t seq, uint64_t timeout_ns) {
  __syncthreads();
  if ((int)threadIdx.x < world) {
    const int r = threadIdx.x;
    __threadfence_system();
    st_release_sys(peers.flags[r] + blockIdx.x * world + rank, seq);
    const uint32_t* mine = peers.flags[rank] + blockIdx.x * world + r;
    uint32_t spins = 0, nap = 32;
    uint64_t t0 = 0;
    while ((int32_t)(ld_acquire_sys(mine) - seq) < 0) {
      if (++spins < 64) continue;                 // the common case: the peer is a few hundred nanoseconds away
      __nanosleep(nap);
      if (nap < 2048) nap <<= 1;
      if ((spins & 1023u) == 0) {
        const uint64_t now = global_timer_ns();
        if (t0 == 0) t0 = now;
        else if (now - t0 > timeout_ns) __trap();  // a peer never arrived: fail the launch instead of hanging the box
      }
    }
  }
  __syncthreads();
}

template <bool NVLS>
__global__ void __launch_bounds__(kCommThreads)
grad_allreduce_kernel(CommPeers peers, float* __restrict__ mc, int64_t offset, int64_t numel, int rank, int world, uint32_t seq,
                      uint64_t timeout_ns) {
  cta_barrier(peers, rank, world, seq, timeout_ns);
  const int64_t vecs = numel >> 2;                                    // 16-byte units; numel % 4 == 0 (host pads)
  const int64_t lo = vecs * rank / world, hi = vecs * (rank + 1) / world;
  const float inv = 1.f / (float)world;
  const int64_t stride = (int64_t)gridDim.x * kCommThreads;
  int64_t i = lo + (int64_t)blockIdx.x * kCommThreads + threadIdx.x;
  if (NVLS) {
    float* base = mc + offset;
    for (; i + 3 * stride < hi; i += 4 * stride) {                    // four requests in flight per thread
      float4 v0 = multimem_ld_reduce_add(base + 4 * i);
      float4 v1 = multimem_ld_reduce_add(base + 4 * (i + stride));
      float4 v2 = multimem_ld_reduce_add(base + 4 * (i + 2 * stride));
      float4 v3 = multimem_ld_reduce_add(base + 4 * (i + 3 * stride));
      multimem_st(base + 4 * i, make_float4(v0.x * inv, v0.y * inv, v0.z * inv, v0.w * inv));
      multimem_st(base + 4 * (i + stride), make_float4(v1.x * inv, v1.y * inv, v1.z * inv, v1.w * inv));
      multimem_st(base + 4 * (i + 2 * stride), make_float4(v2.x * inv, v2.y * inv, v2.z * inv, v2.w * inv));
      multimem_st(base + 4 * (i + 3 * stride), make_float4(v3.x * inv, v3.y * inv, v3.z * inv, v3.w * inv));
    }
    for (; i < hi; i += stride) {
      const float4 v = multimem_ld_reduce_add(base + 4 * i);
      multimem_st(base + 4 * i, make_float4(v.x * inv, v.y * inv, v.z * inv, v.w * inv));
    }
  } else {
    for (; i < hi; i += stride) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      float4 v[kCommMaxWorld];
#pragma unroll
      for (int r = 0; r < kCommMaxWorld; ++r)
        if (r < world) v[r] = ld_sys_v4(peers.data[r] + offset + 4 * i);      // all loads issued before the first use
#pragma unroll
      for (int r = 0; r < kCommMaxWorld; ++r)
        if (r < world) {
          acc.x += v[r].x;
          acc.y += v[r].y;
          acc.z += v[r].z;
          acc.w += v[r].w;
        }
      acc = make_float4(acc.x * inv, acc.y * inv, acc.z * inv, acc.w * inv);
#pragma unroll
      for (int r = 0; r < kCommMaxWorld; ++r)
        if (r < world) *reinterpret_cast<float4*>(peers.data[r] + offset + 4 * i) = acc;
    }
  }
  cta_barrier(peers, rank, world, seq + 1, timeout_ns);
}

}  // namespace w2l

extern "C" int w2l_grad_allreduce(void* const* peer_data_host, void* const* peer_flags_host, void* multicast_base, int64_t offset,
                                  int64_t numel, int32_t rank, int32_t world, uint32_t seq, int32_t ctas, void* stream) {
  using namespace w2l;
  W2L_REQUIRE(peer_data_host && peer_flags_host, "grad_allreduce: null peer tables");
  W2L_REQUIRE(world >= 1 && world <= kCommMaxWorld && rank >= 0 && rank < world, "grad_allreduce: bad rank %d / world %d", rank, world);
  W2L_REQUIRE(numel >= 0 && (numel & 3) == 0 && (offset & 3) == 0, "grad_allreduce: offset %lld / numel %lld must be multiples of 4 floats",
              (long long)offset, (long long)numel);
  W2L_REQUIRE(ctas >= 1 && ctas <= 64, "grad_allreduce: ctas=%d out of range", ctas);
  CommPeers peers;
  for (int r = 0; r < kCommMaxWorld; ++r) {
    peers.data[r] = r < world ? (float*)peer_data_host[r] : nullptr;
    peers.flags[r] = r < world ? (uint32_t*)peer_flags_host[r] : nullptr;
    W2L_REQUIRE(r >= world || (peers.data[r] && peers.flags[r]), "grad_allreduce: null pointer for peer %d", r);
  }
  cudaStream_t st = (cudaStream_t)stream;
  static const uint64_t timeout_ns = [] {
    const char* e = getenv("W2L_COMM_TIMEOUT_S");
    const double sec = e && atof(e) > 0 ? atof(e) : 600.0;
    return (uint64_t)(sec * 1e9);
  }();
  if (multicast_base)
    grad_allreduce_kernel<true><<<ctas, kCommThreads, 0, st>>>(peers, (float*)multicast_base, offset, numel, rank, world, seq, timeout_ns);
  else
    grad_allreduce_kernel<false><<<ctas, kCommThreads, 0, st>>>(peers, nullptr, offset, numel, rank, world, seq, timeout_ns);
  return after_launch("grad_allreduce_kernel");
}
