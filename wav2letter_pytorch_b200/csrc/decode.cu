// Greedy CTC decoding: argmax over classes + collapse (drop blanks / repeats of the previous frame).
// Replaces decoder.py:104-145 (torch.max + a Python loop with two .item() syncs per frame).
//
// Memory-bound: every score is read exactly once.  Pass 1 streams a chunk of frames into shared memory
// with coalesced 128-bit loads, computes one argmax per thread and the number of kept symbols per
// chunk.  Pass 2 turns the per-chunk counts into output positions (ordered compaction, no atomics, so
// the result is deterministic and bit-exact).
#include "common.cuh"

namespace w2l {

constexpr int kDecChunk = 256;  // frames per CTA == threads per CTA

__device__ __forceinline__ int argmax_row(const float* row, int C) {
  // torch.max semantics: first maximal index wins, NaN compares as the maximum.
  float best = row[0];
  int bi = 0;
  for (int c = 1; c < C; ++c) {
    float v = row[c];
    bool best_nan = best != best;
    if (!best_nan && (v != v || v > best)) {
      best = v;
      bi = c;
    }
  }
  return bi;
}

__device__ __forceinline__ int block_exclusive_scan_256(int flag, int* s_warp, int* total) {
  // flag in {0,1}; returns the exclusive prefix within the 256-thread block.
  unsigned ballot = __ballot_sync(0xffffffffu, flag);
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int within = __popc(ballot & ((1u << lane) - 1u));
  if (lane == 0) s_warp[warp] = __popc(ballot);
  __syncthreads();
  int base = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < kDecChunk / 32; ++w) {
    int c = s_warp[w];
    if (w < warp) base += c;
    tot += c;
  }
  *total = tot;
  return base + within;
}

// grid (chunks, N), block 256
__global__ void __launch_bounds__(kDecChunk)
greedy_argmax_kernel(const float* __restrict__ scores, int T, int C, int64_t stride_n, int64_t stride_t,
                     const int32_t* __restrict__ sizes, int blank, int32_t* __restrict__ argmax,
                     int32_t* __restrict__ chunk_counts, int nchunks) {
  extern __shared__ float s_scores[];  // [4 + kDecChunk*C]
  __shared__ int s_warp[kDecChunk / 32];
  __shared__ int s_prev;
  const int n = blockIdx.y, chunk = blockIdx.x;
  const int f0 = chunk * kDecChunk;
  const int frames = min(kDecChunk, T - f0);
  const float* base = scores + (int64_t)n * stride_n + (int64_t)f0 * stride_t;
  int mis = 0;
  if (stride_t == C) {
    // contiguous chunk: flat, 16-byte vectorised, coalesced copy (peel to the alignment of the source)
    const int count = frames * C;
    mis = (int)(((uintptr_t)base >> 2) & 3);          // source misalignment in floats
    const int head = min(count, (4 - mis) & 3);
    const int body4 = (count - head) >> 2;
    const int tail0 = head + (body4 << 2);
    if ((int)threadIdx.x < head) s_scores[mis + threadIdx.x] = base[threadIdx.x];
    const float4* src4 = reinterpret_cast<const float4*>(base + head);
    float4* dst4 = reinterpret_cast<float4*>(s_scores + mis + head);
    for (int i = threadIdx.x; i < body4; i += kDecChunk) dst4[i] = __ldg(src4 + i);
    for (int i = tail0 + threadIdx.x; i < count; i += kDecChunk) s_scores[mis + i] = base[i];
  } else {
    for (int i = threadIdx.x; i < frames * C; i += kDecChunk) {
      int f = i / C, c = i - f * C;
      s_scores[i] = base[(int64_t)f * stride_t + c];
    }
  }
  if (threadIdx.x == 0) {
    // raw argmax of the frame just before this chunk (needed for the repeat test of the first frame)
    s_prev = (f0 > 0) ? argmax_row(base - stride_t, C) : -1;
  }
  __syncthreads();
  const int t = f0 + threadIdx.x;
  int a = -1;
  if ((int)threadIdx.x < frames) {
    a = argmax_row(s_scores + mis + threadIdx.x * C, C);
    argmax[(int64_t)n * T + t] = a;
  }
  // previous frame's argmax: neighbour lane / previous warp via shuffle + smem
  __shared__ int s_last[kDecChunk / 32];
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int prev = __shfl_up_sync(0xffffffffu, a, 1);
  if (lane == 31) s_last[warp] = a;
  __syncthreads();
  if (lane == 0) prev = (warp == 0) ? s_prev : s_last[warp - 1];
  int size = sizes ? max(0, min(T, sizes[n])) : T;
  int keep = ((int)threadIdx.x < frames) && t < size && a != blank && (t == 0 || a != prev);
  int total;
  block_exclusive_scan_256(keep, s_warp, &total);
  if (threadIdx.x == 0) chunk_counts[(int64_t)n * nchunks + chunk] = total;
}

// grid (chunks, N), block 256
__global__ void __launch_bounds__(kDecChunk)
greedy_compact_kernel(const int32_t* __restrict__ argmax, int T, const int32_t* __restrict__ sizes, int blank,
                      const int32_t* __restrict__ chunk_counts, int nchunks, int32_t* __restrict__ tokens,
                      int32_t* __restrict__ offsets, int32_t* __restrict__ counts) {
  __shared__ int s_warp[kDecChunk / 32];
  const int n = blockIdx.y, chunk = blockIdx.x;
  const int f0 = chunk * kDecChunk;
  const int t = f0 + threadIdx.x;
  int base = 0, total_all = 0;
  for (int c = 0; c < nchunks; ++c) {
    int v = chunk_counts[(int64_t)n * nchunks + c];
    if (c < chunk) base += v;
    total_all += v;
  }
  int size = sizes ? max(0, min(T, sizes[n])) : T;
  int a = -1, prev = -1;
  if (t < T) {
    a = argmax[(int64_t)n * T + t];
    if (t > 0) prev = argmax[(int64_t)n * T + t - 1];
  }
  int keep = t < T && t < size && a != blank && (t == 0 || a != prev);
  int total;
  int pos = base + block_exclusive_scan_256(keep, s_warp, &total);
  if (keep) {
    tokens[(int64_t)n * T + pos] = a;
    offsets[(int64_t)n * T + pos] = t;
  }
  if (t < T && t >= total_all) {   // deterministic tail
    tokens[(int64_t)n * T + t] = -1;
    offsets[(int64_t)n * T + t] = -1;
  }
  if (chunk == 0 && threadIdx.x == 0) counts[n] = total_all;
}

}  // namespace w2l

extern "C" {

size_t w2l_greedy_decode_workspace_bytes(int64_t N, int64_t T) {
  int64_t nchunks = (T + w2l::kDecChunk - 1) / w2l::kDecChunk;
  if (nchunks < 1) nchunks = 1;
  return (size_t)(N * nchunks) * sizeof(int32_t);
}

int w2l_greedy_decode(const float* scores, int64_t N, int64_t T, int64_t C, int64_t stride_n, int64_t stride_t,
                      const int32_t* sizes, int32_t blank, int32_t* argmax, int32_t* tokens, int32_t* offsets,
                      int32_t* counts, void* workspace, size_t workspace_bytes, void* stream) {
  using namespace w2l;
  W2L_REQUIRE(N >= 0 && T >= 0 && C >= 1, "greedy_decode: bad shape N=%lld T=%lld C=%lld", (long long)N, (long long)T, (long long)C);
  if (N == 0) return W2L_OK;
  W2L_REQUIRE(scores && argmax && tokens && offsets && counts, "greedy_decode: null pointer");
  W2L_REQUIRE(N <= 65535, "greedy_decode: N=%lld exceeds 65535", (long long)N);
  cudaStream_t st = (cudaStream_t)stream;
  if (T == 0) {
    W2L_CUDA(cudaMemsetAsync(counts, 0, N * sizeof(int32_t), st));
    return W2L_OK;
  }
  W2L_REQUIRE(workspace && workspace_bytes >= w2l_greedy_decode_workspace_bytes(N, T), "greedy_decode: workspace too small");
  const size_t smem = (size_t)(4 + kDecChunk * C) * sizeof(float);
  W2L_REQUIRE(smem <= 200 * 1024, "greedy_decode: C=%lld too large for the shared-memory staging", (long long)C);
  const int nchunks = (int)((T + kDecChunk - 1) / kDecChunk);
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    W2L_CUDA(cudaFuncSetAttribute(greedy_argmax_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  dim3 grid(nchunks, (unsigned)N);
  int32_t* cc = (int32_t*)workspace;
  greedy_argmax_kernel<<<grid, kDecChunk, smem, st>>>(scores, (int)T, (int)C, stride_n, stride_t, sizes, blank, argmax, cc, nchunks);
  int rc = after_launch("greedy_argmax_kernel");
  if (rc) return rc;
  greedy_compact_kernel<<<grid, kDecChunk, 0, st>>>(argmax, (int)T, sizes, blank, cc, nchunks, tokens, offsets, counts);
  return after_launch("greedy_compact_kernel");
}

}  // extern "C"
