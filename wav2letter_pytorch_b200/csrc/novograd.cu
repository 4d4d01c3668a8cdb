// Fused multi-tensor NovoGrad step (replaces the per-parameter Python loop of novograd.py:52-114, which issues
// ~8 elementwise kernels and one host sync per parameter tensor).  Three launches for ALL tensors:
//   1. per-tensor sum of squared gradients (block partials -> fp32 atomics)
//   2. per-tensor second moment v = (v == 0 ? n : b2*v + (1-b2)*n), denominator sqrt(v)+eps   (one thread per tensor)
//   3. g' = g/denom + wd*p [* (1-b1)];  m = b1*m + g';  p -= lr*m;  bf16 shadow of p refreshed in the same pass
#include "common.cuh"

namespace w2l {

// Work is cut into fixed chunks of kNgChunk elements, one CTA per chunk (chunk_prefix[t] = first chunk of tensor t,
// built once by the host from numel): every CTA moves the same number of bytes whatever the tensor sizes are, and each
// thread keeps four 16-byte loads per stream in flight.
constexpr int kNgChunk = 16384;
constexpr int kNgThreads = 256;

__device__ __forceinline__ int chunk_tensor(const int32_t* __restrict__ chunk_prefix, int n_tensors, int chunk) {
  int lo = 0, hi = n_tensors - 1;          // last t with chunk_prefix[t] <= chunk
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (chunk_prefix[mid] <= chunk) lo = mid; else hi = mid - 1;
  }
  return lo;
}

__global__ void __launch_bounds__(kNgThreads)
novograd_norm_kernel(float* const* __restrict__ grads, const int64_t* __restrict__ numel, const int32_t* __restrict__ chunk_prefix,
                     int n_tensors, float* __restrict__ norms) {
  __shared__ float s_part[kNgThreads / 32];
  const int ti = chunk_tensor(chunk_prefix, n_tensors, blockIdx.x);
  const float* g = grads[ti];
  const int64_t n = numel[ti];
  const int64_t begin = (int64_t)(blockIdx.x - chunk_prefix[ti]) * kNgChunk;
  const int64_t end = min(n, begin + kNgChunk);
  float acc = 0.f;
  if ((reinterpret_cast<uintptr_t>(g) & 15) == 0) {
    const int64_t vend = begin + ((end - begin) & ~(int64_t)3);
#pragma unroll 4
    for (int64_t i = begin + threadIdx.x * 4; i < vend; i += kNgThreads * 4) {
      const float4 v = *reinterpret_cast<const float4*>(g + i);
      acc = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, acc))));
    }
    for (int64_t i = vend + threadIdx.x; i < end; i += kNgThreads) acc = fmaf(g[i], g[i], acc);
  } else {
    for (int64_t i = begin + threadIdx.x; i < end; i += kNgThreads) acc = fmaf(g[i], g[i], acc);
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < kNgThreads / 32; ++w) t += s_part[w];
    if (t != 0.f) atomicAdd(norms + ti, t);
  }
}

__global__ void novograd_moment_kernel(float* __restrict__ exp_avg_sq, float* __restrict__ max_exp_avg_sq, float* __restrict__ norms,
                                       int n_tensors, float beta2, float eps) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_tensors) return;
  const float nrm = norms[i];
  float v = exp_avg_sq[i];
  v = (v == 0.f) ? nrm : v * beta2 + (1.f - beta2) * nrm;     // novograd.py:93-96
  exp_avg_sq[i] = v;
  if (max_exp_avg_sq) {                                        // amsgrad: running maximum of the second moment, novograd.py:98-102
    v = fmaxf(max_exp_avg_sq[i], v);
    max_exp_avg_sq[i] = v;
  }
  norms[i] = sqrtf(v) + eps;                                   // denominator, novograd.py:102,104
}

__device__ __forceinline__ float ng_update1(float pv, float gv, float& mv, float inv, float ga, float lr, float beta1, float wd) {
  gv = fmaf(wd, pv, gv * inv) * ga;
  mv = fmaf(beta1, mv, gv);
  return fmaf(-lr, mv, pv);
}

__global__ void __launch_bounds__(kNgThreads)
novograd_update_kernel(float* const* __restrict__ params, float* const* __restrict__ grads, float* const* __restrict__ exp_avg,
                       const float* __restrict__ denom, void* const* __restrict__ shadow, const int64_t* __restrict__ numel,
                       const int32_t* __restrict__ chunk_prefix, int n_tensors, float lr, float beta1, float weight_decay,
                       int grad_averaging) {
  const int ti = chunk_tensor(chunk_prefix, n_tensors, blockIdx.x);
  float* p = params[ti];
  const float* g = grads[ti];
  float* m = exp_avg[ti];
  __nv_bfloat16* sh = shadow ? (__nv_bfloat16*)shadow[ti] : nullptr;
  const int64_t n = numel[ti];
  const int64_t begin = (int64_t)(blockIdx.x - chunk_prefix[ti]) * kNgChunk;
  const int64_t end = min(n, begin + kNgChunk);
  const float inv = 1.f / denom[ti];
  const float ga = grad_averaging ? (1.f - beta1) : 1.f;
  const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m)) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(sh) & 7) == 0;
  int64_t scalar_from = begin;
  if (vec) {
    const int64_t vend = begin + ((end - begin) & ~(int64_t)3);
    scalar_from = vend;
#pragma unroll 4
    for (int64_t i = begin + threadIdx.x * 4; i < vend; i += kNgThreads * 4) {
      float4 pv = *reinterpret_cast<const float4*>(p + i);
      const float4 gv = *reinterpret_cast<const float4*>(g + i);
      float4 mv = *reinterpret_cast<const float4*>(m + i);
      pv.x = ng_update1(pv.x, gv.x, mv.x, inv, ga, lr, beta1, weight_decay);
      pv.y = ng_update1(pv.y, gv.y, mv.y, inv, ga, lr, beta1, weight_decay);
      pv.z = ng_update1(pv.z, gv.z, mv.z, inv, ga, lr, beta1, weight_decay);
      pv.w = ng_update1(pv.w, gv.w, mv.w, inv, ga, lr, beta1, weight_decay);
      *reinterpret_cast<float4*>(m + i) = mv;
      *reinterpret_cast<float4*>(p + i) = pv;
      if (sh) *reinterpret_cast<uint2*>(sh + i) = make_uint2(pack_bf16x2(pv.x, pv.y), pack_bf16x2(pv.z, pv.w));
    }
  }
  for (int64_t i = scalar_from + threadIdx.x; i < end; i += kNgThreads) {
    float mv = m[i];
    const float np = ng_update1(p[i], g[i], mv, inv, ga, lr, beta1, weight_decay);
    m[i] = mv;
    p[i] = np;
    if (sh) sh[i] = __float2bfloat16_rn(np);
  }
}

}  // namespace w2l

extern "C" int32_t w2l_novograd_chunk(void) { return w2l::kNgChunk; }

extern "C" int w2l_novograd_step(float* const* params, float* const* grads, float* const* exp_avg, float* exp_avg_sq,
                                 float* max_exp_avg_sq, void* const* shadow_bf16, const int64_t* numel, const int32_t* chunk_prefix, int32_t n_tensors,
                                 int32_t n_chunks, float lr, float beta1, float beta2, float eps, float weight_decay,
                                 int32_t grad_averaging, float* norms_ws, void* stream) {
  using namespace w2l;
  W2L_REQUIRE(params && grads && exp_avg && exp_avg_sq && numel && norms_ws && chunk_prefix, "novograd_step: null pointer");
  W2L_REQUIRE(n_tensors >= 1 && n_tensors <= 65535, "novograd_step: n_tensors=%d out of range", n_tensors);
  W2L_REQUIRE(n_chunks >= 1, "novograd_step: n_chunks=%d", n_chunks);
  cudaStream_t st = (cudaStream_t)stream;
  W2L_CUDA(cudaMemsetAsync(norms_ws, 0, sizeof(float) * n_tensors, st));
  novograd_norm_kernel<<<n_chunks, kNgThreads, 0, st>>>(grads, numel, chunk_prefix, n_tensors, norms_ws);
  int rc = after_launch("novograd_norm_kernel");
  if (rc) return rc;
  novograd_moment_kernel<<<(n_tensors + 127) / 128, 128, 0, st>>>(exp_avg_sq, max_exp_avg_sq, norms_ws, n_tensors, beta2, eps);
  rc = after_launch("novograd_moment_kernel");
  if (rc) return rc;
  novograd_update_kernel<<<n_chunks, kNgThreads, 0, st>>>(params, grads, exp_avg, norms_ws, shadow_bf16, numel, chunk_prefix, n_tensors,
                                                         lr, beta1, weight_decay, grad_averaging);
  return after_launch("novograd_update_kernel");
}
