// Fused multi-tensor NovoGrad step (replaces the per-parameter Python loop of novograd.py:52-114, which issues
// ~8 elementwise kernels and one host sync per parameter tensor).  Three launches for ALL tensors:
//   1. per-tensor sum of squared gradients (block partials -> fp32 atomics)
//   2. per-tensor second moment v = (v == 0 ? n : b2*v + (1-b2)*n), denominator sqrt(v)+eps   (one thread per tensor)
//   3. g' = g/denom + wd*p [* (1-b1)];  m = b1*m + g';  p -= lr*m;  bf16 shadow of p refreshed in the same pass
#include "common.cuh"

namespace w2l {

constexpr int kNgBlocksPerTensor = 64;

__global__ void novograd_norm_kernel(float* const* __restrict__ grads, const int64_t* __restrict__ numel, float* __restrict__ norms) {
  __shared__ float s_part[8];
  const int ti = blockIdx.y;
  const float* g = grads[ti];
  const int64_t n = numel[ti];
  float acc = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = g[i];
    acc = fmaf(v, v, acc);
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += s_part[w];
    if (t != 0.f) atomicAdd(norms + ti, t);
  }
}

__global__ void novograd_moment_kernel(float* __restrict__ exp_avg_sq, float* __restrict__ norms, int n_tensors, float beta2, float eps) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_tensors) return;
  const float nrm = norms[i];
  float v = exp_avg_sq[i];
  v = (v == 0.f) ? nrm : v * beta2 + (1.f - beta2) * nrm;     // novograd.py:93-96
  exp_avg_sq[i] = v;
  norms[i] = sqrtf(v) + eps;                                   // denominator, novograd.py:104
}

__global__ void novograd_update_kernel(float* const* __restrict__ params, float* const* __restrict__ grads,
                                       float* const* __restrict__ exp_avg, const float* __restrict__ denom,
                                       void* const* __restrict__ shadow, const int64_t* __restrict__ numel, float lr, float beta1,
                                       float weight_decay, int grad_averaging) {
  const int ti = blockIdx.y;
  float* p = params[ti];
  const float* g = grads[ti];
  float* m = exp_avg[ti];
  __nv_bfloat16* sh = shadow ? (__nv_bfloat16*)shadow[ti] : nullptr;
  const int64_t n = numel[ti];
  const float inv = 1.f / denom[ti];
  const float ga = grad_averaging ? (1.f - beta1) : 1.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float pv = p[i];
    float gv = g[i] * inv;
    gv = fmaf(weight_decay, pv, gv) * ga;
    const float mv = fmaf(beta1, m[i], gv);
    m[i] = mv;
    const float np = fmaf(-lr, mv, pv);
    p[i] = np;
    if (sh) sh[i] = __float2bfloat16_rn(np);
  }
}

}  // namespace w2l

extern "C" int w2l_novograd_step(float* const* params, float* const* grads, float* const* exp_avg, float* exp_avg_sq,
                                 void* const* shadow_bf16, const int64_t* numel, int32_t n_tensors, float lr, float beta1,
                                 float beta2, float eps, float weight_decay, int32_t grad_averaging, float* norms_ws, void* stream) {
  using namespace w2l;
  W2L_REQUIRE(params && grads && exp_avg && exp_avg_sq && numel && norms_ws, "novograd_step: null pointer");
  W2L_REQUIRE(n_tensors >= 1 && n_tensors <= 65535, "novograd_step: n_tensors=%d out of range", n_tensors);
  cudaStream_t st = (cudaStream_t)stream;
  W2L_CUDA(cudaMemsetAsync(norms_ws, 0, sizeof(float) * n_tensors, st));
  dim3 grid(kNgBlocksPerTensor, n_tensors);
  novograd_norm_kernel<<<grid, 256, 0, st>>>(grads, numel, norms_ws);
  int rc = after_launch("novograd_norm_kernel");
  if (rc) return rc;
  novograd_moment_kernel<<<(n_tensors + 127) / 128, 128, 0, st>>>(exp_avg_sq, norms_ws, n_tensors, beta2, eps);
  rc = after_launch("novograd_moment_kernel");
  if (rc) return rc;
  novograd_update_kernel<<<grid, 256, 0, st>>>(params, grads, exp_avg, norms_ws, shadow_bf16, numel, lr, beta1, weight_decay,
                                               grad_averaging);
  return after_launch("novograd_update_kernel");
}
