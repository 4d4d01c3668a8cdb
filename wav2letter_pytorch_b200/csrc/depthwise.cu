// Depthwise Conv1d (groups = channels) for the separable Jasper sub-blocks of the shipped model/jasper.yaml
// (jasper.py:318-341: depthwise k-tap conv + pointwise 1x1; the pointwise half runs on the tensor-core GEMM kernel).
// Memory/L1-bound CUDA-core kernels over time-major bf16: a thread owns 8 channels (one 16-byte vector) of one output row;
// neighbouring threads in y walk neighbouring rows, so the k input rows each thread reads are L1 hits after the first.
//   fwd   y[b,t,c]  = sum_j x[b, t*s + j*d - p, c] * w[j,c]            rows >= out_lens[b] written as 0 (consumer's mask)
//   dgrad dx[b,u,c] = sum_j dy[b, u + p - j*d, c] * w[j,c]             (stride 1), dy rows >= dy_lens[b] read as 0
//         dx[b,u,c] = sum_{j : s | u + p - j*d} dy[b, (u + p - j*d)/s, c] * w[j,c]      (stride s > 1, own kernel)
//   wgrad dw[j,c]  += sum_{b,t} dy[b,t,c] * x[b, t*s + j*d - p, c]
// Weights are fp32 [k, C] (the [C,1,k] Parameter is a permuted view of this storage).
#include <stdlib.h>

#include "common.cuh"

namespace w2l {

__device__ __forceinline__ void dw_unpack8(const uint4& q, float (&v)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __bfloat1622float2(h[i]);
    v[2 * i] = f.x;
    v[2 * i + 1] = f.y;
  }
}

// 8 consecutive channels of one row, bf16 or fp32 storage (the fp32-faithful mode of a separable Jasper block): the plain kernels
// below are written once over these; the register-tiled ones stay bf16-only
__device__ __forceinline__ void dw_load8(const __nv_bfloat16* p, float (&v)[8]) { dw_unpack8(__ldg(reinterpret_cast<const uint4*>(p)), v); }
__device__ __forceinline__ void dw_load8(const float* p, float (&v)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
  v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void dw_store8(__nv_bfloat16* p, const float (&v)[8]) {
  uint4 q;
  q.x = pack_bf16x2(v[0], v[1]);
  q.y = pack_bf16x2(v[2], v[3]);
  q.z = pack_bf16x2(v[4], v[5]);
  q.w = pack_bf16x2(v[6], v[7]);
  *reinterpret_cast<uint4*>(p) = q;
}
__device__ __forceinline__ void dw_store8(float* p, const float (&v)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *(reinterpret_cast<float4*>(p) + 1) = make_float4(v[4], v[5], v[6], v[7]);
}

// Generic correlation: out[b,t,c] = sum_j in[b, t*stride + j*dil*dir + off, c] * w[jw(j), c], jw(j) = flip ? k-1-j : j.
// Rows of `in` outside [0, in_rows) or >= in_lens[b] read as zero; rows of `out` >= out_lens[b] are written as zero.
template <typename TA>
__global__ void __launch_bounds__(256)
depthwise_corr_kernel(const TA* __restrict__ in, const float* __restrict__ w, TA* __restrict__ out, int B,
                      int in_rows, int out_rows, int C, int k, int stride, int dil, int off, int flip,
                      const int32_t* __restrict__ in_lens, const int32_t* __restrict__ out_lens) {
  const int c = (blockIdx.x * 32 + threadIdx.x) * 8;
  const int r = blockIdx.y * 8 + threadIdx.y;            // flattened (b, t)
  if (c >= C || r >= B * out_rows) return;
  const int b = r / out_rows, t = r - b * out_rows;
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  if (!out_lens || t < out_lens[b]) {
    const int lim = in_lens ? min(in_rows, max(0, in_lens[b])) : in_rows;
    const TA* ib = in + (int64_t)b * in_rows * C + c;
    for (int j = 0; j < k; ++j) {
      const int u = t * stride + j * dil + off;
      if (u < 0 || u >= lim) continue;
      float xv[8];
      dw_load8(ib + (int64_t)u * C, xv);
      const float* wj = w + (int64_t)(flip ? k - 1 - j : j) * C + c;
      const float4 w0 = __ldg(reinterpret_cast<const float4*>(wj)), w1 = __ldg(reinterpret_cast<const float4*>(wj) + 1);
      acc[0] = fmaf(xv[0], w0.x, acc[0]); acc[1] = fmaf(xv[1], w0.y, acc[1]);
      acc[2] = fmaf(xv[2], w0.z, acc[2]); acc[3] = fmaf(xv[3], w0.w, acc[3]);
      acc[4] = fmaf(xv[4], w1.x, acc[4]); acc[5] = fmaf(xv[5], w1.y, acc[5]);
      acc[6] = fmaf(xv[6], w1.z, acc[6]); acc[7] = fmaf(xv[7], w1.w, acc[7]);
    }
  }
  dw_store8(out + (int64_t)r * C + c, acc);
}

// Backward-data of a STRIDED depthwise conv (a strided separable block that is not the encoder's first one):
// dx[b,u,c] = sum over taps j with u + pad - j*dil = t*stride, 0 <= t < min(y_rows, dy_lens[b]), of dy[b,t,c] * w[j,c].
template <typename TA>
__global__ void __launch_bounds__(256)
depthwise_dgrad_strided_kernel(const TA* __restrict__ dy, const float* __restrict__ w, TA* __restrict__ dx, int B,
                               int x_rows, int y_rows, int C, int k, int stride, int dil, int pad, const int32_t* __restrict__ dy_lens) {
  const int c = (blockIdx.x * 32 + threadIdx.x) * 8;
  const int r = blockIdx.y * 8 + threadIdx.y;            // flattened (b, u)
  if (c >= C || r >= B * x_rows) return;
  const int b = r / x_rows, u = r - b * x_rows;
  const int lim = dy_lens ? min(y_rows, max(0, dy_lens[b])) : y_rows;
  const TA* gb = dy + (int64_t)b * y_rows * C + c;
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  for (int j = 0; j < k; ++j) {
    const int v = u + pad - j * dil;
    if (v < 0) break;                                    // v only decreases with j
    const int t = v / stride;
    if (t * stride != v || t >= lim) continue;
    float g[8];
    dw_load8(gb + (int64_t)t * C, g);
    const float* wj = w + (int64_t)j * C + c;
    const float4 w0 = __ldg(reinterpret_cast<const float4*>(wj)), w1 = __ldg(reinterpret_cast<const float4*>(wj) + 1);
    acc[0] = fmaf(g[0], w0.x, acc[0]); acc[1] = fmaf(g[1], w0.y, acc[1]);
    acc[2] = fmaf(g[2], w0.z, acc[2]); acc[3] = fmaf(g[3], w0.w, acc[3]);
    acc[4] = fmaf(g[4], w1.x, acc[4]); acc[5] = fmaf(g[5], w1.y, acc[5]);
    acc[6] = fmaf(g[6], w1.z, acc[6]); acc[7] = fmaf(g[7], w1.w, acc[7]);
  }
  dw_store8(dx + (int64_t)r * C + c, acc);
}

// block (32, 8); grid (channel blocks, row chunks, tap groups of 4)
template <typename TA>
__global__ void __launch_bounds__(256)
depthwise_wgrad_kernel(const TA* __restrict__ dy, const TA* __restrict__ x, float* __restrict__ dw, int B,
                       int x_rows, int y_rows, int C, int k, int stride, int dil, int pad, const int32_t* __restrict__ dy_lens,
                       int rows_per_block) {
  __shared__ float s_acc[8][4][256 + 8];
  const int c = (blockIdx.x * 32 + threadIdx.x) * 8;
  const int j0 = blockIdx.z * 4;
  float acc[4][8];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[a][i] = 0.f;
  if (c < C) {
    const int rows = B * y_rows;
    const int r_begin = blockIdx.y * rows_per_block, r_end = min(rows, r_begin + rows_per_block);
    for (int r = r_begin + threadIdx.y; r < r_end; r += 8) {
      const int b = r / y_rows, t = r - b * y_rows;
      if (dy_lens && t >= dy_lens[b]) continue;
      float g[8];
      dw_load8(dy + (int64_t)r * C + c, g);
      const TA* xb = x + (int64_t)b * x_rows * C + c;
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const int j = j0 + a, u = t * stride + j * dil - pad;
        if (j < k && u >= 0 && u < x_rows) {
          float xv[8];
          dw_load8(xb + (int64_t)u * C, xv);
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[a][i] = fmaf(g[i], xv[i], acc[a][i]);
        }
      }
    }
  }
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int i = 0; i < 8; ++i) s_acc[threadIdx.y][a][threadIdx.x * 8 + i] = acc[a][i];
  __syncthreads();
  const int tid = threadIdx.y * 32 + threadIdx.x;
  const int cc = blockIdx.x * 256 + tid;
  if (cc < C) {
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      if (j0 + a >= k) break;
      float v = 0.f;
#pragma unroll
      for (int y = 0; y < 8; ++y) v += s_acc[y][a][tid];
      atomicAdd(dw + (int64_t)(j0 + a) * C + cc, v);
    }
  }
}

// ---- register-tiled variants for stride 1, dilation 1 (default; W2L_DW_TILED=0 turns them off) -------------------------------------------------------
// The kernels above issue three 16-byte loads (one activation row, two weight vectors) per 8 FMAs: they are bound by the load pipe,
// not by HBM (every re-read is an L1 hit).  Here a thread keeps a WINDOW of kDwTile activation rows in registers and slides it:
//   correlation: a thread owns 8 channels x kDwTile consecutive output rows; tap j pairs output r with input row r + j, so one new
//                row and one weight vector per tap feed kDwTile x 8 FMAs (3 loads per 64 FMAs instead of per 8);
//   wgrad:       a thread owns 8 channels x kDwTile consecutive taps; row t pairs tap a with input row t + a, so one dy row and one
//                new x row per output row feed kDwTile x 8 FMAs (2 loads per 64 FMAs instead of 5 per 32), and dy is read k/8 times
//                instead of k/4 times.
// Taps are visited in the same order as in the kernels above, so the correlation results are bit-identical to theirs.
constexpr int kDwTile = 8;

__device__ __forceinline__ void dw_load_row(const __nv_bfloat16* base, int u, int lim, int C, float (&dst)[8]) {
  if (u >= 0 && u < lim) {
    dw_unpack8(__ldg(reinterpret_cast<const uint4*>(base + (int64_t)u * C)), dst);
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) dst[i] = 0.f;
  }
}

// out[b,t,c] = sum_j in[b, t + j + off, c] * w[flip ? k-1-j : j, c]; block (32, 8); grid (channel blocks, tiles of kDwTile rows / 8)
__global__ void __launch_bounds__(256)
depthwise_corr_tiled_kernel(const __nv_bfloat16* __restrict__ in, const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int B,
                            int in_rows, int out_rows, int C, int k, int off, int flip, const int32_t* __restrict__ in_lens,
                            const int32_t* __restrict__ out_lens, int tiles_per_utt) {
  const int c = (blockIdx.x * 32 + threadIdx.x) * 8;
  const int tile = blockIdx.y * 8 + threadIdx.y;         // flattened (b, tile of rows)
  if (c >= C || tile >= B * tiles_per_utt) return;
  const int b = tile / tiles_per_utt, t0 = (tile - b * tiles_per_utt) * kDwTile;
  const int olim = out_lens ? min(out_rows, max(0, out_lens[b])) : out_rows;
  const int lim = in_lens ? min(in_rows, max(0, in_lens[b])) : in_rows;
  const __nv_bfloat16* ib = in + (int64_t)b * in_rows * C + c;
  float acc[kDwTile][8], win[kDwTile][8];
#pragma unroll
  for (int r = 0; r < kDwTile; ++r)
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[r][i] = 0.f;
  if (t0 < olim) {
    const int base = t0 + off;
#pragma unroll
    for (int r = 0; r < kDwTile - 1; ++r) dw_load_row(ib, base + r, lim, C, win[r]);
    for (int j0 = 0; j0 < k; j0 += kDwTile) {
#pragma unroll
      for (int jj = 0; jj < kDwTile; ++jj) {
        const int j = j0 + jj;
        if (j < k) {
          dw_load_row(ib, base + j + kDwTile - 1, lim, C, win[(jj + kDwTile - 1) % kDwTile]);   // row m lives in slot m % kDwTile
          const float* wj = w + (int64_t)(flip ? k - 1 - j : j) * C + c;
          const float4 w0 = __ldg(reinterpret_cast<const float4*>(wj)), w1 = __ldg(reinterpret_cast<const float4*>(wj) + 1);
#pragma unroll
          for (int r = 0; r < kDwTile; ++r) {
            const float (&xv)[8] = win[(jj + r) % kDwTile];
            acc[r][0] = fmaf(xv[0], w0.x, acc[r][0]); acc[r][1] = fmaf(xv[1], w0.y, acc[r][1]);
            acc[r][2] = fmaf(xv[2], w0.z, acc[r][2]); acc[r][3] = fmaf(xv[3], w0.w, acc[r][3]);
            acc[r][4] = fmaf(xv[4], w1.x, acc[r][4]); acc[r][5] = fmaf(xv[5], w1.y, acc[r][5]);
            acc[r][6] = fmaf(xv[6], w1.z, acc[r][6]); acc[r][7] = fmaf(xv[7], w1.w, acc[r][7]);
          }
        }
      }
    }
  }
  __nv_bfloat16* ob = out + ((int64_t)b * out_rows + t0) * C + c;
#pragma unroll
  for (int r = 0; r < kDwTile; ++r) {
    if (t0 + r < out_rows) {
      uint4 q = make_uint4(0u, 0u, 0u, 0u);
      if (t0 + r < olim) {
        q.x = pack_bf16x2(acc[r][0], acc[r][1]);
        q.y = pack_bf16x2(acc[r][2], acc[r][3]);
        q.z = pack_bf16x2(acc[r][4], acc[r][5]);
        q.w = pack_bf16x2(acc[r][6], acc[r][7]);
      }
      *reinterpret_cast<uint4*>(ob + (int64_t)r * C) = q;
    }
  }
}

// dw[j,c] += sum_{b,t} dy[b,t,c] * x[b, t + j - pad, c]; block (32, 8); grid (channel blocks, row chunks, tap groups of kDwTile).
// Each of the block's 8 row-threads walks a CONTIGUOUS eighth of the block's rows (the window slides along t), utterance by utterance.
__global__ void __launch_bounds__(256)
depthwise_wgrad_tiled_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x, float* __restrict__ dw, int B,
                             int x_rows, int y_rows, int C, int k, int pad, const int32_t* __restrict__ dy_lens, int rows_per_block) {
  __shared__ float s_acc[8][4][256 + 8];
  const int c = (blockIdx.x * 32 + threadIdx.x) * 8;
  const int j0 = blockIdx.z * kDwTile;
  float acc[kDwTile][8];
#pragma unroll
  for (int a = 0; a < kDwTile; ++a)
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[a][i] = 0.f;
  if (c < C) {
    const int rows = B * y_rows;
    const int per = (rows_per_block + 7) / 8;
    const int r_begin = min(rows, blockIdx.y * rows_per_block + threadIdx.y * per);
    const int r_end = min(min(rows, (blockIdx.y + 1) * rows_per_block), r_begin + per);
    int r = r_begin;
    while (r < r_end) {                       // one utterance's segment [ta, tb) of this thread's rows
      const int b = r / y_rows, ta = r - b * y_rows;
      const int seg_end = min(r_end, (b + 1) * y_rows);
      const int tb = min(seg_end - b * y_rows, dy_lens ? max(0, dy_lens[b]) : y_rows);
      r = seg_end;
      if (ta >= tb) continue;
      const __nv_bfloat16* xb = x + (int64_t)b * x_rows * C + c;
      const __nv_bfloat16* gb = dy + (int64_t)b * y_rows * C + c;
      const int base = ta + j0 - pad;         // x row paired with (t = ta, tap j0)
      float win[kDwTile][8];
#pragma unroll
      for (int a = 0; a < kDwTile - 1; ++a) dw_load_row(xb, base + a, x_rows, C, win[a]);
      for (int q0 = 0; ta + q0 < tb; q0 += kDwTile) {
#pragma unroll
        for (int qq = 0; qq < kDwTile; ++qq) {
          const int t = ta + q0 + qq;
          if (t < tb) {
            dw_load_row(xb, base + q0 + qq + kDwTile - 1, x_rows, C, win[(qq + kDwTile - 1) % kDwTile]);
            float g[8];
            dw_unpack8(__ldg(reinterpret_cast<const uint4*>(gb + (int64_t)t * C)), g);
#pragma unroll
            for (int a = 0; a < kDwTile; ++a) {
              const float (&xv)[8] = win[(qq + a) % kDwTile];
#pragma unroll
              for (int i = 0; i < 8; ++i) acc[a][i] = fmaf(g[i], xv[i], acc[a][i]);
            }
          }
        }
      }
    }
  }
  const int tid = threadIdx.y * 32 + threadIdx.x;
  const int cc = blockIdx.x * 256 + tid;
#pragma unroll
  for (int half = 0; half < kDwTile / 4; ++half) {         // fold the 8 row-threads in shared memory, 4 taps at a time
    __syncthreads();
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int i = 0; i < 8; ++i) s_acc[threadIdx.y][a][threadIdx.x * 8 + i] = acc[half * 4 + a][i];
    __syncthreads();
    if (cc < C) {
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const int j = j0 + half * 4 + a;
        if (j < k) {
          float v = 0.f;
#pragma unroll
          for (int y = 0; y < 8; ++y) v += s_acc[y][a][tid];
          atomicAdd(dw + (int64_t)j * C + cc, v);
        }
      }
    }
  }
}

// Register-tiled kernels are the default where they apply (stride 1, dilation 1): measured on B200 with the shipped separable
// jasper.yaml, 19.4 -> 18.3 ms per training step (profiles/r2_dw_ab.md).  W2L_DW_TILED=0 selects the plain kernels (the tests' A/B switch).
static bool dw_tiled_requested() {
  const char* e = getenv("W2L_DW_TILED");
  return !(e && atoi(e) == 0);
}

static int dw_check(const char* who, int B, int T, int C, int T_out, int k, int stride, int dil, int pad) {
  W2L_REQUIRE(B >= 1 && T >= 1 && T_out >= 1 && k >= 1 && stride >= 1 && dil >= 1 && pad >= 0, "%s: bad geometry", who);
  W2L_REQUIRE(C >= 8 && C % 8 == 0, "%s: C=%d must be a multiple of 8", who, C);
  W2L_REQUIRE((int64_t)B * (T > T_out ? T : T_out) < (1ll << 28), "%s: B*T too large", who);
  return W2L_OK;
}

}  // namespace w2l

extern "C" {

static int dw_fwd_impl(int f32, const void* x, const float* w, void* y, int32_t B, int32_t T, int32_t C, int32_t T_out, int32_t k,
                       int32_t stride, int32_t dilation, int32_t pad, const int32_t* out_lens, void* stream) {
  using namespace w2l;
  int rc = dw_check("depthwise_fwd", B, T, C, T_out, k, stride, dilation, pad);
  if (rc) return rc;
  W2L_REQUIRE(x && w && y, "depthwise_fwd: null pointer");
  dim3 grid((C / 8 + 31) / 32, (B * T_out + 7) / 8), block(32, 8);
  if (f32) {
    depthwise_corr_kernel<float><<<grid, block, 0, (cudaStream_t)stream>>>((const float*)x, w, (float*)y, B, T, T_out, C, k, stride, dilation,
                                                                           -pad, 0, nullptr, out_lens);
    return after_launch("depthwise_corr_kernel<fwd>");
  }
  if (stride == 1 && dilation == 1 && dw_tiled_requested()) {
    const int tiles = (T_out + kDwTile - 1) / kDwTile;
    dim3 tgrid((C / 8 + 31) / 32, (B * tiles + 7) / 8);
    depthwise_corr_tiled_kernel<<<tgrid, block, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, w, (__nv_bfloat16*)y, B, T, T_out, C, k,
                                                                           -pad, 0, nullptr, out_lens, tiles);
    return after_launch("depthwise_corr_tiled_kernel<fwd>");
  }
  depthwise_corr_kernel<__nv_bfloat16><<<grid, block, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, w, (__nv_bfloat16*)y, B, T, T_out, C,
                                                                                 k, stride, dilation, -pad, 0, nullptr, out_lens);
  return after_launch("depthwise_corr_kernel<fwd>");
}
int w2l_depthwise_fwd(const void* x, const float* w, void* y, int32_t B, int32_t T, int32_t C, int32_t T_out, int32_t k, int32_t stride,
                      int32_t dilation, int32_t pad, const int32_t* out_lens, void* stream) {
  return dw_fwd_impl(0, x, w, y, B, T, C, T_out, k, stride, dilation, pad, out_lens, stream);
}
int w2l_depthwise_fwd_f32(const float* x, const float* w, float* y, int32_t B, int32_t T, int32_t C, int32_t T_out, int32_t k, int32_t stride,
                          int32_t dilation, int32_t pad, const int32_t* out_lens, void* stream) {
  return dw_fwd_impl(1, x, w, y, B, T, C, T_out, k, stride, dilation, pad, out_lens, stream);
}

static int dw_dgrad_impl(int f32, const void* dy, const float* w, void* dx, int32_t B, int32_t T, int32_t C, int32_t T_out, int32_t k,
                         int32_t dilation, int32_t pad, const int32_t* dy_lens, void* stream) {
  using namespace w2l;
  int rc = dw_check("depthwise_dgrad", B, T, C, T_out, k, 1, dilation, pad);
  if (rc) return rc;
  W2L_REQUIRE(dy && w && dx, "depthwise_dgrad: null pointer");
  // dx[u] = sum_j dy[u + p - j*d] w[j] = sum_j' dy[u + p - (k-1)d + j'*d] w[k-1-j']
  dim3 grid((C / 8 + 31) / 32, (B * T + 7) / 8), block(32, 8);
  if (f32) {
    depthwise_corr_kernel<float><<<grid, block, 0, (cudaStream_t)stream>>>((const float*)dy, w, (float*)dx, B, T_out, T, C, k, 1, dilation,
                                                                           pad - (k - 1) * dilation, 1, dy_lens, nullptr);
    return after_launch("depthwise_corr_kernel<dgrad>");
  }
  if (dilation == 1 && dw_tiled_requested()) {
    const int tiles = (T + kDwTile - 1) / kDwTile;
    dim3 tgrid((C / 8 + 31) / 32, (B * tiles + 7) / 8);
    depthwise_corr_tiled_kernel<<<tgrid, block, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)dy, w, (__nv_bfloat16*)dx, B, T_out, T, C, k,
                                                                           pad - (k - 1), 1, dy_lens, nullptr, tiles);
    return after_launch("depthwise_corr_tiled_kernel<dgrad>");
  }
  depthwise_corr_kernel<__nv_bfloat16><<<grid, block, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)dy, w, (__nv_bfloat16*)dx, B, T_out, T, C,
                                                                                 k, 1, dilation, pad - (k - 1) * dilation, 1, dy_lens, nullptr);
  return after_launch("depthwise_corr_kernel<dgrad>");
}
int w2l_depthwise_dgrad(const void* dy, const float* w, void* dx, int32_t B, int32_t T, int32_t C, int32_t T_out, int32_t k,
                        int32_t dilation, int32_t pad, const int32_t* dy_lens, void* stream) {
  return dw_dgrad_impl(0, dy, w, dx, B, T, C, T_out, k, dilation, pad, dy_lens, stream);
}
int w2l_depthwise_dgrad_f32(const float* dy, const float* w, float* dx, int32_t B, int32_t T, int32_t C, int32_t T_out, int32_t k,
                            int32_t dilation, int32_t pad, const int32_t* dy_lens, void* stream) {
  return dw_dgrad_impl(1, dy, w, dx, B, T, C, T_out, k, dilation, pad, dy_lens, stream);
}

static int dw_dgrad_strided_impl(int f32, const void* dy, const float* w, void* dx, int32_t B, int32_t T, int32_t C, int32_t T_out, int32_t k,
                                 int32_t stride, int32_t dilation, int32_t pad, const int32_t* dy_lens, void* stream) {
  using namespace w2l;
  int rc = dw_check("depthwise_dgrad_strided", B, T, C, T_out, k, stride, dilation, pad);
  if (rc) return rc;
  W2L_REQUIRE(dy && w && dx, "depthwise_dgrad_strided: null pointer");
  dim3 grid((C / 8 + 31) / 32, (B * T + 7) / 8), block(32, 8);
  if (f32)
    depthwise_dgrad_strided_kernel<float><<<grid, block, 0, (cudaStream_t)stream>>>((const float*)dy, w, (float*)dx, B, T, T_out, C, k, stride,
                                                                                    dilation, pad, dy_lens);
  else
    depthwise_dgrad_strided_kernel<__nv_bfloat16><<<grid, block, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)dy, w, (__nv_bfloat16*)dx, B, T,
                                                                                            T_out, C, k, stride, dilation, pad, dy_lens);
  return after_launch("depthwise_dgrad_strided_kernel");
}
int w2l_depthwise_dgrad_strided(const void* dy, const float* w, void* dx, int32_t B, int32_t T, int32_t C, int32_t T_out, int32_t k,
                                int32_t stride, int32_t dilation, int32_t pad, const int32_t* dy_lens, void* stream) {
  return dw_dgrad_strided_impl(0, dy, w, dx, B, T, C, T_out, k, stride, dilation, pad, dy_lens, stream);
}
int w2l_depthwise_dgrad_strided_f32(const float* dy, const float* w, float* dx, int32_t B, int32_t T, int32_t C, int32_t T_out, int32_t k,
                                    int32_t stride, int32_t dilation, int32_t pad, const int32_t* dy_lens, void* stream) {
  return dw_dgrad_strided_impl(1, dy, w, dx, B, T, C, T_out, k, stride, dilation, pad, dy_lens, stream);
}

static int dw_wgrad_impl(int f32, const void* dy, const void* x, float* dw, int32_t B, int32_t T, int32_t C, int32_t T_out, int32_t k,
                         int32_t stride, int32_t dilation, int32_t pad, const int32_t* dy_lens, void* stream) {
  using namespace w2l;
  int rc = dw_check("depthwise_wgrad", B, T, C, T_out, k, stride, dilation, pad);
  if (rc) return rc;
  W2L_REQUIRE(dy && x && dw, "depthwise_wgrad: null pointer");
  const int rows = B * T_out;
  int rpb = (rows + num_sms() - 1) / num_sms();
  if (rpb < 64) rpb = 64;
  if (!f32 && stride == 1 && dilation == 1 && dw_tiled_requested()) {
    dim3 tgrid((C / 8 + 31) / 32, (rows + rpb - 1) / rpb, (k + kDwTile - 1) / kDwTile), tblock(32, 8);
    depthwise_wgrad_tiled_kernel<<<tgrid, tblock, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)dy, (const __nv_bfloat16*)x, dw, B, T, T_out,
                                                                             C, k, pad, dy_lens, rpb);
    return after_launch("depthwise_wgrad_tiled_kernel");
  }
  dim3 grid((C / 8 + 31) / 32, (rows + rpb - 1) / rpb, (k + 3) / 4), block(32, 8);
  if (f32)
    depthwise_wgrad_kernel<float><<<grid, block, 0, (cudaStream_t)stream>>>((const float*)dy, (const float*)x, dw, B, T, T_out, C, k, stride,
                                                                            dilation, pad, dy_lens, rpb);
  else
    depthwise_wgrad_kernel<__nv_bfloat16><<<grid, block, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)dy, (const __nv_bfloat16*)x, dw, B, T,
                                                                                    T_out, C, k, stride, dilation, pad, dy_lens, rpb);
  return after_launch("depthwise_wgrad_kernel");
}
int w2l_depthwise_wgrad(const void* dy, const void* x, float* dw, int32_t B, int32_t T, int32_t C, int32_t T_out, int32_t k,
                        int32_t stride, int32_t dilation, int32_t pad, const int32_t* dy_lens, void* stream) {
  return dw_wgrad_impl(0, dy, x, dw, B, T, C, T_out, k, stride, dilation, pad, dy_lens, stream);
}
int w2l_depthwise_wgrad_f32(const float* dy, const float* x, float* dw, int32_t B, int32_t T, int32_t C, int32_t T_out, int32_t k,
                            int32_t stride, int32_t dilation, int32_t pad, const int32_t* dy_lens, void* stream) {
  return dw_wgrad_impl(1, dy, x, dw, B, T, C, T_out, k, stride, dilation, pad, dy_lens, stream);
}

}  // extern "C"
